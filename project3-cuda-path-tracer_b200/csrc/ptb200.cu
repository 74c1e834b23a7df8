// ptb200.cu -- implementation of the C ABI in include/ptb200.h: device memory, launch sequences,
// CUDA-graph replay of the per-batch bounce loop.  Host orchestration only; the kernels are in
// ptb_kernels.cuh.  Replaces the host side of src/pathtrace.cu (pathtraceInit :79-101,
// pathtraceFree :103-112, pathtrace :284-393) of the reference.
#include "ptb200.h"
#include "ptb_kernels.cuh"
#include "ptb_bvh.h"

#include <dlfcn.h>
#include <nccl.h>                   // types and prototypes only: libnccl.so.2 is dlopen'ed on first multi-GPU use

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using namespace ptb;

// which BVH traversal the library launches: the per-lane trace_kernel, or trace_pool_kernel (64 ray slots per warp in shared
// memory, any lane runs any slot) -- csrc/ptb_trace.cuh
#ifndef PTB_TRACE_POOL
#define PTB_TRACE_POOL 0
#endif
#if PTB_TRACE_POOL
#define TRACE_KERNEL trace_pool_kernel
constexpr int kTraceStackSmem = kPoolStack, kTraceSlotsPerThread = 2;
constexpr size_t traceKernelSmem() { return poolSmemBytes(); }
#else
#define TRACE_KERNEL trace_kernel
constexpr int kTraceStackSmem = kStackSmem, kTraceSlotsPerThread = 1;
constexpr size_t traceKernelSmem() { return traceSmemBytes(); }
#endif

// the fused pipeline's material sort (bounce_regroup_kernel) halves the instructions of the BVH shading kernel but that kernel waits on DRAM,
// not on issue slots: 2 284 against 2 255 us per iteration of the 100 k-object scene (profiles/r3b_regroup.txt) -- a toggle, off by default
#ifndef PTB_REGROUP_DEFAULT
#define PTB_REGROUP_DEFAULT 0
#endif

namespace {

thread_local std::string g_lastError;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(PTB_ERR_CUDA, "CUDA error (%s:%d): %s: %s", "ptb200.cu", __LINE__, #call,  \
                        cudaGetErrorString(e_));                                                   \
    } while (0)

// Every entry point makes the context's device current for its own duration and restores the caller's (a multi-GPU host
// such as torch keeps its own current device; src/preview.cpp:106-111 of the reference pins device 0 once instead).
struct DeviceScope {
    int prev = -1;
    cudaError_t err;
    explicit DeviceScope(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; err = cudaSetDevice(dev); }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev)                                                                             \
    DeviceScope deviceScope_(dev);                                                                 \
    if (deviceScope_.err != cudaSuccess)                                                           \
        return fail(PTB_ERR_CUDA, "cudaSetDevice(%d): %s", (int)(dev), cudaGetErrorString(deviceScope_.err))

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc(&p, count * sizeof(T));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }                  // an early error return of a ptb_stage_* call frees what it had allocated
};

} // namespace

// ---- NCCL, loaded on first use ----------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) getUniqueId = nullptr;
    decltype(&ncclCommInitRank) commInitRank = nullptr;
    decltype(&ncclCommInitAll) commInitAll = nullptr;
    decltype(&ncclCommDestroy) commDestroy = nullptr;
    decltype(&ncclReduce) reduce = nullptr;
    decltype(&ncclGroupStart) groupStart = nullptr;
    decltype(&ncclGroupEnd) groupEnd = nullptr;
    decltype(&ncclGetErrorString) getErrorString = nullptr;
};
static NcclApi *ncclApi(std::string &err) {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // a copy a host framework (torch) already mapped wins: same soname
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
#define PTB_NCCL_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, #sym))
            PTB_NCCL_SYM(getUniqueId, ncclGetUniqueId); PTB_NCCL_SYM(commInitRank, ncclCommInitRank); PTB_NCCL_SYM(commInitAll, ncclCommInitAll);
            PTB_NCCL_SYM(commDestroy, ncclCommDestroy); PTB_NCCL_SYM(reduce, ncclReduce); PTB_NCCL_SYM(groupStart, ncclGroupStart);
            PTB_NCCL_SYM(groupEnd, ncclGroupEnd); PTB_NCCL_SYM(getErrorString, ncclGetErrorString);
#undef PTB_NCCL_SYM
            if (!api.getUniqueId || !api.commInitRank || !api.commInitAll || !api.commDestroy || !api.reduce || !api.groupStart ||
                !api.groupEnd || !api.getErrorString) { dlclose(api.handle); api.handle = nullptr; }
        }
    }
    if (!api.handle) { err = "libnccl.so.2 could not be loaded (multi-GPU rendering needs NCCL)"; return nullptr; }
    return &api;
}

struct PtbContext {
    int device = 0;
    PtbOptions opt{};
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int numSMs = 148;

    // scene
    int W = 0, H = 0, traceDepth = 0;
    uint32_t numPixels = 0;
    int pixBits = 0;
    int batch = 1;                       // slots per wavefront
    bool useBvh = false;
    bool ext = false;                    // camera / material extensions in use: the EXT kernel instantiations run
    std::vector<int> sceneMaterial;      // scene geom index -> material id
    std::vector<int> slotOfScene;
    DevBuf<GeomRec> geoms;
    DevBuf<GeomRec> geomsScene;          // brute-force table (cubes first, then spheres), also used by ptb_stage_intersect(BRUTE)
    int numCubes = 0;
    DevBuf<BvhNode> bvh;
    int numNodes = 0;
    DevBuf<float4> mats;
    DevBuf<int> sceneMatDev;
    DevBuf<SceneHeader> hdr;
    int numGeoms = 0, numMats = 0;
    int numTriangles = 0;                // triangles of PTB_GEOM_MESH geoms
    int numPrims = 0;                    // BVH primitives: spheres + cubes + triangles
    uint8_t *pinned = nullptr;           // pinned staging for scene uploads (H2D from pinned memory)
    size_t pinnedBytes = 0;
    int64_t h2dBytes = 0;                // bytes copied host->device by the last ptb_create / ptb_update_scene

    // state
    DevBuf<float> image;
    float *imagePtr = nullptr;           // own or external
    DevBuf<float4> A[2], B[2];
    DevBuf<float2> C[2];
    DevBuf<uint16_t> L[2];               // leave codes of the paths (PathPlanes::L; brute-force scenes of the fused pipeline, PtbOptions.leaveTest)
    bool leaveWanted = false;            // ... decided at ptb_create: fused pipeline, brute force, one-launch bounce loop, PtbOptions.leaveTest
    bool leaveTest = false;              // the LEAVE kernel instantiations run: leaveWanted and the scene's matrices are finite and moderate
    bool hitQueue = false;               // bounce_loop_kernel<.., QUEUE>: PtbOptions.hitQueue resolved against the wavefront and the materials
    bool queueWavefront = false, refractive = false;      // ... at least 16 tiles of 256 paths per CTA; a material with REFR > 0 (full BSDF)
    DevBuf<float4> hitTn, fbTn;           // fbTn / fbMeta: first-bounce cache (INSTRUCTION.md:87-89)
    DevBuf<int> fbMeta;
    bool fbValid = false;
    DevBuf<int> hitMeta, sortKeys, sortOrder;
    DevBuf<uint32_t> sortHist, segCounts;
    uint32_t segCap = 0;
    bool loopBounces = false;      // fused pipeline, brute force: one launch for bounces 1 .. traceDepth-1
    bool sharedFirstHit = false;   // fused pipeline: the camera ray is iteration-independent, bounce 0 shares one hit per pixel across the slots
    DevBuf<Control> ctrl;
    DevBuf<uint8_t> rgb8;

    Params base{};
    int gridFused = 0, gridStaged = 0, gridTrace = 0;
    uint32_t slotBlock = 8;              // bounce0_shared_kernel: slots per tile
    size_t smemBrute = 0, smemBvh = 0;
    int bvhDepth = 0;                    // depth of the BVH last uploaded
    DevBuf<int> stackSpill;              // trace_kernel: traversal-stack entries beyond the shared-memory part
    int spillCap = 0;
    DevBuf<float> partial;               // ordered deposit: batch partial images
    uint32_t partialStride = 0;
    // several GPUs: communicator of this context (owned unless it belongs to a PtbMulti), the reduced frame on the root
    ncclComm_t comm = nullptr;
    bool ownComm = false;
    int commRank = 0, commRanks = 1;
    DevBuf<float> frame;
    bool frameValid = false;             // the last ptb_comm_reduce landed here (root of a > 1-rank communicator)
    // ptb_pathtrace: the caller's host image (scene->state.image) is page-locked once so that its per-iteration copy is a DMA at
    // PCIe rate; PtbOptions.lazyImage additionally snapshots the image (device to device) and copies it out on a second stream
    // while the next iteration renders
    float *pinnedHost = nullptr;
    size_t pinnedHostBytes = 0;
    bool pinHostFailed = false;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evImage[2] = {nullptr, nullptr}, evCopy[2] = {nullptr, nullptr};
    DevBuf<float> snap[2];
    int snapIdx = 0;

    cudaGraphExec_t graphExec = nullptr;
    int graphKernels = 0;                // kernels per replay of the captured batch
    bool capturing = false;
    // PtbOptions.timeStages: event pairs of the launches since the last fold, and the per-stage sums
    std::vector<cudaEvent_t> events;
    size_t evUsed = 0;
    std::vector<int> evStage;
    double stageUs[PTB_NUM_STAGES] = {};
    int64_t stageLaunches[PTB_NUM_STAGES] = {};
    int64_t launches = 0;
    int64_t hostIterations = 0;
    // ptb_pathtrace look-ahead (PtbOptions.lookahead): iterations laBase .. laBase + batch - 1 have been traced on aheadStream and set laSet of
    // the prefix images holds the image after each of them; laServed of them have been handed out; the window after it is traced into the
    // other set meanwhile (laNext).  laImage = the image after the last iteration handed out (NULL: the canonical image[] is current)
    bool laValid = false, laNext = false; int laBase = 0, laServed = 0, laSet = 0;
    const float *laImage = nullptr;
    cudaStream_t aheadStream = nullptr;
    cudaEvent_t evWave[2] = {nullptr, nullptr}, evUser = nullptr;
    DevBuf<float> prefix[2];             // batch images each

    ~PtbContext() {
        if (graphExec) cudaGraphExecDestroy(graphExec);
        if (aheadStream) { cudaStreamSynchronize(aheadStream); cudaStreamDestroy(aheadStream); }
        for (int k = 0; k < 2; ++k) { if (evWave[k]) cudaEventDestroy(evWave[k]); prefix[k].release(); }
        if (evUser) cudaEventDestroy(evUser);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        if (comm && ownComm) { std::string err; if (NcclApi *n = ncclApi(err)) n->commDestroy(comm); }
        frame.release();
        if (copyStream) { cudaStreamSynchronize(copyStream); cudaStreamDestroy(copyStream); }
        for (int k = 0; k < 2; ++k) { if (evImage[k]) cudaEventDestroy(evImage[k]); if (evCopy[k]) cudaEventDestroy(evCopy[k]); snap[k].release(); }
        if (pinnedHost) cudaHostUnregister(pinnedHost);
        geoms.release(); geomsScene.release(); bvh.release(); mats.release(); sceneMatDev.release();
        image.release();
        for (int i = 0; i < 2; ++i) { A[i].release(); B[i].release(); C[i].release(); L[i].release(); }
        stackSpill.release(); partial.release();
        hitTn.release(); hitMeta.release(); sortKeys.release(); sortOrder.release(); sortHist.release(); segCounts.release(); fbTn.release(); fbMeta.release();
        ctrl.release(); rgb8.release(); hdr.release();
        if (pinned) cudaFreeHost(pinned);
        if (ownStream && stream) cudaStreamDestroy(stream);
    }
};

namespace {

void dropGraphs(PtbContext *c) {
    if (c->graphExec) { cudaGraphExecDestroy(c->graphExec); c->graphExec = nullptr; }
}

GeomRec packGeom(const PtbGeom &g, int sceneIndex, int order = -1) {
    GeomRec r;
    auto row = [](const float *m, int k) { return make_float4(m[0 + k], m[4 + k], m[8 + k], m[12 + k]); };
    r.inv.r0 = row(g.inverseTransform, 0); r.inv.r1 = row(g.inverseTransform, 1); r.inv.r2 = row(g.inverseTransform, 2);
    r.fwd.r0 = row(g.transform, 0); r.fwd.r1 = row(g.transform, 1); r.fwd.r2 = row(g.transform, 2);
    r.nrm.r0 = row(g.invTranspose, 0); r.nrm.r1 = row(g.invTranspose, 1); r.nrm.r2 = row(g.invTranspose, 2);
    // tag.z = position in the scene's primitive order (tie rule), tag.w = scene geom index (what a hit reports)
    r.tag = make_int4(g.type, g.materialid, order < 0 ? sceneIndex : order, sceneIndex);
    // conservative world-space pruning bound (phase 1 of the brute-force closest hit)
    if (g.type == 1) {
        float lo[3], hi[3];
        geomBounds(g, lo, hi);
        r.b0 = make_float4(lo[0], hi[0], lo[1], hi[1]);          // (lo, hi) pairs per axis: operands of the packed slab test
        r.b1 = make_float4(lo[2], hi[2], 0.f, 0.f);
    } else {
        float c[3], r2, errK;
        sphereBounds(g, c, &r2, &errK);
        r.b0 = make_float4(c[0], c[1], c[2], r2);
        r.b1 = make_float4(errK, 0.f, 0.f, 0.f);
    }
    return r;
}
// One mesh triangle as a BVH primitive: inv / fwd of the mesh geom, the object-space vertices and the world-space
// normal in the nrm rows, the padded world box in b0 / b1.  The normal is normalize(invTranspose * cross(e1, e2))
// in the oracle's (glm's) float operation order -- this TU's host code is built with -ffp-contract=off.
GeomRec packTriangle(const PtbGeom &g, const PtbTriangle &t, int sceneIndex, int order) {
    GeomRec r = packGeom(g, sceneIndex, order);
    r.tag.x = PTB_GEOM_MESH;
    const float e1[3] = {t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2]};
    const float e2[3] = {t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2]};
    const float cx = e1[1] * e2[2] - e2[1] * e1[2], cy = e1[2] * e2[0] - e2[2] * e1[0], cz = e1[0] * e2[1] - e2[0] * e1[1];   // glm::cross
    const float *m = g.invTranspose;
    float n[3];
    for (int k = 0; k < 3; ++k) n[k] = (m[0 + k] * cx + m[4 + k] * cy) + (m[8 + k] * cz + m[12 + k] * 0.0f);            // mat4 * vec4(c, 0)
    const float tx = n[0] * n[0], ty = n[1] * n[1], tz = n[2] * n[2];
    const float inv = 1.0f / sqrtf(tx + ty + tz);                                                                        // glm::normalize
    r.nrm.r0 = make_float4(t.v0[0], t.v0[1], t.v0[2], n[0] * inv);
    r.nrm.r1 = make_float4(t.v1[0], t.v1[1], t.v1[2], n[1] * inv);
    r.nrm.r2 = make_float4(t.v2[0], t.v2[1], t.v2[2], n[2] * inv);
    float lo[3], hi[3];
    triangleBounds(g, t, lo, hi);
    r.b0 = make_float4(lo[0], hi[0], lo[1], hi[1]);
    r.b1 = make_float4(lo[2], hi[2], 0.f, 0.f);
    return r;
}
// brute-force table record: a cube's padded box as [lo, hi, hi, lo] per axis in its (otherwise unused) nrm rows
// (sign-selected 64-bit (near, far) pair loads in closestHitBrute)
GeomRec packGeomBrute(const PtbGeom &g, int sceneIndex) {
    GeomRec r = packGeom(g, sceneIndex);
    if (g.type == 1) {
        static_assert(offsetof(GeomRec, nrm) == kPrunePair0 * 8 && sizeof(Affine) == 12 * 4, "prune pairs");
        const float lo[3] = {r.b0.x, r.b0.z, r.b1.x}, hi[3] = {r.b0.y, r.b0.w, r.b1.y};
        float *w = reinterpret_cast<float *>(&r.nrm);
        for (int ax = 0; ax < 3; ++ax) { w[4 * ax] = lo[ax]; w[4 * ax + 1] = hi[ax]; w[4 * ax + 2] = hi[ax]; w[4 * ax + 3] = lo[ax]; }
    }
    return r;
}

Params makeParams(PtbContext *c, int inBuf) {
    Params p = c->base;
    p.in.A = c->A[inBuf].p; p.in.B = c->B[inBuf].p; p.in.C = c->C[inBuf].p;
    p.out.A = c->A[inBuf ^ 1].p; p.out.B = c->B[inBuf ^ 1].p; p.out.C = c->C[inBuf ^ 1].p;
    // leave codes: only when the LEAVE kernels run (bounce 0 then fills the segments from both ends for bounce_loop_kernel<LEAVE>)
    p.in.L = c->leaveTest ? c->L[inBuf].p : nullptr; p.out.L = c->leaveTest ? c->L[inBuf ^ 1].p : nullptr;
    return p;
}

template <typename K> int setSmem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fail(PTB_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
    }
    return PTB_OK;
}

int validateScene(const PtbSceneDesc *scene) {
    if (scene->numGeoms < 0 || scene->numMaterials <= 0 || (scene->numGeoms > 0 && !scene->geoms) || !scene->materials)
        return fail(PTB_ERR_INVALID, "scene needs geoms[%d] and materials[%d]", scene->numGeoms, scene->numMaterials);
    for (int i = 0; i < scene->numGeoms; ++i) {
        const int ty = scene->geoms[i].type;
        if (ty != 0 && ty != 1 && ty != PTB_GEOM_MESH) return fail(PTB_ERR_INVALID, "geom %d has unknown type %d", i, ty);
        if (ty == PTB_GEOM_MESH) {
            if (!scene->meshRanges || (scene->numTriangles > 0 && !scene->triangles)) return fail(PTB_ERR_INVALID, "geom %d is a mesh but the scene has no triangle arrays", i);
            const int64_t first = scene->meshRanges[2 * i], count = scene->meshRanges[2 * i + 1];
            if (first < 0 || count < 0 || first + count > scene->numTriangles)
                return fail(PTB_ERR_INVALID, "geom %d: triangle range [%lld, +%lld) outside the %d triangles", i, (long long)first, (long long)count, scene->numTriangles);
        }
        if (scene->geoms[i].materialid < 0 || scene->geoms[i].materialid >= scene->numMaterials)
            return fail(PTB_ERR_INVALID, "geom %d references material %d of %d", i, scene->geoms[i].materialid, scene->numMaterials);
    }
    return PTB_OK;
}

// Packs the scene into the device layout inside the pinned staging buffer and copies it to the
// device on the context's stream (src/pathtrace.cu:89-93 uploads the AoS structs instead).
// PtbOptions.hitQueue against what the context knows.  AUTO: the queue measured faster where a warp has many chunks per bounce (wavefronts of
// at least 16 tiles of 256 paths per CTA: cornell 800 x 800 from 32 iterations per wavefront; 2.3 % at 128, 1.7 % at 64) and slower on
// small wavefronts (1.2 % at 16) and on the glass scene (4 %) (profiles/r7a_queue_ab.txt, r7c_queue_ab.txt, r7d_queue_ab.txt)
void resolveHitQueue(PtbContext *c) {
    bool hq = c->opt.hitQueue == 1;
    if (c->opt.hitQueue == PTB_HIT_QUEUE_AUTO) hq = c->queueWavefront && !c->refractive;
    if (hq != c->hitQueue) dropGraphs(c);                                                                 // another kernel: re-capture
    c->hitQueue = hq;
}

int uploadScene(PtbContext *c, const PtbSceneDesc *scene) {
    const int G = scene->numGeoms, M = scene->numMaterials, Gp = std::max(1, G);
    // BVH primitives in scene order, the triangles of a mesh expanded in place: (geom index, triangle index or -1)
    std::vector<std::pair<int, int>> prims;
    prims.reserve((size_t)G + (size_t)std::max(0, scene->numTriangles));
    for (int i = 0; i < G; ++i) {
        if (scene->geoms[i].type == PTB_GEOM_MESH)
            for (int k = 0; k < scene->meshRanges[2 * i + 1]; ++k) prims.emplace_back(i, scene->meshRanges[2 * i] + k);
        else prims.emplace_back(i, -1);
    }
    const int P = (int)prims.size(), Pp = std::max(1, P);
    if (P != c->numPrims) return fail(PTB_ERR_INVALID, "scene has %d primitives, the context was created for %d", P, c->numPrims);
    const bool hasMeshes = P != G || scene->numTriangles > 0;
    const size_t bytesBrute = hasMeshes ? 0 : (size_t)Gp * sizeof(GeomRec);
    const size_t bytesGeoms = (size_t)Pp * sizeof(GeomRec), bytesNodes = (size_t)Pp * sizeof(BvhNode);
    const size_t bytesMats = (size_t)M * 3 * sizeof(float4), bytesSm = (size_t)Gp * sizeof(int);
    const size_t total = bytesBrute + bytesGeoms + bytesNodes + bytesMats + bytesSm + sizeof(SceneHeader);
    if (c->pinnedBytes < total) {
        if (c->pinned) cudaFreeHost(c->pinned);
        c->pinned = nullptr; c->pinnedBytes = 0;
        CU(cudaMallocHost(&c->pinned, total));
        c->pinnedBytes = total;
    }
    uint8_t *q = c->pinned;
    GeomRec *bruteOrder = reinterpret_cast<GeomRec *>(q); q += bytesBrute;
    GeomRec *leafOrder = reinterpret_cast<GeomRec *>(q); q += bytesGeoms;
    BvhNode *nodes = reinterpret_cast<BvhNode *>(q); q += bytesNodes;
    float4 *mats = reinterpret_cast<float4 *>(q); q += bytesMats;
    int *sceneMat = reinterpret_cast<int *>(q); q += bytesSm;
    SceneHeader *hdr = reinterpret_cast<SceneHeader *>(q);
    memset(c->pinned, 0, total);

    c->sceneMaterial.resize(G);
    for (int i = 0; i < G; ++i) sceneMat[i] = c->sceneMaterial[i] = scene->geoms[i].materialid;
    // brute-force table: cubes first, then spheres, each in scene order (tag.z keeps the scene index for the tie rule)
    int numCubes = 0;
    if (!hasMeshes) {
        for (int i = 0; i < G; ++i) if (scene->geoms[i].type == 1) bruteOrder[numCubes++] = packGeomBrute(scene->geoms[i], i);
        for (int i = 0, k = numCubes; i < G; ++i) if (scene->geoms[i].type != 1) bruteOrder[k++] = packGeomBrute(scene->geoms[i], i);
    }
    c->numCubes = numCubes;
    hdr->numCubes = numCubes;
    // The leave test of closestHitBrute<LEAVE> argues with finite arithmetic: it runs when the leave codes are kept (ptb_create) and every
    // inverse transform of the brute-force table is finite and moderate (|m| < 1e9: a ray the kernel accepts as finite -- |o|, |d| < 1e6 --
    // cannot overflow in the object space of any geom)
    bool leaveOk = c->leaveWanted && !hasMeshes && G <= kLeaveMaxGeoms;
    for (int i = 0; i < G && leaveOk; ++i)
        for (int k = 0; k < 16; ++k) leaveOk = leaveOk && std::fabs(scene->geoms[i].inverseTransform[k]) < 1e9f;      // false for NaN
    if (leaveOk != c->leaveTest) dropGraphs(c);                                                           // other kernels: re-capture
    c->leaveTest = leaveOk;
    // the BVH is always built so that ptb_stage_intersect can compare both traversals on any scene;
    // the render path uses it only when useBvh is set
    HostBvh hb;
    int maxLeaf = 1;                     // one primitive per leaf: its padded bound IS the leaf box (measured: 4 -> 7.28, 2 -> 6.62, 1 -> 6.29 ms per iteration of the 100k-object scene)
    if (const char *e = tuningEnv("PTB_BVH_LEAF")) maxLeaf = std::max(1, atoi(e));                       // tuning knob (experiments)
    {
        // Sphere boxes also cover the hits the reference's sphere test reports for rays passing OUTSIDE the sphere, for every ray that
        // starts inside the scene (ptb_bvh.cpp geomBounds / defaultFarPadScene; PTB_BVH_FAR_PAD=0 opts out).
        FarPadScene farPad = defaultFarPadScene(scene->geoms, G, scene->camera.position);
        std::vector<float> lo((size_t)P * 3), hi((size_t)P * 3);
        // triangles first: a mesh reaches as far as its vertices do, not as far as the unit cube under its transform, and rays start there too
        for (int k = 0; k < P; ++k) if (prims[k].second >= 0) {
            triangleBounds(scene->geoms[prims[k].first], scene->triangles[prims[k].second], &lo[3 * (size_t)k], &hi[3 * (size_t)k]);
            if (farPad.on) for (int a = 0; a < 3; ++a) {
                farPad.lo[a] = std::min(farPad.lo[a], (double)lo[3 * (size_t)k + a]);
                farPad.hi[a] = std::max(farPad.hi[a], (double)hi[3 * (size_t)k + a]);
            }
        }
        for (int k = 0; k < P; ++k) if (prims[k].second < 0) {
            const PtbGeom &g = scene->geoms[prims[k].first];
            geomBounds(g, &lo[3 * (size_t)k], &hi[3 * (size_t)k], farPadReach(g, farPad));
        }
        buildBvhBoxes(lo.data(), hi.data(), P, maxLeaf, hb);
    }
    static_assert(sizeof(HostBvhNode) == sizeof(BvhNode), "bvh node");
    if ((int)hb.nodes.size() > Pp) return fail(PTB_ERR_INVALID, "BVH has %zu nodes for %d primitives", hb.nodes.size(), P);
    // device node layout: the planes of BOTH children paired per axis (operands of the packed slab test), then the links
    for (size_t i = 0; i < hb.nodes.size(); ++i) {
        const HostBvhNode &h = hb.nodes[i];
        nodes[i].x = make_float4(h.llo[0], h.rlo[0], h.lhi[0], h.rhi[0]);
        nodes[i].y = make_float4(h.llo[1], h.rlo[1], h.lhi[1], h.rhi[1]);
        nodes[i].z = make_float4(h.llo[2], h.rlo[2], h.lhi[2], h.rhi[2]);
        // child links as the traversal pushes them: interior -> node index, leaf -> ~(first slot | count << 27); an absent child
        // (scenes of 0 or 1 primitives) gets NaN planes: every comparison of the slab test is false for it
        auto code = [](int32_t ref, int32_t count) { return count > 0 ? ~(ref | (count << 27)) : ref; };
        nodes[i].m = make_int4(code(h.lref, h.lcount), code(h.rref, h.rcount), 0, 0);
        const float qnan = std::numeric_limits<float>::quiet_NaN();
        if (h.lref < 0) { nodes[i].x.x = nodes[i].x.z = nodes[i].y.x = nodes[i].y.z = nodes[i].z.x = nodes[i].z.z = qnan; nodes[i].m.x = -1; }
        if (h.rref < 0) { nodes[i].x.y = nodes[i].x.w = nodes[i].y.y = nodes[i].y.w = nodes[i].z.y = nodes[i].z.w = qnan; nodes[i].m.y = -1; }
    }
    for (int sIdx = 0; sIdx < P; ++sIdx) {
        const int k = hb.order[sIdx];
        const PtbGeom &g = scene->geoms[prims[k].first];
        leafOrder[sIdx] = prims[k].second >= 0 ? packTriangle(g, scene->triangles[prims[k].second], prims[k].first, k)
                                               : packGeom(g, prims[k].first, k);
    }
    c->numNodes = (int)hb.nodes.size();
    c->bvhDepth = hb.maxDepth;
    for (int i = 0; i < M; ++i) {
        const PtbMaterial &m = scene->materials[i];
        mats[3 * i + 0] = make_float4(m.color[0], m.color[1], m.color[2], m.emittance);
        mats[3 * i + 1] = make_float4(m.specularColor[0], m.specularColor[1], m.specularColor[2], m.hasReflective);
        mats[3 * i + 2] = make_float4(m.hasRefractive, m.indexOfRefraction, m.specularExponent, 0.f);
    }
    const PtbCamera &cam = scene->camera;
    hdr->cam = CameraParams{cam.position[0], cam.position[1], cam.position[2], cam.view[0], cam.view[1], cam.view[2],
                            cam.up[0], cam.up[1], cam.up[2], cam.right[0], cam.right[1], cam.right[2],
                            cam.pixelLength[0], cam.pixelLength[1], c->W, c->H,
                            c->opt.antialias ? 1 : 0, c->opt.lensRadius > 0.0f ? c->opt.lensRadius : 0.0f, c->opt.focalDistance};
    // the EXT instantiations are needed for a jittered / thin-lens camera and for imperfect-specular materials
    bool ext = c->opt.antialias || c->opt.lensRadius > 0.0f;
    if (c->opt.bsdfMode == PTB_BSDF_FULL)
        for (int i = 0; i < M; ++i) ext = ext || (scene->materials[i].hasReflective > 0.0f && scene->materials[i].specularExponent > 0.0f);
    if (ext != c->ext) dropGraphs(c);                                                                     // other kernels: re-capture
    c->ext = ext;
    c->refractive = false;
    if (c->opt.bsdfMode == PTB_BSDF_FULL) for (int i = 0; i < M; ++i) c->refractive = c->refractive || scene->materials[i].hasRefractive > 0.0f;
    resolveHitQueue(c);
    hdr->numNodes = c->numNodes;

    if (bytesBrute) CU(cudaMemcpyAsync(c->geomsScene.p, bruteOrder, bytesBrute, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->geoms.p, leafOrder, bytesGeoms, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->bvh.p, nodes, bytesNodes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->mats.p, mats, bytesMats, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->sceneMatDev.p, sceneMat, bytesSm, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->hdr.p, hdr, sizeof(SceneHeader), cudaMemcpyHostToDevice, c->stream));
    c->h2dBytes = (int64_t)total;
    return PTB_OK;
}

// ---- launch sequences -------------------------------------------------------------------------
// Every kernel of a batch goes out through LAUNCH: it counts the launch and, with PtbOptions.timeStages, brackets it with a pair
// of CUDA events whose elapsed time is summed per stage (PtbStats.stageUs) the next time the stream is synchronised.
struct StageScope {
    PtbContext *c; bool on;
    StageScope(PtbContext *ctx, int stage) : c(ctx), on(ctx->opt.timeStages && !ctx->capturing) {
        if (!on) return;
        if (c->evUsed + 2 > c->events.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { on = false; return; }
            c->events.push_back(a); c->events.push_back(b);
        }
        c->evStage.push_back(stage);
        cudaEventRecord(c->events[c->evUsed], c->stream);
    }
    ~StageScope() {
        if (!on) return;
        cudaEventRecord(c->events[c->evUsed + 1], c->stream);
        c->evUsed += 2;
    }
};
#define LAUNCH(stage, kernel, grid, block, smem, ...)                                              \
    do {                                                                                           \
        StageScope stageScope_(c, stage);                                                          \
        kernel<<<grid, block, smem, c->stream>>>(__VA_ARGS__);                                     \
        c->launches++;                                                                             \
    } while (0)

// folds the recorded event pairs into the per-stage sums (the stream must be idle)
void foldStageTimers(PtbContext *c) {
    for (size_t i = 0; i + 1 < c->evUsed; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->events[i], c->events[i + 1]) == cudaSuccess) {
            const int stg = c->evStage[i / 2];
            c->stageUs[stg] += (double)ms * 1e3;
            c->stageLaunches[stg]++;
        }
    }
    c->evUsed = 0;
    c->evStage.clear();
}

// fused: traceDepth bounce kernels + the advance kernel.  Buffers ping-pong per bounce.
template <bool EXT, bool MESH>
void launchFusedBounces(PtbContext *c) {
    const size_t smem = c->useBvh ? c->smemBvh : c->smemBrute;
    for (int d = 0; d < c->traceDepth; ++d) {
        // bounce d reads buffer (d-1)&1... bounce 0 reads nothing and writes buffer 0
        Params p = makeParams(c, d == 0 ? 1 : ((d - 1) & 1));
        if (d == 0 && c->sharedFirstHit) {
            // the camera ray does not depend on the iteration: one closest hit per pixel serves every slot of the wavefront
            if (c->useBvh) {
                if (p.fbTn == nullptr) LAUNCH(PTB_STAGE_INTERSECT, (TRACE_KERNEL<SRC_PIXELS, false, MESH>), c->gridTrace, kTraceBlock, smem, p, 0, TraceIO{});
                LAUNCH(PTB_STAGE_SHADE, (bounce0_shared_kernel<true, EXT>), c->gridFused, kBlock, 0, p);
            } else LAUNCH(PTB_STAGE_FUSED_BOUNCE0, (bounce0_shared_kernel<false, EXT>), c->gridFused, kBlock, smem, p);
            continue;
        }
        if (c->useBvh) {
            // closest hit first (persistent threads, R rays multiplexed per lane), then shade / scatter / compact
            const bool cached = d == 0 && p.fbTn != nullptr;
            if (!cached) {
                if (d == 0) LAUNCH(PTB_STAGE_INTERSECT, (TRACE_KERNEL<SRC_CAMERA, EXT, MESH>), c->gridTrace, kTraceBlock, smem, p, d, TraceIO{});
                else        LAUNCH(PTB_STAGE_INTERSECT, (TRACE_KERNEL<SRC_SEGMENTS, false, MESH>), c->gridTrace, kTraceBlock, smem, p, d, TraceIO{});
            }
            if (d == 0) LAUNCH(PTB_STAGE_SHADE, (bounce_kernel<true, true, EXT>), c->gridFused, kBlock, 0, p, d);
            else if (c->opt.regroupMaterials && c->numMats <= kMaxClassTable) LAUNCH(PTB_STAGE_SHADE, (bounce_regroup_kernel<EXT>), c->gridFused, kBlock, 0, p, d);
            else        LAUNCH(PTB_STAGE_SHADE, (bounce_kernel<false, true, EXT>), c->gridFused, kBlock, 0, p, d);
        } else {
            if (d == 0) {
                if (c->leaveTest) LAUNCH(PTB_STAGE_FUSED_BOUNCE0, (bounce_kernel<true, false, EXT, true>), c->gridFused, kBlock, smem, p, d);
                else              LAUNCH(PTB_STAGE_FUSED_BOUNCE0, (bounce_kernel<true, false, EXT>), c->gridFused, kBlock, smem, p, d);
            } else if (c->loopBounces) {
                // bounces 1 .. traceDepth-1 in one launch: every CTA walks its own segment through them
                const bool one = c->numGeoms <= 32;               // a single chunk of the brute-force table
#define PTB_LOOP(L, O, Q) LAUNCH(PTB_STAGE_FUSED_BOUNCES, (bounce_loop_kernel<EXT, L, O, Q>), c->gridFused, kBlock, smem, p, d)
                if (c->hitQueue) { if (c->leaveTest) { if (one) PTB_LOOP(true, true, true); else PTB_LOOP(true, false, true); }
                                   else              { if (one) PTB_LOOP(false, true, true); else PTB_LOOP(false, false, true); } }
                else             { if (c->leaveTest) { if (one) PTB_LOOP(true, true, false); else PTB_LOOP(true, false, false); }
                                   else              { if (one) PTB_LOOP(false, true, false); else PTB_LOOP(false, false, false); } }
#undef PTB_LOOP
                break;
            }
            else        LAUNCH(PTB_STAGE_FUSED_BOUNCES, (bounce_kernel<false, false, EXT>), c->gridFused, kBlock, smem, p, d);
        }
    }
}
// end of a batch: the ordered deposit's partial images into image[] (slot = iteration order), then the control block
void launchBatchTail(PtbContext *c) {
    if (c->base.partial) LAUNCH(PTB_STAGE_GATHER, deposit_reduce_kernel, c->gridStaged, kBlock, 0, c->base);
    LAUNCH(PTB_STAGE_CONTROL, advance_kernel, 1, 32, 0, c->ctrl.p, c->traceDepth, c->numPixels);
}
int launchFusedBatch(PtbContext *c) {
    const bool mesh = c->numTriangles > 0;
    if (c->ext) { if (mesh) launchFusedBounces<true, true>(c); else launchFusedBounces<true, false>(c); }
    else        { if (mesh) launchFusedBounces<false, true>(c); else launchFusedBounces<false, false>(c); }
    launchBatchTail(c);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PTB_ERR_CUDA, "fused batch launch: %s", cudaGetErrorString(e));
    return PTB_OK;
}

// staged: raygen, then per bounce intersect -> [sort] -> shade -> [compact], then gather.
int launchStagedBatch(PtbContext *c) {
    const size_t smem = c->useBvh ? c->smemBvh : c->smemBrute;
    const int grid = c->gridStaged;
    int cur = 0;
    Params p = makeParams(c, cur);
    LAUNCH(PTB_STAGE_RAYGEN, raygen_kernel, grid, kBlock, 0, p);
    const int compaction = c->opt.compaction;
    for (int d = 0; d < c->traceDepth; ++d) {
        p = makeParams(c, cur);
        if (c->useBvh) {
            // raw hit records through the BVH, then the deferred normals in place
            LAUNCH(PTB_STAGE_INTERSECT, (TRACE_KERNEL<SRC_PATHS, false, true>), c->gridTrace, kTraceBlock, smem, p, d, TraceIO{});
            LAUNCH(PTB_STAGE_INTERSECT, resolve_hits_kernel, grid, kBlock, 0, p, &c->ctrl.p->counts[d], 0u, p.hits.tn, p.hits.meta);
        } else LAUNCH(PTB_STAGE_INTERSECT, intersect_kernel, grid, kBlock, smem, p, d);
        const int *order = nullptr;
        if (c->opt.sortMaterials) {
            const int numKeys = c->numMats + 1;
            LAUNCH(PTB_STAGE_SORT, sort_keys_kernel, grid, kBlock, 0, p, d, c->sortKeys.p, c->sortHist.p, numKeys);
            LAUNCH(PTB_STAGE_SORT, sort_hist_dev_kernel, grid, kBlock, sizeof(uint32_t) * numKeys, c->sortKeys.p, c->ctrl.p, d, numKeys, c->sortHist.p);
            LAUNCH(PTB_STAGE_SORT, sort_scan_kernel, 1, 32, 0, c->sortHist.p, numKeys);
            LAUNCH(PTB_STAGE_SORT, sort_scatter_dev_kernel, grid, kBlock, 0, c->sortKeys.p, c->ctrl.p, d, c->sortHist.p, c->sortOrder.p, 0u, numKeys);
            order = c->sortOrder.p;
        }
        LAUNCH(PTB_STAGE_SHADE, shade_kernel, grid, kBlock, 0, p, d, order, compaction ? 1 : 0);
        if (compaction) {
            LAUNCH(PTB_STAGE_COMPACT, compact_kernel, grid, kBlock, 0, p, d);
            cur ^= 1;
        } else {
            LAUNCH(PTB_STAGE_COMPACT, carry_count_kernel, 1, 32, 0, c->ctrl.p, d);
        }
    }
    if (!compaction) {
        p = makeParams(c, cur);
        LAUNCH(PTB_STAGE_GATHER, gather_kernel, grid, kBlock, 0, p);
    }
    launchBatchTail(c);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PTB_ERR_CUDA, "staged batch launch: %s", cudaGetErrorString(e));
    return PTB_OK;
}

int launchBatch(PtbContext *c) {
    return c->opt.pipeline == PTB_PIPELINE_STAGED ? launchStagedBatch(c) : launchFusedBatch(c);
}

int buildGraph(PtbContext *c) {
    if (c->graphExec) return PTB_OK;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int64_t before = c->launches;
    c->capturing = true;
    int rc = launchBatch(c);
    c->capturing = false;
    c->graphKernels = (int)(c->launches - before);    // kernels one replay of the graph launches
    c->launches = before;                         // capture does not execute anything
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != PTB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(PTB_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&c->graphExec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(PTB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    return PTB_OK;
}

// (re)fills the first-bounce cache after a scene / camera upload
int ensureFirstBounceCache(PtbContext *c) {
    if (!c->opt.firstBounceCache || c->opt.pipeline != PTB_PIPELINE_FUSED || c->fbValid) return PTB_OK;
    Params p = c->base;
    if (c->useBvh) {
        // p.hits is free between batches: raw hit records of the pixels' camera rays there, then normals into the cache
        TRACE_KERNEL<SRC_PIXELS, false, true><<<c->gridTrace, kTraceBlock, c->smemBvh, c->stream>>>(p, 0, TraceIO{});
        resolve_hits_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(p, nullptr, c->numPixels, c->fbTn.p, c->fbMeta.p);
        c->launches++;
    } else fb_cache_kernel<<<c->gridStaged, kBlock, c->smemBrute, c->stream>>>(p, c->fbTn.p, c->fbMeta.p);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PTB_ERR_CUDA, "first-bounce cache fill: %s", cudaGetErrorString(e));
    c->fbValid = true;
    return PTB_OK;
}

// trace_kernel keeps kStackSmem traversal-stack entries per lane in shared memory (the stack never holds more entries than the
// tree has levels); a deeper tree gets a global spill area its lanes move blocks of kSpillBlock old entries to and from
// (trace_pool_kernel: the entries beyond kPoolStack live there one by one)
int ensureStackSpill(PtbContext *c) {
#if PTB_TRACE_POOL
    const int need = std::max(0, c->bvhDepth + 2 - kTraceStackSmem);
#else
    const int need = c->bvhDepth + 2 <= kTraceStackSmem ? 0 : (c->bvhDepth + 2 + kSpillBlock - 1) / kSpillBlock * kSpillBlock;
#endif
    if (need > c->spillCap) {
        CU(cudaStreamSynchronize(c->stream));
        CU(c->stackSpill.alloc((size_t)c->gridTrace * kTraceBlock * kTraceSlotsPerThread * need * 2));      // 8-byte entries (child, entry distance)
        c->spillCap = need;
        dropGraphs(c);                                                                        // new launch parameters: re-capture
    }
    c->base.stackSpill = c->stackSpill.p; c->base.spillCap = c->spillCap;
    return PTB_OK;
}

// End of the look-ahead: whatever was traced ahead and not handed out is dropped, and the image after the last iteration handed out becomes
// the canonical image[] again -- before anything else renders, resets or reads image[]
int flushLookahead(PtbContext *c) {
    if (!c->laValid) return PTB_OK;
    c->laValid = false; c->laNext = false;
    CU(cudaStreamSynchronize(c->aheadStream));                     // a window may still be tracing (its partial images are all-zero again when it ends)
    if (c->laImage) CU(cudaMemcpyAsync(c->imagePtr, c->laImage, (size_t)c->numPixels * 3 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    c->laImage = nullptr;
    return PTB_OK;
}

int runBatches(PtbContext *c, int iterBegin, int count, int iterStride, int deferReduce = 0) {
    { int rc = ensureFirstBounceCache(c); if (rc != PTB_OK) return rc; }
    int done = 0;
    while (done < count) {
        const int b = std::min(c->batch, count - done);
        LAUNCH(PTB_STAGE_CONTROL, set_control_kernel, 1, 32, 0, c->ctrl.p, iterBegin + done * iterStride, iterStride, b, deferReduce);
        // full batches replay back to back: advance_kernel moves iterBase on the device
        int reps = (b == c->batch) ? (count - done) / c->batch : 1;
        for (int r = 0; r < reps; ++r) {
            if (c->opt.useGraphs && !c->opt.timeStages) {
                int rc = buildGraph(c);
                if (rc != PTB_OK) return rc;
                CU(cudaGraphLaunch(c->graphExec, c->stream));
                c->launches += c->graphKernels;
            } else {
                int rc = launchBatch(c);
                if (rc != PTB_OK) return rc;
            }
        }
        done += reps * b;
        if (c->opt.timeStages && c->evUsed > 8192) { CU(cudaStreamSynchronize(c->stream)); foldStageTimers(c); }   // bound the event pool
    }
    if (!deferReduce) c->hostIterations += count;
    return PTB_OK;
}

int ensureLookahead(PtbContext *c) {
    if (c->aheadStream) return PTB_OK;
    int least = 0, greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CU(cudaStreamCreateWithPriority(&c->aheadStream, cudaStreamNonBlocking, least));
    for (int k = 0; k < 2; ++k) {
        CU(cudaEventCreateWithFlags(&c->evWave[k], cudaEventDisableTiming));
        CU(c->prefix[k].alloc((size_t)c->batch * c->partialStride));
    }
    CU(cudaEventCreateWithFlags(&c->evUser, cudaEventDisableTiming));
    return PTB_OK;
}

// One wavefront of `batch` iterations starting at `iter` on aheadStream, then the image after each of them into set `set` of the prefix
// images, starting from `base`.  Ordered after everything queued on the caller's stream so far (copies out of that set, resets, uploads).
int launchWindow(PtbContext *c, int iter, int set, const float *base) {
    CU(cudaEventRecord(c->evUser, c->stream));
    CU(cudaStreamWaitEvent(c->aheadStream, c->evUser, 0));
    std::swap(c->stream, c->aheadStream);
    int rc = runBatches(c, iter, c->batch, 1, 1);
    if (rc == PTB_OK) {
        deposit_prefix_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(c->base, base, c->prefix[set].p, (uint32_t)c->batch);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = fail(PTB_ERR_CUDA, "look-ahead window launch failed");
    }
    std::swap(c->stream, c->aheadStream);
    if (rc != PTB_OK) return rc;
    CU(cudaEventRecord(c->evWave[set], c->aheadStream));
    return PTB_OK;
}

} // namespace

// =================================================================================================
extern "C" {

const char *ptb_last_error(void) { return g_lastError.c_str(); }
// for the host-only translation units of the library (ptb_bvh.cpp, host/host_abi.cpp): same thread-local message
int ptb_set_last_error(int code, const char *msg) { g_lastError = msg ? msg : ""; return code; }
int ptb_abi_version(void) { return PTB_ABI_VERSION; }

void ptb_default_options(PtbOptions *o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->device = -1;
    o->bsdfMode = PTB_BSDF_FULL;
    o->pipeline = PTB_PIPELINE_FUSED;
    o->accel = PTB_ACCEL_AUTO;
    o->batchIters = 0;
    o->compaction = 1;
    o->sortMaterials = 0;
    o->firstBounceCache = 0;
    o->useGraphs = 1;
    o->regroupMaterials = PTB_REGROUP_DEFAULT;
    o->lookahead = 1;                    // ptb_pathtrace traces whole wavefronts ahead (bit-identical images; only that entry point is affected)
    o->leaveTest = 1;                    // brute-force scenes: the cube face a bounce ray starts on is settled with one row of the object-space ray
    o->hitQueue = PTB_HIT_QUEUE_AUTO;    // brute-force scenes: hits queued between search and shading where that measured faster
}

int ptb_create(PtbContext **out, const PtbSceneDesc *scene, const PtbOptions *optIn) {
    if (!out || !scene) return fail(PTB_ERR_INVALID, "ptb_create: null argument");
    *out = nullptr;
    PtbOptions opt;
    if (optIn) opt = *optIn; else ptb_default_options(&opt);
    { int vrc = validateScene(scene); if (vrc != PTB_OK) return vrc; }
    const int W = scene->camera.resolution[0], H = scene->camera.resolution[1];
    if (W <= 0 || H <= 0 || (int64_t)W * H > (1ll << 30)) return fail(PTB_ERR_INVALID, "ptb_create: bad resolution %dx%d", W, H);
    if (scene->traceDepth < 1 || scene->traceDepth > 64) return fail(PTB_ERR_INVALID, "ptb_create: traceDepth %d outside [1,64]", scene->traceDepth);
    if (opt.firstBounceCache && (opt.antialias || opt.lensRadius > 0.0f))
        return fail(PTB_ERR_UNSUPPORTED, "ptb_create: firstBounceCache cannot be combined with antialias / lensRadius (camera rays change every iteration)");
    if (opt.lensRadius > 0.0f && !(opt.focalDistance > 0.0f)) return fail(PTB_ERR_INVALID, "ptb_create: lensRadius > 0 needs focalDistance > 0");

    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(PTB_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
    int dev = opt.device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(PTB_ERR_INVALID, "device %d of %d", dev, ndev);
    ON_DEVICE(dev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));

    PtbContext *c = new PtbContext();
    c->device = dev;
    c->opt = opt;
    c->numSMs = prop.multiProcessorCount;
    c->W = W; c->H = H; c->traceDepth = scene->traceDepth;
    c->numPixels = (uint32_t)W * (uint32_t)H;
    c->numGeoms = scene->numGeoms; c->numMats = scene->numMaterials;
    c->numTriangles = 0;
    for (int i = 0; i < scene->numGeoms; ++i) if (scene->geoms[i].type == PTB_GEOM_MESH) c->numTriangles += scene->meshRanges[2 * i + 1];
    c->numPrims = scene->numGeoms + c->numTriangles;
    for (int i = 0; i < scene->numGeoms; ++i) if (scene->geoms[i].type == PTB_GEOM_MESH) c->numPrims--;   // a mesh geom is its triangles
    const bool hasMeshes = c->numPrims != scene->numGeoms || c->numTriangles > 0;
    if (c->numPrims >= (1 << 27)) { const int np = c->numPrims; delete c; return fail(PTB_ERR_UNSUPPORTED, "ptb_create: %d primitives (the BVH leaf encoding holds 2^27)", np); }
    if (hasMeshes && opt.accel == PTB_ACCEL_BRUTE) { delete c; return fail(PTB_ERR_UNSUPPORTED, "ptb_create: scenes with meshes are traced through the BVH (accel = PTB_ACCEL_BRUTE refused)"); }
    int bits = 1;
    while ((1ull << bits) < c->numPixels) ++bits;
    c->pixBits = bits;
    const int maxSlots = 1 << std::min(31 - bits, 12);
    // default wavefront: a power of two between 1 and 64 iterations of about 140 M paths (~112 B per path with both path buffers, the
    // BVH hit records and the partial images: 16 GB) when the device has the room, about 40 M otherwise.  Bigger wavefronts amortise
    // the tails of the per-bounce launches: 100 k-object scene at 1920x1080, 8 / 16 / 32 / 64 iterations per wavefront:
    // 2 454 / 2 265 / 2 171 / 2 111 us per iteration (profiles/r2q_wavefront_size.txt)
    size_t freeBytes = 0, totalBytes = 0;
    if (cudaMemGetInfo(&freeBytes, &totalBytes) != cudaSuccess) { cudaGetLastError(); freeBytes = 0; }
    const uint64_t pathBudget = freeBytes >= (64ull << 30) ? 140000000ull : 40000000ull;
    int autoBatch = 1;
    while (autoBatch < 64 && (uint64_t)(autoBatch * 2) * c->numPixels <= pathBudget) autoBatch <<= 1;
    int batch = opt.batchIters > 0 ? opt.batchIters : autoBatch;
    if (opt.pipeline == PTB_PIPELINE_STAGED && opt.batchIters <= 0) batch = 1;
    // keep the wavefront within 2^31 paths and a sane memory budget (2 x 40 B per path)
    while (batch > 1 && ((uint64_t)batch * c->numPixels > (1ull << 30) || batch > maxSlots)) batch >>= 1;
    c->batch = std::max(1, batch);

    auto bail = [&](int rc) { delete c; return rc; };
#define CUB(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return bail(fail(PTB_ERR_CUDA, "CUDA error (%s:%d): %s: %s", "ptb200.cu", __LINE__, #call, cudaGetErrorString(e_))); \
    } while (0)

    if (opt.stream) { c->stream = (cudaStream_t)opt.stream; c->ownStream = false; }
    else {
        // the library's own stream gets the highest priority: with PtbOptions.lookahead the per-call passes queued on it must not wait behind
        // the wavefront that is being traced ahead on the low-priority stream (priorities only matter between streams of this process)
        int least = 0, greatest = 0;
        CUB(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CUB(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, greatest)); c->ownStream = true;
    }

    // ---- scene upload (src/pathtrace.cu:89-93) in the device layout ----
    c->useBvh = hasMeshes || opt.accel == PTB_ACCEL_BVH || (opt.accel == PTB_ACCEL_AUTO && scene->numGeoms > 64);
    const size_t bruteBytes = hasMeshes ? (size_t)201 * 1024 : (size_t)std::max(1, scene->numGeoms) * sizeof(GeomRec);   // meshes: no brute-force table
    if (!c->useBvh && bruteBytes > 200 * 1024) c->useBvh = true;
    c->leaveWanted = opt.leaveTest && opt.pipeline == PTB_PIPELINE_FUSED && !c->useBvh && !opt.launchPerBounce;
    {
        const size_t Gp = (size_t)std::max(1, scene->numGeoms), Pp = (size_t)std::max(1, c->numPrims);
        CUB(c->geomsScene.alloc(hasMeshes ? 1 : Gp)); CUB(c->geoms.alloc(Pp)); CUB(c->bvh.alloc(Pp)); CUB(c->sceneMatDev.alloc(Gp));
        CUB(c->mats.alloc((size_t)scene->numMaterials * 3)); CUB(c->hdr.alloc(1));
        int urc = uploadScene(c, scene);
        if (urc != PTB_OK) return bail(urc);
    }

    // ---- launch geometry ----
    c->smemBrute = bruteBytes;
    c->smemBvh = traceKernelSmem();
    if (bruteBytes <= 200 * 1024) {                  // brute force is only launchable when the geoms fit in shared memory
        int rc = PTB_OK;
        rc |= setSmem(bounce_kernel<true, false, false>, c->smemBrute); rc |= setSmem(bounce_kernel<false, false, false>, c->smemBrute);
        rc |= setSmem(bounce_kernel<true, false, true>, c->smemBrute); rc |= setSmem(bounce_kernel<false, false, true>, c->smemBrute);
#define PTB_LOOP_SMEM(E, L, O) rc |= setSmem(bounce_loop_kernel<E, L, O, false>, c->smemBrute); rc |= setSmem(bounce_loop_kernel<E, L, O, true>, c->smemBrute)
        PTB_LOOP_SMEM(false, false, false); PTB_LOOP_SMEM(true, false, false); PTB_LOOP_SMEM(false, true, false); PTB_LOOP_SMEM(true, true, false);
        PTB_LOOP_SMEM(false, false, true); PTB_LOOP_SMEM(true, false, true); PTB_LOOP_SMEM(false, true, true); PTB_LOOP_SMEM(true, true, true);
#undef PTB_LOOP_SMEM
        rc |= setSmem(bounce_kernel<true, false, false, true>, c->smemBrute); rc |= setSmem(bounce_kernel<true, false, true, true>, c->smemBrute);
        rc |= setSmem(bounce0_shared_kernel<false, false>, c->smemBrute); rc |= setSmem(bounce0_shared_kernel<false, true>, c->smemBrute);
        rc |= setSmem(intersect_kernel, c->smemBrute); rc |= setSmem(stage_intersect_kernel, c->smemBrute);
        rc |= setSmem(fb_cache_kernel, c->smemBrute);
        if (rc != PTB_OK) return bail(rc);
    }
    {
        int rc = PTB_OK;
        rc |= setSmem(TRACE_KERNEL<SRC_CAMERA, false, false>, c->smemBvh); rc |= setSmem(TRACE_KERNEL<SRC_CAMERA, true, false>, c->smemBvh);
        rc |= setSmem(TRACE_KERNEL<SRC_CAMERA, false, true>, c->smemBvh); rc |= setSmem(TRACE_KERNEL<SRC_CAMERA, true, true>, c->smemBvh);
        rc |= setSmem(TRACE_KERNEL<SRC_SEGMENTS, false, false>, c->smemBvh); rc |= setSmem(TRACE_KERNEL<SRC_SEGMENTS, false, true>, c->smemBvh);
        rc |= setSmem(TRACE_KERNEL<SRC_PATHS, false, true>, c->smemBvh); rc |= setSmem(TRACE_KERNEL<SRC_PIXELS, false, true>, c->smemBvh);
        rc |= setSmem(TRACE_KERNEL<SRC_PIXELS, false, false>, c->smemBvh);
        rc |= setSmem(TRACE_KERNEL<SRC_RAYS, false, true>, c->smemBvh);
        if (rc != PTB_OK) return bail(rc);
    }
    int occ = 0, occTrace = 0;
    CUB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occTrace, TRACE_KERNEL<SRC_SEGMENTS, false, true>, kTraceBlock, c->smemBvh));
    // persistent grid of the BVH traversal: one wave of resident CTAs, each serving several CTA-owned path segments back to back
    c->gridTrace = c->numSMs * std::max(1, occTrace);
    if (c->useBvh) CUB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bounce_kernel<false, true, false>, kBlock, 0));
    else           CUB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bounce_kernel<false, false, false>, kBlock, c->smemBrute));
    // brute-force scenes: bounces 1.. in one launch (bounce_loop_kernel), every CTA walking its own segment through them
    c->loopBounces = !c->useBvh && !opt.launchPerBounce;
    if (const char *e = tuningEnv("PTB_LOOP_BOUNCES")) c->loopBounces = c->loopBounces && atoi(e) != 0;   // tuning knob (experiments)
    // Several CTAs per resident slot: the CTA-owned segments cannot be rebalanced inside a launch, so a few waves of
    // smaller CTAs keep the tail short.  Measured per cornell iteration, one launch per bounce: x1 142 us, x2 138, x4 134.5,
    // x8 135.4, x16 142 (round 1, first build); with the one-launch bounce loop the per-CTA set-up is paid once per
    // wavefront and x8 wins: x2 119.0, x4 114.8, x8 112.9, x16 112.7 (profiles/r01n_loop_ab.txt)
    const size_t wavefrontTiles = ((size_t)c->numPixels * c->batch + kBlock - 1) / kBlock;
    // (one iteration per wavefront -- the per-iteration pathtrace() call: 2 500 tiles at 800 x 800 -- x1 147.6 us, x2 158.2, x4 / x8 176.4,
    // profiles/r2i_b1_grid.txt: the multiplier is the largest that still leaves every CTA four tiles)
    int gridMult = 1;
    if (c->useBvh) gridMult = 4;
    else for (int mlt : {8, 4, 2}) if (wavefrontTiles >= (size_t)c->numSMs * std::max(1, occ) * mlt * 4) { gridMult = mlt; break; }
    if (!c->loopBounces && !c->useBvh) gridMult = std::min(gridMult, 4);
    if (const char *e = tuningEnv("PTB_GRID_MULT")) gridMult = std::max(1, atoi(e));           // tuning knob (experiments)
    c->gridFused = c->numSMs * std::max(1, occ) * gridMult;
    c->gridFused = std::min(c->gridFused, c->gridTrace * kMaxVirtualSegs);      // trace_kernel serves at most kMaxVirtualSegs segments per CTA
    if (const char *e = tuningEnv("PTB_GRID_FUSED")) c->gridFused = std::max(1, atoi(e));       // tuning knob (experiments): the grid itself
    c->gridStaged = c->numSMs * 8;
    c->queueWavefront = wavefrontTiles >= (size_t)c->gridFused * 16;
    resolveHitQueue(c);

    // ---- state (src/pathtrace.cu:84-87, 95-96) ----
    if (opt.externalAccum) c->imagePtr = opt.externalAccum;
    else { CUB(c->image.alloc((size_t)c->numPixels * 3)); c->imagePtr = c->image.p; CUB(cudaMemset(c->imagePtr, 0, (size_t)c->numPixels * 3 * sizeof(float))); }
    const size_t maxPaths = (size_t)c->numPixels * c->batch;
    c->sharedFirstHit = opt.pipeline == PTB_PIPELINE_FUSED && !opt.antialias && !(opt.lensRadius > 0.0f);
    // fused pipeline: every CTA owns a segment of the path buffers that holds the survivors of the tiles it gets at bounce 0:
    // ceil(tiles / grid) tiles of 256 camera paths (bounce_kernel<FIRST>), or of 32 pixels x min(batch, slotBlock) slots (bounce0_shared_kernel)
    c->slotBlock = c->useBvh ? kSlotBlockBvh : kSlotBlockBrute;
    const size_t slotBlock = std::min<size_t>(c->batch, c->slotBlock);
    const size_t sharedTiles = (((size_t)c->numPixels + 31) / 32) * (((size_t)c->batch + slotBlock - 1) / slotBlock);
    const size_t pathTiles = (maxPaths + kBlock - 1) / kBlock;
    const size_t capShared = ((sharedTiles + c->gridFused - 1) / c->gridFused) * 32 * slotBlock;
    const size_t capTiles = ((pathTiles + c->gridFused - 1) / c->gridFused) * kBlock;
    c->segCap = (uint32_t)std::max(capShared, capTiles);
    const size_t pathSlots = std::max(maxPaths, (size_t)c->segCap * c->gridFused);
    for (int i = 0; i < 2; ++i) {
        CUB(c->A[i].alloc(pathSlots)); CUB(c->B[i].alloc(pathSlots)); CUB(c->C[i].alloc(pathSlots));
        if (c->leaveWanted) CUB(c->L[i].alloc(pathSlots));
    }
    CUB(c->segCounts.alloc((size_t)3 * c->gridFused));          // survivors per CTA, ping-pong by depth parity; [2 * grid ..): back of the segments after bounce 0 (LEAVE)
    CUB(cudaMemset(c->segCounts.p, 0, sizeof(uint32_t) * 3 * c->gridFused));
    if (opt.pipeline == PTB_PIPELINE_STAGED || c->useBvh) {
        CUB(c->hitTn.alloc(pathSlots)); CUB(c->hitMeta.alloc(pathSlots));
        if (opt.sortMaterials) { CUB(c->sortKeys.alloc(maxPaths)); CUB(c->sortOrder.alloc(maxPaths)); CUB(c->sortHist.alloc(scene->numMaterials + 2)); }
    }
    if (opt.firstBounceCache && opt.pipeline == PTB_PIPELINE_FUSED) { CUB(c->fbTn.alloc(c->numPixels)); CUB(c->fbMeta.alloc(c->numPixels)); }
    CUB(c->ctrl.alloc(1));
    CUB(cudaMemset(c->ctrl.p, 0, sizeof(Control)));
    // ordered deposit (default): one partial image per slot of the wavefront, all-zero between wavefronts
    if (c->batch > 1 && opt.depositMode == PTB_DEPOSIT_ORDERED) {
        if (reinterpret_cast<uintptr_t>(c->imagePtr) & 15u) return bail(fail(PTB_ERR_INVALID, "ptb_create: externalAccum must be 16-byte aligned"));
        c->partialStride = (c->numPixels * 3u + 3u) & ~3u;
        CUB(c->partial.alloc((size_t)c->batch * c->partialStride));
        CUB(cudaMemset(c->partial.p, 0, (size_t)c->batch * c->partialStride * sizeof(float)));
    }
    { int rc = ensureStackSpill(c); if (rc != PTB_OK) return bail(rc); }

    Params &p = c->base;
    memset(&p, 0, sizeof p);
    p.hdr = c->hdr.p;
    p.geoms = c->useBvh ? c->geoms.p : c->geomsScene.p;
    p.numGeoms = scene->numGeoms;
    p.bvh = c->bvh.p;
    p.mats = c->mats.p; p.numMats = scene->numMaterials;
    p.image = c->imagePtr;
    p.ctrl = c->ctrl.p;
    p.hits.tn = c->hitTn.p; p.hits.meta = c->hitMeta.p;
    p.traceDepth = c->traceDepth;
    p.bsdfMode = opt.bsdfMode;
    p.pixBits = c->pixBits;
    p.numPixels = c->numPixels;
    p.atomicGather = (c->batch > 1 && opt.depositMode != PTB_DEPOSIT_ORDERED) ? 1 : 0;
    p.segCounts = c->segCounts.p;
    p.segCap = c->segCap; p.slotBlock = c->slotBlock;
    p.vgrid = (uint32_t)c->gridFused;
    p.stackSpill = c->stackSpill.p; p.spillCap = c->spillCap;
    p.partial = c->partial.p; p.partialStride = c->partialStride;
    p.fbTn = c->fbTn.p; p.fbMeta = c->fbMeta.p;          // NULL when the cache is off
    CUB(cudaDeviceSynchronize());
#undef CUB
    *out = c;
    return PTB_OK;
}

int ptb_update_scene(PtbContext *c, const PtbSceneDesc *scene) {
    if (!c || !scene) return fail(PTB_ERR_INVALID, "ptb_update_scene: null argument");
    int rc = validateScene(scene);
    if (rc != PTB_OK) return rc;
    int tris = 0;
    for (int i = 0; i < scene->numGeoms; ++i) if (scene->geoms[i].type == PTB_GEOM_MESH) tris += scene->meshRanges[2 * i + 1];
    if (tris != c->numTriangles) return fail(PTB_ERR_INVALID, "ptb_update_scene: %d mesh triangles, the context was created with %d", tris, c->numTriangles);
    if (scene->numGeoms != c->numGeoms || scene->numMaterials != c->numMats || scene->camera.resolution[0] != c->W ||
        scene->camera.resolution[1] != c->H || scene->traceDepth != c->traceDepth)
        return fail(PTB_ERR_INVALID, "ptb_update_scene: geometry/material counts, resolution and depth must match ptb_create");
    ON_DEVICE(c->device);
    { int frc = flushLookahead(c); if (frc != PTB_OK) return frc; }   // the window was traced through the old scene
    CU(cudaStreamSynchronize(c->stream));           // the pinned staging buffer is about to be rewritten
    rc = uploadScene(c, scene);
    if (rc != PTB_OK) return rc;
    c->fbValid = false;                              // new geometry / camera: the cached bounce-0 hits are stale
    rc = ensureStackSpill(c);                        // the rebuilt tree may be deeper
    if (rc != PTB_OK) return rc;
    return ptb_reset(c);
}

int ptb_destroy(PtbContext *c) {
    if (!c) return PTB_OK;                          // pathtraceFree before any pathtraceInit (src/main.cpp:126)
    DeviceScope deviceScope_(c->device);
    if (c->aheadStream) cudaStreamSynchronize(c->aheadStream);
    cudaStreamSynchronize(c->stream);
    delete c;
    return PTB_OK;
}

int ptb_render(PtbContext *c, int iterBegin, int count, int iterStride) {
    if (!c) return fail(PTB_ERR_INVALID, "ptb_render: not initialised");
    if (count < 0 || iterStride < 1 || iterBegin < 0) return fail(PTB_ERR_INVALID, "ptb_render: bad range");
    // makeSeededRandomEngine packs iter into 22 bits and depth into 9 (src/pathtrace.cu:43)
    if ((int64_t)iterBegin + (int64_t)(count - 1) * iterStride >= (1 << 22)) return fail(PTB_ERR_INVALID, "ptb_render: iteration beyond 2^22 aliases the RNG seed");
    ON_DEVICE(c->device);
    { int rc = flushLookahead(c); if (rc != PTB_OK) return rc; }
    return runBatches(c, iterBegin, count, iterStride);
}

int ptb_write_pbo(PtbContext *c, PtbUchar4 *pbo, int samples) {
    if (!c || !pbo) return fail(PTB_ERR_INVALID, "ptb_write_pbo: null argument");
    if (samples <= 0) return fail(PTB_ERR_INVALID, "ptb_write_pbo: samples %d (the divisor of sendImageToPBO must be positive)", samples);
    ON_DEVICE(c->device);
    { int frc = flushLookahead(c); if (frc != PTB_OK) return frc; }
    pbo_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(reinterpret_cast<uchar4 *>(pbo), c->imagePtr, c->numPixels, samples);
    c->launches++;
    CU(cudaGetLastError());
    return PTB_OK;
}

int ptb_download(PtbContext *c, float *hostImage) {
    if (!c || !hostImage) return fail(PTB_ERR_INVALID, "ptb_download: null argument");
    ON_DEVICE(c->device);
    { int frc = flushLookahead(c); if (frc != PTB_OK) return frc; }
    CU(cudaMemcpyAsync(hostImage, c->imagePtr, (size_t)c->numPixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

// page-locks the caller's image buffer the first time it is seen (the reference copies into the pageable std::vector of
// scene->state.image every iteration, src/pathtrace.cu:389-390: a staged copy at a fraction of the PCIe rate)
static void pinHostImage(PtbContext *c, float *hostImage, size_t bytes) {
    if (c->pinnedHost == hostImage && c->pinnedHostBytes == bytes) return;
    if (c->pinnedHost) { cudaStreamSynchronize(c->stream); if (c->copyStream) cudaStreamSynchronize(c->copyStream); cudaHostUnregister(c->pinnedHost); c->pinnedHost = nullptr; }
    if (c->pinHostFailed) return;
    if (cudaHostRegister(hostImage, bytes, cudaHostRegisterDefault) == cudaSuccess) { c->pinnedHost = hostImage; c->pinnedHostBytes = bytes; }
    else { cudaGetLastError(); c->pinHostFailed = true; }       // already pinned by the caller, or not registrable: plain copies still work
}

int ptb_pathtrace(PtbContext *c, PtbUchar4 *pbo, int frame, int iter, float *hostImage) {
    (void)frame;                                     // unused by the reference as well (src/main.cpp:136)
    if (!c) return fail(PTB_ERR_INVALID, "ptb_pathtrace: not initialised");
    if (iter < 1) return fail(PTB_ERR_INVALID, "ptb_pathtrace: iteration %d (the reference's iterations are 1-based, src/main.cpp:132)", iter);
    int rc;
    const bool lookahead = c->opt.lookahead && c->batch > 1 && c->base.partial && c->opt.pipeline == PTB_PIPELINE_FUSED && !c->opt.externalAccum;
    const float *image = c->imagePtr;
    if (lookahead) {
        // Look-ahead: the reference's loop asks for iterations 1, 2, 3, ... (src/main.cpp:132-136) and an iteration depends on nothing but its
        // number (src/pathtrace.cu:41-45), so the call for iteration i has the whole wavefront i .. i + batch - 1 traced at the batched rate
        // (on a second, low-priority stream) and the image AFTER every one of those iterations built in iteration order (deposit_prefix_kernel:
        // the additions of one iteration per call, bit for bit); the window after it is traced while this one is handed out.  A call then
        // costs its image copy: the strict semantic runs at the PCIe rate of the 12 B/pixel copy.  Any other call order, and every other
        // entry point that renders or touches image[], ends the look-ahead (flushLookahead).
        ON_DEVICE(c->device);
        rc = ensureLookahead(c);
        if (rc != PTB_OK) return rc;
        if ((int64_t)iter + 2 * (int64_t)c->batch >= (1 << 22)) return fail(PTB_ERR_INVALID, "ptb_pathtrace: iteration beyond 2^22 aliases the RNG seed");
        if (c->laValid && c->laServed == c->batch && iter == c->laBase + c->batch && c->laNext) {
            c->laBase += c->batch; c->laSet ^= 1; c->laServed = 0; c->laNext = false;      // the window traced ahead becomes the current one
        } else if (!(c->laValid && iter == c->laBase + c->laServed && c->laServed < c->batch)) {
            rc = flushLookahead(c);                                // (also makes image[] current: the base of the new window)
            if (rc != PTB_OK) return rc;
            rc = launchWindow(c, iter, 0, c->imagePtr);
            if (rc != PTB_OK) return rc;
            c->laValid = true; c->laBase = iter; c->laServed = 0; c->laSet = 0; c->laNext = false;
        }
        if (c->laServed == 0) {
            CU(cudaStreamWaitEvent(c->stream, c->evWave[c->laSet], 0));
            if (!c->laNext) {
                // the next window starts from the last image of this one; the copies out of the set it overwrites were queued before this point
                rc = launchWindow(c, c->laBase + c->batch, c->laSet ^ 1, c->prefix[c->laSet].p + (size_t)(c->batch - 1) * c->partialStride);
                if (rc != PTB_OK) return rc;
                c->laNext = true;
            }
        }
        image = c->laImage = c->prefix[c->laSet].p + (size_t)c->laServed * c->partialStride;
        c->laServed++;
        c->hostIterations++;
    } else {
        rc = ptb_render(c, iter, 1, 1);
        if (rc != PTB_OK) return rc;
    }
    if (pbo) {
        ON_DEVICE(c->device);
        pbo_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(reinterpret_cast<uchar4 *>(pbo), image, c->numPixels, iter);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (!hostImage) return PTB_OK;
    ON_DEVICE(c->device);
    const size_t bytes = (size_t)c->numPixels * 3 * sizeof(float);
    pinHostImage(c, hostImage, bytes);
    if (!c->opt.lazyImage || lookahead) {
        // the reference's semantic: the accumulated SUM is in the host image when the call returns.  (Look-ahead + lazy image: the same copy
        // without the wait -- a prefix image is never written again before its window's set is reused: no snapshot needed)
        CU(cudaMemcpyAsync(hostImage, image, bytes, cudaMemcpyDeviceToHost, c->stream));
        if (!c->opt.lazyImage) CU(cudaStreamSynchronize(c->stream));
        return PTB_OK;
    }
    // lazy image: snapshot on the render stream (7.68 MB device to device), copy out on the copy stream while the next
    // iteration renders; two snapshots alternate, each reused only after its previous copy has left
    if (!c->copyStream) {
        CU(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CU(cudaEventCreateWithFlags(&c->evImage[k], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->evCopy[k], cudaEventDisableTiming));
            CU(c->snap[k].alloc((size_t)c->numPixels * 3));
        }
    }
    const int k = c->snapIdx;
    c->snapIdx ^= 1;
    CU(cudaStreamWaitEvent(c->stream, c->evCopy[k], 0));           // (a never-recorded event counts as complete)
    CU(cudaMemcpyAsync(c->snap[k].p, c->imagePtr, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaEventRecord(c->evImage[k], c->stream));
    CU(cudaStreamWaitEvent(c->copyStream, c->evImage[k], 0));
    CU(cudaMemcpyAsync(hostImage, c->snap[k].p, bytes, cudaMemcpyDeviceToHost, c->copyStream));
    CU(cudaEventRecord(c->evCopy[k], c->copyStream));
    return PTB_OK;
}

int ptb_reset(PtbContext *c) {
    if (!c) return fail(PTB_ERR_INVALID, "ptb_reset: not initialised");
    ON_DEVICE(c->device);
    { int frc = flushLookahead(c); if (frc != PTB_OK) return frc; }
    CU(cudaMemsetAsync(c->imagePtr, 0, (size_t)c->numPixels * 3 * sizeof(float), c->stream));
    CU(cudaMemsetAsync(c->ctrl.p, 0, sizeof(Control), c->stream));
    c->hostIterations = 0;
    c->frameValid = false;
    if (c->opt.timeStages) {
        CU(cudaStreamSynchronize(c->stream));
        foldStageTimers(c);
        for (int k = 0; k < PTB_NUM_STAGES; ++k) { c->stageUs[k] = 0.0; c->stageLaunches[k] = 0; }
    }
    return PTB_OK;
}

int ptb_sync(PtbContext *c) {
    if (!c) return fail(PTB_ERR_INVALID, "ptb_sync: not initialised");
    ON_DEVICE(c->device);
    CU(cudaStreamSynchronize(c->stream));
    if (c->copyStream) CU(cudaStreamSynchronize(c->copyStream));   // lazy images of ptb_pathtrace have landed
    return PTB_OK;
}

int ptb_stats(PtbContext *c, PtbStats *out) {
    if (!c || !out) return fail(PTB_ERR_INVALID, "ptb_stats: null argument");
    ON_DEVICE(c->device);
    if (c->aheadStream) CU(cudaStreamSynchronize(c->aheadStream));       // a look-ahead window may still be tracing: its counts belong to the statistics
    Control h;
    CU(cudaMemcpyAsync(&h, c->ctrl.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    foldStageTimers(c);
    memset(out, 0, sizeof *out);
    for (int k = 0; k < PTB_NUM_STAGES; ++k) { out->stageUs[k] = c->stageUs[k]; out->stageLaunches[k] = c->stageLaunches[k]; }
    out->iterations = (int64_t)h.iterations;
    out->segments = (int64_t)h.segments;
    out->kernelLaunches = c->launches;
    for (int d = 0; d < 64; ++d) out->alivePerDepth[d] = (int64_t)h.aliveTotals[d];
    out->traceDepth = c->traceDepth;
    out->numGeoms = c->numGeoms;
    out->bvhNodes = c->useBvh ? c->numNodes : 0;
    out->batchIters = c->batch;
    out->pathStateBytes = (int64_t)c->numPixels * c->batch * 40 * 2;
    out->sceneUploadBytes = c->h2dBytes;
    return PTB_OK;
}

int ptb_accum_device_ptr(PtbContext *c, float **devPtr) {
    if (!c || !devPtr) return fail(PTB_ERR_INVALID, "ptb_accum_device_ptr: null argument");
    { ON_DEVICE(c->device); int frc = flushLookahead(c); if (frc != PTB_OK) return frc; }
    *devPtr = c->imagePtr;
    return PTB_OK;
}

int ptb_resolve_rgb8(PtbContext *c, int samples, uint8_t *hostRgb8) {
    if (!c || !hostRgb8 || samples <= 0) return fail(PTB_ERR_INVALID, "ptb_resolve_rgb8: bad argument");
    ON_DEVICE(c->device);
    { int frc = flushLookahead(c); if (frc != PTB_OK) return frc; }
    if (!c->rgb8.p) CU(c->rgb8.alloc((size_t)c->numPixels * 3));
    resolve_rgb8_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(c->rgb8.p, c->imagePtr, c->W, c->H, (float)samples);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hostRgb8, c->rgb8.p, (size_t)c->numPixels * 3, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

// ---- several GPUs (include/ptb200.h) ---------------------------------------------------------------
#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) return fail(PTB_ERR_CUDA, "NCCL error (%s:%d): %s: %s", "ptb200.cu", __LINE__, #call, nccl->getErrorString(r_)); \
    } while (0)

int ptb_comm_unique_id(uint8_t id[PTB_COMM_ID_BYTES]) {
    if (!id) return fail(PTB_ERR_INVALID, "ptb_comm_unique_id: null argument");
    std::string err;
    NcclApi *nccl = ncclApi(err);
    if (!nccl) return fail(PTB_ERR_UNSUPPORTED, "%s", err.c_str());
    static_assert(sizeof(ncclUniqueId) == PTB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NC(nccl->getUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return PTB_OK;
}

static int ensureFrame(PtbContext *c) {
    if (!c->frame.p) CU(c->frame.alloc((size_t)c->numPixels * 3));
    return PTB_OK;
}

int ptb_comm_init_rank(PtbContext *c, const uint8_t id[PTB_COMM_ID_BYTES], int rank, int nranks) {
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(PTB_ERR_INVALID, "ptb_comm_init_rank: bad argument");
    if (c->comm) return fail(PTB_ERR_INVALID, "ptb_comm_init_rank: the context already has a communicator");
    std::string err;
    NcclApi *nccl = ncclApi(err);
    if (!nccl) return fail(PTB_ERR_UNSUPPORTED, "%s", err.c_str());
    ON_DEVICE(c->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    NC(nccl->commInitRank(&c->comm, nranks, u, rank));
    c->ownComm = true; c->commRank = rank; c->commRanks = nranks;
    return PTB_OK;
}

int ptb_comm_reduce(PtbContext *c, int root) {
    if (!c || !c->comm) return fail(PTB_ERR_INVALID, "ptb_comm_reduce: no communicator (ptb_comm_init_rank / ptb_multi_create)");
    if (root < 0 || root >= c->commRanks) return fail(PTB_ERR_INVALID, "ptb_comm_reduce: root %d of %d", root, c->commRanks);
    std::string err;
    NcclApi *nccl = ncclApi(err);
    if (!nccl) return fail(PTB_ERR_UNSUPPORTED, "%s", err.c_str());
    ON_DEVICE(c->device);
    { int frc = flushLookahead(c); if (frc != PTB_OK) return frc; }
    if (c->commRank == root) { int rc = ensureFrame(c); if (rc != PTB_OK) return rc; }
    // one collective per frame: count = 3 * W * H floats; recvbuff only matters on the root
    NC(nccl->reduce(c->imagePtr, c->commRank == root ? c->frame.p : nullptr, (size_t)c->numPixels * 3, ncclFloat32, ncclSum, root, c->comm, c->stream));
    c->frameValid = c->commRank == root;
    return PTB_OK;
}

static float *framePtr(PtbContext *c) { return (c->frameValid && c->frame.p) ? c->frame.p : c->imagePtr; }

int ptb_frame_device_ptr(PtbContext *c, float **devPtr) {
    if (!c || !devPtr) return fail(PTB_ERR_INVALID, "ptb_frame_device_ptr: null argument");
    *devPtr = framePtr(c);
    return PTB_OK;
}
int ptb_frame_download(PtbContext *c, float *hostImage) {
    if (!c || !hostImage) return fail(PTB_ERR_INVALID, "ptb_frame_download: null argument");
    ON_DEVICE(c->device);
    CU(cudaMemcpyAsync(hostImage, framePtr(c), (size_t)c->numPixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}
static int resolveFrom(PtbContext *c, const float *src, int samples, uint8_t *hostRgb8) {
    if (!c->rgb8.p) CU(c->rgb8.alloc((size_t)c->numPixels * 3));
    resolve_rgb8_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(c->rgb8.p, src, c->W, c->H, (float)samples);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hostRgb8, c->rgb8.p, (size_t)c->numPixels * 3, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}
int ptb_frame_resolve_rgb8(PtbContext *c, int samples, uint8_t *hostRgb8) {
    if (!c || !hostRgb8 || samples <= 0) return fail(PTB_ERR_INVALID, "ptb_frame_resolve_rgb8: bad argument");
    ON_DEVICE(c->device);
    return resolveFrom(c, framePtr(c), samples, hostRgb8);
}

} // extern "C"

struct PtbMulti {
    std::vector<PtbContext *> ctx;
    std::vector<ncclComm_t> comms;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;       // on the root's device: reduce timing
};

extern "C" {

int ptb_multi_destroy(PtbMulti *m) {
    if (!m) return PTB_OK;
    std::string err;
    NcclApi *nccl = ncclApi(err);
    for (size_t g = 0; g < m->ctx.size(); ++g) {
        if (m->ctx[g]) { cudaSetDevice(m->ctx[g]->device); cudaStreamSynchronize(m->ctx[g]->stream); }
    }
    if (nccl) for (ncclComm_t cm : m->comms) if (cm) nccl->commDestroy(cm);
    if (!m->ctx.empty() && m->ctx[0]) {
        DeviceScope ds(m->ctx[0]->device);
        if (m->ev0) cudaEventDestroy(m->ev0);
        if (m->ev1) cudaEventDestroy(m->ev1);
    }
    for (PtbContext *c : m->ctx) if (c) { c->comm = nullptr; ptb_destroy(c); }
    delete m;
    return PTB_OK;
}

int ptb_multi_create(PtbMulti **out, const PtbSceneDesc *scene, const PtbOptions *optIn, const int32_t *devices, int32_t n) {
    if (!out || !scene || n < 1) return fail(PTB_ERR_INVALID, "ptb_multi_create: bad argument");
    *out = nullptr;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    std::vector<int> devs(n);
    for (int g = 0; g < n; ++g) {
        devs[g] = devices ? devices[g] : g;
        if (devs[g] < 0 || devs[g] >= ndev) return fail(PTB_ERR_INVALID, "ptb_multi_create: device %d of %d", devs[g], ndev);
        for (int k = 0; k < g; ++k) if (devs[k] == devs[g]) return fail(PTB_ERR_INVALID, "ptb_multi_create: device %d listed twice", devs[g]);
    }
    std::string err;
    NcclApi *nccl = n > 1 ? ncclApi(err) : nullptr;
    if (n > 1 && !nccl) return fail(PTB_ERR_UNSUPPORTED, "%s", err.c_str());
    PtbMulti *m = new PtbMulti();
    m->ctx.assign(n, nullptr);
    PtbOptions opt;
    if (optIn) opt = *optIn; else ptb_default_options(&opt);
    opt.stream = nullptr; opt.externalAccum = nullptr;
    for (int g = 0; g < n; ++g) {
        opt.device = devs[g];
        int rc = ptb_create(&m->ctx[g], scene, &opt);
        if (rc != PTB_OK) { const std::string keep = g_lastError; ptb_multi_destroy(m); g_lastError = keep; return rc; }
    }
    if (n > 1) {
        m->comms.assign(n, nullptr);
        ncclResult_t r = nccl->commInitAll(m->comms.data(), n, devs.data());
        if (r != ncclSuccess) { ptb_multi_destroy(m); return fail(PTB_ERR_CUDA, "ncclCommInitAll: %s", nccl->getErrorString(r)); }
        for (int g = 0; g < n; ++g) { m->ctx[g]->comm = m->comms[g]; m->ctx[g]->ownComm = false; m->ctx[g]->commRank = g; m->ctx[g]->commRanks = n; }
        // one collective now (over the still all-zero accumulation buffers): NCCL sets its channels up on first use, which would
        // otherwise be charged to the first frame (35 ms on 2 B200s against ~0.1 ms for the reduce itself)
        int rc = PTB_OK;
        { DeviceScope ds(m->ctx[0]->device); rc = ensureFrame(m->ctx[0]); }
        if (rc == PTB_OK) {
            nccl->groupStart();
            for (int g = 0; g < n; ++g) {
                PtbContext *c = m->ctx[g];
                nccl->reduce(c->imagePtr, g == 0 ? c->frame.p : nullptr, (size_t)c->numPixels * 3, ncclFloat32, ncclSum, 0, c->comm, c->stream);
            }
            ncclResult_t r2 = nccl->groupEnd();
            if (r2 != ncclSuccess) rc = fail(PTB_ERR_CUDA, "ncclReduce (set-up): %s", nccl->getErrorString(r2));
            for (int g = 0; g < n && rc == PTB_OK; ++g) rc = ptb_sync(m->ctx[g]);
        }
        if (rc != PTB_OK) { const std::string keep = g_lastError; ptb_multi_destroy(m); g_lastError = keep; return rc; }
    }
    {
        DeviceScope ds(m->ctx[0]->device);
        if (cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess) { ptb_multi_destroy(m); return fail(PTB_ERR_CUDA, "cudaEventCreate"); }
    }
    *out = m;
    return PTB_OK;
}

int ptb_multi_reset(PtbMulti *m) {
    if (!m) return fail(PTB_ERR_INVALID, "ptb_multi_reset: null argument");
    for (PtbContext *c : m->ctx) { int rc = ptb_reset(c); if (rc != PTB_OK) return rc; c->frameValid = false; }
    return PTB_OK;
}

int ptb_multi_render(PtbMulti *m, int iterBegin, int count) {
    if (!m || count < 0) return fail(PTB_ERR_INVALID, "ptb_multi_render: bad argument");
    const int G = (int)m->ctx.size();
    // device g renders iterBegin + g, + G, ...: every prefix of the sample sequence stays balanced.  Asynchronous launches:
    // the devices run concurrently under one host thread (the whole batch loop is CUDA-graph replays).
    for (int g = 0; g < G; ++g) {
        const int mine = count > g ? (count - g + G - 1) / G : 0;
        if (!mine) continue;
        int rc = ptb_render(m->ctx[g], iterBegin + g, mine, G);
        if (rc != PTB_OK) return rc;
    }
    return PTB_OK;
}

int ptb_multi_frame(PtbMulti *m, int samples, float *hostImage, uint8_t *hostRgb8, float *reduceMs) {
    if (!m || (hostRgb8 && samples <= 0)) return fail(PTB_ERR_INVALID, "ptb_multi_frame: bad argument");
    const int G = (int)m->ctx.size();
    PtbContext *root = m->ctx[0];
    if (reduceMs) *reduceMs = 0.f;
    if (G > 1) {
        std::string err;
        NcclApi *nccl = ncclApi(err);
        if (!nccl) return fail(PTB_ERR_UNSUPPORTED, "%s", err.c_str());
        { ON_DEVICE(root->device); int rc = ensureFrame(root); if (rc != PTB_OK) return rc; CU(cudaEventRecord(m->ev0, root->stream)); }
        // ONE grouped collective per frame: ncclReduce(sum, float32, 3*W*H) from every device's private accumulation to the root's frame
        NC(nccl->groupStart());
        for (int g = 0; g < G; ++g) {
            PtbContext *c = m->ctx[g];
            ncclResult_t r = nccl->reduce(c->imagePtr, g == 0 ? root->frame.p : nullptr, (size_t)c->numPixels * 3, ncclFloat32, ncclSum, 0, c->comm, c->stream);
            if (r != ncclSuccess) { nccl->groupEnd(); return fail(PTB_ERR_CUDA, "ncclReduce (device %d): %s", c->device, nccl->getErrorString(r)); }
        }
        NC(nccl->groupEnd());
        root->frameValid = true;
        { ON_DEVICE(root->device); CU(cudaEventRecord(m->ev1, root->stream)); }
    }
    // the senders finish their part of the collective on their own streams
    for (int g = 1; g < G; ++g) { int rc = ptb_sync(m->ctx[g]); if (rc != PTB_OK) return rc; }
    if (hostRgb8) { ON_DEVICE(root->device); int rc = resolveFrom(root, framePtr(root), samples, hostRgb8); if (rc != PTB_OK) return rc; }
    if (hostImage) { int rc = ptb_frame_download(root, hostImage); if (rc != PTB_OK) return rc; }
    { int rc = ptb_sync(root); if (rc != PTB_OK) return rc; }
    if (G > 1 && reduceMs) { ON_DEVICE(root->device); CU(cudaEventElapsedTime(reduceMs, m->ev0, m->ev1)); }
    return PTB_OK;
}

int ptb_multi_stats(PtbMulti *m, PtbStats *sum) {
    if (!m || !sum) return fail(PTB_ERR_INVALID, "ptb_multi_stats: null argument");
    memset(sum, 0, sizeof *sum);
    for (size_t g = 0; g < m->ctx.size(); ++g) {
        PtbStats s;
        int rc = ptb_stats(m->ctx[g], &s);
        if (rc != PTB_OK) return rc;
        sum->iterations += s.iterations; sum->segments += s.segments; sum->kernelLaunches += s.kernelLaunches;
        for (int d = 0; d < 64; ++d) sum->alivePerDepth[d] += s.alivePerDepth[d];
        for (int k = 0; k < PTB_NUM_STAGES; ++k) { sum->stageUs[k] += s.stageUs[k]; sum->stageLaunches[k] += s.stageLaunches[k]; }
        sum->traceDepth = s.traceDepth; sum->numGeoms = s.numGeoms; sum->bvhNodes = s.bvhNodes; sum->batchIters = s.batchIters;
        sum->pathStateBytes += s.pathStateBytes; sum->sceneUploadBytes += s.sceneUploadBytes;
    }
    return PTB_OK;
}

int ptb_multi_context(PtbMulti *m, int index, PtbContext **ctx) {
    if (!m || !ctx || index < 0 || index >= (int)m->ctx.size()) return fail(PTB_ERR_INVALID, "ptb_multi_context: bad argument");
    *ctx = m->ctx[index];
    return PTB_OK;
}

// ---- stage-level entry points ---------------------------------------------------------------------
int ptb_stage_raygen(PtbContext *c, float *origins, float *dirs) { return ptb_stage_raygen_iter(c, 1, origins, dirs); }

int ptb_stage_raygen_iter(PtbContext *c, int iter, float *origins, float *dirs) {
    if (!c || !origins || !dirs) return fail(PTB_ERR_INVALID, "ptb_stage_raygen: null argument");
    ON_DEVICE(c->device);
    const size_t n = (size_t)c->numPixels * 3;
    DevBuf<float> o, d;
    CU(o.alloc(n)); CU(d.alloc(n));
    stage_raygen_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(c->base, iter, o.p, d.p);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(origins, o.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(dirs, d.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    o.release(); d.release();
    return PTB_OK;
}

int ptb_stage_intersect(PtbContext *c, int accel, const float *origins, const float *dirs, int n,
                        int32_t *hit, float *t, float *normal, int32_t *outside) {
    if (!c || !origins || !dirs || !hit || !t || !normal || !outside || n < 0) return fail(PTB_ERR_INVALID, "ptb_stage_intersect: bad argument");
    ON_DEVICE(c->device);
    bool bvh = accel == PTB_ACCEL_BVH || (accel == PTB_ACCEL_AUTO && c->useBvh);
    if (!bvh && c->numPrims != c->numGeoms) return fail(PTB_ERR_UNSUPPORTED, "ptb_stage_intersect: scenes with meshes have no brute-force table");
    if (!bvh && c->smemBrute > 200 * 1024) return fail(PTB_ERR_UNSUPPORTED, "brute force needs %zu B of shared memory", c->smemBrute);
    DevBuf<float> o, d, dt, dn; DevBuf<int> dh, dout;
    DevBuf<float4> rawTn; DevBuf<int> rawMeta;
    CU(o.alloc((size_t)n * 3)); CU(d.alloc((size_t)n * 3)); CU(dt.alloc(n)); CU(dn.alloc((size_t)n * 3)); CU(dh.alloc(n)); CU(dout.alloc(n));
    if (n) {
        CU(cudaMemcpyAsync(o.p, origins, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d.p, dirs, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
        Params p = c->base;
        if (bvh) {
            // raw hit records through trace_kernel (the traversal every BVH render runs), then normals / scene indices
            CU(rawTn.alloc(n)); CU(rawMeta.alloc(n));
            p.geoms = c->geoms.p; p.hits.tn = rawTn.p; p.hits.meta = rawMeta.p;
            // (through LAUNCH: with PtbOptions.timeStages the traversal of the caller's rays shows up in PtbStats.stageUs[PTB_STAGE_INTERSECT],
            // tools/gpu_ray_order.py)
            LAUNCH(PTB_STAGE_INTERSECT, (TRACE_KERNEL<SRC_RAYS, false, true>), c->gridTrace, kTraceBlock, c->smemBvh, p, 0, TraceIO{o.p, d.p, n});
            stage_hits_out_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(p, n, dh.p, dt.p, dn.p, dout.p);
        }
        else { p.geoms = c->geomsScene.p; stage_intersect_kernel<<<c->gridStaged, kBlock, c->smemBrute, c->stream>>>(p, o.p, d.p, n, dh.p, dt.p, dn.p, dout.p); }
        c->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hit, dh.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(t, dt.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(normal, dn.p, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(outside, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    o.release(); d.release(); dt.release(); dn.release(); dh.release(); dout.release();
    return PTB_OK;
}

int ptb_stage_shade(PtbContext *c, int n, const int32_t *hit, const float *t, const float *normal,
                    const int32_t *outside, const int32_t *iter, const int32_t *pixelIndex,
                    const int32_t *depth, float *origins, float *dirs, float *colors, int32_t *remaining) {
    if (!c || n < 0) return fail(PTB_ERR_INVALID, "ptb_stage_shade: bad argument");
    ON_DEVICE(c->device);
    DevBuf<int> dh, dout, dit, dpx, ddp, drem; DevBuf<float> dt, dn, o, d, col;
    CU(dh.alloc(n)); CU(dout.alloc(n)); CU(dit.alloc(n)); CU(dpx.alloc(n)); CU(ddp.alloc(n)); CU(drem.alloc(n));
    CU(dt.alloc(n)); CU(dn.alloc((size_t)n * 3)); CU(o.alloc((size_t)n * 3)); CU(d.alloc((size_t)n * 3)); CU(col.alloc((size_t)n * 3));
    if (n) {
        auto up = [&](void *dst, const void *src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream); };
        CU(up(dh.p, hit, (size_t)n * 4)); CU(up(dout.p, outside, (size_t)n * 4)); CU(up(dit.p, iter, (size_t)n * 4));
        CU(up(dpx.p, pixelIndex, (size_t)n * 4)); CU(up(ddp.p, depth, (size_t)n * 4)); CU(up(drem.p, remaining, (size_t)n * 4));
        CU(up(dt.p, t, (size_t)n * 4)); CU(up(dn.p, normal, (size_t)n * 12)); CU(up(o.p, origins, (size_t)n * 12));
        CU(up(d.p, dirs, (size_t)n * 12)); CU(up(col.p, colors, (size_t)n * 12));
        stage_shade_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(c->base, n, dh.p, c->sceneMatDev.p, dt.p, dn.p, dout.p, dit.p, dpx.p, ddp.p,
                                                                   o.p, d.p, col.p, drem.p);
        c->launches++;
        CU(cudaGetLastError());
        auto down = [&](void *dst, const void *src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream); };
        CU(down(origins, o.p, (size_t)n * 12)); CU(down(dirs, d.p, (size_t)n * 12)); CU(down(colors, col.p, (size_t)n * 12));
        CU(down(remaining, drem.p, (size_t)n * 4));
    }
    CU(cudaStreamSynchronize(c->stream));
    dh.release(); dout.release(); dit.release(); dpx.release(); ddp.release(); drem.release();
    dt.release(); dn.release(); o.release(); d.release(); col.release();
    return PTB_OK;
}

int ptb_stage_compact(PtbContext *c, int n, const int32_t *ids, const int32_t *flags, int32_t *outIds, int32_t *outCount) {
    if (!c || n < 0 || !outCount) return fail(PTB_ERR_INVALID, "ptb_stage_compact: bad argument");
    ON_DEVICE(c->device);
    DevBuf<int> di, df, dout; DevBuf<uint32_t> cnt;
    CU(di.alloc(n)); CU(df.alloc(n)); CU(dout.alloc(n)); CU(cnt.alloc(1));
    CU(cudaMemsetAsync(cnt.p, 0, 4, c->stream));
    if (n) {
        CU(cudaMemcpyAsync(di.p, ids, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(df.p, flags, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        stage_compact_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(n, di.p, df.p, dout.p, cnt.p);
        c->launches++;
        CU(cudaGetLastError());
    }
    uint32_t k = 0;
    CU(cudaMemcpyAsync(&k, cnt.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (k && outIds) CU(cudaMemcpy(outIds, dout.p, (size_t)k * 4, cudaMemcpyDeviceToHost));
    *outCount = (int32_t)k;
    di.release(); df.release(); dout.release(); cnt.release();
    return PTB_OK;
}

int ptb_stage_sort_keys(PtbContext *c, int n, const int32_t *keys, int numKeys, int32_t *outOrder) {
    if (!c || n < 0 || numKeys <= 0 || numKeys > 8192) return fail(PTB_ERR_INVALID, "ptb_stage_sort_keys: bad argument");
    for (int i = 0; i < n; ++i) if (keys[i] < 0 || keys[i] >= numKeys) return fail(PTB_ERR_INVALID, "ptb_stage_sort_keys: key %d out of range", keys[i]);
    ON_DEVICE(c->device);
    DevBuf<int> dk, dord; DevBuf<uint32_t> hist;
    CU(dk.alloc(n)); CU(dord.alloc(n)); CU(hist.alloc(numKeys));
    CU(cudaMemsetAsync(hist.p, 0, sizeof(uint32_t) * numKeys, c->stream));
    if (n) {
        CU(cudaMemcpyAsync(dk.p, keys, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        sort_hist_kernel<<<c->gridStaged, kBlock, sizeof(uint32_t) * numKeys, c->stream>>>(dk.p, (uint32_t)n, numKeys, hist.p);
        sort_scan_kernel<<<1, 32, 0, c->stream>>>(hist.p, numKeys);
        sort_scatter_dev_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(dk.p, nullptr, 0, hist.p, dord.p, (uint32_t)n, numKeys);
        c->launches += 3;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(outOrder, dord.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    dk.release(); dord.release(); hist.release();
    return PTB_OK;
}

int ptb_stage_arith(PtbContext *c, int n, const float *a, const float *b, float *quot, float *root, float *normY) {
    if (!c || n < 0 || !a || !b || !quot || !root || !normY) return fail(PTB_ERR_INVALID, "ptb_stage_arith: bad argument");
    ON_DEVICE(c->device);
    DevBuf<float> da, db, dq, dr, dn;
    CU(da.alloc(n)); CU(db.alloc(n)); CU(dq.alloc(n)); CU(dr.alloc(n)); CU(dn.alloc(n));
    if (n) {
        CU(cudaMemcpyAsync(da.p, a, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(db.p, b, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        stage_arith_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(n, da.p, db.p, dq.p, dr.p, dn.p);
        c->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(quot, dq.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(root, dr.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(normY, dn.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    da.release(); db.release(); dq.release(); dr.release(); dn.release();
    return PTB_OK;
}

int ptb_stage_sincos(PtbContext *c, int n, const float *angles, float *sinOut, float *cosOut) {
    if (!c || n < 0 || !angles || !sinOut || !cosOut) return fail(PTB_ERR_INVALID, "ptb_stage_sincos: bad argument");
    ON_DEVICE(c->device);
    DevBuf<float> da, ds, dc;
    CU(da.alloc(n)); CU(ds.alloc(n)); CU(dc.alloc(n));
    if (n) {
        CU(cudaMemcpyAsync(da.p, angles, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        stage_sincos_kernel<<<c->gridStaged, kBlock, 0, c->stream>>>(n, da.p, ds.p, dc.p);
        c->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(sinOut, ds.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(cosOut, dc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    da.release(); ds.release(); dc.release();
    return PTB_OK;
}

int ptb_stage_rng(PtbContext *c, int iter, int index, int depth, int n, float *out) {
    if (!c || n < 0 || !out) return fail(PTB_ERR_INVALID, "ptb_stage_rng: bad argument");
    ON_DEVICE(c->device);
    DevBuf<float> d;
    CU(d.alloc(n));
    if (n) {
        stage_rng_kernel<<<1, 32, 0, c->stream>>>(iter, index, depth, n, d.p);
        c->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, d.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    d.release();
    return PTB_OK;
}

} // extern "C"

#ifdef PTB_WALK_STATS
// tuning builds only: read and clear the candidate-walk histogram (ptb_kernels.cuh, PTB_WALK_COUNT)
extern "C" int ptb_debug_walk_stats(unsigned long long *out) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, ptb::g_walkStats, sizeof(unsigned long long) * 128) != cudaSuccess) return 1;
    static unsigned long long zeros[128];
    return cudaMemcpyToSymbol(ptb::g_walkStats, zeros, sizeof zeros) != cudaSuccess;
}
#endif

