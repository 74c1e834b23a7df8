/*
 * ptb200.h -- C ABI of the B200-native wavefront path tracer (libptb200.so).
 *
 * Drop-in boundary for the renderer core of CIS565 Project3-CUDA-Path-Tracer.  Every entry point
 * cites the reference interface it replaces (paths relative to the reference repository).  Plain
 * pointers and sizes only: no C++ types, no torch types, no exceptions across the boundary.
 * All functions return PTB_OK (0) or a negative PtbError; ptb_last_error() gives the message.
 *
 * The reference's three C++ symbols
 *     void pathtraceInit(Scene *scene);                       src/pathtrace.h:6
 *     void pathtraceFree();                                   src/pathtrace.h:7
 *     void pathtrace(uchar4 *pbo, int frame, int iteration);  src/pathtrace.h:8
 * are provided on top of this ABI by csrc/host/pathtrace.cpp.
 *
 * There is NO CPU fallback: every call that renders fails with PTB_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_ABI_VERSION 7

/* ---- POD mirrors of src/sceneStructs.h (byte-identical; matrices column-major like glm) ---- */
typedef struct PtbGeom {            /* struct Geom, src/sceneStructs.h:20-29, 236 bytes */
    int32_t type;                   /* enum GeomType: 0 = SPHERE, 1 = CUBE (src/sceneStructs.h:10-13); 2 = MESH (extension, see PtbTriangle) */
    int32_t materialid;
    float   translation[3];
    float   rotation[3];
    float   scale[3];
    float   transform[16];
    float   inverseTransform[16];
    float   invTranspose[16];
} PtbGeom;

typedef struct PtbMaterial {        /* struct Material, src/sceneStructs.h:31-41, 44 bytes */
    float color[3];
    float specularExponent;
    float specularColor[3];
    float hasReflective;
    float hasRefractive;
    float indexOfRefraction;
    float emittance;
} PtbMaterial;

typedef struct PtbCamera {          /* struct Camera, src/sceneStructs.h:43-52, 84 bytes */
    int32_t resolution[2];
    float   position[3];
    float   lookAt[3];
    float   view[3];
    float   up[3];
    float   right[3];
    float   fov[2];
    float   pixelLength[2];
} PtbCamera;

/* Triangle meshes.  The scene format reserves the object type `mesh` (INSTRUCTION.md:291-293) and the
 * assignment names glm::intersectRayTriangle for it (INSTRUCTION.md:123-127); the reference's loader and kernels
 * implement neither.  A geom of type PTB_GEOM_MESH owns triangles[first .. first + count) of the scene's
 * triangle array: object-space vertices under the geom's transform, counter-clockwise = front face. */
#define PTB_GEOM_MESH 2
typedef struct PtbTriangle { float v0[3], v1[3], v2[3]; } PtbTriangle;

typedef struct PtbUchar4 { uint8_t x, y, z, w; } PtbUchar4;   /* CUDA uchar4 (the GL PBO texel) */

/* What pathtraceInit reads from `Scene*` (src/pathtrace.cu:79-93): geoms, materials,
 * state.camera (after main.cpp's camera fix-up, src/main.cpp:102-120) and state.traceDepth. */
typedef struct PtbSceneDesc {
    const PtbGeom     *geoms;
    int32_t            numGeoms;
    const PtbMaterial *materials;
    int32_t            numMaterials;
    PtbCamera          camera;
    int32_t            traceDepth;
    /* meshes (all three zero / NULL when no geom has type PTB_GEOM_MESH) */
    int32_t            numTriangles;
    const PtbTriangle *triangles;
    const int32_t     *meshRanges;  /* 2 ints per geom: first triangle, triangle count (ignored for spheres / cubes) */
} PtbSceneDesc;

enum PtbBsdfMode {
    PTB_BSDF_GOLDEN_DIFFUSE = 0,    /* every non-emissive surface is Lambertian: reproduces img/REFERENCE_cornell.5000samp.png */
    PTB_BSDF_FULL           = 1     /* diffuse + REFL (mirror) + REFR (Schlick dielectric) */
};
enum PtbPipeline {
    PTB_PIPELINE_FUSED  = 0,        /* one kernel per bounce: raygen|load -> intersect -> shade -> compact -> gather */
    PTB_PIPELINE_STAGED = 1         /* the reference's stage structure: raygen, intersect, [sort], shade, [compact], gather */
};
enum PtbAccel { PTB_ACCEL_AUTO = 0, PTB_ACCEL_BRUTE = 1, PTB_ACCEL_BVH = 2 };
enum PtbDepositMode {
    PTB_DEPOSIT_ORDERED = 0,        /* default: every slot of a wavefront deposits into its own partial image and the partial images
                                       are added to image[] in iteration order -- bit for bit the sum the reference's loop builds one
                                       iteration at a time (src/pathtrace.cu:269-278), reproducible run to run, for any batchIters */
    PTB_DEPOSIT_ATOMIC  = 1         /* float atomicAdd straight into image[]: the order of the additions inside one wavefront, and so
                                       the last bits of the sum, vary run to run (|error| <= batchIters * 2^-24 relative) */
};

typedef struct PtbOptions {
    int32_t device;                 /* CUDA ordinal; -1 = current device */
    int32_t bsdfMode;               /* PtbBsdfMode */
    int32_t pipeline;               /* PtbPipeline */
    int32_t accel;                  /* PtbAccel */
    int32_t batchIters;             /* iterations traced per wavefront (fused pipeline); 0 = default: the largest power of two <= 64 that keeps the
                                       wavefront near 140 M paths (~16 GB of state) when the device has >= 64 GB free, near 40 M otherwise */
    int32_t compaction;             /* staged pipeline: 1 = ballot/popc stream compaction, 0 = none (skeleton layout) */
    int32_t sortMaterials;          /* staged pipeline: 1 = sort paths by material id before shading */
    int32_t firstBounceCache;       /* fused pipeline: 1 = cache the bounce-0 hit record of every pixel across iterations
                                       (INSTRUCTION.md:87-89); refilled after ptb_create / ptb_update_scene */
    int32_t useGraphs;              /* 1 = replay the per-batch launch sequence as a CUDA graph */
    int32_t antialias;              /* 1 = stochastic antialiasing: the sample position is jittered uniformly inside the
                                       pixel each iteration (the TODO of src/pathtrace.cu:118-120,134; INSTRUCTION.md:109-113) */
    void   *stream;                 /* cudaStream_t to launch on; NULL = library-owned stream */
    float  *externalAccum;          /* device pointer to W*H*3 floats to accumulate into; NULL = library-owned */
    float   lensRadius;             /* > 0 = thin-lens depth of field (INSTRUCTION.md:107-108, PBRT 6.2.3) */
    float   focalDistance;          /* distance of the plane in focus along camera.view (used when lensRadius > 0) */
    int32_t launchPerBounce;        /* fused pipeline, brute-force scenes: 1 = one bounce_kernel launch per bounce instead of the
                                       one-launch bounce loop (a timing toggle: same arithmetic per path, identical images) */
    int32_t depositMode;            /* PtbDepositMode: how the iterations of one wavefront (batchIters > 1) reach image[] */
    int32_t timeStages;             /* 1 = CUDA-event timers around every kernel launch, summed per stage into PtbStats.stageUs
                                       (launches go out one by one instead of as a CUDA graph; INSTRUCTION.md:339-355 analysis) */
    int32_t lazyImage;              /* ptb_pathtrace only: 1 = the host image copy of iteration i is queued on a copy stream and
                                       overlaps iteration i+1; the host image is valid after ptb_sync (or the next call's return
                                       for the previous iteration).  0 = the reference's semantic: valid on return (src/pathtrace.cu:389-390) */
    int32_t regroupMaterials;       /* fused pipeline, BVH scenes (default 0): 1 = the material sort of the fused path (INSTRUCTION.md:78-86): every warp
                                       regroups 128 paths of its CTA's segment by material class (Lambertian / mirror / dielectric) before
                                       shading, so that each arm of the shader runs with full warps; 0 = shade in path order.  Identical images
                                       and per-depth counts either way; scenes with more than 256 materials are shaded in path order */
    int32_t lookahead;              /* ptb_pathtrace only (default 1; needs the fused pipeline, the ordered deposit, batchIters > 1 and a library-owned
                                       accumulation buffer): the call for iteration i has the whole wavefront i .. i + batchIters - 1 traced at the
                                       batched rate on a second stream, and the image after EACH of those iterations built in iteration order; the
                                       next wavefront is traced while this one is handed out, so a call costs its image copy (the strict semantic of
                                       src/pathtrace.cu:389-390 runs at the PCIe rate of the copy).  The reference's loop asks for consecutive
                                       iterations (src/main.cpp:132-136) and an iteration depends only on its number, so every call leaves the bits
                                       of one iteration per call; any other call order, and every other entry point that renders or touches the
                                       accumulation, ends the look-ahead and drops what was traced ahead.  PtbStats counts traced iterations.
                                       0 = every call traces its own iteration */
    int32_t leaveTest;              /* fused pipeline, brute-force scenes (default 1): every path carries the cube face its ray starts on (2 bytes of
                                       path state) and the closest-hit search settles that cube with one row of the object-space ray instead of
                                       walking it as its first candidate (src/intersections.h:48-78: both slab quotients of that axis are negative,
                                       so the reference's test misses).  A pruning decision only: identical images.  0 = the round-1 search */
    int32_t hitQueue;               /* fused pipeline, brute-force scenes, bounces >= 1: PTB_HIT_QUEUE_AUTO (default), 0 or 1.  1 = a warp queues the
                                       hits of its 32-path chunks in shared memory and shades 32 of them at a time, so the lanes whose ray left the
                                       scene do not idle through normal / RNG / scatter (src/pathtrace.cu:227-266 shades path by path; the order in
                                       which paths are shaded reaches no output).  Identical images.  AUTO: on for wavefronts of at least 16 tiles of
                                       256 paths per CTA on scenes without refractive materials -- where it measured faster (cornell: 2.3 %) */
} PtbOptions;
#define PTB_HIT_QUEUE_AUTO (-1)
/* antialias / lensRadius make the camera rays depend on the iteration, so they cannot be combined with
 * firstBounceCache (INSTRUCTION.md:113: "keep in mind how this influences the first-bounce cache"):
 * ptb_create returns PTB_ERR_UNSUPPORTED.  Materials with REFL > 0 and SPECEX > 0 scatter into a Phong
 * lobe around the mirror direction in PTB_BSDF_FULL mode (imperfect specular, INSTRUCTION.md:202-216);
 * SPECEX 0 stays the perfect mirror.  Exact formulas: DESIGN.md section 8 and oracle/pt_oracle.c. */

/* stages of the per-bounce loop as PtbStats.stageUs indexes them (the reference's kernels: src/pathtrace.cu:122, :149, :224, :269;
 * compaction and material sort: INSTRUCTION.md:78-86, 183-185) */
enum PtbStage {
    PTB_STAGE_RAYGEN = 0,           /* raygen_kernel (staged pipeline) */
    PTB_STAGE_INTERSECT = 1,        /* intersect_kernel (staged) / trace_bvh_kernel (fused, BVH scenes) */
    PTB_STAGE_SORT = 2,             /* sort_keys / hist / scan / scatter */
    PTB_STAGE_SHADE = 3,            /* shade_kernel (staged) / bounce_kernel<USE_BVH> (fused, BVH scenes: shade + compact) */
    PTB_STAGE_COMPACT = 4,          /* compact_kernel / carry_count_kernel */
    PTB_STAGE_GATHER = 5,           /* gather_kernel / the ordered-deposit reduce */
    PTB_STAGE_FUSED_BOUNCE0 = 6,    /* bounce_kernel<FIRST>: raygen + intersect + shade + compact of bounce 0 in one launch */
    PTB_STAGE_FUSED_BOUNCES = 7,    /* bounce_loop_kernel / bounce_kernel: the same for bounces 1.. */
    PTB_STAGE_CONTROL = 8,          /* advance / set_control kernels */
    PTB_NUM_STAGES = 12
};

typedef struct PtbStats {
    int64_t iterations;             /* iterations traced since ptb_init */
    int64_t segments;               /* path segments traced (sum over depth of paths entering that depth) */
    int64_t kernelLaunches;         /* kernels this library launched since ptb_init */
    int64_t alivePerDepth[64];      /* paths entering bounce d, summed over iterations */
    int32_t traceDepth;
    int32_t numGeoms;
    int32_t bvhNodes;               /* 0 when brute force */
    int32_t batchIters;
    int64_t pathStateBytes;         /* bytes of path state resident in HBM */
    int64_t sceneUploadBytes;       /* host->device bytes of the last ptb_create / ptb_update_scene */
    /* PtbOptions.timeStages = 1: device time (CUDA events) and launch count per stage since ptb_create / ptb_reset */
    double  stageUs[PTB_NUM_STAGES];
    int64_t stageLaunches[PTB_NUM_STAGES];
} PtbStats;

enum PtbError {
    PTB_OK = 0,
    PTB_ERR_INVALID = -1,           /* bad argument / not initialised */
    PTB_ERR_CUDA = -2,              /* CUDA runtime failure (no device, launch error, ...) */
    PTB_ERR_UNSUPPORTED = -3,
    PTB_ERR_IO = -4
};

typedef struct PtbContext PtbContext;

void        ptb_default_options(PtbOptions *opts);
const char *ptb_last_error(void);
int         ptb_abi_version(void);

/* pathtraceInit (src/pathtrace.cu:79-101): allocate + upload + zero the accumulation buffer.
 * The scene arrays are copied; the caller keeps ownership. */
int ptb_create(PtbContext **ctx, const PtbSceneDesc *scene, const PtbOptions *opts);
/* Re-upload geoms / materials / camera into an existing context and zero the accumulation: what the
 * app does on a camera move by calling pathtraceFree + pathtraceInit (src/main.cpp:125-128), without
 * reallocating.  Counts, resolution and traceDepth must equal those given to ptb_create. */
int ptb_update_scene(PtbContext *ctx, const PtbSceneDesc *scene);
/* pathtraceFree (src/pathtrace.cu:103-112).  NULL is a no-op (main.cpp:126 frees before the first init). */
int ptb_destroy(PtbContext *ctx);

/* pathtrace(pbo, frame, iter) (src/pathtrace.cu:284-393): trace iteration `iter` (1-based,
 * src/main.cpp:132), add it to the accumulation buffer, write the RGBA8 preview into the device
 * buffer `pbo` when non-NULL (sendImageToPBO, :48-68, :386) and copy the accumulated SUM to the host
 * array `hostImage` (W*H*3 floats, the layout of scene->state.image) when non-NULL (:389-390).
 * `frame` is accepted and ignored exactly like the reference (src/main.cpp:136). */
int ptb_pathtrace(PtbContext *ctx, PtbUchar4 *pbo, int frame, int iter, float *hostImage);

/* Batched form of the same loop for headless use: iterations iterBegin, iterBegin+iterStride, ...
 * (count of them) without any host round trip.  iterStride > 1 is the samples-per-pixel split
 * across GPUs (rank r of G renders r+1, r+1+G, ...).  Asynchronous on the context's stream. */
int ptb_render(PtbContext *ctx, int iterBegin, int count, int iterStride);

/* cudaMemcpy(hst_scene->state.image <- dev_image) (src/pathtrace.cu:389-390), decoupled from the loop.
 * Synchronises the stream. */
int ptb_download(PtbContext *ctx, float *hostImage);
/* sendImageToPBO (src/pathtrace.cu:48-68) on demand; `samples` is the divisor `iter`. */
int ptb_write_pbo(PtbContext *ctx, PtbUchar4 *devicePbo, int samples);
/* zero the accumulation buffer and statistics (what re-running pathtraceInit does, src/pathtrace.cu:84-85) */
int ptb_reset(PtbContext *ctx);
int ptb_sync(PtbContext *ctx);
int ptb_stats(PtbContext *ctx, PtbStats *out);          /* synchronises */
/* device pointer of the W*H*3 float accumulation buffer (for a collective reduce across GPUs) */
int ptb_accum_device_ptr(PtbContext *ctx, float **devPtr);
/* (sum / samples) -> clamp[0,1]*255 -> truncate -> RGB8 with the horizontal flip of saveImage
 * (src/main.cpp:83-89, src/image.cpp:27-30) computed on the device; `hostRgb8` = W*H*3 bytes. */
int ptb_resolve_rgb8(PtbContext *ctx, int samples, uint8_t *hostRgb8);

/* ---- several GPUs: samples-per-pixel split + ONE NCCL sum-reduce of the accumulation buffer per frame -------------------
 * Iterations are independent (the RNG is keyed by (iter, pixelIndex, depth), src/pathtrace.cu:41-45) and the only state they
 * share is the additive accumulation buffer (src/pathtrace.cu:276), so GPU g of G renders iterations g+1, g+1+G, ... into a
 * private buffer and ncclReduce(sum, float32, 3*W*H) brings the frame to the root (the reference is single-GPU:
 * src/preview.cpp:106-111 pins device 0).  NCCL (libnccl.so.2) is loaded on first use; single-GPU use never touches it.
 * The reduced frame lives in a second W*H*3 buffer on the root, so the private partial sums keep accumulating across frames.
 *
 * (1) One process per GPU (torchrun / MPI): every rank creates its own context, rank 0 makes an id, the launcher broadcasts it. */
#define PTB_COMM_ID_BYTES 128
int ptb_comm_unique_id(uint8_t id[PTB_COMM_ID_BYTES]);                         /* ncclGetUniqueId */
int ptb_comm_init_rank(PtbContext *ctx, const uint8_t id[PTB_COMM_ID_BYTES], int rank, int nranks);   /* ncclCommInitRank; collective */
/* ncclReduce of this rank's accumulation buffer into the root's frame buffer, on the context's stream (asynchronous; collective) */
int ptb_comm_reduce(PtbContext *ctx, int root);
/* the reduced frame on the root: device pointer / copy of the SUM to the host / saveImage + savePNG pixels computed on the device
 * (sum / samples -> clamp -> * 255 -> truncate, x-flipped: src/main.cpp:83-89, src/image.cpp:27-30) fused after the reduce.
 * Without a communicator (or after ptb_comm_reduce on a 1-rank communicator) the frame is the context's own accumulation. */
int ptb_frame_device_ptr(PtbContext *ctx, float **devPtr);
int ptb_frame_download(PtbContext *ctx, float *hostImage);
int ptb_frame_resolve_rgb8(PtbContext *ctx, int samples, uint8_t *hostRgb8);

/* (2) One process, several devices: n contexts + ncclCommInitAll behind one handle (what `ptb200_headless --gpus G` uses). */
typedef struct PtbMulti PtbMulti;
/* devices: CUDA ordinals (NULL = 0 .. numDevices-1); opts->device / stream / externalAccum are ignored */
int ptb_multi_create(PtbMulti **multi, const PtbSceneDesc *scene, const PtbOptions *opts, const int32_t *devices, int32_t numDevices);
int ptb_multi_destroy(PtbMulti *multi);
int ptb_multi_reset(PtbMulti *multi);
/* iterations iterBegin .. iterBegin+count-1, interleaved over the devices; asynchronous on every device */
int ptb_multi_render(PtbMulti *multi, int iterBegin, int count);
/* one frame: grouped ncclReduce to device 0, then (each optional) the SUM to hostImage and / or the RGB8 pixels to hostRgb8;
 * milliseconds the reduce took on the root's stream go to *reduceMs when non-NULL.  Synchronises. */
int ptb_multi_frame(PtbMulti *multi, int samples, float *hostImage, uint8_t *hostRgb8, float *reduceMs);
int ptb_multi_stats(PtbMulti *multi, PtbStats *sum);                           /* summed over the devices; synchronises */
int ptb_multi_context(PtbMulti *multi, int index, PtbContext **ctx);           /* borrow device `index`'s context (tests) */

/* ---- stage-level entry points (used by the parity tests; host arrays in, host arrays out) ---- */
/* generateRayFromCamera (src/pathtrace.cu:122-143): origins/dirs = W*H*3 floats each.  With antialias /
 * lensRadius set the rays depend on the iteration: ptb_stage_raygen is iteration 1. */
int ptb_stage_raygen(PtbContext *ctx, float *origins, float *dirs);
int ptb_stage_raygen_iter(PtbContext *ctx, int iter, float *origins, float *dirs);
/* computeIntersections (src/pathtrace.cu:149-213) over n caller-supplied rays.
 * accel: PTB_ACCEL_BRUTE / PTB_ACCEL_BVH / PTB_ACCEL_AUTO.  hit = geom index or -1, t = -1 on miss. */
int ptb_stage_intersect(PtbContext *ctx, int accel, const float *origins, const float *dirs, int n,
                        int32_t *hit, float *t, float *normal, int32_t *outside);
/* one shade/scatter step (shadeFakeMaterial's replacement + scatterRay, src/pathtrace.cu:224-266,
 * src/interactions.h:44-79) on caller-supplied hit records; arrays are updated in place. */
int ptb_stage_shade(PtbContext *ctx, int n, const int32_t *hit, const float *t, const float *normal,
                    const int32_t *outside, const int32_t *iter, const int32_t *pixelIndex,
                    const int32_t *depth, float *origins, float *dirs, float *colors, int32_t *remaining);
/* stream compaction (the stream_compaction submodule's job, INSTRUCTION.md:183-185): keeps ids whose
 * flag is non-zero; returns the kept ids (any order) and their count. */
int ptb_stage_compact(PtbContext *ctx, int n, const int32_t *ids, const int32_t *flags,
                      int32_t *outIds, int32_t *outCount);
/* material-id sort (INSTRUCTION.md:78-86): returns a permutation of 0..n-1 grouping equal keys */
int ptb_stage_sort_keys(PtbContext *ctx, int n, const int32_t *keys, int numKeys, int32_t *outOrder);
/* makeSeededRandomEngine + u01 draws on the device (src/pathtrace.cu:41-45, 240-241) */
int ptb_stage_rng(PtbContext *ctx, int iter, int index, int depth, int n, float *out);

/* the device arithmetic the exact tests are built from (csrc/ptb_math.cuh: IEEE division, square root and
 * glm::normalize with a shared range check) on caller-supplied operands: quot[i] = a[i] / b[i],
 * root[i] = sqrt(|a[i]|), normY[i] = normalize(vec3(a[i], b[i], a[i])).y -- checked against correctly
 * rounded host arithmetic (the glm operations of src/intersections.h:48-144 are built from exactly these) */
int ptb_stage_arith(PtbContext *ctx, int n, const float *a, const float *b, float *quot, float *root, float *normY);

/* sin / cos as the hemisphere and lens samplers compute them on the device (src/interactions.h:38-39:
 * `cos(around)`, `sin(around)` -- the host build of the reference calls glibc's sinf / cosf; the device
 * function restates that algorithm in double precision and is bit-identical to it on [0, 2 pi]) */
int ptb_stage_sincos(PtbContext *ctx, int n, const float *angles, float *sinOut, float *cosOut);

/* ---- host-side pieces of the drop-in surface (C++ in csrc/host/, reachable over the C ABI) ---- */
/* new Scene(filename) (src/scene.cpp:7-33); PTB_ERR_IO when the file cannot be opened */
int  ptb_scene_load(const char *path, void **sceneOut);
int  ptb_scene_counts(void *scene, int32_t *numGeoms, int32_t *numMaterials);
int  ptb_scene_get(void *scene, PtbGeom *geoms, PtbMaterial *materials, PtbCamera *camera,
                   int32_t *iterations, int32_t *traceDepth, char *imageName, int nameCap);
/* triangles of the `mesh <file.obj>` objects of a loaded scene: count, then the array and the (first, count)
 * pair of every geom (2 * numGeoms ints) */
int  ptb_scene_mesh_counts(void *scene, int32_t *numTriangles);
int  ptb_scene_get_meshes(void *scene, PtbTriangle *triangles, int32_t *meshRanges);
void ptb_scene_free(void *scene);
/* the camera rebuild main.cpp performs before the first frame (src/main.cpp:53-67, 102-120) */
void ptb_camera_fixup(PtbCamera *camera);
/* saveImage + image::savePNG / saveHDR (src/main.cpp:78-99, src/image.cpp:22-45): `sum` is the W*H*3
 * accumulated image, divided by `samples`; flipX = 1 reproduces the reference's horizontal flip.
 * ".png" / ".hdr" is appended to baseFilename. */
int  ptb_save_png(const char *baseFilename, int w, int h, const float *sum, float samples, int flipX);
int  ptb_save_hdr(const char *baseFilename, int w, int h, const float *sum, float samples, int flipX);

/* The flat BVH the library builds for large scenes (csrc/ptb_bvh.cpp; the reference's loop is brute force,
 * src/pathtrace.cu:176-199, a hierarchical structure is an assignment ask, INSTRUCTION.md:123-140), built on the host
 * over SPHERE / CUBE geoms exactly as ptb_create builds it (maxLeaf = 1 is what the render path uses).  No CUDA call.
 * farPadDiameter > 0: sphere boxes also cover the hits the reference's sphere test (src/intersections.h:102-144) reports for rays
 *   that pass outside the sphere, for ray origins up to that many world units from the sphere's centre; 0: none; < 0: the library's
 *   own choice -- per sphere, the distance to the farthest corner of the box around all geoms (ptb_create adds the camera position) --
 *   or none when PTB_BVH_FAR_PAD=0 is set in the environment.
 * nodes: nodeCap x 16 words, one 64-byte record per interior node in breadth-first order --
 *   left lo[3], left ref (int32), left hi[3], left count (int32), then the same four fields of the right child;
 *   count > 0: a leaf holding slots [ref, ref + count) of `order`; count == 0: ref = index of the child node; ref < 0: no child.
 * order: numGeoms ints, order[slot] = geom index.  nodes / order may be NULL to query *numNodes (<= max(numGeoms, 1)) first. */
int  ptb_bvh_build_host(const PtbGeom *geoms, int32_t numGeoms, int32_t maxLeaf, double farPadDiameter, float *nodes, int32_t nodeCap,
                        int32_t *order, int32_t *numNodes, int32_t *maxDepth);

/* The conservative pruning bounds of one SPHERE / CUBE geom as the kernels receive them (csrc/ptb_bvh.cpp geomBounds / sphereBounds):
 * the padded world box (reach: see farPadDiameter above, <= 0 for none) and, for a sphere, the padded bounding sphere
 * (centre, squared radius) with the relative slack errK of its discriminant (centre / radius2 / errK may be NULL).  No CUDA call. */
int  ptb_geom_bounds_host(const PtbGeom *geom, double reach, float lo[3], float hi[3], float centre[3], float *radius2, float *errK);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
