// pathtrace.cpp -- pathtraceInit / pathtraceFree / pathtrace (src/pathtrace.h:6-8) as thin C++
// wrappers over the C ABI.  It touches only Scene's public members (geoms, materials, state), so it
// compiles unchanged against the reference's own scene.h/sceneStructs.h (define PTB_USE_REAL_GLM and
// put the reference's src/ first on the include path) -- see INTEGRATION.md.
// Error behaviour follows checkCUDAError (src/pathtrace.cu:21-39): print and exit(EXIT_FAILURE).
#include "pathtrace.h"
#include "ptb200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

static Scene *hst_scene = NULL;
static PtbContext *g_ctx = NULL;
static PtbOptions g_opts;
static bool g_haveOpts = false;

static void check(int rc, const char *what) {
    if (rc == PTB_OK) return;
    fprintf(stderr, "CUDA error (pathtrace.cpp): %s: %s\n", what, ptb_last_error());
    exit(EXIT_FAILURE);
}

void pathtraceSetOptions(const PtbOptions *opts) {
    if (opts) { g_opts = *opts; g_haveOpts = true; } else g_haveOpts = false;
}

void pathtraceSceneDesc(Scene *scene, PtbSceneDesc *desc) {
    PtbSceneDesc &d = *desc;
    memset(&d, 0, sizeof d);
    d.geoms = reinterpret_cast<const PtbGeom *>(scene->geoms.data());
    d.numGeoms = (int)scene->geoms.size();
    d.materials = reinterpret_cast<const PtbMaterial *>(scene->materials.data());
    d.numMaterials = (int)scene->materials.size();
    memcpy(&d.camera, &scene->state.camera, sizeof(PtbCamera));
    d.traceDepth = scene->state.traceDepth;
#ifdef PTB_SCENE_HAS_MESHES                                      // this repo's scene.h; the reference's Scene has no meshes
    d.numTriangles = (int)scene->triangles.size();
    d.triangles = reinterpret_cast<const PtbTriangle *>(scene->triangles.data());
    d.meshRanges = scene->meshRanges.data();
#endif
}

void pathtraceInit(Scene *scene) {
    static_assert(sizeof(Geom) == sizeof(PtbGeom) && sizeof(Material) == sizeof(PtbMaterial) && sizeof(Camera) == sizeof(PtbCamera),
                  "boundary structs must stay byte-identical to src/sceneStructs.h");
    if (g_ctx) { ptb_destroy(g_ctx); g_ctx = NULL; }          // re-init on camera change resets the accumulation (src/main.cpp:125-128)
    hst_scene = scene;
    PtbSceneDesc d;
    pathtraceSceneDesc(scene, &d);
    PtbOptions o;
    if (g_haveOpts) o = g_opts; else ptb_default_options(&o);
    check(ptb_create(&g_ctx, &d, &o), "pathtraceInit");
}

void pathtraceFree() {
    // called before the first pathtraceInit by the app (src/main.cpp:126-127): must be a no-op then
    check(ptb_destroy(g_ctx), "pathtraceFree");
    g_ctx = NULL;
}

void pathtrace(uchar4 *pbo, int frame, int iteration) {
    // post-condition (src/pathtrace.cu:276, 389-390): state.image holds the SUM over iterations
    check(ptb_pathtrace(g_ctx, reinterpret_cast<PtbUchar4 *>(pbo), frame, iteration,
                        reinterpret_cast<float *>(hst_scene->state.image.data())), "pathtrace");
}

void pathtraceBatch(int iterBegin, int count) { check(ptb_render(g_ctx, iterBegin, count, 1), "pathtraceBatch"); }
void pathtraceDownload() { check(ptb_download(g_ctx, reinterpret_cast<float *>(hst_scene->state.image.data())), "pathtraceDownload"); }
void pathtraceSync() { check(ptb_sync(g_ctx), "pathtraceSync"); }
PtbContext *pathtraceContext() { return g_ctx; }
