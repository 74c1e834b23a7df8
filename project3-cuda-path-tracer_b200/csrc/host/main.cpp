// main.cpp -- headless driver: the reference app's main loop (src/main.cpp:33-76 start-up, :101-147
// runCuda, :78-99 saveImage) without GLFW/GLEW/OpenGL.  The reference has no headless mode; the GL
// preview (src/preview.cpp, glslUtility.cpp) is display only and out of scope.
//
//   ptb200_headless SCENEFILE.txt [--iters N] [--depth D] [--bsdf full|golden-diffuse] [--batch B]
//                   [--per-iteration] [--no-lookahead] [--lazy-image] [--out BASENAME] [--hdr] [--staged] [--sort] [--regroup] [--no-compaction]
//                   [--device K] [--gpus G] [--res W H] [--atomic-deposit] [--time-stages] [--quiet]
//
// --gpus G: samples-per-pixel split over devices 0 .. G-1 inside the library (ptb_multi_*: one context per device,
// ncclCommInitAll, interleaved iterations, ONE ncclReduce of the 3*W*H float accumulation buffers to device 0 per frame).
// --res W H overrides the scene file's RES line (BASELINE.json configs[3] is cornell.txt at 3840 x 2160).
//
// Default: renders state.iterations iterations (ITERATIONS of the scene file) through the batched
// loop and writes <FILE>.<UTC start time>.<N>samp.png exactly like saveImage().  --per-iteration uses
// the reference's pathtrace(pbo, frame, iteration) call, one call and one image download per iteration.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <sstream>
#include <string>

#include "ptb200.h"
#include "image.h"
#include "pathtrace.h"
#include "scene.h"

static std::string currentTimeString() {                    // src/preview.cpp:13-19
    time_t now;
    time(&now);
    char buf[sizeof "0000-00-00_00-00-00z"];
    strftime(buf, sizeof buf, "%Y-%m-%d_%H-%M-%Sz", gmtime(&now));
    return std::string(buf);
}

static Scene *scene;
static RenderState *renderState;
static int iteration;
static int width, height;
static std::string startTimeString;

static void saveImage(const std::string &outBase, bool hdr, bool quiet) {   // src/main.cpp:78-99
    float samples = (float)iteration;
    image img(width, height);
    for (int x = 0; x < width; x++)
        for (int y = 0; y < height; y++) {
            int index = x + (y * width);
            glm::vec3 pix = renderState->image[index];
            img.setPixel(width - 1 - x, y, glm::vec3(pix) / samples);
        }
    std::string filename = outBase;
    if (filename.empty()) {
        std::ostringstream ss;
        ss << renderState->imageName << "." << startTimeString << "." << samples << "samp";
        filename = ss.str();
    }
    if (hdr) img.saveHDR(filename, quiet); else img.savePNG(filename, quiet);
}

int main(int argc, char **argv) {
    startTimeString = currentTimeString();
    if (argc < 2) { printf("Usage: %s SCENEFILE.txt [options]\n", argv[0]); return 1; }
    const char *sceneFile = argv[1];
    int iters = -1, depth = -1, batch = 0, device = -1, gpus = 0, resW = 0, resH = 0;
    bool perIteration = false, hdr = false, quiet = false;
    std::string outBase;
    PtbOptions opt;
    ptb_default_options(&opt);
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--iters") iters = atoi(next());
        else if (a == "--depth") depth = atoi(next());
        else if (a == "--batch") batch = atoi(next());
        else if (a == "--device") device = atoi(next());
        else if (a == "--gpus") gpus = atoi(next());
        else if (a == "--res") { resW = atoi(next()); resH = atoi(next()); }
        else if (a == "--atomic-deposit") opt.depositMode = PTB_DEPOSIT_ATOMIC;
        else if (a == "--time-stages") opt.timeStages = 1;
        else if (a == "--lazy-image") opt.lazyImage = 1;              // with --per-iteration: the image copy of iteration i overlaps iteration i + 1
        else if (a == "--out") outBase = next();
        else if (a == "--bsdf") opt.bsdfMode = strcmp(next(), "golden-diffuse") == 0 ? PTB_BSDF_GOLDEN_DIFFUSE : PTB_BSDF_FULL;
        else if (a == "--per-iteration") perIteration = true;
        else if (a == "--hdr") hdr = true;
        else if (a == "--staged") opt.pipeline = PTB_PIPELINE_STAGED;
        else if (a == "--sort") opt.sortMaterials = 1;
        else if (a == "--no-lookahead") opt.lookahead = 0;             // with --per-iteration: every pathtrace() call traces its own iteration
        else if (a == "--regroup") opt.regroupMaterials = 1;           // fused pipeline, BVH scenes: regroup the paths by material class before shading
        else if (a == "--no-compaction") opt.compaction = 0;
        else if (a == "--quiet") quiet = true;
        else if (a == "--cache-first-bounce") opt.firstBounceCache = 1;
        else if (a == "--aa") opt.antialias = 1;                                        // stochastic antialiasing
        else if (a == "--lens") { opt.lensRadius = (float)atof(next()); opt.focalDistance = (float)atof(next()); }   // --lens RADIUS FOCAL_DISTANCE
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    // --per-iteration: pathtrace() per iteration as the reference's main loop calls it; with the look-ahead (default) the library traces whole
    // wavefronts behind that call, without it every call traces one iteration
    opt.batchIters = (perIteration && !opt.lookahead) ? 1 : batch;
    opt.device = device;

    scene = new Scene(sceneFile, !quiet);                      // src/main.cpp:44
    if (!scene->ok) { fprintf(stderr, "Error reading from file - aborting!\n"); return 1; }
    iteration = 0;
    renderState = &scene->state;
    Camera &cam = renderState->camera;
    if (resW > 0 && resH > 0) {
        // what loadCamera derives from RES (src/scene.cpp:133-140): fovx from the aspect ratio, pixelLength, the host image
        cam.resolution.x = resW; cam.resolution.y = resH;
        const float yscaled = tan(cam.fov.y * (PI / 180));
        const float xscaled = (yscaled * resW) / resH;
        cam.fov.x = (atan(xscaled) * 180) / PI;
        cam.pixelLength = glm::vec2(2 * xscaled / (float)resW, 2 * yscaled / (float)resH);
        renderState->image.assign((size_t)resW * resH, glm::vec3());
    }
    width = cam.resolution.x;
    height = cam.resolution.y;
    if (iters >= 0) renderState->iterations = (unsigned)iters;
    if (depth > 0) renderState->traceDepth = depth;

    // src/main.cpp:53-67 + :102-120 -- the camera is rebuilt from (zoom, phi, theta) before frame 1
    ptb_camera_fixup(reinterpret_cast<PtbCamera *>(&cam));

    if (gpus > 0) {
        // several devices: the library's own multi-GPU path (no pathtrace() shim: that is the reference's single-device call)
        PtbSceneDesc desc;
        pathtraceSceneDesc(scene, &desc);
        PtbMulti *multi = NULL;
        if (ptb_multi_create(&multi, &desc, &opt, NULL, gpus) != PTB_OK) { fprintf(stderr, "ptb_multi_create: %s\n", ptb_last_error()); return 1; }
        auto m0 = std::chrono::steady_clock::now();
        float reduceMs = 0.f;
        iteration = (int)renderState->iterations;
        if (ptb_multi_render(multi, 1, iteration) != PTB_OK ||
            ptb_multi_frame(multi, iteration, reinterpret_cast<float *>(renderState->image.data()), NULL, &reduceMs) != PTB_OK) {
            fprintf(stderr, "multi-GPU render: %s\n", ptb_last_error());
            return 1;
        }
        const double msecs = std::chrono::duration<double>(std::chrono::steady_clock::now() - m0).count();
        PtbStats mst;
        ptb_multi_stats(multi, &mst);
        if (!quiet) printf("%d GPUs: %d iterations in %.3f s: %.1f iterations/s, %.4e path segments/s, NCCL reduce of the %.1f MB frame %.3f ms\n",
                           gpus, iteration, msecs, iteration / msecs, mst.segments / msecs, width * height * 12 / 1e6, reduceMs);
        if (iteration > 0) saveImage(outBase, hdr, quiet);
        ptb_multi_destroy(multi);
        return 0;
    }

    pathtraceSetOptions(&opt);
    pathtraceFree();                                           // src/main.cpp:126-127 (free precedes the first init)
    pathtraceInit(scene);

    auto t0 = std::chrono::steady_clock::now();
    if (perIteration) {
        while (iteration < (int)renderState->iterations) {     // src/main.cpp:131-140
            iteration++;
            pathtrace(NULL, 0, iteration);
        }
        pathtraceSync();
    } else {
        pathtraceBatch(1, (int)renderState->iterations);
        iteration = (int)renderState->iterations;
        pathtraceDownload();
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    PtbStats st;
    ptb_stats(pathtraceContext(), &st);
    if (!quiet && opt.timeStages) {
        static const char *names[] = {"raygen", "intersect", "sort", "shade", "compact", "gather", "fused bounce 0", "fused bounces 1..", "control"};
        for (int k = 0; k < 9; ++k) if (st.stageLaunches[k]) printf("stage %-18s %8lld launches %12.1f us\n", names[k], (long long)st.stageLaunches[k], st.stageUs[k]);
    }
    if (!quiet) {
        // (the statistics count traced iterations: the look-ahead traces up to two wavefronts more than it hands out)
        printf("%d iterations in %.3f s: %.1f iterations/s, %.4e path segments/s\n", iteration, secs, iteration / secs,
               (double)st.segments / (double)(st.iterations ? st.iterations : 1) * iteration / secs);
        printf("paths entering each bounce (per iteration):");
        for (int d = 0; d < st.traceDepth; ++d) printf(" %.0f", (double)st.alivePerDepth[d] / (double)(st.iterations ? st.iterations : 1));
        printf("\n");
    }
    if (iteration > 0) saveImage(outBase, hdr, quiet);         // src/main.cpp:141-142
    pathtraceFree();
    return 0;
}
