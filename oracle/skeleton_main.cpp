/*
 * skeleton_main.cpp -- TEST/BENCH INFRASTRUCTURE.  Headless driver for the UNMODIFIED reference
 * renderer (src/pathtrace.cu of the reference, compiled as it lies for sm_100 by oracle/Makefile):
 * the denominator of BASELINE.json's ">= 10x the reference skeleton's own CUDA loop rebuilt for
 * sm_100" target (SURVEY.md 8(d) "Skeleton GPU baseline").
 *
 * It does what the GL app does around the renderer, minus OpenGL: load the scene (src/scene.cpp),
 * rebuild the camera exactly as src/main.cpp:53-67,102-120 does before the first frame, then call
 * pathtraceFree(); pathtraceInit(scene); and pathtrace(pbo, 0, iteration) for iteration = 1..N with a
 * plain cudaMalloc'd uchar4 buffer standing in for the mapped GL PBO (src/main.cpp:125-140).
 * The skeleton traces ONE bounce per iteration (src/pathtrace.cu:376), so segments = pixels * iters.
 * Output: one JSON line.
 */
#include "scene.h"
#include "pathtrace.h"
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

Scene::~Scene() {}            /* declared but never defined by the reference (src/scene.h:21) */

int main(int argc, char **argv) {
    if (argc < 2) { printf("usage: %s SCENEFILE.txt [iterations] [warmup]\n", argv[0]); return 1; }
    const int iters = argc > 2 ? atoi(argv[2]) : 1000;
    const int warm = argc > 3 ? atoi(argv[3]) : 50;
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());                      /* the loader prints several lines per object */
    Scene *scene = new Scene(argv[1]);
    std::cout.rdbuf(old);
    Camera &cam = scene->state.camera;

    /* src/main.cpp:53-67 */
    glm::vec3 view = cam.view;
    glm::vec3 viewXZ = glm::vec3(view.x, 0.0f, view.z);
    glm::vec3 viewZY = glm::vec3(0.0f, view.y, view.z);
    float phi = glm::acos(glm::dot(glm::normalize(viewXZ), glm::vec3(0, 0, -1)));
    float theta = glm::acos(glm::dot(glm::normalize(viewZY), glm::vec3(0, 1, 0)));
    float zoom = glm::length(cam.position - cam.lookAt);
    /* src/main.cpp:102-120 */
    glm::vec3 cameraPosition;
    cameraPosition.x = zoom * sin(phi) * sin(theta);
    cameraPosition.y = zoom * cos(theta);
    cameraPosition.z = zoom * cos(phi) * sin(theta);
    cam.view = -glm::normalize(cameraPosition);
    glm::vec3 v = cam.view, u = glm::vec3(0, 1, 0);
    glm::vec3 r = glm::cross(v, u);
    cam.up = glm::cross(r, v);
    cam.right = r;
    cameraPosition += cam.lookAt;
    cam.position = cameraPosition;

    const int pixels = cam.resolution.x * cam.resolution.y;
    uchar4 *pbo = NULL;
    if (cudaMalloc(&pbo, pixels * sizeof(uchar4)) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 2; }

    pathtraceFree();                                     /* src/main.cpp:126-127 */
    pathtraceInit(scene);
    int iteration = 0;
    for (int i = 0; i < warm; ++i) pathtrace(pbo, 0, ++iteration);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) pathtrace(pbo, 0, ++iteration);
    cudaDeviceSynchronize();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double sum = 0;
    for (int i = 0; i < pixels; ++i) sum += scene->state.image[i].x + scene->state.image[i].y + scene->state.image[i].z;
    pathtraceFree();
    printf("{\"impl\": \"reference-skeleton-cuda\", \"scene\": \"%s\", \"pixels\": %d, \"iterations\": %d, \"seconds\": %.6f, "
           "\"iters_per_s\": %.3f, \"segments_per_s\": %.6e, \"bounces_per_iteration\": 1, \"image_sum\": %.6e}\n",
           scene->state.imageName.c_str(), pixels, iters, secs, iters / secs, (double)pixels * iters / secs, sum);
    return 0;
}
