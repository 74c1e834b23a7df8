// pathtrace.h -- the renderer's three entry points, exactly as the reference declares them
// (src/pathtrace.h:6-8).  Implemented in pathtrace.cpp over the C ABI of libptb200.so.
#pragma once
#include <vector>
#include <cuda_runtime.h>     // uchar4
#include "scene.h"

void pathtraceInit(Scene *scene);
void pathtraceFree();
void pathtrace(uchar4 *pbo, int frame, int iteration);

// Extras for headless use (not in the reference): options for the next pathtraceInit, the batched
// loop without the per-iteration device->host copy, and access to the underlying context.
struct PtbOptions;
struct PtbContext;
struct PtbSceneDesc;
// what pathtraceInit hands to the library: the POD view of `scene` (geoms, materials, camera, depth, meshes); borrows scene's arrays
void pathtraceSceneDesc(Scene *scene, PtbSceneDesc *desc);
void pathtraceSetOptions(const PtbOptions *opts);
void pathtraceBatch(int iterBegin, int count);        // then pathtraceDownload() fills scene->state.image
void pathtraceDownload();
void pathtraceSync();                                 // PtbOptions.lazyImage: scene->state.image is valid after this
PtbContext *pathtraceContext();
