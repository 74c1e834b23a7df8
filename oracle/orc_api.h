/*
 * orc_api.h -- the C interface shared by the two CPU checkers under oracle/.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries, and only as the checker.
 *
 * Two libraries export exactly this interface:
 *   oracle/_ref/libptref.so   built by oracle/Makefile from ref_driver.cpp, which
 *                             #includes the reference's own headers and links its
 *                             scene.cpp / utilities.cpp where they lie under
 *                             /root/reference (kind = "reference").
 *   oracle/libptoracle.so     built from pt_oracle.c, a plain-C restatement with no
 *                             glm / thrust / reference headers (kind = "port").
 * tests/test_oracle_vs_ref.py requires them to agree bit for bit.
 *
 * The POD structs mirror /root/reference/src/sceneStructs.h byte for byte:
 *   OrcGeom     <-> struct Geom      (sceneStructs.h:20-29)  236 B
 *   OrcMaterial <-> struct Material  (sceneStructs.h:31-41)   44 B
 *   OrcCamera   <-> struct Camera    (sceneStructs.h:43-52)   84 B
 * Matrices are column-major (glm): m[4*c + r].
 */
#ifndef ORC_API_H
#define ORC_API_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcGeom {
    int   type;            /* 0 = SPHERE, 1 = CUBE (sceneStructs.h:10-13) */
    int   materialid;
    float translation[3];
    float rotation[3];
    float scale[3];
    float transform[16];
    float inverseTransform[16];
    float invTranspose[16];
} OrcGeom;

typedef struct OrcMaterial {
    float color[3];
    float specExponent;
    float specColor[3];
    float hasReflective;
    float hasRefractive;
    float indexOfRefraction;
    float emittance;
} OrcMaterial;

typedef struct OrcCamera {
    int   resolution[2];
    float position[3];
    float lookAt[3];
    float view[3];
    float up[3];
    float right[3];
    float fov[2];
    float pixelLength[2];
} OrcCamera;

enum { ORC_MODE_GOLDEN_DIFFUSE = 0, ORC_MODE_FULL = 1 };

/* Camera-ray extensions the reference leaves as TODOs in generateRayFromCamera
 * (src/pathtrace.cu:118-120, 134; INSTRUCTION.md:109-113): stochastic antialiasing and a thin
 * lens.  Semantics authored here; all zero = the reference's pinhole ray. */
typedef struct OrcLens {
    int   antialias;       /* 1: jitter the sample position uniformly inside the pixel */
    float lensRadius;      /* > 0: thin-lens depth of field (PBRT 6.2.3) */
    float focalDistance;   /* distance of the plane in focus, measured along cam.view */
} OrcLens;
/* Triangle meshes: the scene format reserves the object type `mesh` (INSTRUCTION.md:291-293) and the
 * assignment names glm::intersectRayTriangle for it (INSTRUCTION.md:123-127); the reference's loader and
 * kernels implement neither.  A geom with type ORC_MESH owns triangles[first .. first+count) of OrcMeshes
 * (object-space vertices, counter-clockwise = front face), ranges = 2 ints (first, count) per geom. */
#define ORC_MESH 2
typedef struct OrcTriangle { float v0[3], v1[3], v2[3]; } OrcTriangle;
typedef struct OrcMeshes {
    const OrcTriangle *triangles;
    int                ntriangles;
    const int         *ranges;
} OrcMeshes;

/* depth key of the camera-ray engine: makeSeededRandomEngine(iter, index, ORC_RAYGEN_DEPTH_KEY).
 * Bounce engines use depth 0 .. traceDepth-1 <= 63; the key field is 9 bits wide (src/pathtrace.cu:43). */
#define ORC_RAYGEN_DEPTH_KEY 511

/* "reference" or "port" */
const char *orc_kind(void);

/* ---- scene file -> POD (src/scene.cpp) ---- */
void *orc_scene_load(const char *path);                 /* NULL if the file cannot be opened */
void  orc_scene_counts(void *scene, int *ngeoms, int *nmaterials);
void  orc_scene_get(void *scene, OrcGeom *geoms, OrcMaterial *materials, OrcCamera *cam,
                    int *iterations, int *traceDepth, char *imageName, int nameCap);
void  orc_scene_free(void *scene);

/* camera fix-up the app performs before frame 1 (src/main.cpp:53-67, 102-120) */
void  orc_camera_fixup(OrcCamera *cam);

/* buildTransformationMatrix + glm::inverse + glm::inverseTranspose
 * (src/utilities.cpp:65-72, src/scene.cpp:82-85) */
void  orc_build_transform(const float t[3], const float r[3], const float s[3],
                          float transform[16], float inverse[16], float invTranspose[16]);

/* ---- hot-path pieces ---- */
/* makeSeededRandomEngine + n draws of u01 (src/pathtrace.cu:41-45, 240-241) */
void  orc_rng_u01(int iter, int index, int depth, int n, float *out);

/* generateRayFromCamera (src/pathtrace.cu:122-143): origins/dirs are P*3 floats */
void  orc_generate_rays(const OrcCamera *cam, float *origins, float *dirs);

/* the same with the OrcLens extensions, for iteration `iter` (lens may be NULL) */
void  orc_generate_rays_lens(const OrcCamera *cam, const OrcLens *lens, int iter, float *origins, float *dirs);

/* computeIntersections closest-hit loop (src/pathtrace.cu:162-211) over arbitrary rays.
 * hit[i] = geom index or -1; t[i] = t_min or -1; normal = 3 floats; outside = flag of the
 * winning test (the reference drops it; refraction needs it). */
void  orc_closest_hit(const OrcGeom *geoms, int ngeoms, const float *origins, const float *dirs,
                      int n, int *hit, float *t, float *normal, int *outside);

/* calculateRandomDirectionInHemisphere (src/interactions.h:10-42) with the engine seeded
 * by (iter[i], index[i], depth[i]) */
void  orc_hemisphere(const float *normals, int n, const int *iter, const int *index,
                     const int *depth, float *outDirs);

/* One shade/scatter step on given inputs:
 *   miss (hit < 0)        : colour = 0, remaining = 0
 *   emissive material     : colour *= m.color * m.emittance, remaining = 0
 *   otherwise             : scatter (new origin/direction, colour *= albedo), remaining -= 1,
 *                           and colour = 0 when remaining reaches 0
 * in/out arrays are per path; the engine is seeded by (iter[i], index[i], depth[i]). */
void  orc_shade(const OrcMaterial *materials, int mode, int n,
                const int *hit, const int *materialIds, const float *t,
                const float *normal, const int *outside,
                const int *iter, const int *index, const int *depth,
                float *origins, float *dirs, float *colors, int *remaining);

/* The normative multi-bounce loop.  image (P*3 floats) is ACCUMULATED
 * into (+=), one iteration after another in ascending order; alivePerDepth[d] += paths that
 * enter bounce d (traceDepth entries).  Returns the number of path segments traced. */
long long orc_render(const OrcGeom *geoms, int ngeoms, const OrcMaterial *materials, int nmaterials,
                     const OrcCamera *cam, int traceDepth, int mode,
                     int iterBegin, int nIters, float *image, long long *alivePerDepth,
                     int nthreads);

/* orc_closest_hit over a scene with ORC_MESH geoms (meshes may be NULL when there are none).  The mesh test
 * (authored) follows the structure of the reference's two tests: ray into object space,
 * direction normalised there, glm::intersectRayTriangle per triangle (front winding, then the reversed
 * winding for the back face), smallest object-space t > 0 wins with the lowest triangle index on ties,
 * hit point = transform * getPointOnRay, normal = normalize(invTranspose * cross(e1, e2)) facing the incoming
 * ray, returned t = world distance. */
void  orc_closest_hit_mesh(const OrcGeom *geoms, int ngeoms, const OrcMeshes *meshes,
                           const float *origins, const float *dirs, int n,
                           int *hit, float *t, float *normal, int *outside);

/* orc_render with the extensions: meshes and / or lens may be NULL (both NULL: identical to orc_render) */
long long orc_render_ex(const OrcGeom *geoms, int ngeoms, const OrcMeshes *meshes,
                        const OrcMaterial *materials, int nmaterials,
                        const OrcCamera *cam, const OrcLens *lens, int traceDepth, int mode,
                        int iterBegin, int nIters, float *image, long long *alivePerDepth,
                        int nthreads);

/* ---- plain-C restatement only (libptoracle.so) ----
 * sin / cos of the samplers: restated = 0 -> libm's sinf / cosf (what the loops above call); restated = 1 ->
 * the double-precision restatement of glibc's algorithm that the CUDA path mirrors.  orc_sincos_exhaustive
 * counts the floats with bit pattern in [loBits, hiBits] on which the two disagree. */
void  orc_sincos(const float *y, int n, int restated, float *sn, float *cs);
long long orc_sincos_exhaustive(unsigned loBits, unsigned hiBits, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
