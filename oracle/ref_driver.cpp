/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE (kind = "reference").  Not part of the product.
 *
 * A thin host driver around the REFERENCE'S OWN __host__ __device__ functions, compiled by
 * oracle/Makefile from the sources where they lie under /root/reference (nothing is copied
 * into this repository):
 *     #include  src/scene.h  src/intersections.h  src/interactions.h      (by -I path)
 *     links     src/scene.cpp  src/utilities.cpp
 * Output: oracle/_ref/libptref.so, exporting the interface in orc_api.h.
 *
 * What comes straight from the reference (called, not restated):
 *     Scene loader, buildTransformationMatrix, glm::inverse / inverseTranspose  (scene.cpp, utilities.cpp)
 *     utilhash, getPointOnRay, multiplyMV, boxIntersectionTest, sphereIntersectionTest (intersections.h)
 *     calculateRandomDirectionInHemisphere                                      (interactions.h:10-42)
 *     thrust::default_random_engine / uniform_real_distribution                 (CUDA toolkit thrust)
 *     glm::reflect / glm::refract / normalize / cross / dot                     (vendored glm 0.9.6.3)
 * What the reference does not have as host-callable code and is therefore restated here with the
 * same glm calls, each citing the lines it follows:
 *     makeSeededRandomEngine   (src/pathtrace.cu:41-45; lives in the .cu next to the kernels)
 *     camera ray generation    (src/pathtrace.cu:122-143; a __global__ kernel)
 *     closest-hit loop         (src/pathtrace.cu:162-211; a __global__ kernel)
 *     camera fix-up            (src/main.cpp:53-67, 102-120; main.cpp needs GL to compile)
 *     the multi-bounce loop and scatterRay body, which the reference leaves as TODOs
 *     (src/pathtrace.cu:339-377, src/interactions.h:69-79) -- SURVEY.md section 9 is the spec.
 *
 * Build flags are pinned in the Makefile: -O2 -ffp-contract=off, no -march=native, no fast-math.
 */
#include "scene.h"                 /* first: brings `using namespace std` (src/scene.h:11) */
#include <thrust/random.h>
#include "intersections.h"
#include "interactions.h"
#include <glm/gtc/matrix_inverse.hpp>
#include <glm/gtx/intersect.hpp>    /* glm::intersectRayTriangle, the function INSTRUCTION.md:123-127 names for meshes */

#include <cfloat>
#include <cstring>
#include <cstdio>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_api.h"

static_assert(sizeof(Geom) == sizeof(OrcGeom), "Geom layout");
static_assert(sizeof(Material) == sizeof(OrcMaterial), "Material layout");
static_assert(sizeof(Camera) == sizeof(OrcCamera), "Camera layout");
static_assert(sizeof(Geom) == 236 && sizeof(Material) == 44 && sizeof(Camera) == 84, "POD sizes");

/* src/scene.h:21 declares ~Scene() but the reference never defines it. */
Scene::~Scene() {}

/* src/pathtrace.cu:41-45 */
static thrust::default_random_engine seededEngine(int iter, int index, int depth) {
    int h = utilhash((1 << 31) | (depth << 22) | iter) ^ utilhash(index);
    return thrust::default_random_engine(h);
}

static inline glm::vec3 v3(const float *p) { return glm::vec3(p[0], p[1], p[2]); }
static inline void put3(float *p, const glm::vec3 &v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

/* The mesh test (authored, see orc_api.h) written like the reference writes its two tests
 * (src/intersections.h:48-144), on the vendored glm::intersectRayTriangle. */
static float meshIntersectionTest(const Geom &mesh, const OrcTriangle *tris, int count, const Ray &r,
                                  glm::vec3 &intersectionPoint, glm::vec3 &normal, bool &outside) {
    Ray q;
    q.origin = multiplyMV(mesh.inverseTransform, glm::vec4(r.origin, 1.0f));
    q.direction = glm::normalize(multiplyMV(mesh.inverseTransform, glm::vec4(r.direction, 0.0f)));
    float tmin = FLT_MAX;
    int best = -1;
    bool bestFront = true;
    for (int i = 0; i < count; i++) {
        glm::vec3 v0 = v3(tris[i].v0), v1 = v3(tris[i].v1), v2 = v3(tris[i].v2);
        glm::vec3 bary(0.0f);
        bool front = true;
        bool hit = glm::intersectRayTriangle(q.origin, q.direction, v0, v1, v2, bary);
        if (!hit) { front = false; hit = glm::intersectRayTriangle(q.origin, q.direction, v0, v2, v1, bary); }
        if (hit && bary.z > 0.0f && bary.z < tmin) { tmin = bary.z; best = i; bestFront = front; }
    }
    if (best < 0) return -1;
    glm::vec3 v0 = v3(tris[best].v0), v1 = v3(tris[best].v1), v2 = v3(tris[best].v2);
    intersectionPoint = multiplyMV(mesh.transform, glm::vec4(getPointOnRay(q, tmin), 1.0f));
    glm::vec3 n = glm::normalize(multiplyMV(mesh.invTranspose, glm::vec4(glm::cross(v1 - v0, v2 - v0), 0.0f)));
    normal = bestFront ? n : -n;
    outside = bestFront;
    return glm::length(r.origin - intersectionPoint);
}

/* src/pathtrace.cu:162-211 */
static int closestHit(const Geom *geoms, int ngeoms, const OrcMeshes *meshes, const Ray &ray,
                      float &tOut, glm::vec3 &nOut, bool &outsideOut) {
    float t;
    glm::vec3 normal;
    float t_min = FLT_MAX;
    int hit_geom_index = -1;
    bool outside = true;
    bool hitOutside = true;
    glm::vec3 tmp_intersect, tmp_normal;
    for (int i = 0; i < ngeoms; i++) {
        const Geom &geom = geoms[i];
        t = -1.0f;
        if (geom.type == CUBE) {
            t = boxIntersectionTest(geom, ray, tmp_intersect, tmp_normal, outside);
        } else if (geom.type == SPHERE) {
            t = sphereIntersectionTest(geom, ray, tmp_intersect, tmp_normal, outside);
        } else if ((int)geom.type == ORC_MESH && meshes) {
            t = meshIntersectionTest(geom, meshes->triangles + meshes->ranges[2 * i], meshes->ranges[2 * i + 1], ray,
                                     tmp_intersect, tmp_normal, outside);
        }
        if (t > 0.0f && t_min > t) {
            t_min = t;
            hit_geom_index = i;
            normal = tmp_normal;
            hitOutside = outside;   /* flag of the winning test; see SURVEY.md 8(a) a14 */
        }
    }
    if (hit_geom_index == -1) {
        tOut = -1.0f;
        nOut = glm::vec3(0.0f);
        outsideOut = true;
    } else {
        tOut = t_min;
        nOut = normal;
        outsideOut = hitOutside;
    }
    return hit_geom_index;
}

/* src/pathtrace.cu:131-138 */
static inline Ray cameraRay(const Camera &cam, int x, int y) {
    Ray r;
    r.origin = cam.position;
    r.direction = glm::normalize(cam.view
        - cam.right * cam.pixelLength.x * ((float)x - (float)cam.resolution.x * 0.5f)
        - cam.up * cam.pixelLength.y * ((float)y - (float)cam.resolution.y * 0.5f));
    return r;
}

/* Camera ray with the OrcLens extensions (orc_api.h); same formulas as pt_oracle.c::cameraRayLens, written
 * with the reference's glm / thrust calls. */
static inline Ray cameraRayLens(const Camera &cam, const OrcLens *lens, int x, int y, int iter) {
    if (!lens || (!lens->antialias && !(lens->lensRadius > 0.0f))) return cameraRay(cam, x, y);
    thrust::default_random_engine rng = seededEngine(iter, x + y * cam.resolution.x, ORC_RAYGEN_DEPTH_KEY);
    thrust::uniform_real_distribution<float> u01(0, 1);
    float sx = (float)x, sy = (float)y;
    if (lens->antialias) {
        sx = sx + (u01(rng) - 0.5f);
        sy = sy + (u01(rng) - 0.5f);
    }
    Ray r;
    r.origin = cam.position;
    r.direction = glm::normalize(cam.view
        - cam.right * cam.pixelLength.x * (sx - (float)cam.resolution.x * 0.5f)
        - cam.up * cam.pixelLength.y * (sy - (float)cam.resolution.y * 0.5f));
    if (lens->lensRadius > 0.0f) {
        float rad = lens->lensRadius * sqrtf(u01(rng));
        float phi = u01(rng) * TWO_PI;
        float lx = rad * cosf(phi), ly = rad * sinf(phi);
        float ft = lens->focalDistance / glm::dot(r.direction, cam.view);
        glm::vec3 focus = r.origin + r.direction * ft;
        r.origin = r.origin + cam.right * lx + cam.up * ly;
        r.direction = glm::normalize(focus - r.origin);
    }
    return r;
}

/* Imperfect specular lobe (INSTRUCTION.md:202-216; GPU Gems 3 ch. 20 eq. 7-9): the tangent frame and the
 * combination are those of calculateRandomDirectionInHemisphere (src/interactions.h:21-41) with
 * cos(theta) = u1^(1/(n+1)) in place of sqrt(u1). */
static inline glm::vec3 phongLobe(const glm::vec3 &axis, float exponent, thrust::default_random_engine &rng) {
    thrust::uniform_real_distribution<float> u01(0, 1);
    float up = powf(u01(rng), 1.0f / (exponent + 1.0f));
    float s = 1 - up * up;
    float over = s > 0.0f ? sqrtf(s) : 0.0f;
    float around = u01(rng) * TWO_PI;
    glm::vec3 directionNotNormal;
    if (abs(axis.x) < SQRT_OF_ONE_THIRD) directionNotNormal = glm::vec3(1, 0, 0);
    else if (abs(axis.y) < SQRT_OF_ONE_THIRD) directionNotNormal = glm::vec3(0, 1, 0);
    else directionNotNormal = glm::vec3(0, 0, 1);
    glm::vec3 p1 = glm::normalize(glm::cross(axis, directionNotNormal));
    glm::vec3 p2 = glm::normalize(glm::cross(axis, p1));
    return up * axis + cos(around) * over * p1 + sin(around) * over * p2;
}

/* The scatterRay body the reference leaves empty (src/interactions.h:44-79), per SURVEY.md section 9.
 * Returns nothing; updates ray + colour in place.  `t` is the hit distance along `ray`. */
static inline void scatter(Ray &ray, glm::vec3 &color, float t, const glm::vec3 &n, bool outside,
                           const Material &m, int mode, thrust::default_random_engine &rng) {
    glm::vec3 P = getPointOnRay(ray, t);
    if (mode == ORC_MODE_FULL && m.hasRefractive > 0.0f) {
        /* dielectric: Schlick-weighted choice between glm::reflect and glm::refract
         * (INSTRUCTION.md:105-108) */
        thrust::uniform_real_distribution<float> u01(0, 1);
        glm::vec3 I = glm::normalize(ray.direction);
        float ior = m.indexOfRefraction;
        float eta = outside ? 1.0f / ior : ior;
        float cosi = -glm::dot(n, I);
        float r0 = (1.0f - ior) / (1.0f + ior);
        r0 = r0 * r0;
        float c1 = 1.0f - cosi;
        float c2 = c1 * c1;
        float c5 = c2 * c2 * c1;
        float R = r0 + (1.0f - r0) * c5;
        float u = u01(rng);
        /* total internal reflection <=> the k of glm::refract is negative
         * (glm/detail/func_geometric.inl:197-199; there sqrt(k)*0 would give NaN) */
        float dotNI = glm::dot(n, I);
        float k = 1.0f - eta * eta * (1.0f - dotNI * dotNI);
        bool tir = !(k >= 0.0f);
        glm::vec3 T = tir ? glm::vec3(0.0f) : glm::refract(I, n, eta);
        if (tir || u < R) {
            ray.direction = glm::reflect(I, n);
            ray.origin = P;
            color *= m.specular.color;
        } else {
            ray.origin = getPointOnRay(ray, t + 0.0002f);   /* step through the surface (old direction) */
            ray.direction = T;
            color *= m.color;
        }
    } else if (mode == ORC_MODE_FULL && m.hasReflective > 0.0f) {
        glm::vec3 R = glm::reflect(ray.direction, n);
        if (m.specular.exponent > 0.0f) {
            R = phongLobe(glm::normalize(R), m.specular.exponent, rng);
            if (glm::dot(R, n) < 0.0f) R = glm::reflect(R, n);
        }
        ray.direction = R;
        ray.origin = P;
        color *= m.specular.color;
    } else {
        ray.direction = calculateRandomDirectionInHemisphere(n, rng);
        ray.origin = P;
        color *= m.color;
    }
}

/* one shade step; returns true while the path is alive */
static inline bool shadeStep(const Material *materials, int mode, int hit, int materialId, float t,
                             const glm::vec3 &n, bool outside, int iter, int index, int depth,
                             Ray &ray, glm::vec3 &color, int &remaining) {
    if (hit < 0) {                                    /* src/pathtrace.cu:262-264 */
        color = glm::vec3(0.0f);
        remaining = 0;
        return false;
    }
    const Material &m = materials[materialId];
    if (m.emittance > 0.0f) {                          /* src/pathtrace.cu:247-249 */
        color *= (m.color * m.emittance);
        remaining = 0;
        return false;
    }
    thrust::default_random_engine rng = seededEngine(iter, index, depth);
    scatter(ray, color, t, n, outside, m, mode, rng);
    remaining--;
    if (remaining == 0) {
        color = glm::vec3(0.0f);
        return false;
    }
    return true;
}

extern "C" {

const char *orc_kind(void) { return "reference"; }

void *orc_scene_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;       /* the reference does a bare `throw;` (src/scene.cpp:12-15) */
    fclose(f);
    /* silence the loader's cout chatter (src/scene.cpp:8-9, 41-62) */
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    Scene *s = new Scene(std::string(path));
    std::cout.rdbuf(old);
    return s;
}

void orc_scene_counts(void *scene, int *ngeoms, int *nmaterials) {
    Scene *s = (Scene *)scene;
    *ngeoms = (int)s->geoms.size();
    *nmaterials = (int)s->materials.size();
}

void orc_scene_get(void *scene, OrcGeom *geoms, OrcMaterial *materials, OrcCamera *cam,
                   int *iterations, int *traceDepth, char *imageName, int nameCap) {
    Scene *s = (Scene *)scene;
    if (geoms) memcpy(geoms, s->geoms.data(), s->geoms.size() * sizeof(Geom));
    if (materials) memcpy(materials, s->materials.data(), s->materials.size() * sizeof(Material));
    if (cam) memcpy(cam, &s->state.camera, sizeof(Camera));
    if (iterations) *iterations = (int)s->state.iterations;
    if (traceDepth) *traceDepth = s->state.traceDepth;
    if (imageName && nameCap > 0) {
        strncpy(imageName, s->state.imageName.c_str(), nameCap - 1);
        imageName[nameCap - 1] = 0;
    }
}

void orc_scene_free(void *scene) { delete (Scene *)scene; }

/* src/main.cpp:53-67 (derive zoom/phi/theta) then :102-120 (rebuild the camera) */
void orc_camera_fixup(OrcCamera *ocam) {
    Camera cam;
    memcpy(&cam, ocam, sizeof(Camera));
    glm::vec3 view = cam.view;
    glm::vec3 viewXZ = glm::vec3(view.x, 0.0f, view.z);
    glm::vec3 viewZY = glm::vec3(0.0f, view.y, view.z);
    float phi = glm::acos(glm::dot(glm::normalize(viewXZ), glm::vec3(0, 0, -1)));
    float theta = glm::acos(glm::dot(glm::normalize(viewZY), glm::vec3(0, 1, 0)));
    float zoom = glm::length(cam.position - cam.lookAt);

    glm::vec3 cameraPosition;
    cameraPosition.x = zoom * sin(phi) * sin(theta);
    cameraPosition.y = zoom * cos(theta);
    cameraPosition.z = zoom * cos(phi) * sin(theta);

    cam.view = -glm::normalize(cameraPosition);
    glm::vec3 v = cam.view;
    glm::vec3 u = glm::vec3(0, 1, 0);
    glm::vec3 r = glm::cross(v, u);
    cam.up = glm::cross(r, v);
    cam.right = r;

    cam.position = cameraPosition;
    cameraPosition += cam.lookAt;
    cam.position = cameraPosition;
    memcpy(ocam, &cam, sizeof(Camera));
}

void orc_build_transform(const float t[3], const float r[3], const float s[3],
                         float transform[16], float inverse[16], float invTranspose[16]) {
    glm::mat4 m = utilityCore::buildTransformationMatrix(v3(t), v3(r), v3(s));
    glm::mat4 inv = glm::inverse(m);
    glm::mat4 it = glm::inverseTranspose(m);
    memcpy(transform, &m[0][0], 64);
    memcpy(inverse, &inv[0][0], 64);
    memcpy(invTranspose, &it[0][0], 64);
}

void orc_rng_u01(int iter, int index, int depth, int n, float *out) {
    thrust::default_random_engine rng = seededEngine(iter, index, depth);
    thrust::uniform_real_distribution<float> u01(0, 1);
    for (int i = 0; i < n; i++) out[i] = u01(rng);
}

void orc_generate_rays(const OrcCamera *ocam, float *origins, float *dirs) {
    Camera cam;
    memcpy(&cam, ocam, sizeof(Camera));
    const int W = cam.resolution.x, H = cam.resolution.y;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            Ray r = cameraRay(cam, x, y);
            size_t i = (size_t)x + (size_t)y * W;
            put3(origins + 3 * i, r.origin);
            put3(dirs + 3 * i, r.direction);
        }
}

void orc_generate_rays_lens(const OrcCamera *ocam, const OrcLens *lens, int iter, float *origins, float *dirs) {
    Camera cam;
    memcpy(&cam, ocam, sizeof(Camera));
    const int W = cam.resolution.x, H = cam.resolution.y;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            Ray r = cameraRayLens(cam, lens, x, y, iter);
            size_t i = (size_t)x + (size_t)y * W;
            put3(origins + 3 * i, r.origin);
            put3(dirs + 3 * i, r.direction);
        }
}

void orc_closest_hit(const OrcGeom *ogeoms, int ngeoms, const float *origins, const float *dirs,
                     int n, int *hit, float *t, float *normal, int *outside) {
    orc_closest_hit_mesh(ogeoms, ngeoms, NULL, origins, dirs, n, hit, t, normal, outside);
}

void orc_closest_hit_mesh(const OrcGeom *ogeoms, int ngeoms, const OrcMeshes *meshes,
                          const float *origins, const float *dirs, int n,
                          int *hit, float *t, float *normal, int *outside) {
    const Geom *geoms = (const Geom *)ogeoms;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n; i++) {
        Ray r;
        r.origin = v3(origins + 3 * (size_t)i);
        r.direction = v3(dirs + 3 * (size_t)i);
        float tt; glm::vec3 nn; bool oo;
        hit[i] = closestHit(geoms, ngeoms, meshes, r, tt, nn, oo);
        t[i] = tt;
        put3(normal + 3 * (size_t)i, nn);
        outside[i] = oo ? 1 : 0;
    }
}

void orc_hemisphere(const float *normals, int n, const int *iter, const int *index,
                    const int *depth, float *outDirs) {
    for (int i = 0; i < n; i++) {
        thrust::default_random_engine rng = seededEngine(iter[i], index[i], depth[i]);
        glm::vec3 d = calculateRandomDirectionInHemisphere(v3(normals + 3 * (size_t)i), rng);
        put3(outDirs + 3 * (size_t)i, d);
    }
}

void orc_shade(const OrcMaterial *omats, int mode, int n,
               const int *hit, const int *materialIds, const float *t,
               const float *normal, const int *outside,
               const int *iter, const int *index, const int *depth,
               float *origins, float *dirs, float *colors, int *remaining) {
    const Material *materials = (const Material *)omats;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        Ray r;
        r.origin = v3(origins + 3 * (size_t)i);
        r.direction = v3(dirs + 3 * (size_t)i);
        glm::vec3 c = v3(colors + 3 * (size_t)i);
        int rem = remaining[i];
        shadeStep(materials, mode, hit[i], hit[i] < 0 ? 0 : materialIds[i], t[i],
                  v3(normal + 3 * (size_t)i), outside[i] != 0, iter[i], index[i], depth[i], r, c, rem);
        put3(origins + 3 * (size_t)i, r.origin);
        put3(dirs + 3 * (size_t)i, r.direction);
        put3(colors + 3 * (size_t)i, c);
        remaining[i] = rem;
    }
}

long long orc_render(const OrcGeom *ogeoms, int ngeoms, const OrcMaterial *omats, int nmaterials,
                     const OrcCamera *ocam, int traceDepth, int mode,
                     int iterBegin, int nIters, float *image, long long *alivePerDepth,
                     int nthreads) {
    return orc_render_ex(ogeoms, ngeoms, NULL, omats, nmaterials, ocam, NULL, traceDepth, mode, iterBegin, nIters,
                         image, alivePerDepth, nthreads);
}

long long orc_render_ex(const OrcGeom *ogeoms, int ngeoms, const OrcMeshes *meshes,
                        const OrcMaterial *omats, int nmaterials,
                        const OrcCamera *ocam, const OrcLens *lens, int traceDepth, int mode,
                        int iterBegin, int nIters, float *image, long long *alivePerDepth,
                        int nthreads) {
    (void)nmaterials;
    const Geom *geoms = (const Geom *)ogeoms;
    const Material *materials = (const Material *)omats;
    Camera cam;
    memcpy(&cam, ocam, sizeof(Camera));
    const int W = cam.resolution.x, H = cam.resolution.y;
    long long segments = 0;
    std::vector<long long> alive(traceDepth > 0 ? traceDepth : 1, 0);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        std::vector<long long> myAlive(alive.size(), 0);
        long long mySeg = 0;
#pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < H; y++) {
            for (int x = 0; x < W; x++) {
                const int index = x + y * W;                       /* src/pathtrace.cu:128 */
                for (int it = 0; it < nIters; it++) {
                    const int iter = iterBegin + it;
                    Ray ray = cameraRayLens(cam, lens, x, y, iter);
                    glm::vec3 color(1.0f, 1.0f, 1.0f);             /* src/pathtrace.cu:132 */
                    int remaining = traceDepth;                    /* src/pathtrace.cu:141 */
                    int depth = 0;
                    while (remaining > 0) {
                        myAlive[depth]++;
                        mySeg++;
                        float t; glm::vec3 n; bool outside;
                        int hit = closestHit(geoms, ngeoms, meshes, ray, t, n, outside);
                        int mid = hit < 0 ? 0 : geoms[hit].materialid;
                        bool go = shadeStep(materials, mode, hit, mid, t, n, outside,
                                            iter, index, depth, ray, color, remaining);
                        depth++;
                        if (!go) break;
                    }
                    /* src/pathtrace.cu:276 */
                    image[3 * (size_t)index + 0] += color.x;
                    image[3 * (size_t)index + 1] += color.y;
                    image[3 * (size_t)index + 2] += color.z;
                }
            }
        }
#pragma omp critical
        {
            for (size_t d = 0; d < alive.size(); d++) alive[d] += myAlive[d];
            segments += mySeg;
        }
    }
    if (alivePerDepth)
        for (int d = 0; d < traceDepth; d++) alivePerDepth[d] += alive[d];
    return segments;
}

} /* extern "C" */
