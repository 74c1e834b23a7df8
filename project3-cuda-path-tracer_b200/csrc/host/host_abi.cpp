// host_abi.cpp -- C-ABI access to the host-side pieces (scene loader, camera fix-up, image writer)
// so that non-C++ callers (the Python tests, other FFI users) reach the same code the headless
// driver uses.  Declared in include/ptb200.h.
#include "ptb200.h"
#include "scene.h"
#include "image.h"
#include <cstring>
#include <string>

extern "C" int ptb_set_last_error(int code, const char *msg);     // ptb200.cu

extern "C" {

// new Scene(filename) (src/scene.cpp:7-33)
int ptb_scene_load(const char *path, void **sceneOut) {
    if (!path || !sceneOut) return ptb_set_last_error(PTB_ERR_INVALID, "ptb_scene_load: null argument");
    Scene *s = new Scene(std::string(path), false);
    if (!s->ok) { delete s; *sceneOut = nullptr; return ptb_set_last_error(PTB_ERR_IO, (std::string("ptb_scene_load: cannot read ") + path).c_str()); }
    *sceneOut = s;
    return PTB_OK;
}
int ptb_scene_counts(void *scene, int32_t *numGeoms, int32_t *numMaterials) {
    if (!scene) return ptb_set_last_error(PTB_ERR_INVALID, "null scene handle");
    Scene *s = static_cast<Scene *>(scene);
    if (numGeoms) *numGeoms = (int32_t)s->geoms.size();
    if (numMaterials) *numMaterials = (int32_t)s->materials.size();
    return PTB_OK;
}
int ptb_scene_get(void *scene, PtbGeom *geoms, PtbMaterial *materials, PtbCamera *camera,
                  int32_t *iterations, int32_t *traceDepth, char *imageName, int nameCap) {
    if (!scene) return ptb_set_last_error(PTB_ERR_INVALID, "null scene handle");
    Scene *s = static_cast<Scene *>(scene);
    if (geoms && !s->geoms.empty()) memcpy(geoms, s->geoms.data(), s->geoms.size() * sizeof(Geom));
    if (materials && !s->materials.empty()) memcpy(materials, s->materials.data(), s->materials.size() * sizeof(Material));
    if (camera) memcpy(camera, &s->state.camera, sizeof(Camera));
    if (iterations) *iterations = (int32_t)s->state.iterations;
    if (traceDepth) *traceDepth = s->state.traceDepth;
    if (imageName && nameCap > 0) { strncpy(imageName, s->state.imageName.c_str(), (size_t)nameCap - 1); imageName[nameCap - 1] = 0; }
    return PTB_OK;
}
int ptb_scene_mesh_counts(void *scene, int32_t *numTriangles) {
    if (!scene) return ptb_set_last_error(PTB_ERR_INVALID, "null scene handle");
    if (numTriangles) *numTriangles = (int32_t)static_cast<Scene *>(scene)->triangles.size();
    return PTB_OK;
}
int ptb_scene_get_meshes(void *scene, PtbTriangle *triangles, int32_t *meshRanges) {
    if (!scene) return ptb_set_last_error(PTB_ERR_INVALID, "null scene handle");
    Scene *s = static_cast<Scene *>(scene);
    static_assert(sizeof(Triangle) == sizeof(PtbTriangle), "triangle layout");
    if (triangles && !s->triangles.empty()) memcpy(triangles, s->triangles.data(), s->triangles.size() * sizeof(Triangle));
    if (meshRanges && !s->meshRanges.empty()) memcpy(meshRanges, s->meshRanges.data(), s->meshRanges.size() * sizeof(int));
    return PTB_OK;
}
void ptb_scene_free(void *scene) { delete static_cast<Scene *>(scene); }

// The camera rebuild the app performs before the first frame: derive (zoom, phi, theta) from the
// loaded camera (src/main.cpp:53-67) and rebuild position/view/up/right from them (:102-120).
// Reproduced as is, including its quirks: `up` from the file is ignored,
// `right` is not normalised, and it is an identity only for horizontal views along -z.
void ptb_camera_fixup(PtbCamera *pc) {
    Camera cam;
    memcpy(&cam, pc, sizeof cam);
    const glm::vec3 view = cam.view;
    const glm::vec3 viewXZ(view.x, 0.0f, view.z), viewZY(0.0f, view.y, view.z);
    const float phi = std::acos(glm::dot(glm::normalize(viewXZ), glm::vec3(0, 0, -1)));
    const float theta = std::acos(glm::dot(glm::normalize(viewZY), glm::vec3(0, 1, 0)));
    const float zoom = glm::length(cam.position - cam.lookAt);
    glm::vec3 cameraPosition;
    cameraPosition.x = zoom * std::sin(phi) * std::sin(theta);
    cameraPosition.y = zoom * std::cos(theta);
    cameraPosition.z = zoom * std::cos(phi) * std::sin(theta);
    cam.view = -glm::normalize(cameraPosition);
    const glm::vec3 v = cam.view, u(0, 1, 0);
    const glm::vec3 r = glm::cross(v, u);
    cam.up = glm::cross(r, v);
    cam.right = r;
    cameraPosition += cam.lookAt;
    cam.position = cameraPosition;
    memcpy(pc, &cam, sizeof cam);
}

// saveImage (src/main.cpp:78-99): pixel (x, y) of the accumulated sum goes to (W-1-x, y) divided by
// the sample count, then image::savePNG.  `baseFilename` gets ".png" appended like the reference.
static int saveImpl(const char *baseFilename, int w, int h, const float *sum, float samples, int flipX, bool hdr) {
    if (!baseFilename || !sum || w <= 0 || h <= 0 || !(samples > 0)) return ptb_set_last_error(PTB_ERR_INVALID, "ptb_save_*: bad argument");
    image img(w, h);
    for (int x = 0; x < w; x++)
        for (int y = 0; y < h; y++) {
            const size_t index = (size_t)x + (size_t)y * w;
            const glm::vec3 pix(sum[3 * index], sum[3 * index + 1], sum[3 * index + 2]);
            img.setPixel(flipX ? w - 1 - x : x, y, pix / samples);
        }
    const bool ok = hdr ? img.saveHDR(baseFilename, true) : img.savePNG(baseFilename, true);
    return ok ? PTB_OK : ptb_set_last_error(PTB_ERR_IO, (std::string("cannot write ") + baseFilename).c_str());
}
int ptb_save_png(const char *baseFilename, int w, int h, const float *sum, float samples, int flipX) {
    return saveImpl(baseFilename, w, h, sum, samples, flipX, false);
}
int ptb_save_hdr(const char *baseFilename, int w, int h, const float *sum, float samples, int flipX) {
    return saveImpl(baseFilename, w, h, sum, samples, flipX, true);
}

} // extern "C"
