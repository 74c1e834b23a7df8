// sceneStructs.h -- the POD scene/ray structs of the renderer boundary, same names, members and byte
// layout as the reference's src/sceneStructs.h (Geom 236 B, Material 44 B, Camera 84 B,
// PathSegment 44 B, ShadeableIntersection 20 B).  On the device the path state
// is NOT kept in these AoS forms (see ptb_kernels.cuh); they are the host-side interface types.
#pragma once
#include <string>
#include <vector>
#include "glm_lite.h"

#ifndef PTB_USE_REAL_GLM
namespace glm = ptbglm;
#endif

#define BACKGROUND_COLOR (glm::vec3(0.0f))

enum GeomType { SPHERE, CUBE, MESH };      // MESH: extension (INSTRUCTION.md:291-293); the reference has SPHERE, CUBE

struct Triangle { glm::vec3 v0, v1, v2; };  // object-space vertices of one mesh triangle (ptb200.h PtbTriangle)

struct Ray { glm::vec3 origin; glm::vec3 direction; };

struct Geom {
    enum GeomType type;
    int materialid;
    glm::vec3 translation;
    glm::vec3 rotation;
    glm::vec3 scale;
    glm::mat4 transform;
    glm::mat4 inverseTransform;
    glm::mat4 invTranspose;
};

struct Material {
    glm::vec3 color;
    struct { float exponent; glm::vec3 color; } specular;
    float hasReflective;
    float hasRefractive;
    float indexOfRefraction;
    float emittance;
};

struct Camera {
    glm::ivec2 resolution;
    glm::vec3 position;
    glm::vec3 lookAt;
    glm::vec3 view;
    glm::vec3 up;
    glm::vec3 right;
    glm::vec2 fov;
    glm::vec2 pixelLength;
};

struct RenderState {
    Camera camera;
    unsigned int iterations;
    int traceDepth;
    std::vector<glm::vec3> image;
    std::string imageName;
};

struct PathSegment { Ray ray; glm::vec3 color; int pixelIndex; int remainingBounces; };
struct ShadeableIntersection { float t; glm::vec3 surfaceNormal; int materialId; };

static_assert(sizeof(Geom) == 236 && sizeof(Material) == 44 && sizeof(Camera) == 84, "boundary struct layout");
static_assert(sizeof(PathSegment) == 44 && sizeof(ShadeableIntersection) == 20 && sizeof(Ray) == 24, "boundary struct layout");
