// scene.h -- scene-file loader with the reference's interface (src/scene.h): `Scene(filename)` and
// the public members geoms / materials / state.
#pragma once
#include <fstream>
#include <string>
#include <vector>
#include "sceneStructs.h"
#include "utilities.h"

class Scene {
private:
    std::ifstream fp_in;
    int loadMaterial(std::string materialid);
    int loadGeom(std::string objectid);
    int loadCamera();
    int loadObj(const std::string &path, std::vector<Triangle> &out);
    std::string baseDir;      // directory of the scene file: `mesh <file.obj>` paths are relative to it
public:
    explicit Scene(std::string filename, bool verbose = false);
    ~Scene() {}

    std::vector<Geom> geoms;
    std::vector<Material> materials;
    // `mesh <file.obj>` objects (extension): all triangles, and (first, count) per geom
#define PTB_SCENE_HAS_MESHES 1
    std::vector<Triangle> triangles;
    std::vector<int> meshRanges;
    RenderState state;
    bool ok = false;          // false when the file could not be opened (the reference aborts, src/scene.cpp:12-15)
    bool verbose = false;     // the reference prints 4+ lines per object (src/scene.cpp:41-62); quiet by default here
};
