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
public:
    explicit Scene(std::string filename, bool verbose = false);
    ~Scene() {}

    std::vector<Geom> geoms;
    std::vector<Material> materials;
    RenderState state;
    bool ok = false;          // false when the file could not be opened (the reference aborts, src/scene.cpp:12-15)
    bool verbose = false;     // the reference prints 4+ lines per object (src/scene.cpp:41-62); quiet by default here
};
