// scene.cpp -- parses the reference's scene text format (INSTRUCTION.md:258-301) with the same
// acceptance rules as src/scene.cpp: MATERIAL blocks of exactly seven property lines, a CAMERA block
// of five fixed lines followed by EYE/LOOKAT/UP until a blank line, OBJECT blocks of a type line
// (compared as a whole line), a material line and TRANS/ROTAT/SCALE until a blank line; ids must be
// sequential; numbers go through atof -> float.
// Extension: the type line `mesh <file.obj>` (the format reserves `mesh`, INSTRUCTION.md:291-293; the
// reference's loader leaves the type uninitialised for it, src/scene.cpp:46-55) loads a Wavefront OBJ
// (v / f records, polygons fanned into triangles) relative to the scene file's directory.
#include "scene.h"
#include <cstdlib>
#include <cstring>
#include <iostream>

using utilityCore::safeGetline;
using utilityCore::tokenizeString;

static glm::vec3 vec3From(const std::vector<std::string> &tok) {
    auto f = [&](size_t i) { return i < tok.size() ? atof(tok[i].c_str()) : 0.0; };
    return glm::vec3(f(1), f(2), f(3));
}

Scene::Scene(std::string filename, bool verbose_) : verbose(verbose_) {
    memset(&state.camera, 0, sizeof state.camera);
    state.iterations = 0; state.traceDepth = 0;
    if (verbose) std::cout << "Reading scene from " << filename << " ..." << std::endl;
    const size_t slash = filename.find_last_of("/\\");
    baseDir = slash == std::string::npos ? std::string() : filename.substr(0, slash + 1);
    fp_in.open(filename.c_str());
    if (!fp_in.is_open()) { if (verbose) std::cout << "Error reading from file - aborting!" << std::endl; return; }
    ok = true;
    while (fp_in.good()) {                                        // src/scene.cpp:16-32
        std::string line;
        safeGetline(fp_in, line);
        if (line.empty()) continue;
        std::vector<std::string> tokens = tokenizeString(line);
        if (tokens.empty()) continue;
        if (tokens[0] == "MATERIAL" && tokens.size() > 1) loadMaterial(tokens[1]);
        else if (tokens[0] == "OBJECT" && tokens.size() > 1) loadGeom(tokens[1]);
        else if (tokens[0] == "CAMERA") loadCamera();
    }
}

int Scene::loadGeom(std::string objectid) {                        // src/scene.cpp:35-90
    const int id = atoi(objectid.c_str());
    if (id != (int)geoms.size()) {
        if (verbose) std::cout << "ERROR: OBJECT ID does not match expected number of geoms" << std::endl;
        return -1;
    }
    Geom g;
    g.type = SPHERE; g.materialid = 0;
    std::string line;
    safeGetline(fp_in, line);                                      // type line, matched whole (:46-55)
    int firstTri = (int)triangles.size(), numTri = 0;
    if (!line.empty() && fp_in.good()) {
        if (line == "sphere") g.type = SPHERE;
        else if (line == "cube") g.type = CUBE;
        else {
            std::vector<std::string> tokens = tokenizeString(line);
            if (tokens.size() > 1 && tokens[0] == "mesh") {
                g.type = MESH;
                const std::string path = (tokens[1][0] == '/') ? tokens[1] : baseDir + tokens[1];
                numTri = loadObj(path, triangles);
                if (numTri < 0) { ok = false; numTri = 0; if (verbose) std::cout << "Error reading mesh " << path << std::endl; }
            }
        }
    }
    meshRanges.push_back(firstTri);
    meshRanges.push_back(numTri);
    safeGetline(fp_in, line);                                      // "material N" (:58-63)
    if (!line.empty() && fp_in.good()) {
        std::vector<std::string> tokens = tokenizeString(line);
        if (tokens.size() > 1) g.materialid = atoi(tokens[1].c_str());
    }
    safeGetline(fp_in, line);                                      // transformations until a blank line (:66-80)
    while (!line.empty() && fp_in.good()) {
        std::vector<std::string> tokens = tokenizeString(line);
        if (!tokens.empty()) {
            if (tokens[0] == "TRANS") g.translation = vec3From(tokens);
            else if (tokens[0] == "ROTAT") g.rotation = vec3From(tokens);
            else if (tokens[0] == "SCALE") g.scale = vec3From(tokens);
        }
        safeGetline(fp_in, line);
    }
    g.transform = utilityCore::buildTransformationMatrix(g.translation, g.rotation, g.scale);   // :82-85
    g.inverseTransform = glm::inverse(g.transform);
    g.invTranspose = glm::inverseTranspose(g.transform);
    geoms.push_back(g);
    return 1;
}

// Wavefront OBJ: `v x y z` and `f a b c ...` records (a = index, index/vt, index/vt/vn or index//vn;
// 1-based, negative = relative to the vertices read so far); polygons become a fan around their first
// vertex; everything else is ignored.  Returns the number of triangles appended, -1 if the file cannot be opened.
int Scene::loadObj(const std::string &path, std::vector<Triangle> &out) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) return -1;
    std::vector<glm::vec3> verts;
    int added = 0;
    while (f.good()) {
        std::string line;
        safeGetline(f, line);
        std::vector<std::string> tok = tokenizeString(line);
        if (tok.empty()) continue;
        if (tok[0] == "v" && tok.size() > 3) verts.push_back(vec3From(tok));
        else if (tok[0] == "f" && tok.size() > 3) {
            std::vector<int> idx;
            for (size_t i = 1; i < tok.size(); ++i) {
                int k = atoi(tok[i].c_str());                   // atoi stops at the first '/'
                if (k < 0) k = (int)verts.size() + k; else k -= 1;
                if (k < 0 || k >= (int)verts.size()) { idx.clear(); break; }
                idx.push_back(k);
            }
            for (size_t i = 2; i < idx.size(); ++i) {
                Triangle t;
                t.v0 = verts[idx[0]]; t.v1 = verts[idx[i - 1]]; t.v2 = verts[idx[i]];
                out.push_back(t);
                ++added;
            }
        }
    }
    return added;
}

int Scene::loadCamera() {                                          // src/scene.cpp:92-151
    Camera &camera = state.camera;
    float fovy = 0.0f;
    for (int i = 0; i < 5; i++) {                                  // five fixed lines (:99)
        std::string line;
        safeGetline(fp_in, line);
        std::vector<std::string> tokens = tokenizeString(line);
        if (tokens.empty()) continue;
        if (tokens[0] == "RES" && tokens.size() > 2) { camera.resolution.x = atoi(tokens[1].c_str()); camera.resolution.y = atoi(tokens[2].c_str()); }
        else if (tokens[0] == "FOVY" && tokens.size() > 1) fovy = (float)atof(tokens[1].c_str());
        else if (tokens[0] == "ITERATIONS" && tokens.size() > 1) state.iterations = atoi(tokens[1].c_str());
        else if (tokens[0] == "DEPTH" && tokens.size() > 1) state.traceDepth = atoi(tokens[1].c_str());
        else if (tokens[0] == "FILE" && tokens.size() > 1) state.imageName = tokens[1];
    }
    std::string line;
    safeGetline(fp_in, line);
    while (!line.empty() && fp_in.good()) {                        // EYE / LOOKAT / UP until a blank line (:117-130)
        std::vector<std::string> tokens = tokenizeString(line);
        if (!tokens.empty()) {
            if (tokens[0] == "EYE") camera.position = vec3From(tokens);
            else if (tokens[0] == "LOOKAT") camera.lookAt = vec3From(tokens);
            else if (tokens[0] == "UP") camera.up = vec3From(tokens);
        }
        safeGetline(fp_in, line);
    }
    // FOVY is a half angle (:133, INSTRUCTION.md:280)
    float yscaled = std::tan(fovy * (PI / 180));
    float xscaled = (yscaled * camera.resolution.x) / camera.resolution.y;
    float fovx = (std::atan(xscaled) * 180) / PI;
    camera.fov = glm::vec2(fovx, fovy);
    // `right` is derived from `view` before `view` is assigned (:138 vs :142): it is NaN here and only
    // becomes valid in the camera fix-up (ptb_camera_fixup / src/main.cpp:102-120).  Kept as is.
    camera.right = glm::normalize(glm::cross(camera.view, camera.up));
    camera.pixelLength = glm::vec2(2 * xscaled / (float)camera.resolution.x, 2 * yscaled / (float)camera.resolution.y);
    camera.view = glm::normalize(camera.lookAt - camera.position);
    state.image.assign((size_t)camera.resolution.x * camera.resolution.y, glm::vec3());
    return 1;
}

int Scene::loadMaterial(std::string materialid) {                  // src/scene.cpp:153-188
    const int id = atoi(materialid.c_str());
    if (id != (int)materials.size()) {
        if (verbose) std::cout << "ERROR: MATERIAL ID does not match expected number of materials" << std::endl;
        return -1;
    }
    Material m;
    m.specular.exponent = 0; m.hasReflective = 0; m.hasRefractive = 0; m.indexOfRefraction = 0; m.emittance = 0;
    for (int i = 0; i < 7; i++) {                                  // exactly seven property lines (:163)
        std::string line;
        safeGetline(fp_in, line);
        std::vector<std::string> tokens = tokenizeString(line);
        if (tokens.empty()) continue;
        auto f1 = [&]() { return tokens.size() > 1 ? (float)atof(tokens[1].c_str()) : 0.0f; };
        if (tokens[0] == "RGB") m.color = vec3From(tokens);
        else if (tokens[0] == "SPECEX") m.specular.exponent = f1();
        else if (tokens[0] == "SPECRGB") m.specular.color = vec3From(tokens);
        else if (tokens[0] == "REFL") m.hasReflective = f1();
        else if (tokens[0] == "REFR") m.hasRefractive = f1();
        else if (tokens[0] == "REFRIOR") m.indexOfRefraction = f1();
        else if (tokens[0] == "EMITTANCE") m.emittance = f1();
    }
    materials.push_back(m);
    return 1;
}
