// ptb_bvh.h -- host-side builder of the flat BVH used for large scenes (BASELINE.json configs[4]).
// The reference has no acceleration structure (its loop is brute force, src/pathtrace.cu:176-199;
// a hierarchical structure is an assignment ask, INSTRUCTION.md:123-140).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "ptb200.h"

// sets the thread-local message ptb_last_error() returns (defined in ptb200.cu); returns `code`
extern "C" int ptb_set_last_error(int code, const char *msg);

namespace ptb {

// Environment knobs of A/B and validation runs (PTB_BVH_*, PTB_LOOP_BOUNCES, PTB_GRID_MULT ...) exist in TUNING builds only
// (make tuning -> libptb200_tuning.so, -DPTB_TUNING).  Several of them change which hits are found (PTB_BVH_TIGHT_SPHERES=1 and
// PTB_BVH_FAR_PAD=0 lose reference hits): the release library never reads the environment.
#ifdef PTB_TUNING
inline const char *tuningEnv(const char *name) { return getenv(name); }
#else
inline const char *tuningEnv(const char *) { return nullptr; }
#endif

struct HostBvhNode {             // mirrors ptb::BvhNode (64 B)
    float llo[3]; int32_t lref; float lhi[3]; int32_t lcount;
    float rlo[3]; int32_t rref; float rhi[3]; int32_t rcount;
};
static_assert(sizeof(HostBvhNode) == 64, "node layout");

struct HostBvh {
    std::vector<HostBvhNode> nodes;      // BFS order: the first 2^k-1 entries are the top k levels
    std::vector<int32_t>     order;      // order[slot] = scene geom index
    int                      maxDepth = 0;
};

// Padded world-space box of one geom (unit cube / sphere under Geom::transform).
// reach > 0: sphere boxes additionally cover the hits the reference's test reports for rays that pass outside the sphere
// (they grow with the square of the distance of the ray origin; see ptb_bvh.cpp).
void geomBounds(const PtbGeom &g, float lo[3], float hi[3], double reach = 0.0);
// Padded world-space bounding sphere of a SPHERE geom: centre, squared radius and the relative slack errK of the pruning
// discriminant (grows with the anisotropy of the ellipsoid, see ptb_bvh.cpp).
void sphereBounds(const PtbGeom &g, float centre[3], float *radius2, float *errK = nullptr, double *sigmaMax = nullptr, double *sigmaMin = nullptr);

// Padded world-space box of one mesh triangle (object-space vertices under the mesh geom's transform).
void triangleBounds(const PtbGeom &mesh, const PtbTriangle &tri, float lo[3], float hi[3]);

// The box of everything a ray can start from (geoms' boxes and the camera position, which may be NULL), and the `reach` of geomBounds it
// implies for one sphere: the distance from the sphere's centre to the farthest corner of that box.  on = false (reach 0) when
// PTB_BVH_FAR_PAD=0 opts out of the rigorous sphere boxes.
struct FarPadScene { bool on = false; double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; };
FarPadScene defaultFarPadScene(const PtbGeom *geoms, int n, const float *cameraPosition);
double farPadReach(const PtbGeom &g, const FarPadScene &scene);

// Binned-SAH top-down build; leaves hold at most `maxLeaf` geoms.  Sphere boxes: per-geom reach from `scene` when given, else `reach` for all.
void buildBvh(const PtbGeom *geoms, int n, int maxLeaf, HostBvh &out, double reach = 0.0, const FarPadScene *scene = nullptr);
// The same over caller-supplied primitive boxes (lo / hi: n x 3 floats); order[slot] = primitive index.
void buildBvhBoxes(const float *lo, const float *hi, int n, int maxLeaf, HostBvh &out);

} // namespace ptb
