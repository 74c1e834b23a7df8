// ptb_bvh.cpp -- see ptb_bvh.h.  Pure host C++.
#include "ptb_bvh.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <queue>

#ifndef PTB_BVH_TIGHT_SPHERES
// Sphere primitives boxed by the ellipsoid's own world extent (0.5 * |row|_2) instead of the cube's (0.5 * |row|_1): 7.6 % fewer
// microseconds on the 100 k-object scene and WRONG -- 1 ray in 1.7e8 loses its hit (tools/gpu_bvh_pathdiff.py, iteration 14,
// pixel 1041646: sphere 31153 at t = 77.27).  The reference's sphere test cancels |ro|^2 - 0.25 against (ro . rd)^2 in object
// space, so at 150 object units from the origin its radicand is good to ~0.003 and it reports hits for rays that pass that far
// OUTSIDE the sphere; the hit point then lies outside any box that hugs the ellipsoid.  The cube's box has the slack this needs
// on every scene tested; off by default, kept as a knob for the record.
#define PTB_BVH_TIGHT_SPHERES 0
#endif

namespace ptb {

// Both reference tests return a hit point computed in float and placed 1e-4 object-space units short
// of the surface (getPointOnRay, src/intersections.h:27-29), so the point can lie slightly outside
// the exact world box of the primitive.  The pad covers that shortfall mapped through the transform
// (<= 1e-4 * sum|column|), float rounding of the object-space round trip, and the pruning slab test.
void geomBounds(const PtbGeom &g, float lo[3], float hi[3], double reach) {
    const float *m = g.transform;                       // column-major
    // Spheres, rigorous part: the reference's test accepts a ray that passes up to c u |ro|^2 object units outside the sphere
    // (see sphereBounds), |ro| <= reach / sigma_min for any ray that starts within `reach` world units of the sphere's centre, so the
    // hit point it reports can lie sigma_max (sqrt(0.25 + c u (reach / sigma_min)^2) - 0.5) outside the ellipsoid.  The box is the cube's box or the ellipsoid's own extent plus
    // that distance, whichever is larger.
    double farPad = 0.0;
    if (g.type == 0 && reach > 0.0) {
        float cc[3], r2; double smax = 1.0, smin = 1.0;
        sphereBounds(g, cc, &r2, nullptr, &smax, &smin);
        // accepted: |closest approach|^2 <= 0.25 + E object units^2, E = c u |ro|^2 (c budgeted at 24)  =>  excess sqrt(0.25 + E) - 0.5
        const double E = 16.0 * 5.9604644775390625e-08 * 1.5 * (reach / smin) * (reach / smin);
        farPad = smax * (std::sqrt(0.25 + E) - 0.5);
        // a singular transform (a SCALE of 0: glm::inverse returns non-finite entries, src/scene.cpp:82-85) has no finite bound; a box of
        // 4 x reach around the centre covers everything a ray can start from several times over
        if (!std::isfinite(farPad)) farPad = 4.0 * reach;
    }
    // knobs of A/B and validation runs (tools/gpu_bvh_pathdiff.py); read per call, the tests flip them between builds
    const bool tight = [] { const char *e = tuningEnv("PTB_BVH_TIGHT_SPHERES"); return e ? atoi(e) != 0 : PTB_BVH_TIGHT_SPHERES != 0; }();
    const double extra = [] { const char *x = tuningEnv("PTB_BVH_EXTRA_PAD"); return x ? atof(x) : 0.0; }();
    for (int r = 0; r < 3; ++r) {
        const double c = m[12 + r];
        const double rowAbs = std::fabs((double)m[0 + r]) + std::fabs((double)m[4 + r]) + std::fabs((double)m[8 + r]);
        // half extent along world axis r: the unit cube reaches 0.5 * |row|_1, the radius-0.5 sphere only 0.5 * |row|_2
        // (an ellipsoid's box is up to sqrt(3) x tighter per axis than the box of the cube around it)
        const double eTight = 0.5 * std::sqrt((double)m[0 + r] * m[0 + r] + (double)m[4 + r] * m[4 + r] + (double)m[8 + r] * m[8 + r]);
        const double e = (g.type == 0 && farPad > 0.0 && !tight) ? std::max(0.5 * rowAbs, eTight + farPad) : (tight && g.type == 0) ? 0.5 * std::sqrt((double)m[0 + r] * m[0 + r] + (double)m[4 + r] * m[4 + r] + (double)m[8 + r] * m[8 + r]) * (1.0 + 1e-6)
                                     : 0.5 * rowAbs;
        const double pad = 4e-4 * rowAbs + 1e-3 + 2e-5 * (std::fabs(c) + e) + extra;
        lo[r] = (float)(c - e - pad);
        hi[r] = (float)(c + e + pad);
    }
}

// The unit-radius-0.5 sphere under the linear part A of Geom::transform lies inside the ball of radius
// 0.5 * sigma_max(A) around the translation; sigma_max by power iteration on A^T A (3x3, double).
// Padded like geomBounds for the 1e-4 fall-short and float rounding of the reference's hit point.
void sphereBounds(const PtbGeom &g, float centre[3], float *radius2, float *errK, double *sigmaMaxOut, double *sigmaMinOut) {
    const float *m = g.transform;
    double A[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r][c] = m[4 * c + r];
    double S[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { S[i][j] = 0; for (int k = 0; k < 3; ++k) S[i][j] += A[k][i] * A[k][j]; }
    double v[3] = {0.577, 0.577, 0.577}, lambda = 0;
    for (int it = 0; it < 64; ++it) {
        double w[3] = {0, 0, 0};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) w[i] += S[i][j] * v[j];
        const double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        if (!(n > 0)) break;
        lambda = n;
        for (int i = 0; i < 3; ++i) v[i] = w[i] / n;
    }
    // power iteration converges from below; the Frobenius norm caps it from above
    const double frob = std::sqrt(S[0][0] + S[1][1] + S[2][2]);
    double sigma = std::min(frob, std::sqrt(lambda) * 1.001);
    double cmag = std::fabs((double)m[12]) + std::fabs((double)m[13]) + std::fabs((double)m[14]);
    const double R = 0.5 * sigma * (1.0 + 1e-4) + 4e-4 * sigma + 1e-3 + 2e-5 * (cmag + sigma);
    centre[0] = m[12]; centre[1] = m[13]; centre[2] = m[14];
    *radius2 = (float)(R * R * (1.0 + 1e-6));
    // Relative slack of the pruning discriminant (ptb_math.cuh, boundEntryRaw: err = errK * (|oc|^2 + b^2)).  The
    // reference's test accepts a ray whose OBJECT-space distance from the centre is up to sqrt(0.25 + c u |ro|^2)
    // (it cancels |ro|^2 - 0.25 against (ro . rd)^2 in float; c ~ 10 rounding errors, u = 2^-24, |ro| the object-space
    // distance of the ray origin).  In world space that is a slack of c u |oc|^2 (sigma_max / sigma_min)^2 on the squared
    // distance between the centre and the ray: isotropic spheres are covered by the 2e-6 that also absorbs the rounding
    // of the bound itself, thin ellipsoids need more.  sigma_min = 1 / sigma_max(A^-1), same power iteration.
    double Ai[3][3];
    const float *mi = g.inverseTransform;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ai[r][c] = mi[4 * c + r];
    double Si[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Si[i][j] = 0; for (int k = 0; k < 3; ++k) Si[i][j] += Ai[k][i] * Ai[k][j]; }
    double vi[3] = {0.577, 0.577, 0.577}, lambdaInv = 0;
    for (int it = 0; it < 64; ++it) {
        double w[3] = {0, 0, 0};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) w[i] += Si[i][j] * vi[j];
        const double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        if (!(n > 0)) break;
        lambdaInv = n;
        for (int i = 0; i < 3; ++i) vi[i] = w[i] / n;
    }
    const double frobInv = std::sqrt(Si[0][0] + Si[1][1] + Si[2][2]);
    const double sigmaInvMax = std::min(frobInv, std::sqrt(lambdaInv) * 1.001);     // = 1 / sigma_min, from above
    const double ratio = sigma * sigmaInvMax;
    if (errK) *errK = (float)std::max(2e-6, 16.0 * 5.9604644775390625e-08 * ratio * ratio * 1.5 + 1e-6);
    if (sigmaMaxOut) *sigmaMaxOut = sigma;
    if (sigmaMinOut) *sigmaMinOut = 1.0 / sigmaInvMax;
}

// A mesh triangle: the reported hit point is transform * (a point 1e-4 object units short of the triangle's
// plane along the ray), computed in float.  Pad like geomBounds: the fall-short mapped through the transform
// (<= 1e-4 * sum|row|), rounding of the object-space round trip (relative to the object- and world-space
// magnitudes involved) and the pruning slab test.
void triangleBounds(const PtbGeom &g, const PtbTriangle &t, float lo[3], float hi[3]) {
    const float *m = g.transform;                       // column-major
    const float *v[3] = {t.v0, t.v1, t.v2};
    double vmax = 0;
    for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) vmax = std::max(vmax, std::fabs((double)v[k][a]));
    for (int r = 0; r < 3; ++r) {
        double wlo = DBL_MAX, whi = -DBL_MAX;
        for (int k = 0; k < 3; ++k) {
            const double w = (double)m[0 + r] * v[k][0] + (double)m[4 + r] * v[k][1] + (double)m[8 + r] * v[k][2] + (double)m[12 + r];
            wlo = std::min(wlo, w); whi = std::max(whi, w);
        }
        const double rowAbs = std::fabs((double)m[0 + r]) + std::fabs((double)m[4 + r]) + std::fabs((double)m[8 + r]);
        const double pad = 4e-4 * rowAbs + 1e-3 + 2e-5 * std::max(std::fabs(wlo), std::fabs(whi)) + 2e-6 * rowAbs * vmax;
        lo[r] = (float)(wlo - pad);
        hi[r] = (float)(whi + pad);
    }
}

#ifndef PTB_BVH_BINS
#define PTB_BVH_BINS 16
#endif
#ifndef PTB_BVH_ALL_AXES
#define PTB_BVH_ALL_AXES 1
#endif

namespace {
struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
    void grow(const Box &b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};
struct Prim { Box box; float cen[3]; int32_t index; };
struct BuildNode { Box box; int first, count; int left = -1, right = -1; };   // leaf when left < 0

struct Builder {
    std::vector<Prim> prims;
    std::vector<BuildNode> nodes;
    int maxLeaf;
    int maxDepth = 0;

    int build(int first, int count, int depth) {
        BuildNode nd;
        nd.first = first; nd.count = count;
        nd.box.reset();
        Box cb; cb.reset();
        for (int i = first; i < first + count; ++i) {
            nd.box.grow(prims[i].box);
            for (int a = 0; a < 3; ++a) { cb.lo[a] = std::min(cb.lo[a], prims[i].cen[a]); cb.hi[a] = std::max(cb.hi[a], prims[i].cen[a]); }
        }
        const int me = (int)nodes.size();
        nodes.push_back(nd);
        maxDepth = std::max(maxDepth, depth);
        if (count <= maxLeaf) return me;                                // leaves hold at most maxLeaf (<= 8) primitives: the 4-bit count of the leaf encoding
        // A degenerate chain (SAH peeling one primitive per level off a pathological centroid distribution) is finished with
        // median splits: this subtree then adds at most log2(count) <= 27 levels, so the tree is never deeper than ~70 levels
        // and the traversal stack (sized from maxDepth, ptb200.cu) stays small.
        const bool medianOnly = depth > 40;
        // binned SAH: the cheapest of the bin boundaries of all three centroid axes (falling back to a median split on
        // the widest axis).  Searching the widest axis only: 4 451 instead of 3 999 us per iteration of the 100 k-object scene.
        int axis = 0;
        float ext = cb.hi[0] - cb.lo[0];
        for (int a = 1; a < 3; ++a) if (cb.hi[a] - cb.lo[a] > ext) { ext = cb.hi[a] - cb.lo[a]; axis = a; }
        int mid = first + count / 2;
        constexpr int NB = PTB_BVH_BINS;
        const bool allAxes = [] { const char *e = tuningEnv("PTB_BVH_ALL_AXES"); return e ? atoi(e) != 0 : PTB_BVH_ALL_AXES != 0; }();   // knob for A/B runs
        float bestCost = FLT_MAX; int bestB = -1, bestAxis = -1; float bestK = 0.f;
        for (int ax = 0; ax < 3 && !medianOnly; ++ax) {
            const float e = cb.hi[ax] - cb.lo[ax];
            if (!(e > 0.f)) continue;
            if (!allAxes && ax != axis) continue;
            Box bb[NB]; int bc[NB];
            for (int b = 0; b < NB; ++b) { bb[b].reset(); bc[b] = 0; }
            const float k = NB * (1.f - 1e-6f) / e;
            for (int i = first; i < first + count; ++i) {
                int b = (int)((prims[i].cen[ax] - cb.lo[ax]) * k);
                b = std::min(std::max(b, 0), NB - 1);
                bb[b].grow(prims[i].box); bc[b]++;
            }
            float leftA[NB], rightA[NB]; int leftC[NB], rightC[NB];
            Box acc; acc.reset(); int c = 0;
            for (int b = 0; b < NB; ++b) { acc.grow(bb[b]); c += bc[b]; leftA[b] = acc.area(); leftC[b] = c; }
            acc.reset(); c = 0;
            for (int b = NB - 1; b >= 0; --b) { acc.grow(bb[b]); c += bc[b]; rightA[b] = acc.area(); rightC[b] = c; }
            for (int b = 0; b + 1 < NB; ++b) {
                if (leftC[b] == 0 || rightC[b + 1] == 0) continue;
                const float cost = leftA[b] * leftC[b] + rightA[b + 1] * rightC[b + 1];
                if (cost < bestCost) { bestCost = cost; bestB = b; bestAxis = ax; bestK = k; }
            }
        }
        if (bestB >= 0) {
            const int ax = bestAxis; const float k = bestK;
            auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const Prim &p) {
                int b = (int)((p.cen[ax] - cb.lo[ax]) * k);
                b = std::min(std::max(b, 0), NB - 1);
                return b <= bestB;
            });
            mid = (int)(it - prims.begin());
        }
        if (mid <= first || mid >= first + count) {
            mid = first + count / 2;
            std::nth_element(prims.begin() + first, prims.begin() + mid, prims.begin() + first + count,
                             [&](const Prim &a, const Prim &b) { return a.cen[axis] < b.cen[axis]; });
        }
        const int l = build(first, mid - first, depth + 1);
        const int r = build(mid, first + count - mid, depth + 1);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
};
} // namespace

// The box of everything a ray can start from -- the camera (when given) and every geom's world box -- from which the rigorous sphere
// boxes of geomBounds take their `reach`: the distance from the sphere's centre to the farthest corner of that box.  Off when
// PTB_BVH_FAR_PAD=0 asks for the round-1 boxes: those are 19 % faster on the 100 k-object scene (4 006 against 4 781 us per iteration with a
// uniform reach = the scene diameter, profiles/r03_bvh_pathdiff.txt) and lose reference hits of rays that graze a thin ellipsoid from more
// than ~100 world units away (5-21 of 1.2 M silhouette rays, tools/cpu_bvh_grazing_fuzz.py --dist 160 230; none of the 2.3e8 real paths
// of that scene).
FarPadScene defaultFarPadScene(const PtbGeom *geoms, int n, const float *cameraPosition) {
    FarPadScene s;
    if (const char *e = tuningEnv("PTB_BVH_FAR_PAD")) if (atoi(e) == 0) return s;
    for (int a = 0; a < 3; ++a) { s.lo[a] = DBL_MAX; s.hi[a] = -DBL_MAX; }
    if (cameraPosition) for (int a = 0; a < 3; ++a) s.lo[a] = s.hi[a] = cameraPosition[a];
    for (int i = 0; i < n; ++i) {
        float l[3], h[3];
        geomBounds(geoms[i], l, h);
        for (int a = 0; a < 3; ++a) { s.lo[a] = std::min(s.lo[a], (double)l[a]); s.hi[a] = std::max(s.hi[a], (double)h[a]); }
    }
    if (!(s.hi[0] >= s.lo[0])) return s;                  // nothing to bound
    s.on = true;
    // A hit the far pad exists for lies OUTSIDE the round-1 box of its sphere, and the next ray of that path starts there: grow the box
    // by the largest pad it implies (one fixed-point step; what is left is second order in pad / diameter, inside the 24u budget).
    double maxPad = 0.0;
    for (int i = 0; i < n; ++i) if (geoms[i].type == 0) {
        float l0[3], h0[3], l1[3], h1[3];
        geomBounds(geoms[i], l0, h0);
        geomBounds(geoms[i], l1, h1, farPadReach(geoms[i], s));
        for (int a = 0; a < 3; ++a) maxPad = std::max(maxPad, std::max((double)l0[a] - l1[a], (double)h1[a] - h0[a]));
    }
    for (int a = 0; a < 3; ++a) { s.lo[a] -= maxPad; s.hi[a] += maxPad; }
    return s;
}

double farPadReach(const PtbGeom &g, const FarPadScene &s) {
    if (!s.on) return 0.0;
    double d2 = 0.0;
    for (int a = 0; a < 3; ++a) {
        const double c = g.transform[12 + a];
        const double d = std::max(std::fabs(c - s.lo[a]), std::fabs(s.hi[a] - c));
        d2 += d * d;
    }
    return std::sqrt(d2);
}

void buildBvh(const PtbGeom *geoms, int n, int maxLeaf, HostBvh &out, double reach, const FarPadScene *scene) {
    std::vector<float> lo((size_t)n * 3), hi((size_t)n * 3);
    for (int i = 0; i < n; ++i) geomBounds(geoms[i], &lo[3 * (size_t)i], &hi[3 * (size_t)i], scene ? farPadReach(geoms[i], *scene) : reach);
    buildBvhBoxes(lo.data(), hi.data(), n, maxLeaf, out);
}

void buildBvhBoxes(const float *lo, const float *hi, int n, int maxLeaf, HostBvh &out) {
    Builder b;
    b.maxLeaf = std::max(1, std::min(maxLeaf, 8));
    b.prims.resize(n);
    for (int i = 0; i < n; ++i) {
        for (int a = 0; a < 3; ++a) {
            b.prims[i].box.lo[a] = lo[3 * (size_t)i + a]; b.prims[i].box.hi[a] = hi[3 * (size_t)i + a];
            b.prims[i].cen[a] = 0.5f * (b.prims[i].box.lo[a] + b.prims[i].box.hi[a]);
        }
        b.prims[i].index = i;
    }
    out.nodes.clear(); out.order.clear();
    if (n == 0) {
        HostBvhNode e{};
        e.lref = -1; e.rref = -1;
        out.nodes.push_back(e);
        return;
    }
    const int root = b.build(0, n, 0);
    out.maxDepth = b.maxDepth;
    out.order.resize(n);
    // scene order inside every leaf keeps ties deterministic without relying on the tie rule alone
    for (const BuildNode &nd : b.nodes)
        if (nd.left < 0)
            std::sort(b.prims.begin() + nd.first, b.prims.begin() + nd.first + nd.count,
                      [](const Prim &x, const Prim &y) { return x.index < y.index; });
    for (int i = 0; i < n; ++i) out.order[i] = b.prims[i].index;

    auto fill = [&](int child, float lo[3], int32_t &ref, float hi[3], int32_t &count, int interiorIndex) {
        const BuildNode &c = b.nodes[child];
        for (int a = 0; a < 3; ++a) { lo[a] = c.box.lo[a]; hi[a] = c.box.hi[a]; }
        if (c.left < 0) { ref = c.first; count = c.count; }
        else { ref = interiorIndex; count = 0; }
    };
    if (b.nodes[root].left < 0) {                         // the whole scene is one leaf
        HostBvhNode e{};
        fill(root, e.llo, e.lref, e.lhi, e.lcount, -1);
        e.rref = -1; e.rcount = 0;
        for (int a = 0; a < 3; ++a) { e.rlo[a] = FLT_MAX; e.rhi[a] = -FLT_MAX; }
        out.nodes.push_back(e);
        return;
    }
    // breadth-first emission of interior nodes: index = order of discovery
    std::vector<int> bfs;                                // build-node ids of interior nodes in BFS order
    std::vector<int> outIndex(b.nodes.size(), -1);
    bfs.push_back(root); outIndex[root] = 0;
    for (size_t h = 0; h < bfs.size(); ++h) {
        const BuildNode &nd = b.nodes[bfs[h]];
        for (int child : {nd.left, nd.right})
            if (b.nodes[child].left >= 0) { outIndex[child] = (int)bfs.size(); bfs.push_back(child); }
    }
    out.nodes.resize(bfs.size());
    for (size_t h = 0; h < bfs.size(); ++h) {
        const BuildNode &nd = b.nodes[bfs[h]];
        HostBvhNode &e = out.nodes[h];
        fill(nd.left, e.llo, e.lref, e.lhi, e.lcount, outIndex[nd.left]);
        fill(nd.right, e.rlo, e.rref, e.rhi, e.rcount, outIndex[nd.right]);
    }
}

} // namespace ptb

// Host-only entry point of the C ABI (include/ptb200.h): the flat BVH ptb_create builds over the SPHERE / CUBE geoms of a scene
// (same bounds, same builder, same PTB_BVH_* knobs), returned as HostBvhNode records.  No CUDA call.
extern "C" int ptb_bvh_build_host(const PtbGeom *geoms, int32_t numGeoms, int32_t maxLeaf, double farPadDiameter, float *nodes, int32_t nodeCap,
                                  int32_t *order, int32_t *numNodes, int32_t *maxDepth) {
    if (numGeoms < 0 || (numGeoms > 0 && !geoms) || !numNodes) return ptb_set_last_error(PTB_ERR_INVALID, "ptb_bvh_build_host: bad argument");
    for (int i = 0; i < numGeoms; ++i) if (geoms[i].type != 0 && geoms[i].type != 1) return ptb_set_last_error(PTB_ERR_INVALID, "ptb_bvh_build_host: only SPHERE / CUBE geoms");
    ptb::HostBvh hb;
    if (farPadDiameter < 0.0) {
        const ptb::FarPadScene scene = ptb::defaultFarPadScene(geoms, numGeoms, nullptr);
        ptb::buildBvh(geoms, numGeoms, maxLeaf, hb, 0.0, &scene);
    } else ptb::buildBvh(geoms, numGeoms, maxLeaf, hb, farPadDiameter);
    *numNodes = (int32_t)hb.nodes.size();
    if (maxDepth) *maxDepth = hb.maxDepth;
    if (nodes) {
        if (nodeCap < (int32_t)hb.nodes.size()) return ptb_set_last_error(PTB_ERR_INVALID, "ptb_bvh_build_host: nodeCap smaller than the node count");
        std::copy(hb.nodes.begin(), hb.nodes.end(), reinterpret_cast<ptb::HostBvhNode *>(nodes));
    }
    if (order) std::copy(hb.order.begin(), hb.order.end(), order);
    return PTB_OK;
}

// Host-only: the pruning bounds of one SPHERE / CUBE geom as ptb_create packs them for the kernels (ptb200.cu packGeom: the padded
// world box of geomBounds with the given reach, and for spheres the bounding sphere and discriminant slack of sphereBounds).
extern "C" int ptb_geom_bounds_host(const PtbGeom *geom, double reach, float lo[3], float hi[3], float centre[3], float *radius2, float *errK) {
    if (!geom || !lo || !hi || (geom->type != 0 && geom->type != 1)) return ptb_set_last_error(PTB_ERR_INVALID, "ptb_geom_bounds_host: bad argument (SPHERE / CUBE geoms only)");
    ptb::geomBounds(*geom, lo, hi, reach > 0.0 ? reach : 0.0);
    if (geom->type == 0 && centre && radius2) ptb::sphereBounds(*geom, centre, radius2, errK);
    return PTB_OK;
}
