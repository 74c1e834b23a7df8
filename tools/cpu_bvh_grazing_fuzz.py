"""CPU search for sphere hits the BVH leaf boxes could lose (open item of trace_bvh_kernel).

The reference's sphere test (src/intersections.h:102-144) cancels |ro|^2 - 0.25 against (ro . rd)^2 in float, so far from a
sphere it reports hits for rays that pass slightly OUTSIDE it; the hit point then lies outside the ellipsoid and a leaf box
has to cover it.  This tool aims rays at the silhouette of random ellipsoids from far away -- the band where that happens,
found per (sphere, origin) by bisection on the oracle's own answer -- and checks every oracle hit against the boxes of
the flat BVH with the traversal kernel's visit rule (tests/test_bvh_host.py::pruned_hits).  No GPU; the oracle
(tests/orclib.py) is the checker, the builder is reached through ptb_bvh_build_host.

    python tools/cpu_bvh_grazing_fuzz.py [--spheres 200] [--scale-lo 0.3] [--scale-hi 1.2] [--dist 50 300] [--seed 1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import orclib                                   # noqa: E402
from common import pkg                          # noqa: E402
from test_bvh_host import pruned_hits, random_spheres, grazing_band, stage_b_rejects, sure_miss_rejects           # noqa: E402

F = np.float32


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=200)
    ap.add_argument("--origins", type=int, default=8, help="ray origins per sphere")
    ap.add_argument("--band", type=int, default=512, help="rays per (sphere, origin) across the silhouette edge")
    ap.add_argument("--scale-lo", type=float, default=0.3)
    ap.add_argument("--scale-hi", type=float, default=1.2)
    ap.add_argument("--dist", type=float, nargs=2, default=(50.0, 300.0))
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cubes", type=int, default=0, help="1: CUBE primitives instead of SPHERE (the 'outside the sphere' line is then meaningless)")
    ap.add_argument("--far-pad", type=int, default=0, help="1: rigorous sphere boxes (geomBounds with the diameter of origins + spheres)")
    ap.add_argument("--in-scene", type=int, default=0, help="1: ray origins next to OTHER spheres of the scene (bounce rays), far pad diameter = "
                    "the diameter of the scene's own boxes, as ptb_create computes it; --dist is ignored")
    a = ap.parse_args()
    orc = orclib.load("port")
    m = pkg()
    rs = np.random.RandomState(a.seed)
    g = random_spheres(orc, rs, a.spheres, a.scale_lo, a.scale_hi, 1 if a.cubes else 0)
    K, Q = a.spheres, a.origins
    o, d, owner_b = grazing_band(orc, g, rs, a.origins, a.band, a.dist, a.scale_hi, bool(a.in_scene))
    hit, t, _, _ = orc.closest_hit(g, o, d)
    pts = np.concatenate([o.astype(np.float64), g["translation"].astype(np.float64) + 2 * a.scale_hi, g["translation"].astype(np.float64) - 2 * a.scale_hi])
    diameter = float(np.linalg.norm(pts.max(0) - pts.min(0))) if a.far_pad else 0.0
    if a.in_scene and a.far_pad:
        root = m.bvh_build_host(g.view(m.GEOM_DT), 1)[0][0]
        diameter = float(np.linalg.norm(np.maximum(root["lhi"], root["rhi"]).astype(np.float64) - np.minimum(root["llo"], root["rlo"])))
    nodes, order, depth = m.bvh_build_host(g.view(m.GEOM_DT), 1, diameter)
    lost, checked = pruned_hits(nodes, order, o, d, hit, t)
    own = hit == owner_b
    # how far outside the ellipsoid do the reported hits of the owner lie?  object-space distance of the ray from the centre - 0.5
    inv = g["inverseTransform"].reshape(K, 4, 4).transpose(0, 2, 1).astype(np.float64)[owner_b]
    qo = np.einsum("nij,nj->ni", inv[:, :3, :3], o.astype(np.float64)) + inv[:, :3, 3]
    qd = np.einsum("nij,nj->ni", inv[:, :3, :3], d.astype(np.float64)); qd /= np.linalg.norm(qd, axis=1, keepdims=True)
    perp = np.linalg.norm(qo - (qo * qd).sum(1, keepdims=True) * qd, axis=1)
    outside = own & (perp > 0.5)
    print("spheres %d (scale %.2f..%.2f), origins/sphere %d at %.0f..%.0f units, band %d rays: %d rays, %d oracle hits, %d on the aimed sphere"
          % (K, a.scale_lo, a.scale_hi, Q, a.dist[0], a.dist[1], a.band, len(o), (hit >= 0).sum(), own.sum()))
    print("hits the reference reports for rays passing OUTSIDE the sphere: %d (max %.3e object units outside, at |ro| = %.1f)"
          % (outside.sum(), (perp[outside] - 0.5).max() if outside.any() else 0.0,
             np.linalg.norm(qo[outside][np.argmax(perp[outside])]) if outside.any() else 0.0))
    print("far pad diameter %.1f; BVH depth %d, %d (ray, box) pairs checked, hits a box on the way prunes: %d" % (diameter, depth, checked, len(lost)))
    rej = stage_b_rejects(m, g, o, d, hit, t)
    print("per-primitive pruning bound (stage B / brute-force phase 1) rejects the oracle's hit: %d" % len(rej))
    sm = sure_miss_rejects(g, o, d, hit)
    print("object-space sure-miss filter calls the oracle's hit a miss: %d" % len(sm))
    rej = np.concatenate([rej, sm])
    for r in rej[:5]:
        print("  rejected: ray %d geom %d t %.6f o %s d %s" % (r, hit[r], t[r], o[r].tolist(), d[r].tolist()))
    for r in lost[:10]:
        print("  lost: ray %d sphere %d t %.6f o %s d %s" % (r, hit[r], t[r], o[r].tolist(), d[r].tolist()))
    return 1 if lost or len(rej) else 0


if __name__ == "__main__":
    sys.exit(main())
