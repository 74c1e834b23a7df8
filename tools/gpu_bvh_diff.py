"""Closest hits of the synthetic 100k-object scene through two differently built BVHs (builder knobs PTB_BVH_ALL_AXES /
PTB_BVH_TIGHT_SPHERES, read when a context is created) on the same rays -- camera rays of the full frame plus secondary
rays leaving the first hits in random directions -- and, for every ray on which they disagree (and a random sample of the
rest), the CPU oracle's brute-force answer over all the geoms.   python tools/gpu_bvh_diff.py [n_secondary_rounds]"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import orclib, scenes
m = g.load_package(); orc = orclib.load("port")
d = tempfile.mkdtemp(); path = os.path.join(d, "synthetic.txt")
open(path, "w").write(scenes.synthetic_text(100000))
scene = m.Scene.load(path); osc = orc.load_scene(path)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def make(env):
    for k, v in env.items(): os.environ[k] = v
    return m.Renderer(scene, device=0, bsdf=1, batch_iters=1)


OLD = {"PTB_BVH_ALL_AXES": "0", "PTB_BVH_TIGHT_SPHERES": "0"}
NEW = {"PTB_BVH_ALL_AXES": "1", "PTB_BVH_TIGHT_SPHERES": "1"}
r_old, r_new = make(OLD), make(NEW)
hits = lambda r, o, dd: r.stage_intersect(o, dd, m.ACCEL_BVH)
o, dd = r_new.stage_raygen()
rs = np.random.RandomState(7)
total = bad = 0
for rnd in range(rounds + 1):
    a = hits(r_old, o, dd)
    b = hits(r_new, o, dd)
    same = (a[0] == b[0]) & (a[1].view(np.uint32) == b[1].view(np.uint32)) & (a[2].view(np.uint32) == b[2].view(np.uint32)).all(1) & (a[3] == b[3])
    diff = np.flatnonzero(~same)
    sample = rs.choice(len(o), 500, replace=False)
    check = np.unique(np.concatenate([diff[:5000], sample]))
    want = orc.closest_hit(osc.geoms, o[check], dd[check])
    okA = (a[0][check] == want[0]) & (a[1][check].view(np.uint32) == want[1].view(np.uint32))
    okB = (b[0][check] == want[0]) & (b[1][check].view(np.uint32) == want[1].view(np.uint32))
    print("round %d: %d rays, %d hits; the two BVHs disagree on %d; against the oracle on %d rays (all disagreements + 500 random): "
          "old builder wrong on %d, new builder wrong on %d" % (rnd, len(o), int((b[0] >= 0).sum()), len(diff), len(check), int((~okA).sum()), int((~okB).sum())), flush=True)
    for i in check[~okB][:5]:
        j = int(np.flatnonzero(check == i)[0])
        print("   ray %d o=%s d=%s  new: hit %d t=%r  oracle: hit %d t=%r  old: hit %d t=%r" % (i, o[i], dd[i], b[0][i], float(b[1][i]), want[0][j], float(want[1][j]), a[0][i], float(a[1][i])))
    total += len(diff); bad += int((~okB).sum())
    # secondary rays: from the hit points (1e-3 off the surface along the normal) in random directions
    h = np.flatnonzero(b[0] >= 0)
    h = h[rs.choice(len(h), min(len(h), 2000000), replace=False)]
    dn = dd[h] / np.linalg.norm(dd[h], axis=1, keepdims=True)
    P = o[h] + dn * (b[1][h][:, None] - 1e-4)
    nd = rs.normal(size=(len(h), 3)).astype(np.float32)
    flip = (nd * b[2][h]).sum(1) < 0
    nd[flip] = -nd[flip]
    o, dd = P.astype(np.float32), nd
print("SUMMARY disagreements %d, new-builder errors against the oracle %d" % (total, bad))
