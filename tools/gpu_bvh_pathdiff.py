"""Which of two BVH builds is right where they differ: walk the REAL paths of the synthetic 100k-object scene bounce by
bounce (stage_raygen / stage_intersect / stage_shade, the render's own RNG) through both BVHs and send every ray on which
they disagree to the CPU oracle's brute-force loop over all the geoms.
   python tools/gpu_bvh_pathdiff.py [first_iter] [n_iters]
OLD_ALL_AXES / NEW_ALL_AXES / NEW_TIGHT / NEW_EXTRA_PAD choose the two builds; OLD_ALL_AXES=1 NEW_EXTRA_PAD=0.3 asks whether the
shipped leaf boxes lose a hit that boxes 0.3 world units fatter would find."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import orclib, scenes
m = g.load_package(); orc = orclib.load("port")
m.LIB_PATH = os.path.join(os.path.dirname(m.__file__), "libptb200_tuning.so")      # the PTB_BVH_* knobs exist in the tuning build only
d = tempfile.mkdtemp(); path = os.path.join(d, "synthetic.txt")
open(path, "w").write(scenes.synthetic_text(100000))
scene = m.Scene.load(path); osc = orc.load_scene(path)
first = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 4


def make(env):
    for k, v in env.items(): os.environ[k] = v
    return m.Renderer(scene, device=0, bsdf=1, batch_iters=1)


r_old = make({"PTB_BVH_ALL_AXES": os.environ.get("OLD_ALL_AXES", "0"), "PTB_BVH_TIGHT_SPHERES": "0", "PTB_BVH_EXTRA_PAD": "0",
              "PTB_BVH_FAR_PAD": os.environ.get("OLD_FAR_PAD", "1")})
r_new = make({"PTB_BVH_ALL_AXES": os.environ.get("NEW_ALL_AXES", "1"), "PTB_BVH_TIGHT_SPHERES": os.environ.get("NEW_TIGHT", "0"),
              "PTB_BVH_EXTRA_PAD": os.environ.get("NEW_EXTRA_PAD", "0"), "PTB_BVH_FAR_PAD": os.environ.get("NEW_FAR_PAD", "1")})
bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
tot_rays = tot_diff = wrong_old = wrong_new = 0
for it in range(first, first + n_it):
    o, dd = r_new.stage_raygen(it)
    P = len(o)
    pix = np.arange(P, dtype=np.int32); col = np.ones((P, 3), np.float32); rem = np.full(P, scene.traceDepth, np.int32)
    for depth in range(scene.traceDepth):
        a = r_old.stage_intersect(o, dd, m.ACCEL_BVH)
        b = r_new.stage_intersect(o, dd, m.ACCEL_BVH)
        same = (a[0] == b[0]) & (bits(a[1]) == bits(b[1])) & (bits(a[2]) == bits(b[2])).all(1) & (a[3] == b[3])
        diff = np.flatnonzero(~same)
        tot_rays += len(o); tot_diff += len(diff)
        hit, t, nrm, outside = b
        if len(diff):
            want = orc.closest_hit(osc.geoms, o[diff], dd[diff])
            okA = (a[0][diff] == want[0]) & (bits(a[1][diff]) == bits(want[1]))
            okB = (b[0][diff] == want[0]) & (bits(b[1][diff]) == bits(want[1]))
            wrong_old += int((~okA).sum()); wrong_new += int((~okB).sum())
            for j, i in enumerate(diff[:8]):
                print("iter %d depth %d pixel %d: old hit %d t=%r | new hit %d t=%r | oracle hit %d t=%r  o=%s d=%s" % (
                    it, depth, pix[i], a[0][i], float(a[1][i]), b[0][i], float(b[1][i]), want[0][j], float(want[1][j]), o[i], dd[i]), flush=True)
            # continue the walk on the oracle's answer
            hit = hit.copy(); t = t.copy(); nrm = nrm.copy(); outside = outside.copy()
            hit[diff], t[diff], nrm[diff], outside[diff] = want
        itv = np.full(len(pix), it, np.int32); dep = np.full(len(pix), depth, np.int32)
        go, gd, gc, grem = r_new.stage_shade(hit, t, nrm, outside, itv, pix, dep, o, dd, col, rem)
        keep = grem > 0
        o, dd, col, rem, pix = go[keep], gd[keep], gc[keep], grem[keep], pix[keep]
        if not len(o): break
    print("iter %d done: rays so far %d, disagreements %d (old wrong %d, new wrong %d)" % (it, tot_rays, tot_diff, wrong_old, wrong_new), flush=True)
print("SUMMARY rays %d disagreements %d old-builder wrong %d new-builder wrong %d" % (tot_rays, tot_diff, wrong_old, wrong_new))
