"""What would re-ordering the rays buy trace_kernel?  Walks the REAL paths of one iteration of the synthetic 100k-object scene bounce by bounce
(stage_raygen / stage_intersect / stage_shade, the render's own RNG) and, at every depth, sends the same rays through the BVH traversal in
several ORDERS -- as the compaction leaves them (pixel order), shuffled, sorted by the Morton code of the origin, by direction octant, by
both -- timing only the trace_kernel launch (PtbOptions.timeStages, stage 'intersect').   python tools/gpu_ray_order.py [max_depth]"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
if os.environ.get("PTB_SO"): m.LIB_PATH = os.path.join(ROOT, os.environ["PTB_SO"])
d = tempfile.mkdtemp(); path = os.path.join(d, "synthetic.txt")
open(path, "w").write(scenes.synthetic_text(100000))
scene = m.Scene.load(path)
max_depth = int(sys.argv[1]) if len(sys.argv) > 1 else 6
r = m.Renderer(scene, device=0, bsdf=1, batch_iters=1, time_stages=1)


def spread(v):                                  # 10 bits -> every third bit
    v = v.astype(np.uint64) & 0x3ff
    v = (v | (v << 16)) & 0x30000ff
    v = (v | (v << 8)) & 0x300f00f
    v = (v | (v << 4)) & 0x30c30c3
    v = (v | (v << 2)) & 0x9249249
    return v


def morton(o):
    lo, hi = o.min(0), o.max(0)
    q = np.clip(((o - lo) / np.maximum(hi - lo, 1e-9) * 1023).astype(np.int64), 0, 1023)
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def timed(o, dd):
    best = 1e30
    for _ in range(3):
        t0 = r.stats()["stage_us"]["intersect"]
        res = r.stage_intersect(o, dd, m.ACCEL_BVH)
        best = min(best, r.stats()["stage_us"]["intersect"] - t0)
    return best, res


rng = np.random.RandomState(1)
o, dd = r.stage_raygen(1)
P = len(o)
pix = np.arange(P, dtype=np.int32); col = np.ones((P, 3), np.float32); rem = np.full(P, scene.traceDepth, np.int32)
print("depth  rays      natural  shuffled   morton   octant  octant+morton   (us of trace_kernel, best of 3)")
for depth in range(max_depth):
    n = len(o)
    oct_ = (dd[:, 0] < 0).astype(np.uint64) | ((dd[:, 1] < 0).astype(np.uint64) << 1) | ((dd[:, 2] < 0).astype(np.uint64) << 2)
    mc = morton(o)
    orders = {"natural": np.arange(n), "shuffled": rng.permutation(n), "morton": np.argsort(mc, kind="stable"),
              "octant": np.argsort(oct_, kind="stable"), "octant+morton": np.argsort((oct_ << np.uint64(30)) | mc, kind="stable")}
    us = {}
    for name, perm in orders.items():
        us[name], res = timed(np.ascontiguousarray(o[perm]), np.ascontiguousarray(dd[perm]))
        if name == "natural": hit, t, nrm, outside = res
    print("%5d  %8d  %7.0f  %8.0f  %7.0f  %7.0f  %13.0f" % (depth, n, us["natural"], us["shuffled"], us["morton"], us["octant"], us["octant+morton"]), flush=True)
    itv = np.full(n, 1, np.int32); dep = np.full(n, depth, np.int32)
    go, gd, gc, grem = r.stage_shade(hit, t, nrm, outside, itv, pix, dep, o, dd, col, rem)
    keep = grem > 0
    o, dd, col, rem, pix = go[keep], gd[keep], gc[keep], grem[keep], pix[keep]
    if not len(o): break
