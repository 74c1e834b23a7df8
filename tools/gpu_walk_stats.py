"""Candidate-walk histogram of the brute-force closest hit (tuning build -DPTB_WALK_STATS, variants/walk.so):
which candidates a lane puts through the object-space filter / the exact test at its 1st, 2nd, 3rd, 4th+ walk step.
   PTB_SO=variants/walk.so python tools/gpu_walk_stats.py [cornell.txt]"""
import os, sys, tempfile, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
m.LIB_PATH = os.path.join(ROOT, os.environ.get("PTB_SO", "variants/walk.so"))
d = tempfile.mkdtemp(); files = scenes.write_all(d)
name = (sys.argv[1:] or ["cornell.txt"])[0]
scene = m.Scene.load(files[name])
L = m.lib()
buf = (C.c_ulonglong * 128)()
with m.Renderer(scene, device=0, bsdf=1, batch_iters=16) as r:
    for label, loop_from in (("bounce 0 only (depth 1)", None), ("all bounces", None)):
        pass
    L.ptb_debug_walk_stats(buf)
    r.render(1, 16); r.sync()
    L.ptb_debug_walk_stats(buf)
    st = r.stats()
h = np.array(list(buf), dtype=np.int64).reshape(4, 8, 4)
print("segments", st["segments"], "alive", st["alive_per_depth"])
print("brute-force table order (cubes first, then spheres); columns: filtered, exact miss, exact hit accepted, exact hit rejected")
for it in range(4):
    print("walk step %d%s: total %d" % (it + 1, "+" if it == 3 else "", int(h[it].sum())))
    for c in range(8):
        if h[it, c].sum():
            print("   slot %d: %s" % (c, " ".join("%10d" % v for v in h[it, c])))
