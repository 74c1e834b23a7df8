"""A tiny render with the hit queue forced on (for compute-sanitizer): cornell 160 x 160, 4 iterations per wavefront, 8 iterations."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
p = os.path.join(tempfile.mkdtemp(), "c.txt"); open(p, "w").write(scenes.cornell_text(res=(160, 160)))
with m.Renderer(m.Scene.load(p), device=0, bsdf=1, batch_iters=4, hit_queue=1) as r:
    r.render(1, 8); r.sync()
    print("rendered", r.stats()["segments"], "segments with the hit queue on", flush=True)
