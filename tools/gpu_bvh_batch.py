"""Renders a few wavefronts of the synthetic 100k-object scene (BVH path) -- the command ncu wraps."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
d = tempfile.mkdtemp(); path = os.path.join(d, "synthetic.txt")
open(path, "w").write(scenes.synthetic_text(100000))
scene = m.Scene.load(path)
with m.Renderer(scene, device=0, bsdf=1, batch_iters=B, use_graphs=0) as r:
    r.render(1, n * B); r.sync()
    print(r.stats()["alive_per_depth"])
