"""Renders a few cornell wavefronts (batch 16) through the fused pipeline: the short command that the ncu
captures wrap.  Usage: python tools/gpu_one_batch.py [batches] [scene] [batch_iters]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
if os.environ.get("PTB_SO"): m.LIB_PATH = os.path.join(ROOT, os.environ["PTB_SO"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name = sys.argv[2] if len(sys.argv) > 2 else "cornell.txt"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
d = tempfile.mkdtemp(); files = scenes.write_all(d)
scene = m.Scene.load(files[name])
with m.Renderer(scene, device=0, bsdf=1, batch_iters=B, use_graphs=0) as r:
    r.render(1, n * B); r.sync()
    print(r.stats()["alive_per_depth"])
