"""Timing of the synthetic 100k-object scene (BVH path) for tuning builds:  PTB_SO=variants/x.so python tools/gpu_time_synth.py"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
if os.environ.get("PTB_SO"): m.LIB_PATH = os.path.join(ROOT, os.environ["PTB_SO"])
d = tempfile.mkdtemp(); path = os.path.join(d, "synthetic.txt")
open(path, "w").write(scenes.synthetic_text(100000))
scene = m.Scene.load(path)
B = int(os.environ.get("PTB_B", "16"))
with m.Renderer(scene, device=0, bsdf=1, batch_iters=B) as r:            # PTB_B=0: the library's default wavefront
    B = r.stats()["batch_iters"]
    n = 2 * B
    r.render(1, n); r.sync()
    best = 1e9
    for _ in range(3):
        r.reset(); r.sync(); t0 = time.perf_counter(); r.render(1, n); r.sync(); best = min(best, time.perf_counter() - t0)
    st = r.stats()
    img = r.download()
    print(os.environ.get("PTB_SO", "default"), "TIMING synthetic B=%d: %.1f us/iter, %.3e seg/s, image sum %.6e, alive %s" % (B, best / n * 1e6, st["segments"] / st["iterations"] * n / best, float(img.astype("float64").sum()), st["alive_per_depth"][:4]), flush=True)
