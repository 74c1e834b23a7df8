"""Timing only: cornell / sphere through the fused pipeline (B=16), best of 3.  python tools/gpu_time.py [scene ...]"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
if os.environ.get("PTB_SO"): m.LIB_PATH = os.path.join(ROOT, os.environ["PTB_SO"])
d = tempfile.mkdtemp(); files = scenes.write_all(d)
for name in (sys.argv[1:] or ["cornell.txt"]):
    scene = m.Scene.load(files[name])
    B = int(os.environ.get("PTB_B", "16"))
    with m.Renderer(scene, device=0, bsdf=1, batch_iters=B, first_bounce_cache=int(os.environ.get("PTB_FBC", "0")), leave_test=int(os.environ.get("PTB_LEAVE", "1")),
                    hit_queue=(int(os.environ["PTB_HQ"]) if "PTB_HQ" in os.environ else None)) as r:
        n = max(8 * B, 256)
        r.render(1, n); r.sync()
        best = 1e9
        for _ in range(3):
            r.reset(); r.sync(); t0 = time.perf_counter(); r.render(1, n); r.sync(); best = min(best, time.perf_counter() - t0)
        st = r.stats()
        print(os.environ.get("PTB_SO", "default"), "TIMING %s B=%d mult=%s leave=%s hq=%s: %.1f us/iter, %.3e seg/s" % (name, B, os.environ.get("PTB_GRID_MULT", "1"), os.environ.get("PTB_LEAVE", "1"), os.environ.get("PTB_HQ", "auto"), best / n * 1e6, st["segments"] / st["iterations"] * n / best), flush=True)
