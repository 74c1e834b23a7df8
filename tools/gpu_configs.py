"""Timings of the other BASELINE.json configs on one GPU (configs[1] sphere 800^2, configs[3] cornell 3840x2160,
configs[4] synthetic 100k objects 1920x1080 depth 16).  One JSON line per config.
   python tools/gpu_configs.py [sphere] [cornell4k] [synthetic]"""
import json, os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
d = tempfile.mkdtemp()
which = sys.argv[1:] or ["sphere", "cornell4k", "synthetic"]
texts = {"sphere": lambda: scenes.sphere_text(), "cornell4k": lambda: scenes.cornell_text(res=(3840, 2160)),
         "synthetic": lambda: scenes.synthetic_text(100000), "cornell": lambda: scenes.cornell_text()}
for name in which:
    path = os.path.join(d, name + ".txt")
    open(path, "w").write(texts[name]())
    t0 = time.perf_counter(); scene = m.Scene.load(path); t_load = time.perf_counter() - t0
    for fbc in (0, 1):
        t0 = time.perf_counter()
        with m.Renderer(scene, device=0, bsdf=1, first_bounce_cache=fbc) as r:
            t_init = time.perf_counter() - t0
            B = r.stats()["batch_iters"]
            n = max(B, 32 if name == "synthetic" else 128)
            n = (n // B) * B
            r.render(1, n); r.sync()
            best = 1e9
            for _ in range(3):
                r.reset(); r.sync(); t0 = time.perf_counter(); r.render(1, n); r.sync(); best = min(best, time.perf_counter() - t0)
            st = r.stats()
        spi = st["segments"] / st["iterations"]
        print(json.dumps({"config": name, "first_bounce_cache": fbc, "pixels": scene.pixels, "geoms": len(scene.geoms), "trace_depth": scene.traceDepth,
                          "bvh_nodes": st["bvh_nodes"], "batch_iters": B, "iterations_timed": n, "us_per_iter": best / n * 1e6,
                          "iters_per_s": n / best, "segments_per_iter": spi, "segments_per_s": spi * n / best,
                          "scene_load_s": t_load, "init_s": t_init}), flush=True)
