"""Exploratory GPU probe (not a test): parity spot checks + rough timings on one B200."""
import os, sys, time, tempfile, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import orclib, scenes
m = g.load_package()
orc = orclib.load("port")
d = tempfile.mkdtemp(); files = scenes.write_all(d)

def timeit(r, n_iters, reps=3):
    r.render(1, n_iters); r.sync()
    best = 1e9
    for _ in range(reps):
        r.reset(); r.sync()
        t0 = time.perf_counter(); r.render(1, n_iters); r.sync(); dt = time.perf_counter() - t0
        best = min(best, dt)
    st = r.stats()
    return best, st

out = {}
g.smoke()
for name, mode in (("cornell.txt", m.BSDF_FULL), ("cornell.txt", m.BSDF_GOLDEN_DIFFUSE), ("sphere.txt", m.BSDF_FULL), ("glass.txt", m.BSDF_FULL)):
    scene = m.Scene.load(files[name]); osc = orc.load_scene(files[name])
    omode = orclib.MODE_FULL if mode == m.BSDF_FULL else orclib.MODE_GOLDEN_DIFFUSE
    oimg, oalive, oseg = orc.render(osc, omode, 1, 1)
    for pipe, kw in (("fused_b1", dict(batch_iters=1)), ("staged_compact", dict(pipeline=m.PIPELINE_STAGED, compaction=1)),
                     ("staged_nocompact", dict(pipeline=m.PIPELINE_STAGED, compaction=0)),
                     ("staged_sort", dict(pipeline=m.PIPELINE_STAGED, compaction=1, sort_materials=1)),
                     ("fused_bvh", dict(batch_iters=1, accel=m.ACCEL_BVH))):
        with m.Renderer(scene, device=0, bsdf=mode, **kw) as r:
            r.render(1, 1); img = r.download(); st = r.stats()
        lit = oimg.sum(1) > 0
        rel = np.abs(img - oimg).max(1)[lit] / np.maximum(oimg.max(1)[lit], 1e-6)
        print(name, mode, pipe, "alive", st["alive_per_depth"], "oracle", list(map(int, oalive)),
              "lit within 1e-4: %.5f" % float((rel <= 1e-4).mean()), "nonlit max", float(np.abs(img[~lit]).max()) if (~lit).any() else 0,
              "exact pixels %.5f" % float((img.view(np.uint32) == oimg.view(np.uint32)).all(1).mean()), flush=True)
# timings
for name in ("cornell.txt", "sphere.txt"):
    scene = m.Scene.load(files[name])
    for B in (1, 4, 8, 16, 32):
        for graphs in (0, 1):
            with m.Renderer(scene, device=0, bsdf=m.BSDF_FULL, batch_iters=B, use_graphs=graphs) as r:
                n = 64 if B < 32 else 128
                dt, st = timeit(r, n)
                seg_per_iter = st["segments"] / st["iterations"]
                print("TIMING %s B=%d graphs=%d: %.1f us/iter, %.3e seg/s (%d seg/iter)" % (name, B, graphs, dt / n * 1e6, seg_per_iter * n / dt, seg_per_iter), flush=True)
    with m.Renderer(scene, device=0, bsdf=m.BSDF_FULL, pipeline=m.PIPELINE_STAGED, compaction=1) as r:
        dt, st = timeit(r, 32)
        print("TIMING %s staged: %.1f us/iter" % (name, dt / 32 * 1e6), flush=True)
