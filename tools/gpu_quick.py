"""Quick GPU check used while optimising: parity on first bounce + iter-1 image + timing."""
import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import orclib, scenes
m = g.load_package(); orc = orclib.load("port")
d = tempfile.mkdtemp(); files = scenes.write_all(d)
for name in ("cornell.txt", "glass.txt", "sphere.txt"):
    scene = m.Scene.load(files[name]); osc = orc.load_scene(files[name])
    with m.Renderer(scene, device=0, bsdf=1, batch_iters=1) as r:
        o, dd = r.stage_raygen()
        want = orc.closest_hit(osc.geoms, o, dd)
        for acc in (1, 2):
            got = r.stage_intersect(o, dd, acc)
            ok = np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)) and np.array_equal(got[2].view(np.uint32), want[2].view(np.uint32)) and np.array_equal(got[3], want[3])
            print(name, "first bounce accel", acc, "bit-exact:", ok)
        rs = np.random.RandomState(3); n = 400000
        ro = rs.uniform(-6, 6, (n, 3)).astype(np.float32) + np.float32([0, 5, 0]); rd = rs.normal(size=(n, 3)).astype(np.float32)
        rd[::50, 0] = 0; rd[::77, 1] = 0; rd[::91] = [0, 0, -1]
        want = orc.closest_hit(osc.geoms, ro, rd)
        for acc in (1, 2):
            got = r.stage_intersect(ro, rd, acc)
            ok = np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)) and np.array_equal(got[2].view(np.uint32), want[2].view(np.uint32)) and np.array_equal(got[3], want[3])
            print(name, "random rays accel", acc, "bit-exact:", ok, "hits", int((want[0] >= 0).sum()))
        r.render(1, 2); img = r.download(); st = r.stats()
    oimg, oalive, oseg = orc.render(osc, 1, 1, 2)
    print(name, "alive", st["alive_per_depth"], "oracle", list(map(int, oalive)), "exact pixels %.6f" % float((img.view(np.uint32) == oimg.view(np.uint32)).all(1).mean()), flush=True)
for name in ("cornell.txt", "sphere.txt"):
    scene = m.Scene.load(files[name])
    for B in (16,):
        with m.Renderer(scene, device=0, bsdf=1, batch_iters=B) as r:
            n = 128
            r.render(1, n); r.sync()
            best = 1e9
            for _ in range(3):
                r.reset(); r.sync(); t0 = time.perf_counter(); r.render(1, n); r.sync(); best = min(best, time.perf_counter() - t0)
            st = r.stats()
            print("TIMING %s B=%d: %.1f us/iter, %.3e seg/s" % (name, B, best / n * 1e6, st["segments"] / st["iterations"] * n / best), flush=True)
