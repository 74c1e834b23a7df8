"""BVH fused path (trace_bvh_kernel + bounce kernel) against brute force on a 1k-object scene (bit for bit),
then timing of the synthetic 100k-object scene."""
import os, sys, time, tempfile, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import scenes
m = g.load_package()
d = tempfile.mkdtemp()
p = os.path.join(d, "syn1k.txt"); open(p, "w").write(scenes.synthetic_text(n_objects=1000, res=(640, 360), depth=8, seed=98))
s = m.Scene.load(p)
imgs = {}
for name, kw in (("brute", dict(accel=m.ACCEL_BRUTE)), ("bvh", dict(accel=m.ACCEL_BVH)), ("bvh_fbc", dict(accel=m.ACCEL_BVH, first_bounce_cache=1)),
                 ("bvh_staged", dict(accel=m.ACCEL_BVH, pipeline=m.PIPELINE_STAGED))):
    with m.Renderer(s, device=0, bsdf=1, batch_iters=1, **kw) as r:
        r.render(1, 3); imgs[name] = r.download(); st = r.stats()
        print(name, st["alive_per_depth"][:8], st["kernel_launches"], flush=True)
for k in ("bvh", "bvh_fbc", "bvh_staged"):
    print(k, "== brute bit for bit:", bool(np.array_equal(imgs[k].view(np.uint32), imgs["brute"].view(np.uint32))), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "nosyn": sys.exit(0)
p = os.path.join(d, "synthetic.txt"); open(p, "w").write(scenes.synthetic_text(100000))
s = m.Scene.load(p)
for fbc in (0, 1):
    with m.Renderer(s, device=0, bsdf=1, first_bounce_cache=fbc) as r:
        B = r.stats()["batch_iters"]; n = 2 * B
        r.render(1, n); r.sync()
        best = 1e9
        for _ in range(2):
            r.reset(); r.sync(); t0 = time.perf_counter(); r.render(1, n); r.sync(); best = min(best, time.perf_counter() - t0)
        st = r.stats(); spi = st["segments"] / st["iterations"]
        print("TIMING synthetic fbc=%d B=%d: %.1f us/iter, %.3e seg/s" % (fbc, B, best / n * 1e6, spi * n / best), flush=True)
