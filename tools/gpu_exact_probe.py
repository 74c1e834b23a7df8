"""How far does bit-exactness of the whole loop reach?  GPU image vs the CPU oracle over several scenes /
modes / iteration counts, printing the number of pixels whose bits differ (expected: 0 whenever the
accumulation order is the oracle's, i.e. sequential iterations or <= 2 deposits per pixel)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib, orclib, scenes
m = importlib.import_module("project3-cuda-path-tracer_b200")
orc = orclib.load("port")
d = tempfile.mkdtemp()
scenes.write_all(d)
for name, mode, iters, kw in [("cornell", 0, 4, dict(batch_iters=1)), ("cornell", 1, 4, dict(batch_iters=1)), ("glass", 1, 4, dict(batch_iters=1)),
                              ("cornell", 1, 4, dict(batch_iters=4)), ("cornell", 1, 16, dict(batch_iters=1, accel=2)),
                              ("glossy", 1, 2, dict(batch_iters=1)), ("sphere", 1, 3, dict(batch_iters=1))]:
    path = os.path.join(d, name + ".txt")
    if not os.path.exists(path):
        continue
    s = m.Scene.load(path); o = orc.load_scene(path)
    with m.Renderer(s, device=0, bsdf=mode, **kw) as r:
        r.render(1, iters); img = r.download(); st = r.stats()
    want, alive, seg = orc.render(o, mode, 1, iters)
    diff = (img.view(np.uint32) != want.view(np.uint32)).any(1)
    rel = np.abs(img - want).max(1) / np.maximum(want.max(1), 1e-12)
    print(name, mode, iters, kw, "pixels differing:", int(diff.sum()), "max rel:", float(rel[want.max(1) > 0].max() if (want.max(1) > 0).any() else 0),
          "alive equal:", list(alive) == st["alive_per_depth"], flush=True)
