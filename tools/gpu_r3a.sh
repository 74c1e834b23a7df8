#!/bin/bash
# pass 3a: material regrouping in the fused BVH shading kernel (bounce_regroup_kernel) on / off; spill-path test with the 8-entry tuning build; GPU suite; ncu
set -u
mkdir -p gpurun_out
TAG=${1:-r3a}
python - <<'PY' | tee gpurun_out/${TAG}_regroup.txt
import os, sys, time, tempfile
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g, scenes, numpy as np
m = g.load_package()
d = tempfile.mkdtemp(); path = os.path.join(d, "synthetic.txt"); open(path, "w").write(scenes.synthetic_text(100000))
scene = m.Scene.load(path)
imgs = {}
for B in (16, 0):
    for rg in (0, 1):
        with m.Renderer(scene, device=0, bsdf=1, batch_iters=B, regroup_materials=rg) as r:
            Bn = r.stats()["batch_iters"]; n = 2 * Bn
            r.render(1, n); r.sync()
            best = 1e9
            for _ in range(3):
                r.reset(); r.sync(); t0 = time.perf_counter(); r.render(1, n); r.sync(); best = min(best, time.perf_counter() - t0)
            st = r.stats(); img = r.download()
            imgs[(B, rg)] = (img, st["alive_per_depth"])
            print("regroup=%d TIMING synthetic B=%d: %.1f us/iter, %.3e seg/s, image sum %.6e" % (rg, Bn, best / n * 1e6, st["segments"] / st["iterations"] * n / best, float(img.astype("float64").sum())), flush=True)
    a, b = imgs[(B, 0)], imgs[(B, 1)]
    print("  identical bits:", bool(np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))), " identical counts:", a[1] == b[1])
PY
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:'bounce_regroup' -s 4 -c 2 -f -o gpurun_out/${TAG}_regroup \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_regroup_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_regroup_ncu.log
