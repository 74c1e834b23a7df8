#!/bin/bash
# pass 3d: ptb_pathtrace look-ahead (PtbOptions.lookahead): parity test, the four drop-in legs of bench.py
set -u
mkdir -p gpurun_out
TAG=${1:-r3d}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lookahead or lazy_image or reference_entry" 2>&1 | tail -4
( time python bench.py --no-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value %.4e' % d['value'])
for k, v in d['e2e_per_iteration_dropin'].items(): print(k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != 'note'})
print(json.dumps(d['skeleton_gpu_baseline']['ours_over_skeleton']))
PY
