#!/bin/bash
# quick parity + bench probe:  bash tools/gpu_quick.sh TAG "pytest -k expression"
TAG=${1:-q}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --other-configs '' > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('$O/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('ms/step', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], 'launches', d['launches_per_step'])
    print(d['roofline']['section_ms']); print([(k['kernel'][:9], k['ms']) for k in d['roofline']['kernels']])
except Exception as e: print('ERR', e); print(open('$O/${TAG}_bench.err').read()[-2000:])
PY
