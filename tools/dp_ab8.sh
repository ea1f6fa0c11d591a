#!/bin/bash
# A/B of data-parallel knobs at N GPUs: bash tools/dp_ab8.sh N "ENV1=.. ENV2=.." "ENV.."   (one bench line per variant)
N=$1; shift
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+i)) bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline --other-configs '' 2>/dev/null | tail -1 > gpurun_out/dpab_$i.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dpab_$i.json")); a=d["all_reduce"]
    print("$v", d["n_gpus"], "ms", round(d["ms_per_step"],4), "without", round(a["ms_per_step_without"],4), "exposed", round(a["exposed_ms"],4), "alone", round(a["alone_ms"],4), "dp", d["dp_check"]["timed_step"]["max_rel_err"])
except Exception as e: print("$v ERR", e)
PY
done
