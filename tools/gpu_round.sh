#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list (+ DRAM bytes), ncu --set full of the decoder kernels,
# compute-sanitizer logs.  Outputs under gpurun_out/$TAG_*.   usage: bash tools/gpu_round.sh TAG [skip_tests]
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
  tail -3 $O/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_cfg3.json 2> $O/${TAG}_bench_cfg3.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$O/${TAG}_bench_cfg3.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launches', d['launches_per_step'])
print(d['roofline']['section_ms']); print([(k['kernel'][:9], k['ms']) for k in d['roofline']['kernels']])
PY
# launch list with DRAM bytes (5 direct steps; the last one is the warm one)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file $O/${TAG}_launches_cfg3.csv python tools/prof_step.py 3 5 > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
# full capture of the decoder kernels of the 4th step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dec -s 9 -c 3 -o $O/${TAG}_kdec -f \
  python tools/prof_step.py 3 5 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/${TAG}_kdec.ncu-rep --page raw --csv > $O/${TAG}_kdec_raw.csv 2>/dev/null
# the same launch list with the caches left alone between kernels (ncu's default flushes them before every kernel, which
# turns every L2-resident hand-off between two kernels of the step into DRAM reads)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 600 --csv \
  --log-file $O/${TAG}_launches_cfg3_warmcache.csv python tools/prof_step.py 3 5 > $O/${TAG}_ncu_list2.log 2>&1; echo "ncu list (cache-control none) rc=$?"
# (compute-sanitizer is closed on this pool: profiles/r02_sanitizer_unavailable.txt)
