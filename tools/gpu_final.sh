O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02final_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/r02final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02final_smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/r02final_bench.json 2> $O/r02final_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$O/r02final_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'launches', d['launches_per_step'])
print(d['roofline']['section_ms']); print([(k['kernel'][:9], k['ms']) for k in d['roofline']['kernels']]); print(d['other_configs']['cfg2']['ms_per_step'], d['other_configs']['cfg2']['value'], d['cpu_baseline']['value'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/r02final_launches_cfg3.csv python tools/prof_step.py 3 5 > $O/r02final_ncu_list.log 2>&1; echo "ncu list rc=$?"
