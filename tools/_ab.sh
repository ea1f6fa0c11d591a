# usage: bash tools/_ab.sh VAR   -> bench with VAR unset, then VAR=1
for v in 0 1; do
  if [ $v = 1 ]; then export $1=1; fi
  timeout 150 python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/ab$v.json
done
python - <<'PY'
import json
for v in range(2):
    try:
        d=json.load(open('gpurun_out/ab%d.json'%v)); print(v, round(d['ms_per_step'],4), [(k['kernel'][:8],k['ms']) for k in d['roofline']['kernels']])
    except Exception as e: print(v,'ERR',e)
PY
