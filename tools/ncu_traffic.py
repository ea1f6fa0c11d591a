"""profiles/ncu_traffic.json from an ncu launch list with DRAM byte counters:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -s <skip> -c <launches of one step> --csv --log-file gpurun_out/traffic.csv python tools/prof_step.py 3 5
    python tools/ncu_traffic.py gpurun_out/traffic.csv 3 profiles/r02_step_traffic_cfg3.txt

Per launch: duration, DRAM read + write bytes.  The decoder entries (k_dec<1..3>, then the first two k_tc_gemm launches
behind k_dec<3>: dAct, dW) go to profiles/ncu_traffic.json together with a hash of the kernel sources, which bench.py
checks before it reports `roofline.traffic` (a capture of other sources is refused)."""
import collections
import csv
import re
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

fn, cfg = sys.argv[1], sys.argv[2]
with open(fn) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
launch = collections.OrderedDict()
for r in rows:
    k = r["ID"]
    ent = launch.setdefault(k, {"name": r["Kernel Name"], "grid": r.get("Grid Size", "")})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        ent["us"] = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        ent[name] = v * scale
# the last complete step of the capture (between the last two k_finalize_step launches): the warm one
items = list(launch.items())
fin = [i for i, (_, e) in enumerate(items) if 'k_finalize_step' in e['name']]
if len(fin) >= 2:
    items = items[fin[-2] + 1:fin[-1] + 1]
out_lines = []
dec, after2, seen, tot_us, tot_b = {}, False, set(), 0.0, 0.0
for k, e in items:
    b = e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
    tot_us += e.get("us", 0.0)
    tot_b += b
    short = e["name"].replace("drv::<unnamed>::", "").replace("void ", "")[:70]
    out_lines.append(f"{e.get('us', 0):8.1f} us  {e.get('dram__bytes_read.sum', 0) / 1e6:8.2f} MB rd  {e.get('dram__bytes_write.sum', 0) / 1e6:8.2f} MB wr  {short}  grid={e['grid']}")
    n = e["name"]
    for ph in (1, 2, 3):
        if f"k_dec<{ph}," in n or f"k_dec<(int){ph}" in n:
            dec[f"k_dec<{ph}>"] = b
            if ph == 2:
                after2, seen = True, set()
    if "k_dec3w" in n:
        dec["k_dec3w"] = b
    mm = re.search(r"k_tc_gemm<\(?(?:bool\))?(\d), \(?(?:bool\))?(\d),", n)
    if mm and after2:
        # behind phase 2: the first GEMM with both operands MN-major is the decoder's dW (log_r half when k_dec3w runs),
        # the first K-major x MN-major one is dAct
        kind = "dW" if mm.group(1) == "1" and mm.group(2) == "1" else ("dAct" if mm.group(1) == "0" and mm.group(2) == "1" else None)
        if kind and kind not in seen:
            dec[kind] = b
            seen.add(kind)
out_lines.append(f"step total {tot_us:.1f} us (serialised, cold caches), {tot_b / 1e6:.1f} MB DRAM traffic, {len(items)} launches")
text = "\n".join(out_lines)
print(text)
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        f.write(text + "\n")
dec["decoder_step_bytes"] = sum(dec.values())
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
cur = {}
if os.path.exists(path):
    with open(path) as f:
        cur = json.load(f)
if cur.get("source_hash") != bench.source_hash():
    cur = {}
cur[f"cfg{cfg}"] = dec
cur["source_hash"] = bench.source_hash()
cur["source"] = f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch ({os.path.basename(fn)}, B200, direct launches of one warm step)"
with open(path, "w") as f:
    json.dump(cur, f, indent=1)
print("wrote", path, dec)
