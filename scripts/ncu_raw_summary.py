"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one block per profiled launch with the
metrics the roofline cites (duration, DRAM bytes, L2/L1 throughput, tensor-pipe activity ...).
usage: ncu_raw_summary.py RAW.csv [--json OUT.json]  (json: kernel -> mean DRAM traffic per launch)"""
import csv
import json
import re
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
traffic = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
    print(f"--- {name}  (launch id {r[col['ID']]})")
    for w in WANT:
        if w in col:
            print(f"    {w:72s} {r[col[w]]:>16s} {units[col[w]]}")
    def f(x):
        return float(r[col[x]].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    t = f("dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]] + \
        f("dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]]
    traffic.setdefault(name, []).append(t)
print()
for k, v in traffic.items():
    print(f"DRAM traffic per launch (read+write, mean of {len(v)}): {k}: {sum(v) / len(v) / 1e6:.1f} MB")
if "--json" in sys.argv:
    out = sys.argv[sys.argv.index("--json") + 1]
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(out, "w"), indent=1)
