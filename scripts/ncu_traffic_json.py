"""One `ncu --set full` capture -> the small JSON bench.py's `roofline.traffic` reads (profiles/<name>.json).

    python scripts/ncu_traffic_json.py gpurun_out/prof.ncu-rep profiles/r02_generic_k1_ncu.json workload=W6 samples=4096
"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
extra = {}
for kv in sys.argv[3:]:
    k, v = kv.split("=", 1)
    extra[k] = int(v) if v.lstrip("-").isdigit() else v
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
d = dict(zip(hdr, r))
u = dict(zip(hdr, units))


def bytes_of(key):
    v = float(d[key].replace(",", ""))
    return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]])


doc = {"kernel": d.get("Kernel Name", "?")[:120], **extra,
       "dram_bytes_read": bytes_of("dram__bytes_read.sum"), "dram_bytes_write": bytes_of("dram__bytes_write.sum"),
       "duration_under_ncu": d["gpu__time_duration.sum"] + " " + u["gpu__time_duration.sum"],
       "pipe_tensor_cycles_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]),
       "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
       "registers_per_thread": float(d["launch__registers_per_thread"]), "source": f"{rep} (ncu --set full --clock-control none)"}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc))
