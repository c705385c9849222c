"""Collects dram__bytes_read/write.sum of the ADMM kernel of every bench configuration from the ncu captures of
tools/run_ncu.sh (gpurun_out/prof_c{2..5}.ncu-rep) into profiles/traffic.json, which bench.py reports as roofline.traffic
when the batch matches.  usage: python tools/make_traffic.py"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BATCH = {"c2": 4096, "c3": 131072, "c4": 65536, "c5": 65536}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for c, batch in BATCH.items():
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{c}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * UNIT[d["dram__bytes_read.sum"][1]]
    wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * UNIT[d["dram__bytes_write.sum"][1]]
    out[c] = {"batch_per_gpu": batch, "kernel": d["Kernel Name"][0].split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr,
              "gpu_time_ns_under_ncu": d["gpu__time_duration.sum"][0] + " " + d["gpu__time_duration.sum"][1],
              "source": f"ncu --set full --clock-control none, tools/run_ncu.sh, bench.py --config {c} (first timed launch)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
