"""Top stall locations of an .ncu-rep captured with --import-source on (SASS page).
usage: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys

rep, N = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    tot = int(r[ix["# Samples"]] or 0)
    data.append((tot, k, r))
total = sum(d[0] for d in data)
agg = {s: sum(int(d[2][ix[s]] or 0) for d in data) for s in stalls}
print("total samples", total, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for tot, k, r in sorted(data, reverse=True)[:N]:
    top = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{100.0 * tot / total:5.2f}%  #{k:5d} exec {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip():60s} {top}")
