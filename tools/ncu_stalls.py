"""Warp-stall breakdown of one kernel of an .ncu-rep (source page): totals per stall reason and the top SASS instructions.
   python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io, os

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", os.environ.get("NCU_SKIP", "0"), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:120])
hdr, data = rows[1], rows[2:]
for _i, _r in enumerate(data):
    if _r and _r[0] == "Kernel Name":
        data = data[:_i]
        break
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall}
for r in data:
    for h in stall:
        try: tot[h] += int(r[ix[h]])
        except ValueError: pass
s = sum(tot.values())
print("instructions", len(data), "samples", s)
print({k[6:]: round(v / s, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]})
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top_n]:
    print(r[ix["# Samples"]], r[ix["Address"]][-5:], r[ix["Source"]][:80], {h[6:]: r[ix[h]] for h in stall if r[ix[h]] not in ("0", "")})
