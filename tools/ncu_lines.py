"""Warp-stall samples of one profiled kernel aggregated per CUDA SOURCE LINE of the kernel's own file (inlined helpers are charged to
the line that calls them).  The .ncu-rep gives samples per SASS address; the object file gives address -> line (nvdisasm inline info).
   python tools/ncu_lines.py report.ncu-rep build/chain_tc.o chain_tc_kernelILb0 [top_n]
The object must be the one the profiled library was linked from."""
import csv, io, os, re, subprocess, sys, tempfile

rep, obj, kname = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of, cur, infn, main_file = {}, None, False, None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        # the LAST annotation of a group names the outermost frame
        cur = (os.path.basename(m.group(3) or m.group(1)), int(m.group(4) or m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", os.environ.get("NCU_SKIP", "0"), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
for _i, _r in enumerate(data):
    if _r and _r[0] == "Kernel Name":
        data = data[:_i]
        break
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ix["Address"]], 16)
agg, inst = {}, {}
for r in data:
    off = int(r[ix["Address"]], 16) - base
    key = line_of.get(off, ("?", 0))
    a = agg.setdefault(key, {})
    for h in stall:
        v = int(r[ix[h]] or 0)
        if v: a[h[6:]] = a.get(h[6:], 0) + v
    inst[key] = inst.get(key, 0) + int(r[ix["Instructions Executed"]] or 0)
tot = sum(sum(a.values()) for a in agg.values())
src = {}
print("kernel %s: %d samples, %d warp instructions" % (rows[0][1][:70], tot, sum(inst.values())))
for key, a in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:top_n]:
    f, n = key
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.isfile(p) else []
    text = src[f][n - 1].strip()[:70] if 0 < n <= len(src[f]) else ""
    s = sum(a.values())
    top = ", ".join("%s %d" % kv for kv in sorted(a.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% inst %8d  %s:%d  [%s]  %s" % (100.0 * s / tot, inst[key], f, n, top, text))
