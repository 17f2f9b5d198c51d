"""Per-CUDA-source-line stall samples of one kernel in an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = {}
fname = None
hdr = None
inst = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        inst += 1; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or inst > 1:
        continue
    try:
        line = int(r[0]); s = int(r[hdr.index("# Samples")])
    except (ValueError, IndexError):
        continue
    if r[2] != "-":      # sass rows under the line
        continue
    key = (fname, line, r[1].strip()[:100])
    agg[key] = agg.get(key, 0) + s
tot = sum(agg.values()) or 1
for (f, l, src), s in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
    print(f"{100*s/tot:5.1f}%  {f}:{l:<4d} {src}")
