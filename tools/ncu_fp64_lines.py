"""FP64 instruction count per CUDA source line (executed warp instructions of D* opcodes) of one kernel."""
import csv, io, subprocess, sys, re
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, allops = {}, {}
fname = hdr = None
inst = 0
cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": inst += 1; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or inst > 1: continue
    ie = hdr.index("Instructions Executed")
    if r[0] != "":                        # a CUDA source line
        try: cur = (fname, int(r[0]), r[1].strip()[:90])
        except ValueError: cur = None
        continue
    if cur is None or len(r) < 4 or r[3] == "...": continue
    sass = r[3]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    if not m: continue
    op = m.group(2)
    try: n = int(r[ie])
    except (ValueError, IndexError): continue
    base = op.split(".")[0]
    allops[base] = allops.get(base, 0) + n
    if base in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "I2F", "F2F", "DMNMX"):
        agg[cur] = agg.get(cur, 0) + n
tot = sum(agg.values()) or 1
allt = sum(allops.values()) or 1
print("opcode mix (warp instructions):", ", ".join(f"{k} {100*v/allt:.1f}%" for k, v in sorted(allops.items(), key=lambda kv: -kv[1])[:14]))
print(f"FP64-class warp instructions: {tot} of {allt} ({100*tot/allt:.1f}%)")
for (f, l, src), s in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
    print(f"{100*s/tot:5.1f}%  {f}:{l:<4d} {src}")
