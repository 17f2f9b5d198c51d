"""Dynamic SASS opcode mix of the kernels in an .ncu-rep (warp-level instructions executed per opcode),
from `ncu --page source --csv` (needs --import-source on / -lineinfo).  Run where ncu is installed."""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for rr in rows:
    if rr and rr[0] == 'Kernel Name':
        cur = {'name': rr[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(rr)
seen = set()
for b in blocks:
    if b['name'] in seen or not b['rows']:
        continue
    seen.add(b['name'])
    h = b['rows'][0]
    try:
        ia, ie = h.index('Source'), h.index('# Instructions Executed') if '# Instructions Executed' in h else h.index('Instructions Executed')
    except ValueError:
        print(b['name'], 'columns:', h); continue
    isamp = h.index('# Samples') if '# Samples' in h else None
    mix, samp = collections.Counter(), collections.Counter()
    for x in b['rows'][1:]:
        if len(x) <= max(ia, ie): continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', x[ia])
        if not m: continue
        op = m.group(2)
        try: mix[op] += int(x[ie].replace(',', ''))
        except ValueError: pass
        if isamp is not None and x[isamp].isdigit(): samp[op] += int(x[isamp])
    tot, ts = sum(mix.values()), sum(samp.values()) or 1
    fp64 = sum(v for k, v in mix.items() if k in ('DADD', 'DMUL', 'DFMA', 'DSETP', 'DMNMX'))
    print('====', b['name'])
    print(f'  warp instructions executed: {tot}   FP64 (DADD+DMUL+DFMA+DSETP): {fp64} = {100*fp64/max(tot,1):.1f}%')
    print('  ' + ', '.join(f'{k} {v} ({100*v/tot:.1f}%, {100*samp[k]/ts:.1f}% of stall samples)' for k, v in mix.most_common(28)))
