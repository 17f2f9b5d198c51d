"""Static SASS instruction count of one kernel per source line of its tile function (code-size profile):
python tools/sass_by_line.py <disassembly from `nvdisasm -gi cubin`> <mangled kernel substring> <line of the tile call>"""
import collections, re, sys
path, kern, call_line = sys.argv[1], sys.argv[2], int(sys.argv[3])
fn, group, fresh = None, [], True
cnt = collections.Counter()
for l in open(path):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        fn = m.group(1); group = []; continue
    if fn is None or kern not in fn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if fresh:
            group = []; fresh = False
        group.append((m.group(1).split('/')[-1], int(m.group(2)), int(m.group(4)) if m.group(4) else None))
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        fresh = True
        key = None
        for f, ln, at in group:
            if at == call_line and f == 'chs_kernels.cuh':
                key = ln
        cnt[key if key is not None else (group[0][:2] if group else None)] += 1
tot = sum(cnt.values())
print(kern, tot, 'instructions')
for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:40]:
    print(f'  {str(k):>30s} {v:6d} {100*v/tot:5.1f}%')
