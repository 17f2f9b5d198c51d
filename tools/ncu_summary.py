"""Prints the key counters / stall reasons / hottest SASS lines of an .ncu-rep (run where ncu is installed)."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__thread_inst_executed_per_inst_executed.ratio']
seen = set()
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d['Kernel Name']
    if name in seen:
        continue
    seen.add(name)
    print('====', name)
    for k in KEYS:
        if k in d:
            print(f'  {k:78s} {d[k]}')
    st = [(k.replace('smsp__pcsamp_warps_issue_stalled_', ''), float(v.replace(',', ''))) for k, v in d.items()
          if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued') and v not in ('', 'n/a')]
    tot = sum(v for _, v in st) or 1
    print('  stalls: ' + ', '.join(f'{k} {100*v/tot:.1f}%' for k, v in sorted(st, key=lambda x: -x[1])[:9]))
    pat = re.search(r'(k_\w+)<', name)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + (pat.group(1) if pat else name)],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    # several kernel instances are concatenated, each introduced by a "Kernel Name" row + header row
    blocks, cur = [], None
    for rr in srows:
        if rr and rr[0] == 'Kernel Name':
            cur = {'name': rr[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(rr)
    for b in blocks:
        if b['name'].replace('chs::', '').replace('(int)', '') .replace(' ', '') != name.replace(' ', ''):
            continue
        h = b['rows'][0]
        data = b['rows'][1:]
        ia, isamp = h.index('Source'), h.index('# Samples')
        tots = sum(int(x[isamp]) for x in data if len(x) > isamp and x[isamp].isdigit())
        top = sorted([(int(x[isamp]), i, x[ia]) for i, x in enumerate(data) if len(x) > isamp and x[isamp].isdigit()], reverse=True)[:top_n]
        print(f'  hottest SASS (of {tots} samples, {len(data)} instructions):')
        for s_, i, t in top:
            print(f'    {100*s_/tots:5.1f}%  #{i:5d}  {t}')
        break
