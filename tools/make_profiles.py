#!/usr/bin/env python
"""turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under profiles/:
   profiles/launches_<tag>.csv/.md   per-launch device times of one bench step (ncu --metrics gpu__time_duration.sum)
   profiles/ncu_<tag>.txt/.json      key metrics of the --set full capture of the hot kernels
   profiles/ncu_traffic.json         dram bytes per launch per kernel (bench.py quotes it as roofline.traffic)
usage: python tools/make_profiles.py <tag> [launches.csv] [prof.ncu-rep]"""
import csv, io, json, os, re, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'gpurun_out', f'launches_{tag}.csv')
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, 'gpurun_out', f'prof_{tag}.ncu-rep')
out = os.path.join(ROOT, 'profiles')
os.makedirs(out, exist_ok=True)

def short(name):
    m = re.match(r'(?:void )?(?:iqw::)?([A-Za-z0-9_]+)', name)
    return m.group(1) if m else name

if os.path.exists(launches):
    text = open(launches).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = collections.OrderedDict()
    ours = [r for r in rows if r.get('Metric Name') == 'gpu__time_duration.sum']
    # keep the last complete step: from the last stft_kernel launch on
    # (the whole-capture launches, not the short per-chunk ones of the host-input (e2e) path that follow)
    def _us(r):
        return float(r['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}.get(r['Metric Unit'], 1)
    stft = [i for i, r in enumerate(ours) if re.search(r'stft\w*_kernel', r['Kernel Name'])]
    longest = max((_us(ours[i]) for i in stft), default=0.0)
    last = max((i for i in stft if _us(ours[i]) >= 0.5 * longest), default=0)
    nxt = min((i for i in stft if i > last), default=len(ours))
    ours = ours[last:nxt]
    with open(os.path.join(out, f'launches_{tag}.csv'), 'w') as f:
        f.write('id,kernel,grid,block,duration_us\n')
        for r in ours:
            us = float(r['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}.get(r['Metric Unit'], 1)
            k = short(r['Kernel Name'])
            f.write(f"{r['ID']},{k},\"{r['Grid Size']}\",\"{r['Block Size']}\",{us:.3f}\n")
            a = per.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
    total = sum(v[1] for v in per.values())
    with open(os.path.join(out, f'launches_{tag}.md'), 'w') as f:
        f.write(f'# per-kernel device time, one profiled run of bench.py (ncu, cold-cache, serialised) -- {tag}\n\n')
        f.write('| kernel | launches | total us | share |\n|---|---|---|---|\n')
        for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| {k} | {n} | {us:.1f} | {us / total:.3f} |\n')
    print('wrote launches summary,', len(ours), 'launches')

if os.path.exists(rep):
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
            'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
            'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
            'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tma.sum']
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    res, traffic = [], {}
    for r in rows[2:]:
        d = {'kernel': short(r[hdr.index('Kernel Name')]), 'kernel_full': r[hdr.index('Kernel Name')]}
        for i, h in enumerate(hdr):
            base = h.split('.', 2)[-1] if h.count('.') >= 2 and h.split('.')[1] == 'TriageCompute' else h
            if base in want:
                d[base] = f'{r[i]} {units[i]}'.strip()
        stalls = {}
        for i, h in enumerate(hdr):
            m = re.match(r'smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio', h)
            if m:
                try: stalls[m.group(1)] = float(r[i])
                except ValueError: pass
        d['top_stalls'] = ', '.join(f'{k} {v:.2f}' for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        res.append(d)
        def nbytes(key):
            v, u = d[key].split()
            return float(v.replace(',', '')) * mult[u]
        traffic.setdefault(d['kernel'], []).append(nbytes('dram__bytes_read.sum') + nbytes('dram__bytes_write.sum'))
    with open(os.path.join(out, f'ncu_{tag}.txt'), 'w') as f:
        f.write(f'ncu --set full --clock-control none, one launch per hot kernel inside bench.py ({tag})\n\n')
        for d in res:
            f.write(f"== {d['kernel_full']}\n")
            for k, v in d.items():
                if k not in ('kernel', 'kernel_full'):
                    f.write(f'   {k}: {v}\n')
            f.write('\n')
    json.dump(res, open(os.path.join(out, f'ncu_{tag}.json'), 'w'), indent=1)
    tr = {k: sum(v) / len(v) for k, v in traffic.items()}
    for k in list(tr):          # bench.py looks kernel 1 up under the name of its profile scope
        if re.fullmatch(r'stft\w*_kernel', k):
            tr.setdefault('stft_kernel', tr[k])
    json.dump(tr, open(os.path.join(out, 'ncu_traffic.json'), 'w'), indent=1)
    print('wrote ncu summary for', [d['kernel'] for d in res])
