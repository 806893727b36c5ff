"""Regenerates the text summaries in profiles/ from ncu artefacts brought back in
gpurun_out/ (scratch, untracked).  usage: python profiles/summarize.py <tag>
Inputs (gpurun_out/): launches_<tag>.csv  = ncu --metrics gpu__time_duration.sum launch list
                      prof_<kernel>_<tag>.ncu-rep = ncu --set full --import-source captures"""
import collections, csv, glob, io, os, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum']


def launches(tag):
    path = os.path.join(REPO, 'gpurun_out', 'launches_%s.csv' % tag)
    if not os.path.exists(path):
        return
    rows = [l for l in open(path) if not l.startswith('==')]
    r = list(csv.DictReader(io.StringIO(''.join(rows))))
    agg = collections.OrderedDict()
    for x in r:
        if x['Metric Name'] != 'gpu__time_duration.sum':
            continue
        k = x['Kernel Name'].split('(')[0]
        v = float(x['Metric Value'].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1}[x['Metric Unit']]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ['# kernel time shares from gpurun_out/launches_%s.csv (ncu --metrics gpu__time_duration.sum '
           '--clock-control none; cold-cache, serialised launches: compare SHARES, not absolutes)' % tag,
           '%-34s %6s %12s %7s' % ('kernel', 'n', 'total_ms', 'share')]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('%-34s %6d %12.3f %7.3f' % (k, a[0], a[1], a[1] / tot))
    out.append('%-34s %6s %12.3f' % ('TOTAL', '', tot))
    open(os.path.join(REPO, 'profiles', '%s_launch_shares.txt' % tag), 'w').write('\n'.join(out) + '\n')
    # keep the raw list too (small)
    open(os.path.join(REPO, 'profiles', '%s_launches.csv' % tag), 'w').write(''.join(rows))


def full(tag):
    for rep in glob.glob(os.path.join(REPO, 'gpurun_out', 'prof_*_%s.ncu-rep' % tag)):
        name = os.path.basename(rep)[len('prof_'):-len('.ncu-rep')]
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        r = list(csv.reader(io.StringIO(raw)))
        h, u, v = r[0], r[1], r[2]
        out = ['# ncu --set full --clock-control none --import-source on, one launch; from gpurun_out/%s' % os.path.basename(rep),
               'kernel: ' + v[h.index('Kernel Name')]]
        d = {}
        for k in KEYS:
            if k in h:
                out.append('%-64s %16s %s' % (k, v[h.index(k)], u[h.index(k)]))
                d[k] = (v[h.index(k)], u[h.index(k)])
        try:
            sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            tr = sum(float(d[k][0]) * sc[d[k][1]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            out.append('%-64s %16.0f byte' % ('dram traffic per launch (read + write)', tr))
        except Exception:
            pass
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(src)))
        hh = rr[1]; rows = [x for x in rr[2:] if len(x) > 5]
        ia = hh.index('Instructions Executed'); isrc = hh.index('Source'); isamp = hh.index('# Samples')
        tot = sum(int(x[ia]) for x in rows if x[ia].isdigit())
        tots = sum(int(x[isamp]) for x in rows if x[isamp].isdigit())
        out.append('\n# hot SASS runs (share of executed warp instructions / of stall samples)')
        runs = []
        for i, x in enumerate(rows):
            c = int(x[ia]) if x[ia].isdigit() else 0
            if c > tot * 0.0004:
                if runs and i - runs[-1][1] <= 4:
                    runs[-1][1] = i; runs[-1][2] += c; runs[-1][3] += int(x[isamp])
                else:
                    runs.append([i, i, c, int(x[isamp])])
        for a in runs:
            if a[2] > tot * 0.02:
                out.append('sass rows %5d-%5d (%4d instr)  inst %.3f  samples %.3f  first: %s' % (
                    a[0], a[1], a[1] - a[0] + 1, a[2] / tot, a[3] / max(tots, 1), rows[a[0]][isrc].strip()[:48]))
        ops = collections.Counter()
        for x in rows:
            c = int(x[ia]) if x[ia].isdigit() else 0
            toks = x[isrc].strip().split()
            op = (toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')).split('.')[0]
            ops[op] += c
        out.append('\n# executed warp instructions by opcode (top 14)')
        for op, c in ops.most_common(14):
            out.append('%-10s %.3f' % (op, c / tot))
        open(os.path.join(REPO, 'profiles', '%s_ncu_full.txt' % name), 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    tag = sys.argv[1]
    launches(tag)
    full(tag)
