"""helper (not a test): summarise an ncu --import-source capture: hot SASS runs by
executed instructions.  usage: python tools/ncu_hot.py rep.ncu-rep [min_frac]"""
import csv, subprocess, sys, io


def main():
    rep = sys.argv[1]
    minf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, u, v = r[0], r[1], r[2]
    for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
              'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
              'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__grid_size'):
        if k in h:
            print(k, v[h.index(k)], u[h.index(k)])
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(src)))
    hh = r[1]; rows = [x for x in r[2:] if len(x) > 5]
    ia = hh.index('Instructions Executed'); isrc = hh.index('Source'); isamp = hh.index('# Samples')
    tot = sum(int(x[ia]) for x in rows if x[ia].isdigit())
    tots = sum(int(x[isamp]) for x in rows if x[isamp].isdigit())
    print('total inst', tot, 'samples', tots)
    runs = []
    for i, x in enumerate(rows):
        c = int(x[ia]) if x[ia].isdigit() else 0
        if c > tot * minf / 50:
            if runs and i - runs[-1][1] <= 4:
                runs[-1][1] = i; runs[-1][2] += c; runs[-1][3] += int(x[isamp])
            else:
                runs.append([i, i, c, int(x[isamp])])
    for a in runs:
        if a[2] > tot * minf:
            print('rows %5d-%5d  n=%4d  inst %.3f  samples %.3f  first: %s' % (
                a[0], a[1], a[1] - a[0] + 1, a[2] / tot, a[3] / max(tots, 1), rows[a[0]][isrc].strip()[:60]))


if __name__ == '__main__':
    main()
