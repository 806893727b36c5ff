"""usage: python tools/make_traffic_json.py <ncu-rep> <reads in the captured launch> [launch index]
Writes profiles/k_align_traffic.json: DRAM bytes per read of the captured k_align<1> launch
(dram__bytes_read.sum + dram__bytes_write.sum from `ncu --set full`), with a hash of the kernel
source so bench.py drops the number once the kernel changes (roofline.traffic -> null)."""
import csv, hashlib, io, json, os, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, reads = sys.argv[1], int(sys.argv[2])
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, u = r[0], r[1]
    rows = [x for x in r[2:] if 'k_align<1>' in x[h.index('Kernel Name')]]
    row = rows[which]
    sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot = 0.0
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        tot += float(row[h.index(k)]) * sc[u[h.index(k)]]
    hs = hashlib.sha256()
    for f in ('dp_row.cuh', 'dp_align.cuh', 'dp_align_kernel.cuh', 'dp_row2.cuh', 'common.cuh'):
        p = os.path.join(REPO, 'tombo_b200', 'csrc', f)
        if os.path.exists(p):
            hs.update(open(p, 'rb').read())
    out = {'dram_bytes_per_read': tot / reads, 'reads_in_launch': reads,
           'dram_bytes_launch': tot, 'duration_ms': row[h.index('gpu__time_duration.sum')],
           'capture': 'ncu --set full --clock-control none, %s' % os.path.basename(rep),
           'source_sha16': hs.hexdigest()[:16]}
    json.dump(out, open(os.path.join(REPO, 'profiles', 'k_align_traffic.json'), 'w'), indent=1)
    print(out)


if __name__ == '__main__':
    main()
