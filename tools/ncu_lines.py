"""helper (not a test): attribute an ncu --import-source capture to CUDA source lines
by joining its SASS rows (in address order) with nvdisasm -g line markers of the
current build.  usage: python tools/ncu_lines.py rep.ncu-rep cubin mangled_kernel_name [top]"""
import csv, io, re, subprocess, sys, collections


def main():
    rep, cubin, fun = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
    # locate the function's text section
    start = None
    for i, l in enumerate(dis):
        if l.strip().startswith('.text.' + fun + ':') or l.strip() == fun + ':':
            start = i
    assert start is not None, 'function not found'
    lines = []          # per instruction (file, line)
    cur = ('?', 0)
    inl = None
    for l in dis[start + 1:]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s*\.section|\s*\.text\.', l) and lines:
            break
        if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+\S', l):
            lines.append(cur)
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(src)))
    hh = r[1]; rows = [x for x in r[2:] if len(x) > 5]
    ia = hh.index('Instructions Executed'); isamp = hh.index('# Samples')
    print('sass rows', len(rows), 'disasm instrs', len(lines))
    n = min(len(rows), len(lines))
    agg = collections.defaultdict(lambda: [0, 0])
    tot = tots = 0
    for i in range(n):
        c = int(rows[i][ia]) if rows[i][ia].isdigit() else 0
        sp = int(rows[i][isamp]) if rows[i][isamp].isdigit() else 0
        agg[lines[i]][0] += c; agg[lines[i]][1] += sp
        tot += c; tots += sp
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%-22s %5d  inst %.3f  samples %.3f' % (k[0], k[1], v[0] / tot, v[1] / max(tots, 1)))


if __name__ == '__main__':
    main()
