"""box probe (not a test): H2D / D2H bandwidth of pinned buffers vs CPU affinity (NUMA)"""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tombo_b200 import _lib
os.system('nvidia-smi topo -m 2>/dev/null | head -8; lscpu | grep -E "NUMA|Socket"')
bus = os.popen('nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0').read().strip().lower()
bus = bus[4:] if len(bus) > 12 else bus
p = '/sys/bus/pci/devices/%s/' % bus
try:
    print('gpu', bus, 'numa_node', open(p + 'numa_node').read().strip(), 'local_cpulist', open(p + 'local_cpulist').read().strip())
    local = open(p + 'local_cpulist').read().strip()
except Exception as e:
    print('sysfs', e); local = None
rt = C.CDLL('libcudart.so.12')
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
ctx = _lib.Context(0)
dev = C.c_void_p()
N = 1 << 30
assert rt.cudaMalloc(C.byref(dev), C.c_size_t(N)) == 0


def parse(cl):
    out = set()
    for part in cl.split(','):
        a, _, b = part.partition('-')
        out.update(range(int(a), int(b or a) + 1))
    return out


def run(tag):
    pa = _lib.PinnedArray((N // 8,), np.float64)
    pa.array[:] = 1.0
    for d, (dst, src, kind) in (('H2D', (dev, C.c_void_p(pa.array.ctypes.data), 1)), ('D2H', (C.c_void_p(pa.array.ctypes.data), dev, 2))):
        rt.cudaMemcpy(dst, src, C.c_size_t(N), kind)
        t = time.time()
        for _ in range(3):
            rt.cudaMemcpy(dst, src, C.c_size_t(N), kind)
        dt = (time.time() - t) / 3
        print('%s %s %.1f GB/s' % (tag, d, N / dt / 1e9))
    pa.free() if hasattr(pa, 'free') else None


all_cpus = os.sched_getaffinity(0)
run('default affinity (%d cpus)' % len(all_cpus))
if local:
    loc = parse(local) & all_cpus
    if loc:
        os.sched_setaffinity(0, loc)
        run('gpu-local cpus (%d)' % len(loc))
        rem = all_cpus - loc
        if rem:
            os.sched_setaffinity(0, rem)
            run('remote cpus (%d)' % len(rem))
