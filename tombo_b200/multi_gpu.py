"""Host-side plumbing for one-process-per-GPU runs.

Reads are independent, so the data path needs no collective: every rank owns a
:class:`tombo_b200._lib.Context` and pulls length-bucketed batches from one shared,
NCCL-free work queue (the reference's equivalent is the multiprocessing queue its
``_resquiggle_worker`` processes read from, resquiggle.py:1868-1928).  The only exchange
the path ever needs is the optional sum of per-position counters when the reads of one
region are spread over ranks (:func:`allreduce_region_counts`, 8(f)-1).

Nothing here touches the device; it is covered by CPU tests (gloo, world_size 2)."""
import fcntl
import mmap
import os
import struct
import threading

import numpy as np


# ---------------------------------------------------------------------------
# NUMA: bind a rank to the node its GPU hangs off before pinned buffers are allocated
# ---------------------------------------------------------------------------
def _parse_cpulist(txt):
    cpus = []
    for part in txt.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(device, sysfs='/sys'):
    """NUMA node of CUDA device ``device`` (honours CUDA_VISIBLE_DEVICES), or None"""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        h = None
        if vis:
            ents = [e.strip() for e in vis.split(',')]
            ent = ents[device] if device < len(ents) else ''
            try:
                h = nv.nvmlDeviceGetHandleByIndex(int(ent))
            except Exception:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(ent.encode())
                except Exception:
                    h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(device)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
    except Exception:
        return None
    dom, rest = bus.lower().split(':', 1)
    path = os.path.join(sysfs, 'bus/pci/devices', '%s:%s' % (dom[-4:], rest), 'numa_node')
    try:
        node = int(open(path).read().strip())
    except Exception:
        return None
    return node if node >= 0 else None


def bind_to_gpu_numa_node(device, sysfs='/sys'):
    """Restrict this process to the CPUs of the GPU's NUMA node (first-touch then places
    pinned staging buffers on that node: H2D / D2H cross no socket link).  Returns a dict
    describing what was done; never raises."""
    info = {'node': None, 'cpus': None, 'bound': False}
    node = gpu_numa_node(device, sysfs)
    if node is None:
        return info
    info['node'] = node
    try:
        cpus = _parse_cpulist(open(os.path.join(sysfs, 'devices/system/node/node%d/cpulist' % node)).read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info['cpus'] = len(allowed)
            info['bound'] = True
    except Exception:
        pass
    return info


# ---------------------------------------------------------------------------
# shared work queue: one integer in /dev/shm, fetch-and-add under flock
# ---------------------------------------------------------------------------
class SharedCounter(object):
    """A process-shared counter (fetch_add) backed by a file; ranks of one node open the
    same ``name``.  No NCCL, no server process."""

    def __init__(self, name, create=False, directory=None):
        directory = directory or ('/dev/shm' if os.path.isdir('/dev/shm') else '/tmp')
        self.path = os.path.join(directory, name)
        flags = os.O_RDWR | (os.O_CREAT if create else 0)
        self.fd = os.open(self.path, flags, 0o600)
        if create:
            fcntl.flock(self.fd, fcntl.LOCK_EX)
            os.ftruncate(self.fd, 8)
            os.pwrite(self.fd, struct.pack('<q', 0), 0)
            fcntl.flock(self.fd, fcntl.LOCK_UN)
        self.mm = mmap.mmap(self.fd, 8)
        self._tlock = threading.Lock()

    def fetch_add(self, n=1):
        # flock excludes other processes; threads of this process share the open file
        # description (and so the lock), hence the thread lock around it
        with self._tlock:
            fcntl.flock(self.fd, fcntl.LOCK_EX)
            try:
                v = struct.unpack_from('<q', self.mm, 0)[0]
                struct.pack_into('<q', self.mm, 0, v + n)
            finally:
                fcntl.flock(self.fd, fcntl.LOCK_UN)
        return v

    def reset(self, v=0):
        fcntl.flock(self.fd, fcntl.LOCK_EX)
        struct.pack_into('<q', self.mm, 0, v)
        fcntl.flock(self.fd, fcntl.LOCK_UN)

    def value(self):
        return struct.unpack_from('<q', self.mm, 0)[0]

    def close(self, unlink=False):
        try:
            self.mm.close()
            os.close(self.fd)
        finally:
            if unlink:
                try:
                    os.unlink(self.path)
                except OSError:
                    pass


def length_buckets(raw_lens, n_bases, target_samples, max_reads=65536, tail_fraction=0.0,
                   tail_divisor=4):
    """Cut a read set into batches of similar shape.  The DP cost of a read is set by its
    events E ~ raw length, its bases B and the bandwidth; with one bandwidth per run, sorting
    by (raw length, bases) groups reads of equal (E, B, bw).  Buckets are runs of the sorted
    order holding at most ``target_samples`` raw samples / ``max_reads`` reads, returned
    longest first (the expensive buckets are pulled first, the tail is cheap).  The last
    ``tail_fraction`` of the samples is cut ``tail_divisor`` times finer (guided scheduling:
    the ranks then finish within a small bucket of each other)."""
    raw_lens = np.asarray(raw_lens, dtype=np.int64)
    order = np.lexsort((np.asarray(n_bases, dtype=np.int64), raw_lens))[::-1]
    total = int(raw_lens.sum())
    tail_from = total - int(total * float(tail_fraction))
    small = max(1, int(target_samples) // max(1, int(tail_divisor)))
    buckets, cur, acc, seen = [], [], 0, 0
    for i in order:
        n = int(raw_lens[i])
        cap = small if seen >= tail_from else target_samples
        if cur and (acc + n > cap or len(cur) >= max_reads):
            buckets.append(np.array(cur, dtype=np.int64))
            cur, acc = [], 0
        cur.append(int(i))
        acc += n
        seen += n
    if cur:
        buckets.append(np.array(cur, dtype=np.int64))
    return buckets


class WorkQueue(object):
    """Ranks call :meth:`next` until it returns None; every bucket index is handed out
    exactly once across all processes sharing ``name``."""

    def __init__(self, name, n_items, create=False, directory=None):
        self.n_items = int(n_items)
        self.counter = SharedCounter(name, create=create, directory=directory)

    def next(self):
        i = self.counter.fetch_add(1)
        return i if i < self.n_items else None

    def close(self, unlink=False):
        self.counter.close(unlink)


# ---------------------------------------------------------------------------
# the one exchange of the path: summing per-position counters over ranks
# ---------------------------------------------------------------------------
def allreduce_region_counts(counts, dist=None, device=None):
    """Element-wise sum of the int32 counter array of tb2_region_counts_get over all ranks
    (torch.distributed: NCCL when ``device`` is a CUDA device index -- the counters travel
    over NVLink -- else the group's CPU backend).  Returns a numpy int32 array to hand to
    tb2_region_counts_set."""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return counts
    import torch
    t = torch.from_numpy(counts.copy())
    if device is not None and dist.get_backend() == 'nccl':
        t = t.cuda(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
