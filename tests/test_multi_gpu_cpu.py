"""CPU (gloo, world_size 2): the host side of the N > 1 path -- the shared NCCL-free work
queue hands every length bucket to exactly one rank, region counters sum over ranks, NUMA
binding parses sysfs."""
import os
import subprocess
import sys
import textwrap

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_length_buckets_group_similar_reads_and_cover_all():
    from tombo_b200 import multi_gpu as mg
    rs = np.random.RandomState(0)
    raw = rs.randint(2000, 20000, 5000)
    nb = raw // 9
    b = mg.length_buckets(raw, nb, target_samples=2_000_000)
    allidx = np.concatenate(b)
    assert np.array_equal(np.sort(allidx), np.arange(5000))
    assert all(raw[x].sum() <= 2_000_000 or len(x) == 1 for x in b)
    # longest first, and each bucket spans a narrow length range
    firsts = [raw[x].max() for x in b]
    assert firsts == sorted(firsts, reverse=True)
    assert np.median([raw[x].max() / raw[x].min() for x in b]) < 1.15


def test_numa_binding_reads_sysfs(tmp_path, monkeypatch):
    from tombo_b200 import multi_gpu as mg
    (tmp_path / 'devices/system/node/node1').mkdir(parents=True)
    (tmp_path / 'devices/system/node/node1/cpulist').write_text('0-1,3\n')
    monkeypatch.setattr(mg, 'gpu_numa_node', lambda device, sysfs='/sys': 1)
    before = os.sched_getaffinity(0)
    try:
        info = mg.bind_to_gpu_numa_node(0, str(tmp_path))
        assert info['node'] == 1 and info['bound']
        assert os.sched_getaffinity(0) <= {0, 1, 3}
    finally:
        os.sched_setaffinity(0, before)
    assert mg._parse_cpulist('0-3,8,10-11') == [0, 1, 2, 3, 8, 10, 11]


WORKER = textwrap.dedent('''
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %(repo)r)
    import torch.distributed as dist
    from tombo_b200 import multi_gpu as mg
    rank = int(os.environ['RANK'])
    dist.init_process_group('gloo', rank=rank, world_size=2)
    if rank == 0:
        q = mg.WorkQueue(%(name)r, 37, create=True)
    dist.barrier()
    if rank != 0:
        q = mg.WorkQueue(%(name)r, 37)
    mine = []
    while True:
        i = q.next()
        if i is None:
            break
        mine.append(i)
    counts = np.arange(12, dtype=np.int32) * (rank + 1)
    tot = mg.allreduce_region_counts(counts, dist)
    dist.barrier()
    q.close(unlink=(rank == 0))
    print(json.dumps({'rank': rank, 'mine': mine, 'tot': tot.tolist()}))
    dist.destroy_process_group()
''')


def test_work_queue_and_counter_allreduce_world_size_2(tmp_path):
    import json
    name = 'tb2_test_queue_%d' % os.getpid()
    script = tmp_path / 'w.py'
    script.write_text(WORKER % {'repo': REPO, 'name': name})
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29611', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    got = sorted(outs[0]['mine'] + outs[1]['mine'])
    assert got == list(range(37))                      # every bucket exactly once
    assert outs[0]['tot'] == outs[1]['tot'] == (np.arange(12) * 3).tolist()


def test_shared_counter_is_thread_safe(tmp_path):
    """two worker threads of one rank pull from the same queue object (bench.py --queue
    --queue-threads 2): every index is handed out exactly once"""
    import threading
    from tombo_b200 import multi_gpu as mg
    q = mg.WorkQueue('tb2_thread_q', 2000, create=True, directory=str(tmp_path))
    got = [[] for _ in range(4)]

    def worker(k):
        while True:
            i = q.next()
            if i is None:
                break
            got[k].append(i)
    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    q.close(unlink=True)
    flat = sorted(i for g in got for i in g)
    assert flat == list(range(2000))


def test_length_buckets_tail_is_cut_finer():
    import numpy as np
    from tombo_b200 import multi_gpu as mg
    rng = np.random.RandomState(2)
    lens = rng.randint(2000, 20001, size=5000)
    nb = lens // 9
    plain = mg.length_buckets(lens, nb, target_samples=2000000)
    guided = mg.length_buckets(lens, nb, target_samples=2000000, tail_fraction=0.25, tail_divisor=4)
    for b in (plain, guided):
        assert sorted(np.concatenate(b).tolist()) == list(range(5000))     # every read exactly once
        assert all(lens[x].sum() <= 2000000 for x in b)
    assert len(guided) > len(plain)
    sizes = [int(lens[x].sum()) for x in guided]
    assert max(sizes[-5:]) <= 500000 and max(sizes[:5]) > 1500000
    # longest first
    assert lens[guided[0]].min() >= lens[guided[-1]].max()
