"""CPU (gloo, world_size 2): the multi-rank plumbing of bench.py -- reads shard
across ranks with no data-path collective; only the barrier and the max-over-ranks
time / sum-over-ranks read count use torch.distributed."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_reduce_and_barrier():
    env = dict(os.environ)
    env['MASTER_ADDR'] = '127.0.0.1'
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
         '--master-addr', '127.0.0.1', '--master-port', '29531',
         os.path.join(REPO, 'tests', 'dist_helper.py')],
        capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith('{')][-1]
    d = json.loads(line)
    assert d == {'world': 2, 'max_t': 1.0, 'sum_n': 300.0}


def test_reference_arm_runs_on_cpu():
    """bench.py --impl reference: the reference's own CPU path on a tiny sample"""
    out = subprocess.run(
        [sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1',
         '--warmup', '1', '--cpu-sample', '8'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert d['impl'] == 'reference' and d['value'] > 0 and d['unit'] == 'reads/s'
    assert d['cpu_baseline']['kind'] in ('reference', 'port')
    assert d['e2e']['h2d_bytes_per_step'] == 0
