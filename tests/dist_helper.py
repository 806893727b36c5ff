"""helper for test_multi_rank_cpu.py: exercises bench.py's rank plumbing on gloo"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, world, local, dist = bench.dist_setup(world_size := int(os.environ.get('WORLD_SIZE', '1')))
bench.barrier(dist)
# every rank "processes" (rank+1)*100 reads in (rank+1)*0.5 s
t = bench.reduce_max(dist, (rank + 1) * 0.5)
n = bench.reduce_sum(dist, (rank + 1) * 100.0)
bench.barrier(dist)
if rank == 0:
    print(json.dumps({'world': world, 'max_t': t, 'sum_n': n}))
if dist is not None:
    dist.destroy_process_group()
