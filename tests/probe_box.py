"""box probe (not a test): memory first-touch speed, pinned alloc speed, cores"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t = time.time(); a = np.ones(50_000_000); print('np.ones 400MB first touch %.3fs' % (time.time() - t))
t = time.time(); a[:] = 2; print('rewrite %.3fs' % (time.time() - t))
from tombo_b200 import _lib
t = time.time(); p = _lib.PinnedArray((125_000_000,), np.float64); print('pinned alloc 1GB %.3fs' % (time.time() - t))
t = time.time(); p.array[:] = 1; print('pinned first write %.3fs' % (time.time() - t))
print('cores', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())
os.system('free -g | head -2; nvidia-smi -L; lscpu | grep -E "Model name|Socket|^CPU\\(s\\)"')
