"""ad-hoc GPU debug script (not a test): lane-chunk engine counters on mixed reads"""
import ctypes as C
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import bench
from tombo_b200 import _lib, synthetic as syn
kmer_ref, cpos, raw, raw_off, seq, seq_off = bench.make_workload(3000, 3, mixed=True)
means, sds = syn.kmer_table(kmer_ref)
ctx = _lib.Context(0)
ctx.set_model(means, sds, 6, cpos)
rp, sp = bench.RP(bench.ALN_MIXED), bench.RP(bench.ALN_MIXED, save=True)
pol = _lib.make_policy('DNA')
fn = ctx.lib.tb2_debug_dp_counters
fn.restype = C.c_int
out = (C.c_uint64 * 8)()
fn(ctx.handle, out, C.c_int(1))
res = ctx.resquiggle_batch(raw, raw_off, seq, seq_off, rp, sp, pol)
fn(ctx.handle, out, C.c_int(1))
rows = max(1, out[0])
print('rows %d  rounds/row %.2f  rewalked cells/row %.1f  rows>2rounds %.3f' % (out[0], out[1] / rows, out[2] / rows, out[3] / rows))
print('status ok', int((res['status'] == 0).sum()), 'of', len(res['status']), 'flags static', int(((res['flags'] & 4) != 0).sum()))
print('timing', ctx.last_timing())
