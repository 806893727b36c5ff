"""ad-hoc GPU debug script (not a test)"""
import sys, os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'oracle')); sys.path.insert(0, os.path.join(REPO, 'tests'))
import ctypes as C
import numpy as np, oracle as orc
from tombo_b200 import synthetic as syn, _lib
from conftest import RP
from test_pipeline_gpu import _levels
kmer_ref, cpos = syn.make_kmer_ref('DNA', 0)
means, sds = syn.kmer_table(kmer_ref)
aln = (4.2, 4.2, 120, 1500, 20.0, 40, 300, 2500, 100)
rp, sp = RP(aln), RP(aln, save=True)
r = syn.make_read(kmer_ref, cpos, 600, 7100)
rm, rsd = _levels(r.genome_seq, means, sds, 6)
pol = orc.policy('DNA')
o = orc.resquiggle_read(r.raw, rm, rsd, rp, pol, key=0)
sv2 = o['scale_values']
st, norm2, svb = orc.normalize_raw_signal(r.raw, scale_values=sv2)
ne = max(len(r.raw)//5, int(600*1.1))
st, cpts2 = orc.valid_cpts_w_cap(norm2, 3, 5, ne)
em2 = orc.new_means(norm2, cpts2)
nb = 600
es0 = np.zeros(nb, dtype=np.int64); tb0 = np.zeros(nb+1, dtype=np.int64)
orc.lib().orc_set_debug_buffers(es0.ctypes.data_as(C.POINTER(C.c_int64)), tb0.ctypes.data_as(C.POINTER(C.c_int64)))
st0, segs0, rs0, dbg0, epb0 = orc.find_adaptive_base_assignment(cpts2, em2, rp, rm, rsd)
ctx = _lib.Context(0)
st1, segs1, rs1, dbg1 = ctx.find_adaptive_base_assignment(cpts2, em2, rp, rm, rsd)
es1 = np.zeros(nb, dtype=np.int64); tb1 = np.zeros(nb+1, dtype=np.int64)
fn = ctx.lib.tb2_debug_last_assignment; fn.restype = C.c_int
fn(ctx.handle, C.c_int64(nb), es1.ctypes.data_as(C.POINTER(C.c_int64)), tb1.ctypes.data_as(C.POINTER(C.c_int64)))
print('status', st0, st1, 'dbg', dbg0, dbg1, 'epb', epb0)
print('starts equal', np.array_equal(es0, es1), 'first diff', np.where(es0 != es1)[0][:10])
print('es0', es0[:12]); print('es1', es1[:12])
print('tb0', tb0[:12]); print('tb1', tb1[:12])
print('tb diff idx', np.where(tb0 != tb1)[0][:20])
print('segs diff idx', np.where(segs0 != segs1)[0][:20])
