"""CPU: the banded-DP device source (tombo_b200/csrc/dp_*.cuh, the files nvcc compiles for
sm_100a) executed on the host emulation of tests/emul and compared bit for bit with the
oracle's find_adaptive_base_assignment -- static band, start search, masked start,
adaptive rows, traceback.  An algorithm check of the kernel source; the GPU parity tests
remain the proof for the compiled kernel."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emul'))


def _events(orc, read, means, sds, rp, k=6):
    """first-call inputs of the assignment: changepoints + event means (oracle stages)"""
    from tombo_b200 import synthetic as syn
    codes = syn.seq_to_codes(read.genome_seq).astype(np.int64)
    nb = codes.shape[0] - k + 1
    kidx = np.zeros(nb, dtype=np.int64)
    for j in range(k):
        kidx = kidx * 4 + codes[j:j + nb]
    rm, rs = means[kidx], sds[kidx]
    st, norm, sv = orc.normalize_raw_signal(np.asarray(read.raw, dtype=np.float64), 5.0)
    assert st == 0
    n_ev = max(norm.shape[0] // rp.mean_obs_per_event, int(nb * 1.1))
    st, cp = orc.valid_cpts_w_cap(norm, rp.min_obs_per_base, rp.running_stat_width, n_ev)
    assert st == 0
    cp = np.sort(cp)
    em = orc.new_means(norm, cp)
    return cp.astype(np.int32), em, rm, rs


CASES = [
    # name, aln params, n_bases list, klass
    ('static4k', (4.2, 4.2, 200, 1500, 20.0, 40, 750, 2500, 250), [444, 444, 300, 61], 1),
    ('adapt4k', (4.2, 4.2, 200, 1500, 20.0, 40, 300, 2500, 100), [444, 444, 380], 2),
    ('adapt_bw400', (4.2, 4.2, 400, 1500, 20.0, 40, 750, 2500, 250), [900, 700], 2),
    ('adapt_bw120_narrow', (4.2, 4.2, 120, 1500, 20.0, 40, 300, 2500, 100), [500, 444], 0),
]


@pytest.mark.parametrize('name,aln,nbs,klass', CASES)
def test_device_source_matches_oracle(orc, dna_model, RPcls, name, aln, nbs, klass):
    import emul
    from tombo_b200 import synthetic as syn
    kmer_ref, cpos = dna_model
    means, sds = syn.kmer_table(kmer_ref)
    rp = RPcls(aln)
    reads = syn.make_reads(kmer_ref, cpos, len(nbs), nbs, seed0=4100)
    ins = [_events(orc, r, means, sds, rp) for r in reads]
    res = emul.align_batch(ins, rp, klass=klass)
    for (cp, em, rm, rs), o in zip(ins, res):
        st, segs, rsrtr, dbg, epb = orc.find_adaptive_base_assignment(cp, em, rp, rm, rs)
        assert o['status'] == st
        if st == 0:
            assert np.array_equal(o['segs'], segs)
            assert o['rsrtr'] == rsrtr
            assert o['dbg'][0] == dbg[0]
