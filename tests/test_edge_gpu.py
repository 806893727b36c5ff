"""GPU edge cases and size-independent properties of the batched hot path (through the
C ABI): empty / degenerate / ragged inputs, batch-order invariance and structural
invariants of the results at a BASELINE-sized read shape."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(ctx, raw, raw_off, codes, seq_off, RPcls, aln, **kw):
    from tombo_b200 import _lib
    rp, sp = RPcls(aln), RPcls(aln, save=True)
    pol = _lib.make_policy('DNA')
    res = ctx.resquiggle_batch(raw, raw_off, codes, seq_off, rp, sp, pol, **kw)
    return {k: np.array(v, copy=True) for k, v in res.items()}


def test_empty_batch(ctx, dna_model, RPcls):
    import bench
    from tombo_b200 import synthetic as syn
    kmer_ref, cpos = dna_model
    ctx.set_model(*syn.kmer_table(kmer_ref), 6, cpos)
    res = _run(ctx, np.zeros(0), np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.uint8),
               np.zeros(1, dtype=np.int64), RPcls, bench.ALN_DNA)
    assert res['status'].shape == (0,) and res['segs'].shape == (0,)


def test_degenerate_reads_fail_alone(ctx, orc, dna_model, RPcls):
    """a batch holding reads that cannot be processed (no signal, a sequence shorter
    than the k-mer, a one-base mapping, a handful of samples) reports them per read, with
    the oracle's status, and leaves the healthy reads' results untouched"""
    import bench
    from tombo_b200 import synthetic as syn
    kmer_ref, cpos = dna_model
    means, sds = syn.kmer_table(kmer_ref)
    ctx.set_model(means, sds, 6, cpos)
    good = syn.make_reads(kmer_ref, cpos, 3, 444, seed0=4100)
    raws = [good[0].raw, np.zeros(0), good[1].raw, good[1].raw[:30], good[2].raw, good[2].raw[:900]]
    seqs = [good[0].genome_seq, good[1].genome_seq, good[1].genome_seq[:3], good[1].genome_seq,
            good[2].genome_seq, good[2].genome_seq[:6]]
    raw = np.concatenate(raws)
    raw_off = np.concatenate([[0], np.cumsum([r.shape[0] for r in raws])]).astype(np.int64)
    codes = [syn.seq_to_codes(q) for q in seqs]
    seq_off = np.concatenate([[0], np.cumsum([c.shape[0] for c in codes])]).astype(np.int64)
    res = _run(ctx, raw, raw_off, np.concatenate(codes), seq_off, RPcls, bench.ALN_DNA)
    solo = [_run(ctx, g.raw, np.array([0, g.raw.shape[0]]), syn.seq_to_codes(g.genome_seq),
                 np.array([0, len(g.genome_seq)]), RPcls, bench.ALN_DNA) for g in good]
    for i, k in ((0, 0), (2, 4)):
        a, b = res['seg_off'][k], res['seg_off'][k + 1]
        assert res['status'][k] == 0
        assert np.array_equal(res['segs'][a:b], solo[i]['segs'])
        assert res['sig_match_score'][k] == solo[i]['sig_match_score'][0]
    for k in (1, 2, 3):
        assert res['status'][k] != 0, k
    # the one-base mapping: whatever the reference does, the oracle does the same
    rp, sp = RPcls(bench.ALN_DNA), RPcls(bench.ALN_DNA, save=True)
    c5 = syn.seq_to_codes(seqs[5]).astype(np.int64)
    kidx = int(np.polyval(c5, 4))
    o = orc.run_read(np.asarray(raws[5], dtype=np.float64), means[[kidx]], sds[[kidx]], rp, sp,
                     orc.policy('DNA'), read_index=5)
    assert res['status'][5] == o['status']
    if o['status'] == 0:
        a, b = res['seg_off'][5], res['seg_off'][5 + 1]
        assert np.array_equal(res['segs'][a:b], o['segs'])


def test_order_invariance_and_invariants_at_baseline_shape(ctx, dna_model, RPcls):
    """configs[1] read shape, a few thousand reads: every read's result is independent
    of its position in the batch, segs start at 0, increase strictly and end inside the
    raw signal, per-base means are finite"""
    import bench
    from tombo_b200 import synthetic as syn
    kmer_ref, cpos = dna_model
    ctx.set_model(*syn.kmer_table(kmer_ref), 6, cpos)
    n = 3000
    raw, raw_off, codes, seq_off = syn.make_read_batch(kmer_ref, n, bench.N_BASES, 31)
    a = _run(ctx, raw, raw_off, codes, seq_off, RPcls, bench.ALN_DNA)
    perm = np.random.RandomState(3).permutation(n)
    rl, sl = np.diff(raw_off), np.diff(seq_off)
    raw_p = np.concatenate([raw[raw_off[i]:raw_off[i + 1]] for i in perm])
    codes_p = np.concatenate([codes[seq_off[i]:seq_off[i + 1]] for i in perm])
    ro_p = np.concatenate([[0], np.cumsum(rl[perm])]).astype(np.int64)
    so_p = np.concatenate([[0], np.cumsum(sl[perm])]).astype(np.int64)
    b = _run(ctx, raw_p, ro_p, codes_p, so_p, RPcls, bench.ALN_DNA)
    assert (a['status'] == 0).mean() > 0.95
    for q, i in enumerate(perm):
        assert a['status'][i] == b['status'][q]
        if a['status'][i] != 0:
            continue
        sa = a['segs'][a['seg_off'][i]:a['seg_off'][i + 1]]
        sb = b['segs'][b['seg_off'][q]:b['seg_off'][q + 1]]
        assert np.array_equal(sa, sb)
        assert a['read_start_rel_to_raw'][i] == b['read_start_rel_to_raw'][q]
        assert a['sig_match_score'][i] == b['sig_match_score'][q]
        assert np.array_equal(a['scale_values'][i, :4], b['scale_values'][q, :4])
        assert sa[0] == 0 and np.all(np.diff(sa) > 0)
        assert a['read_start_rel_to_raw'][i] + sa[-1] <= rl[i]
        nm = a['norm_mean'][a['base_off'][i]:a['base_off'][i + 1]]
        assert np.all(np.isfinite(nm))
