"""GPU: thousands of fresh-seed reads per BASELINE.json shape against the C oracle, bit for
bit -- so the rare paths (Theil-Sen bracket miss -> full-pair retry, skipped-base windows,
band-edge failures, rescue) are hit by real traffic, not one hand-built case each.
north_star: "bit-exact segmentation on the synthetic read set"."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALN_C1 = (4.2, 4.2, 200, 1500, 20.0, 40, 750, 2500, 250)
ALN_C3 = (4.2, 4.2, 400, 1500, 20.0, 40, 750, 2500, 250)
ALN_C5 = (4.2, 4.2, 1200, 1500, 20.0, 40, 750, 2500, 250)
ALN_NARROW = (4.2, 4.2, 100, 1500, 20.0, 40, 300, 2500, 100)


def _sweep(ctx, orc, RPcls, kind, aln, seg, nbases, n_reads, seed, int16=False, min_ok=0.9):
    from tombo_b200 import _lib, synthetic as syn
    kmer_ref, cpos = syn.make_kmer_ref(kind, 0)
    k = len(kmer_ref[0][0])
    means, sds = syn.kmer_table(kmer_ref)
    ctx.set_model(means, sds, k, cpos)
    rna = kind == 'RNA'
    rp, sp = RPcls(aln, seg, rna=rna), RPcls(aln, seg, rna=rna, save=True)
    raw, raw_off, seq, seq_off = syn.make_read_batch(kmer_ref, n_reads, nbases, seed, kind=kind,
                                                     int16=int16)
    pol = _lib.make_policy(kind, subsample_seed=seed)
    res = ctx.resquiggle_batch(raw, raw_off, seq, seq_off, rp, sp, pol)
    out = orc.run_batch(raw, raw_off, seq, seq_off, means, sds, k, rp, sp,
                        orc.policy(kind, subsample_seed=seed))
    bad = orc.compare_batch(res, out)
    assert not bad, bad[:10]
    ok = float((res['status'] == 0).mean())
    assert ok >= min_ok, ok
    return res


def test_sweep_config1_shape_2048_reads(ctx, orc, RPcls):
    """configs[1]: 4k-sample DNA reads, static band"""
    _sweep(ctx, orc, RPcls, 'DNA', ALN_C1, (5, 3, 1, 5), 444, 2048, 101)


def test_sweep_config1_int16_1024_reads(ctx, orc, RPcls):
    """the DAC dtype: int16 raw with the pinned tie rule"""
    _sweep(ctx, orc, RPcls, 'DNA', ALN_C1, (5, 3, 1, 5), 444, 1024, 102, int16=True)


def test_sweep_config3_shape_2048_mixed_reads(ctx, orc, RPcls):
    """configs[2]: 2k-20k samples, bandwidth 400 adaptive band + rescue"""
    nbs = np.random.RandomState(103).randint(222, 2223, 2048)
    res = _sweep(ctx, orc, RPcls, 'DNA', ALN_C3, (5, 3, 1, 5), nbs, 2048, 103)
    assert (res['flags'] & 4 == 0).mean() > 0.5     # most reads took the adaptive path


def test_sweep_narrow_band_rescues_and_failures_1024_reads(ctx, orc, RPcls):
    """a deliberately narrow band (100): many reads leave it, are rescued with the save
    bandwidth or fail with the reference's message -- statuses and rescued results equal"""
    nbs = np.random.RandomState(104).randint(300, 900, 1024)
    res = _sweep(ctx, orc, RPcls, 'DNA', ALN_NARROW, (5, 3, 1, 5), nbs, 1024, 104, min_ok=0.0)
    assert (res['flags'] & 2).any(), 'no read exercised the rescue path'


def test_sweep_config5_shape_64_long_reads(ctx, orc, RPcls):
    """configs[4]: 50k-sample reads, bandwidth 1200"""
    _sweep(ctx, orc, RPcls, 'DNA', ALN_C5, (5, 3, 1, 5), 5555, 64, 105)


def test_sweep_config4_shape_512_rna_reads(ctx, orc, RPcls):
    """configs[3]: direct RNA, 8k samples (t-test segmentation, stalls, event scaling)"""
    import golden_util as gu
    _sweep(ctx, orc, RPcls, 'RNA', gu.RNA_ALN, gu.RNA_SEG, 270, 512, 106)
