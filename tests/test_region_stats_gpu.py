"""GPU: SURVEY.md 8(f)-1 (per-position aggregation) and 8(f)-2 (de novo / sample-compare
per-read tests) against goldens recorded from the unmodified reference
(tests/golden/make_stats_golden.py).  Counts, positions and fractions are exact; p-values
go through device erfc / log / exp and are pinned within rtol 1e-7."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
RTOL = 1e-7


def _close(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=RTOL, atol=0)


@pytest.fixture(scope='module')
def g():
    return gu.load('region_stats')


def _model():
    from tombo_b200 import synthetic as syn, tombo_stats as ts
    kmer_ref, cpos = syn.make_kmer_ref('DNA', 0)
    return ts.TomboModel(kmer_ref=kmer_ref, central_pos=cpos), kmer_ref, cpos


def test_window_fisher_matches_scipy_chi2(ctx, g):
    from tombo_b200 import tombo_stats as ts
    for lag in (1, 2, 4):
        _close(ts.calc_window_fishers_method(g['fw_p'], lag), g['fw_lag%d' % lag])


class _Reg(object):
    def __init__(self, start, end):
        self.start, self.end, self.chrm, self.strand = int(start), int(end), 'chr', '+'


def _seams(monkeypatch, means, seq):
    from tombo_b200 import tombo_helper as th
    bases = np.array(list(seq), dtype='S1')
    monkeypatch.setattr(th, 'get_multiple_slots_read_centric',
                        lambda r, slots, grp=None: (means.copy(), bases) if len(slots) == 2
                        else (means.copy(),))
    monkeypatch.setattr(th, 'get_raw_read_slot', lambda r: (_ for _ in ()).throw(KeyError()))


def test_de_novo_and_sample_compare_mirrors(ctx, g, monkeypatch):
    from tombo_b200 import tombo_helper as th, tombo_stats as ts
    std_ref, _, _ = _model()
    reg = _Reg(*g['reg'])
    n_checked = 0
    for i in range(int(g['dn_n'])):
        a, b = int(g['dn_off'][i]), int(g['dn_off'][i + 1])
        means, seq = g['dn_means'][a:b], str(g['dn_seq'][i])
        r_data = th.readData(start=int(g['dn_start'][i]), end=int(g['dn_start'][i]) + (b - a),
                             filtered=False, read_start_rel_to_raw=0,
                             strand=str(g['dn_strand'][i]), fn='x', corr_group='g', rna=False)
        _seams(monkeypatch, means, seq)
        for fm in (0, 1, 2):
            for tag, rg in (('whole', None), ('reg', reg)):
                key = 'dn_r%d_fm%d_%s' % (i, fm, tag)
                if str(g[key + '_err']):
                    with pytest.raises(th.TomboError):
                        ts.compute_de_novo_read_stats(r_data, std_ref, fm, rg)
                    continue
                pv, ps, _ = ts.compute_de_novo_read_stats(r_data, std_ref, fm, rg)
                assert np.array_equal(ps['de_novo'], g[key + '_pos'])
                _close(pv['de_novo'], g[key + '_p'])
                n_checked += 1
            key = 'sc_r%d_fm%d' % (i, fm)
            if str(g[key + '_err']):
                continue
            pv, ps, _ = ts.compute_sample_compare_read_stats(
                r_data, g['sc_cm_fm%d' % fm], g['sc_cs_fm%d' % fm], fm, reg)
            assert np.array_equal(ps['sample_compare'], g[key + '_pos'])
            _close(pv['sample_compare'], g[key + '_p'])
            n_checked += 1
    assert n_checked >= 40


def test_de_novo_batch_kernel_on_plus_strand_reads(ctx, g):
    """tb2_de_novo_read_stats_batch: levels looked up on the device, whole reads"""
    from tombo_b200 import synthetic as syn
    _, kmer_ref, cpos = _model()
    means, sds = syn.kmer_table(kmer_ref)
    ctx.set_model(means, sds, 6, cpos)
    idx = [i for i in range(int(g['dn_n'])) if str(g['dn_strand'][i]) == '+']
    # the C ABI takes B + K - 1 base codes and B means per read; the stored read has B bases
    # and B means, so the first / last (K - 1) means positions only provide k-mer context:
    # pass the stored bases as the sequence and the matching inner means
    for fm in (0, 1, 2):
        nm, mo, sq, so, st = [], [0], [], [0], []
        for i in idx:
            a, b = int(g['dn_off'][i]), int(g['dn_off'][i + 1])
            seq = syn.seq_to_codes(str(g['dn_seq'][i]))
            nm.append(g['dn_means'][a:b]); mo.append(mo[-1] + (b - a))
            # library layout: seq holds cpos extra codes before and K-1-cpos after the means
            sq.append(np.concatenate([np.zeros(cpos, np.uint8), seq, np.zeros(6 - cpos - 1, np.uint8)]))
            so.append(so[-1] + sq[-1].shape[0]); st.append(int(g['dn_start'][i]))
        pv, pos, off = ctx.de_novo_read_stats_batch(np.concatenate(nm), mo, np.concatenate(sq), so,
                                                    st, fm)
        for k, i in enumerate(idx):
            key = 'dn_r%d_fm%d_whole' % (i, fm)
            assert np.array_equal(pos[off[k]:off[k + 1]], g[key + '_pos'])
            _close(pv[off[k]:off[k + 1]], g[key + '_p'])


@pytest.mark.parametrize('name,stat_type', [('alt_lower', 'model_compare'),
                                            ('alt_abs', 'model_compare'), ('denovo', 'de_novo')])
def test_collate_reg_stats_matches_reference(ctx, g, name, stat_type):
    from tombo_b200 import tombo_stats as ts
    off = g['rg_off']
    stats = [g['rg_stats'][off[i]:off[i + 1]] for i in range(off.shape[0] - 1)]
    locs = [g['rg_locs'][off[i]:off[i + 1]] for i in range(off.shape[0] - 1)]
    thr, lower = g['rg_%s_params' % name]
    r = ts.collate_reg_stats(stats, locs, None, None, _Reg(5000, 6000), float(thr),
                             None if np.isnan(lower) else float(lower), stat_type, stat_type, None)
    assert np.array_equal(r.reg_poss, g['rg_%s_pos' % name])
    assert np.array_equal(r.reg_cov, g['rg_%s_cov' % name])
    assert np.array_equal(r.valid_cov, g['rg_%s_valid' % name])
    assert np.array_equal(r.reg_frac_standard_base, g['rg_%s_frac' % name], equal_nan=True)
    # dampened fraction: the device's fused evaluation and the host mirror
    ctx2 = ctx
    keep = ~np.isnan(g['rg_stats'])
    ctx2.region_stats_begin(5000, 1000)
    ctx2.region_stats_add(g['rg_stats'], g['rg_locs'], float(thr),
                          None if np.isnan(lower) else float(lower),
                          0 if stat_type == 'model_compare' else 1)
    fin = ctx2.region_stats_finalize(unmod_count=2, mod_count=0)
    assert np.array_equal(fin['pos'], g['rg_%s_pos' % name])
    assert np.array_equal(fin['damp_frac'], g['rg_%s_damp' % name], equal_nan=True)
    with np.errstate(all='ignore'):
        host = ts.calc_damp_fraction({'unmod': 2, 'mod': 0}, r.reg_frac_standard_base,
                                     r.valid_cov.astype(np.float64))
    assert np.array_equal(host, g['rg_%s_damp' % name], equal_nan=True)
    assert keep.sum() == fin['cov'].sum()


def test_region_counters_add_across_shards(ctx, g):
    """reads of one region split over two shards: summed counters == one pass (what an
    all-reduce over GPUs does, tombo_b200.multi_gpu.allreduce_region_counts)"""
    n = g['rg_stats'].shape[0]
    half = n // 2
    parts = []
    for sl in (slice(0, half), slice(half, n)):
        ctx.region_stats_begin(5000, 1000)
        ctx.region_stats_add(g['rg_stats'][sl], g['rg_locs'][sl], 2.5, -1.5, 0)
        parts.append(ctx.region_counts_get())
    ctx.region_stats_begin(5000, 1000)
    ctx.region_counts_set(parts[0] + parts[1])
    fin = ctx.region_stats_finalize()
    assert np.array_equal(fin['pos'], g['rg_alt_lower_pos'])
    assert np.array_equal(fin['valid_cov'], g['rg_alt_lower_valid'])
    assert np.array_equal(fin['frac'], g['rg_alt_lower_frac'], equal_nan=True)


def test_resident_llr_and_fused_region_stats(ctx, RPcls):
    """resquiggle -> LLR -> per-position counts without leaving HBM == the host-array route"""
    from tombo_b200 import _lib, synthetic as syn
    kmer_ref, cpos = syn.make_kmer_ref('DNA', 0)
    means, sds = syn.kmer_table(kmer_ref)
    alt = np.full((4 ** 6, 6), np.nan)
    code = {'A': 0, 'C': 1, 'G': 2, 'T': 3}
    for km, pos, m, sd in syn.make_alt_kmer_ref(kmer_ref, 'C', seed=1):
        idx = 0
        for b in km:
            idx = idx * 4 + code[b]
        alt[idx, pos] = m
    ctx.set_model(means, sds, 6, cpos)
    ctx.set_alt_model(alt, 6)
    aln = (4.2, 4.2, 200, 1500, 20.0, 40, 750, 2500, 250)
    rp, sp = RPcls(aln), RPcls(aln, save=True)
    n = 200
    raw, raw_off, seq, seq_off = syn.make_read_batch(kmer_ref, n, 300, 31)
    # one hopeless read: its sites must not reach the counters
    raw[raw_off[7]:raw_off[8]] = 480.0
    pol = _lib.make_policy('DNA')
    ctx.batch_upload(raw, raw_off, seq, seq_off, rp, pol)
    ctx.batch_compute(rp, sp, pol)
    res = ctx.batch_download()
    assert res['status'][7] != 0 and (res['status'] == 0).sum() >= n - 5
    read_start = (np.arange(n, dtype=np.int64) * 37) % 900 + 10000
    tot = ctx.batch_alt_llr(read_start, 1)
    llr, pos, site_off = ctx.batch_llr_download()
    # host-array route on the successful reads
    ok = res['status'] == 0
    nm = res['norm_mean'].copy()
    llr_h, pos_h, so_h = ctx.alt_model_llr_batch(nm, res['base_off'], seq, seq_off, read_start, 1)
    keep = np.repeat(ok, np.diff(so_h))
    assert tot == keep.sum() and site_off[8] == site_off[7]
    assert np.array_equal(pos, pos_h[keep]) and np.array_equal(llr, llr_h[keep])
    ctx.region_stats_begin(10000, 2000)
    ctx.region_stats_add_batch_llr(2.5, -1.5, 0)
    fused = ctx.region_stats_finalize(2, 0)
    ctx.region_stats_begin(10000, 2000)
    ctx.region_stats_add(llr, pos, 2.5, -1.5, 0)
    host = ctx.region_stats_finalize(2, 0)
    for k in fused:
        assert np.array_equal(fused[k], host[k], equal_nan=True), k
    # numpy restatement of apply_per_read_thresh on the same pairs
    up = np.unique(pos)
    assert np.array_equal(fused['pos'], up)
    for q in (0, len(up) // 2, len(up) - 1):
        s = llr[pos == up[q]]
        v = s[(s <= -1.5) | (s >= 2.5)]
        assert fused['cov'][q] == s.shape[0] and fused['valid_cov'][q] == v.shape[0]
        if v.shape[0]:
            assert fused['frac'][q] == (v >= 2.5).sum() / v.shape[0]
