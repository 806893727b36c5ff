#!/usr/bin/env python
"""Goldens for SURVEY.md 8(f)-1 / 8(f)-2 from the UNMODIFIED reference (oracle/_ref):
collate_reg_stats / apply_per_read_thresh / calc_damp_fraction on per-read LLRs, and
compute_de_novo_read_stats / compute_sample_compare_read_stats / calc_window_fishers_method
(FAST5 reads replaced at the reference's own seams, as in make_golden.py).

    python oracle/build_ref.py && python tests/golden/make_stats_golden.py
"""
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))

import ref_harness as rh  # noqa: E402
from tombo_b200 import synthetic as syn  # noqa: E402


class Reg(object):
    def __init__(self, start, end, chrm='chr', strand='+'):
        self.start, self.end, self.chrm, self.strand = start, end, chrm, strand


def main():
    m = rh.load_reference()
    th, ts = m['th'], m['ts']
    rs = np.random.RandomState(77)
    out = {}
    kmer_ref, cpos = syn.make_kmer_ref('DNA', 0)
    std_ref, _ = rh.make_models(kmer_ref, cpos)
    K = 6

    # ---------------- 8(f)-2: synthetic "resquiggled reads" (per-base means + bases) ------
    n_reads, reads = 6, []
    for i in range(n_reads):
        nb = int(rs.randint(60, 140))
        codes = rs.randint(0, 4, nb)
        seq = ''.join('ACGT'[c] for c in codes)
        lv, _ = std_ref.get_exp_levels_from_seq(seq)           # nb - K + 1 levels
        means = np.concatenate([rs.normal(0, 1, cpos), lv + rs.normal(0, 0.25, lv.shape[0]),
                                rs.normal(0, 1, K - cpos - 1)])
        means[rs.randint(cpos, nb - K, 3)] += rs.choice([-1.5, 1.5], 3)     # a few outliers
        start = int(1000 + rs.randint(0, 80))
        strand = '+' if i % 3 else '-'
        reads.append((means, seq, start, strand))
    out['dn_n'] = n_reads
    out['dn_means'] = np.concatenate([r[0] for r in reads])
    out['dn_off'] = np.concatenate([[0], np.cumsum([r[0].shape[0] for r in reads])]).astype(np.int64)
    out['dn_seq'] = np.array([r[1] for r in reads])
    out['dn_start'] = np.array([r[2] for r in reads], dtype=np.int64)
    out['dn_strand'] = np.array([r[3] for r in reads])
    reg = Reg(1040, 1110)
    out['reg'] = np.array([reg.start, reg.end], dtype=np.int64)
    # control levels for the sample comparison: region [start - fm, end + fm), some NaN
    for fm in (0, 1, 2):
        size = reg.end - reg.start + 2 * fm
        cm = rs.normal(0, 1.2, size)
        cs = rs.uniform(0.15, 0.4, size)
        cm[rs.randint(0, size, 4)] = np.nan
        out['sc_cm_fm%d' % fm], out['sc_cs_fm%d' % fm] = cm, cs

    def with_seams(means, seq, fn):
        bases = np.array(list(seq), dtype='S1')
        orig = (th.get_multiple_slots_read_centric, th.get_raw_read_slot,
                th.get_single_slot_read_centric)
        th.get_multiple_slots_read_centric = lambda *a, **k: (means.copy(), bases)
        th.get_single_slot_read_centric = lambda *a, **k: means.copy()
        th.get_raw_read_slot = lambda *a, **k: mock.MagicMock()
        try:
            with rh.ref_errstate():
                return fn()
        finally:
            (th.get_multiple_slots_read_centric, th.get_raw_read_slot,
             th.get_single_slot_read_centric) = orig

    for i, (means, seq, start, strand) in enumerate(reads):
        r_data = th.readData(start=start, end=start + means.shape[0], filtered=False,
                             read_start_rel_to_raw=0, strand=strand, fn='x', corr_group='g',
                             rna=False)
        for fm in (0, 1, 2):
            for tag, rg in (('whole', None), ('reg', reg)):
                key = 'dn_r%d_fm%d_%s' % (i, fm, tag)
                try:
                    pv, ps, _ = with_seams(means, seq, lambda: ts.compute_de_novo_read_stats(
                        r_data, std_ref, fm, rg))
                    out[key + '_p'], out[key + '_pos'] = pv['de_novo'], ps['de_novo'].astype(np.int64)
                    out[key + '_err'] = ''
                except th.TomboError as e:
                    out[key + '_err'] = str(e)
            key = 'sc_r%d_fm%d' % (i, fm)
            try:
                pv, ps, _ = with_seams(means, seq, lambda: ts.compute_sample_compare_read_stats(
                    r_data, out['sc_cm_fm%d' % fm], out['sc_cs_fm%d' % fm], fm, reg))
                out[key + '_p'] = pv['sample_compare']
                out[key + '_pos'] = ps['sample_compare'].astype(np.int64)
                out[key + '_err'] = ''
            except th.TomboError as e:
                out[key + '_err'] = str(e)
    pv = np.concatenate([10.0 ** rs.uniform(-60, 0, 60), [np.nan, 0.5, 1.0, 0.0, 1e-300]])
    rs.shuffle(pv)
    out['fw_p'] = pv
    with rh.ref_errstate(), np.errstate(divide='ignore'):
        for lag in (1, 2, 4):
            out['fw_lag%d' % lag] = ts.calc_window_fishers_method(pv.copy(), lag)

    # ---------------- 8(f)-1: region aggregation of per-read statistics -------------------
    n_r = 40
    stats, locs = [], []
    for i in range(n_r):
        s0 = int(rs.randint(5000, 5400))
        npos = int(rs.randint(20, 200))
        p = np.sort(rs.choice(np.arange(s0, s0 + 400), npos, replace=False)).astype(np.int64)
        v = rs.normal(0.5, 3.0, npos)
        v[rs.randint(0, npos, 2)] = np.nan
        stats.append(v); locs.append(p)
    out['rg_stats'] = np.concatenate(stats)
    out['rg_locs'] = np.concatenate(locs)
    out['rg_off'] = np.concatenate([[0], np.cumsum([s.shape[0] for s in stats])]).astype(np.int64)
    reg2 = Reg(5000, 6000)
    cases = {'alt_lower': (2.5, -1.5, 'model_compare'), 'alt_abs': (2.0, None, 'model_compare'),
             'denovo': (0.7, None, 'de_novo')}
    with rh.ref_errstate(), np.errstate(invalid='ignore'):
        for name, (thr, lower, st) in cases.items():
            r = ts.collate_reg_stats([s.copy() for s in stats], [l.copy() for l in locs], None, None,
                                     reg2, thr, lower, st, st, None)
            out['rg_%s_frac' % name] = r.reg_frac_standard_base
            out['rg_%s_pos' % name] = r.reg_poss.astype(np.int64)
            out['rg_%s_cov' % name] = np.asarray(r.reg_cov, dtype=np.int64)
            out['rg_%s_valid' % name] = np.asarray(r.valid_cov, dtype=np.int64)
            out['rg_%s_params' % name] = np.array([thr, np.nan if lower is None else lower])
            damp = {'unmod': 2, 'mod': 0}
            with np.errstate(all='ignore'):
                out['rg_%s_damp' % name] = ts.calc_damp_fraction(
                    damp, r.reg_frac_standard_base, np.asarray(r.valid_cov, dtype=np.float64))
    np.savez_compressed(os.path.join(HERE, 'region_stats.npz'), **out)
    print('region_stats.npz:', len(out), 'arrays;',
          sorted(set(str(out[k]) for k in out if k.endswith('_err'))))


if __name__ == '__main__':
    main()
