#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run in the build container only:

    python oracle/build_ref.py && python tests/golden/make_golden.py

The reference holds no golden vectors of its own, so these
files -- outputs of the reference's own code on seeded synthetic inputs -- are what
pins the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_golden_gpu.py).  Inputs are regenerated from the stored seeds with
tombo_b200.synthetic (numpy RandomState streams are frozen); a checksum of every
raw signal is stored to detect generator drift.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))

import ref_harness as rh  # noqa: E402
from tombo_b200 import synthetic as syn  # noqa: E402

CONFIGS = {
    # name: (kind, sig_aln_params, n_bases list / scalar, n_reads, seed0, extra)
    'dna_static4k': ('DNA', (4.2, 4.2, 200, 1500, 20.0, 40, 750, 2500, 250), 444, 16, 11000, {}),
    'dna_adapt4k': ('DNA', (4.2, 4.2, 200, 1500, 20.0, 40, 300, 2500, 100), 444, 16, 12000, {}),
    'dna_adapt_bw400': ('DNA', (4.2, 4.2, 400, 1500, 20.0, 40, 750, 2500, 250), 2222, 3, 13000, {}),
    'dna_long_subsample': ('DNA', (4.2, 4.2, 400, 1500, 20.0, 40, 750, 2500, 250), 1300, 3, 14000, {}),
    'dna_int16_stable': ('DNA', (4.2, 4.2, 200, 1500, 20.0, 40, 750, 2500, 250), 444, 8, 15000,
                         {'int16': True}),
    'dna_rescue': ('DNA', (4.2, 4.2, 120, 1500, 20.0, 40, 300, 2500, 100), 600, 6, 16000,
                   {'stall': True}),
    'rna_8k': ('RNA', None, 270, 6, 17000, {}),
    # --- round 2: the shapes of BASELINE.json configs[2..4] ---
    # configs[4]: 50k-sample reads, bandwidth 1200; read 2 carries a planted stall that leaves
    # the 1200 band and is rescued with the save bandwidth (1500)
    'dna_c5_bw1200': ('DNA', (4.2, 4.2, 1200, 1500, 20.0, 40, 750, 2500, 250), 5555, 3, 19000,
                      {'stall_at': {2: (2500, 11000)}}),
    # configs[2]: a 20k-sample bandwidth-400 read that only aligns with the save bandwidth
    'dna_c3_rescue_long': ('DNA', (4.2, 4.2, 400, 1500, 20.0, 40, 750, 2500, 250),
                           [2222, 2222, 1800], 3, 20000,
                           {'stall_at': {0: (1200, 3500), 2: (900, 2500)}}),
    # configs[3]: direct RNA, 8k samples, caller-supplied const_scale (median_const_scale branch)
    'rna_const_scale': ('RNA', None, 270, 3, 21000, {'const_scale': 95.0}),
}


def run_config(name):
    kind, aln, nbases, nreads, seed0, extra = CONFIGS[name]
    kmer_ref, cpos = syn.make_kmer_ref(kind, 0)
    std_ref, _ = rh.make_models(kmer_ref, cpos)
    sst, p, sp = rh.make_params(kind, aln)
    out = dict(kind=kind, aln=np.array(aln if aln is not None else [], dtype=np.float64),
               seed0=seed0, nreads=nreads,
               nbases=np.broadcast_to(np.asarray(nbases), (nreads,)).astype(np.int64),
               int16=bool(extra.get('int16')), stall=bool(extra.get('stall')))
    segs, seg_off, scal, msgs, chk = [], [0], [], [], []
    stall_at = np.full((nreads, 2), -1, dtype=np.int64)
    for i, v in extra.get('stall_at', {}).items():
        stall_at[i] = v
    out['stall_at'] = stall_at
    out['const_scale'] = float(extra.get('const_scale', np.nan))
    for i in range(nreads):
        kw = {}
        if extra.get('int16'):
            kw['int16'] = True
        if extra.get('stall'):
            kw['stall'] = (300 + i, 1500)
        if stall_at[i, 0] >= 0:
            kw['stall'] = (int(stall_at[i, 0]), int(stall_at[i, 1]))
        r = syn.make_read(kmer_ref, cpos, int(out['nbases'][i]), seed0 + i, kind=kind, **kw)
        res, err, info = rh.run_read(r.raw, r.genome_seq, std_ref, sst, p, sp, read_index=i,
                                     stable_ties=bool(extra.get('int16')),
                                     const_scale=extra.get('const_scale'))
        chk.append(float(np.sum(np.asarray(r.raw, dtype=np.float64))))
        if res is None:
            msgs.append(err)
            scal.append([np.nan] * 6 + [info['calls'], int(info['rescued']), -1, 0])
            seg_off.append(seg_off[-1])
            continue
        msgs.append('')
        segs.append(res.segs.astype(np.int64))
        seg_off.append(seg_off[-1] + res.segs.shape[0])
        sv = res.scale_values
        scal.append([sv.shift, sv.scale, sv.lower_lim, sv.upper_lim, res.sig_match_score,
                     res.read_start_rel_to_raw, info['calls'], int(info['rescued']),
                     info['n_iters'], int(res.norm_params_changed)])
    out.update(segs=np.concatenate(segs) if segs else np.zeros(0, np.int64),
               seg_off=np.array(seg_off, dtype=np.int64), scalars=np.array(scal, dtype=np.float64),
               messages=np.array(msgs), raw_checksum=np.array(chk))
    return out


def kernel_kats():
    """Known-answer vectors of the Cython entry points on small random inputs."""
    m = rh.load_reference()
    cdp, ch = m['cdp'], m['ch']
    rs = np.random.RandomState(424242)
    out = {}
    with rh.ref_errstate():
        z = 5.0 - np.minimum(20.0, np.abs(rs.normal(0, 3, (60, 90))))
        es = np.cumsum(rs.randint(0, 3, 60)).astype(np.int64)
        fwd, tb = cdp.c_banded_forward_pass(z, es, 4.2, 4.2)
        out.update(bfp_z=z, bfp_es=es, bfp_fwd=fwd, bfp_tb=tb[1:])
        top = int(np.argmax(fwd[-1]))
        out['bfp_traceback'] = cdp.c_banded_traceback(tb, es, top, -1)
        # adaptive
        nb, bw, ssp = 120, 64, 20
        rm = rs.normal(0, 1.4826, nb)
        rsd = rs.uniform(0.15, 0.3, nb)
        em = np.repeat(rm, np.maximum(1, rs.poisson(2.0, nb))) + rs.normal(0, 0.2, 1)[0]
        em = em + rs.normal(0, 0.2, em.shape[0])
        es0 = (np.arange(ssp) * 2).astype(np.int64)
        z0 = np.empty((ssp, bw))
        for r in range(ssp):
            z0[r] = 5.0 - np.minimum(20.0, np.abs(em[es0[r]:es0[r] + bw] - rm[r]) / rsd[r])
        f0, t0 = cdp.c_banded_forward_pass(z0, es0, 4.2, 4.2)
        fwd = np.zeros((nb + 1, bw)); tbm = np.zeros((nb + 1, bw), dtype=np.int64)
        esf = np.zeros(nb, dtype=np.int64)
        fwd[:ssp + 1] = f0; tbm[:ssp + 1] = t0; esf[:ssp] = es0
        out.update(ad_seed_fwd=fwd.copy(), ad_seed_tb=tbm.copy(), ad_seed_es=esf.copy(),
                   ad_em=em, ad_rm=rm, ad_rs=rsd, ad_ssp=ssp)
        try:
            cdp.c_adaptive_banded_forward_pass(fwd, tbm, esf, em, rm, rsd, 5.0, 4.2, 4.2, ssp,
                                               -15.0, True, 20.0)
            out['ad_ok'] = 1
        except NotImplementedError:
            out['ad_ok'] = 0
        out.update(ad_fwd=fwd, ad_tb=tbm, ad_es=esf)
        # helpers
        sig = rs.normal(0, 1, 3000)
        segs = np.sort(rs.choice(np.arange(1, 3000), 400, replace=False)).astype(np.int64)
        out.update(h_sig=sig, h_segs=segs, h_means=ch.c_new_means(sig, segs))
        mm, ss = ch.c_new_mean_stds(sig, segs)
        out.update(h_mean_stds_m=mm, h_mean_stds_s=ss)
        cp = ch.c_valid_cpts_w_cap(sig, 3, 5, 500)
        cp.sort()
        out['h_cpts'] = cp
        cpt = ch.c_valid_cpts_w_cap_t_test(sig, 6, 12, 150)
        cpt.sort()
        out['h_cpts_t'] = cpt
        ev = rs.normal(0, 1, 40); md = ev * 1.05 + 0.1 + rs.normal(0, 0.1, 40)
        out.update(h_ev=ev, h_md=md, h_slopes=ch.c_compute_slopes(ev, md))
        # likelihood ratios
        m7 = rs.normal(0, 1, 6); r7 = rs.normal(0, 1, 6); a7 = r7 + rs.normal(0, 0.3, 6)
        a7[2] = r7[2]
        out.update(l_m=m7, l_r=r7, l_a=a7,
                   l_scaled=ch.c_calc_scaled_llh_ratio_const_var(m7, r7, a7, 0.04, 4.0, 1.0, 0.2),
                   l_const=ch.c_calc_llh_ratio_const_var(m7, r7, a7, 0.04),
                   l_full=ch.c_calc_llh_ratio(m7, r7, a7, np.full(6, 0.04), np.full(6, 0.05)))
    return out


def llr_config(kind='DNA', seed0=18000, nbases=300):
    """compute_alt_model_read_stats on resquiggled synthetic reads (5mC alt model); DNA
    (6-mers) or direct RNA (5-mers, BASELINE.json configs[3])."""
    m = rh.load_reference()
    th, ts = m['th'], m['ts']
    kmer_ref, cpos = syn.make_kmer_ref(kind, 0)
    alt_rows = syn.make_alt_kmer_ref(kmer_ref, 'C', seed=1)
    std_ref, alt_ref = rh.make_models(kmer_ref, cpos, alt_rows, 'C')
    sst, p, sp = rh.make_params(kind, CONFIGS['dna_static4k'][1] if kind == 'DNA' else None)
    out = dict(seed0=seed0, nreads=4, nbases=nbases, kind=kind)
    llr_s, llr_p, pos, off = [], [], [], [0]
    for i in range(4):
        r = syn.make_read(kmer_ref, cpos, nbases, seed0 + i, kind=kind)
        res, err, info = rh.run_read(r.raw, r.genome_seq, std_ref, sst, p, sp, read_index=i)
        assert res is not None
        norm_mean = ts.compute_base_means(res.raw_signal, res.segs)
        bases = np.array(list(res.genome_seq), dtype='S1')
        r_data = th.readData(start=1000 * i, end=1000 * i + len(res.genome_seq), filtered=False,
                             read_start_rel_to_raw=0, strand='+', fn='x', corr_group='g',
                             rna=(kind == 'RNA'))
        orig = (th.get_multiple_slots_read_centric, th.get_raw_read_slot)
        from unittest import mock
        th.get_multiple_slots_read_centric = lambda *a, **k: (norm_mean, bases)
        th.get_raw_read_slot = lambda *a, **k: mock.MagicMock()
        try:
            with rh.ref_errstate():
                a, pp, _ = ts.compute_alt_model_read_stats(r_data, std_ref, [('5mC', alt_ref)])
                b, _, _ = ts.compute_alt_model_read_stats(r_data, std_ref, [('5mC', alt_ref)],
                                                          use_standard_llhr=True)
        finally:
            th.get_multiple_slots_read_centric, th.get_raw_read_slot = orig
        llr_s.append(a['5mC']); llr_p.append(b['5mC']); pos.append(pp['5mC'])
        off.append(off[-1] + a['5mC'].shape[0])
    out.update(llr_scaled=np.concatenate(llr_s), llr_standard=np.concatenate(llr_p),
               pos=np.concatenate(pos).astype(np.int64), site_off=np.array(off, dtype=np.int64))
    return out


if __name__ == '__main__':
    only = sys.argv[1:]
    for name in CONFIGS:
        if only and name not in only:
            continue
        o = run_config(name)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **o)
        ok = int(np.sum(o['messages'] == ''))
        print(name, 'reads', o['nreads'], 'ok', ok, 'rescued', int(o['scalars'][:, 7].sum()),
              'msgs', sorted(set(o['messages'])))
    if not only or 'kats' in only:
        np.savez_compressed(os.path.join(HERE, 'kernel_kats.npz'), **kernel_kats())
        print('kernel_kats done')
    if not only or 'llr' in only:
        np.savez_compressed(os.path.join(HERE, 'llr_5mc.npz'), **llr_config())
        print('llr done')
    if not only or 'llr_rna' in only:
        np.savez_compressed(os.path.join(HERE, 'llr_rna_5mc.npz'),
                            **llr_config('RNA', seed0=22000, nbases=270))
        print('llr_rna done')
