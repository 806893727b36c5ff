"""GPU parity of the banded DP kernels against the C oracle (bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_z(rs, nb, bw):
    return 5.0 - np.minimum(20.0, np.abs(rs.normal(0, 3.0, (nb, bw))))


@pytest.mark.parametrize('nb,bw,kind', [
    (40, 64, 'diag'), (250, 750, 'diag'), (120, 200, 'flat'), (300, 748, 'ramp'),
    (64, 33, 'diag'), (30, 2500, 'diag'), (50, 1200, 'ramp'), (17, 1, 'flat'),
    (10, 5, 'ramp')])
def test_banded_forward_pass_bit_exact(ctx, orc, nb, bw, kind):
    rs = np.random.RandomState(nb * 1000 + bw)
    z = _rand_z(rs, nb, bw)
    if kind == 'diag':
        es = np.arange(nb, dtype=np.int64)
    elif kind == 'flat':
        es = np.zeros(nb, dtype=np.int64)
    else:
        es = np.cumsum(rs.randint(0, 4, nb)).astype(np.int64)
        if bw < 8:
            es = np.cumsum(rs.randint(0, 2, nb)).astype(np.int64)
    f0, t0 = orc.banded_forward_pass(z, es, 4.2, 4.2)
    f1, t1 = ctx.banded_forward_pass(z, es, 4.2, 4.2)
    assert np.array_equal(f0, f1)
    assert np.array_equal(t0[1:], t1[1:])
    top = int(np.argmax(f0[-1]))
    s0, p0 = orc.banded_traceback(t0, es, top, -1)
    s1, p1 = ctx.banded_traceback(t1, es, top, -1)
    assert s0 == s1 == 0
    assert np.array_equal(p0, p1)
    # with a boundary threshold (may raise in both)
    s0, p0 = orc.banded_traceback(t0, es, top, 5)
    s1, p1 = ctx.banded_traceback(t1, es, top, 5)
    assert s0 == s1
    if s0 == 0:
        assert np.array_equal(p0, p1)


def _adaptive_case(seed, nb, bw, n_ev, ssp):
    rs = np.random.RandomState(seed)
    rm = rs.normal(0, 1.4826, nb)
    rsd = np.full(nb, 0.2) if seed % 2 else rs.uniform(0.1, 0.4, nb)
    # events follow the sequence with ~ n_ev / nb events per base
    per = np.maximum(1, rs.poisson(n_ev / nb, nb))
    em = np.repeat(rm, per) + rs.normal(0, 0.2, per.sum())
    em = em[:n_ev] if em.shape[0] >= n_ev else np.concatenate(
        [em, rs.normal(0, 1.5, n_ev - em.shape[0])])
    return rm, rsd, em


@pytest.mark.parametrize('seed,nb,bw,n_ev,ssp', [
    (1, 300, 200, 650, 51), (2, 400, 300, 800, 60), (3, 200, 400, 520, 101),
    (4, 150, 1200, 1500, 1), (5, 500, 96, 1100, 30), (6, 300, 200, 350, 20)])
def test_adaptive_forward_pass_bit_exact(ctx, orc, seed, nb, bw, n_ev, ssp):
    rm, rsd, em = _adaptive_case(seed, nb, bw, n_ev, ssp)
    z_shift = 4.2 + float(np.sqrt(2 / np.pi))
    # seed rows from a static pass
    es0 = (np.arange(ssp) * (n_ev / nb)).astype(np.int64)
    z0 = np.empty((ssp, bw))
    for r in range(ssp):
        seg = em[es0[r]:es0[r] + bw]
        z0[r, :seg.shape[0]] = z_shift - np.minimum(20.0, np.abs(seg - rm[r]) / rsd[r])
        z0[r, seg.shape[0]:] = -15.0
    f_seed, t_seed = orc.banded_forward_pass(z0, es0, 4.2, 4.2)

    def fresh():
        fwd = np.zeros((nb + 1, bw))
        tb = np.zeros((nb + 1, bw), dtype=np.int64)
        es = np.zeros(nb, dtype=np.int64)
        fwd[:ssp + 1] = f_seed
        tb[:ssp + 1] = t_seed
        es[:ssp] = es0
        return fwd, tb, es
    fa, ta, ea = fresh()
    sa, _ = orc.adaptive_banded_forward_pass(fa, ta, ea, em, rm, rsd, z_shift, 4.2,
                                             4.2, ssp, -15.0, True, 20.0)
    fb, tbb, eb = fresh()
    sb = ctx.adaptive_banded_forward_pass(fb, tbb, eb, em, rm, rsd, z_shift, 4.2, 4.2,
                                          ssp, -15.0, True, 20.0)
    assert sa == sb
    if sa == 0:
        assert np.array_equal(ea, eb)
        assert np.array_equal(fa, fb)
        assert np.array_equal(ta, tbb)


def _events_from_read(orc, read, std_means, std_sds, kmer, rp):
    """normalise + segment with the oracle to obtain DP inputs"""
    from tombo_b200 import synthetic as syn
    codes = syn.seq_to_codes(read.genome_seq).astype(np.int64)
    nb = codes.shape[0] - kmer + 1
    kidx = np.zeros(nb, dtype=np.int64)
    for j in range(kmer):
        kidx = kidx * 4 + codes[j:j + nb]
    rm, rsd = std_means[kidx], std_sds[kidx]
    st, norm, sv = orc.normalize_raw_signal(read.raw, outlier_thresh=5.0)
    ne = max(read.raw.shape[0] // rp.mean_obs_per_event, int(nb * 1.1))
    st, cpts = orc.valid_cpts_w_cap(norm, rp.min_obs_per_base, rp.running_stat_width, ne)
    assert st == 0
    em = orc.new_means(norm, cpts)
    return cpts, em, rm, rsd


@pytest.mark.parametrize('aln,nbases,seed', [
    ((4.2, 4.2, 200, 1500, 20.0, 40, 750, 2500, 250), 444, 1000),   # static path
    ((4.2, 4.2, 200, 1500, 20.0, 40, 300, 2500, 100), 444, 2000),   # adaptive, bw 200
    ((4.2, 4.2, 400, 1500, 20.0, 40, 750, 2500, 250), 1500, 3000),  # adaptive, bw 400
    ((4.2, 4.2, 300, 1500, 20.0, 40, 750, 2500, 250), 90, 4000),    # tiny read
])
def test_find_adaptive_base_assignment_bit_exact(ctx, orc, dna_model, RPcls, aln,
                                                 nbases, seed):
    from tombo_b200 import synthetic as syn
    kmer_ref, cpos = dna_model
    means, sds = syn.kmer_table(kmer_ref)
    rp = RPcls(aln)
    for i in range(4):
        read = syn.make_read(kmer_ref, cpos, nbases, seed + i)
        cpts, em, rm, rsd = _events_from_read(orc, read, means, sds, 6, rp)
        s0, segs0, r0, dbg0, _ = orc.find_adaptive_base_assignment(cpts, em, rp, rm, rsd)
        s1, segs1, r1, dbg1 = ctx.find_adaptive_base_assignment(cpts, em, rp, rm, rsd)
        assert s0 == s1, (s0, s1)
        assert np.array_equal(dbg0, dbg1), (dbg0, dbg1)
        if s0 == 0:
            assert r0 == r1
            assert np.array_equal(segs0, segs1)
