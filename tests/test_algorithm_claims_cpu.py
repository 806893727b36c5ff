"""CPU checks of the exactness arguments the CUDA stages rely on:
each restructuring used on the device is replayed here in numpy against the
straightforward reference formulation, on random inputs including ties.

These are properties of the algorithms, not of the kernels: the kernels themselves are
compared with the oracle and the golden vectors in the -m gpu tests."""
import numpy as np
import pytest


# ---------------------------------------------------------------------------
# k_normalize: the outlier-clipping medians in closed form
# (normalize_raw_signal tombo_stats.py:541-563)
# ---------------------------------------------------------------------------
def _raw(rs, n, kind):
    if kind == 'int16':                       # DAC values: many exact ties
        return rs.randint(380, 620, n).astype(np.float64)
    return 500.0 + 40.0 * rs.standard_normal(n)


@pytest.mark.parametrize('kind', ['float', 'int16'])
def test_normalisation_medians_closed_form(kind):
    rs = np.random.RandomState(11)
    for n in list(range(5, 60)) + [4149, 4150, 4151]:
        raw = _raw(rs, n, kind)
        shift = np.median(raw)
        scale = np.median(np.abs(raw - shift))
        if scale == 0:
            continue
        norm = (raw - shift) / scale
        med = np.median(norm)
        if n & 1:
            assert med == 0.0 and not np.signbit(med)
            assert np.median(np.abs(norm - med)) == 1.0
        else:
            srt = np.sort(raw)
            a, b = srt[n // 2 - 1], srt[n // 2]
            assert shift == (a + b) / 2.0
            assert med == (((a - shift) / scale) + ((b - shift) / scale)) / 2.0


# ---------------------------------------------------------------------------
# k_cpts: ranked greedy with an exclusion zone == Jacobi fixed point on sets
# (c_valid_cpts_w_cap _c_helper.pyx:99-118)
# ---------------------------------------------------------------------------
def _greedy_sorted(scores, m):
    """the reference: candidates by score descending (ties: larger position first, the
    pinned rule), accepted unless an accepted one sits within +-(m - 1)"""
    n = scores.shape[0]
    order = sorted(range(n), key=lambda i: (scores[i], i), reverse=True)
    acc = np.zeros(n, dtype=bool)
    blocked = np.zeros(n, dtype=bool)
    for i in order:
        if blocked[i]:
            continue
        acc[i] = True
        blocked[max(0, i - m + 1):i + m] = True
    return acc


def _greedy_fixed_point(scores, m):
    """what k_cpts does, on boolean arrays instead of 32-candidate words"""
    n = scores.shape[0]
    acc = np.zeros(n, dtype=bool)
    dec = np.zeros(n, dtype=bool)
    idx = np.arange(n)
    rounds = 0
    while not dec.all():
        rounds += 1
        accnb = np.zeros(n, dtype=bool)
        blocked = np.zeros(n, dtype=bool)
        for o in range(1, m):
            for sgn in (1, -1):
                k = idx + sgn * o
                ok = (k >= 0) & (k < n)
                kk = np.clip(k, 0, n - 1)
                accnb |= ok & acc[kk]
                outranks = (scores[kk] > scores) | ((scores[kk] == scores) & (kk > idx))
                blocked |= ok & ~dec[kk] & outranks
        und = ~dec
        new_rej = und & accnb
        new_acc = und & ~accnb & ~blocked
        assert (new_rej | new_acc).any(), 'no progress'
        acc |= new_acc
        dec |= new_rej | new_acc
    return acc, rounds


@pytest.mark.parametrize('m', [2, 3, 6])
def test_greedy_exclusion_fixed_point_equals_ranked_greedy(m):
    rs = np.random.RandomState(5 + m)
    for trial in range(12):
        n = rs.randint(1, 400)
        scores = np.abs(rs.standard_normal(n))
        if trial % 3 == 0:
            scores = np.round(scores * 4) / 4          # heavy ties
        a = _greedy_sorted(scores, m)
        b, rounds = _greedy_fixed_point(scores, m)
        assert np.array_equal(a, b)
        assert rounds <= n + 1


# ---------------------------------------------------------------------------
# k_theil_sen: symmetry of the slope, and the guarded fp32 screen
# (c_compute_slopes _c_helper.pyx:362-377)
# ---------------------------------------------------------------------------
def _slope(ev, md, i, j):
    de = ev[i] - ev[j]
    return 1000.0 if de == 0 else (md[i] - md[j]) / de


def test_slope_is_symmetric_so_sorting_by_ev_keeps_the_multiset():
    rs = np.random.RandomState(2)
    ev = rs.standard_normal(120)
    ev[7] = ev[31]                                    # an exact tie
    md = 0.9 * ev + 0.3 * rs.standard_normal(120)
    for i in range(0, 120, 7):
        for j in range(120):
            if i != j:
                assert _slope(ev, md, i, j) == _slope(ev, md, j, i)
    o = np.argsort(ev, kind='stable')
    a = sorted(_slope(ev, md, i, j) for i in range(120) for j in range(i + 1, 120))
    b = sorted(_slope(ev[o], md[o], i, j) for i in range(120) for j in range(i + 1, 120))
    assert a == b


def test_fp32_screen_never_contradicts_the_fp64_slope():
    """pairs a < b of points sorted by ev; Q_T(k) = md_k - T ev_k in fp32.  Wherever the
    screen of k_theil_sen decides (difference beyond the guard, distinct fp32 ev), the
    fp64 slope the reference computes is on the same side of T."""
    f32 = np.float32
    rs = np.random.RandomState(9)
    checked = 0
    for trial in range(30):
        n = 200
        scale = 10.0 ** rs.uniform(-1, 2)
        ev = np.sort(scale * rs.standard_normal(n))
        md = rs.uniform(0.5, 1.5) * ev + scale * 0.2 * rs.standard_normal(n)
        if trial % 4 == 0:
            ev[50] = ev[51]
        M = max(1.0, float(max(np.abs(ev.astype(f32)).max(), np.abs(md.astype(f32)).max())))
        for T in (np.median((md[1:] - md[:-1]) / np.where(ev[1:] == ev[:-1], 1, ev[1:] - ev[:-1])),
                  0.7, 1.3, -0.2):
            Tf = f32(T)
            g = f32(1e-5) * f32(M) * (f32(1.0) + abs(Tf))
            evf, mdf = ev.astype(f32), md.astype(f32)
            # fmaf(-T_f, ev_f, md_f): exact product and sum in fp64, one rounding to fp32
            q = (mdf.astype(np.float64) - Tf.astype(np.float64) * evf.astype(np.float64)).astype(f32)
            for a in range(0, n, 3):
                b = np.arange(a + 1, n)
                diff = q[a] - q[b]
                distinct = evf[a] != evf[b]
                de = ev[a] - ev[b]
                with np.errstate(divide='ignore', invalid='ignore'):
                    s = np.where(de == 0, 1000.0, (md[a] - md[b]) / de)
                low = distinct & (diff > g)            # screen: certainly slope < T
                high = distinct & (diff < -g)          # screen: certainly slope >= T
                assert np.all(s[low] < T)
                assert np.all(s[high] >= T)
                checked += int(low.sum() + high.sum())
    assert checked > 100000


# ---------------------------------------------------------------------------
# wavefront engine: the predicated ("lean") cell update == the reference row update
# (c_banded_forward_pass inner loop, _c_dynamic_programming.pyx:213-234)
# ---------------------------------------------------------------------------
def _row_reference(prev, z, d, stay, skip):
    W = prev.shape[0]
    NEG = -np.inf
    out = np.empty(W)
    codes = np.empty(W, dtype=np.int64)
    x = 0.0
    for j in range(W):
        p = j + d
        u = prev[p] if p < W else NEG
        ul = prev[p - 1] if 1 <= p <= W else NEG
        if j == 0:
            if d == 0:
                nx, code = u - skip, 1            # band did not move: skip, no stay / diag
            else:
                nx, code = ul + z[j], 2           # band moved: diagonal only
        else:
            a = (x - stay) + z[j]
            cc, cf = ul + z[j], 2
            sk = u - skip
            if sk > cc:
                cc, cf = sk, 1
            if cc > a:
                nx, code = cc, cf
            else:
                nx, code = a, 0
        out[j], codes[j] = nx, code
        x = nx
    return out, codes


def _row_lean(prev, z, d, stay, skip):
    """one formula for every cell: x starts at -inf, the cells above / above-left are
    replaced by -inf outside their validity ranges (dp_row.cuh TB2_WF_LEAN_STEP)"""
    W = prev.shape[0]
    NEG = -np.inf
    julo = 1 if d >= 1 else 0
    jllo = 1 - julo
    n_u = max(W - d - julo, 0)
    n_ul = max(W - d - jllo + 1, 0)
    out = np.empty(W)
    codes = np.empty(W, dtype=np.int64)
    x = NEG
    for j in range(W):
        p = j + d
        u = prev[p] if 0 <= j - julo < n_u else NEG
        ul = prev[p - 1] if 0 <= j - jllo < n_ul else NEG
        a = (x - stay) + z[j]
        cc, code = ul + z[j], 2
        sk = u - skip
        if sk > cc:
            cc, code = sk, 1
        nx = a
        if cc > a:
            nx = cc
        else:
            code = 0
        out[j], codes[j] = nx, code
        x = nx
    return out, codes


def test_lean_wavefront_step_equals_reference_row_update():
    rs = np.random.RandomState(21)
    for trial in range(400):
        W = rs.randint(2, 40)
        d = rs.randint(0, min(6, W))
        prev = 5.0 * rs.standard_normal(W)
        z = 3.0 - np.abs(2.0 * rs.standard_normal(W))
        if trial % 5 == 0:
            z = np.round(z)                       # exact ties between the three moves
            prev = np.round(prev)
        a, ca = _row_reference(prev, z, d, 4.0, 4.0)
        b, cb = _row_lean(prev, z, d, 4.0, 4.0)
        assert np.array_equal(a, b), (trial, W, d)
        assert np.array_equal(ca, cb), (trial, W, d)
