// dp_row.cuh -- warp-per-read banded dynamic programming engine.
//
// One warp owns one read.  A band row (W cells, fp64) is split into 32
// contiguous chunks, one per lane; the row lives in shared memory in a
// lane-transposed layout (cell j -> buf[(j % chunk) * 32 + j / chunk]) so that
// the lanes' serial walks are bank-conflict free.
//
// Exactness (SURVEY.md s7 hard parts 1,2): the reference recurrence
//     x_j = max_tie( (x_{j-1} - stay_pen) + z_j ,  diag_j , skip_j )
// (_c_dynamic_programming.pyx:213-234) is a serial chain inside a row.  We do
// NOT re-associate it.  Every lane first walks its chunk speculatively with an
// unknown left neighbour (x_in = -inf), then lanes whose true x_in (the left
// lane's last cell) differs re-walk their prefix until the recomputed value is
// bit-identical to the stored one.  f_j is monotone in x_{j-1}, so from the first
// bit-identical cell on the speculative suffix is exactly what the serial
// reference computes; rounds repeat until no lane's input changed (<= 32, usually
// 2).  Every cell is evaluated with the reference's fp64 operations in the
// reference's order, so fwd values, move codes and arg-max are bit-exact.
#pragma once
#include "common.cuh"

struct DpConsts {
    double z_shift, stay_pen, skip_pen, mhz;
    int winsor;
};

// how to evaluate the shifted z-score of band cell j of one row
struct RowSpec {
    const double *em;    // event means (index e0 + j)
    const double *zrow;  // explicit z-scores of this row (mirror API) or nullptr
    double mu, sd;
    int e0;              // event index of band position 0 (may be negative)
    int lo, hi;          // cells outside [lo, hi) hold maskval
    double maskval;
};

__device__ __forceinline__ double tb2_zscore(const RowSpec &rs, const DpConsts &c, int j)
{
    if (rs.zrow) return rs.zrow[j];
    if (j < rs.lo || j >= rs.hi) return rs.maskval;
    // z_shift - min(max_half_z, |ev - mean| / sd)
    // (_c_dynamic_programming.pyx:366-372; resquiggle.py:574-582)
    double a = fabs(__ldg(rs.em + (rs.e0 + j)) - rs.mu) / rs.sd;
    if (c.winsor) a = (c.mhz < a) ? c.mhz : a;
    return c.z_shift - a;
}

// One lane walks its chunk left to right.  FIX=false: speculative first walk.
// FIX=true: re-walk from the true left input until bit-identical.
template <int WPL, bool FIX>
__device__ __forceinline__ void tb2_walk(const double *prev, double *cur, int W, int chunk,
                                         int lane, int nvalid, const RowSpec &rs,
                                         const DpConsts &c, int d, bool first_skip, double x,
                                         uint32_t (&codes)[WPL], double &best, int &best_idx,
                                         double &x_end)
{
    const int j0 = lane * chunk;
    int p = j0 + d;  // previous-row band position under cell j (skip source)
    int lane_p = p / chunk;
    int i_p = p - lane_p * chunk;
    double pm1 = tb2_neg_inf();  // previous-row value at p-1 (diag source)
    if (p >= 1 && p - 1 < W) {
        const int q = p - 1, lq = q / chunk;
        pm1 = prev[(q - lq * chunk) * 32 + lq];
    }
    bool done = false;
#pragma unroll
    for (int w = 0; w < WPL; ++w) {
        const int ibeg = w * 16;
        const int iend = min(nvalid, ibeg + 16);
        uint32_t cw = FIX ? codes[w] : 0u;
        for (int i = ibeg; i < iend && !done; ++i) {
            const int j = j0 + i;
            const double z = tb2_zscore(rs, c, j);
            const double pv = (p < W) ? prev[i_p * 32 + lane_p] : tb2_neg_inf();
            double nx;
            uint32_t code;
            if (j == 0) {
                // band position 0: skip if the band did not move, else diag
                // (_c_dynamic_programming.pyx:261-270, 393-401); never a stay
                if (first_skip) { nx = pv - c.skip_pen; code = 1u; }
                else { nx = pm1 + z; code = 2u; }
            } else {
                const double a = (x - c.stay_pen) + z;   // stay  (code 0)
                double cc = pm1 + z;                     // diag  (code 2)
                uint32_t cf = 2u;
                const double sk = pv - c.skip_pen;       // skip  (code 1)
                if (sk > cc) { cc = sk; cf = 1u; }
                if (cc > a) { nx = cc; code = cf; }
                else { nx = a; code = 0u; }
            }
            const int sh = 2 * (i - ibeg);
            if (FIX) {
                const double old = cur[i * 32 + lane];
                cw = (cw & ~(3u << sh)) | (code << sh);
                if (nx == old) done = true;
            } else {
                cw |= code << sh;
            }
            cur[i * 32 + lane] = nx;
            if (nx > best || (FIX && nx == best && j < best_idx)) { best = nx; best_idx = j; }
            x = nx;
            pm1 = pv;
            ++p;
            if (++i_p == chunk) { i_p = 0; ++lane_p; }
        }
        codes[w] = cw;
    }
    if (!done && nvalid > 0) x_end = x;
}

// One band row for the whole warp.  On return cur holds the row, codes the 2-bit
// moves of this lane's chunk, (best, best_idx) this lane's first arg-max.
template <int WPL>
__device__ __forceinline__ void tb2_dp_row(const double *prev, double *cur, int W, int chunk,
                                           int lane, const RowSpec &rs, const DpConsts &c,
                                           int d, bool first_skip, uint32_t (&codes)[WPL],
                                           double &best, int &best_idx)
{
    const int nvalid = max(0, min(chunk, W - lane * chunk));
    best = tb2_neg_inf();
    best_idx = 0x7fffffff;
    double x_end = tb2_neg_inf();
    tb2_walk<WPL, false>(prev, cur, W, chunk, lane, nvalid, rs, c, d, first_skip, tb2_neg_inf(),
                         codes, best, best_idx, x_end);
    double last_in = tb2_neg_inf();
    for (;;) {
        const double xin = __shfl_up_sync(TB2_FULL_MASK, x_end, 1);
        const bool need = (lane > 0) && (nvalid > 0) && (xin != last_in);
        if (!__any_sync(TB2_FULL_MASK, need)) break;
        if (need) {
            tb2_walk<WPL, true>(prev, cur, W, chunk, lane, nvalid, rs, c, d, first_skip, xin,
                                codes, best, best_idx, x_end);
            last_in = xin;
        }
    }
}

// warp arg-max with first-maximum semantics (c_argmax :186-197, np.argmax)
__device__ __forceinline__ int tb2_warp_argmax(double best, int best_idx)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(TB2_FULL_MASK, best, off);
        const int oi = __shfl_xor_sync(TB2_FULL_MASK, best_idx, off);
        if (ob > best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
    }
    return best_idx;
}

// ---------------------------------------------------------------------------
// Pass description: a run of consecutive rows sharing one band width.
// ---------------------------------------------------------------------------
enum { TB2_MODE_PLAIN = 0, TB2_MODE_MASKED = 1, TB2_MODE_ADAPTIVE = 2, TB2_MODE_EXPLICIT = 3 };

struct PassCtx {
    // geometry
    int W, chunk;
    // row buffers (lane-transposed, >= chunk*32 doubles each)
    double *buf0, *buf1;
    // inputs
    const double *em;   // event means (already offset by events_start_clip)
    int n_em;
    const double *rm, *rs_;  // reference levels per base
    const double *zmat;      // EXPLICIT mode: n_rows x W z-scores
    // masked-start parameters (resquiggle.py:607-683)
    int mso;                 // mapped_start_offset
    double msp_start, msp_stop;  // mask_start_pos = linspace(msp_start, msp_stop, MASK_BASES)
    double mask_shifted;     // (mask_fill - z_shift) + z_shift
    double mask_fill;        // MASK_FILL_Z_SCORE (adaptive padding)
    // outputs
    int *starts;             // band event starts per base (global)
    uint32_t *tb;            // packed moves: row r at tb[r * wpl * 32 + w * 32 + lane]
    // mirror-API dumps (may be null)
    double *dbg_fwd;         // (n_bases + 1) x W
    long long *dbg_tb;       // (n_bases + 1) x W
};

#define TB2_MASK_BASES 50

// Runs rows [r_begin, r_end).  `prev` must hold fwd row r_begin on entry (buf
// selected by *cur_sel), starts[r_begin-1] valid if r_begin > 0.  Returns status
// (TB2_OK or TB2_ERR_ADAPTIVE_BEYOND_SIGNAL); on return *cur_sel selects the
// buffer holding fwd row r_end and (best_idx_out) its first arg-max.
template <int WPL>
__device__ int tb2_run_rows(const PassCtx &pc, const DpConsts &c, int mode, int r_begin,
                            int r_end, int n_bases_total, int *cur_sel, int *argmax_out)
{
    const int lane = tb2_lane();
    const int W = pc.W, chunk = pc.chunk;
    double *prev = (*cur_sel) ? pc.buf1 : pc.buf0;
    double *cur = (*cur_sel) ? pc.buf0 : pc.buf1;
    int last_argmax = *argmax_out;
    int prev_start = (r_begin > 0) ? pc.starts[r_begin - 1] : 0;
    const int half_bw = W / 2;
    for (int r = r_begin; r < r_end; ++r) {
        RowSpec rs;
        rs.em = pc.em; rs.zrow = nullptr;
        rs.mu = pc.rm ? __ldg(pc.rm + r) : 0.0;
        rs.sd = pc.rs_ ? __ldg(pc.rs_ + r) : 1.0;
        rs.lo = 0; rs.hi = W; rs.maskval = pc.mask_fill;
        int cur_start;
        if (mode == TB2_MODE_ADAPTIVE) {
            // _c_dynamic_programming.pyx:344-358
            cur_start = prev_start + last_argmax - half_bw + 1;
            if (cur_start < prev_start) cur_start = prev_start;
            if (cur_start >= pc.n_em) {
                if (r < n_bases_total - 2) return TB2_ERR_ADAPTIVE_BEYOND_SIGNAL;
                cur_start = pc.n_em - 1;
            }
            if (lane == 0) pc.starts[r] = cur_start;
            rs.hi = min(W, pc.n_em - cur_start);
        } else {
            cur_start = pc.starts[r];
            if (mode == TB2_MODE_MASKED) {
                // get_start_mask_z_score resquiggle.py:647-673 (validated by caller)
                const int sml = max(pc.mso - cur_start, 0);
                int eml = 0;
                if (r < TB2_MASK_BASES) {
                    const int msp = (int)tb2_linspace_at(pc.msp_start, pc.msp_stop,
                                                         TB2_MASK_BASES, r);
                    eml = W - (msp - cur_start);
                }
                if (cur_start + W - eml > pc.n_em) eml = cur_start + W - pc.n_em;
                rs.lo = sml; rs.hi = W - eml; rs.maskval = pc.mask_shifted;
            } else if (mode == TB2_MODE_EXPLICIT) {
                rs.zrow = pc.zmat + (size_t)r * W;
            }
        }
        rs.e0 = cur_start;
        const int d = (r > 0) ? cur_start - prev_start : 0;
        const bool first_skip = (r == 0) || (d == 0);
        uint32_t codes[WPL];
        double best; int best_idx;
        tb2_dp_row<WPL>(prev, cur, W, chunk, lane, rs, c, d, first_skip, codes, best, best_idx);
#pragma unroll
        for (int w = 0; w < WPL; ++w)
            pc.tb[(size_t)r * (WPL * 32) + w * 32 + lane] = codes[w];
        if (mode == TB2_MODE_ADAPTIVE || r == r_end - 1 || r == n_bases_total - 1)
            last_argmax = tb2_warp_argmax(best, best_idx);
        __syncwarp();
        if (pc.dbg_fwd) {
            const int nvalid = max(0, min(chunk, W - lane * chunk));
            for (int i = 0; i < nvalid; ++i) {
                const int j = lane * chunk + i;
                pc.dbg_fwd[(size_t)(r + 1) * W + j] = cur[i * 32 + lane];
                uint32_t cwv = codes[0];
#pragma unroll
                for (int k = 1; k < WPL; ++k) if ((i >> 4) == k) cwv = codes[k];
                pc.dbg_tb[(size_t)(r + 1) * W + j] = (cwv >> (2 * (i & 15))) & 3u;
            }
        }
        double *t = prev; prev = cur; cur = t;
        *cur_sel ^= 1;
        prev_start = cur_start;
    }
    *argmax_out = last_argmax;
    return TB2_OK;
}

// fwd row 0 = zeros (_c_dynamic_programming.pyx:253-254)
__device__ __forceinline__ void tb2_init_row0(const PassCtx &pc, int *cur_sel)
{
    const int lane = tb2_lane();
    for (int i = 0; i < pc.chunk; ++i) pc.buf0[i * 32 + lane] = 0.0;
    *cur_sel = 0;
    __syncwarp();
}

// load an explicit fwd row (mirror API of the in-place adaptive pass)
__device__ __forceinline__ void tb2_load_row(const PassCtx &pc, const double *row, int *cur_sel)
{
    const int lane = tb2_lane();
    for (int j = lane; j < pc.W; j += 32) {
        const int lj = j / pc.chunk;
        pc.buf0[(j - lj * pc.chunk) * 32 + lj] = row[j];
    }
    *cur_sel = 0;
    __syncwarp();
}

// ---------------------------------------------------------------------------
// Traceback (c_banded_traceback _c_dynamic_programming.pyx:281-310), warp
// cooperative: every lane loads its own packed words of a row (coalesced, position
// independent, four rows in flight), cells are fetched with shuffles.
// ---------------------------------------------------------------------------
template <int WPL>
__device__ __forceinline__ uint32_t tb2_tb_code(const uint32_t (&w)[WPL], int bp, int chunk)
{
    const int lj = bp / chunk;
    const int i = bp - lj * chunk;
    uint32_t v = w[0];
#pragma unroll
    for (int k = 1; k < WPL; ++k) if ((i >> 4) == k) v = w[k];
    v = __shfl_sync(TB2_FULL_MASK, v, lj);
    return (v >> (2 * (i & 15))) & 3u;
}

template <int WPL>
__device__ int tb2_traceback(const uint32_t *tb, const int *starts, int n_bases, int W,
                             int chunk, int band_pos, int thresh, int *read_tb)
{
    const int lane = tb2_lane();
    int cur_event = band_pos + starts[n_bases - 1];
    if (lane == 0) read_tb[n_bases] = cur_event + 1;
    int sp = n_bases;
    while (sp > 0) {
        const int nblk = min(4, sp);
        uint32_t w[4][WPL];
        int st[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b < nblk) {
                const int row = sp - 1 - b;
#pragma unroll
                for (int k = 0; k < WPL; ++k)
                    w[b][k] = tb[(size_t)row * (WPL * 32) + k * 32 + lane];
                st[b] = starts[row];
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b < nblk) {
                int bp = cur_event - st[b];
                if (bp < 0 || bp >= W) return TB2_ERR_UNEXPECTED;
                uint32_t code = tb2_tb_code<WPL>(w[b], bp, chunk);
                while (code == 0u) {  // 0: stay in the current base
                    --bp;
                    if (bp < 0) return TB2_ERR_UNEXPECTED;
                    code = tb2_tb_code<WPL>(w[b], bp, chunk);
                }
                if (code == 2u) --bp;  // diagonal
                if (thresh >= 0 && min(bp, W - bp - 1) < thresh)
                    return TB2_ERR_BEYOND_BANDWIDTH;
                cur_event = st[b] + bp;
                if (lane == 0) read_tb[sp - 1 - b] = cur_event + 1;
            }
        }
        sp -= nblk;
    }
    __syncwarp();
    return TB2_OK;
}
