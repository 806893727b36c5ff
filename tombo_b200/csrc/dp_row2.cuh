// dp_row2.cuh -- register-resident engine for the ADAPTIVE band
// (c_adaptive_banded_forward_pass _c_dynamic_programming.pyx:314-412).
//
// The band start of row r+1 is a function of the arg-max of the complete row r
// (:344-346), so rows cannot be skewed against each other; the parallelism is inside a
// row.  Round 1 kept rows in shared memory in band-relative coordinates: every row
// re-read its W event means (stride-CH gathers), two row buffers and two fix-up caches,
// and the shared-memory footprint (13 KB / warp) capped the SM at 16 latency-bound warps.
//
// Here the row lives in REGISTERS in ABSOLUTE event coordinates.  Events are cut into
// chunks of CH; chunk c belongs to lane c % 32 and a lane holds one chunk at a time (the
// band [s, s+W) spans at most 32 chunks because W <= 31*CH + 1).  Consequences:
//   * cell (r, e) and cell (r-1, e) live in the same register of the same lane: the
//     "skip" source is free, the "diagonal" source is the neighbouring register (one
//     shuffle per row brings the last cell of the previous lane);
//   * an event mean is loaded from HBM once per read -- when its chunk enters the band --
//     and then stays in a register while the band slides over it (traffic: 8 bytes per
//     event per read instead of per row);
//   * no shared memory at all for the adaptive rows; moves leave as one coalesced
//     128-byte line per row (2 bits / cell, lane-major).
// The arithmetic of a cell is the reference's, in the reference's order (stay, then the
// better of diagonal / skip; first maximum wins), so every value is bit-identical.  The
// serial stay chain inside a row is resolved as in round 1: every lane first walks its
// chunk assuming nothing arrives from the left, then lanes whose true input is larger
// re-walk a prefix.  Because a larger input can only turn cells into "stay", the re-walk
// needs one add pair and one compare per cell and stops at the first cell the old value
// survives (f is monotone in its input; proof in DESIGN.md).
#pragma once
#include "dp_row.cuh"
#include "kernels.h"

#ifdef TB2_DP_COUNTERS
// tuning counters: [0] rows, [1] fix-up rounds, [2] re-walk steps (max over lanes, summed),
// [3] chunk reloads
#define TB2_CNT(i, v) do { if (tb2_lane() == 0) atomicAdd(&g_tb2_dp_counters[i], (unsigned long long)(v)); } while (0)
#else
#define TB2_CNT(i, v) do { } while (0)
#endif

// chunk width for a band of W cells: the smallest instantiated CH with W <= 31*CH + 1
__host__ __device__ __forceinline__ int tb2_abs_chunk(int W) { return tb2_abs_chunk_host(W); }
// packed move words per row and lane
__host__ __device__ __forceinline__ int tb2_abs_wpr(int ch) { return ch > 16 ? 2 : 1; }

// one speculative walk over the lane's chunk (chain input xc: -inf = nothing arrives from
// the left).  vmask: cells
// inside the band; skmask: cells that may take a skip (band position 0 may not once the
// band has moved, _c_dynamic_programming.pyx:261-270, 393-401); tmask (TAIL rows only):
// cells beyond the signal, whose z-score is the mask fill (:366-372).
// Moves are kept as two bit sets per lane: bit i of `msk` = skip beat the diagonal,
// bit i of `mcd` = that candidate beat the stay (code: mcd ? (msk ? 1 : 2) : 0).
template <int CH, bool TAIL>
__device__ __forceinline__ void tb2_abs_walk(const double (&em)[CH], double (&x)[CH], double (&z)[CH],
                                             double pm1, double xc, double mu, double sd,
                                             tb2_rcp inv_sd, double zs, double mhz, double stay,
                                             double skip, double maskval, uint32_t vmask,
                                             uint32_t skmask, uint32_t tmask, uint32_t &msk,
                                             uint32_t &mcd)
{
    const double NEG = tb2_neg_inf();
    uint32_t ms = 0u, mc = 0u;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const double pv = x[i];
        double a0 = tb2_div_by(fabs(em[i] - mu), sd, inv_sd);
        a0 = (mhz < a0) ? mhz : a0;
        double zz = zs - a0;
        if (TAIL) { if (tmask & (1u << i)) zz = maskval; }
        if (!(vmask & (1u << i))) zz = NEG;                  // outside the band: no such cell
        double sk = pv - skip;                               // skip  (code 1)
        if (!(skmask & (1u << i))) sk = NEG;
        double cand = pm1 + zz;                              // diag  (code 2)
        if (sk > cand) { cand = sk; ms |= 1u << i; }
        const double a = (xc - stay) + zz;                   // stay  (code 0)
        double nx = a;
        if (cand > a) { nx = cand; mc |= 1u << i; }
        x[i] = nx;
        z[i] = zz;
        xc = nx;
        pm1 = pv;
    }
    msk = ms; mcd = mc;
}

// rows [r_begin, r_end) of the adaptive pass.  On entry `rowbuf` holds fwd row r_begin
// (W doubles, band-relative, band start pc.starts[r_begin-1]) and *argmax_io its first
// arg-max; r_begin >= 1.  Moves of row r go to tb[((r - r_begin) * WPR + w) * 32 + lane]
// (CH <= 16: one word, skip set in the low half, candidate set in the high half; else two).
// On return *argmax_io is the first arg-max (band position) of fwd row r_end.
template <int CH>
__device__ __noinline__ int tb2_adaptive_rows_abs(const PassCtx &pc, const DpConsts &c, int r_begin,
                                                  int r_end, int nb_total, const double *rowbuf,
                                                  uint32_t *tb, int *argmax_io)
{
    constexpr int WPR = (CH > 16) ? 2 : 1;
    const int lane = tb2_lane();
    const int W = pc.W, half_bw = W / 2, n_em = pc.n_em;
    const double NEG = tb2_neg_inf();
    const double stay = c.stay_pen, skip = c.skip_pen, zs = c.z_shift;
    const double mhz = c.winsor ? c.mhz : __longlong_as_double(0x7ff0000000000000LL);   // +inf: no clamp
    const double maskval = pc.mask_fill;
    const double *em_g = pc.em;
    const int left_lane = (lane + 31) & 31;
    int prev_start = pc.starts[r_begin - 1];
    int last_argmax = *argmax_io;

    double em[CH], x[CH], z[CH];
    // chunk owned by this lane: the one in [c_lo, c_lo + 31] congruent to the lane
    int c_lo = prev_start / CH;
    int chunk = c_lo + ((lane - c_lo) & 31);
    {
        const int e0 = chunk * CH;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int e = e0 + i, j = e - prev_start;
            x[i] = ((unsigned)j < (unsigned)W) ? rowbuf[j] : NEG;
            em[i] = (e < n_em) ? __ldg(em_g + e) : 0.0;
            z[i] = 0.0;
        }
    }
    for (int r = r_begin; r < r_end; ++r) {
        // ---- band placement (_c_dynamic_programming.pyx:344-358) ----
        int cur_start = prev_start + last_argmax - half_bw + 1;
        if (cur_start < prev_start) cur_start = prev_start;
        if (cur_start >= n_em) {
            if (r < nb_total - 2) return TB2_ERR_ADAPTIVE_BEYOND_SIGNAL;
            cur_start = n_em - 1;
        }
        if (lane == 0) pc.starts[r] = cur_start;
        const int d = cur_start - prev_start;
        const double mu = __ldg(pc.rm + r), sd = __ldg(pc.rs_ + r);
        const tb2_rcp inv_sd = tb2_rcp_of(sd);
        // the cell left of this lane's chunk in the previous row (diagonal source of cell 0),
        // taken before chunks that left the band are recycled
        const double pm1 = __shfl_sync(TB2_FULL_MASK, x[CH - 1], left_lane);
        c_lo = cur_start / CH;
        if (chunk < c_lo) {
            // this lane's chunk is left of the band for good: take the chunk 32 further on
            chunk += 32;
            const int e0n = chunk * CH;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                x[i] = NEG;
                em[i] = (e0n + i < n_em) ? __ldg(em_g + e0n + i) : 0.0;
            }
            TB2_CNT(3, 1);
        }
        const int e0 = chunk * CH;
        const bool is_first = chunk == c_lo;                 // holds band position 0
        // valid cells of this lane: i in [lo, hi)
        const int lo = min(max(cur_start - e0, 0), CH);
        const int hi = min(max(cur_start + W - e0, 0), CH);
        const uint32_t vmask = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
        const uint32_t skmask = (is_first && d >= 1) ? (vmask & ~(1u << lo)) : vmask;
        uint32_t msk, mcd;
        if (cur_start + W > n_em) {
            // band reaches beyond the signal: those cells take the mask fill
            const int tl = min(max(n_em - e0, 0), CH);
            tb2_abs_walk<CH, true>(em, x, z, pm1, NEG, mu, sd, inv_sd, zs, mhz, stay, skip, maskval, vmask,
                                   skmask, ~((1u << tl) - 1u), msk, mcd);
        } else {
            tb2_abs_walk<CH, false>(em, x, z, pm1, NEG, mu, sd, inv_sd, zs, mhz, stay, skip, maskval, vmask,
                                    skmask, 0u, msk, mcd);
        }
        // ---- fix-up: lanes whose true left input is larger re-walk a prefix.  A larger
        // input can only turn cells into "stay": new value = max(stay', old), ties to stay ----
        double x_end = x[CH - 1], last_in = NEG;
        int n_rounds = 0, n_steps = 0;
        for (;;) {
            const double xin = __shfl_sync(TB2_FULL_MASK, x_end, left_lane);
            const bool need = !is_first && (xin > last_in);
            if (!__any_sync(TB2_FULL_MASK, need)) break;
            ++n_rounds;
            bool run = need;
            double xx = xin;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                // no early exit: in almost every round some lane right of the path walks its
                // whole chunk (measured: 11 of 13 steps), a vote per step costs more than it saves
                // (cells right of the band need no test: z = -inf there, so a = old = -inf, the
                // "tie" branch rewrites -inf over -inf and x_end does not change)
                ++n_steps;
                const double a = (xx - stay) + z[i];
                const double old = x[i];
                // stay now wins (ties go to stay).  On an exact tie the walk goes on: the cells
                // to the right are recomputed to the values they already hold (idempotent)
                // until one keeps its candidate -- cheaper than a second compare per step
                run = run && (a >= old);
                if (run) { x[i] = a; mcd &= ~(1u << i); }
                xx = a;
            }
            x_end = x[CH - 1];
            if (need) last_in = xin;
        }
        TB2_CNT(0, 1); TB2_CNT(1, n_rounds); TB2_CNT(2, n_steps);
#ifdef TB2_DP_COUNTERS
        TB2_CNT(4 + min(n_rounds, 3), 1);
#endif
        (void)n_steps; (void)n_rounds;
        // ---- moves out ----
        if (WPR == 1) tb[(size_t)(r - r_begin) * 32 + lane] = msk | (mcd << 16);
        else {
            tb[((size_t)(r - r_begin) * 2) * 32 + lane] = msk;
            tb[((size_t)(r - r_begin) * 2 + 1) * 32 + lane] = mcd;
        }
        // ---- first arg-max of the row (c_argmax :186-197): first local maximum with its
        // index, warp maximum, then the smallest event holding it (band order = event order) ----
        double best = x[0];
        int best_i = 0;
#pragma unroll
        for (int i = 1; i < CH; ++i)
            if (x[i] > best) { best = x[i]; best_i = i; }
        double wbest = best;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(TB2_FULL_MASK, wbest, off);
            wbest = (ob > wbest) ? ob : wbest;
        }
        const int win_e = __reduce_min_sync(TB2_FULL_MASK, (best == wbest) ? e0 + best_i : 0x7fffffff);
        last_argmax = win_e - cur_start;
        // a row of -inf only: the reference's arg-max is position 0
        if (last_argmax < 0 || last_argmax >= W) last_argmax = 0;
        prev_start = cur_start;
    }
    *argmax_io = last_argmax;
    return TB2_OK;
}

// traceback over rows [row_lo, row_hi) stored by tb2_adaptive_rows_abs (tb row 0 = fwd
// row row_lo); c_banded_traceback _c_dynamic_programming.pyx:281-310
template <int CH>
__device__ __noinline__ int tb2_tb_seg_abs(const uint32_t *tb, const int *starts, int row_hi, int row_lo,
                                           int W, int thresh, int *cur_event_io, int *read_tb)
{
    constexpr int WPR = (CH > 16) ? 2 : 1;
    const int lane = tb2_lane();
    int cur_event = *cur_event_io;
    int sp = row_hi;
    while (sp > row_lo) {
        const int nblk = min(4, sp - row_lo);
        uint32_t w[4][WPR];
        int st[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b < nblk) {
                const int row = sp - 1 - b;
#pragma unroll
                for (int k = 0; k < WPR; ++k)
                    w[b][k] = tb[((size_t)(row - row_lo) * WPR + k) * 32 + lane];
                st[b] = starts[row];
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b < nblk) {
                int bp = cur_event - st[b];
                if (bp < 0 || bp >= W) return TB2_ERR_UNEXPECTED;
                bool diag;
                for (;;) {
                    const int e = st[b] + bp;
                    const int ch = e / CH, i = e - ch * CH;
                    uint32_t ms = __shfl_sync(TB2_FULL_MASK, w[b][0], ch & 31), mc;
                    if (WPR == 2) mc = __shfl_sync(TB2_FULL_MASK, w[b][WPR - 1], ch & 31);
                    else { mc = ms >> 16; }
                    if ((mc >> i) & 1u) { diag = !((ms >> i) & 1u); break; }
                    --bp;                                   // stay in the current base
                    if (bp < 0) return TB2_ERR_UNEXPECTED;
                }
                if (diag) --bp;                             // diagonal
                if (thresh >= 0 && min(bp, W - bp - 1) < thresh) return TB2_ERR_BEYOND_BANDWIDTH;
                cur_event = st[b] + bp;
                if (lane == 0) read_tb[sp - 1 - b] = cur_event + 1;
            }
        }
        sp -= nblk;
    }
    *cur_event_io = cur_event;
    __syncwarp();
    return TB2_OK;
}

// ---------------------------------------------------------------------------
// Wide bands (bandwidth 1200, the save bandwidth 1500): NS chunks per lane.  The band
// spans up to 32*NS chunks; chunk c lives in slot c % (32*NS) = (slab (c/32) % NS, lane
// c % 32).  A row is NS passes in band order; in a pass the 32 lanes hold 32 consecutive
// chunks exactly as in the single-chunk engine (state comes from / goes back to the lane's
// slab in shared memory: [slab][cell][lane], conflict free), and the chain value leaving the
// pass's last chunk enters the next pass's first chunk exactly -- no speculation between
// passes.  Shared memory per warp: NS * CH * 32 doubles of row values.  The event means of a
// chunk are re-read from global memory in every pass (13-17 loads per lane, 104-136 bytes
// apart between lanes; the 3-4 KB a pass touches stay in L1 for the ~1000 rows the chunk is
// inside the band): keeping them in a second slab halved the resident warps (8 per SM) of a
// kernel that is latency-bound at this band width.
// ---------------------------------------------------------------------------
template <int CH, int NS>
__device__ __noinline__ int tb2_adaptive_rows_abs_ms(const PassCtx &pc, const DpConsts &c, int r_begin,
                                                     int r_end, int nb_total, double *st_s,
                                                     const double *rowbuf, uint32_t *tb, int *argmax_io)
{
    constexpr int WPR = (CH > 16) ? 2 : 1;
    constexpr int NSL = 32 * NS;
    const int lane = tb2_lane();
    const int W = pc.W, half_bw = W / 2, n_em = pc.n_em;
    const double NEG = tb2_neg_inf();
    const double stay = c.stay_pen, skip = c.skip_pen, zs = c.z_shift;
    const double mhz = c.winsor ? c.mhz : __longlong_as_double(0x7ff0000000000000LL);
    const double maskval = pc.mask_fill;
    const double *em_g = pc.em;
    const int left_lane = (lane + 31) & 31;
    double *x_s = st_s + lane;                                       // [(slab * CH + i) * 32]
    int prev_start = pc.starts[r_begin - 1];
    int last_argmax = *argmax_io;
    int c_lo_prev = prev_start / CH;
    {
        // initial state: fwd row r_begin from the wavefront row buffer, which shares the
        // shared memory with the slabs -- read everything first, then write
        double tmp[NS][CH];
        const int k = (lane - c_lo_prev) & 31;
#pragma unroll
        for (int p = 0; p < NS; ++p)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int j = (c_lo_prev + 32 * p + k) * CH + i - prev_start;
                tmp[p][i] = ((unsigned)j < (unsigned)W) ? rowbuf[j] : NEG;
            }
        __syncwarp();
#pragma unroll
        for (int p = 0; p < NS; ++p) {
            const int ch = c_lo_prev + 32 * p + k, sb = ((ch >> 5) % NS) * CH;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                x_s[(sb + i) * 32] = tmp[p][i];
            }
        }
        __syncwarp();
    }
    for (int r = r_begin; r < r_end; ++r) {
        int cur_start = prev_start + last_argmax - half_bw + 1;          // :344-358
        if (cur_start < prev_start) cur_start = prev_start;
        if (cur_start >= n_em) {
            if (r < nb_total - 2) return TB2_ERR_ADAPTIVE_BEYOND_SIGNAL;
            cur_start = n_em - 1;
        }
        if (lane == 0) pc.starts[r] = cur_start;
        const int d = cur_start - prev_start;
        const double mu = __ldg(pc.rm + r), sd = __ldg(pc.rs_ + r);
        const tb2_rcp inv_sd = tb2_rcp_of(sd);
        const int c_lo = cur_start / CH;
        const int k = (lane - c_lo) & 31;                  // position of this lane in a pass
        const int last_lane = (c_lo + 31) & 31;            // lane holding a pass's last chunk
        const bool tail = cur_start + W > n_em;
        // previous-row cell left of the band's first chunk (diagonal source of its cell 0):
        // still in its slot if that chunk was in the previous row's window, else outside
        double carry_pm1 = NEG;
        if (c_lo > c_lo_prev) {
            const int cp = c_lo - 1;
            carry_pm1 = st_s[((((cp >> 5) % NS) * CH + CH - 1) * 32) + (cp & 31)];
        }
        double carry_x = NEG;
        double lbest = NEG;
        int lbest_e = 0x7fffffff;
#pragma unroll
        for (int p = 0; p < NS; ++p) {
            const int ch = c_lo + 32 * p + k, sb = ((ch >> 5) % NS) * CH, e0 = ch * CH;
            const bool is_new = ch >= c_lo_prev + NSL;     // entered the window with this row
            double em[CH], x[CH], z[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                em[i] = (e0 + i < n_em) ? __ldg(em_g + e0 + i) : 0.0;
                x[i] = is_new ? NEG : x_s[(sb + i) * 32];
            }
            double pm1 = __shfl_sync(TB2_FULL_MASK, x[CH - 1], left_lane);
            const double next_pm1 = __shfl_sync(TB2_FULL_MASK, x[CH - 1], last_lane);
            const bool head = k == 0;                      // first chunk of the pass: exact inputs
            if (head) pm1 = carry_pm1;
            const int lo = min(max(cur_start - e0, 0), CH);
            const int hi = min(max(cur_start + W - e0, 0), CH);
            const uint32_t vmask = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
            const uint32_t skmask = (p == 0 && head && d >= 1) ? (vmask & ~(1u << lo)) : vmask;
            const double xc0 = head ? carry_x : NEG;
            uint32_t msk, mcd;
            if (tail) {
                const int tl = min(max(n_em - e0, 0), CH);
                tb2_abs_walk<CH, true>(em, x, z, pm1, xc0, mu, sd, inv_sd, zs, mhz, stay, skip, maskval,
                                       vmask, skmask, ~((1u << tl) - 1u), msk, mcd);
            } else {
                tb2_abs_walk<CH, false>(em, x, z, pm1, xc0, mu, sd, inv_sd, zs, mhz, stay, skip, maskval,
                                        vmask, skmask, 0u, msk, mcd);
            }
            double x_end = x[CH - 1], last_in = NEG;
            for (;;) {
                const double xin = __shfl_sync(TB2_FULL_MASK, x_end, left_lane);
                const bool need = !head && (xin > last_in);
                if (!__any_sync(TB2_FULL_MASK, need)) break;
                bool run = need;
                double xx = xin;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const double a = (xx - stay) + z[i];
                    const double old = x[i];
                    run = run && (a >= old);
                    if (run) { x[i] = a; mcd &= ~(1u << i); }
                    xx = a;
                }
                x_end = x[CH - 1];
                if (need) last_in = xin;
            }
#pragma unroll
            for (int i = 0; i < CH; ++i) x_s[(sb + i) * 32] = x[i];
            carry_x = __shfl_sync(TB2_FULL_MASK, x_end, last_lane);
            carry_pm1 = next_pm1;
            uint32_t *trow = tb + ((size_t)(r - r_begin) * NS + p) * WPR * 32 + lane;
            if (WPR == 1) trow[0] = msk | (mcd << 16);
            else { trow[0] = msk; trow[32] = mcd; }
#pragma unroll
            for (int i = 0; i < CH; ++i)
                if (x[i] > lbest) { lbest = x[i]; lbest_e = e0 + i; }
        }
        // first arg-max of the row: warp maximum, then the smallest event holding it
        double wbest = lbest;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(TB2_FULL_MASK, wbest, off);
            wbest = (ob > wbest) ? ob : wbest;
        }
        const int win_e = __reduce_min_sync(TB2_FULL_MASK, (lbest == wbest) ? lbest_e : 0x7fffffff);
        last_argmax = win_e - cur_start;
        if (last_argmax < 0 || last_argmax >= W) last_argmax = 0;
        prev_start = cur_start;
        c_lo_prev = c_lo;
        __syncwarp();
    }
    *argmax_io = last_argmax;
    return TB2_OK;
}

template <int CH, int NS>
__device__ __noinline__ int tb2_tb_seg_abs_ms(const uint32_t *tb, const int *starts, int row_hi, int row_lo,
                                              int W, int thresh, int *cur_event_io, int *read_tb)
{
    constexpr int WPR = (CH > 16) ? 2 : 1;
    const int lane = tb2_lane();
    int cur_event = *cur_event_io;
    for (int sp = row_hi; sp > row_lo; --sp) {
        const int row = sp - 1;
        const int st = starts[row];
        const int c_lo = st / CH;
        uint32_t w[NS][WPR];
#pragma unroll
        for (int p = 0; p < NS; ++p)
#pragma unroll
            for (int q = 0; q < WPR; ++q)
                w[p][q] = tb[(((size_t)(row - row_lo) * NS + p) * WPR + q) * 32 + lane];
        int bp = cur_event - st;
        if (bp < 0 || bp >= W) return TB2_ERR_UNEXPECTED;
        bool diag;
        for (;;) {
            const int e = st + bp;
            const int ch = e / CH, i = e - ch * CH, p = (ch - c_lo) >> 5;
            uint32_t a = w[0][0], b = w[0][WPR - 1];
#pragma unroll
            for (int q = 1; q < NS; ++q) if (p == q) { a = w[q][0]; b = w[q][WPR - 1]; }
            const uint32_t ms = __shfl_sync(TB2_FULL_MASK, a, ch & 31);
            uint32_t mc;
            if (WPR == 2) mc = __shfl_sync(TB2_FULL_MASK, b, ch & 31);
            else mc = ms >> 16;
            if ((mc >> i) & 1u) { diag = !((ms >> i) & 1u); break; }
            --bp;
            if (bp < 0) return TB2_ERR_UNEXPECTED;
        }
        if (diag) --bp;
        if (thresh >= 0 && min(bp, W - bp - 1) < thresh) return TB2_ERR_BEYOND_BANDWIDTH;
        cur_event = st + bp;
        if (lane == 0) read_tb[row] = cur_event + 1;
    }
    *cur_event_io = cur_event;
    __syncwarp();
    return TB2_OK;
}

__device__ int tb2_adaptive_rows_abs_ms_dyn(int ch, const PassCtx &pc, const DpConsts &c, int r_begin,
                                            int r_end, int nb_total, double *st_s, const double *rowbuf,
                                            uint32_t *tb, int *amax)
{
    switch (ch) {
    case 13: return tb2_adaptive_rows_abs_ms<13, TB2_ABS_MS_SLABS>(pc, c, r_begin, r_end, nb_total, st_s, rowbuf, tb, amax);
    case 17: return tb2_adaptive_rows_abs_ms<17, TB2_ABS_MS_SLABS>(pc, c, r_begin, r_end, nb_total, st_s, rowbuf, tb, amax);
    default: return TB2_ERR_CAPACITY;
    }
}

__device__ int tb2_tb_seg_abs_ms_dyn(int ch, const uint32_t *tb, const int *starts, int row_hi,
                                     int row_lo, int W, int thresh, int *cur_event, int *read_tb)
{
    switch (ch) {
    case 13: return tb2_tb_seg_abs_ms<13, TB2_ABS_MS_SLABS>(tb, starts, row_hi, row_lo, W, thresh, cur_event, read_tb);
    case 17: return tb2_tb_seg_abs_ms<17, TB2_ABS_MS_SLABS>(tb, starts, row_hi, row_lo, W, thresh, cur_event, read_tb);
    default: return TB2_ERR_CAPACITY;
    }
}

__device__ int tb2_adaptive_rows_abs_dyn(int ch, const PassCtx &pc, const DpConsts &c, int r_begin,
                                         int r_end, int nb_total, const double *rowbuf, uint32_t *tb,
                                         int *amax)
{
    switch (ch) {
    case 7: return tb2_adaptive_rows_abs<7>(pc, c, r_begin, r_end, nb_total, rowbuf, tb, amax);
    case 10: return tb2_adaptive_rows_abs<10>(pc, c, r_begin, r_end, nb_total, rowbuf, tb, amax);
    case 13: return tb2_adaptive_rows_abs<13>(pc, c, r_begin, r_end, nb_total, rowbuf, tb, amax);
    case 17: return tb2_adaptive_rows_abs<17>(pc, c, r_begin, r_end, nb_total, rowbuf, tb, amax);
    default: return TB2_ERR_CAPACITY;
    }
}

__device__ int tb2_tb_seg_abs_dyn(int ch, const uint32_t *tb, const int *starts, int row_hi,
                                  int row_lo, int W, int thresh, int *cur_event, int *read_tb)
{
    switch (ch) {
    case 7: return tb2_tb_seg_abs<7>(tb, starts, row_hi, row_lo, W, thresh, cur_event, read_tb);
    case 10: return tb2_tb_seg_abs<10>(tb, starts, row_hi, row_lo, W, thresh, cur_event, read_tb);
    case 13: return tb2_tb_seg_abs<13>(tb, starts, row_hi, row_lo, W, thresh, cur_event, read_tb);
    case 17: return tb2_tb_seg_abs<17>(tb, starts, row_hi, row_lo, W, thresh, cur_event, read_tb);
    default: return TB2_ERR_CAPACITY;
    }
}
