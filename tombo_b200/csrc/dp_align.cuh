// dp_align.cuh -- event -> sequence assignment of one read by one warp:
// find_adaptive_base_assignment (resquiggle.py:866-1050, start_clip_bases=None)
// with find_seq_start_in_events (:685-752), _get_masked_start_fwd_pass (:607-683),
// find_static_base_assignment (:547-600), c_adaptive_banded_forward_pass,
// c_banded_traceback, _trim_traceback (:754-764), get_rel_raw_coords (:858-864).
//
// Static bands (short reads, start search, masked start) run on the wavefront
// engine, adaptive rows on the lane-chunk engine (dp_row.cuh).
#pragma once
#include "dp_row2.cuh"

// packed-move words per lane (lane-chunk layout) are instantiated for
// {1,2,3,4,5,8,16}: band widths up to 16*16*32 = 8192 cells
#define TB2_MAX_WPL 16
#define TB2_MASK_FILL_Z_SCORE (-15.0)

// per-warp resources
struct WarpRes {
    double *smem_rows;    // smem_cap doubles of shared memory
    int smem_cap;
    double *ring;         // TB2_WF_RING doubles of shared memory (wavefront exchange)
    double *grow;         // optional global row scratch, 2 * grow_cap doubles
    int grow_cap;
    uint32_t *tb;         // packed move scratch
    size_t tb_words;      // capacity in words
};

// one read's view
struct AlignRead {
    const int *cpts;      // valid changepoints (n_cpts)
    int n_cpts;
    const double *em;     // event means (n_cpts - 1)
    const double *rm, *rs;  // reference levels (nb)
    int nb;
    int *starts;          // scratch, nb ints
    int *read_tb;         // scratch, nb + 1 ints
    int *segs;            // out, nb + 1
    int *rsrtr;           // out
    int *dbg;             // out (may be null): path, mapped_start, clip
};

// two lane-transposed row buffers for the lane-chunk engine
__device__ __forceinline__ bool tb2_setup_geom(PassCtx &pc, const WarpRes &wr, int W)
{
    pc.W = W;
    pc.chunk = (W + 31) / 32;
    const int cells = pc.chunk * 32;
    pc.zbuf = nullptr; pc.cbuf = nullptr;
    if (2 * cells <= wr.smem_cap) {
        pc.buf0 = wr.smem_rows;
        pc.buf1 = wr.smem_rows + cells;
        if (4 * cells <= wr.smem_cap) { pc.zbuf = wr.smem_rows + 2 * cells; pc.cbuf = wr.smem_rows + 3 * cells; }
        return true;
    }
    if (wr.grow != nullptr && cells <= wr.grow_cap) {
        pc.buf0 = wr.grow;
        pc.buf1 = wr.grow + wr.grow_cap;
        return true;
    }
    return false;
}

// one plain row buffer (W doubles) for the wavefront engine
__device__ __forceinline__ double *tb2_wf_rowbuf(const WarpRes &wr, int W)
{
    if (W <= wr.smem_cap) return wr.smem_rows;
    if (wr.grow != nullptr && W <= 2 * wr.grow_cap) return wr.grow;
    return nullptr;
}

__device__ __forceinline__ int tb2_wpl_of(int chunk)
{
    const int w = (chunk + 15) / 16;
    return w <= 5 ? w : (w <= 8 ? 8 : (w <= 16 ? 16 : w));
}

__device__ int tb2_run_rows_dyn(int wpl, const PassCtx &pc, const DpConsts &c, int mode,
                                int r_begin, int r_end, int nb_total, int *cur_sel, int *amax)
{
    switch (wpl) {
    case 1: return tb2_run_rows<1>(pc, c, mode, r_begin, r_end, nb_total, cur_sel, amax);
    case 2: return tb2_run_rows<2>(pc, c, mode, r_begin, r_end, nb_total, cur_sel, amax);
    case 3: return tb2_run_rows<3>(pc, c, mode, r_begin, r_end, nb_total, cur_sel, amax);
    case 4: return tb2_run_rows<4>(pc, c, mode, r_begin, r_end, nb_total, cur_sel, amax);
    case 5: return tb2_run_rows<5>(pc, c, mode, r_begin, r_end, nb_total, cur_sel, amax);
    case 8: return tb2_run_rows<8>(pc, c, mode, r_begin, r_end, nb_total, cur_sel, amax);
    case 16: return tb2_run_rows<16>(pc, c, mode, r_begin, r_end, nb_total, cur_sel, amax);
    default: return TB2_ERR_CAPACITY;
    }
}

__device__ int tb2_tb_seg_chunk_dyn(int wpl, const uint32_t *tb, const int *starts, int row_hi,
                                    int row_lo, int W, int chunk, int thresh, int *cur_event,
                                    int *read_tb)
{
    switch (wpl) {
    case 1: return tb2_tb_seg_chunk<1>(tb, starts, row_hi, row_lo, W, chunk, thresh, cur_event, read_tb);
    case 2: return tb2_tb_seg_chunk<2>(tb, starts, row_hi, row_lo, W, chunk, thresh, cur_event, read_tb);
    case 3: return tb2_tb_seg_chunk<3>(tb, starts, row_hi, row_lo, W, chunk, thresh, cur_event, read_tb);
    case 4: return tb2_tb_seg_chunk<4>(tb, starts, row_hi, row_lo, W, chunk, thresh, cur_event, read_tb);
    case 5: return tb2_tb_seg_chunk<5>(tb, starts, row_hi, row_lo, W, chunk, thresh, cur_event, read_tb);
    case 8: return tb2_tb_seg_chunk<8>(tb, starts, row_hi, row_lo, W, chunk, thresh, cur_event, read_tb);
    case 16: return tb2_tb_seg_chunk<16>(tb, starts, row_hi, row_lo, W, chunk, thresh, cur_event, read_tb);
    default: return TB2_ERR_CAPACITY;
    }
}

__device__ __forceinline__ void tb2_pc_defaults(PassCtx &pc, const AlignRead &a,
                                                const WarpRes &wr, const DpConsts &c)
{
    pc.em = a.em; pc.n_em = a.n_cpts - 1;
    pc.rm = a.rm; pc.rs_ = a.rs; pc.zmat = nullptr;
    pc.mso = 0; pc.msp_start = 0; pc.msp_stop = 0;
    pc.mask_fill = TB2_MASK_FILL_Z_SCORE;
    pc.mask_shifted = (TB2_MASK_FILL_Z_SCORE - c.z_shift) + c.z_shift;  // resquiggle.py:666,678
    pc.starts = a.starts; pc.tb = wr.tb; pc.dbg_fwd = nullptr; pc.dbg_tb = nullptr;
    pc.buf0 = nullptr; pc.buf1 = nullptr; pc.zbuf = nullptr; pc.cbuf = nullptr; pc.chunk = 0; pc.W = 0;
    pc.ring = wr.ring;
}

// static-band forward pass + traceback over rows [0, n_rows) (wavefront engine)
__device__ int tb2_static_pass(PassCtx &pc, const WarpRes &wr, const DpConsts &c, int W,
                               int n_rows, int *read_tb)
{
    pc.W = W;
    double *rowbuf = tb2_wf_rowbuf(wr, W);
    if (rowbuf == nullptr) return TB2_ERR_CAPACITY;
    const long long words = tb2_wf_total_words(pc.starts, n_rows, W);
    if (words < 0) return TB2_ERR_UNEXPECTED;
    if ((size_t)words > wr.tb_words) return TB2_ERR_CAPACITY;
    int amax = 0;
    int st = tb2_wavefront_rows(pc, c, TB2_MODE_PLAIN, n_rows, rowbuf, wr.tb, &amax);
    if (st != TB2_OK) return st;
    int cur_event = amax + pc.starts[n_rows - 1];
    if (tb2_lane() == 0) read_tb[n_rows] = cur_event + 1;
    return tb2_tb_seg_wf(wr.tb, words, pc.starts, n_rows, W, -1, &cur_event, read_tb);
}

// find_seq_start_in_events resquiggle.py:685-752
__device__ __noinline__ int tb2_start_find(const AlignRead &a, const WarpRes &wr, const DpConsts &c,
                              int num_bases, int num_events, bool check_score,
                              double sig_match_thresh, int *start_loc, double *epb)
{
    const int lane = tb2_lane();
    const int n_em = a.n_cpts - 1;
    if (n_em < num_events + num_bases) return TB2_ERR_READ_TOO_SHORT_START;
    if (a.nb < num_bases) return TB2_ERR_MAP_TOO_SHORT_START;
    PassCtx pc;
    tb2_pc_defaults(pc, a, wr, c);
    for (int r = lane; r < num_bases; r += 32) a.starts[r] = r;  // :721
    __syncwarp();
    int st = tb2_static_pass(pc, wr, c, num_events, num_bases, a.read_tb);
    if (st != TB2_OK) return st;
    // scoring scratch (one double per base): the now free row buffer, or -- start windows
    // narrower than the start bases, never a default -- the move scratch, free as well after
    // the traceback
    double *t = tb2_wf_rowbuf(wr, num_events);
    if (num_bases + 1 > num_events) {
        t = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(wr.tb) + 7) & ~(uintptr_t)7);
        if ((size_t)(num_bases + 2) * 2 > wr.tb_words) return TB2_ERR_CAPACITY;
    }
    int sloc = 0;
    double e = 0;
    if (lane == 0) {
        sloc = a.read_tb[0];
        e = (double)(a.read_tb[num_bases] - a.read_tb[0]) / (double)(num_bases + 1);  // :749
        if (check_score) {
            // score_valid_bases tombo_stats.py:2340-2362 (np.mean = pairwise sum / n)
            int nv = 0;
            for (int i = 0; i < num_bases; ++i) {
                const int s0 = a.read_tb[i], s1 = a.read_tb[i + 1];
                if (s1 != s0) {
                    const double m = tb2_pairwise_sum(a.em + s0, s1 - s0) / (double)(s1 - s0);
                    t[nv++] = fabs((m - a.rm[i]) / a.rs[i]);
                }
            }
            if (nv == 0) st = TB2_ERR_INVALID_START_PATH;
            else if (tb2_pairwise_sum(t, nv) / (double)nv > sig_match_thresh)
                st = TB2_ERR_POOR_START_MATCH;
        }
    }
    st = __shfl_sync(TB2_FULL_MASK, st, 0);
    *start_loc = __shfl_sync(TB2_FULL_MASK, sloc, 0);
    *epb = __shfl_sync(TB2_FULL_MASK, e, 0);
    __syncwarp();
    return st;
}

// cpts[read_tb] - cpts[read_tb[0]]  (get_rel_raw_coords resquiggle.py:858-864)
__device__ int tb2_emit_segs(const AlignRead &a, const int *cpts, int n_idx)
{
    const int lane = tb2_lane();
    const int nb = a.nb;
    int bad = 0;
    const int t0 = a.read_tb[0];
    const int base = (t0 >= 0 && t0 < n_idx) ? cpts[t0] : 0;
    for (int i = lane; i <= nb; i += 32) {
        const int t = a.read_tb[i];
        if (t < 0 || t >= n_idx) { bad = 1; continue; }
        a.segs[i] = cpts[t] - base;
    }
    if (__any_sync(TB2_FULL_MASK, bad) || t0 < 0 || t0 >= n_idx) return TB2_ERR_UNEXPECTED;
    if (lane == 0) *a.rsrtr = base;
    return TB2_OK;
}

// find_static_base_assignment resquiggle.py:547-600 + get_short_read_results
__device__ __noinline__ int tb2_static_assign(const AlignRead &a, const WarpRes &wr, const DpConsts &c,
                                 bool emit_segs = true)
{
    const int lane = tb2_lane();
    const int n_em = a.n_cpts - 1, nb = a.nb;
    const int mask_len = min(nb, n_em) / 4;
    const int W = n_em - mask_len;
    if (W <= 0 || nb <= 0) return TB2_ERR_UNEXPECTED;
    PassCtx pc;
    tb2_pc_defaults(pc, a, wr, c);
    const int n0 = nb - 2 * mask_len;
    for (int r = lane; r < nb; r += 32)   // :567-569
        a.starts[r] = (r < n0) ? 0
                               : (int)tb2_linspace_at(0.0, (double)mask_len, 2 * mask_len, r - n0);
    __syncwarp();
    int st = tb2_static_pass(pc, wr, c, W, nb, a.read_tb);
    if (st != TB2_OK) return st;
    if (!emit_segs) return TB2_OK;
    return tb2_emit_segs(a, a.cpts, a.n_cpts);
}

__device__ int tb2_align_read(const AlignRead &a, const WarpRes &wr, const tb2_params &p,
                              double sig_match_thresh)
{
    const int lane = tb2_lane();
    DpConsts c;
    c.z_shift = p.z_shift; c.stay_pen = p.stay_pen; c.skip_pen = p.skip_pen;
    c.winsor = !isnan(p.max_half_z_score);
    c.mhz = c.winsor ? p.max_half_z_score : 0.0;
    const int n_em = a.n_cpts - 1, nb = a.nb;
    if (n_em < 1 || nb < 1) return TB2_ERR_UNEXPECTED;
    const int start_bw = (int)p.start_bw, start_save_bw = (int)p.start_save_bw,
              start_n = (int)p.start_n_bases, bw = (int)p.bandwidth;
    if (a.dbg && lane == 0) { a.dbg[0] = 0; a.dbg[1] = -1; a.dbg[2] = -1; }
    // short reads: one static band over the whole read (:986-989)
    if (n_em < start_bw + start_n || nb < start_n) return tb2_static_assign(a, wr, c);
    int mapped_start = 0;
    double epb = 0;
    int st = tb2_start_find(a, wr, c, start_n, start_bw, true, sig_match_thresh, &mapped_start,
                            &epb);
    if (st != TB2_OK && st != TB2_ERR_UNEXPECTED && st != TB2_ERR_CAPACITY) {  // except TomboError
        if (n_em < start_save_bw + start_n) return tb2_static_assign(a, wr, c);
        st = tb2_start_find(a, wr, c, start_n, start_save_bw, false, 0.0, &mapped_start, &epb);
    }
    if (st != TB2_OK) return st;
    if (epb == 0) return TB2_ERR_OPEN_PORE;  // :1008
    const int half_bw = bw / 2;
    int clip, mso;
    if (mapped_start < half_bw) { clip = 0; mso = mapped_start; }
    else { clip = mapped_start - half_bw; mso = half_bw; }
    if (a.dbg && lane == 0) { a.dbg[1] = mapped_start; a.dbg[2] = clip; }
    if ((int)((double)(half_bw + 1) / epb) >= nb || (n_em - mso - clip < bw))  // :1024-1027
        return tb2_static_assign(a, wr, c);
    if (a.dbg && lane == 0) a.dbg[0] = 1;

    // ---- _get_masked_start_fwd_pass :607-683 (static rows: wavefront engine) ----
    const int n_emc = n_em - clip;
    if (n_emc - mso < bw) return TB2_ERR_START_TOO_FAR;
    PassCtx pc;
    tb2_pc_defaults(pc, a, wr, c);
    pc.em = a.em + clip; pc.n_em = n_emc; pc.mso = mso;
    // adaptive rows: register engine (dp_row2.cuh) for bands up to 528 cells, the
    // shared-memory lane-chunk engine for wider ones
    const int ach = tb2_abs_chunk(bw);
    // three chunks per lane for bands up to 1616 cells when the warp's shared memory holds
    // the slabs (3 * CH * 32 doubles)
    int mch = ach ? 0 : tb2_abs_ms_chunk_host(bw);
    if (mch && (wr.smem_cap < TB2_ABS_MS_SLABS * mch * 32 || !__isShared(wr.smem_rows))) mch = 0;
    int wpl;
    if (ach || mch) {
        pc.W = bw; pc.chunk = 0;
        pc.buf0 = tb2_wf_rowbuf(wr, bw);
        if (pc.buf0 == nullptr) return TB2_ERR_CAPACITY;
        wpl = tb2_abs_words_per_row(bw);
    } else {
        if (!tb2_setup_geom(pc, wr, bw)) return TB2_ERR_CAPACITY;
        wpl = tb2_wpl_of(pc.chunk);
        if (wpl > TB2_MAX_WPL) return TB2_ERR_CAPACITY;
    }
    const int bes0 = (half_bw <= mso) ? 0 : mso - half_bw;
    const int t2 = (int)((double)(half_bw + 1) / epb);
    const int tmp_len = max(max(half_bw, TB2_MASK_BASES), t2) + 1;
    const double ls_stop = (double)bes0 + ((double)tmp_len * epb);
    int first = 0x7fffffff;
    for (int i = lane; i < tmp_len; i += 32) {
        const int v = (int)tb2_linspace_at((double)bes0, ls_stop, tmp_len, i);
        if (i < nb) a.starts[i] = v;
        if (v >= mso && i < first) first = i;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        first = min(first, __shfl_xor_sync(TB2_FULL_MASK, first, off));
    __syncwarp();
    if (first == 0x7fffffff) return TB2_ERR_UNEXPECTED;
    int mask_seq_len = max(TB2_MASK_BASES, first + 2);
    if (mask_seq_len > tmp_len) mask_seq_len = tmp_len;
    if (mask_seq_len > nb) return TB2_ERR_UNEXPECTED;
    // move scratch: wavefront rows first, then the lane-chunk rows (indexed by row)
    const long long wf_words_ll = tb2_wf_total_words(a.starts, mask_seq_len, bw);
    if (wf_words_ll < 0) return TB2_ERR_UNEXPECTED;
    const size_t wf_words = (size_t)wf_words_ll;
    if (wf_words + (size_t)nb * wpl * 32 > wr.tb_words) return TB2_ERR_CAPACITY;
    uint32_t *tb_wf = wr.tb, *tb_chunk = wr.tb + wf_words;
    pc.tb = tb_chunk;
    pc.msp_start = (double)(mso + 1);
    pc.msp_stop = (double)(a.starts[TB2_MASK_BASES - 1] + bw);
    // validate every masked row (the reference raises from inside the loop)
    int bad = 0;
    for (int r = lane; r < mask_seq_len; r += 32) {
        const int ep = a.starts[r];
        const int sml = max(mso - ep, 0);
        int eml = 0;
        if (r < TB2_MASK_BASES)
            eml = bw - ((int)tb2_linspace_at(pc.msp_start, pc.msp_stop, TB2_MASK_BASES, r) - ep);
        if (ep + bw - eml > n_emc) eml = ep + bw - n_emc;
        const int aa = ep + sml;
        int bb = ep + bw - eml;
        if (bb > n_emc) bb = n_emc;
        const int nv = max(bb - aa, 0);
        if (aa < 0 || eml < 0 || sml + nv + eml != bw) bad = 1;
    }
    if (__any_sync(TB2_FULL_MASK, bad)) return TB2_ERR_MASKED_TOO_FEW;
    // the wavefront row buffer is the first half of the two lane-chunk buffers
    double *rowbuf = pc.buf0;
    int amax = 0;
    st = tb2_wavefront_rows(pc, c, TB2_MODE_MASKED, mask_seq_len, rowbuf, tb_wf, &amax);
    if (st != TB2_OK) return st;
    const int thresh = (int)p.band_bound_thresh;
    int cur_event;
    if (ach) {
        // ---- adaptive rows :314-412 (register engine; moves indexed from mask_seq_len) ----
        st = tb2_adaptive_rows_abs_dyn(ach, pc, c, mask_seq_len, nb, nb, rowbuf, tb_chunk, &amax);
        if (st != TB2_OK) return st;
        __syncwarp();
        cur_event = amax + a.starts[nb - 1];
        if (lane == 0) a.read_tb[nb] = cur_event + 1;
        st = tb2_tb_seg_abs_dyn(ach, tb_chunk, a.starts, nb, mask_seq_len, bw, thresh, &cur_event,
                                a.read_tb);
        if (st != TB2_OK) return st;
    } else if (mch) {
        // ---- adaptive rows, wide band: three chunks per lane, state in shared memory ----
        st = tb2_adaptive_rows_abs_ms_dyn(mch, pc, c, mask_seq_len, nb, nb, wr.smem_rows, rowbuf,
                                          tb_chunk, &amax);
        if (st != TB2_OK) return st;
        __syncwarp();
        cur_event = amax + a.starts[nb - 1];
        if (lane == 0) a.read_tb[nb] = cur_event + 1;
        st = tb2_tb_seg_abs_ms_dyn(mch, tb_chunk, a.starts, nb, mask_seq_len, bw, thresh, &cur_event,
                                   a.read_tb);
        if (st != TB2_OK) return st;
    } else {
        // last masked row -> lane-transposed buffer 1 (source and destination disjoint)
        for (int j = lane; j < bw; j += 32) {
            const int lj = j / pc.chunk;
            pc.buf1[(j - lj * pc.chunk) * 32 + lj] = rowbuf[j];
        }
        __syncwarp();
        int sel = 1;
        // ---- adaptive rows :314-412 (lane-chunk engine) ----
        st = tb2_run_rows_dyn(wpl, pc, c, TB2_MODE_ADAPTIVE, mask_seq_len, nb, nb, &sel, &amax);
        if (st != TB2_OK) return st;
        cur_event = amax + a.starts[nb - 1];
        if (lane == 0) a.read_tb[nb] = cur_event + 1;
        st = tb2_tb_seg_chunk_dyn(wpl, tb_chunk, a.starts, nb, mask_seq_len, bw, pc.chunk, thresh,
                                  &cur_event, a.read_tb);
        if (st != TB2_OK) return st;
    }
    st = tb2_tb_seg_wf(tb_wf, wf_words_ll, a.starts, mask_seq_len, bw, thresh, &cur_event, a.read_tb);
    if (st != TB2_OK) return st;
    // _trim_traceback :754-764
    if (lane == 0) {
        int i = 0;
        while (i <= nb && a.read_tb[i] < 0) a.read_tb[i++] = 0;
        int e = 1;
        while (e <= nb + 1 && a.read_tb[nb + 1 - e] > n_emc) { a.read_tb[nb + 1 - e] = n_emc; ++e; }
    }
    __syncwarp();
    return tb2_emit_segs(a, a.cpts + clip, a.n_cpts - clip);
}
