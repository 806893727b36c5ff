// dp_align_kernel.cuh -- the persistent warp-per-read assignment kernel (device code
// only, so that tests/emul can run the same source on the host)
#pragma once
#include "dp_align.cuh"
#include "kernels.h"

#define ALIGN_WARPS 4

// ---------------------------------------------------------------------------
// production kernel: persistent warps, one read per warp at a time
// ---------------------------------------------------------------------------
// KLASS 1 is the lean kernel for reads that can only take the static-band path
// (find_adaptive_base_assignment resquiggle.py:986-989): wavefront engine only, so
// fewer registers and twice the resident warps of the general kernel.
// resident CTAs per SM the register allocation is bounded for: 8 for the lean static-band
// kernel, 4 for the general one
template <int KLASS>
__global__ void __launch_bounds__(ALIGN_WARPS * 32, (KLASS == 1) ? 8 : 4)
k_align(AlignBatch b, AlignLaunchCfg cfg, uint32_t *tb_pool, double *grow_pool, int *counter)
{
    TB2_DYN_SMEM(double, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t slot = (size_t)blockIdx.x * ALIGN_WARPS + warp;
    WarpRes wr;
    wr.smem_rows = smem + (size_t)warp * (2 * cfg.smem_cells + TB2_WF_RING);
    wr.smem_cap = 2 * cfg.smem_cells;
    wr.ring = wr.smem_rows + 2 * cfg.smem_cells;
    wr.grow = cfg.grow_cells > 0 ? grow_pool + slot * 2 * (size_t)cfg.grow_cells : nullptr;
    wr.grow_cap = cfg.grow_cells;
    wr.tb = tb_pool + slot * cfg.tb_words;
    wr.tb_words = cfg.tb_words;
    for (;;) {
        int r = 0;
        if (lane == 0) r = atomicAdd(counter, 1);
        r = __shfl_sync(TB2_FULL_MASK, r, 0);
        if (r >= b.n_reads) break;
        if (b.order) r = b.order[r];
        const size_t ix = (size_t)r * b.stride;
        if (b.status[ix] != TB2_OK) continue;
        if (b.active && !b.active[ix]) continue;
        if (KLASS != 0) {
            const int nbr = (int)(b.base_off[r + 1] - b.base_off[r]);
            const bool is_short = (b.num_events[ix] - 1 < b.params.start_bw + b.params.start_n_bases) ||
                                  (nbr < b.params.start_n_bases);
            if (is_short != (KLASS == 1)) continue;
        }
        AlignRead a;
        const long long eo = b.ev_off[r], bo = b.base_off[r];
        a.cpts = b.cpts + eo;
        a.n_cpts = b.n_cpts[ix];
        a.em = b.em + eo;
        a.rm = b.rm + bo;
        a.rs = b.rs + bo;
        a.nb = (int)(b.base_off[r + 1] - bo);
        a.starts = b.starts + bo;
        a.read_tb = b.read_tb + bo + r;
        a.segs = b.segs + bo + r;
        a.rsrtr = b.rsrtr + ix;
        a.dbg = b.dbg ? b.dbg + 3 * (size_t)r : nullptr;
        int st;
        if (KLASS == 1) {
            DpConsts c;
            c.z_shift = b.params.z_shift; c.stay_pen = b.params.stay_pen;
            c.skip_pen = b.params.skip_pen;
            c.winsor = !isnan(b.params.max_half_z_score);
            c.mhz = c.winsor ? b.params.max_half_z_score : 0.0;
            if (a.dbg && lane == 0) { a.dbg[0] = 0; a.dbg[1] = -1; a.dbg[2] = -1; }
            st = (a.n_cpts - 1 < 1 || a.nb < 1) ? TB2_ERR_UNEXPECTED : tb2_static_assign(a, wr, c);
        } else {
            st = tb2_align_read(a, wr, b.params, b.sig_match_thresh);
        }
        __syncwarp();
        if (lane == 0) b.status[ix] = st;
    }
}

