// llr.cu -- per-read alternative-model log-likelihood ratios:
// compute_alt_model_read_stats tombo_stats.py:3972-4082 (whole read, '+' strand,
// single-base motif TomboMotif(alt_base, 1)) with trim_seq_and_means (:3888-3970),
// c_calc_scaled_llh_ratio_const_var _c_helper.pyx:313-358 (default) and
// c_calc_llh_ratio_const_var :298-311.  Also c_new_mean_stds :38-57.
#include "batch.h"
#include <vector>

namespace {
enum { L_MEAN = 50, L_MOFF, L_SEQ, L_SOFF, L_START, L_CNT, L_SITEOFF, L_LLR, L_POS, L_A, L_B, L_C, L_D };

struct LlrArgs {
    int n_reads, K, cpos, alt_code, use_std;
    double sf, hf, hp;
    const double *norm_mean;
    const long long *mean_off, *seq_off, *read_start;
    const unsigned char *seq;
    const double *kmeans, *ksds, *alt;   // alt[code * K + pos]
    const int *status;                   // resident batch: reads that failed hold no sites
    int status_stride;
};

__device__ __forceinline__ int kmer_code(const unsigned char *bases, int K)
{
    int c = 0;
    for (int j = 0; j < K; ++j) c = c * 4 + (bases[j] & 3);
    return c;
}

// FILL = false: count sites per read; FILL = true: write llr / pos
template <bool FILL>
__global__ void __launch_bounds__(256)
k_llr(LlrArgs a, int *counts, const long long *site_off, double *llr_out, long long *pos_out)
{
    __shared__ unsigned int warp_tot[8];
    const int r = blockIdx.x, tid = threadIdx.x;
    const long long mo = a.mean_off[r];
    const int nb = (int)(a.mean_off[r + 1] - mo);
    const int K = a.K;
    // trimmed read sequence: base i = seq[so + cpos + i]
    const unsigned char *bases = a.seq + a.seq_off[r] + a.cpos;
    const double *means = a.norm_mean + mo;
    int testable = nb - 2 * (K - 1);            // len(motif_search_seq)
    if (a.status && a.status[(size_t)r * a.status_stride] != TB2_OK) testable = 0;
    if (testable <= 0) { if (!FILL && tid == 0) counts[r] = 0; return; }
    const int per = (testable + 255) / 256;
    const int i0 = min(testable, tid * per), i1 = min(testable, i0 + per);
    unsigned int mine = 0;
    for (int i = i0; i < i1; ++i) mine += (bases[i + K - 1] == a.alt_code);
    const int lane = tid & 31, warp = tid >> 5;
    unsigned int inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned int o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    unsigned int base = 0, total = 0;
    for (int q = 0; q < 8; ++q) { if (q < warp) base += warp_tot[q]; total += warp_tot[q]; }
    if (!FILL) { if (tid == 0) counts[r] = (int)total; return; }
    long long o = site_off[r] + base + inc - mine;
    for (int i = i0; i < i1; ++i) {
        if (bases[i + K - 1] != a.alt_code) continue;
        // alt_pos = i: k-mers i .. i+K-1 of the trimmed sequence, means' = means[cpos + .]
        const double const_var = a.ksds[kmer_code(bases + i, K)];
        const double cv = const_var * const_var;                 // np.square(r_ref_sds)[alt_pos]
        double acc = 0.0;
        for (int t = 0; t < K; ++t) {
            const int code = kmer_code(bases + i + t, K);
            const double obs = means[a.cpos + i + t];
            const double ref_mean = a.kmeans[code];
            const double alt_mean = a.alt[(size_t)code * K + (K - 1 - t)];
            if (a.use_std) {
                const double rd = obs - ref_mean, ad = obs - alt_mean;
                acc += ((ad * ad) - (rd * rd)) / cv;
            } else {
                if (ref_mean == alt_mean) continue;
                const double scale_mean = (alt_mean + ref_mean) / 2;
                const double ref_diff = obs - ref_mean, alt_diff = obs - alt_mean;
                const double scale_diff = obs - scale_mean;
                double means_diff = alt_mean - ref_mean;
                if (means_diff < 0) means_diff = means_diff * -1;
                acc += exp(-(scale_diff * scale_diff) / (a.sf * cv)) *
                       ((alt_diff * alt_diff) - (ref_diff * ref_diff)) /
                       (cv * pow(means_diff, a.hp) * a.hf);
            }
        }
        llr_out[o] = acc;
        pos_out[o] = a.read_start[r] + (K - 1) + i;
        ++o;
    }
}

__global__ void k_mean_stds(const double *sig, const long long *segs, long long n_segs,
                            double *means, double *sds)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_segs) return;
    const long long a = segs[i], z = segs[i + 1];
    double s = 0, v = 0;
    for (long long k = a; k < z; ++k) s += sig[k];
    const double m = s / (double)(z - a);
    means[i] = m;
    for (long long k = a; k < z; ++k) { const double d = sig[k] - m; v += d * d; }
    sds[i] = sqrt(v / (double)(z - a));
}
}  // namespace

extern "C" int tb2_set_alt_model(tb2_ctx *ctx, const double *alt_means, int kmer_width)
{
    int rc = tb2_use(ctx);
    if (rc) return rc;
    if (!alt_means || kmer_width < 1 || kmer_width > 12) return TB2_ERR_INVALID_ARG;
    const size_t n = ((size_t)1 << (2 * kmer_width)) * kmer_width;
    TB2_CUDA_TRY(ctx, ctx->alt_means.reserve(n * 8));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->alt_means.p, alt_means, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    TB2_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->alt_kmer_width = kmer_width;
    return TB2_OK;
}

extern "C" int tb2_alt_model_llr_batch(tb2_ctx *ctx, int64_t n_reads, const double *norm_mean,
                                       const int64_t *mean_off, const uint8_t *seq,
                                       const int64_t *seq_off, const int64_t *read_start,
                                       int alt_base_code, int use_standard_llhr,
                                       double scale_factor, double height_factor,
                                       double height_power, double *llr_out, int64_t *pos_out,
                                       int64_t *site_off)
{
    int rc = tb2_use(ctx);
    if (rc) return rc;
    if (n_reads < 0 || !mean_off || !seq_off || !read_start || !site_off || alt_base_code < 0 ||
        alt_base_code > 3)
        return TB2_ERR_INVALID_ARG;
    if (ctx->kmer_width <= 0 || ctx->alt_kmer_width != ctx->kmer_width) {
        ctx->err = "standard and alternative models must be set with the same k-mer width";
        return TB2_ERR_INVALID_ARG;
    }
    site_off[0] = 0;
    if (n_reads == 0) return TB2_OK;
    if (!norm_mean || !seq || !llr_out || !pos_out) return TB2_ERR_INVALID_ARG;
    const int n = (int)n_reads;
    auto &P = ctx->pool;
    cudaStream_t s = ctx->stream;
    const size_t tm = (size_t)mean_off[n], ts = (size_t)seq_off[n];
    TB2_CUDA_TRY(ctx, P[L_MEAN].reserve(tm * 8 + 8));
    TB2_CUDA_TRY(ctx, P[L_MOFF].reserve((n + 1) * 8));
    TB2_CUDA_TRY(ctx, P[L_SEQ].reserve(ts + 8));
    TB2_CUDA_TRY(ctx, P[L_SOFF].reserve((n + 1) * 8));
    TB2_CUDA_TRY(ctx, P[L_START].reserve((size_t)n * 8));
    TB2_CUDA_TRY(ctx, P[L_CNT].reserve((size_t)n * 4));
    TB2_CUDA_TRY(ctx, P[L_SITEOFF].reserve((n + 1) * 8));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_MEAN].p, norm_mean, tm * 8, cudaMemcpyHostToDevice, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_MOFF].p, mean_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_SEQ].p, seq, ts, cudaMemcpyHostToDevice, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_SOFF].p, seq_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_START].p, read_start, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    LlrArgs a;
    a.status = nullptr; a.status_stride = 0;
    a.n_reads = n; a.K = ctx->kmer_width; a.cpos = ctx->central_pos; a.alt_code = alt_base_code;
    a.use_std = use_standard_llhr ? 1 : 0;
    a.sf = scale_factor; a.hf = height_factor; a.hp = height_power;
    a.norm_mean = P[L_MEAN].as<double>();
    a.mean_off = P[L_MOFF].as<long long>();
    a.seq_off = P[L_SOFF].as<long long>();
    a.read_start = P[L_START].as<long long>();
    a.seq = P[L_SEQ].as<unsigned char>();
    a.kmeans = ctx->model_means.as<double>();
    a.ksds = ctx->model_sds.as<double>();
    a.alt = ctx->alt_means.as<double>();
    k_llr<false><<<n, 256, 0, s>>>(a, P[L_CNT].as<int>(), nullptr, nullptr, nullptr);
    TB2_CHECK_LAUNCH(ctx);
    std::vector<int> cnt((size_t)n);
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(cnt.data(), P[L_CNT].p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    TB2_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    for (int r = 0; r < n; ++r) site_off[r + 1] = site_off[r] + cnt[r];
    const size_t total = (size_t)site_off[n];
    if (total == 0) return TB2_OK;
    TB2_CUDA_TRY(ctx, P[L_LLR].reserve(total * 8));
    TB2_CUDA_TRY(ctx, P[L_POS].reserve(total * 8));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_SITEOFF].p, site_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    k_llr<true><<<n, 256, 0, s>>>(a, nullptr, P[L_SITEOFF].as<long long>(), P[L_LLR].as<double>(),
                                  P[L_POS].as<long long>());
    TB2_CHECK_LAUNCH(ctx);
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(llr_out, P[L_LLR].p, total * 8, cudaMemcpyDeviceToHost, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(pos_out, P[L_POS].p, total * 8, cudaMemcpyDeviceToHost, s));
    TB2_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return TB2_OK;
}

// ---------------------------------------------------------------------------
// the same scoring on the RESIDENT batch (after tb2_batch_compute): sequence, per-base
// means and per-read status are already in HBM, LLRs and positions stay there for
// tb2_region_stats_add_batch_llr / tb2_batch_llr_download
// ---------------------------------------------------------------------------
namespace {
// exclusive scan of the per-read site counts (one block; n is a batch, <= ~1e6)
__global__ void __launch_bounds__(1024) k_scan_sites(const int *cnt, long long *site_off, int n)
{
    __shared__ long long warp_tot[32];
    __shared__ long long base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { base_s = 0; site_off[0] = 0; }
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int i = c0 + tid;
        long long v = i < n ? cnt[i] : 0, inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += o;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        long long before = 0, total = 0;
        for (int q = 0; q < 32; ++q) { if (q < warp) before += warp_tot[q]; total += warp_tot[q]; }
        if (i < n) site_off[i + 1] = base_s + before + inc;
        __syncthreads();
        if (tid == 0) base_s += total;
        __syncthreads();
    }
}
enum { L_RSTART = 63, L_RCNT, L_RSITEOFF, L_RLLR, L_RPOS, L_RTOTAL };
}  // namespace

extern "C" int tb2_batch_alt_llr(tb2_ctx *ctx, const int64_t *read_start, int alt_base_code,
                                 int use_standard_llhr, double scale_factor, double height_factor,
                                 double height_power, int64_t *n_sites_total)
{
    int rc = tb2_use(ctx);
    if (rc) return rc;
    if (!read_start || alt_base_code < 0 || alt_base_code > 3) return TB2_ERR_INVALID_ARG;
    if (ctx->kmer_width <= 0 || ctx->alt_kmer_width != ctx->kmer_width) {
        ctx->err = "standard and alternative models must be set with the same k-mer width";
        return TB2_ERR_INVALID_ARG;
    }
    BatchResultView v;
    if ((rc = tb2_batch_result_view(ctx, &v))) return rc;
    const int n = v.n_reads;
    auto &P = ctx->pool;
    cudaStream_t s = ctx->stream;
    TB2_CUDA_TRY(ctx, P[L_RSTART].reserve((size_t)n * 8));
    TB2_CUDA_TRY(ctx, P[L_RCNT].reserve((size_t)n * 4));
    TB2_CUDA_TRY(ctx, P[L_RSITEOFF].reserve((size_t)(n + 1) * 8));
    // every site is a base: the batch's base count bounds the site count (no size round trip)
    TB2_CUDA_TRY(ctx, P[L_RLLR].reserve((size_t)v.total_bases * 8 + 8));
    TB2_CUDA_TRY(ctx, P[L_RPOS].reserve((size_t)v.total_bases * 8 + 8));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_RSTART].p, read_start, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    LlrArgs a;
    a.n_reads = n; a.K = ctx->kmer_width; a.cpos = ctx->central_pos; a.alt_code = alt_base_code;
    a.use_std = use_standard_llhr ? 1 : 0;
    a.sf = scale_factor; a.hf = height_factor; a.hp = height_power;
    a.norm_mean = v.norm_mean; a.mean_off = v.base_off; a.seq_off = v.seq_off;
    a.read_start = P[L_RSTART].as<long long>();
    a.seq = v.seq;
    a.kmeans = ctx->model_means.as<double>(); a.ksds = ctx->model_sds.as<double>();
    a.alt = ctx->alt_means.as<double>();
    a.status = v.status; a.status_stride = v.stride;
    k_llr<false><<<n, 256, 0, s>>>(a, P[L_RCNT].as<int>(), nullptr, nullptr, nullptr);
    TB2_CHECK_LAUNCH(ctx);
    k_scan_sites<<<1, 1024, 0, s>>>(P[L_RCNT].as<int>(), P[L_RSITEOFF].as<long long>(), n);
    TB2_CHECK_LAUNCH(ctx);
    k_llr<true><<<n, 256, 0, s>>>(a, nullptr, P[L_RSITEOFF].as<long long>(), P[L_RLLR].as<double>(),
                                  P[L_RPOS].as<long long>());
    TB2_CHECK_LAUNCH(ctx);
    long long total = 0;
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(&total, P[L_RSITEOFF].as<long long>() + n, 8, cudaMemcpyDeviceToHost, s));
    TB2_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    ctx->resident_llr_sites = total;
    ctx->resident_llr_reads = n;
    if (n_sites_total) *n_sites_total = total;
    return TB2_OK;
}

extern "C" int tb2_batch_llr_download(tb2_ctx *ctx, double *llr_out, int64_t *pos_out, int64_t *site_off)
{
    int rc = tb2_use(ctx);
    if (rc) return rc;
    if (ctx->resident_llr_reads <= 0 || !site_off) return TB2_ERR_INVALID_ARG;
    auto &P = ctx->pool;
    cudaStream_t s = ctx->stream;
    const size_t t = (size_t)ctx->resident_llr_sites;
    if (t && (!llr_out || !pos_out)) return TB2_ERR_INVALID_ARG;
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(site_off, P[L_RSITEOFF].p, (size_t)(ctx->resident_llr_reads + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (t) {
        TB2_CUDA_TRY(ctx, cudaMemcpyAsync(llr_out, P[L_RLLR].p, t * 8, cudaMemcpyDeviceToHost, s));
        TB2_CUDA_TRY(ctx, cudaMemcpyAsync(pos_out, P[L_RPOS].p, t * 8, cudaMemcpyDeviceToHost, s));
    }
    TB2_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return TB2_OK;
}

// the resident LLRs feed the region counters without leaving the device
extern "C" int tb2_region_stats_add_batch_llr(tb2_ctx *ctx, double single_read_thresh,
                                              double lower_thresh, int stat_type)
{
    int rc = tb2_use(ctx);
    if (rc) return rc;
    if (ctx->resident_llr_reads <= 0 || isnan(single_read_thresh)) return TB2_ERR_INVALID_ARG;
    return tb2_region_accumulate_dev(ctx, ctx->resident_llr_sites, ctx->pool[L_RLLR].as<double>(),
                                     ctx->pool[L_RPOS].as<long long>(), single_read_thresh,
                                     lower_thresh, stat_type);
}

extern "C" int tb2_new_mean_stds(tb2_ctx *ctx, const double *sig, int64_t n_sig,
                                 const int64_t *segs, int64_t n_segs, double *means_out,
                                 double *sds_out)
{
    int rc = tb2_use(ctx);
    if (rc) return rc;
    if (!sig || !segs || !means_out || !sds_out || n_sig < 1 || n_segs < 1) return TB2_ERR_INVALID_ARG;
    for (int64_t i = 0; i <= n_segs; ++i)
        if (segs[i] < 0 || segs[i] > n_sig) return TB2_ERR_INVALID_ARG;
    auto &P = ctx->pool;
    cudaStream_t s = ctx->stream;
    TB2_CUDA_TRY(ctx, P[L_A].reserve((size_t)n_sig * 8));
    TB2_CUDA_TRY(ctx, P[L_B].reserve((size_t)(n_segs + 1) * 8));
    TB2_CUDA_TRY(ctx, P[L_C].reserve((size_t)n_segs * 8));
    TB2_CUDA_TRY(ctx, P[L_D].reserve((size_t)n_segs * 8));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_A].p, sig, (size_t)n_sig * 8, cudaMemcpyHostToDevice, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_B].p, segs, (size_t)(n_segs + 1) * 8, cudaMemcpyHostToDevice, s));
    k_mean_stds<<<(unsigned)((n_segs + 127) / 128), 128, 0, s>>>(
        P[L_A].as<double>(), P[L_B].as<long long>(), n_segs, P[L_C].as<double>(), P[L_D].as<double>());
    TB2_CHECK_LAUNCH(ctx);
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(means_out, P[L_C].p, (size_t)n_segs * 8, cudaMemcpyDeviceToHost, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(sds_out, P[L_D].p, (size_t)n_segs * 8, cudaMemcpyDeviceToHost, s));
    TB2_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return TB2_OK;
}

// ---------------------------------------------------------------------------
// batched mirrors of the three Cython scorers over explicit windows:
// mode 0 c_calc_scaled_llh_ratio_const_var (_c_helper.pyx:313-358)
// mode 1 c_calc_llh_ratio_const_var (:298-311), mode 2 c_calc_llh_ratio (:277-296)
// means / ref_means / alt_means: n x K row-major; var_a: const_var[n] (modes 0,1)
// or ref_vars[n x K] (mode 2); var_b: alt_vars[n x K] (mode 2)
// ---------------------------------------------------------------------------
namespace {
__global__ void k_llh_windows(int mode, long long n, int K, const double *m, const double *rm,
                              const double *am, const double *va, const double *vb, double sf,
                              double hf, double hp, double *out)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double *pm = m + s * K, *pr = rm + s * K, *pa = am + s * K;
    double acc = 0.0;
    if (mode == 2) {
        double rz = 0, rl = 0, az = 0, al = 0;
        for (int i = 0; i < K; ++i) {
            const double rd = pm[i] - pr[i];
            rz += (rd * rd) / va[s * K + i];
            rl += log(va[s * K + i]);
            const double ad = pm[i] - pa[i];
            az += (ad * ad) / vb[s * K + i];
            al += log(vb[s * K + i]);
        }
        out[s] = az + al - rz - rl;
        return;
    }
    const double cv = va[s];
    for (int i = 0; i < K; ++i) {
        const double obs = pm[i], ref_mean = pr[i], alt_mean = pa[i];
        if (mode == 1) {
            const double rd = obs - ref_mean, ad = obs - alt_mean;
            acc += ((ad * ad) - (rd * rd)) / cv;
        } else {
            if (ref_mean == alt_mean) continue;
            const double scale_mean = (alt_mean + ref_mean) / 2;
            const double ref_diff = obs - ref_mean, alt_diff = obs - alt_mean;
            const double scale_diff = obs - scale_mean;
            double means_diff = alt_mean - ref_mean;
            if (means_diff < 0) means_diff = means_diff * -1;
            acc += exp(-(scale_diff * scale_diff) / (sf * cv)) *
                   ((alt_diff * alt_diff) - (ref_diff * ref_diff)) / (cv * pow(means_diff, hp) * hf);
        }
    }
    out[s] = acc;
}
}  // namespace

extern "C" int tb2_calc_llh_ratio_windows(tb2_ctx *ctx, int mode, int64_t n_sites, int kmer_width,
                                          const double *means, const double *ref_means,
                                          const double *alt_means, const double *var_a,
                                          const double *var_b, double scale_factor,
                                          double height_factor, double height_power,
                                          double *llr_out)
{
    int rc = tb2_use(ctx);
    if (rc) return rc;
    if (mode < 0 || mode > 2 || n_sites < 0 || kmer_width < 1 || !var_a || (mode == 2 && !var_b))
        return TB2_ERR_INVALID_ARG;
    if (n_sites == 0) return TB2_OK;
    if (!means || !ref_means || !alt_means || !llr_out) return TB2_ERR_INVALID_ARG;
    auto &P = ctx->pool;
    cudaStream_t s = ctx->stream;
    const size_t nk = (size_t)n_sites * kmer_width * 8, nv = (size_t)n_sites * 8;
    TB2_CUDA_TRY(ctx, P[L_MEAN].reserve(nk));
    TB2_CUDA_TRY(ctx, P[L_A].reserve(nk));
    TB2_CUDA_TRY(ctx, P[L_B].reserve(nk));
    TB2_CUDA_TRY(ctx, P[L_C].reserve(mode == 2 ? nk : nv));
    TB2_CUDA_TRY(ctx, P[L_D].reserve(mode == 2 ? nk : 8));
    TB2_CUDA_TRY(ctx, P[L_LLR].reserve(nv));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_MEAN].p, means, nk, cudaMemcpyHostToDevice, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_A].p, ref_means, nk, cudaMemcpyHostToDevice, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_B].p, alt_means, nk, cudaMemcpyHostToDevice, s));
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_C].p, var_a, mode == 2 ? nk : nv, cudaMemcpyHostToDevice, s));
    if (mode == 2) TB2_CUDA_TRY(ctx, cudaMemcpyAsync(P[L_D].p, var_b, nk, cudaMemcpyHostToDevice, s));
    k_llh_windows<<<(unsigned)((n_sites + 127) / 128), 128, 0, s>>>(
        mode, n_sites, kmer_width, P[L_MEAN].as<double>(), P[L_A].as<double>(), P[L_B].as<double>(),
        P[L_C].as<double>(), P[L_D].as<double>(), scale_factor, height_factor, height_power,
        P[L_LLR].as<double>());
    TB2_CHECK_LAUNCH(ctx);
    TB2_CUDA_TRY(ctx, cudaMemcpyAsync(llr_out, P[L_LLR].p, nv, cudaMemcpyDeviceToHost, s));
    TB2_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return TB2_OK;
}
