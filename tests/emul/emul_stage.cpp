// emul_stage.cpp -- TEST INFRASTRUCTURE ONLY: the stage kernels' device source
// (tombo_b200/csrc/stage_kernels.cuh) run on the host through tests/emul/cuda_emul.h, one
// read at a time, for comparison with the oracle in the CPU test-suite.
#include "cuda_emul.h"
#include "../../tombo_b200/csrc/stage_kernels.cuh"
#include <vector>

extern "C" {

// k_resolve on one read (resolve_skipped_bases_with_raw resquiggle.py:402-540)
int emul_resolve(const int *segs_dp, int nb, const double *rm, const double *rs, const double *norm,
                 int n_norm, const tb2_params *p, long long max_raw_cpts, long long cap_doubles,
                 long long big_cap_doubles, int *segs_out, int *status_out)
{
    BatchView b;
    memset(&b, 0, sizeof(b));
    long long raw_off[2] = {0, n_norm}, base_off[2] = {0, nb};
    ReadState st;
    memset(&st, 0, sizeof(st));
    st.active = 1; st.status = TB2_OK; st.rsrtr = 0;
    std::vector<int> starts(nb + 8), read_tb(nb + 8), sd(segs_dp, segs_dp + nb + 1), so(nb + 1);
    std::vector<double> n2(norm, norm + n_norm), rmv(rm, rm + nb), rsv(rs, rs + nb);
    b.n_reads = 1; b.raw_off = raw_off; b.base_off = base_off;
    b.norm = n2.data(); b.rm = rmv.data(); b.rs = rsv.data();
    b.starts = starts.data(); b.read_tb = read_tb.data();
    b.segs_dp = sd.data(); b.segs = so.data(); b.st = &st;
    StagePolicy pol;
    memset(&pol, 0, sizeof(pol));
    pol.max_raw_cpts = max_raw_cpts;
    std::vector<double> pool((size_t)cap_doubles * 4 + 8), big((size_t)big_cap_doubles + 8);
    int counter[4] = {0, 0, 0, 0};
    emul::launch(emul::Idx3{1, 1, 1}, 128, 0, [&]() {
        k_resolve(b, *p, pol, pool.data(), (size_t)cap_doubles, counter, big.data(),
                  (unsigned long long)big_cap_doubles, (unsigned long long *)(counter + 2));
    });
    for (int i = 0; i <= nb; ++i) segs_out[i] = so[i];
    *status_out = st.status;
    return 0;
}

// k_theil_sen on one read (calc_kmer_fitted_shift_scale tombo_stats.py:370-450)
int emul_theil_sen(const double *bm, const double *rm, int nb, double prev_shift, double prev_scale,
                   unsigned int key, double *out4 /* shift, scale, shc, scc */, int *status_out)
{
    BatchView b;
    memset(&b, 0, sizeof(b));
    long long base_off[2] = {0, nb};
    ReadState st;
    memset(&st, 0, sizeof(st));
    st.active = 1; st.status = TB2_OK;
    st.sv.shift = prev_shift; st.sv.scale = prev_scale;
    std::vector<double> bmv(bm, bm + nb), rmv(rm, rm + nb);
    b.n_reads = 1; b.base_off = base_off; b.bm = bmv.data(); b.rm = rmv.data(); b.st = &st;
    StagePolicy pol;
    memset(&pol, 0, sizeof(pol));
    pol.outlier_thresh = 5.0; pol.subsample_seed = key; pol.literal_key = 1;
    emul::launch(emul::Idx3{1, 1, 1}, ST_THREADS, sizeof(TsSmem), [&]() { k_theil_sen(b, pol, 0); });
    out4[0] = st.sv.shift; out4[1] = st.sv.scale; out4[2] = st.shc; out4[3] = st.scc;
    *status_out = st.status;
    return 0;
}

// tb2_block_select2 (select.cuh) on one array: values of ascending rank k and k + 1
void emul_select2(const double *v, int n, int k, double *out2)
{
    static SelectSmem sm;
    emul::launch(emul::Idx3{1, 1, 1}, TB2_SEL_THREADS, 0, [&]() {
        double a = 0, b = 0;
        auto f = [&](int i) { return v[i]; };
        tb2_block_select2(f, PredAll(), n, k, true, &a, &b, sm);
        if (threadIdx.x == 0) { out2[0] = a; out2[1] = b; }
    });
}

void emul_ts_counters(unsigned long long *out8, int reset)
{
    for (int i = 0; i < 8; ++i) out8[i] = g_tb2_counters[i];
    if (reset) for (int i = 0; i < 8; ++i) g_tb2_counters[i] = 0;
}
}
