// Host-side state of one engine handle and the launcher entry points shared by the translation units of
// libnmma_b200.so (api.cu: C ABI, tables, two-stage kernels; launch_fused.cu / launch_tc.cu: one
// throughput kernel family each, so that they compile in parallel).
#pragma once
#include "../../include/nmma_b200.h"

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "backend.cuh"

namespace nmma {
struct PriorPlan;
}
struct nmma_b200_handle {
    int device = 0;
    int sm_count = 0;
    std::string err;
    // ---- host copies of the configuration ----
    int F = 0, d = 0, K = 0, T = 0;
    std::vector<double> tt, pmin, pmax, VA, mins, maxs;
    bool have_svd = false;
    int kind = -1;  // 0 mlp, 1 gp
    int H = 0, Kout = 0;
    std::vector<float> W1, b1, W2, b2;
    int Ntr = 0;
    std::vector<double> gpX, gpAlpha, gpC2, gpRa, gpRl, gpYm, gpYs;
    std::vector<double> samp;  // empty: default to tt[0]
    int P = 0;
    bool have_layout = false;
    std::vector<nmma::ParamSrc> xsrc;
    nmma::ParamSrc dl{-1, 0, 1e-5}, ts{-1, 0, 0.0}, zsrc{-1, 0, 0.0};
    int zmode = 0;
    std::vector<double> zd, zz;
    int G = 0;
    bool have_obs = false;
    std::vector<int> g_nh, g_h, g_off;
    std::vector<double> o_t, o_m, o_s, g_lim;
    bool have_sys = false;
    std::vector<int> sy_mode, sy_nn, sy_off;
    std::vector<double> sy_budget, sy_t;
    std::vector<nmma::ParamSrc> sy_src;
    // Constraint priors and extinction (api.cu: nmma_b200_set_constraints / nmma_b200_set_extinction)
    std::vector<nmma::ParamSrc> con_src;
    std::vector<double> con_lo, con_hi;
    int ext_law = 0;
    nmma::ParamSrc ebv{-1, 0, 0.0};
    std::vector<double> ext_nu, ext_coef;
    // ---- priors on the device (prior.cu) ----
    int prP = 0;
    int pr_tab_total = 0;                   // entries in each half (cdf | grid) of pr_tab_dev
    nmma::PriorPlan* pr_dev = nullptr;      // device copy of the per-column prior plan
    double* pr_tab_dev = nullptr;           // INTERPED tables: cdf values then grid values
    double* sweep_scratch = nullptr;        // L2-sized block of drawn points for nmma_b200_logl_sweep
    size_t sweep_cap = 0;
    // ---- device state ----
    bool dirty = true;
    std::vector<void*> dev_allocs;
    nmma::DevCfg cfg{};
    bool fused_supported = false;
    bool tc_supported = false;
    bool fast_backend_ok = false;      // direct filter maps on the uniform training grid: the fp32 per-observation term applies
    bool tc_front_supported = false;   // tensor-core front end in coefficient mode (any n_coeff <= 16, any filter mapping)
    long long opt_tc_front_min = 128;  // two-stage path: coefficients from the tensor-core kernel from this batch size
    double* coeff_scratch = nullptr;
    double* tc_parts = nullptr;       // per-part sums of a filter-split tensor-core launch (launch_tc.cu)
    size_t tc_parts_cap = 0;
    double* gp_parts = nullptr;       // fused GP kernel: per-(tile, filter) sums and per-tile tickets (launch_gp.cu)
    unsigned int* gp_tickets = nullptr;
    size_t gp_parts_cap = 0, gp_tickets_cap = 0;
    bool gp_fused_supported = false;
    bool gp_alpha_ok = false;         // every RationalQuadratic alpha in (0, 1e5]: gf_pow applies
    size_t coeff_cap = 0;
    double* stage_in_dev = nullptr;
    double* stage_out_dev = nullptr;
    double* stage_in_host = nullptr;
    double* stage_out_host = nullptr;
    size_t stage_cap_in = 0, stage_cap_out = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_in_stream = nullptr, copy_out_stream = nullptr;   // nmma_b200_logl_host copy/compute pipeline
    std::vector<cudaEvent_t> pipe_events;
    // CUDA graphs of the latency path (nmma_b200_logl_host, N <= 256): copy + two kernels replayed with one launch call
    struct LatGraph { int64_t N; unsigned long long epoch; cudaGraphExec_t exec; long long launches; };
    std::vector<LatGraph> lat_graphs;
    std::vector<int64_t> lat_warm;          // batch sizes that have run once un-captured at this epoch (allocations done)
    unsigned long long cfg_epoch = 0;       // bumped whenever finalize() rebuilds the device configuration
    int opt_graphs = 1;                     // set_option "cuda_graphs"
    int opt_pipeline = 6;             // row blocks of the host-buffer pipeline (set_option "pipeline_blocks"; 1 = serial)
    // ---- knobs / counters ----
    int opt_path = 0;
    long long opt_fused_min = 2048;
    long long opt_tc_min = 1;         // tensor-core path from the first point: with the filters of a super-tile split over CTAs
                                      // (launch_tc.cu) one call takes 39 us at N = 1 and 40-46 us up to 4096 points, the two-stage
                                      // kernels 47 us + 0.6 us per point (tools/latency_breakdown.py, tools/latency.py)
    long long opt_latency_max = 1024; // up to this batch size the hidden layer is split over CTAs too (path 5: launch_tc_coeff_parts +
                                      // backend_logl_parts_kernel); above, the filter-split fused kernel (path 3)
    long long opt_gp_min = 4096;      // fused GP kernel (thread = point) from this batch size; below, the two-stage kernels (lanes = training rows)
    int opt_max_ctas = 0;
    int opt_no_fast = 0;
    int opt_zero_copy = 1;            // set_option "zero_copy": small host batches are read / written in place (pinned memory)
    int opt_no_fsplit = 0;            // set_option "no_filter_split": keep one CTA per super-tile
    int last_ctas_per_sm = 0;
    int opt_pt = 0;
    long long launches = 0;
    int last_path = 0;
};

namespace nmma {
// records the message on the handle (or for a failed create when h == NULL) and returns `code`
int fail(nmma_b200_t* h, int code, const char* fmt, ...);

// throughput launchers; *_has: is the kernel instantiated for this (d, n_coeff)?
int launch_fused(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st);
bool fused_has(int d, int K);
int launch_tc(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st);
int launch_tc_coeff(nmma_b200_t* h, const double* pts, long long N, double* coeff, cudaStream_t st);
int launch_tc_coeff_parts(nmma_b200_t* h, const double* pts, long long N, float* parts, int* hsplit_out, cudaStream_t st);
int launch_gp(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st);
bool gp_fused_has(int d, int K);
}  // namespace nmma

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return nmma::fail(h, NMMA_B200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
