// Launcher of fused_tc_logl_kernel (tcgen05 throughput path, tc_kernel.cuh).
#include <algorithm>

#include "handle.h"
#include "tc_kernel.cuh"

namespace nmma {

namespace {
// Adds the per-part sums of a filter-split launch in part order (deterministic); any failed part -> sentinel.
__global__ void combine_parts_kernel(const double* __restrict__ parts, int fsplit, long long N, double* __restrict__ out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double s = 0.0;
    for (int p = 0; p < fsplit; ++p) s += parts[n * fsplit + p];
    out[n] = isfinite(s) ? s : NMMA_SENTINEL;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel function, not to the handle or the launcher: one
// high-water mark per instantiation (two launchers with their own marks lowered each other's setting), and one driver
// call per configuration instead of one per launch (one-point latency).
template <int K, bool FAST, bool SPLIT, bool COEFF>
int ensure_tc_smem(nmma_b200_t* h, size_t smem) {
    static size_t mark = 0;
    if (mark < smem) {
        CU(cudaFuncSetAttribute(fused_tc_logl_kernel<K, FAST, SPLIT, COEFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mark = smem;
    }
    return NMMA_B200_OK;
}

template <bool FAST>
int launch_tc_f(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st, bool per_filter = false) {
    constexpr int K = 10;
    const size_t smem = tc_smem_bytes(K, h->T, h->cfg.S, h->cfg.nobs);
    if (int rc = ensure_tc_smem<K, FAST, true, false>(h, smem)) return rc;
    if (int rc = ensure_tc_smem<K, FAST, false, false>(h, smem)) return rc;
    const long long super = (long long)kTcTile * kTcTiles;
    const long long nsuper = (N + super - 1) / super;
    long long grid = h->sm_count;  // one CTA per SM: each CTA owns all 512 TMEM columns of its SM
    if (h->opt_max_ctas > 0) grid = std::min<long long>(grid, h->opt_max_ctas);
    // small batches: a 256-point super-tile takes ~205 us on one SM whatever N is, so when the super-tiles leave SMs idle
    // the filters of each are spread over several CTAs (per-part sums, combined in a fixed order)
    // large batches whose last wave would leave more than half of the SMs idle: whole waves with the un-split kernel, the
    // rest with ONE FILTER PER WORK ITEM spread over all SMs (e.g. 10^6 points = 26 waves + 59 super-tiles -> 26.5 instead of
    // 27 rounds).  One part per filter, added in filter order, is the association of the un-split kernel's running sum:
    // a row's log-likelihood does not depend on which side of the cut it is (test_bu2019lm_device_tensor_and_large_batch).
    if (!h->opt_no_fsplit && !per_filter && nsuper > grid && (nsuper % grid) > 0 && (nsuper % grid) * 2 <= grid) {
        const long long n_main = (nsuper / grid) * grid * super;
        if (int rc = launch_tc_f<FAST>(h, pts, n_main, out, st)) return rc;
        return launch_tc_f<FAST>(h, pts + n_main * h->cfg.P, N - n_main, out + n_main, st, true);
    }
    int fsplit = 1;
    if (per_filter) fsplit = h->F;
    else if (!h->opt_no_fsplit && nsuper * 2 <= grid) fsplit = (int)std::min<long long>(h->F, grid / nsuper);
    double* dst = out;
    if (fsplit > 1) {
        const size_t need = (size_t)N * fsplit;
        if (need > h->tc_parts_cap) {
            if (h->tc_parts) cudaFree(h->tc_parts);
            h->tc_parts = nullptr; h->tc_parts_cap = 0;
            CU(cudaMalloc((void**)&h->tc_parts, need * sizeof(double)));
            h->tc_parts_cap = need;
        }
        dst = h->tc_parts;
    }
    grid = std::max<long long>(1, std::min(grid, nsuper * fsplit));
    if (fsplit > 1) fused_tc_logl_kernel<K, FAST, true><<<(unsigned)grid, kTcThreads, smem, st>>>(h->cfg, pts, N, dst, fsplit, 1);
    else fused_tc_logl_kernel<K, FAST, false><<<(unsigned)grid, kTcThreads, smem, st>>>(h->cfg, pts, N, dst, 1, 1);
    CU(cudaGetLastError());
    if (fsplit > 1) {
        combine_parts_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(h->tc_parts, fsplit, N, out);
        CU(cudaGetLastError());
        h->launches += 1;
    }
    h->launches += 1;
    h->last_ctas_per_sm = 1;
    return NMMA_B200_OK;
}
}  // namespace

// Front end only: coefficients [N][F][K] (fp64) for the generic back end -- any n_coeff <= 16, any filter mapping.
int launch_tc_coeff(nmma_b200_t* h, const double* pts, long long N, double* coeff, cudaStream_t st) {
    constexpr int K = kTcN2;
    const size_t smem = tc_smem_bytes(K, h->T, h->cfg.S, h->cfg.nobs);
    if (int rc = ensure_tc_smem<K, false, true, true>(h, smem)) return rc;
    if (int rc = ensure_tc_smem<K, false, false, true>(h, smem)) return rc;
    const long long super = (long long)kTcTile * kTcTiles;
    const long long nsuper = (N + super - 1) / super;
    long long grid = h->sm_count;
    if (h->opt_max_ctas > 0) grid = std::min<long long>(grid, h->opt_max_ctas);
    int fsplit = 1;   // small batches: the filters of a super-tile over several CTAs (each part writes its own filters)
    if (!h->opt_no_fsplit && nsuper * 2 <= grid) fsplit = (int)std::min<long long>(h->F, grid / nsuper);
    grid = std::max<long long>(1, std::min(grid, nsuper * fsplit));
    if (fsplit > 1) fused_tc_logl_kernel<K, false, true, true><<<(unsigned)grid, kTcThreads, smem, st>>>(h->cfg, pts, N, coeff, fsplit, 0);
    else fused_tc_logl_kernel<K, false, false, true><<<(unsigned)grid, kTcThreads, smem, st>>>(h->cfg, pts, N, coeff, 1, 0);
    CU(cudaGetLastError());
    h->launches += 1;
    h->last_ctas_per_sm = 1;
    return NMMA_B200_OK;
}

// Latency path (one point per call from bilby / pymultinest, small live-point batches): the filters AND the hidden layer of
// each 256-point super-tile are spread over the SMs; `parts` receives fp32 partial coefficient sums
// [N][F][hsplit][K] that backend_logl_parts_kernel (api.cu) adds in range order.  Returns hsplit through *hsplit_out.
int launch_tc_coeff_parts(nmma_b200_t* h, const double* pts, long long N, float* parts, int* hsplit_out, cudaStream_t st) {
    constexpr int K = kTcN2;
    const size_t smem = tc_smem_bytes(K, h->T, h->cfg.S, h->cfg.nobs);
    if (int rc = ensure_tc_smem<K, false, true, true>(h, smem)) return rc;
    const long long super = (long long)kTcTile * kTcTiles;
    const long long nsuper = (N + super - 1) / super;
    const long long grid_max = h->opt_max_ctas > 0 ? std::min<long long>(h->sm_count, h->opt_max_ctas) : h->sm_count;
    const int fsplit = (int)std::max<long long>(1, std::min<long long>(h->F, grid_max / nsuper));
    // hidden ranges: a power of two that leaves every range whole accumulation groups (in parts mode: kTcBufs chunks)
    int hsplit = 1;
    while (nsuper * fsplit * (hsplit * 2) <= grid_max && h->cfg.tc_nch % (hsplit * 2 * kTcBufs) == 0) hsplit *= 2;
    const long long grid = std::max<long long>(1, std::min(grid_max, nsuper * fsplit * hsplit));
    fused_tc_logl_kernel<K, false, true, true><<<(unsigned)grid, kTcThreads, smem, st>>>(
        h->cfg, pts, N, reinterpret_cast<double*>(parts), fsplit, hsplit);
    CU(cudaGetLastError());
    h->launches += 1;
    h->last_ctas_per_sm = 1;
    *hsplit_out = hsplit;
    return NMMA_B200_OK;
}

int launch_tc(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    const bool fast = h->cfg.single_stage && h->cfg.uniform && !h->opt_no_fast;
    return fast ? launch_tc_f<true>(h, pts, N, out, st) : launch_tc_f<false>(h, pts, N, out, st);
}

}  // namespace nmma
