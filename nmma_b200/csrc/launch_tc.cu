// Launcher of fused_tc_logl_kernel (tcgen05 throughput path, tc_kernel.cuh).
#include <algorithm>

#include "handle.h"
#include "tc_kernel.cuh"

namespace nmma {

namespace {
template <bool FAST>
int launch_tc_f(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    constexpr int K = 10;
    auto kern = fused_tc_logl_kernel<K, FAST>;
    const size_t smem = tc_smem_bytes(K, h->T, h->cfg.S, h->cfg.nobs);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long super = (long long)kTcTile * kTcTiles;
    const long long nsuper = (N + super - 1) / super;
    long long grid = h->sm_count;  // one CTA per SM: each CTA owns all 512 TMEM columns of its SM
    if (h->opt_max_ctas > 0) grid = std::min<long long>(grid, h->opt_max_ctas);
    grid = std::max<long long>(1, std::min(grid, nsuper));
    kern<<<(unsigned)grid, kTcThreads, smem, st>>>(h->cfg, pts, N, out);
    CU(cudaGetLastError());
    h->launches += 1;
    h->last_ctas_per_sm = 1;
    return NMMA_B200_OK;
}
}  // namespace

int launch_tc(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    const bool fast = h->cfg.single_stage && h->cfg.uniform && !h->opt_no_fast;
    return fast ? launch_tc_f<true>(h, pts, N, out, st) : launch_tc_f<false>(h, pts, N, out, st);
}

}  // namespace nmma
