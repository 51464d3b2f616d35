// Launcher of fused_hy_logl_kernel (FFMA layer 1 + tcgen05 layer 2, hy_kernel.cuh).
#include <algorithm>

#include "handle.h"
#include "hy_kernel.cuh"

namespace nmma {

namespace {
template <int D, bool FAST>
int launch_hy_df(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    constexpr int K = 10;
    auto kern = fused_hy_logl_kernel<D, K, FAST>;
    const size_t smem = hy_smem_bytes(D, K, h->T, h->cfg.S, h->cfg.nobs);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long nsuper = (N + kHySuper - 1) / kHySuper;
    long long grid = h->sm_count;  // one CTA per SM: each CTA owns all 512 TMEM columns of its SM
    if (h->opt_max_ctas > 0) grid = std::min<long long>(grid, h->opt_max_ctas);
    grid = std::max<long long>(1, std::min(grid, nsuper));
    kern<<<(unsigned)grid, kHyThreads, smem, st>>>(h->cfg, pts, N, out);
    CU(cudaGetLastError());
    h->launches += 1;
    h->last_ctas_per_sm = 1;
    return NMMA_B200_OK;
}
}  // namespace

bool hy_has(int d, int K) { return (d == 3 || d == 4 || d == 7) && K == 10; }

int launch_hy(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    const bool fast = h->cfg.single_stage && h->cfg.uniform && !h->opt_no_fast;
    switch (h->d) {
#ifndef NMMA_DEV_BUILD
        case 3: return fast ? launch_hy_df<3, true>(h, pts, N, out, st) : launch_hy_df<3, false>(h, pts, N, out, st);
        case 7: return fast ? launch_hy_df<7, true>(h, pts, N, out, st) : launch_hy_df<7, false>(h, pts, N, out, st);
#endif
        case 4: return fast ? launch_hy_df<4, true>(h, pts, N, out, st) : launch_hy_df<4, false>(h, pts, N, out, st);
        default: return fail(h, NMMA_B200_ERR_UNSUPPORTED, "hybrid kernel not instantiated for d=%d", h->d);
    }
}

}  // namespace nmma
