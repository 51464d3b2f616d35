// Launcher of fused_mlp_logl_kernel (FFMA throughput path, kernels.cuh).  The file is compiled once per input
// dimension with -DNMMA_FUSED_D=<d> (the instantiations of that d: PT = 1/2/4 x FAST) and once without (the dispatcher),
// so that the 36 kernel variants build in parallel.
#include <algorithm>

#include "handle.h"
#include "kernels.cuh"

namespace nmma {

template <int D>
int launch_fused_d(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st);

#ifdef NMMA_FUSED_D
namespace {
template <int D, int PT, bool FAST>
int launch_fused_dk(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    constexpr int K = 10;
    auto kern = fused_mlp_logl_kernel<D, K, PT, FAST>;
    const size_t smem = fused_smem_bytes(D, K, h->T, h->cfg.S, h->cfg.nobs);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFusedThreads, smem));
    if (per_sm < 1) return fail(h, NMMA_B200_ERR_CUDA, "fused kernel does not fit on an SM (smem %zu B)", smem);
    const long long tile = (long long)kFusedThreads * PT;
    const long long ntiles = (N + tile - 1) / tile;
    long long grid = (long long)h->sm_count * per_sm;
    if (h->opt_max_ctas > 0) grid = std::min<long long>(grid, h->opt_max_ctas);
    grid = std::max<long long>(1, std::min(grid, ntiles));
    kern<<<(unsigned)grid, kFusedThreads, smem, st>>>(h->cfg, pts, N, out);
    CU(cudaGetLastError());
    h->launches += 1;
    h->last_ctas_per_sm = per_sm;
    return NMMA_B200_OK;
}

template <int D, bool FAST>
int launch_fused_df(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    // points per thread: as many as keep every SM busy (a weight fetched from shared memory is
    // reused PT times; PT = 4 makes the kernel FMA-bound instead of LDS-bound)
    int pt = h->opt_pt;
    const long long per_wave = (long long)h->sm_count * kFusedThreads;
    // measured on B200 (profiles/): PT = 2 at two CTAs per SM beats PT = 4 at one CTA per SM
    if (pt == 0) pt = (N >= 2 * per_wave) ? 2 : 1;
    if (pt == 1) return launch_fused_dk<D, 1, FAST>(h, pts, N, out, st);
    if (pt == 2) return launch_fused_dk<D, 2, FAST>(h, pts, N, out, st);
    return launch_fused_dk<D, 4, FAST>(h, pts, N, out, st);
}

}  // namespace

template <int D>
int launch_fused_d(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    const bool fast = h->cfg.single_stage && h->cfg.uniform && !h->opt_no_fast;
    return fast ? launch_fused_df<D, true>(h, pts, N, out, st) : launch_fused_df<D, false>(h, pts, N, out, st);
}

template int launch_fused_d<NMMA_FUSED_D>(nmma_b200_t*, const double*, long long, double*, cudaStream_t);

#else  // dispatcher

int launch_fused(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    switch (h->d) {
#ifndef NMMA_DEV_BUILD  // development builds instantiate d = 4 only (compile time)
        case 2: return launch_fused_d<2>(h, pts, N, out, st);
        case 3: return launch_fused_d<3>(h, pts, N, out, st);
        case 5: return launch_fused_d<5>(h, pts, N, out, st);
        case 6: return launch_fused_d<6>(h, pts, N, out, st);
        case 7: return launch_fused_d<7>(h, pts, N, out, st);
#endif
        case 4: return launch_fused_d<4>(h, pts, N, out, st);
        default: return fail(h, NMMA_B200_ERR_UNSUPPORTED, "fused kernel not instantiated for d=%d", h->d);
    }
}

bool fused_has(int d, int K) { return d >= 2 && d <= 7 && K == 10; }

#endif  // NMMA_FUSED_D

}  // namespace nmma
