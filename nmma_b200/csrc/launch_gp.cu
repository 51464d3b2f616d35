// Launcher of fused_gp_logl_kernel (sklearn_gp throughput path, gp_kernel.cuh).  Like launch_fused.cu the file is compiled
// once per input dimension with -DNMMA_GP_D=<d> (FAST and generic back end) and once without (the dispatcher).
#include <algorithm>

#include "gp_kernel.cuh"
#include "handle.h"

namespace nmma {

template <int D>
int launch_gp_d(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st);

#ifdef NMMA_GP_D
namespace {
template <int D, bool FAST>
int launch_gp_df(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    constexpr int K = 10;
    auto kern = fused_gp_logl_kernel<D, K, FAST>;
    const size_t smem = gf_smem_bytes(h->Ntr, D);
    static size_t attr_smem = 0;   // per instantiation: the attribute belongs to the function, not to the handle
    if (attr_smem < smem) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    const long long ntiles = (N + 31) / 32;
    // per-tile scratch: F x 32 per-filter sums + one ticket (zeroed once, the kernel resets what it used)
    const size_t need = (size_t)ntiles * h->F * 32;
    if (need > h->gp_parts_cap || (size_t)ntiles > h->gp_tickets_cap) {
        if (h->gp_parts) cudaFree(h->gp_parts);
        if (h->gp_tickets) cudaFree(h->gp_tickets);
        h->gp_parts = nullptr; h->gp_tickets = nullptr; h->gp_parts_cap = 0; h->gp_tickets_cap = 0;
        CU(cudaMalloc((void**)&h->gp_parts, need * sizeof(double)));
        CU(cudaMalloc((void**)&h->gp_tickets, (size_t)ntiles * sizeof(unsigned int)));
        CU(cudaMemsetAsync(h->gp_tickets, 0, (size_t)ntiles * sizeof(unsigned int), st));
        h->gp_parts_cap = need; h->gp_tickets_cap = (size_t)ntiles;
    }
    const long long nitems = ntiles * h->F;
    long long grid = h->sm_count;   // one 16-warp CTA per SM (the replicated tables take 64 KB of its shared memory)
    if (h->opt_max_ctas > 0) grid = std::min<long long>(grid, h->opt_max_ctas);
    grid = std::max<long long>(1, std::min(grid, (nitems + kGfWarps - 1) / kGfWarps));
    kern<<<(unsigned)grid, kGfWarps * 32, smem, st>>>(h->cfg, pts, N, out, h->gp_parts, h->gp_tickets);
    CU(cudaGetLastError());
    h->launches += 1;
    h->last_ctas_per_sm = 1;
    return NMMA_B200_OK;
}
}  // namespace

template <int D>
int launch_gp_d(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    const bool fast = h->cfg.single_stage && h->cfg.uniform && !h->opt_no_fast;
    return fast ? launch_gp_df<D, true>(h, pts, N, out, st) : launch_gp_df<D, false>(h, pts, N, out, st);
}

template int launch_gp_d<NMMA_GP_D>(nmma_b200_t*, const double*, long long, double*, cudaStream_t);

#else  // dispatcher

int launch_gp(nmma_b200_t* h, const double* pts, long long N, double* out, cudaStream_t st) {
    switch (h->d) {
#ifndef NMMA_DEV_BUILD  // development builds instantiate d = 3 only (compile time)
        case 2: return launch_gp_d<2>(h, pts, N, out, st);
        case 4: return launch_gp_d<4>(h, pts, N, out, st);
        case 5: return launch_gp_d<5>(h, pts, N, out, st);
        case 6: return launch_gp_d<6>(h, pts, N, out, st);
        case 7: return launch_gp_d<7>(h, pts, N, out, st);
#endif
        case 3: return launch_gp_d<3>(h, pts, N, out, st);
        default: return fail(h, NMMA_B200_ERR_UNSUPPORTED, "fused GP kernel not instantiated for d=%d", h->d);
    }
}

bool gp_fused_has(int d, int K) {
#ifdef NMMA_DEV_BUILD
    return d == 3 && K == 10;
#else
    return d >= 2 && d <= 7 && K == 10;
#endif
}

#endif  // NMMA_GP_D

}  // namespace nmma
