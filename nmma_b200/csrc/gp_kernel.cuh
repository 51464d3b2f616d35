// Fused sklearn_gp path: GP mean -> SVD coefficients -> light curve -> log-likelihood in ONE launch, no coefficient scratch.
//
//   c_{f,k}(x) = ystd * sum_t C^2 alpha_t (1 + |x - X_t|^2 q)^(-a) + ymean ,   q = 1 / (2 a l^2)
//   (sklearn RationalQuadratic.__call__ + GaussianProcessRegressor.predict, nmma/em/lightcurve_generation.py:200-211)
//
// F K Ntr = 29 610 kernel values per evaluation for Ka2017 (9 filters x 10 coefficients x 329 training rows), each an
// fp64 pow (the alpha-weighted sum cancels up to 3e5 : 1, so the values need ~1e-13 relative accuracy): the path is bound
// by the fp64 pipe (DFMA issues every 2 cycles per SM sub-partition, profiles/r01_fp64_rate.txt).  The two-stage
// coeff_gp_kernel reached 47 % of that pipe (profiles/r02_gp_backend_before_summary.json): 24 fp64 + 49 other
// instructions per value, three int<->fp64 conversions on the XU pipe, three random 8-byte shared-memory gathers with
// ~5-way bank conflicts (70 % of the shared-memory wavefront peak).  This kernel:
//   * thread = point, warp = work item (32-point tile, model filter), one persistent 16-warp CTA per SM; the K coefficient
//     sums of the filter are K independent dependency chains per thread (ILP instead of occupancy); all per-pair
//     constants and the alpha values are warp-uniform; r^2 is computed once per training row and reused by the K pairs;
//   * gf_pow: 17 fp64 instructions per value.  log2: base = 2^e m, u = m r_i - 1 with r_i an fp32 reciprocal of the
//     centre of the i-th of 256 mantissa cells (|u| <= 2^-9; the fma is exact), log2 m = -log2 r_i + u P3(u) with an
//     interpolated (near-minimax) cubic, error 1.1e-15 absolute.  exp2: t = -256 a log2(base); k = rint(t) and the
//     remainder by the 1.5 * 2^52 trick (no F2I / I2F), 2^(k/256) = 2^(k >> 8) * table[k & 255] * (1 + xr E3(xr)), error
//     5e-18.  One conversion (the exponent e) is left on the XU pipe.
//   * the two tables are replicated per bank group (8 x 16-byte log entries per quarter warp, 16 x 8-byte exp entries per
//     half warp): every gather is conflict-free, 4 + 2 wavefronts per warp-level value instead of ~15;
//   * the coefficients never leave registers: each warp scores its filter's observations right away through the same
//     fused_filter_logl back end as the MLP kernels (basis rows read through L1); the per-filter sums of a tile (8 B per
//     point and filter) are added in filter order by the warp that finishes the tile's last filter.
// Measured (profiles/r02_gp_fused_*): 10.5 -> 22 M evals/s on config 4; fp64 pipe 59 %.  What bounds it now is not the pipe
// but operand delivery and issue slots: a DFMA with three distinct register operands issues every 3 cycles, not 2
// (tools/fp64_rate.cu: "DFMA3r"), and the ~20 integer / load instructions per value cost almost a full slot each next to
// the fp64 stream; 8, 12, 16 or 20 warps per SM and 2- or 10-way interleaving all land within 5 % of each other.  The
// per-pair constants cannot be moved off the register file: ptxas loads constant-bank / kernel-parameter operands into
// vector registers inside the loop (tried: parameter struct indexed by a uniform filter index, 16 loop instances with
// compile-time offsets, and the polynomial coefficients in __constant__ memory: 20.9, 19.3 and 20.0 M evals/s -- every
// constant that lands in a vector register turns a 2-operand DFMA into a 3-operand one; 128-bit loads of the alphas:
// 20.8 M, the kernel sits at the 128-register limit and ten more live registers spill).
// Also tried: a warp-uniform fast path without the exponent clamp (host-computed r^2 bound per filter, __all_sync per training
// row, two instances of the row body): 9.05 -> 10.46 ms per 2e5 evaluations -- the vote, the branch and the doubled loop body
// cost more than the one integer min per value they save.
// gf_pow is restated operation by operation in tests/test_rq_pow.py (accuracy vs a 40-digit reference).
#pragma once
#include "kernels.cuh"

namespace nmma {

constexpr int kGfWarps = 16;                   // one CTA per SM: 4 warps per sub-partition at <= 128 registers

inline size_t gf_smem_bytes(int Ntr, int d) {
    return (size_t)kGfLogBytes + kGfExpBytes + ((size_t)Ntr * d * sizeof(double) + 127) / 128 * 128;
}

// The same arithmetic for the K pairs of one training row, written stage by stage so that the K dependency chains are in
// flight together (a dependent DFMA waits ~8 cycles, tools/fp64_rate.cu "DFMA chain1").
template <int K>
__device__ __forceinline__ void gf_pow_row(double r2, const double (&qk)[K], const double (&nak)[K],
                                           const double* __restrict__ a, double (&acc)[K],
                                           const unsigned char* __restrict__ ltab, const unsigned char* __restrict__ etab) {
    double base[K], u[K], l2[K], ed[K], p[K];
#pragma unroll
    for (int k = 0; k < K; ++k) base[k] = fma(r2, qk[k], 1.0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int hi = __double2hiint(base[k]);
        const double2 ent = *reinterpret_cast<const double2*>(ltab + ((hi >> 5) & 0x7f80));
        const double rs = __hiloint2double(__double2hiint(ent.x) + 0x3ff00000 - (hi & 0x7ff00000), __double2loint(ent.x));
        ed[k] = (double)((hi >> 20) - 1023);
        l2[k] = ent.y;
        u[k] = fma(base[k], rs, -1.0);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = fma(-0.36067471452205946, u[k], 0.4808994921226281);
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = fma(p[k], u[k], -0.7213475204440083);
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = fma(p[k], u[k], 1.4426950408883954);
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = fma(p[k], u[k], l2[k]) + ed[k];     // log2(base)
    double s[K], xr[K], g[K];
    int kk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = fma(nak[k], p[k], 6755399441055744.0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        xr[k] = fma(nak[k], p[k], -(s[k] - 6755399441055744.0));
        kk[k] = max(__double2loint(s[k]), -1020 * 256);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = fma(2.2393953277407236e-12, xr[k], 3.308302983832675e-9);
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = fma(g[k], xr[k], 3.665565596910102e-6);
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = fma(g[k], xr[k], 0.0027076061740622769) * xr[k];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double e2 = *reinterpret_cast<const double*>(etab + ((kk[k] << 7) & 0x7f80));
        const double v = fma(e2, g[k], e2);
        const double vs = __hiloint2double(__double2hiint(v) + ((kk[k] >> 8) << 20), __double2loint(v));
        acc[k] = fma(vs, __ldg(a + k), acc[k]);
    }
}

// Work item = (32-point tile, model filter); a warp owns an item, the thread a point.  The per-filter sums of a tile go
// through `parts` [tile][F][32]; the warp that finishes a tile's last filter (global ticket `tickets[tile]`, which it
// resets for the next launch) adds them in filter order, so the result does not depend on the schedule.
// FAST = sample grid is the (uniform) training grid itself (fp32 back end), else the generic fp64 back end.
template <int D, int K, bool FAST>
__global__ void __launch_bounds__(kGfWarps * 32, 1)
fused_gp_logl_kernel(const DevCfg cfg, const double* __restrict__ pts, long long N, double* __restrict__ out,
                     double* __restrict__ parts, unsigned int* __restrict__ tickets) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Ntr = cfg.Ntr, F = cfg.F;
    double* Xs = reinterpret_cast<double*>(smem + kGfLogBytes + kGfExpBytes);
    gf_tabs_fill(smem, tid, blockDim.x);
    for (int i = tid; i < Ntr * D; i += blockDim.x) Xs[i] = cfg.gpX[i];
    __syncthreads();
    const unsigned char* ltab = smem + (lane & 7) * 16;
    const unsigned char* etab = smem + kGfLogBytes + (lane & 15) * 8;

    const long long ntiles = (N + 31) / 32;
    const long long nitems = ntiles * F;
    for (long long item = (long long)blockIdx.x * kGfWarps + warp; item < nitems; item += (long long)gridDim.x * kGfWarps) {
        const long long tile = item / F;
        const int f = (int)(item - tile * F);
        const long long n = tile * 32 + lane;
        const bool live = n < N;
        const double* row = pts + (live ? n : N - 1) * cfg.P;
        const PointScal ps = point_setup(cfg, row);
        bool ok = !ps.bad && !cfg.static_fail;
        // the scaled input depends on the filter only through param_mins/maxs, which training.py:216-230 shares across
        // filters (checked on the host)
        double x[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            x[i] = scaled_input(cfg, 0, i, row);
            ok = ok && isfinite(x[i]);
        }
        double qk[K], nak[K], acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            qk[k] = cfg.gp_q[f * K + k];
            nak[k] = -256.0 * cfg.gp_ra[f * K + k];
            acc[k] = 0.0;
        }
        const double* __restrict__ a = cfg.gpAT + (size_t)f * Ntr * K;   // [Ntr][K]: the K alphas of a row are adjacent
        const double* __restrict__ xt = Xs;
#pragma unroll 1
        for (int t = 0; t < Ntr; ++t, a += K, xt += D) {
            double r2 = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const double df = x[i] - xt[i];
                r2 = fma(df, df, r2);
            }
            gf_pow_row<K>(r2, qk, nak, a, acc, ltab, etab);
        }
        double c[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            c[k] = cfg.gp_ys[f * K + k] * acc[k] + cfg.gp_ym[f * K + k];
            ok = ok && isfinite(c[k]);
        }
        double lsum = CUDART_NAN;
        if (ok) lsum = fused_filter_logl<K, FAST>(cfg, f, c, ps, row, basis_src<FAST>(cfg, f, K), cfg.o_pack, cfg.samp);
        // ---- hand the filter's sum in; the last filter of the tile to arrive adds them up in filter order ----
        double* tp = parts + (size_t)tile * F * 32;
        __stcg(tp + f * 32 + lane, lsum);
        __threadfence();
        __syncwarp();
        unsigned int old = 0;
        if (lane == 0) old = atomicAdd(&tickets[tile], 1u);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == (unsigned)(F - 1)) {
            __threadfence();
            double s = 0.0;
            for (int j = 0; j < F; ++j) s += __ldcg(tp + j * 32 + lane);
            if (live) out[n] = isfinite(s) ? s : NMMA_SENTINEL;
            if (lane == 0) tickets[tile] = 0;   // ready for the next launch on this handle
        }
    }
}

}  // namespace nmma
