// Hybrid throughput kernel: layer 1 of the surrogate MLP on the CUDA cores (exact fp32 FFMA, like Keras), layer 2 on
// tcgen05 tensor cores (3xTF32 split, accumulators and the activation operand in TMEM), fused with the fp64 likelihood.
//
// Why hybrid.  The all-tensor-core kernel (tc_kernel.cuh) is a latency chain: layer-1 MMA -> tcgen05.ld -> ReLU/split
// -> tcgen05.st -> layer-2 MMA, two chunks in flight per tile, tensor pipe 35 % busy
// (profiles/r01_fused_tc_v2_summary.json).  Layer 1 is only d+1 = 5 of the 15 multiply-adds per hidden unit, so here
// the CUDA cores compute relu(W1 x + b1) directly in registers (packed fma.rn.f32x2 over hidden-unit pairs, the input
// as the scalar-broadcast operand) and only *write* the split activations to TMEM; nothing waits for an MMA result
// except the once-per-64-hidden partial read.  The 10-wide layer-2 contraction - the part that made the FFMA kernel
// shared-memory-broadcast bound - stays on the tensor cores.
//
// Numerics.  h = h_hi + h_lo, W2 = W_hi + W_lo with hi = the 19 bits kind::tf32 reads; h_hi*W_hi + h_lo*W_hi +
// h_hi*W_lo accumulate in TMEM in chains of 24 MMAs (one 64-hidden group) that the CUDA cores sum with round-to-nearest
// adds (the tensor core's accumulator add rounds toward zero, profiles/r01_tc_probe.txt).  Layer 1 is bit-identical
// to the FFMA kernels.
//
// Work decomposition (one persistent CTA per SM, 19 warps, 8 tiles of 128 points = 8 x 64 TMEM columns in flight):
//   warps 0-15        compute warps.  Warp w: TMEM lane quadrant q = w & 3, chunk parity par = (w >> 2) & 1, tile set
//                     g = w >> 3 (tiles 4g..4g+3).  Thread = lane l of the four tiles.  Layer 1: the warp handles the
//                     8-hidden chunks of its parity for all four tiles (weights as warp-uniform LDS.128 broadcasts
//                     amortised over 4 points, 4 hidden pairs x 4 points of FFMA2, ReLU, hi/lo split, one
//                     tcgen05.st.x16 per tile into the A buffer of its parity).  Layer-2 partials and the fp64 back end
//                     (fused_filter_logl): the warp owns tiles 4g+2par and 4g+2par+1.
//   warps 16 / 17     MMA issuer of tile set 0 / 1 (one elected lane): per chunk and tile 3 MMAs 128x16x8.
//   warp 18           TMA producer (9.5 KB weight groups through a shared-memory ring, basis packs); TMEM allocation.
// All hand-offs are mbarriers (tcgen05.commit for MMA completion).
#pragma once
#include "tc_kernel.cuh"

namespace nmma {

constexpr int kHyThreads = 608;
constexpr int kHyActThreads = 512;
constexpr int kHyOwn = 2;                    // tiles whose partials / back end a compute warp owns
constexpr int kHyPT = 4;                     // tiles (= points) per compute thread
constexpr int kHyWgs = 2;
constexpr int kHyTiles = kHyPT * kHyWgs;     // 8 tiles x 64 TMEM columns
constexpr int kHySuper = kHyTiles * kTcTile; // 1024 points per CTA pass
constexpr int kHyChunk = 8;                  // hidden units per hand-off (one K = 8 MMA step)
constexpr int kHyGroup = 8;                  // chunks per ring slot / per D2main chain (64 hidden units)
constexpr int kHyStages = 6;
// TMEM columns of one tile (tile T at column 64 T)
constexpr uint32_t kHyColA = 0;              // 2 x [h (hi read by the tensor core) 8 | h_lo 8], double buffered by chunk
constexpr uint32_t kHyColMain = 32;          // 2 x 16  layer-2 partials, double buffered by group
constexpr int kHyB2Floats = kHyGroup * kTcN2 * kHyChunk;  // 1024 floats per hi / lo half of a group

__host__ __device__ constexpr int hy_w1_floats(int D) { return kHyGroup * kHyChunk * (D + 1); }
__host__ __device__ constexpr int hy_slot_floats(int D) { return hy_w1_floats(D) + 2 * kHyB2Floats; }
__host__ __device__ inline size_t hy_smem_bytes(int D, int K, int T, int S, int nobs) {
    const size_t w = (size_t)kHyStages * hy_slot_floats(D) * sizeof(float);
    const size_t o = ((size_t)nobs * kObsRec * sizeof(double) + 127) / 128 * 128;
    const size_t sg = ((size_t)S * sizeof(double) + 127) / 128 * 128;
    const size_t ps = (size_t)kHyOwn * 4 * kHyActThreads * sizeof(double);
    const size_t cb = (size_t)kHyOwn * K * kHyActThreads * sizeof(float);
    return w + 2 * fused_bslot(K, T) + o + sg + ps + cb + 512;
}

struct HyBars {
    uint64_t w_full[kHyStages], w_free[kHyStages];
    uint64_t b_full[2], b_free[2];
    uint64_t a2_full[kHyWgs][2], a2_free[kHyWgs][2];
    uint64_t d2_full[kHyWgs][2], d2_free[kHyWgs][2];
    uint32_t tmem_base;
};

// shared-window (32-bit) addressing for the hot loop: keeps ptxas from re-deriving the window base (S2UR) per access
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int D, int K, bool FAST>
__global__ void __launch_bounds__(kHyThreads, 1)
fused_hy_logl_kernel(const DevCfg cfg, const double* __restrict__ pts, long long N, double* __restrict__ out) {
    static_assert(K <= kTcN2, "n_coeff must fit the N=16 layer-2 MMA");
    constexpr int SLOT = hy_slot_floats(D);
    constexpr int W1F = hy_w1_floats(D);
    constexpr int CW = kHyChunk * (D + 1);   // W1 floats per chunk: 4 hidden pairs x [b, b', w0, w0', ...]
    static_assert(CW % 4 == 0, "chunk weights are read as float4");
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t wbytes = (size_t)kHyStages * SLOT * sizeof(float);
    const size_t bslot = fused_bslot(K, cfg.T);
    const uint32_t bbytes = (uint32_t)(cfg.T * (K + 2) * sizeof(double));
    const size_t obytes = ((size_t)cfg.nobs * kObsRec * sizeof(double) + 127) / 128 * 128;
    const size_t sbytes = ((size_t)cfg.S * sizeof(double) + 127) / 128 * 128;
    const size_t psbytes = (size_t)kHyOwn * 4 * kHyActThreads * sizeof(double);
    const size_t cbytes = (size_t)kHyOwn * K * kHyActThreads * sizeof(float);
    float* wring = reinterpret_cast<float*>(smem);
    unsigned char* s_basis0 = smem + wbytes;
    double* s_obs = reinterpret_cast<double*>(smem + wbytes + 2 * bslot);
    double* s_samp = reinterpret_cast<double*>(smem + wbytes + 2 * bslot + obytes);
    double* s_ps = reinterpret_cast<double*>(smem + wbytes + 2 * bslot + obytes + sbytes);      // [own][field][thread]
    float* s_c = reinterpret_cast<float*>(smem + wbytes + 2 * bslot + obytes + sbytes + psbytes);  // [own][k][thread]
    HyBars* bars = reinterpret_cast<HyBars*>(smem + wbytes + 2 * bslot + obytes + sbytes + psbytes + cbytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int F = cfg.F, NG = cfg.hy_ngrp;
    const long long nsuper = (N + kHySuper - 1) / kHySuper;
    const long long my_super = (nsuper > blockIdx.x) ? (nsuper - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == 0) {
        for (int i = 0; i < kHyStages; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_free[i], 16 + kHyWgs); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_free[i], 16); }
        for (int g = 0; g < kHyWgs; ++g) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(&bars->a2_full[g][b], 4);
                mbar_init(&bars->a2_free[g][b], 1);
                mbar_init(&bars->d2_full[g][b], 1);
                mbar_init(&bars->d2_free[g][b], 8);
            }
        }
        mbar_fence_init();
    }
    if (warp == 18) tmem_alloc(&bars->tmem_base, 512);
    for (int i = tid; i < cfg.nobs * kObsRec; i += kHyThreads) s_obs[i] = cfg.o_pack[i];
    for (int i = tid; i < cfg.S; i += kHyThreads) s_samp[i] = cfg.samp[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp < 16) {
        // =====================================================================================================
        // compute warps
        // =====================================================================================================
        const int q = warp & 3, par = (warp >> 2) & 1, g = warp >> 3;
        const int pidx = q * 32 + lane;                     // TMEM lane = point within each tile
        const int at = tid;                                 // 0..511: private column of s_ps / s_c
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * kHyPT * g);
        const uint32_t wring_s = smem_u32(wring);
        const uint32_t a2_full_s = smem_u32(&bars->a2_full[g][par]), a2_free_s = smem_u32(&bars->a2_free[g][par]);
        uint32_t gq = 0;                                    // (super-tile, filter, group) sequence number
        uint32_t vseq = 0;
        for (long long it = 0; it < my_super; ++it) {
            const long long sup = blockIdx.x + it * gridDim.x;
            const long long n0 = sup * kHySuper + (long long)(kHyPT * g) * kTcTile + pidx;   // point of tile 4g
            double logl[kHyOwn];
            unsigned okmask = 0;                            // bit t: point of tile 4g + t still valid
#pragma unroll
            for (int o = 0; o < kHyOwn; ++o) {
                const long long n = n0 + (long long)(kHyOwn * par + o) * kTcTile;
                const double* row = pts + (n < N ? n : 0) * cfg.P;
                const PointScal ps = point_setup(cfg, row);
                s_ps[(o * 4 + 0) * kHyActThreads + at] = ps.z1;
                s_ps[(o * 4 + 1) * kHyActThreads + at] = ps.ts;
                s_ps[(o * 4 + 2) * kHyActThreads + at] = ps.dm;
                s_ps[(o * 4 + 3) * kHyActThreads + at] = ps.zc;
                if (!ps.bad && !cfg.static_fail) okmask |= 1u << (kHyOwn * par + o);
                logl[o] = 0.0;
            }
            for (int f = 0; f < F; ++f, ++vseq) {
                // ---- scaled inputs of the four points (fp64 scaling, fp32 cast like Keras) ----
                float x[kHyPT][D];
#pragma unroll
                for (int t = 0; t < kHyPT; ++t) {
                    const long long n = n0 + (long long)t * kTcTile;
                    const double* row = pts + (n < N ? n : 0) * cfg.P;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        const double xs = scaled_input(cfg, f, i, row);
                        if (!isfinite(xs)) okmask &= ~(1u << t);
                        x[t][i] = (float)xs;
                    }
                }
                float acc[kHyOwn][K];
#pragma unroll
                for (int o = 0; o < kHyOwn; ++o)
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[o][k] = 0.f;

                auto read_partial = [&](uint32_t G) {   // layer-2 partial of group G, own tiles -> acc (RN adds)
                    const uint32_t p = G & 1;
                    mbar_wait(&bars->d2_full[g][p], (G >> 1) & 1);
                    tc_fence_after();
                    uint32_t part[kHyOwn][16];
#pragma unroll
                    for (int o = 0; o < kHyOwn; ++o)
                        tmem_ld16(tbase + 64 * (kHyOwn * par + o) + kHyColMain + 16 * p, part[o]);
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->d2_free[g][p]);
#pragma unroll
                    for (int o = 0; o < kHyOwn; ++o)
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[o][k] += __uint_as_float(part[o][k]);
                };

                for (int grp = 0; grp < NG; ++grp, ++gq) {
                    const uint32_t st = gq % kHyStages;
                    mbar_wait(&bars->w_full[st], (gq / kHyStages) & 1);
                    const uint32_t wv = wring_s + st * (uint32_t)(SLOT * 4);
                    bool pending = false;
#pragma unroll 1
                    for (int ch = par; ch < kHyGroup; ch += 2) {
                        const uint32_t u = (gq * kHyGroup + ch) >> 1;   // use count of this warp's A buffer
                        float w[CW];
#pragma unroll
                        for (int j = 0; j < CW / 4; ++j) {
                            const float4 v = lds128(wv + (uint32_t)(ch * CW + 4 * j) * 4);  // warp-uniform: broadcast
                            w[4 * j + 0] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
                        }
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            uint32_t o[2][16];
#pragma unroll
                            for (int tt = 0; tt < 2; ++tt) {
                                const int t = 2 * half + tt;
#pragma unroll
                                for (int p = 0; p < kHyChunk / 2; ++p) {
                                    const float* wp = w + p * 2 * (D + 1);
                                    float2 h = make_float2(wp[0], wp[1]);
#pragma unroll
                                    for (int i = 0; i < D; ++i)
                                        h = __ffma2_rn(make_float2(x[t][i], x[t][i]), make_float2(wp[2 + 2 * i], wp[3 + 2 * i]), h);
                                    h.x = fmaxf(h.x, 0.f);
                                    h.y = fmaxf(h.y, 0.f);
                                    const float2 hh = make_float2(__uint_as_float(__float_as_uint(h.x) & 0xFFFFE000u),
                                                                  __uint_as_float(__float_as_uint(h.y) & 0xFFFFE000u));
                                    const float2 lo = __ffma2_rn(hh, make_float2(-1.f, -1.f), h);   // h - hi, exact
                                    o[tt][2 * p] = __float_as_uint(h.x);
                                    o[tt][2 * p + 1] = __float_as_uint(h.y);
                                    o[tt][8 + 2 * p] = __float_as_uint(lo.x);
                                    o[tt][8 + 2 * p + 1] = __float_as_uint(lo.y);
                                }
                            }
                            if (half == 0) {
                                if (pending) {
                                    // publish the previous chunk: its tcgen05.st were issued half a chunk of FFMA work ago
                                    tmem_wait_st();
                                    tc_fence_before();
                                    __syncwarp();
                                    if (lane == 0) mbar_arrive_s(a2_full_s);
                                }
                                if (u >= 1) {
                                    // the MMAs of this buffer's previous chunk have finished reading it
                                    mbar_wait_s(a2_free_s, (u - 1) & 1);
                                    tc_fence_after();
                                }
                            }
                            tmem_st16(tbase + 64 * (2 * half) + kHyColA + 16 * par, o[0]);
                            tmem_st16(tbase + 64 * (2 * half + 1) + kHyColA + 16 * par, o[1]);
                        }
                        pending = true;
                    }
                    // last chunk of the group: publish now (the partial read below needs the group's MMAs anyway)
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive_s(a2_full_s);
                        mbar_arrive(&bars->w_free[st]);
                    }
                    if (grp >= 1) read_partial(gq - 1);
                }
                read_partial(gq - 1);
                // ---- coefficients of the own tiles: summed partials + b2 (fp32, Keras Dense) -> shared memory ----
#pragma unroll
                for (int o = 0; o < kHyOwn; ++o)
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        s_c[(o * K + k) * kHyActThreads + at] = acc[o][k] + cfg.b2[f * K + k];
                // ---- fp64 back end for the two own points ----
                const int slot = (int)(vseq & 1);
                mbar_wait(&bars->b_full[slot], (vseq >> 1) & 1);
                const double* basis = reinterpret_cast<const double*>(s_basis0 + slot * bslot);
#pragma unroll 1
                for (int o = 0; o < kHyOwn; ++o) {
                    const int t = kHyOwn * par + o;
                    double cp[K];
                    bool fin = true;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const float cf = s_c[(o * K + k) * kHyActThreads + at];
                        fin = fin && isfinite(cf);
                        cp[k] = (double)cf;
                    }
                    if (!fin) okmask &= ~(1u << t);
                    if ((okmask >> t) & 1u) {
                        PointScal ps;
                        ps.z1 = s_ps[(o * 4 + 0) * kHyActThreads + at];
                        ps.ts = s_ps[(o * 4 + 1) * kHyActThreads + at];
                        ps.dm = s_ps[(o * 4 + 2) * kHyActThreads + at];
                        ps.zc = s_ps[(o * 4 + 3) * kHyActThreads + at];
                        ps.bad = false;
                        point_fast_fields(cfg, ps);
                        const long long nn = n0 + (long long)t * kTcTile;
                        const double* row = pts + (nn < N ? nn : 0) * cfg.P;
                        const double v = fused_filter_logl<K, FAST>(cfg, f, cp, ps, row, basis, s_obs, s_samp);
#pragma unroll
                        for (int u = 0; u < kHyOwn; ++u)
                            if (u == o) logl[u] += v;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->b_free[slot]);
            }
#pragma unroll
            for (int o = 0; o < kHyOwn; ++o) {
                const int t = kHyOwn * par + o;
                const long long n = n0 + (long long)t * kTcTile;
                if (n < N) out[n] = (((okmask >> t) & 1u) && isfinite(logl[o])) ? logl[o] : NMMA_SENTINEL;
            }
        }
    } else if (warp < 16 + kHyWgs) {
        // =====================================================================================================
        // MMA issuer of warpgroup g
        // =====================================================================================================
        const int g = warp - 16;
        constexpr uint32_t id2 = tc_idesc(kTcN2);
        const uint32_t wbase = smem_u32(wring);
        const uint64_t dB = tc_smem_desc(wbase + W1F * 4, kTcN2 * 16, 128);
        const uint32_t bhi = (uint32_t)(dB >> 32);
        const uint32_t blo0 = (uint32_t)dB;
        constexpr uint32_t kSlotStep = (uint32_t)(SLOT * 4) >> 4;
        constexpr uint32_t kChunkStep = (uint32_t)(kTcN2 * kHyChunk * 4) >> 4;   // 512 B per chunk tile
        constexpr uint32_t kLoOff = (uint32_t)(kHyB2Floats * 4) >> 4;
        const uint32_t tb = tmem + (uint32_t)(64 * kHyPT * g);
        uint32_t gq = 0;
        const long long total_v = my_super * F;
        for (long long vv = 0; vv < total_v; ++vv) {
            for (int grp = 0; grp < NG; ++grp, ++gq) {
                const uint32_t st = gq % kHyStages;
                const uint32_t p = gq & 1;
                mbar_wait(&bars->w_full[st], (gq / kHyStages) & 1);
                if (gq >= 2) mbar_wait(&bars->d2_free[g][p], ((gq >> 1) - 1) & 1);
                const uint32_t bslot_lo = blo0 + st * kSlotStep;
#pragma unroll 1
                for (int ch = 0; ch < kHyGroup; ++ch) {
                    const uint32_t cq = gq * kHyGroup + ch;
                    mbar_wait(&bars->a2_full[g][cq & 1], (cq >> 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t bh = bslot_lo + ch * kChunkStep;
                        const uint32_t bl = bh + kLoOff;
#pragma unroll
                        for (int t = 0; t < kHyPT; ++t) {
                            const uint32_t d2 = tb + 64 * t + kHyColMain + 16 * p;
                            const uint32_t a = tb + 64 * t + kHyColA + 16 * (cq & 1);
                            mma_tf32_ts(d2, a, bh, bhi, id2, ch > 0 ? 1u : 0u);
                            mma_tf32_ts(d2, a + 8, bh, bhi, id2, 1u);
                            mma_tf32_ts(d2, a, bl, bhi, id2, 1u);
                        }
                        tc_commit(&bars->a2_free[g][cq & 1]);
                        if (ch == kHyGroup - 1) {
                            tc_commit(&bars->d2_full[g][p]);
                            tc_commit(&bars->w_free[st]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 18) {
        // =====================================================================================================
        // TMA producer
        // =====================================================================================================
        uint32_t st = 0, ph = 0, vseq = 0;
        int f = 0;
        const long long total_v = my_super * F;
        constexpr uint32_t kSlotBytes = (uint32_t)(SLOT * 4);
        for (long long vv = 0; vv < total_v; ++vv, ++vseq) {
            const int slot = (int)(vseq & 1);
            mbar_wait(&bars->b_free[slot], ((vseq >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars->b_full[slot], bbytes);
                bulk_g2s(s_basis0 + slot * bslot, basis_src<FAST>(cfg, f, K), bbytes, &bars->b_full[slot]);
            }
            __syncwarp();
            const float* src = cfg.hypack + (size_t)f * NG * SLOT;
            for (int grp = 0; grp < NG; ++grp) {
                mbar_wait(&bars->w_free[st], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bars->w_full[st], kSlotBytes);
                    bulk_g2s(wring + (size_t)st * SLOT, src + (size_t)grp * SLOT, kSlotBytes, &bars->w_full[st]);
                }
                __syncwarp();
                if (++st == kHyStages) { st = 0; ph ^= 1; }
            }
            if (++f == F) f = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 18) tmem_dealloc(tmem, 512);
}

}  // namespace nmma
