// Tensor-core throughput kernel: the surrogate MLP on tcgen05 (5th-gen tensor cores, accumulators and the
// activation operand in TMEM), split-precision fp16 operands (kind::f16), fused with the fp64 likelihood back end.
//
// Why split precision: single-pass TF32/BF16 misses the 1e-3 mag budget by two orders of magnitude
// (profiles/r01_tc_numerics.txt); with a = a_hi + a_lo, b = b_hi + b_lo (fp16 halves: 11 + 11 significant bits) the
// three products a_hi*b_hi + a_lo*b_hi + a_hi*b_lo carry ~2^-21 relative error, the order of fp32 FFMA
// (tools/tc_numerics.py "f16x3": on the trained fixture weights it is as close to fp64 as NumPy's fp32).  fp16 has a
// 5-bit exponent, so every operand is brought into range by EXACT power-of-two scalings:
//   * layer 1: row i of [W1; b1] is staged times 2^-r_i so that its largest entry is in [2^9, 2^10) and input i is
//     multiplied by 2^r_i (cfg.tc_xs); per point and filter the whole input row is scaled by 2^e with
//     2^e sum_i |x_i 2^r_i| in [8, 16), so 2^e relu(v) < 2^14 whatever the point is (ReLU is positively homogeneous:
//     the coefficients are multiplied by 2^-e at the end);
//   * layer 2: column k of W2 is staged times 2^q_k (largest entry in [8, 16), cfg.tc_s2inv = 2^-q_k), its fp16
//     remainder times 2^11; products of two fp16 values are exact in the fp32 accumulator.
// The hi/lo split of h costs 2 CUDA-core instructions per hidden unit and point (3.5 with tf32 operands, round 1):
//   hi2 = cvt.rz.relu.f16x2.f32 (v1, v0)         ReLU and round-toward-zero in one F2FP, so that v - hi >= 0 for v > 0
//   l   = fma.rn.f32.f16 (hi, -1, v)             FHFMA: fp16 operand taken straight from the packed register
//   lo2 = cvt.rn.relu.f16x2.f32 (l1, l0)         for v < 0: hi = 0, l = v < 0 -> 0
// The tensor core adds into its fp32 accumulator with round-toward-zero (profiles/r01_tc_probe.txt), so layer 2
// accumulates in chains of kTcGroup chunks whose partials the CUDA cores sum with round-to-nearest adds.
//
// What bounds it (profiles/r01_tc_experiments.md, r02_tc_f16_experiments.md): not the tensor pipe (25 % active), not TMEM
// bandwidth, not the weight stream, but the activation warps' dispatch-port time (3.0 issue cycles per hidden unit and point
// for the split -- F2FP holds the port for 2 -- plus their tcgen05.ld / st) and the per-chunk hand-offs.  Hence: h_hi / h_lo go
// back IN PLACE over the layer-1 accumulator they came from (one tcgen05.st per 32 hidden units), [W_hi | 2^11 W_lo] is ONE
// N = 32 B tile (h_hi needs one MMA per 16 hidden units for both terms), and K = 16 per MMA: 9 MMAs per 64-hidden chunk
// instead of 23 with tf32 operands.  More activation-warp sets, more TMEM buffers, other chunk widths (TCV_SETS, TCV_BUFS,
// TCV_CHUNK, TCV_GROUP: compile-time experiments, only the defaults are parity-tested on every path) measured no faster.
//
// Work decomposition (one persistent CTA per SM, 20 warps, two 128-point tiles in flight):
//   warps 0-3 / 4-7    activation warps of tile 0 / 1: thread = one parameter point = one TMEM lane.  Per filter they
//                      write the scaled, split inputs as the layer-1 A operand (TMEM); per chunk they read the
//                      layer-1 accumulator (tcgen05.ld), apply ReLU + split, write [h_hi | h_lo] back in place
//                      (tcgen05.st, the layer-2 A operands) and sum the layer-2 group partials; the K
//                      coefficients go to shared memory.
//   warps 8 / 9        MMA issuer of tile 0 / 1 (one elected lane): per chunk D2 += h_hi . [W_hi | W_lo']^T (N = 32) and
//                      D2[0:16] += h_lo . W_hi^T (N = 16), and, two chunks ahead, D1 = A1 . B1^T (1-2 MMAs, N = chunk);
//                      A from TMEM, B from shared memory.
//   warp 10            TMA producer: 8 KB weight chunks (cp.async.bulk) into a shared-memory ring, basis packs per filter.
//   warp 11            TMEM allocation / release, watchdog.
//   warps 12-15/16-19  back-end warps of tile 0 / 1: thread = one point; reconstruction, interpolation and likelihood
//                      (kernels.cuh: fused_filter_logl) for the filter whose coefficients the activation warps just
//                      finished, overlapping the next filter's MLP.
// All hand-offs are mbarriers (tcgen05.commit for MMA completion); nothing spins on memory.
#pragma once
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace nmma {

#ifndef TCV_SETS
#define TCV_SETS 1
#endif
#ifndef TCV_BUFS
#define TCV_BUFS 2
#endif
constexpr int kTcSets = TCV_SETS;            // activation-warp sets per tile: set s takes the chunks c = s (mod kTcSets)
constexpr int kTcBufs = TCV_BUFS;            // layer-1 accumulator / layer-2 operand buffers per tile: chunk c uses buffer c % kTcBufs
constexpr int kTcActWarps = 8 * kTcSets;     // sets x two tiles x four TMEM lane quadrants
constexpr int kTcProdWarp = kTcActWarps + 2;
constexpr int kTcOwnerWarp = kTcActWarps + 3;
constexpr int kTcBackWarp0 = kTcActWarps + 4;
constexpr int kTcThreads = 32 * (kTcBackWarp0 + 8);
constexpr int kTcTile = 128;                 // points per tile = TMEM lanes
constexpr int kTcTiles = 2;                  // tiles in flight per CTA
#ifndef TCV_CHUNK
#define TCV_CHUNK 64
#endif
constexpr int kTcChunk = TCV_CHUNK;          // hidden units per chunk = per act <-> issuer hand-off (N of the layer-1 MMA)
#ifndef TCV_BLK
#define TCV_BLK ((TCV_CHUNK % 32 == 0) ? 32 : 16)
#endif
constexpr int kTcBlk = TCV_BLK;   // columns per ReLU/split/store block: [h_hi pairs | h_lo pairs] in place
constexpr int kTcKSteps = kTcChunk / 16;     // kind::f16 layer-2 k-steps (K = 16) per chunk
constexpr int kTcN2 = 16;                    // layer-2 MMA N per term (n_coeff padded)
constexpr int kTcK1Max = 2;                  // layer-1 k-steps staged (3 (d + 1) slots: d <= 4 uses one, d <= 7 two)
constexpr int kTcB1Halfs = kTcChunk * 16;    // one layer-1 B tile: [chunk hidden] x [16 K slots] fp16
constexpr int kTcB2Halfs = 2 * kTcN2 * 16;   // one layer-2 B tile: [W_hi (16 coeff) | 2^11 W_lo (16 coeff)] x [16 hidden] fp16
constexpr int kTcChunkFloats = (kTcK1Max * kTcB1Halfs + kTcKSteps * kTcB2Halfs) / 2;   // B1[0] | B1[1] | B2[k-steps] = 8 KB at 64
constexpr uint32_t kTcChunkBytes = kTcChunkFloats * 4;
#ifndef TCV_RING_KB
#define TCV_RING_KB 48
#endif
constexpr int kTcStages = (TCV_RING_KB * 1024) / (int)kTcChunkBytes;   // weight ring depth (layer 1 runs kTcBufs chunks ahead of layer 2)
// TMEM columns of one tile (tile t at column 256 t): everything layer 2 reads is written IN PLACE over the layer-1
// accumulator block it came from (block of kTcBlk columns -> kTcBlk / 2 columns of h_hi fp16 pairs, then kTcBlk / 2 of h_lo).
constexpr uint32_t kColD1 = 0;                       // bufs x chunk  layer-1 accumulators / layer-2 A operands
constexpr uint32_t kColD2 = kTcBufs * kTcChunk;      // 2 x 32        layer-2 partials [h W_hi | h_hi W_lo'], double buffered by group
constexpr uint32_t kColA1 = kTcBufs * kTcChunk + 64; // 16            layer-1 A operand: fp16 pairs, 8 columns per k-step
static_assert(kColA1 + 16 <= 256, "TMEM columns per tile");
static_assert(kTcChunk % 16 == 0 && kTcChunk >= 16 && kTcChunk <= 256, "layer-1 MMA N");
#ifndef TCV_GROUP
#define TCV_GROUP 4
#endif
constexpr int kTcGroup = TCV_GROUP;          // chunks per layer-2 accumulation chain (2 kTcKSteps MMAs each, RZ accumulate)
// The partial of the group that chunk e closes is read while chunk e + kTcBufs is processed (its d1_full implies that layer 2
// of chunk e is complete), by the set that owns that chunk: it must be the last set (it keeps the sums), and the read must come
// before the group two further on restarts the accumulator at chunk e + kTcGroup + 1.
static_assert(kTcBufs <= kTcGroup, "a group partial would be overwritten before it is read");
static_assert(kTcSets == 1 || (kTcSets == 2 && kTcGroup % 2 == 0 && kTcBufs % 2 == 0), "groups must close on the last set's chunks");
static_assert((kTcBufs & (kTcBufs - 1)) == 0 && (kTcGroup & (kTcGroup - 1)) == 0 && kTcGroup <= 8, "powers of two");
constexpr int kTcUnit = kTcGroup;            // chunk counts (per filter, per hidden range) are multiples of this (>= kTcBufs)
static_assert(kTcStages >= kTcBufs + 2, "weight ring too shallow");
// TMEM column (relative to the chunk buffer) of the h_hi pairs of layer-2 k-step s; its h_lo pairs are kTcBlk / 2 further
__host__ __device__ constexpr uint32_t tc_a2_col(int s) { return (uint32_t)(kTcBlk * (s / (kTcBlk / 16)) + 8 * (s % (kTcBlk / 16))); }

// ---- tcgen05 wrappers (PTX forms as in cute/arch/mma_sm100_umma.hpp, copy_sm100.hpp, tmem_allocator_sm100.hpp) ----
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::tf32, M = 128.  The matrix descriptor is passed as two 32-bit halves: only
// the low half (start address) changes between MMAs, so the issuer's address arithmetic stays 32-bit.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ReLU + hi/lo split of two layer-1 outputs into fp16 pairs (.x = low half = even hidden unit), 4 instructions:
// hi = RZ(relu(v)) (one F2FP.RELU..RZ), v - hi in fp32 with the fp16 operand taken from the packed register (two FHFMA;
// mixed-precision fma, PTX ISA 8.6 / sm_100), lo = RN(relu(v - hi)) (one F2FP.RELU): RZ makes v - hi >= 0 for v > 0, and
// for v <= 0 hi = 0 and v - hi = v <= 0 -> 0.
__device__ __forceinline__ void relu_split_f16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    float l0, l1;
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    asm("{\n\t.reg .b16 a, b, m;\n\t"
        "mov.b32 {a, b}, %2;\n\t"
        "mov.b16 m, 0xBC00;\n\t"   // -1.0
        "fma.rn.f32.f16 %0, a, m, %3;\n\t"
        "fma.rn.f32.f16 %1, b, m, %4;\n\t}"
        : "=f"(l0), "=f"(l1)
        : "r"(hi), "f"(v0), "f"(v1));
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
}
// round-to-nearest hi/lo split of a (range-checked) fp32 value into two fp16 values
__device__ __forceinline__ void split_f16(float a, __half& hi, __half& lo) {
    hi = __float2half_rn(a);
    lo = __float2half_rn(a - __half2float(hi));
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(addr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(addr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr)
        : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1): core matrix = 8 rows of
// 16 B, rows 16 B apart; lbo = bytes between the two 16-byte K halves, sbo = bytes between 8-row groups.
__host__ __device__ inline uint64_t tc_smem_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// kind::tf32 instruction descriptor: fp32 accumulate, A/B tf32, both K-major, M = 128 (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t tc_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// kind::f16 instruction descriptor: fp32 accumulate, A/B fp16, both K-major, M = 128, K = 16.
__host__ __device__ constexpr uint32_t tc_idesc_f16(int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// half index of element (n, k) of an N x 16 fp16 B-operand tile (same core-matrix geometry: 8 rows x 16 bytes)
__host__ __device__ constexpr int tc_b_index16(int N, int n, int k) { return (k >> 3) * (N * 8) + n * 8 + (k & 7); }
// float index of element (n, k) of an N x 8 B-operand tile
__host__ __device__ constexpr int tc_b_index(int N, int n, int k) { return (k >> 2) * (N * 4) + n * 4 + (k & 3); }

__host__ __device__ inline size_t tc_cbuf_bytes() { return (size_t)kTcTiles * 2 * kTcN2 * kTcTile * sizeof(float); }
__host__ __device__ inline size_t tc_smem_bytes(int K, int T, int S, int nobs) {
    const size_t w = (size_t)kTcStages * kTcChunkBytes;
    const size_t o = ((size_t)nobs * kObsRec * sizeof(double) + 127) / 128 * 128;
    const size_t sg = ((size_t)S * sizeof(double) + 127) / 128 * 128;
    return w + 2 * fused_bslot(K, T) + o + sg + tc_cbuf_bytes() + 512 /* barriers, tmem base */;
}

struct TcBars {
    uint64_t w_full[kTcStages], w_free[kTcStages];
    uint64_t b_full[2], b_free[2];
    uint64_t a1_full[kTcTiles];
    uint64_t d1_full[kTcTiles][kTcBufs];
    uint64_t a2_full[kTcTiles][kTcBufs], a2_free[kTcTiles][kTcBufs];
    uint64_t c_full[kTcTiles][2], c_free[kTcTiles][2];
    uint32_t tmem_base;
    uint32_t progress;   // bumped by the back-end warps once per (tile, filter): the watchdog's sign of life
    uint32_t finished;   // role warps that have left their loops
};

#ifdef TCV_TIMELINE   // debug build (tools/build_variants.py): clock64 stamps of one filter pass of CTA 0, printed at the end
__device__ long long g_tc_tl[kTcTiles][64][10];
#define TC_STAMP(t, c, k) do { if (blockIdx.x == 0 && vseq == 11 && (c) < 64 && lane == 0 && (warp & 3) == 0) g_tc_tl[t][c][k] = clock64(); } while (0)
#define TC_STAMP_I(t, c, k) do { if (blockIdx.x == 0 && vseq == 11 && (c) < 64 && lane == 0) g_tc_tl[t][c][k] = clock64(); } while (0)
#else
#define TC_STAMP(t, c, k) do { } while (0)
#define TC_STAMP_I(t, c, k) do { } while (0)
#endif
#ifdef TCV_SLEEP_PROD
#define TC_WAIT_PROD(bar, par) mbar_wait_sleep(bar, par, TCV_SLEEP_PROD)
#else
#define TC_WAIT_PROD(bar, par) mbar_wait_spin(bar, par)
#endif
#ifdef TCV_SLEEP_BACK
#define TC_WAIT_BACK(bar, par) mbar_wait_sleep(bar, par, TCV_SLEEP_BACK)
#else
#define TC_WAIT_BACK(bar, par) mbar_wait_spin(bar, par)
#endif

// COEFF = true: front end only.  The back-end warps write the coefficients [N][F][cfg.K] (fp64, + b2) to `out` instead of
// scoring them: the generic back end (backend_logl_kernel / backend_mags_kernel) follows in a second launch.  This is the
// path of configurations the fused back end does not cover (averaged filters, n_coeff != 10): K is then the padded
// width kTcN2 and cfg.K the real one.
template <int K, bool FAST, bool SPLIT, bool COEFF = false>
__global__ void __launch_bounds__(kTcThreads, 1)
fused_tc_logl_kernel(const DevCfg cfg, const double* __restrict__ pts, long long N, double* __restrict__ out, int fsplit_arg,
                     int hsplit_arg) {
    const int fsplit = SPLIT ? fsplit_arg : 1;   // SPLIT = false: the throughput instantiation, index arithmetic folds away
    // COEFF + SPLIT only (one-point / small-batch latency, launch_tc.cu: launch_tc_coeff_parts): the hidden layer of a
    // filter is cut into hsplit chunk ranges as well, each CTA hands out the PARTIAL coefficient sums of its range as fp32
    // [N][F][hsplit][cfg.K] (no b2), and backend_logl_parts_kernel adds them in range order before it scores them
    const bool parts_mode = SPLIT && COEFF && hsplit_arg > 0;   // hsplit_arg = 0: plain coefficient mode (fp64, + b2)
    const int hsplit = parts_mode ? hsplit_arg : 1;
    // chunks per layer-2 accumulation chain: kTcGroup; in parts mode the minimum (kTcBufs), so that the hidden layer can be cut
    // into twice as many ranges (one-point latency).  Folds to the constant in the throughput instantiations.
    const int gsz = (COEFF && SPLIT && parts_mode) ? kTcBufs : kTcGroup;
    const int gsh = gsz == 1 ? 0 : (gsz == 2 ? 1 : (gsz == 4 ? 2 : 3));
    const int nparts = fsplit * hsplit;
    // Work item = (256-point super-tile, filter part): with fsplit > 1 (small batches, launch_tc.cu) the filters of one
    // super-tile are spread over fsplit CTAs and `out` receives the per-part sums [N][fsplit] (NaN = failed) that
    // combine_parts_kernel adds in a fixed order; with fsplit == 1 `out` is the final log-likelihood.
    static_assert(K <= kTcN2, "n_coeff must fit the N=16 layer-2 MMA");
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t wbytes = (size_t)kTcStages * kTcChunkBytes;
    const size_t bslot = fused_bslot(K, cfg.T);
    const uint32_t bbytes = (uint32_t)(cfg.T * (K + 2) * sizeof(double));
    const size_t obytes = ((size_t)cfg.nobs * kObsRec * sizeof(double) + 127) / 128 * 128;
    const size_t sbytes = ((size_t)cfg.S * sizeof(double) + 127) / 128 * 128;
    float* wring = reinterpret_cast<float*>(smem);
    unsigned char* s_basis0 = smem + wbytes;
    double* s_obs = reinterpret_cast<double*>(smem + wbytes + 2 * bslot);
    double* s_samp = reinterpret_cast<double*>(smem + wbytes + 2 * bslot + obytes);
    float* s_c = reinterpret_cast<float*>(smem + wbytes + 2 * bslot + obytes + sbytes);  // [tile][slot][k][point]
    TcBars* bars = reinterpret_cast<TcBars*>(smem + wbytes + 2 * bslot + obytes + sbytes + tc_cbuf_bytes());

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int F = cfg.F, NCHT = cfg.tc_nch, NCH = NCHT / hsplit;   // chunks per filter, per (filter, hidden range)
    const int Kr = COEFF ? cfg.K : K;   // real n_coeff
    constexpr int SUPER = kTcTile * kTcTiles;
    const long long nsuper = ((N + SUPER - 1) / SUPER) * nparts;   // work items
    const long long my_super = (nsuper > blockIdx.x) ? (nsuper - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // filters [f0, f1) of work item w: part p = w % fsplit takes p F / fsplit .. (p + 1) F / fsplit
#define TC_PART_F0(w) (SPLIT ? (int)(((w) % nparts / hsplit) * F / fsplit) : 0)
#define TC_PART_F1(w) (SPLIT ? (int)(((w) % nparts / hsplit + 1) * F / fsplit) : F)
#define TC_PART_HS(w) ((SPLIT && COEFF) ? (int)((w) % nparts % hsplit) : 0)
    const uint32_t uses = (uint32_t)NCH / kTcBufs;  // NCH is a multiple of kTcBufs: TMEM buffer = c % kTcBufs, its use count = vseq * uses + c / kTcBufs

    if (tid == 0) {
        for (int i = 0; i < kTcStages; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_free[i], kTcTiles); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_free[i], 8); }
        for (int t = 0; t < kTcTiles; ++t) {
            mbar_init(&bars->a1_full[t], 4);
            for (int b = 0; b < kTcBufs; ++b) {
                mbar_init(&bars->d1_full[t][b], 1);
                mbar_init(&bars->a2_full[t][b], 4);
                mbar_init(&bars->a2_free[t][b], 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(&bars->c_full[t][b], 4);
                mbar_init(&bars->c_free[t][b], 4);
            }
        }
        bars->progress = 0;
        bars->finished = 0;
        mbar_fence_init();
    }
    if (warp == kTcOwnerWarp) tmem_alloc(&bars->tmem_base, 512);
    if constexpr (!COEFF) {   // the coefficient mode has no back end: nothing to stage
        for (int i = tid; i < cfg.nobs * kObsRec; i += kTcThreads) s_obs[i] = cfg.o_pack[i];
        for (int i = tid; i < cfg.S; i += kTcThreads) s_samp[i] = cfg.samp[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp < kTcActWarps) {
        // =====================================================================================================
        // activation warps: set 0 owns the even chunks (TMEM buffer 0) and writes the layer-1 A operand, set 1 owns the odd
        // chunks (buffer 1), sums the layer-2 partials (accumulation groups close on odd chunks) and hands the coefficients on
        // =====================================================================================================
        const int set = kTcSets > 1 ? warp >> 3 : 0;
        const int t = (warp >> 2) & 1;
        const int pidx = (warp & 3) * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(256 * t);
        uint32_t vseq = 0;  // (super-tile, filter) sequence number of this CTA
        for (long long it = 0; it < my_super; ++it) {
            const long long item = blockIdx.x + it * gridDim.x;
            const long long n = (item / nparts) * SUPER + (long long)t * kTcTile + pidx;
            const double* row = pts + (n < N ? n : 0) * cfg.P;
            for (int f = TC_PART_F0(item), f1 = TC_PART_F1(item); f < f1; ++f, ++vseq) {
                const uint32_t ubase = vseq * uses;
                // ---- layer-1 A operand (fp64 scaling, fp32 cast like Keras; then the exact power-of-two scalings of the
                //      header comment and the fp16 hi/lo split): K slot 3 i + {0, 1, 2} = {hi_i, lo_i, hi_i} ----
                bool okx = true;
                float inv_sc = 1.f;
                {
                    float xv[8];
                    float S = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        xv[i] = 0.f;
                        if (i < cfg.d) {
                            const double xs = scaled_input(cfg, f, i, row);
                            okx = okx && isfinite(xs);
                            xv[i] = (float)xs * cfg.tc_xs[f * 8 + i];
                        } else if (i == cfg.d) {
                            xv[i] = cfg.tc_xs[f * 8 + i];
                        }
                        S += fabsf(xv[i]);
                    }
                    // 2^e S in [8, 16): e = 3 - ilogb(S), from the exponent field (clamped: a degenerate S only costs accuracy)
                    okx = okx && isfinite(S);
                    uint32_t eb = (__float_as_uint(S) >> 23) & 0xFFu;
                    eb = eb < 24u ? 24u : (eb > 230u ? 230u : eb);
                    inv_sc = __uint_as_float((eb - 3u) << 23);
                    if (set == 0) {
                        const float sc = okx ? __uint_as_float((257u - eb) << 23) : 0.f;
                        __half hs[16 * kTcK1Max];
#pragma unroll
                        for (int i = 0; i < 16 * kTcK1Max; ++i) hs[i] = __ushort_as_half((unsigned short)0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __half hi, lo;
                            split_f16(okx ? xv[i] * sc : 0.f, hi, lo);
                            hs[3 * i] = hi; hs[3 * i + 1] = lo; hs[3 * i + 2] = hi;
                        }
                        uint32_t a1[8 * kTcK1Max];
#pragma unroll
                        for (int i = 0; i < 8 * kTcK1Max; ++i)
                            a1[i] = (uint32_t)__half_as_ushort(hs[2 * i]) | ((uint32_t)__half_as_ushort(hs[2 * i + 1]) << 16);
                        if (kTcSets > 1 && vseq > 0) {   // the previous filter's last layer-1 MMA (chunk NCH - 1) has read the old operand
                            mbar_wait_spin(&bars->d1_full[t][kTcBufs - 1], (ubase - 1) & 1);
                            tc_fence_after();
                        }
                        tmem_st16(tbase + kColA1, a1);
                        tmem_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->a1_full[t]);
                    }
                }
                float acc[K];
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] = 0.f;
                for (int c = set; c < NCH; c += kTcSets) {
                    const int b = c & (kTcBufs - 1);
                    const uint32_t u = ubase + (uint32_t)c / kTcBufs;
                    const uint32_t dbuf = tbase + kColD1 + kTcChunk * b;
                    TC_STAMP(t, c, 0);
                    mbar_wait_spin(&bars->d1_full[t][b], u & 1);
                    tc_fence_after();
                    TC_STAMP(t, c, 1);
                    // ReLU + hi/lo split in kTcBlk-column blocks, [h_hi pairs | h_lo pairs] back in place.  The load of block
                    // blk + 1 is issued before block blk is computed (ptxas tracks each tcgen05.ld's registers on its own
                    // scoreboard: the wait below costs nothing for a load that is not consumed yet), the store of block blk is
                    // in flight while block blk + 1 is computed.
                    constexpr int NB = kTcChunk / kTcBlk;
                    uint32_t v[2][kTcBlk];
                    auto load_blk = [&](int blk, uint32_t* dst) {
                        if constexpr (kTcBlk == 32) tmem_ld32(dbuf + kTcBlk * blk, dst);
                        else tmem_ld16(dbuf + kTcBlk * blk, dst);
                    };
                    load_blk(0, v[0]);
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
                        uint32_t o[kTcBlk];
                        const uint32_t* vc = v[blk & 1];
#ifdef TCV_PREFETCH
                        if (blk + 1 < NB) load_blk(blk + 1, v[(blk + 1) & 1]);
#endif
                        tmem_wait_ld();
#ifndef TCV_NO_ALU   // TCV_*: compile-time timing experiments (tools/build_variants.py, profiles/r01_tc_experiments.md);
                     // a library built with any of them returns wrong numbers and only serves to time the skeleton
#pragma unroll
                        for (int j = 0; j < kTcBlk / 2; ++j)
                            relu_split_f16x2(__uint_as_float(vc[2 * j]), __uint_as_float(vc[2 * j + 1]), o[j], o[kTcBlk / 2 + j]);
#else
#pragma unroll
                        for (int j = 0; j < kTcBlk; ++j) o[j] = vc[j];
#endif
                        if constexpr (kTcBlk == 32) tmem_st32(dbuf + kTcBlk * blk, o);
                        else tmem_st16(dbuf + kTcBlk * blk, o);
#ifndef TCV_PREFETCH
                        if (blk + 1 < NB) load_blk(blk + 1, v[(blk + 1) & 1]);
#endif
                    }
                    TC_STAMP(t, c, 3);
                    if (c >= kTcBufs && ((c - kTcBufs) & (gsz - 1)) == gsz - 1) {
                        // layer 2 of chunk c - kTcBufs is complete (layer 1 of this chunk was queued behind it and d1_full has
                        // fired); it closed a group: add its partials.  The issuer reuses that accumulator only after
                        // this warp's a2_full of a later chunk (it takes the chunks in order).
                        mbar_wait_spin(&bars->a2_free[t][b], (u - 1) & 1);
                        tc_fence_after();
                        uint32_t part[16], px[16];
                        const uint32_t d2 = tbase + kColD2 + 32 * (((c - kTcBufs) >> gsh) & 1);
                        tmem_ld16(d2, part);
                        tmem_ld16(d2 + 16, px);
                        tmem_wait_ld();
                        // one running sum per coefficient: the W_lo term (staged times 2^11) joins its group's partial first
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] += fmaf(__uint_as_float(px[k]), 1.0f / 2048.0f, __uint_as_float(part[k]));
                    }
                    TC_STAMP(t, c, 4);
                    tmem_wait_st();
                    tc_fence_before();  // orders the D1 / D2 loads and the operand stores before the issuer's next MMAs
                    __syncwarp();
                    TC_STAMP(t, c, 5);
                    if (lane == 0) mbar_arrive(&bars->a2_full[t][b]);
                    TC_STAMP(t, c, 6);
                }
                if (set != kTcSets - 1) continue;
                // drain: the last group's partial (NCH is a multiple of the group size; the tensor pipe completes in order, so
                // layer 2 of chunk NCH - 1 done means every MMA of the filter is done)
                mbar_wait_spin(&bars->a2_free[t][kTcBufs - 1], (ubase + uses - 1) & 1);
                tc_fence_after();
                {
                    uint32_t part[16], px[16];
                    const uint32_t d2 = tbase + kColD2 + 32 * (((NCH - 1) >> gsh) & 1);
                    tmem_ld16(d2, part);
                    tmem_ld16(d2 + 16, px);
                    tmem_wait_ld();
                    // undo the scalings (all exact powers of two): 2^-11 of the W_lo term, 2^-e of the point, 2^-q_k of the column
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        acc[k] = (acc[k] + fmaf(__uint_as_float(px[k]), 1.0f / 2048.0f, __uint_as_float(part[k]))) * inv_sc *
                                 cfg.tc_s2inv[f * kTcN2 + k];
                }
                tc_fence_before();
                // ---- hand the coefficients (+ b2 in fp32, Keras Dense) to the back-end warps ----
                const int slot = (int)(vseq & 1);
                mbar_wait_spin(&bars->c_free[t][slot], ((vseq >> 1) & 1) ^ 1);
                float* cb = s_c + ((size_t)(t * 2 + slot) * kTcN2) * kTcTile + pidx;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (!COEFF || k < Kr) cb[k * kTcTile] = okx ? acc[k] + (parts_mode ? 0.f : cfg.b2[f * Kr + k]) : CUDART_NAN_F;
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->c_full[t][slot]);
            }
        }
    } else if (warp < kTcActWarps + kTcTiles) {
        // =====================================================================================================
        // MMA issuer of tile t.  Ring cursors advance incrementally; descriptors differ only in their low word.
        // =====================================================================================================
        const int t = warp - kTcActWarps;
        constexpr uint32_t id1 = tc_idesc_f16(kTcChunk), id2 = tc_idesc_f16(2 * kTcN2), id2l = tc_idesc_f16(kTcN2);
        const uint32_t wbase = smem_u32(wring);
        const uint64_t dB1 = tc_smem_desc(wbase, kTcChunk * 16, 128);                              // [k half][chunk rows][8 halfs]
        const uint64_t dB2 = tc_smem_desc(wbase + kTcK1Max * kTcB1Halfs * 2, 2 * kTcN2 * 16, 128);   // [k half][32 rows][8 halfs]
        const uint32_t hi1 = (uint32_t)(dB1 >> 32), hi2 = (uint32_t)(dB2 >> 32);
        constexpr uint32_t kSlotStep = kTcChunkBytes >> 4;
        const int nk1 = 3 * (cfg.d + 1) <= 16 ? 1 : 2;   // layer-1 k-steps in use
        uint32_t s1 = 0, p1 = 0;                    // ring slot / phase parity of the next chunk layer 1 consumes
        uint32_t lo1 = (uint32_t)dB1;               // low descriptor word of that slot's B1 tile
        uint32_t s2 = 0, lo2 = (uint32_t)dB2;       // ... layer 2 (its chunk was waited for by layer 1 two chunks earlier)
        const uint32_t tb = tmem + 256 * t;
        auto adv1 = [&]() { lo1 += kSlotStep; if (++s1 == kTcStages) { s1 = 0; p1 ^= 1; lo1 = (uint32_t)dB1; } };
        auto adv2 = [&]() { lo2 += kSlotStep; if (++s2 == kTcStages) { s2 = 0; lo2 = (uint32_t)dB2; } };
        auto l1 = [&](int b) {  // D1[b] = A1 . B1(slot s1)   (elected lane only)
            mma_f16_ts(tb + kColD1 + kTcChunk * b, tb + kColA1, lo1, hi1, id1, 0u);
            if (nk1 > 1) mma_f16_ts(tb + kColD1 + kTcChunk * b, tb + kColA1 + 8, lo1 + ((kTcB1Halfs * 2) >> 4), hi1, id1, 1u);
            tc_commit(&bars->d1_full[t][b]);
        };
        uint32_t vseq = 0;
        long long total_v = my_super * F;
        if constexpr (SPLIT) {
            total_v = 0;
            for (long long it = 0; it < my_super; ++it)
                total_v += TC_PART_F1(blockIdx.x + it * gridDim.x) - TC_PART_F0(blockIdx.x + it * gridDim.x);
        }
        for (long long vv = 0; vv < total_v; ++vv, ++vseq) {
            mbar_wait_spin(&bars->a1_full[t], vseq & 1);
#pragma unroll
            for (int b = 0; b < kTcBufs; ++b) {  // prologue: layer 1 of the first kTcBufs chunks
                mbar_wait_spin(&bars->w_full[s1], p1);
                tc_fence_after();
                if (elect_one()) l1(b);
                __syncwarp();
                adv1();
            }
            const uint32_t ubase = vseq * uses;
            for (int c = 0; c < NCH; c += kTcBufs) {
#pragma unroll
                for (int b = 0; b < kTcBufs; ++b) {
                    const bool more = c + b + kTcBufs < NCH;
                    if (more) mbar_wait_spin(&bars->w_full[s1], p1);
                    mbar_wait_spin(&bars->a2_full[t][b], (ubase + (uint32_t)c / kTcBufs) & 1);
                    tc_fence_after();
                    TC_STAMP_I(t, c + b, 7);
                    if (elect_one()) {
                        const int cc = c + b;
                        const uint32_t d2 = tb + kColD2 + 32 * ((cc >> gsh) & 1);
                        const uint32_t a2 = tb + kColD1 + kTcChunk * b;
                        const uint32_t gfirst = (cc & (gsz - 1)) == 0 ? 0u : 1u;
#pragma unroll
                        for (int s = 0; s < kTcKSteps; ++s) {   // k-step = 16 hidden units = 8 columns of fp16 pairs, a 1 KB B tile
                            mma_f16_ts(d2, a2 + tc_a2_col(s), lo2 + s * ((kTcB2Halfs * 2) >> 4), hi2, id2, s > 0 ? 1u : gfirst);   // h_hi [W_hi | W_lo']
#ifndef TCV_NO_L2X
                            mma_f16_ts(d2, a2 + tc_a2_col(s) + kTcBlk / 2, lo2 + s * ((kTcB2Halfs * 2) >> 4), hi2, id2l, 1u);      // h_lo W_hi
#endif
                        }
                        tc_commit(&bars->a2_free[t][b]);
                        tc_commit(&bars->w_free[s2]);
                        if (more) l1(b);  // the activation warps read D1[b] before they signalled a2_full[b]
                    }
                    __syncwarp();
                    TC_STAMP_I(t, c + b, 8);
                    adv2();
                    if (more) adv1();
                }
            }
        }
    } else if (warp == kTcProdWarp) {
        // =====================================================================================================
        // TMA producer
        // =====================================================================================================
        uint32_t st = 0, ph = 0, vseq = 0;
        long long total_v = my_super * F;
        if constexpr (SPLIT) {
            total_v = 0;
            for (long long it = 0; it < my_super; ++it)
                total_v += TC_PART_F1(blockIdx.x + it * gridDim.x) - TC_PART_F0(blockIdx.x + it * gridDim.x);
        }
        long long item = blockIdx.x;
        int f = TC_PART_F0(item), f1 = TC_PART_F1(item);
        for (long long vv = 0; vv < total_v; ++vv, ++vseq) {
            const int slot = (int)(vseq & 1);
            if constexpr (!COEFF) {
                TC_WAIT_PROD(&bars->b_free[slot], ((vseq >> 1) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bars->b_full[slot], bbytes);
                    bulk_g2s(s_basis0 + slot * bslot, basis_src<FAST>(cfg, f, K), bbytes, &bars->b_full[slot]);
                }
                __syncwarp();
            }
            const float* src = cfg.tcpack + ((size_t)f * NCHT + (size_t)TC_PART_HS(item) * NCH) * kTcChunkFloats;
            for (int c = 0; c < NCH; ++c) {
                TC_WAIT_PROD(&bars->w_free[st], ph ^ 1);
                if (elect_one()) {
#ifdef TCV_NO_WLOAD
                    mbar_arrive(&bars->w_full[st]);
#else
                    mbar_arrive_expect_tx(&bars->w_full[st], kTcChunkBytes);
                    bulk_g2s(wring + (size_t)st * kTcChunkFloats, src + (size_t)c * kTcChunkFloats, kTcChunkBytes,
                             &bars->w_full[st]);
#endif
                }
                __syncwarp();
                if (++st == kTcStages) { st = 0; ph ^= 1; }
            }
            if (++f == f1) {   // next work item of this CTA
                item += gridDim.x;
                f = TC_PART_F0(item); f1 = TC_PART_F1(item);
            }
        }
    } else if (warp >= kTcBackWarp0) {
        // =====================================================================================================
        // back-end warps: likelihood of filter f (kernels.cuh: fused_filter_logl) while the tensor cores work on filter f + 1
        // =====================================================================================================
        const int t = (warp - kTcBackWarp0) >> 2;
        const int pidx = ((warp - kTcBackWarp0) & 3) * 32 + lane;
        // shared addresses of the back end's tables, opaque to the optimiser (kernels.cuh: sa_opaque)
        const uint32_t basis_sa0 = sa_opaque(s_basis0), basis_sa1 = sa_opaque(s_basis0 + bslot);
        const uint32_t obs_sa = sa_opaque(s_obs), samp_sa = sa_opaque(s_samp);
        uint32_t vseq = 0;
        for (long long it = 0; it < my_super; ++it) {
            const long long item = blockIdx.x + it * gridDim.x;
            const long long n = (item / nparts) * SUPER + (long long)t * kTcTile + pidx;
            const bool live = n < N;
            const double* row = pts + (live ? n : 0) * cfg.P;
            if constexpr (COEFF) {
                for (int f = TC_PART_F0(item), f1 = TC_PART_F1(item); f < f1; ++f, ++vseq) {
                    const int slot = (int)(vseq & 1);
                    TC_WAIT_BACK(&bars->c_full[t][slot], (vseq >> 1) & 1);
                    const float* cb = s_c + ((size_t)(t * 2 + slot) * kTcN2) * kTcTile + pidx;
                    if (live && parts_mode) {
                        float* dst = reinterpret_cast<float*>(out) + (((size_t)n * F + f) * hsplit + TC_PART_HS(item)) * Kr;
                        for (int k = 0; k < Kr; ++k) dst[k] = cb[k * kTcTile];
                    } else if (live) {
                        double* dst = out + ((size_t)n * F + f) * Kr;
                        for (int k = 0; k < Kr; ++k) dst[k] = (double)cb[k * kTcTile];
                    }
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&bars->c_free[t][slot]);
                        atomicAdd(&bars->progress, 1u);
                    }
                }
                continue;
            }
            const PointScal ps = point_setup(cfg, row);
            bool ok = !ps.bad && !cfg.static_fail;
            double logl = 0.0;
            for (int f = TC_PART_F0(item), f1 = TC_PART_F1(item); f < f1; ++f, ++vseq) {
                const int slot = (int)(vseq & 1);
                const uint32_t par = (vseq >> 1) & 1;
                TC_WAIT_BACK(&bars->c_full[t][slot], par);
                const float* cb = s_c + ((size_t)(t * 2 + slot) * kTcN2) * kTcTile + pidx;
                float cf[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    cf[k] = cb[k * kTcTile];
                    ok = ok && isfinite(cf[k]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->c_free[t][slot]);
                TC_WAIT_BACK(&bars->b_full[slot], par);
                if (ok) {
                    const double* basis = reinterpret_cast<const double*>(s_basis0 + slot * bslot);
#ifdef TCV_NO_BACKEND
                    if (true) {
                        logl += cf[0];
                    } else
#endif
                    if constexpr (FAST) {
#ifndef TCV_NO_SA
                        const uint32_t sa[3] = {slot ? basis_sa1 : basis_sa0, obs_sa, samp_sa};
                        logl += fused_filter_logl<K, true, float, true>(cfg, f, cf, ps, row, basis, s_obs, s_samp, sa);
#else
                        logl += fused_filter_logl<K, true>(cfg, f, cf, ps, row, basis, s_obs, s_samp);
#endif
                    } else {
                        double cp[K];
#pragma unroll
                        for (int k = 0; k < K; ++k) cp[k] = (double)cf[k];
                        logl += fused_filter_logl<K, false>(cfg, f, cp, ps, row, basis, s_obs, s_samp);
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bars->b_free[slot]);
                    atomicAdd(&bars->progress, 1u);
                }
            }
            if (live) {
                const bool good = ok && isfinite(logl);
                if (fsplit == 1) out[n] = good ? logl : NMMA_SENTINEL;
                else out[n * fsplit + item % fsplit] = good ? logl : CUDART_NAN;
            }
        }
    }
    if (warp != kTcOwnerWarp) {
        __syncwarp();
        if (lane == 0) atomicAdd(&bars->finished, 1u);
    } else if (lane == 0) {
        // watchdog (the TMEM-owner warp has nothing else to do): the hand-off waits above spin without a bound, so a
        // phase bug would hang the sampler process; no progress of any back-end warp for kMbarTimeoutNs -> trap
        volatile uint32_t* fin = &bars->finished;
        volatile uint32_t* prog = &bars->progress;
        uint32_t last = *prog;
        unsigned long long t_last = global_timer_ns();
        while (*fin < (uint32_t)(kTcThreads / 32 - 1)) {
            __nanosleep(500);   // short: the CTA cannot retire before this warp has seen `finished` (a 20 us nap cost a
                                // one-point call 26 us, tools/latency.py)
            const uint32_t p = *prog;
            const unsigned long long now = global_timer_ns();
            if (p != last) { last = p; t_last = now; }
            else if (now - t_last > kMbarTimeoutNs) __trap();
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef TCV_TIMELINE
    if (blockIdx.x == 0 && tid == 0) {
        const long long t0 = g_tc_tl[0][0][0];
        for (int c = 0; c < NCH && c < 64; ++c)
            for (int t = 0; t < kTcTiles; ++t) {
                printf("tl tile %d chunk %2d:", t, c);
                for (int k = 0; k < 9; ++k) printf(" %6lld", g_tc_tl[t][c][k] - t0);
                printf("\n");
            }
    }
#endif
    if (warp == kTcOwnerWarp) tmem_dealloc(tmem, 512);
#undef TC_PART_F0
#undef TC_PART_F1
#undef TC_PART_HS
}

}  // namespace nmma
