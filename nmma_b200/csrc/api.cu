// C-ABI implementation (include/nmma_b200.h): handle, one-time staging of the
// surrogate / observation tables into device buffers, and kernel dispatch.
#define NMMA_TWO_STAGE_TU 1  // this translation unit owns the non-template two-stage kernels
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "handle.h"
#include "gp_kernel.cuh"  // shared-memory footprint of the fused GP kernel
#include "tc_kernel.cuh"  // layout constants of the tensor-core kernel (its weight pack is built here)

using namespace nmma;

namespace {
thread_local std::string g_create_error;
constexpr int kVersion = 100;  // 0.1.0
constexpr long long kTwoStageChunk = 1 << 18;  // points per coefficient-scratch chunk
}  // namespace

int nmma::fail(nmma_b200_t* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

namespace {

bool all_finite(const double* p, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(p[i])) return false;
    return true;
}

void free_dev(nmma_b200_t* h) {
    for (void* p : h->dev_allocs) cudaFree(p);
    h->dev_allocs.clear();
}

template <typename T>
int upload(nmma_b200_t* h, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CU(cudaMalloc(&p, bytes + 256));  // slack: 16-byte bulk copies may be rounded up
    h->dev_allocs.push_back(p);
    if (!v.empty()) CU(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T*>(p);
    return NMMA_B200_OK;
}

int check_src(nmma_b200_t* h, const ParamSrc& s, const char* what) {
    if (s.col >= h->P) return fail(h, NMMA_B200_ERR_ARG, "%s: column %d out of range (P=%d)", what, s.col, h->P);
    if (s.xf < 0 || s.xf > 5) return fail(h, NMMA_B200_ERR_ARG, "%s: unknown transform %d", what, s.xf);
    return NMMA_B200_OK;
}

ParamSrc to_src(const nmma_b200_param_src& s) { return ParamSrc{s.col, s.transform, s.value}; }

// hi = the 19 bits kind::tf32 reads, lo = remainder (exact in fp32; the tensor core truncates it again)
inline void tf32_split(float w, float* hi, float* lo) {
    uint32_t u;
    std::memcpy(&u, &w, 4);
    u &= 0xFFFFE000u;
    std::memcpy(hi, &u, 4);
    *lo = std::isfinite(w) ? (w - *hi) : 0.f;
}

int ensure_scratch(nmma_b200_t* h, size_t n_doubles) {
    if (n_doubles <= h->coeff_cap) return NMMA_B200_OK;
    for (auto& g : h->lat_graphs) cudaGraphExecDestroy(g.exec);   // captured graphs hold the old scratch pointer
    h->lat_graphs.clear();
    h->lat_warm.clear();
    if (h->coeff_scratch) cudaFree(h->coeff_scratch);
    h->coeff_scratch = nullptr;
    h->coeff_cap = 0;
    CU(cudaMalloc((void**)&h->coeff_scratch, n_doubles * sizeof(double)));
    h->coeff_cap = n_doubles;
    return NMMA_B200_OK;
}

int launch_frontend(nmma_b200_t* h, const double* pts, long long N, double* coeff, cudaStream_t st) {
    if (h->kind == 0 && h->tc_front_supported && N >= h->opt_tc_front_min && h->opt_path != 2) {
        // "path" = 2 forces the plain two-stage kernels (parity tests compare the two front ends)
        return launch_tc_coeff(h, pts, N, coeff, st);
    } else if (h->kind == 0) {
        dim3 grid((unsigned)h->F, (unsigned)std::min<long long>(N, 32768));
        coeff_mlp_kernel<<<grid, kCoeffThreads, 0, st>>>(h->cfg, pts, N, coeff);
    } else {
        const long long ntiles = (N + kGpPts - 1) / kGpPts;
        // gf_pow needs 256 a log2(base) < 2^31: a <= 1e5 (finalize), and its tables 64 KB next to the r^2 tile
        // Large batches only: the tables cap the residency at 3 CTAs per SM, which small grids of this lanes-over-rows
        // kernel (4 chains per warp) pay for in latency hiding (profiles/r02_gp_threshold.txt: break-even at ~4096 points)
        const bool gf = h->gp_alpha_ok && ntiles >= 2048 &&
                        (size_t)kGpPts * h->Ntr * sizeof(double) + kGfLogBytes + kGfExpBytes <= 227 * 1024;
        const size_t smem = (size_t)kGpPts * h->Ntr * sizeof(double) + (gf ? kGfLogBytes + kGfExpBytes : 0);
        static size_t attr_smem[2] = {0, 0};
        if (attr_smem[gf] < smem) {
            if (gf) CU(cudaFuncSetAttribute(coeff_gp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else CU(cudaFuncSetAttribute(coeff_gp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem[gf] = smem;
        }
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ntiles, (long long)h->sm_count * (gf ? 3 : 8)));
        // few tiles: spread the F K (filter, coefficient) pairs of each over several CTAs, one pair per warp at most
        const long long pair_groups = ((long long)h->F * h->K + kGpThreads / 32 - 1) / (kGpThreads / 32);
        const unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>(pair_groups, (long long)h->sm_count * 4 / std::max<long long>(1, ntiles)));
        if (gf) coeff_gp_kernel<true><<<dim3(grid, gy), kGpThreads, smem, st>>>(h->cfg, pts, N, coeff);
        else coeff_gp_kernel<false><<<dim3(grid, gy), kGpThreads, smem, st>>>(h->cfg, pts, N, coeff);
    }
    CU(cudaGetLastError());
    h->launches += 1;
    return NMMA_B200_OK;
}

// Builds every derived table and uploads the configuration (lazy, after any set_*).
int finalize(nmma_b200_t* h, bool need_obs) {
    if (!h->have_svd) return fail(h, NMMA_B200_ERR_STATE, "nmma_b200_set_svd has not been called");
    if (h->kind < 0) return fail(h, NMMA_B200_ERR_STATE, "no surrogate front end: call nmma_b200_set_mlp or nmma_b200_set_gp");
    if (!h->have_layout) return fail(h, NMMA_B200_ERR_STATE, "nmma_b200_set_param_layout has not been called");
    if (need_obs && !h->have_obs) return fail(h, NMMA_B200_ERR_STATE, "nmma_b200_set_observations has not been called");
    if (!h->dirty) return NMMA_B200_OK;
    CU(cudaSetDevice(h->device));
    free_dev(h);
    const int F = h->F, d = h->d, K = h->K, T = h->T;
    DevCfg& c = h->cfg;
    c = DevCfg{};
    c.F = F; c.d = d; c.K = K; c.T = T; c.P = h->P; c.kind = h->kind;

    if ((int)h->xsrc.size() != d) return fail(h, NMMA_B200_ERR_ARG, "param layout lists %zu model parameters, surrogate has d=%d", h->xsrc.size(), d);
    for (int i = 0; i < d; ++i) {
        if (int rc = check_src(h, h->xsrc[i], "model parameter")) return rc;
        c.xsrc[i] = h->xsrc[i];
    }
    if (int rc = check_src(h, h->dl, "luminosity_distance")) return rc;
    if (int rc = check_src(h, h->ts, "timeshift")) return rc;
    if (int rc = check_src(h, h->zsrc, "redshift")) return rc;
    c.dl = h->dl; c.ts = h->ts; c.zsrc = h->zsrc; c.zmode = h->zmode;
    if (h->zmode == NMMA_B200_Z_TABLE && h->zd.size() < 1)
        return fail(h, NMMA_B200_ERR_STATE, "z_mode is Z_TABLE but nmma_b200_set_redshift_table has not been called");
    c.nz = (int)h->zd.size();
    if (int rc = upload(h, h->zd, &c.zd)) return rc;
    if (int rc = upload(h, h->zz, &c.zz)) return rc;
    c.ncon = (int)h->con_src.size();
    for (int i = 0; i < c.ncon; ++i) {
        if (int rc = check_src(h, h->con_src[i], "constraint")) return rc;
        c.con_src[i] = h->con_src[i]; c.con_lo[i] = h->con_lo[i]; c.con_hi[i] = h->con_hi[i];
    }
    c.ext_law = h->ext_law;
    c.ebv = h->ebv;
    if (h->ext_law != 0) {
        if (int rc = check_src(h, h->ebv, "Ebv")) return rc;
        if ((int)h->ext_nu.size() != F || (int)h->ext_coef.size() != F)
            return fail(h, NMMA_B200_ERR_ARG, "extinction tables describe %zu filters, surrogate has F=%d", h->ext_nu.size(), F);
        if (int rc = upload(h, h->ext_nu, &c.ext_nu)) return rc;
        if (int rc = upload(h, h->ext_coef, &c.ext_coef)) return rc;
    }

    // ---- basis pack + input scaling ----
    std::vector<double> pden((size_t)F * d), bpack((size_t)F * T * (K + 2));
    for (size_t i = 0; i < pden.size(); ++i) pden[i] = h->pmax[i] - h->pmin[i];
    for (int f = 0; f < F; ++f)
        for (int j = 0; j < T; ++j) {
            double* r = &bpack[((size_t)f * T + j) * (K + 2)];
            for (int i = 0; i < K; ++i) r[i] = h->VA[((size_t)f * T + j) * K + i];
            r[K] = h->maxs[(size_t)f * T + j] - h->mins[(size_t)f * T + j];
            r[K + 1] = h->mins[(size_t)f * T + j];
        }
    if (int rc = upload(h, h->pmin, &c.pmin)) return rc;
    if (int rc = upload(h, pden, &c.pden)) return rc;
    if (int rc = upload(h, bpack, &c.bpack)) return rc;
    {   // fp32 row-pair pack of the FAST back end (backend.cuh: DevCfg::bpack32)
        std::vector<float2> b32((size_t)F * (K + 2) * T);
        for (int f = 0; f < F; ++f)
            for (int i = 0; i < K + 2; ++i)
                for (int j = 0; j < T; ++j) {
                    const int j1 = std::min(j + 1, T - 1);
                    b32[((size_t)f * (K + 2) + i) * T + j] =
                        make_float2((float)bpack[((size_t)f * T + j) * (K + 2) + i], (float)bpack[((size_t)f * T + j1) * (K + 2) + i]);
                }
        if (int rc = upload(h, b32, &c.bpack32)) return rc;
    }

    // ---- stage 1 tables: np.interp(sample_times, tt_f, ., left=inf, right=inf) ----
    std::vector<double> samp = h->samp;
    if (samp.empty()) samp.assign(h->tt.begin(), h->tt.begin() + T);
    const int S = (int)samp.size();
    c.S = S;
    for (int s = 0; s < S; ++s) {
        if (!std::isfinite(samp[s])) return fail(h, NMMA_B200_ERR_ARG, "sample_times[%d] is not finite", s);
        if (s > 0 && !(samp[s] >= samp[s - 1])) return fail(h, NMMA_B200_ERR_ARG, "sample_times must be non-decreasing");
    }
    std::vector<int> s_lo(F), s_hi(F), s1_j((size_t)F * S, 0);
    std::vector<double> s1_dx((size_t)F * S, 0.0), s1_dt((size_t)F * S, 0.0);
    bool single = true;
    int static_fail = 0;
    for (int f = 0; f < F; ++f) {
        const double* tf = &h->tt[(size_t)f * T];
        int lo = S, hi = -1;
        for (int s = 0; s < S; ++s) {
            const double x = samp[s];
            if (x < tf[0] || x > tf[T - 1]) continue;
            lo = std::min(lo, s); hi = std::max(hi, s);
            int j = (int)(std::upper_bound(tf, tf + T, x) - tf) - 1;  // last tt[j] <= x
            const size_t idx = (size_t)f * S + s;
            s1_j[idx] = j;
            if (j == T - 1 || tf[j] == x) {
                s1_dt[idx] = 0.0;
            } else {
                s1_dx[idx] = x - tf[j];
                s1_dt[idx] = tf[j + 1] - tf[j];
            }
            if (!(s1_dt[idx] == 0.0 && j == s)) single = false;
        }
        if (hi < lo) { lo = 0; hi = -1; }
        if (!(lo == 0 && hi == S - 1)) single = false;
        if (hi - lo + 1 < 2) static_fail = 1;
        s_lo[f] = lo; s_hi[f] = hi;
    }
    c.single_stage = single ? 1 : 0;
    c.static_fail = static_fail;
    // uniform sample grid -> O(1) interval guess
    c.uniform = 0;
    if (S >= 2) {
        const double ds = (samp[S - 1] - samp[0]) / (S - 1);
        bool uni = ds > 0;
        for (int s = 0; s < S && uni; ++s)
            if (std::fabs(samp[s] - (samp[0] + s * ds)) > 1e-6 * ds) uni = false;
        if (uni) { c.uniform = 1; c.uni_s0 = samp[0]; c.uni_inv_ds = 1.0 / ds; }
        // fp32 index guess: |error| <= ~4 ulp(S) + the 1e-6 non-uniformity admitted above; 16x margin, >= 2^-10
        c.fast_delta = std::max(1.0f / 1024.0f, (float)S * 8e-6f);
        if (c.fast_delta >= 0.25f) c.uniform = 0;   // grid too long for an fp32 guess: exact path only
    }
    if (int rc = upload(h, samp, &c.samp)) return rc;
    if (int rc = upload(h, s_lo, &c.s_lo)) return rc;
    if (int rc = upload(h, s_hi, &c.s_hi)) return rc;
    if (int rc = upload(h, s1_j, &c.s1_j)) return rc;
    if (int rc = upload(h, s1_dx, &c.s1_dx)) return rc;
    if (int rc = upload(h, s1_dt, &c.s1_dt)) return rc;

    // ---- front end ----
    if (h->kind == 0) {
        if (h->Kout != K)
            return fail(h, NMMA_B200_ERR_ARG, "network emits %d coefficients but n_coeff=%d (np.dot(VA[:, :n], cAproj) would not align)", h->Kout, K);
        const int H = h->H;
        const int RW = (d + 1 + K + 3) / 4 * 4;
        const int HP = (H + kHC - 1) / kHC * kHC;
        c.H = H; c.HP = HP; c.RW = RW;
        std::vector<float> wpack((size_t)F * HP * RW, 0.f);
        for (int f = 0; f < F; ++f)
            for (int j = 0; j < H; ++j) {
                float* r = &wpack[((size_t)f * HP + j) * RW];
                for (int i = 0; i < d; ++i) r[i] = h->W1[((size_t)f * d + i) * H + j];
                r[d] = h->b1[(size_t)f * H + j];
                for (int k = 0; k < K; ++k) r[d + 1 + k] = h->W2[((size_t)f * H + j) * K + k];
            }
        if (int rc = upload(h, wpack, &c.wpack)) return rc;
        if (int rc = upload(h, h->b2, &c.b2)) return rc;
        // tensor-core operand tiles (tc_kernel.cuh): K-major, no swizzle, hi/lo split
        c.tc_nch = 0;
        c.tcpack = nullptr;
        // fp16 operands, brought into range by exact power-of-two scalings (tc_kernel.cuh header): row i of [W1; b1]
        // times 2^-r_i (largest entry in [2^9, 2^10)), column k of W2 times 2^q_k (largest entry in [8, 16)), the
        // remainder of W2 times 2^11; every weight = hi + lo with two fp16 values (lo may be subnormal: its absolute
        // error is then 2^-25, i.e. < 2^-34 of the row / column maximum)
        bool w_ok = true;
        for (float w : h->W1) w_ok = w_ok && std::isfinite(w);
        for (float w : h->b1) w_ok = w_ok && std::isfinite(w);
        for (float w : h->W2) w_ok = w_ok && std::isfinite(w);
        // ... and the per-point scale is an exponent-field computation clamped to 2^-100 .. 2^100: rows / columns whose
        // largest entry is outside 2^-60 .. 2^60 keep the tensor-core path off as well (the FFMA kernel takes over)
        auto sane = [](float vmax) { return !(vmax > 0.f) || (vmax > 8.7e-19f && vmax < 1.1e18f); };
        for (int f = 0; f < F && w_ok; ++f) {
            for (int i = 0; i <= d && w_ok; ++i) {
                float vmax = 0.f;
                for (int j = 0; j < H; ++j)
                    vmax = std::max(vmax, std::fabs(i < d ? h->W1[((size_t)f * d + i) * H + j] : h->b1[(size_t)f * H + j]));
                w_ok = sane(vmax);
            }
            for (int o = 0; o < K && w_ok; ++o) {
                float vmax = 0.f;
                for (int j = 0; j < H; ++j) vmax = std::max(vmax, std::fabs(h->W2[((size_t)f * H + j) * K + o]));
                w_ok = sane(vmax);
            }
        }
        if (d + 1 <= 8 && K <= kTcN2 && w_ok) {
            int nch = (H + kTcChunk - 1) / kTcChunk;
            nch = (nch + kTcUnit - 1) / kTcUnit * kTcUnit;  // whole layer-2 accumulation groups (a multiple of the TMEM buffers per tile)
            c.tc_nch = nch;
            std::vector<float> tp((size_t)F * nch * kTcChunkFloats, 0.f), xs((size_t)F * 8, 0.f), s2inv((size_t)F * kTcN2, 1.f);
            auto pow2_scale = [](float vmax, int top) {   // 2^p with vmax 2^p in [2^(top-1), 2^top); 1 for an all-zero row
                if (!(vmax > 0.f)) return 1.0f;
                int e; std::frexp(vmax, &e);              // vmax = m 2^e, m in [0.5, 1)
                return std::ldexp(1.0f, top - e);
            };
            for (int f = 0; f < F; ++f) {
                float rs[8];   // 2^-r_i
                for (int i = 0; i <= d; ++i) {
                    float vmax = 0.f;
                    for (int j = 0; j < H; ++j)
                        vmax = std::max(vmax, std::fabs(i < d ? h->W1[((size_t)f * d + i) * H + j] : h->b1[(size_t)f * H + j]));
                    rs[i] = pow2_scale(vmax, 10);
                    xs[(size_t)f * 8 + i] = 1.0f / rs[i];
                }
                float cs[kTcN2];   // 2^q_k
                for (int o = 0; o < K; ++o) {
                    float vmax = 0.f;
                    for (int j = 0; j < H; ++j) vmax = std::max(vmax, std::fabs(h->W2[((size_t)f * H + j) * K + o]));
                    cs[o] = pow2_scale(vmax, 4);
                    s2inv[(size_t)f * kTcN2 + o] = 1.0f / cs[o];
                }
                for (int j = 0; j < H; ++j) {
                    __half* ch = reinterpret_cast<__half*>(&tp[((size_t)f * nch + j / kTcChunk) * kTcChunkFloats]);
                    const int n = j % kTcChunk;
                    for (int i = 0; i <= d; ++i) {   // K slot 3 i + {0, 1, 2} pairs with the A slots {x_hi, x_lo, x_hi}
                        const float w = rs[i] * (i < d ? h->W1[((size_t)f * d + i) * H + j] : h->b1[(size_t)f * H + j]);
                        const __half whi = __float2half_rn(w), wlo = __float2half_rn(w - __half2float(whi));
                        const __half slot[3] = {whi, whi, wlo};
                        for (int q = 0; q < 3; ++q) {
                            const int k = 3 * i + q;
                            ch[(k / 16) * kTcB1Halfs + tc_b_index16(kTcChunk, n, k % 16)] = slot[q];
                        }
                    }
                    __half* b2t = ch + kTcK1Max * kTcB1Halfs + (n / 16) * kTcB2Halfs;   // [W_hi rows 0..15 | W_lo' rows 16..31] x 16 hidden
                    for (int o = 0; o < K; ++o) {
                        const float w = cs[o] * h->W2[((size_t)f * H + j) * K + o];
                        const __half whi = __float2half_rn(w);
                        b2t[tc_b_index16(2 * kTcN2, o, n % 16)] = whi;
                        b2t[tc_b_index16(2 * kTcN2, kTcN2 + o, n % 16)] = __float2half_rn((w - __half2float(whi)) * 2048.0f);
                    }
                }
            }
            if (int rc = upload(h, xs, &c.tc_xs)) return rc;
            if (int rc = upload(h, s2inv, &c.tc_s2inv)) return rc;
            if (int rc = upload(h, tp, &c.tcpack)) return rc;
        }
    } else {
        c.Ntr = h->Ntr;
        // GP inputs are scaled with filter 0's param_mins/maxs: training shares them (em/training.py:216-230)
        for (int f = 1; f < F; ++f)
            for (int i = 0; i < d; ++i)
                if (h->pmin[f * d + i] != h->pmin[i] || h->pmax[f * d + i] != h->pmax[i])
                    return fail(h, NMMA_B200_ERR_UNSUPPORTED, "GP path needs identical param_mins/maxs across filters");
        const size_t n = (size_t)F * K;
        h->gp_alpha_ok = true;   // gf_pow takes rint(256 a log2(base)) from the low word of a double: a <= 1e5, the sklearn bound
        for (double a : h->gpRa) h->gp_alpha_ok = h->gp_alpha_ok && a > 0.0 && a <= 1e5;
        std::vector<double> A(n * h->Ntr), q(n);
        for (size_t p = 0; p < n; ++p) {
            q[p] = 1.0 / (2.0 * h->gpRa[p] * h->gpRl[p] * h->gpRl[p]);
            for (int t = 0; t < h->Ntr; ++t) A[p * h->Ntr + t] = h->gpC2[p] * h->gpAlpha[p * h->Ntr + t];
        }
        if (int rc = upload(h, h->gpX, &c.gpX)) return rc;
        if (int rc = upload(h, A, &c.gpA)) return rc;
        std::vector<double> AT(n * h->Ntr);
        for (int f = 0; f < F; ++f)
            for (int k = 0; k < K; ++k)
                for (int t = 0; t < h->Ntr; ++t) AT[((size_t)f * h->Ntr + t) * K + k] = A[((size_t)f * K + k) * h->Ntr + t];
        if (int rc = upload(h, AT, &c.gpAT)) return rc;
        if (int rc = upload(h, q, &c.gp_q)) return rc;
        if (int rc = upload(h, h->gpRa, &c.gp_ra)) return rc;
        if (int rc = upload(h, h->gpYm, &c.gp_ym)) return rc;
        if (int rc = upload(h, h->gpYs, &c.gp_ys)) return rc;
    }

    // ---- observations + systematics ----
    h->fused_supported = false;
    h->tc_supported = false;
    h->gp_fused_supported = false;
    h->fast_backend_ok = false;
    // coefficient mode of the tensor-core kernel: no observations needed (generate_lightcurve / coeffs use it too)
    if (!h->have_obs) { c.G = 0; c.nobs = 0; }
    h->tc_front_supported = (h->kind == 0) && K <= kTcN2 && c.tc_nch > 0 &&
                            tc_smem_bytes(kTcN2, T, c.S, h->have_obs ? h->g_off[h->G] : 0) <= 227 * 1024;
    if (h->have_obs) {
        const int G = h->G;
        const int nobs = h->g_off[G];
        c.G = G; c.nobs = nobs;
        for (int g = 0; g < G; ++g)
            for (int k = 0; k < h->g_nh[g]; ++k)
                if (h->g_h[g * 3 + k] < 0 || h->g_h[g * 3 + k] >= F)
                    return fail(h, NMMA_B200_ERR_ARG, "observed filter %d maps to model filter %d, but F=%d", g, h->g_h[g * 3 + k], F);
        std::vector<int> sy_mode = h->sy_mode, sy_nn = h->sy_nn, sy_off = h->sy_off;
        std::vector<double> sy_budget = h->sy_budget, sy_t = h->sy_t;
        std::vector<ParamSrc> sy_src = h->sy_src;
        if (!h->have_sys) {  // FilterSystematicsHandler default: error_budget = 1.0 (systematics.py:203-210)
            sy_mode.assign(G, 0); sy_nn.assign(G, 0); sy_off.assign(G, 0); sy_budget.assign(G, 1.0);
            sy_t.clear(); sy_src.clear();
        } else if ((int)sy_mode.size() != G) {
            return fail(h, NMMA_B200_ERR_ARG, "systematics describe %zu filters, observations %d", sy_mode.size(), G);
        }
        for (auto& s : sy_src)
            if (int rc = check_src(h, s, "systematics parameter")) return rc;
        std::vector<int> o_g(nobs);
        std::vector<double> o_sig(nobs), o_lsc(nobs), o_pack((size_t)nobs * kObsRec);
        for (int g = 0; g < G; ++g)
            for (int k = h->g_off[g]; k < h->g_off[g + 1]; ++k) {
                o_g[k] = g;
                const double sg = std::sqrt(h->o_s[k] * h->o_s[k] + sy_budget[g] * sy_budget[g]);
                o_sig[k] = sg;
                o_lsc[k] = std::log(sg) + NMMA_NORM_PDF_LOGC;
                double* rec = &o_pack[(size_t)k * kObsRec];
                rec[0] = h->o_t[k]; rec[1] = h->o_m[k]; rec[2] = h->o_s[k];
                rec[3] = sg; rec[4] = 1.0 / sg; rec[5] = o_lsc[k];
                float* rf = reinterpret_cast<float*>(rec + 6);
                rf[0] = (float)h->o_t[k]; rf[1] = (float)h->o_m[k]; rf[2] = (float)(1.0 / sg); rf[3] = (float)o_lsc[k];
                // observation class of the FAST back end (kernels.cuh: kObsGeneral / kObsSimple / kObsSampled)
                const double lim = h->g_lim[g], tk = h->o_t[k];
                const bool det = std::isfinite(h->o_s[k]) && std::isfinite(h->o_m[k]) && std::isfinite(tk);
                int cls = kObsGeneral, i0 = -1, i1 = -1;
                float wgt = 0.f;
                const bool upper = !std::isfinite(h->o_s[k]) && std::isfinite(h->o_m[k]) && std::isfinite(tk);
                if (det && sy_mode[g] == 0 && lim == INFINITY) {
                    cls = kObsSimple;
                } else if (upper || (det && h->o_m[k] <= lim && !std::isnan(lim) && lim > -INFINITY)) {
                    // detections: the support test m <= limit does not depend on the point; m > limit stays on the exact
                    // path (-inf).  Upper limits (norm.logsf) share the sigma_sys plumbing of the sampled class.
                    const int fast_cls = upper ? kObsUpper : kObsSampled;
                    if (sy_mode[g] == 0) {
                        cls = fast_cls;      // constant budget (behind a finite detection limit / for an upper limit)
                    } else if (sy_mode[g] == 1) {
                        cls = fast_cls; i0 = i1 = sy_off[g];
                    } else if (sy_mode[g] == 2 && sy_nn[g] >= 2) {
                        // np.interp(t, nodes, values) with 'constant' ends (em/utils.py:667-670): fixed bracket and weight
                        const double* tn = &sy_t[sy_off[g]];
                        const int nn = sy_nn[g];
                        bool sorted = true;
                        for (int i = 0; i + 1 < nn; ++i) sorted = sorted && tn[i + 1] > tn[i];
                        if (sorted) {
                            cls = fast_cls;
                            if (tk <= tn[0]) { i0 = i1 = sy_off[g]; }
                            else if (tk >= tn[nn - 1]) { i0 = i1 = sy_off[g] + nn - 1; }
                            else {
                                int j = (int)(std::upper_bound(tn, tn + nn, tk) - tn) - 1;
                                i0 = sy_off[g] + j; i1 = i0 + 1;
                                wgt = (float)((tk - tn[j]) / (tn[j + 1] - tn[j]));
                            }
                        }
                    }
                }
                int32_t* ri = reinterpret_cast<int32_t*>(rec + 8);
                ri[0] = cls; ri[1] = i0;
                ri[2] = i1; std::memcpy(&ri[3], &wgt, sizeof(float));
                float* rg = reinterpret_cast<float*>(rec + 10);
                rg[0] = (float)(h->o_s[k] * h->o_s[k]); rg[1] = (float)lim;
                rg[2] = (float)sy_budget[g]; rg[3] = 0.f;
            }
        std::vector<int> f_goff(F + 1, 0), f_glist;
        bool direct = true;
        for (int g = 0; g < G; ++g) direct = direct && (h->g_nh[g] == 1);
        for (int f = 0; f < F; ++f) {
            for (int g = 0; g < G; ++g)
                if (h->g_nh[g] == 1 && h->g_h[g * 3] == f) f_glist.push_back(g);
            f_goff[f + 1] = (int)f_glist.size();
        }
        if (int rc = upload(h, h->g_off, &c.g_off)) return rc;
        if (int rc = upload(h, h->g_nh, &c.g_nh)) return rc;
        if (int rc = upload(h, h->g_h, &c.g_h)) return rc;
        if (int rc = upload(h, h->g_lim, &c.g_lim)) return rc;
        if (int rc = upload(h, o_g, &c.o_g)) return rc;
        if (int rc = upload(h, h->o_t, &c.o_t)) return rc;
        if (int rc = upload(h, h->o_m, &c.o_m)) return rc;
        if (int rc = upload(h, h->o_s, &c.o_s)) return rc;
        if (int rc = upload(h, o_sig, &c.o_sig)) return rc;
        if (int rc = upload(h, o_lsc, &c.o_lsc)) return rc;
        if (int rc = upload(h, o_pack, &c.o_pack)) return rc;
        if (int rc = upload(h, sy_mode, &c.sy_mode)) return rc;
        if (int rc = upload(h, sy_budget, &c.sy_budget)) return rc;
        if (int rc = upload(h, sy_nn, &c.sy_nn)) return rc;
        if (int rc = upload(h, sy_off, &c.sy_off)) return rc;
        if (int rc = upload(h, sy_src, &c.sy_src)) return rc;
        if (int rc = upload(h, sy_t, &c.sy_t)) return rc;
        if (int rc = upload(h, f_goff, &c.f_goff)) return rc;
        if (int rc = upload(h, f_glist, &c.f_glist)) return rc;
        h->fast_backend_ok = direct && c.single_stage && c.uniform;   // what fused_filter_logl<FAST> / fast_obs_term assume
        h->fused_supported = (h->kind == 0) && direct && fused_has(d, K) &&
                             fused_smem_bytes(d, K, T, c.S, nobs) <= 227 * 1024;
        h->tc_supported = (h->kind == 0) && direct && K == 10 && c.tc_nch > 0 &&
                          tc_smem_bytes(K, T, c.S, nobs) <= 227 * 1024;
        // fused GP kernel: gf_pow takes rint(256 a log2(base)) from the low word of a double, i.e. needs it below 2^31:
        // a <= 1e5 (the sklearn bound of RationalQuadratic.alpha) leaves room for base < 2^80
        h->gp_fused_supported = (h->kind == 1) && direct && gp_fused_has(d, K) && h->gp_alpha_ok &&
                                gf_smem_bytes(h->Ntr, d) <= 227 * 1024;
    }
    h->dirty = false;
    h->cfg_epoch += 1;   // captured graphs hold the old DevCfg by value
    return NMMA_B200_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int nmma_b200_version(void) { return kVersion; }

const char* nmma_b200_last_error(const nmma_b200_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int nmma_b200_create(int device, nmma_b200_t** out) {
    nmma_b200_t* h = nullptr;
    if (!out) return fail(nullptr, NMMA_B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, NMMA_B200_ERR_CUDA, "no CUDA device available (%s); nmma_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(nullptr, NMMA_B200_ERR_ARG, "device %d out of range (have %d)", device, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, NMMA_B200_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, NMMA_B200_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    device, prop.major, prop.minor);
    h = new nmma_b200_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete h;
        return fail(nullptr, NMMA_B200_ERR_CUDA, "stream creation failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return NMMA_B200_OK;
}

int nmma_b200_destroy(nmma_b200_t* h) {
    if (!h) return NMMA_B200_OK;
    cudaSetDevice(h->device);
    free_dev(h);
    if (h->coeff_scratch) cudaFree(h->coeff_scratch);
    for (auto& g : h->lat_graphs) cudaGraphExecDestroy(g.exec);
    if (h->tc_parts) cudaFree(h->tc_parts);
    if (h->gp_parts) cudaFree(h->gp_parts);
    if (h->gp_tickets) cudaFree(h->gp_tickets);
    if (h->pr_dev) cudaFree(h->pr_dev);
    if (h->pr_tab_dev) cudaFree(h->pr_tab_dev);
    if (h->sweep_scratch) cudaFree(h->sweep_scratch);
    if (h->stage_in_dev) cudaFree(h->stage_in_dev);
    if (h->stage_out_dev) cudaFree(h->stage_out_dev);
    if (h->stage_in_host) cudaFreeHost(h->stage_in_host);
    if (h->stage_out_host) cudaFreeHost(h->stage_out_host);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_in_stream) cudaStreamDestroy(h->copy_in_stream);
    if (h->copy_out_stream) cudaStreamDestroy(h->copy_out_stream);
    for (cudaEvent_t ev : h->pipe_events) cudaEventDestroy(ev);
    delete h;
    return NMMA_B200_OK;
}

int nmma_b200_set_svd(nmma_b200_t* h, int F, int d, int K, int T, const double* tt, const double* param_mins,
                      const double* param_maxs, const double* VA, const double* mins, const double* maxs) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (F < 1 || d < 1 || K < 1 || T < 2) return fail(h, NMMA_B200_ERR_ARG, "set_svd: need F>=1, d>=1, K>=1, T>=2");
    if (d > kMaxD || K > kMaxK) return fail(h, NMMA_B200_ERR_UNSUPPORTED, "set_svd: d<=%d and K<=%d supported", kMaxD, kMaxK);
    if (!tt || !param_mins || !param_maxs || !VA || !mins || !maxs) return fail(h, NMMA_B200_ERR_ARG, "set_svd: NULL array");
    const size_t FT = (size_t)F * T;
    if (!all_finite(tt, FT) || !all_finite(VA, FT * K) || !all_finite(mins, FT) || !all_finite(maxs, FT) ||
        !all_finite(param_mins, (size_t)F * d) || !all_finite(param_maxs, (size_t)F * d))
        return fail(h, NMMA_B200_ERR_UNSUPPORTED, "set_svd: non-finite entries in the SVD model are not supported");
    for (int f = 0; f < F; ++f)
        for (int j = 1; j < T; ++j)
            if (!(tt[(size_t)f * T + j] > tt[(size_t)f * T + j - 1]))
                return fail(h, NMMA_B200_ERR_ARG, "set_svd: tt must be strictly increasing (filter %d, node %d)", f, j);
    h->F = F; h->d = d; h->K = K; h->T = T;
    h->tt.assign(tt, tt + FT);
    h->pmin.assign(param_mins, param_mins + (size_t)F * d);
    h->pmax.assign(param_maxs, param_maxs + (size_t)F * d);
    h->VA.assign(VA, VA + FT * K);
    h->mins.assign(mins, mins + FT);
    h->maxs.assign(maxs, maxs + FT);
    h->have_svd = true;
    h->kind = -1;
    h->ext_law = 0;   // per-filter extinction tables belong to the previous filter set
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_mlp(nmma_b200_t* h, int H, int K_out, const float* W1, const float* b1, const float* W2, const float* b2) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (!h->have_svd) return fail(h, NMMA_B200_ERR_STATE, "set_mlp: call nmma_b200_set_svd first");
    if (H < 1 || K_out < 1 || K_out > kMaxK) return fail(h, NMMA_B200_ERR_ARG, "set_mlp: bad H=%d / K_out=%d", H, K_out);
    if (!W1 || !b1 || !W2 || !b2) return fail(h, NMMA_B200_ERR_ARG, "set_mlp: NULL array");
    const size_t F = h->F, d = h->d;
    h->H = H; h->Kout = K_out;
    h->W1.assign(W1, W1 + F * d * H);
    h->b1.assign(b1, b1 + F * H);
    h->W2.assign(W2, W2 + F * H * K_out);
    h->b2.assign(b2, b2 + F * K_out);
    h->kind = 0;
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_gp(nmma_b200_t* h, int Ntr, const double* X, const double* alpha, const double* c2,
                     const double* rq_alpha, const double* rq_len, const double* ymean, const double* ystd) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (!h->have_svd) return fail(h, NMMA_B200_ERR_STATE, "set_gp: call nmma_b200_set_svd first");
    if (Ntr < 1) return fail(h, NMMA_B200_ERR_ARG, "set_gp: Ntr must be positive");
    if (!X || !alpha || !c2 || !rq_alpha || !rq_len || !ymean || !ystd) return fail(h, NMMA_B200_ERR_ARG, "set_gp: NULL array");
    if ((size_t)kGpPts * Ntr * sizeof(double) > 200 * 1024)
        return fail(h, NMMA_B200_ERR_UNSUPPORTED, "set_gp: Ntr=%d exceeds the shared-memory tile (max %d)", Ntr, (int)(200 * 1024 / (kGpPts * 8)));
    const size_t n = (size_t)h->F * h->K;
    h->Ntr = Ntr;
    h->gpX.assign(X, X + (size_t)Ntr * h->d);
    h->gpAlpha.assign(alpha, alpha + n * Ntr);
    h->gpC2.assign(c2, c2 + n);
    h->gpRa.assign(rq_alpha, rq_alpha + n);
    h->gpRl.assign(rq_len, rq_len + n);
    h->gpYm.assign(ymean, ymean + n);
    h->gpYs.assign(ystd, ystd + n);
    h->kind = 1;
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_sample_grid(nmma_b200_t* h, int S, const double* sample_times) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (S < 0) return fail(h, NMMA_B200_ERR_ARG, "set_sample_grid: S < 0");
    if (S == 0 || !sample_times) h->samp.clear();
    else h->samp.assign(sample_times, sample_times + S);
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_param_layout(nmma_b200_t* h, int P, const nmma_b200_param_src* model_params,
                               const nmma_b200_param_src* luminosity_distance, const nmma_b200_param_src* timeshift,
                               const nmma_b200_param_src* redshift, int z_mode) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (!h->have_svd) return fail(h, NMMA_B200_ERR_STATE, "set_param_layout: call nmma_b200_set_svd first");
    if (P < 1 || !model_params) return fail(h, NMMA_B200_ERR_ARG, "set_param_layout: P < 1 or NULL model_params");
    if (z_mode < 0 || z_mode > 2) return fail(h, NMMA_B200_ERR_ARG, "set_param_layout: unknown z_mode %d", z_mode);
    h->P = P;
    h->xsrc.clear();
    for (int i = 0; i < h->d; ++i) h->xsrc.push_back(to_src(model_params[i]));
    h->dl = luminosity_distance ? to_src(*luminosity_distance) : ParamSrc{-1, 0, 1e-5};   // 10 pc, nmma/em/model.py:291-293
    h->ts = timeshift ? to_src(*timeshift) : ParamSrc{-1, 0, 0.0};
    h->zsrc = redshift ? to_src(*redshift) : ParamSrc{-1, 0, 0.0};
    h->zmode = z_mode;
    h->have_layout = true;
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_redshift_table(nmma_b200_t* h, int n, const double* dist_grid, const double* z_grid) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (n < 0 || (n > 0 && (!dist_grid || !z_grid))) return fail(h, NMMA_B200_ERR_ARG, "set_redshift_table: bad arguments");
    for (int i = 1; i < n; ++i)
        if (!(dist_grid[i] >= dist_grid[i - 1])) return fail(h, NMMA_B200_ERR_ARG, "set_redshift_table: dist_grid must be sorted");
    h->zd.assign(dist_grid, dist_grid + n);
    h->zz.assign(z_grid, z_grid + n);
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_observations(nmma_b200_t* h, int G, const int32_t* n_helpers, const int32_t* helper_idx,
                               const int32_t* offsets, const double* t, const double* mag, const double* sigma_obs,
                               const double* det_limit) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (G < 1 || !n_helpers || !helper_idx || !offsets || !t || !mag || !sigma_obs || !det_limit)
        return fail(h, NMMA_B200_ERR_ARG, "set_observations: bad arguments");
    if (offsets[0] != 0) return fail(h, NMMA_B200_ERR_ARG, "set_observations: offsets[0] must be 0");
    for (int g = 0; g < G; ++g) {
        if (offsets[g + 1] < offsets[g]) return fail(h, NMMA_B200_ERR_ARG, "set_observations: offsets must be non-decreasing");
        if (n_helpers[g] < 1 || n_helpers[g] > NMMA_B200_MAX_HELPERS)
            return fail(h, NMMA_B200_ERR_ARG, "set_observations: n_helpers[%d]=%d outside 1..3", g, n_helpers[g]);
    }
    const int n = offsets[G];
    for (int k = 0; k < n; ++k)
        if (!std::isfinite(t[k])) return fail(h, NMMA_B200_ERR_ARG, "set_observations: observation time %d is not finite", k);
    h->G = G;
    h->g_nh.assign(n_helpers, n_helpers + G);
    h->g_h.assign(helper_idx, helper_idx + (size_t)G * 3);
    h->g_off.assign(offsets, offsets + G + 1);
    h->o_t.assign(t, t + n);
    h->o_m.assign(mag, mag + n);
    h->o_s.assign(sigma_obs, sigma_obs + n);
    h->g_lim.assign(det_limit, det_limit + G);
    h->have_obs = true;
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_systematics(nmma_b200_t* h, int G, const int32_t* mode, const double* budget,
                              const int32_t* n_nodes, const int32_t* node_offset,
                              const nmma_b200_param_src* node_src, const double* node_times) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (G < 1 || !mode || !budget || !n_nodes || !node_offset) return fail(h, NMMA_B200_ERR_ARG, "set_systematics: bad arguments");
    int total = 0;
    for (int g = 0; g < G; ++g) {
        if (mode[g] < 0 || mode[g] > 2) return fail(h, NMMA_B200_ERR_ARG, "set_systematics: unknown mode %d", mode[g]);
        const int nn = mode[g] == 0 ? 0 : (mode[g] == 1 ? 1 : n_nodes[g]);
        if (mode[g] == 2 && (nn < 1 || nn > kMaxSysNodes))
            return fail(h, NMMA_B200_ERR_UNSUPPORTED, "set_systematics: %d time nodes (1..%d supported)", nn, kMaxSysNodes);
        if (nn > 0) total = std::max(total, node_offset[g] + nn);
    }
    if (total > 0 && (!node_src || !node_times)) return fail(h, NMMA_B200_ERR_ARG, "set_systematics: NULL node arrays");
    h->sy_mode.assign(mode, mode + G);
    h->sy_budget.assign(budget, budget + G);
    h->sy_nn.assign(n_nodes, n_nodes + G);
    h->sy_off.assign(node_offset, node_offset + G);
    h->sy_src.clear();
    for (int i = 0; i < total; ++i) h->sy_src.push_back(to_src(node_src[i]));
    h->sy_t.assign(node_times, node_times + total);
    h->have_sys = true;
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_constraints(nmma_b200_t* h, int n, const nmma_b200_param_src* src, const double* minimum,
                              const double* maximum) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (n < 0 || n > kMaxCon) return fail(h, NMMA_B200_ERR_UNSUPPORTED, "set_constraints: %d constraints (0..%d supported)", n, kMaxCon);
    if (n > 0 && (!src || !minimum || !maximum)) return fail(h, NMMA_B200_ERR_ARG, "set_constraints: NULL array");
    h->con_src.clear(); h->con_lo.clear(); h->con_hi.clear();
    for (int i = 0; i < n; ++i) {
        if (std::isnan(minimum[i]) || std::isnan(maximum[i])) return fail(h, NMMA_B200_ERR_ARG, "set_constraints: NaN bound");
        h->con_src.push_back(to_src(src[i]));
        h->con_lo.push_back(minimum[i]);
        h->con_hi.push_back(maximum[i]);
    }
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_set_extinction(nmma_b200_t* h, int law, const nmma_b200_param_src* ebv, const double* nu0, const double* coef) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (!h->have_svd) return fail(h, NMMA_B200_ERR_STATE, "set_extinction: call nmma_b200_set_svd first");
    if (law < NMMA_B200_EXT_NONE || law > NMMA_B200_EXT_LINEAR) return fail(h, NMMA_B200_ERR_ARG, "set_extinction: unknown law %d", law);
    const int F = h->F;
    h->ext_law = law;
    h->ebv = ebv ? to_src(*ebv) : ParamSrc{-1, 0, 0.0};
    h->ext_nu.assign(F, 0.0);
    h->ext_coef.assign(F, 0.0);
    if (law == NMMA_B200_EXT_P92_SMC_HOST) {
        if (!nu0) return fail(h, NMMA_B200_ERR_ARG, "set_extinction: P92_SMC_host needs the filter frequencies nu0[F]");
        for (int f = 0; f < F; ++f) {
            if (std::isnan(nu0[f]) || nu0[f] < 0.0 || std::isinf(nu0[f])) return fail(h, NMMA_B200_ERR_ARG, "set_extinction: nu0[%d] must be finite and >= 0", f);
            h->ext_nu[f] = nu0[f];
        }
    } else if (law == NMMA_B200_EXT_LINEAR) {
        if (!coef) return fail(h, NMMA_B200_ERR_ARG, "set_extinction: the linear law needs coef[F] = A_f / E(B-V)");
        for (int f = 0; f < F; ++f) {
            if (!std::isfinite(coef[f])) return fail(h, NMMA_B200_ERR_ARG, "set_extinction: coef[%d] is not finite", f);
            h->ext_coef[f] = coef[f];
        }
    }
    h->dirty = true;
    return NMMA_B200_OK;
}

int nmma_b200_logl(nmma_b200_t* h, const double* points_dev, int64_t N, double* out_dev, void* stream) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (N < 0) return fail(h, NMMA_B200_ERR_ARG, "logl: N < 0");
    if (int rc = finalize(h, true)) return rc;
    if (N == 0) return NMMA_B200_OK;
    if (!points_dev || !out_dev) return fail(h, NMMA_B200_ERR_ARG, "logl: NULL device pointer");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int path = h->opt_path;
    if (path == 1 && !h->fused_supported)
        return fail(h, NMMA_B200_ERR_UNSUPPORTED, "fused kernel unavailable for this configuration (GP path, averaged filters, or d/K not instantiated)");
    if (path == 3 && !h->tc_supported)
        return fail(h, NMMA_B200_ERR_UNSUPPORTED, "tensor-core kernel unavailable for this configuration (GP path, averaged filters, d > 7 or n_coeff != 10)");
    if (path == 4 && !h->gp_fused_supported)
        return fail(h, NMMA_B200_ERR_UNSUPPORTED, "fused GP kernel unavailable for this configuration (MLP path, averaged filters, d outside 2..7, n_coeff != 10 or alpha > 1e5)");
    if (path == 5 && !h->tc_front_supported)
        return fail(h, NMMA_B200_ERR_UNSUPPORTED, "latency path unavailable for this configuration (GP path, d > 7 or n_coeff > 16)");
    if (path == 0) {
        if (h->tc_front_supported && N <= h->opt_latency_max) path = 5;
        else if (h->tc_supported && N >= h->opt_tc_min) path = 3;
        else if (h->gp_fused_supported && N >= h->opt_gp_min) path = 4;
        else path = (h->fused_supported && N >= h->opt_fused_min) ? 1 : 2;
    }
    h->last_path = path;
    if (path == 3) return launch_tc(h, points_dev, N, out_dev, st);
    if (path == 1) return launch_fused(h, points_dev, N, out_dev, st);
    if (path == 4) return launch_gp(h, points_dev, N, out_dev, st);
    if (path == 5) {
        // one point per call / small batches: filters and hidden ranges of each super-tile over all SMs (fp32 partial
        // coefficient sums), then one warp per point adds and scores them -- 2 launches
        const size_t FK = (size_t)h->F * h->K;
        if (int rc = ensure_scratch(h, (size_t)N * FK * 16 / 2 + 16)) return rc;   // <= 16 hidden ranges of fp32
        int hsplit = 1;
        float* parts = reinterpret_cast<float*>(h->coeff_scratch);
        if (int rc = launch_tc_coeff_parts(h, points_dev, N, parts, &hsplit, st)) return rc;
        const long long wpb = kBackThreads / 32;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((N + wpb - 1) / wpb, (long long)h->sm_count * 16));
        // FAST per-observation terms (lanes over observations) where the fused kernels would use them too; else generic fp64
        const bool fast = h->tc_supported && h->cfg.single_stage && h->cfg.uniform && !h->opt_no_fast;
        if (fast) backend_logl_parts_fast_kernel<10, true><<<(unsigned)std::min<long long>(N, (long long)h->sm_count * 8), kLatThreads,
                                                        ((FK + 1) & ~size_t(1)) * sizeof(float) + (kLatThreads / 32) * sizeof(double), st>>>(
                      h->cfg, points_dev, parts, hsplit, N, out_dev);
        else backend_logl_parts_kernel<<<grid, kBackThreads, wpb * FK * sizeof(double), st>>>(h->cfg, points_dev, parts, hsplit, N, out_dev);
        CU(cudaGetLastError());
        h->launches += 1;
        return NMMA_B200_OK;
    }
    const size_t FK = (size_t)h->F * h->K;
    const long long chunk = std::min<long long>(N, kTwoStageChunk);
    if (int rc = ensure_scratch(h, (size_t)chunk * FK)) return rc;
    for (long long n0 = 0; n0 < N; n0 += chunk) {
        const long long nn = std::min<long long>(chunk, N - n0);
        const double* pts = points_dev + n0 * h->P;
        if (int rc = launch_frontend(h, pts, nn, h->coeff_scratch, st)) return rc;
        const long long warps_per_block = kBackThreads / 32;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nn + warps_per_block - 1) / warps_per_block, (long long)h->sm_count * 16));
        // small batches of configurations the FAST back end covers: one CTA per point, one thread per observation
        const bool fast_lat = nn <= h->opt_latency_max && h->fast_backend_ok && h->K == 10 && !h->opt_no_fast;
        if (fast_lat)
            backend_logl_parts_fast_kernel<10, false><<<(unsigned)std::min<long long>(nn, (long long)h->sm_count * 8), kLatThreads,
                                                       ((FK + 1) & ~size_t(1)) * sizeof(float) + (kLatThreads / 32) * sizeof(double), st>>>(
                h->cfg, pts, reinterpret_cast<const float*>(h->coeff_scratch), 1, nn, out_dev + n0);
        else
            backend_logl_kernel<<<grid, kBackThreads, 0, st>>>(h->cfg, pts, h->coeff_scratch, nn, out_dev + n0);
        CU(cudaGetLastError());
        h->launches += 1;
    }
    return NMMA_B200_OK;
}

// True when `p` points into page-locked host memory (cudaMallocHost / cudaHostRegister / torch pin_memory).
static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
static bool is_device_mem(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice;
}

static void drop_lat_graphs(nmma_b200_t* h) {
    for (auto& g : h->lat_graphs) cudaGraphExecDestroy(g.exec);
    h->lat_graphs.clear();
    h->lat_warm.clear();
}

constexpr int64_t kZeroCopyMax = 256;   // rows up to which nmma_b200_logl_host skips the staging copies

// Shared body of nmma_b200_logl_host (out_dev == false: `out_host` is host memory) and nmma_b200_logl_host_to_device
// (out_dev == true: `out_host` is a device pointer, the result stays on the GPU and no D2H copy is made).
static int logl_host_impl(nmma_b200_t* h, const double* points_host, int64_t N, double* out_host, bool out_dev) {
    const size_t nin = (size_t)N * h->P, nout = (size_t)N;
    // page-locked caller buffers are copied directly; pageable ones go through pinned staging
    const bool in_pinned = is_pinned_host(points_host);
    const bool out_pinned = out_dev || is_pinned_host(out_host);
    if (nin > h->stage_cap_in || (!out_dev && nout > h->stage_cap_out)) drop_lat_graphs(h);
    if (nin > h->stage_cap_in) {
        if (h->stage_in_dev) cudaFree(h->stage_in_dev);
        if (h->stage_in_host) cudaFreeHost(h->stage_in_host);
        h->stage_in_dev = nullptr; h->stage_in_host = nullptr; h->stage_cap_in = 0;
        CU(cudaMalloc((void**)&h->stage_in_dev, nin * sizeof(double)));
        CU(cudaMallocHost((void**)&h->stage_in_host, nin * sizeof(double)));
        h->stage_cap_in = nin;
    }
    if (!out_dev && nout > h->stage_cap_out) {
        if (h->stage_out_dev) cudaFree(h->stage_out_dev);
        if (h->stage_out_host) cudaFreeHost(h->stage_out_host);
        h->stage_out_dev = nullptr; h->stage_out_host = nullptr; h->stage_cap_out = 0;
        CU(cudaMalloc((void**)&h->stage_out_dev, nout * sizeof(double)));
        CU(cudaMallocHost((void**)&h->stage_out_host, nout * sizeof(double)));
        h->stage_cap_out = nout;
    }
    double* dst = out_pinned ? out_host : h->stage_out_host;
    double* res = out_dev ? out_host : h->stage_out_dev;
    // Copy/compute pipeline: the batch is cut into row blocks (whole waves of the persistent throughput kernels);
    // block c+1 crosses PCIe (and, for pageable callers, is staged into pinned memory) while block c computes, and
    // block c-1 returns.  Three streams, one event pair per block; only the first H2D and the last D2H are exposed, so the
    // first block is ONE wave (its copy is what the first kernel waits for: 1.8 MB instead of 9 MB of a 10^6-row batch).
    long long nblk = 1;
    const long long wave = (long long)h->sm_count * 256;
    if (h->opt_pipeline > 1 && N >= 4 * wave) nblk = std::min<long long>(h->opt_pipeline, N / (2 * wave));
    const long long first = nblk > 1 ? wave : 0;                 // rows of the short first block (0 = no pipeline)
    long long rows = (N - first + nblk - 1) / nblk;              // rows of the others
    if (nblk > 1) rows = (rows + wave - 1) / wave * wave;
    nblk = (N - first + rows - 1) / rows + (first ? 1 : 0);
    if (nblk == 1 && N <= kZeroCopyMax && !out_dev && h->opt_zero_copy) {
        // latency path (one point per call from bilby / pymultinest, small live-point batches): the kernels read the rows
        // from and write the results to page-locked host memory directly (unified addressing), which removes two copy
        // calls and their DMA round trips from a call whose kernels take ~40 us (tools/latency_breakdown.py)
        const double* src = points_host;
        if (!in_pinned) { std::memcpy(h->stage_in_host, points_host, nin * sizeof(double)); src = h->stage_in_host; }
        // The latency path (path 5) reads each row from hundreds of threads in ~150 CTAs: over PCIe that costs more than
        // it saves (73 vs 60 us per one-point call), so its rows are copied (one small DMA); the result is still written in
        // place.  The filter-split fused kernel (9 CTAs, one reader per point) keeps reading in place.
        if (h->tc_front_supported && N <= h->opt_latency_max && (h->opt_path == 0 || h->opt_path == 5)) {
            if (h->opt_graphs) {
                // One-point calls are launch-bound (a DMA and two kernels of ~10 us each): the sequence is captured once per
                // batch size into a CUDA graph over the handle's own staging buffers and replayed with one launch call.
                // The first call of a size runs un-captured (allocations, function attributes), the second captures.
                if (!h->lat_graphs.empty() && h->lat_graphs[0].epoch != h->cfg_epoch) drop_lat_graphs(h);
                nmma_b200_handle::LatGraph* g = nullptr;
                for (auto& c : h->lat_graphs) if (c.N == N) g = &c;
                const bool warm = std::find(h->lat_warm.begin(), h->lat_warm.end(), N) != h->lat_warm.end();
                if (!g && warm) {
                    cudaGraph_t graph = nullptr;
                    cudaGraphExec_t exec = nullptr;
                    const long long l0 = h->launches;
                    bool ok = cudaStreamBeginCapture(h->own_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                    if (ok) {
                        ok = cudaMemcpyAsync(h->stage_in_dev, h->stage_in_host, nin * sizeof(double), cudaMemcpyHostToDevice, h->own_stream) == cudaSuccess;
                        ok = ok && nmma_b200_logl(h, h->stage_in_dev, N, h->stage_out_host, h->own_stream) == NMMA_B200_OK;
                        ok = (cudaStreamEndCapture(h->own_stream, &graph) == cudaSuccess) && ok && graph != nullptr;
                    }
                    if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
                    if (graph) cudaGraphDestroy(graph);
                    if (ok) {
                        h->lat_graphs.push_back({N, h->cfg_epoch, exec, h->launches - l0});
                        h->launches = l0;
                        g = &h->lat_graphs.back();
                    } else {
                        cudaGetLastError();
                        h->opt_graphs = 0;   // fall back to plain launches for the rest of this handle's life
                    }
                }
                if (g) {
                    if (src != h->stage_in_host) std::memcpy(h->stage_in_host, src, nin * sizeof(double));
                    CU(cudaGraphLaunch(g->exec, h->own_stream));
                    CU(cudaStreamSynchronize(h->own_stream));
                    std::memcpy(out_host, h->stage_out_host, nout * sizeof(double));
                    h->launches += g->launches;
                    h->last_path = 5;
                    return NMMA_B200_OK;
                }
                if (!warm) h->lat_warm.push_back(N);
            }
            CU(cudaMemcpyAsync(h->stage_in_dev, src, nin * sizeof(double), cudaMemcpyHostToDevice, h->own_stream));
            src = h->stage_in_dev;
        }
        if (int rc = nmma_b200_logl(h, src, N, dst, h->own_stream)) return rc;
        CU(cudaStreamSynchronize(h->own_stream));
    } else if (nblk == 1) {
        const double* src = points_host;
        if (!in_pinned) { std::memcpy(h->stage_in_host, points_host, nin * sizeof(double)); src = h->stage_in_host; }
        CU(cudaMemcpyAsync(h->stage_in_dev, src, nin * sizeof(double), cudaMemcpyHostToDevice, h->own_stream));
        if (int rc = nmma_b200_logl(h, h->stage_in_dev, N, res, h->own_stream)) return rc;
        if (!out_dev) CU(cudaMemcpyAsync(dst, res, nout * sizeof(double), cudaMemcpyDeviceToHost, h->own_stream));
        CU(cudaStreamSynchronize(h->own_stream));
    } else {
        if (!h->copy_in_stream) CU(cudaStreamCreateWithFlags(&h->copy_in_stream, cudaStreamNonBlocking));
        if (!h->copy_out_stream) CU(cudaStreamCreateWithFlags(&h->copy_out_stream, cudaStreamNonBlocking));
        while ((long long)h->pipe_events.size() < 2 * nblk) {
            cudaEvent_t ev;
            CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            h->pipe_events.push_back(ev);
        }
        for (long long c = 0; c < nblk; ++c) {
            const long long r0 = c == 0 ? 0 : first + (c - 1) * rows, nr = c == 0 ? first : std::min<long long>(rows, N - r0);
            const size_t o = (size_t)r0 * h->P, nb = (size_t)nr * h->P * sizeof(double);
            const double* src = points_host + o;
            if (!in_pinned) { std::memcpy(h->stage_in_host + o, points_host + o, nb); src = h->stage_in_host + o; }
            CU(cudaMemcpyAsync(h->stage_in_dev + o, src, nb, cudaMemcpyHostToDevice, h->copy_in_stream));
            CU(cudaEventRecord(h->pipe_events[2 * c], h->copy_in_stream));
            CU(cudaStreamWaitEvent(h->own_stream, h->pipe_events[2 * c], 0));
            if (int rc = nmma_b200_logl(h, h->stage_in_dev + o, nr, res + r0, h->own_stream)) return rc;
            if (out_dev) continue;
            CU(cudaEventRecord(h->pipe_events[2 * c + 1], h->own_stream));
            CU(cudaStreamWaitEvent(h->copy_out_stream, h->pipe_events[2 * c + 1], 0));
            CU(cudaMemcpyAsync(dst + r0, res + r0, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost,
                               h->copy_out_stream));
        }
        CU(cudaStreamSynchronize(out_dev ? h->own_stream : h->copy_out_stream));
    }
    if (!out_pinned) std::memcpy(out_host, h->stage_out_host, nout * sizeof(double));
    return NMMA_B200_OK;
}

int nmma_b200_logl_host(nmma_b200_t* h, const double* points_host, int64_t N, double* out_host) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (N < 0) return fail(h, NMMA_B200_ERR_ARG, "logl_host: N < 0");
    if (int rc = finalize(h, true)) return rc;
    if (N == 0) return NMMA_B200_OK;
    if (!points_host || !out_host) return fail(h, NMMA_B200_ERR_ARG, "logl_host: NULL pointer");
    CU(cudaSetDevice(h->device));
    if (is_device_mem(points_host) || is_device_mem(out_host))
        return fail(h, NMMA_B200_ERR_ARG, "logl_host: device pointer passed (use nmma_b200_logl, or nmma_b200_logl_host_to_device for a device result)");
    return logl_host_impl(h, points_host, N, out_host, false);
}

int nmma_b200_logl_host_to_device(nmma_b200_t* h, const double* points_host, int64_t N, double* out_dev) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (N < 0) return fail(h, NMMA_B200_ERR_ARG, "logl_host_to_device: N < 0");
    if (int rc = finalize(h, true)) return rc;
    if (N == 0) return NMMA_B200_OK;
    if (!points_host || !out_dev) return fail(h, NMMA_B200_ERR_ARG, "logl_host_to_device: NULL pointer");
    CU(cudaSetDevice(h->device));
    if (is_device_mem(points_host) || !is_device_mem(out_dev))
        return fail(h, NMMA_B200_ERR_ARG, "logl_host_to_device: points must be host memory and out a device pointer on this handle's GPU");
    return logl_host_impl(h, points_host, N, out_dev, true);
}

int nmma_b200_coeffs(nmma_b200_t* h, const double* points_dev, int64_t N, double* coeffs_dev, void* stream) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (N < 0) return fail(h, NMMA_B200_ERR_ARG, "coeffs: N < 0");
    if (int rc = finalize(h, false)) return rc;
    if (N == 0) return NMMA_B200_OK;
    if (!points_dev || !coeffs_dev) return fail(h, NMMA_B200_ERR_ARG, "coeffs: NULL device pointer");
    CU(cudaSetDevice(h->device));
    return launch_frontend(h, points_dev, N, coeffs_dev, static_cast<cudaStream_t>(stream));
}

int nmma_b200_mags(nmma_b200_t* h, const double* points_dev, int64_t N, int apparent, double* mags_dev,
                   double* tobs_dev, void* stream) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (N < 0) return fail(h, NMMA_B200_ERR_ARG, "mags: N < 0");
    if (int rc = finalize(h, false)) return rc;
    if (N == 0) return NMMA_B200_OK;
    if (!points_dev || !mags_dev) return fail(h, NMMA_B200_ERR_ARG, "mags: NULL device pointer");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t FK = (size_t)h->F * h->K;
    const long long chunk = std::min<long long>(N, kTwoStageChunk);
    if (int rc = ensure_scratch(h, (size_t)chunk * FK)) return rc;
    const int S = h->cfg.S;
    for (long long n0 = 0; n0 < N; n0 += chunk) {
        const long long nn = std::min<long long>(chunk, N - n0);
        const double* pts = points_dev + n0 * h->P;
        if (int rc = launch_frontend(h, pts, nn, h->coeff_scratch, st)) return rc;
        const long long total = nn * h->F * S;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)h->sm_count * 32));
        backend_mags_kernel<<<grid, 256, 0, st>>>(h->cfg, pts, h->coeff_scratch, nn, apparent,
                                                  mags_dev + (size_t)n0 * h->F * S,
                                                  tobs_dev ? tobs_dev + (size_t)n0 * S : nullptr);
        CU(cudaGetLastError());
        h->launches += 1;
    }
    return NMMA_B200_OK;
}

int nmma_b200_set_option(nmma_b200_t* h, const char* key, int64_t value) {
    if (!h || !key) return NMMA_B200_ERR_ARG;
    const std::string k(key);
    drop_lat_graphs(h);   // any knob may change which kernels a captured latency graph would have launched
    if (k == "path") { if (value < 0 || value > 5) return fail(h, NMMA_B200_ERR_ARG, "path must be 0 (auto), 1 (fused FFMA), 2 (two-stage), 3 (tensor core), 4 (fused GP) or 5 (latency: hidden-split tensor core + back end)"); h->opt_path = (int)value; }
    else if (k == "fused_min_points") h->opt_fused_min = value;
    else if (k == "tc_min_points") h->opt_tc_min = value;
    else if (k == "gp_min_points") h->opt_gp_min = value;
    else if (k == "tc_front_min_points") h->opt_tc_front_min = value;
    else if (k == "latency_max_points") h->opt_latency_max = value;
    else if (k == "max_ctas") h->opt_max_ctas = (int)value;
    else if (k == "pipeline_blocks") { if (value < 1 || value > 64) return fail(h, NMMA_B200_ERR_ARG, "pipeline_blocks must be 1..64"); h->opt_pipeline = (int)value; }
    else if (k == "packed_fma") { /* kept for compatibility: two points per thread always use FFMA2 */ }
    else if (k == "no_fast_backend") h->opt_no_fast = value ? 1 : 0;
    else if (k == "no_filter_split") h->opt_no_fsplit = value ? 1 : 0;
    else if (k == "zero_copy") h->opt_zero_copy = value ? 1 : 0;
    else if (k == "cuda_graphs") h->opt_graphs = value ? 1 : 0;
    else if (k == "points_per_thread") { if (value != 0 && value != 1 && value != 2 && value != 4) return fail(h, NMMA_B200_ERR_ARG, "points_per_thread must be 0 (auto), 1, 2 or 4"); h->opt_pt = (int)value; }
    else return fail(h, NMMA_B200_ERR_ARG, "unknown option '%s'", key);
    return NMMA_B200_OK;
}

int nmma_b200_get_info(nmma_b200_t* h, const char* key, int64_t* value) {
    if (!h || !key || !value) return NMMA_B200_ERR_ARG;
    const std::string k(key);
    if (k == "launches") *value = h->launches;
    else if (k == "last_path") *value = h->last_path;
    else if (k == "sm_count") *value = h->sm_count;
    else if (k == "ctas_per_sm") *value = h->last_ctas_per_sm;
    else if (k == "fused_supported") { if (int rc = finalize(h, true)) return rc; *value = h->fused_supported ? 1 : 0; }
    else if (k == "tc_supported") { if (int rc = finalize(h, true)) return rc; *value = h->tc_supported ? 1 : 0; }
    else if (k == "tc_front_supported") { if (int rc = finalize(h, true)) return rc; *value = h->tc_front_supported ? 1 : 0; }
    else if (k == "gp_fused_supported") { if (int rc = finalize(h, true)) return rc; *value = h->gp_fused_supported ? 1 : 0; }
    else if (k == "algorithmic_flop_per_eval") {
        // F * [2 H (d + K) + 2 T K] (MLP) or Ntr*3d + F*K*Ntr*8 + F*2*T*K (GP)
        if (h->kind == 0) *value = (int64_t)h->F * (2LL * h->H * (h->d + h->K) + 2LL * h->T * h->K);
        else if (h->kind == 1) *value = (int64_t)h->Ntr * 3 * h->d + (int64_t)h->F * h->K * h->Ntr * 8 + (int64_t)h->F * 2 * h->T * h->K;
        else return fail(h, NMMA_B200_ERR_STATE, "no surrogate configured");
    } else if (k == "tc_executed_flop_per_eval") {
        // tensor-core kernel: per chunk and point 1-2 layer-1 MMAs (N = chunk, K = 16) + per 16 hidden units one N = 32 MMA
        // (h_hi [W_hi | W_lo']) and one N = 16 MMA (h_lo W_hi), K = 16, kind::f16
        if (int rc = finalize(h, false)) return rc;
        *value = (int64_t)h->F * h->cfg.tc_nch * ((3 * (h->d + 1) <= 16 ? 1LL : 2LL) * 2 * kTcChunk * 16 + 1LL * kTcKSteps * 2 * (3 * kTcN2) * 16);
    } else return fail(h, NMMA_B200_ERR_ARG, "unknown info key '%s'", key);
    return NMMA_B200_OK;
}

int nmma_b200_ffma_peak(nmma_b200_t* h, int variant, int iters, double* flops_per_s) {
    if (!h || !flops_per_s || iters < 1) return NMMA_B200_ERR_ARG;
    CU(cudaSetDevice(h->device));
    float* sink = nullptr;
    CU(cudaMalloc((void**)&sink, sizeof(float)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const unsigned grid = (unsigned)h->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, h->own_stream));
        if (variant == 1) ffma_peak_kernel<true><<<grid, 256, 0, h->own_stream>>>(iters, 1.0f, sink);
        else ffma_peak_kernel<false><<<grid, 256, 0, h->own_stream>>>(iters, 1.0f, sink);
        CU(cudaGetLastError());
        CU(cudaEventRecord(e1, h->own_stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        h->launches += 1;
        const double fl = 2.0 * 16.0 * (double)iters * 256.0 * grid;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *flops_per_s = best;
    return NMMA_B200_OK;
}

int nmma_b200_obs_terms(nmma_b200_t* h, int n, const double* mag, const double* model_mag, const double* sigma_obs,
                        const double* sigma_sys, const double* det_limit, double* out) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (n < 0 || (n > 0 && (!mag || !model_mag || !sigma_obs || !sigma_sys || !det_limit || !out)))
        return fail(h, NMMA_B200_ERR_ARG, "obs_terms: bad arguments");
    if (n == 0) return NMMA_B200_OK;
    CU(cudaSetDevice(h->device));
    double* buf = nullptr;
    const size_t nb = (size_t)n * sizeof(double);
    CU(cudaMalloc((void**)&buf, 6 * nb));
    const double* src[5] = {mag, model_mag, sigma_obs, sigma_sys, det_limit};
    for (int i = 0; i < 5; ++i) {
        cudaError_t e = cudaMemcpy(buf + (size_t)i * n, src[i], nb, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(buf); return fail(h, NMMA_B200_ERR_CUDA, "cudaMemcpy: %s", cudaGetErrorString(e)); }
    }
    obs_terms_kernel<<<(n + 255) / 256, 256, 0, h->own_stream>>>(n, buf, buf + n, buf + 2 * (size_t)n, buf + 3 * (size_t)n,
                                                                buf + 4 * (size_t)n, buf + 5 * (size_t)n);
    h->launches += 1;
    cudaError_t e = cudaStreamSynchronize(h->own_stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, buf + 5 * (size_t)n, nb, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(h, NMMA_B200_ERR_CUDA, "obs_terms: %s", cudaGetErrorString(e));
    return NMMA_B200_OK;
}

}  // extern "C"
