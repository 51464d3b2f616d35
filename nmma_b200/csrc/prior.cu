// Priors on the device (include/nmma_b200.h, "priors on the device"): unit cube -> physical point
// (bilby PriorDict.rescale, one analytic prior per column of points[N,P]), counter-based sampling of the
// unit cube (Philox4x32-10), and the host-traffic-free prior sweep.  HBM-bound element-wise work: one
// thread per point, the P columns of a row read / written as consecutive doubles (rows are 8 P bytes, so a
// warp touches 32 consecutive rows = one contiguous 256 P-byte span; the stores coalesce in L2).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "handle.h"
#include "device_math.cuh"

namespace nmma {

struct PriorCol {
    int kind;
    int tab_off, tab_n;
    double a, b, c, e;   // derived constants, see prior_rescale
};
struct PriorPlan {
    int P;
    PriorCol col[NMMA_B200_MAX_P];
};

// bilby.core.prior.analytical.*.rescale in the operation order of the Python expressions (header comment of
// include/nmma_b200.h); the constants a..e are precomputed in nmma_b200_set_priors with the same expressions.
__device__ __forceinline__ double prior_rescale(const PriorCol& p, double u, const double* __restrict__ tab_cdf,
                                                const double* __restrict__ tab_grid) {
    switch (p.kind) {
        case NMMA_B200_PR_UNIFORM:      // minimum + val * (maximum - minimum);   a = min, b = max - min
            return __dadd_rn(p.a, __dmul_rn(u, p.b));
        case NMMA_B200_PR_DELTA:
            return p.a;
        case NMMA_B200_PR_SINE:         // arccos(cos(min) - val / norm);          a = cos(min), b = norm
            return acos(__dsub_rn(p.a, __ddiv_rn(u, p.b)));
        case NMMA_B200_PR_COSINE:       // arcsin(val / norm + sin(min));          a = sin(min), b = norm
            return asin(__dadd_rn(__ddiv_rn(u, p.b), p.a));
        case NMMA_B200_PR_GAUSSIAN:     // mu + erfinv(2 val - 1) * 2**0.5 * sigma; a = mu, b = sigma
            return __dadd_rn(p.a, __dmul_rn(__dmul_rn(erfinv(__dsub_rn(__dmul_rn(2.0, u), 1.0)), 1.4142135623730951), p.b));
        case NMMA_B200_PR_TRUNC_GAUSS:  // erfinv(2 val norm + erf_lo) * 2**0.5 * sigma + mu; c = norm, e = erf_lo
            return __dadd_rn(__dmul_rn(__dmul_rn(erfinv(__dadd_rn(__dmul_rn(__dmul_rn(2.0, u), p.c), p.e)), 1.4142135623730951), p.b), p.a);
        case NMMA_B200_PR_POWERLAW:
            if (p.a == -1.0)            // minimum * exp(val * log(maximum / minimum));  b = min, c = log(max / min)
                return __dmul_rn(p.b, exp(__dmul_rn(u, p.c)));
            // (min**a1 + val * (max**a1 - min**a1)) ** (1 / a1);  b = min**a1, c = max**a1 - min**a1, e = 1 / a1
            return pow(__dadd_rn(p.b, __dmul_rn(u, p.c)), p.e);
        case NMMA_B200_PR_TRIANGULAR: { // a = min, b = max, c = mode, e = (mode - min) / (max - min)
            if (u < p.e) return __dadd_rn(p.a, sqrt(__dmul_rn(__dmul_rn(fmax(u, 0.0), __dsub_rn(p.b, p.a)), __dsub_rn(p.c, p.a))));
            return __dsub_rn(p.b, sqrt(__dmul_rn(__dmul_rn(fmax(__dsub_rn(1.0, u), 0.0), __dsub_rn(p.b, p.a)), __dsub_rn(p.b, p.c))));
        }
        case NMMA_B200_PR_INTERPED:     // np.interp(val, cdf, grid)
            return np_interp(u, tab_cdf + p.tab_off, tab_grid + p.tab_off, p.tab_n, tab_grid[p.tab_off],
                             tab_grid[p.tab_off + p.tab_n - 1]);
    }
    return CUDART_NAN;
}

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011; Random123 constants) ----
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 53 random bits -> double in [0, 1) (the construction of numpy's / torch's next_double)
__device__ __forceinline__ double u01_53(uint32_t a, uint32_t b) {
    return (double)(((unsigned long long)(a >> 5) << 26) | (unsigned long long)(b >> 6)) * (1.0 / 9007199254740992.0);
}

constexpr int kPriorThreads = 256;

// SAMPLE: draw the unit cube from (seed, first + i); else read it from `unit`.
template <bool SAMPLE>
__global__ void __launch_bounds__(kPriorThreads) prior_kernel(const PriorPlan* __restrict__ plan_g,
                                                             const double* __restrict__ tab, int tab_total,
                                                             const double* __restrict__ unit, unsigned long long seed,
                                                             long long first, long long N, double* __restrict__ pts,
                                                             double* __restrict__ unit_out) {
    __shared__ PriorPlan plan;
    for (int i = threadIdx.x; i < (int)(sizeof(PriorPlan) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&plan)[i] = reinterpret_cast<const uint32_t*>(plan_g)[i];
    __syncthreads();
    const int P = plan.P;
    const double* tab_cdf = tab;
    const double* tab_grid = tab + tab_total;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long gidx = (unsigned long long)(first + i);
        double* row = pts + i * P;
        for (int j0 = 0; j0 < P; j0 += 2) {
            double u0, u1 = 0.0;
            if (SAMPLE) {
                uint32_t r[4];
                philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)(j0 >> 1), 0u, (uint32_t)seed,
                              (uint32_t)(seed >> 32), r);
                u0 = u01_53(r[0], r[1]);
                u1 = u01_53(r[2], r[3]);
                if (unit_out) {
                    unit_out[i * P + j0] = u0;
                    if (j0 + 1 < P) unit_out[i * P + j0 + 1] = u1;
                }
            } else {
                u0 = unit[i * P + j0];
                if (j0 + 1 < P) u1 = unit[i * P + j0 + 1];
            }
            row[j0] = prior_rescale(plan.col[j0], u0, tab_cdf, tab_grid);
            if (j0 + 1 < P) row[j0 + 1] = prior_rescale(plan.col[j0 + 1], u1, tab_cdf, tab_grid);
        }
    }
}

}  // namespace nmma

using namespace nmma;

namespace {
int launch_prior(nmma_b200_t* h, bool sample, const double* unit, unsigned long long seed, long long first, long long N,
                 double* pts, double* unit_out, cudaStream_t st) {
    const long long blocks = (N + kPriorThreads - 1) / kPriorThreads;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(blocks, (long long)h->sm_count * 8));
    const int tab_total = h->pr_tab_total;   // the grid half of the table starts after all cdf entries
    if (sample)
        prior_kernel<true><<<grid, kPriorThreads, 0, st>>>(h->pr_dev, h->pr_tab_dev, tab_total, nullptr, seed, first, N, pts, unit_out);
    else
        prior_kernel<false><<<grid, kPriorThreads, 0, st>>>(h->pr_dev, h->pr_tab_dev, tab_total, unit, 0ull, 0, N, pts, nullptr);
    CU(cudaGetLastError());
    h->launches += 1;
    return NMMA_B200_OK;
}

int prior_ready(nmma_b200_t* h, const char* what) {
    if (!h->pr_dev) return fail(h, NMMA_B200_ERR_STATE, "%s: nmma_b200_set_priors has not been called", what);
    return NMMA_B200_OK;
}
}  // namespace

extern "C" {

int nmma_b200_set_priors(nmma_b200_t* h, int P, const int32_t* kind, const double* par, const int32_t* tab_offset,
                         const double* tab_cdf, const double* tab_grid) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (P < 1 || P > NMMA_B200_MAX_P) return fail(h, NMMA_B200_ERR_ARG, "set_priors: P=%d outside 1..%d", P, NMMA_B200_MAX_P);
    if (!kind || !par) return fail(h, NMMA_B200_ERR_ARG, "set_priors: NULL kind/par");
    if (h->have_layout && h->P != P)
        return fail(h, NMMA_B200_ERR_ARG, "set_priors: P=%d differs from the parameter layout's P=%d", P, h->P);
    PriorPlan plan;
    std::memset(&plan, 0, sizeof plan);
    plan.P = P;
    int tab_total = 0;
    for (int j = 0; j < P; ++j) {
        PriorCol& c = plan.col[j];
        const double* q = par + 4 * j;
        c.kind = kind[j];
        switch (kind[j]) {
            case NMMA_B200_PR_UNIFORM:
                if (!(q[1] >= q[0]) || !std::isfinite(q[0]) || !std::isfinite(q[1]))
                    return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d Uniform needs finite minimum <= maximum", j);
                c.a = q[0]; c.b = q[1] - q[0];
                break;
            case NMMA_B200_PR_DELTA:
                c.a = q[0];
                break;
            case NMMA_B200_PR_SINE:
                c.a = std::cos(q[0]); c.b = 1.0 / (std::cos(q[0]) - std::cos(q[1]));
                if (!std::isfinite(c.b)) return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d Sine has an empty range", j);
                break;
            case NMMA_B200_PR_COSINE:
                c.a = std::sin(q[0]); c.b = 1.0 / (std::sin(q[1]) - std::sin(q[0]));
                if (!std::isfinite(c.b)) return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d Cosine has an empty range", j);
                break;
            case NMMA_B200_PR_GAUSSIAN:
                if (!(q[1] > 0)) return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d Gaussian needs sigma > 0", j);
                c.a = q[0]; c.b = q[1];
                break;
            case NMMA_B200_PR_TRUNC_GAUSS: {
                if (!(q[1] > 0) || !(q[3] > q[2]))
                    return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d TruncatedGaussian needs sigma > 0, maximum > minimum", j);
                c.a = q[0]; c.b = q[1];
                const double s2 = std::pow(2.0, 0.5);
                const double elo = std::erf((q[2] - q[0]) / s2 / q[1]), ehi = std::erf((q[3] - q[0]) / s2 / q[1]);
                c.c = (ehi - elo) / 2; c.e = elo;
                break;
            }
            case NMMA_B200_PR_POWERLAW:
                if (!(q[1] > 0) || !(q[2] > q[1]))
                    return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d PowerLaw needs 0 < minimum < maximum", j);
                c.a = q[0];
                if (q[0] == -1.0) { c.b = q[1]; c.c = std::log(q[2] / q[1]); }
                else { const double a1 = 1 + q[0]; c.b = std::pow(q[1], a1); c.c = std::pow(q[2], a1) - std::pow(q[1], a1); c.e = 1.0 / a1; }
                break;
            case NMMA_B200_PR_TRIANGULAR:
                if (!(q[2] > q[1]) || q[0] < q[1] || q[0] > q[2])
                    return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d Triangular needs minimum <= mode <= maximum", j);
                c.a = q[1]; c.b = q[2]; c.c = q[0]; c.e = (q[0] - q[1]) / (q[2] - q[1]);
                break;
            case NMMA_B200_PR_INTERPED: {
                if (!tab_offset || !tab_cdf || !tab_grid) return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d Interped without tables", j);
                const int n = tab_offset[j + 1] - tab_offset[j];
                if (n < 2 || tab_offset[j] < 0) return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d Interped table needs >= 2 nodes", j);
                for (int i = 1; i < n; ++i)
                    if (!(tab_cdf[tab_offset[j] + i] >= tab_cdf[tab_offset[j] + i - 1]))
                        return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d Interped cdf is not non-decreasing", j);
                c.tab_off = tab_offset[j]; c.tab_n = n;
                tab_total = std::max(tab_total, tab_offset[j + 1]);
                break;
            }
            default:
                return fail(h, NMMA_B200_ERR_ARG, "set_priors: column %d has unknown prior kind %d", j, kind[j]);
        }
    }
    CU(cudaSetDevice(h->device));
    if (!h->pr_dev) CU(cudaMalloc((void**)&h->pr_dev, sizeof(PriorPlan)));
    CU(cudaMemcpy(h->pr_dev, &plan, sizeof plan, cudaMemcpyHostToDevice));
    if (h->pr_tab_dev) { cudaFree(h->pr_tab_dev); h->pr_tab_dev = nullptr; }
    h->pr_tab_total = tab_total;
    if (tab_total > 0) {
        CU(cudaMalloc((void**)&h->pr_tab_dev, (size_t)2 * tab_total * sizeof(double)));
        CU(cudaMemcpy(h->pr_tab_dev, tab_cdf, (size_t)tab_total * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->pr_tab_dev + tab_total, tab_grid, (size_t)tab_total * sizeof(double), cudaMemcpyHostToDevice));
    }
    h->prP = P;
    return NMMA_B200_OK;
}

int nmma_b200_prior_transform(nmma_b200_t* h, const double* unit_dev, int64_t N, double* points_dev, void* stream) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (N < 0) return fail(h, NMMA_B200_ERR_ARG, "prior_transform: N < 0");
    if (int rc = prior_ready(h, "prior_transform")) return rc;
    if (N == 0) return NMMA_B200_OK;
    if (!unit_dev || !points_dev) return fail(h, NMMA_B200_ERR_ARG, "prior_transform: NULL device pointer");
    CU(cudaSetDevice(h->device));
    return launch_prior(h, false, unit_dev, 0ull, 0, N, points_dev, nullptr, static_cast<cudaStream_t>(stream));
}

int nmma_b200_prior_sample(nmma_b200_t* h, uint64_t seed, int64_t first_index, int64_t N, double* points_dev,
                           double* unit_dev, void* stream) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (N < 0 || first_index < 0) return fail(h, NMMA_B200_ERR_ARG, "prior_sample: N or first_index < 0");
    if (int rc = prior_ready(h, "prior_sample")) return rc;
    if (N == 0) return NMMA_B200_OK;
    if (!points_dev) return fail(h, NMMA_B200_ERR_ARG, "prior_sample: NULL device pointer");
    CU(cudaSetDevice(h->device));
    return launch_prior(h, true, nullptr, seed, first_index, N, points_dev, unit_dev, static_cast<cudaStream_t>(stream));
}

int nmma_b200_logl_sweep(nmma_b200_t* h, uint64_t seed, int64_t first_index, int64_t N, double* out_dev,
                         double* points_dev, void* stream) {
    if (!h) return NMMA_B200_ERR_ARG;
    if (N < 0 || first_index < 0) return fail(h, NMMA_B200_ERR_ARG, "logl_sweep: N or first_index < 0");
    if (int rc = prior_ready(h, "logl_sweep")) return rc;
    if (!h->have_layout) return fail(h, NMMA_B200_ERR_STATE, "logl_sweep: parameter layout not set");
    if (h->prP != h->P) return fail(h, NMMA_B200_ERR_ARG, "logl_sweep: priors have P=%d, layout has P=%d", h->prP, h->P);
    if (N == 0) return NMMA_B200_OK;
    if (!out_dev) return fail(h, NMMA_B200_ERR_ARG, "logl_sweep: NULL device pointer");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // block ~ 2^20 points: 8 P MB of drawn points (48 MB at P = 6) stay in the 126 MB L2 between the two kernels; whole
    // waves of the persistent throughput kernels (256 points per SM), so that no block ends in a partly filled round
    const long long wave = (long long)h->sm_count * 256;
    const long long block = std::min<long long>(N, std::max<long long>(wave, (1ll << 20) / wave * wave));
    if (!points_dev && (size_t)block * h->P > h->sweep_cap) {
        if (h->sweep_scratch) cudaFree(h->sweep_scratch);
        h->sweep_scratch = nullptr; h->sweep_cap = 0;
        CU(cudaMalloc((void**)&h->sweep_scratch, (size_t)block * h->P * sizeof(double)));
        h->sweep_cap = (size_t)block * h->P;
    }
    for (long long n0 = 0; n0 < N; n0 += block) {
        const long long nn = std::min<long long>(block, N - n0);
        double* pts = points_dev ? points_dev + n0 * h->P : h->sweep_scratch;
        if (int rc = launch_prior(h, true, nullptr, seed, first_index + n0, nn, pts, nullptr, st)) return rc;
        if (int rc = nmma_b200_logl(h, pts, nn, out_dev + n0, st)) return rc;
    }
    return NMMA_B200_OK;
}

}  // extern "C"
