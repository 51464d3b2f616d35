// Device-side restatement of the reference's scalar numerics (fp64).
//
// Every helper names the reference / third-party routine whose *semantics* it
// follows; the arithmetic order is kept where an index or a mask depends on it
// (np.interp interval search, detector-frame time grid), using __dmul_rn /
// __dadd_rn so that nvcc cannot contract a*b+c into an FMA there.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace nmma {

#define NMMA_SENTINEL (-1.7976931348623157e308)  // np.nan_to_num(-np.inf), nmma/core/base.py:82
#define NMMA_PI 3.141592653589793
#define NMMA_SQRT1_2 0.7071067811865476
#define NMMA_NORM_PDF_LOGC 0.9189385332046727  // log(2*pi)/2, scipy _norm_pdf_logC

struct ParamSrc {
    int32_t col;
    int32_t xf;
    double val;
};

// scipy.special.ndtr (cephes ndtr.c): 0.5*erfc(-x/sqrt2) split at |x|/sqrt2 < sqrt(1/2).
__device__ __forceinline__ double ndtr(double a) {
    if (isnan(a)) return a;
    const double x = a * NMMA_SQRT1_2;
    const double z = fabs(x);
    if (z < NMMA_SQRT1_2) return 0.5 + 0.5 * erf(x);
    double y = 0.5 * erfc(z);
    if (x > 0) y = 1.0 - y;
    return y;
}

// scipy.special.log_ndtr (xsf): log(erfcx(-t)/2) - t^2 for a < -1, log1p(-erfc(t)/2) otherwise.
__device__ __forceinline__ double log_ndtr(double a) {
    const double t = a * NMMA_SQRT1_2;
    if (a < -1.0) return log(erfcx(-t) / 2) - t * t;
    return log1p(-erfc(t) / 2);
}

// np.interp (numpy/_core/src/multiarray/compiled_base.c, arr_interp) for one x.
// xp is sorted ascending with n >= 1 entries; fp are the values.
__device__ __forceinline__ double np_interp(double x, const double* __restrict__ xp,
                                            const double* __restrict__ fp, int n, double left,
                                            double right) {
    if (isnan(x)) return x;
    if (x > xp[n - 1]) return right;
    if (x < xp[0]) return left;
    int lo = 0, hi = n - 1;  // invariant xp[lo] <= x, answer j = last index with xp[j] <= x
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    int j = (xp[hi] <= x) ? hi : lo;
    if (j == n - 1) return fp[j];
    if (xp[j] == x) return fp[j];
    const double slope = __ddiv_rn(__dsub_rn(fp[j + 1], fp[j]), __dsub_rn(xp[j + 1], xp[j]));
    double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[j])), fp[j]);
    if (isnan(r)) {
        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[j + 1])), fp[j + 1]);
        if (isnan(r) && fp[j] == fp[j + 1]) r = fp[j];
    }
    return r;
}

// One per-point scalar: column or constant, then the reference's conversion
// (nmma/core/conversion.py:119-126, nmma/em/model.py:276-283).
__device__ __forceinline__ double eval_src(const ParamSrc& s, const double* __restrict__ row) {
    double v = (s.col >= 0) ? row[s.col] : s.val;
    switch (s.xf) {
        case 1: v = __ddiv_rn(__dmul_rn(v, 180.0), NMMA_PI); break;
        case 2: v = log10(v); break;
        case 3: v = pow(10.0, v); break;
        case 5: v = acos(v);  // fallthrough: theta_jn from cos_theta_jn
        case 4: {
            const double w = __dsub_rn(NMMA_PI, v);
            v = (isnan(v) || isnan(w)) ? CUDART_NAN : (v < w ? v : w);  // np.minimum propagates NaN
            v = __ddiv_rn(__dmul_rn(v, 180.0), NMMA_PI);
            break;
        }
        default: break;
    }
    return v;
}

// One photometric point, nmma/em/em_likelihood.py:224-256:
//   finite sigma  -> scipy.stats.truncnorm.logpdf(m, -inf, (lim-mu)/sigma, loc=mu, scale=sigma)
//   infinite/NaN sigma (upper limit) -> scipy.stats.norm.logsf(m, mu, sigma_sys)
// including the generic-distribution wrappers' masks (rv_continuous.logpdf / logsf):
// NaN for invalid args, -inf outside the support, 0.0 for logsf at x <= -inf.
__device__ __forceinline__ double obs_term(double m, double mu, double sobs, double ssys, double lim) {
    const double sig = sqrt(sobs * sobs + ssys * ssys);
    if (isfinite(sig)) {
        const double x = __ddiv_rn(__dsub_rn(m, mu), sig);
        const double b = __ddiv_rn(__dsub_rn(lim, mu), sig);
        const bool cond0 = (b > -CUDART_INF) && (sig > 0.0);  // _argcheck a < b (false for NaN b), scale > 0
        if (!cond0 || isnan(x)) return CUDART_NAN;
        if (!(x <= b)) return -CUDART_INF;  // _support_mask (a = -inf <= x always)
        // _log_gauss_mass(-inf, b): log_ndtr(b) in the left case, log1p(-ndtr(-inf) - ndtr(-b)) centrally
        const double mass = (b <= 0.0) ? log_ndtr(b) : log1p(-0.0 - ndtr(-b));
        return -(x * x) / 2.0 - NMMA_NORM_PDF_LOGC - mass - log(sig);
    }
    const double x = __ddiv_rn(__dsub_rn(m, mu), ssys);
    if (!(ssys > 0.0) || isnan(x)) return CUDART_NAN;
    if (x <= -CUDART_INF) return 0.0;
    if (!(x < CUDART_INF)) return -CUDART_INF;
    return log_ndtr(-x);
}

// Same as obs_term for a detection whose total sigma is known on the host
// (constant error budget): sigma, log(sigma) + log(2 pi)/2 are staged once.
__device__ __forceinline__ double obs_term_static_det(double m, double mu, double sig, double logsig_c,
                                                      double lim) {
    const double x = __ddiv_rn(__dsub_rn(m, mu), sig);
    const double b = __ddiv_rn(__dsub_rn(lim, mu), sig);
    if (!(b > -CUDART_INF) || !(sig > 0.0) || isnan(x)) return CUDART_NAN;
    if (!(x <= b)) return -CUDART_INF;
    double r = -(x * x) / 2.0 - logsig_c;
    if (b < CUDART_INF) r -= (b <= 0.0) ? log_ndtr(b) : log1p(-0.0 - ndtr(-b));
    return r;
}

}  // namespace nmma
