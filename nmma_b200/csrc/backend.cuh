// Device configuration block and the fp64 back end shared by all kernels:
// SVD reconstruction on the bracketing grid nodes, the two np.interp stages,
// systematics, and the per-observation likelihood term.
#pragma once
#include "device_math.cuh"

namespace nmma {

constexpr int kMaxD = 16;
constexpr int kMaxK = 16;
constexpr int kMaxSysNodes = 16;
constexpr int kMaxCon = 8;   // Constraint priors evaluated per point (nmma_b200_set_constraints)

// Passed by value to every kernel (fits the 4 KB parameter space).
struct DevCfg {
    int F, d, K, T, S, H, HP, RW, Ntr, P, G, nobs, kind;  // kind: 0 = MLP, 1 = GP
    int single_stage;  // sample grid == tt of every filter: stage 1 is the identity
    int uniform;       // sample grid is uniform: O(1) interval guess instead of bisection
    int static_fail;   // some model filter has < 2 sample nodes inside its training range
    double uni_s0, uni_inv_ds;
    // surrogate basis, per filter
    const double* pmin;   // F*d
    const double* pden;   // F*d   param_maxs - param_mins
    const double* bpack;  // F*T*(K+2): row j = [VA[j,0..K-1], maxs[j]-mins[j], mins[j]]
    // fp32 copy for the FAST back end (kernels.cuh: fused_filter_logl): F*(K+2)*T float2, entry (i, j) = column i of
    // rows j and min(j+1, T-1) -- the two grid nodes that bracket an observation come back in one 8-byte load, and
    // lanes with different j hit different banks.  Same byte count per filter as bpack.
    const float2* bpack32;
    float fast_delta;     // FAST back end: |frac(index guess) - {0,1}| below this -> settle the interval in fp64
    // stage 1: tt -> sample grid (np.interp with +inf outside the training range)
    const double* samp;   // S
    const int* s_lo;      // F  first sample node inside [tt[0], tt[-1]]
    const int* s_hi;      // F  last one
    const int* s1_j;      // F*S  left tt node of sample node s
    const double* s1_dx;  // F*S  sample[s] - tt[j]
    const double* s1_dt;  // F*S  tt[j+1] - tt[j]; 0 marks "exactly on node j"
    // MLP front end
    const float* wpack;   // F*HP*RW: row j = [W1[0..d-1][j], b1[j], W2[j][0..K-1], 0-pad]
    const float* b2;      // F*K
    // tensor-core front end (tc_kernel.cuh): per (filter, chunk of hidden units) kTcChunkFloats of fp16 B tiles
    // B1[2 k-steps] | B2[k-steps] = [W_hi | 2^11 W_lo], all scaled into the fp16 range by exact powers of two
    const float* tcpack;  // F*tc_nch*kTcChunkFloats
    const float* tc_xs;   // F*8   2^r_i: multiplier of scaled input i (i = d: the bias "input" 1) matching row i of the staged W1
    const float* tc_s2inv;// F*16  2^-q_k: undoes the scaling of column k of the staged W2
    int tc_nch;           // chunks per filter (even)
    // GP front end
    const double* gpX;    // Ntr*d
    const double* gpA;    // F*K*Ntr  constant_value * alpha_
    const double* gpAT;   // F*Ntr*K  the same, the K values of a training row adjacent (fused GP kernel)
    const double* gp_q;   // F*K      1 / (2 * alpha * length_scale^2)
    const double* gp_ra;  // F*K      alpha
    const double* gp_ym;  // F*K
    const double* gp_ys;  // F*K
    // per-point scalars
    ParamSrc xsrc[kMaxD];
    ParamSrc dl, ts, zsrc;
    int zmode, nz;
    const double* zd;
    const double* zz;
    // observations grouped by observed filter
    const int* g_off;     // G+1
    const int* g_nh;      // G    number of model filters averaged (1 = direct)
    const int* g_h;       // G*3  their indices
    const double* g_lim;  // G    detection limit
    const int* o_g;       // nobs observed-filter index of each observation
    const double* o_t;
    const double* o_m;
    const double* o_s;    // sigma_obs (+inf: upper limit)
    const double* o_sig;  // budget mode: sqrt(sigma_obs^2 + budget^2)
    const double* o_lsc;  // budget mode: log(o_sig) + log(2 pi)/2
    const double* o_pack; // nobs*6: [t, mag, sigma_obs, o_sig, 1/o_sig, o_lsc] staged into shared memory by the fused kernel
    // systematics per observed filter
    const int* sy_mode;
    const double* sy_budget;
    const int* sy_nn;
    const int* sy_off;
    const ParamSrc* sy_src;
    const double* sy_t;
    // fused-kernel schedule: observed filters that map directly onto model filter f
    const int* f_goff;    // F+1
    const int* f_glist;
    // Constraint priors (nmma/core/base.py:67-68): the point fails unless lo < value < hi for every entry
    int ncon;
    ParamSrc con_src[kMaxCon];
    double con_lo[kMaxCon], con_hi[kMaxCon];
    // extinction (nmma/em/model.py:323-350): 0 none, 1 P92 SMC in the host frame, 2 per-filter coefficient x Ebv
    int ext_law;
    ParamSrc ebv;
    const double* ext_nu;    // F  observer-frame frequency nu_0 = c / wave_eff of each model filter [Hz]; 0 = no entry
    const double* ext_coef;  // F  law 2: A_f / E(B-V) (e.g. the G23 Milky-Way curve at the observer-frame wavelength)
};

struct PointScal {
    double z1, ts, dm, zc, ebv;
    float ga, gb, dmz;  // FAST back end: index guess = fma((float)t, ga, gb) on a uniform grid; dm + zc in fp32
    bool bad;
};

// FAST back end scalars: detector-frame node j sits at (samp[0] + j ds) z1 + ts  ->  j(t) = (t - ts) / (z1 ds) - samp[0] / ds
__device__ __forceinline__ void point_fast_fields(const DevCfg& cfg, PointScal& ps) {
    const double inv = cfg.uni_inv_ds / ps.z1;
    ps.ga = (float)inv;
    ps.gb = (float)(-(ps.ts * inv + cfg.uni_s0 * cfg.uni_inv_ds));
    ps.dmz = (float)(ps.dm + ps.zc);
}

// em_parameter_setup (nmma/em/model.py:288-303) + redshift_from_dlum (:259-263) +
// the scalars of gen_detector_lc / combine_detector_data (:374,393).
__device__ __forceinline__ PointScal point_setup(const DevCfg& cfg, const double* __restrict__ row) {
    PointScal ps;
    const double dl = eval_src(cfg.dl, row);
    ps.ts = eval_src(cfg.ts, row);
    double z = 0.0;
    if (cfg.zmode == 1) z = eval_src(cfg.zsrc, row);
    else if (cfg.zmode == 2) z = np_interp(dl, cfg.zd, cfg.zz, cfg.nz, cfg.zz[0], cfg.zz[cfg.nz - 1]);
    ps.z1 = 1.0 + z;
    ps.dm = 5.0 * (5 + log10(dl));
    ps.zc = -2.5 * log10(ps.z1);
    ps.bad = !(isfinite(ps.z1) && isfinite(ps.ts));
    ps.ebv = 0.0;
#ifndef TCV_NO_EXT_CON   // timing experiment: round-1 point_setup (no extinction, no constraints)
    if (cfg.ext_law != 0) {
        ps.ebv = eval_src(cfg.ebv, row);
        ps.bad = ps.bad || !isfinite(ps.ebv);   // NaN magnitudes in every filter -> sanity_check fails
    }
    // evaluate_constraints (nmma/core/base.py:67-68): Constraint.prob = (val > minimum) & (val < maximum)
    for (int i = 0; i < cfg.ncon; ++i) {
        const double v = eval_src(cfg.con_src[i], row);
        ps.bad = ps.bad || !(v > cfg.con_lo[i] && v < cfg.con_hi[i]);
    }
#endif
    point_fast_fields(cfg, ps);
    return ps;
}

// get_extinction_mags for model filter f (nmma/em/model.py:323-342), added to the absolute magnitudes before the
// distance modulus (apply_extinction_correction, :344-350).
//   law 1, extinctionFactorP92SMC (nmma/em/utils.py:373-433): Pei (1992) SMC curve xi(lambda) = sum_i a_i /
//   ((lambda/lambda_i)^n_i + (lambda_i/lambda)^n_i + b_i) with the six terms of his Table 4 (inline at :398-423),
//   amplitudes referred to A_V through A_B/A_V = 1/3.08 + 1 (dust_extinction P92.AbAv), evaluated at the HOST-frame
//   wavelength c / (nu_0 (1+z)) where nu_host lies in [c * 10 cm^-1, 2e16 Hz]; A_V = 2.93 Ebv;
//   ext_mag = -2.5 log10(10^(-0.4 xi A_V)).
//   law 2, extinctionFactorG23MW (:436-466): observer-frame, redshift-independent: ext_mag = coef_f * Ebv with
//   coef_f = R_V A(lambda_f)/A(V) staged by the host.
__device__ __forceinline__ double p92_term(double lam, double amp, double cen, double b, bool quartic) {
    const double l = lam / cen;
    double p = l * l, q = 1.0 / p;   // np.power(l, n), np.power(l, -n) for n = 2.0
    if (quartic) { p = p * p; q = 1.0 / p; }
    return amp / (p + q + b);
}
__device__ __forceinline__ double ext_mag(const DevCfg& cfg, int f, const PointScal& ps) {
    if (cfg.ext_law == 0 || ps.ebv == 0.0) return 0.0;
    if (cfg.ext_law == 2) return -2.5 * log10(pow(10.0, -0.4 * cfg.ext_coef[f] * ps.ebv));
    const double nu = cfg.ext_nu[f];
    if (!(nu > 0.0)) return 0.0;                         // filter without a wavelength entry: left uncorrected
    const double c_cgs = 29979245800.0;
    const double nu_host = nu * ps.z1;
    if (!(nu_host >= 1e-3 * 1e4 * c_cgs && nu_host <= 2e16)) return -2.5 * log10(1.0);
    const double lam = 1.0 / (1.0 / ((c_cgs / nu_host) * 1e4));   // cm -> micron -> 1/micron -> micron, as dust_extinction does
    const double abav = 1.0 / 3.08 + 1.0;
    const double ax = p92_term(lam, 185.0 * abav, 0.042, 90.0, false) + p92_term(lam, 27 * abav, 0.08, 5.5, true) +
                      p92_term(lam, 0.005 * abav, 0.22, -1.95, false) + p92_term(lam, 0.010 * abav, 9.7, -1.95, false) +
                      p92_term(lam, 0.012 * abav, 18.0, -1.80, false) + p92_term(lam, 0.030 * abav, 25.0, 0.0, false);
    const double av = 2.93 * ps.ebv;
    return -2.5 * log10(pow(10.0, -0.4 * ax * av));
}

// x' = (x - param_mins) / (param_maxs - param_mins), lightcurve_generation.py:193-194.
__device__ __forceinline__ double scaled_input(const DevCfg& cfg, int f, int i, const double* __restrict__ row) {
    const double x = eval_src(cfg.xsrc[i], row);
    return __ddiv_rn(__dsub_rn(x, cfg.pmin[f * cfg.d + i]), cfg.pden[f * cfg.d + i]);
}

// svd_back[j] = dot(VA[j, :K], c) * (maxs - mins)[j] + mins[j], lightcurve_generation.py:214-216.
template <typename CT>
__device__ __forceinline__ double node_mag(const double* __restrict__ bp, int K, int j, const CT* c) {
    const double* r = bp + (size_t)j * (K + 2);
    double acc = 0.0;
    for (int i = 0; i < K; ++i) acc = fma(r[i], (double)c[i], acc);
    return __dadd_rn(__dmul_rn(acc, r[K]), r[K + 1]);
}
template <int K, typename CT>
__device__ __forceinline__ double node_mag_k(const double* __restrict__ bp, int j, const CT* c) {
    const double* r = bp + (size_t)j * (K + 2);
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) acc = fma(r[i], (double)c[i], acc);
    return __dadd_rn(__dmul_rn(acc, r[K]), r[K + 1]);
}

// Stage 1, calc_svd_lc -> autocomplete_data(sample_times, tt, mag_back, inf)
// (lightcurve_generation.py:177): absolute mag at a sample node inside the training range.
template <typename NodeFn>
__device__ __forceinline__ double sample_mag(const DevCfg& cfg, int f, int s, NodeFn node) {
    if (cfg.single_stage) return node(s);
    const int idx = f * cfg.S + s;
    const int j = cfg.s1_j[idx];
    const double dt = cfg.s1_dt[idx];
    const double m0 = node(j);
    if (dt == 0.0) return m0;
    const double m1 = node(j + 1);
    const double slope = __ddiv_rn(__dsub_rn(m1, m0), dt);
    return __dadd_rn(__dmul_rn(slope, cfg.s1_dx[idx]), m0);
}

// observable_times[s] = sample_times[s] * (1 + z) + timeshift, em/model.py:374 (two roundings).
__device__ __forceinline__ double tobs_at(const DevCfg& cfg, int s, double z1, double ts) {
    return __dadd_rn(__dmul_rn(cfg.samp[s], z1), ts);
}

// Interval search of np.interp on the per-point detector-frame grid restricted to
// [lo, hi]: last j with tobs[j] <= t.  Caller guarantees tobs[lo] <= t <= tobs[hi].
__device__ __forceinline__ int locate(const DevCfg& cfg, int lo, int hi, double t, double z1, double ts) {
    if (cfg.uniform) {
        // guess from the inverse map, then settle with the exact rounded comparisons
        double g = ((t - ts) / z1 - cfg.uni_s0) * cfg.uni_inv_ds;
        int j = (g >= (double)hi) ? hi : ((g <= (double)lo) ? lo : (int)g);
        for (int it = 0; it < 4 && j < hi && tobs_at(cfg, j + 1, z1, ts) <= t; ++it) ++j;
        for (int it = 0; it < 4 && j > lo && tobs_at(cfg, j, z1, ts) > t; ++it) --j;
        const bool ok = tobs_at(cfg, j, z1, ts) <= t && (j == hi || tobs_at(cfg, j + 1, z1, ts) > t);
        if (ok) return j;
    }
    int a = lo, b = hi;
    while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (tobs_at(cfg, mid, z1, ts) <= t) a = mid; else b = mid;
    }
    return (tobs_at(cfg, b, z1, ts) <= t) ? b : a;
}

// Stage 2, update_lightcurve_reference (em_likelihood.py:313-335): apparent magnitude of
// model filter f at observation time t.  `abs_at(s)` returns the stage-1 absolute mag.
template <typename AbsFn>
__device__ __forceinline__ double interp_obs(const DevCfg& cfg, int f, double t, const PointScal& ps, double ext, AbsFn abs_at) {
    const int lo = cfg.s_lo[f], hi = cfg.s_hi[f];
    const double tlo = tobs_at(cfg, lo, ps.z1, ps.ts), thi = tobs_at(cfg, hi, ps.z1, ps.ts);
    if (t < tlo || t > thi) return CUDART_INF;  // left = right = +inf
    const int j = locate(cfg, lo, hi, t, ps.z1, ps.ts);
    const double tj = tobs_at(cfg, j, ps.z1, ps.ts);
    // (mags + ext_mag) + distmod + redshift_correction, em/model.py:347,396
    const double aj = __dadd_rn(__dadd_rn(__dadd_rn(abs_at(j), ext), ps.dm), ps.zc);
    if (j == hi || tj == t) return aj;
    const double tj1 = tobs_at(cfg, j + 1, ps.z1, ps.ts);
    const double aj1 = __dadd_rn(__dadd_rn(__dadd_rn(abs_at(j + 1), ext), ps.dm), ps.zc);
    const double slope = __ddiv_rn(__dsub_rn(aj1, aj), __dsub_rn(tj1, tj));
    double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(t, tj)), aj);
    if (isnan(r)) {
        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(t, tj1)), aj1);
        if (isnan(r) && aj == aj1) r = aj;
    }
    return r;
}

// FilterSystematicsHandler.__call__ for one observation (systematics.py:279-291;
// 'constant' extrapolation through autocomplete_data, em/utils.py:634-639,667-670).
__device__ __forceinline__ double sys_sigma(const DevCfg& cfg, int g, double t, const double* __restrict__ row) {
    const int mode = cfg.sy_mode[g];
    if (mode == 0) return cfg.sy_budget[g];
    const int off = cfg.sy_off[g];
    if (mode == 1) return eval_src(cfg.sy_src[off], row);
    const int nn = cfg.sy_nn[g];
    double tn[kMaxSysNodes], v[kMaxSysNodes];
    int m = 0;
    for (int i = 0; i < nn; ++i) {  // finite-mask compaction
        const double vi = eval_src(cfg.sy_src[off + i], row);
        if (isfinite(vi)) { tn[m] = cfg.sy_t[off + i]; v[m] = vi; ++m; }
    }
    if (m < 2) return CUDART_INF;
    return np_interp(t, tn, v, m, v[0], v[m - 1]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace nmma
