/*
 * nmma_b200.h -- C ABI of the B200-native kilonova likelihood engine.
 *
 * Drop-in boundary for NMMA's inner likelihood loop.  The
 * reference (nmma v1.0.1) is pure Python and has no FFI layer; each entry point
 * below names the reference interface it replaces (paths relative to the
 * reference checkout).  A maintainer binds these with ctypes;
 * nmma_b200/engine.py is that binding.
 *
 * Conventions: extern "C", plain pointers and sizes, no C++/torch types; every
 * call returns an int status (0 = NMMA_B200_OK) and never throws or aborts the
 * sampler process; the message for the last failure is available from
 * nmma_b200_last_error().  The opaque handle owns all device memory.  Host
 * arrays are borrowed for the duration of the call only.  Device pointers
 * passed in stay caller-owned and must live on the handle's device.  Work is
 * enqueued on the caller-supplied cudaStream_t (passed as void*).  A handle is
 * not thread-safe and owns per-handle scratch buffers: keep ONE stream in flight per
 * handle (calls on a second stream must wait for the first to finish); distinct
 * handles are independent.  A device-side wait that exceeds 8 s (which only a bug
 * could cause) traps: the call then fails with NMMA_B200_ERR_CUDA instead of hanging.
 *
 * Layout symbols: F model filters, d model inputs, K SVD coefficients kept
 * (n_coeff), T model time-grid nodes, S sample-grid nodes, H hidden units,
 * Ntr GP training points, P columns of a point, G observed filters, n
 * observations.  All matrices are dense row-major (C order).
 */
#ifndef NMMA_B200_H
#define NMMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nmma_b200_handle nmma_b200_t;

enum {
    NMMA_B200_OK = 0,
    NMMA_B200_ERR_ARG = 1,        /* invalid argument / inconsistent configuration */
    NMMA_B200_ERR_CUDA = 2,       /* CUDA runtime failure (message carries cudaGetErrorString) */
    NMMA_B200_ERR_STATE = 3,      /* compute requested before the configuration is complete */
    NMMA_B200_ERR_UNSUPPORTED = 4 /* configuration valid for the reference but outside this build */
};

/* Where a per-point scalar comes from: a column of points[N,P] (col >= 0) or a
 * constant (col < 0, e.g. a DeltaFunction prior or a reference default), then a
 * transform.  Transforms restate nmma/core/conversion.py:119-126
 * (observation_angle_conversion) and nmma/em/model.py:276-283 (log10 twins). */
enum {
    NMMA_B200_XF_NONE = 0,        /* value used as is                                       */
    NMMA_B200_XF_RAD2DEG = 1,     /* KNtheta = inclination_EM * 180.0 / pi                  */
    NMMA_B200_XF_LOG10 = 2,       /* log10_x = log10(x)                                     */
    NMMA_B200_XF_POW10 = 3,       /* x = 10 ** log10_x                                      */
    NMMA_B200_XF_THETAJN_DEG = 4, /* min(theta_jn, pi - theta_jn) * 180.0 / pi              */
    NMMA_B200_XF_COSTHETAJN_DEG = 5 /* theta_jn = arccos(cos_theta_jn), then as above       */
};

typedef struct {
    int32_t col;        /* column in points[N,P], or -1 for the constant `value` */
    int32_t transform;  /* NMMA_B200_XF_*                                        */
    double value;       /* used when col < 0                                     */
} nmma_b200_param_src;

/* Redshift source, nmma/em/model.py:249-267 + nmma/core/conversion.py:57-64. */
enum {
    NMMA_B200_Z_ZERO = 0,   /* no distance information: z = 0                                  */
    NMMA_B200_Z_PARAM = 1,  /* `redshift` is a sampled / fixed parameter                       */
    NMMA_B200_Z_TABLE = 2   /* z = np.interp(luminosity_distance, dist_grid, z_grid)           */
};

/* Systematic-error mode of one observed filter, nmma/em/systematics.py:51,279-291. */
enum {
    NMMA_B200_SYS_BUDGET = 0, /* from_budget: constant error budget                            */
    NMMA_B200_SYS_PARAM = 1,  /* from_param / from_single_params: one sampled value            */
    NMMA_B200_SYS_INTERP = 2  /* from_interpolated_params: time nodes, 'constant' extrapolation */
};

/* Extinction law of gen_detector_lc, nmma/em/model.py:201,323-342. */
enum {
    NMMA_B200_EXT_NONE = 0,         /* Ebv == 0 / not used                                            */
    NMMA_B200_EXT_P92_SMC_HOST = 1, /* extinctionFactorP92SMC: Pei (1992) SMC curve in the host frame */
    NMMA_B200_EXT_LINEAR = 2        /* ext_mag_f = coef_f * Ebv (extinctionFactorG23MW, observer frame) */
};

#define NMMA_B200_MAX_CONSTRAINTS 8
#define NMMA_B200_MAX_D 16
#define NMMA_B200_MAX_K 16
#define NMMA_B200_MAX_HELPERS 3
#define NMMA_B200_MAX_SYS_NODES 16

/* ---- lifetime ---------------------------------------------------------- */
int nmma_b200_create(int device, nmma_b200_t** out);
int nmma_b200_destroy(nmma_b200_t* h);
/* Message of the last failed call on `h` (h == NULL: last failed create). */
const char* nmma_b200_last_error(const nmma_b200_t* h);
int nmma_b200_version(void);

/* ---- surrogate: replaces SVDLightCurveModel.__init__/load_filt_model weight
 * loading, nmma/em/model.py:568-696; arrays are the svd_mag_model[filt] entries
 * written by nmma/em/training.py:229-263 (VA already cut to its first K columns). */
int nmma_b200_set_svd(nmma_b200_t* h, int F, int d, int K, int T,
                      const double* tt /* F*T */, const double* param_mins /* F*d */,
                      const double* param_maxs /* F*d */, const double* VA /* F*T*K */,
                      const double* mins /* F*T */, const double* maxs /* F*T */);
/* Keras Dense(H, relu) -> Dense(K_out) per filter (nmma/em/training.py:353-364);
 * kernels are (in, out) as Keras stores them.  Selects the `tensorflow`/`keras` path. */
int nmma_b200_set_mlp(nmma_b200_t* h, int H, int K_out, const float* W1 /* F*d*H */,
                      const float* b1 /* F*H */, const float* W2 /* F*H*K_out */,
                      const float* b2 /* F*K_out */);
/* scikit-learn GaussianProcessRegressor list per filter, kernel C^2 *
 * RationalQuadratic(alpha, length_scale) (nmma/em/training.py:429-453): X_train_,
 * alpha_, constant_value, alpha, length_scale, _y_train_mean, _y_train_std.
 * Selects the `sklearn_gp` path. */
int nmma_b200_set_gp(nmma_b200_t* h, int Ntr, const double* X /* Ntr*d */,
                     const double* alpha /* F*K*Ntr */, const double* c2 /* F*K */,
                     const double* rq_alpha /* F*K */, const double* rq_len /* F*K */,
                     const double* ymean /* F*K */, const double* ystd /* F*K */);
/* model_times / --em-tmin,--em-tmax,--em-tstep grid (nmma/em/utils.py:72-93,
 * nmma/em/model.py:230-232).  S == 0 or NULL: use the surrogate's own grid tt[0]. */
int nmma_b200_set_sample_grid(nmma_b200_t* h, int S, const double* sample_times);

/* ---- per-point parameters: replaces the dict plumbing of
 * LightCurveModelContainer.parameter_conversion / em_parameter_setup /
 * combine_lc_params, nmma/em/model.py:272-303,701-705. */
int nmma_b200_set_param_layout(nmma_b200_t* h, int P,
                               const nmma_b200_param_src* model_params /* d, model order */,
                               const nmma_b200_param_src* luminosity_distance /* NULL: const 1e-5 Mpc (10 pc) */,
                               const nmma_b200_param_src* timeshift /* NULL: const 0 */,
                               const nmma_b200_param_src* redshift /* read when z_mode == Z_PARAM; NULL: const 0 */,
                               int z_mode);
/* The three scalars are passed by pointer, not by value: a {int32, int32, double} struct by value is split over an
 * integer and an SSE register, and ctypes/libffi (CPython 3.12) hands every such argument the LAST struct's double --
 * a binder would silently evaluate at the wrong distance. */
/* dL -> z lookup built by check_vs_priors, nmma/em/model.py:249-267 (50 points). */
int nmma_b200_set_redshift_table(nmma_b200_t* h, int n, const double* dist_grid, const double* z_grid);

/* ---- data: replaces MultiFilterTransient.__init__ state,
 * nmma/em/em_likelihood.py:164-178,290-303.  Observations are grouped by observed
 * filter g (offsets[g]..offsets[g+1]); t is days since trigger; sigma_obs = +inf
 * marks an upper limit; helper_idx lists the model filter(s) the observed filter
 * maps to (1 = direct map, 2-3 = arithmetic mean, nmma/em/utils.py:549-584). */
int nmma_b200_set_observations(nmma_b200_t* h, int G, const int32_t* n_helpers /* G */,
                               const int32_t* helper_idx /* G*3 */, const int32_t* offsets /* G+1 */,
                               const double* t, const double* mag, const double* sigma_obs,
                               const double* det_limit /* G, +inf = none */);
/* FilterSystematicsHandler.__call__, nmma/em/systematics.py:54,279-296.  For
 * mode SYS_BUDGET `budget[g]` is used; SYS_PARAM reads node_src[node_offset[g]];
 * SYS_INTERP reads n_nodes[g] sources/times starting at node_offset[g]. */
int nmma_b200_set_systematics(nmma_b200_t* h, int G, const int32_t* mode, const double* budget,
                              const int32_t* n_nodes, const int32_t* node_offset,
                              const nmma_b200_param_src* node_src, const double* node_times);

/* Constraint priors, NMMALikelihoodMixin.evaluate_constraints (nmma/core/base.py:67-68,77-82;
 * bilby Constraint.prob = (val > minimum) & (val < maximum)): a point whose value (a column or a
 * constant, after the transform: KNtheta from inclination_EM, log10 twins) violates any of the n
 * constraints gets the sentinel.  n == 0 clears them. */
int nmma_b200_set_constraints(nmma_b200_t* h, int n, const nmma_b200_param_src* src /* n */,
                              const double* minimum /* n */, const double* maximum /* n */);
/* Extinction, LightCurveModelContainer.get_extinction_mags / apply_extinction_correction
 * (nmma/em/model.py:323-350) with extinctionFactorP92SMC / extinctionFactorG23MW
 * (nmma/em/utils.py:373-466): ext_mag_f is added to the absolute magnitudes of model filter f
 * before the distance modulus.  `ebv` is the Ebv column or constant.  law P92_SMC_HOST: nu0[F] =
 * c / wave_eff of each model filter in Hz (0 = filter unknown to get_default_filts_lambdas: left
 * uncorrected), evaluated per point at nu0 (1+z); coef ignored.  law LINEAR: coef[F] =
 * A_f / E(B-V) (R_V x the curve at the observer-frame wavelength); nu0 ignored.  law NONE: both
 * may be NULL.  Call after nmma_b200_set_svd (which resets it). */
int nmma_b200_set_extinction(nmma_b200_t* h, int law, const nmma_b200_param_src* ebv /* NULL: const 0 */,
                             const double* nu0 /* F */,
                             const double* coef /* F */);

/* ---- compute ----------------------------------------------------------- */
/* EMTransientLikelihood.log_likelihood for N points (nmma/core/base.py:77-82,178-182
 * -> nmma/em/em_likelihood.py:186-204): out[i] is log L or the reference's sentinel
 * -1.7976931348623157e308.  points/out are DEVICE pointers; asynchronous on `stream`. */
int nmma_b200_logl(nmma_b200_t* h, const double* points_dev /* N*P */, int64_t N,
                   double* out_dev /* N */, void* stream);
/* Same through HOST buffers: pinned staging, H2D, kernels, D2H, synchronised on return.
 * This is the call an unmodified one-point-at-a-time sampler ends up in.  Large batches
 * are cut into row blocks (one wave first, then option "pipeline_blocks", default 6) whose copies overlap the
 * kernels of their neighbours on separate streams.  Both pointers must be host memory. */
int nmma_b200_logl_host(nmma_b200_t* h, const double* points_host, int64_t N, double* out_host);
/* Same input side, but the result stays on the GPU in out_dev[N] (no D2H copy): the sharded
 * path gathers it with NCCL, the analogue of the reference's MPI result gather
 * (nmma/core/mpi_setup.py:651-683).  Synchronised on return. */
int nmma_b200_logl_host_to_device(nmma_b200_t* h, const double* points_host, int64_t N, double* out_dev);
/* SVDLightCurveModel.generate_lightcurve (apparent == 0, nmma/em/model.py:707-728:
 * absolute mags on the sample grid, +inf outside the training time range) or
 * gen_detector_lc (apparent == 1, nmma/em/model.py:352-404: detector-frame times and
 * apparent mags).  mags_dev is N*F*S; tobs_dev (N*S) may be NULL. */
int nmma_b200_mags(nmma_b200_t* h, const double* points_dev, int64_t N, int apparent,
                   double* mags_dev, double* tobs_dev, void* stream);
/* eval_svd_model front end only (nmma/em/lightcurve_generation.py:193-211): the
 * K projection coefficients per filter, N*F*K, for parity checks of the surrogate. */
int nmma_b200_coeffs(nmma_b200_t* h, const double* points_dev, int64_t N,
                     double* coeffs_dev, void* stream);

/* ---- priors on the device ------------------------
 * Replaces bilby's PriorDict.rescale / PriorDict.sample for the sampled columns of
 * points[N,P] (the reference builds the dict in nmma/em/prior.py:221-244 from
 * priors/*.prior and calls bilby.core.prior.PriorDict.rescale once per sampler
 * point; em_syserr* entries come from nmma/em/systematics.py:57-113).  Each column
 * is one analytic bilby prior, restated from bilby.core.prior.analytical (bilby is
 * a third-party dependency absent offline; floors only in pyproject.toml:30-53):
 *   UNIFORM      min + u (max - min)                                   par = min, max
 *   DELTA        peak                                                  par = peak
 *   SINE         arccos(cos(min) - u (cos(min) - cos(max)))            par = min, max
 *   COSINE       arcsin(u (sin(max) - sin(min)) + sin(min))            par = min, max
 *   GAUSSIAN     mu + erfinv(2u - 1) sqrt(2) sigma                     par = mu, sigma
 *   TRUNC_GAUSS  erfinv(2 u norm + erf((min-mu)/(sqrt2 sigma))) sqrt(2) sigma + mu
 *                                                                      par = mu, sigma, min, max
 *   POWERLAW     (min^(1+a) + u (max^(1+a) - min^(1+a)))^(1/(1+a)); a == -1 (LogUniform):
 *                min exp(u log(max/min))                               par = alpha, min, max
 *   TRIANGULAR   inverse CDF of the triangular density                 par = mode, min, max
 *   INTERPED     np.interp(u, cdf, grid) of a tabulated density (the Ebv prior,
 *                nmma/em/prior.py:209-216)                             table = cdf[n], grid[n]
 */
enum {
    NMMA_B200_PR_UNIFORM = 0,
    NMMA_B200_PR_DELTA = 1,
    NMMA_B200_PR_SINE = 2,
    NMMA_B200_PR_COSINE = 3,
    NMMA_B200_PR_GAUSSIAN = 4,
    NMMA_B200_PR_TRUNC_GAUSS = 5,
    NMMA_B200_PR_POWERLAW = 6,
    NMMA_B200_PR_TRIANGULAR = 7,
    NMMA_B200_PR_INTERPED = 8
};
#define NMMA_B200_MAX_P 32
/* kind[P], par[P*4]; tab_offset[P+1] indexes tab_cdf/tab_grid (equal entries for
 * non-INTERPED columns; all three may be NULL when no column is INTERPED).
 * P must equal the P of nmma_b200_set_param_layout when both are set. */
int nmma_b200_set_priors(nmma_b200_t* h, int P, const int32_t* kind, const double* par /* P*4 */,
                         const int32_t* tab_offset /* P+1 */, const double* tab_cdf, const double* tab_grid);
/* PriorDict.rescale for N points: unit_dev[N,P] in [0,1] -> points_dev[N,P] (may alias). */
int nmma_b200_prior_transform(nmma_b200_t* h, const double* unit_dev, int64_t N, double* points_dev,
                              void* stream);
/* PriorDict.sample for N points: counter-based Philox4x32-10 (Salmon et al. 2011, the
 * generator of cuRAND/torch), key = seed, counter = (global point index, column pair), so
 * point `first_index + i` is the same on any rank and for any sharding; uniform doubles
 * from 53 random bits, then the transform above.  unit_dev (N*P, optional) receives the
 * unit-cube draws. */
int nmma_b200_prior_sample(nmma_b200_t* h, uint64_t seed, int64_t first_index, int64_t N,
                           double* points_dev, double* unit_dev, void* stream);
/* Prior sweep without host traffic (BASELINE.json configs[4]): draws points
 * first_index .. first_index+N-1 on the device in L2-sized blocks and evaluates
 * nmma_b200_logl on each; out_dev[N]; points_dev (N*P) may be NULL (blocks are then
 * drawn into a scratch buffer that never leaves L2). */
int nmma_b200_logl_sweep(nmma_b200_t* h, uint64_t seed, int64_t first_index, int64_t N,
                         double* out_dev, double* points_dev, void* stream);

/* ---- knobs / introspection ------------------------------------------- */
/* keys: "path" (0 auto, 1 fused FFMA kernel, 2 two-stage front end + back end, 3 tensor-core kernel, 4 fused GP kernel,
 *       5 latency path: tensor-core kernel in coefficient mode with filters and hidden ranges over all SMs + back end),
 *       "tc_min_points" / "fused_min_points" / "gp_min_points" / "tc_front_min_points" / "latency_max_points" (thresholds of the automatic path), "max_ctas" (0 = one per SM),
 *       "points_per_thread" (FFMA kernel: 0 auto, 1, 2, 4), "no_fast_backend" (1 = generic fp64 back end),
 *       "no_filter_split" (1 = tensor-core kernel keeps one CTA per 256-point super-tile on small batches),
 *       "pipeline_blocks" (row blocks of the nmma_b200_logl_host copy/compute pipeline, 1 = serial),
 *       "zero_copy" (1 = nmma_b200_logl_host lets the kernels read / write page-locked host memory for <= 256 rows),
 *       "cuda_graphs" (1 = nmma_b200_logl_host replays a captured CUDA graph per batch size on the latency path). */
int nmma_b200_set_option(nmma_b200_t* h, const char* key, int64_t value);
/* keys: "launches" (kernels launched by this handle so far), "last_path", "sm_count", "ctas_per_sm",
 *       "fused_supported", "tc_supported", "tc_front_supported", "gp_fused_supported", "algorithmic_flop_per_eval", "tc_executed_flop_per_eval". */
int nmma_b200_get_info(nmma_b200_t* h, const char* key, int64_t* value);
/* FP32 FFMA throughput micro-benchmark used as the roofline denominator
 * (SURVEY.md 8d): runs `iters` dependent-chain FMA rounds on every SM and returns
 * the achieved FLOP/s for scalar (variant 0) or packed f32x2 (variant 1) FMAs. */
int nmma_b200_ffma_peak(nmma_b200_t* h, int variant, int iters, double* flops_per_s);
/* Dense tcgen05 kind::tf32 rate: `iters` x 2 MMAs of 128x128x8 per SM, A from TMEM, B from shared
 * memory; FLOP/s over all SMs.  (Roofline denominator of the round-1 tf32-split kernel; the current
 * kernel runs kind::f16 MMAs and is reported against the dense bf16/fp16 rate, twice this one.) */
int nmma_b200_tf32_peak(nmma_b200_t* h, int iters, double* flops_per_s);
/* fp64 FMA rate (the GP front end's roofline denominator), same construction as ffma_peak. */
int nmma_b200_dfma_peak(nmma_b200_t* h, int iters, double* flops_per_s);

/* Diagnostic: the per-observation term of chisquare_gaussianlog_from_lc_data
 * (nmma/em/em_likelihood.py:224-256: truncnorm.logpdf for finite sigma, norm.logsf for
 * upper limits) for n independent (mag, model mag, sigma_obs, sigma_sys, limit) tuples.
 * HOST arrays.  Used by the parity tests for the SciPy edge semantics. */
int nmma_b200_obs_terms(nmma_b200_t* h, int n, const double* mag, const double* model_mag,
                        const double* sigma_obs, const double* sigma_sys, const double* det_limit,
                        double* out);

#ifdef __cplusplus
}
#endif
#endif /* NMMA_B200_H */
