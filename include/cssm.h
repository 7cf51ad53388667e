/*
 * cssm.h -- C ABI of libcssm_gpu.so: the B200 (sm_100a) particle-filter hot path that sits
 * behind the Scala API of jonnylaw/ComposableStateSpaceModels.
 *
 * The reference has no FFI of its own (it is 100 % Scala).  Every entry point below names the
 * reference interface it replaces; paths are relative to the reference checkout,
 * "model/X.scala" = src/main/scala/com/github/jonnylaw/model/X.scala.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns an int status: 0 = ok, <0 = error (cssm_status); the message of
 *     the last error on the calling thread is returned by cssm_last_error().
 *   - the caller owns every host buffer; the library owns all device memory and streams.
 *   - a handle is not thread-affine (cudaSetDevice on every entry); one handle must not be used
 *     from two threads at once; different handles may be used concurrently.
 *   - there is NO CPU fallback: with no usable CUDA device every compute entry point fails with
 *     CSSM_ERR_CUDA.
 *
 * Particle state layout (host side of the ABI): structure-of-arrays, x[k*N + i] = coordinate k
 * (leaves in Tree.flatten order, model/Tree.scala:49-53, components in order inside a leaf) of
 * particle i.  Injected noise uses the same [d][N] layout.
 */
#ifndef CSSM_H
#define CSSM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSSM_VERSION 100 /* 0.1.0 */
#define CSSM_MAX_DIM 32  /* total latent dimension (sum of leaf dimensions) */
#define CSSM_MAX_RANKS 8 /* GPUs one particle cloud can be sharded over (one NVSwitch domain) */
#define CSSM_SHARD_BLOB_BYTES 1024

typedef enum {
  CSSM_OK = 0,
  CSSM_ERR_INVALID = -1,     /* bad argument / malformed model descriptor                    */
  CSSM_ERR_CUDA = -2,        /* CUDA runtime error or no device (there is no CPU fallback)   */
  CSSM_ERR_NOMEM = -3,       /* device or host allocation failed                             */
  CSSM_ERR_UNSUPPORTED = -4, /* valid request that this build does not implement             */
  CSSM_ERR_STATE = -5,       /* call order violated (e.g. step before init)                  */
  CSSM_ERR_COMM = -6         /* multi-GPU communicator failure                               */
} cssm_status;

/* model/Sde.scala:69-163 */
typedef enum { CSSM_SDE_BROWNIAN = 0, CSSM_SDE_GEN_BROWNIAN = 1, CSSM_SDE_OU = 2 } cssm_sde_kind;
/* Model.f: first component (model/Model.scala:184,250,271,328,366) or seasonal (:217-225) */
typedef enum { CSSM_F_FIRST = 0, CSSM_F_SEASONAL = 1 } cssm_f_kind;
/* observation model of the LEFT-MOST model of a composition (model/Model.scala:118-132) */
typedef enum {
  CSSM_OBS_POISSON = 0,   /* model/Model.scala:266-274 */
  CSSM_OBS_NEGBIN = 1,    /* :168-196 */
  CSSM_OBS_NORMAL = 2,    /* LinearModel :241-259 and SeasonalModel :204-234 */
  CSSM_OBS_BERNOULLI = 3, /* :315-337 */
  CSSM_OBS_LGCP = 4,      /* :363-369 via FilterLgcp, model/ParticleFilter.scala:169-227 */
  CSSM_OBS_STUDENT_T = 5, /* StudentsTModel :144-162 (df in obs_df; the log-density is multiplied by 1/v as written there) */
  CSSM_OBS_ZIP = 6,       /* ZeroInflatedPoisson :281-309 (scale = logit of the extra-zero probability) */
  CSSM_OBS_BETA = 7       /* BetaModel :339-353: Beta(exp(-gamma), 1.0).logPdf(y) -- the scale is not used by the likelihood */
} cssm_obs_kind;
/* exact transition (the stepFunction overrides) or Euler-Maruyama (model/Sde.scala:23-43) */
typedef enum { CSSM_STEP_EXACT = 0, CSSM_STEP_EULER = 1 } cssm_step_mode;
/* model/Resampling.scala:63-96 */
typedef enum { CSSM_RESAMPLE_SYSTEMATIC = 0, CSSM_RESAMPLE_STRATIFIED = 1, CSSM_RESAMPLE_MULTINOMIAL = 2 } cssm_resample_kind;
typedef enum { CSSM_F32 = 0, CSSM_F64 = 1 } cssm_dtype;
/* which particle an output takes when cumulative weights repeat (a weight too small to change the running sum):
 *   CSSM_TIE_REFERENCE  the reference: `tree ++ (cumulative sums zip items)` keeps the LAST particle inserted under a
 *                       repeated key (model/Resampling.scala:52-58), so the offspring of a particle go to the last particle
 *                       of the run of vanishing weights that follows it.  Default; ancestors bit-exact with the reference.
 *   CSSM_TIE_FIRST      the textbook inverse CDF: the FIRST index whose cumulative weight reaches k.  Not the reference's
 *                       result: with 2^20+ particles and degenerate weights the reference's rule hands about 1 % of the
 *                       offspring per step to particles of negligible weight and the log-likelihood estimate drifts down
 *                       like log N (measured: DESIGN.md section 2a); this rule does not. */
typedef enum { CSSM_TIE_REFERENCE = 0, CSSM_TIE_FIRST = 1 } cssm_tie_rule;

/*
 * One leaf of the composed model = one (Model, Sde) pair of the reference.
 * All arrays have length `dim` and hold EFFECTIVE values: already cyclically repeated by
 * Sde.buildParamRepeat (model/Sde.scala:177-179) and already transformed as the SDE
 * constructors do (c0, sigma -> exp; phi -> logistic; model/Sde.scala:70-73,99-102,133-137).
 * Unused arrays (phi for Brownian motion, mu for plain Brownian motion) may be NULL.
 * A seasonal leaf must have dim == 2*harmonics (model/Model.scala:217-225).
 */
typedef struct {
  int32_t sde_kind; /* cssm_sde_kind */
  int32_t dim;
  int32_t f_kind; /* cssm_f_kind */
  int32_t period; /* seasonal only */
  int32_t harmonics;
  const double* m0;
  const double* c0;
  const double* phi;
  const double* mu;
  const double* sigma;
} cssm_leaf_t;

/*
 * A parameterised composed model: what UnparamModel.run(params) returns in the reference
 * (model/Model.scala:110-136), flattened.  `scale` is the RAW ParamNode.scale of the left-most
 * leaf (log of the NegBin size, log of the Normal / Student-t scale, logit of the ZIP zero probability).
 */
typedef struct {
  int32_t n_leaves;
  const cssm_leaf_t* leaves; /* Tree.flatten order */
  int32_t obs_kind;          /* cssm_obs_kind */
  int32_t has_scale;
  double scale;
  int32_t step_mode;      /* cssm_step_mode */
  int32_t lgcp_precision; /* FilterLgcp.precision: sub-step = 10^-precision */
  int32_t obs_df;         /* StudentsTModel.df (model/Model.scala:144); 0 otherwise */
  int32_t reserved;
} cssm_model_desc_t;

typedef struct cssm_filter cssm_filter_t;

/* ------------------------------------------------------------------------------------------
 * library
 * ---------------------------------------------------------------------------------------- */
int cssm_version(void);
const char* cssm_last_error(void); /* thread-local, never NULL */
int cssm_device_count(int* n_out);

/* ------------------------------------------------------------------------------------------
 * filter life cycle
 * ---------------------------------------------------------------------------------------- */

/* Replaces constructing Filter(mod, resample) / FilterLgcp(mod, resample, precision)
 * (model/ParticleFilter.scala:169-172,233-235) for `n_particles` particles.
 * `seed`/`stream_id` key the in-register Philox4x32-7 generator (-DCSSM_PHILOX_ROUNDS=10 for the cuRAND round count); runs with the same
 * (seed, stream_id) and the same call sequence are reproducible and independent of the
 * launch geometry. */
int cssm_filter_create(const cssm_model_desc_t* model, int64_t n_particles, int resample_kind,
                       int dtype, int device, uint64_t seed, uint64_t stream_id,
                       cssm_filter_t** out);

/* ------------------------------------------------------------------------------------------
 * one filter sharded over several GPUs  (no reference analogue: a reference filter is one JVM
 * thread).  Rank r of `world` owns the global particle slots
 * [r*n_local, (r+1)*n_local).  Every rank runs the same three kernels per step as a single-GPU
 * filter; the per-step exchanges (max log-weight, sum w / sum w^2, "resampling done") are small
 * stores with a release flag into the peers' memory over NVLink, ancestor indices are scattered to
 * the rank that owns the offspring slot and parent states are gathered from the rank that owns
 * the parent, all through peer-mapped pointers.  Because every weight sum is exact fixed point,
 * log-likelihood, ESS, ancestors and states are bit-identical to the unsharded filter of
 * world*n_local particles with the same seed, for every `world`.
 *
 *   1. every rank:  cssm_filter_create_sharded(...)           (its own device)
 *   2. every rank:  cssm_filter_shard_export(f, blob)         CSSM_SHARD_BLOB_BYTES bytes
 *   3. the host all-gathers the blobs in rank order (torch.distributed, MPI, a JVM socket ...)
 *   4. every rank:  cssm_filter_shard_connect(f, blobs, world)
 *   5. every rank makes the SAME sequence of init / step / ll calls (SPMD).  Log-likelihood and
 *      ESS are the global ones on every rank; get_particles / sample_one / mean_state return the
 *      rank's own slots.  A rank whose peers do not arrive within 4 s fails with CSSM_ERR_COMM.
 * Ranks may be separate processes (CUDA IPC) or, for tests, several handles of one process (on
 * one device or several; use the cssm_group_* drivers below, which launch in lock-step).
 * Multinomial resampling and the per-time sampled states of cssm_filter_run are not available.
 * ---------------------------------------------------------------------------------------- */
int cssm_filter_create_sharded(const cssm_model_desc_t* model, int64_t n_local, int resample_kind,
                               int dtype, int device, uint64_t seed, uint64_t stream_id, int rank,
                               int world, cssm_filter_t** out);
int cssm_filter_shard_export(cssm_filter_t* f, void* blob_out);
int cssm_filter_shard_connect(cssm_filter_t* f, const void* blobs, int world);
int cssm_filter_shard_info(const cssm_filter_t* f, int32_t* rank_out, int32_t* world_out,
                           int64_t* slot0_out);

/* In-process group drivers: `shards[r]` = rank r, all connected.  One host thread launches every
 * phase of a step on every shard in lock-step (virtual ranks on one GPU, or one process driving
 * several GPUs).  z0 / z / outputs are GLOBAL arrays ([d][world*n_local] etc.), u holds one
 * uniform per global output for stratified resampling.  cssm_group_ll is llFilter for the group;
 * ms_out = slowest shard's device time. */
int cssm_group_init(cssm_filter_t* const* shards, int world, double t0);
int cssm_group_init_injected(cssm_filter_t* const* shards, int world, double t0, const double* z0);
int cssm_group_step_injected(cssm_filter_t* const* shards, int world, double t, int has_obs,
                             double y, const double* z, const double* u, double* x_prop_out,
                             double* logw_out, double* w1_out, int32_t* anc_out, double* ll_out,
                             int32_t* ess_out);
int cssm_group_get_particles(cssm_filter_t* const* shards, int world, double* x_out);
int cssm_group_ll(cssm_filter_t* const* shards, int world, const double* t, const double* y,
                  const uint8_t* has_obs, int64_t T, double* ll_out, float* ms_out);

/* New parameter values, same shapes: what PMMH does through model.run(p) per iteration
 * (examples/DetermineParameters.scala:70-72, model/PMMH.scala:71).  No reallocation. */
int cssm_filter_set_params(cssm_filter_t* f, const cssm_model_desc_t* model);

/* Re-key the generator (a fresh, independent likelihood evaluation). */
int cssm_filter_reseed(cssm_filter_t* f, uint64_t seed, uint64_t stream_id);

/* Run on a caller-provided CUDA stream (a cudaStream_t passed as void*), e.g. torch's current
 * stream; NULL restores the filter's own stream. */
int cssm_filter_set_stream(cssm_filter_t* f, void* cuda_stream);

int cssm_filter_destroy(cssm_filter_t* f);

int cssm_filter_dim(const cssm_filter_t* f, int32_t* d_out);
int cssm_filter_n_particles(const cssm_filter_t* f, int64_t* n_out);

/* ------------------------------------------------------------------------------------------
 * stepping API  (ParticleFilter.initialiseState / stepFilter)
 * ---------------------------------------------------------------------------------------- */

/* initialiseState, model/ParticleFilter.scala:105-108: draws every particle from
 * Sde.initialState (model/Sde.scala:75-80,104-108,152-156,206-209); ll = 0, ess = N. */
int cssm_filter_init(cssm_filter_t* f, double t0);

/* FilterInit.initialiseState, model/ParticleFilter.scala:257-260: all particles = x0[d]. */
int cssm_filter_init_state(cssm_filter_t* f, double t0, const double* x0);

/* stepFilter, model/ParticleFilter.scala:116-132 (FilterLgcp: :210-226).
 * has_obs == 0 is `observation = None`: propagate only, ll and ess unchanged.
 * *ll_out is the accumulated log-likelihood after the step, *ess_out the ESS. */
int cssm_filter_step(cssm_filter_t* f, double t, int has_obs, double y, double* ll_out,
                     int32_t* ess_out);

/* ------------------------------------------------------------------------------------------
 * whole-series API  (ParticleFilter.llFilter / filter)
 * ---------------------------------------------------------------------------------------- */

/* llFilter, model/ParticleFilter.scala:137-140: t0 = min(t), initialiseState, fold stepFilter
 * over the T data, return the log-likelihood.  The whole T-loop runs on the device without a
 * host round trip.  Host buffers in, one double out (this is the end-to-end entry point). */
int cssm_filter_ll(cssm_filter_t* f, const double* t, const double* y, const uint8_t* has_obs,
                   int64_t T, double* ll_out);

/* Split form of cssm_filter_ll for data that stays resident on the device between likelihood
 * evaluations (PMMH evaluates the same series at many parameter values):
 *   cssm_filter_load_series  builds and uploads the per-observation constant table,
 *   cssm_filter_ll_resident  runs init + T steps on it.  ess_out/ll_steps_out may be NULL;
 *   otherwise they receive the T per-step values (ll accumulated). */
int cssm_filter_load_series(cssm_filter_t* f, const double* t, const double* y,
                            const uint8_t* has_obs, int64_t T);
int cssm_filter_ll_resident(cssm_filter_t* f, double* ll_out, double* ll_steps_out,
                            int32_t* ess_out);
/* number of data of the series the handle holds (loaded by cssm_filter_load_series, _ll, _run or rebuilt by
 * _set_params): the length cssm_filter_ll_resident writes to ll_steps_out / ess_out.  0: none loaded. */
int cssm_filter_series_len(const cssm_filter_t* f, int64_t* T_out);

/* filter, model/ParticleFilter.scala:152-158: as llFilter but also returns, for the initial
 * state and each of the T steps, ONE particle sampled uniformly from the cloud
 * (Resampling.sampleOne, model/Resampling.scala:151-154): states_out[(T+1)][d]. */
int cssm_filter_run(cssm_filter_t* f, const double* t, const double* y, const uint8_t* has_obs,
                    int64_t T, double* ll_out, double* states_out);

/* How a whole-series call (cssm_filter_ll / _ll_resident / _run) is executed.  AUTO: a cloud small
 * enough for one resident grid (a few hundred 512-particle tiles; not LGCP, multinomial or sharded)
 * runs the whole foldLeft of llFilter (model/ParticleFilter.scala:139) inside ONE cooperative kernel
 * with grid barriers between the stages of a step -- the PMMH likelihood evaluation
 * (model/PMMH.scala:71) is latency bound, not bandwidth bound; larger clouds use three launches per
 * observation.  Both return the same bits.  THREE_LAUNCH / SINGLE_LAUNCH force one of the two
 * (SINGLE_LAUNCH fails with CSSM_ERR_UNSUPPORTED where it does not apply). */
#define CSSM_SERIES_AUTO 0
#define CSSM_SERIES_THREE_LAUNCH 1
#define CSSM_SERIES_SINGLE_LAUNCH 2
int cssm_filter_series_mode(cssm_filter_t* f, int mode);

/* device time (ms, CUDA events on the filter's stream) of the last whole-series call */
int cssm_filter_last_elapsed_ms(const cssm_filter_t* f, float* ms_out);
/* number of kernels the last whole-series / step call launched */
int cssm_filter_last_launches(const cssm_filter_t* f, int64_t* n_out);

/* Per-kernel device timing for the roofline report: with stride > 0 every stride-th stepFilter
 * brackets each of its kernels with CUDA events on the launching stream (0 switches it off and
 * clears the sums).  Classes: 0 gather+propagate+weight, 1 exact weight sums, 2 CDF scan +
 * ancestor search, 3 multinomial search, 5 the single-launch series kernel (one sample per
 * whole-series call).  ms_sum_out[8], count_out[8]. */
int cssm_filter_profile(cssm_filter_t* f, int stride);
int cssm_filter_profile_read(cssm_filter_t* f, double* ms_sum_out, int64_t* count_out);

/* ------------------------------------------------------------------------------------------
 * reading the cloud back  (PfState.particles, model/ParticleFilter.scala:32-37)
 * ---------------------------------------------------------------------------------------- */
/* resampled particles after the last step, x_out[d][N] (SoA) */
int cssm_filter_get_particles(cssm_filter_t* f, double* x_out);
/* ParticleFilter.getIntervals (model/ParticleFilter.scala:415-424) of the current cloud, on the device
 * (radix select, no sort, no N x d copy to the host):
 *   state_mean[d]              meanState (:478-480)
 *   state_lower/upper[d]       getCredibleInterval (:490-505): sorted(n - index - 1), sorted(index - 1),
 *                              index = floor(interval * n), per coordinate -- elements of the cloud, bit for bit
 *   gamma_out[2]               the order statistics ordered(n - index), ordered(index), index = floor(n * interval)
 *                              (getOrderStatistic :455-460) of gamma_i = f(x_i, t).  The caller applies Model.link:
 *                              eta_i = link(gamma_i) is what the reference orders, and every link of the reference is
 *                              monotone (increasing; decreasing for the Beta model, for which the mirrored ranks
 *                              n - 1 - index and index - 1 are returned, so that link(gamma_out[1]) and
 *                              link(gamma_out[0]) are the reference's lower and upper eta).
 * interval is 0.975 in the reference.  CSSM_ERR_INVALID where the reference would throw IndexOutOfBounds. */
int cssm_filter_intervals(cssm_filter_t* f, double t, double interval, double* state_mean,
                          double* state_lower, double* state_upper, double* gamma_out);
/* ParticleFilter.getForecast / getMeanForecast (model/ParticleFilter.scala:368-412) and one step of
 * SimulateData.forecast (model/Data.scala:186-217), on the device.  Every particle of the current (resampled)
 * cloud is advanced to time t without touching the filter -- x1 = stepFunction(t - s.t)(x).draw, gamma = f(x1, t),
 * eta = link(gamma) -- and two observations are drawn from Model.observation(gamma) (getForecast's, and the second
 * draw getMeanForecast summarises); the result is kept in a forecast cloud next to the filter's own.
 *   chain != 0     continue from the previous forecast cloud instead (Data.forecast's scan over times); the forecast
 *                  cloud is dropped by every filter step / initialisation (CSSM_ERR_STATE if there is none)
 *   state_mean[d], state_lower[d], state_upper[d]   meanState and getallCredibleIntervals of x1 (:398-399)
 *   eta_out[3]     mean(eta) and getOrderStatistic(eta, interval): mean, lower, upper (:400-401)
 *   obs_out[3]     the same for the second observation draw (:402-404)
 * All five outputs NULL: only the forecast cloud is produced (read it with cssm_filter_forecast_cloud).
 * Samplers: Poisson (CDF inversion / Hoermann's PTRS), Gamma (Marsaglia-Tsang), Student's t, Beta, Bernoulli, Normal
 * from the filter's Philox stream; the reference's Breeze RNG streams are not reproduced.  LGCP: CSSM_ERR_UNSUPPORTED
 * (`observation = ???`, model/Model.scala:364). */
int cssm_filter_forecast(cssm_filter_t* f, double t, double interval, int chain, double* state_mean,
                         double* state_lower, double* state_upper, double* eta_out, double* obs_out);
/* the last forecast cloud: x_out[d][N], gamma_out[N], eta_out[N], obs_out[N] (getForecast's draw), obs2_out[N] (the draw
 * getMeanForecast summarises); NULL to skip */
int cssm_filter_forecast_cloud(cssm_filter_t* f, double* x_out, double* gamma_out, double* eta_out,
                               double* obs_out, double* obs2_out);
/* How the scan + search kernel of large fp32 clouds (2048-particle tiles, systematic resampling; every rank of a sharded
 * filter likewise) works.
 * AUTO: on clouds of more than 1024 tiles per rank (smaller ones last as long as their slowest block, and a tile handed
 * from one path to the other costs both; CSSM_K3_FAST=1 in the environment lifts the limit) every tile is first scanned
 * in fp64 with CERTIFIED offspring counts -- a count is taken only where the error
 * bound of the fp64 prefix cannot change it, a repeated-key decision only where it cannot flip -- and a tile with one
 * undecided particle is recomputed by the exact 128-bit fixed-point path; the ancestors are bit-identical to EXACT,
 * which runs the exact path on every tile.  cssm_filter_scan_stats: tiles settled by either path since the last
 * initialisation (diagnostics). */
#define CSSM_SCAN_AUTO 0
#define CSSM_SCAN_EXACT 1
int cssm_filter_scan_mode(cssm_filter_t* f, int mode);
int cssm_filter_scan_stats(cssm_filter_t* f, int64_t* fast_tiles_out, int64_t* exact_tiles_out);
/* the rule for repeated cumulative weights in systematic / stratified resampling (cssm_tie_rule); for a sharded filter
 * set the same rule on every shard */
int cssm_filter_set_tie_rule(cssm_filter_t* f, int rule);
/* FilterInterpolate (model/ParticleFilter.scala:273-311): particles are PATHS and an observed step resamples whole
 * paths.  With path storage enabled the streaming entry points (cssm_filter_init*, cssm_filter_step, cssm_filter_step_injected)
 * keep the propagated cloud of every step and the ancestors of every resampling on the device (max_steps + 1 clouds);
 * a step beyond max_steps fails with CSSM_ERR_STATE.  max_steps == 0 frees the storage.  The whole-series calls
 * (cssm_filter_ll, _run, _ll_resident) do not record. */
int cssm_filter_paths_enable(cssm_filter_t* f, int64_t max_steps);
/* steps recorded since the last initialisation (a path has len + 1 states); -1: nothing recorded */
int cssm_filter_paths_len(const cssm_filter_t* f, int64_t* len_out);
/* the paths of the particles idx[0..n_idx) of the current (resampled) cloud -- idx NULL: particles 0..n_idx-1 --
 * by a walk through the ancestor tree: out[n_idx][len + 1][d], OLDEST state first (the reference's List is newest first) */
int cssm_filter_get_paths(cssm_filter_t* f, const int32_t* idx, int64_t n_idx, double* out);
/* Resampling.sampleOne of the current cloud, x_out[d] */
int cssm_filter_sample_one(cssm_filter_t* f, double* x_out);
/* PfState.ll / PfState.ess */
int cssm_filter_get_ll(cssm_filter_t* f, double* ll_out, int32_t* ess_out);
/* per-coordinate mean of the current cloud (ParticleFilter.meanState, :477-479), mean_out[d] */
int cssm_filter_mean_state(cssm_filter_t* f, double* mean_out);

/* ------------------------------------------------------------------------------------------
 * Resample[A]  (model/package.scala:23; model/Resampling.scala:63-96)
 * ---------------------------------------------------------------------------------------- */
/* weights w[n] (unnormalised, >= 0), uniforms u: systematic 1, stratified n, multinomial n.
 * ancestors_out[n]: the index into the input vector each output position takes. */
int cssm_resample(int kind, const double* w, int64_t n, const double* u, int64_t n_u,
                  int32_t* ancestors_out, int device);

/* ------------------------------------------------------------------------------------------
 * parity hooks: identical kernels, noise supplied by the caller instead of Philox
 * ---------------------------------------------------------------------------------------- */
/* z0[d][N] standard normals (as double; converted to the filter dtype on upload) */
int cssm_filter_init_injected(cssm_filter_t* f, double t0, const double* z0);

/* One stepFilter with injected noise.
 *   z   [n_sub][d][N] standard normals; n_sub = 1 except for LGCP (ceil(dt/10^-precision),
 *       0 when dt == 0) -- see cssm_filter_n_substeps.
 *   u   uniforms for the resampler (1 / N / N), ignored when has_obs == 0.
 * Optional outputs (NULL to skip), all as the step computed them on the device:
 *   x_prop_out[d][N]  propagated particles before resampling   (model/ParticleFilter.scala:118)
 *   logw_out[N]       log-weights                               (:123)
 *   w1_out[N]         exp(logw - max)                           (:125)
 *   anc_out[N]        ancestor indices chosen by the resampler  (:126)
 */
int cssm_filter_step_injected(cssm_filter_t* f, double t, int has_obs, double y, const double* z,
                              const double* u, double* x_prop_out, double* logw_out,
                              double* w1_out, int32_t* anc_out, double* ll_out,
                              int32_t* ess_out);
/* number of sub-steps FilterLgcp.calcWeight takes for time increment dt (:190); 1 otherwise */
int cssm_filter_n_substeps(const cssm_filter_t* f, double dt, int64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* CSSM_H */
