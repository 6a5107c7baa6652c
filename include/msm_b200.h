/* msm_b200.h -- C ABI of the B200-native MSM time-evolution path.
 *
 * This is the drop-in boundary for the `simulator` crate's hot loop of andillio/MSM
 * (multi-stream split-step Schroedinger-Poisson integrator).  The reference has no FFI seam of its own:
 * `SimulationObject::update()` (simulator/src/simulation_object.rs:475-661 static box, :669-873 expanding
 * box) calls ArrayFire free functions directly on the `pub` grid fields.  The entry points below are the
 * narrowest cut that contains the whole path; each one names the reference code it replaces.
 *
 * Two levels are exported by the same shared library (libmsm_b200.so):
 *
 *   msm_*      grid level.  Owns the device arrays of `SimulationGrid` (simulation_object.rs:42-64) for a batch
 *              of streams and runs the fused CUDA passes.  The host supplies per-step scalars.
 *   msm_sim_*  host-logic level.  A C++ mirror of `SimulationObject` / `SimulationParameters`
 *              (simulation_object.rs:67-184): adaptive time step, dump bookkeeping, time snapping, the
 *              scale-factor solver and t<->tau conversion, batched over streams.  It calls only msm_*.
 *
 * Conventions
 *   - all functions return MSM_OK (0) or a negative MSM_E_* code; nothing throws or aborts across the boundary
 *     (the reference panics on aliasing, simulation_object.rs:607-617; here the host decides).
 *   - the library owns all device memory; the caller owns every host buffer it passes.  Input buffers are only
 *     read during the call, output buffers are completely written before the call returns.  No host pointer is
 *     retained across calls (mirrors `Array::new(&data, dims)` ics.rs:726 and `array.host(&mut v)` io.rs:46-47).
 *   - grids are linear, dim-0-fastest exactly like the reference's ArrayFire arrays: NumPy [i][j][k] (C order)
 *     <-> linear i*n*n + j*n + k (pinned by utils/fft.rs:185-214 and io.rs:63-66).  Complex data is either
 *     interleaved (re,im,re,im,... = `Vec<Complex<f64>>`) or split into two planes (= the two NPY files of a dump).
 *   - all arithmetic is fp64 (`type T = f64`, simulator/src/main.rs:19).
 *   - one host thread per context; calls on one context are not re-entrant.  Multi-GPU = one context per rank.
 *   - there is NO CPU fallback: every entry point that computes fails with MSM_E_CUDA if no sm_100 device is usable.
 */
#ifndef MSM_B200_H
#define MSM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSM_OK            0
#define MSM_E_ARG        -1   /* bad argument / unsupported size                                             */
#define MSM_E_CUDA       -2   /* CUDA runtime error (text via msm_last_error)                                 */
#define MSM_E_NCCL       -3   /* NCCL error                                                                  */
#define MSM_E_ALIASING   -4   /* RuntimeError::FourierAliasing (utils/error.rs:5-27); per stream, see state  */
#define MSM_E_NAN        -5   /* RuntimeError::NanOrInf                                                      */
#define MSM_E_STATE      -6   /* call sequence violated (e.g. step before any psi was set)                   */
#define MSM_E_NOMEM      -7   /* device allocation failed                                                    */
#define MSM_E_IO         -8   /* RuntimeError::IOError (dump writer)                                         */

#define MSM_COUPLING_INDEPENDENT 0  /* reference behaviour: every stream has its own potential (main.rs:43)   */
#define MSM_COUPLING_SUMMED      1  /* north-star variant: rho = (A/S) sum_s |psi_s|^2, shared potential      */

#define MSM_SCHEME_NONE    0
#define MSM_SCHEME_POISSON 1        /* common/src/ics.rs:33-37 SamplingScheme                                 */
#define MSM_SCHEME_WIGNER  2
#define MSM_SCHEME_HUSIMI  3

typedef struct msm_ctx msm_ctx;
typedef struct msm_sim msm_sim;

/* ------------------------------------------------------------------------------------------------------------
 * grid level
 * ---------------------------------------------------------------------------------------------------------- */

/* Resolved scalars of one batch of streams.  Replaces the grid-related part of `SimulationParameters::new`
 * (simulation_object.rs:223-315); `spec_grid` (:273, utils/fft.rs:123-161) is never materialised -- k^2 is
 * computed from indices inside the kernels. */
typedef struct msm_config {
    int32_t struct_size;        /* = sizeof(msm_config); ABI guard                                           */
    int32_t dims;               /* 1, 2 or 3   (utils/grid.rs:272-278 Dimensions)                            */
    int32_t size;               /* cells per axis, power of two in [2, 1024]  (fft.rs:105 requires even)     */
    int32_t n_streams;          /* streams owned by THIS context (rank-local)                                */
    int32_t coupling;           /* MSM_COUPLING_*                                                            */
    int32_t device;             /* CUDA device ordinal                                                       */
    int32_t chunk_streams;      /* streams per launch group / scratch sizing; 0 = automatic                  */
    int32_t rank;               /* coupled mode only: rank in the communicator                               */
    int32_t nranks;             /* 1 = no communicator                                                       */
    int32_t n_streams_global;   /* coupled mode: S in rho = (A/S) sum ; 0 = n_streams                        */
    double  dx;                 /* cell size (:260-262); dk = dx in the reference (:263)                     */
    double  density_prefactor;  /* A of rho = A |psi|^2   (calculate_density :1031-1063)                     */
    double  poisson_coeff;      /* c of phi_k = c rho_k / k^2: -POIS_CONST static, -1 expanding (:1079-1086) */
    double  k2_cutoff;          /* alias check bound in [0,1] (:1262-1269)                                   */
    const void* nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks, NULL when nranks == 1          */
} msm_config;

const char* msm_version(void);
const char* msm_strerror(int code);

/* Fill `out128` with a fresh ncclUniqueId (rank 0 calls this and broadcasts the bytes out of band). */
int msm_nccl_unique_id(void* out128);

/* Allocate device state for cfg->n_streams streams.  Replaces `SimulationGrid::new` (simulation_object.rs:203-209)
 * and the device-side half of `SimulationParameters::new`. */
int msm_create(const msm_config* cfg, msm_ctx** out);
void msm_destroy(msm_ctx* ctx);
/* Text of the last error on this context (or of the last failed msm_create when ctx == NULL). */
const char* msm_last_error(const msm_ctx* ctx);
/* Bytes of device memory held by the context. */
int msm_device_bytes(const msm_ctx* ctx, uint64_t* bytes);

/* Upload the wavefunction of one stream.  Replaces `Array::new(&data, dim4)` (ics.rs:726) and, lazily at the
 * next compute call, `psi_k = forward(psi)` of the first step (simulation_object.rs:477-479). */
int msm_set_psi(msm_ctx* ctx, int32_t stream, const double* psi_interleaved /* 2*n^dims doubles */);
int msm_set_psi_planes(msm_ctx* ctx, int32_t stream, const double* re, const double* im /* n^dims each */);

/* Download psi of one stream (replaces `array.host()` + the re/im split of `complex_array_to_disk`,
 * utils/io.rs:46-47,58-60,72-74).  Either plane pointer may be NULL. */
int msm_get_psi(msm_ctx* ctx, int32_t stream, double* re, double* im);
int msm_get_psi_interleaved(msm_ctx* ctx, int32_t stream, double* out /* 2*n^dims */);
/* Dump of several streams: the inverse transform and plane split of stream i+1 run on the compute
 * stream while the planes of stream i cross PCIe on a copy stream (two staging buffers).  re[i] / im[i] receive stream
 * streams[i]; use pinned host memory for real overlap.  The reference copies synchronously (utils/io.rs:46-47). */
int msm_get_psi_many(msm_ctx* ctx, int32_t n, const int32_t* streams, double* const* re, double* const* im);
/* Asynchronous transfers: the same upload / download as msm_set_psi / msm_get_psi, but the call only enqueues work --
 * uploads on a dedicated stream (the compute stream waits for a stream's upload the next time it touches that
 * stream), downloads after everything already enqueued for that stream, through two dedicated staging buffers on
 * a copy stream -- so that PCIe traffic of some streams overlaps the step kernels of others (the reference copies
 * synchronously: `Array::new` ics.rs:726, `array.host` utils/io.rs:46-47).  EXCEPTION to the buffer convention above:
 * the host buffers are retained until msm_transfers_wait returns (pinned memory is needed for real overlap).
 * msm_chunk_streams returns the resolved streams-per-launch-group of the context. */
int msm_upload_begin(msm_ctx* ctx, int32_t stream, const double* psi_interleaved);
int msm_download_begin(msm_ctx* ctx, int32_t stream, double* re, double* im);
int msm_transfers_wait(msm_ctx* ctx);
int msm_chunk_streams(const msm_ctx* ctx, int32_t* chunk);
/* A ticket for every download enqueued so far; msm_download_wait blocks the CALLING thread (any thread -- the one
 * exception to "one host thread per context": the dump writer threads of msm_sim_dump use it) until those downloads
 * have landed in their host buffers, without touching the compute stream (the reference's dump is a synchronous
 * `array.host()`, utils/io.rs:46-47). */
int msm_download_ticket(msm_ctx* ctx, uint64_t* ticket);
int msm_download_wait(msm_ctx* ctx, uint64_t ticket);
/* Pinned host memory for asynchronous transfers (cudaMallocHost / cudaFreeHost). */
int msm_host_alloc(msm_ctx* ctx, size_t bytes, void** out);
int msm_host_free(msm_ctx* ctx, void* p);
/* psi_k as the reference holds it after `update()` (second drift applied, :574). */
int msm_get_psik_interleaved(msm_ctx* ctx, int32_t stream, double* out /* 2*n^dims */);

/* `calculate_potential()` at the current psi followed by `max_all(abs(phi))` (simulation_object.rs:497 -> :1066-1110
 * and :905 / :954).  active == NULL means all streams; otherwise active[s] != 0 selects stream s.
 * max_abs_phi has n_streams entries; entries of inactive streams are left untouched.  Blocking (the reference's
 * `max_all` is its per-step host sync, too). */
int msm_potential_max(msm_ctx* ctx, const int32_t* active, double* max_abs_phi);

/* `calculate_potential()` of one stream, downloaded (dump with output_potential, :1167-1180).  phi is real; the
 * reference's imaginary file is all zeros. */
int msm_get_potential(msm_ctx* ctx, int32_t stream, double* phi /* n^dims */);

/* One split step for every active stream (simulation_object.rs:504-581 static, :699-787 expanding):
 *     psi_k <- psi_k * exp(-i drift_coeff k^2);  psi <- F^-1 psi_k;  phi <- Phi(psi);
 *     psi   <- psi * exp(-i kick_coeff phi);     psi_k <- F psi;     psi_k <- psi_k * exp(-i drift_coeff k^2)
 * followed by `check_alias` (:1249-1293): alias_mass[s] = sum_{k^2 > k2_cutoff*k2_max} |psi_k|^2 * dx^dims.
 *   drift_coeff[s] = dt*hbar_/4 (static, :508) or dtau/4 (expanding, :701)
 *   kick_coeff[s]  = dt/hbar_   (static, :537) or (dtau/2)*(a1+a2) (expanding: the two half kicks of :726-760
 *                    use the same phi, so they are one multiply)
 * Arrays have n_streams entries; entries of inactive streams are ignored / untouched.  Blocking: alias_mass is
 * final on return (the reference's `sum_all`, :1280, is its second per-step host sync).  alias_mass may be NULL
 * (then the call only enqueues work and the masses are returned by the next msm_read_alias). */
int msm_step(msm_ctx* ctx, const int32_t* active, const double* drift_coeff, const double* kick_coeff,
             double* alias_mass);
int msm_read_alias(msm_ctx* ctx, double* alias_mass);
/* MAX of one host scalar over the ranks of the communicator (in place; a no-op without one).  The summed-density mode
 * uses it to agree on the largest alias mass: the streams share one potential, so when any stream on any rank crosses
 * alias_threshold every rank must stop at the same step (the reference panics, :607-617) -- otherwise the ranks would
 * disagree about the next collective. */
int msm_allreduce_max(msm_ctx* ctx, double* value);
int msm_synchronize(msm_ctx* ctx);

/* Stand-alone unitary FFT of `batch` host arrays (utils/fft.rs:6-98 forward / inverse, scale size^(-dims/2) in both
 * directions).  Runs the same pass kernels as the integrator; used by the parity tests of the FFT layer. */
int msm_fft(int32_t device, int32_t dims, int32_t size, int32_t inverse, int32_t batch, double* data_interleaved);

/* k^2 exactly as the kernels compute it from indices (parity with `spec_grid`, utils/fft.rs:123-161). */
int msm_spec_grid(int32_t device, int32_t dims, int32_t size, double dx, double* k2_out /* n^dims */);

/* On-device initial conditions (restates simulator/src/ics.rs).  `msm_ic_*` overwrite
 * psi of one stream; msm_sample_perturbation applies `sample_quantum_perturbation` (ics.rs:434-648) with a
 * counter-based Philox-4x32-10 keyed (seed, cell): Box-Muller normals for Wigner / Husimi (not ArrayFire's stream:
 * unpinned), Knuth / PTRS Poisson variates for the Poisson scheme (the reference's own draw is unseeded, ics.rs:497,
 * so only its distribution can be matched). */
int msm_ic_cold_gauss(msm_ctx* ctx, int32_t stream, const double* mean, const double* std);
int msm_ic_spherical_tophat(msm_ctx* ctx, int32_t stream, double axis_length, double radius, double delta,
                            double slope);
/* cold_gauss_kspace (ics.rs:282-431): Gaussian in k with uniform random phases exp(2 pi i u), then the forward transform.
 * The phases come from the same counter-based Philox as the sampler (draw slot 7); only dims == 3 is meaningful in the
 * reference (it hard-codes a 3-D phase array, ics.rs:401-406). */
int msm_ic_cold_gauss_kspace(msm_ctx* ctx, int32_t stream, const double* mean, const double* std, uint64_t phase_seed);
int msm_ic_copy(msm_ctx* ctx, int32_t dst_stream, int32_t src_stream);
/* Keep one wavefunction aside / put it back into a stream (device-to-device, enqueued on the compute stream): the
 * un-sampled initial condition that every stream of a TOML starts from before its seed is applied
 * (`new_from_params`, simulation_object.rs:404-435 builds it anew for every stream). */
int msm_ic_store(msm_ctx* ctx, int32_t stream);
int msm_ic_load(msm_ctx* ctx, int32_t stream);
int msm_sample_perturbation(msm_ctx* ctx, int32_t stream, int32_t scheme, uint64_t seed, double n_tot);

/* Ensemble statistics over streams (restates the reductions of the `synthesizer` crate,
 * synthesizer/src/lib.rs:106-342 `analyze_sims` with the field closures of synthesizer/src/main.rs:63-93):
 * SUMS over the selected local streams of   psi, |psi|^2, psi_k, |psi_k|^2   where psi_k is the UN-normalised forward
 * DFT of psi (ndrustfft `ndfft` per axis, lib.rs:206-213 -- unlike the simulator's unitary transform).  Done on the
 * device from the resident arrays, so S x dumps x 2 GiB of disk traffic disappears.  The caller divides by the global
 * stream count (after summing over ranks) and forms  Qx = sum_cells(<|psi|^2> - |<psi>|^2) * dx^dims  (main.rs:161-173).
 * msm_ensemble_get downloads one field (0 psi, 1 psi2, 2 psik, 3 psik2) as re / im planes in the host's linear layout;
 * psi2 and psik2 are real (the reference writes an all-zero imaginary file); either pointer may be NULL. */
int msm_ensemble_accumulate(msm_ctx* ctx, const int32_t* active);
/* Sum the four accumulators over the ranks (ncclAllReduce, in place; every rank ends up with the totals): the
 * multi-rank form of the `+=` into the shared accumulators of synthesizer/src/lib.rs:217-240.  Needs a context
 * created with nranks > 1 and an nccl_unique_id (any coupling mode); a no-op for nranks == 1. */
int msm_ensemble_allreduce(msm_ctx* ctx);
int msm_ensemble_get(msm_ctx* ctx, int32_t field, double* re, double* im);

/* Per-kernel timing (CUDA events around every launch on the context's stream).  msm_profile_read returns up to
 * `cap` records; names are static strings. */
typedef struct msm_profile_record {
    const char* name;           /* kernel family, e.g. "fft_pass<512,inv,drift,rho>"                          */
    uint64_t launches;
    double   ms_total;          /* summed CUDA-event time                                                     */
    double   algorithmic_bytes; /* summed algorithmic HBM bytes                         */
} msm_profile_record;
int msm_profile_enable(msm_ctx* ctx, int32_t on);
int msm_profile_read(msm_ctx* ctx, msm_profile_record* out, int32_t cap, int32_t* n_out);
/* CUDA-event stopwatch on the context's own stream (torch.cuda.Event would only see torch's stream). */
int msm_timer_start(msm_ctx* ctx);
int msm_timer_stop(msm_ctx* ctx, double* elapsed_ms);   /* records, synchronises, returns the elapsed time */
/* Number of kernel launches issued by this context so far. */
int msm_launch_count(const msm_ctx* ctx, uint64_t* launches);

/* ------------------------------------------------------------------------------------------------------------
 * host-logic level: mirror of SimulationObject (simulation_object.rs:145-184)
 * ---------------------------------------------------------------------------------------------------------- */

/* Resolved `SimulationParameters` (simulation_object.rs:67-140) as produced by `SimulationIter::next`
 * (utils/io.rs:164-245) from the TOML; all streams of one msm_sim share them (they differ only by seed/name). */
typedef struct msm_sim_params {
    int32_t struct_size;
    int32_t dims;
    int32_t size;
    int32_t n_streams;          /* rank-local streams evolved by this object                                  */
    int32_t expanding;          /* 0 = static box, 1 = cargo feature `expanding`                              */
    int32_t coupling;           /* MSM_COUPLING_*                                                             */
    int32_t device;
    int32_t chunk_streams;
    uint32_t num_data_dumps;
    int32_t has_max_dloga;
    int32_t rank, nranks, n_streams_global;
    int32_t reserved;
    double axis_length;
    double time;                /* start time (the reference only supports 0, :627)                           */
    double final_sim_time;
    double cfl;
    double total_mass;
    double particle_mass;
    double hbar_;
    double k2_cutoff;
    double alias_threshold;
    /* [cosmology] (common/src/parameters.rs:71-86); ignored unless expanding */
    double omega_matter_now, omega_radiation_now, h, z0, max_dloga;
    const void* nccl_unique_id;
} msm_sim_params;

/* Scalars of one stream after the last msm_sim_update (the fields of `SimulationParameters` that change). */
typedef struct msm_stream_state {
    double time;                /* parameters.time                                                            */
    double tau;                 /* parameters.tau (expanding)                                                 */
    double dt;                  /* dt (static) or dtau (expanding) of the last step                           */
    double potential_max;       /* max|phi| that set the last step                                            */
    double alias_mass;          /* p_mass of the last check_alias                                             */
    double scale_factor;        /* a(t) (expanding), else 1                                                   */
    uint64_t n_steps;
    uint32_t current_dumps;
    int32_t dumped;             /* 1 if the last update took the dump branch (:620 / :828)                    */
    int32_t finished;           /* !not_finished() (:1226-1228)                                               */
    int32_t aliased;            /* alias_mass > alias_threshold (the reference would have panicked)           */
} msm_stream_state;

/* Derived scalars (dx, dk, k2_max, comoving box, tau, final tau, A, c) exactly as `SimulationParameters::new`,
 * `calculate_density` and `calculate_potential` derive them. */
typedef struct msm_derived {
    double dx, dk, k2_max, n_tot, comoving_boxsize, tau0, final_sim_tau, density_prefactor, poisson_coeff;
} msm_derived;

int msm_sim_create(const msm_sim_params* p, msm_sim** out);          /* new_from_params without the IC (:404)  */
void msm_sim_destroy(msm_sim* sim);
const char* msm_sim_last_error(const msm_sim* sim);
msm_ctx* msm_sim_ctx(msm_sim* sim);                                  /* borrow the grid-level context          */
int msm_sim_derived(const msm_sim* sim, msm_derived* out);
int msm_sim_set_psi(msm_sim* sim, int32_t stream, const double* psi_interleaved);
/* One `update()` (:475 / :669) for every stream that is not finished.  Returns MSM_E_ALIASING if any stream
 * crossed alias_threshold (its state says which); the other streams have still been advanced. */
int msm_sim_update(msm_sim* sim);
/* The same for a subset: subset[s] != 0 selects stream s (n_streams entries, NULL = all).  Independent coupling only. */
int msm_sim_update_streams(msm_sim* sim, const int32_t* subset);
/* The reference's outer loop over the streams of one TOML (simulator/src/main.rs:43-85): for stream streams[i],
 * upload the initial wavefunction psi_in[i] (interleaved, 2*n^dims doubles; a fresh SimulationObject: time, dumps and
 * scale factor restart.  NULL entry / NULL array = keep the stream's current state), `while not_finished() { update() }`
 * (main.rs:65-69; at most max_updates calls per stream when max_updates > 0), then write the final psi as re / im planes
 * to re_out[i] / im_out[i] (either may be NULL).  The reference runs its streams one after another; here they advance
 * in groups, and the uploads of the next group and the downloads of the previous one overlap the step kernels of the
 * current one (pinned host memory is needed for that overlap).  Host buffers are read / written until the call
 * returns.  A stream that crosses alias_threshold stops there (the reference panics, :607-617) and the call returns
 * MSM_E_ALIASING after finishing the others.  Independent coupling only. */
int msm_sim_run_streams(msm_sim* sim, int32_t n, const int32_t* streams, const double* const* psi_in,
                        double* const* re_out, double* const* im_out, uint64_t max_updates);
/* The same loop with the initial conditions built ON THE DEVICE: every listed stream starts from the
 * wavefunction saved by msm_ic_store (the un-sampled initial condition: build it once with msm_ic_* / msm_set_psi on any
 * stream of msm_sim_ctx(sim), then msm_ic_store) and receives `sample_quantum_perturbation` (ics.rs:434-648) with
 * seeds[i] and the given MSM_SCHEME_* (n_tot = total_mass / particle_mass); seeds[i] == MSM_SEED_NONE leaves the stream
 * un-sampled (the trailing mean-field run of `SimulationIter`, utils/io.rs:214-240).  Nothing but scalars crosses PCIe on
 * the way in, so an 8-rank run is not limited by the host's upload bandwidth. */
#define MSM_SEED_NONE UINT64_MAX
int msm_sim_run_streams_seeded(msm_sim* sim, int32_t n, const int32_t* streams, int32_t scheme, const uint64_t* seeds,
                               double* const* re_out, double* const* im_out, uint64_t max_updates);
/* The stream groups msm_sim_run_streams forms for n streams with the given launch chunk: writes the group boundaries
 * (groups + 1 entries, bounds[0] = 0, last = n; at most cap entries) and returns the number of groups.  Host logic only. */
int msm_run_groups(int32_t n, int32_t chunk, int32_t* bounds, int32_t cap);
int msm_sim_not_finished(const msm_sim* sim);                        /* 1 while any stream has time < final    */
int msm_sim_state(const msm_sim* sim, int32_t stream, msm_stream_state* out);
int msm_sim_get_psi(msm_sim* sim, int32_t stream, double* re, double* im);
/* Write `sim-data/<sim_name>/psi_%05d_real|_imag` (two extension-less NPY v1 files, f64, shape (n,n|1,n|1,1)),
 * the layout of `complex_array_to_disk` (utils/io.rs:34-88, simulation_object.rs:1155-1158).  SURVEY row f-2. */
int msm_sim_dump(msm_sim* sim, int32_t stream, const char* root_dir, const char* sim_name, uint32_t dump_index);
/* The `output_potential` branch of dump() (simulation_object.rs:1167-1180): `calculate_potential()` of the stream, then
 * `potential_%05d_real|_imag` in the same layout; phi is real, so the imaginary file is all zeros, as in the reference. */
int msm_sim_dump_potential(msm_sim* sim, int32_t stream, const char* root_dir, const char* sim_name,
                           uint32_t dump_index);
/* msm_sim_dump only ENQUEUES the inverse transform, plane split and device-to-host copy (pinned staging pool, copy
 * stream) and hands the planes to writer threads; the step loop is not stalled (the reference blocks in `array.host()`,
 * utils/io.rs:46-47).  A writer that fails (RuntimeError::IOError, utils/error.rs:5-27; the reference panics in
 * `expect("write to disk failed")`) makes the next msm_sim_dump* or msm_sim_wait_io return MSM_E_IO.
 * msm_sim_wait_io joins the background NPY writers (the reference joins its I/O threads when a stream finishes, :651-655). */
int msm_sim_wait_io(msm_sim* sim);
/* Pin up to n staging buffers of the dump pipeline now (2 * n^dims doubles each; the pool is capped by
 * MSM_B200_DUMP_BUFFERS, default 4).  Pinning costs 0.5-1 s per GiB; without this call the pool grows on demand, i.e.
 * inside the step loop at the first dumps. */
int msm_sim_reserve_dump_buffers(msm_sim* sim, int32_t n);

/* host scalars, exported for the parity tests of rows a7/a8/a16 */
double msm_get_tau(double target_time, double omega_matter_now, double omega_radiation_now, double h, double z0,
                   double max_dloga, int32_t has_max_dloga);                     /* simulation_object.rs:1408-1453 */
double msm_supercomoving_boxsize(double hbar_, double omega_matter_now, double h, double z0,
                                 double axis_length);                            /* common parameters.rs:205-220   */
double msm_scale_factor_after(double t, double omega_matter_now, double omega_radiation_now, double h, double z0,
                              double max_dloga);                                 /* expanding.rs:99-105            */

#ifdef __cplusplus
}
#endif
#endif /* MSM_B200_H */
