/*
 * mcl_b200.h -- C ABI of the B200-native trapped-charge kinetics path (libmcl_b200.so).
 *
 * The reference (HarrisNH/MCLuminescence, pure Python/NumPy) has no FFI of its own; its seams are
 * Python call signatures.  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference root).  A maintainer binds them with ctypes -- see
 * INTEGRATION.md for the stub.
 *
 * Conventions
 *   - plain C types only; every pointer is either HOST or DEVICE memory as marked;
 *   - caller owns every buffer; the library keeps no state between calls except the
 *     thread-local last-error string and mcl_objective's cached scratch slab (mcl_release_scratch);
 *   - return value: 0 on success, negative MCL_ERR_* otherwise (message via mcl_last_error());
 *   - calls enqueue on the CUDA stream in args->stream (0 = legacy default stream) and do not
 *     synchronise unless stated.
 */
#ifndef MCL_B200_H
#define MCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCL_ABI_VERSION 1

/* protocols: which reference loop a replica follows */
#define MCL_PROTO_SIMULATE 0   /* src/class/simulate.py:46-92  (dt cap, `.any()` rule, Lum record) */
#define MCL_PROTO_TL_LAB   1   /* src/class/tl_trap_lab.py:75-111 (one lab row)                     */
#define MCL_PROTO_ISO_LAB  2   /* src/class/tl_trap_lab.py:135-172 (one isothermal experiment)      */
#define MCL_PROTO_TL_LEGACY 3  /* src/est_params/functions.py:270-360: the PRE-REFACTOR TL row loop that produced
                                  results/lab_sims/result_*.csv (one channel draw per step, exact caches after a fill,
                                  the oldest electron recombines, re-trapping re-adds a pair).  Boxes of <= 124 traps,
                                  native mode, no histograms; n_h0 = int(holes) + int(density * (V_b - V))            */

/* random-number modes */
#define MCL_MODE_PHILOX 0      /* native: Philox4x32-10 keyed (seed, replica), counter (slot, step); FP32 + SFU */
#define MCL_MODE_REPLAY 1      /* parity: consumes the reference's uniform draws in its order; FP64 */

/* status codes (also written per replica into args->status) */
#define MCL_OK             0
#define MCL_ERR_STEPS     -1   /* a replica needs more than max_steps records (reference: IndexError, simulate.py:64) */
#define MCL_ERR_NOHOLES   -2   /* nearest-hole search over zero holes (reference: ValueError from np.min) */
#define MCL_ERR_STREAM    -3   /* replay stream exhausted before the replica finished                 */
#define MCL_ERR_ALLOC     -4
#define MCL_ERR_NOEVENT   -5   /* TL_lab row with zero steps (reference: IndexError, tl_trap_lab.py:111) */
#define MCL_ERR_ARG       -6
#define MCL_ERR_CUDA      -7
#define MCL_ERR_CAPACITY  -8   /* replica does not fit the kernel's per-block capacity              */
#define MCL_ERR_INTERNAL  -9   /* a kernel self-check failed (test modes only)                       */

/* One leg of a temperature / dose schedule.  The reference has exactly one per replica
 * (simulate.py:40-45,53-54); several legs chain irradiation -> hold -> readout on one box. */
typedef struct mcl_segment {
    double T_start;    /* deg C: simulate T0 (simulate.py:53); TL_lab row.T_start (tl_trap_lab.py:78); ISO temp (:138) */
    double T_rate;     /* deg C / s (simulate.py:43; tl_trap_lab.py:79)                              */
    double duration;   /* s (simulate.py:40; tl_trap_lab.py:90)                                      */
    double dose_rate;  /* D in Gy/s for the filling clock (tl_trap_lab.py:53-60)                     */
    double dt_cap;     /* s, max_dt / T_rate or 1e20 (simulate.py:44-45); unused by lab protocols    */
    double A_opt;      /* 1/s optical excitation into the tunnelling state; 0 = reference physics    */
} mcl_segment;

/* Per-replica physics + geometry: the Physics record (engine.py:44-60) and the numbers
 * TLTrapSim.__init__ derives (tl_trap_lab.py:33-39).  The host computes side/n_e0/n_h0 with the
 * reference's own Python expressions so that int() truncation matches. */
typedef struct mcl_replica {
    double alpha, b, s, E_cb, E_loc_1, E_loc_2, D0, Retrap, k_b;
    double side;             /* core cube edge in metres: (holes / rho) ** (1/3)                     */
    double boundary_factor;
    int32_t N_e;             /* int(mc.N_e)                                                          */
    int32_t n_e0;            /* int(N_e * e_ratio_start)                                             */
    int32_t n_h0;            /* int(holes * boundary_factor ** 3)                                    */
    int32_t protocol;        /* MCL_PROTO_*                                                          */
    int32_t seg_begin, seg_count;
    int32_t obs_begin, obs_count;   /* ISO_lab observation times (tl_trap_lab.py:144-172)           */
} mcl_replica;

/* Fused ensemble reduction (replaces the per-replica loops of plots.py:50-74 for ensembles):
 * leg `sg` of replica r adds its events / occupancy to row `hist_group[r] + sg` of integer
 * histograms on a common axis (`hist_group[r]` = row of the replica's FIRST leg; without
 * hist_group every replica starts at row 0).  The buffers have `n_groups` rows; mcl_run rejects
 * (MCL_ERR_ARG) any replica whose rows [hist_group[r], hist_group[r] + seg_count) do not fit.
 * Integer accumulation => results do not depend on block order or on the GPU count. */
#define MCL_AXIS_TIME_LIN  0   /* bin k covers [lo + k*w, lo + (k+1)*w)                  */
#define MCL_AXIS_TIME_LOG  1   /* log10(t) linear between log10(lo) and log10(hi)        */
#define MCL_AXIS_TEMP      2   /* T_start + T_rate * t of the active segment, deg C      */
typedef struct mcl_hist_spec {
    int32_t axis;
    int32_t n_bins;
    int32_t n_groups;            /* ROWS of the three histogram buffers */
    int32_t reserved;
    double  lo, hi;
} mcl_hist_spec;

typedef struct mcl_run_args {
    /* ---- inputs, HOST memory (copied to the device inside the call) ---- */
    const mcl_replica *replicas;   int32_t n_replicas;
    const mcl_segment *segments;   int32_t n_segments;
    const double      *obs_time;   int32_t n_obs;
    int32_t  max_steps;            /* record capacity per replica (cfg `steps`, simulate.py:26)     */
    int32_t  mode;                 /* MCL_MODE_*                                                    */
    uint64_t seed;                 /* Philox key word 0/1                                           */
    uint64_t replica_id0;          /* global id of replicas[0]: stream of replica r is keyed by id0+r,
                                      so results are invariant to how replicas are sharded           */
    /* ---- replay inputs, DEVICE memory ---- */
    const double  *replay_u;       /* uniforms in the reference's draw order                        */
    const int64_t *replay_off;     /* [n_replicas+1] slice of replay_u owned by each replica        */
    /* ---- outputs, DEVICE memory, each may be NULL ---- */
    int32_t *event;                /* [R, max_steps] 1 = recombination (Lum, simulate.py:85)        */
    int32_t *n_e;                  /* [R, max_steps] electrons after the step (simulate.py:88)      */
    double  *t;                    /* [R, max_steps] time after the step (x_ax, simulate.py:64)     */
    int32_t *kind;                 /* [R, max_steps] 0 none / 1 fill / 2 recombination              */
    int32_t *e_idx;                /* [R, max_steps] replay only: electron index of the event       */
    int32_t *h_idx;                /* [R, max_steps] replay only: hole index of the event           */
    int32_t *steps_used;           /* [R]                                                           */
    int32_t *final_n_e;            /* [R]                                                           */
    int64_t *esteps;               /* [R] electron-steps: sum over steps of n_e before the event    */
    int64_t *consumed;             /* [R] replay only: uniforms consumed                            */
    int32_t *status;               /* [R] MCL_OK or MCL_ERR_*                                       */
    int32_t *obs_n_e;              /* [n_obs] ISO_lab: n_e at each observation crossing             */
    /* ---- fused ensemble histograms, DEVICE memory, optional ---- */
    const mcl_hist_spec *hist;     /* HOST; NULL = no histogram                                     */
    const int32_t *hist_group;     /* HOST [R] row of each replica's first leg; NULL = row 0        */
    int64_t *hist_events;          /* DEVICE [n_groups, n_bins] recombinations per bin (ADDED to)   */
    int64_t *hist_occ;             /* DEVICE [n_groups, n_bins] sum of n_e at each bin's left edge  */
    int64_t *hist_occ_sq;          /* DEVICE [n_groups, n_bins] sum of n_e^2 (for the spread)       */
    /* ---- scratch, DEVICE memory ---- */
    void    *workspace;            /* at least mcl_workspace_bytes(args) bytes, 256-byte aligned    */
    size_t   workspace_bytes;
    void    *stream;               /* cudaStream_t                                                  */
} mcl_run_args;

/* Replaces the per-replica body of simulate() (src/class/simulate.py:46-92) and the per-row /
 * per-experiment bodies of TLTrapSim.TL_lab / ISO_lab (src/class/tl_trap_lab.py:75-111,135-172),
 * i.e. Box.seed/_rebuild/add_electron/remove_pair (engine.py:113-175), Physics.lifetime
 * (engine.py:65-77), _update_lifetimes/_filling_time (tl_trap_lab.py:48-60), for a whole batch
 * of independent replicas in one launch. */
int mcl_run(const mcl_run_args *args);

/* Scratch bytes mcl_run needs for these args (uses replicas, n_replicas, max_steps, mode). */
size_t mcl_workspace_bytes(const mcl_run_args *args);

/* Same as mcl_run but every output pointer is HOST memory and replay_u / replay_off are HOST
 * memory: allocates device buffers, runs, copies back, synchronises.  This is the call a
 * non-torch host (or the e2e benchmark leg) makes. */
int mcl_run_host(const mcl_run_args *args);

/* Replaces optimizer.objective mapped over a population (src/class/optimizer.py:49-84 with
 * tl_trap_lab.py:65-123,125-179): P is HOST [10, S] row-major in the reference's parameter order
 * (rho_prime, E_cb, E_loc_1, E_loc_2, D0, s, b, alpha, holes, retrap); lab rows are described by
 * `rows` (one mcl_segment per TL row or ISO experiment, plus targets); mse is HOST [S].
 * Synchronises. */
typedef struct mcl_lab {
    int32_t protocol;              /* MCL_PROTO_TL_LAB or MCL_PROTO_ISO_LAB                         */
    int32_t n_rows;                /* TL rows or ISO experiments                                    */
    const mcl_segment *rows;       /* HOST [n_rows]                                                 */
    const double *e_ratio_start;   /* HOST [n_rows]                                                 */
    const int32_t *obs_begin;      /* HOST [n_rows+1] (ISO) or NULL                                 */
    const double *obs_time;        /* HOST [n_obs] (ISO)                                            */
    const double *target;          /* HOST: TL [n_rows] Fill; ISO [n_obs] e_ratio                   */
    double N_e, boundary_factor, D, k_b;  /* fixed experiment fields (TLlab.yaml / lab_TL.yaml)    */
    int32_t max_steps;
    int32_t flags;                 /* bit 0: legacy est_params semantics (TL only; MCL_PROTO_TL_LEGACY)   */
} mcl_lab;
#define MCL_LAB_LEGACY 1
int mcl_objective(const double *P, int32_t S, const mcl_lab *lab, uint64_t seed,
                  uint64_t candidate_id0, double *mse, int64_t *esteps_total, void *stream);
/* mcl_objective keeps its (multi-GB) device scratch between calls (per device, grow-only) and is
 * therefore serialised internally: concurrent calls from several host threads are safe but run one
 * after another.  mcl_release_scratch frees the slabs of every device (waiting for a running call)
 * and leaves the caller's current device unchanged. */
void mcl_release_scratch(void);
/* Device time (ms, CUDA events around the kernel launch) of this thread's last mcl_objective call. */
float mcl_objective_last_kernel_ms(void);

/* Issue-rate microbenchmarks for the roofline denominators (SFU, FP32 FMA, INT32 multiply-add,
 * LOP3), measured on the current device.  Results in giga lane-ops per second. */
typedef struct mcl_peaks {
    double mufu_gops, ffma_gops, imad_gops, lop3_gops, sm_clock_mhz;
    int32_t n_sm, reserved;
} mcl_peaks;
int mcl_device_peaks(mcl_peaks *out);

/* Test hook: the native kernel's exponential draw for raw Philox words, evaluated on the device:
 * neg_lg2_u[i] = -lg2.approx(u01(words[i])) (the waiting time in units of ln 2 / rate) and
 * lg2_of_that[i] = lg2.approx(neg_lg2_u[i]) (what enters the log-domain argmin).  All pointers HOST.
 * tests/test_gpu_philox.py pins the small-waiting-time tail with it. */
int mcl_debug_exp_draws(const uint32_t *words, int32_t n, float *neg_lg2_u, float *lg2_of_that);

/* Kernels this library has launched in this process so far (replay, block and one-warp kernels; the roofline
 * microbenchmarks of mcl_device_peaks are not counted).  bench.py reads it around its timed region for `gpu_launches`. */
int64_t mcl_launch_count(void);

const char *mcl_last_error(void);
int mcl_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MCL_B200_H */
