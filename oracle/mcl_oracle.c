/*
 * mcl_oracle.c -- CPU restatement (plain C, float64) of MCLuminescence's trapped-charge kinetics loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * mcluminescence_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product never does (it has no CPU fallback).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against golden
 * vectors produced by the UNMODIFIED reference run in the build container
 * (oracle/ref_harness/gen_golden.py -> tests/golden/): integer event / n_e traces and the
 * structural (electron index, hole index) logs bit-exact, event times to 1e-12 relative
 * (the reference's np.exp is NumPy's AVX512 kernel here, which differs from libm exp by 1 ulp on
 * ~5 % of arguments).
 *
 * What is restated (reference file:line, relative to /root/reference):
 *   Physics.rate_cb / rate_tunnel / lifetime     src/class/engine.py:65-77
 *   Box.seed / _rebuild / nearest                src/class/engine.py:113-129,185-188
 *   Box.add_electron (stale incremental cache)   src/class/engine.py:133-152
 *   Box.remove_pair (shift-then-mask rescan)     src/class/engine.py:154-175
 *   TLTrapSim._update_lifetimes / _filling_time  src/class/tl_trap_lab.py:48-60
 *   simulate() step loop                         src/class/simulate.py:46-92
 *   TLTrapSim.TL_lab / ISO_lab loops             src/class/tl_trap_lab.py:75-111,135-172
 * The random stream is NumPy's legacy global RandomState (MT19937, 53-bit doubles;
 * exponential(scale) == scale * -log(1.0 - U)), restated here from its published algorithm
 * (Matsumoto & Nishimura init_genrand / genrand_res53).
 *
 * Unlike the reference no dense n_e x n_h matrix is kept: distances are recomputed from the
 * coordinates with the identical expression sqrt((dx*dx + dy*dy) + dz*dz), so the cached values
 * the reference would read are reproduced bit-for-bit.  Arrays are kept compact (deleted rows
 * shift down) so every index means what it means in the reference.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread -shared -fPIC).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MCLO_PROTO_SIMULATE 0
#define MCLO_PROTO_TL_LAB   1
#define MCLO_PROTO_ISO_LAB  2
#define MCLO_PROTO_TL_LEGACY 3  /* pre-refactor TL loop: src/est_params/functions.py:270-360 (see run_legacy_tl below) */

#define MCLO_OK              0
#define MCLO_ERR_STEPS      -1   /* replica needs more than max_steps records (reference: IndexError) */
#define MCLO_ERR_NOHOLES    -2   /* nearest-hole search over zero holes (reference: ValueError in np.min) */
#define MCLO_ERR_STREAM     -3   /* external uniform stream exhausted */
#define MCLO_ERR_ALLOC      -4
#define MCLO_ERR_NOEVENT    -5   /* TL_lab row finished with zero steps (reference: IndexError sim_ratio[-1]) */

/* ------------------------------------------------------------------------------------------ */
/* RNG: MT19937 exactly as numpy.random.seed(int) / random_sample                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t key[624];
    int pos;
    /* optional external stream (replay of recorded uniforms) */
    const double *ext;
    int64_t ext_n;
    int64_t consumed;
    int exhausted;
} mclo_rng;

void mclo_rng_seed(mclo_rng *r, uint32_t seed)
{
    for (int i = 0; i < 624; i++) {
        r->key[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    r->pos = 624;
    r->ext = NULL;
    r->ext_n = 0;
    r->consumed = 0;
    r->exhausted = 0;
}

void mclo_rng_external(mclo_rng *r, const double *u, int64_t n)
{
    memset(r, 0, sizeof(*r));
    r->ext = u;
    r->ext_n = n;
}

static void mt_refill(mclo_rng *r)
{
    uint32_t *mt = r->key;
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < 623; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    r->pos = 0;
}

static inline uint32_t mt_u32(mclo_rng *r)
{
    if (r->pos == 624) mt_refill(r);
    uint32_t y = r->key[r->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static inline double rng_uniform(mclo_rng *r)
{
    r->consumed++;
    if (r->ext) {
        if (r->consumed > r->ext_n) { r->exhausted = 1; return 0.5; }
        return r->ext[r->consumed - 1];
    }
    uint32_t a = mt_u32(r) >> 5, b = mt_u32(r) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* numpy legacy: exponential(scale) = scale * -log(1.0 - U) */
static inline double rng_exponential(mclo_rng *r, double scale)
{
    return scale * (-log(1.0 - rng_uniform(r)));
}

void mclo_rng_fill(mclo_rng *r, double *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = rng_uniform(r);
}

int64_t mclo_rng_consumed(const mclo_rng *r) { return r->consumed; }
size_t mclo_rng_sizeof(void) { return sizeof(mclo_rng); }

/* ------------------------------------------------------------------------------------------ */
/* Replica description                                                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double T_start;     /* deg C.  simulate: T0; TL_lab: row.T_start; ISO_lab: sub.temp           */
    double T_rate;      /* deg C / s                                                              */
    double duration;    /* s                                                                      */
    double dose_rate;   /* D, Gy/s (0 => filling clock is the 1e-20 sentinel)                     */
    double dt_cap;      /* s; 1e20 when T_rate == 0 (simulate.py:45); ignored by lab protocols    */
    double A_opt;       /* optical excitation rate, 1/s (extension, parity unpinned); 0 = off     */
} mclo_segment;

typedef struct {
    double alpha, b, s, E_cb, E_loc_1, E_loc_2, D0, Retrap, k_b;
    double side;            /* core cube edge, m: (holes/rho)^(1/3), computed by the host         */
    double boundary_factor;
    int32_t N_e;            /* trap capacity int(mc.N_e)                                          */
    int32_t n_e0;           /* int(N_e * e_ratio_start)                                           */
    int32_t n_h0;           /* int(holes * bf**3)                                                 */
    int32_t protocol;
    int32_t seg_begin, seg_count;   /* into the segment table (lab protocols: exactly one)        */
    int32_t obs_begin, obs_count;   /* ISO_lab observation times                                  */
} mclo_replica;

typedef struct {
    int n_e, n_h, cap_e, cap_h;
    double *ex, *ey, *ez, *min_d;
    int32_t *nearest;
    double *hx, *hy, *hz;
    double *wait;
} box_t;

static int box_alloc(box_t *bx, int cap_e, int cap_h)
{
    memset(bx, 0, sizeof(*bx));
    bx->cap_e = cap_e; bx->cap_h = cap_h;
    bx->ex = malloc(sizeof(double) * cap_e); bx->ey = malloc(sizeof(double) * cap_e);
    bx->ez = malloc(sizeof(double) * cap_e); bx->min_d = malloc(sizeof(double) * cap_e);
    bx->wait = malloc(sizeof(double) * cap_e);
    bx->nearest = malloc(sizeof(int32_t) * cap_e);
    bx->hx = malloc(sizeof(double) * cap_h); bx->hy = malloc(sizeof(double) * cap_h);
    bx->hz = malloc(sizeof(double) * cap_h);
    if (!bx->ex || !bx->ey || !bx->ez || !bx->min_d || !bx->wait || !bx->nearest ||
        !bx->hx || !bx->hy || !bx->hz) return MCLO_ERR_ALLOC;
    return MCLO_OK;
}

static void box_free(box_t *bx)
{
    free(bx->ex); free(bx->ey); free(bx->ez); free(bx->min_d); free(bx->wait);
    free(bx->nearest); free(bx->hx); free(bx->hy); free(bx->hz);
}

static int box_grow_holes(box_t *bx)
{
    int cap = bx->cap_h * 2 + 16;
    double *x = realloc(bx->hx, sizeof(double) * cap); if (!x) return MCLO_ERR_ALLOC; bx->hx = x;
    double *y = realloc(bx->hy, sizeof(double) * cap); if (!y) return MCLO_ERR_ALLOC; bx->hy = y;
    double *z = realloc(bx->hz, sizeof(double) * cap); if (!z) return MCLO_ERR_ALLOC; bx->hz = z;
    bx->cap_h = cap;
    return MCLO_OK;
}

static int box_grow_electrons(box_t *bx)
{
    int cap = bx->cap_e * 2 + 16;
#define GROW(p, T) { T *q = realloc(bx->p, sizeof(T) * cap); if (!q) return MCLO_ERR_ALLOC; bx->p = q; }
    GROW(ex, double) GROW(ey, double) GROW(ez, double) GROW(min_d, double) GROW(wait, double)
    GROW(nearest, int32_t)
#undef GROW
    bx->cap_e = cap;
    return MCLO_OK;
}

/* np.linalg.norm(e - h, axis=-1): sqrt(add.reduce(x*x)) => ((dx*dx + dy*dy) + dz*dz), no FMA */
static inline double dist(double ax, double ay, double az, double bx, double by, double bz)
{
    double dx = ax - bx, dy = ay - by, dz = az - bz;
    return sqrt((dx * dx + dy * dy) + dz * dz);
}

/* first-minimum scan of electron i against all current holes (np.min / np.argmin of a row) */
static int scan_nearest(const box_t *bx, double x, double y, double z, int n_h, double *dmin, int32_t *arg)
{
    if (n_h <= 0) return MCLO_ERR_NOHOLES;
    double best = dist(x, y, z, bx->hx[0], bx->hy[0], bx->hz[0]);
    int32_t bi = 0;
    for (int j = 1; j < n_h; j++) {
        double d = dist(x, y, z, bx->hx[j], bx->hy[j], bx->hz[j]);
        if (d < best) { best = d; bi = j; }
    }
    *dmin = best; *arg = bi;
    return MCLO_OK;
}

/* engine.py:124-129 then the lazy _rebuild of engine.py:113-119 */
static int box_seed(box_t *bx, const mclo_replica *rp, mclo_rng *rng)
{
    double core = rp->side, bnd = rp->side * rp->boundary_factor;
    bx->n_e = rp->n_e0; bx->n_h = rp->n_h0;
    for (int i = 0; i < bx->n_e; i++) {
        bx->ex[i] = rng_uniform(rng) * core;
        bx->ey[i] = rng_uniform(rng) * core;
        bx->ez[i] = rng_uniform(rng) * core;
    }
    for (int j = 0; j < bx->n_h; j++) {
        bx->hx[j] = rng_uniform(rng) * bnd;
        bx->hy[j] = rng_uniform(rng) * bnd;
        bx->hz[j] = rng_uniform(rng) * bnd;
    }
    for (int i = 0; i < bx->n_e; i++) {
        int rc = scan_nearest(bx, bx->ex[i], bx->ey[i], bx->ez[i], bx->n_h, &bx->min_d[i], &bx->nearest[i]);
        if (rc) return rc;
    }
    return MCLO_OK;
}

/* engine.py:133-152: the new electron sees the OLD holes only; nobody sees the new hole */
static int box_add_electron(box_t *bx, const mclo_replica *rp, mclo_rng *rng)
{
    double core = rp->side, bnd = rp->side * rp->boundary_factor;
    if (bx->n_e + 1 > bx->cap_e) { int rc = box_grow_electrons(bx); if (rc) return rc; }
    if (bx->n_h + 1 > bx->cap_h) { int rc = box_grow_holes(bx); if (rc) return rc; }
    double x = rng_uniform(rng) * core, y = rng_uniform(rng) * core, z = rng_uniform(rng) * core;
    double hx = rng_uniform(rng) * bnd, hy = rng_uniform(rng) * bnd, hz = rng_uniform(rng) * bnd;
    int e = bx->n_e;
    bx->ex[e] = x; bx->ey[e] = y; bx->ez[e] = z;
    int rc = scan_nearest(bx, x, y, z, bx->n_h, &bx->min_d[e], &bx->nearest[e]);
    if (rc) return rc;
    bx->hx[bx->n_h] = hx; bx->hy[bx->n_h] = hy; bx->hz[bx->n_h] = hz;
    bx->n_e++; bx->n_h++;
    return MCLO_OK;
}

/* engine.py:154-175 */
static int box_remove_pair(box_t *bx, int e, int h)
{
    int ne = bx->n_e, nh = bx->n_h;
    size_t te = (size_t)(ne - 1 - e);
    memmove(bx->ex + e, bx->ex + e + 1, te * sizeof(double));
    memmove(bx->ey + e, bx->ey + e + 1, te * sizeof(double));
    memmove(bx->ez + e, bx->ez + e + 1, te * sizeof(double));
    memmove(bx->min_d + e, bx->min_d + e + 1, te * sizeof(double));
    memmove(bx->nearest + e, bx->nearest + e + 1, te * sizeof(int32_t));
    size_t th = (size_t)(nh - 1 - h);
    memmove(bx->hx + h, bx->hx + h + 1, th * sizeof(double));
    memmove(bx->hy + h, bx->hy + h + 1, th * sizeof(double));
    memmove(bx->hz + h, bx->hz + h + 1, th * sizeof(double));
    bx->n_e = --ne; bx->n_h = --nh;
    /* shift first (:168), THEN take the mask (:171): catches old index h and old index h+1 */
    for (int i = 0; i < ne; i++) if (bx->nearest[i] > h) bx->nearest[i]--;
    for (int i = 0; i < ne; i++) {
        if (bx->nearest[i] == h) {
            int rc = scan_nearest(bx, bx->ex[i], bx->ey[i], bx->ez[i], nh, &bx->min_d[i], &bx->nearest[i]);
            if (rc) return rc;
        }
    }
    return MCLO_OK;
}

/* tl_trap_lab.py:48-51 with engine.py:65-77.  Draw order: n_e selectors, then n_e exponentials. */
static void update_lifetimes(box_t *bx, const mclo_replica *rp, double T, double A_opt, mclo_rng *rng)
{
    int n = bx->n_e;
    double kT = rp->k_b * T;
    double k_cb = rp->s * exp(-rp->E_cb / kT);
    /* selectors first (np.random.rand(*d.shape) inside rate_tunnel), lifetimes kept in wait[] */
    for (int i = 0; i < n; i++) {
        double u = rng_uniform(rng);
        double E_loc = (u < rp->Retrap) ? rp->E_loc_2 : rp->E_loc_1;
        double k_tun;
        if (A_opt == 0.0)
            k_tun = rp->b * exp(-E_loc / kT - rp->alpha * bx->min_d[i]);
        else   /* extension: optical excitation into the tunnelling state, SURVEY 8f-1 */
            k_tun = (A_opt + rp->b * exp(-E_loc / kT)) * exp(-rp->alpha * bx->min_d[i]);
        bx->wait[i] = 1.0 / (k_cb + k_tun);
    }
    for (int i = 0; i < n; i++) bx->wait[i] = rng_exponential(rng, bx->wait[i]);
}

/* tl_trap_lab.py:53-60 -- always consumes one uniform */
static double filling_time(const box_t *bx, const mclo_replica *rp, double D, mclo_rng *rng)
{
    double lam;
    if (bx->n_e == rp->N_e || D == 0.0) lam = 1e-20;
    else lam = (D / rp->D0) * (double)(rp->N_e - bx->n_e);
    if (lam > 0) return rng_exponential(rng, 1.0 / lam);
    return 1e20;
}

static inline void wait_min(const box_t *bx, double *mn, int *arg, int *any)
{
    double best = 0; int bi = -1, a = 0;
    for (int i = 0; i < bx->n_e; i++) {
        double w = bx->wait[i];
        if (w != 0.0) a = 1;                       /* ndarray.any(): non-zero (inf, nan count) */
        if (bi < 0 || w < best) { best = w; bi = i; }   /* np.min/argmin: first minimum */
    }
    *mn = best; *arg = bi; *any = a;
}

/* ------------------------------------------------------------------------------------------ */
/* Legacy (pre-refactor) semantics: src/est_params/functions.py, the code that produced         */
/* results/lab_sims/result_*.csv.  Pinned by tests/golden/legacy_*.npz, generated from the      */
/* UNMODIFIED legacy function sim_lab_TL_residuals (oracle/ref_harness/gen_golden.py).          */
/* ------------------------------------------------------------------------------------------ */
/* functions.py:138-148 -- lifetime D0/D*1/(N-e), 1e20 when full or D == 0; always one uniform */
static double legacy_filling_time(const box_t *bx, const mclo_replica *rp, double D, mclo_rng *rng)
{
    double lifetime;
    if (bx->n_e == rp->N_e || D == 0.0) lifetime = 1e20;
    else lifetime = rp->D0 / D * 1 / (double)(rp->N_e - bx->n_e);
    return rng_exponential(rng, lifetime);
}

/* functions.py:115-136 -- ONE channel draw per call (`rand(1) > Retrap` -> E_loc_1), then one exponential per electron */
static void legacy_lifetimes(box_t *bx, const mclo_replica *rp, double T, mclo_rng *rng)
{
    double kT = rp->k_b * T;
    double k_cb = rp->s * exp(-rp->E_cb / kT);
    double u = rng_uniform(rng);
    double E_loc = (u > rp->Retrap) ? rp->E_loc_1 : rp->E_loc_2;
    for (int i = 0; i < bx->n_e; i++)
        bx->wait[i] = 1.0 / (k_cb + rp->b * exp(-E_loc / kT - rp->alpha * bx->min_d[i]));
    for (int i = 0; i < bx->n_e; i++) bx->wait[i] = rng_exponential(rng, bx->wait[i]);
}

/* functions.py:214-228 + 150-161 + 104-113: append a pair, then EVERY electron's nearest hole is exact again
 * (the new electron sees its twin hole too; np.argmin keeps the old hole on an exact tie) */
static int legacy_add_electron(box_t *bx, const mclo_replica *rp, mclo_rng *rng)
{
    double core = rp->side, bnd = rp->side * rp->boundary_factor;
    if (bx->n_e + 1 > bx->cap_e) { int rc = box_grow_electrons(bx); if (rc) return rc; }
    if (bx->n_h + 1 > bx->cap_h) { int rc = box_grow_holes(bx); if (rc) return rc; }
    double x = rng_uniform(rng) * core, y = rng_uniform(rng) * core, z = rng_uniform(rng) * core;
    double hx = rng_uniform(rng) * bnd, hy = rng_uniform(rng) * bnd, hz = rng_uniform(rng) * bnd;
    int e = bx->n_e, h = bx->n_h;
    bx->ex[e] = x; bx->ey[e] = y; bx->ez[e] = z;
    bx->hx[h] = hx; bx->hy[h] = hy; bx->hz[h] = hz;
    bx->n_e++; bx->n_h++;
    for (int i = 0; i < e; i++) {
        double d = dist(bx->ex[i], bx->ey[i], bx->ez[i], hx, hy, hz);
        if (d < bx->min_d[i]) { bx->min_d[i] = d; bx->nearest[i] = h; }
    }
    return scan_nearest(bx, x, y, z, bx->n_h, &bx->min_d[e], &bx->nearest[e]);
}

/* functions.py:241-261 with :176-199,201-212.  `np.where(time >= recombination + e_timer)[0]` broadcasts against the
 * default e_timer of shape (1,1), so the ROW indices it returns are all zero: whatever the waiting times say, the
 * electron that recombines is electron 0 (the oldest), with its cached hole.  Electrons that shared the hole re-scan. */
static int legacy_remove_first(box_t *bx)
{
    int ne = bx->n_e, nh = bx->n_h;
    int h = bx->nearest[0];
    memmove(bx->ex, bx->ex + 1, (size_t)(ne - 1) * sizeof(double));
    memmove(bx->ey, bx->ey + 1, (size_t)(ne - 1) * sizeof(double));
    memmove(bx->ez, bx->ez + 1, (size_t)(ne - 1) * sizeof(double));
    memmove(bx->min_d, bx->min_d + 1, (size_t)(ne - 1) * sizeof(double));
    memmove(bx->nearest, bx->nearest + 1, (size_t)(ne - 1) * sizeof(int32_t));
    size_t th = (size_t)(nh - 1 - h);
    memmove(bx->hx + h, bx->hx + h + 1, th * sizeof(double));
    memmove(bx->hy + h, bx->hy + h + 1, th * sizeof(double));
    memmove(bx->hz + h, bx->hz + h + 1, th * sizeof(double));
    bx->n_e = --ne; bx->n_h = --nh;
    /* the electrons to refresh were selected BEFORE the shift (electrons_new_distances): exactly those cached on h */
    for (int i = 0; i < ne; i++) {
        if (bx->nearest[i] == h) {
            int rc = scan_nearest(bx, bx->ex[i], bx->ey[i], bx->ez[i], nh, &bx->min_d[i], &bx->nearest[i]);
            if (rc) return rc;
        } else if (bx->nearest[i] > h) {
            bx->nearest[i]--;
        }
    }
    return MCLO_OK;
}

/* Per-replica outputs (all caller-allocated, row r at offset r*max_steps) */
typedef struct {
    int32_t *event;      /* [R,max_steps] 1 = recombination (Lum), 0 otherwise              */
    int32_t *n_e;        /* [R,max_steps] electrons after the step                          */
    double  *t;          /* [R,max_steps] time after the step (x_ax)                        */
    int32_t *kind;       /* [R,max_steps] optional: 0 none, 1 fill, 2 recombination         */
    int32_t *e_idx;      /* [R,max_steps] optional: electron index of the event (fill: n_e before) */
    int32_t *h_idx;      /* [R,max_steps] optional: hole index of the event (fill: n_h before)     */
    int32_t *steps_used; /* [R]                                                             */
    int32_t *final_n_e;  /* [R]                                                             */
    int64_t *consumed;   /* [R] uniforms consumed by this replica                           */
    int64_t *esteps;     /* [R] sum over steps of n_e before the event (electron-steps)     */
    int32_t *obs_n_e;    /* [sum obs_count] ISO_lab: n_e at each observation crossing       */
    double  *obs_t;      /* [sum obs_count] optional: t_cur at the crossing                 */
    int32_t *status;     /* [R]                                                             */
} mclo_out;

#define REC(o, r, i, ev, kd, ei, hi, ne_, tt) do {                                   \
        size_t _p = (size_t)(r) * (size_t)max_steps + (size_t)(i);                     \
        if ((o)->event) (o)->event[_p] = (ev);                                         \
        if ((o)->n_e) (o)->n_e[_p] = (ne_);                                            \
        if ((o)->t) (o)->t[_p] = (tt);                                                 \
        if ((o)->kind) (o)->kind[_p] = (kd);                                           \
        if ((o)->e_idx) (o)->e_idx[_p] = (ei);                                         \
        if ((o)->h_idx) (o)->h_idx[_p] = (hi);                                         \
    } while (0)

static int run_one(const mclo_replica *rp, const mclo_segment *segs, const double *obs_time,
                   int r, int max_steps, mclo_rng *rng, mclo_out *out)
{
    box_t bx;
    int rc = box_alloc(&bx, (rp->N_e > rp->n_e0 ? rp->N_e : rp->n_e0) + 8, rp->n_h0 + rp->N_e + 64);
    if (rc) return rc;
    int64_t c0 = rng->consumed, esteps = 0;
    int i = 0;              /* record index */
    rc = box_seed(&bx, rp, rng);
    if (rc) goto done;

    if (rp->protocol == MCLO_PROTO_TL_LEGACY) {
        /* src/est_params/functions.py:289-349 (sim_lab_TL_residuals, one lab row, one replica) */
        const mclo_segment *S = &segs[rp->seg_begin];
        double D = S->dose_rate, T = S->T_start + 273.15, t_cur = 0.0;
        double dt_filling = legacy_filling_time(&bx, rp, D, rng);
        if (bx.n_e > 0 && bx.n_h > 0) legacy_lifetimes(&bx, rp, T, rng);      /* `if distances.size != 0` */
        while (t_cur < S->duration) {
            double wmin; int arg, any;
            wait_min(&bx, &wmin, &arg, &any);
            double dt_recomb = bx.n_e > 0 ? wmin : dt_filling;
            double dt = dt_recomb < dt_filling ? dt_recomb : dt_filling;
            if (i >= max_steps) { rc = MCLO_ERR_STEPS; break; }
            esteps += bx.n_e;
            int ev = 0, kd, ei, hi;
            if (dt == dt_filling) {
                if (dt < 0) dt = 0;
                T = T + dt * S->T_rate;
                t_cur = t_cur + dt;
                ei = bx.n_e; hi = bx.n_h; kd = 1;
                rc = legacy_add_electron(&bx, rp, rng);
                if (rc) break;
                legacy_lifetimes(&bx, rp, T, rng);
                dt_filling = legacy_filling_time(&bx, rp, D, rng);
            } else {
                if (dt < 0) dt = 0;
                T = T + dt * S->T_rate;
                t_cur = t_cur + dt;
                ei = 0; hi = bx.nearest[0]; kd = 2; ev = 1;
                rc = legacy_remove_first(&bx);
                if (rc) break;
                if (rng_uniform(rng) < rp->Retrap) {            /* functions.py:328-331: re-trapping adds a fresh pair */
                    rc = legacy_add_electron(&bx, rp, rng);
                    if (rc) break;
                }
                legacy_lifetimes(&bx, rp, T, rng);
                dt_filling = legacy_filling_time(&bx, rp, D, rng);
            }
            REC(out, r, i, ev, kd, ei, hi, bx.n_e, t_cur);
            i++;
        }
    } else if (rp->protocol == MCLO_PROTO_SIMULATE) {
        /* simulate.py:46-92, one pass per schedule segment (the reference has exactly one) */
        double t_off = 0.0;
        for (int sg = 0; sg < rp->seg_count && !rc; sg++) {
            const mclo_segment *S = &segs[rp->seg_begin + sg];
            double t_cur = 0.0;
            while (t_cur <= S->duration) {
                double T_now = S->T_start + S->T_rate * t_cur + 273.15;
                update_lifetimes(&bx, rp, T_now, S->A_opt, rng);
                double dt_fill = filling_time(&bx, rp, S->dose_rate, rng);
                double wmin; int arg, any;
                wait_min(&bx, &wmin, &arg, &any);
                double dt_recomb = any ? wmin : dt_fill;
                /* python min(a,b,c): first minimal element */
                double dt = dt_fill;
                if (dt_recomb < dt) dt = dt_recomb;
                if (S->dt_cap < dt) dt = S->dt_cap;
                if (i >= max_steps) { rc = MCLO_ERR_STEPS; break; }
                esteps += bx.n_e;
                t_cur += dt;
                int ev = 0, kd = 0, ei = -1, hi = -1;
                if (dt == dt_fill) {
                    ei = bx.n_e; hi = bx.n_h; kd = 1;
                    rc = box_add_electron(&bx, rp, rng);
                    if (rc) break;
                } else if (dt == dt_recomb) {
                    ei = arg; hi = bx.nearest[arg]; kd = 2; ev = 1;
                    rc = box_remove_pair(&bx, ei, hi);
                    if (rc) break;
                }
                REC(out, r, i, ev, kd, ei, hi, bx.n_e, t_off + t_cur);
                i++;
                if (S->duration != 0.0 && t_cur >= S->duration) break;
            }
            t_off += t_cur;
        }
    } else {
        /* tl_trap_lab.py:75-111 (TL_lab row) and :135-172 (ISO_lab experiment) */
        const mclo_segment *S = &segs[rp->seg_begin];
        int iso = (rp->protocol == MCLO_PROTO_ISO_LAB);
        double T0K = S->T_start + 273.15;
        double D = S->dose_rate;
        update_lifetimes(&bx, rp, T0K, 0.0, rng);
        double tf = filling_time(&bx, rp, D, rng);
        double t_cur = 0.0;
        int obs_idx = 0;
        const double *obs = iso ? obs_time + rp->obs_begin : NULL;
        for (;;) {
            if (iso) { if (!(obs_idx < rp->obs_count)) break; }
            else     { if (!(t_cur < S->duration)) break; }
            double wmin; int arg, any;
            wait_min(&bx, &wmin, &arg, &any);
            double dt_recomb = bx.n_e ? wmin : tf;       /* `.size`, not `.any()` */
            double dt = (tf < dt_recomb) ? tf : dt_recomb; /* python min(dt_recomb, tf) */
            double T_now = iso ? T0K : T0K + S->T_rate * (t_cur + dt);
            if (i >= max_steps) { rc = MCLO_ERR_STEPS; break; }
            esteps += bx.n_e;
            t_cur += dt;
            int ev = 0, kd, ei, hi;
            if (dt == tf) {
                ei = bx.n_e; hi = bx.n_h; kd = 1;
                rc = box_add_electron(&bx, rp, rng);
                if (rc) break;
            } else {
                ei = arg; hi = bx.nearest[arg]; kd = 2; ev = 1;
                rc = box_remove_pair(&bx, ei, hi);
                if (rc) break;
            }
            tf = filling_time(&bx, rp, D, rng);
            update_lifetimes(&bx, rp, T_now, 0.0, rng);
            REC(out, r, i, ev, kd, ei, hi, bx.n_e, t_cur);
            i++;
            if (iso) {
                while (obs_idx < rp->obs_count && t_cur >= obs[obs_idx]) {
                    if (out->obs_n_e) out->obs_n_e[rp->obs_begin + obs_idx] = bx.n_e;
                    if (out->obs_t) out->obs_t[rp->obs_begin + obs_idx] = t_cur;
                    obs_idx++;
                }
            }
        }
        if (!rc && !iso && i == 0) rc = MCLO_ERR_NOEVENT;
    }
done:
    if (!rc && rng->exhausted) rc = MCLO_ERR_STREAM;
    if (out->steps_used) out->steps_used[r] = i;
    if (out->final_n_e) out->final_n_e[r] = bx.n_e;
    if (out->consumed) out->consumed[r] = rng->consumed - c0;
    if (out->esteps) out->esteps[r] = esteps;
    if (out->status) out->status[r] = rc;
    box_free(&bx);
    return rc;
}

/*
 * Sequential run: all R replicas draw from ONE continuing stream, in order -- exactly what the
 * reference does with the global np.random state (simulate.py:36,46 never re-seeds).
 * Stops at the first failing replica and returns its status.
 */
int mclo_run_sequential(const mclo_replica *reps, int R, const mclo_segment *segs,
                        const double *obs_time, int max_steps, mclo_rng *rng, mclo_out *out)
{
    for (int r = 0; r < R; r++) {
        int rc = run_one(&reps[r], segs, obs_time, r, max_steps, rng, out);
        if (rc) return rc;
    }
    return MCLO_OK;
}

/*
 * Independent run: replica r uses its own MT19937 stream seeded with seed0 + r; replicas are
 * spread over host threads (pthreads).  This is the CPU baseline leg (not a parity mode: the
 * reference cannot run replicas concurrently).
 */
typedef struct {
    const mclo_replica *reps; int R; const mclo_segment *segs; const double *obs_time;
    int max_steps; uint32_t seed0; mclo_out *out;
    volatile int next; volatile int worst;
} par_job;

static void *par_worker(void *arg)
{
    par_job *J = (par_job *)arg;
    mclo_rng *rng = malloc(sizeof(mclo_rng));
    if (!rng) return NULL;
    for (;;) {
        int r = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (r >= J->R) break;
        mclo_rng_seed(rng, J->seed0 + (uint32_t)r);
        int rc = run_one(&J->reps[r], J->segs, J->obs_time, r, J->max_steps, rng, J->out);
        if (rc) __atomic_store_n(&J->worst, rc, __ATOMIC_RELAXED);
    }
    free(rng);
    return NULL;
}

int mclo_run_parallel(const mclo_replica *reps, int R, const mclo_segment *segs,
                      const double *obs_time, int max_steps, uint32_t seed0, int n_threads,
                      mclo_out *out)
{
    par_job J = { reps, R, segs, obs_time, max_steps, seed0, out, 0, MCLO_OK };
    if (n_threads < 1) n_threads = 1;
    if (n_threads > R) n_threads = R > 0 ? R : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
    if (!th) return MCLO_ERR_ALLOC;
    int started = 0;
    for (int k = 0; k < n_threads - 1; k++)
        if (pthread_create(&th[started], NULL, par_worker, &J) == 0) started++;
    par_worker(&J);
    for (int k = 0; k < started; k++) pthread_join(th[k], NULL);
    free(th);
    return J.worst;
}

int mclo_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
