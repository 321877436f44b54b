// mcl_smallbox.cu -- the Optimizer path: lab-protocol replicas with at most 124 traps, ONE WARP per replica (sm_100a).
//
// What runs here: TLTrapSim.TL_lab rows and ISO_lab experiments (reference src/class/tl_trap_lab.py:75-111,135-172) --
// boxes of <= 100 electrons and a few hundred to a few thousand holes that start empty or full, fill by irradiation
// and recombine while the temperature falls.  optimizer.objective (src/class/optimizer.py:82-84) runs 11 such rows per
// parameter candidate; a differential-evolution population is tens of thousands of them.
//
// Why its own kernel: in the block kernel (mcl_philox.cu) such a replica is latency-bound on a serial chain -- a
// 300-instruction step loop built for 10^4-electron boxes, grid searches that are chains of dependent L2 round trips,
// candidate lists that fills invalidate.  Here the whole box lives on chip:
//   * electrons in REGISTERS: lane l owns slots l, l+32, l+64, l+96 (distance to the cached nearest hole, that hole's slot, the
//     coordinates): no shared memory and no barrier for anything per-electron; "which of my electrons cached the dead
//     hole" is four register compares;
//   * holes in SHARED memory as three float arrays (12 bytes per hole); every nearest-hole search is a brute-force pass
//     of the warp over all slots -- no grid, no dependent loads, ~10 instructions per 32 holes;
//   * one Philox call per lane and step serves its four electrons (top 23 bits of a word: the exponential draw, low 9 bits:
//     the tunnelling-channel selector, ties settled by one more call of the warp);
//   * hole slots are reused lowest-first, so a replica needs n_h0 + N_e slots whatever its history.
// Index semantics follow the reference statistically: the hole order is the generation order (no spatial sorting), the
// "hole after h" of Box.remove_pair's shift-then-mask quirk (engine.py:168-171) is the next alive slot, new electrons see
// the old holes only (engine.py:147-152), existing electrons keep their stale cache until they are re-scanned.
// Which replicas come here is decided per replica from its own fields (mcl_abi.cu: lab protocol, N_e <= 124, hole
// capacity), never from the launch, so results depend only on (seed, global replica id).
#include <math_constants.h>
#include <cstdlib>
#include "mcl_common.cuh"
#include "mcl_rng.cuh"

namespace mcl {

namespace {

constexpr float F_INF = __builtin_huge_valf();
constexpr float DEAD_X = 1e30f;
constexpr float LN2F = 0.69314718055994530942f;
constexpr double L2E = 1.4426950408889634074;

// Element k (0..3, a run-time value) of four registers.  Written in PTX so that the compiler cannot turn the select chain
// into a dynamically indexed array -- which would move the electron state of the whole kernel to local memory.
__device__ __forceinline__ float pick4(float a0, float a1, float a2, float a3, int k)
{
    float r;
    asm("{\n\t.reg .pred p1, p2, p3;\n\tsetp.eq.s32 p1, %5, 1;\n\tsetp.eq.s32 p2, %5, 2;\n\tsetp.eq.s32 p3, %5, 3;\n\t"
        "selp.f32 %0, %2, %1, p1;\n\tselp.f32 %0, %3, %0, p2;\n\tselp.f32 %0, %4, %0, p3;\n\t}"
        : "=f"(r) : "f"(a0), "f"(a1), "f"(a2), "f"(a3), "r"(k));
    return r;
}
__device__ __forceinline__ int pick4(int a0, int a1, int a2, int a3, int k)
{
    return __float_as_int(pick4(__int_as_float(a0), __int_as_float(a1), __int_as_float(a2), __int_as_float(a3), k));
}

// a[k] = v for a run-time k, same reason
__device__ __forceinline__ void set4(float (&a)[4], int k, float v)
{
    asm("{\n\t.reg .pred p0, p1, p2, p3;\n\tsetp.eq.s32 p0, %4, 0;\n\tsetp.eq.s32 p1, %4, 1;\n\tsetp.eq.s32 p2, %4, 2;\n\tsetp.eq.s32 p3, %4, 3;\n\t"
        "selp.f32 %0, %5, %0, p0;\n\tselp.f32 %1, %5, %1, p1;\n\tselp.f32 %2, %5, %2, p2;\n\tselp.f32 %3, %5, %3, p3;\n\t}"
        : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]) : "r"(k), "f"(v));
}
__device__ __forceinline__ void set4(int (&a)[4], int k, int v)
{
    asm("{\n\t.reg .pred p0, p1, p2, p3;\n\tsetp.eq.s32 p0, %4, 0;\n\tsetp.eq.s32 p1, %4, 1;\n\tsetp.eq.s32 p2, %4, 2;\n\tsetp.eq.s32 p3, %4, 3;\n\t"
        "selp.b32 %0, %5, %0, p0;\n\tselp.b32 %1, %5, %1, p1;\n\tselp.b32 %2, %5, %2, p2;\n\tselp.b32 %3, %5, %3, p3;\n\t}"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]) : "r"(k), "r"(v));
}

// Nearest alive hole of (x, y, z) among slots [0, n_hs): (squared distance, slot), ties to the smaller slot; slot = -1 when
// no hole is alive.  WANT_FREE also returns the lowest dead slot (n_hs if none) for a fill.  All lanes get the results.
// The arrays are padded to a multiple of 128 slots and every slot at or beyond n_hs reads as dead (x = 1e30), so the pass
// needs neither index clamps nor bound predicates: three loads, six FP32 operations and the running minimum per hole.
template <bool WANT_FREE>
__device__ __forceinline__ void sb_search(const float *__restrict__ hx, const float *__restrict__ hy, const float *__restrict__ hz,
                                          int n_hs, float x, float y, float z, int lane, float &d2_out, int &slot_out, int &free_out)
{
    float bd = F_INF;
    int bj = 0x7fffffff, fd = 0x7fffffff;
    for (int j0 = lane; j0 < n_hs; j0 += 128) {
        float hxv[4], hyv[4], hzv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { hxv[u] = hx[j0 + 32 * u]; hyv[u] = hy[j0 + 32 * u]; hzv[u] = hz[j0 + 32 * u]; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = j0 + 32 * u;
            const float dx = x - hxv[u], dy = y - hyv[u], dz = z - hzv[u];
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));   // a dead hole (x = 1e30) gives +inf
            if (d2 < bd) { bd = d2; bj = j; }
            if (WANT_FREE && hxv[u] > 0.5f * DEAD_X && j < fd) fd = j;       // (the padding is dead: the minimum is <= n_hs)
        }
    }
    const uint32_t m = warp_min_u32(__float_as_uint(bd));           // d2 >= 0: unsigned order is float order
    const uint32_t s = warp_min_u32(__float_as_uint(bd) == m && bd < F_INF ? (uint32_t)bj : 0x7fffffffu);
    d2_out = __uint_as_float(m);
    slot_out = s == 0x7fffffffu ? -1 : (int)s;
    if (WANT_FREE) { const uint32_t f = warp_min_u32((uint32_t)fd); free_out = min((int)f, n_hs); }
}

template <bool TRACE>
__global__ void __launch_bounds__(32, 20) smallbox_kernel(const LaunchParams p, const RoundKeys K, const int *__restrict__ order, int hcap)
{
    const int r = order[blockIdx.x];
    const int lane = threadIdx.x;
    const mcl_replica rp = p.replicas[r];
    const mcl_segment S = p.segments[rp.seg_begin];
    const unsigned long long rid = p.replica_id0 + (unsigned long long)r;
    const uint32_t rid_lo = (uint32_t)rid, rid_hi = (uint32_t)(rid >> 32) & 0x0fffffffu;

    extern __shared__ __align__(16) float sb_smem[];
    float *hx = sb_smem, *hy = hx + hcap, *hz = hy + hcap;

    const float core_s = (float)(rp.side * rp.alpha * L2E);
    const float bnd_s = (float)(rp.side * rp.boundary_factor * rp.alpha * L2E);
    int status = MCL_OK;
    int n_e = rp.n_e0, n_slots = rp.n_e0, n_hs = rp.n_h0;
    if (rp.n_e0 > 124 || rp.N_e > 124 || rp.n_h0 + rp.N_e + 8 > hcap) status = MCL_ERR_CAPACITY;
    if (n_e > 0 && n_hs <= 0) status = MCL_ERR_NOHOLES;

    // electron state of this lane's four slots
    float c[4], px[4], py[4], pz[4];
    int nh[4];
    int birth[4];                  // creation order (legacy semantics: the OLDEST electron is the one that recombines)
#pragma unroll
    for (int k = 0; k < 4; k++) { c[k] = F_INF; nh[k] = -1; px[k] = py[k] = pz[k] = 0.f; birth[k] = lane + 32 * k; }
    int next_birth = rp.n_e0;

    // every slot that holds no hole reads as dead (hcap is a multiple of 128: a search pass never needs a bound check)
    for (int j = lane; j < hcap; j += 32) { hx[j] = DEAD_X; hy[j] = 0.f; hz[j] = 0.f; }
    __syncwarp();
    if (status == MCL_OK) {
        // ---------------- Box.seed (engine.py:124-129): holes in generation order, electrons slot i = electron i
        for (int j = lane; j < n_hs; j += 32) {
            uint32_t c0 = (uint32_t)j, c1 = 0u, c2 = rid_lo, c3 = rid_hi | (DOM_SEED_H << 28);
            philox4x32_10(c0, c1, c2, c3, K);
            hx[j] = u01(c0) * bnd_s; hy[j] = u01(c1) * bnd_s; hz[j] = u01(c2) * bnd_s;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = lane + 32 * k;
            if (i < n_e) {
                uint32_t c0 = (uint32_t)i, c1 = 0u, c2 = rid_lo, c3 = rid_hi | (DOM_SEED_E << 28);
                philox4x32_10(c0, c1, c2, c3, K);
                px[k] = u01(c0) * core_s; py[k] = u01(c1) * core_s; pz[k] = u01(c2) * core_s;
            }
        }
        __syncwarp();
        // Box._rebuild (engine.py:113-119): nearest hole of every electron
        for (int i = 0; i < n_e; i++) {
            const int own = i & 31, k = i >> 5;
            const float sx = pick4(px[0], px[1], px[2], px[3], k), sy = pick4(py[0], py[1], py[2], py[3], k),
                        sz = pick4(pz[0], pz[1], pz[2], pz[3], k);
            const float x = __shfl_sync(0xffffffffu, sx, own), y = __shfl_sync(0xffffffffu, sy, own), z = __shfl_sync(0xffffffffu, sz, own);
            float d2; int j, fr;
            sb_search<false>(hx, hy, hz, n_hs, x, y, z, lane, d2, j, fr);
            if (lane == own) { set4(c, k, sqrtf(d2)); set4(nh, k, j); }
        }
    }

    // ---------------- per-replica constants of the rate law, log2 domain (same formulation as mcl_philox.cu)
    const float lb = (float)log2(rp.b), ls = (float)log2(rp.s);
    const float eb1 = (float)(rp.E_loc_1 * L2E / rp.k_b), eb2 = (float)(rp.E_loc_2 * L2E / rp.k_b);
    const float ecb = (float)(rp.E_cb * L2E / rp.k_b);
    const bool one_ch_2 = rp.Retrap >= 1.0, one_ch_1 = rp.Retrap <= 0.0;
    const double sel_scaled = (one_ch_1 || one_ch_2) ? 0.0 : rp.Retrap * 512.0;
    const uint32_t sel_T9 = (uint32_t)sel_scaled;
    const uint32_t sel_frac = (uint32_t)fmin((sel_scaled - (double)sel_T9) * 4294967296.0, 4294967295.0);
    const uint32_t sel_tie = sel_frac ? sel_T9 : 0xffffu;
    const float cr_far = bnd_s * 1.7320508f;

    const bool iso = rp.protocol == MCL_PROTO_ISO_LAB;
    // Legacy (pre-refactor) TL semantics, reference src/est_params/functions.py:270-360: ONE channel draw per step for all
    // electrons (:125), every cache exact again after a pair is added (:316-317), the electron that recombines is the
    // OLDEST one whatever the waiting times say (np.where(...)[0] on a (1, n) array returns row indices, :246), and a
    // recombination re-adds a fresh pair with probability Retrap (:328-331).
    const bool legacy = rp.protocol == MCL_PROTO_TL_LEGACY;
    const float retrap_f = (float)rp.Retrap;
    const double *obs = p.obs_time + rp.obs_begin;
    const float dose_over_D0 = (float)(S.dose_rate / rp.D0);
    const bool dose_on = S.dose_rate != 0.0;
    const double T0K = S.T_start + 273.15;
    const size_t rec_base = (size_t)r * (size_t)p.max_steps;
    int obs_idx = 0, rec_i = 0;
    bool ever_filled = false;
    long long esteps = 0;
    double t_cur = 0.0;
    uint32_t sd0 = 0u, sd1 = 0u, sd2 = 0u, sd3 = 0u;        // step scalars of step (rec_i & ~31) + lane

    while (status == MCL_OK) {
        // ---------------- loop condition (tl_trap_lab.py:90,147)
        if (iso) { if (!(obs_idx < rp.obs_count)) break; }
        else { if (!(t_cur < S.duration)) break; }
        if (rec_i >= p.max_steps) { status = MCL_ERR_STEPS; break; }
        if ((rec_i & 31) == 0) {
            sd0 = 0u; sd1 = (uint32_t)(rec_i + lane); sd2 = rid_lo; sd3 = rid_hi | (DOM_SCALAR << 28);
            philox4x32_10(sd0, sd1, sd2, sd3, K);
        }
        // ---------------- rate-law prefactors at T(t_cur): the lifetimes of a step are drawn at the temperature the previous
        // step ended at (tl_trap_lab.py:93,105); ISO_lab is isothermal
        float A1, A2, g;
        bool has_cb;
        uint32_t leg_retrap_word = 0u;
        {
            const float T_now = (float)(iso ? T0K : (S.T_start + S.T_rate * t_cur + 273.15));
            float invT;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(invT) : "f"(T_now));
            A1 = fmaf(-eb1, invT, lb); A2 = fmaf(-eb2, invT, lb);
            if (one_ch_2) A1 = A2;
            if (one_ch_1) A2 = A1;
            if (legacy) {            // all electrons share the channel of this step: `rand(1) > Retrap` -> E_loc_1
                uint32_t s0 = 2u, s1 = (uint32_t)rec_i, s2 = rid_lo, s3 = rid_hi | (DOM_SCALAR << 28);
                philox4x32_10(s0, s1, s2, s3, K);
                leg_retrap_word = s1;
                A1 = A2 = (u01(s0) > retrap_f) ? A1 : A2;
            }
            g = fmaf(-ecb, invT, ls);
            has_cb = g > fminf(A1, A2) - cr_far - 30.0f;
        }
        // ---------------- per-electron clocks (engine.py:65-77, tl_trap_lab.py:51), log2 domain
        float best = F_INF;
        int bslot = 0x7fffffff;
        const int kmax = (n_slots + 31) >> 5;            // slot s lives in lane s & 31, register row s >> 5: rows in use
        if (kmax > 0) {
            uint32_t w0 = (uint32_t)lane, w1 = (uint32_t)rec_i, w2 = rid_lo, w3 = rid_hi | (DOM_STEP1 << 28);
            philox4x32_10(w0, w1, w2, w3, K);
            const uint32_t w[4] = {w0, w1, w2, w3};
            uint32_t ch2 = 0u, tie = 0u;
            if (A1 != A2) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t sel = w[k] & 0x1ffu;
                    ch2 |= sel < sel_T9 ? (1u << k) : 0u;
                    tie |= (sel == sel_tie && c[k] < F_INF) ? (1u << k) : 0u;
                }
                if (tie) {                          // 2^-9 per electron: one more call settles this lane's ties
                    uint32_t t0 = (uint32_t)lane, t1 = (uint32_t)rec_i, t2 = rid_lo, t3 = rid_hi | (DOM_SEL << 28);
                    philox4x32_10(t0, t1, t2, t3, K);
                    const uint32_t tw[4] = {t0, t1, t2, t3};
#pragma unroll
                    for (int k = 0; k < 4; k++) ch2 |= (((tie >> k) & 1u) && tw[k] < sel_frac) ? (1u << k) : 0u;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k >= kmax) break;                           // warp-uniform: register rows beyond the slots in use hold nothing
                const float le = lg2_fast(-lg2_fast(u01(w[k])));
                const float Ak = ((ch2 >> k) & 1u) ? A2 : A1;
                float l;
                if (has_cb) {
                    const float a = Ak - c[k];
                    const float kk = fmaxf(a, g) + lg2_fast(1.0f + ex2_fast(-fabsf(a - g)));
                    l = (le - kk) + (c[k] - c[k]);              // (c - c) turns an empty slot into NaN
                } else {
                    l = (le + c[k]) - Ak;
                }
                if (l < best) { best = l; bslot = lane + 32 * k; }
            }
        }
        // ---------------- warp argmin; equal clocks go to the smaller slot
        const float vmin = warp_min_f32(best);
        const int smin = (int)warp_min_u32(best == vmin ? (uint32_t)bslot : 0x7fffffffu);   // 0x7fffffff: no clock at all
        // ---------------- filling clock (tl_trap_lab.py:53-60) and dt (tl_trap_lab.py:91-92,150-151)
        const uint32_t u_fill = __shfl_sync(0xffffffffu, sd0, rec_i & 31);
        const float lam = (n_e == rp.N_e || !dose_on) ? 1e-20f : dose_over_D0 * (float)(rp.N_e - n_e);
        const float e2 = fmaxf(-lg2_fast(u01(u_fill)), 5.0e-8f);         // (see mcl_philox.cu: the SFU's worst near u -> 1)
        const float dt_fill = lam > 0.0f ? __fdividef(e2 * LN2F, lam) : 1e20f;
        const float dt_rec = n_e > 0 ? ex2_fast(vmin) * LN2F : dt_fill;
        const bool is_fill = dt_fill <= dt_rec;
        const float dt = is_fill ? dt_fill : dt_rec;
        esteps += n_e;
        t_cur += (double)dt;

        // Adds one electron + its twin hole.  Current semantics (Box.add_electron, engine.py:133-152): the new electron sees the
        // OLD holes only and nobody sees the new hole.  `exact` (legacy): afterwards every cache is exact, twin included.
        auto add_pair = [&](float nx, float ny, float nz, float qx, float qy, float qz, bool exact) -> bool {
            float d2; int j, hs;
            sb_search<true>(hx, hy, hz, n_hs, nx, ny, nz, lane, d2, j, hs);
            if (j < 0 && !exact) { status = MCL_ERR_NOHOLES; return false; }
            uint32_t fm = 0u;                                   // lowest free electron slot
#pragma unroll
            for (int k = 0; k < 4; k++) fm |= !(c[k] < F_INF) ? (1u << k) : 0u;
            const int es = (int)warp_min_u32(fm ? (uint32_t)(lane + 32 * (__ffs(fm) - 1)) : 0x7fffffffu);
            if (es >= 124 || hs >= hcap) { status = MCL_ERR_CAPACITY; return false; }
            float cnew = sqrtf(d2);
            int jn = j;
            if (exact) {
                const float tx = nx - qx, ty = ny - qy, tz = nz - qz;
                const float dt2 = fmaf(tx, tx, fmaf(ty, ty, tz * tz));
                if (j < 0 || dt2 < d2) { cnew = sqrtf(dt2); jn = hs; }       // np.argmin keeps the older hole on an exact tie
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float ex_ = px[k] - qx, ey_ = py[k] - qy, ez_ = pz[k] - qz;
                    const float dd = sqrtf(fmaf(ex_, ex_, fmaf(ey_, ey_, ez_ * ez_)));
                    if (c[k] < F_INF && dd < c[k]) { c[k] = dd; nh[k] = hs; }
                }
            }
            if (lane == (es & 31)) {
                const int k = es >> 5;
                set4(c, k, cnew); set4(nh, k, jn); set4(px, k, nx); set4(py, k, ny); set4(pz, k, nz); set4(birth, k, next_birth);
            }
            if (lane == 0) { hx[hs] = qx; hy[hs] = qy; hz[hs] = qz; }
            next_birth++;
            n_slots = max(n_slots, es + 1);
            n_hs = max(n_hs, hs + 1);
            n_e++;
            __syncwarp();
            return true;
        };

        int ev = 0;
        if (!is_fill) {
            // ---------------- Box.remove_pair (engine.py:154-175) / legacy recomber (functions.py:241-261)
            ev = 1;
            int victim = smin;
            if (legacy) {                    // the oldest alive electron
                uint32_t key = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < 4; k++) if (c[k] < F_INF) key = min(key, ((uint32_t)birth[k] << 7) | (uint32_t)(lane + 32 * k));
                victim = (int)(warp_min_u32(key) & 127u);
            }
            const int own = victim & 31, ks = victim >> 5;
            const int hsel = pick4(nh[0], nh[1], nh[2], nh[3], ks);
            const int h = __shfl_sync(0xffffffffu, hsel, own);
            if (lane == own) { set4(c, ks, F_INF); set4(nh, ks, -1); }
            n_e--;
            if (lane == 0) hx[h] = DEAD_X;
            __syncwarp();
            // stale-cache mode: electrons cached on the hole that FOLLOWS the removed one in index order are refreshed too
            // (shift-then-mask, engine.py:168-171); without fills every cache is exact and the refresh finds the same hole.
            // The legacy code selects the electrons to refresh before it shifts: no such quirk there.
            int h2 = -1;
            if (ever_filled && !legacy) {
                for (int base = h + 1; base < n_hs && h2 < 0; base += 32) {
                    const int j = base + lane;
                    const unsigned m = __ballot_sync(0xffffffffu, j < n_hs && hx[j] < 0.5f * DEAD_X);
                    if (m) h2 = base + __ffs(m) - 1;
                }
            }
            uint32_t hits = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) hits |= (nh[k] >= 0 && (nh[k] == h || nh[k] == h2)) ? (1u << k) : 0u;
            unsigned need = __ballot_sync(0xffffffffu, hits != 0u);
            while (need) {
                const int src = __ffs(need) - 1;
                const uint32_t hsrc = __shfl_sync(0xffffffffu, hits, src);
                const int k = __ffs(hsrc) - 1;
                const float sx = pick4(px[0], px[1], px[2], px[3], k), sy = pick4(py[0], py[1], py[2], py[3], k),
                            sz = pick4(pz[0], pz[1], pz[2], pz[3], k);
                const float x = __shfl_sync(0xffffffffu, sx, src), y = __shfl_sync(0xffffffffu, sy, src), z = __shfl_sync(0xffffffffu, sz, src);
                float d2; int j, fr;
                sb_search<false>(hx, hy, hz, n_hs, x, y, z, lane, d2, j, fr);
                if (j < 0) { status = MCL_ERR_NOHOLES; break; }
                if (lane == src) { set4(c, k, sqrtf(d2)); set4(nh, k, j); hits &= hits - 1u; }
                need = __ballot_sync(0xffffffffu, hits != 0u);
            }
            if (status != MCL_OK) break;
            if (legacy && u01(leg_retrap_word) < retrap_f) {
                // re-trapping (functions.py:328-331): a fresh pair, positions from two more scalar calls of this step
                uint32_t a0 = 3u, a1 = (uint32_t)rec_i, a2 = rid_lo, a3 = rid_hi | (DOM_SCALAR << 28);
                philox4x32_10(a0, a1, a2, a3, K);
                uint32_t b0 = 4u, b1 = (uint32_t)rec_i, b2 = rid_lo, b3 = rid_hi | (DOM_SCALAR << 28);
                philox4x32_10(b0, b1, b2, b3, K);
                if (!add_pair(u01(a0) * core_s, u01(a1) * core_s, u01(a2) * core_s, u01(b0) * bnd_s, u01(b1) * bnd_s, u01(b2) * bnd_s, true)) break;
            }
        } else {
            // ---------------- Box.add_electron (engine.py:133-152) / legacy add_electron + full refresh (functions.py:214-228,316-317)
            ever_filled = true;
            const int src = rec_i & 31;
            const float nx = u01(__shfl_sync(0xffffffffu, sd1, src)) * core_s, ny = u01(__shfl_sync(0xffffffffu, sd2, src)) * core_s,
                        nz = u01(__shfl_sync(0xffffffffu, sd3, src)) * core_s;
            uint32_t d0 = 1u, d1 = (uint32_t)rec_i, d2w = rid_lo, d3 = rid_hi | (DOM_SCALAR << 28);
            philox4x32_10(d0, d1, d2w, d3, K);
            if (!add_pair(nx, ny, nz, u01(d0) * bnd_s, u01(d1) * bnd_s, u01(d2w) * bnd_s, legacy)) break;
        }
        // ---------------- record (tl_trap_lab.py:107-108) and ISO observations (:153-172)
        if (TRACE && lane == 0) {
            if (p.event) p.event[rec_base + rec_i] = ev;
            if (p.n_e) p.n_e[rec_base + rec_i] = n_e;
            if (p.t) p.t[rec_base + rec_i] = t_cur;
        }
        rec_i++;
        if (iso) {
            while (obs_idx < rp.obs_count && t_cur >= obs[obs_idx]) {
                if (lane == 0 && p.obs_n_e) p.obs_n_e[rp.obs_begin + obs_idx] = n_e;
                obs_idx++;
            }
        }
    }
    if (status == MCL_OK && rp.protocol == MCL_PROTO_TL_LAB && rec_i == 0) status = MCL_ERR_NOEVENT;     // (the legacy loop returns its start value)
    if (lane == 0) {
        if (p.steps_used) p.steps_used[r] = rec_i;
        if (p.final_n_e) p.final_n_e[r] = n_e;
        if (p.esteps) p.esteps[r] = esteps;
        if (p.consumed) p.consumed[r] = 0;
        if (p.status) p.status[r] = status;
    }
}

}  // namespace

// Per-replica eligibility (never a property of the launch): lab protocol, at most 124 traps, holes that fit shared memory.
bool smallbox_eligible(const mcl_replica &rp)
{
    if (rp.protocol != MCL_PROTO_TL_LAB && rp.protocol != MCL_PROTO_ISO_LAB && rp.protocol != MCL_PROTO_TL_LEGACY) return false;
    if (rp.N_e > 124 || rp.n_e0 > 124 || rp.N_e < 0) return false;
    return smallbox_hole_capacity(rp) <= kSmallboxMaxHoles;
}

cudaError_t launch_smallbox(const LaunchParams &p, const int *order_dev, int count, int hcap, cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    const RoundKeys K = make_round_keys(p.seed);
    hcap = (hcap + 127) & ~127;                          // padded with dead slots: the search pass runs four loads deep without bound checks
    const size_t smem = sizeof(float) * 3 * (size_t)hcap;
    const bool trace = p.event || p.n_e || p.t;
    cudaError_t e;
    if (trace) {
        e = cudaFuncSetAttribute(smallbox_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smallbox_kernel<true><<<count, 32, smem, stream>>>(p, K, order_dev, hcap);
    } else {
        e = cudaFuncSetAttribute(smallbox_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smallbox_kernel<false><<<count, 32, smem, stream>>>(p, K, order_dev, hcap);
    }
    return cudaGetLastError();
}

}  // namespace mcl
