// mcl_philox.cu -- native throughput kernel of the trapped-charge kinetics loop (sm_100a).
//
// One CTA (NT = 32..512 threads) per replica; one electron per shared-memory slot; one Philox call serves a chunk of
// four slots (top 23 bits of a word: the exponential draw; its 9 low bits: the tunnelling-channel selector).
// What one step does (reference src/class/simulate.py:51-92, tl_trap_lab.py:90-108):
//   1. SWEEP.  Every alive electron i draws a selector and an exponential and forms its waiting time
//        wait_i = E_i / (k_cb + b*exp(-E_loc/kT - alpha*r_i))          (engine.py:65-77, tl_trap_lab.py:51)
//      in the log2 domain:  l_i = lg2(-lg2 u_i) - lg2 k_i,  wait_i = ln2 * 2^l_i,
//      so neither the tunnelling factor (alpha*r up to several hundred) nor E_cb/kT can under/overflow
//      FP32.  SFU ops per electron-step: lg2, lg2 (+ ex2, lg2 when the conduction-band rate matters).
//      A thread owns whole chunks of slots (chunk b -> thread b % NT); outside barrier-protected phases
//      only the owner touches cr[] / near[] of its slots.
//   2. ARGMIN.  Per-thread running min -> CREDUX.MIN.F32 + ballot per warp -> one smem row per warp
//      (value, slot, its hole), double-buffered by step parity -> ONE __syncthreads -> every warp
//      re-reduces the <= 16 rows and derives the same decision from uniform inputs.
//   3. EVENT.  dt = min(dt_fill, dt_recomb, dt_cap).  Recombination (engine.py:154-175): the owner
//      tombstones the slot; every thread scans the near[] entries of its own chunks for the dead hole (wide CTAs
//      keep per-hole masks that tell whether anybody, and which warp group, can have a hit at all); a hit
//      is re-targeted by its owner from the electron's K nearest-hole candidate list (exact while no hole
//      was ever added), else by a grid search of the owner's warp.  No CTA barrier.  Fill
//      (engine.py:133-152, stale-cache semantics): warp 0 places the pair; one extra barrier.
//   4. OUTPUT.  The (event, n_e, t) record is staged in smem and flushed 32 steps at a time with
//      coalesced stores; fused integer histograms of events / occupancy replace the trace for ensembles.
// Seeding (Box.seed/_rebuild, engine.py:113-129): holes and electrons are counting-sorted into a uniform
// cell grid; one warp per occupied cell loads the ~95 surrounding holes into registers once and builds the
// K-nearest lists of the cell's electrons.
// Random numbers: Philox4x32-10, key = seed (launch-wide round keys live in the constant bank),
// counter = (chunk or slot pair | element index, step, replica id lo, replica id hi | domain).  Results
// depend only on (seed, global replica id), never on the launch shape or the GPU count.
//
// Electron state: cr[slot] = alpha*log2(e)*r (FP32, +inf = empty slot), near[slot] = hole slot, both in
// shared memory, plus a 1-bit-per-hole alive bitmap; coordinates (pre-scaled by alpha*log2 e) and candidate
// lists in the replica's HBM slab, touched only on events.  Holes: [0,n_h0) sorted by grid cell
// (+cell_start table), [n_h0, ...) added by fills.
#include <math_constants.h>
#include <type_traits>
#include <cstdlib>
#include "mcl_common.cuh"
#include "mcl_rng.cuh"

namespace mcl {

namespace {

#ifndef MCL_SCAN_UNROLL
#define MCL_SCAN_UNROLL 4
#endif
#ifndef MCL_SCAN_UNROLL_NARROW
#define MCL_SCAN_UNROLL_NARROW 1
#endif
#ifndef MCL_FAST_LOOP
#define MCL_FAST_LOOP 1
#endif
#ifndef MCL_ONE_CHAINS
#define MCL_ONE_CHAINS 2
#endif
// Two clocks that are EQUAL in FP32 at the step's minimum (~2^-18 per step on a 10^4-electron box): 1 = the lowest lane, then
// the lowest warp wins, i.e. the winner depends on which thread owns which slot and therefore on the CTA width; 0 = the
// smallest slot wins whatever the width (exact launch-shape invariance) at a measured 2.3 % of C2's throughput (6.20e11
// vs 6.06e11: one more branch in each of the two serial reduction stages of every step).  Either way the result is a valid
// sample; only bit-identity across CTA widths is at stake.  Default: 1.
// How many sweeps the sweep team of the pipelined loop may be ahead of the decision warp (a power of two, <= 4: one named
// barrier and one set of per-thread minima per step in flight).
#ifndef MCL_PIPE_DEPTH
#define MCL_PIPE_DEPTH 2
#endif
#ifndef MCL_PIPE_SLEEP
#define MCL_PIPE_SLEEP 50
#endif
#ifndef MCL_OLD_TIEBREAK
#define MCL_OLD_TIEBREAK 1
#endif
// Resident CTAs per SM the narrow kernels are compiled for (= their register cap).  Measured on B200 (C5, 2000-electron boxes,
// 64 threads): 8 CTAs at 128 registers beat 16 CTAs at 64 registers by 9 % -- the spills of the 64-register build sit in
// the serial part of every step.  The 256-thread kernel (C2) is the other way round: 3 CTAs at 80 registers beat 2 at 128.
#ifndef MCL_NT64_MINB
#define MCL_NT64_MINB 8
#endif
#ifndef MCL_NT128_MINB
#define MCL_NT128_MINB 4
#endif
#ifndef MCL_NT32_MINB
#define MCL_NT32_MINB 16
#endif
#ifndef MCL_NT256_MINB
#define MCL_NT256_MINB 3
#endif
constexpr float F_INF = __builtin_huge_valf();
constexpr float DEAD_X = 1e30f;
constexpr float LN2F = 0.69314718055994530942f;
constexpr double L2E = 1.4426950408889634074;

#ifdef MCL_PROFILE_SKEW
// Profiling build only (scripts/build_variant.sh skew -DMCL_PROFILE_SKEW; scripts/skew_probe.py): cycles per warp index
// spent in the sweep / waiting at the step barrier / between the barrier and the next sweep, and the step count.
__device__ unsigned long long g_prof[10][32];      // rows 4..9: parts of the time after the barrier (see the marks)
#endif

struct Cfg {
    int cap_slots;     // even; smem slots per replica
    int g_max;         // largest grid edge the slab has room for
    int cap_cells;     // g_max^3 + 1
    int bm_words;      // smem words of the hole alive bitmap
    int share_bm;      // 1: per-hole sharing masks (targeted by >= 2 electrons; by which warp groups) let events skip the scan
    int ref_words;     // smem words of the warp-group reference mask (4 bits per hole)
    size_t off_holes;  // byte offset of the hole table inside the slab (16-byte aligned)
    size_t off_cand;   // byte offset of the candidate lists inside the slab
    int fast;          // 1: legs that qualify run the specialised step loop (0: always the general one; same results)
    size_t off_hpos2;  // byte offset (global slab) of the second hole table used by a regrid
    size_t off_hmap;   // byte offset (global slab) of the old-slot -> new-slot table of a regrid
    int fill_extra;    // fill-region slots beyond N_e: a fill that finds N_e + fill_extra of them in use regrids first
    int has_regrid;    // the slab has the regrid scratch (launches with a dosed leg somewhere)
    int relist;        // 1: a dose-free simulate leg that follows fills regrids and rebuilds the candidate lists
    int pipe;          // 1: isothermal legs that qualify for the specialised loop run its pipelined form (wide CTAs; same results)
    int off_pipe;      // byte offset (dynamic shared memory) of the sweep team's per-thread minima; inside ref4[] when that is large enough
    // shared-memory slab (small boxes): word offsets from the start of dynamic shared memory
    int sm_hpos, sm_exyz, sm_cstart, sm_cfill;
};


__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

struct Holes {
    float4 *pos;                 // (x, y, z, original index bits); x = 1e30 marks a dead hole
    const int *cell_start;
    int G, n_h0, n_slots;        // n_slots = high-water mark including the fill region
    float inv_w, w;
};

__device__ __forceinline__ int cell_index(const Holes &H, float x, float y, float z)
{
    const int cx = min(H.G - 1, (int)(x * H.inv_w)), cy = min(H.G - 1, (int)(y * H.inv_w)), cz = min(H.G - 1, (int)(z * H.inv_w));
    return (cx * H.G + cy) * H.G + cz;
}

// Warp-cooperative nearest alive hole of (x,y,z).  Returns (bits(d2) << 32 | slot) to all lanes.
__device__ unsigned long long warp_nearest(const Holes &H, float x, float y, float z, int lane, int exclude = -1)
{
    const int G = H.G;
    int cx = min(G - 1, max(0, (int)(x * H.inv_w)));
    int cy = min(G - 1, max(0, (int)(y * H.inv_w)));
    int cz = min(G - 1, max(0, (int)(z * H.inv_w)));
    unsigned long long best = ~0ull;
    for (int R = 1;; R++) {
        const int side = 2 * R + 1, ncb = side * side * side;
        for (int idx = lane; idx < ncb; idx += 32) {
            int oz = idx % side - R, oy = (idx / side) % side - R, ox = idx / (side * side) - R;
            if (R > 1 && max(abs(ox), max(abs(oy), abs(oz))) < R) continue;   // inner block already done
            int ax = cx + ox, ay = cy + oy, az = cz + oz;
            if ((unsigned)ax >= (unsigned)G || (unsigned)ay >= (unsigned)G || (unsigned)az >= (unsigned)G) continue;
            int c = (ax * G + ay) * G + az;
            int j0 = H.cell_start[c], j1 = H.cell_start[c + 1];
            // four holes of the cell per round trip to L2 / HBM (a cell holds ~3.5): the loads are what a search costs
            for (int jb = j0; jb < j1; jb += 4) {
                float4 hp[4];
#pragma unroll
                for (int u = 0; u < 4; u++) hp[u] = H.pos[min(jb + u, j1 - 1)];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int j = jb + u;
                    float dx = x - hp[u].x, dy = y - hp[u].y, dz = z - hp[u].z;
                    float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)j;
                    best = (j < j1 && key < best && j != exclude) ? key : best;
                }
            }
        }
        best = warp_min_u64(best);
        // everything outside the scanned block is at least `bound` away
        float bound = F_INF;
        bool covered = true;
        if (cx - R > 0) { bound = fminf(bound, x - (cx - R) * H.w); covered = false; }
        if (cx + R < G - 1) { bound = fminf(bound, (cx + R + 1) * H.w - x); covered = false; }
        if (cy - R > 0) { bound = fminf(bound, y - (cy - R) * H.w); covered = false; }
        if (cy + R < G - 1) { bound = fminf(bound, (cy + R + 1) * H.w - y); covered = false; }
        if (cz - R > 0) { bound = fminf(bound, z - (cz - R) * H.w); covered = false; }
        if (cz + R < G - 1) { bound = fminf(bound, (cz + R + 1) * H.w - z); covered = false; }
        float d2b = __uint_as_float((uint32_t)(best >> 32));
#ifdef MCL_PROFILE_SKEW
        if (lane == 0) atomicAdd(&g_prof[9][min(R, 31)], 1ull);       // searches that scanned ring R
#endif
        if (covered || d2b <= bound * bound) break;
    }
    // holes added by fills live outside the grid (again four loads in flight per lane)
    for (int jb = H.n_h0 + lane; jb < H.n_slots; jb += 128) {
        float4 hp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) hp[u] = H.pos[min(jb + 32 * u, H.n_slots - 1)];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + 32 * u;
            float dx = x - hp[u].x, dy = y - hp[u].y, dz = z - hp[u].z;
            float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)j;
            best = (j < H.n_slots && key < best && j != exclude) ? key : best;
        }
    }
    return warp_min_u64(best);
}

constexpr int KC = 4;      // candidate holes remembered per electron
constexpr int kFillRegrid = 512;
constexpr int TOMB_DIV = 16;   // compaction threshold (measured: 8 -> 16 is +1.7 % on 10^4-electron boxes, -1 % on 2000)

// Warp-cooperative search for the KC nearest alive holes of (x,y,z) among the INITIAL holes (the
// grid region).  out[k] (sorted, all lanes) = bits(d2) << 32 | slot, ~0 when fewer exist.
// While no hole is ever added (no fills), the nearest alive hole of an electron at any later time
// is the first still-alive entry of this list, so a re-search after a recombination is a lookup.
__device__ void warp_nearest_k(const Holes &H, float x, float y, float z, int lane, unsigned long long out[KC])
{
    const int G = H.G;
    int cx = min(G - 1, max(0, (int)(x * H.inv_w)));
    int cy = min(G - 1, max(0, (int)(y * H.inv_w)));
    int cz = min(G - 1, max(0, (int)(z * H.inv_w)));
    unsigned long long k0 = ~0ull, k1 = ~0ull, k2 = ~0ull, k3 = ~0ull;     // this lane's sorted best four
    for (int R = 1;; R++) {
        const int side = 2 * R + 1, ncb = side * side * side;
        for (int idx = lane; idx < ncb; idx += 32) {
            int oz = idx % side - R, oy = (idx / side) % side - R, ox = idx / (side * side) - R;
            if (R > 1 && max(abs(ox), max(abs(oy), abs(oz))) < R) continue;
            int ax = cx + ox, ay = cy + oy, az = cz + oz;
            if ((unsigned)ax >= (unsigned)G || (unsigned)ay >= (unsigned)G || (unsigned)az >= (unsigned)G) continue;
            int c = (ax * G + ay) * G + az;
            int j0 = H.cell_start[c], j1 = H.cell_start[c + 1];
            for (int j = j0; j < j1; j++) {
                const float4 hp = H.pos[j];
                float dx = x - hp.x, dy = y - hp.y, dz = z - hp.z;
                float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)j;
                if (key < k3) {
                    k3 = key;
                    if (k3 < k2) { unsigned long long t = k2; k2 = k3; k3 = t; }
                    if (k2 < k1) { unsigned long long t = k1; k1 = k2; k2 = t; }
                    if (k1 < k0) { unsigned long long t = k0; k0 = k1; k1 = t; }
                }
            }
        }
        // merge the per-lane lists: KC rounds of "global minimum, its owner pops"
        unsigned long long a0 = k0, a1 = k1, a2 = k2, a3 = k3;
#pragma unroll
        for (int k = 0; k < KC; k++) {
            unsigned long long m = warp_min_u64(a0);
            out[k] = m;
            if (a0 == m && m != ~0ull) { a0 = a1; a1 = a2; a2 = a3; a3 = ~0ull; }
        }
        float bound = F_INF;
        bool covered = true;
        if (cx - R > 0) { bound = fminf(bound, x - (cx - R) * H.w); covered = false; }
        if (cx + R < G - 1) { bound = fminf(bound, (cx + R + 1) * H.w - x); covered = false; }
        if (cy - R > 0) { bound = fminf(bound, y - (cy - R) * H.w); covered = false; }
        if (cy + R < G - 1) { bound = fminf(bound, (cy + R + 1) * H.w - y); covered = false; }
        if (cz - R > 0) { bound = fminf(bound, z - (cz - R) * H.w); covered = false; }
        if (cz + R < G - 1) { bound = fminf(bound, (cz + R + 1) * H.w - z); covered = false; }
        float d2k = __uint_as_float((uint32_t)(out[KC - 1] >> 32));     // NaN bits when the list is short
        if (covered || d2k <= bound * bound) break;
    }
}

template <int NT>
__device__ __forceinline__ void cta_sync()
{
    if (NT > 32) __syncthreads(); else __syncwarp();
}

// Named barriers of the pipelined step loop (ids 1..4; __syncthreads is barrier 0): the producer side arrives without waiting,
// the consumer side waits -- PTX's producer / consumer pattern.  `n` counts ALL participating threads (arrivers + waiters).
// (immediate ids: with an id in a register ptxas reserves all 16 barriers for the CTA)
template <int ID> __device__ __forceinline__ void named_sync_id(int n) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(n) : "memory"); }
template <int ID> __device__ __forceinline__ void named_arrive_id(int n) { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(n) : "memory"); }
constexpr int BAR_FULL = 1;                    // + step % depth
__device__ __forceinline__ void named_sync(int id, int n)
{
    if (id == 1) named_sync_id<1>(n); else if (id == 2) named_sync_id<2>(n); else if (id == 3) named_sync_id<3>(n); else named_sync_id<4>(n);
}
__device__ __forceinline__ void named_arrive(int id, int n)
{
    if (id == 1) named_arrive_id<1>(n); else if (id == 2) named_arrive_id<2>(n); else if (id == 3) named_arrive_id<3>(n); else named_arrive_id<4>(n);
}

#ifdef MCL_PIPE_STATS
// Profiling build only: [0] steps applied by the pipelined loop, [1] entries, [2] hand-backs: leg end / max steps, [3] compaction
// due, [4] filling clock, [5] hit overflow, [6] repair sweeps (a row whose winner was a stale slot), [7] steps with a scan,
// [8] hits, [9] warp searches
// [16..] cycles: sweep team (warp 0): 16 sweep, 17 wait DONE, 18 entry + arrive; decision warp: 20 wait FULL, 22 minimum +
// re-evaluations, 23 decision + event (up to DONE), 24 histogram
__device__ unsigned long long g_pipe[32];
__device__ unsigned long long g_pipe_h[4][8];    // [0]/[1]: count / cycles of decision-warp steps (FULL -> next FULL wait) by duration bucket; [2]/[3]: the same for warp 0's flag waits
__device__ __forceinline__ int pipe_bucket(long long c) { return c < 3000 ? 0 : c < 4500 ? 1 : c < 6000 ? 2 : c < 8000 ? 3 : c < 12000 ? 4 : c < 20000 ? 5 : c < 40000 ? 6 : 7; }
#define MCL_PSTAT(i, v) { if (lane == 0) atomicAdd(&g_pipe[i], (unsigned long long)(v)); }
#define MCL_PTIME(i) { const long long now_ = clock64(); if (lane == 0 && (warp == 0 || warp == NW - 1)) atomicAdd(&g_pipe[i], (unsigned long long)(now_ - pt_)); pt_ = now_; }
#else
#define MCL_PSTAT(i, v)
#define MCL_PTIME(i)
#endif

// NearT: type of the per-electron nearest-hole slot kept in shared memory.  uint16_t whenever the
// hole capacity allows (halves the footprint and the traffic of the post-event scan); the all-ones
// value marks an empty electron slot.
template <typename NearT> struct NearTraits;
template <> struct NearTraits<uint16_t> { static constexpr uint32_t DEAD = 0xffffu; };
template <> struct NearTraits<uint32_t> { static constexpr uint32_t DEAD = 0xffffffffu; };

// Seeding, part 3 (Box._rebuild, engine.py:113-119): the KC nearest holes of every electron.
// One warp per occupied cell: lanes 0..26 read the hole ranges of the 27 surrounding cells, the
// (<= 128) holes are spread 4 per lane and loaded once, then each electron of the cell needs 4
// distance evaluations per lane and KC warp-min rounds.
template <typename NearT>
__device__ __forceinline__ void seed_candidate_lists_impl(const Holes &H, const int *e_start, const float *ex, const float *ey,
                                                          const float *ez, float4 *cand_d, NearT *cand_j, float *cr, NearT *near,
                                                          int n_cells, int warp, int lane, int NW)
{
    constexpr uint32_t NEAR_DEAD = NearTraits<NearT>::DEAD;
    const float4 *hpos = H.pos;
    const int *cell_start = H.cell_start;
    // all lanes hold the same (d2, slot) list; lane k < KC writes entry k
    auto store_lists = [&](int i, const float d2[KC], const int jj[KC]) {
        float myd = F_INF; int myj = -1;
#pragma unroll
        for (int k = 0; k < KC; k++) if (lane == k) { myd = d2[k]; myj = jj[k]; }
        float d;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(myd));
        if (myj < 0) d = F_INF;
        if (lane < KC) {
            reinterpret_cast<float *>(cand_d + i)[lane] = d;
            cand_j[(size_t)i * KC + lane] = myj >= 0 ? (NearT)myj : (NearT)NEAR_DEAD;
            if (lane == 0) { cr[i] = d; near[i] = myj >= 0 ? (NearT)myj : (NearT)NEAR_DEAD; }
        }
    };
    auto slow_lists = [&](int i) {            // generic ring search (crowded cell or a list that is not yet exact)
        unsigned long long b[KC];
        warp_nearest_k(H, ex[i], ey[i], ez[i], lane, b);
        float d2[KC]; int jj[KC];
#pragma unroll
        for (int k = 0; k < KC; k++) {
            const bool ok = b[k] != ~0ull;
            d2[k] = ok ? __uint_as_float((uint32_t)(b[k] >> 32)) : F_INF;
            jj[k] = ok ? (int)(uint32_t)b[k] : -1;
        }
        store_lists(i, d2, jj);
    };
    const int G = H.G;
    for (int c = warp; c < n_cells; c += NW) {
        const int i0 = e_start[c], i1 = e_start[c + 1];
        if (i0 == i1) continue;                                     // warp-uniform
        const int ax = c / (G * G), ay = (c / G) % G, az = c % G;
        const int nx_ = ax + lane / 9 - 1, ny_ = ay + (lane / 3) % 3 - 1, nz_ = az + lane % 3 - 1;
        const bool valid = lane < 27 && (unsigned)nx_ < (unsigned)G && (unsigned)ny_ < (unsigned)G && (unsigned)nz_ < (unsigned)G;
        const int nc = valid ? (nx_ * G + ny_) * G + nz_ : 0;
        const int j0 = valid ? cell_start[nc] : 0;
        const int cnt = valid ? cell_start[nc + 1] - j0 : 0;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        const int off = inc - cnt;
        if (total > 128) { for (int i = i0; i < i1; i++) slow_lists(i); continue; }
        float4 hp[4]; int hj[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int item = lane + 32 * k;
            int pos = 0;                                            // last lane whose range starts at or before item
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                const int t = __shfl_sync(0xffffffffu, off, min(pos + sft, 31));
                if (pos + sft < 32 && t <= item) pos += sft;
            }
            const int jbase = __shfl_sync(0xffffffffu, j0, pos), obase = __shfl_sync(0xffffffffu, off, pos);
            const bool on = item < total;
            hj[k] = on ? jbase + (item - obase) : -1;
            hp[k] = on ? hpos[hj[k]] : make_float4(DEAD_X, 0.f, 0.f, 0.f);
        }
        // everything outside the 3x3x3 block is at least this far from an electron of cell c
        const float lo_x = (ax > 0) ? (ax - 1) * H.w : -F_INF, hi_x = (ax < G - 1) ? (ax + 2) * H.w : F_INF;
        const float lo_y = (ay > 0) ? (ay - 1) * H.w : -F_INF, hi_y = (ay < G - 1) ? (ay + 2) * H.w : F_INF;
        const float lo_z = (az > 0) ? (az - 1) * H.w : -F_INF, hi_z = (az < G - 1) ? (az + 2) * H.w : F_INF;
        // the cell's electrons are fetched 32 at a time (one round trip to L2 / HBM per cell instead of one per electron)
        for (int ib = i0; ib < i1; ib += 32) {
        const int mine_e = ib + lane;
        float xl = 0.f, yl = 0.f, zl = 0.f;
        if (mine_e < i1) { xl = ex[mine_e]; yl = ey[mine_e]; zl = ez[mine_e]; }
        const int i_end = min(i1, ib + 32);
        for (int i = ib; i < i_end; i++) {
            const float x = __shfl_sync(0xffffffffu, xl, i - ib), y = __shfl_sync(0xffffffffu, yl, i - ib), z = __shfl_sync(0xffffffffu, zl, i - ib);
            float d2[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float dx = x - hp[k].x, dy = y - hp[k].y, dz = z - hp[k].z;
                d2[k] = fmaf(dx, dx, fmaf(dy, dy, dz * dz));                   // inf for an unused register
            }
            float od[KC]; int oj[KC];
#pragma unroll
            for (int r_ = 0; r_ < KC; r_++) {
                const float m = fminf(fminf(d2[0], d2[1]), fminf(d2[2], d2[3]));
                const float g_ = warp_min_f32(m);
                const unsigned who = __ballot_sync(0xffffffffu, m == g_);
                const int owner = who ? __ffs(who) - 1 : 0;
                // branch-free: every lane picks its own best register; only the owner retires it
                int ksel = 3;
#pragma unroll
                for (int k = 2; k >= 0; k--) if (d2[k] == m) ksel = k;
                int mine_j = hj[3];
#pragma unroll
                for (int k = 0; k < 3; k++) if (k == ksel) mine_j = hj[k];
                const bool retire = (lane == owner);
#pragma unroll
                for (int k = 0; k < 4; k++) if (retire && k == ksel) d2[k] = F_INF;
                oj[r_] = (g_ < F_INF) ? __shfl_sync(0xffffffffu, mine_j, owner) : -1;
                od[r_] = g_;
            }
            // a face of the block that lies inside the grid limits how far the list is provably complete
            const float bound = fminf(fminf(fminf(x - lo_x, hi_x - x), fminf(y - lo_y, hi_y - y)), fminf(z - lo_z, hi_z - z));
            if (od[KC - 1] <= bound * bound) store_lists(i, od, oj);
            else slow_lists(i);
        }
        }
    }
}

// Out of line for the wide CTAs (keeps its register needs away from the allocation of their step loop),
// inline for the narrow ones (where the call ABI costs more than it saves).
template <typename NearT>
__device__ __noinline__ void seed_candidate_lists_call(const Holes &H, const int *e_start, const float *ex, const float *ey,
                                                       const float *ez, float4 *cand_d, NearT *cand_j, float *cr, NearT *near,
                                                       int n_cells, int warp, int lane, int NW)
{
    seed_candidate_lists_impl<NearT>(H, e_start, ex, ey, ez, cand_d, cand_j, cr, near, n_cells, warp, lane, NW);
}

// SLAB_SMEM: the hole table, the cell tables and the electron coordinates of the replica live in shared memory instead of
// its HBM slab (small boxes -- the Optimizer path: every nearest-hole search is then a handful of shared-memory reads
// instead of dependent L2 / HBM round trips).  Same algorithm, same results.
// REGRID: the launch has a dosed leg somewhere (mcl_run sized the slabs for it): regrids and candidate-list rebuilds are
// compiled in.  Dose-free launches (the BASELINE ensembles) run the instantiation without them, whose hole grid is a
// set of launch constants again -- fewer live registers in the step loop.
template <int NT, int MINB, typename NearT, int PPC, bool SLAB_SMEM = false, bool REGRID = true>
__global__ void __launch_bounds__(NT, MINB) philox_kernel(const LaunchParams p, const RoundKeys K, const Cfg cfg)
{
    constexpr int NW = NT / 32;
    constexpr uint32_t NEAR_DEAD = NearTraits<NearT>::DEAD;
    constexpr int SPC = 2 * PPC;              // slots per chunk: a thread owns whole chunks (chunk b -> thread b % NT)
    static_assert(PPC == 2, "a chunk is four slots: one 16-byte load of cr[], one Philox call when the channels are identical");
    // block b runs replica order[b] in slab b.  The index is re-read where it is needed (start, record flushes, end)
    // instead of being carried through the step loop in a register.
    auto replica_index = [&]() { return p.order ? p.order[blockIdx.x] : (int)blockIdx.x; };
    const int r = replica_index();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const mcl_replica rp = p.replicas[r];
    const unsigned long long rid = p.replica_id0 + (unsigned long long)r;
    const uint32_t rid_lo = (uint32_t)rid, rid_hi = (uint32_t)(rid >> 32) & 0x0fffffffu;

    // ---------------- shared memory
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *cr = reinterpret_cast<float *>(smem_raw);                  // [cap_slots]
    NearT *near = reinterpret_cast<NearT *>(cr + cfg.cap_slots);      // [cap_slots]
    uint32_t *hole_bm = reinterpret_cast<uint32_t *>(near + cfg.cap_slots);   // [bm_words] 1 = hole slot alive
    // Optional (wide CTAs): ref4 = 4 bits per hole, bit g set when some slot OWNED BY WARP GROUP g (NW / 4 warps each) has
    // the hole as its cached nearest; multi_bm = the hole is the cached nearest of >= 2 electrons.  Electrons leave a
    // hole only when it dies, so for an ALIVE hole both are exact as long as no electron is added (stale ref4 bits of
    // recombined electrons only cost a scan).  An event whose hole has no second electron skips the post-event scan
    // altogether; otherwise only the warp groups that reference the hole scan.  Slot ownership changes in a
    // compaction, which rebuilds ref4.
    uint32_t *multi_bm = hole_bm + cfg.bm_words;                      // [bm_words]
    uint32_t *ref4 = multi_bm + cfg.bm_words;                         // [ref_words] 8 holes per word
    constexpr bool SHARE_K = NT >= 256;             // narrow CTAs never use the masks: keep their code out of those kernels
    const bool share_bm = SHARE_K && cfg.share_bm != 0;
    const bool verify_skip = SHARE_K && cfg.share_bm == 2;     // test mode: scan anyway and flag a hit that the masks ruled out
    constexpr int GROUP_SHIFT = NW >= 4 ? (NW == 4 ? 0 : (NW == 8 ? 1 : 2)) : 0;      // warps per group = NW / 4 (NW >= 4)
    const int my_group = min(3, warp >> GROUP_SHIFT);
    auto group_of_slot = [&](int sl) { return min(3, (((sl >> 2) % NT) >> 5) >> GROUP_SHIFT); };   // owner = chunk % NT
    auto mark_target = [&](uint32_t j, int sl) {
        const int sh = 4 * (j & 7);
        const uint32_t old_ = atomicOr(&ref4[j >> 3], 1u << (sh + group_of_slot(sl)));
        if ((old_ >> sh) & 0xfu) atomicOr(&multi_bm[j >> 5], 1u << (j & 31));      // somebody was there already
    };
    __shared__ int4 red_row[2][32];            // per warp: (bits of its minimum, the slot, the slot's hole, -)
    __shared__ uint32_t stepdraw[2][32][4];     // step scalars of 32 consecutive steps, double-buffered per block of 32
    __shared__ int rec_ev[32], rec_ne[32];
    __shared__ double rec_t[32];
    __shared__ int s_scan[33];
    __shared__ int s_err;              // self-check failures (verify mode)
    // pipelined loop: s_cmd[0] = how far the decision warp has got (the sweep team polls it), and the state handed back at the end
    __shared__ int s_cmd[2], s_pipe_i[4];
    __shared__ double s_pipe_d[2];
    __shared__ long long s_pipe_ll;
    if (threadIdx.x == 0) s_err = 0;

    // ---------------- HBM slab
    unsigned char *ws = p.ws + (size_t)blockIdx.x * p.ws_stride;
    const size_t ce = (size_t)p.cap_e, ch = (size_t)p.cap_h;
    float *ex = reinterpret_cast<float *>(ws), *ey = ex + ce, *ez = ex + 2 * ce;
    float4 *hpos = reinterpret_cast<float4 *>(ws + cfg.off_holes);    // [cap_h] (x, y, z, original index)
    int *cell_start = reinterpret_cast<int *>(hpos + ch);             // [cap_cells]
    int *cell_fill = cell_start + cfg.cap_cells;                      // [cap_cells] seeding / regrid only
    int *e_id_buf = cell_fill + cfg.cap_cells;                        // [cap_e] seeding only
    if (SLAB_SMEM) {
        uint32_t *sm = reinterpret_cast<uint32_t *>(smem_raw);
        hpos = reinterpret_cast<float4 *>(sm + cfg.sm_hpos);
        ex = reinterpret_cast<float *>(sm + cfg.sm_exyz); ey = ex + ce; ez = ex + 2 * ce;
        cell_start = reinterpret_cast<int *>(sm + cfg.sm_cstart);
        cell_fill = reinterpret_cast<int *>(sm + cfg.sm_cfill);
    }
    float4 *hpos2 = reinterpret_cast<float4 *>(ws + cfg.off_hpos2);   // [cap_h] regrid only
    int *hmap = reinterpret_cast<int *>(ws + cfg.off_hmap);           // [cap_h] regrid only
    float4 *cand_d = reinterpret_cast<float4 *>(ws + cfg.off_cand);   // [cap_e] cr of the KC nearest initial holes
    NearT *cand_j = reinterpret_cast<NearT *>(cand_d + ce);           // [cap_e][KC] their slots (NEAR_DEAD = none)

#ifdef MCL_PROFILE_SKEW
    long long pf_sweep = 0, pf_wait = 0, pf_rest = 0, pf_steps = 0, pf_t = 0, pf_m = 0, pf_part[6] = {0, 0, 0, 0, 0, 0};
#define MCL_MARK(i) { const long long now_ = clock64(); pf_part[i] += now_ - pf_m; pf_m = now_; }
#else
#define MCL_MARK(i)
#endif
    int status = MCL_OK;
    const float core_s = (float)(rp.side * rp.alpha * L2E);
    const float bnd_s = (float)(rp.side * rp.boundary_factor * rp.alpha * L2E);
    const int n_h0 = rp.n_h0;         // INITIAL holes; H.n_h0 is the size of the grid region (changes in a regrid)
    int n_e = rp.n_e0;
    if (n_e > cfg.cap_slots - 4 || n_e > p.cap_e || n_h0 > p.cap_h) status = MCL_ERR_CAPACITY;
    if (n_e > 0 && n_h0 <= 0) status = MCL_ERR_NOHOLES;

    Holes H;
    H.pos = hpos; H.cell_start = cell_start; H.n_h0 = n_h0; H.n_slots = n_h0;
    H.G = max(1, min(cfg.g_max, (int)cbrtf((float)n_h0 * (1.0f / 3.0f))));
    H.w = bnd_s / (float)H.G; H.inv_w = (float)H.G / bnd_s;
    const int n_cells = H.G * H.G * H.G;

    for (int s = tid; s < cfg.cap_slots; s += NT) { cr[s] = F_INF; near[s] = (NearT)NEAR_DEAD; }
    for (int w = tid; w < cfg.bm_words; w += NT) {          // slots [0, n_h0) start alive, the fill region empty
        const int lo = 32 * w;
        hole_bm[w] = lo + 32 <= rp.n_h0 ? 0xffffffffu : (lo >= rp.n_h0 ? 0u : ((1u << (rp.n_h0 - lo)) - 1u));
    }

    if (status == MCL_OK) {
        // ---------------- Box.seed (engine.py:124-129): holes, binned into the cell grid
        for (int c = tid; c <= n_cells; c += NT) { cell_start[c] = 0; cell_fill[c] = 0; }
        cta_sync<NT>();
        auto hole_pos = [&](int j, float &x, float &y, float &z) {
            uint32_t c0 = (uint32_t)j, c1 = 0u, c2 = rid_lo, c3 = rid_hi | (DOM_SEED_H << 28);
            philox4x32_10(c0, c1, c2, c3, K);
            x = u01(c0) * bnd_s; y = u01(c1) * bnd_s; z = u01(c2) * bnd_s;
        };
        auto cell_of = [&](float x, float y, float z) { return cell_index(H, x, y, z); };
        for (int j = tid; j < n_h0; j += NT) {
            float x, y, z; hole_pos(j, x, y, z);
            atomicAdd(&cell_fill[cell_of(x, y, z)], 1);
        }
        cta_sync<NT>();
        // exclusive scan of the counts by warp 0 (cell_start[c] = holes in cells < c)
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < n_cells; base += 32) {
                int c = base + lane;
                int v = c < n_cells ? cell_fill[c] : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                if (c < n_cells) { cell_start[c] = run + inc - v; cell_fill[c] = 0; }
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) cell_start[n_cells] = run;
        }
        cta_sync<NT>();
        for (int j = tid; j < n_h0; j += NT) {
            float x, y, z; hole_pos(j, x, y, z);
            int c = cell_of(x, y, z);
            int q = cell_start[c] + atomicAdd(&cell_fill[c], 1);
            hpos[q] = make_float4(x, y, z, __int_as_float(j));
        }
        cta_sync<NT>();
        // deterministic order inside each cell (by original index): atomics above are unordered
        for (int c = tid; c < n_cells; c += NT) {
            int j0 = cell_start[c], j1 = cell_start[c + 1];
            for (int a = j0 + 1; a < j1; a++) {
                const float4 v = hpos[a];
                const int id = __float_as_int(v.w);
                int b = a - 1;
                while (b >= j0 && __float_as_int(hpos[b].w) > id) { hpos[b + 1] = hpos[b]; b--; }
                hpos[b + 1] = v;
            }
        }
        cta_sync<NT>();
        // ---------------- electrons, stored in grid-cell order (slot order carries no meaning in Philox mode;
        // it only has to be deterministic).  Grouping them by cell lets one warp fetch the ~95 holes around a
        // cell ONCE and serve every electron of the cell from registers.
        auto electron_pos = [&](int i, float &x, float &y, float &z) {
            uint32_t c0 = (uint32_t)i, c1 = 0u, c2 = rid_lo, c3 = rid_hi | (DOM_SEED_E << 28);
            philox4x32_10(c0, c1, c2, c3, K);
            x = u01(c0) * core_s; y = u01(c1) * core_s; z = u01(c2) * core_s;
        };
        int *e_start = cell_fill;                 // per-cell electron counts -> exclusive starts
        int *e_id = e_id_buf;                     // original electron index held by every sorted slot
        for (int c = tid; c <= n_cells; c += NT) e_start[c] = 0;
        for (int i = tid; i < n_e; i += NT) e_id[i] = -1;
        cta_sync<NT>();
        for (int i = tid; i < n_e; i += NT) {
            float x, y, z; electron_pos(i, x, y, z);
            atomicAdd(&e_start[cell_of(x, y, z)], 1);
        }
        cta_sync<NT>();
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < n_cells; base += 32) {
                int c = base + lane;
                int v = c < n_cells ? e_start[c] : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                if (c < n_cells) e_start[c] = run + inc - v;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) e_start[n_cells] = run;
        }
        cta_sync<NT>();
        for (int i = tid; i < n_e; i += NT) {     // unordered scatter into the cell's range ...
            float x, y, z; electron_pos(i, x, y, z);
            int q = e_start[cell_of(x, y, z)];
            while (atomicCAS(&e_id[q], -1, i) != -1) q++;
        }
        cta_sync<NT>();
        for (int c = tid; c < n_cells; c += NT) { // ... then a tiny insertion sort per cell makes it deterministic
            const int j0 = e_start[c], j1 = e_start[c + 1];
            for (int a = j0 + 1; a < j1; a++) {
                const int id = e_id[a];
                int b = a - 1;
                while (b >= j0 && e_id[b] > id) { e_id[b + 1] = e_id[b]; b--; }
                e_id[b + 1] = id;
            }
        }
        cta_sync<NT>();
        for (int i = tid; i < n_e; i += NT) {
            float x, y, z; electron_pos(e_id[i], x, y, z);
            ex[i] = x; ey[i] = y; ez[i] = z;
        }
        cta_sync<NT>();
        // ---------------- Box._rebuild (engine.py:113-119): the KC nearest holes of every electron (kept out of
        // line so that its register needs do not disturb the allocation of the step loop)
        if (share_bm) for (int w = tid; w < cfg.bm_words + cfg.ref_words; w += NT) multi_bm[w] = 0u;
        if (NT >= 256) seed_candidate_lists_call<NearT>(H, cell_fill, ex, ey, ez, cand_d, cand_j, cr, near, n_cells, warp, lane, NW);
        else seed_candidate_lists_impl<NearT>(H, cell_fill, ex, ey, ez, cand_d, cand_j, cr, near, n_cells, warp, lane, NW);
        cta_sync<NT>();
        if (share_bm) {
            for (int i = tid; i < n_e; i += NT) { const uint32_t j = near[i]; if (j != NEAR_DEAD) mark_target(j, i); }
            cta_sync<NT>();
        }
    }

    // ---------------- per-replica constants of the rate law, log2 domain
    const float lb = (float)log2(rp.b), ls = (float)log2(rp.s);
    const float eb1 = (float)(rp.E_loc_1 * L2E / rp.k_b), eb2 = (float)(rp.E_loc_2 * L2E / rp.k_b);
    const float ecb = (float)(rp.E_cb * L2E / rp.k_b);
    // Channel selector U < Retrap (engine.py:72)  <=>  word < thr ; Retrap >= 1 / <= 0 collapse the two channels.
    // (The one-warp kernel of the Optimizer path takes the selector from the 9 spare low bits of the word that carries the
    // exponential draw, ties settled by one more call: one Philox call per four electrons.  Here that variant MEASURED
    // SLOWER than a second call per chunk -- 4.17e11 vs 4.53e11 electron-steps/s on a two-channel C2: the first chunks a
    // thread sweeps cannot yet rule a tie out, so four warp-steps in ten paid for the settling call -- profiles/README.md.)
    const bool one_ch_2 = rp.Retrap >= 1.0, one_ch_1 = rp.Retrap <= 0.0;
    const uint32_t thr = (one_ch_1 || one_ch_2) ? 0u : (uint32_t)fmin(rp.Retrap * 4294967296.0, 4294967295.0);
    const float cr_far = bnd_s * 1.7320508f;        // no electron-hole distance exceeds the box diagonal

    const bool lab = rp.protocol != MCL_PROTO_SIMULATE;
    const bool iso = rp.protocol == MCL_PROTO_ISO_LAB;
    const double *obs = p.obs_time + rp.obs_begin;
    int obs_idx = 0;
    int n_slots = n_e;             // electron slots in use (alive or tombstoned)
    int n_fill_alive = 0;          // alive holes in the fill region
    bool ever_filled = false;
    bool lists_valid_ = true;      // (REGRID kernels; the others use !ever_filled) the candidate lists name the K nearest of ALL holes that were ever alive since they were built
    bool draws_valid = false;      // stepdraw holds the block of 32 steps that contains rec_i
    int rec_i = 0;
    long long esteps = 0;
    uint32_t es32 = 0u;            // electron-steps since the last carry into `esteps`
    double t_off = 0.0;

    const bool trace = (p.event != nullptr) || (p.n_e != nullptr) || (p.t != nullptr);
    // histogram
    const int hgroup = (p.hist.n_bins > 0 && p.hist_group) ? p.hist_group[r] : 0;
    // rows are validated on the host (plan_layout); a replica whose rows do not fit writes no histogram at all
    const bool hist_on = p.hist.n_bins > 0 && hgroup >= 0 && hgroup + rp.seg_count <= p.hist.n_groups;

    auto flush_records = [&](int count) {       // warp 0 only; count <= 32 records ending at rec_i
        if (!trace || warp != 0) return;
        __syncwarp();
        int first = rec_i - count;
        if (lane < count) {
            size_t q = (size_t)replica_index() * (size_t)p.max_steps + (size_t)(first + lane);
            if (p.event) p.event[q] = rec_ev[lane];
            if (p.n_e) p.n_e[q] = rec_ne[lane];
            if (p.t) p.t[q] = rec_t[lane];
        }
        __syncwarp();
    };

    // ---------------- regrid: every alive hole (initial or added by a fill) is re-binned into a fresh cell grid and the
    // hole slots are renumbered densely in cell order; the fill region is empty afterwards.  Called (by all threads)
    // when a fill finds N_e + fill_extra fill-region slots in use -- which bounds the hole capacity a replica needs at
    // n_h0 + 2 N_e + fill_extra whatever its history -- and at the start of a dose-free leg that follows fills, before
    // the candidate lists are rebuilt.  The trigger depends on the replica's own state only, never on the launch.
    // Cached nearest holes stay what they were (the reference's cache is stale by design, engine.py:147-152); only
    // their slot numbers change.  alive holes - n_e is invariant (pairs are added and removed together).
    // (Folding the fill region into the grid every kFillRegrid slots keeps the linear part of a nearest-hole search short:
    // C3 1.43e11 -> 1.55e11.  A constant of the replica's stream, like TOMB_DIV: WHEN slots are renumbered decides which hole
    // "follows" a removed one.)
    const int fill_cap = max(4, min(rp.N_e + cfg.fill_extra, kFillRegrid));
    bool can_regrid = false;       // only replicas with a dosed leg regrid (the launch then has the scratch for it)
    for (int sg = 0; sg < rp.seg_count; sg++) can_regrid |= p.segments[rp.seg_begin + sg].dose_rate != 0.0;
    can_regrid = REGRID && can_regrid && cfg.has_regrid;
    auto regrid = [&]() {
        cta_sync<NT>();
        const int n_old = H.n_slots;
        const int n_alive = n_h0 - rp.n_e0 + n_e;
        H.G = max(1, min(cfg.g_max, (int)cbrtf((float)n_alive * (1.0f / 3.0f))));
        H.w = bnd_s / (float)H.G; H.inv_w = (float)H.G / bnd_s;
        const int nc = H.G * H.G * H.G;
        for (int c = tid; c <= nc; c += NT) { cell_start[c] = 0; cell_fill[c] = 0; }
        cta_sync<NT>();
        for (int j = tid; j < n_old; j += NT) {
            const float4 v = hpos[j];
            if (v.x < 0.5f * DEAD_X) atomicAdd(&cell_fill[cell_index(H, v.x, v.y, v.z)], 1);
        }
        cta_sync<NT>();
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < nc; base += 32) {
                const int c = base + lane;
                const int v = c < nc ? cell_fill[c] : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                if (c < nc) { cell_start[c] = run + inc - v; cell_fill[c] = 0; }
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) cell_start[nc] = run;
        }
        cta_sync<NT>();
        for (int j = tid; j < n_old; j += NT) {
            const float4 v = hpos[j];
            if (v.x < 0.5f * DEAD_X) {
                const int c = cell_index(H, v.x, v.y, v.z);
                const int q = cell_start[c] + atomicAdd(&cell_fill[c], 1);
                hpos2[q] = make_float4(v.x, v.y, v.z, __int_as_float(j));
            }
        }
        cta_sync<NT>();
        for (int c = tid; c < nc; c += NT) {          // deterministic order inside each cell: by old slot
            const int j0 = cell_start[c], j1 = cell_start[c + 1];
            for (int a = j0 + 1; a < j1; a++) {
                const float4 v = hpos2[a];
                const int id = __float_as_int(v.w);
                int b = a - 1;
                while (b >= j0 && __float_as_int(hpos2[b].w) > id) { hpos2[b + 1] = hpos2[b]; b--; }
                hpos2[b + 1] = v;
            }
        }
        cta_sync<NT>();
        for (int q = tid; q < n_old; q += NT) {
            if (q < n_alive) {
                float4 v = hpos2[q];
                hmap[__float_as_int(v.w)] = q;
                v.w = __int_as_float(q);
                hpos[q] = v;
            } else {
                hpos[q].x = DEAD_X;
            }
        }
        for (int w = tid; w < cfg.bm_words; w += NT) {
            const int lo = 32 * w;
            hole_bm[w] = lo + 32 <= n_alive ? 0xffffffffu : (lo >= n_alive ? 0u : ((1u << (n_alive - lo)) - 1u));
        }
        cta_sync<NT>();
        for (int sl = tid; sl < n_slots; sl += NT) {
            const uint32_t nn = near[sl];
            if (nn != NEAR_DEAD) near[sl] = (NearT)hmap[nn];
        }
        H.n_h0 = n_alive; H.n_slots = n_alive; n_fill_alive = 0;
        lists_valid_ = false;                      // the lists name old slots
        cta_sync<NT>();
    };
    // The KC nearest alive holes of every electron, searched in the grid (after a regrid: all holes).  cr[] / near[] are
    // NOT touched: the cache stays stale where it is stale; the lists only say what a re-search would find.
    auto rebuild_lists = [&]() {
        for (int sl = warp; sl < n_slots; sl += NW) {
            if (!(cr[sl] < F_INF)) continue;                        // warp-uniform
            unsigned long long b[KC];
            warp_nearest_k(H, ex[sl], ey[sl], ez[sl], lane, b);
            if (lane < KC) {
                unsigned long long mine = b[0];
#pragma unroll
                for (int k = 1; k < KC; k++) if (lane == k) mine = b[k];
                const bool ok = mine != ~0ull;
                reinterpret_cast<float *>(cand_d + sl)[lane] = ok ? sqrtf(__uint_as_float((uint32_t)(mine >> 32))) : F_INF;
                cand_j[(size_t)sl * KC + lane] = ok ? (NearT)(uint32_t)mine : (NearT)NEAR_DEAD;
            }
        }
        cta_sync<NT>();
        lists_valid_ = true;
    };

    // New nearest hole of slot `sl` from its candidate list: the first remembered candidate that is still alive and is not
    // `h_dead` (the hole dying in the event at hand, whose bitmap bit may not be visible yet).  Distances and slots are
    // fetched together: ONE round trip to L2 / HBM.  false: list exhausted.
    auto retarget_from_list = [&](int sl, int h_dead, bool mark = true) -> bool {
        const float4 d4 = cand_d[sl];
        uint32_t cj[KC];
        if (sizeof(NearT) == 2) {
            const uint2 v = *reinterpret_cast<const uint2 *>(cand_j + (size_t)sl * KC);
            cj[0] = v.x & 0xffffu; cj[1] = v.x >> 16; cj[2] = v.y & 0xffffu; cj[3] = v.y >> 16;
        } else {
            const uint4 v = *reinterpret_cast<const uint4 *>(cand_j + (size_t)sl * KC);
            cj[0] = v.x; cj[1] = v.y; cj[2] = v.z; cj[3] = v.w;
        }
        const float dk[KC] = {d4.x, d4.y, d4.z, d4.w};
        bool fixed = false;
#pragma unroll
        for (int c = 0; c < KC; c++) {
            const uint32_t j = cj[c];
            if (!fixed && j != NEAR_DEAD && j != (uint32_t)h_dead && ((hole_bm[j >> 5] >> (j & 31)) & 1u)) {
                cr[sl] = dk[c]; near[sl] = (NearT)j; fixed = true;
                if (mark && share_bm && !ever_filled) mark_target(j, sl);
            }
        }
        return fixed;
    };

    for (int sg = 0; sg < rp.seg_count && status == MCL_OK; sg++) {
        const mcl_segment S = p.segments[rp.seg_begin + sg];
        const float dose_over_D0 = (float)(S.dose_rate / rp.D0);
        const bool dose_on = S.dose_rate != 0.0;
        const float dt_cap = (float)fmin(S.dt_cap, 3.0e38);
        const double T0K = S.T_start + 273.15;
        const float A_opt = lab ? 0.0f : (float)S.A_opt;
        double t_cur = 0.0;
        // histogram cursor for this leg
        const int hrow = hist_on ? (hgroup + sg) : 0;             // leg sg -> row hist_group[r] + sg (mcl_run validates the range)
        int hbin_next = 0;        // next bin whose left edge has not been passed yet

        // Histogram axis of this leg.  Bin edges are visited in order, so the cursor (hbin_next, hedge_next)
        // also tells which bin an event falls in: no per-event logarithm.  On the log-time axis the next edge
        // is the previous one times a constant ratio.
        const bool h_log = hist_on && p.hist.axis == MCL_AXIS_TIME_LOG;
        const bool h_temp = hist_on && p.hist.axis == MCL_AXIS_TEMP;
        const bool h_mono = hist_on && (!h_temp || S.T_rate > 0.0);       // cooling legs: events only, via bin_of
        const double h_ratio = h_log ? exp10((log10(p.hist.hi) - log10(p.hist.lo)) / (double)p.hist.n_bins) : 1.0;
        const double h_step = (p.hist.hi - p.hist.lo) / (double)(hist_on ? p.hist.n_bins : 1);
        auto edge_after = [&](int k, double prev) -> double {   // left edge of bin k, given the edge of bin k-1
            if (h_log) return prev * h_ratio;
            const double v = p.hist.lo + (double)k * h_step;
            return h_temp ? (v - S.T_start) / S.T_rate : v;
        };
        auto bin_of = [&](double t) -> int {                    // only for non-monotonic (cooling) legs
            const double f = (S.T_start + S.T_rate * t - p.hist.lo) / (p.hist.hi - p.hist.lo);
            if (!(f >= 0.0) || !(f < 1.0)) return -1;
            return min(p.hist.n_bins - 1, (int)(f * p.hist.n_bins));
        };
        double hedge_next = CUDART_INF;
        if (h_mono) hedge_next = h_log ? p.hist.lo : edge_after(0, 0.0);

        // Rate-law prefactors at temperature T(t) (log2 domain).  Isothermal legs (C2, ISO_lab) evaluate them once.
        float A1, A2, g; bool has_cb;
        const bool T_const = (lab && iso) || S.T_rate == 0.0;
        auto set_T = [&](double t) {
            const float T_now = (float)((lab && iso) ? T0K : (S.T_start + S.T_rate * t + 273.15));
            float invT;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(invT) : "f"(T_now));
            A1 = fmaf(-eb1, invT, lb); A2 = fmaf(-eb2, invT, lb);
            if (A_opt != 0.0f) {       // (A_opt + b e^{-E/kT}) e^{-alpha r}: fold the sum into the prefactor
                A1 = lg2_fast(A_opt + ex2_fast(A1));
                A2 = lg2_fast(A_opt + ex2_fast(A2));
            }
            if (one_ch_2) A1 = A2;
            if (one_ch_1) A2 = A1;
            g = fmaf(-ecb, invT, ls);                                   // lg2 k_cb
            // conduction-band channel can be skipped when it is < 2^-30 of the slowest tunnelling rate
            has_cb = g > fminf(A1, A2) - cr_far - 30.0f;
        };
        set_T(0.0);
        // A dose-free leg after fills (irradiation -> read-out): nothing can be added during the leg, so fresh candidate
        // lists stay exact for all of it and re-targeting is a lookup again instead of a grid search per hit.
        bool relist_pending = REGRID && cfg.relist && can_regrid && !lab && !dose_on && ever_filled && !lists_valid_ && n_e > 0;

        // The step loop exists twice.  FAST: the simulate protocol without a dose, without per-step records and while no
        // hole was ever added -- what BASELINE-sized ensembles run; protocol, dose, trace and fill-mode branches are compiled
        // out, which shortens the serial part of a step (C2 +6.7 %, C5 +9.3 %).  The general loop takes over, repeating the
        // step in hand, the moment the filling clock could matter.  Same records either way (scripts/compare_libs.py).
        // Returns 0: leg finished (or error), 1: continue in the general loop, 2: `budget` steps done (the pipelined loop asked
        // for a step in order; budget < 0: no limit).
        auto step_loop = [&](auto fast_tag, int budget) -> int {
            constexpr bool FAST = decltype(fast_tag)::value;
            const bool lab_o = lab, iso_o = iso, trace_o = trace, dose_o = dose_on, verify_o = verify_skip;
            if (FAST) { __builtin_assume(!ever_filled); __builtin_assume(lists_valid_); }
            // without regrids the lists are exact until the first fill and useless afterwards
#define lists_valid (REGRID ? lists_valid_ : !ever_filled)
            {
            const bool lab = FAST ? false : lab_o, iso = FAST ? false : iso_o, trace = FAST ? false : trace_o;
            const bool dose_on = FAST ? false : dose_o, verify_skip = FAST ? false : verify_o;
            for (;;) {
                // ---------------- loop condition (simulate.py:51; tl_trap_lab.py:90,147)
                if (!lab) { if (!(t_cur <= S.duration)) break; }
                else if (iso) { if (!(obs_idx < rp.obs_count)) break; }
                else { if (!(t_cur < S.duration)) break; }
                if (rec_i >= p.max_steps) { status = MCL_ERR_STEPS; break; }
                if (!FAST && REGRID) {
                    // the ONE place a regrid happens: before a step that could need a fill-region slot when all fill_cap of them
                    // hold alive holes, and at the start of a dose-free leg after fills (then the lists are rebuilt as well)
                    const bool fill_full = can_regrid && dose_on && H.n_slots - H.n_h0 >= fill_cap && n_fill_alive == H.n_slots - H.n_h0;
                    if (__builtin_expect(fill_full || relist_pending, 0)) {
                        regrid();
                        if (relist_pending) rebuild_lists();
                        relist_pending = false;
                    }
                }

                // ---------------- temperature-dependent scalars (uniform; FP32 from an FP64 clock)
                if (!T_const) set_T(t_cur);
                const int par = rec_i & 1;

#ifdef MCL_PROFILE_SKEW
                { const long long now = clock64(); if (pf_t) pf_rest += now - pf_t; pf_t = now; }
#endif
                // ---------------- per-electron clocks + running argmin
                float best = F_INF; int bslot = -1;
                // A thread owns whole CHUNKS of SPC = 4 slots (chunk b = slots 4b.. belongs to thread b % NT): one 16-byte
                // load feeds the clocks of a chunk and the post-event scan reads its nearest-hole slots with one load.
                const int n_chunks = (n_slots + SPC - 1) / SPC;
                auto pair_loop = [&](auto with_cb, auto one_channel) {
                    constexpr bool CB = decltype(with_cb)::value;
                    constexpr bool ONE = decltype(one_channel)::value;     // both tunnelling channels identical
                    if constexpr (ONE) {
                        // Identical channels: the selector draw cannot change anything, so no word is spent on it.  One
                        // Philox call serves the FOUR slots of a chunk (word k -> slot 4b + k); MCL_ONE_CHAINS chunks of
                        // the same owner per iteration keep that many independent Philox chains in flight.
                        const float4 *cr4 = reinterpret_cast<const float4 *>(cr);
                        auto chunks = [&](auto n_chains, int b0) {
                            constexpr int NCH = decltype(n_chains)::value;
                            float cs[NCH][4];
                            uint32_t w[NCH][4];
#pragma unroll
                            for (int q = 0; q < NCH; q++) {
                                const int b = b0 + q * NT;
                                const float4 cq = cr4[b];
                                cs[q][0] = cq.x; cs[q][1] = cq.y; cs[q][2] = cq.z; cs[q][3] = cq.w;
                                w[q][0] = (uint32_t)b; w[q][1] = (uint32_t)rec_i; w[q][2] = rid_lo; w[q][3] = rid_hi | (DOM_STEP1 << 28);
                            }
#pragma unroll
                            for (int q = 0; q < NCH; q++) philox4x32_10(w[q][0], w[q][1], w[q][2], w[q][3], K);
#pragma unroll
                            for (int q = 0; q < NCH; q++) {
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    const float le = lg2_fast(-lg2_fast(u01(w[q][k])));
                                    float l;
                                    if (CB) {
                                        // lg2(2^a + 2^g) = max + lg2(1 + 2^-|a-g|); (c - c) turns an empty slot into NaN
                                        const float a = A1 - cs[q][k];
                                        const float kk = fmaxf(a, g) + lg2_fast(1.0f + ex2_fast(-fabsf(a - g)));
                                        l = (le - kk) + (cs[q][k] - cs[q][k]);
                                    } else {
                                        l = le + cs[q][k];              // the uniform prefactor A1 is subtracted after the loop
                                    }
                                    if (l < best) { best = l; bslot = 4 * (b0 + q * NT) + k; }
                                }
                            }
                        };
                        int b0 = tid;
                        for (; b0 + (MCL_ONE_CHAINS - 1) * NT < n_chunks; b0 += MCL_ONE_CHAINS * NT)
                            chunks(std::integral_constant<int, MCL_ONE_CHAINS>{}, b0);
                        for (; b0 < n_chunks; b0 += NT) chunks(std::integral_constant<int, 1>{}, b0);       // tail: no wasted calls
                        if (!CB) best -= A1;
                    } else {
                        // Two distinct channels: one Philox call per slot PAIR -- words 0 / 2 pick the channels, words 1 / 3 are the
                        // exponential draws
                        for (int b = tid; b < n_chunks; b += NT) {
                            float cs[SPC];
                            const float4 cq = reinterpret_cast<const float4 *>(cr)[b];
                            cs[0] = cq.x; cs[1] = cq.y; cs[2] = cq.z; cs[3] = cq.w;
                            float l[SPC];
#pragma unroll
                            for (int i = 0; i < PPC; i++) {
                                uint32_t c0 = (uint32_t)(PPC * b + i), c1 = (uint32_t)rec_i, c2 = rid_lo, c3 = rid_hi | (DOM_STEP << 28);
                                philox4x32_10(c0, c1, c2, c3, K);
                                const float a0 = ((c0 < thr) ? A2 : A1) - cs[2 * i];
                                const float a1 = ((c2 < thr) ? A2 : A1) - cs[2 * i + 1];
                                const float le0 = lg2_fast(-lg2_fast(u01(c1)));
                                const float le1 = lg2_fast(-lg2_fast(u01(c3)));
                                if (CB) {
                                    const float k0 = fmaxf(a0, g) + lg2_fast(1.0f + ex2_fast(-fabsf(a0 - g)));
                                    const float k1 = fmaxf(a1, g) + lg2_fast(1.0f + ex2_fast(-fabsf(a1 - g)));
                                    l[2 * i] = (le0 - k0) + (cs[2 * i] - cs[2 * i]);
                                    l[2 * i + 1] = (le1 - k1) + (cs[2 * i + 1] - cs[2 * i + 1]);
                                } else {
                                    l[2 * i] = le0 - a0;
                                    l[2 * i + 1] = le1 - a1;
                                }
                            }
#pragma unroll
                            for (int i = 0; i < SPC; i++) if (l[i] < best) { best = l[i]; bslot = SPC * b + i; }
                        }
                    }
                };
                if (A1 == A2) { if (has_cb) pair_loop(std::true_type{}, std::true_type{}); else pair_loop(std::false_type{}, std::true_type{}); }
                else          { if (has_cb) pair_loop(std::true_type{}, std::false_type{}); else pair_loop(std::false_type{}, std::false_type{}); }
                // warp argmin -> one row per warp
                {
                    // ties on the FP32 clock go to the SMALLEST SLOT (a thread keeps its lowest slot already: ascending
                    // chunks, strict <), so the winner does not depend on which thread owns which slot -- i.e. on NT
                    float wv = warp_min_f32(best);
                    const unsigned m = __ballot_sync(0xffffffffu, best == wv);
                    int ws_;
#if MCL_OLD_TIEBREAK
                    ws_ = __shfl_sync(0xffffffffu, bslot, m ? (__ffs(m) - 1) : 0);
#else
                    if (m & (m - 1u)) ws_ = (int)warp_min_u32(best == wv ? (uint32_t)bslot : 0xffffffffu);   // tie (2^-18 per step), or no clock at all: -1
                    else ws_ = __shfl_sync(0xffffffffu, bslot, m ? (__ffs(m) - 1) : 0);
#endif
                    __syncwarp();
                    if (lane == 0) red_row[par][warp] = make_int4(__float_as_int(wv), ws_, ws_ >= 0 ? (int)near[ws_] : -1, 0);
                }
                if (warp == 0 && ((rec_i & 31) == 0 || !draws_valid)) {
                    // step scalars (fill clock + coordinates of a would-be new electron) of the 32 steps of this block
                    // of records, one Philox call per lane: counter-based, so step k gets the same words as if it were
                    // drawn on its own.  Double-buffered: the other buffer may still be read by a warp that is late in
                    // the previous step.
                    uint32_t c0 = 0u, c1 = (uint32_t)((rec_i & ~31) + lane), c2 = rid_lo, c3 = rid_hi | (DOM_SCALAR << 28);
                    philox4x32_10(c0, c1, c2, c3, K);
                    uint32_t *d = stepdraw[(rec_i >> 5) & 1][lane];
                    d[0] = c0; d[1] = c1; d[2] = c2; d[3] = c3;
                }
                draws_valid = true;
#ifdef MCL_PROFILE_SKEW
                { const long long now = clock64(); pf_sweep += now - pf_t; pf_t = now; }
#endif
                cta_sync<NT>();                                   // ===== B1
#ifdef MCL_PROFILE_SKEW
                { const long long now = clock64(); pf_wait += now - pf_t; pf_t = now; pf_steps++; pf_m = now; }
#endif
                float vmin; int smin, hmin;
                {
                    const int4 row = lane < NW ? red_row[par][lane] : make_int4(__float_as_int(F_INF), -1, -1, 0);
                    const float v = __int_as_float(row.x);
                    const int s = row.y, hh = row.z;
                    vmin = warp_min_f32(v);
                    unsigned m = __ballot_sync(0xffffffffu, v == vmin);
                    if (!MCL_OLD_TIEBREAK && (m & (m - 1u))) {             // tied rows (or no clock anywhere): the smallest slot wins, whatever warp holds it
                        const int s_low = (int)warp_min_u32(v == vmin ? (uint32_t)s : 0xffffffffu);
                        m = __ballot_sync(0xffffffffu, v == vmin && s == s_low);
                    }
                    const int src = m ? (__ffs(m) - 1) : 0;
                    smin = __shfl_sync(0xffffffffu, s, src);
                    hmin = __shfl_sync(0xffffffffu, hh, src);
                }
                // ---------------- filling clock (tl_trap_lab.py:53-60) and dt (simulate.py:58-60)
                // Without a dose the clock is exponential(1e20 s).  The smallest draw of the 24-bit uniform is -lg2(1 - 2^-24) =
                // 8.6e-8, but `lg2.approx` is only good to ~2^-23 ABSOLUTE there (measured: 1.2e-8 comes out,
                // tests/test_gpu_bench_shapes.py), so the clock's draw is clamped at 5e-8: dt_fill >= 5e-8 ln2 1e20 = 3.47e12 s.
                // It can only matter (the reference's spurious fill) when nothing else happens before 3e12 s --
                // otherwise it is not evaluated.
                const float dt_rec0 = n_e > 0 ? ex2_fast(vmin) * LN2F : F_INF;
                float dt_fill = F_INF;
                if (FAST && !(fminf(dt_rec0, dt_cap) < 3.0e12f)) {
                    // the filling clock could matter (spurious fill): the general loop repeats this step
                    cta_sync<NT>();
                    return 1;
                }
                if (dose_on || !((lab ? dt_rec0 : fminf(dt_rec0, dt_cap)) < 3.0e12f)) {       // (the lab loops have no step cap)
                    float lam = (n_e == rp.N_e || !dose_on) ? 1e-20f : dose_over_D0 * (float)(rp.N_e - n_e);
                    const float e2 = fmaxf(-lg2_fast(u01(stepdraw[(rec_i >> 5) & 1][rec_i & 31][0])), 5.0e-8f);
                    dt_fill = lam > 0.0f ? __fdividef(e2 * LN2F, lam) : 1e20f;
                }
                const float dt_rec = n_e > 0 ? dt_rec0 : dt_fill;
                float dt; bool is_fill, is_rec;
                if (!lab) {
                    dt = fminf(fminf(dt_fill, dt_rec), dt_cap);
                    is_fill = FAST ? false : (dt == dt_fill);
                    is_rec = !is_fill && (dt == dt_rec);
                } else {
                    is_fill = (dt_fill <= dt_rec);
                    is_rec = !is_fill;
                    dt = is_fill ? dt_fill : dt_rec;
                }
                es32 += (uint32_t)n_e;
                if (es32 > 0xC0000000u) { esteps += es32; es32 = 0u; }      // the 64-bit total lives in local memory
                const int n_before = n_e;
                const double t_new = t_cur + (double)dt;

                MCL_MARK(0)      // reduce + decision
                // ---------------- fused occupancy histogram: edges passed while n_e was n_before
                if (hedge_next <= t_new) {
                    // hbin_next <= n_bins; edge n_bins (the right end of the axis) closes the last bin
                    while (hedge_next <= t_new) {
                        if (tid == 0 && p.hist_occ && hbin_next < p.hist.n_bins) {
                            size_t q = (size_t)hrow * p.hist.n_bins + hbin_next;
                            atomicAdd(&p.hist_occ[q], (unsigned long long)n_before);
                            if (p.hist_occ_sq) atomicAdd(&p.hist_occ_sq[q], (unsigned long long)n_before * (unsigned long long)n_before);
                        }
                        hbin_next++;
                        hedge_next = hbin_next <= p.hist.n_bins ? edge_after(hbin_next, hedge_next) : CUDART_INF;
                    }
                }
                t_cur = t_new;

                MCL_MARK(1)      // histogram
                int ev = 0;
                if (is_rec) {
                    // ---------------- Box.remove_pair (engine.py:154-175)
                    ev = 1;
                    const int h = hmin;                                    // nearest hole of the winner (read before B1)
                    if (tid == ((smin / SPC) % NT)) { cr[smin] = F_INF; near[smin] = (NearT)NEAR_DEAD; }   // owner tombstones it
                    n_e--;
                    int h2 = -1;
                    if (ever_filled) {
                        // stale-cache mode: the reference also refreshes electrons cached on the hole that
                        // FOLLOWS the removed one in index order (shift-then-mask, engine.py:168-171)
                        const int last_w = (H.n_slots - 1) >> 5;
                        int w_ = (h + 1) >> 5;
                        uint32_t bits = w_ <= last_w ? (hole_bm[w_] & (0xffffffffu << ((h + 1) & 31))) : 0u;
                        while (!bits && w_ < last_w) bits = hole_bm[++w_];
                        if (bits) { const int j = 32 * w_ + __ffs(bits) - 1; if (j < H.n_slots) h2 = j; }
                    }
                    if (tid == 0) {
                        hpos[h].x = DEAD_X;
                        // Bit h is the only bit of the bitmap that changes in this phase, and every concurrent reader
                        // (retarget below, other warps) skips hole h before it looks at a word: reading the old or the
                        // new word gives the same answer.  compute-sanitizer racecheck flags exactly this read / RMW
                        // overlap; it is benign by construction.  The next barrier publishes the bit.
                        hole_bm[h >> 5] &= ~(1u << (h & 31));
                        if (hist_on && p.hist_events) {
                            // edges up to t_cur have been passed: the event sits in the bin before the cursor
                            const int b = h_mono ? (hbin_next - 1) : bin_of(t_cur);
                            if (b >= 0 && b < p.hist.n_bins) atomicAdd(&p.hist_events[(size_t)hrow * p.hist.n_bins + b], 1ull);
                        }
                    }
                    if (h >= H.n_h0) n_fill_alive--;
                    // Which of MY other electrons were cached on h (or h2)?  A thread owns the pairs it sweeps
                    // (q = tid, tid + NT, ...), and only the owner ever touches cr[] / near[] of a pair outside
                    // barrier-protected phases -- so re-targeting needs no CTA barrier at all.
                    // no other electron can be cached on h if h was never the target of a second one
                    const bool lone = share_bm && !ever_filled &&
                                      (!((multi_bm[h >> 5] >> (h & 31)) & 1u) ||                    // nobody else at all, or
                                       !((ref4[h >> 3] >> (4 * (h & 7) + my_group)) & 1u));         // nobody in MY warp group (warp-uniform)
                    int redo = -1;                      // a slot of mine that needs the warp-cooperative search
                    auto retarget = [&](int sl) { return lists_valid ? retarget_from_list(sl, h) : false; };
                    auto scan = [&](auto two_targets) {
                        constexpr bool TWO = decltype(two_targets)::value;
                        constexpr int NWORD = SPC * (int)sizeof(NearT) / 4;       // 32-bit words of near[] per chunk
                        // chunks whose loads are in flight together (narrow CTAs: few chunks per thread, and their 16 CTAs per SM
                        // at different phases feel every extra kilobyte of code in the instruction cache)
                        constexpr int SU = NT >= 256 ? MCL_SCAN_UNROLL : MCL_SCAN_UNROLL_NARROW;
                        const uint32_t hh = (uint32_t)h * 0x00010001u, hh2 = (uint32_t)h2 * 0x00010001u;
                        for (int b0 = tid; b0 < n_chunks; b0 += SU * NT) {
                            uint32_t w[SU][NWORD];
#pragma unroll
                            for (int u = 0; u < SU; u++) {
                                const int b = b0 + u * NT;
                                if (u > 0 && b >= n_chunks) {             // past the end: the empty-slot pattern matches no hole
#pragma unroll
                                    for (int k = 0; k < NWORD; k++) w[u][k] = 0xffffffffu;
                                } else if (NWORD == 4) {
                                    const uint4 v = reinterpret_cast<const uint4 *>(near)[b];
                                    w[u][0] = v.x; w[u][1 % NWORD] = v.y; w[u][2 % NWORD] = v.z; w[u][3 % NWORD] = v.w;
                                } else {
                                    const uint2 v = reinterpret_cast<const uint2 *>(near)[b];
                                    w[u][0] = v.x; w[u][1 % NWORD] = v.y;
                                }
                            }
                            uint32_t hits = 0u;                           // bit u: chunk b0 + u * NT may hold a match
#pragma unroll
                            for (int u = 0; u < SU; u++) {
                                bool hit;
                                if (sizeof(NearT) == 2) {
                                    // 16-bit slots: zero-halfword test (false positives possible, re-checked below)
                                    auto zh = [](uint32_t t) { return (t - 0x00010001u) & ~t & 0x80008000u; };
                                    uint32_t z = 0u;
#pragma unroll
                                    for (int k = 0; k < NWORD; k++) { z |= zh(w[u][k] ^ hh); if (TWO) z |= zh(w[u][k] ^ hh2); }
                                    hit = z != 0u;
                                } else {
                                    hit = false;
#pragma unroll
                                    for (int k = 0; k < NWORD; k++) { hit |= (w[u][k] == (uint32_t)h); if (TWO) hit |= (w[u][k] == (uint32_t)h2); }
                                }
                                hits |= hit ? (1u << u) : 0u;
                            }
                            while (hits) {
                                const int b = b0 + (__ffs(hits) - 1) * NT;
                                hits &= hits - 1u;
                                for (int k = 0; k < SPC; k++) {
                                    const int sl = SPC * b + k;
                                    const uint32_t nn = near[sl];
                                    if ((nn == (uint32_t)h || (TWO && nn == (uint32_t)h2)) && sl != smin && lone) atomicOr(&s_err, 1);
                                    if ((nn == (uint32_t)h || (TWO && nn == (uint32_t)h2)) && sl != smin) {
                                        if (!retarget(sl)) {
                                            // rare: park it until the warp search below
                                            if (redo >= 0) cr[sl] = -1.0f;      // more than one: mark, found again below
                                            else redo = sl;
                                        }
                                    }
                                }
                            }
                        }
                    };
                    MCL_MARK(2)  // retire the pair
                    if (!lone || verify_skip) { if (h2 >= 0) scan(std::true_type{}); else scan(std::false_type{}); }
                    MCL_MARK(3)  // scan + re-target from the lists
                    // Exhausted lists and fill mode (new holes become visible on a re-search, engine.py:171-175):
                    // the owner's WARP searches the cell grid cooperatively; still no CTA barrier.
                    {
                        unsigned need = __ballot_sync(0xffffffffu, redo >= 0);
                        while (need) {
                            const int src = __ffs(need) - 1;
                            need &= need - 1;
                            const int sl = __shfl_sync(0xffffffffu, redo, src);
                            unsigned long long b = warp_nearest(H, ex[sl], ey[sl], ez[sl], lane, h);
                            if (lane == src) {
                                cr[sl] = sqrtf(__uint_as_float((uint32_t)(b >> 32))); near[sl] = (NearT)(uint32_t)b;
                                if (share_bm && !ever_filled) mark_target((uint32_t)b, sl);
                                // further parked slots of this lane were marked with cr = -1
                                redo = -1;
                                for (int b = tid; b < n_chunks && redo < 0; b += NT)
                                    for (int k = 0; k < SPC && redo < 0; k++)
                                        if (cr[SPC * b + k] == -1.0f) redo = SPC * b + k;
                            }
                            need |= __ballot_sync(0xffffffffu, lane == src && redo >= 0) ;
                        }
                    }
                    MCL_MARK(4)  // warp searches
                    // ---------------- compaction: keep tombstones below 1/TOMB_DIV of the slots in use.  The Philox counter of a
                    // clock is its slot, so WHEN slots move is part of the stream definition: TOMB_DIV is one constant for
                    // every CTA width (results must not depend on the launch shape).
                    if ((n_slots - n_e) * TOMB_DIV > n_slots && n_slots >= 64) {
                        cta_sync<NT>();
                        int run = 0;
                        for (int base = 0; base < n_slots; base += NT) {
                            int s = base + tid;
                            float c = s < n_slots ? cr[s] : F_INF;
                            NearT nn = s < n_slots ? near[s] : (NearT)NEAR_DEAD;
                            bool alive = c < F_INF;
                            float x = 0.f, y = 0.f, z = 0.f;
                            float4 cd = make_float4(0.f, 0.f, 0.f, 0.f);
                            NearT cj[KC];
#pragma unroll
                            for (int k = 0; k < KC; k++) cj[k] = (NearT)NEAR_DEAD;
                            if (alive) {
                                x = ex[s]; y = ey[s]; z = ez[s];
                                if (lists_valid) {
                                    cd = cand_d[s];
#pragma unroll
                                    for (int k = 0; k < KC; k++) cj[k] = cand_j[(size_t)s * KC + k];
                                }
                            }
                            unsigned m = __ballot_sync(0xffffffffu, alive);
                            int wpre = __popc(m & ((1u << lane) - 1u));
                            if (lane == 0) s_scan[warp] = __popc(m);
                            cta_sync<NT>();
                            int woff = 0, tot = 0;
#pragma unroll
                            for (int k = 0; k < NW; k++) { int v = s_scan[k]; if (k < warp) woff += v; tot += v; }
                            int dst = run + woff + wpre;
                            cta_sync<NT>();                       // all reads of this tile done
                            if (alive) {
                                cr[dst] = c; near[dst] = nn; ex[dst] = x; ey[dst] = y; ez[dst] = z;
                                if (lists_valid) {
                                    cand_d[dst] = cd;
#pragma unroll
                                    for (int k = 0; k < KC; k++) cand_j[(size_t)dst * KC + k] = cj[k];
                                }
                            }
                            run += tot;
                        }
                        cta_sync<NT>();
                        for (int s = run + tid; s < n_slots; s += NT) { cr[s] = F_INF; near[s] = (NearT)NEAR_DEAD; }
                        n_slots = run;
                        cta_sync<NT>();
                        if (share_bm && !ever_filled) {           // slots changed owners: rebuild the warp-group reference mask
                            for (int w = tid; w < cfg.ref_words; w += NT) ref4[w] = 0u;
                            cta_sync<NT>();
                            for (int sl = tid; sl < n_slots; sl += NT) {
                                const uint32_t j = near[sl];
                                if (j != NEAR_DEAD) atomicOr(&ref4[j >> 3], 1u << (4 * (j & 7) + group_of_slot(sl)));
                            }
                            cta_sync<NT>();
                        }
                    }
                } else if (is_fill) {
                    // ---------------- Box.add_electron (engine.py:133-152), done by warp 0
                    ever_filled = true;
                    lists_valid_ = false;
                    int es = -1;
                    if (n_e == n_slots) es = n_slots;             // no tombstone to reuse
                    const bool append_e = (es >= 0);
                    if (append_e && es >= cfg.cap_slots - 4) { status = MCL_ERR_CAPACITY; break; }
                    const bool append_h = (n_fill_alive == H.n_slots - H.n_h0);       // (a full fill region was folded into the grid at the top of the step)
                    if (append_h && H.n_slots >= p.cap_h) { status = MCL_ERR_CAPACITY; break; }
                    if (warp == 0) {
                        if (!append_e) {
                            for (int base = 0; base < n_slots && es < 0; base += 32) {
                                int s = base + lane;
                                unsigned m = __ballot_sync(0xffffffffu, s < n_slots && !(cr[s] < F_INF));
                                if (m) es = base + __ffs(m) - 1;
                            }
                        }
                        int hs = H.n_slots;
                        if (!append_h) {
                            hs = -1;
                            for (int base = H.n_h0; base < H.n_slots && hs < 0; base += 32) {
                                int j = base + lane;
                                unsigned m = __ballot_sync(0xffffffffu, j < H.n_slots && !((hole_bm[j >> 5] >> (j & 31)) & 1u));
                                if (m) hs = base + __ffs(m) - 1;
                            }
                        }
                        const uint32_t *sd = stepdraw[(rec_i >> 5) & 1][rec_i & 31];
                        float nx = u01(sd[1]) * core_s, ny = u01(sd[2]) * core_s, nz = u01(sd[3]) * core_s;
                        uint32_t d0 = 1u, d1 = (uint32_t)rec_i, d2 = rid_lo, d3 = rid_hi | (DOM_SCALAR << 28);
                        philox4x32_10(d0, d1, d2, d3, K);
                        float qx = u01(d0) * bnd_s, qy = u01(d1) * bnd_s, qz = u01(d2) * bnd_s;
                        unsigned long long b = warp_nearest(H, nx, ny, nz, lane);     // OLD holes only
                        if (lane == 0) {
                            ex[es] = nx; ey[es] = ny; ez[es] = nz;
                            cr[es] = sqrtf(__uint_as_float((uint32_t)(b >> 32))); near[es] = (NearT)(uint32_t)b;
                            hpos[hs] = make_float4(qx, qy, qz, __int_as_float(hs));
                            hole_bm[hs >> 5] |= 1u << (hs & 31);
                        }
                    }
                    if (append_e) n_slots++;
                    if (append_h) H.n_slots++;
                    n_fill_alive++;
                    n_e++;
                    cta_sync<NT>();
                }

                MCL_MARK(5)      // compaction / fill
                // ---------------- record (simulate.py:64,85-89): staged, flushed 32 at a time
                if (trace && tid == 0) { rec_ev[rec_i & 31] = ev; rec_ne[rec_i & 31] = n_e; rec_t[rec_i & 31] = t_off + t_cur; }
                rec_i++;
                if (trace && (rec_i & 31) == 0) flush_records(32);
                if (iso) {
                    while (obs_idx < rp.obs_count && t_cur >= obs[obs_idx]) {
                        if (tid == 0 && p.obs_n_e) p.obs_n_e[rp.obs_begin + obs_idx] = n_e;
                        obs_idx++;
                    }
                }
                if (!lab && S.duration != 0.0 && t_cur >= S.duration) break;          // simulate.py:91-92
                if (FAST && --budget == 0) return 2;
            }
            }
            return 0;
#undef lists_valid
        };

        // ---------------- The specialised loop, PIPELINED (isothermal legs, wide CTAs).  In the loop above a warp sweeps, waits for
        // the slowest warp, and then EVERY warp walks the same ~300-instruction decision / event path before the next sweep
        // can start: more than half of a step is serial.  Here the roles are split: warps 0 .. NW-2 (the sweep team) SWEEP
        // step k+1 while the last warp DECIDES step k and applies its event.  Two facts make that exact:
        //  * At constant temperature the clocks of step k+1 depend on event k only through the slots the event touches, and
        //    clocks are counter-based (slot, step): whoever re-evaluates a slot gets THE clock the loop in order would use.
        //  * LAZY RE-TARGETING.  An electron cached on a hole that has died keeps its stale (smaller) distance, so its clock
        //    is a LOWER BOUND of the true one (same draw, larger rate).  If it does not win the step with the bound it cannot
        //    win with the true clock either, so nothing has to be done; if it does win, the decision warp re-targets it
        //    (candidate list, else a grid search), re-evaluates the clocks of the thread that reported it and takes the
        //    minimum again.  No post-event scan, no sharing masks; each step's winner and waiting time are those of the
        //    loop in order (an exact FP32 tie at the minimum may go to the other slot).
        //   sweep team:     sweep(k) -> wait until step k-2 is done (a flag) -> every thread's best (clock, slot) and one row per warp
        //                   (clock, slot, reporting lane) -> arrive FULL[k]
        //   decision warp:  wait FULL[k] -> minimum of the rows -> winner retired while the sweep ran, stale, or written while
        //                   the sweep ran?  re-evaluate the reporting thread's chunks (one per lane), rebuild that warp's row
        //                   from its threads' entries, again -> decision -> event -> step k done (flag) -> histograms
        // Steps the pipeline cannot take (compaction due, last step of the leg, filling clock relevant) are handed back BEFORE
        // anything of them is applied; on the way out every stale electron is re-targeted and the sharing masks are rebuilt, so
        // the loop in order finds the state it would have produced itself.
        // Ramps qualify too while both channels are identical and the conduction-band term is off: the clocks are then
        // lg2(-lg2 u) + cr - A1(T), the temperature enters as ONE uniform offset, and the sweep does not need it at all.
        // (64-thread CTAs -- one sweep warp, one decision warp -- measured slower than their loop in order: C5 3.22e11 vs 3.40e11.)
        constexpr bool PIPE_K = NT >= 128 && !SLAB_SMEM;
        auto pipe_loop = [&]() -> int {               // returns the number of steps it took
            // (Measured, no gain: a double share of chunks for the team warp that shares a scheduler with the decision warp -- it
            // took twice as long, 7.74e11 -> 7.09e11; CTAs of 288 threads = 8 team warps + the decision warp at 72 registers:
            // 7.93e11 vs 7.85e11; the team 4 sweeps ahead instead of 2; a named barrier or a sleeping poll instead of the spin.)
            constexpr int NTS = PIPE_K ? NT - 32 : 32, DW = NW - 1;         // (narrow CTAs never come here: PIPE_K)
            constexpr int DEPTH = MCL_PIPE_DEPTH;
            constexpr unsigned FULLM = 0xffffffffu;
            if (!PIPE_K) return 0;
            const int k0 = rec_i;
            const int n_chunks = (n_slots + SPC - 1) / SPC;
            const bool one = (A1 == A2);
            int4 *team_row = &red_row[0][0];                  // [DEPTH][NW]
            static_assert(MCL_PIPE_DEPTH * NW <= 64, "team rows live in red_row[2][32]");
            float2 *team_best = reinterpret_cast<float2 *>(smem_raw + cfg.off_pipe);      // [DEPTH][NTS] (clock, slot bits); may alias ref4[]
            // raw clocks (identical channels without a conduction-band term: before the uniform prefactor A1 is subtracted)
            // of chunks b_first, b_first + stride, ... : the expressions of pair_loop
            auto sweep_impl = [&](auto with_cb, auto one_channel, int b_first, int stride, int step, float &best, int &bslot) {
                constexpr bool CB = decltype(with_cb)::value;
                constexpr bool ONE = decltype(one_channel)::value;
                const float4 *cr4 = reinterpret_cast<const float4 *>(cr);
                if constexpr (ONE) {
                    auto chunks = [&](auto n_chains, int b0) {
                        constexpr int NCH = decltype(n_chains)::value;
                        float cs[NCH][4];
                        uint32_t w[NCH][4];
#pragma unroll
                        for (int q = 0; q < NCH; q++) {
                            const int b = b0 + q * stride;
                            const float4 cq = cr4[b];
                            cs[q][0] = cq.x; cs[q][1] = cq.y; cs[q][2] = cq.z; cs[q][3] = cq.w;
                            w[q][0] = (uint32_t)b; w[q][1] = (uint32_t)step; w[q][2] = rid_lo; w[q][3] = rid_hi | (DOM_STEP1 << 28);
                        }
#pragma unroll
                        for (int q = 0; q < NCH; q++) philox4x32_10(w[q][0], w[q][1], w[q][2], w[q][3], K);
#pragma unroll
                        for (int q = 0; q < NCH; q++) {
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const float le = lg2_fast(-lg2_fast(u01(w[q][k])));
                                float l;
                                if (CB) {
                                    const float a = A1 - cs[q][k];
                                    const float kk = fmaxf(a, g) + lg2_fast(1.0f + ex2_fast(-fabsf(a - g)));
                                    l = (le - kk) + (cs[q][k] - cs[q][k]);
                                } else {
                                    l = le + cs[q][k];
                                }
                                if (l < best) { best = l; bslot = 4 * (b0 + q * stride) + k; }
                            }
                        }
                    };
                    int b0 = b_first;
                    for (; b0 + (MCL_ONE_CHAINS - 1) * stride < n_chunks; b0 += MCL_ONE_CHAINS * stride)
                        chunks(std::integral_constant<int, MCL_ONE_CHAINS>{}, b0);
                    for (; b0 < n_chunks; b0 += stride) chunks(std::integral_constant<int, 1>{}, b0);
                } else {
                    for (int b = b_first; b < n_chunks; b += stride) {
                        float cs[SPC];
                        const float4 cq = cr4[b];
                        cs[0] = cq.x; cs[1] = cq.y; cs[2] = cq.z; cs[3] = cq.w;
                        float l[SPC];
#pragma unroll
                        for (int i = 0; i < PPC; i++) {
                            uint32_t c0 = (uint32_t)(PPC * b + i), c1 = (uint32_t)step, c2 = rid_lo, c3 = rid_hi | (DOM_STEP << 28);
                            philox4x32_10(c0, c1, c2, c3, K);
                            const float a0 = ((c0 < thr) ? A2 : A1) - cs[2 * i];
                            const float a1 = ((c2 < thr) ? A2 : A1) - cs[2 * i + 1];
                            const float le0 = lg2_fast(-lg2_fast(u01(c1)));
                            const float le1 = lg2_fast(-lg2_fast(u01(c3)));
                            if (CB) {
                                const float k0_ = fmaxf(a0, g) + lg2_fast(1.0f + ex2_fast(-fabsf(a0 - g)));
                                const float k1_ = fmaxf(a1, g) + lg2_fast(1.0f + ex2_fast(-fabsf(a1 - g)));
                                l[2 * i] = (le0 - k0_) + (cs[2 * i] - cs[2 * i]);
                                l[2 * i + 1] = (le1 - k1_) + (cs[2 * i + 1] - cs[2 * i + 1]);
                            } else {
                                l[2 * i] = le0 - a0;
                                l[2 * i + 1] = le1 - a1;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < SPC; i++) if (l[i] < best) { best = l[i]; bslot = SPC * b + i; }
                    }
                }
            };
            auto sweep = [&](int b_first, int stride, int step, float &best, int &bslot) {
                if (one) { if (has_cb) sweep_impl(std::true_type{}, std::true_type{}, b_first, stride, step, best, bslot); else sweep_impl(std::false_type{}, std::true_type{}, b_first, stride, step, best, bslot); }
                else     { if (has_cb) sweep_impl(std::true_type{}, std::false_type{}, b_first, stride, step, best, bslot); else sweep_impl(std::false_type{}, std::false_type{}, b_first, stride, step, best, bslot); }
            };
            auto hole_alive = [&](uint32_t j) { return ((hole_bm[j >> 5] >> (j & 31)) & 1u) != 0u; };

            volatile int *s_done = &s_cmd[0];       // steps < s_done are applied; -(j+1): step j was handed back
            if (tid == 0) s_cmd[0] = k0;
            cta_sync<NT>();
            MCL_PSTAT(1, warp == DW)
#ifdef MCL_PIPE_STATS
            long long pt_ = clock64();
#endif
            if (warp != DW) {
                // ===== the sweep team
                for (int k = k0;; k++) {
                    const int par = k & (DEPTH - 1);
                    float best = F_INF; int bslot = -1;
                    const int d_early = *s_done;          // read before the sweep, looked at after it: usually step k-2 is long done
#ifdef MCL_PIPE_STATS
                    const long long sw0_ = clock64();
#endif
                    sweep(tid, NTS, k, best, bslot);
#ifdef MCL_PIPE_STATS
                    if (lane == 0 && warp < 7) atomicAdd(&g_pipe[25 + warp], (unsigned long long)(clock64() - sw0_));     // [25..31]: sweep cycles of team warp 0..6
#endif
                    const float wv = warp_min_f32(best);
                    const unsigned wm_ = __ballot_sync(FULLM, best == wv);
                    const int wl_ = wm_ ? (__ffs(wm_) - 1) : 0;
                    const int ws_ = __shfl_sync(FULLM, bslot, wl_);
                    MCL_PTIME(16)
                    if (k >= k0 + DEPTH) {
                        // step k-2 applied (or handed back) and its entries read?  A flag, not a barrier: the team's warps
                        // must not wait for each other here (they meet at FULL only through the decision warp).
                        int d = d_early, a = d < 0 ? -d : d;
#ifdef MCL_PIPE_STATS
                        const long long fw0_ = clock64();
#endif
                        // (a sleeping poll: a team that is held up by its decision warp -- small boxes -- must not take issue slots from it)
                        if (a < k - DEPTH + 1) { d = *s_done; a = d < 0 ? -d : d; }
                        while (a < k - DEPTH + 1) { __nanosleep(MCL_PIPE_SLEEP); d = *s_done; a = d < 0 ? -d : d; }
#ifdef MCL_PIPE_STATS
                        if (lane == 0 && warp == 0) { const long long c_ = clock64() - fw0_; atomicAdd(&g_pipe_h[2][pipe_bucket(c_)], 1ull); atomicAdd(&g_pipe_h[3][pipe_bucket(c_)], (unsigned long long)c_); }
#endif
                        if (d < 0 && -d - 1 <= k - DEPTH) break;          // step -d-1 was handed back: leave
                    }
                    MCL_PTIME(17)
                    // every thread's minimum (read only when a winner has to be re-evaluated) and one row per warp (what the decision
                    // warp reads every step): (clock, slot, the lane that reported it)
                    team_best[par * NTS + tid] = make_float2(best, __int_as_float(bslot));
                    if (lane == 0) team_row[par * NW + warp] = make_int4(__float_as_int(wv), ws_, wl_, 0);
                    named_arrive(BAR_FULL + par, NT);                     // (barrier instructions order the shared-memory accesses before them)
                    MCL_PTIME(18)
                }
            } else {
                // ===== the decision warp
                // lane i: the i-th slot re-targeted in this step / 1 .. DEPTH steps ago (written while the sweep at hand may have run)
                int f_cur = -1, f_old[DEPTH];
#pragma unroll
                for (int i = 0; i < DEPTH; i++) f_old[i] = -1;
                int nf_cur = 0;
                int k = k0;
                unsigned ev_acc = 0u;                 // events since the histogram cursor last moved (monotonic axes)
                auto flush_events = [&]() {
                    const int b = hbin_next - 1;
                    if (ev_acc && lane == 0 && b >= 0 && b < p.hist.n_bins) atomicAdd(&p.hist_events[(size_t)hrow * p.hist.n_bins + b], (unsigned long long)ev_acc);
                    ev_acc = 0u;
                };
                for (;;) {
                    const int par = k & (DEPTH - 1);
                    named_sync(BAR_FULL + par, NT);
                    MCL_PTIME(20)
#ifdef MCL_PIPE_STATS
                    const long long dw0_ = clock64();
#endif
                    int hand_back = 0;
                    if (!T_const) { set_T(t_cur); if (has_cb) hand_back = 7; }      // (a ramp: the sweep team's clocks carry no temperature)
                    float v = F_INF; int s_ = -1, l_ = 0;             // lane w < DW: the row of team warp w
                    if (lane < DW) { const int4 row = team_row[par * NW + lane]; v = __int_as_float(row.x); s_ = row.y; l_ = row.z; }
#pragma unroll
                    for (int i = DEPTH - 1; i > 0; i--) f_old[i] = f_old[i - 1];
                    f_old[0] = f_cur; f_cur = -1; nf_cur = 0;
                    uint32_t fresh = 0u;              // bit w: thread (w, my lane) of the team was re-evaluated in this step, from the state as it is
                    float vraw; int smin;
                    for (;;) {
                        vraw = warp_min_f32(v);
                        const unsigned m = __ballot_sync(FULLM, v == vraw);
                        const int wsrc = m ? (__ffs(m) - 1) : 0;                      // the team warp that reported the minimum ...
                        smin = __shfl_sync(FULLM, s_, wsrc);
                        const int tl = __shfl_sync(FULLM, l_, wsrc);                  // ... and its lane
                        if (smin < 0) break;                                          // no clock anywhere
                        const bool is_fresh = (__shfl_sync(FULLM, fresh, tl) >> wsrc) & 1u;
                        bool again = false;
                        if (!(cr[smin] < F_INF)) {
                            again = true;                                             // retired while the sweep ran
                        } else if (!hole_alive(near[smin])) {
                            // a stale electron wins with its lower bound: re-target it now
                            bool ok = true;
                            if (lane == 0) ok = retarget_from_list(smin, -1, false);
                            ok = __shfl_sync(FULLM, (int)ok, 0) != 0;
                            if (!ok) {
                                const unsigned long long b = warp_nearest(H, ex[smin], ey[smin], ez[smin], lane);
                                if (lane == 0) { cr[smin] = sqrtf(__uint_as_float((uint32_t)(b >> 32))); near[smin] = (NearT)(uint32_t)b; }
                                MCL_PSTAT(9, 1)
                            }
                            __syncwarp();
                            if (lane == nf_cur) f_cur = smin;
                            nf_cur++;
                            again = true;
                            MCL_PSTAT(8, 1)
                        } else if (!is_fresh) {
#pragma unroll
                            for (int i = 0; i < DEPTH; i++)
                                for (unsigned mm = __ballot_sync(FULLM, f_old[i] >= 0); mm; mm &= mm - 1) again |= (smin == __shfl_sync(FULLM, f_old[i], __ffs(mm) - 1));
                        }
                        if (!again || hand_back) break;
                        if (nf_cur >= 32) { hand_back = 5; break; }
                        // the clocks of the thread that reported it, from the state as it is (one chunk per lane), then that
                        // warp's row again from its threads' entries
                        const int t = 32 * wsrc + tl;
                        float b_ = F_INF; int bs_ = -1;
                        sweep(t + lane * NTS, 32 * NTS, k, b_, bs_);
                        const float tv = warp_min_f32(b_);
                        const unsigned m2 = __ballot_sync(FULLM, b_ == tv);
                        const int ts_ = __shfl_sync(FULLM, bs_, m2 ? (__ffs(m2) - 1) : 0);
                        if (lane == tl) { team_best[par * NTS + t] = make_float2(tv, __int_as_float(ts_)); fresh |= 1u << wsrc; }
                        __syncwarp();
                        const float2 e = team_best[par * NTS + 32 * wsrc + lane];
                        const float wv = warp_min_f32(e.x);
                        const unsigned m3 = __ballot_sync(FULLM, e.x == wv);
                        const int wl_ = m3 ? (__ffs(m3) - 1) : 0;
                        const int ws_ = __shfl_sync(FULLM, __float_as_int(e.y), wl_);
                        if (lane == wsrc) { v = wv; s_ = ws_; l_ = wl_; }
                        MCL_PSTAT(6, 1)
                    }
                    MCL_PTIME(22)
                    // ---------------- decision: the expressions of the loop in order (dt_fill = inf, no fill)
                    const float vmin = (one && !has_cb) ? vraw - A1 : vraw;
                    const int hmin = smin >= 0 ? (int)near[smin] : -1;
                    const float dt_rec0 = ex2_fast(vmin) * LN2F;               // n_e > 0 in this loop
                    const float dt = fminf(dt_rec0, dt_cap);
                    const bool is_rec = (dt == dt_rec0);
                    const double t_new = t_cur + (double)dt;
                    if (!hand_back) {
                        if (!(dt < 3.0e12f) || smin < 0) hand_back = 4;                                   // the filling clock could matter
                        else if (!(t_new < S.duration) || k + 1 >= p.max_steps) hand_back = 2;            // last step of the leg
                        else if (is_rec && ((n_slots - n_e + 1) * TOMB_DIV > n_slots || n_e - 1 < 4 * NT)) hand_back = 3;   // compaction due
                    }
                    if (hand_back) {
                        MCL_PSTAT(hand_back, 1)
                        __threadfence_block();
                        if (lane == 0) *s_done = -(k + 1);
                        // the team is up to DEPTH sweeps ahead: its entries of steps k+1 .. k+DEPTH-1 are on their way
                        for (int i = 1; i < DEPTH; i++) named_sync(BAR_FULL + ((k + i) & (DEPTH - 1)), NT);
                        break;
                    }
                    // ---------------- apply
                    es32 += (uint32_t)n_e;
                    if (es32 > 0xC0000000u) { esteps += es32; es32 = 0u; }
                    const int n_before = n_e;
                    if (is_rec) {
                        const int h = hmin;
                        if (lane == 0) { cr[smin] = F_INF; near[smin] = (NearT)NEAR_DEAD; hole_bm[h >> 5] &= ~(1u << (h & 31)); }
                        n_e--;
                    }
                    __threadfence_block();                  // (only shared-memory stores are pending here: the global ones follow the flag)
                    if (lane == 0) *s_done = k + 1;
                    MCL_PTIME(23)
                    if (is_rec && lane == 0) hpos[hmin].x = DEAD_X;       // read by this warp's grid searches and after the pipeline only
                    // ---------------- fused histograms (off the sweep team's critical path)
                    if (hedge_next <= t_new) {
                        flush_events();                     // the cursor moves: the events counted so far sit in the bin before it
                        while (hedge_next <= t_new) {
                            if (lane == 0 && p.hist_occ && hbin_next < p.hist.n_bins) {
                                size_t q = (size_t)hrow * p.hist.n_bins + hbin_next;
                                atomicAdd(&p.hist_occ[q], (unsigned long long)n_before);
                                if (p.hist_occ_sq) atomicAdd(&p.hist_occ_sq[q], (unsigned long long)n_before * (unsigned long long)n_before);
                            }
                            hbin_next++;
                            hedge_next = hbin_next <= p.hist.n_bins ? edge_after(hbin_next, hedge_next) : CUDART_INF;
                        }
                    }
                    t_cur = t_new;
                    if (is_rec && hist_on && p.hist_events) {
                        if (h_mono) ev_acc++;               // one atomic per bin, not per event
                        else if (lane == 0) { const int b = bin_of(t_cur); if (b >= 0 && b < p.hist.n_bins) atomicAdd(&p.hist_events[(size_t)hrow * p.hist.n_bins + b], 1ull); }
                    }
                    k++;
#ifdef MCL_PIPE_STATS
                    if (lane == 0) { const long long c_ = clock64() - dw0_; atomicAdd(&g_pipe_h[0][pipe_bucket(c_)], 1ull); atomicAdd(&g_pipe_h[1][pipe_bucket(c_)], (unsigned long long)c_); }
#endif
                    MCL_PTIME(24)
                    MCL_PSTAT(0, 1)
                }
                flush_events();
                if (lane == 0) {
                    s_pipe_d[0] = t_cur; s_pipe_d[1] = hedge_next; s_pipe_ll = esteps;
                    s_pipe_i[0] = hbin_next; s_pipe_i[1] = n_e; s_pipe_i[2] = k; s_pipe_i[3] = (int)es32;
                }
            }
            cta_sync<NT>();
            t_cur = s_pipe_d[0]; hedge_next = s_pipe_d[1]; esteps = s_pipe_ll;
            hbin_next = s_pipe_i[0]; n_e = s_pipe_i[1]; rec_i = s_pipe_i[2]; es32 = (uint32_t)s_pipe_i[3];
            draws_valid = false;
            // ---------------- on the way out: every electron still cached on a dead hole is re-targeted (the loop in order keeps
            // no stale electron), then the sharing masks are counted afresh (team_best may have lived in ref4[])
            for (int base = 0; base < n_slots; base += NT) {
                const int sl = base + tid;
                bool ok = true;
                if (sl < n_slots) { const uint32_t nn = near[sl]; if (nn != NEAR_DEAD && !hole_alive(nn)) ok = retarget_from_list(sl, -1, false); }
                for (unsigned need = __ballot_sync(FULLM, !ok); need; need &= need - 1) {
                    const int src = __ffs(need) - 1;
                    const int sl2 = __shfl_sync(FULLM, sl, src);
                    const unsigned long long b = warp_nearest(H, ex[sl2], ey[sl2], ez[sl2], lane);
                    if (lane == src) { cr[sl2] = sqrtf(__uint_as_float((uint32_t)(b >> 32))); near[sl2] = (NearT)(uint32_t)b; }
                }
            }
            if (share_bm) {
                for (int w = tid; w < cfg.bm_words + cfg.ref_words; w += NT) multi_bm[w] = 0u;
                cta_sync<NT>();
                for (int sl = tid; sl < n_slots; sl += NT) { const uint32_t j = near[sl]; if (j != NEAR_DEAD) mark_target(j, sl); }
            }
            cta_sync<NT>();
            return rec_i - k0;
        };

        {
            const bool fast_ok = MCL_FAST_LOOP && cfg.fast && !lab && !trace && !verify_skip && !dose_on && !ever_filled;
            int rc = 1;
            if (fast_ok) {
                // Pipelined while it pays: the pipeline hands back before every step it cannot take (compaction due, end of
                // the leg, ...), that ONE step runs in order, and it starts again; a leg on which it keeps handing back after
                // a few steps (a depleted box whose events re-target dozens of electrons) finishes in order.  (One call site
                // per loop: they must stay inlined, their state in registers.)
                const bool use_pipe = PIPE_K && cfg.pipe && (T_const || A1 == A2) && (!REGRID || lists_valid_);
                int short_runs = 0;
                for (;;) {
                    int budget = -1;
                    if (use_pipe && !T_const) set_T(t_cur);
                    if (use_pipe && (T_const || !has_cb) && n_e >= 4 * NT && short_runs < 8) {
                        const int done = pipe_loop();
                        short_runs = done < 8 ? short_runs + 1 : 0;
                        budget = 1;
                    }
                    rc = step_loop(std::true_type{}, budget);
                    if (rc != 2) break;
                }
            }
            if (rc == 1) step_loop(std::false_type{}, -1);
        }
        t_off += t_cur;
        if (lab) break;
    }
#ifdef MCL_PROFILE_SKEW
    if (lane == 0) {
        atomicAdd(&g_prof[0][warp], (unsigned long long)pf_sweep); atomicAdd(&g_prof[1][warp], (unsigned long long)pf_wait);
        atomicAdd(&g_prof[2][warp], (unsigned long long)pf_rest); atomicAdd(&g_prof[3][warp], (unsigned long long)pf_steps);
        for (int i = 0; i < 5; i++) atomicAdd(&g_prof[4 + i][warp], (unsigned long long)pf_part[i]);
    }
#endif
    flush_records(rec_i & 31);
    if (status == MCL_OK && rp.protocol == MCL_PROTO_TL_LAB && rec_i == 0) status = MCL_ERR_NOEVENT;
    cta_sync<NT>();
    if (s_err && status == MCL_OK) status = MCL_ERR_INTERNAL;
    if (tid == 0) {
        const int r_out = replica_index();
        if (p.steps_used) p.steps_used[r_out] = rec_i;
        if (p.final_n_e) p.final_n_e[r_out] = n_e;
        if (p.esteps) p.esteps[r_out] = esteps + (long long)es32;
        if (p.consumed) p.consumed[r_out] = 0;
        if (p.status) p.status[r_out] = status;
    }
}


}  // namespace

int philox_max_slots() { return 24000; }

// Test hook (include/mcl_b200.h): the kernel's exponential draw -lg2(u01(word)) for given Philox words, so that the
// small-waiting-time tail of the SFU logarithm can be pinned against float64 on the device it runs on.
namespace {
__global__ void exp_draw_kernel(const uint32_t *w, int n, float *neg_lg2_u, float *lg2_of_that)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float e2 = -lg2_fast(u01(w[i]));
    neg_lg2_u[i] = e2;
    lg2_of_that[i] = lg2_fast(e2);
}
}  // namespace

extern "C" int mcl_debug_exp_draws(const uint32_t *words, int32_t n, float *neg_lg2_u, float *lg2_of_that)
{
    if (!words || n <= 0 || !neg_lg2_u || !lg2_of_that) { set_error("mcl_debug_exp_draws: null argument"); return MCL_ERR_ARG; }
    uint32_t *dw = nullptr; float *da = nullptr, *db = nullptr;
    int rc = MCL_OK;
    if (cudaMalloc(&dw, sizeof(uint32_t) * (size_t)n) != cudaSuccess || cudaMalloc(&da, sizeof(float) * (size_t)n) != cudaSuccess ||
        cudaMalloc(&db, sizeof(float) * (size_t)n) != cudaSuccess) { set_error("mcl_debug_exp_draws: cudaMalloc failed"); rc = MCL_ERR_ALLOC; }
    if (rc == MCL_OK) {
        cudaMemcpy(dw, words, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice);
        exp_draw_kernel<<<(n + 255) / 256, 256>>>(dw, n, da, db);
        cudaMemcpy(neg_lg2_u, da, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaMemcpy(lg2_of_that, db, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("mcl_debug_exp_draws: %s", cudaGetErrorString(e)); rc = MCL_ERR_CUDA; }
    }
    cudaFree(dw); cudaFree(da); cudaFree(db);
    return rc;
}

#ifdef MCL_PIPE_STATS
extern "C" int mcl_debug_pipe_stats(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out, g_pipe, sizeof(unsigned long long) * 32);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out + 32, g_pipe_h, sizeof(unsigned long long) * 32);
    if (reset) { static unsigned long long z[32]; cudaMemcpyToSymbol(g_pipe, z, sizeof(z)); cudaMemcpyToSymbol(g_pipe_h, z, sizeof(z)); }
    return (int)e;
}
#endif

#ifdef MCL_PROFILE_SKEW
extern "C" int mcl_debug_prof(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * 10 * 32);
    if (reset) { static unsigned long long z[10 * 32]; cudaMemcpyToSymbol(g_prof, z, sizeof(z)); }
    return (int)e;
}
#endif

static int grid_edge_max(int n_h0_max)
{
    int g = (int)floor(cbrt((double)n_h0_max / 3.0)) + 1;
    return g < 1 ? 1 : g;
}

struct PhiloxPlan {
    int nt; int cap_slots; int g_max; int cap_cells; int bm_words; int share_bm; int ref_words; size_t smem; size_t off_holes; size_t off_cand;
    size_t off_hpos2, off_hmap; size_t stride; bool near16; int off_pipe;
    bool slab_smem; int sm_hpos, sm_exyz, sm_cstart, sm_cfill;
};

static int g_nt_override = 0;
void philox_set_block_threads(int nt) { g_nt_override = nt; }

static PhiloxPlan make_plan(int cap_e, int cap_h, int nt_override, int n_replicas = 1 << 30, bool with_regrid = true)
{
    PhiloxPlan pl;
    pl.cap_slots = (int)align_up((size_t)cap_e + 2, 64);
    pl.g_max = grid_edge_max(cap_h);          // n_h0 <= cap_h: always enough room for the cell tables
    pl.cap_cells = pl.g_max * pl.g_max * pl.g_max + 1;
    pl.near16 = cap_h <= 65534;
    if (const char *env = getenv("MCL_PHILOX_NEAR32")) { if (atoi(env) == 1) pl.near16 = false; }   // test knob: force the 32-bit slot path
    pl.bm_words = (cap_h + 31) / 32;
    pl.smem = (size_t)pl.cap_slots * (4 + (pl.near16 ? 2 : 4)) + 4 * (size_t)pl.bm_words;
    pl.off_holes = align_up(sizeof(float) * 3 * (size_t)cap_e, 16);
    size_t b = pl.off_holes + 16 * (size_t)cap_h + sizeof(int) * (2 * (size_t)pl.cap_cells + (size_t)cap_e);
    pl.off_cand = align_up(b, 16);
    pl.off_hpos2 = align_up(pl.off_cand + (size_t)cap_e * (16 + 4 * (pl.near16 ? 2 : 4)), 16);
    pl.off_hmap = pl.off_hpos2 + (with_regrid ? 16 * (size_t)cap_h : 0);
    pl.stride = align_up(pl.off_hmap + (with_regrid ? 4 * (size_t)cap_h : 0), 256);
    // CTA width: throughput optimum measured on B200 (2000 electrons: 64 threads, 10^4: 256).  Results do not
    // depend on it.  With fewer replicas than SMs the launch is latency-bound: widen the CTAs instead.
    int nt;
    if (cap_e <= 256) nt = 32;
    else if (cap_e <= 3072) nt = 64;
    else if (cap_e <= 6144) nt = 128;
    else nt = 256;
    if (n_replicas <= 148) nt = nt * 4 > 512 ? 512 : nt * 4;
    else if (n_replicas <= 296) nt = nt * 2 > 512 ? 512 : nt * 2;
    if (const char *env = getenv("MCL_PHILOX_NT")) nt_override = atoi(env);     // tuning knob
    if (nt_override == 32 || nt_override == 64 || nt_override == 128 || nt_override == 256 || nt_override == 512)
        nt = nt_override;
    pl.nt = nt;
    // large boxes have the shared memory to spare (they are limited to 3 CTAs per SM either way)
    pl.share_bm = nt >= 256 ? 1 : 0;          // (the kernels narrower than 256 threads are compiled without the masks)
    if (const char *env = getenv("MCL_PHILOX_SHARE_BM")) { int v = atoi(env); pl.share_bm = v < 0 ? 0 : (v > 2 ? 2 : v); }   // knob: 0 off, 1 on, 2 on + self-check
    pl.ref_words = (cap_h + 7) / 8;
    if (nt < 256) pl.share_bm = 0;
    // pipelined loop: (clock, slot) of every sweep-team thread.  It is idle whenever ref4[] is in use and the other way
    // round (the masks are counted afresh when the pipeline hands back), so a large enough ref4[] doubles as its storage --
    // 10^4-electron boxes have no shared memory to spare at three CTAs per SM.
    pl.off_pipe = 0;
    if (nt >= 128) {
        const size_t need = 8 * MCL_PIPE_DEPTH * (size_t)nt;
        const size_t ref_off = align_up(pl.smem + 4 * (size_t)pl.bm_words, 8);
        if (pl.share_bm && ref_off + need <= pl.smem + 4 * ((size_t)pl.bm_words + (size_t)pl.ref_words)) pl.off_pipe = (int)ref_off;
        else { pl.off_pipe = (int)align_up(pl.smem + (pl.share_bm ? 4 * ((size_t)pl.bm_words + (size_t)pl.ref_words) : 0), 8); }
    }
    if (pl.share_bm) pl.smem += 4 * ((size_t)pl.bm_words + (size_t)pl.ref_words);
    if (nt >= 128 && (size_t)pl.off_pipe + 8 * MCL_PIPE_DEPTH * (size_t)nt > pl.smem) pl.smem = (size_t)pl.off_pipe + 8 * MCL_PIPE_DEPTH * (size_t)nt;
    // Small boxes (one warp per replica): hole table, cell tables and electron coordinates in shared memory when at least
    // four such CTAs fit an SM.  MCL_PHILOX_SMEM_SLAB=0 keeps them in the HBM slab (same results; test knob).
    pl.slab_smem = false; pl.sm_hpos = pl.sm_exyz = pl.sm_cstart = pl.sm_cfill = 0;
    if (nt == 32) {
        size_t o = align_up(pl.smem, 16);
        const size_t o_hpos = o; o += 16 * (size_t)cap_h;
        const size_t o_exyz = o; o += 12 * (size_t)cap_e;
        const size_t o_cstart = o; o += 4 * (size_t)pl.cap_cells;
        const size_t o_cfill = o; o += 4 * (size_t)pl.cap_cells;
        bool want = o <= 56 * 1024;
        if (const char *env = getenv("MCL_PHILOX_SMEM_SLAB")) want = want && atoi(env) != 0;
        if (want) {
            pl.slab_smem = true; pl.smem = o;
            pl.sm_hpos = (int)(o_hpos / 4); pl.sm_exyz = (int)(o_exyz / 4); pl.sm_cstart = (int)(o_cstart / 4); pl.sm_cfill = (int)(o_cfill / 4);
        }
    }
    return pl;
}

size_t philox_ws_stride(int cap_e, int cap_h, bool with_regrid)
{
    return make_plan(cap_e, cap_h, 0, 1 << 30, with_regrid).stride;
}

template <int NT, int MINB, typename NearT, int PPC, bool SLAB_SMEM, bool REGRID>
static cudaError_t launch_two(const LaunchParams &p, const RoundKeys &K, const Cfg &cfg, size_t smem, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(philox_kernel<NT, MINB, NearT, PPC, SLAB_SMEM, REGRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    philox_kernel<NT, MINB, NearT, PPC, SLAB_SMEM, REGRID><<<p.n_launch, NT, smem, stream>>>(p, K, cfg);
    return cudaGetLastError();
}
template <int NT, int MINB, typename NearT, int PPC, bool SLAB_SMEM = false>
static cudaError_t launch_one(const LaunchParams &p, const RoundKeys &K, const Cfg &cfg, size_t smem, cudaStream_t stream)
{
    if constexpr (SLAB_SMEM) {
        return launch_two<NT, MINB, NearT, PPC, true, true>(p, K, cfg, smem, stream);
    } else {
        if (p.with_regrid) return launch_two<NT, MINB, NearT, PPC, false, true>(p, K, cfg, smem, stream);
        return launch_two<NT, MINB, NearT, PPC, false, false>(p, K, cfg, smem, stream);
    }
}

cudaError_t launch_philox(const LaunchParams &p, cudaStream_t stream, int /*max_slots*/)
{
    if (p.n_launch <= 0) return cudaSuccess;
    PhiloxPlan pl = make_plan(p.cap_e, p.cap_h, g_nt_override, p.n_launch, p.with_regrid != 0);
    int fast = 1;
    if (const char *env = getenv("MCL_PHILOX_FAST")) fast = atoi(env) != 0;        // knob: 0 = general step loop only
    int fill_extra = kFillExtra, relist = 1;
    if (const char *env = getenv("MCL_PHILOX_FILL_EXTRA")) { int v = atoi(env); fill_extra = v > kFillExtra ? kFillExtra : v; }   // test knob: regrid early (may be negative)
    int pipe = 1;
    if (const char *env = getenv("MCL_PHILOX_PIPE")) pipe = atoi(env) != 0;        // knob: 0 = isothermal legs run the specialised loop in order
    if (const char *env = getenv("MCL_PHILOX_RELIST")) relist = atoi(env) != 0;    // knob: 0 = read-out legs after fills keep searching the grid
    Cfg cfg{pl.cap_slots, pl.g_max, pl.cap_cells, pl.bm_words, pl.share_bm, pl.ref_words, pl.off_holes, pl.off_cand, fast,
            pl.off_hpos2, pl.off_hmap, fill_extra, p.with_regrid != 0, relist, pipe, pl.off_pipe, pl.sm_hpos, pl.sm_exyz, pl.sm_cstart, pl.sm_cfill};
    const RoundKeys K = make_round_keys(p.seed);
    // MINB caps the register count at 64 per thread (32 resident warps per SM when smem allows)
#define MCL_CASE(NT_, MINB_, PPC_)                                                                  \
    case NT_:                                                                                      \
        return pl.near16 ? launch_one<NT_, MINB_, uint16_t, PPC_>(p, K, cfg, pl.smem, stream)      \
                         : launch_one<NT_, MINB_, uint32_t, PPC_>(p, K, cfg, pl.smem, stream)
#if defined(MCL_ONLY_NT)    // build-time probe: ONE instantiation (-DMCL_ONLY_NT=96 -DMCL_ONLY_MINB=6), for quick A/B builds; run it with MCL_PHILOX_NT set to the same width
    return launch_two<MCL_ONLY_NT, MCL_ONLY_MINB, uint16_t, 2, false, false>(p, K, cfg, pl.smem, stream);
#elif defined(MCL_ONLY_C2)  // build-time probe (scripts/regs_probe.sh): only the BASELINE C2 instantiation, for quick ptxas -v runs
    return launch_two<256, MCL_NT256_MINB, uint16_t, 2, false, false>(p, K, cfg, pl.smem, stream);
#else
    if (pl.slab_smem)       // (implies nt == 32) few CTAs per SM: no register cap worth the name
        return pl.near16 ? launch_one<32, 12, uint16_t, 2, true>(p, K, cfg, pl.smem, stream)
                         : launch_one<32, 12, uint32_t, 2, true>(p, K, cfg, pl.smem, stream);
    switch (pl.nt) {
        MCL_CASE(32, MCL_NT32_MINB, 2);
        MCL_CASE(64, MCL_NT64_MINB, 2);
        MCL_CASE(128, MCL_NT128_MINB, 2);
        MCL_CASE(256, MCL_NT256_MINB, 2);
        default: break;
    }
    return pl.near16 ? launch_one<512, 2, uint16_t, 2>(p, K, cfg, pl.smem, stream)
                     : launch_one<512, 2, uint32_t, 2>(p, K, cfg, pl.smem, stream);
#endif
#undef MCL_CASE
}

}  // namespace mcl
