// mcl_replay.cu -- FP64 replay kernel: consumes the reference's uniform draws in the reference's
// order and reproduces its integer event / n_e traces bit-for-bit (parity mode, not the
// throughput path).
//
// One CTA per replica.  State lives in the replica's scratch slab in HBM (float64, compact arrays
// exactly like the reference's NumPy arrays, so every index means what it means there):
//   ex,ey,ez,min_d [cap_e] f64 | nearest [cap_e] i32 | hx,hy,hz [cap_h] f64 | flagged [cap_e] i32
// Reference semantics restated (file:line relative to the reference root):
//   Box.seed / _rebuild              src/class/engine.py:113-129
//   Box.add_electron (stale cache)   src/class/engine.py:133-152
//   Box.remove_pair (shift, then mask: old index h AND old index h+1 rescan)  :154-175
//   Physics.rate_cb / rate_tunnel / lifetime    src/class/engine.py:65-77
//   _update_lifetimes / _filling_time           src/class/tl_trap_lab.py:48-60
//   simulate() loop                             src/class/simulate.py:46-92
//   TL_lab / ISO_lab loops                      src/class/tl_trap_lab.py:75-111,135-172
// All arithmetic that feeds a comparison uses explicit round-to-nearest intrinsics so that nvcc
// cannot contract a multiply into an add (NumPy never does).
#include <math_constants.h>
#include "mcl_common.cuh"

namespace mcl {

namespace {

constexpr int RT = 256;                // threads per CTA
constexpr int RW = RT / 32;

struct MinIdx { double v; int i; };

__device__ __forceinline__ bool better(double av, int ai, double bv, int bi)
{
    // np.argmin: first minimum.  Empty lanes carry i == INT_MAX.
    return (av < bv) || (av == bv && ai < bi) || (bi == 0x7fffffff && ai != 0x7fffffff);
}

__device__ MinIdx block_argmin(double v, int i, double *sv, int *si)
{
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, v, o);
        int oi = __shfl_down_sync(0xffffffffu, i, o);
        if (better(ov, oi, v, i)) { v = ov; i = oi; }
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                   // protect sv/si from the previous use
    if (l == 0) { sv[w] = v; si[w] = i; }
    __syncthreads();
    v = sv[0]; i = si[0];
    for (int k = 1; k < RW; k++) if (better(sv[k], si[k], v, i)) { v = sv[k]; i = si[k]; }
    return MinIdx{v, i};
}

__device__ __forceinline__ double dist(double ax, double ay, double az, double bx, double by, double bz)
{
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

struct Box {
    double *ex, *ey, *ez, *min_d, *hx, *hy, *hz;
    int *nearest, *flagged;
    int n_e, n_h;
};

// whole-CTA scan of one electron against holes [0, n_h): first minimum
__device__ MinIdx scan_holes(const Box &bx, double x, double y, double z, int n_h, double *sv, int *si)
{
    double best = CUDART_INF; int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < n_h; j += RT) {
        double d = dist(x, y, z, bx.hx[j], bx.hy[j], bx.hz[j]);
        if (bi == 0x7fffffff || d < best) { best = d; bi = j; }
    }
    return block_argmin(best, bi, sv, si);
}

// order-preserving delete of element `pos` from arrays of length n (np.delete): chunked so that
// every chunk reads before it writes
template <typename T>
__device__ void shift_down(T *a, int pos, int n)
{
    for (int base = pos; base < n - 1; base += RT) {
        int i = base + threadIdx.x;
        T v = T();
        if (i < n - 1) v = a[i + 1];
        __syncthreads();
        if (i < n - 1) a[i] = v;
        __syncthreads();
    }
}

struct Stream {
    const double *u; long long pos, end; bool bad;
    __device__ __forceinline__ bool have(long long n) { if (pos + n > end) { bad = true; return false; } return true; }
};

__global__ void __launch_bounds__(RT) replay_kernel(LaunchParams p)
{
    const int r = blockIdx.x;
    const int tid = threadIdx.x;
    const mcl_replica rp = p.replicas[r];
    __shared__ double sv[RW];
    __shared__ int si[RW];
    __shared__ int s_nflag;
    __shared__ int s_any;

    unsigned char *ws = p.ws + (size_t)r * p.ws_stride;
    Box bx;
    {
        size_t ce = (size_t)p.cap_e, ch = (size_t)p.cap_h;
        double *d = reinterpret_cast<double *>(ws);
        bx.ex = d; bx.ey = d + ce; bx.ez = d + 2 * ce; bx.min_d = d + 3 * ce;
        bx.hx = d + 4 * ce; bx.hy = d + 4 * ce + ch; bx.hz = d + 4 * ce + 2 * ch;
        int *ip = reinterpret_cast<int *>(d + 4 * ce + 3 * ch);
        bx.nearest = ip; bx.flagged = ip + ce;
    }
    Stream st;
    st.u = p.replay_u; st.pos = p.replay_off[r]; st.end = p.replay_off[r + 1]; st.bad = false;
    const long long pos0 = st.pos;

    int status = MCL_OK;
    int rec_i = 0;
    long long esteps = 0;
    const size_t rec_base = (size_t)r * (size_t)p.max_steps;

    const double core = rp.side, bnd = __dmul_rn(rp.side, rp.boundary_factor);
    bx.n_e = rp.n_e0; bx.n_h = rp.n_h0;
    if (bx.n_e > p.cap_e || bx.n_h > p.cap_h) status = MCL_ERR_CAPACITY;

    // ---- Box.seed: electrons first, then holes, row-major (engine.py:125-128)
    if (status == MCL_OK && !st.have(3LL * (bx.n_e + bx.n_h))) status = MCL_ERR_STREAM;
    if (status == MCL_OK) {
        for (int i = tid; i < bx.n_e; i += RT) {
            bx.ex[i] = __dmul_rn(st.u[st.pos + 3LL * i + 0], core);
            bx.ey[i] = __dmul_rn(st.u[st.pos + 3LL * i + 1], core);
            bx.ez[i] = __dmul_rn(st.u[st.pos + 3LL * i + 2], core);
        }
        st.pos += 3LL * bx.n_e;
        for (int j = tid; j < bx.n_h; j += RT) {
            bx.hx[j] = __dmul_rn(st.u[st.pos + 3LL * j + 0], bnd);
            bx.hy[j] = __dmul_rn(st.u[st.pos + 3LL * j + 1], bnd);
            bx.hz[j] = __dmul_rn(st.u[st.pos + 3LL * j + 2], bnd);
        }
        st.pos += 3LL * bx.n_h;
        __syncthreads();
        // ---- _rebuild: thread-per-electron brute force (engine.py:113-119)
        if (bx.n_e > 0 && bx.n_h <= 0) status = MCL_ERR_NOHOLES;
        else
            for (int i = tid; i < bx.n_e; i += RT) {
                double x = bx.ex[i], y = bx.ey[i], z = bx.ez[i];
                double best = dist(x, y, z, bx.hx[0], bx.hy[0], bx.hz[0]); int bi = 0;
                for (int j = 1; j < bx.n_h; j++) {
                    double d = dist(x, y, z, bx.hx[j], bx.hy[j], bx.hz[j]);
                    if (d < best) { best = d; bi = j; }
                }
                bx.min_d[i] = best; bx.nearest[i] = bi;
            }
        __syncthreads();
    }

    const bool lab = rp.protocol != MCL_PROTO_SIMULATE;
    const bool iso = rp.protocol == MCL_PROTO_ISO_LAB;
    const double *obs = p.obs_time + rp.obs_begin;
    int obs_idx = 0;
    double t_off = 0.0;

    // clocks of the coming step: only min / argmin / any of the waits are ever used
    double wmin = 0.0; int warg = -1; int wany = 0;
    double tf = 0.0;

    // _update_lifetimes(T): n selectors then n exponentials (engine.py:72, tl_trap_lab.py:51)
    auto draw_waits = [&](double T, double A_opt) -> bool {
        const int n = bx.n_e;
        if (!st.have(2LL * n)) return false;
        const double kT = __dmul_rn(rp.k_b, T);
        const double k_cb = __dmul_rn(rp.s, exp(__ddiv_rn(-rp.E_cb, kT)));
        double best = CUDART_INF; int bi = 0x7fffffff; int any = 0;
        for (int i = tid; i < n; i += RT) {
            double us = st.u[st.pos + i];
            double ue = st.u[st.pos + n + i];
            double E_loc = (us < rp.Retrap) ? rp.E_loc_2 : rp.E_loc_1;
            double k_tun;
            if (A_opt == 0.0)
                k_tun = __dmul_rn(rp.b, exp(__dsub_rn(__ddiv_rn(-E_loc, kT), __dmul_rn(rp.alpha, bx.min_d[i]))));
            else   // extension (parity unpinned): optical excitation into the tunnelling state
                k_tun = __dmul_rn(__dadd_rn(A_opt, __dmul_rn(rp.b, exp(__ddiv_rn(-E_loc, kT)))),
                                  exp(-__dmul_rn(rp.alpha, bx.min_d[i])));
            double tau = __ddiv_rn(1.0, __dadd_rn(k_cb, k_tun));
            double w = __dmul_rn(tau, -log(__dsub_rn(1.0, ue)));
            if (w != 0.0) any = 1;
            if (bi == 0x7fffffff || w < best) { best = w; bi = i; }
        }
        st.pos += 2LL * n;
        if (tid == 0) s_any = 0;
        MinIdx m = block_argmin(best, bi, sv, si);      // its barriers order the s_any reset
        if (any) atomicOr(&s_any, 1);
        __syncthreads();
        wmin = m.v; warg = (m.i == 0x7fffffff) ? -1 : m.i; wany = s_any;
        __syncthreads();                                // nobody still reads s_any when it is reset
        return true;
    };
    // _filling_time(D): one uniform whenever lam > 0 (tl_trap_lab.py:53-60)
    auto draw_fill = [&](double D) -> bool {
        double lam = (bx.n_e == rp.N_e || D == 0.0)
                         ? 1e-20 : __dmul_rn(__ddiv_rn(D, rp.D0), (double)(rp.N_e - bx.n_e));
        if (lam > 0) {
            if (!st.have(1)) return false;
            tf = __dmul_rn(__ddiv_rn(1.0, lam), -log(__dsub_rn(1.0, st.u[st.pos])));
            st.pos += 1;
        } else tf = 1e20;
        return true;
    };

    for (int sg = 0; sg < rp.seg_count && status == MCL_OK; sg++) {
        const mcl_segment S = p.segments[rp.seg_begin + sg];
        const double T0K = __dadd_rn(S.T_start, 273.15);      // lab: row.T_start + 273.15
        double t_cur = 0.0;
        double T_now = T0K;
        if (lab) {      // tl_trap_lab.py:81-84 / :140-141
            if (!draw_waits(T0K, 0.0) || !draw_fill(S.dose_rate)) { status = MCL_ERR_STREAM; break; }
        }
        for (;;) {
            // ---------------- loop condition
            if (!lab) { if (!(t_cur <= S.duration)) break; }
            else if (iso) { if (!(obs_idx < rp.obs_count)) break; }
            else { if (!(t_cur < S.duration)) break; }

            // ---------------- simulate: draw at the top of every step, order sel, exp, fill
            if (!lab) {
                T_now = __dadd_rn(__dadd_rn(S.T_start, __dmul_rn(S.T_rate, t_cur)), 273.15);
                if (!draw_waits(T_now, S.A_opt) || !draw_fill(S.dose_rate)) { status = MCL_ERR_STREAM; break; }
            }

            // ---------------- choose dt
            double dt, dt_recomb;
            bool is_fill, is_recomb;
            if (!lab) {
                dt_recomb = wany ? wmin : tf;                       // simulate.py:59 (.any())
                dt = tf;                                            // python min(dt_fill, dt_recomb, dt_cap)
                if (dt_recomb < dt) dt = dt_recomb;
                if (S.dt_cap < dt) dt = S.dt_cap;
                is_fill = (dt == tf);
                is_recomb = !is_fill && (dt == dt_recomb);
            } else {
                dt_recomb = bx.n_e ? wmin : tf;                     // tl_trap_lab.py:91 (.size)
                dt = (tf < dt_recomb) ? tf : dt_recomb;
                is_fill = (dt == tf);
                is_recomb = !is_fill;
                if (!iso) T_now = __dadd_rn(T0K, __dmul_rn(S.T_rate, __dadd_rn(t_cur, dt)));
            }
            if (rec_i >= p.max_steps) { status = MCL_ERR_STEPS; break; }
            esteps += bx.n_e;
            t_cur = __dadd_rn(t_cur, dt);

            int ev = 0, kd = 0, ei = -1, hi = -1;
            if (is_fill) {
                // ---------------- Box.add_electron (engine.py:133-152)
                kd = 1; ei = bx.n_e; hi = bx.n_h;
                if (bx.n_e + 1 > p.cap_e || bx.n_h + 1 > p.cap_h) { status = MCL_ERR_CAPACITY; break; }
                if (!st.have(6)) { status = MCL_ERR_STREAM; break; }
                if (bx.n_h <= 0) { status = MCL_ERR_NOHOLES; break; }
                double x = __dmul_rn(st.u[st.pos + 0], core), y = __dmul_rn(st.u[st.pos + 1], core),
                       z = __dmul_rn(st.u[st.pos + 2], core);
                double qx = __dmul_rn(st.u[st.pos + 3], bnd), qy = __dmul_rn(st.u[st.pos + 4], bnd),
                       qz = __dmul_rn(st.u[st.pos + 5], bnd);
                st.pos += 6;
                MinIdx m = scan_holes(bx, x, y, z, bx.n_h, sv, si);   // OLD holes only
                if (tid == 0) {
                    int e = bx.n_e, h = bx.n_h;
                    bx.ex[e] = x; bx.ey[e] = y; bx.ez[e] = z;
                    bx.min_d[e] = m.v; bx.nearest[e] = m.i;
                    bx.hx[h] = qx; bx.hy[h] = qy; bx.hz[h] = qz;
                }
                bx.n_e++; bx.n_h++;
                __syncthreads();
            } else if (is_recomb) {
                // ---------------- Box.remove_pair (engine.py:154-175)
                kd = 2; ev = 1; ei = warg; hi = bx.nearest[warg];
                __syncthreads();                       // everyone has read nearest[warg]
                const int e = ei, h = hi;
                shift_down(bx.ex, e, bx.n_e); shift_down(bx.ey, e, bx.n_e); shift_down(bx.ez, e, bx.n_e);
                shift_down(bx.min_d, e, bx.n_e); shift_down(bx.nearest, e, bx.n_e);
                shift_down(bx.hx, h, bx.n_h); shift_down(bx.hy, h, bx.n_h); shift_down(bx.hz, h, bx.n_h);
                bx.n_e--; bx.n_h--;
                if (tid == 0) s_nflag = 0;
                __syncthreads();
                // shift first (:168) THEN mask (:171): old index h and old index h+1 both rescan
                for (int i = tid; i < bx.n_e; i += RT) {
                    int nn = bx.nearest[i];
                    if (nn > h) { nn--; bx.nearest[i] = nn; }
                    if (nn == h) bx.flagged[atomicAdd(&s_nflag, 1)] = i;
                }
                __syncthreads();
                const int nflag = s_nflag;
                if (nflag > 0 && bx.n_h <= 0) { status = MCL_ERR_NOHOLES; break; }
                for (int f = 0; f < nflag; f++) {
                    int i = bx.flagged[f];
                    MinIdx m = scan_holes(bx, bx.ex[i], bx.ey[i], bx.ez[i], bx.n_h, sv, si);
                    if (tid == 0) { bx.min_d[i] = m.v; bx.nearest[i] = m.i; }
                }
                __syncthreads();
            }

            // ---------------- lab: redraw AFTER the event, order fill, sel, exp (tl_trap_lab.py:104-105)
            if (lab) {
                if (!draw_fill(S.dose_rate) || !draw_waits(T_now, 0.0)) { status = MCL_ERR_STREAM; break; }
            }

            // ---------------- record (simulate.py:64,85-89)
            if (tid == 0) {
                size_t q = rec_base + (size_t)rec_i;
                if (p.event) p.event[q] = ev;
                if (p.n_e) p.n_e[q] = bx.n_e;
                if (p.t) p.t[q] = __dadd_rn(t_off, t_cur);
                if (p.kind) p.kind[q] = kd;
                if (p.e_idx) p.e_idx[q] = ei;
                if (p.h_idx) p.h_idx[q] = hi;
            }
            rec_i++;
            if (iso) {
                while (obs_idx < rp.obs_count && t_cur >= obs[obs_idx]) {
                    if (tid == 0 && p.obs_n_e) p.obs_n_e[rp.obs_begin + obs_idx] = bx.n_e;
                    obs_idx++;
                }
            }
            if (!lab && S.duration != 0.0 && t_cur >= S.duration) break;     // simulate.py:91-92
        }
        t_off = __dadd_rn(t_off, t_cur);
        if (lab) break;        // lab protocols have exactly one leg
    }
    if (status == MCL_OK && rp.protocol == MCL_PROTO_TL_LAB && rec_i == 0) status = MCL_ERR_NOEVENT;
    if (st.bad && status == MCL_OK) status = MCL_ERR_STREAM;
    if (tid == 0) {
        if (p.steps_used) p.steps_used[r] = rec_i;
        if (p.final_n_e) p.final_n_e[r] = bx.n_e;
        if (p.esteps) p.esteps[r] = esteps;
        if (p.consumed) p.consumed[r] = st.pos - pos0;
        if (p.status) p.status[r] = status;
    }
}

}  // namespace

size_t replay_ws_stride(int cap_e, int cap_h)
{
    size_t b = sizeof(double) * (4 * (size_t)cap_e + 3 * (size_t)cap_h) + sizeof(int) * 2 * (size_t)cap_e;
    return align_up(b, 256);
}

cudaError_t launch_replay(const LaunchParams &p, cudaStream_t stream)
{
    replay_kernel<<<p.n_replicas, RT, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace mcl
