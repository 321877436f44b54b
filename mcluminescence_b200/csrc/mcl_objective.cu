// mcl_objective.cu -- batched Optimizer objective: S parameter candidates x n_rows lab rows in ONE
// launch of the native kernel, MSE per candidate back on the host.
//
// Replaces, for a whole differential-evolution population at once, the reference chain
//   optimizer.objective -> cfg_with_params -> run_one_sim -> TLTrapSim.TL_lab / ISO_lab
//   (src/class/optimizer.py:49-84, src/class/tl_trap_lab.py:27-43,65-123,125-179).
// Geometry is derived with the reference's own expressions in C doubles (Python's float `**` is
// C pow), so int() truncations agree with the Python host path.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "mcl_common.cuh"

namespace {

// NumPy's float64 add.reduce over a contiguous vector of n < 128 values (pairwise_sum's 8-lane
// unrolled block), so that the mean matches LabTable.mse() / np.mean bit for bit.
double numpy_sum(const double *a, size_t n)
{
    if (n < 8) { double r = 0.0; for (size_t i = 0; i < n; i++) r += a[i]; return r; }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        size_t i;
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    size_t n2 = n / 2; n2 -= n2 % 8;
    return numpy_sum(a, n2) + numpy_sum(a + n2, n - n2);
}


// The scratch slab of an Optimizer population is gigabytes; allocating and freeing it on every call
// costs anything from 1 ms to 0.5 s.  It is kept (grow-only, per device) between calls and handed back
// by mcl_release_scratch().  This is the only state the library keeps.
struct ScratchCache {
    std::mutex mu;
    std::mutex call_mu;          // held for a whole mcl_objective call: calls share the slab, so they are serialised
    void *ptr[64] = {nullptr};
    size_t bytes[64] = {0};
    void *get(size_t need)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
        std::lock_guard<std::mutex> lk(mu);
        if (bytes[dev] < need) {
            if (ptr[dev]) cudaFree(ptr[dev]);
            ptr[dev] = nullptr; bytes[dev] = 0;
            size_t want = need + need / 8;
            if (cudaMalloc(&ptr[dev], want) != cudaSuccess) {
                cudaGetLastError();
                if (cudaMalloc(&ptr[dev], need) != cudaSuccess) { ptr[dev] = nullptr; return nullptr; }
                want = need;
            }
            bytes[dev] = want;
        }
        return ptr[dev];
    }
    void release()
    {
        std::lock_guard<std::mutex> call_lk(call_mu);        // never under a running call
        std::lock_guard<std::mutex> lk(mu);
        int cur = -1;
        const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
        for (int d = 0; d < 64; d++) if (ptr[d]) { cudaSetDevice(d); cudaFree(ptr[d]); ptr[d] = nullptr; bytes[d] = 0; }
        if (have_cur) cudaSetDevice(cur);                       // leave the caller's current device as it was
    }
};
ScratchCache g_scratch;
thread_local float g_last_kernel_ms = 0.f;

}  // namespace

extern "C" int mcl_objective(const double *P, int32_t S, const mcl_lab *lab, uint64_t seed,
                             uint64_t candidate_id0, double *mse, int64_t *esteps_total, void *stream)
{
    using namespace mcl;
    if (!P || S <= 0 || !lab || !mse) { set_error("mcl_objective: null argument"); return MCL_ERR_ARG; }
    if (lab->protocol != MCL_PROTO_TL_LAB && lab->protocol != MCL_PROTO_ISO_LAB) { set_error("mcl_objective: protocol must be TL_LAB or ISO_LAB"); return MCL_ERR_ARG; }
    if (lab->n_rows <= 0 || !lab->rows || !lab->e_ratio_start || !lab->target) { set_error("mcl_objective: lab table incomplete"); return MCL_ERR_ARG; }
    const bool iso = lab->protocol == MCL_PROTO_ISO_LAB;
    const bool legacy = (lab->flags & MCL_LAB_LEGACY) != 0;
    if (legacy && iso) { set_error("mcl_objective: legacy semantics exist for the TL protocol only"); return MCL_ERR_ARG; }
    if (iso && (!lab->obs_begin || !lab->obs_time)) { set_error("mcl_objective: ISO needs obs_begin / obs_time"); return MCL_ERR_ARG; }
    const int n_rows = lab->n_rows;
    const int n_obs = iso ? lab->obs_begin[n_rows] : 0;
    const size_t R = (size_t)S * (size_t)n_rows;
    if (R > 0x7fffffffu) { set_error("mcl_objective: too many replicas"); return MCL_ERR_ARG; }

    // One call at a time: concurrent callers would run kernels in the same cached slab (and a larger request
    // would free it under the other call).  Documented in mcl_b200.h.
    std::lock_guard<std::mutex> call_lock(g_scratch.call_mu);
    // MCL_OBJECTIVE_TIMING=1: wall-clock breakdown of the call on stderr (tables / plan / scratch / launch / wait / mse)
    const bool timing = getenv("MCL_OBJECTIVE_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    double t_ms[6] = {0, 0, 0, 0, 0, 0};
    auto lap = [&](int i) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        t_ms[i] = std::chrono::duration<double, std::milli>(now - t_prev).count();
        t_prev = now;
    };
    std::vector<mcl_replica> reps(R);
    std::vector<mcl_segment> segs(lab->rows, lab->rows + n_rows);
    for (int k = 0; k < n_rows; k++) {
        segs[k].dt_cap = 1e20; segs[k].A_opt = 0.0;
        if (!iso) segs[k].dose_rate = lab->D;                  // tl_trap_lab.py:83
    }
    // every candidate re-uses the same observation table; obs outputs are per replica, so the
    // table is tiled per candidate
    std::vector<double> obs((size_t)n_obs * (size_t)(iso ? S : 0));
    for (int c = 0; c < S; c++) {
        const double rho_prime = P[0 * (size_t)S + c], E_cb = P[1 * (size_t)S + c], E1 = P[2 * (size_t)S + c],
                     E2 = P[3 * (size_t)S + c], D0 = P[4 * (size_t)S + c], s = P[5 * (size_t)S + c],
                     b = P[6 * (size_t)S + c], alpha = P[7 * (size_t)S + c], holes = P[8 * (size_t)S + c],
                     retrap = P[9 * (size_t)S + c];
        const double rho = rho_prime * (3.0 / (4.0 * M_PI) * pow(alpha, 3.0));     // tl_trap_lab.py:33
        const double side = pow(holes / rho, 1.0 / 3.0);                              // :34
        int n_h0 = (int)(holes * pow(lab->boundary_factor, 3.0));                     // engine.py:127
        double side_c = side;
        if (legacy) {
            // initialize_box_bg (src/est_params/functions.py:51-80), same operation order as the Python expressions
            const int h = (int)holes;
            const double d = pow((double)h / rho, 1.0 / 3.0);
            const double lb_ = d * lab->boundary_factor;
            const double vol = d * d * d, vol_b = lb_ * lb_ * lb_;
            const double density = (double)h / (d * d * d);
            n_h0 = h + (int)(density * (vol_b - vol));
            side_c = d;
        }
        for (int k = 0; k < n_rows; k++) {
            mcl_replica &rp = reps[(size_t)c * n_rows + k];
            rp.alpha = alpha; rp.b = b; rp.s = s; rp.E_cb = E_cb; rp.E_loc_1 = E1; rp.E_loc_2 = E2;
            rp.D0 = D0; rp.Retrap = retrap; rp.k_b = lab->k_b; rp.side = side_c;
            rp.boundary_factor = lab->boundary_factor;
            rp.N_e = (int)lab->N_e;
            rp.n_e0 = (int)(lab->N_e * lab->e_ratio_start[k]);                        // tl_trap_lab.py:38
            rp.n_h0 = n_h0;
            rp.protocol = legacy ? MCL_PROTO_TL_LEGACY : lab->protocol;
            rp.seg_begin = k; rp.seg_count = 1;
            rp.obs_begin = iso ? c * n_obs + lab->obs_begin[k] : 0;
            rp.obs_count = iso ? lab->obs_begin[k + 1] - lab->obs_begin[k] : 0;
        }
        if (iso) for (int o = 0; o < n_obs; o++) obs[(size_t)c * n_obs + o] = lab->obs_time[o];
    }

    // the per-replica outputs are carved from the cached scratch slab too (behind mcl_run's own workspace): no cudaMalloc
    // / cudaFree on the call path (a cudaFree synchronises the device and was seen to take hundreds of milliseconds)
    const size_t o_final = 0, o_status = align_up(o_final + sizeof(int32_t) * R, 256), o_esteps = align_up(o_status + sizeof(int32_t) * R, 256),
                 o_obs = align_up(o_esteps + sizeof(int64_t) * R, 256), out_bytes = align_up(o_obs + sizeof(int32_t) * obs.size(), 256);
    mcl_run_args a;
    memset(&a, 0, sizeof(a));
    a.replicas = reps.data(); a.n_replicas = (int32_t)R;
    a.segments = segs.data(); a.n_segments = n_rows;
    a.obs_time = iso ? obs.data() : nullptr; a.n_obs = (int32_t)obs.size();
    a.max_steps = lab->max_steps; a.mode = MCL_MODE_PHILOX;
    a.seed = seed; a.replica_id0 = candidate_id0 * (uint64_t)n_rows;
    a.stream = stream;
    lap(0);
    size_t need = 0;
    PlannedRun *plan = mcl_plan(&a, &need);
    lap(1);
    if (!plan) return MCL_ERR_ARG;
    need = align_up(need, 256);
    void *scratch = g_scratch.get(need + out_bytes);
    if (!scratch) { mcl_plan_discard(plan); set_error("mcl_objective: cudaMalloc(workspace %zu) failed", need + out_bytes); return MCL_ERR_ALLOC; }
    a.workspace = scratch; a.workspace_bytes = need;
    unsigned char *outs = (unsigned char *)scratch + need;
    int32_t *d_final = (int32_t *)(outs + o_final), *d_status = (int32_t *)(outs + o_status), *d_obs = (int32_t *)(outs + o_obs);
    int64_t *d_esteps = (int64_t *)(outs + o_esteps);
    a.final_n_e = d_final; a.esteps = d_esteps; a.status = d_status;
    a.obs_n_e = iso ? d_obs : nullptr;
    lap(2);
    // kernel time of this call (H2D of the tables excluded), for the benchmark's roofline line
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    const bool timed = cudaEventCreate(&ev0) == cudaSuccess && cudaEventCreate(&ev1) == cudaSuccess;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = mcl_run_planned(&a, plan, timed ? (void *)ev0 : nullptr, timed ? (void *)ev1 : nullptr);
    if (rc) { if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); return rc; }

    lap(3);
    std::vector<int32_t> final_n(R), status(R), obs_n(obs.size());
    std::vector<int64_t> est(R);
    cudaMemcpyAsync(final_n.data(), d_final, sizeof(int32_t) * R, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(status.data(), d_status, sizeof(int32_t) * R, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(est.data(), d_esteps, sizeof(int64_t) * R, cudaMemcpyDeviceToHost, st);
    if (iso) cudaMemcpyAsync(obs_n.data(), d_obs, sizeof(int32_t) * obs.size(), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (timed) {
        float ms = 0.f;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) g_last_kernel_ms = ms;
    }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (e != cudaSuccess) { set_error("mcl_objective: %s", cudaGetErrorString(e)); return MCL_ERR_CUDA; }

    lap(4);
    int64_t total = 0;
    std::vector<double> se((size_t)(iso ? n_obs : n_rows));
    for (int c = 0; c < S; c++) {
        bool failed = false;
        for (int k = 0; k < n_rows; k++) {
            size_t r = (size_t)c * n_rows + k;
            total += est[r];
            if (status[r] != MCL_OK) failed = true;
        }
        if (failed) { mse[c] = INFINITY; continue; }       // a candidate the reference would crash on
        if (!iso) {
            for (int k = 0; k < n_rows; k++) {
                double err = (double)final_n[(size_t)c * n_rows + k] / lab->N_e - lab->target[k];    // tl_trap_lab.py:111
                se[k] = err * err;
            }
        } else {
            for (int o = 0; o < n_obs; o++) {
                double err = (double)obs_n[(size_t)c * n_obs + o] / lab->N_e - lab->target[o];      // tl_trap_lab.py:165-167
                se[o] = err * err;
            }
        }
        mse[c] = numpy_sum(se.data(), se.size()) / (double)se.size();
    }
    if (esteps_total) *esteps_total = total;
    lap(5);
    if (timing) fprintf(stderr, "mcl_objective S=%d: tables+alloc %.2f ms, plan %.2f, scratch %.2f, upload+launch %.2f, kernel+download %.2f (kernel %.2f), mse %.2f\n",
                        S, t_ms[0], t_ms[1], t_ms[2], t_ms[3], t_ms[4], g_last_kernel_ms, t_ms[5]);
    return MCL_OK;
}

extern "C" void mcl_release_scratch(void) { g_scratch.release(); }
extern "C" float mcl_objective_last_kernel_ms(void) { return g_last_kernel_ms; }
