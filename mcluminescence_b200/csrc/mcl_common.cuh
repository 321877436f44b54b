// mcl_common.cuh -- shared device/host declarations for libmcl_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mcl_b200.h"

namespace mcl {

// Device-side view of one launch: every pointer is device memory.
struct LaunchParams {
    const mcl_replica *replicas;
    const mcl_segment *segments;
    const double *obs_time;
    int32_t n_replicas, max_steps;
    uint64_t seed, replica_id0;
    // replay
    const double *replay_u;
    const int64_t *replay_off;
    // outputs
    int32_t *event, *n_e;
    double *t;
    int32_t *kind, *e_idx, *h_idx;
    int32_t *steps_used, *final_n_e;
    int64_t *esteps, *consumed;
    int32_t *status;
    int32_t *obs_n_e;
    // histogram
    mcl_hist_spec hist;          // n_bins == 0 => disabled
    const int32_t *hist_group;   // device [R] or nullptr
    unsigned long long *hist_events, *hist_occ, *hist_occ_sq;
    // per-replica scratch: replica r owns [ws + r*ws_stride, +ws_stride)
    unsigned char *ws;
    size_t ws_stride;
    int32_t cap_e, cap_h;        // per-replica element capacities inside the scratch
    int32_t with_regrid;         // the slabs carry the regrid scratch (some replica of the launch has a dosed leg)
    // which replicas this launch runs: block b runs replica order[b] (nullptr: b) and owns slab b; n_launch blocks
    const int32_t *order;
    int32_t n_launch;
};

// Fill-region slots beyond N_e: a fill that finds N_e + kFillExtra of them in use folds them into the cell grid first, so a
// replica never needs more than n_h0 + 2 N_e + kFillExtra hole slots (alive holes <= n_h0 + N_e at any time).
constexpr int kFillExtra = 64;

void set_error(const char *fmt, ...);
int mcl_run_timed(const mcl_run_args *a, void *ev0, void *ev1);
struct PlannedRun;
PlannedRun *mcl_plan(const mcl_run_args *a, size_t *bytes);                      // nullptr on error (mcl_last_error)
int mcl_run_planned(const mcl_run_args *a, PlannedRun *pr, void *ev0, void *ev1);   // consumes pr
void mcl_plan_discard(PlannedRun *pr);   // mcl_run with CUDA events recorded around the kernel launch

// kernels' host launchers (defined in the .cu files)
cudaError_t launch_replay(const LaunchParams &p, cudaStream_t stream);
cudaError_t launch_philox(const LaunchParams &p, cudaStream_t stream, int max_slots);
size_t replay_ws_stride(int cap_e, int cap_h);
size_t philox_ws_stride(int cap_e, int cap_h, bool with_regrid);
int philox_max_slots();       // largest per-replica electron capacity the block kernel supports
// one-warp-per-replica kernel of the Optimizer path (mcl_smallbox.cu)
constexpr int kSmallboxMaxHoles = 4096;
static inline int smallbox_hole_capacity(const mcl_replica &rp) { return rp.n_h0 + rp.N_e + 8; }   // slots are reused lowest-first
bool smallbox_eligible(const mcl_replica &rp);
cudaError_t launch_smallbox(const LaunchParams &p, const int *order_dev, int count, int hcap, cudaStream_t stream);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mcl
