// mcl_abi.cu -- extern "C" entry points of libmcl_b200.so (see include/mcl_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <atomic>
#include <vector>
#include "mcl_common.cuh"

namespace mcl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<int64_t> g_launches{0};      // kernels launched by this process (mcl_launch_count)

struct Layout {
    size_t off_rep, off_seg, off_obs, off_grp, off_order, off_slabs, total;
    size_t stride;
    int cap_e, cap_h;
    bool any_dose;
    // Which kernel runs which replica, decided per replica from its own fields: `small` -> one warp per replica
    // (mcl_smallbox.cu, the Optimizer path), `big` -> one CTA per replica (mcl_philox.cu / mcl_replay.cu).  Both lists are
    // ordered longest-expected-work first: blocks are dispatched in index order, so the hardware scheduler is the work
    // queue and the long replicas do not end up in the tail of the launch.
    std::vector<int32_t> small, big;
    int small_hcap;
    // the small list is launched in hole-capacity classes (shared memory per warp = 12 bytes x class capacity, so the
    // small boxes of a population are not held to the occupancy of its largest one): [begin, end) into `small` + capacity
    struct SmallClass { int begin, end, hcap; };
    std::vector<SmallClass> small_classes;
};
static const int kSmallClassCaps[] = {384, 640, 896, 1152, 1536, 2048, 3072, kSmallboxMaxHoles};

// expected work of a replica, up to a constant: what the launch order sorts by
static double work_estimate(const mcl_run_args *a, const mcl_replica &rp)
{
    const double n = (double)std::max(rp.N_e, rp.n_e0);
    if (rp.protocol == MCL_PROTO_ISO_LAB)
        return n * (rp.obs_count > 0 && a->obs_time ? a->obs_time[rp.obs_begin + rp.obs_count - 1] : 0.0);
    double w = 0.0;
    for (int s = 0; s < rp.seg_count; s++) {
        const mcl_segment &sg = a->segments[rp.seg_begin + s];
        w += std::min(sg.duration, 1e9) * (sg.dose_rate != 0.0 ? 2.0 : 1.0);
    }
    return n * w;
}

static int plan_layout(const mcl_run_args *a, Layout *L)
{
    if (!a || !a->replicas || a->n_replicas <= 0 || !a->segments || a->n_segments <= 0) {
        set_error("mcl_run: replicas/segments missing"); return MCL_ERR_ARG;
    }
    if (a->max_steps <= 0) { set_error("mcl_run: max_steps must be positive"); return MCL_ERR_ARG; }
    if (a->mode != MCL_MODE_PHILOX && a->mode != MCL_MODE_REPLAY) { set_error("mcl_run: unknown mode %d", a->mode); return MCL_ERR_ARG; }
    if (a->hist && (a->hist->n_bins <= 0 || a->hist->n_groups <= 0 || !(a->hist->hi > a->hist->lo))) { set_error("mcl_run: bad histogram spec"); return MCL_ERR_ARG; }
    int ne_max = 0, nh_max = 0, seg_max = 1;
    bool any_dose = false;
    L->small.clear(); L->big.clear(); L->small_hcap = 0;
    L->small.reserve((size_t)a->n_replicas);
    bool small_ok = a->mode == MCL_MODE_PHILOX && !a->hist && !a->kind && !a->e_idx && !a->h_idx;
    if (const char *env = getenv("MCL_SMALLBOX")) small_ok = small_ok && atoi(env) != 0;       // test knob: 0 = block kernel for every replica
    for (int r = 0; r < a->n_replicas; r++) {
        const mcl_replica &rp = a->replicas[r];
        if (rp.N_e < 0 || rp.n_e0 < 0 || rp.n_h0 < 0) { set_error("replica %d: negative sizes", r); return MCL_ERR_ARG; }
        if (rp.protocol < MCL_PROTO_SIMULATE || rp.protocol > MCL_PROTO_TL_LEGACY) { set_error("replica %d: bad protocol %d", r, rp.protocol); return MCL_ERR_ARG; }
        if (rp.protocol == MCL_PROTO_TL_LEGACY && !(small_ok && smallbox_eligible(rp))) {
            set_error("replica %d: the legacy TL semantics exist for native-mode boxes of at most 124 traps without histograms", r); return MCL_ERR_ARG;
        }
        if (rp.seg_count < 1 || rp.seg_begin < 0 || rp.seg_begin + rp.seg_count > a->n_segments) {
            set_error("replica %d: segment range [%d,%d) outside table of %d", r, rp.seg_begin, rp.seg_begin + rp.seg_count, a->n_segments);
            return MCL_ERR_ARG;
        }
        if (rp.protocol != MCL_PROTO_SIMULATE && rp.seg_count != 1) { set_error("replica %d: lab protocols take exactly one segment", r); return MCL_ERR_ARG; }
        if (rp.obs_count < 0 || rp.obs_begin < 0 || rp.obs_begin + rp.obs_count > a->n_obs) { set_error("replica %d: observation range outside table", r); return MCL_ERR_ARG; }
        if (rp.obs_count > 0 && !a->obs_time) { set_error("replica %d: obs_time missing", r); return MCL_ERR_ARG; }
        if (small_ok && smallbox_eligible(rp)) {          // sized separately: these replicas never see a slab
            L->small.push_back(r);
            L->small_hcap = std::max(L->small_hcap, smallbox_hole_capacity(rp));
            continue;
        }
        L->big.push_back(r);
        ne_max = std::max(ne_max, std::max(rp.N_e, rp.n_e0));
        nh_max = std::max(nh_max, rp.n_h0);
        seg_max = std::max(seg_max, rp.seg_count);
        for (int s = 0; s < rp.seg_count; s++) if (a->segments[rp.seg_begin + s].dose_rate != 0.0) any_dose = true;
        if (a->hist) {
            // leg sg of replica r writes histogram row hist_group[r] + sg: every such row must exist
            const int row0 = a->hist_group ? a->hist_group[r] : 0;
            if (row0 < 0 || row0 + rp.seg_count > a->hist->n_groups) {
                set_error("replica %d: histogram rows [%d,%d) outside the %d rows of the buffers", r, row0, row0 + rp.seg_count, a->hist->n_groups);
                return MCL_ERR_ARG;
            }
        }
    }
    L->cap_e = ne_max + 8 + seg_max;
    if (a->mode == MCL_MODE_REPLAY) L->cap_h = nh_max + L->cap_e + 8;
    else L->cap_h = nh_max + (any_dose ? 2 * ne_max + kFillExtra : 0) + 64 + seg_max;     // see kFillExtra (regrid)
    L->any_dose = any_dose;
    L->stride = a->mode == MCL_MODE_REPLAY ? replay_ws_stride(L->cap_e, L->cap_h) : philox_ws_stride(L->cap_e, L->cap_h, any_dose);
    size_t o = 0;
    L->off_rep = o; o = align_up(o + sizeof(mcl_replica) * (size_t)a->n_replicas, 256);
    L->off_seg = o; o = align_up(o + sizeof(mcl_segment) * (size_t)a->n_segments, 256);
    L->off_obs = o; o = align_up(o + sizeof(double) * (size_t)std::max(a->n_obs, 1), 256);
    L->off_grp = o; o = align_up(o + sizeof(int32_t) * (size_t)a->n_replicas, 256);
    L->off_order = o; o = align_up(o + sizeof(int32_t) * (size_t)a->n_replicas, 256);
    L->off_slabs = o; o += L->stride * L->big.size();
    if (a->mode == MCL_MODE_PHILOX) {
        // longest expected work first.  Populations have a handful of distinct weights (one per lab row), ensembles one:
        // a counting sort over the distinct values; a comparison sort only when there are many.
        std::vector<double> w((size_t)a->n_replicas);
        std::vector<double> distinct;
        for (int r = 0; r < a->n_replicas; r++) {
            w[r] = work_estimate(a, a->replicas[r]);
            if (distinct.size() <= 64 && std::find(distinct.begin(), distinct.end(), w[r]) == distinct.end()) distinct.push_back(w[r]);
        }
        auto order_list = [&](std::vector<int32_t> &list) {
            if (distinct.size() <= 1 || list.size() < 2) return;
            if (distinct.size() > 64) {
                std::stable_sort(list.begin(), list.end(), [&](int32_t x, int32_t y) { return w[x] > w[y]; });
                return;
            }
            std::vector<double> d(distinct);
            std::sort(d.begin(), d.end(), std::greater<double>());
            std::vector<std::vector<int32_t>> bucket(d.size());
            for (int32_t r : list) bucket[std::find(d.begin(), d.end(), w[r]) - d.begin()].push_back(r);
            size_t o2 = 0;
            for (auto &b : bucket) for (int32_t r : b) list[o2++] = r;
        };
        order_list(L->big);
        // small boxes: by capacity class first, longest work first inside a class -- one pass into (class, weight) buckets
        L->small_classes.clear();
        if (!L->small.empty()) {
            constexpr int NCLS = (int)(sizeof(kSmallClassCaps) / sizeof(kSmallClassCaps[0]));
            const bool few = distinct.size() <= 64;
            std::vector<double> d(distinct);
            std::sort(d.begin(), d.end(), std::greater<double>());
            const size_t nw = few ? d.size() : 1;
            std::vector<std::vector<int32_t>> bucket((size_t)NCLS * nw);
            int hmax[NCLS] = {0};
            for (int32_t r : L->small) {
                const int hc = smallbox_hole_capacity(a->replicas[r]);
                int cls = 0;
                while (cls < NCLS - 1 && hc > kSmallClassCaps[cls]) cls++;
                const size_t wb = few ? (size_t)(std::find(d.begin(), d.end(), w[r]) - d.begin()) : 0;
                bucket[(size_t)cls * nw + wb].push_back(r);
                hmax[cls] = std::max(hmax[cls], hc);
            }
            L->small.clear();
            for (int cls = 0; cls < NCLS; cls++) {
                const int b = (int)L->small.size();
                for (size_t wb = 0; wb < nw; wb++) {
                    auto &v = bucket[(size_t)cls * nw + wb];
                    L->small.insert(L->small.end(), v.begin(), v.end());
                }
                const int e = (int)L->small.size();
                if (e == b) continue;
                if (!few) std::stable_sort(L->small.begin() + b, L->small.begin() + e, [&](int32_t x, int32_t y) { return w[x] > w[y]; });
                L->small_classes.push_back(Layout::SmallClass{b, e, hmax[cls]});
            }
        }
    }
    L->total = o;
    return MCL_OK;
}

#define CUDA_TRY(x)                                                                         \
    do {                                                                                    \
        cudaError_t _e = (x);                                                               \
        if (_e != cudaSuccess) { set_error("%s: %s", #x, cudaGetErrorString(_e)); return MCL_ERR_CUDA; } \
    } while (0)

static int run_device(const mcl_run_args *a, cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr, const Layout *planned = nullptr)
{
    Layout L_own;
    if (!planned) {
        int rc = plan_layout(a, &L_own);
        if (rc) return rc;
        planned = &L_own;
    }
    const Layout &L = *planned;
    if (!a->workspace || a->workspace_bytes < L.total) {
        set_error("mcl_run: workspace too small (%zu < %zu)", a->workspace_bytes, L.total); return MCL_ERR_ARG;
    }
    if (a->mode == MCL_MODE_REPLAY && (!a->replay_u || !a->replay_off)) { set_error("mcl_run: replay mode needs replay_u / replay_off"); return MCL_ERR_ARG; }
    if (a->mode == MCL_MODE_PHILOX && L.cap_e + 2 > philox_max_slots()) {
        set_error("mcl_run: %d electrons per replica exceed the kernel capacity %d", L.cap_e, philox_max_slots()); return MCL_ERR_CAPACITY;
    }
    cudaStream_t st = (cudaStream_t)a->stream;
    unsigned char *ws = (unsigned char *)a->workspace;
    CUDA_TRY(cudaMemcpyAsync(ws + L.off_rep, a->replicas, sizeof(mcl_replica) * (size_t)a->n_replicas, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ws + L.off_seg, a->segments, sizeof(mcl_segment) * (size_t)a->n_segments, cudaMemcpyHostToDevice, st));
    if (a->n_obs > 0) CUDA_TRY(cudaMemcpyAsync(ws + L.off_obs, a->obs_time, sizeof(double) * (size_t)a->n_obs, cudaMemcpyHostToDevice, st));
    if (a->hist && a->hist_group) CUDA_TRY(cudaMemcpyAsync(ws + L.off_grp, a->hist_group, sizeof(int32_t) * (size_t)a->n_replicas, cudaMemcpyHostToDevice, st));

    LaunchParams p;
    memset(&p, 0, sizeof(p));
    p.replicas = (const mcl_replica *)(ws + L.off_rep);
    p.segments = (const mcl_segment *)(ws + L.off_seg);
    p.obs_time = (const double *)(ws + L.off_obs);
    p.n_replicas = a->n_replicas; p.max_steps = a->max_steps;
    p.seed = a->seed; p.replica_id0 = a->replica_id0;
    p.replay_u = a->replay_u; p.replay_off = a->replay_off;
    p.event = a->event; p.n_e = a->n_e; p.t = a->t;
    p.kind = a->kind; p.e_idx = a->e_idx; p.h_idx = a->h_idx;
    p.steps_used = a->steps_used; p.final_n_e = a->final_n_e;
    p.esteps = a->esteps; p.consumed = a->consumed; p.status = a->status; p.obs_n_e = a->obs_n_e;
    if (a->hist) {
        p.hist = *a->hist;
        p.hist_group = a->hist_group ? (const int32_t *)(ws + L.off_grp) : nullptr;
        p.hist_events = (unsigned long long *)a->hist_events;
        p.hist_occ = (unsigned long long *)a->hist_occ;
        p.hist_occ_sq = (unsigned long long *)a->hist_occ_sq;
    }
    p.ws = ws + L.off_slabs; p.ws_stride = L.stride; p.cap_e = L.cap_e; p.cap_h = L.cap_h; p.with_regrid = L.any_dose ? 1 : 0;
    p.order = nullptr; p.n_launch = a->n_replicas;
    int32_t *order_dev = (int32_t *)(ws + L.off_order);
    if (a->mode == MCL_MODE_PHILOX) {
        // [small list | big list] in one upload (pageable host memory: the copy is staged before the call returns)
        std::vector<int32_t> both(L.small);
        both.insert(both.end(), L.big.begin(), L.big.end());
        CUDA_TRY(cudaMemcpyAsync(order_dev, both.data(), sizeof(int32_t) * both.size(), cudaMemcpyHostToDevice, st));
        p.order = order_dev + L.small.size(); p.n_launch = (int32_t)L.big.size();
    }
    if (ev0) cudaEventRecord(ev0, st);
    cudaError_t e = cudaSuccess;
    if (a->mode == MCL_MODE_REPLAY) { e = launch_replay(p, st); g_launches++; }
    else {
        for (const auto &c : L.small_classes)
            if (e == cudaSuccess) { e = launch_smallbox(p, order_dev + c.begin, c.end - c.begin, c.hcap, st); g_launches++; }
        if (e == cudaSuccess && !L.big.empty()) { e = launch_philox(p, st, 0); g_launches++; }
    }
    if (ev1) cudaEventRecord(ev1, st);
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return MCL_ERR_CUDA; }
    return MCL_OK;
}

int mcl_run_timed(const mcl_run_args *a, void *ev0, void *ev1) { return run_device(a, (cudaEvent_t)ev0, (cudaEvent_t)ev1); }

// plan once, then allocate and run: what mcl_objective does (planning a population of 45 000 replicas takes milliseconds)
struct PlannedRun { Layout L; };
PlannedRun *mcl_plan(const mcl_run_args *a, size_t *bytes)
{
    PlannedRun *pr = new PlannedRun;
    if (plan_layout(a, &pr->L)) { delete pr; return nullptr; }
    *bytes = pr->L.total;
    return pr;
}
int mcl_run_planned(const mcl_run_args *a, PlannedRun *pr, void *ev0, void *ev1)
{
    const int rc = run_device(a, (cudaEvent_t)ev0, (cudaEvent_t)ev1, &pr->L);
    delete pr;
    return rc;
}
void mcl_plan_discard(PlannedRun *pr) { delete pr; }

}  // namespace mcl

using namespace mcl;

namespace {
struct HostMirror {
    struct Buf { void *dev; void *host; size_t bytes; bool in; };
    std::vector<Buf> bufs;
    void *ws = nullptr;
    ~HostMirror() { for (auto &b : bufs) cudaFree(b.dev); if (ws) cudaFree(ws); }
    // replaces *field (a HOST pointer) by a fresh device buffer of `bytes`
    template <typename T> int add(T **field, size_t bytes, bool in)
    {
        if (!*field) return MCL_OK;
        void *dev = nullptr;
        if (cudaMalloc(&dev, bytes ? bytes : 1) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", bytes); return MCL_ERR_ALLOC; }
        bufs.push_back(Buf{dev, (void *)*field, bytes, in});
        *field = (T *)dev;
        return MCL_OK;
    }
};
}  // namespace

extern "C" {

int mcl_abi_version(void) { return MCL_ABI_VERSION; }
int64_t mcl_launch_count(void) { return g_launches.load(); }
const char *mcl_last_error(void) { return g_err; }

size_t mcl_workspace_bytes(const mcl_run_args *args)
{
    Layout L;
    if (plan_layout(args, &L)) return 0;
    return L.total;
}

int mcl_run(const mcl_run_args *args) { return run_device(args); }

int mcl_run_host(const mcl_run_args *args)
{
    Layout L;
    int rc = plan_layout(args, &L);
    if (rc) return rc;
    const size_t R = (size_t)args->n_replicas, MS = (size_t)args->max_steps;
    mcl_run_args d = *args;
    HostMirror M;
    cudaStream_t st = (cudaStream_t)args->stream;
    int64_t n_u = 0;
    if (args->mode == MCL_MODE_REPLAY && args->replay_off) n_u = args->replay_off[args->n_replicas];
#define MIRROR(f, bytes, in) if ((rc = M.add(&d.f, (bytes), (in))) != MCL_OK) return rc
    MIRROR(replay_u, sizeof(double) * (size_t)n_u, true);
    MIRROR(replay_off, sizeof(int64_t) * (R + 1), true);
    const size_t n_inputs = M.bufs.size();
    MIRROR(event, sizeof(int32_t) * R * MS, false);
    MIRROR(n_e, sizeof(int32_t) * R * MS, false);
    MIRROR(t, sizeof(double) * R * MS, false);
    MIRROR(kind, sizeof(int32_t) * R * MS, false);
    MIRROR(e_idx, sizeof(int32_t) * R * MS, false);
    MIRROR(h_idx, sizeof(int32_t) * R * MS, false);
    MIRROR(steps_used, sizeof(int32_t) * R, false);
    MIRROR(final_n_e, sizeof(int32_t) * R, false);
    MIRROR(esteps, sizeof(int64_t) * R, false);
    MIRROR(consumed, sizeof(int64_t) * R, false);
    MIRROR(status, sizeof(int32_t) * R, false);
    MIRROR(obs_n_e, sizeof(int32_t) * (size_t)std::max(args->n_obs, 1), false);
    if (args->hist) {
        size_t hb = sizeof(int64_t) * (size_t)args->hist->n_groups * (size_t)args->hist->n_bins;
        MIRROR(hist_events, hb, true);      // histograms are ADDED to: upload the caller's values
        MIRROR(hist_occ, hb, true);
        MIRROR(hist_occ_sq, hb, true);
    }
#undef MIRROR
    if (cudaMalloc(&M.ws, L.total) != cudaSuccess) { set_error("cudaMalloc(workspace %zu) failed", L.total); M.ws = nullptr; return MCL_ERR_ALLOC; }
    d.workspace = M.ws; d.workspace_bytes = L.total;
    for (auto &b : M.bufs) {
        cudaError_t e = b.in ? cudaMemcpyAsync(b.dev, b.host, b.bytes, cudaMemcpyHostToDevice, st)
                             : cudaMemsetAsync(b.dev, 0, b.bytes, st);
        if (e != cudaSuccess) { set_error("upload: %s", cudaGetErrorString(e)); return MCL_ERR_CUDA; }
    }
    rc = run_device(&d);
    if (rc) return rc;
    for (size_t i = n_inputs; i < M.bufs.size(); i++) {
        auto &b = M.bufs[i];
        cudaError_t e = cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) { set_error("download: %s", cudaGetErrorString(e)); return MCL_ERR_CUDA; }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("kernel: %s", cudaGetErrorString(e)); return MCL_ERR_CUDA; }
    return MCL_OK;
}

}  // extern "C"
