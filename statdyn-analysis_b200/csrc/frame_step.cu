// frame_step.cu -- one call per frame: the whole device-side sequence of the hot loop body
// (reference read/_read.py:134-153 for every active origin) enqueued from native code, so the host
// spends one FFI call per frame instead of one per kernel group plus tensor copies.
#include <cstdlib>
#include <mutex>
#include <vector>

#include "sdb_internal.cuh"

// ---- optional per-kernel-group timing (CUDA events on the launching stream) --------------------------
// Groups: 0 overlap prepare, 1 dyn_step + finalize, 2 struct, 3 overlap resolve.
namespace {
struct StepEvents {
    cudaEvent_t ev[5];
    int n_origins;
};
std::mutex g_prof_mutex;
bool g_prof_on = false;
std::vector<StepEvents> g_prof_steps;
std::vector<StepEvents> g_prof_free;

// ---- side stream (one per device): the overlap bookkeeping and the row sums are chains of small,
// latency-bound kernels that depend only on the descriptors / on dyn_step_kernel's partial rows; they run
// next to the increments pass and next to the struct kernel instead of in front of / behind them.
constexpr int SDB_MAX_GROUPS = 4;
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, prepared = nullptr, stepped = nullptr, joined = nullptr;
    // origin groups (see sdb_frame_step): dyn_step_kernel of group g >= 1 runs on group_stream[g - 1]
    cudaStream_t group_stream[SDB_MAX_GROUPS - 1] = {};
    cudaEvent_t group_ready = nullptr, group_done[SDB_MAX_GROUPS - 1] = {};
    int state = 0;  // 0 untried, 1 ready, -1 unavailable
};
SideStream g_side[64];
std::mutex g_side_mutex;

SideStream* side_stream()
{
    static const bool enabled = [] {
        const char* e = getenv("SDB_SIDE_STREAM");
        return !(e && e[0] == '0');
    }();
    if (!enabled) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_side_mutex);
    SideStream& ss = g_side[dev];
    if (ss.state == 0) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
        bool ok = cudaStreamCreateWithPriority(&ss.stream, cudaStreamNonBlocking, hi) == cudaSuccess;
        for (cudaEvent_t* e : {&ss.fork, &ss.prepared, &ss.stepped, &ss.joined, &ss.group_ready})
            ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
        for (int g = 0; g < SDB_MAX_GROUPS - 1; ++g) {
            ok = ok && cudaStreamCreateWithFlags(&ss.group_stream[g], cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&ss.group_done[g], cudaEventDisableTiming) == cudaSuccess;
        }
        ss.state = ok ? 1 : -1;
        if (!ok) cudaGetLastError();
    }
    return ss.state == 1 ? &ss : nullptr;
}

StepEvents* prof_begin(int n_origins)
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (!g_prof_on) return nullptr;
    StepEvents se;
    if (!g_prof_free.empty()) {
        se = g_prof_free.back();
        g_prof_free.pop_back();
    } else {
        for (auto& e : se.ev)
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    }
    se.n_origins = n_origins;
    g_prof_steps.push_back(se);
    return &g_prof_steps.back();
}
}  // namespace

// Start (on != 0) or stop collecting per-group kernel times of sdb_frame_step calls made by this process.
extern "C" void sdb_profile_enable(int32_t on)
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_on = on != 0;
}

// Sum of the collected group times in ms (h_ms[4]), number of steps and of origin-steps; synchronises
// on the recorded events and clears the collection.
extern "C" int sdb_profile_collect(double* h_ms, int64_t* h_steps, int64_t* h_origin_steps)
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (h_ms) h_ms[0] = h_ms[1] = h_ms[2] = h_ms[3] = 0.0;
    int64_t steps = 0, origin_steps = 0;
    for (auto& se : g_prof_steps) {
        if (sdb_check_cuda(cudaEventSynchronize(se.ev[4]), "sdb_profile_collect")) return SDB_ECUDA;
        for (int g = 0; g < 4; ++g) {
            float ms = 0.f;
            if (sdb_check_cuda(cudaEventElapsedTime(&ms, se.ev[g], se.ev[g + 1]), "sdb_profile_collect")) return SDB_ECUDA;
            if (h_ms) h_ms[g] += ms;
        }
        ++steps;
        origin_steps += se.n_origins;
        g_prof_free.push_back(se);
    }
    g_prof_steps.clear();
    if (h_steps) *h_steps = steps;
    if (h_origin_steps) *h_origin_steps = origin_steps;
    return SDB_OK;
}

int sdb_frame_step_impl(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, float* d_cur_compact,
                        const void* h_desc, void* d_desc, int32_t desc_bytes, int32_t segments_offset, int32_t n_origins,
                        int32_t n_segments, int32_t max_segment_count, double* d_partials, double* d_sums, double* h_sums,
                        void* d_ov_ws, int32_t ov_ws_origins, int32_t overlap_k, const float* h_sites_xy, int32_t n_sites,
                        double struct_dist, int32_t struct_mode, const void* const* h_inc_prev, const uint32_t* h_inc_flags,
                        int32_t inc_sets, void* d_inc_ws, int32_t step_flags, int32_t desc_on_sm, void* stream);

extern "C" int sdb_frame_step(const SdbDynParams* h_params, const float* d_pos, const float* d_quat,
                              float* d_cur_compact, const void* h_desc, void* d_desc, int32_t desc_bytes,
                              int32_t segments_offset, int32_t n_origins, int32_t n_segments,
                              int32_t max_segment_count, double* d_partials, double* d_sums, double* h_sums,
                              void* d_ov_ws, int32_t ov_ws_origins, int32_t overlap_k, const float* h_sites_xy,
                              int32_t n_sites, double struct_dist, int32_t struct_mode, const void* const* h_inc_prev,
                              const uint32_t* h_inc_flags, int32_t inc_sets, void* d_inc_ws, int32_t step_flags, void* stream)
{
    return sdb_frame_step_impl(h_params, d_pos, d_quat, d_cur_compact, h_desc, d_desc, desc_bytes, segments_offset, n_origins,
                               n_segments, max_segment_count, d_partials, d_sums, h_sums, d_ov_ws, ov_ws_origins, overlap_k,
                               h_sites_xy, n_sites, struct_dist, struct_mode, h_inc_prev, h_inc_flags, inc_sets, d_inc_ws,
                               step_flags, 0, stream);
}

// (the native engine of csrc/engine.cu calls this directly)
int sdb_frame_step_impl(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, float* d_cur_compact,
                        const void* h_desc, void* d_desc, int32_t desc_bytes, int32_t segments_offset, int32_t n_origins,
                        int32_t n_segments, int32_t max_segment_count, double* d_partials, double* d_sums, double* h_sums,
                        void* d_ov_ws, int32_t ov_ws_origins, int32_t overlap_k, const float* h_sites_xy, int32_t n_sites,
                        double struct_dist, int32_t struct_mode, const void* const* h_inc_prev, const uint32_t* h_inc_flags,
                        int32_t inc_sets, void* d_inc_ws, int32_t step_flags, int32_t desc_on_sm, void* stream)
{
    if (!h_params || !h_desc || !d_desc || desc_bytes <= 0 || segments_offset <= 0 || segments_offset >= desc_bytes) {
        sdb_set_error("sdb_frame_step: bad descriptor arguments");
        return SDB_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc;
    if (desc_on_sm && sdb_small_copies_on_sm()) {  // (h_desc is then mapped pinned memory with 16 bytes of slack: the engine's)
        rc = sdb_small_copy(d_desc, h_desc, (size_t)desc_bytes, stream);
    } else {
        rc = sdb_check_cuda(cudaMemcpyAsync(d_desc, h_desc, (size_t)desc_bytes, cudaMemcpyHostToDevice, s),
                            "sdb_frame_step: descriptor upload");
    }
    if (rc) return rc;
    const SdbOrigin* d_origins = static_cast<const SdbOrigin*>(d_desc);
    const SdbSegment* d_segments =
        reinterpret_cast<const SdbSegment*>(static_cast<const char*>(d_desc) + segments_offset);
    const bool sums = h_params->do_sums != 0;
    const bool use_ov = sums && d_ov_ws != nullptr && overlap_k > 0;
    StepEvents* prof = prof_begin(n_origins);
    StepEvents prof_copy;
    if (prof) {
        prof_copy = *prof;  // the vector may grow under another thread; the handles stay valid
        cudaEventRecord(prof_copy.ev[0], s);
    }
    // small systems: every kernel of the frame takes a few microseconds and the six event calls of the fork / join
    // cost the host more than the overlap gains (SDB_SIDE_MIN_N overrides the threshold)
    static const int side_min_n = [] { const char* e = getenv("SDB_SIDE_MIN_N"); return e ? atoi(e) : 0; }();
    SideStream* side = (sums && h_params->n >= side_min_n) ? side_stream() : nullptr;
    if (use_ov) {
        // overlap prepare needs only the descriptors (and, for sampled brackets, the state before the update)
        void* ps = stream;
        if (side) {
            cudaEventRecord(side->fork, s);
            cudaStreamWaitEvent(side->stream, side->fork, 0);
            ps = side->stream;
        }
        rc = sdb_overlap_prepare_ex(h_params, d_pos, d_quat, d_origins, n_origins, d_segments, n_segments,
                                    max_segment_count, overlap_k, d_ov_ws, ov_ws_origins,
                                    (step_flags & SDB_STEP_PREDICTED) != 0, ps);
        if (rc) return rc;
        if (side) cudaEventRecord(side->prepared, side->stream);
    }
    if (prof) cudaEventRecord(prof_copy.ev[1], s);
    const bool use_inc = inc_sets > 0 && d_inc_ws != nullptr && h_inc_prev != nullptr && h_inc_flags != nullptr && d_pos;
    if (use_inc) {  // phase 1 once per frame and previous sample (it also writes the new previous planes)
        rc = sdb_dyn_increments(h_params, d_pos, d_quat, d_cur_compact, h_inc_prev, h_inc_flags, inc_sets, d_inc_ws, stream);
        if (rc) return rc;
    }
    if (use_ov && side) cudaStreamWaitEvent(s, side->prepared, 0);

    // ---- origin groups (experiment, off by default: SDB_STEP_GROUPS=1) -----------------------------------
    // dyn_step_kernel leaves DRAM bandwidth unused (it is issue bound) and the struct kernel, which needs
    // the finished planes of an origin, comes after it.  With SDB_STEP_GROUPS=G > 1 the origins are split
    // into up to G groups of whole segments: every group's dyn_step launch goes to its own stream and
    // struct of group g starts as soon as dyn_step of group g is done, next to dyn_step of the later groups.
    // Measured on B200 (1 M molecules): 200 origins 1.655 ms (G=1) / 1.69 (G=2) / 1.68 (G=3) per step,
    // 25 origins 0.283 / 0.284 / 0.298 -- the mix of the two kernels is no faster than running them in turn.
    static const int want_groups = [] {
        const char* e = getenv("SDB_STEP_GROUPS");
        const int g = e ? atoi(e) : 1;
        return g < 1 ? 1 : (g > SDB_MAX_GROUPS ? SDB_MAX_GROUPS : g);
    }();
    const bool do_struct = sums && h_sites_xy && n_sites > 0;
    int n_groups = 1;
    int seg_lo[SDB_MAX_GROUPS + 1] = {0, n_segments}, org_lo[SDB_MAX_GROUPS + 1] = {0, n_origins};
    if (side && do_struct && !prof && want_groups > 1 && n_segments > 1) {
        const SdbSegment* h_segments = reinterpret_cast<const SdbSegment*>(static_cast<const char*>(h_desc) + segments_offset);
        bool contiguous = true;  // segments must tile [0, n_origins) in order
        int next = 0;
        for (int i = 0; i < n_segments && contiguous; ++i) {
            contiguous = h_segments[i].first == next && h_segments[i].count > 0;
            next += h_segments[i].count;
        }
        if (contiguous && next == n_origins) {
            const int target = want_groups < n_segments ? want_groups : n_segments;
            int g = 0, i = 0;
            while (g < target) {
                seg_lo[g] = i;
                org_lo[g] = h_segments[i].first;
                const int until = (int)((long)n_origins * (g + 1) / target);  // origins covered after this group
                do ++i;
                while (i < n_segments - (target - 1 - g) && h_segments[i - 1].first + h_segments[i - 1].count < until);
                ++g;
            }
            seg_lo[g] = n_segments;
            org_lo[g] = n_origins;
            n_groups = g;
            if (seg_lo[g - 1] >= n_segments) n_groups = 1;  // (cannot happen: every group takes >= 1 segment)
        }
        if (n_groups == 1) {
            seg_lo[0] = 0; seg_lo[1] = n_segments;
            org_lo[0] = 0; org_lo[1] = n_origins;
        }
    }
    if (n_groups > 1) {
        cudaEventRecord(side->group_ready, s);  // descriptors uploaded, increments computed, brackets ready
        for (int g = 1; g < n_groups; ++g) cudaStreamWaitEvent(side->group_stream[g - 1], side->group_ready, 0);
    }
    for (int g = 0; g < n_groups; ++g) {
        void* gs = g == 0 ? stream : static_cast<void*>(side->group_stream[g - 1]);
        // (without the increments pass the first segment of the frame also writes the new previous-sample planes)
        rc = sdb_dyn_step_launch(h_params, d_pos, d_quat, (use_inc || g > 0) ? nullptr : d_cur_compact, d_origins, n_origins,
                                 d_segments + seg_lo[g], seg_lo[g + 1] - seg_lo[g], d_partials, d_sums,
                                 use_ov ? d_ov_ws : nullptr, ov_ws_origins, use_inc ? d_inc_ws : nullptr,
                                 use_inc ? inc_sets : 0, (side || use_ov) ? 0 : 1, gs);
        if (rc) return rc;
        if (g > 0) cudaEventRecord(side->group_done[g - 1], side->group_stream[g - 1]);
    }
    if (prof) cudaEventRecord(prof_copy.ev[2], s);
    // row sums + overlap resolve (side stream)  ||  struct (this stream)
    void* rs = stream;
    if (side) {
        cudaEventRecord(side->stepped, s);
        cudaStreamWaitEvent(side->stream, side->stepped, 0);
        for (int g = 1; g < n_groups; ++g) cudaStreamWaitEvent(side->stream, side->group_done[g - 1], 0);
        rs = side->stream;
        if (!use_ov) {
            rc = sdb_dyn_finalize(d_partials, h_params->n, d_sums, n_origins, 0, rs);
            if (rc) return rc;
        }
    }
    auto resolve = [&]() -> int {  // row sums + selections + intersection in one launch
        if (!use_ov) return SDB_OK;
        return sdb_overlap_resolve_fused(d_origins, n_origins, h_params->n, h_params->stride, overlap_k, d_ov_ws,
                                         ov_ws_origins, d_partials, d_sums, rs);
    };
    if (side) {
        rc = resolve();
        if (rc) return rc;
        cudaEventRecord(side->joined, side->stream);
    }
    if (do_struct) {
        for (int g = 0; g < n_groups; ++g) {
            if (g > 0) cudaStreamWaitEvent(s, side->group_done[g - 1], 0);
            // (dyn_step_kernel has cleared the struct slots of every origin of the frame)
            rc = sdb_struct_launch(d_origins + org_lo[g], org_lo[g + 1] - org_lo[g], h_params->n, h_params->stride, h_sites_xy,
                                   n_sites, struct_dist, struct_mode, d_sums + (size_t)org_lo[g] * SDB_NSUM, 0, stream);
            if (rc) return rc;
        }
    }
    if (prof) cudaEventRecord(prof_copy.ev[3], s);
    if (side) {
        cudaStreamWaitEvent(s, side->joined, 0);
    } else {
        rc = resolve();
        if (rc) return rc;
    }
    if (prof) cudaEventRecord(prof_copy.ev[4], s);
    if (sums && h_sums) {
        rc = sdb_check_cuda(cudaMemcpyAsync(h_sums, d_sums, sizeof(double) * SDB_NSUM * (size_t)n_origins,
                                            cudaMemcpyDeviceToHost, s),
                            "sdb_frame_step: sums download");
        if (rc) return rc;
    }
    return SDB_OK;
}
