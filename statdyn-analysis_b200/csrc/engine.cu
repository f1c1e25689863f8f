// engine.cu -- the native engine: every piece of host work of the hot loop behind one call per frame.
//
// Reference: read/_read.py:128-168 keeps a dict of (Dynamics, Relaxations) objects per keyframe and calls
// compute_all + Relaxations.add on each active one for every frame.  Round 1 did that bookkeeping in Python
// around sdb_frame_step (~0.25 ms per frame, as long as the kernels of a 25-origin rank and 10x the kernels
// of a 32 k-molecule frame).  Here it is plain C++: per-origin records, the reference-counted pool of
// previous samples (TrackedMotion.previous_*, _util.py:155-156), grouping of the active origins into
// segments, descriptor rings in pinned memory, staging of host frames (pageable -> pinned by a few host
// threads -> device on a copy stream), the result-row ring.  No torch, no Python between the kernels.
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "sdb_internal.cuh"

int sdb_frame_step_impl(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, float* d_cur_compact,
                        const void* h_desc, void* d_desc, int32_t desc_bytes, int32_t segments_offset, int32_t n_origins,
                        int32_t n_segments, int32_t max_segment_count, double* d_partials, double* d_sums, double* h_sums,
                        void* d_ov_ws, int32_t ov_ws_origins, int32_t overlap_k, const float* h_sites_xy, int32_t n_sites,
                        double struct_dist, int32_t struct_mode, const void* const* h_inc_prev, const uint32_t* h_inc_flags,
                        int32_t inc_sets, void* d_inc_ws, int32_t step_flags, int32_t desc_on_sm, void* stream);

namespace {

// ---- a few persistent host threads for the pageable -> pinned copy of a frame -------------------------
// A 1 M-molecule frame is 28 MB: one core copies it in ~3 ms, twice the time of the kernels of that frame.
class CopyPool {
public:
    static CopyPool& instance()
    {
        static CopyPool pool;
        return pool;
    }
    // memcpy(dst, src, bytes) split over the workers and the calling thread
    void copy(void* dst, const void* src, size_t bytes)
    {
        const size_t min_chunk = size_t(1) << 20;
        if (bytes < 4 * min_chunk || workers_.empty()) {
            memcpy(dst, src, bytes);
            return;
        }
        std::lock_guard<std::mutex> one_caller(callers_);  // (engines of different host threads share the pool)
        std::unique_lock<std::mutex> lock(mutex_);
        const int parts = (int)std::min<size_t>(workers_.size() + 1, bytes / min_chunk);
        const size_t chunk = ((bytes + parts - 1) / parts + 63) & ~size_t(63);
        dst_ = static_cast<char*>(dst);
        src_ = static_cast<const char*>(src);
        bytes_ = bytes;
        chunk_ = chunk;
        next_ = 1;  // part 0 is the caller's
        parts_ = parts;
        done_ = 0;
        ++generation_;
        lock.unlock();
        wake_.notify_all();
        memcpy(dst_, src_, std::min(chunk, bytes));
        lock.lock();
        ++done_;
        finished_.wait(lock, [&] { return done_ == parts_; });
    }

private:
    CopyPool()
    {
        const int hw = (int)std::thread::hardware_concurrency();
        int want = hw >= 16 ? 7 : 3;
        if (const char* e = getenv("SDB_COPY_THREADS")) want = atoi(e);
        if (hw > 0) want = std::min(want, std::max(hw - 1, 0));
        for (int i = 0; i < want; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            stop_ = true;
        }
        wake_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run()
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lock(mutex_);
        for (;;) {
            wake_.wait(lock, [&] { return stop_ || (generation_ != seen && next_ < parts_); });
            if (stop_) return;
            while (next_ < parts_) {
                const int part = next_++;
                const size_t off = (size_t)part * chunk_;
                const size_t len = off < bytes_ ? std::min(chunk_, bytes_ - off) : 0;
                char* d = dst_ + off;
                const char* s = src_ + off;
                lock.unlock();
                if (len) memcpy(d, s, len);
                lock.lock();
                if (++done_ == parts_) finished_.notify_all();
            }
            seen = generation_;
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mutex_, callers_;
    std::condition_variable wake_, finished_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, chunk_ = 0;
    int next_ = 0, parts_ = 0, done_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

struct PoolFrame {
    float* planes = nullptr;  // float[4][stride]: x, y, qw, qz
    int refs = 0;
    bool has_orientation = true;
    int64_t stamp = -1;  // step in which `group` was assigned
    int group = 0;
};

struct Origin {
    int64_t key = 0, t0 = 0;
    char* block = nullptr;  // one allocation: delta planes + history | status | flags
    float* delta = nullptr;
    uint8_t* flags = nullptr;
    uint32_t* status = nullptr;
    PoolFrame* prev = nullptr;
    int64_t max_td = -1, updates = 0, last_ts = 0;
    int ov_run = 0;  // consecutive earlier steps with `overlap` at increasing times
    int64_t order = 0;  // creation order
    int position = 0;   // index in the origin order of the step being built
};

struct ResultSlot {
    char* h_desc = nullptr;  // pinned
    char* d_desc = nullptr;
    double* d_sums = nullptr;
    double* h_sums = nullptr;  // pinned
    double* d_isf = nullptr;
    double* h_isf = nullptr;  // pinned
    cudaEvent_t done = nullptr;
    bool pending = false;  // `done` has been recorded and not yet waited for
    int64_t ticket = -1;
    int n_rows = 0;
    bool host_rows = false, has_isf = false, has_sums = false;
    std::vector<int64_t> timediffs;  // caller order
    std::vector<int32_t> perm;       // position of keys[i] in the engine's origin order
    bool identity = true;
    std::vector<double> stash;  // rows saved across a capacity change
};

struct FrameSlot {
    float* h = nullptr;  // pinned staging copy (SDB_MEM_HOST)
    float* d = nullptr;  // device frame: pos at 0, quat at qoff
    cudaEvent_t uploaded = nullptr, consumed = nullptr;
    bool upload_pending = false, consumed_valid = false;
};

}  // namespace

struct SdbEngine {
    SdbEngineConfig cfg{};
    SdbDynParams params{};
    bool dry = false;
    int device = 0;
    int n = 0, stride = 0, n_parts = 0;
    double distance = 0.0;  // pi / (2 wave_number)
    int depth = 3;
    int cap = 0;
    int desc_bytes_cap = 0;
    std::vector<ResultSlot> slots;
    std::vector<FrameSlot> frame_slots;
    int frame_i = 0;
    size_t qoff = 0;
    double* d_partials = nullptr;
    void* d_ov_ws = nullptr;
    void* d_inc_ws = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::unordered_map<int64_t, Origin*> origins;
    std::vector<Origin*> free_origins;  // dropped origins whose block is recycled
    std::vector<PoolFrame*> pool, pool_free;
    int64_t tickets = 0, created = 0;
    int64_t attempts = 0;  // step() calls, successful or not: stamps the group index cached in a PoolFrame
    int64_t frames = 0, updates = 0, h2d_bytes = 0, d2h_bytes = 0, device_bytes = 0;
    int last_segments = 0, last_inc_sets = 0, last_step_flags = 0;
    // descriptor block of the last step (test hook)
    std::vector<char> last_desc;
    int last_n_origins = 0, last_n_segments = 0, last_seg_off = 0;
    // scratch reused every step
    std::vector<Origin*> ordered, fresh;
    struct Group {
        PoolFrame* prev;
        std::vector<Origin*> members;
    };
    std::vector<Group> groups;
    std::vector<Origin*> by_key;
    std::vector<uint32_t> seg_set;
    struct SegRow {
        const float* prev;
        uint32_t flags;
        int first, count;
    };
    std::vector<SegRow> seg_rows;

    // ---- memory -----------------------------------------------------------------------------------------
    // Per-origin blocks and previous-sample frames come out of large slabs: a run with hundreds of origins then
    // costs a handful of cudaMalloc / cudaFree calls instead of hundreds (cudaFree synchronises the device, a few
    // ms each at the end of every process_file).  Blocks are recycled through free lists, never returned.
    struct Slab {
        char* base;
        size_t size, used;
    };
    std::vector<Slab> slabs;
    int arena_alloc(void** p, size_t bytes)
    {
        bytes = (bytes + 255) & ~size_t(255);
        if (slabs.empty() || slabs.back().used + bytes > slabs.back().size) {
            // geometric growth: an engine with one origin (the per-object Dynamics / Relaxations API creates one
            // engine per object) stays at two blocks' worth; a run with hundreds of origins gets ~log2 slabs
            size_t total = 0;
            for (const Slab& sl : slabs) total += sl.size;
            size_t want = std::min<size_t>(std::max<size_t>(2 * bytes, total), size_t(1) << 30);
            want = std::max(want, bytes);
            Slab sl{nullptr, want, 0};
            int rc = dev_alloc(reinterpret_cast<void**>(&sl.base), want);
            if (rc && want > bytes) {  // not enough memory for a whole slab: exactly what is needed
                sl.size = bytes;
                rc = dev_alloc(reinterpret_cast<void**>(&sl.base), bytes);
            }
            if (rc) return rc;
            slabs.push_back(sl);
        }
        *p = slabs.back().base + slabs.back().used;
        slabs.back().used += bytes;
        return SDB_OK;
    }

    int dev_alloc(void** p, size_t bytes)
    {
        if (bytes == 0) bytes = 16;
        if (dry) {
            *p = calloc(1, bytes);
            return *p ? SDB_OK : SDB_EINVAL;
        }
        int rc = sdb_check_cuda(cudaMalloc(p, bytes), "sdb_engine: cudaMalloc");
        if (!rc) device_bytes += (int64_t)bytes;
        return rc;
    }
    void dev_free(void* p)
    {
        if (!p) return;
        if (dry) free(p);
        else cudaFree(p);
    }
    int pin_alloc(void** p, size_t bytes)
    {
        if (bytes == 0) bytes = 16;
        if (dry) {
            *p = calloc(1, bytes);
            return *p ? SDB_OK : SDB_EINVAL;
        }
        return sdb_check_cuda(cudaHostAlloc(p, bytes, cudaHostAllocMapped | cudaHostAllocPortable), "sdb_engine: cudaHostAlloc");
    }
    void pin_free(void* p)
    {
        if (!p) return;
        if (dry) free(p);
        else cudaFreeHost(p);
    }
    int new_event(cudaEvent_t* e)
    {
        if (dry) {
            *e = nullptr;
            return SDB_OK;
        }
        return sdb_check_cuda(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "sdb_engine: cudaEventCreate");
    }

    size_t origin_block_layout(size_t* off_status, size_t* off_flags) const
    {
        auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
        size_t off = align(sizeof(float) * (3 * (size_t)stride + SDB_HISTORY_WORDS));
        *off_status = off;
        off = align(off + sizeof(uint32_t) * (size_t)SDB_MAX_TRACKERS * stride);
        *off_flags = off;
        return align(off + (size_t)stride);
    }

    Origin* new_origin(int64_t key, int64_t timestep, cudaStream_t s, int* rc)
    {
        Origin* o;
        if (!free_origins.empty()) {
            o = free_origins.back();
            free_origins.pop_back();
            char* block = o->block;
            *o = Origin{};
            o->block = block;
        } else {
            o = new Origin{};
            size_t off_status, off_flags;
            const size_t bytes = origin_block_layout(&off_status, &off_flags);
            *rc = arena_alloc(reinterpret_cast<void**>(&o->block), bytes);
            if (*rc) {
                delete o;
                return nullptr;
            }
        }
        size_t off_status, off_flags;
        origin_block_layout(&off_status, &off_flags);
        o->key = key;
        o->t0 = timestep;
        o->delta = reinterpret_cast<float*>(o->block);
        o->status = reinterpret_cast<uint32_t*>(o->block + off_status);
        o->flags = reinterpret_cast<uint8_t*>(o->block + off_flags);
        o->order = created++;
        // the threshold history behind the planes starts empty; the planes, flag bytes and passage times are
        // initialised by the kernel (SDB_ORIGIN_NEW)
        if (dry) memset(o->delta + 3 * (size_t)stride, 0, sizeof(uint32_t) * SDB_HISTORY_WORDS);
        else
            *rc = sdb_check_cuda(cudaMemsetAsync(o->delta + 3 * (size_t)stride, 0, sizeof(uint32_t) * SDB_HISTORY_WORDS, s),
                                 "sdb_engine: history clear");
        return o;
    }

    PoolFrame* new_pool_frame(int* rc)
    {
        if (!pool_free.empty()) {
            PoolFrame* f = pool_free.back();
            pool_free.pop_back();
            return f;
        }
        PoolFrame* f = new PoolFrame{};
        *rc = arena_alloc(reinterpret_cast<void**>(&f->planes), sizeof(float) * 4 * (size_t)stride);
        if (*rc) {
            delete f;
            return nullptr;
        }
        pool.push_back(f);
        return f;
    }
    void release(PoolFrame* f, int count = 1)
    {
        if (!f) return;
        f->refs -= count;
        if (f->refs == 0) pool_free.push_back(f);
    }

    // ---- rings sized for `cap` origins ------------------------------------------------------------------
    // Small per-origin buffers (descriptors, rows; a few hundred KB for 1024 origins) grow by factors of eight:
    // re-allocating page-locked memory in the middle of a run costs anything from a millisecond to hundreds of
    // them, so it happens at most twice on the way to 1000 origins, while an engine with a single origin (the
    // per-object API) stays small.  The two workspaces that scale with origins x molecules (per-tile partial sums,
    // overlap candidates) grow by factors of four.
    int small_cap = 0;
    int ensure_small(int n_active)
    {
        if (n_active <= small_cap) return SDB_OK;
        int want = 16;
        while (want < n_active) want <<= 3;  // 16 -> 128 -> 1024 -> ...: at most two re-allocations for 1000 origins
        if (!dry && small_cap > 0) {
            int rc = sdb_check_cuda(cudaDeviceSynchronize(), "sdb_engine: synchronize before growing");
            if (rc) return rc;
        }
        for (auto& sl : slots) {  // results of earlier steps still in the rings keep their rows
            if (sl.ticket >= 0 && sl.host_rows && sl.has_sums && sl.stash.empty() && sl.h_sums)
                sl.stash.assign(sl.h_sums, sl.h_sums + (size_t)sl.n_rows * SDB_NSUM);
            sl.pending = false;
        }
        const int n_desc = want * (int)sizeof(SdbOrigin) + (want + 8) * (int)sizeof(SdbSegment) + 64;
        for (auto& sl : slots) {
            pin_free(sl.h_desc);
            dev_free(sl.d_desc);
            dev_free(sl.d_sums);
            pin_free(sl.h_sums);
            dev_free(sl.d_isf);
            pin_free(sl.h_isf);
            sl.h_desc = sl.d_desc = nullptr;
            sl.d_sums = sl.h_sums = sl.d_isf = sl.h_isf = nullptr;
            int rc = pin_alloc(reinterpret_cast<void**>(&sl.h_desc), n_desc);
            if (!rc) rc = dev_alloc(reinterpret_cast<void**>(&sl.d_desc), n_desc);
            if (!rc) rc = dev_alloc(reinterpret_cast<void**>(&sl.d_sums), sizeof(double) * SDB_NSUM * want);
            if (!rc) rc = pin_alloc(reinterpret_cast<void**>(&sl.h_sums), sizeof(double) * SDB_NSUM * want);
            if (!rc && cfg.compute_isf) {
                rc = dev_alloc(reinterpret_cast<void**>(&sl.d_isf), sizeof(double) * want);
                if (!rc) rc = pin_alloc(reinterpret_cast<void**>(&sl.h_isf), sizeof(double) * want);
            }
            if (rc) return rc;
            if (!sl.done) {
                rc = new_event(&sl.done);
                if (rc) return rc;
            }
        }
        small_cap = want;
        desc_bytes_cap = n_desc;
        return SDB_OK;
    }

    int ensure_capacity(int n_active)
    {
        int rc = ensure_small(n_active);
        if (rc) return rc;
        if (n_active <= cap) return SDB_OK;
        int want = 16;
        while (want < n_active) want <<= 2;
        if (!dry && cap > 0) {
            rc = sdb_check_cuda(cudaDeviceSynchronize(), "sdb_engine: synchronize before growing");
            if (rc) return rc;
        }
        dev_free(d_partials);
        dev_free(d_ov_ws);
        d_partials = nullptr;
        d_ov_ws = nullptr;
        rc = dev_alloc(reinterpret_cast<void**>(&d_partials), sizeof(double) * SDB_NSUM * (size_t)n_parts * want);
        if (rc) return rc;
        if (cfg.compute_sums && cfg.overlap_k > 0) {
            rc = dev_alloc(&d_ov_ws, sdb_overlap_ws_bytes(want, n));
            if (rc) return rc;
        }
        cap = want;
        return SDB_OK;
    }

    // ---- frame staging ----------------------------------------------------------------------------------
    int stage(const SdbFrame& f, cudaStream_t s, const float** d_pos, const float** d_quat, FrameSlot** used)
    {
        *used = nullptr;
        if (f.where == SDB_MEM_DEVICE) {
            if ((reinterpret_cast<uintptr_t>(f.pos) & 15) || (reinterpret_cast<uintptr_t>(f.quat) & 15)) {
                sdb_set_error("sdb_engine_step: device frame pointers must be 16-byte aligned");
                return SDB_EINVAL;
            }
            *d_pos = f.pos;
            *d_quat = f.quat;
            h2d_bytes = 0;
            return SDB_OK;
        }
        if (f.where != SDB_MEM_HOST && f.where != SDB_MEM_PINNED) {
            sdb_set_error("sdb_engine_step: unknown memory kind %d", f.where);
            return SDB_EINVAL;
        }
        if (frame_slots.empty()) {
            frame_slots.resize(depth);
            for (auto& fs : frame_slots) {
                int rc = dev_alloc(reinterpret_cast<void**>(&fs.d), sizeof(float) * (qoff + 4 * (size_t)n));
                if (!rc) rc = new_event(&fs.uploaded);
                if (!rc) rc = new_event(&fs.consumed);
                if (rc) return rc;
            }
            if (!dry) {
                int rc = sdb_check_cuda(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking), "sdb_engine: copy stream");
                if (rc) return rc;
            }
        }
        frame_i = (frame_i + 1) % depth;
        FrameSlot& fs = frame_slots[frame_i];
        const float* src_pos = f.pos;
        const float* src_quat = f.quat;
        const size_t pos_bytes = sizeof(float) * 3 * (size_t)n, quat_bytes = sizeof(float) * 4 * (size_t)n;
        if (f.where == SDB_MEM_HOST) {
            if (!fs.h) {
                int rc = pin_alloc(reinterpret_cast<void**>(&fs.h), sizeof(float) * (qoff + 4 * (size_t)n));
                if (rc) return rc;
            }
            if (fs.upload_pending && !dry) {  // the upload that last read this pinned buffer
                int rc = sdb_check_cuda(cudaEventSynchronize(fs.uploaded), "sdb_engine: staging buffer");
                if (rc) return rc;
            }
            CopyPool::instance().copy(fs.h, f.pos, pos_bytes);
            if (f.quat) CopyPool::instance().copy(fs.h + qoff, f.quat, quat_bytes);
            src_pos = fs.h;
            src_quat = f.quat ? fs.h + qoff : nullptr;
        }
        if (!dry) {
            if (fs.consumed_valid) cudaStreamWaitEvent(copy_stream, fs.consumed, 0);  // kernels that read this buffer
            int rc = sdb_check_cuda(cudaMemcpyAsync(fs.d, src_pos, pos_bytes, cudaMemcpyHostToDevice, copy_stream),
                                    "sdb_engine: frame upload");
            if (!rc && src_quat)
                rc = sdb_check_cuda(cudaMemcpyAsync(fs.d + qoff, src_quat, quat_bytes, cudaMemcpyHostToDevice, copy_stream),
                                    "sdb_engine: frame upload");
            if (rc) return rc;
            cudaEventRecord(fs.uploaded, copy_stream);
            cudaStreamWaitEvent(s, fs.uploaded, 0);
        }
        fs.upload_pending = true;
        *d_pos = fs.d;
        *d_quat = src_quat ? fs.d + qoff : nullptr;
        *used = &fs;
        h2d_bytes = (int64_t)(pos_bytes + (src_quat ? quat_bytes : 0));
        return SDB_OK;
    }

    // ---- one frame ----------------------------------------------------------------------------------------
    int64_t step(const SdbFrame& f, cudaStream_t s)
    {
        const int n_active = f.n_keys;
        if (n_active < 0 || (n_active > 0 && !f.keys)) {
            sdb_set_error("sdb_engine_step: bad key list");
            return SDB_EINVAL;
        }
        const bool zero_step = (f.flags & SDB_FRAME_ZERO_STEP) != 0;
        const bool do_sums = cfg.compute_sums && !(f.flags & SDB_FRAME_NO_SUMS);
        const bool to_host = !(f.flags & SDB_FRAME_ROWS_DEVICE);
        if (!zero_step && n_active > 0 && !f.pos) {
            sdb_set_error("sdb_engine_step: no position array");
            return SDB_EINVAL;
        }
        // refuse before anything is created or re-sized: a failed step leaves the engine as it was
        for (int i = 0; i < n_active; ++i) {
            auto it = origins.find(f.keys[i]);
            if (it == origins.end()) {
                if (zero_step) {
                    sdb_set_error("sdb_engine_step: zero step on an origin that does not exist");
                    return SDB_EINVAL;
                }
                continue;
            }
            const int64_t td = it->second->updates == 0 ? 0 : f.timestep - it->second->t0;
            if (td < 0 || td >= (int64_t)SDB_NO_STATUS) {
                sdb_set_error("sdb_engine_step: time difference %lld outside [0, 2**32-1): passage times are 32-bit on the device",
                              (long long)td);
                return SDB_EINVAL;
            }
        }
        int rc = ensure_capacity(std::max(n_active, 1));
        if (rc) return rc;
        const int64_t ticket = tickets;
        ResultSlot& sl = slots[(size_t)(ticket % depth)];
        if (sl.pending && !dry) {
            // the descriptor upload and the row download of the step that used this slot `depth` steps ago
            rc = sdb_check_cuda(cudaEventSynchronize(sl.done), "sdb_engine: ring slot");
            if (rc) return rc;
        }
        sl.pending = false;
        sl.stash.clear();
        sl.n_rows = n_active;
        sl.has_sums = do_sums;
        sl.host_rows = do_sums && to_host;
        sl.has_isf = false;
        sl.timediffs.resize((size_t)n_active);
        sl.perm.resize((size_t)n_active);
        sl.identity = true;
        if (n_active == 0) {
            sl.ticket = ticket;
            ++tickets;
            return ticket;
        }

        // -- origins, grouped by previous sample (first-seen order) ------------------------------------------
        ordered.clear();
        fresh.clear();
        groups.clear();
        by_key.resize((size_t)n_active);
        const int64_t stamp = ++attempts;
        int orphan_group = -1;  // origins without a previous sample (only after a failed step)
        for (int i = 0; i < n_active; ++i) {
            auto it = origins.find(f.keys[i]);
            if (it == origins.end()) {
                if (zero_step) {
                    sdb_set_error("sdb_engine_step: zero step on an origin that does not exist");
                    return SDB_EINVAL;
                }
                Origin* o = new_origin(f.keys[i], f.timestep, s, &rc);
                if (!o) return rc ? rc : SDB_EINVAL;
                origins.emplace(f.keys[i], o);
                fresh.push_back(o);
                by_key[(size_t)i] = o;
            } else {
                Origin* o = it->second;
                by_key[(size_t)i] = o;
                if (o->updates == 0) {  // created by a step that failed before its launch: still new, and it starts now
                    o->t0 = f.timestep;
                    fresh.push_back(o);
                    continue;
                }
                int g;
                if (!o->prev) {
                    if (orphan_group < 0) {
                        orphan_group = (int)groups.size();
                        groups.push_back(Group{nullptr, {}});
                    }
                    g = orphan_group;
                } else if (o->prev->stamp == stamp) {
                    g = o->prev->group;
                } else {
                    o->prev->stamp = stamp;
                    o->prev->group = g = (int)groups.size();
                    groups.push_back(Group{o->prev, {}});
                }
                groups[(size_t)g].members.push_back(o);
            }
        }

        // -- stage the frame ------------------------------------------------------------------------------------
        const float *d_pos = nullptr, *d_quat = nullptr;
        FrameSlot* staged = nullptr;
        if (!zero_step) {
            rc = stage(f, s, &d_pos, &d_quat, &staged);
            if (rc) return rc;
        }
        const bool has_quat = d_quat != nullptr;

        // -- segments: runs of origins sharing a previous sample, split to fill the machine ----------------
        seg_rows.clear();
        const long target_blocks = 148L * 16;
        const long blocks_per_segment = (n_parts + 3) / 4;
        // a segment of fewer origins amortises a CTA's set-up (tracker table, pipeline fill, phase 1) over fewer
        // iterations, but small systems need more than (tiles / 4) x (origins / 8) CTAs to fill 148 SMs
        static const long seg_floor_env = [] { const char* e = getenv("SDB_SEG_MIN"); return e ? atol(e) : 0L; }();
        const long seg_floor = seg_floor_env > 0 ? seg_floor_env : (blocks_per_segment >= 592 ? 8L : (blocks_per_segment >= 148 ? 4L : 2L));
        auto split = [&](const std::vector<Origin*>& members, const float* prev, uint32_t flags) {
            const long len = (long)members.size();
            long size = cfg.chunk > 0 ? cfg.chunk : std::max(seg_floor, (len * blocks_per_segment + target_blocks - 1) / target_blocks);
            const long pieces = (len + size - 1) / size;
            size = (len + pieces - 1) / pieces;  // equal shares
            for (long first = 0; first < len; first += size) {
                const long count = std::min(size, len - first);
                seg_rows.push_back(SegRow{prev, flags, (int)ordered.size(), (int)count});
                ordered.insert(ordered.end(), members.begin() + first, members.begin() + first + count);
            }
        };
        for (auto& g : groups) {
            uint32_t flags = 0;
            if (zero_step) flags |= SDB_SEG_ZERO_STEP;
            else if (!has_quat || !g.prev || !g.prev->has_orientation) flags |= SDB_SEG_NO_ROTATION;
            split(g.members, g.prev ? g.prev->planes : nullptr, flags);
        }
        if (!fresh.empty()) {
            seg_rows.push_back(SegRow{nullptr, has_quat ? 0u : SDB_SEG_NO_ROTATION, (int)ordered.size(), (int)fresh.size()});
            ordered.insert(ordered.end(), fresh.begin(), fresh.end());
        }
        const int n_segments = (int)seg_rows.size();

        // -- previous samples shared by several segments: their increments are computed once per frame ------
        const void* inc_prev[SDB_MAX_INC_SETS] = {};
        uint32_t inc_flags[SDB_MAX_INC_SETS] = {};
        int inc_sets = 0;
        seg_set.assign((size_t)n_segments, 0u);
        const int inc_max = cfg.inc_max_per_segment > 0 ? cfg.inc_max_per_segment : 40;
        if (!(cfg.flags & SDB_ENGINE_NO_INCREMENTS) && !zero_step) {
            for (int i = 0; i < n_segments; ++i) {
                const SegRow& r = seg_rows[(size_t)i];
                if (!r.prev || seg_set[(size_t)i]) continue;
                int cnt = 0, members = 0;
                for (int j = 0; j < n_segments; ++j)
                    if (seg_rows[(size_t)j].prev == r.prev && seg_rows[(size_t)j].flags == r.flags) {
                        ++cnt;
                        members += seg_rows[(size_t)j].count;
                    }
                bool seen = false;
                for (int j = 0; j < i; ++j) seen = seen || (seg_rows[(size_t)j].prev == r.prev && seg_rows[(size_t)j].flags == r.flags);
                if (seen || cnt <= 1 || members > inc_max * cnt || inc_sets >= SDB_MAX_INC_SETS) continue;
                inc_prev[inc_sets] = r.prev;
                inc_flags[inc_sets] = r.flags;
                ++inc_sets;
                for (int j = 0; j < n_segments; ++j)
                    if (seg_rows[(size_t)j].prev == r.prev && seg_rows[(size_t)j].flags == r.flags) seg_set[(size_t)j] = (uint32_t)inc_sets;
            }
        }

        // -- descriptors ------------------------------------------------------------------------------------
        const uint32_t origin_flags = (cfg.flags & SDB_ENGINE_NO_PREDICT) ? 0u : SDB_ORIGIN_HISTORY;
        const bool use_ov = do_sums && cfg.overlap_k > 0 && d_ov_ws != nullptr;
        SdbOrigin* od = reinterpret_cast<SdbOrigin*>(sl.h_desc);
        bool all_history = use_ov && origin_flags && !zero_step;
        for (int i = 0; i < n_active; ++i) {
            Origin* o = ordered[(size_t)i];
            const int64_t td = f.timestep - o->t0;
            if (td < 0 || td >= (int64_t)SDB_NO_STATUS) {
                sdb_set_error("sdb_engine_step: time difference %lld outside [0, 2**32-1): passage times are 32-bit on the device",
                              (long long)td);
                // origins created by this call stay (state uninitialised but flagged new on their next step)
                return SDB_EINVAL;
            }
            od[i].d_delta = o->delta;
            od[i].d_flags = params.n_trackers > 0 ? o->flags : nullptr;
            od[i].d_status = params.n_trackers > 0 ? o->status : nullptr;
            od[i].timediff = (uint32_t)td;
            od[i].flags = origin_flags | (o->updates == 0 ? SDB_ORIGIN_NEW : (td < o->max_td ? SDB_ORIGIN_LITERAL : 0u));
            all_history = all_history && o->updates > 0 && o->ov_run >= 4 && f.timestep > o->last_ts;
            o->position = i;
        }
        const int seg_off = (int)((sizeof(SdbOrigin) * (size_t)n_active + 15) / 16 * 16);
        SdbSegment* sd = reinterpret_cast<SdbSegment*>(sl.h_desc + seg_off);
        int seg_count_max = 0;
        for (int i = 0; i < n_segments; ++i) {
            const SegRow& r = seg_rows[(size_t)i];
            sd[i].d_prev = r.prev;
            sd[i].first = r.first;
            sd[i].count = r.count;
            sd[i].flags = r.flags;
            sd[i].reserved = seg_set[(size_t)i];
            seg_count_max = std::max(seg_count_max, r.count);
        }
        const int desc_bytes = seg_off + (int)sizeof(SdbSegment) * n_segments;
        for (int i = 0; i < n_active; ++i) {
            const int p = by_key[(size_t)i]->position;
            sl.perm[(size_t)i] = p;
            sl.timediffs[(size_t)i] = (int64_t)od[p].timediff;
            sl.identity = sl.identity && p == i;
        }
        last_desc.assign(sl.h_desc, sl.h_desc + desc_bytes);
        last_n_origins = n_active;
        last_n_segments = n_segments;
        last_seg_off = seg_off;

        // -- the frame that becomes everybody's previous sample ------------------------------------------------
        PoolFrame* new_prev = nullptr;
        if (!zero_step) {
            new_prev = new_pool_frame(&rc);
            if (!new_prev) return rc ? rc : SDB_EINVAL;
            new_prev->has_orientation = has_quat;
        }
        // a step that fails from here on hands its (still unreferenced) pool frame back
        auto give_back = [&](int code) {
            if (new_prev) pool_free.push_back(new_prev);
            return code;
        };
        int step_flags = 0;
        if (all_history) step_flags |= SDB_STEP_PREDICTED;
        if (inc_sets > 0 && !d_inc_ws) {
            rc = dev_alloc(&d_inc_ws, sdb_dyn_increments_bytes(n, stride, SDB_MAX_INC_SETS));
            if (rc) return give_back(rc);
        }
        params.do_sums = do_sums ? 1 : 0;
        const bool do_struct = do_sums && cfg.n_sites > 0 && cfg.wave_number > 0;
        if (!dry) {
            rc = sdb_frame_step_impl(&params, d_pos, d_quat, new_prev ? new_prev->planes : nullptr, sl.h_desc, sl.d_desc, desc_bytes,
                                     seg_off, n_active, n_segments, seg_count_max, d_partials, sl.d_sums,
                                     sl.host_rows ? sl.h_sums : nullptr, use_ov ? d_ov_ws : nullptr, cap, use_ov ? cfg.overlap_k : 0,
                                     do_struct ? cfg.sites_xy : nullptr, do_struct ? cfg.n_sites : 0, distance, cfg.struct_mode,
                                     inc_sets ? inc_prev : nullptr, inc_sets ? inc_flags : nullptr, inc_sets, d_inc_ws, step_flags, 1, s);
            if (rc) return give_back(rc);
            if (do_sums && cfg.compute_isf && cfg.wave_number > 0) {
                rc = sdb_isf_origins(reinterpret_cast<const SdbOrigin*>(sl.d_desc), n_active, n, stride, cfg.wave_number,
                                     cfg.angular_resolution, sl.d_isf, s);
                if (rc) return give_back(rc);
                if (sl.host_rows) {
                    rc = sdb_check_cuda(cudaMemcpyAsync(sl.h_isf, sl.d_isf, sizeof(double) * (size_t)n_active, cudaMemcpyDeviceToHost, s),
                                        "sdb_engine: isf download");
                    if (rc) return give_back(rc);
                }
                sl.has_isf = true;
            }
            if (do_sums && f.d_rows_out) {  // rows in caller order straight to their destination (peer memory allowed)
                if (sl.identity && sdb_small_copies_on_sm()) {
                    rc = sdb_small_copy(f.d_rows_out, sl.d_sums, sizeof(double) * SDB_NSUM * (size_t)n_active, s);
                } else if (sl.identity) {
                    rc = sdb_check_cuda(cudaMemcpyAsync(f.d_rows_out, sl.d_sums, sizeof(double) * SDB_NSUM * (size_t)n_active,
                                                        cudaMemcpyDeviceToDevice, s),
                                        "sdb_engine: row copy");
                } else {
                    for (int i = 0; i < n_active && !rc; ++i)
                        rc = sdb_check_cuda(cudaMemcpyAsync(f.d_rows_out + (size_t)i * SDB_NSUM, sl.d_sums + (size_t)sl.perm[(size_t)i] * SDB_NSUM,
                                                            sizeof(double) * SDB_NSUM, cudaMemcpyDeviceToDevice, s),
                                            "sdb_engine: row copy");
                }
                if (rc) return give_back(rc);
            }
            if (staged) {
                cudaEventRecord(staged->consumed, s);
                staged->consumed_valid = true;
            }
            cudaEventRecord(sl.done, s);
            sl.pending = true;
        }
        last_segments = n_segments;
        last_inc_sets = inc_sets;
        last_step_flags = step_flags;
        d2h_bytes = sl.host_rows ? (int64_t)n_active * SDB_NSUM * 8 : 0;

        // -- bookkeeping ----------------------------------------------------------------------------------------
        if (!zero_step) {
            for (int i = 0; i < n_active; ++i) {
                Origin* o = ordered[(size_t)i];
                const int64_t td = (int64_t)od[i].timediff;
                release(o->prev);
                o->prev = new_prev;
                const bool forward = o->updates == 0 || f.timestep > o->last_ts;
                o->ov_run = (use_ov && forward) ? o->ov_run + 1 : 0;
                o->updates += 1;
                o->last_ts = f.timestep;
                if (td > o->max_td) o->max_td = td;
            }
            new_prev->refs += n_active;
            ++frames;
            updates += (int64_t)n_active * n;
        } else {
            for (Origin* o : ordered) o->ov_run = 0;
        }
        sl.ticket = ticket;
        ++tickets;
        return ticket;
    }

    int wait(int64_t ticket, double* h_rows, int64_t* h_timediffs, double* h_isf)
    {
        if (ticket < 0 || ticket >= tickets || ticket < tickets - depth) {
            sdb_set_error("sdb_engine_wait: ticket %lld is no longer in the ring (depth %d)", (long long)ticket, depth);
            return SDB_EINVAL;
        }
        ResultSlot& sl = slots[(size_t)(ticket % depth)];
        if (sl.ticket != ticket) {
            sdb_set_error("sdb_engine_wait: ticket %lld was overwritten", (long long)ticket);
            return SDB_EINVAL;
        }
        if (sl.pending && !dry && (h_rows || h_isf)) {  // (time differences alone need no device result)
            int rc = sdb_check_cuda(cudaEventSynchronize(sl.done), "sdb_engine_wait");
            if (rc) return rc;
            sl.pending = false;
        }
        for (int i = 0; i < sl.n_rows; ++i) {
            if (h_timediffs) h_timediffs[i] = sl.timediffs[(size_t)i];
            if (h_rows && sl.host_rows) {
                const double* src = sl.stash.empty() ? sl.h_sums : sl.stash.data();
                memcpy(h_rows + (size_t)i * SDB_NSUM, src + (size_t)sl.perm[(size_t)i] * SDB_NSUM, sizeof(double) * SDB_NSUM);
            }
            if (h_isf && sl.has_isf && sl.host_rows && sl.h_isf) h_isf[i] = sl.h_isf[sl.perm[(size_t)i]];
        }
        if (h_rows && !sl.host_rows && sl.n_rows > 0) {
            sdb_set_error("sdb_engine_wait: step %lld has no host rows", (long long)ticket);
            return SDB_EINVAL;
        }
        return SDB_OK;
    }

    void destroy()
    {
        if (!dry) cudaDeviceSynchronize();
        for (auto& sl : slots) {
            pin_free(sl.h_desc);
            dev_free(sl.d_desc);
            dev_free(sl.d_sums);
            pin_free(sl.h_sums);
            dev_free(sl.d_isf);
            pin_free(sl.h_isf);
            if (sl.done) cudaEventDestroy(sl.done);
        }
        for (auto& fs : frame_slots) {
            pin_free(fs.h);
            dev_free(fs.d);
            if (fs.uploaded) cudaEventDestroy(fs.uploaded);
            if (fs.consumed) cudaEventDestroy(fs.consumed);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        dev_free(d_partials);
        dev_free(d_ov_ws);
        dev_free(d_inc_ws);
        for (auto& kv : origins) delete kv.second;
        for (Origin* o : free_origins) delete o;
        for (PoolFrame* f : pool) delete f;
        for (Slab& sl : slabs) dev_free(sl.base);
    }
};

namespace {
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const SdbEngine* e)
    {
        if (e->dry) return;
        int cur = 0;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != e->device) {
            prev = cur;
            cudaSetDevice(e->device);
        }
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace

extern "C" int sdb_engine_create(const SdbEngineConfig* c, SdbEngine** out)
{
    if (!c || !out) {
        sdb_set_error("sdb_engine_create: null argument");
        return SDB_EINVAL;
    }
    *out = nullptr;
    if (c->n <= 0) {
        sdb_set_error("sdb_engine_create: Position must contain values, has length of 0.");
        return SDB_EINVAL;
    }
    if (c->n_trackers < 0 || c->n_trackers > SDB_MAX_TRACKERS || c->n_sites < 0 || c->n_sites > 8 || c->overlap_k < 0 ||
        c->overlap_k > c->n) {
        sdb_set_error("sdb_engine_create: bad configuration (n_trackers=%d n_sites=%d overlap_k=%d)", c->n_trackers, c->n_sites,
                      c->overlap_k);
        return SDB_EINVAL;
    }
    SdbEngine* e = new SdbEngine{};
    e->cfg = *c;
    e->dry = (c->flags & SDB_ENGINE_DRY) != 0;
    e->n = c->n;
    e->stride = (c->n + 3) / 4 * 4;
    e->n_parts = sdb_dyn_num_parts(c->n);
    e->qoff = (3 * (size_t)c->n + 3) / 4 * 4;
    e->depth = std::max(2, c->ring_depth > 0 ? c->ring_depth : 3);
    e->distance = c->wave_number > 0 ? 3.14159265358979323846 / (2.0 * c->wave_number) : 0.0;
    if (!e->dry && sdb_check_cuda(cudaGetDevice(&e->device), "sdb_engine_create: cudaGetDevice")) {
        delete e;
        return SDB_ECUDA;
    }
    SdbDynParams& p = e->params;
    p.n = e->n;
    p.stride = e->stride;
    p.lx = c->box[0];
    p.ly = c->box[1];
    p.xy = c->box[3];
    p.com_cut_lt = c->wave_number > 0 ? c->com_cut_lt : -1.0f;
    p.n_trackers = c->n_trackers;
    p.do_sums = c->compute_sums ? 1 : 0;
    for (int i = 0; i < c->n_trackers; ++i) p.trackers[i] = c->trackers[i];
    e->slots.resize((size_t)e->depth);
    *out = e;
    return SDB_OK;
}

extern "C" void sdb_engine_destroy(SdbEngine* e)
{
    if (!e) return;
    {
        DeviceGuard guard(e);
        e->destroy();
    }
    delete e;
}

extern "C" int sdb_engine_set_trackers(SdbEngine* e, const SdbTracker* trackers, int32_t n_trackers)
{
    if (!e || n_trackers < 0 || n_trackers > SDB_MAX_TRACKERS || (n_trackers > 0 && !trackers)) {
        sdb_set_error("sdb_engine_set_trackers: bad argument");
        return SDB_EINVAL;
    }
    DeviceGuard guard(e);
    e->params.n_trackers = n_trackers;
    for (int i = 0; i < n_trackers; ++i) e->params.trackers[i] = trackers[i];
    e->cfg.n_trackers = n_trackers;
    // every origin's passage state starts afresh, its accumulated motion is kept (relaxations.py:108-133)
    if (!e->dry && !e->origins.empty()) {  // rare call: simply let the device drain, then rewrite the state
        int rc = sdb_check_cuda(cudaDeviceSynchronize(), "sdb_engine_set_trackers");
        if (rc) return rc;
    }
    for (auto& kv : e->origins) {
        Origin* o = kv.second;
        if (o->updates == 0) continue;  // initialised by the kernel on its first step
        if (e->dry) {
            memset(o->flags, 0, (size_t)e->stride);
            memset(o->status, 0xff, sizeof(uint32_t) * (size_t)SDB_MAX_TRACKERS * e->stride);
        } else {
            int rc = sdb_check_cuda(cudaMemset(o->flags, 0, (size_t)e->stride), "sdb_engine_set_trackers");
            if (!rc) rc = sdb_check_cuda(cudaMemset(o->status, 0xff, sizeof(uint32_t) * (size_t)SDB_MAX_TRACKERS * e->stride),
                                         "sdb_engine_set_trackers");
            if (rc) return rc;
        }
    }
    if (!e->dry && !e->origins.empty()) return sdb_check_cuda(cudaDeviceSynchronize(), "sdb_engine_set_trackers");
    return SDB_OK;
}

extern "C" int64_t sdb_engine_step(SdbEngine* e, const SdbFrame* frame, void* stream)
{
    if (!e || !frame) {
        sdb_set_error("sdb_engine_step: null argument");
        return SDB_EINVAL;
    }
    DeviceGuard guard(e);
    return e->step(*frame, static_cast<cudaStream_t>(stream));
}

extern "C" int sdb_engine_wait(SdbEngine* e, int64_t ticket, double* h_rows, int64_t* h_timediffs, double* h_isf)
{
    if (!e) return SDB_EINVAL;
    DeviceGuard guard(e);
    return e->wait(ticket, h_rows, h_timediffs, h_isf);
}

extern "C" int sdb_engine_device_rows(SdbEngine* e, int64_t ticket, double** d_rows, int32_t* h_perm)
{
    if (!e || ticket < 0 || ticket >= e->tickets || ticket < e->tickets - e->depth) {
        sdb_set_error("sdb_engine_device_rows: bad ticket");
        return SDB_EINVAL;
    }
    ResultSlot& sl = e->slots[(size_t)(ticket % e->depth)];
    if (sl.ticket != ticket || !sl.has_sums) {
        sdb_set_error("sdb_engine_device_rows: step has no rows");
        return SDB_EINVAL;
    }
    if (d_rows) *d_rows = sl.d_sums;
    if (h_perm)
        for (int i = 0; i < sl.n_rows; ++i) h_perm[i] = sl.perm[(size_t)i];
    return SDB_OK;
}

extern "C" int sdb_engine_run(SdbEngine* e, const SdbFrame* frames, int32_t n_frames, double* h_rows, int64_t* h_timediffs,
                              double* h_isf, void* stream)
{
    if (!e || (n_frames > 0 && !frames)) {
        sdb_set_error("sdb_engine_run: null argument");
        return SDB_EINVAL;
    }
    DeviceGuard guard(e);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    std::vector<size_t> offset((size_t)n_frames + 1, 0);
    for (int i = 0; i < n_frames; ++i) offset[(size_t)i + 1] = offset[(size_t)i] + (size_t)std::max(frames[i].n_keys, 0);
    std::vector<int64_t> ticket((size_t)n_frames, -1);
    auto collect = [&](int i) -> int {
        const SdbFrame& f = frames[i];
        const bool rows = e->cfg.compute_sums && !(f.flags & (SDB_FRAME_NO_SUMS | SDB_FRAME_ROWS_DEVICE));
        return e->wait(ticket[(size_t)i], rows && h_rows ? h_rows + offset[(size_t)i] * SDB_NSUM : nullptr,
                       h_timediffs ? h_timediffs + offset[(size_t)i] : nullptr, rows && h_isf ? h_isf + offset[(size_t)i] : nullptr);
    };
    int collected = 0;
    for (int i = 0; i < n_frames; ++i) {
        // keep depth - 1 frames in flight: the slot of frame i is that of frame i - depth
        while (collected <= i - (e->depth - 1)) {
            int rc = collect(collected++);
            if (rc) return rc;
        }
        const int64_t t = e->step(frames[i], s);
        if (t < 0) return (int)t;
        ticket[(size_t)i] = t;
    }
    while (collected < n_frames) {
        int rc = collect(collected++);
        if (rc) return rc;
    }
    return SDB_OK;
}

extern "C" int sdb_engine_drop(SdbEngine* e, int64_t key)
{
    if (!e) return SDB_EINVAL;
    auto it = e->origins.find(key);
    if (it == e->origins.end()) {
        sdb_set_error("sdb_engine_drop: unknown key %lld", (long long)key);
        return SDB_EINVAL;
    }
    Origin* o = it->second;
    e->release(o->prev);
    o->prev = nullptr;
    e->origins.erase(it);
    e->free_origins.push_back(o);  // (stream order keeps a recycled block safe: one stream per engine)
    return SDB_OK;
}

extern "C" int sdb_engine_origin(SdbEngine* e, int64_t key, SdbOriginInfo* out)
{
    if (!e || !out) return SDB_EINVAL;
    auto it = e->origins.find(key);
    if (it == e->origins.end()) {
        sdb_set_error("sdb_engine_origin: unknown key %lld", (long long)key);
        return SDB_EINVAL;
    }
    const Origin* o = it->second;
    out->timestep = o->t0;
    out->updates = o->updates;
    out->max_timediff = o->max_td;
    out->d_delta = o->delta;
    out->d_flags = e->params.n_trackers > 0 ? o->flags : nullptr;
    out->d_status = e->params.n_trackers > 0 ? o->status : nullptr;
    out->d_prev = o->prev ? o->prev->planes : nullptr;
    out->prev_has_orientation = o->prev ? (int)o->prev->has_orientation : 0;
    out->stride = e->stride;
    return SDB_OK;
}

extern "C" int32_t sdb_engine_num_origins(SdbEngine* e) { return e ? (int32_t)e->origins.size() : 0; }

extern "C" int32_t sdb_engine_keys(SdbEngine* e, int64_t* h_keys, int32_t capacity)
{
    if (!e) return 0;
    std::vector<const Origin*> all;
    all.reserve(e->origins.size());
    for (auto& kv : e->origins) all.push_back(kv.second);
    std::sort(all.begin(), all.end(), [](const Origin* a, const Origin* b) { return a->order < b->order; });
    const int32_t count = (int32_t)all.size();
    if (h_keys)
        for (int32_t i = 0; i < count && i < capacity; ++i) h_keys[i] = all[(size_t)i]->key;
    return count;
}

extern "C" int sdb_engine_stats(SdbEngine* e, SdbEngineStats* out)
{
    if (!e || !out) return SDB_EINVAL;
    out->frames = e->frames;
    out->updates = e->updates;
    out->h2d_bytes = e->h2d_bytes;
    out->d2h_bytes = e->d2h_bytes;
    out->pool_frames = (int64_t)e->pool.size();
    out->pool_free = (int64_t)e->pool_free.size();
    out->device_bytes = e->device_bytes;
    out->last_segments = e->last_segments;
    out->last_inc_sets = e->last_inc_sets;
    out->last_step_flags = e->last_step_flags;
    out->capacity = e->cap;
    out->d_ov_ws = e->d_ov_ws;
    out->d_inc_ws = e->d_inc_ws;
    return SDB_OK;
}

extern "C" int32_t sdb_engine_last_descriptors(SdbEngine* e, void* h_out, int32_t capacity, int32_t* n_origins, int32_t* n_segments,
                                               int32_t* segments_offset)
{
    if (!e) return 0;
    if (n_origins) *n_origins = e->last_n_origins;
    if (n_segments) *n_segments = e->last_n_segments;
    if (segments_offset) *segments_offset = e->last_seg_off;
    const int32_t bytes = (int32_t)std::min<size_t>(e->last_desc.size(), (size_t)std::max(capacity, 0));
    if (h_out && bytes > 0) memcpy(h_out, e->last_desc.data(), (size_t)bytes);
    return bytes;
}

// Test / diagnostics hook: the engine's multi-threaded host copy (pageable -> pinned staging) on caller buffers.
extern "C" void sdb_test_host_copy(void* dst, const void* src, size_t bytes) { CopyPool::instance().copy(dst, src, bytes); }
