// sdb_internal.cuh -- shared declarations of the library's translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>

#include "../../include/sdb200.h"
#include "sdb_math.cuh"

// ---- tiling of the fused dynamics kernel ------------------------------------------------------
// A warp owns 256 consecutive molecules: two groups of 128, each lane 4 consecutive molecules per
// group, so every per-origin plane (dx, dy, dtheta: f32; flags: u8) is read and written with fully
// coalesced 16-byte (4-byte for flags) accesses.  Warps never synchronise with each other.
constexpr int SDB_LANE_MOLS = 4;
constexpr int SDB_GROUPS = 2;
constexpr int SDB_WARP_MOLS = 32 * SDB_LANE_MOLS * SDB_GROUPS;  // 256
constexpr int SDB_BLOCK_WARPS = 4;
constexpr int SDB_BLOCK_THREADS = 32 * SDB_BLOCK_WARPS;
constexpr int SDB_NFSUM = 10;  // fp64 sums reduced through shared memory (SDB_S_DISP..SDB_S_D2R2)

void sdb_set_error(const char* fmt, ...);
int sdb_check_cuda(cudaError_t err, const char* what);
void sdb_count_launch(int n = 1);
const double (*sdb_device_atan_table())[2];  // device pointer to the 1024-entry cos/sin table
const double (*sdb_host_atan_table())[2];

// sdb_dyn_step in two halves (the second can run on another stream): csrc/dyn_step.cu
int sdb_dyn_step_launch(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, float* d_cur_compact,
                        const SdbOrigin* d_origins, int32_t n_origins, const SdbSegment* d_segments, int32_t n_segments,
                        double* d_partials, double* d_sums, void* d_ov_ws, int32_t ov_ws_origins, void* d_inc_ws,
                        int32_t inc_sets, int32_t finalize, void* stream);
int sdb_dyn_finalize(const double* d_partials, int32_t n, double* d_sums, int32_t n_origins, int32_t with_ov, void* stream);

// tiny copies on the SMs instead of the copy engines (SDB_SMALL_COPIES=kernel): csrc/api.cu
bool sdb_small_copies_on_sm();
int sdb_small_copy(void* dst, const void* src, size_t bytes, void* stream);

// sdb_struct with the clear of the SDB_S_STRUCT slots optional: csrc/struct.cu
int sdb_struct_launch(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride, const float* h_sites_xy,
                      int32_t n_sites, double dist, int32_t mode, double* d_sums, int32_t clear, void* stream);

// row sums + overlap selection + intersection in one launch: csrc/overlap.cu
int sdb_overlap_resolve_fused(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride, int32_t k, void* d_ws,
                              int32_t ws_origins, const double* d_partials, double* d_sums, void* stream);

int sdb_overlap_prepare_ex(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, const SdbOrigin* d_origins,
                           int32_t n_origins, const SdbSegment* d_segments, int32_t n_segments, int32_t max_segment_count,
                           int32_t k, void* d_ws, int32_t ws_origins, int32_t skip_sample, void* stream);

#define SDB_CHECK_LAUNCH(what)                                       \
    do {                                                             \
        sdb_count_launch();                                          \
        int _rc = sdb_check_cuda(cudaGetLastError(), what);          \
        if (_rc) return _rc;                                         \
    } while (0)

__device__ __forceinline__ float4 sdb_ldg4(const float* p)
{
    return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming (evict-first) 16-byte accesses for state that is touched once per frame
__device__ __forceinline__ float4 sdb_ld_stream4(const float* p)
{
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void sdb_st_stream4(float* p, float4 v)
{
    __stcs(reinterpret_cast<float4*>(p), v);
}

// ---- freud Box (any dimension, triclinic) -----------------------------------------------------------
struct SdbBox3 {
    float lx, ly, lz, xy, xz, yz;
    int is2d;
};

// freud Box::wrap (oracle/freud_shim.py), general triclinic form
__device__ __forceinline__ void sdb_wrap3(const SdbBox3& b, float vx, float vy, float vz, float& ox, float& oy, float& oz)
{
    const float lox = -__fmul_rn(b.lx, 0.5f), loy = -__fmul_rn(b.ly, 0.5f), loz = -__fmul_rn(b.lz, 0.5f);
    float dx = __fsub_rn(vx, lox), dy = __fsub_rn(vy, loy), dz = __fsub_rn(vz, loz);
    dx = __fsub_rn(dx, __fadd_rn(__fmul_rn(__fsub_rn(b.xz, __fmul_rn(b.yz, b.xy)), vz), __fmul_rn(b.xy, vy)));
    dy = __fsub_rn(dy, __fmul_rn(b.yz, vz));
    const float fx = sdb_frac_wrap(dx, b.lx), fy = sdb_frac_wrap(dy, b.ly);
    const float fz = b.is2d ? 0.0f : sdb_frac_wrap(dz, b.lz);
    oz = __fadd_rn(loz, __fmul_rn(fz, b.lz));
    oy = __fadd_rn(loy, __fmul_rn(fy, b.ly));
    ox = __fadd_rn(lox, __fmul_rn(fx, b.lx));
    ox = __fadd_rn(ox, __fadd_rn(__fmul_rn(b.xy, oy), __fmul_rn(b.xz, oz)));
    oy = __fadd_rn(oy, __fmul_rn(b.yz, oz));
    if (b.is2d) oz = 0.0f;
}

// ---- TrackedMotion.add for one molecule and one origin (dynamics/_util.py:149-153) ---------------
// Shared by the fused kernel and the overlap sample kernel so both see bit-identical keys.
__device__ __forceinline__ void sdb_apply_increment(float& dx, float& dy, float& dt, float wx, float wy, double alpha,
                                                    float& d2, float& th)
{
    dx = SDB_FSUB(dx, wx);
    dy = SDB_FSUB(dy, wy);
    dt = (float)((double)dt + alpha);
    // np.square(delta).sum(axis=1) in f32, one rounding per operation
    d2 = SDB_FADD(SDB_FMUL(dx, dx), SDB_FMUL(dy, dy));
    th = fabsf(dt);
}

// ---- workspace of the bracketed `overlap` selection (csrc/overlap.cu) -----------------------------
// Keys: translation d2 = |dr|^2 (f32, monotone in the reference's |dr|), rotation |dtheta|; both >= 0.
constexpr int SDB_OV_SAMPLES = 16384;  // target number of sampled molecules per origin
constexpr int SDB_OV_STAGE = 96;       // candidates one warp can stage per (origin, 256-molecule tile)

struct SdbOvWorkspace {
    float* brackets;     // [origins][4]: lo_r, hi_r, lo_t, hi_t (inclusive candidate intervals)
    float* samples;      // [origins][2][n_samples]
    float* increments;   // [segments <= origins][n_samples][4]: wx, wy, alpha (f64 in two words)
    uint32_t* cand_cnt;  // [origins][2]: number of candidates, overflow flag
    uint32_t* cand_idx;  // [origins][cand_cap]  molecule index        } dense, unordered
    uint32_t* cand_r;    // [origins][cand_cap]  bits of d2            }
    uint32_t* cand_t;    // [origins][cand_cap]  bits of |dtheta|      }
    uint32_t* select;    // [origins][8]: thresholds and index cuts found by the resolve kernel
    uint32_t* fail;      // [1 + origins]: count, then origin indices whose bracket failed
    uint32_t* exact;     // scratch of the exact (radix select) kernel, one origin at a time
    int n_samples, sample_stride, n_parts, cand_cap;
};

// ---- threshold history of one origin (SDB_ORIGIN_HISTORY): words after the three delta planes -------
//   [0] thr_r1  [1] thr_t1  [2] td1   exact thresholds (f32 bits) and time difference of the last frame
//   [3] thr_r0  [4] thr_t0  [5] td0   ... of the frame before
//   [6] valid   number of consecutive successful selections recorded (0, 1, 2)
//   [7] widen   log2 of the bracket widening factor (grows when a predicted bracket misses the threshold)
enum { SDB_H_R1 = 0, SDB_H_T1 = 1, SDB_H_TD1 = 2, SDB_H_R0 = 3, SDB_H_T0 = 4, SDB_H_TD0 = 5, SDB_H_VALID = 6, SDB_H_WIDEN = 7 };

__device__ __forceinline__ uint32_t* sdb_history(const SdbOrigin& org, long stride)
{
    return (org.flags & SDB_ORIGIN_HISTORY) ? reinterpret_cast<uint32_t*>(org.d_delta + 3 * stride) : nullptr;
}

// Bracket predicted by extrapolating (in time) the last two exact thresholds of one key array.  Two
// models are evaluated and the bracket is their envelope widened by 2^widen max(|step| / 10, 8 thr / sqrt(n)):
//   linear in the key            -- |dr|^2 of diffusing molecules, |dtheta| of freely rotating ones;
//   linear in the "other power"  -- key^2 for |dtheta| (rotational diffusion: |dtheta| ~ sqrt(t)),
//                                   sqrt(key) for |dr|^2 (ballistic motion: |dr|^2 ~ t^2).
// Young origins, whose thresholds move by a large fraction of their value per frame, get a bracket of
// ~0.3 |step| instead of the |step| a single linear model needs to be safe; old origins reach the noise
// floor (threshold jitter <= 4 / sqrt(n) relative).  Any bracket is safe -- the selection verifies that
// the k-th largest key lies inside and falls back to the exact kernel otherwise, and `widen` grows
// whenever a bracket misses.
__device__ __forceinline__ bool sdb_predict_bracket(const uint32_t* hist, uint32_t td, int arr, float inv_sqrt_n,
                                                    float& lo, float& hi)
{
    if (hist == nullptr || hist[SDB_H_VALID] < 2u) return false;
    const uint32_t td1 = hist[SDB_H_TD1], td0 = hist[SDB_H_TD0];
    if (!(td > td1 && td1 > td0)) return false;
    const float t1 = __uint_as_float(hist[arr ? SDB_H_T1 : SDB_H_R1]), t0 = __uint_as_float(hist[arr ? SDB_H_T0 : SDB_H_R0]);
    if (!(t1 > 0.0f) || !(t0 > 0.0f) || !(t1 < INFINITY)) return false;
    const float ratio = (float)(td - td1) / (float)(td1 - td0);
    const float step = (t1 - t0) * ratio;
    const float lin = t1 + step;
    float alt;
    if (arr) {
        alt = sqrtf(fmaxf(t1 * t1 + (t1 * t1 - t0 * t0) * ratio, 0.0f));
    } else {
        const float s1 = sqrtf(t1), s = s1 + (s1 - sqrtf(t0)) * ratio;
        alt = s > 0.0f ? s * s : 0.0f;
    }
    const float w = fmaxf(0.1f * fabsf(step), 8.0f * inv_sqrt_n * t1) * (float)(1u << min(hist[SDB_H_WIDEN], 6u));
    lo = fmaxf(fminf(lin, alt) - w, 0.0f);
    hi = fmaxf(lin, alt) + w;
    return hi > lo && hi < INFINITY;
}

inline size_t sdb_ov_exact_words(int n) { return 2 * 2048 + 2 * 8 + 2 * (size_t)((n + 255) / 256) + 64; }

inline size_t sdb_ov_layout(void* base, int max_origins, int n, SdbOvWorkspace* w)
{
    const int n_parts = (n + SDB_WARP_MOLS - 1) / SDB_WARP_MOLS;
    // the sample: groups of 4 consecutive molecules, one group every 4 * stride molecules
    const int stride = n / SDB_OV_SAMPLES > 1 ? n / SDB_OV_SAMPLES : 1;
    const int n_groups = (n + 4 * stride - 1) / (4 * stride);
    const int n_samples = 4 * n_groups;
    // expected candidates: ~2 x 4.3 % of n; 4x head room (and never less than one staging buffer)
    const int cand_cap = ((n / 3 > 4 * SDB_OV_STAGE ? n / 3 : 4 * SDB_OV_STAGE) + 63) & ~63;
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    size_t off = 0;
    char* b = static_cast<char*>(base);
    const size_t o_br = off;   off = align(off + sizeof(float) * 4 * (size_t)max_origins);
    const size_t o_sm = off;   off = align(off + sizeof(float) * 2 * (size_t)n_samples * max_origins);
    const size_t o_in = off;   off = align(off + sizeof(float) * 4 * (size_t)n_samples * max_origins);
    const size_t o_cc = off;   off = align(off + sizeof(uint32_t) * 2 * (size_t)max_origins);
    const size_t o_ci = off;   off = align(off + sizeof(uint32_t) * (size_t)cand_cap * max_origins);
    const size_t o_cr = off;   off = align(off + sizeof(uint32_t) * (size_t)cand_cap * max_origins);
    const size_t o_ct = off;   off = align(off + sizeof(uint32_t) * (size_t)cand_cap * max_origins);
    const size_t o_se = off;   off = align(off + sizeof(uint32_t) * 8 * (size_t)max_origins);
    const size_t o_fl = off;   off = align(off + sizeof(uint32_t) * (size_t)(1 + max_origins));
    const size_t o_ex = off;   off = align(off + sizeof(uint32_t) * sdb_ov_exact_words(n));
    if (w) {
        w->brackets = reinterpret_cast<float*>(b + o_br);
        w->samples = reinterpret_cast<float*>(b + o_sm);
        w->increments = reinterpret_cast<float*>(b + o_in);
        w->cand_cnt = reinterpret_cast<uint32_t*>(b + o_cc);
        w->cand_idx = reinterpret_cast<uint32_t*>(b + o_ci);
        w->cand_r = reinterpret_cast<uint32_t*>(b + o_cr);
        w->cand_t = reinterpret_cast<uint32_t*>(b + o_ct);
        w->select = reinterpret_cast<uint32_t*>(b + o_se);
        w->fail = reinterpret_cast<uint32_t*>(b + o_fl);
        w->exact = reinterpret_cast<uint32_t*>(b + o_ex);
        w->n_samples = n_samples;
        w->sample_stride = stride;
        w->n_parts = n_parts;
        w->cand_cap = cand_cap;
    }
    return off;
}
