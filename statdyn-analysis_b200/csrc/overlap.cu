// overlap.cu -- `overlap` of Dynamics.compute_all (reference dynamics/dynamics.py:380-382,431-445):
//   k = int(N * 0.1); |argsort(|dr|)[-k:]  intersect  argsort(|dtheta|)[-k:]| / k
//
// The reference sorts both arrays completely for every (frame, origin).  Here the k-th largest value
// of each array is found exactly without sorting and without an extra pass over the state:
//
//   1. sample   a strided sample (~16k molecules) of the *updated* keys of every active origin is
//               computed from the old state and the frame increments (same arithmetic as the fused
//               kernel, so the keys are bit-identical);
//   2. bracket  one block per (origin, array) selects two order statistics of the sample that enclose
//               the k-th largest of the full array with probability 1 - 1e-9: the bracket [lo, hi];
//   3. fused    dyn_step_kernel (dyn_step.cu) counts the keys above hi and appends the ~3 % of
//               molecules inside a bracket to a per-warp-tile candidate list;
//   4. resolve  one block per origin selects the exact threshold among the candidates (radix select
//               on the key offset from lo), breaks ties by molecule index, counts the intersection;
//   5. exact    origins whose bracket failed (all keys equal, overflow, rare sampling misses) are
//               redone by a kernel that radix-selects over the whole array, one block per origin
//               (3 digits of the f32 bit pattern).  It also serves mobile_overlap() on caller arrays.
//
// Keys: translation d2 = |dr|^2 = fl32(dx^2 + dy^2) -- monotone in the reference's |dr| = sqrt(d2) --
// and rotation |dtheta|.  Members of a top-k set are the keys above the threshold plus, among keys
// equal to it, those with the largest molecule indices (numpy's introsort leaves the choice among
// equal keys unspecified -- parity trap T4; for all-equal arrays this rule gives identical sets for
// both arrays, i.e. overlap = 1, like the reference).
#include <cstdlib>

#include "sdb_internal.cuh"

namespace {

constexpr int OV_BINS = 2048;

// =================================================================================================
// 1. sample
// =================================================================================================
struct SampleArgs {
    SdbDynParams p;
    const float* pos;
    const float* quat;
    const SdbOrigin* origins;
    const SdbSegment* segments;
    const double (*atan_tab)[2];
    SdbOvWorkspace w;
};

// The sample consists of groups of 4 consecutive molecules (one 16-byte vector per state plane),
// group g covering molecules 4 g stride .. 4 g stride + 3; sample slot i = 4 g + e.
// Stage 1, grid = (group tiles, segments): frame increments (wx, wy, alpha) of the sampled
// molecules, once per segment.
__global__ void __launch_bounds__(128) ov_increment_kernel(const SampleArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.w.n_samples) return;
    const SdbSegment seg = a.segments[blockIdx.y];
    const long j = (long)(i >> 2) * 4 * a.w.sample_stride + (i & 3);
    const long stride = a.p.stride;
    const bool zero_step = (seg.flags & SDB_SEG_ZERO_STEP) != 0 || a.pos == nullptr;
    float wx = 0.f, wy = 0.f;
    double alpha = 0.0;
    if (!zero_step && j < a.p.n) {
        const float cx = __ldg(a.pos + 3 * j), cy = __ldg(a.pos + 3 * j + 1);
        float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
        if (a.quat) q = sdb_ldg4(a.quat + 4 * j);
        float px = cx, py = cy, pw = q.x, pz = q.w;
        const float* prev = static_cast<const float*>(seg.d_prev);
        if (prev) {
            px = __ldg(prev + j);
            py = __ldg(prev + stride + j);
            pw = __ldg(prev + 2 * stride + j);
            pz = __ldg(prev + 3 * stride + j);
        }
        sdb_wrap2d(a.p.lx, a.p.ly, a.p.xy, SDB_FSUB(px, cx), SDB_FSUB(py, cy), wx, wy);
        if ((seg.flags & SDB_SEG_NO_ROTATION) == 0) {
            int bad;
            alpha = sdb_rot_increment(q.x, q.w, pw, pz, a.atan_tab, bad);
        }
    }
    float4 out;
    out.x = wx;
    out.y = wy;
    out.z = __int_as_float(__double2loint(alpha));
    out.w = __int_as_float(__double2hiint(alpha));
    reinterpret_cast<float4*>(a.w.increments)[(long)blockIdx.y * a.w.n_samples + i] = out;
}

// Stage 2, grid = (group tiles, origins of the longest segment, segments): the updated keys of one
// sampled group of one origin (same arithmetic as dyn_step_kernel: bit-identical keys).  Molecules
// past the end of the array (last group only) get key 0, which can only widen the bracket.
__global__ void __launch_bounds__(128) ov_sample_kernel(const SampleArgs a)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const SdbSegment seg = a.segments[blockIdx.z];
    if (4 * g >= a.w.n_samples || (int)blockIdx.y >= seg.count) return;
    const int o = seg.first + blockIdx.y;
    const SdbOrigin org = a.origins[o];
    if (org.flags & SDB_ORIGIN_NEW) return;
    const long stride = a.p.stride;
    if (a.w.sample_stride > 1) {  // the bracket of this origin is predicted from its history: no sample needed
        float lo, hi;
        if (sdb_predict_bracket(sdb_history(org, stride), org.timediff, 0, rsqrtf((float)a.p.n), lo, hi)) return;
    }
    const long j = (long)g * 4 * a.w.sample_stride;
    const float4 vx = sdb_ldg4(org.d_delta + j), vy = sdb_ldg4(org.d_delta + stride + j), vt = sdb_ldg4(org.d_delta + 2 * stride + j);
    float dx[4] = {vx.x, vx.y, vx.z, vx.w}, dy[4] = {vy.x, vy.y, vy.z, vy.w}, dt[4] = {vt.x, vt.y, vt.z, vt.w};
    const float4* incs = reinterpret_cast<const float4*>(a.w.increments) + (long)blockIdx.z * a.w.n_samples + 4 * g;
    float d2[4], th[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float4 inc = __ldg(incs + e);
        const double alpha = __hiloint2double(__float_as_int(inc.w), __float_as_int(inc.z));
        sdb_apply_increment(dx[e], dy[e], dt[e], inc.x, inc.y, alpha, d2[e], th[e]);
        if (j + e >= a.p.n) d2[e] = th[e] = 0.f;
    }
    float* out = a.w.samples + (long)o * 2 * a.w.n_samples + 4 * g;
    *reinterpret_cast<float4*>(out) = make_float4(d2[0], d2[1], d2[2], d2[3]);
    *reinterpret_cast<float4*>(out + a.w.n_samples) = make_float4(th[0], th[1], th[2], th[3]);
}

// =================================================================================================
// block-wide helpers
// =================================================================================================
// Given hist[0..nbins) (shared, nbins <= 2 * blockDim) find the bin b with
//   count(bins > b) < remaining <= count(bins >= b)
// Returns (in every thread) b, the count above b and hist[b].  `scratch` holds blockDim words.
__device__ void block_pick(const uint32_t* hist, int nbins, uint32_t remaining, uint32_t* scratch, uint32_t& bin,
                           uint32_t& above, uint32_t& at_bin)
{
    __shared__ uint32_t res[3];
    const int per = (nbins + blockDim.x - 1) / blockDim.x;
    const int first = threadIdx.x * per;
    uint32_t mine = 0;
    for (int b = first; b < first + per && b < nbins; ++b) mine += hist[b];
    // inclusive suffix sum over the threads: shuffles inside a warp, one shared-memory hop across warps
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    uint32_t suffix = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t up = __shfl_down_sync(0xffffffffu, suffix, off);
        if (lane + off < 32) suffix += up;
    }
    if (lane == 0) scratch[warp] = suffix;  // total of this warp
    __syncthreads();
    uint32_t later = 0;  // total of the warps after this one
    for (int w = warp + 1; w < n_warps; ++w) later += scratch[w];
    const uint32_t higher0 = suffix - mine + later;
    uint32_t higher = higher0;
    if (higher < remaining && higher + mine >= remaining) {
        for (int b = min(first + per, nbins) - 1; b >= first; --b) {
            const uint32_t h = hist[b];
            if (higher + h >= remaining) {
                res[0] = (uint32_t)b;
                res[1] = higher;
                res[2] = h;
                break;
            }
            higher += h;
        }
    }
    __syncthreads();
    bin = res[0];
    above = res[1];
    at_bin = res[2];
    __syncthreads();
}

// =================================================================================================
// 2. bracket: two order statistics of the sample, one block per (array, origin)
// =================================================================================================
struct BracketArgs {
    const SdbOrigin* origins;
    long stride;
    SdbOvWorkspace w;
    uint32_t rank_hi, rank_lo;  // ranks counted from the largest sample value (1 = maximum)
    int no_sample;              // the sample kernels were skipped: an origin without a predicted bracket gets an
                                // open one (every molecule a candidate -> overflow -> exact kernel)
};

__global__ void __launch_bounds__(512) ov_bracket_kernel(const BracketArgs a)
{
    __shared__ uint32_t hist[2][OV_BINS];
    __shared__ uint32_t scratch[512];
    const int arr = blockIdx.x, o = blockIdx.y;
    const SdbOrigin org = a.origins[o];
    // clear this frame's candidate counter, select record and the fail list (consumed by the resolve step)
    if (arr == 0 && threadIdx.x < 8) {
        a.w.select[8L * o + threadIdx.x] = 0u;
        if (threadIdx.x < 2) a.w.cand_cnt[2 * o + threadIdx.x] = 0u;
        if (o == 0 && threadIdx.x == 2) a.w.fail[0] = 0u;
    }
    if (org.flags & SDB_ORIGIN_NEW) return;
    if (a.w.sample_stride > 1) {
        // both arrays are predicted or neither (the sample kernel tests array 0 only)
        float lo, hi, lo0, hi0;
        const uint32_t* hist = sdb_history(org, a.stride);
        const float inv_sqrt_n = rsqrtf((float)(a.w.n_parts * SDB_WARP_MOLS));
        if (sdb_predict_bracket(hist, org.timediff, 0, inv_sqrt_n, lo0, hi0)) {
            if (arr == 0 || !sdb_predict_bracket(hist, org.timediff, 1, inv_sqrt_n, lo, hi)) {
                if (arr == 0) { lo = lo0; hi = hi0; }
                else { lo = 0.0f; hi = __uint_as_float(0x7f7fffffu); }  // open bracket: overflow -> exact path
            }
            if (threadIdx.x == 0) {
                a.w.brackets[o * 4 + 2 * arr + 0] = lo;
                a.w.brackets[o * 4 + 2 * arr + 1] = hi;
            }
            return;
        }
    }
    if (a.no_sample) {
        if (threadIdx.x == 0) {
            a.w.brackets[o * 4 + 2 * arr + 0] = 0.0f;
            a.w.brackets[o * 4 + 2 * arr + 1] = __uint_as_float(0x7f7fffffu);
        }
        return;
    }
    const uint32_t* keys = reinterpret_cast<const uint32_t*>(a.w.samples + ((long)o * 2 + arr) * a.w.n_samples);
    const int ns = a.w.n_samples;
    uint32_t prefix[2] = {0u, 0u};
    uint32_t remaining[2] = {a.rank_hi, a.rank_lo};
    const bool active[2] = {a.rank_hi >= 1u && a.rank_hi <= (uint32_t)ns, a.rank_lo >= 1u && a.rank_lo <= (uint32_t)ns};
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 20 : (pass == 1 ? 9 : 0);
        const int upper = pass == 0 ? 31 : (pass == 1 ? 20 : 9);
        const uint32_t mask = pass == 2 ? 511u : 2047u;
        for (int b = threadIdx.x; b < 2 * OV_BINS; b += blockDim.x) (&hist[0][0])[b] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < ns; i += blockDim.x) {
            const uint32_t key = keys[i];
#pragma unroll
            for (int s = 0; s < 2; ++s)
                if (active[s] && (upper >= 31 || (key >> upper) == (prefix[s] >> upper)))
                    atomicAdd(&hist[s][(key >> shift) & mask], 1u);
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (!active[s]) continue;
            uint32_t bin, above, at_bin;
            block_pick(hist[s], OV_BINS, remaining[s], scratch, bin, above, at_bin);
            prefix[s] |= bin << shift;
            remaining[s] -= above;
        }
    }
    if (threadIdx.x == 0) {
        // inactive selections: the bracket is open on that side
        const uint32_t hi = active[0] ? prefix[0] : 0x7f800000u;
        const uint32_t lo = active[1] ? prefix[1] : 0u;
        a.w.brackets[o * 4 + 2 * arr + 0] = __uint_as_float(lo);
        a.w.brackets[o * 4 + 2 * arr + 1] = __uint_as_float(hi);
    }
}

// =================================================================================================
// 4. resolve: exact threshold among the candidates of one origin
// =================================================================================================
struct ResolveArgs {
    const SdbOrigin* origins;
    SdbOvWorkspace w;
    double* sums;
    uint32_t k;
    long stride;
};

// Record this frame's exact thresholds in the origin's history (one thread).  `cause`: 0 = the bracket
// held, 1 = it was too wide (candidate overflow), 2 = it missed the threshold (-> widen the next ones).
__device__ void ov_record_history(const SdbOrigin& org, long stride, uint32_t cause, uint32_t thr_r, uint32_t thr_t)
{
    uint32_t* h = sdb_history(org, stride);
    if (h == nullptr) return;
    if (org.flags & SDB_ORIGIN_NEW) {
        h[SDB_H_VALID] = 0u;
        h[SDB_H_WIDEN] = 0u;
        return;
    }
    if (cause & 2u) h[SDB_H_WIDEN] = min(h[SDB_H_WIDEN] + 1u, 6u);
    else if (cause == 1u && h[SDB_H_WIDEN] > 0u) h[SDB_H_WIDEN] -= 1u;
    const uint32_t valid = h[SDB_H_VALID];
    if (valid >= 1u && org.timediff > h[SDB_H_TD1]) {
        h[SDB_H_R0] = h[SDB_H_R1];
        h[SDB_H_T0] = h[SDB_H_T1];
        h[SDB_H_TD0] = h[SDB_H_TD1];
        h[SDB_H_VALID] = 2u;
    } else if (valid == 0u || org.timediff < h[SDB_H_TD1]) {
        h[SDB_H_VALID] = 1u;  // first record, or time moved backwards: start afresh
    }  // same time difference again (zero step): overwrite the last record
    h[SDB_H_R1] = thr_r;
    h[SDB_H_T1] = thr_t;
    h[SDB_H_TD1] = org.timediff;
}

// Strided walk over an array with 8 independent loads in flight per thread (the walks are latency
// bound otherwise: ~80k candidates per origin, 512 threads).
template <class F>
__device__ __forceinline__ void walk8(const uint32_t* __restrict__ v, uint32_t n, F f)
{
    const uint32_t step = blockDim.x;
    uint32_t i = threadIdx.x;
    for (; i + 7 * step < n; i += 8 * step) {
        uint32_t x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = __ldg(v + i + u * step);
#pragma unroll
        for (int u = 0; u < 8; ++u) f(x[u], i + u * step);
    }
    for (; i < n; i += step) f(__ldg(v + i), i);
}

// select record per origin (8 words): thr_r, cut_r, thr_t, cut_t, -, -, fail_r, fail_t
enum { RS_THR = 0, RS_CUT = 1, RS_FAIL = 6 };

// One block per (array, origin): the need-th largest candidate key inside the bracket (radix select
// on the offset from lo, 11 bits per walk over the dense candidate array) and, when several keys
// equal it, the index cut that keeps the `rem` largest molecule indices.
__global__ void __launch_bounds__(512) ov_select_kernel(const ResolveArgs a)
{
    __shared__ uint32_t hist[OV_BINS];
    __shared__ uint32_t scratch[512];
    __shared__ uint32_t tot;
    const int x = blockIdx.x, o = blockIdx.y;
    uint32_t* sel = a.w.select + 8L * o;
    if (a.origins[o].flags & SDB_ORIGIN_NEW) return;
    const double* sums = a.sums + (long)o * SDB_NSUM;
    const uint32_t plo = __float_as_uint(a.w.brackets[o * 4 + 2 * x]), phi = __float_as_uint(a.w.brackets[o * 4 + 2 * x + 1]);
    const long need = (long)a.k - (long)sums[x == 0 ? SDB_S_OV_RABOVE : SDB_S_OV_TABOVE];
    const uint32_t n_cand = a.w.cand_cnt[2 * o];
    const bool overflow = a.w.cand_cnt[2 * o + 1] != 0 || n_cand > (uint32_t)a.w.cand_cap;
    const uint32_t* keys = (x == 0 ? a.w.cand_r : a.w.cand_t) + (long)o * a.w.cand_cap;
    const uint32_t* idx = a.w.cand_idx + (long)o * a.w.cand_cap;

    // walk 0: candidates inside this bracket
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    uint32_t inside = 0;
    if (!overflow) walk8(keys, n_cand, [&](uint32_t kv, uint32_t) { inside += kv >= plo && kv <= phi; });
    inside = __reduce_add_sync(0xffffffffu, inside);
    if ((threadIdx.x & 31) == 0 && inside) atomicAdd(&tot, inside);
    __syncthreads();
    const uint32_t total = tot;
    if (overflow || need < 1 || need > (long)total) {
        if (threadIdx.x == 0) sel[RS_FAIL + x] = overflow ? 1u : 2u;
        return;
    }

    const int bits = phi > plo ? 32 - __clz(phi - plo) : 0;
    uint32_t value = 0, rem = (uint32_t)need, eq = total;
    for (int done = 0; done < bits;) {
        const int nb = min(11, bits - done), shift = bits - done - nb;
        for (int b = threadIdx.x; b < OV_BINS; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        walk8(keys, n_cand, [&](uint32_t key, uint32_t) {
            if (key < plo || key > phi) return;
            const uint32_t kv = key - plo;
            if (done > 0 && (kv >> (shift + nb)) != (value >> (shift + nb))) return;
            atomicAdd(&hist[(kv >> shift) & ((1u << nb) - 1u)], 1u);
        });
        __syncthreads();
        uint32_t bin, above, at_bin;
        block_pick(hist, 1 << nb, rem, scratch, bin, above, at_bin);
        value |= bin << shift;
        rem -= above;
        eq = at_bin;
        done += nb;
    }
    const uint32_t thr = value + plo;

    uint32_t cut = 0;
    if (eq > rem) {  // keep the rem largest molecule indices among the keys equal to the threshold
        const int idx_bits = 32 - __clz(max(a.w.n_parts * SDB_WARP_MOLS - 1, 1));
        uint32_t r = rem;
        for (int done = 0; done < idx_bits;) {
            const int nb = min(11, idx_bits - done), shift = idx_bits - done - nb;
            for (int b = threadIdx.x; b < OV_BINS; b += blockDim.x) hist[b] = 0;
            __syncthreads();
            walk8(keys, n_cand, [&](uint32_t key, uint32_t i) {
                if (key != thr) return;
                const uint32_t id = __ldg(idx + i);
                if (done > 0 && (id >> (shift + nb)) != (cut >> (shift + nb))) return;
                atomicAdd(&hist[(id >> shift) & ((1u << nb) - 1u)], 1u);
            });
            __syncthreads();
            uint32_t bin, above, at_bin;
            block_pick(hist, 1 << nb, r, scratch, bin, above, at_bin);
            cut |= bin << shift;
            r -= above;
            done += nb;
        }
    }
    if (threadIdx.x == 0) {
        sel[2 * x + RS_THR] = thr;
        sel[2 * x + RS_CUT] = cut;
    }
}

// grid = (chunks, origins): count the candidates inside both top-k sets; add the molecules above
// both brackets (counted by the fused kernel).  Failed origins are queued for the exact kernel.
__global__ void __launch_bounds__(256) ov_intersect_kernel(const ResolveArgs a)
{
    __shared__ uint32_t warp_counts[8];
    const int o = blockIdx.y;
    double* sums = a.sums + (long)o * SDB_NSUM;
    const uint32_t* sel = a.w.select + 8L * o;
    const SdbOrigin org = a.origins[o];
    if (org.flags & SDB_ORIGIN_NEW) {
        // all keys are zero: ties are broken by index for both arrays -> identical sets
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            sums[SDB_S_OVERLAP] = (double)a.k;
            ov_record_history(org, a.stride, 0u, 0u, 0u);
        }
        return;
    }
    if (sel[RS_FAIL] | sel[RS_FAIL + 1]) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const uint32_t slot = atomicAdd(a.w.fail, 1u);
            a.w.fail[1 + slot] = (uint32_t)o;  // the exact kernel records the thresholds
        }
        return;
    }
    const uint32_t thr_r = sel[RS_THR], cut_r = sel[RS_CUT], thr_t = sel[2 + RS_THR], cut_t = sel[2 + RS_CUT];
    if (blockIdx.x == 0 && threadIdx.x == 0) ov_record_history(org, a.stride, 0u, thr_r, thr_t);
    const uint32_t n_cand = a.w.cand_cnt[2 * o];
    const long row = (long)o * a.w.cand_cap;
    uint32_t both = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += gridDim.x * blockDim.x) {
        const uint32_t id = __ldg(a.w.cand_idx + row + i), r = __ldg(a.w.cand_r + row + i), t = __ldg(a.w.cand_t + row + i);
        const bool in_r = r > thr_r || (r == thr_r && id >= cut_r);
        const bool in_t = t > thr_t || (t == thr_t && id >= cut_t);
        both += in_r && in_t;
    }
    both = __reduce_add_sync(0xffffffffu, both);
    if ((threadIdx.x & 31) == 0) warp_counts[threadIdx.x >> 5] = both;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int w = 0; w < 8; ++w) total += warp_counts[w];
        double add = (double)total;
        if (blockIdx.x == 0) add += sums[SDB_S_OV_BOTH];
        // integer-valued doubles: exact and independent of the arrival order
        if (add != 0.0) atomicAdd(sums + SDB_S_OVERLAP, add);
    }
}

// =================================================================================================
// 4b. fused resolve (sdb_frame_step): row sums + both selections + intersection of one origin in ONE
//     block of 1024 threads.  The chain finalize -> select -> intersect of the kernels above is bound by
//     launch gaps and dependent global-memory round trips (~130 us whatever the number of origins);
//     here the two selections run side by side on the two halves of the block (named barriers) and the
//     intersection follows in place.  The candidates (~10 k per origin with predicted brackets) are walked
//     in L2: staging them in 186 KB of shared memory per block was measured and dropped -- it did not
//     shorten the kernel and slowed the struct kernel running beside it by 10-30 %.
// =================================================================================================
constexpr int RES_THREADS = 1024;
constexpr int RES_HALF = 512;

struct FusedArgs {
    const SdbOrigin* origins;
    SdbOvWorkspace w;
    const double* partials;  // [origins][n_parts][SDB_NSUM]
    double* sums;
    int n_parts;
    uint32_t k;
    long stride;
};

__device__ __forceinline__ void half_sync(int h)
{
    if (h) asm volatile("bar.sync 2, 512;" ::: "memory");
    else asm volatile("bar.sync 1, 512;" ::: "memory");
}

// block_pick for one half of the block (512 threads, named barrier 1 + h); `t` = thread index in the half.
__device__ __forceinline__ void half_pick(const uint32_t* hist, int nbins, uint32_t remaining, uint32_t* scratch,
                                          uint32_t* res, int h, int t, uint32_t& bin, uint32_t& above, uint32_t& at_bin)
{
    const int per = (nbins + RES_HALF - 1) / RES_HALF;
    const int first = t * per;
    uint32_t mine = 0;
    for (int b = first; b < first + per && b < nbins; ++b) mine += hist[b];
    const int lane = t & 31, warp = t >> 5;
    uint32_t suffix = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t up = __shfl_down_sync(0xffffffffu, suffix, off);
        if (lane + off < 32) suffix += up;
    }
    if (lane == 0) scratch[warp] = suffix;
    half_sync(h);
    uint32_t later = 0;
    for (int w = warp + 1; w < RES_HALF / 32; ++w) later += scratch[w];
    uint32_t higher = suffix - mine + later;
    if (higher < remaining && higher + mine >= remaining) {
        for (int b = min(first + per, nbins) - 1; b >= first; --b) {
            const uint32_t hv = hist[b];
            if (higher + hv >= remaining) {
                res[0] = (uint32_t)b;
                res[1] = higher;
                res[2] = hv;
                break;
            }
            higher += hv;
        }
    }
    half_sync(h);
    bin = res[0];
    above = res[1];
    at_bin = res[2];
    half_sync(h);
}

// Strided walk of one half of the block, 8 independent loads in flight per thread.
template <class F>
__device__ __forceinline__ void half_walk8(const uint32_t* v, uint32_t n, uint32_t t, F f)
{
    uint32_t i = t;
    for (; i + 7 * RES_HALF < n; i += 8 * RES_HALF) {
        uint32_t x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = v[i + u * RES_HALF];
#pragma unroll
        for (int u = 0; u < 8; ++u) f(x[u], i + u * RES_HALF);
    }
    for (; i < n; i += RES_HALF) f(v[i], i);
}

// (launch bounds: 32 registers per thread, so that the struct kernel's CTAs fit next to this block)
__global__ void __launch_bounds__(RES_THREADS, 2) ov_resolve_fused_kernel(const FusedArgs a)
{
    __shared__ uint32_t hist[2][OV_BINS];
    __shared__ double fin[64][17];
    __shared__ double osum[SDB_NSUM];
    __shared__ uint32_t scratch[2][RES_HALF / 32], res[2][3], tot[2], out_thr[2], out_cut[2], out_fail[2], both_total;
    const int o = blockIdx.x, tid = threadIdx.x;

    // ---- row sums of this origin (dyn_finalize_kernel) -------------------------------------------
    {
        const int v = tid & 15, r = tid >> 4;
        const double* base = a.partials + (long)o * a.n_parts * SDB_NSUM + v;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int p = r;
        for (; p + 192 < a.n_parts; p += 256) {
            s0 += base[(long)p * SDB_NSUM];
            s1 += base[(long)(p + 64) * SDB_NSUM];
            s2 += base[(long)(p + 128) * SDB_NSUM];
            s3 += base[(long)(p + 192) * SDB_NSUM];
        }
        for (; p < a.n_parts; p += 64) s0 += base[(long)p * SDB_NSUM];
        fin[r][v] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (tid < 16) {
            double t = 0.0;
#pragma unroll 8
            for (int i = 0; i < 64; ++i) t += fin[i][tid];
            osum[tid] = t;
            if (tid != SDB_S_STRUCT && tid != SDB_S_OVERLAP) {
                const bool produced = tid < SDB_NFSUM || tid == SDB_S_COMSTRUCT || tid == SDB_S_BADQUAT || tid >= SDB_S_OV_RABOVE;
                a.sums[(long)o * SDB_NSUM + tid] = produced ? t : 0.0;
            }
        }
        if (tid < 2) {
            tot[tid] = 0;
            out_fail[tid] = 0;
            out_thr[tid] = 0;
            out_cut[tid] = 0;
        }
        if (tid == 0) both_total = 0;
    }
    const SdbOrigin org = a.origins[o];
    double* sums = a.sums + (long)o * SDB_NSUM;
    uint32_t* sel = a.w.select + 8L * o;
    if (org.flags & SDB_ORIGIN_NEW) {
        // all keys are zero: ties are broken by index for both arrays -> identical sets
        if (tid == 0) {
            sums[SDB_S_OVERLAP] = (double)a.k;
            ov_record_history(org, a.stride, 0u, 0u, 0u);
        }
        return;
    }
    const uint32_t n_cand = a.w.cand_cnt[2 * o];
    const bool overflow = a.w.cand_cnt[2 * o + 1] != 0 || n_cand > (uint32_t)a.w.cand_cap;
    const long row = (long)o * a.w.cand_cap;
    __syncthreads();
    const uint32_t* idx = a.w.cand_idx + row;
    const uint32_t* keys_r = a.w.cand_r + row;
    const uint32_t* keys_t = a.w.cand_t + row;

    // ---- the two selections, one per half of the block ---------------------------------------------
    {
        const int h = tid >> 9, t = tid & (RES_HALF - 1);
        const uint32_t* keys = h ? keys_t : keys_r;
        uint32_t* hh = hist[h];
        const uint32_t plo = __float_as_uint(a.w.brackets[o * 4 + 2 * h]), phi = __float_as_uint(a.w.brackets[o * 4 + 2 * h + 1]);
        const long need = (long)a.k - (long)osum[h ? SDB_S_OV_TABOVE : SDB_S_OV_RABOVE];
        uint32_t inside = 0;
        if (!overflow) half_walk8(keys, n_cand, t, [&](uint32_t kv, uint32_t) { inside += kv >= plo && kv <= phi; });
        inside = __reduce_add_sync(0xffffffffu, inside);
        if ((t & 31) == 0 && inside) atomicAdd(&tot[h], inside);
        half_sync(h);
        const uint32_t total = tot[h];
        if (overflow || need < 1 || need > (long)total) {
            if (t == 0) out_fail[h] = overflow ? 1u : 2u;
        } else {
            const int bits = phi > plo ? 32 - __clz(phi - plo) : 0;
            uint32_t value = 0, rem = (uint32_t)need, eq = total;
            for (int done = 0; done < bits;) {
                const int nb = min(11, bits - done), shift = bits - done - nb;
                for (int b = t; b < OV_BINS; b += RES_HALF) hh[b] = 0;
                half_sync(h);
                half_walk8(keys, n_cand, t, [&](uint32_t key, uint32_t) {
                    if (key < plo || key > phi) return;
                    const uint32_t kv = key - plo;
                    if (done > 0 && (kv >> (shift + nb)) != (value >> (shift + nb))) return;
                    atomicAdd(&hh[(kv >> shift) & ((1u << nb) - 1u)], 1u);
                });
                half_sync(h);
                uint32_t bin, above, at_bin;
                half_pick(hh, 1 << nb, rem, scratch[h], res[h], h, t, bin, above, at_bin);
                value |= bin << shift;
                rem -= above;
                eq = at_bin;
                done += nb;
            }
            const uint32_t thr = value + plo;
            uint32_t cut = 0;
            if (eq > rem) {  // keep the rem largest molecule indices among the keys equal to the threshold
                const int idx_bits = 32 - __clz(max(a.w.n_parts * SDB_WARP_MOLS - 1, 1));
                uint32_t r = rem;
                for (int done = 0; done < idx_bits;) {
                    const int nb = min(11, idx_bits - done), shift = idx_bits - done - nb;
                    for (int b = t; b < OV_BINS; b += RES_HALF) hh[b] = 0;
                    half_sync(h);
                    half_walk8(keys, n_cand, t, [&](uint32_t key, uint32_t i) {
                        if (key != thr) return;
                        const uint32_t id = idx[i];
                        if (done > 0 && (id >> (shift + nb)) != (cut >> (shift + nb))) return;
                        atomicAdd(&hh[(id >> shift) & ((1u << nb) - 1u)], 1u);
                    });
                    half_sync(h);
                    uint32_t bin, above, at_bin;
                    half_pick(hh, 1 << nb, r, scratch[h], res[h], h, t, bin, above, at_bin);
                    cut |= bin << shift;
                    r -= above;
                    done += nb;
                }
            }
            if (t == 0) {
                out_thr[h] = thr;
                out_cut[h] = cut;
            }
        }
    }
    __syncthreads();
    if (tid < 2) {
        sel[2 * tid + RS_THR] = out_thr[tid];
        sel[2 * tid + RS_CUT] = out_cut[tid];
        sel[RS_FAIL + tid] = out_fail[tid];
    }
    if (out_fail[0] | out_fail[1]) {
        if (tid == 0) {
            const uint32_t slot = atomicAdd(a.w.fail, 1u);
            a.w.fail[1 + slot] = (uint32_t)o;  // the exact kernel records the thresholds
        }
        return;
    }

    // ---- intersection ---------------------------------------------------------------------------
    const uint32_t thr_r = out_thr[0], cut_r = out_cut[0], thr_t = out_thr[1], cut_t = out_cut[1];
    uint32_t both = 0;
    for (uint32_t i = tid; i < n_cand; i += RES_THREADS) {
        const uint32_t id = idx[i], r = keys_r[i], t = keys_t[i];
        const bool in_r = r > thr_r || (r == thr_r && id >= cut_r);
        const bool in_t = t > thr_t || (t == thr_t && id >= cut_t);
        both += in_r && in_t;
    }
    both = __reduce_add_sync(0xffffffffu, both);
    if ((tid & 31) == 0 && both) atomicAdd(&both_total, both);
    __syncthreads();
    if (tid == 0) {
        sums[SDB_S_OVERLAP] = (double)both_total + osum[SDB_S_OV_BOTH];
        ov_record_history(org, a.stride, 0u, thr_r, thr_t);
    }
}

// =================================================================================================
// 5. exact selection over the whole array: one block per origin of the list
// =================================================================================================
struct ExactArgs {
    const SdbOrigin* origins;
    const uint32_t* list;  // [0] = count, [1..] = origin indices
    const uint32_t* select;  // select records of the bracketed path (failure causes) or nullptr
    double* sums;
    int n;
    long stride;
    uint32_t k;
    int key_mode;  // 0: keys from the (dx, dy, dtheta) planes; 1: plane 0 = displacement, plane 2 = rotation
};

__device__ __forceinline__ void ex_keys4(const float* d, long stride, int j, uint32_t kr[4], uint32_t kt[4], int key_mode)
{
    const float4 x = sdb_ldg4(d + j), y = sdb_ldg4(d + stride + j), t = sdb_ldg4(d + 2 * stride + j);
    const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w}, ts[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float d2 = __fadd_rn(__fmul_rn(xs[e], xs[e]), __fmul_rn(ys[e], ys[e]));
        kr[e] = key_mode ? __float_as_uint(xs[e]) : __float_as_uint(d2);
        kt[e] = __float_as_uint(fabsf(ts[e]));
    }
}

// Visit (key_r, key_t, index) of every molecule of one origin with the whole block.
template <class F>
__device__ __forceinline__ void ex_for_each(const ExactArgs& a, const float* d, F f)
{
    for (int j = threadIdx.x * 4; j < a.n; j += blockDim.x * 4) {
        uint32_t kr[4], kt[4];
        ex_keys4(d, a.stride, j, kr, kt, a.key_mode);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (j + e < a.n) f(kr[e], kt[e], (uint32_t)(j + e));
    }
}

// Radix select (most significant digit first, 11 bits per pass) entirely inside one block: the
// k-th largest f32 bit pattern of each array, then -- if several keys equal it -- the index cut that
// keeps the largest molecule indices, then the intersection count.  Rare path: it serves origins
// whose sampled bracket failed (e.g. all keys equal) and mobile_overlap() on caller arrays.
__global__ void __launch_bounds__(1024) ov_exact_kernel(const ExactArgs a)
{
    __shared__ uint32_t hist[2][OV_BINS];
    __shared__ uint32_t scratch[1024];
    __shared__ uint32_t total;
    const uint32_t n_list = a.list[0];
    for (uint32_t li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int o = (int)a.list[1 + li];
        const float* d = a.origins[o].d_delta;
        uint32_t thr[2] = {0u, 0u}, rem[2] = {a.k, a.k}, eq[2] = {0u, 0u};
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = pass == 0 ? 20 : (pass == 1 ? 9 : 0);
            const int upper = pass == 0 ? 31 : (pass == 1 ? 20 : 9);
            const uint32_t mask = pass == 2 ? 511u : 2047u;
            for (int i = threadIdx.x; i < 2 * OV_BINS; i += blockDim.x) (&hist[0][0])[i] = 0;
            __syncthreads();
            ex_for_each(a, d, [&](uint32_t kr, uint32_t kt, uint32_t) {
                if (upper >= 31 || (kr >> upper) == (thr[0] >> upper)) atomicAdd(&hist[0][(kr >> shift) & mask], 1u);
                if (upper >= 31 || (kt >> upper) == (thr[1] >> upper)) atomicAdd(&hist[1][(kt >> shift) & mask], 1u);
            });
            __syncthreads();
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                uint32_t bin, above, at_bin;
                block_pick(hist[x], OV_BINS, rem[x], scratch, bin, above, at_bin);
                thr[x] |= bin << shift;
                rem[x] -= above;
                eq[x] = at_bin;
            }
        }
        // ties across the cut: keep the rem largest molecule indices (radix select on the index)
        uint32_t cut[2] = {0u, 0u};
        const int idx_bits = 32 - __clz(max(a.n - 1, 1));
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            if (eq[x] <= rem[x]) continue;  // block-uniform
            uint32_t r = rem[x];
            for (int done = 0; done < idx_bits;) {
                const int nb = min(11, idx_bits - done), shift = idx_bits - done - nb;
                for (int i = threadIdx.x; i < OV_BINS; i += blockDim.x) hist[0][i] = 0;
                __syncthreads();
                ex_for_each(a, d, [&](uint32_t kr, uint32_t kt, uint32_t idx) {
                    if ((x == 0 ? kr : kt) != thr[x]) return;
                    if (done > 0 && (idx >> (shift + nb)) != (cut[x] >> (shift + nb))) return;
                    atomicAdd(&hist[0][(idx >> shift) & ((1u << nb) - 1u)], 1u);
                });
                __syncthreads();
                uint32_t bin, above, at_bin;
                block_pick(hist[0], 1 << nb, r, scratch, bin, above, at_bin);
                cut[x] |= bin << shift;
                r -= above;
                done += nb;
            }
        }
        // intersection
        if (threadIdx.x == 0) total = 0;
        __syncthreads();
        uint32_t both = 0;
        ex_for_each(a, d, [&](uint32_t kr, uint32_t kt, uint32_t idx) {
            const bool in_r = kr > thr[0] || (kr == thr[0] && idx >= cut[0]);
            const bool in_t = kt > thr[1] || (kt == thr[1] && idx >= cut[1]);
            both += in_r && in_t;
        });
        both = __reduce_add_sync(0xffffffffu, both);
        if ((threadIdx.x & 31) == 0 && both) atomicAdd(&total, both);
        __syncthreads();
        if (threadIdx.x == 0) {
            a.sums[(long)o * SDB_NSUM + SDB_S_OVERLAP] = (double)total;
            if (a.select && a.key_mode == 0)
                ov_record_history(a.origins[o], a.stride, a.select[8L * o + RS_FAIL] | a.select[8L * o + RS_FAIL + 1],
                                  thr[0], thr[1]);
        }
        __syncthreads();
    }
}

__global__ void ov_fill_list_kernel(uint32_t* list, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) list[0] = (uint32_t)n;
    if (i < n) list[1 + i] = (uint32_t)i;
}

int launch_exact(const SdbOrigin* d_origins, const uint32_t* d_list, const uint32_t* d_select, double* d_sums, int n,
                 int stride, int k, int key_mode, int max_list, cudaStream_t s)
{
    ExactArgs a;
    a.origins = d_origins;
    a.list = d_list;
    a.select = d_select;
    a.sums = d_sums;
    a.n = n;
    a.stride = stride;
    a.k = (uint32_t)k;
    a.key_mode = key_mode;
    const int blocks = max_list < 148 ? (max_list < 1 ? 1 : max_list) : 148;
    ov_exact_kernel<<<blocks, 1024, 0, s>>>(a);
    SDB_CHECK_LAUNCH("ov_exact_kernel");
    return SDB_OK;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
extern "C" size_t sdb_overlap_ws_bytes(int32_t max_origins, int32_t n)
{
    if (max_origins <= 0 || n <= 0) return 0;
    return sdb_ov_layout(nullptr, max_origins, n, nullptr);
}

extern "C" int sdb_overlap_prepare(const SdbDynParams* h_params, const float* d_pos, const float* d_quat,
                                   const SdbOrigin* d_origins, int32_t n_origins, const SdbSegment* d_segments,
                                   int32_t n_segments, int32_t max_segment_count, int32_t k, void* d_ws,
                                   int32_t ws_origins, void* stream)
{
    return sdb_overlap_prepare_ex(h_params, d_pos, d_quat, d_origins, n_origins, d_segments, n_segments, max_segment_count, k,
                                  d_ws, ws_origins, 0, stream);
}

// skip_sample != 0: the caller knows that every origin has a threshold history (two or more consecutive
// frames with increasing time), so the brackets are predicted and the sample kernels are not launched.
int sdb_overlap_prepare_ex(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, const SdbOrigin* d_origins,
                           int32_t n_origins, const SdbSegment* d_segments, int32_t n_segments, int32_t max_segment_count,
                           int32_t k, void* d_ws, int32_t ws_origins, int32_t skip_sample, void* stream)
{
    if (!h_params || !d_origins || !d_segments || !d_ws || n_origins > ws_origins || k <= 0 || k > h_params->n ||
        max_segment_count <= 0 || max_segment_count > 65535 || n_segments > 65535) {
        sdb_set_error("sdb_overlap_prepare: bad argument");
        return SDB_EINVAL;
    }
    if (n_origins <= 0) return SDB_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SampleArgs sa;
    sa.p = *h_params;
    sa.pos = d_pos;
    sa.quat = d_quat;
    sa.origins = d_origins;
    sa.segments = d_segments;
    sa.atan_tab = sdb_device_atan_table();
    if (!sa.atan_tab) return SDB_ECUDA;
    sdb_ov_layout(d_ws, ws_origins, h_params->n, &sa.w);
    if (n_segments > ws_origins) {
        sdb_set_error("sdb_overlap_prepare: more segments than workspace rows");
        return SDB_EINVAL;
    }
    // (the candidate counters, select records and the fail list are cleared by ov_bracket_kernel)
    skip_sample = skip_sample && sa.w.sample_stride > 1;  // tiny systems: the sample is the whole array
    if (!skip_sample) {
        ov_increment_kernel<<<dim3((sa.w.n_samples + 127) / 128, n_segments), 128, 0, s>>>(sa);
        SDB_CHECK_LAUNCH("ov_increment_kernel");
        ov_sample_kernel<<<dim3((sa.w.n_samples / 4 + 127) / 128, max_segment_count, n_segments), 128, 0, s>>>(sa);
        SDB_CHECK_LAUNCH("ov_sample_kernel");
    }

    // ranks (from the top) of the two sample order statistics that bracket the k-th largest key
    BracketArgs ba;
    ba.origins = d_origins;
    ba.stride = h_params->stride;
    ba.w = sa.w;
    const double n = h_params->n, ns = sa.w.n_samples, p = k / n;
    if (sa.w.sample_stride == 1) {
        ba.rank_hi = ba.rank_lo = (uint32_t)k;  // the sample is the whole array: lo = hi = threshold
    } else {
        const double mu = p * ns, sigma = sqrt(ns * p * (1.0 - p));
        const double width = 10.0 * sigma + 4.0;  // 5 sigma with a 2x allowance: the sample consists of
                                                   // groups of 4 neighbours, which may be correlated
        const double hi = floor(mu - width), lo = ceil(mu + width);
        ba.rank_hi = hi < 1.0 ? 0u : (uint32_t)hi;                // 0: open above
        ba.rank_lo = lo > ns ? (uint32_t)ns + 1u : (uint32_t)lo;  // > ns: open below
    }
    ba.no_sample = skip_sample && sa.w.sample_stride > 1;
    ov_bracket_kernel<<<dim3(2, n_origins), 512, 0, s>>>(ba);
    SDB_CHECK_LAUNCH("ov_bracket_kernel");
    return SDB_OK;
}

extern "C" int sdb_overlap_resolve(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride, int32_t k,
                                   void* d_ws, int32_t ws_origins, double* d_sums, void* stream)
{
    if (!d_origins || !d_ws || !d_sums || n <= 0 || k <= 0 || k > n || n_origins > ws_origins) {
        sdb_set_error("sdb_overlap_resolve: bad argument");
        return SDB_EINVAL;
    }
    if (n_origins <= 0) return SDB_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ResolveArgs ra;
    ra.origins = d_origins;
    ra.sums = d_sums;
    ra.k = (uint32_t)k;
    ra.stride = stride;
    sdb_ov_layout(d_ws, ws_origins, n, &ra.w);
    int rc = sdb_check_cuda(cudaMemset2DAsync(d_sums + SDB_S_OVERLAP, SDB_NSUM * sizeof(double), 0, sizeof(double),
                                              (size_t)n_origins, s),
                            "sdb_overlap_resolve: clear");
    if (rc) return rc;
    ov_select_kernel<<<dim3(2, n_origins), 512, 0, s>>>(ra);
    SDB_CHECK_LAUNCH("ov_select_kernel");
    ov_intersect_kernel<<<dim3(8, n_origins), 256, 0, s>>>(ra);
    SDB_CHECK_LAUNCH("ov_intersect_kernel");
    // origins whose bracket failed: exact selection over the whole array (no-op when the list is empty)
    return launch_exact(d_origins, ra.w.fail, ra.w.select, d_sums, n, stride, k, 0, n_origins, s);
}

// sdb_dyn_finalize + sdb_overlap_resolve in one launch per frame (plus the exact kernel for failed brackets):
// the resolve step of sdb_frame_step.  Expects dyn_step_kernel's partial rows (not yet summed).
int sdb_overlap_resolve_fused(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride, int32_t k, void* d_ws,
                              int32_t ws_origins, const double* d_partials, double* d_sums, void* stream)
{
    if (!d_origins || !d_ws || !d_sums || !d_partials || n <= 0 || k <= 0 || k > n || n_origins > ws_origins) {
        sdb_set_error("sdb_overlap_resolve_fused: bad argument");
        return SDB_EINVAL;
    }
    if (n_origins <= 0) return SDB_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    FusedArgs fa;
    fa.origins = d_origins;
    fa.partials = d_partials;
    fa.sums = d_sums;
    fa.k = (uint32_t)k;
    fa.stride = stride;
    sdb_ov_layout(d_ws, ws_origins, n, &fa.w);
    fa.n_parts = fa.w.n_parts;
    ov_resolve_fused_kernel<<<n_origins, RES_THREADS, 0, s>>>(fa);
    SDB_CHECK_LAUNCH("ov_resolve_fused_kernel");
    return launch_exact(d_origins, fa.w.fail, fa.w.select, d_sums, n, stride, k, 0, n_origins, s);
}

// Diagnostics of the last sdb_overlap_resolve on this workspace (synchronises the stream):
// h_out[0] = number of origins that took the exact path, h_out[1 + o] = candidates collected for origin o.
extern "C" int sdb_overlap_stats(void* d_ws, int32_t ws_origins, int32_t n, int32_t n_origins, uint32_t* h_out,
                                 void* stream)
{
    if (!d_ws || !h_out || n_origins > ws_origins || n <= 0) {
        sdb_set_error("sdb_overlap_stats: bad argument");
        return SDB_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SdbOvWorkspace w;
    sdb_ov_layout(d_ws, ws_origins, n, &w);
    int rc = sdb_check_cuda(cudaMemcpyAsync(h_out, w.fail, sizeof(uint32_t), cudaMemcpyDeviceToHost, s), "stats: fail");
    if (rc) return rc;
    rc = sdb_check_cuda(cudaMemcpy2DAsync(h_out + 1, sizeof(uint32_t), w.cand_cnt, 2 * sizeof(uint32_t), sizeof(uint32_t),
                                          (size_t)n_origins, cudaMemcpyDeviceToHost, s),
                        "stats: candidates");
    if (rc) return rc;
    return sdb_check_cuda(cudaStreamSynchronize(s), "stats: sync");
}

extern "C" size_t sdb_overlap_scratch_bytes(int32_t n_origins, int32_t n)
{
    if (n_origins <= 0 || n <= 0) return 0;
    (void)n;
    return sizeof(uint32_t) * (1 + (size_t)n_origins + 64);
}

extern "C" int sdb_overlap(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride, int32_t key_mode,
                           int32_t k, void* d_scratch, double* d_sums, void* stream)
{
    if (!d_origins || !d_scratch || !d_sums || n <= 0 || k <= 0 || k > n) {
        sdb_set_error("sdb_overlap: bad argument (n=%d k=%d)", n, k);
        return SDB_EINVAL;
    }
    if (n_origins <= 0) return SDB_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint32_t* list = static_cast<uint32_t*>(d_scratch);
    ov_fill_list_kernel<<<(n_origins + 255) / 256, 256, 0, s>>>(list, n_origins);
    SDB_CHECK_LAUNCH("ov_fill_list_kernel");
    return launch_exact(d_origins, list, nullptr, d_sums, n, stride, k, key_mode, n_origins, s);
}
