// overlap.cu -- exact `overlap` of Dynamics.compute_all (reference dynamics/dynamics.py:380-382,431-445):
//   k = int(N * 0.1); |argsort(|dr|)[-k:]  intersect  argsort(|dtheta|)[-k:]| / k
//
// Exact selection of the k largest keys of two f32 arrays per origin without sorting: a three-digit
// (11 + 11 + 9 bit) most-significant-digit radix select on the f32 bit patterns of the non-negative
// keys finds the k-th largest value t; members of the top-k set are the keys > t plus, of the keys
// == t, the ones with the largest molecule indices (numpy's introsort leaves the choice among equal
// keys unspecified -- parity trap T4 -- this is our deterministic rule; for all-equal arrays it
// gives identical sets for both arrays, hence overlap = 1 like the reference).
// Keys are exactly the reference's: |dr| = fl32(sqrt(fl32(dx^2 + dy^2))), |dtheta| = |dth|.
#include "sdb_internal.cuh"

namespace {

constexpr int OV_BINS = 2048;
constexpr int OV_SEL = 8;  // words per (origin, array) selection record
enum { SEL_PREFIX = 0, SEL_REMAINING = 1, SEL_COUNT_EQ = 2, SEL_IDXCUT = 3, SEL_NEED_TIE = 4 };
constexpr int OV_THREADS = 256;
constexpr int OV_ITERS = 4;
constexpr int OV_BLOCK_MOLS = OV_THREADS * 4 * OV_ITERS;  // 4096

struct OvArgs {
    const SdbOrigin* origins;
    uint32_t* scratch;
    double* sums;
    int n;
    long stride;
    int n_parts;  // 256-molecule parts (tie counting)
    long words_per_origin;
    uint32_t k;
    int key_mode;  // 0: keys from the (dx, dy, dtheta) planes; 1: plane 0 = |dr|, plane 2 = rotation, given
};

__device__ __forceinline__ uint32_t* ov_hist(const OvArgs& a, int o) { return a.scratch + o * a.words_per_origin; }
__device__ __forceinline__ uint32_t* ov_sel(const OvArgs& a, int o) { return ov_hist(a, o) + 2 * OV_BINS; }
__device__ __forceinline__ uint32_t* ov_tie(const OvArgs& a, int o) { return ov_sel(a, o) + 2 * OV_SEL; }

__device__ __forceinline__ void ov_keys4(const float* d, long stride, int j, uint32_t kr[4], uint32_t kt[4],
                                         int key_mode = 0)
{
    const float4 x = sdb_ldg4(d + j), y = sdb_ldg4(d + stride + j), t = sdb_ldg4(d + 2 * stride + j);
    const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w}, ts[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float d2 = __fadd_rn(__fmul_rn(xs[e], xs[e]), __fmul_rn(ys[e], ys[e]));
        kr[e] = key_mode ? __float_as_uint(xs[e]) : __float_as_uint(__fsqrt_rn(d2));
        kt[e] = __float_as_uint(fabsf(ts[e]));
    }
}

__device__ __forceinline__ int ov_shift(int pass) { return pass == 0 ? 20 : (pass == 1 ? 9 : 0); }

__global__ void ov_init_kernel(const OvArgs a)
{
    const int o = blockIdx.x;
    uint32_t* h = ov_hist(a, o);
    for (int i = threadIdx.x; i < 2 * OV_BINS + 2 * OV_SEL; i += blockDim.x) h[i] = 0;
    __syncthreads();
    if (threadIdx.x < 2) ov_sel(a, o)[threadIdx.x * OV_SEL + SEL_REMAINING] = a.k;
}

__global__ void __launch_bounds__(OV_THREADS) ov_hist_kernel(const OvArgs a, int pass)
{
    __shared__ uint32_t sh[2][OV_BINS];
    const int o = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * OV_BINS; i += OV_THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();
    const float* d = a.origins[o].d_delta;
    const uint32_t* sel = ov_sel(a, o);
    const uint32_t prefix[2] = {sel[SEL_PREFIX], sel[OV_SEL + SEL_PREFIX]};
    const int shift = ov_shift(pass);
    const int upper = pass == 0 ? 31 : (pass == 1 ? 20 : 9);  // bits above the current digit
    const uint32_t mask = pass == 2 ? 511u : 2047u;
#pragma unroll
    for (int it = 0; it < OV_ITERS; ++it) {
        const int j = blockIdx.x * OV_BLOCK_MOLS + (it * OV_THREADS + threadIdx.x) * 4;
        if (j >= a.n) break;
        uint32_t kr[4], kt[4];
        ov_keys4(d, a.stride, j, kr, kt, a.key_mode);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (j + e >= a.n) break;
            const uint32_t keys[2] = {kr[e], kt[e]};
#pragma unroll
            for (int arr = 0; arr < 2; ++arr) {
                const uint32_t key = keys[arr];
                const bool match = upper >= 31 || (key >> upper) == (prefix[arr] >> upper);
                if (match) atomicAdd(&sh[arr][(key >> shift) & mask], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t* h = ov_hist(a, o);
    for (int i = threadIdx.x; i < 2 * OV_BINS; i += OV_THREADS) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(h + i, c);
    }
}

// One block per (array, origin): locate the digit of the k-th largest key, clear the histogram.
__global__ void __launch_bounds__(256) ov_pick_kernel(const OvArgs a, int pass)
{
    __shared__ uint32_t totals[256];
    const int arr = blockIdx.x, o = blockIdx.y;
    uint32_t* h = ov_hist(a, o) + arr * OV_BINS;
    uint32_t* sel = ov_sel(a, o) + arr * OV_SEL;
    uint32_t local[8];
    uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        local[i] = h[threadIdx.x * 8 + i];
        mine += local[i];
    }
    totals[threadIdx.x] = mine;
    __syncthreads();
    // exclusive suffix sum over threads (count held by higher bins)
    for (int off = 1; off < 256; off <<= 1) {
        const uint32_t add = threadIdx.x + off < 256 ? totals[threadIdx.x + off] : 0;
        __syncthreads();
        totals[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t above = totals[threadIdx.x] - mine;  // keys in bins of higher threads
    const uint32_t remaining = sel[SEL_REMAINING];
    __syncthreads();
    if (above < remaining && above + mine >= remaining) {
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            if (above < remaining && above + local[i] >= remaining) {
                const uint32_t bin = threadIdx.x * 8 + i;
                sel[SEL_PREFIX] |= bin << ov_shift(pass);
                sel[SEL_REMAINING] = remaining - above;
                sel[SEL_COUNT_EQ] = local[i];
                if (pass == 2) {
                    sel[SEL_NEED_TIE] = local[i] > remaining - above;
                    sel[SEL_IDXCUT] = 0;
                }
                above = 0xffffffffu;  // done
                break;
            }
            above += local[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) h[threadIdx.x * 8 + i] = 0;
}

// Ties across the cut: count, per 256-molecule part, the keys equal to the threshold.
__global__ void __launch_bounds__(256) ov_tiecount_kernel(const OvArgs a)
{
    const int o = blockIdx.y;
    const uint32_t* sel = ov_sel(a, o);
    const bool need[2] = {sel[SEL_NEED_TIE] != 0, sel[OV_SEL + SEL_NEED_TIE] != 0};
    if (!need[0] && !need[1]) return;
    const uint32_t thr[2] = {sel[SEL_PREFIX], sel[OV_SEL + SEL_PREFIX]};
    const float* d = a.origins[o].d_delta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int part = blockIdx.x * 8 + warp;
    if (part >= a.n_parts) return;
    int c[2] = {0, 0};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int j = part * 256 + g * 128 + lane * 4;
        if (j < a.n) {
            uint32_t kr[4], kt[4];
            ov_keys4(d, a.stride, j, kr, kt, a.key_mode);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (j + e < a.n) {
                    c[0] += kr[e] == thr[0];
                    c[1] += kt[e] == thr[1];
                }
            }
        }
    }
    c[0] = __reduce_add_sync(0xffffffffu, c[0]);
    c[1] = __reduce_add_sync(0xffffffffu, c[1]);
    if (lane == 0) {
        ov_tie(a, o)[part] = c[0];
        ov_tie(a, o)[a.n_parts + part] = c[1];
    }
}

// One block per (array, origin): smallest molecule index still inside the top-k among the ties.
__global__ void __launch_bounds__(256) ov_tiepick_kernel(const OvArgs a)
{
    __shared__ uint32_t sh[256];
    __shared__ int found_part;
    __shared__ uint32_t found_rem;
    const int arr = blockIdx.x, o = blockIdx.y;
    uint32_t* sel = ov_sel(a, o) + arr * OV_SEL;
    if (!sel[SEL_NEED_TIE]) return;
    const uint32_t* tie = ov_tie(a, o) + arr * a.n_parts;
    const uint32_t remaining = sel[SEL_REMAINING];
    // walk the parts from the end in chunks of 256, carrying the count found so far
    uint32_t carried = 0;
    if (threadIdx.x == 0) found_part = -1;
    __syncthreads();
    for (int hi = a.n_parts; hi > 0 && found_part < 0; hi -= 256) {
        const int p = hi - 1 - threadIdx.x;  // thread 0 owns the last part of the chunk
        const uint32_t mine = p >= 0 ? tie[p] : 0;
        sh[threadIdx.x] = mine;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {  // inclusive prefix over threads (= suffix over parts)
            const uint32_t add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        const uint32_t incl = carried + sh[threadIdx.x], excl = incl - mine;
        if (p >= 0 && excl < remaining && incl >= remaining) {
            found_part = p;
            found_rem = remaining - excl;
        }
        carried += sh[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t idxcut = 0;
        if (found_part >= 0) {
            const float* d = a.origins[o].d_delta;
            const uint32_t thr = sel[SEL_PREFIX];
            uint32_t need = found_rem;
            const int lo = found_part * 256;
            int hi = lo + 255;
            if (hi >= a.n) hi = a.n - 1;
            for (int j = hi; j >= lo; --j) {
                uint32_t key;
                if (arr == 0) {
                    const float dx = d[j], dy = d[a.stride + j];
                    key = a.key_mode ? __float_as_uint(dx)
                                     : __float_as_uint(__fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
                } else {
                    key = __float_as_uint(fabsf(d[2 * a.stride + j]));
                }
                if (key == thr && --need == 0) {
                    idxcut = (uint32_t)j;
                    break;
                }
            }
        }
        sel[SEL_IDXCUT] = idxcut;
    }
}

__global__ void __launch_bounds__(OV_THREADS) ov_count_kernel(const OvArgs a)
{
    __shared__ int warp_counts[OV_THREADS / 32];
    const int o = blockIdx.y;
    const float* d = a.origins[o].d_delta;
    const uint32_t* sel = ov_sel(a, o);
    const uint32_t thr[2] = {sel[SEL_PREFIX], sel[OV_SEL + SEL_PREFIX]};
    const uint32_t cut[2] = {sel[SEL_IDXCUT], sel[OV_SEL + SEL_IDXCUT]};
    int count = 0;
#pragma unroll
    for (int it = 0; it < OV_ITERS; ++it) {
        const int j = blockIdx.x * OV_BLOCK_MOLS + (it * OV_THREADS + threadIdx.x) * 4;
        if (j >= a.n) break;
        uint32_t kr[4], kt[4];
        ov_keys4(d, a.stride, j, kr, kt, a.key_mode);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t idx = (uint32_t)(j + e);
            const bool in_r = kr[e] > thr[0] || (kr[e] == thr[0] && idx >= cut[0]);
            const bool in_t = kt[e] > thr[1] || (kt[e] == thr[1] && idx >= cut[1]);
            count += (j + e < a.n) && in_r && in_t;
        }
    }
    count = __reduce_add_sync(0xffffffffu, count);
    if ((threadIdx.x & 31) == 0) warp_counts[threadIdx.x >> 5] = count;
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int w = 0; w < OV_THREADS / 32; ++w) total += warp_counts[w];
        if (total) atomicAdd(a.sums + (long)o * SDB_NSUM + SDB_S_OVERLAP, (double)total);
    }
}

long ov_words_per_origin(int n) { return 2L * OV_BINS + 2L * OV_SEL + 2L * ((n + 255) / 256); }

}  // namespace

extern "C" size_t sdb_overlap_scratch_bytes(int32_t n_origins, int32_t n)
{
    return (size_t)n_origins * (size_t)ov_words_per_origin(n) * sizeof(uint32_t);
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
    OvArgs a;
    a.origins = d_origins;
    a.scratch = static_cast<uint32_t*>(d_scratch);
    a.sums = d_sums;
    a.n = n;
    a.stride = stride;
    a.n_parts = (n + 255) / 256;
    a.words_per_origin = ov_words_per_origin(n);
    a.k = (uint32_t)k;
    a.key_mode = key_mode;
    int rc = sdb_check_cuda(cudaMemset2DAsync(d_sums + SDB_S_OVERLAP, SDB_NSUM * sizeof(double), 0, sizeof(double),
                                              (size_t)n_origins, s),
                            "sdb_overlap: clear");
    if (rc) return rc;
    ov_init_kernel<<<n_origins, 256, 0, s>>>(a);
    SDB_CHECK_LAUNCH("ov_init_kernel");
    const dim3 grid_n((n + OV_BLOCK_MOLS - 1) / OV_BLOCK_MOLS, n_origins);
    for (int pass = 0; pass < 3; ++pass) {
        ov_hist_kernel<<<grid_n, OV_THREADS, 0, s>>>(a, pass);
        SDB_CHECK_LAUNCH("ov_hist_kernel");
        ov_pick_kernel<<<dim3(2, n_origins), 256, 0, s>>>(a, pass);
        SDB_CHECK_LAUNCH("ov_pick_kernel");
    }
    ov_tiecount_kernel<<<dim3((a.n_parts + 7) / 8, n_origins), 256, 0, s>>>(a);
    SDB_CHECK_LAUNCH("ov_tiecount_kernel");
    ov_tiepick_kernel<<<dim3(2, n_origins), 256, 0, s>>>(a);
    SDB_CHECK_LAUNCH("ov_tiepick_kernel");
    ov_count_kernel<<<grid_n, OV_THREADS, 0, s>>>(a);
    SDB_CHECK_LAUNCH("ov_count_kernel");
    return SDB_OK;
}
