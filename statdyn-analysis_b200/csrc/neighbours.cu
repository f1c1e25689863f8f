// neighbours.cu -- k nearest neighbours within a cutoff in a periodic 2D box (cell list) and the
// orientational order parameter.  Replaces freud.locality.NearestNeighbors(rmax, k, strict_cut=True)
// as used by reference order.py:174-207,265-281,284-309 and the rowan arithmetic of
// order.py:28-86 (_neighbour_relative_angle / _orientational_order).
//
// Pipeline (all on one stream):
//   1. bin:      fractional coordinate -> cell id, per-cell counts (atomics)
//   2. scan:     exclusive prefix sum of the counts (two-level)
//   3. scatter:  (x, y, index) of every molecule into cell order
//   4. search:   one thread per molecule (in cell order, so neighbouring threads share candidate
//                cells in L1): 3x3 stencil, distance = dot(box.wrap(r_j - r_i)) evaluated exactly as
//                freud does (f32, one rounding per operation), strict rsq < rmax^2, top-k kept
//                sorted by (rsq, index) in registers; writes the list, r^2, the capped count and
//                the order parameter  mean_k cos^2(2 * intrinsic_distance(q_nb, q_i)).
#include <cstdlib>

#include "sdb_internal.cuh"

namespace {

struct NbArgs {
    const float* pos;
    const float* quat;
    int n;
    float lx, ly, xy;
    int ncx, ncy;
    float rmaxsq;
    float inv_lx, inv_ly, filt;  // two-phase search: cheap minimum image and its acceptance bound (rmax + eps)^2
    int k;      // capacity of the search: the count saturates at k
    int k_out;  // width of the list / r^2 rows and number of neighbours in the order parameter (<= k)
    uint32_t* cell_of;     // [n]
    uint32_t* cell_count;  // [ncells + 1] -> after scan: cell_start
    uint32_t* cell_fill;   // [ncells]
    uint32_t* block_sums;  // scan scratch
    float4* sorted;        // [n] (x, y, index bits, 0)
    uint32_t* nlist;
    float* rsq;
    uint32_t* counts;
    double* order;
};

__device__ __forceinline__ int cell_coord(float v, float L, int nc)
{
    // fraction in [0, 1): same arithmetic family as Box::makeFraction + wrap, any consistent
    // assignment works because the search visits the full 3x3 neighbourhood
    float f = (v + 0.5f * L) / L;
    f -= floorf(f);
    int c = (int)(f * (float)nc);
    return min(max(c, 0), nc - 1);
}

__global__ void nb_bin_kernel(const NbArgs a)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n) return;
    const float x = a.pos[3L * j], y = a.pos[3L * j + 1];
    const int cy = cell_coord(y, a.ly, a.ncy);
    const int cx = cell_coord(x - a.xy * y, a.lx, a.ncx);
    const uint32_t c = (uint32_t)(cy * a.ncx + cx);
    a.cell_of[j] = c;
    atomicAdd(a.cell_count + c, 1u);
}

// exclusive scan, level 1: 1024 elements per block
__global__ void __launch_bounds__(256) nb_scan_local(uint32_t* data, int n, uint32_t* block_sums)
{
    __shared__ uint32_t sh[256];
    const int base = blockIdx.x * 1024 + threadIdx.x * 4;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = base + i < n ? data[base + i] : 0;
        sum += v[i];
    }
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const uint32_t add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t run = sh[threadIdx.x] - sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 255) block_sums[blockIdx.x] = sh[255];
}

__global__ void __launch_bounds__(1024) nb_scan_blocks(uint32_t* block_sums, int n_blocks)
{
    __shared__ uint32_t sh[1024];
    uint32_t carry = 0;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const uint32_t add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < n_blocks) block_sums[i] = carry + sh[threadIdx.x] - v;
        carry += sh[1023];
        __syncthreads();
    }
}

// Exclusive scan of up to 64 k counters by ONE block (the cell counts of a 32 k-molecule frame are ~13 k
// numbers: three launches of the two-level scan cost more than the work).
__global__ void __launch_bounds__(1024) nb_scan_single(uint32_t* data, int n)
{
    __shared__ uint32_t warp_sums[32];
    const int per = (n + 1023) / 1024;
    const int begin = min(threadIdx.x * per, n), end = min(begin + per, n);
    uint32_t local = 0;
    for (int i = begin; i < end; ++i) local += data[i];
    // inclusive scan of the per-thread sums across the block
    uint32_t v = local;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    uint32_t run = v - local + (warp ? warp_sums[warp - 1] : 0u);  // exclusive prefix of this thread's range
    for (int i = begin; i < end; ++i) {
        const uint32_t c = data[i];
        data[i] = run;
        run += c;
    }
}

__global__ void nb_scan_add(uint32_t* data, int n, const uint32_t* block_sums)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += block_sums[i / 1024];
}

__global__ void nb_scatter_kernel(const NbArgs a)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n) return;
    const uint32_t c = a.cell_of[j];
    const uint32_t slot = a.cell_count[c] + atomicAdd(a.cell_fill + c, 1u);
    a.sorted[slot] = make_float4(a.pos[3L * j], a.pos[3L * j + 1], __uint_as_float((uint32_t)j), 0.0f);
}

template <int K>
__global__ void __launch_bounds__(128) nb_search_kernel(const NbArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const float4 me = a.sorted[t];
    const uint32_t i = __float_as_uint(me.z);
    // the centre cell is recomputed from the position exactly as in nb_bin_kernel
    const int cy = cell_coord(me.y, a.ly, a.ncy);
    const int cx = cell_coord(me.x - a.xy * me.y, a.lx, a.ncx);

    float best_r[K];
    uint32_t best_i[K];
#pragma unroll
    for (int s = 0; s < K; ++s) {
        best_r[s] = INFINITY;
        best_i[s] = 0xFFFFFFFFu;
    }
    int found = 0;

    const int span_x = min(a.ncx, 3), span_y = min(a.ncy, 3);
    for (int oy = 0; oy < span_y; ++oy) {
        int yy = cy + oy - (span_y == 3 ? 1 : 0);
        yy = (yy + a.ncy) % a.ncy;
        for (int ox = 0; ox < span_x; ++ox) {
            int xx = cx + ox - (span_x == 3 ? 1 : 0);
            xx = (xx + a.ncx) % a.ncx;
            const uint32_t c = (uint32_t)(yy * a.ncx + xx);
            const uint32_t begin = a.cell_count[c], end = a.cell_count[c + 1];
            for (uint32_t p = begin; p < end; ++p) {
                const float4 other = __ldg(a.sorted + p);
                const uint32_t j = __float_as_uint(other.z);
                if (j == i) continue;
                float wx, wy;
                sdb_wrap2d(a.lx, a.ly, a.xy, __fsub_rn(other.x, me.x), __fsub_rn(other.y, me.y), wx, wy);
                // dot(rij, rij) = (x*x + y*y) + z*z with z = 0
                const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), 0.0f);
                if (!(r2 < a.rmaxsq)) continue;
                ++found;
                // insert (r2, j) keeping the list ascending by (r2, index)
                float cr = r2;
                uint32_t ci = j;
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    const bool before = cr < best_r[s] || (cr == best_r[s] && ci < best_i[s]);
                    const float tr = best_r[s];
                    const uint32_t ti = best_i[s];
                    best_r[s] = before ? cr : tr;
                    best_i[s] = before ? ci : ti;
                    cr = before ? tr : cr;
                    ci = before ? ti : ci;
                }
            }
        }
    }

    const int kk = a.k_out;  // K is the compiled capacity, kk <= a.k <= K the requested width
    if (a.nlist) {
#pragma unroll
        for (int s = 0; s < K; ++s)
            if (s < kk) a.nlist[(long)i * kk + s] = best_i[s];
    }
    if (a.rsq) {
#pragma unroll
        for (int s = 0; s < K; ++s)
            if (s < kk) a.rsq[(long)i * kk + s] = best_i[s] == 0xFFFFFFFFu ? -1.0f : best_r[s];
    }
    if (a.counts) a.counts[i] = (uint32_t)min(found, a.k);
    if (a.order && a.quat) {
        const float4 q = sdb_ldg4(a.quat + 4L * i);
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < K; ++s) {
            if (s < kk) {
                const uint32_t nb = best_i[s] == 0xFFFFFFFFu ? i : best_i[s];
                const float4 p = sdb_ldg4(a.quat + 4L * nb);
                // multiply(conjugate(p), q) with f32 products (rowan on f32 inputs), then
                // cos^2(2 acos(w/|r|)) = (2 w^2/|r|^2 - 1)^2 in f64
                const float cx_ = -p.y, cy_ = -p.z, cz_ = -p.w;
                const float rw = __fsub_rn(__fmul_rn(p.x, q.x),
                                           __fadd_rn(__fadd_rn(__fmul_rn(cx_, q.y), __fmul_rn(cy_, q.z)), __fmul_rn(cz_, q.w)));
                const float rx = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.y), __fmul_rn(q.x, cx_)),
                                           __fsub_rn(__fmul_rn(cy_, q.w), __fmul_rn(cz_, q.z)));
                const float ry = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.z), __fmul_rn(q.x, cy_)),
                                           __fsub_rn(__fmul_rn(cz_, q.y), __fmul_rn(cx_, q.w)));
                const float rz = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.w), __fmul_rn(q.x, cz_)),
                                           __fsub_rn(__fmul_rn(cx_, q.z), __fmul_rn(cy_, q.y)));
                const double w2 = (double)rw * (double)rw;
                const double n2 = w2 + (double)rx * (double)rx + (double)ry * (double)ry + (double)rz * (double)rz;
                const double c2 = 2.0 * w2 / n2 - 1.0;
                acc += c2 * c2;
            }
        }
        a.order[i] = acc / (double)kk;
    }
}

// ---- two-phase search (orthorhombic boxes with at least 3 x 3 cells) ------------------------------------
// The exact distance (freud's wrap: two IEEE divisions, ~45 instructions) and the sorted insertion
// (~50) are needed only for the ~8 of ~22 candidates that lie inside the cutoff, but in the
// one-phase kernel above a warp pays for them at every candidate (some lane always accepts).  Here
// phase A walks the three cell rows (the 3 cells of a row are one contiguous range of the sorted array)
// with a cheap minimum image d - L rint(d / L) and keeps the indices of the candidates within
// rmax + eps in a per-thread list in shared memory (eps covers the f32 rounding of either distance:
// nothing the exact test accepts is dropped); phase B evaluates exactly those, so its cost follows
// the number of real neighbours.  The results are identical to nb_search_kernel.
constexpr int NB_LIST = 24;

template <int K>
__device__ __forceinline__ void nb_exact_insert(const NbArgs& a, const float4 me, const uint32_t i, const uint32_t p,
                                                float (&best_r)[K], uint32_t (&best_i)[K], int& found)
{
    const float4 other = __ldg(a.sorted + p);
    const uint32_t j = __float_as_uint(other.z);
    if (j == i) return;
    float wx, wy;
    sdb_wrap2d(a.lx, a.ly, a.xy, __fsub_rn(other.x, me.x), __fsub_rn(other.y, me.y), wx, wy);
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), 0.0f);
    if (!(r2 < a.rmaxsq)) return;
    ++found;
    float cr = r2;
    uint32_t ci = j;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        const bool before = cr < best_r[s] || (cr == best_r[s] && ci < best_i[s]);
        const float tr = best_r[s];
        const uint32_t ti = best_i[s];
        best_r[s] = before ? cr : tr;
        best_i[s] = before ? ci : ti;
        cr = before ? tr : cr;
        ci = before ? ti : ci;
    }
}

template <int K>
__device__ __forceinline__ void nb_write_results(const NbArgs& a, const uint32_t i, const float (&best_r)[K],
                                                 const uint32_t (&best_i)[K], const int found)
{
    const int kk = a.k_out;  // K is the compiled capacity, kk <= a.k <= K the requested width
    if (a.nlist) {
#pragma unroll
        for (int s = 0; s < K; ++s)
            if (s < kk) a.nlist[(long)i * kk + s] = best_i[s];
    }
    if (a.rsq) {
#pragma unroll
        for (int s = 0; s < K; ++s)
            if (s < kk) a.rsq[(long)i * kk + s] = best_i[s] == 0xFFFFFFFFu ? -1.0f : best_r[s];
    }
    if (a.counts) a.counts[i] = (uint32_t)min(found, a.k);
    if (a.order && a.quat) {
        const float4 q = sdb_ldg4(a.quat + 4L * i);
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < K; ++s) {
            if (s < kk) {
                const uint32_t nb = best_i[s] == 0xFFFFFFFFu ? i : best_i[s];
                const float4 p = sdb_ldg4(a.quat + 4L * nb);
                // multiply(conjugate(p), q) with f32 products (rowan on f32 inputs), then
                // cos^2(2 acos(w/|r|)) = (2 w^2/|r|^2 - 1)^2 in f64
                const float cx_ = -p.y, cy_ = -p.z, cz_ = -p.w;
                const float rw = __fsub_rn(__fmul_rn(p.x, q.x),
                                           __fadd_rn(__fadd_rn(__fmul_rn(cx_, q.y), __fmul_rn(cy_, q.z)), __fmul_rn(cz_, q.w)));
                const float rx = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.y), __fmul_rn(q.x, cx_)),
                                           __fsub_rn(__fmul_rn(cy_, q.w), __fmul_rn(cz_, q.z)));
                const float ry = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.z), __fmul_rn(q.x, cy_)),
                                           __fsub_rn(__fmul_rn(cz_, q.y), __fmul_rn(cx_, q.w)));
                const float rz = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.w), __fmul_rn(q.x, cz_)),
                                           __fsub_rn(__fmul_rn(cx_, q.z), __fmul_rn(cy_, q.y)));
                const double w2 = (double)rw * (double)rw;
                const double n2 = w2 + (double)rx * (double)rx + (double)ry * (double)ry + (double)rz * (double)rz;
                const double c2 = 2.0 * w2 / n2 - 1.0;
                acc += c2 * c2;
            }
        }
        a.order[i] = acc / (double)kk;
    }
}

template <int K>
__global__ void __launch_bounds__(128) nb_search2_kernel(const NbArgs a)
{
    __shared__ uint32_t surv[NB_LIST][128];  // [slot][thread]: conflict-free
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const float4 me = a.sorted[t];
    const uint32_t i = __float_as_uint(me.z);
    const int cy = cell_coord(me.y, a.ly, a.ncy);
    const int cx = cell_coord(me.x, a.lx, a.ncx);  // xy == 0

    float best_r[K];
    uint32_t best_i[K];
#pragma unroll
    for (int s = 0; s < K; ++s) {
        best_r[s] = INFINITY;
        best_i[s] = 0xFFFFFFFFu;
    }
    int found = 0, cnt = 0;

    // phase A: candidates of the 3 x 3 cells; a row of cells is one range of the sorted array, or two
    // when it crosses the periodic boundary in x
    for (int oy = -1; oy <= 1; ++oy) {
        int yy = cy + oy;
        yy = yy < 0 ? yy + a.ncy : (yy >= a.ncy ? yy - a.ncy : yy);
        const uint32_t* row = a.cell_count + (long)yy * a.ncx;
        uint32_t b0, e0, b1 = 0, e1 = 0;
        if (cx >= 1 && cx + 1 < a.ncx) {
            b0 = row[cx - 1];
            e0 = row[cx + 2];
        } else if (cx == 0) {
            b0 = row[0];
            e0 = row[2];
            b1 = row[a.ncx - 1];
            e1 = row[a.ncx];
        } else {
            b0 = row[a.ncx - 2];
            e0 = row[a.ncx];
            b1 = row[0];
            e1 = row[1];
        }
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            const uint32_t begin = part ? b1 : b0, end = part ? e1 : e0;
            for (uint32_t p = begin; p < end; ++p) {
                const float4 other = __ldg(a.sorted + p);
                const float dx = other.x - me.x, dy = other.y - me.y;
                const float ax = fmaf(-a.lx, rintf(dx * a.inv_lx), dx);
                const float ay = fmaf(-a.ly, rintf(dy * a.inv_ly), dy);
                if (fmaf(ax, ax, ay * ay) < a.filt) {
                    if (cnt < NB_LIST) surv[cnt++][threadIdx.x] = p;
                    else nb_exact_insert<K>(a, me, i, p, best_r, best_i, found);  // list full: evaluate in place
                }
            }
        }
    }
    // phase B: exact distance and sorted insertion of the survivors
    for (int s = 0; s < cnt; ++s) nb_exact_insert<K>(a, me, i, surv[s][threadIdx.x], best_r, best_i, found);
    nb_write_results<K>(a, i, best_r, best_i, found);
}

__global__ void order_from_list_kernel(const uint32_t* __restrict__ nlist, int n, int k, const float* __restrict__ quat,
                                       double* __restrict__ order, double* __restrict__ angles)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q = sdb_ldg4(quat + 4L * i);
    double acc = 0.0;
    for (int s = 0; s < k; ++s) {
        uint32_t nb = nlist[(long)i * k + s];
        if (nb >= (uint32_t)n) nb = i;  // order.py:58-59
        const float4 p = sdb_ldg4(quat + 4L * nb);
        const float cx_ = -p.y, cy_ = -p.z, cz_ = -p.w;
        const float rw = __fsub_rn(__fmul_rn(p.x, q.x),
                                   __fadd_rn(__fadd_rn(__fmul_rn(cx_, q.y), __fmul_rn(cy_, q.z)), __fmul_rn(cz_, q.w)));
        const float rx = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.y), __fmul_rn(q.x, cx_)),
                                   __fsub_rn(__fmul_rn(cy_, q.w), __fmul_rn(cz_, q.z)));
        const float ry = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.z), __fmul_rn(q.x, cy_)),
                                   __fsub_rn(__fmul_rn(cz_, q.y), __fmul_rn(cx_, q.w)));
        const float rz = __fadd_rn(__fadd_rn(__fmul_rn(p.x, q.w), __fmul_rn(q.x, cz_)),
                                   __fsub_rn(__fmul_rn(cx_, q.z), __fmul_rn(cy_, q.y)));
        const double w2 = (double)rw * (double)rw;
        const double n2 = w2 + (double)rx * (double)rx + (double)ry * (double)ry + (double)rz * (double)rz;
        const double c2 = 2.0 * w2 / n2 - 1.0;
        acc += c2 * c2;
        // 2 * intrinsic_distance = 2 * acos(w / |r|) in [0, 2 pi]  (order.py:61-65)
        if (angles) angles[(long)i * k + s] = 2.0 * acos(fmin(fmax((double)rw / sqrt(n2), -1.0), 1.0));
    }
    if (order) order[i] = acc / (double)k;
}

struct NbLayout {
    int ncx, ncy, ncells, scan_blocks;
    size_t off_cell_of, off_count, off_fill, off_block, off_sorted, total;
};

bool nb_layout(const float* box, int n, float rmax, NbLayout& L)
{
    const float lx = box[0], ly = box[1], xy = box[3];
    if (!(lx > 0.f) || !(ly > 0.f) || !(rmax > 0.f)) return false;
    // cell widths measured perpendicular to the cell faces must be >= rmax
    const float shear = sqrtf(1.0f + xy * xy);
    L.ncx = (int)floorf(lx / (rmax * shear * 1.0001f));
    L.ncy = (int)floorf(ly / (rmax * 1.0001f));
    L.ncx = L.ncx < 1 ? 1 : L.ncx;
    L.ncy = L.ncy < 1 ? 1 : L.ncy;
    // no point in having (many) more cells than molecules
    while ((long)L.ncx * L.ncy > 4L * n + 16 && (L.ncx > 3 || L.ncy > 3)) {
        if (L.ncx > 3) L.ncx = (L.ncx + 1) / 2;
        if (L.ncy > 3) L.ncy = (L.ncy + 1) / 2;
    }
    L.ncells = L.ncx * L.ncy;
    L.scan_blocks = (L.ncells + 1 + 1023) / 1024;
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    size_t off = 0;
    L.off_cell_of = off; off = align(off + sizeof(uint32_t) * (size_t)n);
    L.off_count = off;   off = align(off + sizeof(uint32_t) * (size_t)(L.ncells + 1));
    L.off_fill = off;    off = align(off + sizeof(uint32_t) * (size_t)L.ncells);
    L.off_block = off;   off = align(off + sizeof(uint32_t) * (size_t)L.scan_blocks);
    L.off_sorted = off;  off = align(off + sizeof(float4) * (size_t)n);
    L.total = off;
    return true;
}

}  // namespace

extern "C" size_t sdb_neighbours_scratch_bytes(const float* h_box, int32_t n, float rmax)
{
    NbLayout L;
    if (!h_box || n <= 0 || !nb_layout(h_box, n, rmax, L)) return 0;
    return L.total;
}

extern "C" int sdb_neighbours(const float* h_box, int32_t n, const float* d_pos, const float* d_quat, float rmax, int32_t k,
                              uint32_t* d_nlist, float* d_rsq, uint32_t* d_counts, double* d_order, void* d_scratch,
                              size_t scratch_bytes, void* stream)
{
    return sdb_neighbours_ex(h_box, n, d_pos, d_quat, rmax, k, k, d_nlist, d_rsq, d_counts, d_order, d_scratch, scratch_bytes,
                             stream);
}

extern "C" int sdb_neighbours_ex(const float* h_box, int32_t n, const float* d_pos, const float* d_quat, float rmax, int32_t k,
                                 int32_t k_list, uint32_t* d_nlist, float* d_rsq, uint32_t* d_counts, double* d_order,
                                 void* d_scratch, size_t scratch_bytes, void* stream)
{
    NbLayout L;
    if (!h_box || !d_pos || !d_scratch || n <= 0 || k <= 0 || k > SDB_MAX_K || k_list <= 0 || k_list > k ||
        !nb_layout(h_box, n, rmax, L)) {
        sdb_set_error("sdb_neighbours: bad argument (n=%d k=%d)", n, k);
        return SDB_EINVAL;
    }
    if (scratch_bytes < L.total) {
        sdb_set_error("sdb_neighbours: scratch too small (%zu < %zu)", scratch_bytes, L.total);
        return SDB_EINVAL;
    }
    if (d_order && !d_quat) {
        sdb_set_error("sdb_neighbours: order parameter requested without orientations");
        return SDB_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(d_scratch);
    NbArgs a;
    a.pos = d_pos;
    a.quat = d_quat;
    a.n = n;
    a.lx = h_box[0];
    a.ly = h_box[1];
    a.xy = h_box[3];
    a.ncx = L.ncx;
    a.ncy = L.ncy;
    a.rmaxsq = rmax * rmax;
    a.k = k;
    a.k_out = k_list;
    a.cell_of = reinterpret_cast<uint32_t*>(base + L.off_cell_of);
    a.cell_count = reinterpret_cast<uint32_t*>(base + L.off_count);
    a.cell_fill = reinterpret_cast<uint32_t*>(base + L.off_fill);
    a.block_sums = reinterpret_cast<uint32_t*>(base + L.off_block);
    a.sorted = reinterpret_cast<float4*>(base + L.off_sorted);
    a.nlist = d_nlist;
    a.rsq = d_rsq;
    a.counts = d_counts;
    a.order = d_order;
    // counts, fill and block sums are contiguous: one clear
    int rc = sdb_check_cuda(cudaMemsetAsync(base + L.off_count, 0, L.off_sorted - L.off_count, s), "sdb_neighbours: clear");
    if (rc) return rc;
    const int nblk = (n + 255) / 256;
    nb_bin_kernel<<<nblk, 256, 0, s>>>(a);
    SDB_CHECK_LAUNCH("nb_bin_kernel");
    if (L.ncells + 1 <= 65536) {
        nb_scan_single<<<1, 1024, 0, s>>>(a.cell_count, L.ncells + 1);
        SDB_CHECK_LAUNCH("nb_scan_single");
    } else {
        nb_scan_local<<<L.scan_blocks, 256, 0, s>>>(a.cell_count, L.ncells + 1, a.block_sums);
        SDB_CHECK_LAUNCH("nb_scan_local");
    }
    if (L.ncells + 1 > 65536 && L.scan_blocks > 1) {
        nb_scan_blocks<<<1, 1024, 0, s>>>(a.block_sums, L.scan_blocks);
        SDB_CHECK_LAUNCH("nb_scan_blocks");
        nb_scan_add<<<(L.ncells + 1 + 255) / 256, 256, 0, s>>>(a.cell_count, L.ncells + 1, a.block_sums);
        SDB_CHECK_LAUNCH("nb_scan_add");
    }
    nb_scatter_kernel<<<nblk, 256, 0, s>>>(a);
    SDB_CHECK_LAUNCH("nb_scatter_kernel");
    const int sblk = (n + 127) / 128;
    static const bool two_phase_on = [] { const char* e = getenv("SDB_NB_TWO_PHASE"); return !(e && e[0] == '0'); }();
    if (two_phase_on && a.xy == 0.0f && L.ncx >= 3 && L.ncy >= 3) {
        // bound of the cheap distance: rmax plus the f32 rounding of both evaluations (a few ulp of L)
        const float eps = 8.0f * ((nextafterf(a.lx, INFINITY) - a.lx) + (nextafterf(a.ly, INFINITY) - a.ly));
        a.inv_lx = 1.0f / a.lx;
        a.inv_ly = 1.0f / a.ly;
        a.filt = (rmax + eps) * (rmax + eps);
        if (k <= 8)
            nb_search2_kernel<8><<<sblk, 128, 0, s>>>(a);
        else if (k <= 9)
            nb_search2_kernel<9><<<sblk, 128, 0, s>>>(a);
        else
            nb_search2_kernel<SDB_MAX_K><<<sblk, 128, 0, s>>>(a);
    } else if (k <= 8)
        nb_search_kernel<8><<<sblk, 128, 0, s>>>(a);
    else if (k <= 9)
        nb_search_kernel<9><<<sblk, 128, 0, s>>>(a);
    else
        nb_search_kernel<SDB_MAX_K><<<sblk, 128, 0, s>>>(a);
    SDB_CHECK_LAUNCH("nb_search_kernel");
    return SDB_OK;
}

extern "C" int sdb_orientational_order(const uint32_t* d_nlist, int32_t n, int32_t k, const float* d_quat, double* d_order,
                                       double* d_angles, void* stream)
{
    if (!d_nlist || !d_quat || (!d_order && !d_angles) || n <= 0 || k <= 0) {
        sdb_set_error("sdb_orientational_order: bad argument");
        return SDB_EINVAL;
    }
    order_from_list_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_nlist, n, k, d_quat, d_order,
                                                                                           d_angles);
    SDB_CHECK_LAUNCH("order_from_list_kernel");
    return SDB_OK;
}
