// struct.cu -- the all-site structural relaxation `struct` of Dynamics.compute_all
// (reference dynamics/dynamics.py:305-326 with dynamics/_util.py:42-54).
//
// mode 0 reproduces the reference exactly, including its two quirks:
//   * the z-angle is fed to rowan.from_euler as gamma, i.e. sites are rotated about x by dtheta_z;
//   * rotated sites are concatenated site-major, translations repeated molecule-major, so row
//     i = s*N + j (site s of molecule j) is translated by delta_translation[i / n_sites].
// mode 1 is the physical definition: every molecule's own sites rotated in-plane by its own angle
// and translated by its own displacement.
//
// One thread per (origin, molecule j); every row is decided by an f32 filter unless it lies within a
// safety margin of the cutoff, in which case the exact f64 sequence of rowan.rotate is evaluated.
#include <cstdlib>

#include <algorithm>
#include <cstdlib>

#include "sdb_internal.cuh"

namespace {

struct StructArgs {
    const SdbOrigin* origins;
    double* sums;
    int n;
    long stride;
    int n_sites;
    int mode;
    double cut2;  // sum of squares < cut2  <=>  fl64(sqrt(sum)) < dist
    float sx[8], sy[8];
};

// Exact reference row (f64, operation order of rowan.rotate; see oracle/rowan_shim.py).
__device__ __noinline__ bool struct_row_exact(float theta, double tx, double ty, double vx, double vy, double cut2)
{
    // rowan.from_axis_angle: half angle, cos and sin in f32; basis axis in f64
    const float half = theta / 2.0f;
    const double c = (double)cosf(half), s = (double)sinf(half);
    // inner = (0, v) (x) conj(q) = (vx s, c vx, c vy, vy s) ; outer = q (x) inner
    const double iw = __dmul_rn(vx, s), ix = __dmul_rn(c, vx), iy = __dmul_rn(c, vy), iz = __dmul_rn(vy, s);
    const double rx = __dadd_rn(__dmul_rn(c, ix), __dmul_rn(iw, s));
    const double ry = __dadd_rn(__dmul_rn(c, iy), -__dmul_rn(s, iz));
    const double rz = __dadd_rn(__dmul_rn(c, iz), __dmul_rn(s, iy));
    const double ex = __dadd_rn(__dadd_rn(rx, tx), -vx);
    const double ey = __dadd_rn(__dadd_rn(ry, ty), -vy);
    const double ss = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(rz, rz));
    return ss < cut2;
}

// 2 (1 - cos x) = 4 sin^2(x/2) with ~1e-7 absolute error (Cody-Waite reduction + odd polynomial);
// only used by the f32 filter below, whose margin is two orders of magnitude wider.
__device__ __forceinline__ float two_one_minus_cos(float x)
{
    if (fabsf(x) > 8192.0f) {
        const float sh = sinf(0.5f * x);
        return 4.0f * sh * sh;
    }
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(-k, 6.28125f, x);
    r = fmaf(-k, 1.9353071795864769e-3f, r);
    const float h = 0.5f * r, h2 = h * h;
    float p = fmaf(h2, 1.6059043836821613e-10f, -2.5052108385441720e-8f);
    p = fmaf(h2, p, 2.7557319223985893e-6f);
    p = fmaf(h2, p, -1.9841269841269841e-4f);
    p = fmaf(h2, p, 8.3333333333333332e-3f);
    p = fmaf(h2, p, -1.6666666666666666e-1f);
    const float sn = fmaf(h * h2, p, h);
    return 4.0f * sn * sn;
}

// mode 0.  A site (vx, vy, 0) rotated about x by theta and translated by t moves by
//   e = (tx, ty - vy (1 - cos theta), vy sin theta)  =>  |e|^2 = |t|^2 + 2 (1 - cos theta) vy (vy - ty)
// up to terms of relative size 1e-7 (f32 cos/sin of the half angle are not exactly normalised).
// This f32 value decides every row that is not within a safety margin of the cutoff; the rows
// inside the margin (about one in 1e5) are decided by the exact f64 sequence above, so the count is
// identical to evaluating every row exactly.
//
// Each thread owns 4 consecutive rotation molecules j0..j0+3 (one float4 of dtheta).  Row
// i = site * n + j pairs with translation molecule i / n_sites: for 4 consecutive rows that is at
// most two consecutive molecules, fetched once per site.
template <int NSITES>
__global__ void __launch_bounds__(256) struct_kernel(const StructArgs a)
{
    const int o = blockIdx.y;
    const float* d = a.origins[o].d_delta;
    const unsigned n = (unsigned)a.n;
    const unsigned n_sites = NSITES > 0 ? (unsigned)NSITES : (unsigned)a.n_sites;
    const float cut = (float)a.cut2;
    int count = 0;
    for (unsigned j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u; j0 < n; j0 += gridDim.x * blockDim.x * 4u) {
        const float4 t4 = sdb_ldg4(d + 2 * a.stride + j0);
        const float theta[4] = {t4.x, t4.y, t4.z, t4.w};
        if (a.mode == 0) {
            float g[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) g[e] = two_one_minus_cos(theta[e]);
#pragma unroll
            for (int site = 0; site < (NSITES > 0 ? NSITES : 8); ++site) {
                if (site >= (int)n_sites) break;
                const unsigned row0 = (unsigned)site * n + j0;  // host guarantees n * n_sites < 2^32
                const unsigned m0 = row0 / n_sites, r0 = row0 - m0 * n_sites;
                const unsigned m1 = min(m0 + 1u, n - 1u);
                const float tx0 = __ldg(d + m0), ty0 = __ldg(d + a.stride + m0);
                const float tx1 = __ldg(d + m1), ty1 = __ldg(d + a.stride + m1);
                const float vx = a.sx[site], vy = a.sy[site];
                const float vmargin = 1.0f + 8.0f * vy * vy;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float tx, ty;
                    if (NSITES >= 3) {
                        const bool second = r0 + (unsigned)e >= n_sites;  // (r0 + e) / n_sites is 0 or 1
                        tx = second ? tx1 : tx0;
                        ty = second ? ty1 : ty0;
                    } else {  // generic site count: one lookup per row
                        const unsigned m = min((row0 + (unsigned)e) / n_sites, n - 1u);
                        tx = __ldg(d + m);
                        ty = __ldg(d + a.stride + m);
                    }
                    const float t2 = fmaf(tx, tx, ty * ty);
                    const float e2 = fmaf(g[e], vy * (vy - ty), t2);
                    const float margin = 8e-6f * (t2 + vmargin);
                    bool inside = e2 < cut;
                    if (!(fabsf(e2 - cut) > margin))
                        inside = struct_row_exact(theta[e], (double)tx, (double)ty, (double)vx, (double)vy, a.cut2);
                    count += inside && (j0 + (unsigned)e < n);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (j0 + (unsigned)e >= n) break;
                float sn, cs;
                sincosf(theta[e], &sn, &cs);
                const double c = cs, s = sn;
                const double tx = (double)d[j0 + e], ty = (double)d[a.stride + j0 + e];
                for (int site = 0; site < (int)n_sites; ++site) {
                    const double vx = (double)a.sx[site], vy = (double)a.sy[site];
                    const double ex = (c * vx - s * vy) + tx - vx;
                    const double ey = (s * vx + c * vy) + ty - vy;
                    count += (ex * ex + ey * ey) < a.cut2;
                }
            }
        }
    }
    count = __reduce_add_sync(0xffffffffu, count);
    __shared__ int warp_counts[8];
    if ((threadIdx.x & 31) == 0) warp_counts[threadIdx.x >> 5] = count;
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int w = 0; w < 8; ++w) total += warp_counts[w];
        // integer-valued doubles: the sum is exact, hence independent of the arrival order
        if (total) atomicAdd(a.sums + (long)o * SDB_NSUM + SDB_S_STRUCT, (double)total);
    }
}

// Fast kernel for three sites, mode 0.  Each thread owns 4 consecutive TRANSLATION molecules
// m0..m0+3, i.e. the 12 consecutive rows i = 3 m0 .. 3 m0 + 11 (row i is translated by molecule
// i / 3).  Row i = s n + j is site s of rotation molecule j: away from the two places where s
// changes the 12 rotation angles are 12 consecutive floats (three 16-byte loads when n is a
// multiple of 4).  Per row:  |e|^2 = |t|^2 + (1 - cos theta) 2 vy (vy - ty).
//
// Work mapping: a group is 4 translation molecules (12 rows).  Row i = s n + j reads the angle of rotation
// molecule j, so every angle is needed three times, by groups a third of the array apart.  A thread
// therefore handles, back to back, the three groups g_s + u (s = 0, 1, 2; g_s = the group holding row
// s n): their angles are the same 12 floats shifted by at most 11 elements, so the second and third
// read hit L1 and the dtheta plane crosses DRAM once instead of three times (12 B per update
// instead of 16-20).
__device__ __forceinline__ int struct3_group(const StructArgs& a, const float* __restrict__ d, const float* __restrict__ dth,
                                             const unsigned n, const float cut, const unsigned m0)
{
    int count = 0;
    {
        const float4 tx4 = sdb_ldg4(d + m0), ty4 = sdb_ldg4(d + a.stride + m0);
        const float tx[4] = {tx4.x, tx4.y, tx4.z, tx4.w};
        const float ty[4] = {ty4.x, ty4.y, ty4.z, ty4.w};
        const unsigned i0 = 3u * m0;  // host guarantees 3 n < 2^32
        const unsigned s0 = (i0 >= n) + (i0 >= 2u * n);
        const unsigned s_last = (i0 + 11u >= n) + (i0 + 11u >= 2u * n);
        const unsigned j0 = i0 - s0 * n;
        if (m0 + 4u <= n && s0 == s_last && (j0 & 3u) == 0u) {
            const float vx = s0 == 0u ? a.sx[0] : (s0 == 1u ? a.sx[1] : a.sx[2]);
            const float vy = s0 == 0u ? a.sy[0] : (s0 == 1u ? a.sy[1] : a.sy[2]);
            // |e|^2 - cut = (|t|^2 + g - cut) - g cos(theta),  g = 2 vy (vy - ty): one cos.approx (absolute error
            // < 2.5e-6 for |theta| <= 16, argument reduction included) and one fma per row
            const float c2 = 2.0f * vy * vy, b2 = 2.0f * vy;
            const float vmargin = 8e-6f * (1.0f + 8.0f * vy * vy);
            float g[4], base[4], mg[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float t2 = fmaf(tx[e], tx[e], ty[e] * ty[e]);
                g[e] = fmaf(-b2, ty[e], c2);
                base[e] = (t2 + g[e]) - cut;
                mg[e] = fmaf(8e-6f, t2, vmargin);
            }
            float th[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 v = sdb_ldg4(dth + j0 + 4 * q);
                th[4 * q] = v.x; th[4 * q + 1] = v.y; th[4 * q + 2] = v.z; th[4 * q + 3] = v.w;
            }
            // rows the fast filter cannot decide (within the margin, or |theta| > 16) are rare (~1e-5): the
            // group is then redone row by row.  (A NaN row is "not inside" on both paths.)
            float th_max = 0.0f;
#pragma unroll
            for (int r = 0; r < 12; ++r) th_max = fmaxf(th_max, fabsf(th[r]));
            bool undecided = th_max > 16.0f;
            int fast = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d0 = fmaf(-__cosf(th[3 * e]), g[e], base[e]);
                const float d1 = fmaf(-__cosf(th[3 * e + 1]), g[e], base[e]);
                const float d2 = fmaf(-__cosf(th[3 * e + 2]), g[e], base[e]);
                fast += (d0 < 0.0f) + (d1 < 0.0f) + (d2 < 0.0f);
                undecided |= !(fminf(fminf(fabsf(d0), fabsf(d1)), fabsf(d2)) > mg[e]);
            }
            if (!undecided) {
                count += fast;
            } else {
                for (unsigned r = 0; r < 12u; ++r) {
                    const unsigned m = m0 + r / 3u;
                    const float theta = __ldg(dth + j0 + r), txm = __ldg(d + m), tym = __ldg(d + a.stride + m);
                    const float t2m = fmaf(txm, txm, tym * tym);
                    const float e2 = fmaf(two_one_minus_cos(theta), vy * (vy - tym), t2m);
                    bool inside = e2 < cut;
                    if (!(fabsf(e2 - cut) > 8e-6f * (t2m + 1.0f + 8.0f * vy * vy)))
                        inside = struct_row_exact(theta, (double)txm, (double)tym, (double)vx, (double)vy, a.cut2);
                    count += inside;
                }
            }
        } else {  // a site boundary, the last molecules, or n not a multiple of 4: row by row
            for (unsigned r = 0; r < 12u; ++r) {
                const unsigned i = i0 + r, m = m0 + r / 3u;
                if (m >= n) break;
                const unsigned s = (i >= n) + (i >= 2u * n);
                const float theta = __ldg(dth + (i - s * n));
                const float vx = s == 0u ? a.sx[0] : (s == 1u ? a.sx[1] : a.sx[2]);
                const float vy = s == 0u ? a.sy[0] : (s == 1u ? a.sy[1] : a.sy[2]);
                const float txm = __ldg(d + m), tym = __ldg(d + a.stride + m);
                const float t2 = fmaf(txm, txm, tym * tym);
                const float e2 = fmaf(two_one_minus_cos(theta), vy * (vy - tym), t2);
                bool inside = e2 < cut;
                if (!(fabsf(e2 - cut) > 8e-6f * (t2 + 1.0f + 8.0f * vy * vy)))
                    inside = struct_row_exact(theta, (double)txm, (double)tym, (double)vx, (double)vy, a.cut2);
                count += inside;
            }
        }
    }
    return count;
}

__global__ void __launch_bounds__(256) struct3_kernel(const StructArgs a)
{
    const int o = blockIdx.y;
    const float* __restrict__ d = a.origins[o].d_delta;
    const float* __restrict__ dth = d + 2 * a.stride;
    const unsigned n = (unsigned)a.n;
    const float cut = (float)a.cut2;
    // first group of every site (host guarantees 3 n < 2^32) and the number of groups
    const unsigned g1 = n / 12u, g2 = (unsigned)((2ull * n) / 12ull), g3 = (n + 3u) / 4u;
    const unsigned span = max(g1, max(g2 - g1, g3 - g2));
    int count = 0;
    for (unsigned u = blockIdx.x * blockDim.x + threadIdx.x; u < span; u += gridDim.x * blockDim.x) {
        if (u < g1) count += struct3_group(a, d, dth, n, cut, 4u * u);
        if (g1 + u < g2) count += struct3_group(a, d, dth, n, cut, 4u * (g1 + u));
        if (g2 + u < g3) count += struct3_group(a, d, dth, n, cut, 4u * (g2 + u));
    }
    count = __reduce_add_sync(0xffffffffu, count);
    __shared__ int warp_counts[8];
    if ((threadIdx.x & 31) == 0) warp_counts[threadIdx.x >> 5] = count;
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int w = 0; w < 8; ++w) total += warp_counts[w];
        if (total) atomicAdd(a.sums + (long)o * SDB_NSUM + SDB_S_STRUCT, (double)total);
    }
}

}  // namespace

extern "C" int sdb_struct(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride,
                          const float* h_sites_xy, int32_t n_sites, double dist, int32_t mode, double* d_sums,
                          void* stream)
{
    return sdb_struct_launch(d_origins, n_origins, n, stride, h_sites_xy, n_sites, dist, mode, d_sums, 1, stream);
}

// clear == 0: the SDB_S_STRUCT slots have been zeroed on this stream already (dyn_step_kernel does it in
// sdb_frame_step: one launch less per frame)
int sdb_struct_launch(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride, const float* h_sites_xy,
                      int32_t n_sites, double dist, int32_t mode, double* d_sums, int32_t clear, void* stream)
{
    if (!d_origins || !h_sites_xy || !d_sums || n <= 0 || n_sites <= 0 || n_sites > 8 || mode < 0 || mode > 1 ||
        (long long)n * n_sites >= 0xffffffffLL) {
        sdb_set_error("sdb_struct: bad argument");
        return SDB_EINVAL;
    }
    if (n_origins <= 0) return SDB_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StructArgs a;
    a.origins = d_origins;
    a.sums = d_sums;
    a.n = n;
    a.stride = stride;
    a.n_sites = n_sites;
    a.mode = mode;
    // smallest double x with sqrt(x) >= dist: (norm < dist) == (sumsq < cut2), sqrt being monotone
    double cut2 = dist * dist;
    while (sqrt(cut2) >= dist) cut2 = nextafter(cut2, -INFINITY);
    while (sqrt(cut2) < dist) cut2 = nextafter(cut2, INFINITY);
    a.cut2 = cut2;
    for (int i = 0; i < n_sites; ++i) {
        a.sx[i] = h_sites_xy[2 * i];
        a.sy[i] = h_sites_xy[2 * i + 1];
    }
    if (clear) {
        int rc = sdb_check_cuda(cudaMemset2DAsync(d_sums + SDB_S_STRUCT, SDB_NSUM * sizeof(double), 0, sizeof(double),
                                                  (size_t)n_origins, s),
                                "sdb_struct: clear");
        if (rc) return rc;
    }
    // grid.x blocks per origin, each striding over the 1024-molecule tiles; sized so that the grid is a
    // whole number of waves (148 SMs x 5 resident blocks of 256 threads at 48 registers)
    // (struct3_kernel: one block iteration covers 256 groups of each of the three sites = 3072 molecules)
    const bool fast3 = n_sites == 3 && mode == 0;
    const int tiles = fast3 ? (n / 12 + 1 + 255) / 256 : (n + 1023) / 1024;
    const int base = tiles < 48 ? tiles : 48;
    const long wave = 148L * 5;
    const long waves = ((long)base * n_origins + wave - 1) / wave;
    // with few origins a grid of two waves leaves a long tail (blocks of 5 and of 6 iterations): at least
    // SDB_STRUCT_MIN_WAVES (default 8) waves of shorter blocks
    // (measured at 25 origins x 1 M molecules: 0.096 ms with 2 waves, 0.086 with 6, 0.084 with 12)
    static const long min_waves = [] { const char* e = getenv("SDB_STRUCT_MIN_WAVES"); return e ? atol(e) : 8L; }();
    long gx = std::max(waves, min_waves) * wave / n_origins;
    if (gx < 1) gx = 1;
    if (gx > tiles) gx = tiles;
    dim3 grid((unsigned)gx, n_origins);
    {
        // SDB_STRUCT_CARVEOUT=1 (experiment, off): the L1 / shared-memory split of dyn_step_kernel and the resolve
        // kernel.  Measured: no gain, struct itself 5-10 % slower
        static std::atomic<bool> configured[64];
        static const bool want = [] { const char* e = getenv("SDB_STRUCT_CARVEOUT"); return e && e[0] == '1'; }();
        int dev = 0;
        cudaGetDevice(&dev);
        if (want && dev < 64 && !configured[dev].load(std::memory_order_acquire)) {
            cudaFuncSetAttribute(struct3_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            configured[dev].store(true, std::memory_order_release);
        }
    }
    if (n_sites == 3 && mode == 0)
        struct3_kernel<<<grid, 256, 0, s>>>(a);
    else if (n_sites == 3)
        struct_kernel<3><<<grid, 256, 0, s>>>(a);
    else
        struct_kernel<0><<<grid, 256, 0, s>>>(a);
    SDB_CHECK_LAUNCH("struct_kernel");
    return SDB_OK;
}
