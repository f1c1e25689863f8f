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
// One thread per (origin, molecule j): one sincos of dtheta_j/2 (f32, like numpy on an f32 array),
// then n_sites rows in f64 in the operation order of rowan.rotate (oracle/rowan_shim.py).
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

__global__ void __launch_bounds__(256) struct_kernel(const StructArgs a)
{
    const int o = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const float* d = a.origins[o].d_delta;
    int count = 0;
    if (j < a.n) {
        const float theta = d[2 * a.stride + j];
        if (a.mode == 0) {
            // rowan.from_axis_angle: half angle, cos and sin in f32; basis axis in f64
            const float half = theta / 2.0f;
            const double c = (double)cosf(half), s = (double)sinf(half);
            for (int site = 0; site < a.n_sites; ++site) {
                const long row = (long)site * a.n + j;
                const long m = row / a.n_sites;
                const double tx = (double)d[m], ty = (double)d[a.stride + m];
                const double vx = (double)a.sx[site], vy = (double)a.sy[site];
                // inner = (0, v) (x) conj(q) = (vx s, c vx, c vy, vy s) ; outer = q (x) inner
                const double iw = __dmul_rn(vx, s), ix = __dmul_rn(c, vx), iy = __dmul_rn(c, vy), iz = __dmul_rn(vy, s);
                const double rx = __dadd_rn(__dmul_rn(c, ix), __dmul_rn(iw, s));
                const double ry = __dadd_rn(__dmul_rn(c, iy), -__dmul_rn(s, iz));
                const double rz = __dadd_rn(__dmul_rn(c, iz), __dmul_rn(s, iy));
                const double ex = __dadd_rn(__dadd_rn(rx, tx), -vx);
                const double ey = __dadd_rn(__dadd_rn(ry, ty), -vy);
                const double ss = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(rz, rz));
                count += ss < a.cut2;
            }
        } else {
            float sn, cs;
            sincosf(theta, &sn, &cs);
            const double c = cs, s = sn;
            const double tx = (double)d[j], ty = (double)d[a.stride + j];
            for (int site = 0; site < a.n_sites; ++site) {
                const double vx = (double)a.sx[site], vy = (double)a.sy[site];
                const double ex = (c * vx - s * vy) + tx - vx;
                const double ey = (s * vx + c * vy) + ty - vy;
                count += (ex * ex + ey * ey) < a.cut2;
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

}  // namespace

extern "C" int sdb_struct(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride,
                          const float* h_sites_xy, int32_t n_sites, double dist, int32_t mode, double* d_sums,
                          void* stream)
{
    if (!d_origins || !h_sites_xy || !d_sums || n <= 0 || n_sites <= 0 || n_sites > 8 || mode < 0 || mode > 1) {
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
    int rc = sdb_check_cuda(cudaMemset2DAsync(d_sums + SDB_S_STRUCT, SDB_NSUM * sizeof(double), 0, sizeof(double),
                                              (size_t)n_origins, s),
                            "sdb_struct: clear");
    if (rc) return rc;
    dim3 grid((n + 255) / 256, n_origins);
    struct_kernel<<<grid, 256, 0, s>>>(a);
    SDB_CHECK_LAUNCH("struct_kernel");
    return SDB_OK;
}
