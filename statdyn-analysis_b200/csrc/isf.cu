// isf.cu -- intermediate scattering function, Dynamics.compute_isf (reference dynamics/dynamics.py:233-235
// with the wave vectors of :105-112):
//     mean over M angles phi_i = 2 pi i / M and over molecules j of cos(k (cos phi_i, sin phi_i) . dr_j)
// The reference evaluates an (M x 2) . (2 x N) product and M N cosines per (frame, origin); it is off by
// default (`--scattering-function`, main.py:131-137).
//
// For one molecule the angular mean is a function of x = k |dr| only up to terms J_M(x), J_2M(x), ...:
//     (1/M) sum_i cos(x cos(phi_i - psi)) = J0(x) + 2 sum_{m>=1} (-1)^{mM/2} J_{mM}(x) cos(m M psi)   (M even)
// and |J_M(x)| < 1e-30 for x <= 0.4 M, so the mean is J0(x) to double precision there (M = 360: |dr| up
// to ~50 at k = 2.9); beyond that the M cosines are evaluated literally.  One Bessel evaluation per
// molecule instead of M cosines; everything in fp64 like the reference (f64 wave vectors x f32 dr).
#include "sdb_internal.cuh"

namespace {

__device__ double isf_angular_mean(double dx, double dy, double k, int m)
{
    const double x = k * sqrt(dx * dx + dy * dy);
    if ((m & 1) == 0 && x <= 0.4 * m) return j0(x);
    double s = 0.0;
    for (int i = 0; i < m; ++i) {
        double sn, cs;
        sincospi(2.0 * i / m, &sn, &cs);
        s += cos(k * (cs * dx + sn * dy));
    }
    return s / m;
}

struct IsfArgs {
    const SdbOrigin* origins;  // multi-origin planar layout, or nullptr
    const float* dx;           // single array: x and y components, `elem_stride` floats apart per molecule
    const float* dy;
    long elem_stride;
    long stride;  // plane stride of the planar layout
    int n;
    double k;
    int m;
    double* out;  // [origins]: sum over molecules of the angular mean
};

__global__ void __launch_bounds__(256) isf_kernel(const IsfArgs a)
{
    const int o = blockIdx.y;
    const float* px = a.origins ? a.origins[o].d_delta : a.dx;
    const float* py = a.origins ? a.origins[o].d_delta + a.stride : a.dy;
    const long es = a.origins ? 1 : a.elem_stride;
    double s = 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += gridDim.x * blockDim.x)
        s += isf_angular_mean((double)px[j * es], (double)py[j * es], a.k, a.m);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    __shared__ double warp_sums[8];
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += warp_sums[w];
        atomicAdd(a.out + o, t);
    }
}

int launch_isf(IsfArgs a, int n_origins, cudaStream_t s)
{
    int rc = sdb_check_cuda(cudaMemsetAsync(a.out, 0, sizeof(double) * (size_t)n_origins, s), "sdb_isf: clear");
    if (rc) return rc;
    const int tiles = (a.n + 255) / 256;
    isf_kernel<<<dim3(tiles < 296 ? tiles : 296, n_origins), 256, 0, s>>>(a);
    SDB_CHECK_LAUNCH("isf_kernel");
    return SDB_OK;
}

}  // namespace

extern "C" int sdb_isf_origins(const SdbOrigin* d_origins, int32_t n_origins, int32_t n, int32_t stride,
                               double wave_number, int32_t angular_resolution, double* d_out, void* stream)
{
    if (!d_origins || !d_out || n <= 0 || stride < n || angular_resolution <= 0) {
        sdb_set_error("sdb_isf_origins: bad argument");
        return SDB_EINVAL;
    }
    if (n_origins <= 0) return SDB_OK;
    IsfArgs a{};
    a.origins = d_origins;
    a.stride = stride;
    a.n = n;
    a.k = wave_number;
    a.m = angular_resolution;
    a.out = d_out;
    return launch_isf(a, n_origins, static_cast<cudaStream_t>(stream));
}

extern "C" int sdb_isf(const float* d_dx, const float* d_dy, int32_t elem_stride, int32_t n, double wave_number,
                       int32_t angular_resolution, double* d_out, void* stream)
{
    if (!d_dx || !d_dy || !d_out || n <= 0 || elem_stride <= 0 || angular_resolution <= 0) {
        sdb_set_error("sdb_isf: bad argument");
        return SDB_EINVAL;
    }
    IsfArgs a{};
    a.dx = d_dx;
    a.dy = d_dy;
    a.elem_stride = elem_stride;
    a.n = n;
    a.k = wave_number;
    a.m = angular_resolution;
    a.out = d_out;
    return launch_isf(a, 1, static_cast<cudaStream_t>(stream));
}
