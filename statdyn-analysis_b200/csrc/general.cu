// general.cu -- general (3D box / arbitrary quaternion) tracked motion and its reductions.
//
// Not the hot path: the reference's data are 2D Trimers, handled by dyn_step.cu.  These kernels
// keep the drop-in API complete for the cases the reference tests exercise (3D boxes,
// rotations about all three axes: test/tracked_motion_test.py, test/dynamics_test.py:373-393) with
// state in the reference's own (N,3) layout.  One origin per call.
//
// The f64 sequence follows oracle/rowan_shim.py (rowan.divide -> rowan.to_euler) operation by
// operation; atan2/asin/cos come from the CUDA math library (<= 2 ulp).
#include "sdb_internal.cuh"

namespace {

using Box3 = SdbBox3;
#define wrap3 sdb_wrap3

__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }

// rowan.to_euler(rowan.divide(q, qp)) -> (alpha, beta, gamma); bad when rowan would raise
__device__ void euler_of_quotient(const float4 q, const float4 qp, double out[3], int& bad)
{
    // inverse(qp): f64 copy, conj / |qp|^2 with |qp|^2 = (sqrt(sum sq))^2
    const double pw = qp.x, px = qp.y, py = qp.z, pz = qp.w;
    const double nrm = sqrt(da(da(da(dm(pw, pw), dm(px, px)), dm(py, py)), dm(pz, pz)));
    const double nsq = dm(nrm, nrm);
    double iw = pw, ix = -px, iy = -py, iz = -pz;
    if (nsq > 0.0) {
        iw = iw / nsq; ix = ix / nsq; iy = iy / nsq; iz = iz / nsq;
    }
    // multiply(q, inverse)
    const double w = q.x, x = q.y, y = q.z, z = q.w;
    const double rw = da(dm(w, iw), -da(da(dm(x, ix), dm(y, iy)), dm(z, iz)));
    const double rx = da(da(dm(w, ix), dm(iw, x)), da(dm(y, iz), -dm(z, iy)));
    const double ry = da(da(dm(w, iy), dm(iw, y)), da(dm(z, ix), -dm(x, iz)));
    const double rz = da(da(dm(w, iz), dm(iw, z)), da(dm(x, iy), -dm(y, ix)));
    // to_matrix: s = 1 / |r|
    const double rn = sqrt(da(da(da(dm(rw, rw), dm(rx, rx)), dm(ry, ry)), dm(rz, rz)));
    bad = !(fabs(rn - 1.0) <= 1e-8 + 1e-5) || !(rn > 0.0);
    const double s2 = dm(2.0, 1.0 / rn);
    auto clip = [](double v) { return fmin(fmax(v, -1.0), 1.0); };
    const double m00 = clip(da(1.0, -dm(s2, da(dm(ry, ry), dm(rz, rz)))));
    const double m01 = clip(dm(s2, da(dm(rx, ry), -dm(rz, rw))));
    const double m10 = clip(dm(s2, da(dm(rx, ry), dm(rz, rw))));
    const double m11 = clip(da(1.0, -dm(s2, da(dm(rx, rx), dm(rz, rz)))));
    const double m20 = clip(dm(s2, da(dm(rx, rz), -dm(ry, rw))));
    const double m21 = clip(dm(s2, da(dm(ry, rz), dm(rx, rw))));
    const double m22 = clip(da(1.0, -dm(s2, da(dm(rx, rx), dm(ry, ry)))));
    const double beta = asin(-m20);
    const bool gimbal = fabs(cos(beta)) <= 1e-3;
    out[0] = gimbal ? atan2(-m01, m11) : atan2(m10, m00);
    out[1] = beta;
    out[2] = gimbal ? 0.0 : atan2(m21, m22);
}

__global__ void motion_general_kernel(const Box3 box, int n, const float* __restrict__ pos, const float* __restrict__ quat,
                                      float* __restrict__ prev_pos, float* __restrict__ prev_quat,
                                      float* __restrict__ dtrans, float* __restrict__ drot, int rotate,
                                      uint32_t* __restrict__ bad_out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (j < n) {
        float c[3], w[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) c[a] = pos[3L * j + a];
        wrap3(box, __fsub_rn(prev_pos[3L * j], c[0]), __fsub_rn(prev_pos[3L * j + 1], c[1]),
              __fsub_rn(prev_pos[3L * j + 2], c[2]), w[0], w[1], w[2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            dtrans[3L * j + a] = __fsub_rn(dtrans[3L * j + a], w[a]);
            prev_pos[3L * j + a] = c[a];
        }
        if (quat != nullptr) {
            const float4 q = *reinterpret_cast<const float4*>(quat + 4L * j);
            if (rotate) {
                const float4 qp = *reinterpret_cast<const float4*>(prev_quat + 4L * j);
                double e[3];
                euler_of_quotient(q, qp, e, bad);
#pragma unroll
                for (int a = 0; a < 3; ++a) drot[3L * j + a] = (float)((double)drot[3L * j + a] + e[a]);
            }
            *reinterpret_cast<float4*>(prev_quat + 4L * j) = q;
        }
    }
    const int total = __reduce_add_sync(0xffffffffu, bad);
    if (bad_out && total && (threadIdx.x & 31) == 0) atomicAdd(bad_out, (uint32_t)total);
}

// Reductions of Dynamics.compute_* on (N,3) accumulators; also emits |dr| and |dtheta| per molecule.
__global__ void __launch_bounds__(256) sums_general_kernel(int n, const float* __restrict__ dtrans,
                                                            const float* __restrict__ drot, float com_thr,
                                                            float* __restrict__ disp_out, float* __restrict__ rot_out,
                                                            double* __restrict__ partials)
{
    __shared__ double sh[8][SDB_NSUM];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[SDB_NSUM];
#pragma unroll
    for (int v = 0; v < SDB_NSUM; ++v) acc[v] = 0.0;
    if (j < n) {
        const float x = dtrans[3L * j], y = dtrans[3L * j + 1], z = dtrans[3L * j + 2];
        const float a = drot[3L * j], b = drot[3L * j + 1], c = drot[3L * j + 2];
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        const float t2 = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
        const float disp = __fsqrt_rn(d2), th = __fsqrt_rn(t2);
        if (disp_out) disp_out[j] = disp;
        if (rot_out) rot_out[j] = th;
        const float cs = cosf(th);
        acc[SDB_S_DISP] = disp;
        acc[SDB_S_DISP2] = d2;
        acc[SDB_S_DISP4] = __fmul_rn(d2, d2);
        acc[SDB_S_COMSTRUCT] = disp < com_thr;
        acc[SDB_S_ROT] = th;
        acc[SDB_S_ROT2] = t2;
        acc[SDB_S_COS] = cs;
        acc[SDB_S_COS2] = __fsub_rn(__fmul_rn(2.0f, __fmul_rn(cs, cs)), 1.0f);
        acc[SDB_S_ROT4] = __fmul_rn(t2, t2);
        acc[SDB_S_D2R2] = __fmul_rn(d2, t2);
    }
#pragma unroll
    for (int v = 0; v < SDB_NSUM; ++v) {
        double s = acc[v];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][v] = s;
    }
    __syncthreads();
    if (threadIdx.x < SDB_NSUM) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
        partials[(long)blockIdx.x * SDB_NSUM + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) sums_general_finalize(const double* __restrict__ partials, int n_parts,
                                                              double* __restrict__ sums)
{
    __shared__ double sh[16][17];
    const int v = threadIdx.x & 15, r = threadIdx.x >> 4;
    double s = 0.0;
    for (int p = r; p < n_parts; p += 16) s += partials[(long)p * SDB_NSUM + v];
    sh[r][v] = s;
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) t += sh[i][threadIdx.x];
        sums[threadIdx.x] = t;
    }
}

}  // namespace

extern "C" int sdb_motion_general(const float* h_box, int32_t is2d, int32_t n, const float* d_pos, const float* d_quat,
                                  float* d_prev_pos, float* d_prev_quat, int32_t rotate, float* d_dtrans, float* d_drot,
                                  uint32_t* d_bad, void* stream)
{
    if (!h_box || !d_pos || !d_prev_pos || !d_dtrans || !d_drot || n <= 0 || (d_quat && !d_prev_quat)) {
        sdb_set_error("sdb_motion_general: bad argument");
        return SDB_EINVAL;
    }
    Box3 b;
    b.lx = h_box[0]; b.ly = h_box[1]; b.lz = is2d ? 0.0f : h_box[2];
    b.xy = h_box[3]; b.xz = is2d ? 0.0f : h_box[4]; b.yz = is2d ? 0.0f : h_box[5];
    b.is2d = is2d;
    motion_general_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        b, n, d_pos, d_quat, d_prev_pos, d_prev_quat, d_dtrans, d_drot, rotate, d_bad);
    SDB_CHECK_LAUNCH("motion_general_kernel");
    return SDB_OK;
}

extern "C" int32_t sdb_sums_general_parts(int32_t n) { return (n + 255) / 256; }

extern "C" int sdb_sums_general(int32_t n, const float* d_dtrans, const float* d_drot, float com_threshold,
                                float* d_disp, float* d_rot, double* d_partials, double* d_sums, void* stream)
{
    if (!d_dtrans || !d_drot || !d_partials || !d_sums || n <= 0) {
        sdb_set_error("sdb_sums_general: bad argument");
        return SDB_EINVAL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int parts = sdb_sums_general_parts(n);
    sums_general_kernel<<<parts, 256, 0, s>>>(n, d_dtrans, d_drot, com_threshold, d_disp, d_rot, d_partials);
    SDB_CHECK_LAUNCH("sums_general_kernel");
    sums_general_finalize<<<1, 256, 0, s>>>(d_partials, parts, d_sums);
    SDB_CHECK_LAUNCH("sums_general_finalize");
    return SDB_OK;
}
