// dyn_step.cu -- fused per-frame dynamics + first/last-passage step for all active time-origins.
//
// Replaces the body of the reference hot loop (read/_read.py:134-153): for every active origin,
// TrackedMotion.add (dynamics/_util.py:133-156), the reductions of Dynamics.compute_all
// (dynamics/dynamics.py:328-394) and Relaxations.add (dynamics/relaxations.py:135-162).
//
// Work decomposition: grid = (warp tiles / 4, segments).  A segment is a run of origins that share
// the previous sample; the minimum-image step w = wrap(prev - cur) (f32, freud op order) and the
// rotation increment alpha (f64) depend only on (segment, molecule), so each warp computes them
// once for its 256 molecules and then streams the per-origin state planes through registers:
//     dx -= wx ; dy -= wy ; dtheta = f32(f64(dtheta) + alpha)
// followed by the moment sums (fp64), the passage trackers (one flag byte per molecule and origin;
// the 4-byte passage times are only ever written, never read, while time moves forward) and the
// store of the three planes.  HBM traffic per update is 12 B read + 12 B write + 1 B flags.
#include "sdb_internal.cuh"

namespace {

struct StepArgs {
    SdbDynParams p;
    const float* pos;
    const float* quat;
    float* cur_compact;  // float[4][stride] planes (x, y, qw, qz) or nullptr
    const SdbOrigin* origins;
    const SdbSegment* segments;
    double* partials;    // double[n_origins][n_parts][SDB_NSUM]
    const double (*atan_tab)[2];
    int n_parts;
};

constexpr int NM = SDB_LANE_MOLS * SDB_GROUPS;  // molecules per lane

__global__ void __launch_bounds__(SDB_BLOCK_THREADS) dyn_step_kernel(const StepArgs a)
{
    __shared__ double red[SDB_BLOCK_WARPS][SDB_NFSUM][33];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int part = blockIdx.x * SDB_BLOCK_WARPS + warp;  // warp tile index
    const int n = a.p.n;
    const long stride = a.p.stride;
    if (part * SDB_WARP_MOLS >= n) return;

    const SdbSegment seg = a.segments[blockIdx.y];
    const float* prev = static_cast<const float*>(seg.d_prev);
    const bool zero_step = (seg.flags & SDB_SEG_ZERO_STEP) != 0 || a.pos == nullptr;

    // ---- current frame (GSD layout) for this lane's 8 molecules -------------------------------
    int j0[SDB_GROUPS];
    float cx[NM], cy[NM], cw[NM], cz[NM];
    int nonplanar = 0;
#pragma unroll
    for (int g = 0; g < SDB_GROUPS; ++g) {
        j0[g] = part * SDB_WARP_MOLS + g * (32 * SDB_LANE_MOLS) + lane * SDB_LANE_MOLS;
        if (zero_step) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                cx[4 * g + e] = 0.f; cy[4 * g + e] = 0.f; cw[4 * g + e] = 1.f; cz[4 * g + e] = 0.f;
            }
        } else if (j0[g] + SDB_LANE_MOLS <= n) {
            const float4 p0 = sdb_ldg4(a.pos + 3L * j0[g]);
            const float4 p1 = sdb_ldg4(a.pos + 3L * j0[g] + 4);
            const float4 p2 = sdb_ldg4(a.pos + 3L * j0[g] + 8);
            cx[4 * g + 0] = p0.x; cy[4 * g + 0] = p0.y;
            cx[4 * g + 1] = p0.w; cy[4 * g + 1] = p1.x;
            cx[4 * g + 2] = p1.z; cy[4 * g + 2] = p1.w;
            cx[4 * g + 3] = p2.y; cy[4 * g + 3] = p2.z;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
                if (a.quat) q = sdb_ldg4(a.quat + 4L * (j0[g] + e));
                cw[4 * g + e] = q.x; cz[4 * g + e] = q.w;
                nonplanar += (q.y != 0.f) | (q.z != 0.f);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0[g] + e;
                const bool ok = j < n;
                cx[4 * g + e] = ok ? __ldg(a.pos + 3L * j) : 0.f;
                cy[4 * g + e] = ok ? __ldg(a.pos + 3L * j + 1) : 0.f;
                float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
                if (ok && a.quat) q = sdb_ldg4(a.quat + 4L * j);
                cw[4 * g + e] = q.x; cz[4 * g + e] = q.w;
                nonplanar += (q.y != 0.f) | (q.z != 0.f);
            }
        }
    }
    if (a.cur_compact && blockIdx.y == 0 && !zero_step) {
#pragma unroll
        for (int g = 0; g < SDB_GROUPS; ++g) {
            if (j0[g] < n) {  // planes are padded to a multiple of 4: whole-vector store is safe
                float* c = a.cur_compact + j0[g];
                *reinterpret_cast<float4*>(c) = make_float4(cx[4 * g], cx[4 * g + 1], cx[4 * g + 2], cx[4 * g + 3]);
                *reinterpret_cast<float4*>(c + stride) = make_float4(cy[4 * g], cy[4 * g + 1], cy[4 * g + 2], cy[4 * g + 3]);
                *reinterpret_cast<float4*>(c + 2 * stride) = make_float4(cw[4 * g], cw[4 * g + 1], cw[4 * g + 2], cw[4 * g + 3]);
                *reinterpret_cast<float4*>(c + 3 * stride) = make_float4(cz[4 * g], cz[4 * g + 1], cz[4 * g + 2], cz[4 * g + 3]);
            }
        }
    }

    // ---- per-segment increments: w = box.wrap(prev - cur), alpha = to_euler(q / q_prev)[0] ------
    float wx[NM], wy[NM];
    double alpha[NM];
    int bad = nonplanar;
    {
        float px[NM], py[NM], pw[NM], pz[NM];
#pragma unroll
        for (int g = 0; g < SDB_GROUPS; ++g) {
            if (prev != nullptr && j0[g] < n && !zero_step) {
                const float4 vx = sdb_ldg4(prev + j0[g]);
                const float4 vy = sdb_ldg4(prev + stride + j0[g]);
                const float4 vw = sdb_ldg4(prev + 2 * stride + j0[g]);
                const float4 vz = sdb_ldg4(prev + 3 * stride + j0[g]);
                px[4 * g] = vx.x; px[4 * g + 1] = vx.y; px[4 * g + 2] = vx.z; px[4 * g + 3] = vx.w;
                py[4 * g] = vy.x; py[4 * g + 1] = vy.y; py[4 * g + 2] = vy.z; py[4 * g + 3] = vy.w;
                pw[4 * g] = vw.x; pw[4 * g + 1] = vw.y; pw[4 * g + 2] = vw.z; pw[4 * g + 3] = vw.w;
                pz[4 * g] = vz.x; pz[4 * g + 1] = vz.y; pz[4 * g + 2] = vz.z; pz[4 * g + 3] = vz.w;
            } else {  // origin created at this frame: previous sample == current sample
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    px[4 * g + e] = cx[4 * g + e]; py[4 * g + e] = cy[4 * g + e];
                    pw[4 * g + e] = cw[4 * g + e]; pz[4 * g + e] = cz[4 * g + e];
                }
            }
        }
        const bool rotate = (seg.flags & SDB_SEG_NO_ROTATION) == 0 && !zero_step;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            sdb_wrap2d(a.p.lx, a.p.ly, a.p.xy, SDB_FSUB(px[m], cx[m]), SDB_FSUB(py[m], cy[m]), wx[m], wy[m]);
            alpha[m] = 0.0;
            if (rotate) {
                int b;
                alpha[m] = sdb_rot_increment(cw[m], cz[m], pw[m], pz[m], a.atan_tab, b);
                const int g = m / SDB_LANE_MOLS, e = m % SDB_LANE_MOLS;
                bad += (b != 0) & (j0[g] + e < n);
            }
        }
    }
    const int bad_warp = __reduce_add_sync(0xffffffffu, bad);

    // ---- stream the per-origin state ----------------------------------------------------------
    for (int oi = 0; oi < seg.count; ++oi) {
        const int o = seg.first + oi;
        const SdbOrigin org = a.origins[o];
        const bool is_new = org.flags & SDB_ORIGIN_NEW;
        const bool literal = org.flags & SDB_ORIGIN_LITERAL;
        float* dbase = org.d_delta;
        uint8_t* fbase = org.d_flags;
        uint32_t* sbase = org.d_status;

        double acc[SDB_NFSUM];
#pragma unroll
        for (int v = 0; v < SDB_NFSUM; ++v) acc[v] = 0.0;
        int com_count = 0;

#pragma unroll
        for (int g = 0; g < SDB_GROUPS; ++g) {
            if (j0[g] >= n) continue;
            float4 dx = make_float4(0.f, 0.f, 0.f, 0.f), dy = dx, dt = dx;
            uint32_t fl4 = 0;
            if (!is_new) {
                dx = sdb_ld_stream4(dbase + j0[g]);
                dy = sdb_ld_stream4(dbase + stride + j0[g]);
                dt = sdb_ld_stream4(dbase + 2 * stride + j0[g]);
                if (a.p.n_trackers) fl4 = __ldcs(reinterpret_cast<const uint32_t*>(fbase + j0[g]));
            }
            float ddx[4] = {dx.x, dx.y, dx.z, dx.w};
            float ddy[4] = {dy.x, dy.y, dy.z, dy.w};
            float ddt[4] = {dt.x, dt.y, dt.z, dt.w};
            uint32_t fl4_new = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = 4 * g + e;
                const int j = j0[g] + e;
                // TrackedMotion.add, _util.py:149-153
                ddx[e] = SDB_FSUB(ddx[e], wx[m]);
                ddy[e] = SDB_FSUB(ddy[e], wy[m]);
                ddt[e] = (float)((double)ddt[e] + alpha[m]);
                // np.square(delta).sum(axis=1) / np.linalg.norm(delta, axis=1) in f32
                const float d2 = SDB_FADD(SDB_FMUL(ddx[e], ddx[e]), SDB_FMUL(ddy[e], ddy[e]));
                const float th = fabsf(ddt[e]);
                const bool live = j < n;
                if (a.p.do_sums && live) {
                    const float disp = __fsqrt_rn(d2);
                    const float t2 = SDB_FMUL(ddt[e], ddt[e]);
                    const float c = cosf(th);
                    acc[SDB_S_DISP] += (double)disp;
                    acc[SDB_S_DISP2] += (double)d2;
                    acc[SDB_S_DISP4] += (double)SDB_FMUL(d2, d2);
                    acc[SDB_S_ROT] += (double)th;
                    acc[SDB_S_ROT2] += (double)t2;
                    acc[SDB_S_COS] += (double)c;
                    acc[SDB_S_COS2] += (double)SDB_FSUB(SDB_FMUL(2.0f, SDB_FMUL(c, c)), 1.0f);
                    acc[SDB_S_ROT4] += (double)SDB_FMUL(t2, t2);
                    acc[SDB_S_D2R2] += (double)SDB_FMUL(d2, t2);
                    com_count += d2 < a.p.com_cut_lt;
                }
                // Relaxations.add, relaxations.py:152-162 with MolecularRelaxation.add :209-218 and
                // LastMolecularRelaxation.add :238-289
                uint32_t fl = (fl4 >> (8 * e)) & 0xffu;
                if (live) {
#pragma unroll
                    for (int t = 0; t < SDB_MAX_TRACKERS; ++t) {
                        if (t >= a.p.n_trackers) break;
                        const SdbTracker tr = a.p.trackers[t];
                        const float val = tr.is_rotation ? th : d2;
                        uint32_t* st_word = sbase + (long)t * stride + j;
                        if (is_new) *st_word = SDB_NO_STATUS;
                        if (!tr.is_last) {
                            if (val >= tr.cut_gt) {
                                const bool fired = (fl >> tr.bit) & 1u;
                                if (!fired || (literal && *st_word > org.timediff)) *st_word = org.timediff;
                                fl |= 1u << tr.bit;
                            }
                        } else {
                            uint32_t s = (fl >> tr.bit) & 3u;
                            if (s == 0 && val >= tr.cut_gt) {
                                s = 1;
                                *st_word = org.timediff;
                            } else if (s == 1 && val < tr.cut_lt) {
                                s = 0;
                            }
                            if (s == 1 && val >= tr.cut_irr) s = 2;
                            fl = (fl & ~(3u << tr.bit)) | (s << tr.bit);
                        }
                    }
                }
                fl4_new |= fl << (8 * e);
            }
            sdb_st_stream4(dbase + j0[g], make_float4(ddx[0], ddx[1], ddx[2], ddx[3]));
            sdb_st_stream4(dbase + stride + j0[g], make_float4(ddy[0], ddy[1], ddy[2], ddy[3]));
            sdb_st_stream4(dbase + 2 * stride + j0[g], make_float4(ddt[0], ddt[1], ddt[2], ddt[3]));
            if (a.p.n_trackers && (is_new || fl4_new != fl4))
                *reinterpret_cast<uint32_t*>(fbase + j0[g]) = fl4_new;
        }

        if (a.p.do_sums) {
            // warp reduction of the 10 fp64 sums through shared memory: lane (v, third) adds every
            // third entry of row v, two shuffles combine the thirds.
            __syncwarp();
#pragma unroll
            for (int v = 0; v < SDB_NFSUM; ++v) red[warp][v][lane] = acc[v];
            __syncwarp();
            double s = 0.0;
            const int v = lane % SDB_NFSUM, third = lane / SDB_NFSUM;
            if (lane < 3 * SDB_NFSUM) {
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                    const int idx = third + 3 * i;
                    if (idx < 32) s += red[warp][v][idx];
                }
            }
            s += __shfl_down_sync(0xffffffffu, s, SDB_NFSUM) + __shfl_down_sync(0xffffffffu, s, 2 * SDB_NFSUM);
            const int com_warp = __reduce_add_sync(0xffffffffu, com_count);
            double* out = a.partials + ((long)o * a.n_parts + part) * SDB_NSUM;
            if (lane < SDB_NFSUM) {
                out[lane] = s;
            } else if (lane == SDB_NFSUM) {
                out[SDB_S_COMSTRUCT] = (double)com_warp;
            } else if (lane == SDB_NFSUM + 1) {
                out[SDB_S_BADQUAT] = (double)bad_warp;
            }
        }
    }
}

// Sum the per-warp-tile partials of one origin: block = 256 threads = 16 sums x 16 strided readers.
__global__ void __launch_bounds__(256) dyn_finalize_kernel(const double* __restrict__ partials, int n_parts,
                                                            double* __restrict__ sums)
{
    __shared__ double sh[16][17];
    const int o = blockIdx.x;
    const int v = threadIdx.x & 15, r = threadIdx.x >> 4;
    const double* base = partials + (long)o * n_parts * SDB_NSUM;
    double s = 0.0;
    for (int p = r; p < n_parts; p += 16) s += base[(long)p * SDB_NSUM + v];
    sh[r][v] = s;
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) t += sh[i][threadIdx.x];
        const int vv = threadIdx.x;
        // the struct / overlap slots are owned by their own kernels; spare slots are zeroed
        if (vv != SDB_S_STRUCT && vv != SDB_S_OVERLAP) {
            const bool produced = vv < SDB_NFSUM || vv == SDB_S_COMSTRUCT || vv == SDB_S_BADQUAT;
            sums[(long)o * SDB_NSUM + vv] = produced ? t : 0.0;
        }
    }
}

// Pack a GSD-layout frame into (x, y, qw, qz) planes.
__global__ void pack_frame_kernel(const float* __restrict__ pos, const float* __restrict__ quat, int n, long stride,
                                  float* __restrict__ out, uint32_t* __restrict__ bad)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int nonplanar = 0;
    if (j < n) {
        float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
        if (quat) q = sdb_ldg4(quat + 4L * j);
        out[j] = pos[3L * j];
        out[stride + j] = pos[3L * j + 1];
        out[2 * stride + j] = q.x;
        out[3 * stride + j] = q.w;
        nonplanar = (q.y != 0.f) | (q.z != 0.f);
    }
    const int total = __reduce_add_sync(0xffffffffu, nonplanar);
    if (bad && total && (threadIdx.x & 31) == 0) atomicAdd(bad, (uint32_t)total);
}

__global__ void norms_kernel(const float* __restrict__ delta, int n, long stride, float* __restrict__ disp,
                             float* __restrict__ rot)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float dx = delta[j], dy = delta[stride + j];
    if (disp) disp[j] = __fsqrt_rn(SDB_FADD(SDB_FMUL(dx, dx), SDB_FMUL(dy, dy)));
    if (rot) rot[j] = fabsf(delta[2 * stride + j]);
}

}  // namespace

extern "C" int sdb_dyn_norms(const float* d_delta, int32_t n, int32_t stride, float* d_disp, float* d_rot, void* stream)
{
    if (!d_delta || n <= 0 || stride < n) {
        sdb_set_error("sdb_dyn_norms: bad argument");
        return SDB_EINVAL;
    }
    norms_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_delta, n, stride, d_disp, d_rot);
    SDB_CHECK_LAUNCH("norms_kernel");
    return SDB_OK;
}

extern "C" int32_t sdb_dyn_num_parts(int32_t n) { return (n + SDB_WARP_MOLS - 1) / SDB_WARP_MOLS; }

extern "C" int sdb_dyn_step(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, float* d_cur_compact,
                            const SdbOrigin* d_origins, int32_t n_origins, const SdbSegment* d_segments,
                            int32_t n_segments, double* d_partials, double* d_sums, void* stream)
{
    if (!h_params || !d_origins || !d_segments) {
        sdb_set_error("sdb_dyn_step: null argument");
        return SDB_EINVAL;
    }
    const SdbDynParams& p = *h_params;
    if (p.n <= 0 || p.stride < p.n || (p.stride & 3) || p.n_trackers < 0 || p.n_trackers > SDB_MAX_TRACKERS) {
        sdb_set_error("sdb_dyn_step: bad sizes n=%d stride=%d n_trackers=%d", p.n, p.stride, p.n_trackers);
        return SDB_EINVAL;
    }
    if (p.do_sums && (!d_partials || !d_sums)) {
        sdb_set_error("sdb_dyn_step: sums requested without output arrays");
        return SDB_EINVAL;
    }
    if (n_origins <= 0 || n_segments <= 0) return SDB_OK;
    StepArgs a;
    a.p = p;
    a.pos = d_pos;
    a.quat = d_quat;
    a.cur_compact = d_cur_compact;
    a.origins = d_origins;
    a.segments = d_segments;
    a.partials = d_partials;
    a.atan_tab = sdb_device_atan_table();
    a.n_parts = sdb_dyn_num_parts(p.n);
    if (!a.atan_tab) return SDB_ECUDA;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid((a.n_parts + SDB_BLOCK_WARPS - 1) / SDB_BLOCK_WARPS, n_segments);
    dyn_step_kernel<<<grid, SDB_BLOCK_THREADS, 0, s>>>(a);
    SDB_CHECK_LAUNCH("dyn_step_kernel");
    if (p.do_sums) {
        dyn_finalize_kernel<<<n_origins, 256, 0, s>>>(d_partials, a.n_parts, d_sums);
        SDB_CHECK_LAUNCH("dyn_finalize_kernel");
    }
    return SDB_OK;
}

extern "C" int sdb_pack_frame(const float* d_pos, const float* d_quat, int32_t n, int32_t stride, float* d_compact,
                              uint32_t* d_bad, void* stream)
{
    if (!d_pos || !d_compact || n <= 0 || stride < n) {
        sdb_set_error("sdb_pack_frame: bad argument");
        return SDB_EINVAL;
    }
    pack_frame_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_pos, d_quat, n, stride,
                                                                                      d_compact, d_bad);
    SDB_CHECK_LAUNCH("pack_frame_kernel");
    return SDB_OK;
}
