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
    double* sums;        // double[n_origins][SDB_NSUM] or nullptr: the warp of tile 0 clears the `struct` slot of every
                         // origin it passes (the struct kernel then needs no separate clear)
    const double (*atan_tab)[2];
    int n_parts;
    int full_parts;  // number of complete 256-molecule tiles (a partial last tile may follow)
    int ov_on;  // bracketed overlap selection: count keys above the brackets, collect candidates
    SdbOvWorkspace ov;
    // per-frame precomputed increments (dyn_increments_kernel), one set per distinct previous sample:
    // wx, wy float[sets][stride], alpha double[sets][stride], bad uint32[sets][n_parts]; null: none
    const float* inc_wx;
    const float* inc_wy;
    const double* inc_al;
    const uint32_t* inc_bad;
};

// Layout of the increments workspace.
struct IncLayout {
    size_t off_wy, off_al, off_bad, bytes;
};
inline IncLayout sdb_inc_layout(int n, int stride, int n_sets)
{
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t n_parts = (size_t)((n + SDB_WARP_MOLS - 1) / SDB_WARP_MOLS);
    IncLayout l;
    l.off_wy = align(sizeof(float) * (size_t)stride * n_sets);
    l.off_al = l.off_wy + align(sizeof(float) * (size_t)stride * n_sets);
    l.off_bad = l.off_al + align(sizeof(double) * (size_t)stride * n_sets);
    l.bytes = l.off_bad + align(sizeof(uint32_t) * n_parts * n_sets);
    return l;
}


// cos(x) for x >= 0 with ~1e-7 absolute error: two-constant Cody-Waite reduction to [-pi, pi], then
// cos r = 1 - 2 sin^2(r/2) with an odd polynomial for sin on [-pi/2, pi/2].  Used for the rot1/rot2
// means only (rtol 1e-5 on a mean of N terms; the special-function unit's cos.approx is biased by
// ~5e-8, which shows in rot2 when it is close to zero).
__device__ __forceinline__ float sdb_cos_reduced(float x)
{
    // |x| up to ~1e5 rad keeps the stated accuracy (k * 6.28125 is exact for k < 2^16)
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
    return fmaf(-2.0f * sn, sn, 1.0f);
}

// The same polynomial on two molecules at once with Blackwell's packed f32x2 arithmetic
// (FMUL2 / FFMA2: one issue slot for both lanes of the pair).
__device__ __forceinline__ float2 sdb_cos_reduced2(float2 x)
{
    const float2 q = __fmul2_rn(x, make_float2(0.15915494309189535f, 0.15915494309189535f));
    const float2 k = make_float2(rintf(q.x), rintf(q.y));
    float2 r = __ffma2_rn(k, make_float2(-6.28125f, -6.28125f), x);
    r = __ffma2_rn(k, make_float2(-1.9353071795864769e-3f, -1.9353071795864769e-3f), r);
    const float2 h = __fmul2_rn(r, make_float2(0.5f, 0.5f)), h2 = __fmul2_rn(h, h);
    float2 p = __ffma2_rn(h2, make_float2(1.6059043836821613e-10f, 1.6059043836821613e-10f),
                          make_float2(-2.5052108385441720e-8f, -2.5052108385441720e-8f));
    p = __ffma2_rn(h2, p, make_float2(2.7557319223985893e-6f, 2.7557319223985893e-6f));
    p = __ffma2_rn(h2, p, make_float2(-1.9841269841269841e-4f, -1.9841269841269841e-4f));
    p = __ffma2_rn(h2, p, make_float2(8.3333333333333332e-3f, 8.3333333333333332e-3f));
    p = __ffma2_rn(h2, p, make_float2(-1.6666666666666666e-1f, -1.6666666666666666e-1f));
    const float2 sn = __ffma2_rn(__fmul2_rn(h, h2), p, h);
    return __ffma2_rn(__fmul2_rn(sn, make_float2(-2.0f, -2.0f)), sn, make_float2(1.0f, 1.0f));
}

__device__ __forceinline__ float sdb_sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Slow path of the passage trackers (something may change for this molecule):
// MolecularRelaxation.add relaxations.py:209-218, LastMolecularRelaxation.add :238-289.
// Not inlined: it runs for the few molecules that are near a threshold; `tr` points into the kernel
// parameter block (__grid_constant__).
__device__ __noinline__ uint32_t sdb_track(const SdbTracker* __restrict__ tr, int n_trackers, uint32_t fl, float d2,
                                           float th, uint32_t* status, long stride, uint32_t timediff, bool literal)
{
#pragma unroll 1
    for (int t = 0; t < n_trackers; ++t, ++tr) {
        const float val = tr->is_rotation ? th : d2;
        uint32_t* st_word = status + (long)t * stride;
        if (!tr->is_last) {
            if (val >= tr->cut_gt) {
                const bool fired = (fl >> tr->bit) & 1u;
                if (!fired || (literal && *st_word > timediff)) *st_word = timediff;
                fl |= 1u << tr->bit;
            }
        } else {
            uint32_t s = (fl >> tr->bit) & 3u;
            if (s == 0 && val >= tr->cut_gt) {
                s = 1;
                *st_word = timediff;
            } else if (s == 1 && val < tr->cut_lt) {
                s = 0;
            }
            if (s == 1 && val >= tr->cut_irr) s = 2;
            fl = (fl & ~(3u << tr->bit)) | (s << tr->bit);
        }
    }
    return fl;
}

// ---- bulk asynchronous copies (TMA, 1-D) completing on a shared-memory mbarrier ----------------------
__device__ __forceinline__ uint32_t sdb_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sdb_mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sdb_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sdb_mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "SDB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra SDB_DONE;\n"
        "bra SDB_WAIT;\n"
        "SDB_DONE:\n"
        "}" ::"r"(sdb_smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// Per-warp shared memory: a ring of state tiles filled by bulk copies (one origin per stage), the
// reduction scratch and the staged overlap candidates.  Warps never synchronise with each other.
#ifndef SDB_STAGES_OVERRIDE
#define SDB_STAGES_OVERRIDE 3
#endif
#ifndef SDB_MIN_BLOCKS
#define SDB_MIN_BLOCKS 4
#endif
constexpr int SDB_STAGES = SDB_STAGES_OVERRIDE;  // (build-time knobs for experiments: -DSDB_STAGES_OVERRIDE=2 -DSDB_MIN_BLOCKS=5)
constexpr int SDB_PLANE_BYTES = SDB_WARP_MOLS * 4;                       // 1 KB: one f32 plane of a tile
constexpr int SDB_STAGE_BYTES = 3 * SDB_PLANE_BYTES + SDB_WARP_MOLS;     // dx, dy, dtheta planes + flag bytes
struct alignas(128) StepWarpShared {
    alignas(128) unsigned char stage[SDB_STAGES][SDB_STAGE_BYTES];
    float red[SDB_NFSUM][33];
    uint32_t cand[3][SDB_OV_STAGE];  // staged overlap candidates (idx, d2, |dtheta|)
    // descriptors of the next origins (SdbOrigin + overlap bracket), fetched with cp.async three
    // iterations ahead: with ~220 KB of shared memory per SM there is hardly any L1 left, so a plain load
    // of a descriptor costs an L2 round trip at the top of every iteration
    alignas(16) uint32_t desc[4][12];
    alignas(8) uint64_t bar[SDB_STAGES];
};

// Per flag byte: the interval of each key in which NO tracker of the molecule changes state,
//   quiet  <=>  lo_d2 <= d2 < hi_d2  &&  lo_th <= |dtheta| < hi_th
// (first passage, not fired: key < cut_gt; last passage: state 0 key < cut_gt, state 1 cut_lt <= key <
// cut_irr, state 2 anything).  Built once per CTA; lets molecules that sit between two thresholds for
// many frames skip the tracker slow path.
__device__ __forceinline__ float4 sdb_quiet_interval(const SdbTracker* tr, int n_trackers, uint32_t fl)
{
    float4 q = make_float4(INFINITY, 0.0f, INFINITY, 0.0f);  // hi_d2, lo_d2, hi_th, lo_th
    for (int t = 0; t < n_trackers; ++t) {
        float hi = INFINITY, lo = 0.0f;
        if (!tr[t].is_last) {
            if (!((fl >> tr[t].bit) & 1u)) hi = tr[t].cut_gt;
        } else {
            const uint32_t st = (fl >> tr[t].bit) & 3u;
            if (st == 0u) hi = tr[t].cut_gt;
            else if (st == 1u) { hi = tr[t].cut_irr; lo = tr[t].cut_lt; }
        }
        if (tr[t].is_rotation) { q.z = fminf(q.z, hi); q.w = fmaxf(q.w, lo); }
        else { q.x = fminf(q.x, hi); q.y = fmaxf(q.y, lo); }
    }
    return q;
}

// Phase 1 of a warp tile: the increments of this lane's 8 molecules for one previous sample.
//   w = box.wrap(prev - cur)  (f32, freud operation order)     _util.py:149
//   alpha = to_euler(q / q_prev)[0]  (f64)                     _util.py:150-153
// Returns the lane's count of quaternion pairs rowan would reject (or that are not planar).
template <bool TAIL>
__device__ __forceinline__ int sdb_tile_increments(const StepArgs& a, const float* prev, uint32_t seg_flags, bool zero_step,
                                                   bool write_compact, int part, int lane, int (&j0)[SDB_GROUPS],
                                                   float (&wx)[SDB_GROUPS][4], float (&wy)[SDB_GROUPS][4],
                                                   double (&al)[SDB_GROUPS][4])
{
    int bad = 0;
    const int n = a.p.n;
    const long stride = a.p.stride;
    const bool rotate = (seg_flags & SDB_SEG_NO_ROTATION) == 0 && !zero_step;
#pragma unroll
    for (int g = 0; g < SDB_GROUPS; ++g) {
        const int jg = part * SDB_WARP_MOLS + g * (32 * SDB_LANE_MOLS) + lane * SDB_LANE_MOLS;
        j0[g] = jg;
        float cx[4], cy[4], cw[4], cz[4];
        if (zero_step || jg >= n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                cx[e] = 0.f; cy[e] = 0.f; cw[e] = 1.f; cz[e] = 0.f;
            }
        } else if (!TAIL || jg + SDB_LANE_MOLS <= n) {
            const float4 p0 = sdb_ldg4(a.pos + 3L * jg);
            const float4 p1 = sdb_ldg4(a.pos + 3L * jg + 4);
            const float4 p2 = sdb_ldg4(a.pos + 3L * jg + 8);
            cx[0] = p0.x; cy[0] = p0.y;
            cx[1] = p0.w; cy[1] = p1.x;
            cx[2] = p1.z; cy[2] = p1.w;
            cx[3] = p2.y; cy[3] = p2.z;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
                if (a.quat) q = sdb_ldg4(a.quat + 4L * (jg + e));
                cw[e] = q.x; cz[e] = q.w;
                bad += (q.y != 0.f) | (q.z != 0.f);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = jg + e;
                const bool ok = j < n;
                cx[e] = ok ? __ldg(a.pos + 3L * j) : 0.f;
                cy[e] = ok ? __ldg(a.pos + 3L * j + 1) : 0.f;
                float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
                if (ok && a.quat) q = sdb_ldg4(a.quat + 4L * j);
                cw[e] = q.x; cz[e] = q.w;
                bad += (q.y != 0.f) | (q.z != 0.f);
            }
        }
        if (a.cur_compact && write_compact && !zero_step && jg < n) {
            // planes are padded to a multiple of 4: whole-vector stores are safe
            float* c = a.cur_compact + jg;
            *reinterpret_cast<float4*>(c) = make_float4(cx[0], cx[1], cx[2], cx[3]);
            *reinterpret_cast<float4*>(c + stride) = make_float4(cy[0], cy[1], cy[2], cy[3]);
            *reinterpret_cast<float4*>(c + 2 * stride) = make_float4(cw[0], cw[1], cw[2], cw[3]);
            *reinterpret_cast<float4*>(c + 3 * stride) = make_float4(cz[0], cz[1], cz[2], cz[3]);
        }
        float px[4], py[4], pw[4], pz[4];
        if (prev != nullptr && jg < n && !zero_step) {
            const float4 vx = sdb_ldg4(prev + jg);
            const float4 vy = sdb_ldg4(prev + stride + jg);
            const float4 vw = sdb_ldg4(prev + 2 * stride + jg);
            const float4 vz = sdb_ldg4(prev + 3 * stride + jg);
            px[0] = vx.x; px[1] = vx.y; px[2] = vx.z; px[3] = vx.w;
            py[0] = vy.x; py[1] = vy.y; py[2] = vy.z; py[3] = vy.w;
            pw[0] = vw.x; pw[1] = vw.y; pw[2] = vw.z; pw[3] = vw.w;
            pz[0] = vz.x; pz[1] = vz.y; pz[2] = vz.z; pz[3] = vz.w;
        } else {  // origin created at this frame: previous sample == current sample
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                px[e] = cx[e]; py[e] = cy[e]; pw[e] = cw[e]; pz[e] = cz[e];
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sdb_wrap2d(a.p.lx, a.p.ly, a.p.xy, SDB_FSUB(px[e], cx[e]), SDB_FSUB(py[e], cy[e]), wx[g][e], wy[g][e]);
            al[g][e] = 0.0;
            if (rotate) {
                int b;
                al[g][e] = sdb_rot_increment(cw[e], cz[e], pw[e], pz[e], a.atan_tab, b);
                bad += (b != 0) & (jg + e < n);
            }
        }
    }
    return bad;
}

// One warp tile (256 molecules) of one segment.  TAIL: the tile is the partial last one (bounds checks).
template <bool SUMS, bool OV, bool TRACK, bool PRE, bool TAIL>
__device__ __forceinline__ void dyn_step_tile(const StepArgs& a, StepWarpShared& sh, const float4* quiet_tab,
                                              const int part)
{
    const int lane = threadIdx.x & 31;
    const int n = a.p.n;
    const long stride = a.p.stride;

    const SdbSegment seg = a.segments[blockIdx.y];
    const float* prev = static_cast<const float*>(seg.d_prev);
    const bool zero_step = (seg.flags & SDB_SEG_ZERO_STEP) != 0 || a.pos == nullptr;
    const long tile0 = (long)part * SDB_WARP_MOLS;
    // rows of this tile inside the padded planes (a multiple of 4)
    const uint32_t plane_bytes = 4u * (uint32_t)min((long)SDB_WARP_MOLS, stride - tile0);
    const uint32_t flag_bytes = (TRACK && !TAIL) ? (uint32_t)SDB_WARP_MOLS : 0u;

    // ---- start the pipeline: bulk copies of the first origins' state tiles -----------------------
    // Executed by the whole (converged) warp with warp-uniform operands; one elected lane issues.
    auto issue = [&](int oi, int slot) {
        const SdbOrigin* od = reinterpret_cast<const SdbOrigin*>(sh.desc[oi & 3]);
        const float* dsrc = od->d_delta + tile0;
        const uint32_t bar = sdb_smem_addr(&sh.bar[slot]);
        const uint32_t dst = sdb_smem_addr(sh.stage[slot]);
        if (TRACK && !TAIL) {
            const uint8_t* fsrc = od->d_flags + tile0;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "elect.sync _|p, 0xffffffff;\n"
                "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%6], %10, [%0];\n"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%3], [%7], %10, [%0];\n"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%4], [%8], %10, [%0];\n"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%5], [%9], %11, [%0];\n"
                "}" ::"r"(bar),
                "r"(3u * plane_bytes + flag_bytes), "r"(dst), "r"(dst + SDB_PLANE_BYTES), "r"(dst + 2 * SDB_PLANE_BYTES),
                "r"(dst + 3 * SDB_PLANE_BYTES), "l"(dsrc), "l"(dsrc + stride), "l"(dsrc + 2 * stride), "l"(fsrc),
                "r"(plane_bytes), "r"(flag_bytes)
                : "memory");
        } else {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "elect.sync _|p, 0xffffffff;\n"
                "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%5], %8, [%0];\n"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%3], [%6], %8, [%0];\n"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%4], [%7], %8, [%0];\n"
                "}" ::"r"(bar),
                "r"(3u * plane_bytes), "r"(dst), "r"(dst + SDB_PLANE_BYTES), "r"(dst + 2 * SDB_PLANE_BYTES), "l"(dsrc),
                "l"(dsrc + stride), "l"(dsrc + 2 * stride), "r"(plane_bytes)
                : "memory");
        }
    };
    // lanes 0, 1: the two halves of the SdbOrigin record; lane 2: the overlap bracket of that origin
    auto fetch_desc = [&](int oi) {
        if (lane < (OV && SUMS ? 3 : 2)) {
            const char* src = lane < 2 ? reinterpret_cast<const char*>(a.origins + seg.first + oi) + 16 * lane
                                       : reinterpret_cast<const char*>(a.ov.brackets + 4L * (seg.first + oi));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sdb_smem_addr(&sh.desc[oi & 3][4 * lane])), "l"(src)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SDB_STAGES; ++s) sdb_mbar_init(&sh.bar[s], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
#pragma unroll
    for (int s = 0; s < SDB_STAGES; ++s)
        if (s < seg.count) fetch_desc(s);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int s = 0; s < SDB_STAGES - 1; ++s)
        if (s < seg.count) issue(s, s);

    // ---- phase 1: per-segment increments of this lane's 8 molecules, kept in registers ----------
    int j0[SDB_GROUPS];
    float wx[SDB_GROUPS][4], wy[SDB_GROUPS][4];
    double al[SDB_GROUPS][4];
    int bad_warp;
    if (PRE && seg.reserved != 0u && !zero_step) {
        // the increments of this previous sample were computed once for the whole frame
        const long set_off = (long)(seg.reserved - 1u) * stride;
#pragma unroll
        for (int g = 0; g < SDB_GROUPS; ++g) {
            const int jg = part * SDB_WARP_MOLS + g * (32 * SDB_LANE_MOLS) + lane * SDB_LANE_MOLS;
            j0[g] = jg;
            float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vy = vx;
            double2 a0 = make_double2(0.0, 0.0), a1 = a0;
            if (jg < n) {
                vx = sdb_ldg4(a.inc_wx + set_off + jg);
                vy = sdb_ldg4(a.inc_wy + set_off + jg);
                a0 = __ldg(reinterpret_cast<const double2*>(a.inc_al + set_off + jg));
                a1 = __ldg(reinterpret_cast<const double2*>(a.inc_al + set_off + jg + 2));
            }
            wx[g][0] = vx.x; wx[g][1] = vx.y; wx[g][2] = vx.z; wx[g][3] = vx.w;
            wy[g][0] = vy.x; wy[g][1] = vy.y; wy[g][2] = vy.z; wy[g][3] = vy.w;
            al[g][0] = a0.x; al[g][1] = a0.y; al[g][2] = a1.x; al[g][3] = a1.y;
        }
        bad_warp = (int)__ldg(a.inc_bad + (long)(seg.reserved - 1u) * a.n_parts + part);
    } else {
        const int bad = sdb_tile_increments<TAIL>(a, prev, seg.flags, zero_step, blockIdx.y == 0, part, lane, j0, wx, wy, al);
        bad_warp = __reduce_add_sync(0xffffffffu, bad);
    }

    // ---- tracker pre-digestion (uniform across the grid) ------------------------------------------
    // A group of 4 molecules cannot change its flag bytes when (a) all four are below every "greater
    // than" cut and no last-passage tracker is in state 1, or (b) every tracker of all four has
    // reached its final state.
    float min_cut_d2 = INFINITY, min_cut_th = INFINITY;
    uint32_t state1_bits = 0, done_pattern = 0;
    for (int t = 0; t < a.p.n_trackers; ++t) {
        const SdbTracker tr = a.p.trackers[t];
        if (tr.is_rotation) min_cut_th = fminf(min_cut_th, tr.cut_gt);
        else min_cut_d2 = fminf(min_cut_d2, tr.cut_gt);
        if (tr.is_last) {
            state1_bits |= 1u << tr.bit;
            done_pattern |= 2u << tr.bit;
        } else {
            done_pattern |= 1u << tr.bit;
        }
    }
    const uint32_t state1_x4 = state1_bits * 0x01010101u, done_x4 = done_pattern * 0x01010101u;
    constexpr bool track = TRACK, do_sums = SUMS, ov = OV && SUMS;
    int live_count = 0;
#pragma unroll
    for (int g = 0; g < SDB_GROUPS; ++g) live_count += max(0, min(SDB_LANE_MOLS, n - j0[g]));
    const int live_warp = __reduce_add_sync(0xffffffffu, live_count);
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    const float com_cut = a.p.com_cut_lt;

    // ---- stream the per-origin state ----------------------------------------------------------
    int slot = 0;            // stage of origin oi
    uint32_t parity = 0;     // phase parity of that stage's barrier
    int pf_slot = SDB_STAGES - 1;  // stage the next prefetch goes to
    for (int oi = 0; oi < seg.count; ++oi) {
        const int o = seg.first + oi;
        // the stage consumed in the previous iteration is free: prefetch origin oi + STAGES - 1 into it
        // (its descriptor was fetched during the previous iteration), then fetch the next descriptor
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (oi + SDB_STAGES - 1 < seg.count) issue(oi + SDB_STAGES - 1, pf_slot);
        pf_slot = pf_slot + 1 == SDB_STAGES ? 0 : pf_slot + 1;
        if (oi + SDB_STAGES < seg.count) fetch_desc(oi + SDB_STAGES);
        const SdbOrigin org = *reinterpret_cast<const SdbOrigin*>(sh.desc[oi & 3]);
        float* dbase = org.d_delta;
        uint8_t* fbase = org.d_flags;
        uint32_t* sbase = org.d_status;
        double* out = a.partials + ((long)o * a.n_parts + part) * SDB_NSUM;
        if (SUMS && part == 0 && lane == 0 && a.sums) a.sums[(long)o * SDB_NSUM + SDB_S_STRUCT] = 0.0;
        const unsigned char* st = sh.stage[slot];
        sdb_mbar_wait(&sh.bar[slot], parity);
        slot = slot + 1 == SDB_STAGES ? 0 : slot + 1;
        parity ^= (slot == 0);

        if (org.flags & SDB_ORIGIN_NEW) {
            // Dynamics.from_frame / Relaxations.from_frame followed by the first compute_all / add with
            // the same frame: zero motion, nothing relaxed (read/_read.py:135-153).
#pragma unroll
            for (int g = 0; g < SDB_GROUPS; ++g) {
                if (j0[g] >= n) continue;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                sdb_st_stream4(dbase + j0[g], zero);
                sdb_st_stream4(dbase + stride + j0[g], zero);
                sdb_st_stream4(dbase + 2 * stride + j0[g], zero);
                if (track) {
                    *reinterpret_cast<uint32_t*>(fbase + j0[g]) = 0u;
                    for (int t = 0; t < a.p.n_trackers; ++t)
                        *reinterpret_cast<uint4*>(sbase + (long)t * stride + j0[g]) =
                            make_uint4(SDB_NO_STATUS, SDB_NO_STATUS, SDB_NO_STATUS, SDB_NO_STATUS);
                }
            }
            if (do_sums && lane < SDB_NSUM) {
                double v = 0.0;  // all sums vanish except cos(0) = 1 terms and the |dr| < d count
                if (lane == SDB_S_COS || lane == SDB_S_COS2) v = (double)live_warp;
                if (lane == SDB_S_COMSTRUCT && com_cut > 0.f) v = (double)live_warp;
                if (lane == SDB_S_BADQUAT) v = (double)bad_warp;
                if (lane < SDB_NFSUM || lane == SDB_S_COMSTRUCT || lane == SDB_S_BADQUAT || lane >= SDB_S_OV_RABOVE)
                    out[lane] = v;
            }
            __syncwarp();
            continue;
        }
        const bool literal = org.flags & SDB_ORIGIN_LITERAL;

        float acc[SDB_NFSUM];
        float2 acc2[SDB_NFSUM];  // full tiles: one partial sum per half of the molecule pairs
#pragma unroll
        for (int v = 0; v < SDB_NFSUM; ++v) {
            acc[v] = 0.f;
            acc2[v] = make_float2(0.f, 0.f);
        }
        int com_count = 0;
        // overlap: keys above the bracket of each array, and molecules inside a bracket (candidates)
        float4 br = make_float4(0.f, INFINITY, 0.f, INFINITY);  // lo_r, hi_r, lo_t, hi_t
        if (ov) br = *reinterpret_cast<const float4*>(&sh.desc[oi & 3][8]);
        const int lo_r = __float_as_int(br.x), hi_r = __float_as_int(br.y);
        const int lo_t = __float_as_int(br.z), hi_t = __float_as_int(br.w);
        const uint32_t span_r = (uint32_t)(hi_r - lo_r), span_t = (uint32_t)(hi_t - lo_t);
        int n_ar = 0, n_at = 0, n_both = 0, n_cand = 0;
        // Tracker slow path, compacted: molecules for which some tracker may change state (typically 1-2 % of
        // a tile once the origins have aged: crossings of a threshold) are queued in shared memory -- in the
        // part of this origin's stage whose data are already in registers -- and processed after the group
        // loop one molecule per lane, instead of one divergent call per molecule and lane.
        // (Measured alternative: queueing whole lane-groups with one ballot per group instead of one per molecule
        // executes fewer instructions but costs 24 more bytes of spills in the loop: 1.47 -> 1.56 ms.)
        uint32_t fl4_g[SDB_GROUPS], need_g[SDB_GROUPS];
        int n_q = 0;  // queued molecules (warp-uniform)
        unsigned char* const qbase = const_cast<unsigned char*>(st);
        uint32_t* const q_id = reinterpret_cast<uint32_t*>(qbase);                      // local index | flag byte << 8
        float* const q_d2 = reinterpret_cast<float*>(qbase + SDB_PLANE_BYTES);
        float* const q_th = reinterpret_cast<float*>(qbase + 2 * SDB_PLANE_BYTES);
        unsigned char* const q_fl = qbase + 3 * SDB_PLANE_BYTES;                       // new flag byte per molecule

#pragma unroll
        for (int g = 0; g < SDB_GROUPS; ++g) {
            const bool valid = !TAIL || j0[g] < n;  // only the partial last tile has invalid groups
            float4 dx = make_float4(0.f, 0.f, 0.f, 0.f), dy = dx, dt = dx;
            uint32_t fl4 = 0;
            if (valid) {
                const int off = (g * 32 + lane) * 16;
                dx = *reinterpret_cast<const float4*>(st + off);
                dy = *reinterpret_cast<const float4*>(st + SDB_PLANE_BYTES + off);
                dt = *reinterpret_cast<const float4*>(st + 2 * SDB_PLANE_BYTES + off);
                if (track) {
                    if (TAIL) fl4 = __ldcs(reinterpret_cast<const uint32_t*>(fbase + j0[g]));
                    else fl4 = *reinterpret_cast<const uint32_t*>(st + 3 * SDB_PLANE_BYTES + (g * 32 + lane) * 4);
                }
            }
            float ddx[4] = {dx.x, dx.y, dx.z, dx.w};
            float ddy[4] = {dy.x, dy.y, dy.z, dy.w};
            float ddt[4] = {dt.x, dt.y, dt.z, dt.w};
            float d2[4], th[4];
            if (!TAIL && do_sums) {
                // full tile: molecules in pairs, packed f32x2 arithmetic (same roundings as the scalar path:
                // every packed operation rounds each half to nearest like its scalar counterpart)
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int e0 = 2 * p, e1 = 2 * p + 1;
                    const float2 x = __fadd2_rn(make_float2(ddx[e0], ddx[e1]), make_float2(-wx[g][e0], -wx[g][e1]));
                    const float2 y = __fadd2_rn(make_float2(ddy[e0], ddy[e1]), make_float2(-wy[g][e0], -wy[g][e1]));
                    const float2 dd = __fadd2_rn(__fmul2_rn(x, x), __fmul2_rn(y, y));
                    ddx[e0] = x.x; ddx[e1] = x.y;
                    ddy[e0] = y.x; ddy[e1] = y.y;
                    ddt[e0] = (float)((double)ddt[e0] + al[g][e0]);
                    ddt[e1] = (float)((double)ddt[e1] + al[g][e1]);
                    d2[e0] = dd.x; d2[e1] = dd.y;
                    th[e0] = fabsf(ddt[e0]); th[e1] = fabsf(ddt[e1]);
                    const float2 tt = make_float2(th[e0], th[e1]);
                    const float2 t2 = __fmul2_rn(tt, tt);
                    const float2 c = sdb_cos_reduced2(tt);
                    const float2 sq = make_float2(sdb_sqrt_approx(dd.x), sdb_sqrt_approx(dd.y));
                    acc2[SDB_S_DISP] = __fadd2_rn(acc2[SDB_S_DISP], sq);
                    acc2[SDB_S_DISP2] = __fadd2_rn(acc2[SDB_S_DISP2], dd);
                    acc2[SDB_S_DISP4] = __ffma2_rn(dd, dd, acc2[SDB_S_DISP4]);
                    acc2[SDB_S_ROT] = __fadd2_rn(acc2[SDB_S_ROT], tt);
                    acc2[SDB_S_ROT2] = __fadd2_rn(acc2[SDB_S_ROT2], t2);
                    acc2[SDB_S_COS] = __fadd2_rn(acc2[SDB_S_COS], c);
                    acc2[SDB_S_COS2] = __ffma2_rn(c, c, acc2[SDB_S_COS2]);  // sum cos^2; 2 s - n at the end
                    acc2[SDB_S_ROT4] = __ffma2_rn(t2, t2, acc2[SDB_S_ROT4]);
                    acc2[SDB_S_D2R2] = __ffma2_rn(dd, t2, acc2[SDB_S_D2R2]);
                    com_count += (dd.x < com_cut) + (dd.y < com_cut);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool live = !TAIL || j0[g] + e < n;
                if (TAIL || !do_sums) {
                    sdb_apply_increment(ddx[e], ddy[e], ddt[e], wx[g][e], wy[g][e], al[g][e], d2[e], th[e]);
                    if (do_sums && live) {
                        const float t2 = th[e] * th[e];
                        const float c = sdb_cos_reduced(th[e]);
                        acc[SDB_S_DISP] += sdb_sqrt_approx(d2[e]);
                        acc[SDB_S_DISP2] += d2[e];
                        acc[SDB_S_DISP4] = fmaf(d2[e], d2[e], acc[SDB_S_DISP4]);
                        acc[SDB_S_ROT] += th[e];
                        acc[SDB_S_ROT2] += t2;
                        acc[SDB_S_COS] += c;
                        acc[SDB_S_COS2] = fmaf(c, c, acc[SDB_S_COS2]);  // sum cos^2; 2 s - n at the end
                        acc[SDB_S_ROT4] = fmaf(t2, t2, acc[SDB_S_ROT4]);
                        acc[SDB_S_D2R2] = fmaf(d2[e], t2, acc[SDB_S_D2R2]);
                        com_count += d2[e] < com_cut;
                    }
                }
                if (ov) {
                    // keys are non-negative floats: their bit patterns order like the values (and like the
                    // exact kernel's radix select), so "key > hi" is the sign of an integer difference
                    const int kr = __float_as_int(d2[e]), kt = __float_as_int(th[e]);
                    const int ar = hi_r - kr, at = hi_t - kt;  // negative <=> above the bracket
                    const bool in_r = (uint32_t)(kr - lo_r) <= span_r, in_t = (uint32_t)(kt - lo_t) <= span_t;
                    const bool is_cand = live && (in_r || in_t);
                    if (live) {
                        n_ar += (uint32_t)ar >> 31;
                        n_at += (uint32_t)at >> 31;
                        n_both += (uint32_t)(ar & at) >> 31;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, is_cand);
                    if (m) {  // warp-uniform; with history-predicted brackets ~1 % of the molecules qualify
                        const int cs = n_cand + __popc(m & lanemask_lt);
                        if (is_cand && cs < SDB_OV_STAGE) {
                            sh.cand[0][cs] = (uint32_t)(j0[g] + e);
                            sh.cand[1][cs] = (uint32_t)kr;
                            sh.cand[2][cs] = (uint32_t)kt;
                        }
                        n_cand += __popc(m);
                    }
                }
            }
            fl4_g[g] = fl4;
            need_g[g] = 0u;
            if (track) {
                const float m_d2 = fmaxf(fmaxf(d2[0], d2[1]), fmaxf(d2[2], d2[3]));
                const float m_th = fmaxf(fmaxf(th[0], th[1]), fmaxf(th[2], th[3]));
                const bool quiet = (m_d2 < min_cut_d2 && m_th < min_cut_th && (fl4 & state1_x4) == 0u) ||
                                   (fl4 == done_x4 && !literal);
                uint32_t need = 0u;
                if (valid && !quiet) {  // some molecule of this group is above a cut or in an intermediate state
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (TAIL && j0[g] + e >= n) break;
                        const float4 q = quiet_tab[(fl4 >> (8 * e)) & 0xffu];
                        const bool quiet_e = d2[e] < q.x && d2[e] >= q.y && th[e] < q.z && th[e] >= q.w && !literal;
                        need |= quiet_e ? 0u : (1u << e);
                    }
                }
                need_g[g] = need;
                if (__any_sync(0xffffffffu, need != 0u)) {  // warp-uniform
                    __syncwarp();  // every lane holds this group's part of the stage in registers: it is free
                    if (need) *reinterpret_cast<uint32_t*>(q_fl + (g * 32 + lane) * 4) = fl4;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool want = (need >> e) & 1u;
                        const unsigned m = __ballot_sync(0xffffffffu, want);
                        if (want) {
                            // group 0 queues at most 128 entries: they stay inside its own half of the planes
                            const int qs = n_q + __popc(m & lanemask_lt);
                            q_id[qs] = (uint32_t)(g * 128 + lane * 4 + e) | (((fl4 >> (8 * e)) & 0xffu) << 8);
                            q_d2[qs] = d2[e];
                            q_th[qs] = th[e];
                        }
                        n_q += __popc(m);
                    }
                }
            }
            if (valid) {
                sdb_st_stream4(dbase + j0[g], make_float4(ddx[0], ddx[1], ddx[2], ddx[3]));
                sdb_st_stream4(dbase + stride + j0[g], make_float4(ddy[0], ddy[1], ddy[2], ddy[3]));
                sdb_st_stream4(dbase + 2 * stride + j0[g], make_float4(ddt[0], ddt[1], ddt[2], ddt[3]));
            }
        }
        if (track && n_q > 0) {  // warp-uniform
            __syncwarp();
            for (int i = lane; i < n_q; i += 32) {
                const uint32_t id = q_id[i];
                const uint32_t local = id & 0xffu;
                const uint32_t fl_new = sdb_track(a.p.trackers, a.p.n_trackers, (id >> 8) & 0xffu, q_d2[i], q_th[i],
                                                  sbase + tile0 + local, stride, org.timediff, literal);
                q_fl[local] = (unsigned char)fl_new;
            }
            __syncwarp();
#pragma unroll
            for (int g = 0; g < SDB_GROUPS; ++g) {
                if (need_g[g]) {
                    const uint32_t fl4_new = *reinterpret_cast<const uint32_t*>(q_fl + (g * 32 + lane) * 4);
                    if (fl4_new != fl4_g[g]) *reinterpret_cast<uint32_t*>(fbase + j0[g]) = fl4_new;
                }
            }
            // generic-proxy writes to this stage precede the next bulk copy (async proxy) into it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        // Append the staged candidates to the origin's dense list: the slot reservation (one atomic per
        // warp tile) is issued here and its result is only consumed after the reduction below.
        uint32_t cand_base = 0;
        const bool cand_fits = n_cand <= SDB_OV_STAGE;
        if (ov && n_cand > 0 && lane == 0) {
            if (cand_fits) cand_base = atomicAdd(a.ov.cand_cnt + 2 * o, (uint32_t)n_cand);
            else a.ov.cand_cnt[2 * o + 1] = 1u;  // staging overflow: this origin takes the exact path
        }

        if (do_sums) {
            // Per-lane f32 partials (8 molecules) are summed over a third of the warp in f32 (11 terms,
            // pairwise) and from there on in fp64: the thirds by shuffle, the warp tiles by
            // dyn_finalize_kernel.
#pragma unroll
            for (int v = 0; v < SDB_NFSUM; ++v) sh.red[v][lane] = TAIL ? acc[v] : acc2[v].x + acc2[v].y;
            __syncwarp();
            double s = 0.0;
            const int v = lane % SDB_NFSUM, third = lane / SDB_NFSUM;
            if (lane < 3 * SDB_NFSUM) {
                const float* row = &sh.red[v][third];
                const float p0 = (row[0] + row[3]) + (row[6] + row[9]);
                const float p1 = (row[12] + row[15]) + (row[18] + row[21]);
                float p2 = (row[24] + row[27]);
                if (third + 30 < 32) p2 += row[30];
                s = (double)((p0 + p1) + p2);
            }
            s += __shfl_down_sync(0xffffffffu, s, SDB_NFSUM) + __shfl_down_sync(0xffffffffu, s, 2 * SDB_NFSUM);
            if (lane == SDB_S_COS2) s = 2.0 * s - (double)live_warp;  // sum (2 cos^2 - 1)
            const int com_warp = __reduce_add_sync(0xffffffffu, com_count);
            if (lane < SDB_NFSUM) {
                out[lane] = s;
            } else if (lane == SDB_NFSUM) {
                out[SDB_S_COMSTRUCT] = (double)com_warp;
            } else if (lane == SDB_NFSUM + 1) {
                out[SDB_S_BADQUAT] = (double)bad_warp;
            }
            if (ov) {
                const int w_ar = __reduce_add_sync(0xffffffffu, n_ar);
                const int w_at = __reduce_add_sync(0xffffffffu, n_at);
                const int w_both = __reduce_add_sync(0xffffffffu, n_both);
                if (lane == 0) {
                    out[SDB_S_OV_RABOVE] = (double)w_ar;
                    out[SDB_S_OV_TABOVE] = (double)w_at;
                    out[SDB_S_OV_BOTH] = (double)w_both;
                }
            }
        }
        if (ov && n_cand > 0 && cand_fits) {  // warp-uniform
            const uint32_t base = __shfl_sync(0xffffffffu, cand_base, 0);
            if (base + (uint32_t)n_cand <= (uint32_t)a.ov.cand_cap) {
                const long row = (long)o * a.ov.cand_cap + base;
                for (int s = lane; s < n_cand; s += 32) {
                    a.ov.cand_idx[row + s] = sh.cand[0][s];
                    a.ov.cand_r[row + s] = sh.cand[1][s];
                    a.ov.cand_t[row + s] = sh.cand[2][s];
                }
            } else if (lane == 0) {
                a.ov.cand_cnt[2 * o + 1] = 1u;
            }
        }
        __syncwarp();  // every lane is done with this stage, the scratch and the candidate buffer
    }
}

// grid = (warp tiles / 4, segments).  The partial last tile (if any) takes the bounds-checked path.
template <bool SUMS, bool OV, bool TRACK, bool PRE>
__global__ void __launch_bounds__(SDB_BLOCK_THREADS, SDB_MIN_BLOCKS) dyn_step_kernel(const __grid_constant__ StepArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StepWarpShared& sh = reinterpret_cast<StepWarpShared*>(smem_raw)[threadIdx.x >> 5];
    float4* quiet_tab = reinterpret_cast<float4*>(smem_raw + sizeof(StepWarpShared) * SDB_BLOCK_WARPS);
    if (TRACK) {
        for (uint32_t fl = threadIdx.x; fl < 256u; fl += SDB_BLOCK_THREADS)
            quiet_tab[fl] = sdb_quiet_interval(a.p.trackers, a.p.n_trackers, fl);
        __syncthreads();
    }
    const int part = blockIdx.x * SDB_BLOCK_WARPS + (threadIdx.x >> 5);  // warp tile index
    if (part >= a.n_parts) return;
    if (part < a.full_parts)
        dyn_step_tile<SUMS, OV, TRACK, PRE, false>(a, sh, quiet_tab, part);
    else
        dyn_step_tile<SUMS, OV, TRACK, PRE, true>(a, sh, quiet_tab, part);
}

// Increments of every molecule for up to SDB_MAX_INC_SETS previous samples, once per frame: with several
// segments per previous sample (origins are split into segments to fill the GPU) phase 1 -- two f32
// divisions, an f64 atan2 -- would otherwise be repeated by every segment.  grid = (tiles / 4, sets).
struct IncArgs {
    StepArgs a;
    const float* prev[SDB_MAX_INC_SETS];
    uint32_t flags[SDB_MAX_INC_SETS];
    float* wx;
    float* wy;
    double* al;
    uint32_t* bad;
};

__global__ void __launch_bounds__(SDB_BLOCK_THREADS) dyn_increments_kernel(const __grid_constant__ IncArgs ia)
{
    const StepArgs& a = ia.a;
    const int lane = threadIdx.x & 31;
    const int part = blockIdx.x * SDB_BLOCK_WARPS + (threadIdx.x >> 5);
    if (part >= a.n_parts) return;
    const int set = blockIdx.y;
    int j0[SDB_GROUPS];
    float wx[SDB_GROUPS][4], wy[SDB_GROUPS][4];
    double al[SDB_GROUPS][4];
    int bad;
    if (part < a.full_parts)
        bad = sdb_tile_increments<false>(a, ia.prev[set], ia.flags[set], false, set == 0, part, lane, j0, wx, wy, al);
    else
        bad = sdb_tile_increments<true>(a, ia.prev[set], ia.flags[set], false, set == 0, part, lane, j0, wx, wy, al);
    const long set_off = (long)set * a.p.stride;
#pragma unroll
    for (int g = 0; g < SDB_GROUPS; ++g) {
        if (j0[g] >= a.p.n) continue;
        *reinterpret_cast<float4*>(ia.wx + set_off + j0[g]) = make_float4(wx[g][0], wx[g][1], wx[g][2], wx[g][3]);
        *reinterpret_cast<float4*>(ia.wy + set_off + j0[g]) = make_float4(wy[g][0], wy[g][1], wy[g][2], wy[g][3]);
        *reinterpret_cast<double2*>(ia.al + set_off + j0[g]) = make_double2(al[g][0], al[g][1]);
        *reinterpret_cast<double2*>(ia.al + set_off + j0[g] + 2) = make_double2(al[g][2], al[g][3]);
    }
    const int bad_warp = __reduce_add_sync(0xffffffffu, bad);
    if (lane == 0) ia.bad[(long)set * a.n_parts + part] = (uint32_t)bad_warp;
}

// Sum the per-warp-tile partials of one origin: block = 1024 threads = 16 sums x 64 strided readers,
// four independent rows in flight per reader (the walk is latency bound: ~4k rows of 128 bytes).
__global__ void __launch_bounds__(1024) dyn_finalize_kernel(const double* __restrict__ partials, int n_parts,
                                                             double* __restrict__ sums, int with_ov)
{
    __shared__ double sh[64][17];
    const int o = blockIdx.x;
    const int v = threadIdx.x & 15, r = threadIdx.x >> 4;
    const double* base = partials + (long)o * n_parts * SDB_NSUM + v;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int p = r;
    for (; p + 192 < n_parts; p += 256) {
        s0 += base[(long)p * SDB_NSUM];
        s1 += base[(long)(p + 64) * SDB_NSUM];
        s2 += base[(long)(p + 128) * SDB_NSUM];
        s3 += base[(long)(p + 192) * SDB_NSUM];
    }
    for (; p < n_parts; p += 64) s0 += base[(long)p * SDB_NSUM];
    sh[r][v] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) t += sh[i][threadIdx.x];
        const int vv = threadIdx.x;
        // the struct / overlap slots are owned by their own kernels; spare slots are zeroed
        if (vv != SDB_S_STRUCT && vv != SDB_S_OVERLAP) {
            const bool produced = vv < SDB_NFSUM || vv == SDB_S_COMSTRUCT || vv == SDB_S_BADQUAT ||
                                  (with_ov && vv >= SDB_S_OV_RABOVE);
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

extern "C" size_t sdb_dyn_increments_bytes(int32_t n, int32_t stride, int32_t n_sets)
{
    if (n <= 0 || stride < n || n_sets <= 0 || n_sets > SDB_MAX_INC_SETS) return 0;
    return sdb_inc_layout(n, stride, n_sets).bytes;
}

static void sdb_inc_pointers(void* d_ws, int n, int stride, int n_sets, float** wx, float** wy, double** al, uint32_t** bad)
{
    const IncLayout l = sdb_inc_layout(n, stride, n_sets);
    char* b = static_cast<char*>(d_ws);
    *wx = reinterpret_cast<float*>(b);
    *wy = reinterpret_cast<float*>(b + l.off_wy);
    *al = reinterpret_cast<double*>(b + l.off_al);
    *bad = reinterpret_cast<uint32_t*>(b + l.off_bad);
}

extern "C" int sdb_dyn_increments(const SdbDynParams* h_params, const float* d_pos, const float* d_quat,
                                  float* d_cur_compact, const void* const* h_prev, const uint32_t* h_seg_flags,
                                  int32_t n_sets, void* d_ws, void* stream)
{
    if (!h_params || !d_pos || !h_prev || !h_seg_flags || !d_ws || n_sets <= 0 || n_sets > SDB_MAX_INC_SETS) {
        sdb_set_error("sdb_dyn_increments: bad argument");
        return SDB_EINVAL;
    }
    const SdbDynParams& p = *h_params;
    if (p.n <= 0 || p.stride < p.n || (p.stride & 3)) {
        sdb_set_error("sdb_dyn_increments: bad sizes n=%d stride=%d", p.n, p.stride);
        return SDB_EINVAL;
    }
    IncArgs ia{};
    ia.a.p = p;
    ia.a.pos = d_pos;
    ia.a.quat = d_quat;
    ia.a.cur_compact = d_cur_compact;
    ia.a.atan_tab = sdb_device_atan_table();
    if (!ia.a.atan_tab) return SDB_ECUDA;
    ia.a.n_parts = sdb_dyn_num_parts(p.n);
    ia.a.full_parts = p.n / SDB_WARP_MOLS;
    for (int i = 0; i < n_sets; ++i) {
        ia.prev[i] = static_cast<const float*>(h_prev[i]);
        ia.flags[i] = h_seg_flags[i];
    }
    sdb_inc_pointers(d_ws, p.n, p.stride, n_sets, &ia.wx, &ia.wy, &ia.al, &ia.bad);
    const dim3 grid((ia.a.n_parts + SDB_BLOCK_WARPS - 1) / SDB_BLOCK_WARPS, n_sets);
    dyn_increments_kernel<<<grid, SDB_BLOCK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(ia);
    SDB_CHECK_LAUNCH("dyn_increments_kernel");
    return SDB_OK;
}

// Sum of the per-tile partial rows of every origin (second half of sdb_dyn_step; sdb_frame_step runs it
// on its side stream next to the struct kernel).
int sdb_dyn_finalize(const double* d_partials, int32_t n, double* d_sums, int32_t n_origins, int32_t with_ov, void* stream)
{
    if (n_origins <= 0) return SDB_OK;
    dyn_finalize_kernel<<<n_origins, 1024, 0, static_cast<cudaStream_t>(stream)>>>(d_partials, sdb_dyn_num_parts(n), d_sums,
                                                                                   with_ov);
    SDB_CHECK_LAUNCH("dyn_finalize_kernel");
    return SDB_OK;
}

extern "C" int sdb_dyn_step(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, float* d_cur_compact,
                            const SdbOrigin* d_origins, int32_t n_origins, const SdbSegment* d_segments,
                            int32_t n_segments, double* d_partials, double* d_sums, void* d_ov_ws,
                            int32_t ov_ws_origins, void* d_inc_ws, int32_t inc_sets, void* stream)
{
    return sdb_dyn_step_launch(h_params, d_pos, d_quat, d_cur_compact, d_origins, n_origins, d_segments, n_segments,
                               d_partials, d_sums, d_ov_ws, ov_ws_origins, d_inc_ws, inc_sets, 1, stream);
}

int sdb_dyn_step_launch(const SdbDynParams* h_params, const float* d_pos, const float* d_quat, float* d_cur_compact,
                        const SdbOrigin* d_origins, int32_t n_origins, const SdbSegment* d_segments, int32_t n_segments,
                        double* d_partials, double* d_sums, void* d_ov_ws, int32_t ov_ws_origins, void* d_inc_ws,
                        int32_t inc_sets, int32_t finalize, void* stream)
{
    if (!h_params || !d_origins || !d_segments || (d_ov_ws && n_origins > ov_ws_origins)) {
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
    a.sums = p.do_sums ? d_sums : nullptr;
    a.atan_tab = sdb_device_atan_table();
    a.n_parts = sdb_dyn_num_parts(p.n);
    if (!a.atan_tab) return SDB_ECUDA;
    a.ov_on = d_ov_ws != nullptr && p.do_sums;
    if (a.ov_on) sdb_ov_layout(d_ov_ws, ov_ws_origins, p.n, &a.ov);
    else a.ov = SdbOvWorkspace{};
    a.inc_wx = a.inc_wy = nullptr;
    a.inc_al = nullptr;
    a.inc_bad = nullptr;
    if (d_inc_ws && inc_sets > 0 && inc_sets <= SDB_MAX_INC_SETS) {
        float *wx, *wy;
        double* al;
        uint32_t* bad;
        sdb_inc_pointers(d_inc_ws, p.n, p.stride, inc_sets, &wx, &wy, &al, &bad);
        a.inc_wx = wx;
        a.inc_wy = wy;
        a.inc_al = al;
        a.inc_bad = bad;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    a.full_parts = p.n / SDB_WARP_MOLS;
    const int variant = (a.inc_wx ? 8 : 0) | (p.do_sums ? 4 : 0) | (a.ov_on ? 2 : 0) | (p.n_trackers > 0 ? 1 : 0);
    const dim3 grid((a.n_parts + SDB_BLOCK_WARPS - 1) / SDB_BLOCK_WARPS, n_segments);
    constexpr size_t smem = sizeof(StepWarpShared) * SDB_BLOCK_WARPS + 256 * sizeof(float4);
    auto launch = [&](void (*kernel)(const StepArgs)) -> int {
        static std::atomic<bool> configured[16][64];  // [variant][device]: opt in to > 48 KB of shared memory once
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 64 || !configured[variant][dev].load(std::memory_order_acquire)) {
            int rc = sdb_check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                    "dyn_step_kernel: shared memory size");
            if (rc) return rc;
            if (dev < 64) configured[variant][dev].store(true, std::memory_order_release);
        }
        kernel<<<grid, SDB_BLOCK_THREADS, smem, s>>>(a);
        return SDB_OK;
    };
    int lrc;
    switch (variant) {
        case 0: lrc = launch(dyn_step_kernel<false, false, false, false>); break;
        case 1: lrc = launch(dyn_step_kernel<false, false, true, false>); break;
        case 4: lrc = launch(dyn_step_kernel<true, false, false, false>); break;
        case 5: lrc = launch(dyn_step_kernel<true, false, true, false>); break;
        case 6: lrc = launch(dyn_step_kernel<true, true, false, false>); break;
        case 7: lrc = launch(dyn_step_kernel<true, true, true, false>); break;
        case 8: lrc = launch(dyn_step_kernel<false, false, false, true>); break;
        case 9: lrc = launch(dyn_step_kernel<false, false, true, true>); break;
        case 12: lrc = launch(dyn_step_kernel<true, false, false, true>); break;
        case 13: lrc = launch(dyn_step_kernel<true, false, true, true>); break;
        case 14: lrc = launch(dyn_step_kernel<true, true, false, true>); break;
        default: lrc = launch(dyn_step_kernel<true, true, true, true>); break;
    }
    if (lrc) return lrc;
    SDB_CHECK_LAUNCH("dyn_step_kernel");
    if (p.do_sums && finalize) return sdb_dyn_finalize(d_partials, p.n, d_sums, n_origins, a.ov_on, stream);
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
