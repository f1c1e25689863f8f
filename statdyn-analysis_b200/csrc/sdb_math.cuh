// sdb_math.cuh -- arithmetic helpers shared by the kernels; compiled for host too (test hooks).
//
// Everything that feeds a threshold comparison is evaluated with one IEEE rounding per operation
// and no fused multiply-add, in the operation order of the reference's numpy/freud code, so that
// per-molecule displacements are bit-identical to the CPU oracle (oracle/freud_shim.py,
// oracle/dynamics.py).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDA_ARCH__)
#define SDB_HD __host__ __device__ __forceinline__
#define SDB_FADD(a, b) __fadd_rn((a), (b))
#define SDB_FSUB(a, b) __fsub_rn((a), (b))
#define SDB_FMUL(a, b) __fmul_rn((a), (b))
#define SDB_FDIV(a, b) __fdiv_rn((a), (b))
#define SDB_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define SDB_DMUL(a, b) __dmul_rn((a), (b))
#define SDB_DADD(a, b) __dadd_rn((a), (b))
#else
#ifdef __CUDACC__
#define SDB_HD __host__ __device__ inline
#else
#define SDB_HD inline
#endif
// Host build: baseline x86-64 has no FMA instruction, so a*b+c is never contracted.
#define SDB_FADD(a, b) ((a) + (b))
#define SDB_FSUB(a, b) ((a) - (b))
#define SDB_FMUL(a, b) ((a) * (b))
#define SDB_FDIV(a, b) ((a) / (b))
#define SDB_DFMA(a, b, c) fma((a), (b), (c))
#define SDB_DMUL(a, b) ((a) * (b))
#define SDB_DADD(a, b) ((a) + (b))
#endif

// ---- freud Box::wrap, 2D (oracle/freud_shim.py Box.wrap; reference call dynamics/_util.py:149) ----
// fraction = (v - lo - tilt) / L ; fmodf(fraction, 1) ; +1 if negative ; lo + fraction * L (+ tilt)
SDB_HD float sdb_frac_wrap(float delta, float L)
{
    float f = SDB_FDIV(delta, L);
    f = SDB_FSUB(f, truncf(f));  // == fmodf(f, 1.0f), exact
    if (f < 0.0f) f = SDB_FADD(f, 1.0f);
    return f;
}

SDB_HD void sdb_wrap2d(float lx, float ly, float xy, float vx, float vy, float& ox, float& oy)
{
    const float lox = -SDB_FMUL(lx, 0.5f), loy = -SDB_FMUL(ly, 0.5f);
    // makeFraction: delta = v - lo ; delta.x -= xy * v.y   (xz = yz = 0 in a 2D box)
    float dx = SDB_FSUB(vx, lox);
    float dy = SDB_FSUB(vy, loy);
    dx = SDB_FSUB(dx, SDB_FADD(0.0f, SDB_FMUL(xy, vy)));
    const float fx = sdb_frac_wrap(dx, lx);
    const float fy = sdb_frac_wrap(dy, ly);
    // makeCoordinates: v = lo + f * L ; v.x += xy * v.y
    oy = SDB_FADD(loy, SDB_FMUL(fy, ly));
    ox = SDB_FADD(lox, SDB_FMUL(fx, lx));
    ox = SDB_FADD(ox, SDB_FADD(SDB_FMUL(xy, oy), 0.0f));
}

// ---- rotation increment about z (rowan.to_euler(rowan.divide(q, q_prev))[:, 0]) ------------------
// Reference: dynamics/_util.py:150-153.  For planar quaternions q = (w,0,0,z), q' = (w',0,0,z'):
//   p = q (x) conj(q') / |q'|^2 = (a, 0, 0, b) / n',  a = w w' + z z',  b = z w' - w z',  n' = w'^2 + z'^2
//   rowan.to_matrix uses the scale s = 1/|p| (not 1/|p|^2):
//   m10 = 2 s p_z p_w ,  m00 = 1 - 2 s p_z^2 ,  alpha = atan2(clip(m10), clip(m00)).
// Multiplying both arguments by the positive number n'^2 / (2 s) = n' sqrt(n n') / 2 =: h gives
//   alpha = atan2(clamp(a b, -h, h), clamp(h - b^2, -h, h)),   n = w^2 + z^2,
// which needs no division.  Products of two f32 are exact in f64, so a, b, n, n' are correctly
// rounded.  The result agrees with the oracle's literal f64 evaluation to ~1 ulp(f64); it is then
// added to the f32 accumulator and rounded once, so accumulators differ from the oracle's only when
// the f64 sum falls within ~1e-16 relative of an f32 rounding boundary (probability ~1e-8 per update).
struct SdbAtanTable {
    double cs[1024][2];  // cos, sin of 2*pi*i/1024
};

#define SDB_TWO_PI_OVER_1024 6.1359231515425647e-03
#define SDB_1024_OVER_TWO_PI_F 162.97466f

SDB_HD float sdb_fast_atan2f(float y, float x)
{
    // ~1e-5 rad accurate; only used to pick a table entry
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mx > 0.0f ? mn / mx : 0.0f;
    const float s = t * t;
    float a = -0.01172120f;
    a = a * s + 0.05265332f;
    a = a * s - 0.11643287f;
    a = a * s + 0.19354346f;
    a = a * s - 0.33262347f;
    a = a * s + 0.99997726f;
    a = a * t;
    if (ay > ax) a = 1.57079637f - a;
    if (x < 0.0f) a = 3.14159274f - a;
    return copysignf(a, y);
}

// atan2 for finite (y, x) != (0, 0), ~1 ulp.  tab: 1024-entry cos/sin table.
SDB_HD double sdb_atan2_tab(double y, double x, const double (*tab)[2])
{
    const float phi0 = sdb_fast_atan2f((float)y, (float)x);
    const int i = (int)rintf(phi0 * SDB_1024_OVER_TWO_PI_F);  // [-512, 512]
    const double c = tab[i & 1023][0], s = tab[i & 1023][1];
    const double num = SDB_DFMA(y, c, -SDB_DMUL(x, s));
    const double den = SDB_DFMA(x, c, SDB_DMUL(y, s));
    const double r = num / den;  // tan(alpha - theta_i), |r| < 3.2e-3
    const double r2 = SDB_DMUL(r, r);
    double p = SDB_DFMA(r2, -1.0 / 7.0, 0.2);
    p = SDB_DFMA(r2, p, -1.0 / 3.0);
    p = SDB_DMUL(SDB_DMUL(r2, r), p);  // r^3 (-1/3 + r^2/5 - r^4/7)
    return SDB_DADD(SDB_DMUL((double)i, SDB_TWO_PI_OVER_1024), SDB_DADD(r, p));
}

// Returns alpha (f64).  *bad is set when rowan would raise: |p| not within np.allclose of 1
// (rtol 1e-5, atol 1e-8), or either quaternion has zero norm.
SDB_HD double sdb_rot_increment(float w, float z, float pw, float pz, const double (*tab)[2], int& bad)
{
    const double dw = w, dz = z, dpw = pw, dpz = pz;
    const double n = SDB_DFMA(dw, dw, SDB_DMUL(dz, dz));
    const double np_ = SDB_DFMA(dpw, dpw, SDB_DMUL(dpz, dpz));
    const double a = SDB_DFMA(dw, dpw, SDB_DMUL(dz, dpz));
    const double b = SDB_DFMA(dz, dpw, -SDB_DMUL(dw, dpz));
    const double u = SDB_DMUL(n, np_);  // (|q| |q'|)^2
    // |p| = |q|/|q'| ; unit test of rowan._validate_unit: | |p| - 1 | <= 1e-8 + 1e-5
    // |p|^2 = n / n'  ->  compare n against n' (1 +- tol)^2 without dividing
    const double tol_hi = 1.00001000001 * 1.00001000001, tol_lo = 0.99998999999 * 0.99998999999;
    bad = !(n <= np_ * tol_hi && n >= np_ * tol_lo) || !(np_ > 0.0);
    // sqrt(u) by series around 1 when u is close to 1 (always, for f32-normalised quaternions)
    const double e = u - 1.0;
    double root;
    if (fabs(e) < 1e-4) {
        // sqrt(1+e) = 1 + e/2 - e^2/8 + e^3/16 ; |e^4 term| < 4e-18
        root = SDB_DFMA(e, SDB_DFMA(e, SDB_DFMA(e, 0.0625, -0.125), 0.5), 1.0);
    } else {
        root = sqrt(u);
    }
    const double h = SDB_DMUL(SDB_DMUL(np_, root), 0.5);
    double yy = SDB_DMUL(a, b);
    double xx = SDB_DFMA(-b, b, h);
    yy = fmin(fmax(yy, -h), h);
    xx = fmax(xx, -h);
    if (yy == 0.0 && xx > 0.0) return yy;  // exact zero rotation (keeps the sign of zero)
    return sdb_atan2_tab(yy, xx, tab);
}
