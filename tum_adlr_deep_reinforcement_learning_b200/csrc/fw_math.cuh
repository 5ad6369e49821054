// fw_math.cuh — branch-free FP64 exp / asin / atan2 for the RHS of the flight dynamics.
//
// Why: the attempt kernel is latency bound (one FP64 issue per ~6 cycles per warp, 2 warps per scheduler, profiles/
// r01_v3): the CUDA math library's asin/atan2/exp are correct to <= 2 ulp but are built from short basic blocks
// (special-case branches) around dependent Horner chains, which stops the compiler from interleaving them with each
// other and with the independent kinematics of the same RHS.  The versions below are straight-line code (selects
// instead of branches, even/odd split polynomials), accurate to ~1 ulp on the domain the dynamics can reach;
// the full domain of each function is covered without a library fallback.  Coefficients: tools/gen_math_coeffs.py (Chebyshev interpolation
// in 60-digit arithmetic).  Accuracy is tested on the GPU against the host libm (tests/test_gpu_math.py).
#pragma once
#include <cuda_runtime.h>

namespace fw {

// Polynomial coefficients live in constant memory: as 64-bit immediates every coefficient costs
// two UMOV / IMAD.MOV issue slots on sm_100a (11 % of the RHS instruction stream, profiles/r01_v7); from the constant
// bank two coefficients arrive with one LDCU.128.
// (non-const on purpose: a const table with a visible initialiser is folded back into immediates)
#define FW_TAB static __device__ __constant__ double
FW_TAB EXP_PE[7] = {1.0 / 479001600.0, 1.0 / 3628800.0, 1.0 / 40320.0, 1.0 / 720.0, 1.0 / 24.0, 0.5, 1.0};
FW_TAB EXP_PO[7] = {1.0 / 6227020800.0, 1.0 / 39916800.0, 1.0 / 362880.0, 1.0 / 5040.0, 1.0 / 120.0, 1.0 / 6.0, 1.0};
FW_TAB ASIN_PE[7] = {-0.01924167174674304, 0.0030448799094556773, 0.009621842970100282, 0.01396378001220357,
                     0.02237215744350722, 0.044642857142551895, 0.16666666666666666};
FW_TAB ASIN_PO[7] = {0.02961201126495512, 0.019554513336123378, 0.009319560794767446, 0.011566459612121669,
                     0.017352816540325496, 0.03038194447553234, 0.07500000000000118};
FW_TAB ATAN_PE[6] = {-0.034570561981427744, -0.05230454270650244, -0.06666424885738255, -0.09090908753500877,
                     -0.14285714285659828, -0.3333333333333333};
FW_TAB ATAN_PO[6] = {0.016285756855221028, 0.04551593220626549, 0.05878928997834775, 0.07692296375032143,
                     0.11111111105155447, 0.19999999999999804};

// The same coefficients as compile-time constants (folded into immediates): the one-wave init kernel runs every RHS
// once per launch with a cold constant cache and a 128-register budget, where the table loads cost 1.7 us and 140 B of spills.
#define FW_IMM static __device__ const double
FW_IMM EXP_PE_I[7] = {1.0 / 479001600.0, 1.0 / 3628800.0, 1.0 / 40320.0, 1.0 / 720.0, 1.0 / 24.0, 0.5, 1.0};
FW_IMM EXP_PO_I[7] = {1.0 / 6227020800.0, 1.0 / 39916800.0, 1.0 / 362880.0, 1.0 / 5040.0, 1.0 / 120.0, 1.0 / 6.0, 1.0};
FW_IMM ASIN_PE_I[7] = {-0.01924167174674304, 0.0030448799094556773, 0.009621842970100282, 0.01396378001220357,
                     0.02237215744350722, 0.044642857142551895, 0.16666666666666666};
FW_IMM ASIN_PO_I[7] = {0.02961201126495512, 0.019554513336123378, 0.009319560794767446, 0.011566459612121669,
                     0.017352816540325496, 0.03038194447553234, 0.07500000000000118};
FW_IMM ATAN_PE_I[6] = {-0.034570561981427744, -0.05230454270650244, -0.06666424885738255, -0.09090908753500877,
                     -0.14285714285659828, -0.3333333333333333};
FW_IMM ATAN_PO_I[6] = {0.016285756855221028, 0.04551593220626549, 0.05878928997834775, 0.07692296375032143,
                     0.11111111105155447, 0.19999999999999804};

#define FW_COEF(name, k) (CT ? name[k] : name##_I[k])

// 1/sqrt(x) for normal-range x > 0: hardware seed (rsqrt.approx.ftz.f64, ~2^-21) and ONE third-order correction
// r (1 + e/2 + 3 e^2 / 8), e = 1 - x r^2 (truncation 5/16 e^3 < 2^-60): ~1 ulp like ::rsqrt, but straight-line — the
// library version carries a slow-path CALL for denormals / specials that splits the RHS into basic blocks the scheduler
// cannot interleave across.  x = 0 gives inf (as ::rsqrt); denormal x is flushed to 0 (callers guard with selects).
__device__ __forceinline__ double rsqrt_fast(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double t = x * r;
    const double e = fma(-t, r, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double w = r * e;
    return fma(w, p, r);
}

// exp(x) for |x| <= 700: n = rint(x log2 e), r = x - n ln2 (two-term Cody-Waite), Taylor degree 13 on |r| <= 0.3466
// (truncation 4e-18), scaled by 2^n through the exponent field.
template <bool CT = true> __device__ __forceinline__ double exp_bf(double x) {
    const double t = x * 1.4426950408889634074;
    const double n = (t + 6755399441055744.0) - 6755399441055744.0;       // rint via the 1.5 * 2^52 trick
    double r = fma(-n, 6.93147180369123816490e-01, x);
    r = fma(-n, 1.90821492927058770002e-10, r);
    const double r2 = r * r;
    // even / odd Horner halves of sum r^k / k!
    double pe = FW_COEF(EXP_PE, 0), po = FW_COEF(EXP_PO, 0);
#pragma unroll
    for (int k = 1; k < 7; ++k) { pe = fma(pe, r2, FW_COEF(EXP_PE, k)); po = fma(po, r2, FW_COEF(EXP_PO, k)); }
    const double p = fma(po, r, pe);
    const int ni = (int)n;
    return __hiloint2double(__double2hiint(p) + ni * 1048576, __double2loint(p));
}

// log(x) for normal-range x > 0 (callers clamp), straight line: x = 2^k m with m in [sqrt(1/2), sqrt(2)), f = m - 1,
// s = f / (2 + f), log(1 + f) = f - hfsq + s (hfsq + R(s^2)) with the degree-7 minimax R of fdlibm's e_log.c
// (|error| < 2^-58.45 there), result k ln2_hi + (log(1 + f) + k ln2_lo): < 1 ulp (tests/test_gpu_math.py).
FW_TAB LOG_LG[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
                    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01};
FW_IMM LOG_LG_I[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
                      1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01};
__device__ __forceinline__ double rcp_fast(double x);
template <bool CT = true> __device__ __forceinline__ double log_bf(double x) {
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;          // m >= sqrt(2): halve it, k + 1
    k += i >> 20;
    const double m = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
    const double f = m - 1.0;
    const double s = f * rcp_fast(2.0 + f);
    const double z = s * s, w = z * z;
    double t1 = FW_COEF(LOG_LG, 5), t2 = FW_COEF(LOG_LG, 6);
    t1 = fma(t1, w, FW_COEF(LOG_LG, 3)); t2 = fma(t2, w, FW_COEF(LOG_LG, 4));
    t1 = fma(t1, w, FW_COEF(LOG_LG, 1)); t2 = fma(t2, w, FW_COEF(LOG_LG, 2));
    t2 = fma(t2, w, FW_COEF(LOG_LG, 0));
    const double R = fma(t1, w, 0.0) + t2 * z;        // Lg2 w + Lg4 w^2 + Lg6 w^3  +  z (Lg1 + Lg3 w + Lg5 w^2 + Lg7 w^3)
    const double hfsq = 0.5 * f * f, dk = (double)k;
    return dk * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
}

// 1/x for normal-range x: hardware seed (rcp.approx.ftz.f64, ~2^-23) + two Newton steps -> ~1 ulp, 6 instructions
// instead of the ~35 of an IEEE-rounded division with its special-case fix-ups.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// asin(x), |x| <= 1, straight line.  |x| <= 1/2: x + x z g(z), z = x^2.  Otherwise asin = pi/2 - 2 asin(sqrt((1-|x|)/2))
// with the same polynomial on z = (1-|x|)/2; the square root comes from rsqrt plus its exact residual (fma) so that the
// doubled term keeps ~1 ulp.  Worst error 1.6 ulp (tools/gen_math_coeffs.py grid; tests/test_gpu_math.py).
template <bool CT = true> __device__ __forceinline__ double asin_bf(double x) {
    const double ax = fabs(x);
    const bool big = ax > 0.5;
    const double z = big ? (1.0 - ax) * 0.5 : x * x;
    const double rs = rsqrt_fast(z);
    double sq = z * rs;                                   // ~sqrt(z)
    double corr = fma(-sq, sq, z) * (0.5 * rs);           // sqrt(z) - sq to first order
    if (!(z > 0.0)) { sq = 0.0; corr = 0.0; }
    const double sv = big ? sq : ax;
    const double cv = big ? corr : 0.0;
    const double z2 = z * z;
    double pe, po;
    pe = FW_COEF(ASIN_PE, 0); po = FW_COEF(ASIN_PO, 0);
#pragma unroll
    for (int k = 1; k < 7; ++k) { pe = fma(pe, z2, FW_COEF(ASIN_PE, k)); po = fma(po, z2, FW_COEF(ASIN_PO, k)); }
    const double g = fma(po, z, pe);
    const double pp = sv + fma(sv * z, g, cv);
    const double res = big ? 1.57079632679489655800e+00 - (2.0 * pp - 6.12323399573676603587e-17) : pp;
    return copysign(res, x);
}

// atan2(y, x): one reciprocal.  m = min/max of |x|, |y|; if m > tan(pi/8) the argument is folded with
// atan(m) = pi/4 + atan((m - 1)/(m + 1)) which is formed directly as (mn - mx)/(mn + mx); then |t| <= tan(pi/8) and
// atan(t) = t + t w q(w), w = t^2.  Octant / quadrant fix-ups are selects with hi/lo split constants.
template <bool CT = true> __device__ __forceinline__ double atan2_bf(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    const bool big = mn > 0.41421356237309503 * mx;
    const double num = big ? mn - mx : mn;
    const double den = big ? mn + mx : mx;
    double t = num * rcp_fast(den);
    if (!(mx > 1e-290)) t = 0.0;                   // atan2(0, 0) = 0 (and the flushed-denormal corner)
    const double z = t * t;
    const double z2 = z * z;
    double pe, po;
    pe = FW_COEF(ATAN_PE, 0); po = FW_COEF(ATAN_PO, 0);
#pragma unroll
    for (int k = 1; k < 6; ++k) { pe = fma(pe, z2, FW_COEF(ATAN_PE, k)); po = fma(po, z2, FW_COEF(ATAN_PO, k)); }
    const double q = fma(po, z, pe);
    double a = fma(t * z, q, t);
    if (big) a = 7.85398163397448278999e-01 + (a + 3.06161699786838301793e-17);        // + pi/4
    if (ay > ax) a = 1.57079632679489655800e+00 - (a - 6.12323399573676603587e-17);    // pi/2 - a
    if (x < 0.0) a = 3.14159265358979311600e+00 - (a - 1.22464679914735317723e-16);    // pi - a
    return copysign(a, y);
}

// x^p for x > 0 as exp(p log x), x clamped into [1e-300, 1e300]: the step-size controller's err^-0.2 and the initial
// step's (0.01 / d)^0.2 (scipy rk.py:139-166, common.py:118-126), where the library pow is a ~300-instruction call once
// per attempt.  |p log x| <= 140 keeps the result within ~8 ulp of pow (tests/test_gpu_math.py); both uses then pass
// through min / max against constants, and the parity suite sees identical step-size decisions on every fixture step.
template <bool CT = true> __device__ __forceinline__ double pow_hot_bf(double x, double p) {
    const double xc = fmin(fmax(x, 1e-300), 1e300);
    const double r = exp_bf<CT>(p * log_bf<CT>(xc));
    return (x != x) ? x : r;                       // NaN in, NaN out (the clamp would swallow it)
}

}  // namespace fw
