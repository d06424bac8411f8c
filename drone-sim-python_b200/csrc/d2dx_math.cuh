// Branch-free fp64 elementary functions for the rollout kernels (sm_100a).
//
// Why not CUDA's libdevice versions: ncu on the round-1 kernel (profiles/r1a_rollout_dfff_circle.md) showed 24 % of
// all issued instructions to be UMOV pairs materialising the fp64 literals of libdevice's polynomials, plus ~10 %
// BRA/BSSY/BSYNC around slow paths that this workload never takes (|angle| < 1e5, no denormals).  Here every
// coefficient lives in one __constant__ table (ptxas fetches two per LDCU.128) and every function is straight-line.
// Accuracy: polynomial approximation errors 2e-17 (sin), 9e-19 (cos), 4.5e-18 (atan core) relative (checked with
// mpmath, 200 bits); results within 1-2 ulp of the correctly rounded value -- far inside the 1e-9 parity budget and
// of the same class as libdevice (sincos 1-2 ulp, atan2 2 ulp).
#pragma once
#include <cuda_runtime.h>

namespace d2dx {
namespace fm {

// layout chosen so that consecutive uses are adjacent (16-byte pairs)
static __constant__ __align__(16) double kTab[76] = {
    /* 0*/ 0.6366197723675814, 6755399441055744.0,                    // 2/pi, 1.5*2^52
    /* 2*/ -1.5707963267948966, -6.123233995736766e-17,               // -pi/2 split in three
    /* 4*/ 1.4973849048591698e-33, 0.0,
    /* 6*/ 1.590307857061102704e-10, -2.505091138364548653e-08,       // sin: r + r^3 P(r^2)
    /* 8*/ 2.755731498463002875e-06, -1.984126983447703004e-04,
    /*10*/ 8.333333333329348558e-03, -1.666666666666666297e-01,
    /*12*/ -1.136781730462628422e-11, 2.087588337859780049e-09,       // cos: 1 + r^2 Q(r^2)
    /*14*/ -2.755731554299955694e-07, 2.480158729361868326e-05,
    /*16*/ -1.388888888888066683e-03, 4.166666666666663660e-02,
    /*18*/ -5.000000000000000000e-01, 0.0,
    /*20*/ 1.62858201153657823623e-02, -3.65315727442169155270e-02,   // atan core, highest power first (fdlibm aT[10..0])
    /*22*/ 4.97687799461593236017e-02, -5.83357013379057348645e-02,
    /*24*/ 6.66107313738753120669e-02, -7.69187620504482999495e-02,
    /*26*/ 9.09088713343650656196e-02, -1.11111104054623557880e-01,
    /*28*/ 1.42857142725034663711e-01, -1.99999999998764832476e-01,
    /*30*/ 3.33333333333329318027e-01, 0.0,
    /*32*/ 4.63647609000806093515e-01, 2.26987774529616870924e-17,    // atan(0.5) hi, lo
    /*34*/ 7.85398163397448278999e-01, 3.06161699786838301793e-17,    // atan(1)
    /*36*/ 9.82793723247329054082e-01, 1.39033110312309984516e-17,    // atan(1.5)
    /*38*/ 1.57079632679489655800e+00, 6.12323399573676603587e-17,    // atan(inf)
    // tan(x) = x P(x^2) / Q(x^2), Pade [4/5] of tan(x)/x in x^2: 7e-17 relative on |x| <= 1.15 with these doubles
    /*40*/ 8.400421197118823e-08, -3.9313971202516095e-05,            // P, highest power first
    /*42*/ 0.0043343653250774, -0.14035087719298245,
    /*44*/ -1.5273493085670588e-09, 2.2681137232220826e-06,           // Q, highest power first
    /*46*/ -0.00048159614723082214, 0.02889576883384933,
    /*48*/ -0.47368421052631576, 0.0,
    // small-angle sin / cos (|d| < 0.1): d (1 + d^2 S(d^2)), 1 + d^2 C(d^2)
    /*50*/ 2.7557319223985893e-06, -1.984126984126984e-04,            // 1/9!, -1/7!
    /*52*/ 8.333333333333333e-03, -1.6666666666666666e-01,            // 1/5!, -1/3!
    /*54*/ 2.48015873015873e-05, -1.3888888888888889e-03,             // 1/8!, -1/6!
    /*56*/ 4.1666666666666664e-02, -0.5,                              // 1/4!, -1/2!
    /*58*/ 0.0, 0.0,
    // exp_neg: log2 e, -ln2 hi, -ln2 lo, then the Taylor coefficients 1/13! .. 1/3! (the magic constant is kTab[1])
    /*60*/ 1.4426950408889634, -6.93147180369123816490e-01,
    /*62*/ -1.90821492927058770002e-10, 1.6059043836821613e-10,
    /*64*/ 2.08767569878681e-09, 2.505210838544172e-08,
    /*66*/ 2.755731922398589e-07, 2.7557319223985893e-06,
    /*68*/ 2.48015873015873e-05, 1.984126984126984e-04,
    /*70*/ 1.3888888888888889e-03, 8.333333333333333e-03,
    /*72*/ 4.1666666666666664e-02, 1.6666666666666666e-01,
    /*74*/ 0.0, 0.0,
};

// 1/b to ~1 ulp: MUFU.RCP64H seed y0 (it looks at the upper 32 bits of b only: relative error e0 <= 1.04 * 2^-20, measured by
// d2dx_math_probe row 8), then y0 (1 + e0 + e0^2) in three DFMA: e = 1 - b y0, t = e + e^2, y = y0 + y0 t -> error e0^3 < 2^-59.
// No slow path (b normal, non-zero).
__device__ __forceinline__ double rcp_seed(double b) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
  return y;
}
__device__ __forceinline__ double rcp(double b) {
  const double y = rcp_seed(b);
  const double e = fma(-b, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

// a/b with a final residual correction (<= 1 ulp)
__device__ __forceinline__ double div(double a, double b) {
  const double y = rcp(b);
  const double q = a * y;
  return fma(fma(-b, q, a), y, q);
}

// 1/sqrt(a), sqrt(a) for normal positive a
__device__ __forceinline__ double rsqrt_seed(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  return y;
}
__device__ __forceinline__ double rsqrt(double a) {
  const double y = rsqrt_seed(a);
  // e = 1 - a y^2 (|e| <= 1.9 * 2^-20, measured: probe row 9);  1/sqrt(1 - e) = 1 + e/2 + 3 e^2/8 + 5 e^3/16 + ...: the cubic term is < 2^-61
  const double e = fma(-a * y, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

__device__ __forceinline__ double sqrt(double a) {
  const double y = rsqrt(a);
  const double g = a * y;
  return fma(fma(-g, g, a), 0.5 * y, g);    // one correction step in the residual
}

// sin and cos of x, |x| < ~1e5 (3-term Cody-Waite reduction to [-pi/4, pi/4])
__device__ __forceinline__ void sincos(double x, double& s, double& c) {
  const double t = fma(x, kTab[0], kTab[1]);
  const int q = __double2loint(t);
  const double j = t - kTab[1];
  double r = fma(j, kTab[2], x);
  r = fma(j, kTab[3], r);
  r = fma(j, kTab[4], r);
  const double z = r * r;
  double ps = fma(kTab[6], z, kTab[7]);
  double pc = fma(kTab[12], z, kTab[13]);
  ps = fma(ps, z, kTab[8]);  pc = fma(pc, z, kTab[14]);
  ps = fma(ps, z, kTab[9]);  pc = fma(pc, z, kTab[15]);
  ps = fma(ps, z, kTab[10]); pc = fma(pc, z, kTab[16]);
  ps = fma(ps, z, kTab[11]); pc = fma(pc, z, kTab[17]);
  pc = fma(pc, z, kTab[18]);
  const double sn = fma(ps * z, r, r);
  const double cs = fma(pc, z, 1.0);
  const double a = (q & 1) ? cs : sn;
  const double b = (q & 1) ? sn : cs;
  // quadrant signs straight into the sign bit (one integer op each instead of a negate and two selects)
  s = __hiloint2double(__double2hiint(a) ^ ((q & 2) << 30), __double2loint(a));
  c = __hiloint2double(__double2hiint(b) ^ (((q + 1) & 2) << 30), __double2loint(b));
}

// tan(x) as a ratio: tan x = tn / td with tn = x P(x^2), td = Q(x^2); valid for |x| <= 1.15 (bank angles).  The caller
// folds td into a division it performs anyway (g tan(phi) / v = g tn / (v td)).
__device__ __forceinline__ void tan_ratio(double x, double& tn, double& td) {
  const double u = x * x;
  double p = fma(kTab[40], u, kTab[41]);
  double q = fma(kTab[44], u, kTab[45]);
  p = fma(p, u, kTab[42]); q = fma(q, u, kTab[46]);
  p = fma(p, u, kTab[43]); q = fma(q, u, kTab[47]);
  p = fma(p, u, 1.0);      q = fma(q, u, kTab[48]);
  q = fma(q, u, 1.0);
  tn = x * p; td = q;
}

// sin and cos of a small increment |d| < 0.1 (absolute error < 3e-17)
__device__ __forceinline__ void sincos_small(double d, double& s, double& c) {
  const double u = d * d;
  double ps = fma(kTab[50], u, kTab[51]);
  double pc = fma(kTab[54], u, kTab[55]);
  ps = fma(ps, u, kTab[52]); pc = fma(pc, u, kTab[56]);
  ps = fma(ps, u, kTab[53]); pc = fma(pc, u, kTab[57]);
  s = fma(ps * u, d, d);
  c = fma(pc, u, 1.0);
}

// exp(y) for y <= 0 (Gaussian-type penalties): k = rint(y log2 e), r = y - k ln2 (two-term), degree-13 Taylor polynomial
// on |r| <= ln2/2 (truncation 4e-18 relative), scaling through the exponent field (saturating at 2^-1000).
// The polynomial is evaluated by Estrin's scheme: 16 fp64 instructions instead of Horner's 13, but a dependency depth of 5
// instead of 13 -- the collision kernels are bound by the latency of this chain, not by its instruction count.
__device__ __forceinline__ double exp_neg(double y) {
  const double t = fma(y, kTab[60], kTab[1]);
  const int k = __double2loint(t);
  const double j = t - kTab[1];
  double r = fma(j, kTab[61], y);
  r = fma(j, kTab[62], r);
  const double r2 = r * r;
  const double p01 = 1.0 + r;                                   // c0 + c1 r
  const double p23 = fma(kTab[73], r, 0.5);                     // c2 + c3 r       (1/2!, 1/3!)
  const double p45 = fma(kTab[71], r, kTab[72]);                // 1/4! + r/5!
  const double p67 = fma(kTab[69], r, kTab[70]);                // 1/6! + r/7!
  const double p89 = fma(kTab[67], r, kTab[68]);                // 1/8! + r/9!
  const double pab = fma(kTab[65], r, kTab[66]);                // 1/10! + r/11!
  const double pcd = fma(kTab[63], r, kTab[64]);                // 1/12! + r/13!
  const double r4 = r2 * r2;
  const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(pab, r2, p89);
  const double r8 = r4 * r4;
  const double s0 = fma(q1, r4, q0), s1 = fma(pcd, r4, q2);
  const double p = fma(s1, r8, s0);
  // exponent clamped at -1000 instead of a compare and two selects: below e^-693 the result is p 2^-1000 ~ 1e-301 rather than
  // the true (sub-1e-301) value -- the same thing to every sum these penalties enter
  const int hi = __double2hiint(p) + (max(k, -1000) << 20);
  return __hiloint2double(hi, __double2loint(p));
}

// atan(num/den) for den > 0 ... folded into atan2 below: one division in total.
// atan2(y, x) (fdlibm's 4-interval reduction atan(t) = atan(c) + atan((t-c)/(1+ct)), c in {0, .5, 1, 1.5, inf}, with
// t = |y|/|x| never formed: (t-c)/(1+ct) = (|y| - c|x|)/(|x| + c|y|)).  x = y = 0 returns 0 like np.arctan2.
__device__ __forceinline__ double atan2(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  // interval id from a 20-bit estimate of t = ay/ax (MUFU seed x one multiply), compared on the HIGH WORD with integer
  // instructions: thresholds 7/16, 11/16, 19/16, 39/16 have zero low words, and the reduction is valid on either side of a
  // threshold, so an estimate that is off in the 20th bit only moves the switch point.  (Four DSETP + five DMUL on the fp64
  // pipe before: the rollout kernel is bound by that pipe.)  ax = 0 or ay/ax overflowing gives inf / NaN -> last interval.
  double rax;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rax) : "d"(ax));
  const int qh = __double2hiint(ay * rax) & 0x7fffffff;
  const bool g0 = qh >= 0x3FDC0000, g1 = qh >= 0x3FE60000, g2 = qh >= 0x3FF30000, g3 = qh >= 0x40038000;
  const double cc = g2 ? 1.5 : (g1 ? 1.0 : (g0 ? 0.5 : 0.0));
  double num = fma(-cc, ax, ay), den = fma(cc, ay, ax);
  if (g3) { num = -ax; den = ay; }
  const double hi = g3 ? kTab[38] : (g2 ? kTab[36] : (g1 ? kTab[34] : (g0 ? kTab[32] : 0.0)));   // uniform loads + selects
  const bool den0 = ((__double2hiint(den) & 0x7fffffff) | __double2loint(den)) == 0;          // den == 0.0 on the integer pipe
  const double t = den0 ? 0.0 : div(num, den);
  const double z = t * t;
  double p = fma(kTab[20], z, kTab[21]);
  p = fma(p, z, kTab[22]); p = fma(p, z, kTab[23]); p = fma(p, z, kTab[24]); p = fma(p, z, kTab[25]);
  p = fma(p, z, kTab[26]); p = fma(p, z, kTab[27]); p = fma(p, z, kTab[28]); p = fma(p, z, kTab[29]);
  p = fma(p, z, kTab[30]);
  const double ts = t * (z * p);
  double r = hi - (ts - t);                  // in [0, pi/2]; the low words of atan(c) (< 0.5 ulp of the result) are not carried
  if (__double2hiint(x) < 0) r = 3.141592653589793 - (r - 1.2246467991473532e-16);   // sign bit: x < 0, and x = -0.0 like np.arctan2
  return copysign(r, y);
}

// atan(v): the same reduction with x = 1 (thresholds are compile-time constants)
__device__ __forceinline__ double atan(double v) {
  const double ay = fabs(v);
  // the four thresholds have zero low words: comparing the high word is the exact comparison, on the integer pipe (NaN -> g3)
  const int qh = __double2hiint(ay);
  const bool g0 = qh >= 0x3FDC0000, g1 = qh >= 0x3FE60000, g2 = qh >= 0x3FF30000, g3 = qh >= 0x40038000;
  const double cc = g2 ? 1.5 : (g1 ? 1.0 : (g0 ? 0.5 : 0.0));
  double num = ay - cc, den = fma(cc, ay, 1.0);
  if (g3) { num = -1.0; den = ay; }
  const double hi = g3 ? kTab[38] : (g2 ? kTab[36] : (g1 ? kTab[34] : (g0 ? kTab[32] : 0.0)));
  const double t = div(num, den);
  const double z = t * t;
  double p = fma(kTab[20], z, kTab[21]);
  p = fma(p, z, kTab[22]); p = fma(p, z, kTab[23]); p = fma(p, z, kTab[24]); p = fma(p, z, kTab[25]);
  p = fma(p, z, kTab[26]); p = fma(p, z, kTab[27]); p = fma(p, z, kTab[28]); p = fma(p, z, kTab[29]);
  p = fma(p, z, kTab[30]);
  const double ts = t * (z * p);
  return copysign(hi - (ts - t), v);
}

}  // namespace fm
}  // namespace d2dx
