// pb_math.cuh - branch-free fp64 exp / reciprocal for the flux kernels.
//
// Why not libdevice exp()/division: both inline a slow-path *branch* (BSSY/BSYNC), which
// splits the layer body into many basic blocks and stops ptxas from interleaving the five
// independent exponentials and four divisions of a layer step - the kernels were
// latency-bound on those dependent DFMA chains (ncu r1: "stall_wait" dominant, fp64 pipe
// 44 % busy), and 17 % of all issued instructions were UMOVs materialising 64-bit
// polynomial immediates.  These versions are straight-line code with coefficients taken
// from the constant bank, accurate to ~1 ulp (exp) / ~1 ulp (rcp) - parity tolerance of
// the path is 1e-6 - and they propagate NaN and handle +-inf / underflow like libm.
#pragma once

#include <cuda_runtime.h>

namespace pbm {

// Taylor coefficients 1/k!, k = 2..13, for exp(r), |r| <= ln2/2.
// Remainder bound |r|^14/14! = 4.2e-18 (< 0.04 ulp).
__constant__ double kExpC[12] = {
    1.0 / 2.0,           1.0 / 6.0,           1.0 / 24.0,           1.0 / 120.0,
    1.0 / 720.0,         1.0 / 5040.0,        1.0 / 40320.0,        1.0 / 362880.0,
    1.0 / 3628800.0,     1.0 / 39916800.0,    1.0 / 479001600.0,    1.0 / 6227020800.0};

// exp(x).  All range / special-value handling is done on the high word with integer ALU
// ops and selects, so the fp64 pipe (the bottleneck of the flux kernels) only sees the 17
// arithmetic instructions.  Results below 2^-1021 (x <= -708) are flushed to zero - libm
// would return a value < 3.4e-308 there; every such value in this code path multiplies an
// O(1) quantity - x >= 709.4 saturates to +inf (libm: 709.78; the reference clips its
// growing exponents at 35), NaN gives NaN.
__device__ __forceinline__ double exp(double x)
{
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: round-to-nearest-integer shift
    const double t = fma(x, 1.4426950408889634074, kMagic);
    const int n = __double2loint(t);
    const double nf = t - kMagic;
    double r = fma(nf, -6.93147180369123816490e-01, x);  // ln2 hi (fdlibm split)
    r = fma(nf, -1.90821492927058770002e-10, r);         // ln2 lo
    // two interleaved Horner chains in r^2 (even / odd powers) halve the dependent chain
    const double r2 = r * r;
    double pe = kExpC[10];
    double po = kExpC[11];
    pe = fma(pe, r2, kExpC[8]);
    po = fma(po, r2, kExpC[9]);
    pe = fma(pe, r2, kExpC[6]);
    po = fma(po, r2, kExpC[7]);
    pe = fma(pe, r2, kExpC[4]);
    po = fma(po, r2, kExpC[5]);
    pe = fma(pe, r2, kExpC[2]);
    po = fma(po, r2, kExpC[3]);
    pe = fma(pe, r2, kExpC[0]);
    po = fma(po, r2, kExpC[1]);
    // exp(r) = 1 + r + r^2 * (pe + r * po)
    const double q = fma(po, r, pe);
    const double p = fma(q, r2, r) + 1.0;
    // p in [0.70, 1.42): multiply by 2^n by adding n to the exponent field; valid for
    // -1021 <= n <= 1023, i.e. -708 < x < 709.4 (tested on the high word, integer ALU)
    const int hx = __double2hiint(x);
    const int ax = hx & 0x7fffffff;
    const int lim = hx < 0 ? 0x40862000 /* 708.0 */ : 0x40862b33 /* 709.4 */;
    const bool in_range = ax < lim;
    const bool is_nan = ax > 0x7ff00000 || (ax == 0x7ff00000 && __double2loint(x) != 0);
    const int sat = hx < 0 ? 0 : 0x7ff00000;  // underflow -> 0, overflow -> +inf
    int hi = in_range ? __double2hiint(p) + (n << 20) : sat;
    int lo = in_range ? __double2loint(p) : 0;
    hi = is_nan ? (hx | 0x00080000) : hi;
    lo = is_nan ? __double2loint(x) : lo;
    return __hiloint2double(hi, lo);
}

// 1/x: MUFU.RCP64H seed + two Newton steps, no slow-path branch; the special cases
// (0, subnormal, +-inf, NaN) take the seed, which already holds the libm answer.
__device__ __forceinline__ double rcp(double x)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    // y0 (1 + e)(1 + e^2) = y0 (1 - e^4) / (1 - e): the second step reuses e (e^2 in parallel with the first update)
    // instead of a second residual - dependent depth 3 instead of 4, x y = 1 - e^4 with |e| < 2^-20
    const double e = fma(-x, y0, 1.0);
    double y = fma(y0, e, y0);
    y = fma(y, e * e, y);
    const unsigned ex = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;  // biased exponent
    const bool regular = (ex - 1u) < 0x7feu;                           // 1 .. 0x7fe
    return regular ? y : y0;
}

__device__ __forceinline__ double div(double a, double b) { return a * rcp(b); }

// ---- table exp (toon_reflected_toa5.cuh, toon_thermal.cu) ----------------------------------------------
// A CTA keeps 2^(j/64), j = 0..63, in shared memory, 16 interleaved copies (entry j of copy c at [16 j + c]): a
// thread reads copy (lane & 15), so the 16 lanes of each half-warp phase of an LDS.64 hit 16 different bank
// pairs whatever their j - conflict-free for any argument pattern.  8 KB.
constexpr int kExpTabDoubles = 64 * 16;
__device__ __forceinline__ void exp_tab_fill(double *tabw, int tid, int nthreads)
{
    for (int i = tid; i < kExpTabDoubles; i += nthreads) tabw[i] = ::exp2((double)(i >> 4) * (1.0 / 64.0));
}

// exp(x) for x <= 709 (every argument in this kernel is -dtau/u <= 0 or the clipped lambda dtau <= 40).
// tab = shared table base + (lane & 15); entry j at tab[16 j] holds 2^(j/64).
// 9 fp64 instructions, no branch, dependent depth 8: x = (64 k + j) ln2/64 + r, exp(x) = 2^k 2^(j/64) exp(r); the power of two goes
// into the exponent field of the table entry.  x <= -1024 (optically very thick layers, -inf) is clamped to -1024
// on the integer pipe and k to the normal range, so such results come out as ~1e-308 instead of 0 - every one of
// them multiplies an O(1) quantity; NaN propagates.  Relative error <= 1.2e-16 |x| + 1 ulp.
__constant__ double kExpTabC[5] = {92.33248261689366 /* 64/ln2 */, -0.010830424696249145 /* -ln2/64 */,
                                 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};
__device__ __forceinline__ double exp_tab(double x, const double *tab)
{
    const double kMagic = 6755399441055744.0;               // 1.5 * 2^52
    {
        const unsigned hx = (unsigned)__double2hiint(x);    // -inf <= x <= -1024  <=>  0xC0900000 <= hx <= 0xFFF00000
        const bool big = (hx - 0xC0900000u) <= (0xFFF00000u - 0xC0900000u);
        x = __hiloint2double(big ? (int)0xC0900000u : (int)hx, big ? 0 : __double2loint(x));
    }
    const double t = fma(x, kExpTabC[0], kMagic);
    const int n = __double2loint(t);
    const double nf = t - kMagic;
    const double r = fma(nf, kExpTabC[1], x);                 // |r| <= 5.42e-3
    // exp(r) - 1 = r + r^2 (1/2 + r/6 + r^2 (1/24 + r/120)), Estrin: dependent depth 4 instead of Horner's 6
    const double r2 = r * r;
    const double a = fma(r, kExpTabC[4], 0.5);
    const double b = fma(r, kExpTabC[2], kExpTabC[3]);
    const double c = fma(r2, b, a);
    const double pm1 = fma(r2, c, r);
    const double T = tab[(n & 63) << 4];
    const int k = max(min(n >> 6, 1023), -1022);
    const double Ts = __hiloint2double(__double2hiint(T) + (k << 20), __double2loint(T));
    return fma(Ts, pm1, Ts);
}


// What the kernels call.  Measured on B200 (r1 A/B, reflected 60x10000x5, us per launch):
// libdevice exp + pbm::rcp 104.2 | libdevice both 108.4 | pbm::exp + libdevice div 110.7 |
// pbm both 112.6.  The branch-free exp costs 18 fp64-pipe instructions against libdevice's
// ~15 (degree-13 Taylor vs degree-11 minimax) and the fp64 pipe is the bottleneck, so the
// default is libdevice exp + branch-free reciprocal.  -DPB_CUSTOM_EXP / -DPB_LIBM_RCP flip it.
#ifndef PB_CUSTOM_EXP
__device__ __forceinline__ double kexp(double x) { return ::exp(x); }
#else
__device__ __forceinline__ double kexp(double x) { return pbm::exp(x); }
#endif
#ifdef PB_LIBM_RCP
__device__ __forceinline__ double krcp(double x) { return 1.0 / x; }
#else
__device__ __forceinline__ double krcp(double x) { return pbm::rcp(x); }
#endif

} // namespace pbm
