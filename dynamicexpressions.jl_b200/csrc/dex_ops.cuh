// dex_ops.cuh — device implementations of the builtin operators (include/dex_ops.def)
// and of their partial derivatives.
//
// Values follow Julia Base scalar semantics for Float32/Float64 — what the reference
// calls as `op(...)` inside its loop kernels (/root/reference/src/Evaluate.jl:366-392)
// — with the documented deviation that domain errors produce NaN instead of throwing.
// Partials are the analytic form of the ChainRules scalar rules behind
// `_zygote_gradient` (/root/reference/ext/DynamicExpressionsZygoteExt.jl:7-15).
//
// Each operator is one row of an X-macro so that the interpreter kernels can expand
// "case OPCODE: for k in 0..K-1: r[k] = EXPR" — the op switch is executed once per
// tape instruction, not once per sample.
//   DEX_UNARY_OPS(X)    X(SYM, value(x),      d/dx(x, v))
//   DEX_BINARY_OPS(X)   X(SYM, value(x,y),    d/dx(x,y,v), d/dy(x,y,v))
//   DEX_TERNARY_OPS(X)  X(SYM, value(x,y,z),  d/dx, d/dy, d/dz)
// with `v` = the value, all of element type T.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/dex_wire.h"

namespace dex {

#define DEX_M1(name)                                                            \
    __device__ __forceinline__ float m_##name(float x) { return name##f(x); }   \
    __device__ __forceinline__ double m_##name(double x) { return name(x); }
#define DEX_M2(name)                                                                     \
    __device__ __forceinline__ float m_##name(float x, float y) { return name##f(x, y); } \
    __device__ __forceinline__ double m_##name(double x, double y) { return name(x, y); }
DEX_M1(fabs) DEX_M1(sqrt) DEX_M1(cbrt) DEX_M1(exp) DEX_M1(exp2) DEX_M1(exp10) DEX_M1(expm1)
DEX_M1(log) DEX_M1(log2) DEX_M1(log10) DEX_M1(log1p) DEX_M1(tan)
DEX_M1(asin) DEX_M1(acos) DEX_M1(atan) DEX_M1(sinh) DEX_M1(cosh) DEX_M1(tanh) DEX_M1(asinh)
DEX_M1(acosh) DEX_M1(atanh) DEX_M1(rint) DEX_M1(floor) DEX_M1(ceil) DEX_M1(trunc) DEX_M1(erf)
DEX_M1(erfc)
DEX_M2(pow) DEX_M2(fmod) DEX_M2(atan2) DEX_M2(copysign)
__device__ __forceinline__ float m_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double m_fma(double a, double b, double c) { return fma(a, b, c); }
#undef DEX_M1
#undef DEX_M2

// ---- sin / cos ------------------------------------------------------------------------
// Float32: an inline fast path (three-constant Cody-Waite reduction by multiples of pi/2 carried in
// FMAs, one sine polynomial on [-pi/2, pi/2]; <= 1.83 ulp) for |x| <= 105615, and an integer
// Payne-Hanek reduction (large_sincosf below) for everything beyond — no library call.
// Float64: the library functions, out of line to keep the hot loop inside the instruction cache.
static __device__ __noinline__ double slow_sin(double x) { return sin(x); }
static __device__ __noinline__ double slow_cos(double x) { return cos(x); }

template <int QADD> __device__ __forceinline__ float fast_sincosf(float x) {
    // x = q pi/2 + r with q even (sin: q = 2 rint(x/pi)) or odd (cos: q = 2 rint(x/pi - 1/2) + 1):
    // r lies in [-pi/2, pi/2] and the result is +-sin(r), the sign being the parity of the rounded
    // integer (low mantissa bit of the 1.5 * 2^23 magic-number sum, exact for |x/pi| < 2^22).
    // sin(r) = r + r^3 (s0 + z (s1 + z (s2 + z s3))): own least-squares fit, <= 1.83 ulp against
    // float64 for |x| <= 105615.  Must stay operation-for-operation identical to the packed forms
    // (sincos_packed in dex_eval.cu, sincos() in gen_interp_ptx.py): results are bit-identical
    // whichever code path a sample takes.
    float m, q;
    if (QADD) {
        const float u = fmaf(x, 0.318309886183790672f, -0.5f);
        m = u + 12582912.0f;
        q = fmaf(m - 12582912.0f, 2.0f, 1.0f);
    } else {
        m = fmaf(x, 0.318309886183790672f, 12582912.0f);
        const float j = m - 12582912.0f;
        q = j + j;
    }
    float r = fmaf(q, -1.5707962513e+00f, x);
    r = fmaf(q, -7.5497894159e-08f, r);
    r = fmaf(q, -5.3903029534e-15f, r);
    const float z = r * r;
    float sp = fmaf(z, 0x1.5dbce6p-19f, -0x1.9f6feep-13f);
    sp = fmaf(sp, z, 0x1.110ed4p-7f);
    sp = fmaf(sp, z, -0x1.55554cp-3f);
    sp = fmaf(sp * z, r, r);
    const unsigned par = (__float_as_uint(m) & 1u) ^ (QADD ? 1u : 0u);
    return __uint_as_float(__float_as_uint(sp) ^ (par << 31));
}
// Arguments beyond the Cody-Waite range (|x| > 105615, up to the largest float): Payne-Hanek
// reduction in INTEGER arithmetic.  The 24 + shift mantissa bits of x are multiplied by a 96-bit
// window of the bits of 2/pi chosen by the exponent (three 32-bit words of k_inv_pio4, which holds
// the hex expansion of 2/pi = 0.A2F9836E 4E441529 FC2757D1 ... in overlapping windows one byte
// apart); the top two bits of the 64-bit product are the quadrant and the rest, as a signed fixed
// point number times pi/2 * 2^-62, is the reduced argument in [-pi/4, pi/4] (absolute error
// < 1.3e-16 for every float from 1e2 to 3.4e38, checked against 400-bit arithmetic).  This is the
// published large-argument scheme of the ARM optimized routines / glibc sinf.  Branch-free and
// convergent: the CUDA library's per-lane Payne-Hanek path cost ~10 % of a whole population launch
// (cos(exp(...)) reaches such arguments in a few percent of the warps).  Then the classic sine /
// cosine kernels chosen by the quadrant.  Inf and NaN give NaN.  The packed PTX form
// (gen_interp_ptx.py `sincos_large`) is operation-for-operation identical.
static __device__ const uint32_t k_inv_pio4[24] = {
    0xa2u,       0xa2f9u,     0xa2f983u,   0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,
    0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u,
    0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u};

template <int QADD> __device__ __forceinline__ float large_sincosf(float x) {
    const uint32_t xi = __float_as_uint(x);
    const uint32_t* arr = k_inv_pio4 + ((xi >> 26) & 15u);
    const uint32_t a0 = __ldg(arr), a4 = __ldg(arr + 4), a8 = __ldg(arr + 8);
    const uint32_t m = ((xi & 0xffffffu) | 0x800000u) << ((xi >> 23) & 7u);
    const uint32_t r0 = m * a0;
    const unsigned long long r1 = (unsigned long long)m * a4, r2 = (unsigned long long)m * a8;
    unsigned long long res = ((unsigned long long)r0 << 32) | (r2 >> 32);
    res += r1;
    const unsigned long long n = (res + (1ull << 61)) >> 62;
    res -= n << 62;
    const double r = __dmul_rn(__ll2double_rn((long long)res), 0x1.921FB54442D18p-62);
    const float rf = __double2float_rn(r);
    const float z = rf * rf;
    float sp = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, z, -1.6666654611e-1f);
    sp = sp * z;
    sp = fmaf(sp, rf, rf);
    float cp = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, z, 4.166664568298827e-2f);
    cp = cp * z;
    const float t2 = fmaf(z, -0.5f, 1.0f);
    cp = fmaf(cp, z, t2);
    const uint32_t q = (uint32_t)n + (uint32_t)QADD;
    const float v = (q & 1u) ? cp : sp;
    // quadrant sign; sin is odd in x (the reduction worked on |x|), cos is even
    uint32_t bits = __float_as_uint(v) ^ ((q & 2u) << 30);
    if (!QADD) bits ^= xi & 0x80000000u;
    if ((xi & 0x7f800000u) == 0x7f800000u) bits = 0x7fffffffu;   // Inf, NaN
    return __uint_as_float(bits);
}
__device__ __forceinline__ float m_sin(float x) {
    if (fabsf(x) > 105615.0f) return large_sincosf<0>(x);   // false for NaN: the fast path propagates it
    return fast_sincosf<0>(x);
}
__device__ __forceinline__ float m_cos(float x) {
    if (fabsf(x) > 105615.0f) return large_sincosf<1>(x);
    return fast_sincosf<1>(x);
}
#ifndef DEX_F64_SINCOS_INLINE
#define DEX_F64_SINCOS_INLINE 1
#endif
#if DEX_F64_SINCOS_INLINE
// inline: the K samples of a thread are independent, and the library sequences are long chains of
// dependent DFMAs — interleaved they hide each other's latency, behind a call they run one by one
__device__ __forceinline__ double m_sin(double x) { return sin(x); }
__device__ __forceinline__ double m_cos(double x) { return cos(x); }
#else
__device__ __forceinline__ double m_sin(double x) { return slow_sin(x); }
__device__ __forceinline__ double m_cos(double x) { return slow_cos(x); }
#endif

template <typename T> __device__ __forceinline__ T t_nan();
template <> __device__ __forceinline__ float t_nan<float>() { return CUDART_NAN_F; }
template <> __device__ __forceinline__ double t_nan<double>() { return CUDART_NAN; }
template <typename T> __device__ __forceinline__ T t_inf();
template <> __device__ __forceinline__ float t_inf<float>() { return CUDART_INF_F; }
template <> __device__ __forceinline__ double t_inf<double>() { return CUDART_INF; }

__device__ __forceinline__ bool t_signbit(float x) { return (__float_as_uint(x) >> 31) != 0u; }
__device__ __forceinline__ bool t_signbit(double x) { return __double2hiint(x) < 0; }
__device__ __forceinline__ bool t_finite(float x) { return (__float_as_uint(x) & 0x7f800000u) != 0x7f800000u; }
__device__ __forceinline__ bool t_finite(double x) { return ((unsigned)__double2hiint(x) & 0x7ff00000u) != 0x7ff00000u; }

// Julia max/min: NaN if either is NaN; max(-0.0, 0.0) == 0.0
template <typename T> __device__ __forceinline__ T j_max(T x, T y) {
    if (x != x || y != y) return t_nan<T>();
    if (x > y) return x;
    if (y > x) return y;
    return t_signbit(x) ? y : x;
}
template <typename T> __device__ __forceinline__ T j_min(T x, T y) {
    if (x != x || y != y) return t_nan<T>();
    if (x < y) return x;
    if (y < x) return y;
    return t_signbit(x) ? x : y;
}
// Julia mod: floored, result takes the sign of y
template <typename T> __device__ __forceinline__ T j_mod(T x, T y) {
    T r = m_fmod(x, y);
    if (r != r) return r;
    if (r == T(0)) return m_copysign(r, y);
    if ((r > T(0)) != (y > T(0))) return r + y;
    return r;
}
template <typename T> __device__ __forceinline__ T j_sign(T x) {
    return x > T(0) ? T(1) : (x < T(0) ? T(-1) : x);
}
template <typename T> __device__ __forceinline__ T b2t(bool b) { return b ? T(1) : T(0); }

#define DEX_LN2 T(0.693147180559945309417232121458176568)
#define DEX_LN10 T(2.30258509299404568401799145468436421)
#define DEX_2_SQRTPI T(1.12837916709551257389615890312154517)

// clang-format off
#define DEX_UNARY_OPS(X) \
    X(NEG,        -x,                                   T(-1)) \
    X(ABS,        m_fabs(x),                            j_sign(x)) \
    X(ABS2,       x * x,                                T(2) * x) \
    X(SQUARE,     x * x,                                T(2) * x) \
    X(CUBE,       x * x * x,                            T(3) * x * x) \
    X(INV,        T(1) / x,                             T(-1) / (x * x)) \
    X(SQRT,       m_sqrt(x),                            T(1) / (T(2) * v)) \
    X(CBRT,       m_cbrt(x),                            T(1) / (T(3) * v * v)) \
    X(EXP,        m_exp(x),                             v) \
    X(EXP2,       m_exp2(x),                            v * DEX_LN2) \
    X(EXP10,      m_exp10(x),                           v * DEX_LN10) \
    X(EXPM1,      m_expm1(x),                           m_exp(x)) \
    X(LOG,        m_log(x),                             T(1) / x) \
    X(LOG2,       m_log2(x),                            T(1) / (x * DEX_LN2)) \
    X(LOG10,      m_log10(x),                           T(1) / (x * DEX_LN10)) \
    X(LOG1P,      m_log1p(x),                           T(1) / (T(1) + x)) \
    X(SIN,        m_sin(x),                             m_cos(x)) \
    X(COS,        m_cos(x),                             -m_sin(x)) \
    X(TAN,        m_tan(x),                             T(1) + v * v) \
    X(ASIN,       m_asin(x),                            T(1) / m_sqrt(T(1) - x * x)) \
    X(ACOS,       m_acos(x),                            T(-1) / m_sqrt(T(1) - x * x)) \
    X(ATAN,       m_atan(x),                            T(1) / (T(1) + x * x)) \
    X(SINH,       m_sinh(x),                            m_cosh(x)) \
    X(COSH,       m_cosh(x),                            m_sinh(x)) \
    X(TANH,       m_tanh(x),                            T(1) - v * v) \
    X(ASINH,      m_asinh(x),                           T(1) / m_sqrt(x * x + T(1))) \
    X(ACOSH,      m_acosh(x),                           T(1) / m_sqrt(x * x - T(1))) \
    X(ATANH,      m_atanh(x),                           T(1) / (T(1) - x * x)) \
    X(ROUND,      m_rint(x),                            T(0)) \
    X(FLOOR,      m_floor(x),                           T(0)) \
    X(CEIL,       m_ceil(x),                            T(0)) \
    X(TRUNC,      m_trunc(x),                           T(0)) \
    X(SIGN,       j_sign(x),                            T(0)) \
    X(RELU,       (x < T(0) ? T(0) : x),                (x < T(0) ? T(0) : T(1))) \
    X(IDENTITY,   x,                                    T(1)) \
    X(SAFE_LOG,   (x <= T(0) ? t_nan<T>() : m_log(x)),      (x <= T(0) ? T(0) : T(1) / x)) \
    X(SAFE_LOG2,  (x <= T(0) ? t_nan<T>() : m_log2(x)),     (x <= T(0) ? T(0) : T(1) / (x * DEX_LN2))) \
    X(SAFE_LOG10, (x <= T(0) ? t_nan<T>() : m_log10(x)),    (x <= T(0) ? T(0) : T(1) / (x * DEX_LN10))) \
    X(SAFE_LOG1P, (x <= T(-1) ? t_nan<T>() : m_log1p(x)),   (x <= T(-1) ? T(0) : T(1) / (T(1) + x))) \
    X(SAFE_SQRT,  (x < T(0) ? t_nan<T>() : m_sqrt(x)),      (x < T(0) ? T(0) : T(1) / (T(2) * v))) \
    X(SAFE_ACOSH, (x < T(1) ? t_nan<T>() : m_acosh(x)),     (x < T(1) ? T(0) : T(1) / m_sqrt(x * x - T(1)))) \
    X(COS2,       sq(m_cos(x)),                         T(-2) * m_cos(x) * m_sin(x)) \
    X(ERF,        m_erf(x),                             DEX_2_SQRTPI * m_exp(-x * x)) \
    X(ERFC,       m_erfc(x),                            -DEX_2_SQRTPI * m_exp(-x * x))

#define DEX_BINARY_OPS(X) \
    X(ADD,        x + y,                                T(1),                       T(1)) \
    X(SUB,        x - y,                                T(1),                       T(-1)) \
    X(MUL,        x * y,                                y,                          x) \
    X(DIV,        x / y,                                T(1) / y,                   -(v / y)) \
    X(POW,        m_pow(x, y),                          y * m_pow(x, y - T(1)),     ((x == T(0) && y > T(0)) ? T(0) : v * m_log(m_fabs(x)))) \
    X(MAX,        j_max(x, y),                          b2t<T>(x > y),              b2t<T>(!(x > y))) \
    X(MIN,        j_min(x, y),                          b2t<T>(!(x > y)),           b2t<T>(x > y)) \
    X(MOD,        j_mod(x, y),                          (m_floor(x / y) == x / y ? t_nan<T>() : T(1)), (m_floor(x / y) == x / y ? t_nan<T>() : -m_floor(x / y))) \
    X(ATAN2,      m_atan2(x, y),                        y / (x * x + y * y),        -x / (x * x + y * y)) \
    X(COPYSIGN,   m_copysign(x, y),                     (t_signbit(x) == t_signbit(y) ? T(1) : T(-1)), T(0)) \
    X(GREATER,    b2t<T>(x > y),                        T(0),                       T(0)) \
    X(LESS,       b2t<T>(x < y),                        T(0),                       T(0)) \
    X(POW_ABS,    m_exp(y * m_log(m_fabs(x))),          v * y / x,                  v * m_log(m_fabs(x))) \
    X(COND,       (x > T(0) ? y : T(0)),                T(0),                       b2t<T>(x > T(0))) \
    X(LOGICAL_OR, b2t<T>(x > T(0) || y > T(0)),         T(0),                       T(0)) \
    X(LOGICAL_AND,b2t<T>(x > T(0) && y > T(0)),         T(0),                       T(0)) \
    X(GREATER_EQ, b2t<T>(x >= y),                       T(0),                       T(0)) \
    X(LESS_EQ,    b2t<T>(x <= y),                       T(0),                       T(0))

#define DEX_TERNARY_OPS(X) \
    X(FMA,        m_fma(x, y, z),                       y, x, T(1)) \
    X(MULADD,     m_fma(x, y, z),                       y, x, T(1)) \
    X(CLAMP,      (x > z ? z : (x < y ? y : x)),        b2t<T>(!((x < y) || (z < x))), b2t<T>(x < y), b2t<T>(z < x)) \
    X(MAX3,       j_max(j_max(x, y), z),                b2t<T>((j_max(x, y) > z) && (x > y)), b2t<T>((j_max(x, y) > z) && !(x > y)), b2t<T>(!(j_max(x, y) > z))) \
    X(MIN3,       j_min(j_min(x, y), z),                b2t<T>(!(j_min(x, y) > z) && !(x > y)), b2t<T>(!(j_min(x, y) > z) && (x > y)), b2t<T>(j_min(x, y) > z)) \
    X(ADD3,       (x + y) + z,                          T(1), T(1), T(1)) \
    X(MUL3,       (x * y) * z,                          y * z, x * z, x * y)
// clang-format on

template <typename T> __device__ __forceinline__ T sq(T c) { return c * c; }

}  // namespace dex
