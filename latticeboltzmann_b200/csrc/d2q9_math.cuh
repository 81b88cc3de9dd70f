// D2Q9 BGK arithmetic for the B200 kernels.
//
// Two flavours, selected by the EXACT template flag:
//   EXACT = true : every operation is an individually rounded IEEE operation
//                  (__dadd_rn / __dmul_rn / __ddiv_rn ... are never contracted
//                  into FMAs by nvcc), in the expression order of the reference
//                  (c/d2q9.h:59-81 equilibrium, c/d2q9.h:121-131 collide).  The
//                  result is bit-identical to the reference's baseline x86-64
//                  build (no -march => mulsd/addsd, no FMA).
//   EXACT = false: same formulas written as plain C++ so that nvcc contracts
//                  a*b+c into FMAs, with one reciprocal of rho and constant
//                  divisions turned into multiplications.
//
// Three reference divisions are folded away exactly in EXACT mode:
//   4*rho/9  == 4*(rho/9)   and   rho/36 == (rho/9)/4   (power-of-two scaling
//   commutes with rounding),  x/2 == x*0.5.
// That leaves 4 IEEE divisions per cell (rho/9, uu/6, ux/rho, uy/rho).
#pragma once

namespace lbm {

// Channel order and velocities: PyLB/Streaming.py:28-29.
//   i :  0   1   2   3   4   5   6   7   8
//        .   E   N   W   S   NE  NW  SW  SE      (cavity_opt2.py:70)
enum { Q0 = 0, QE = 1, QN = 2, QW = 3, QS = 4, QNE = 5, QNW = 6, QSW = 7, QSE = 8 };
__host__ __device__ constexpr int cx_of(int i) { return i == 1 || i == 5 || i == 8 ? 1 : (i == 3 || i == 6 || i == 7 ? -1 : 0); }
__host__ __device__ constexpr int cy_of(int i) { return i == 2 || i == 5 || i == 6 ? 1 : (i == 4 || i == 7 || i == 8 ? -1 : 0); }
__host__ __device__ constexpr int opp_of(int i) { return i == 0 ? 0 : (i <= 4 ? ((i - 1 + 2) % 4) + 1 : ((i - 5 + 2) % 4) + 5); }

// Individually rounded operations (never contracted).
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double rn_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
// IEEE division.  nvcc's inline fast path (MUFU.RCP64H + 2 Newton steps + Markstein correction, ~12
// instructions) bails out to a ~100-instruction subroutine when the numerator is zero or tiny -- and a
// fluid at rest has ux = uy = uu = 0 EXACTLY in every cell, so a cavity spends most of its time there.
// 0/b is +-0 with sign(a)^sign(b); b is a density or a positive constant here, never 0/inf/nan... if it
// were, the general path below still handles it because only a == 0 with finite non-zero b is shortcut.
__device__ __forceinline__ double rn_div(double a, double b)
{
    if (a == 0.0 && b == b && fabs(b) <= 1.79769313486231570e308 && b != 0.0)
        return __longlong_as_double((__double_as_longlong(a) ^ __double_as_longlong(b)) & (long long)0x8000000000000000ull);
    return __ddiv_rn(a, b);
}
__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float rn_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float rn_div(float a, float b)
{
    if (a == 0.0f && b == b && fabsf(b) <= 3.402823466e38f && b != 0.0f)
        return __int_as_float((__float_as_int(a) ^ __float_as_int(b)) & (int)0x80000000u);
    return __fdiv_rn(a, b);
}

// Correctly rounded x / C for the small integer constants C = 9 and C = 6 in three operations:
//     y = RN(1/C);  q = RN(x*y);  r = x - C*q (exact, one FMA);  q' = RN(q + r*y).
// Proof sketch (p-bit binary floating point, no over/underflow -- hence the range guard):
// |q - x/C| < 2 ulp(q), so r = x - C*q is a multiple of ulp(q) smaller than 2C ulp(q) and is
// exactly representable; the FMA therefore rounds x/C + (r/C)*d with |d| <= 2^-p, a perturbation
// below 2^-(p-1) ulp(q).  x/C = q + (m/C) ulp(q) with m an integer, and since x itself is
// representable (a multiple of 4 or 8 ulp(q)) the quotient can neither be a rounding tie nor come
// closer than ulp(q)/(2C) to one, so the perturbation cannot change the rounding: q' = RN(x/C).
// (checked against __ddiv_rn / __fdiv_rn on 2^28 random and edge inputs in
// tests/test_gpu_parity.py::test_exact_constant_division.)
template <int C>
__device__ __forceinline__ double rn_div_const(double x)
{
    const double y = 1.0 / C;
    const double ax = fabs(x);
    if (ax >= 0x1p-900 && ax <= 0x1p900) {
        const double q = __dmul_rn(x, y);
        const double r = __fma_rn(-double(C), q, x);
        return __fma_rn(r, y, q);
    }
    return rn_div(x, double(C));
}
template <int C>
__device__ __forceinline__ float rn_div_const(float x)
{
    const float y = 1.0f / C;
    const float ax = fabsf(x);
    if (ax >= 0x1p-100f && ax <= 0x1p100f) {
        const float q = __fmul_rn(x, y);
        const float r = __fmaf_rn(-float(C), q, x);
        return __fmaf_rn(r, y, q);
    }
    return rn_div(x, float(C));
}

// c/d2q9.h:59-81
template <typename T, bool EXACT>
__device__ __forceinline__ void d2q9_equilibrium(T rho, T ux, T uy, T (&e)[9])
{
    if (EXACT) {
        const T w1 = rn_div_const<9>(rho);         // rho/9, correctly rounded
        const T w0 = rn_mul(T(4), w1);             // == (4*rho)/9 exactly
        const T w5 = rn_mul(T(0.25), w1);          // == rho/36 exactly
        ux = rn_mul(ux, T(3));
        uy = rn_mul(uy, T(3));
        const T cu5 = rn_add(ux, uy);
        const T cu6 = rn_add(-ux, uy);
        const T cu7 = rn_sub(-ux, uy);
        const T cu8 = rn_sub(ux, uy);
        const T uu = rn_div_const<6>(rn_add(rn_mul(ux, ux), rn_mul(uy, uy)));
        const T hx = rn_mul(rn_mul(ux, ux), T(0.5));
        const T hy = rn_mul(rn_mul(uy, uy), T(0.5));
        e[0] = rn_mul(w0, rn_sub(T(1), uu));
        e[1] = rn_mul(w1, rn_sub(rn_add(rn_add(T(1), ux), hx), uu));
        e[2] = rn_mul(w1, rn_sub(rn_add(rn_add(T(1), uy), hy), uu));
        e[3] = rn_mul(w1, rn_sub(rn_add(rn_sub(T(1), ux), hx), uu));
        e[4] = rn_mul(w1, rn_sub(rn_add(rn_sub(T(1), uy), hy), uu));
        e[5] = rn_mul(w5, rn_sub(rn_add(rn_add(T(1), cu5), rn_mul(rn_mul(cu5, cu5), T(0.5))), uu));
        e[6] = rn_mul(w5, rn_sub(rn_add(rn_add(T(1), cu6), rn_mul(rn_mul(cu6, cu6), T(0.5))), uu));
        e[7] = rn_mul(w5, rn_sub(rn_add(rn_add(T(1), cu7), rn_mul(rn_mul(cu7, cu7), T(0.5))), uu));
        e[8] = rn_mul(w5, rn_sub(rn_add(rn_add(T(1), cu8), rn_mul(rn_mul(cu8, cu8), T(0.5))), uu));
    } else {
        const T w1 = rho * T(1.0 / 9.0);
        const T w0 = T(4) * w1;
        const T w5 = T(0.25) * w1;
        ux *= T(3);
        uy *= T(3);
        const T cu5 = ux + uy, cu6 = uy - ux;
        const T base = T(1) - (ux * ux + uy * uy) * T(1.0 / 6.0);
        const T bx = base + T(0.5) * ux * ux, by = base + T(0.5) * uy * uy;
        const T b5 = base + T(0.5) * cu5 * cu5, b6 = base + T(0.5) * cu6 * cu6;
        e[0] = w0 * base;
        e[1] = w1 * (bx + ux);
        e[2] = w1 * (by + uy);
        e[3] = w1 * (bx - ux);
        e[4] = w1 * (by - uy);
        e[5] = w5 * (b5 + cu5);
        e[6] = w5 * (b6 + cu6);
        e[7] = w5 * (b5 - cu5);
        e[8] = w5 * (b6 - cu6);
    }
}

// c/d2q9.h:125-129 for one cell.  rho follows Eigen 3.4.0's unrolled redux tree
// for a fixed-size 9-vector: ((f0+f1)+(f2+f3)) + ((f4+f5)+(f6+(f7+f8))).
template <typename T, bool EXACT>
__device__ __forceinline__ void d2q9_collide(T (&f)[9], T omega)
{
    T e[9];
    if (EXACT) {
        const T rho = rn_add(rn_add(rn_add(f[0], f[1]), rn_add(f[2], f[3])),
                             rn_add(rn_add(f[4], f[5]), rn_add(f[6], rn_add(f[7], f[8]))));
        const T sx = rn_add(rn_sub(rn_sub(rn_add(rn_sub(f[1], f[3]), f[5]), f[6]), f[7]), f[8]);
        const T sy = rn_sub(rn_sub(rn_add(rn_add(rn_sub(f[2], f[4]), f[5]), f[6]), f[7]), f[8]);
        d2q9_equilibrium<T, true>(rho, rn_div(sx, rho), rn_div(sy, rho), e);
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = rn_add(f[i], rn_mul(omega, rn_sub(e[i], f[i])));
    } else {
        const T rho = ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + (f[7] + f[8])));
        const T sx = (f[1] - f[3]) + (f[5] - f[6]) + (f[8] - f[7]);
        const T sy = (f[2] - f[4]) + (f[5] - f[8]) + (f[6] - f[7]);
        const T inv = T(1) / rho;
        d2q9_equilibrium<T, false>(rho, sx * inv, sy * inv, e);
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = f[i] + omega * (e[i] - f[i]);
    }
}

// ---- simple_flows flavour -------------------------------------------------
// simulators/simple_flows/PoiseuilleFlow.py:25-42 (== slidingLid.py:33-50): a different algebraic
// form of the same equilibrium.  numpy evaluates every operation separately in double, left to
// right, so individually rounded operations in the same order are bit-identical.  2*rho/9, rho/18
// and rho/36 are exact power-of-two multiples of rho/9 (one IEEE division instead of three).
template <typename T>
__device__ __forceinline__ void sf_equilibrium(T rho, T ux, T uy, T (&e)[9])
{
    const T p3 = rn_mul(T(3), rn_add(ux, uy));
    const T m3 = rn_mul(T(3), rn_sub(ux, uy));
    const T uu = rn_mul(T(3), rn_add(rn_mul(ux, ux), rn_mul(uy, uy)));
    const T ux6 = rn_mul(T(6), ux), uy6 = rn_mul(T(6), uy);
    const T uxx9 = rn_mul(rn_mul(T(9), ux), ux);
    const T uyy9 = rn_mul(rn_mul(T(9), uy), uy);
    const T uxy9 = rn_mul(rn_mul(T(9), ux), uy);
    const T r9 = rn_div_const<9>(rho);
    const T a = rn_mul(T(2), r9), b = rn_mul(T(0.5), r9), c = rn_mul(T(0.25), r9);
    e[0] = rn_mul(a, rn_sub(T(2), uu));
    e[1] = rn_mul(b, rn_sub(rn_add(rn_add(T(2), ux6), uxx9), uu));
    e[2] = rn_mul(b, rn_sub(rn_add(rn_add(T(2), uy6), uyy9), uu));
    e[3] = rn_mul(b, rn_sub(rn_add(rn_sub(T(2), ux6), uxx9), uu));
    e[4] = rn_mul(b, rn_sub(rn_add(rn_sub(T(2), uy6), uyy9), uu));
    e[5] = rn_mul(c, rn_add(rn_add(rn_add(T(1), p3), uxy9), uu));
    e[6] = rn_mul(c, rn_add(rn_sub(rn_sub(T(1), m3), uxy9), uu));
    e[7] = rn_mul(c, rn_add(rn_add(rn_sub(T(1), p3), uxy9), uu));
    e[8] = rn_mul(c, rn_add(rn_sub(rn_add(T(1), m3), uxy9), uu));
}

// PoiseuilleFlow.py:49-53: rho = np.sum(grid, axis=0) (f0..f8 in order),
// ux = ((f1+f5+f8) - (f3+f6+f7))/rho, uy = ((f2+f5+f6) - (f4+f7+f8))/rho.
template <typename T>
__device__ __forceinline__ void sf_moments(const T (&f)[9], T &rho, T &ux, T &uy)
{
    rho = rn_add(rn_add(rn_add(rn_add(rn_add(rn_add(rn_add(rn_add(f[0], f[1]), f[2]), f[3]), f[4]), f[5]), f[6]), f[7]), f[8]);
    ux = rn_div(rn_sub(rn_add(rn_add(f[1], f[5]), f[8]), rn_add(rn_add(f[3], f[6]), f[7])), rho);
    uy = rn_div(rn_sub(rn_add(rn_add(f[2], f[5]), f[6]), rn_add(rn_add(f[4], f[7]), f[8])), rho);
}

// PoiseuilleFlow.py:45-47: grid -= relaxation * (grid - feq)
template <typename T>
__device__ __forceinline__ void sf_collide(T (&f)[9], T omega)
{
    T rho, ux, uy, e[9];
    sf_moments<T>(f, rho, ux, uy);
    sf_equilibrium<T>(rho, ux, uy, e);
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = rn_sub(f[i], rn_mul(omega, rn_sub(f[i], e[i])));
}

// Moments as the reference's drivers compute them for output
// (cavity_opt2.py:280-281): rho = np.sum(f, axis=0) accumulates f0..f8 in order.
template <typename T>
__device__ __forceinline__ void d2q9_moments(const T (&f)[9], T &rho, T &ux, T &uy)
{
    rho = rn_add(rn_add(rn_add(rn_add(rn_add(rn_add(rn_add(rn_add(f[0], f[1]), f[2]), f[3]), f[4]), f[5]), f[6]), f[7]), f[8]);
    const T sx = rn_add(rn_sub(rn_sub(rn_add(rn_sub(f[1], f[3]), f[5]), f[6]), f[7]), f[8]);
    const T sy = rn_sub(rn_sub(rn_add(rn_add(rn_sub(f[2], f[4]), f[5]), f[6]), f[7]), f[8]);
    ux = rn_div(sx, rho);
    uy = rn_div(sy, rho);
}

}  // namespace lbm
