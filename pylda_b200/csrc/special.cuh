// fp64 special functions for the VB E-step (sm_100a).
//
// The reference evaluates scipy.special.psi (cephes/xsf digamma) at
// inferencer.py:17-18 and variational_bayes.py:177.  Here both psi(x) and the fused
// exp(psi(x) - c) are computed branch-free for x > 0:
//   * recurrence  psi(x) = psi(x+m) - sum_{i<m} 1/(x+i), with the sum folded into ONE
//     rational Q(x)/P(x), P = prod (x+i), Q = P'  (all terms positive: no cancellation);
//   * psi(y)      = log y - 1/(2y) - u p(u),      u = 1/y^2,  y = x+6  >= 6
//   * exp(psi(y)) = z + g(u)/z,                   u = 1/z^2,  z = y - 1/2, y = x+4 >= 4
//     (so the inner loop needs no log at all);
//   * a single fp64 division serves both 1/(x+m) and Q/P.
// p and g are degree-8 near-minimax fits (Chebyshev-node interpolation in 40-digit
// arithmetic, tools/fit_special.py); measured against mpmath: |psi err| <= 1e-15 max(1,|psi|),
// exp(psi) relative error <= 3 eps max(1,|psi|) -- the conditioning of exp itself.
#pragma once
#include <cuda_runtime.h>

namespace pylda {

__device__ __forceinline__ double digamma_pos(double x) {
    // P = x(x+1)...(x+5), Q = dP/dx
    double P = x, Q = 1.0;
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        const double f = x + (double)i;
        Q = fma(Q, f, P);
        P = P * f;
    }
    const double y = x + 6.0;
    const double r = 1.0 / (P * y);
    const double invy = r * P;
    const double qp = Q * (r * y);
    const double u = invy * invy;
    double p = 0x1.3d3a1cf7d5ce2p+0;
    p = fma(p, u, -0x1.78afabe0fbee1p-2);
    p = fma(p, u, 0x1.4d9027aef3165p-4);
    p = fma(p, u, -0x1.591b310557adfp-6);
    p = fma(p, u, 0x1.f077969fc7922p-8);
    p = fma(p, u, -0x1.11110af266f00p-8);
    p = fma(p, u, 0x1.041040ffe2d6dp-8);
    p = fma(p, u, -0x1.1111111110839p-7);
    p = fma(p, u, 0x1.5555555555555p-4);
    return log(y) - 0.5 * invy - u * p - qp;
}

// exp(psi(x) + negc) for x > 0.  negc = -c shifts the exponent BEFORE exp is taken so
// that tiny Dirichlet parameters (psi ~ -1/x) do not underflow against a large c.
__device__ __forceinline__ double exp_digamma_shifted(double x, double negc) {
    const double x1 = x + 1.0, x2 = x + 2.0, x3 = x + 3.0;
    const double a = x * x1, da = x + x1;
    const double b = x2 * x3, db = x2 + x3;
    const double P = a * b;
    const double Q = fma(da, b, a * db);
    const double z = x + 3.5;
    const double r = 1.0 / (P * z);
    const double invz = r * P;
    const double qp = Q * (r * z);
    const double u = invz * invz;
    double g = 0x1.72c2625e26025p-2;
    g = fma(g, u, -0x1.ab037fd41fbcdp-3);
    g = fma(g, u, 0x1.16a7995f48852p-4);
    g = fma(g, u, -0x1.4a0ddd7f70d64p-6);
    g = fma(g, u, 0x1.e1ae396a755f1p-8);
    g = fma(g, u, -0x1.0315dff6af42ap-8);
    g = fma(g, u, 0x1.d1a17ce364565p-9);
    g = fma(g, u, -0x1.a4fa4f9f36231p-8);
    g = fma(g, u, 0x1.55555555553dap-5);
    const double G = fma(invz, g, z);
    return G * exp(negc - qp);
}

// 1/x for normal positive x: MUFU.RCP64H seed (>= 20 good bits) + two Newton steps.  Relative
// error <= 2 ulp; no slow path, no branches (the IEEE division costs ~2x the instructions and
// carries a divergent fix-up branch).
__device__ __forceinline__ double rcp_nr(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double t = fma(-x, r, 1.0);
    r = fma(r, t, r);
    t = fma(-x, r, 1.0);
    r = fma(r, t, r);
    return r;
}

// Polynomial coefficients live in constant memory so that every Horner step is ONE DFMA with a
// constant-bank operand (as immediates each 64-bit coefficient costs two extra IMAD.MOV; measured:
// 18 % of all instructions of the register-tile kernel, profiles/r1b_estep_rt_W1_by_line.txt).
// Generated / checked by tools/fit_special.py.
//   c_exp: exp(r) on |r| <= ln2/2, degree 11, max relative error 1.7e-17
//   c_g  : exp(psi(y)) = z + g(u)/z, u = 1/z^2, z = y - 1/2 >= 3.5, highest degree first
__constant__ double c_exp[12] = {0x1.0000000000000p+0, 0x1.0000000000000p+0, 0x1.0000000000011p-1, 0x1.555555555555ap-3,
                                 0x1.555555554f0cfp-5, 0x1.111111110f225p-7, 0x1.6c16c187fbe02p-10, 0x1.a01a01b14378fp-13,
                                 0x1.a01991ac8730ap-16, 0x1.71ddf5749d126p-19, 0x1.28b4057f44145p-22, 0x1.af631d0059becp-26};
__constant__ double c_g[9] = {0x1.72c2625e26025p-2, -0x1.ab037fd41fbcdp-3, 0x1.16a7995f48852p-4, -0x1.4a0ddd7f70d64p-6,
                              0x1.e1ae396a755f1p-8, -0x1.0315dff6af42ap-8, 0x1.d1a17ce364565p-9, -0x1.a4fa4f9f36231p-8,
                              0x1.55555555553dap-5};

// exp(y) for y <= 0, branch-free: y = n ln2 + r, degree-11 polynomial (Estrin form: the block-level
// ncu view showed the exp(psi) code 61 % in fixed-latency dependency waits -- register pressure keeps
// ptxas from interleaving the four evaluations of a lane, so the parallelism has to be inside one),
// exponent patched in.  Results below 2^-1020 are flushed to 0 (the reference's exp() would return
// a denormal there; such e_k are 300 orders of magnitude below anything that reaches gamma, phi or
// the ELBO).
__device__ __forceinline__ double exp_nonpos(double y) {
    const double SHIFT = 6755399441055744.0;                 // 1.5 * 2^52: n lands in the low word of t
    const double t = fma(y, 0x1.71547652b82fep+0, SHIFT);
    const double n = t - SHIFT;
    double r = fma(n, -0x1.62e42fefa39efp-1, y);
    r = fma(n, -0x1.abc9e3b39803fp-56, r);
    const double r2 = r * r;
    const double a0 = fma(c_exp[1], r, c_exp[0]);
    const double a1 = fma(c_exp[3], r, c_exp[2]);
    const double a2 = fma(c_exp[5], r, c_exp[4]);
    const double a3 = fma(c_exp[7], r, c_exp[6]);
    const double a4 = fma(c_exp[9], r, c_exp[8]);
    const double a5 = fma(c_exp[11], r, c_exp[10]);
    const double r4 = r2 * r2;
    const double b0 = fma(a1, r2, a0);
    const double b1 = fma(a3, r2, a2);
    const double b2 = fma(a5, r2, a4);
    const double p = fma(fma(b2, r4, b1), r4, b0);
    const int ni = __double2loint(t);
    const double res = __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
    return y < -707.0 ? 0.0 : res;
}

// 1/x with ONE cubic Newton step on the MUFU.RCP64H seed: r (1 + e + e^2), e = 1 - x r; the seed has
// >= 20 good bits, so the result is good to 2^-60 with a 3-deep dependency chain instead of 4.
__device__ __forceinline__ double rcp_cubic(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double e2 = fma(e, e, e);
    return fma(r, e2, r);
}

// exp(psi(x)) for x > 0 (x >= 1e-300): same construction as exp_digamma_shifted with c = 0, the
// Newton reciprocal and the branch-free exp, both polynomials in Estrin form.  Underflows to 0 where
// exp(psi(x)) < 1e-307 (x <~ 1.4e-3).
__device__ __forceinline__ double exp_digamma(double x) {
    const double x1 = x + 1.0, x2 = x + 2.0, x3 = x + 3.0;
    const double a = x * x1, da = x + x1;
    const double b = x2 * x3, db = x2 + x3;
    const double P = a * b;
    const double Q = fma(da, b, a * db);
    const double z = x + 3.5;
    const double r = rcp_cubic(P * z);
    const double invz = r * P;
    const double qp = Q * (r * z);
    const double ex = exp_nonpos(-qp);
    const double u = invz * invz;
    // g(u) = sum_i c_g[8-i] u^i, degree 8
    const double u2 = u * u;
    const double g01 = fma(c_g[7], u, c_g[8]);
    const double g23 = fma(c_g[5], u, c_g[6]);
    const double g45 = fma(c_g[3], u, c_g[4]);
    const double g67 = fma(c_g[1], u, c_g[2]);
    const double u4 = u2 * u2;
    const double h0 = fma(g23, u2, g01);
    const double h1 = fma(g67, u2, g45);
    const double g = fma(fma(c_g[0], u4, h1), u4, h0);
    const double G = fma(invz, g, z);
    return G * ex;
}

// Same function with the coefficients as instruction immediates and the library exp(): more
// instructions, but no uniform-register traffic -- measured faster in the register-starved
// multi-warp register-tile kernels (estep_rt, W >= 2), slower everywhere else.
__device__ __forceinline__ double exp_digamma_imm(double x) {
    const double x1 = x + 1.0, x2 = x + 2.0, x3 = x + 3.0;
    const double a = x * x1, da = x + x1;
    const double b = x2 * x3, db = x2 + x3;
    const double P = a * b;
    const double Q = fma(da, b, a * db);
    const double z = x + 3.5;
    const double r = rcp_nr(P * z);
    const double invz = r * P;
    const double qp = Q * (r * z);
    const double u = invz * invz;
    double g = 0x1.72c2625e26025p-2;
    g = fma(g, u, -0x1.ab037fd41fbcdp-3);
    g = fma(g, u, 0x1.16a7995f48852p-4);
    g = fma(g, u, -0x1.4a0ddd7f70d64p-6);
    g = fma(g, u, 0x1.e1ae396a755f1p-8);
    g = fma(g, u, -0x1.0315dff6af42ap-8);
    g = fma(g, u, 0x1.d1a17ce364565p-9);
    g = fma(g, u, -0x1.a4fa4f9f36231p-8);
    g = fma(g, u, 0x1.55555555553dap-5);
    const double G = fma(invz, g, z);
    return G * exp(-qp);
}

// cheap stand-in for psi(x), |psi(x) - approx| < 0.12 for all x > 0; only used to pick
// the per-iteration exponent shift c (any c gives the same phi after normalisation).
__device__ __forceinline__ double digamma_rough(double x) {
    return log(x + 0.5) - 1.0 / x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace pylda
