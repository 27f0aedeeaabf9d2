#!/usr/bin/env python
"""Coefficient generator / checker for csrc/special.cuh (needs mpmath; build-time tool only).

  exp(r) on |r| <= ln2/2            degree-11 interpolant at Chebyshev nodes (c_exp in special.cuh)
  g(u)  : exp(psi(y)) = z + g(u)/z  u = 1/z^2, z = y - 1/2 >= 3.5  (checked, not re-fitted)

Prints C initialisers and the maximum relative error measured against mpmath on a dense grid.
"""
import mpmath as mp

mp.mp.dps = 50


def cheb_interp(f, a, b, deg):
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    return [c for c in mp.lu_solve(A, y)]


def horner(cs, x):
    r = mp.mpf(0)
    for c in reversed(cs):
        r = r * x + c
    return r


def main():
    a = mp.log(2) / 2
    cs = cheb_interp(mp.exp, -a, a, 11)
    cs_d = [mp.mpf(float(c)) for c in cs]          # rounded to double
    err = max(abs(horner(cs_d, x) / mp.exp(x) - 1) for x in mp.linspace(-a, a, 4001))
    print("// exp(r), |r| <= ln2/2, degree 11, max rel err %s" % mp.nstr(err, 3))
    print("__constant__ double c_exp[12] = {%s};" % ", ".join(float(c).hex() for c in cs))

    g = [float.fromhex(h) for h in (
        "0x1.55555555553dap-5", "-0x1.a4fa4f9f36231p-8", "0x1.d1a17ce364565p-9", "-0x1.0315dff6af42ap-8",
        "0x1.e1ae396a755f1p-8", "-0x1.4a0ddd7f70d64p-6", "0x1.16a7995f48852p-4", "-0x1.ab037fd41fbcdp-3",
        "0x1.72c2625e26025p-2")]
    worst = 0
    for y in mp.linspace(4, 200, 3000):
        z = y - mp.mpf(1) / 2
        u = 1 / z ** 2
        approx = z + horner([mp.mpf(c) for c in g], u) / z
        worst = max(worst, abs(approx / mp.exp(mp.digamma(y)) - 1))
    print("// exp(psi(y)) = z + g(u)/z on y in [4, 200]: max rel err %s" % mp.nstr(worst, 3))


if __name__ == "__main__":
    main()
