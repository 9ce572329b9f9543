#!/usr/bin/env python3
"""Generate straight-line fp64 register FFT codelets (spheremanopt_b200/csrc/codelets.cuh).

Each codelet ``fftR_<dir>(double (&xr)[R], double (&xi)[R])`` is an in-place, natural-order
complex DFT of length R held entirely in registers: X[k] = sum_n x[n] exp(-+2*pi*i*n*k/R)
(``fwd`` = minus sign, ``inv`` = plus sign, both unnormalised).  They are the leaves of the
shared-memory multi-stage FFT passes (fft_pass.cuh, xpass.cuh, sh23.cuh).

The generator does a decimation-in-time Cooley-Tukey recursion down to radix 2/3/4/5 butterflies
at generation time, so that all index permutations vanish and twiddles that are 1, -1, +-i or
(+-1+-i)/sqrt(2) cost no multiplications.  Twiddle constants are emitted as 17-significant-digit
literals computed in extended precision.

Usage: python tools/gen_codelets.py > spheremanopt_b200/csrc/codelets.cuh
"""
import sys
from fractions import Fraction
import mpmath

mpmath.mp.dps = 40

SIZES = [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 18, 20, 24]
FACTOR = {6: (3, 2), 8: (4, 2), 9: (3, 3), 10: (5, 2), 12: (4, 3), 15: (5, 3), 16: (4, 4), 18: (3, 6), 20: (4, 5), 24: (4, 6)}


class Emit:
    def __init__(self):
        self.lines = []
        self.n = 0

    def tmp(self):
        self.n += 1
        return "t%d" % self.n

    def let(self, expr):
        v = self.tmp()
        self.lines.append("  const double %s = %s;" % (v, expr))
        return v


def lit(x):
    return mpmath.nstr(x, 17, min_fixed=-1, max_fixed=1, strip_zeros=False).replace("e", "e")


def cmul_const(E, a, frac, sign):
    """multiply complex var pair a=(re,im) by exp(sign*2*pi*i*frac), frac a Fraction in [0,1)"""
    frac = frac % 1
    re, im = a
    if frac == 0:
        return a
    if frac == Fraction(1, 2):
        return (E.let("-%s" % re), E.let("-%s" % im))
    # exp(sign*i*pi/2) = sign*i ;  (re+i im)*(i) = -im + i re
    if frac == Fraction(1, 4):
        if sign > 0:
            return (E.let("-%s" % im), re)
        return (im, E.let("-%s" % re))
    if frac == Fraction(3, 4):
        if sign > 0:
            return (im, E.let("-%s" % re))
        return (E.let("-%s" % im), re)
    ang = 2 * mpmath.pi * mpmath.mpf(frac.numerator) / frac.denominator
    c = mpmath.cos(ang)
    s = mpmath.sin(ang) * sign
    if frac.denominator == 8:
        # |c| == |s| == 1/sqrt(2)
        h = lit(mpmath.sqrt(mpmath.mpf(1) / 2))
        sc = 1 if c > 0 else -1
        ss = 1 if s > 0 else -1
        # (re + i im)(sc + i ss) h = h*((sc re - ss im) + i (ss re + sc im))
        r_expr = "%s(%s%s %s %s)" % ("%s*" % h, "" if sc > 0 else "-", re, "-" if ss > 0 else "+", im)
        i_expr = "%s(%s%s %s %s)" % ("%s*" % h, "" if ss > 0 else "-", re, "+" if sc > 0 else "-", im)
        return (E.let(r_expr), E.let(i_expr))
    cl, sl = lit(c), lit(s)
    return (E.let("%s*(%s) - %s*(%s)" % (re, cl, im, sl)), E.let("%s*(%s) + %s*(%s)" % (re, sl, im, cl)))


def add(E, a, b):
    return (E.let("%s + %s" % (a[0], b[0])), E.let("%s + %s" % (a[1], b[1])))


def sub(E, a, b):
    return (E.let("%s - %s" % (a[0], b[0])), E.let("%s - %s" % (a[1], b[1])))


def mul_i(E, a, sign):
    """a * (sign*i)"""
    if sign > 0:
        return (E.let("-%s" % a[1]), a[0])
    return (a[1], E.let("-%s" % a[0]))


def bf2(E, x, sign):
    return [add(E, x[0], x[1]), sub(E, x[0], x[1])]


def bf4(E, x, sign):
    a = add(E, x[0], x[2]); b = sub(E, x[0], x[2])
    c = add(E, x[1], x[3]); d = sub(E, x[1], x[3])
    di = mul_i(E, d, sign)           # forward: -i d ; inverse: +i d
    return [add(E, a, c), add(E, b, di), sub(E, a, c), sub(E, b, di)]


def bf3(E, x, sign):
    s = add(E, x[1], x[2]); d = sub(E, x[1], x[2])
    c3 = lit(mpmath.mpf(-1) / 2)
    s3 = lit(mpmath.sqrt(3) / 2)
    X0 = add(E, x[0], s)
    m = (E.let("%s + (%s)*%s" % (x[0][0], c3, s[0])), E.let("%s + (%s)*%s" % (x[0][1], c3, s[1])))
    # sign*i*s3*d
    if sign > 0:
        e = (E.let("-(%s*%s)" % (s3, d[1])), E.let("%s*%s" % (s3, d[0])))
    else:
        e = (E.let("%s*%s" % (s3, d[1])), E.let("-(%s*%s)" % (s3, d[0])))
    return [X0, add(E, m, e), sub(E, m, e)]


def bf5(E, x, sign):
    c1 = mpmath.cos(2 * mpmath.pi / 5); c2 = mpmath.cos(4 * mpmath.pi / 5)
    s1 = mpmath.sin(2 * mpmath.pi / 5); s2 = mpmath.sin(4 * mpmath.pi / 5)
    a1 = add(E, x[1], x[4]); b1 = sub(E, x[1], x[4])
    a2 = add(E, x[2], x[3]); b2 = sub(E, x[2], x[3])
    X0 = (E.let("%s + %s + %s" % (x[0][0], a1[0], a2[0])), E.let("%s + %s + %s" % (x[0][1], a1[1], a2[1])))
    m1 = (E.let("%s + (%s)*%s + (%s)*%s" % (x[0][0], lit(c1), a1[0], lit(c2), a2[0])),
          E.let("%s + (%s)*%s + (%s)*%s" % (x[0][1], lit(c1), a1[1], lit(c2), a2[1])))
    m2 = (E.let("%s + (%s)*%s + (%s)*%s" % (x[0][0], lit(c2), a1[0], lit(c1), a2[0])),
          E.let("%s + (%s)*%s + (%s)*%s" % (x[0][1], lit(c2), a1[1], lit(c1), a2[1])))
    # n1 = s1*b1 + s2*b2 ; n2 = s2*b1 - s1*b2 ; X1 = m1 + sign*i*n1 ; X4 = m1 - sign*i*n1 ; X2 = m2 + sign*i*n2 ; X3 = m2 - ...
    n1 = (E.let("(%s)*%s + (%s)*%s" % (lit(s1), b1[0], lit(s2), b2[0])), E.let("(%s)*%s + (%s)*%s" % (lit(s1), b1[1], lit(s2), b2[1])))
    n2 = (E.let("(%s)*%s - (%s)*%s" % (lit(s2), b1[0], lit(s1), b2[0])), E.let("(%s)*%s - (%s)*%s" % (lit(s2), b1[1], lit(s1), b2[1])))
    e1 = mul_i(E, n1, sign); e2 = mul_i(E, n2, sign)
    return [X0, add(E, m1, e1), add(E, m2, e2), sub(E, m2, e2), sub(E, m1, e1)]


BASE = {2: bf2, 3: bf3, 4: bf4, 5: bf5}


def fft(E, x, sign):
    N = len(x)
    if N == 1:
        return x
    if N in BASE:
        return BASE[N](E, x, sign)
    N1, N2 = FACTOR[N]
    # n = n2 + N2*n1 ; k = k1 + N1*k2
    Y = []
    for n2 in range(N2):
        sub_in = [x[n2 + N2 * n1] for n1 in range(N1)]
        y = fft(E, sub_in, sign)
        y = [cmul_const(E, y[k1], Fraction(n2 * k1, N), sign) for k1 in range(N1)]
        Y.append(y)
    out = [None] * N
    for k1 in range(N1):
        z = fft(E, [Y[n2][k1] for n2 in range(N2)], sign)
        for k2 in range(N2):
            out[k1 + N1 * k2] = z[k2]
    return out


def gen(N, sign):
    E = Emit()
    x = [("xr[%d]" % n, "xi[%d]" % n) for n in range(N)]
    # load into temporaries first so that in-place stores cannot alias
    xin = [(E.let(a), E.let(b)) for a, b in x]
    out = fft(E, xin, sign)
    name = "fft%d_%s" % (N, "inv" if sign > 0 else "fwd")
    s = ["SMO_HD void %s(double (&xr)[%d], double (&xi)[%d]) {" % (name, N, N)]
    s += E.lines
    for k in range(N):
        s.append("  xr[%d] = %s; xi[%d] = %s;" % (k, out[k][0], k, out[k][1]))
    s.append("}")
    return "\n".join(s)


def main():
    print("// GENERATED by tools/gen_codelets.py - do not edit.  Register FFT codelets (fp64, unnormalised).")
    print("// fwd: X[k] = sum_n x[n] exp(-2 pi i n k / R);  inv: exp(+2 pi i n k / R).")
    print("#pragma once")
    print('#include "smo_common.cuh"')
    print("namespace smo {")
    for N in SIZES:
        print(gen(N, -1))
        print(gen(N, +1))
    print("// DIR = -1 forward, +1 inverse")
    print("template <int R, int DIR> struct RegFFT;")
    for N in SIZES:
        print("template <> struct RegFFT<%d, -1> { SMO_HD static void run(double (&xr)[%d], double (&xi)[%d]) { fft%d_fwd(xr, xi); } };" % (N, N, N, N))
        print("template <> struct RegFFT<%d, +1> { SMO_HD static void run(double (&xr)[%d], double (&xi)[%d]) { fft%d_inv(xr, xi); } };" % (N, N, N, N))
    print("}  // namespace smo")


if __name__ == "__main__":
    main()
