#!/usr/bin/env python3
"""genbutterfly.py -- build-time generator of straight-line radix-r DFT
butterflies for the sm_100a Stockham kernels.

Role: this replaces, for the GPU, what the reference's OCaml `genfft` does for
its CPU codelets (reference: genfft/fft.ml:53-79 small-prime rule, :241-287
Cooley-Tukey / split-radix rules, :288-307 dispatch; genfft/algsimp.ml
algebraic simplifier; genfft/schedule.ml scheduler).  It is NOT a translation
of that code: it is a small symbolic DAG builder written for this project:

  * complex values are pairs of hash-consed real expression nodes (CSE for free)
  * constant folding: x*0, x*1, x*(-1), -(-x), multiplication by +-i is a
    swap, multiplication by (1+-i)/sqrt(2) costs 2 add + 2 mul
  * power-of-two sizes use split-radix decimation in time, odd primes use the
    symmetric (x_j +- x_{p-j}) real-matrix rule, other composites Cooley-Tukey
  * the unparser does a depth-first schedule from the outputs (keeps live
    ranges short, like genfft's bisection scheduler does for CPU registers;
    ptxas reschedules anyway) and leaves FMA contraction to nvcc (-fmad=true).

Emitted code is a set of `template<typename T> __host__ __device__ void
bflyN(T (&re)[N], T (&im)[N])` in-place forward (sign -1) butterflies on
natural-order data.  Backward transforms swap re/im at the call site, exactly
like the reference does (kernel/extract-reim.c:27-36).

Run:  python genbutterfly.py --out ../csrc/device/butterflies_gen.cuh
      python genbutterfly.py --selftest
"""
import argparse
import math
import sys

import numpy as np

RADICES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 32]


# ----------------------------------------------------------------------------
# expression DAG
# ----------------------------------------------------------------------------
class Node:
    __slots__ = ("op", "args", "val", "id")

    def __init__(self, op, args, val=None):
        self.op, self.args, self.val, self.id = op, args, val, None


class Dag:
    def __init__(self):
        self.tab = {}
        self.count = 0

    def _mk(self, op, args, val=None):
        key = (op, tuple(id(a) for a in args), val)
        n = self.tab.get(key)
        if n is None:
            n = Node(op, args, val)
            n.id = self.count
            self.count += 1
            self.tab[key] = n
        return n

    def inp(self, name):
        return self._mk("in", (), name)

    def const(self, c):
        return self._mk("const", (), float(c))

    def neg(self, a):
        if a.op == "neg":
            return a.args[0]
        if a.op == "const":
            return self.const(-a.val)
        if a.op == "sub":
            return self.sub(a.args[1], a.args[0])
        return self._mk("neg", (a,))

    def add(self, a, b):
        if a.op == "const" and a.val == 0.0:
            return b
        if b.op == "const" and b.val == 0.0:
            return a
        if b.op == "neg":
            return self.sub(a, b.args[0])
        if a.op == "neg":
            return self.sub(b, a.args[0])
        if id(a) > id(b):          # commutative canonical order -> better CSE
            a, b = b, a
        return self._mk("add", (a, b))

    def sub(self, a, b):
        if b.op == "const" and b.val == 0.0:
            return a
        if a.op == "const" and a.val == 0.0:
            return self.neg(b)
        if b.op == "neg":
            return self.add(a, b.args[0])
        if a is b:
            return self.const(0.0)
        return self._mk("sub", (a, b))

    def mulc(self, c, a):
        """constant * node"""
        if c == 0.0:
            return self.const(0.0)
        if c == 1.0:
            return a
        if c == -1.0:
            return self.neg(a)
        if a.op == "neg":
            return self.mulc(-c, a.args[0])
        if a.op == "const":
            return self.const(c * a.val)
        if c < 0:                   # keep constants positive: -c*a == neg(c*a), shares the product
            return self.neg(self._mk("mul", (self.const(-c), a)))
        return self._mk("mul", (self.const(c), a))


class Cx:
    """complex value = (re node, im node)"""
    __slots__ = ("d", "re", "im")

    def __init__(self, d, re, im):
        self.d, self.re, self.im = d, re, im

    def __add__(self, o):
        return Cx(self.d, self.d.add(self.re, o.re), self.d.add(self.im, o.im))

    def __sub__(self, o):
        return Cx(self.d, self.d.sub(self.re, o.re), self.d.sub(self.im, o.im))

    def mul_i(self, s):
        """multiply by s*i, s = +1 or -1"""
        if s > 0:
            return Cx(self.d, self.d.neg(self.im), self.re)
        return Cx(self.d, self.im, self.d.neg(self.re))

    def scale(self, c):
        return Cx(self.d, self.d.mulc(c, self.re), self.d.mulc(c, self.im))

    def mul_const(self, wr, wi):
        d = self.d
        eps = 1e-15
        if abs(wi) < eps:
            return self.scale(_snap(wr))
        if abs(wr) < eps:
            return self.scale(_snap(wi)).mul_i(+1)
        if abs(abs(wr) - abs(wi)) < eps:
            # (a+ib) * c(sr + i si), c = |wr|: 2 add + 2 mul
            c = abs(wr)
            sr, si = (1 if wr > 0 else -1), (1 if wi > 0 else -1)
            # (a + ib)(sr + i si) = (a sr - b si) + i (a si + b sr)
            a, b = self.re, self.im
            re = d.sub(d.mulc(sr, a), d.mulc(si, b))
            im = d.add(d.mulc(si, a), d.mulc(sr, b))
            return Cx(d, d.mulc(c, re), d.mulc(c, im))
        a, b = self.re, self.im
        re = d.sub(d.mulc(wr, a), d.mulc(wi, b))
        im = d.add(d.mulc(wi, a), d.mulc(wr, b))
        return Cx(d, re, im)


def _snap(x):
    for v in (0.0, 1.0, -1.0, 0.5, -0.5):
        if abs(x - v) < 1e-15:
            return v
    return x


def root(k, n):
    """exp(-2 pi i k / n) as accurately rounded doubles (octant reduction, long double)."""
    k %= n
    neg_s = neg_c = swap = False
    m, q = k, n
    if 2 * m > q:
        m, neg_s = q - m, True
    if 4 * m > q:
        m, q, neg_c = q - 2 * m, 2 * q, True
    if 8 * m > q:
        m, q, swap = q - 4 * m, 4 * q, True
    t = 2 * np.pi * np.longdouble(m) / np.longdouble(q)
    t = np.longdouble(2) * np.longdouble("3.14159265358979323846264338327950288") * np.longdouble(m) / np.longdouble(q)
    c, s = np.cos(t), np.sin(t)
    if swap:
        c, s = s, c
    if neg_c:
        c = -c
    if neg_s:
        s = -s
    return _snap(float(c)), _snap(float(-s))


# ----------------------------------------------------------------------------
# DFT rules
# ----------------------------------------------------------------------------
def smallest_factor(n):
    p = 2
    while p * p <= n:
        if n % p == 0:
            return p
        p += 1
    return n


def dft(xs):
    n = len(xs)
    if n == 1:
        return list(xs)
    if n == 2:
        return [xs[0] + xs[1], xs[0] - xs[1]]
    if n & (n - 1) == 0:
        return dft_split_radix(xs)
    p = smallest_factor(n)
    if p == n:
        return dft_prime(xs)
    return dft_ct(xs, p)


def dft_split_radix(xs):
    n = len(xs)
    if n == 4:
        a, b = xs[0] + xs[2], xs[0] - xs[2]
        c, e = xs[1] + xs[3], (xs[1] - xs[3]).mul_i(-1)
        return [a + c, b + e, a - c, b - e]
    E = dft(xs[0::2])
    O1 = dft(xs[1::4])
    O3 = dft(xs[3::4])
    q = n // 4
    out = [None] * n
    for k in range(q):
        t1 = O1[k].mul_const(*root(k, n))
        t3 = O3[k].mul_const(*root(3 * k, n))
        s = t1 + t3
        dm = (t1 - t3).mul_i(-1)
        out[k] = E[k] + s
        out[k + 2 * q] = E[k] - s
        out[k + q] = E[k + q] + dm
        out[k + 3 * q] = E[k + q] - dm
    return out


def dft_prime(xs):
    p = len(xs)
    d = xs[0].d
    h = (p - 1) // 2
    plus = [xs[j] + xs[p - j] for j in range(1, h + 1)]
    minus = [xs[j] - xs[p - j] for j in range(1, h + 1)]
    out = [None] * p
    acc = xs[0]
    for v in plus:
        acc = acc + v
    out[0] = acc
    for k in range(1, h + 1):
        a = xs[0]
        b = None
        for j in range(1, h + 1):
            c, s = root(j * k, p)          # c = cos, s = -sin
            a = a + plus[j - 1].scale(c)
            t = minus[j - 1].scale(-s)     # sin * (x_j - x_{p-j})
            b = t if b is None else b + t
        # X_k = A - i B ; X_{p-k} = A + i B
        ib = b.mul_i(+1)
        out[k] = a - ib
        out[p - k] = a + ib
    return out


def dft_ct(xs, r):
    """n = r*m decimation in time: r interleaved sub-DFTs of size m, then radix-r."""
    n = len(xs)
    m = n // r
    subs = [dft(xs[q::r]) for q in range(r)]
    out = [None] * n
    for k in range(m):
        col = [subs[q][k].mul_const(*root(q * k, n)) for q in range(r)]
        res = dft(col)
        for j in range(r):
            out[j * m + k] = res[j]
    return out


# ----------------------------------------------------------------------------
# unparser
# ----------------------------------------------------------------------------
def emit(n, fname=None):
    d = Dag()
    xs = [Cx(d, d.inp("re[%d]" % i), d.inp("im[%d]" % i)) for i in range(n)]
    ys = dft(xs)
    lines = []
    names = {}
    counter = [0]
    stats = {"add": 0, "mul": 0, "neg": 0}

    def ref(node):
        if node in names:
            return names[node]
        if node.op == "in":
            # inputs are copied to temporaries first (outputs overwrite them)
            raise AssertionError("input not preloaded")
        if node.op == "const":
            s = "T(%s)" % repr(node.val)
            names[node] = s
            return s
        args = [ref(a) for a in node.args]
        nm = "t%d" % counter[0]
        counter[0] += 1
        if node.op == "add":
            expr = "%s + %s" % (args[0], args[1]); stats["add"] += 1
        elif node.op == "sub":
            expr = "%s - %s" % (args[0], args[1]); stats["add"] += 1
        elif node.op == "mul":
            expr = "%s * %s" % (args[0], args[1]); stats["mul"] += 1
        elif node.op == "neg":
            expr = "-%s" % args[0]; stats["neg"] += 1
        else:
            raise AssertionError(node.op)
        lines.append("    const T %s = %s;" % (nm, expr))
        names[node] = nm
        return nm

    pre = []
    for i in range(n):
        pre.append("    const T xr%d = re[%d], xi%d = im[%d];" % (i, i, i, i))
        names[xs[i].re] = "xr%d" % i
        names[xs[i].im] = "xi%d" % i
    sys.setrecursionlimit(100000)
    outs = []
    for i, y in enumerate(ys):
        outs.append("    re[%d] = %s;" % (i, ref(y.re)))
        outs.append("    im[%d] = %s;" % (i, ref(y.im)))
    # interleave: emit each output assignment right after the code it needs is
    # not possible in place (later outputs may still read earlier inputs), so
    # inputs were copied to xr/xi and all stores go last.
    name = fname or ("bfly%d" % n)
    head = ("// radix-%d forward DFT, %d add/sub, %d mul, %d neg (before nvcc FMA contraction)\n"
            "template <typename T>\n__host__ __device__ __forceinline__ void %s(T (&re)[%d], T (&im)[%d])\n{"
            % (n, stats["add"], stats["mul"], stats["neg"], name, n, n))
    return "\n".join([head] + pre + lines + outs + ["}"]) + "\n", stats


def evaluate(n, x):
    """numerically evaluate the generated DAG (self-test)"""
    d = Dag()
    xs = [Cx(d, d.inp(("re", i)), d.inp(("im", i))) for i in range(n)]
    ys = dft(xs)
    memo = {}

    def ev(node):
        if node in memo:
            return memo[node]
        if node.op == "in":
            kind, i = node.val
            v = x[i].real if kind == "re" else x[i].imag
        elif node.op == "const":
            v = node.val
        else:
            a = [ev(t) for t in node.args]
            v = {"add": lambda: a[0] + a[1], "sub": lambda: a[0] - a[1],
                 "mul": lambda: a[0] * a[1], "neg": lambda: -a[0]}[node.op]()
        memo[node] = v
        return v

    sys.setrecursionlimit(100000)
    return np.array([ev(y.re) + 1j * ev(y.im) for y in ys])


def selftest():
    rng = np.random.default_rng(0)
    ok = True
    for n in RADICES + [14, 15, 17, 20, 25, 64]:
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        y = evaluate(n, x)
        err = np.linalg.norm(y - np.fft.fft(x)) / np.linalg.norm(y)
        _, st = emit(n)
        print("radix %3d  rel err %.2e   add %4d mul %4d neg %3d" % (n, err, st["add"], st["mul"], st["neg"]))
        ok &= err < 1e-14
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    ap.add_argument("--selftest", action="store_true")
    a = ap.parse_args()
    if a.selftest:
        sys.exit(0 if selftest() else 1)
    parts = ["// GENERATED by fftw3_b200/gen/genbutterfly.py -- do not edit.\n"
             "// Straight-line forward (sign -1) radix-r DFT butterflies, in place, natural order.\n"
             "#pragma once\n"
             "#ifndef __CUDACC__\n#ifndef __host__\n#define __host__\n#endif\n#ifndef __device__\n#define __device__\n#endif\n"
             "#ifndef __forceinline__\n#define __forceinline__ inline\n#endif\n#endif\n"]
    for n in RADICES:
        src, _ = emit(n)
        parts.append(src)
    # compile-time roots of unity W_E^m = exp(-2 pi i m / E), used by the specialised
    # kernels to derive per-butterfly twiddles from one loaded base twiddle
    for e in (4, 8, 10, 16, 32):
        cs = [root(m, e) for m in range(e)]
        cases = "\n".join("    case %d: re = T(%s); im = T(%s); break;" % (m, repr(c[0]), repr(c[1]))
                          for m, c in enumerate(cs))
        parts.append("template <typename T>\n__host__ __device__ __forceinline__ void unit_root%d(int m, T &re, T &im)\n"
                     "{\n    switch (m) {\n%s\n    default: re = T(1.0); im = T(0.0); break;\n    }\n}\n" % (e, cases))
    # dispatcher
    parts.append("template <int R, typename T> struct Butterfly;\n")
    for n in RADICES:
        parts.append("template <typename T> struct Butterfly<%d, T> { static __host__ __device__ __forceinline__ "
                     "void run(T (&re)[%d], T (&im)[%d]) { bfly%d(re, im); } };\n" % (n, n, n, n))
    text = "\n".join(parts)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
