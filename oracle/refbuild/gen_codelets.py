#!/usr/bin/env python3
"""gen_codelets.py -- TEST INFRASTRUCTURE ONLY (oracle side, never part of the product).

The reference's generated codelets are not in its git tree and genfft needs OCaml, which this image lacks, so
oracle/_ref used to be built with EMPTY codelet tables: correct but ~65x slower than a real FFTW, which made
the CPU arm of bench.py a weak baseline.  This script emits, with THIS project's own butterfly
generator (fftw3_b200/gen/genbutterfly.py -- not genfft), scalar C codelets in the reference's codelet ABI:

    n1_R   no-twiddle DFT of size R      kdft   (dft/codelet-dft.h:59-61, genus dft/scalar/n.c)
    t1_R   DIT twiddle pass of radix R   kdftw  (dft/codelet-dft.h:83-87, genus dft/scalar/t.c; twiddles
           {TW_FULL, 0, R}: W[2(k-1)], W[2(k-1)+1] = cos, sin(2 pi j k / n), input k times conj(W_k),
           kernel/twiddle.c:142-151)

for R in 2,3,4,5,6,7,8,9,10,11,12,13,16,32 and the solvtab X(solvtab_dft_standard) registering them.  The
reference's planner, Cooley-Tukey solvers (dft/ct.c, dft/dftw-direct.c, dft/direct.c), threads and API stay
exactly as they are; only the leaves are ours.  Constants are doubles: the long-double build keeps the empty table
(it is the high-precision pin of oracle/oracle_dft.c and must not be limited by 53-bit constants).

    python oracle/refbuild/gen_codelets.py > oracle/refbuild/codelets_gen.c
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "fftw3_b200", "gen"))
import genbutterfly as G  # noqa: E402

RADICES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 32]


def mul(d, a, b):
    """node * node (the twiddle products; genbutterfly only multiplies by constants)"""
    return d._mk("mul", (a, b))


def unparse(outs, names):
    """depth-first straight-line C from output nodes; `names` maps input nodes to C lvalues already loaded"""
    lines, counter, stats = [], [0], {"add": 0, "mul": 0, "neg": 0}

    def ref(node):
        if node in names:
            return names[node]
        if node.op == "const":
            s = "((E) %s)" % repr(node.val)
            names[node] = s
            return s
        assert node.op != "in", "input not preloaded"
        args = [ref(a) for a in node.args]
        nm = "T%d" % counter[0]
        counter[0] += 1
        if node.op == "add":
            e = "%s + %s" % tuple(args); stats["add"] += 1
        elif node.op == "sub":
            e = "%s - %s" % tuple(args); stats["add"] += 1
        elif node.op == "mul":
            e = "%s * %s" % tuple(args); stats["mul"] += 1
        else:
            e = "-%s" % args[0]; stats["neg"] += 1
        lines.append("\t  const E %s = %s;" % (nm, e))
        names[node] = nm
        return nm

    sys.setrecursionlimit(100000)
    refs = [ref(o) for o in outs]
    return lines, refs, stats


def n1(r):
    d = G.Dag()
    xs = [G.Cx(d, d.inp(("re", i)), d.inp(("im", i))) for i in range(r)]
    ys = G.dft(xs)
    names, pre = {}, []
    for i in range(r):
        pre.append("\t  const E xr%d = ri[WS(is, %d)], xi%d = ii[WS(is, %d)];" % (i, i, i, i))
        names[xs[i].re], names[xs[i].im] = "xr%d" % i, "xi%d" % i
    lines, refs, st = unparse([v for y in ys for v in (y.re, y.im)], names)
    post = []
    for i in range(r):
        post.append("\t  ro[WS(os, %d)] = %s;" % (i, refs[2 * i]))
        post.append("\t  io[WS(os, %d)] = %s;" % (i, refs[2 * i + 1]))
    body = "\n".join(pre + lines + post)
    return """
static void n1_%(r)d(const R *ri, const R *ii, R *ro, R *io, stride is, stride os, INT v, INT ivs, INT ovs)
{
     INT i;
     for (i = v; i > 0; --i, ri += ivs, ii += ivs, ro += ovs, io += ovs) {
%(body)s
     }
}
static const kdft_desc desc_n1_%(r)d = { %(r)d, "n1_%(r)d", { %(add)d, %(mul)d, 0, 0 }, &X(dft_n_genus), 0, 0, 0, 0 };
static void reg_n1_%(r)d(planner *p) { X(kdft_register)(p, n1_%(r)d, &desc_n1_%(r)d); }
""" % {"r": r, "body": body, "add": st["add"], "mul": st["mul"]}


def t1(r):
    d = G.Dag()
    raw = [G.Cx(d, d.inp(("re", i)), d.inp(("im", i))) for i in range(r)]
    tw = [None] + [(d.inp(("wr", i)), d.inp(("wi", i))) for i in range(1, r)]
    xs = [raw[0]]
    for k in range(1, r):
        wr, wi = tw[k]
        xr, xi = raw[k].re, raw[k].im
        # x * conj(w) = (xr wr + xi wi) + i (xi wr - xr wi)
        xs.append(G.Cx(d, d.add(mul(d, xr, wr), mul(d, xi, wi)), d.sub(mul(d, xi, wr), mul(d, xr, wi))))
    ys = G.dft(xs)
    names, pre = {}, []
    for i in range(r):
        pre.append("\t  const E xr%d = ri[WS(rs, %d)], xi%d = ii[WS(rs, %d)];" % (i, i, i, i))
        names[raw[i].re], names[raw[i].im] = "xr%d" % i, "xi%d" % i
    for k in range(1, r):
        pre.append("\t  const E wr%d = W[%d], wi%d = W[%d];" % (k, 2 * (k - 1), k, 2 * (k - 1) + 1))
        names[tw[k][0]], names[tw[k][1]] = "wr%d" % k, "wi%d" % k
    lines, refs, st = unparse([v for y in ys for v in (y.re, y.im)], names)
    post = []
    for i in range(r):
        post.append("\t  ri[WS(rs, %d)] = %s;" % (i, refs[2 * i]))
        post.append("\t  ii[WS(rs, %d)] = %s;" % (i, refs[2 * i + 1]))
    body = "\n".join(pre + lines + post)
    return """
static void t1_%(r)d(R *ri, R *ii, const R *W, stride rs, INT mb, INT me, INT ms)
{
     INT m;
     for (m = mb, W = W + mb * %(tw)d; m < me; ++m, ri += ms, ii += ms, W += %(tw)d) {
%(body)s
     }
}
static const tw_instr twinstr_t1_%(r)d[] = { { TW_FULL, 0, %(r)d }, { TW_NEXT, 1, 0 } };
static const ct_desc desc_t1_%(r)d = { %(r)d, "t1_%(r)d", twinstr_t1_%(r)d, &X(dft_t_genus), { %(add)d, %(mul)d, 0, 0 }, 0, 0, 0 };
static void reg_t1_%(r)d(planner *p) { X(kdft_dit_register)(p, t1_%(r)d, &desc_t1_%(r)d); }
""" % {"r": r, "body": body, "tw": 2 * (r - 1), "add": st["add"], "mul": st["mul"]}


def main():
    out = ["""/* GENERATED by oracle/refbuild/gen_codelets.py -- TEST INFRASTRUCTURE ONLY, do not edit.
 * Scalar DFT codelets in the reference's codelet ABI, emitted by this project's own butterfly generator, plus
 * the solver tables the reference's conf.c files expect (dft/conf.c:43, rdft/conf.c:57-59).  The real-data
 * tables stay empty (rdft leaves are not generated); the long-double build keeps all tables empty. */
#include "kernel/ifftw.h"
#include "dft/codelet-dft.h"
#include "dft/scalar/n.h"
#undef GENUS
#include "dft/scalar/t.h"
#undef GENUS
extern const solvtab X(solvtab_rdft_r2cf);
extern const solvtab X(solvtab_rdft_r2cb);
extern const solvtab X(solvtab_rdft_r2r);
const solvtab X(solvtab_rdft_r2cf) = { SOLVTAB_END };
const solvtab X(solvtab_rdft_r2cb) = { SOLVTAB_END };
const solvtab X(solvtab_rdft_r2r) = { SOLVTAB_END };
#if defined(FFTW_LDOUBLE) || defined(FFTW_QUAD)
const solvtab X(solvtab_dft_standard) = { SOLVTAB_END };
#else
"""]
    for r in RADICES:
        out.append(n1(r))
    for r in RADICES:
        out.append(t1(r))
    out.append("const solvtab X(solvtab_dft_standard) = {")
    for r in RADICES:
        out.append("     SOLVTAB(reg_n1_%d)," % r)
    for r in RADICES:
        out.append("     SOLVTAB(reg_t1_%d)," % r)
    out.append("     SOLVTAB_END\n};\n#endif")
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
