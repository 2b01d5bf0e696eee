"""Generates ``streamoptima_b200/csrc/so_dct_ducc.cuh``: straight-line FP64 DCT-II / DCT-III of length 2, 4, 8, 16
whose floating-point operation order reproduces ``scipy.fftpack.dct/idct(norm='ortho')`` bit for bit.

Why: the reference rounds the float64 transform output to integers (Encoder.py:783, 815).  About one residual block
in ten has a coefficient that is exactly k+0.5 in real arithmetic; which side SciPy lands on depends on
its rounding errors, so only the *same sequence of IEEE operations* gives the same levels.

What is reproduced (established empirically against SciPy 1.18.1 / ``_duccfft``; 100 % bit-equal on random int and
float inputs, see tests/test_dct_model.py):
  * the DCT-II/III <-> real-FFT reduction of pocketfft/ducc0 (``T_dcst23``): pre/post butterflies, a *backward*
    real FFT for type 2 and a *forward* one for type 3, ``fct = sqrt(1/(2N))`` applied to every output of the FFT,
    ``c[0] *= sqrt2*0.5`` (type 2) / ``c[0] *= sqrt2`` (type 3);
  * real-FFT passes radix 4 / radix 2 in FFTPACK order, factor list {2:[2], 4:[4], 8:[2,4], 16:[4,4]};
  * twiddles from ducc0's ``UnityRoots`` (two-table product of libm sin/cos of ``x * (0.25*pi/n)`` in double), which
    are NOT correctly rounded (e.g. cos(pi/4) comes out one ulp low) -- they are evaluated here with Python's ``math``
    (the same libm) and embedded as hex-float literals;
  * no FMA contraction anywhere (the SciPy wheel targets baseline x86-64): the CUDA code uses __dadd_rn/__dsub_rn/
    __dmul_rn, which the compiler never fuses.

The program is obtained by *tracing* a direct Python transcription of the algorithm with symbolic operands, so the
emitted code has exactly the dataflow that was validated.  usage: python tools/dctgen/gen_dct.py
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from roots import UnityRoots  # noqa: E402

SQRT2 = 1.414213562373095048801688724209698
HSQT2 = 0.707106781186547524400844362104849

_roots = {}


def cs(k, n):
    if n not in _roots:
        _roots[n] = UnityRoots(n)
    return _roots[n][k]


def factorize(n):
    f = []
    while n % 4 == 0:
        f.append(4)
        n //= 4
    if n % 2 == 0:
        n //= 2
        f.append(2)
        f[0], f[-1] = f[-1], f[0]
    assert n == 1
    return f


def twiddles(N, facts):
    tws = []
    l1 = 1
    for k, ip in enumerate(facts):
        ido = N // (l1 * ip)
        tw = None
        if k < len(facts) - 1:
            tw = [0.0] * ((ip - 1) * (ido - 1))
            for j in range(1, ip):
                for i in range(1, (ido - 1) // 2 + 1):
                    c, s = cs(j * l1 * i, N)
                    tw[(j - 1) * (ido - 1) + 2 * i - 2] = c
                    tw[(j - 1) * (ido - 1) + 2 * i - 1] = s
        tws.append(tw)
        l1 *= ip
    return tws


# ---- the algorithm, written once over generic operands (floats for checking, Sym for code generation) ----------
def mulpm(c, d, e, f):
    return c * e + d * f, c * f - d * e


def radf2(ido, l1, cc, wa):
    ch = [0.0] * len(cc)
    CC = lambda a, b, c: cc[a + ido * (b + l1 * c)]
    CH = lambda a, b, c: a + ido * (b + 2 * c)
    WA = lambda x, i: wa[i + x * (ido - 1)]
    for k in range(l1):
        ch[CH(0, 0, k)] = CC(0, k, 0) + CC(0, k, 1)
        ch[CH(ido - 1, 1, k)] = CC(0, k, 0) - CC(0, k, 1)
    if ido % 2 == 0:
        for k in range(l1):
            ch[CH(0, 1, k)] = -CC(ido - 1, k, 1)
            ch[CH(ido - 1, 0, k)] = CC(ido - 1, k, 0)
    if ido <= 2:
        return ch
    for k in range(l1):
        for i in range(2, ido, 2):
            ic = ido - i
            tr2, ti2 = mulpm(WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
            ch[CH(i - 1, 0, k)] = CC(i - 1, k, 0) + tr2
            ch[CH(ic - 1, 1, k)] = CC(i - 1, k, 0) - tr2
            ch[CH(i, 0, k)] = ti2 + CC(i, k, 0)
            ch[CH(ic, 1, k)] = ti2 - CC(i, k, 0)
    return ch


def radf4(ido, l1, cc, wa):
    ch = [0.0] * len(cc)
    CC = lambda a, b, c: cc[a + ido * (b + l1 * c)]
    CH = lambda a, b, c: a + ido * (b + 4 * c)
    WA = lambda x, i: wa[i + x * (ido - 1)]
    for k in range(l1):
        tr1 = CC(0, k, 3) + CC(0, k, 1)
        ch[CH(0, 2, k)] = CC(0, k, 3) - CC(0, k, 1)
        tr2 = CC(0, k, 0) + CC(0, k, 2)
        ch[CH(ido - 1, 1, k)] = CC(0, k, 0) - CC(0, k, 2)
        ch[CH(0, 0, k)] = tr2 + tr1
        ch[CH(ido - 1, 3, k)] = tr2 - tr1
    if ido % 2 == 0:
        for k in range(l1):
            ti1 = (-HSQT2) * (CC(ido - 1, k, 1) + CC(ido - 1, k, 3))
            tr1 = HSQT2 * (CC(ido - 1, k, 1) - CC(ido - 1, k, 3))
            ch[CH(ido - 1, 0, k)] = CC(ido - 1, k, 0) + tr1
            ch[CH(ido - 1, 2, k)] = CC(ido - 1, k, 0) - tr1
            ch[CH(0, 3, k)] = ti1 + CC(ido - 1, k, 2)
            ch[CH(0, 1, k)] = ti1 - CC(ido - 1, k, 2)
    if ido <= 2:
        return ch
    for k in range(l1):
        for i in range(2, ido, 2):
            ic = ido - i
            cr2, ci2 = mulpm(WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
            cr3, ci3 = mulpm(WA(1, i - 2), WA(1, i - 1), CC(i - 1, k, 2), CC(i, k, 2))
            cr4, ci4 = mulpm(WA(2, i - 2), WA(2, i - 1), CC(i - 1, k, 3), CC(i, k, 3))
            tr1, tr4 = cr4 + cr2, cr4 - cr2
            ti1, ti4 = ci2 + ci4, ci2 - ci4
            tr2, tr3 = CC(i - 1, k, 0) + cr3, CC(i - 1, k, 0) - cr3
            ti2, ti3 = CC(i, k, 0) + ci3, CC(i, k, 0) - ci3
            ch[CH(i - 1, 0, k)] = tr2 + tr1
            ch[CH(ic - 1, 3, k)] = tr2 - tr1
            ch[CH(i, 0, k)] = ti1 + ti2
            ch[CH(ic, 3, k)] = ti1 - ti2
            ch[CH(i - 1, 2, k)] = tr3 + ti4
            ch[CH(ic - 1, 1, k)] = tr3 - ti4
            ch[CH(i, 2, k)] = tr4 + ti3
            ch[CH(ic, 1, k)] = tr4 - ti3
    return ch


def radb2(ido, l1, cc, wa):
    ch = [0.0] * len(cc)
    CC = lambda a, b, c: cc[a + ido * (b + 2 * c)]
    CH = lambda a, b, c: a + ido * (b + l1 * c)
    WA = lambda x, i: wa[i + x * (ido - 1)]
    for k in range(l1):
        ch[CH(0, k, 0)] = CC(0, 0, k) + CC(ido - 1, 1, k)
        ch[CH(0, k, 1)] = CC(0, 0, k) - CC(ido - 1, 1, k)
    if ido % 2 == 0:
        for k in range(l1):
            ch[CH(ido - 1, k, 0)] = 2.0 * CC(ido - 1, 0, k)
            ch[CH(ido - 1, k, 1)] = (-2.0) * CC(0, 1, k)
    if ido <= 2:
        return ch
    for k in range(l1):
        for i in range(2, ido, 2):
            ic = ido - i
            ch[CH(i - 1, k, 0)] = CC(i - 1, 0, k) + CC(ic - 1, 1, k)
            tr2 = CC(i - 1, 0, k) - CC(ic - 1, 1, k)
            ti2 = CC(i, 0, k) + CC(ic, 1, k)
            ch[CH(i, k, 0)] = CC(i, 0, k) - CC(ic, 1, k)
            a, b = mulpm(WA(0, i - 2), WA(0, i - 1), ti2, tr2)
            ch[CH(i, k, 1)] = a
            ch[CH(i - 1, k, 1)] = b
    return ch


def radb4(ido, l1, cc, wa):
    ch = [0.0] * len(cc)
    CC = lambda a, b, c: cc[a + ido * (b + 4 * c)]
    CH = lambda a, b, c: a + ido * (b + l1 * c)
    WA = lambda x, i: wa[i + x * (ido - 1)]
    for k in range(l1):
        tr2, tr1 = CC(0, 0, k) + CC(ido - 1, 3, k), CC(0, 0, k) - CC(ido - 1, 3, k)
        tr3 = 2.0 * CC(ido - 1, 1, k)
        tr4 = 2.0 * CC(0, 2, k)
        ch[CH(0, k, 0)] = tr2 + tr3
        ch[CH(0, k, 2)] = tr2 - tr3
        ch[CH(0, k, 3)] = tr1 + tr4
        ch[CH(0, k, 1)] = tr1 - tr4
    if ido % 2 == 0:
        for k in range(l1):
            ti1, ti2 = CC(0, 3, k) + CC(0, 1, k), CC(0, 3, k) - CC(0, 1, k)
            tr2, tr1 = CC(ido - 1, 0, k) + CC(ido - 1, 2, k), CC(ido - 1, 0, k) - CC(ido - 1, 2, k)
            ch[CH(ido - 1, k, 0)] = tr2 + tr2
            ch[CH(ido - 1, k, 1)] = SQRT2 * (tr1 - ti1)
            ch[CH(ido - 1, k, 2)] = ti2 + ti2
            ch[CH(ido - 1, k, 3)] = (-SQRT2) * (tr1 + ti1)
    if ido <= 2:
        return ch
    for k in range(l1):
        for i in range(2, ido, 2):
            ic = ido - i
            tr2, tr1 = CC(i - 1, 0, k) + CC(ic - 1, 3, k), CC(i - 1, 0, k) - CC(ic - 1, 3, k)
            ti1, ti2 = CC(i, 0, k) + CC(ic, 3, k), CC(i, 0, k) - CC(ic, 3, k)
            tr4, ti3 = CC(i, 2, k) + CC(ic, 1, k), CC(i, 2, k) - CC(ic, 1, k)
            tr3, ti4 = CC(i - 1, 2, k) + CC(ic - 1, 1, k), CC(i - 1, 2, k) - CC(ic - 1, 1, k)
            ch[CH(i - 1, k, 0)] = tr2 + tr3
            cr3 = tr2 - tr3
            ch[CH(i, k, 0)] = ti2 + ti3
            ci3 = ti2 - ti3
            cr4, cr2 = tr1 + tr4, tr1 - tr4
            ci2, ci4 = ti1 + ti4, ti1 - ti4
            a, b = mulpm(WA(0, i - 2), WA(0, i - 1), ci2, cr2)
            ch[CH(i, k, 1)], ch[CH(i - 1, k, 1)] = a, b
            a, b = mulpm(WA(1, i - 2), WA(1, i - 1), ci3, cr3)
            ch[CH(i, k, 2)], ch[CH(i - 1, k, 2)] = a, b
            a, b = mulpm(WA(2, i - 2), WA(2, i - 1), ci4, cr4)
            ch[CH(i, k, 3)], ch[CH(i - 1, k, 3)] = a, b
    return ch


def rfft_forward(c, N, facts, tws, fct):
    p = list(c)
    l1 = N
    for k1 in range(len(facts)):
        k = len(facts) - k1 - 1
        ip = facts[k]
        ido = N // l1
        l1 //= ip
        p = radf4(ido, l1, p, tws[k]) if ip == 4 else radf2(ido, l1, p, tws[k])
    return [v * fct for v in p]


def rfft_backward(c, N, facts, tws, fct):
    p = list(c)
    l1 = 1
    for k, ip in enumerate(facts):
        ido = N // (ip * l1)
        p = radb4(ido, l1, p, tws[k]) if ip == 4 else radb2(ido, l1, p, tws[k])
        l1 *= ip
    return [v * fct for v in p]


def plan(N):
    facts = factorize(N)
    tws = twiddles(N, facts)
    tw = [cs(i + 1, 4 * N)[0] for i in range(N)]
    fct = math.sqrt(1.0 / (2 * N))
    return facts, tws, tw, fct


def dct2(x):
    """scipy.fftpack.dct(x, type=2, norm='ortho')"""
    N = len(x)
    facts, tws, tw, fct = plan(N)
    c = list(x)
    c[0] = c[0] * 2.0
    c[N - 1] = c[N - 1] * 2.0
    for k in range(1, N - 1, 2):
        t = c[k + 1]
        c[k + 1] = t - c[k]
        c[k] = t + c[k]
    c = rfft_backward(c, N, facts, tws, fct)
    NS2 = (N + 1) // 2
    k, kc = 1, N - 1
    while k < NS2:
        t1 = tw[k - 1] * c[kc] + tw[kc - 1] * c[k]
        t2 = tw[k - 1] * c[k] - tw[kc - 1] * c[kc]
        c[k] = 0.5 * (t1 + t2)
        c[kc] = 0.5 * (t1 - t2)
        k += 1
        kc -= 1
    c[NS2] = c[NS2] * tw[NS2 - 1]
    c[0] = c[0] * (SQRT2 * 0.5)
    return c


def dct3(x):
    """scipy.fftpack.idct(x, type=2, norm='ortho')  (== DCT-III)"""
    N = len(x)
    facts, tws, tw, fct = plan(N)
    c = list(x)
    c[0] = c[0] * SQRT2
    NS2 = (N + 1) // 2
    k, kc = 1, N - 1
    while k < NS2:
        t1, t2 = c[k] + c[kc], c[k] - c[kc]
        c[k] = tw[k - 1] * t2 + tw[kc - 1] * t1
        c[kc] = tw[k - 1] * t1 - tw[kc - 1] * t2
        k += 1
        kc -= 1
    c[NS2] = c[NS2] * (2.0 * tw[NS2 - 1])
    c = rfft_forward(c, N, facts, tws, fct)
    for k in range(1, N - 1, 2):
        t = c[k]
        c[k] = t - c[k + 1]
        c[k + 1] = t + c[k + 1]
    return c


# ---- tracing ---------------------------------------------------------------------------------------------------
class Sym:
    __array_ufunc__ = None
    counter = 0
    prog = None

    def __init__(self, expr=None, name=None):
        if name is None:
            name = f"t{Sym.counter}"
            Sym.counter += 1
            Sym.prog.append((name, expr))
        self.name = name

    @staticmethod
    def _lit(v):
        return float(v).hex()

    def __add__(self, o):
        return Sym(("add", self.name, o.name))

    def __sub__(self, o):
        return Sym(("sub", self.name, o.name))

    def __mul__(self, o):
        assert not isinstance(o, Sym)
        return Sym(("mul", self.name, Sym._lit(o)))

    __rmul__ = __mul__

    def __neg__(self):
        return Sym(("neg", self.name))


def trace(fn, N):
    Sym.counter = 0
    Sym.prog = []
    xs = [Sym(name=f"x{i}") for i in range(N)]
    out = fn(xs)
    return list(Sym.prog), [o.name for o in out]


def emit(fn_name, prog, outs, N, flavour):
    lines = []
    if flavour == "cuda":
        lines.append(f"__device__ __forceinline__ void {fn_name}(double* v, int stride) {{")
        A, S, M = "__dadd_rn", "__dsub_rn", "__dmul_rn"
    else:
        lines.append(f"static void {fn_name}(double* v, long stride) {{")
        A = S = M = None
    for i in range(N):
        lines.append(f"    const double x{i} = v[{i} * stride];")
    for name, e in prog:
        if e[0] == "add":
            rhs = f"{A}({e[1]}, {e[2]})" if A else f"{e[1]} + {e[2]}"
        elif e[0] == "sub":
            rhs = f"{S}({e[1]}, {e[2]})" if S else f"{e[1]} - {e[2]}"
        elif e[0] == "mul":
            rhs = f"{M}({e[1]}, {e[2]})" if M else f"{e[1]} * {e[2]}"
        else:
            rhs = f"-{e[1]}"
        lines.append(f"    const double {name} = {rhs};")
    for i, o in enumerate(outs):
        lines.append(f"    v[{i} * stride] = {o};")
    lines.append("}")
    return "\n".join(lines)


HEADER = """// GENERATED by tools/dctgen/gen_dct.py -- do not edit.
// Straight-line FP64 DCT-II / DCT-III (N = 2, 4, 8, 16) reproducing scipy.fftpack.dct/idct(norm='ortho') of
// SciPy 1.18.1 (ducc0 backend) bit for bit: same operation order, same (not correctly rounded) twiddles, no FMA.
// The reference calls these at Encoder.py:781 and Encoder.py:812.
"""


def generate():
    cuda = [HEADER, "#pragma once\n"]
    cpu = [HEADER, "/* CPU build of the same program: test infrastructure (oracle/dct_ducc_c.c includes it). */\n"]
    for N in (2, 4, 8, 16):
        for nm, fn in (("dct2", dct2), ("dct3", dct3)):
            prog, outs = trace(fn, N)
            cuda.append(emit(f"ducc_{nm}_{N}", prog, outs, N, "cuda") + "\n")
            cpu.append(emit(f"ducc_{nm}_{N}", prog, outs, N, "c") + "\n")
    root = os.path.dirname(os.path.dirname(HERE))
    with open(os.path.join(root, "streamoptima_b200", "csrc", "so_dct_ducc.cuh"), "w") as f:
        f.write("\n".join(cuda))
    with open(os.path.join(root, "oracle", "dct_ducc_generated.h"), "w") as f:
        f.write("\n".join(cpu))


if __name__ == "__main__":
    generate()
    print("generated")
