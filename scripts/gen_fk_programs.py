#!/usr/bin/env python
"""URDF -> straight-line CUDA: emits csrc/fk_programs_gen.h with one fully unrolled, constant-folded kinematic chain per
packaged robot (Panda, Kuka, Baxter), plus the raw program tables the runtime matches an incoming hrp_fk_program
against (bitwise) before it selects a generated chain. Any other URDF runs through the table interpreter.

The chain is produced from the SAME compiled program (hrp_b200.urdf.compile_program) the interpreter executes, so
numbers are never copied by hand. Folding happens here, on float32 table entries: an entry with
|v| < 1e-12 is a structural zero (cos(pi/2) evaluated in float64 leaves 6e-17 in the reference's origins; dropping it
moves a keypoint by < 1e-16 m), entries within 1e-12 of +-1 are exact signs. What remains for a z-axis joint behind a
signed-permutation origin is two FMAs per rotated column instead of two dense 3x3 products.

  python scripts/gen_fk_programs.py            # rewrite the header
  python scripts/gen_fk_programs.py --check    # exit 1 if the committed header is stale (used by tests/test_host.py)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hrp_b200  # noqa: E402,F401
from hrp_b200 import urdf  # noqa: E402

OUT = os.path.join(ROOT, "holistic-robot-pose-estimation-study_b200", "csrc", "fk_programs_gen.h")
ROBOTS = ("panda", "kuka", "baxter")
EPS = 1e-12


def lit(v):
    return "%sf" % float(np.float32(v)).hex()


class Emitter:
    """Tiny expression builder: a value is a python float (compile-time constant) or the name of a float variable."""

    def __init__(self):
        self.lines = []
        self.n = 0

    def tmp(self, expr):
        name = "t%d" % self.n
        self.n += 1
        self.lines.append("  const float %s = %s;" % (name, expr))
        return name

    @staticmethod
    def const(v):
        v = float(np.float32(v))
        if abs(v) < EPS:
            return 0.0
        if abs(v - 1.0) < EPS:
            return 1.0
        if abs(v + 1.0) < EPS:
            return -1.0
        return v

    def neg(self, a):
        if isinstance(a, float):
            return -a
        return self.tmp("-%s" % a)

    def mul(self, a, b):
        if isinstance(a, float) and isinstance(b, float):
            return float(np.float32(a) * np.float32(b))
        if isinstance(b, float):
            a, b = b, a
        if isinstance(a, float):
            if a == 0.0:
                return 0.0
            if a == 1.0:
                return b
            if a == -1.0:
                return self.neg(b)
            return self.tmp("%s * %s" % (lit(a), b))
        return self.tmp("%s * %s" % (a, b))

    def add(self, a, b):
        if isinstance(a, float) and isinstance(b, float):
            return float(np.float32(a) + np.float32(b))
        if isinstance(b, float):
            a, b = b, a
        if isinstance(a, float):
            if a == 0.0:
                return b
            return self.tmp("%s + %s" % (b, lit(a)))
        return self.tmp("%s + %s" % (a, b))

    def fma(self, a, b, c):
        """a*b + c with folding; emits fmaf when all three survive."""
        if isinstance(a, float) and isinstance(b, float):
            return self.add(self.mul(a, b), c)
        if isinstance(b, float):
            a, b = b, a
        if isinstance(a, float) and a in (0.0, 1.0, -1.0):
            return self.add(self.mul(a, b), c) if a != -1.0 else self.sub(c, b)
        if isinstance(c, float) and c == 0.0:
            return self.mul(a, b)
        sa = lit(a) if isinstance(a, float) else a
        sc = lit(c) if isinstance(c, float) else c
        return self.tmp("fmaf(%s, %s, %s)" % (sa, b, sc))

    def sub(self, a, b):
        if isinstance(b, float):
            return self.add(a, -b)
        if isinstance(a, float):
            if a == 0.0:
                return self.neg(b)
            return self.tmp("%s - %s" % (lit(a), b))
        return self.tmp("%s - %s" % (a, b))

    def dot3(self, a, b, c=0.0):
        acc = c
        for x, y in zip(a, b):
            acc = self.fma(x, y, acc)
        return acc


def rodrigues(E, axis, sn, cs):
    """cos*I + (1-cos)*a a^T + sin*[a]x (urdf.py:2451-2463) with a constant axis; exact form for coordinate axes."""
    a = [E.const(v) for v in axis]
    nz = [i for i in range(3) if a[i] != 0.0]
    if len(nz) == 1 and abs(a[nz[0]]) == 1.0:
        k = nz[0]
        i, j = (k + 1) % 3, (k + 2) % 3
        s = sn if a[k] > 0 else E.neg(sn)
        R = [[0.0] * 3 for _ in range(3)]
        R[k][k] = 1.0
        R[i][i] = cs
        R[j][j] = cs
        R[i][j] = E.neg(s)
        R[j][i] = s
        return R
    oc = E.sub(1.0, cs)
    R = [[None] * 3 for _ in range(3)]
    cross = [[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]]
    for i in range(3):
        for j in range(3):
            v = E.mul(E.mul(a[i], a[j]), oc)
            if i == j:
                v = E.add(v, cs)
            R[i][j] = E.fma(cross[i][j], sn, v)
    return R


def gen_chain(name, P):
    E = Emitter()
    A = P.arrays()
    nsteps = len(P.step_type)
    frames = {}                       # step -> (R 3x3, t 3) of Exprs
    kp_lines = []
    kp_ptr = 0

    def emit_kp(step, T):
        nonlocal kp_ptr
        while kp_ptr < P.nkpt and int(A["kp_step"][kp_ptr]) == step:
            off = [E.const(v) for v in A["kp_offset"][kp_ptr]]
            o = int(A["kp_index"][kp_ptr]) * 3
            for r in range(3):
                if T is None:
                    v = off[r]
                else:
                    v = E.dot3(T[0][r], off, T[1][r])
                sv = lit(v) if isinstance(v, float) else v
                E.lines.append("  kp[%d] = %s;" % (o + r, sv))
            kp_ptr += 1

    emit_kp(-1, None)
    prev = None
    root = None
    for s in range(nsteps):
        q = "q[%d]" % int(A["step_q"][s])
        mul, off = E.const(A["step_mul"][s]), E.const(A["step_off"][s])
        qv = E.add(E.mul(mul, q), off)
        if not isinstance(qv, str) or qv == q:
            qv = E.tmp(q) if qv == q else qv
        O = A["step_origin"][s].reshape(3, 4)
        Or = [[E.const(O[i][j]) for j in range(3)] for i in range(3)]
        Ot = [E.const(O[i][3]) for i in range(3)]
        if int(A["step_type"][s]) == urdf.REVOLUTE:
            E.lines.append("  float sn%d, cs%d;" % (s, s))
            E.lines.append("  fk_sincos(%s, sn%d, cs%d);" % (qv, s, s))
            R = rodrigues(E, A["step_axis"][s], "sn%d" % s, "cs%d" % s)
            Mr = [[E.dot3(Or[i], [R[0][j], R[1][j], R[2][j]]) for j in range(3)] for i in range(3)]
            Mt = Ot
        else:
            d = [E.mul(E.const(v), qv) for v in A["step_axis"][s]]
            Mr = Or
            Mt = [E.dot3(Or[i], d, Ot[i]) for i in range(3)]
        par = int(A["step_parent"][s])
        if par == urdf.PARENT_BASE:
            T = (Mr, Mt)
        else:
            Pm = prev if par == urdf.PARENT_PREV else frames[("slot", par)]
            Tr_ = [[E.dot3(Pm[0][i], [Mr[0][j], Mr[1][j], Mr[2][j]]) for j in range(3)] for i in range(3)]
            Tt_ = [E.dot3(Pm[0][i], Mt, Pm[1][i]) for i in range(3)]
            T = (Tr_, Tt_)
        sv = int(A["step_save"][s])
        if sv >= 0:
            frames[("slot", sv)] = T
        if s == P.root_step:
            root = T
        prev = T
        emit_kp(s, T)
    assert kp_ptr == P.nkpt
    if P.root_kp != 0:
        F = np.asarray(A["root_fixed"]).reshape(3, 4)
        Fr = [[E.const(F[i][j]) for j in range(3)] for i in range(3)]
        Ft = [E.const(F[i][3]) for i in range(3)]
        if root is None:
            Rr, Rt = Fr, Ft
        else:
            Rr = [[E.dot3(root[0][i], [Fr[0][j], Fr[1][j], Fr[2][j]]) for j in range(3)] for i in range(3)]
            Rt = [E.dot3(root[0][i], Ft, root[1][i]) for i in range(3)]
        for i in range(3):
            for j in range(3):
                v = Rr[i][j]
                E.lines.append("  Tr[%d] = %s;" % (i * 3 + j, lit(v) if isinstance(v, float) else v))
        for i in range(3):
            v = Rt[i]
            E.lines.append("  Tr[%d] = %s;" % (9 + i, lit(v) if isinstance(v, float) else v))
    head = ["// %s: dof %d, %d keypoints, %d movable steps, root keypoint %d" % (name, P.dof, P.nkpt, nsteps, P.root_kp),
            "__device__ __forceinline__ void fk_chain_%s(const float (&q)[%d], float (&kp)[%d], float (&Tr)[12]) {" % (name, P.dof, P.nkpt * 3)]
    return "\n".join(head + E.lines + ["}"])


def gen_raw(name, P):
    A = P.arrays()

    def ints(k):
        return ", ".join(str(int(v)) for v in A[k].reshape(-1)) or "0"

    def flts(k):
        return ", ".join("0x%08xu" % int(np.float32(v).view(np.uint32)) for v in A[k].reshape(-1)) or "0"

    n = len(P.step_type)
    return ("static const FkRaw fk_raw_%s = {%d, %d, %d, %d, %d, %d,\n  {%s}, {%s}, {%s}, {%s},\n  {%s},\n  {%s},\n  {%s},\n  {%s},\n  {%s}, {%s},\n  {%s},\n  {%s}};"
            % (name, P.dof, P.nkpt, n, P.n_slots, P.root_kp, P.root_step, ints("step_type"), ints("step_parent"), ints("step_save"),
               ints("step_q"), flts("step_mul"), flts("step_off"), flts("step_origin"), flts("step_axis"), ints("kp_step"),
               ints("kp_index"), flts("kp_offset"), flts("root_fixed")))


def generate():
    parts = ["// GENERATED by scripts/gen_fk_programs.py from the packaged URDFs -- do not edit; `--check` runs in the CPU tests.",
             "// Included by fk_project.cu inside namespace hrp (needs fk_sincos and FkRaw).", ""]
    for name in ROBOTS:
        _, P = urdf.load_robot(name)
        parts.append(gen_chain(name, P))
        parts.append("")
        parts.append(gen_raw(name, P))
        parts.append("template <> struct FkGen<FK_%s> {" % name.upper())
        # CTAs of 128 poses that fit one SM's 227 KB with one input tile + one output staging tile each (register budget follows)
        min_ctas = min(5, (227 * 1024) // (128 * 4 * (P.dof + 18 + P.nkpt * 5) + 1024))
        parts.append("  static constexpr int DOF = %d, NK = %d, ROOT_KP = %d, MIN_CTAS = %d;" % (P.dof, P.nkpt, P.root_kp, min_ctas))
        parts.append("  static __device__ __forceinline__ void chain(const float (&q)[DOF], float (&kp)[NK * 3], float (&Tr)[12]) { fk_chain_%s(q, kp, Tr); }" % name)
        parts.append("};")
        parts.append("")
    return "\n".join(parts)


if __name__ == "__main__":
    text = generate()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            print("fk_programs_gen.h is stale: run scripts/gen_fk_programs.py")
            sys.exit(1)
        print("fk_programs_gen.h is up to date")
    else:
        with open(OUT, "w") as f:
            f.write(text)
        print("wrote", OUT, "(%d lines)" % text.count("\n"))
