#!/usr/bin/env python3
"""Generate straight-line CUDA device code for the IAPWS-97 power sums.

The reference evaluates integer powers through `powertable` objects whose
multiplication chains are fixed at configure time (src/powertable.F90:95-245)
and then sums coefficient * power * power terms in array order
(src/IAPWS.F90:503-542 region 1, :596-639 region 2, :689-727 region 3,
:412-443 viscosity).  On the GPU a table in local memory would spill, so this
script replays the reference's configuration rule offline and prints the
resulting chain as straight-line register code: every power is one multiply of
two earlier powers, in exactly the product order the reference would use, and
each sum runs in the reference's term order (zero-coefficient terms dropped:
adding an exact zero does not change the sum).

Run:  python tools/gen_iapws.py > waiwera_b200/csrc/wb_iapws_gen.cuh
"""
import sys


def nint(x):
    # Fortran nint: half away from zero
    return int(x + 0.5) if x >= 0 else -int(-x + 0.5)


class PowerTable:
    """Replays powertable_configure (src/powertable.F90:135-245)."""

    def __init__(self):
        self.lower = self.upper = None
        self.product = {}
        self.required = {}
        self.configured_once = False
        self.plist = []

    def _conf(self, i):
        return abs(i) <= 1 or (self.product.get(i, (0, 0))[0] != 0 and self.product.get(i, (0, 0))[1] != 0)

    def _configure_product(self, i):
        if abs(i) > 1 and not self._conf(i):
            s = 1 if i >= 0 else -1
            i2 = nint(i / 2.0)
            c = i2
            while (c >= s) if s > 0 else (c <= s):
                j = i - c
                if self._conf(c) and self._conf(j):
                    self.product[i] = (c, j)
                    break
                c -= s
            if not self._conf(i):
                if i % 2 == 0:
                    self._configure_product(i2)
                    self.product[i] = (i2, i2)
                    self.required[i2] = 2
                else:
                    j = i - s
                    self._configure_product(j)
                    self.product[i] = (s, j)
                    self.required[j] = 2

    def configure(self, powers):
        minp = min([0] + list(powers))
        maxp = max([1] + list(powers))
        if self.configured_once:
            old_required = self.required
            self.lower = min(self.lower, minp)
            self.upper = max(self.upper, maxp)
            self.required = {k: 1 for k, v in old_required.items() if v == 1}
        else:
            self.lower, self.upper = minp, maxp
            self.required = {}
        self.product = {}
        self.configured_once = True
        for p in powers:
            if abs(p) > 1:
                self.required[p] = 1
        for s in (1, -1):
            u = self.upper if s > 0 else self.lower
            i = 2 * s
            while (i <= u) if s > 0 else (i >= u):
                if self.required.get(i, 0) > 0:
                    self._configure_product(i)
                i += s
        self.plist = []
        for s in (1, -1):
            u = self.upper if s > 0 else self.lower
            p = 2 * s
            while (p <= u) if s > 0 else (p >= u):
                if self.required.get(p, 0) > 0:
                    self.plist.append((self.product[p][0], self.product[p][1], p))
                p += s


def pname(prefix, i):
    return "%s_%s%d" % (prefix, "m" if i < 0 else "p", abs(i))


def emit_chain(out, tbl, prefix, val, used):
    """Emit the chain for the powers in `used` (and what they depend on)."""
    need = set()

    def mark(i):
        if i in need or i == 0:
            return
        need.add(i)
        if abs(i) > 1:
            a, b = tbl.product[i]
            mark(a)
            mark(b)
        elif i == -1:
            pass

    for u in used:
        mark(u)
    out.append("  const double %s = %s;" % (pname(prefix, 1), val))
    if any(i < 0 for i in need):
        out.append("  const double %s = 1.0 / %s;" % (pname(prefix, -1), pname(prefix, 1)))
    for a, b, p in tbl.plist:
        if p in need:
            out.append("  const double %s = %s * %s;" % (pname(prefix, p), pname(prefix, a), pname(prefix, b)))


def pref(prefix, i):
    return "1.0" if i == 0 else pname(prefix, i)


def lit(x):
    return repr(float(x))


def term(coef, facs):
    """(coef * f1) * f2 with unit factors dropped (x*1.0 == x exactly)."""
    e = lit(coef)
    for f in facs:
        if f != "1.0":
            e = "%s * %s" % (e, f)
    return e


# ---- coefficient tables (src/IAPWS.F90:48-233) ----
r1_n = [0.14632971213167, -0.84548187169114, -0.37563603672040e1, 0.33855169168385e1,
        -0.95791963387872, 0.15772038513228, -0.16616417199501e-1, 0.81214629983568e-3,
        0.28319080123804e-3, -0.60706301565874e-3, -0.18990068218419e-1, -0.32529748770505e-1,
        -0.21841717175414e-1, -0.52838357969930e-4, -0.47184321073267e-3, -0.30001780793026e-3,
        0.47661393906987e-4, -0.44141845330846e-5, -0.72694996297594e-15, -0.31679644845054e-4,
        -0.28270797985312e-5, -0.85205128120103e-9, -0.22425281908000e-5, -0.65171222895601e-6,
        -0.14341729937924e-12, -0.40516996860117e-6, -0.12734301741641e-8, -0.17424871230634e-9,
        -0.68762131295531e-18, 0.14478307828521e-19, 0.26335781662795e-22, -0.11947622640071e-22,
        0.18228094581404e-23, -0.93537087292458e-25]
r1_I = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 8, 8, 21, 23, 29, 30, 31, 32]
r1_J = [-2, -1, 0, 1, 2, 3, 4, 5, -9, -7, -1, 0, 1, 3, -3, 0, 1, 3, 17, -4, 0, 6, -5, -2, 10, -8, -11, -6,
        -29, -31, -38, -39, -40, -41]
r2_n0 = [-0.96927686500217e1, 0.10086655968018e2, -0.56087911283020e-2, 0.71452738081455e-1,
         -0.40710498223928, 0.14240819171444e1, -0.43839511319450e1, -0.28408632460772, 0.21268463753307e-1]
r2_J0 = [0, 1, -5, -4, -3, -2, -1, 2, 3]
r2_n = [-0.17731742473213e-2, -0.17834862292358e-1, -0.45996013696365e-1, -0.57581259083432e-1,
        -0.50325278727930e-1, -0.33032641670203e-4, -0.18948987516315e-3, -0.39392777243355e-2,
        -0.43797295650573e-1, -0.26674547914087e-4, 0.20481737692309e-7, 0.43870667284435e-6,
        -0.32277677238570e-4, -0.15033924542148e-2, -0.40668253562649e-1, -0.78847309559367e-9,
        0.12790717852285e-7, 0.48225372718507e-6, 0.22922076337661e-5, -0.16714766451061e-10,
        -0.21171472321355e-2, -0.23895741934104e2, -0.59059564324270e-17, -0.12621808899101e-5,
        -0.38946842435739e-1, 0.11256211360459e-10, -0.82311340897998e1, 0.19809712802088e-7,
        0.10406965210174e-18, -0.10234747095929e-12, -0.10018179379511e-8, -0.80882908646985e-10,
        0.10693031879409, -0.33662250574171, 0.89185845355421e-24, 0.30629316876232e-12,
        -0.42002467698208e-5, -0.59056029685639e-25, 0.37826947613457e-5, -0.12768608934681e-14,
        0.73087610595061e-28, 0.55414715350778e-16, -0.94369707241210e-6]
r2_I = [1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 5, 6, 6, 6, 7, 7, 7, 8, 8, 9, 10, 10, 10, 16, 16,
        18, 20, 20, 20, 21, 22, 23, 24, 24, 24]
r2_J = [0, 1, 2, 3, 6, 1, 2, 4, 7, 36, 0, 1, 3, 6, 35, 1, 2, 3, 7, 3, 16, 35, 0, 11, 25, 8, 36, 13, 4, 10, 14,
        29, 50, 57, 20, 35, 48, 21, 53, 39, 26, 40, 58]
r3_n = [0.10658070028513e1, -0.15732845290239e2, 0.20944396974307e2, -0.76867707878716e1,
        0.26185947787954e1, -0.28080781148620e1, 0.12053369696517e1, -0.84566812812502e-2,
        -0.12654315477714e1, -0.11524407806681e1, 0.88521043984318, -0.64207765181607,
        0.38493460186671, -0.85214708824206, 0.48972281541877e1, -0.30502617256965e1,
        0.39420536879154e-1, 0.12558408424308, -0.27999329698710, 0.13899799569460e1,
        -0.20189915023570e1, -0.82147637173963e-2, -0.47596035734923, 0.43984074473500e-1,
        -0.44476435428739, 0.90572070719733, 0.70522450087967, 0.10770512626332,
        -0.32913623258954, -0.50871062041158, -0.22175400873096e-1, 0.94260751665092e-1,
        0.16436278447961, -0.13503372241348e-1, -0.14834345352472e-1, 0.57922953628084e-3,
        0.32308904703711e-2, 0.80964802996215e-4, -0.16557679795037e-3, -0.44923899061815e-4]
r3_I = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 6, 6, 6, 7, 8,
        9, 9, 10, 10, 11]
r3_J = [0, 0, 1, 2, 7, 10, 12, 23, 2, 6, 15, 17, 0, 2, 6, 7, 22, 26, 0, 2, 4, 16, 26, 0, 2, 4, 26, 1, 3, 26, 0,
        2, 26, 2, 26, 2, 26, 0, 1, 26]
visc_h0 = [1.67752, 2.20462, 0.6366564, -0.241605]
visc_h1 = [5.20094e-1, 8.50895e-2, -1.08374, -2.89555e-1, 2.22531e-1, 9.99115e-1, 1.88797, 1.26613,
           1.20573e-1, -2.81378e-1, -9.06851e-1, -7.72479e-1, -4.89837e-1, -2.57040e-1, 1.61913e-1,
           2.57399e-1, -3.25372e-2, 6.98452e-2, 8.72102e-3, -4.35673e-3, -5.93264e-4]
visc_I = [0, 1, 2, 3, 0, 1, 2, 3, 5, 0, 1, 2, 3, 4, 0, 1, 0, 3, 4, 3, 5]
visc_J = [0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 4, 4, 5, 6, 6]


def table(*power_sets):
    t = PowerTable()
    for ps in power_sets:
        t.configure(ps)
    return t


def sums(out, name, coefs, tI, pI, tJ, pJ, dI, dJ, weight):
    """name += (n*w) * PI[I+dI] * PJ[J+dJ] in array order."""
    out.append("  double %s = 0.0;" % name)
    for n, I, J in coefs:
        w = n * weight(I, J)
        if w == 0.0:
            continue
        out.append("  %s += %s;" % (name, term(w, [pref(pI, I + dI), pref(pJ, J + dJ)])))


def main():
    o = []
    o.append("// GENERATED by tools/gen_iapws.py -- do not edit.")
    o.append("// IAPWS-97 power sums as straight-line register code; the multiplication")
    o.append("// chains replay the reference's powertable configuration (src/powertable.F90:95-245),")
    o.append("// the sums run in the reference's term order (src/IAPWS.F90:412-443,503-542,596-639,689-727).")
    o.append("#pragma once")
    o.append("")

    # ---- region 1 ----
    tI = table(r1_I, [i - 1 for i in r1_I])
    tJ = table(r1_J, [j - 1 for j in r1_J])
    o.append("// src/IAPWS.F90:503-542: a = 7.1 - pi, b = tau - 1.222; s1 = sum nI a^(I-1) b^J, s2 = sum nJ a^I b^(J-1)")
    o.append("WB_HD void wb_iapws_r1_sums(double a, double b, double &s1_out, double &s2_out) {")
    usedI = set([i - 1 for n, i in zip(r1_n, r1_I) if n * i != 0] + [i for n, i, j in zip(r1_n, r1_I, r1_J) if n * j != 0])
    usedJ = set([j for n, i, j in zip(r1_n, r1_I, r1_J) if n * i != 0] + [j - 1 for n, j in zip(r1_n, r1_J) if n * j != 0])
    emit_chain(o, tI, "a", "a", usedI)
    emit_chain(o, tJ, "b", "b", usedJ)
    c = list(zip(r1_n, r1_I, r1_J))
    sums(o, "s1", c, tI, "a", tJ, "b", -1, 0, lambda I, J: I)
    sums(o, "s2", c, tI, "a", tJ, "b", 0, -1, lambda I, J: J)
    o.append("  s1_out = s1; s2_out = s2;")
    o.append("}")
    o.append("")

    # ---- region 2 ----
    tJ0 = table(r2_J0, [j - 1 for j in r2_J0])
    tI = table(r2_I, [i - 1 for i in r2_I], [-1])
    tJ = table(r2_J, [j - 1 for j in r2_J])
    o.append("// src/IAPWS.F90:596-639: gamt0 = sum n0 J0 tau^(J0-1); gampir = sum nI pi^(I-1) c^J; gamtr = sum nJ pi^I c^(J-1); c = tau-0.5")
    o.append("WB_HD void wb_iapws_r2_sums(double pi, double tau, double c, double &gamt0_out, double &gampir_out, double &gamtr_out, double &pim1_out) {")
    used0 = set(j - 1 for n, j in zip(r2_n0, r2_J0) if n * j != 0)
    usedI = set([i - 1 for i in r2_I] + list(r2_I) + [-1])
    usedJ = set([j for j in r2_J] + [j - 1 for n, j in zip(r2_n, r2_J) if n * j != 0])
    emit_chain(o, tJ0, "t", "tau", used0)
    emit_chain(o, tI, "q", "pi", usedI)
    emit_chain(o, tJ, "c", "c", usedJ)
    o.append("  double gamt0 = 0.0;")
    for n, j in zip(r2_n0, r2_J0):
        w = n * j
        if w != 0.0:
            o.append("  gamt0 += %s;" % term(w, [pref("t", j - 1)]))
    c2 = list(zip(r2_n, r2_I, r2_J))
    sums(o, "gampir", c2, tI, "q", tJ, "c", -1, 0, lambda I, J: I)
    sums(o, "gamtr", c2, tI, "q", tJ, "c", 0, -1, lambda I, J: J)
    o.append("  gamt0_out = gamt0; gampir_out = gampir; gamtr_out = gamtr; pim1_out = %s;" % pname("q", -1))
    o.append("}")
    o.append("")

    # ---- region 3 ----
    tI = table(r3_I, [i - 1 for i in r3_I])
    tJ = table(r3_J, [j - 1 for j in r3_J])
    o.append("// src/IAPWS.F90:689-727: s1 = sum nI delta^(I-1) tau^J ; s2 = sum nJ delta^I tau^(J-1); dm1 = 1/delta")
    o.append("WB_HD void wb_iapws_r3_sums(double delta, double tau, double &s1_out, double &s2_out, double &dm1_out) {")
    usedI = set([i - 1 for n, i in zip(r3_n, r3_I) if n * i != 0] + [i for n, i, j in zip(r3_n, r3_I, r3_J) if n * j != 0] + [-1])
    usedJ = set([j for n, i, j in zip(r3_n, r3_I, r3_J) if n * i != 0] + [j - 1 for n, j in zip(r3_n, r3_J) if n * j != 0])
    emit_chain(o, tI, "d", "delta", usedI)
    emit_chain(o, tJ, "t", "tau", usedJ)
    c3 = list(zip(r3_n, r3_I, r3_J))
    sums(o, "s1", c3, tI, "d", tJ, "t", -1, 0, lambda I, J: I)
    sums(o, "s2", c3, tI, "d", tJ, "t", 0, -1, lambda I, J: J)
    o.append("  s1_out = s1; s2_out = s2; dm1_out = %s;" % pname("d", -1))
    o.append("}")
    o.append("")

    # ---- viscosity ----
    tI = table(visc_I)
    tJ = table(visc_J)
    tK = table([0, 1, 2, 3])
    o.append("// src/IAPWS.F90:412-443: s0 = sum h0_k (1/tau)^k ; s1 = sum ((1/tau-1)^I * h1) * (del-1)^J")
    o.append("WB_HD void wb_iapws_visc_sums(double rtau, double dm1, double &s0_out, double &s1_out) {")
    emit_chain(o, tK, "k", "rtau", set([1, 2, 3]))
    o.append("  const double e = %s - 1.0;" % pname("k", 1))
    emit_chain(o, tI, "e", "e", set(visc_I))
    emit_chain(o, tJ, "f", "dm1", set(visc_J))
    o.append("  double s0 = 0.0;")
    for k, h in enumerate(visc_h0):
        o.append("  s0 += %s;" % term(h, [pref("k", k)]))
    o.append("  double s1 = 0.0;")
    for h, I, J in zip(visc_h1, visc_I, visc_J):
        # reference order: PI * h1 * PJ
        e = pref("e", I)
        if e == "1.0":
            expr = lit(h)
        else:
            expr = "%s * %s" % (e, lit(h))
        if J != 0:
            expr = "%s * %s" % (expr, pref("f", J))
        o.append("  s1 += %s;" % expr)
    o.append("  s0_out = s0; s1_out = s1;")
    o.append("}")
    sys.stdout.write("\n".join(o) + "\n")


if __name__ == "__main__":
    main()
