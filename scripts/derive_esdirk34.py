"""Derive the ESDIRK3(4) tableau used by csrc/device/psi_bdf.cuh (solver `Esdirk34`, the reference's
OdeSolver::Sdirk(SdirkTableau::Esdirk34), ode/mod.rs:59-84 -> diffsol `esdirk34`) from its defining conditions, with
40-digit arithmetic.  diffsol is third-party and not vendored under /root/reference; the method it names is the ESDIRK34
of Jorgensen, Kristensen & Thomsen, "A family of ESDIRK integration methods" (2018): 4 stages, explicit first stage,
stiffly accurate, L-stable, stage order 2, advancing order 3 with an embedded 4th-order solution.

  stage order 2      c2 = 2 gamma;  a31 + a32 + gamma = c3;  a32 c2 + gamma c3 = c3^2 / 2
  order 3            sum b = 1;  b.c = 1/2;  b.c^2 = 1/3                 (b4 = gamma, c4 = 1: stiffly accurate)
  L-stability        gamma^3 - 3 gamma^2 + 3/2 gamma - 1/6 = 0           (the root 0.4358665...)
  embedded order 4   sum bh = 1; bh.c = 1/2; bh.c^2 = 1/3; bh.c^3 = 1/4; bh.A.c^2 = 1/12   -> pins c3

The solution is unique for c3 in (0, 1); the printed values are the literals in Esdirk34Tab.

    python scripts/derive_esdirk34.py
"""
import mpmath as mp

mp.mp.dps = 40
g = mp.findroot(lambda x: x**3 - 3 * x**2 + mp.mpf(3) / 2 * x - mp.mpf(1) / 6, 0.4358665)


def build(c3):
    c2 = 2 * g
    a32 = (c3**2 / 2 - g * c3) / c2
    a31 = c3 - a32 - g
    b = mp.lu_solve(mp.matrix([[1, 1, 1], [0, c2, c3], [0, c2**2, c3**2]]), mp.matrix([1 - g, mp.mpf(1) / 2 - g, mp.mpf(1) / 3 - g]))
    A = mp.matrix(4, 4)
    A[1, 0] = g; A[1, 1] = g
    A[2, 0] = a31; A[2, 1] = a32; A[2, 2] = g
    A[3, 0] = b[0]; A[3, 1] = b[1]; A[3, 2] = b[2]; A[3, 3] = g
    return A, mp.matrix([0, c2, c3, 1])


def embedded(c3):
    A, c = build(c3)
    V = mp.matrix([[1, 1, 1, 1], [c[i] for i in range(4)], [c[i]**2 for i in range(4)], [c[i]**3 for i in range(4)]])
    bh = mp.lu_solve(V, mp.matrix([1, mp.mpf(1) / 2, mp.mpf(1) / 3, mp.mpf(1) / 4]))
    Ac2 = A * mp.matrix([c[i]**2 for i in range(4)])
    return sum(bh[i] * Ac2[i] for i in range(4)) - mp.mpf(1) / 12, bh


if __name__ == "__main__":
    c3 = mp.findroot(lambda x: embedded(x)[0], 0.5)
    A, c = build(c3)
    bh = embedded(c3)[1]
    print("gamma =", mp.nstr(g, 22))
    print("c     =", [mp.nstr(c[i], 22) for i in range(4)])
    for i in range(4):
        print(f"A[{i}]  =", [mp.nstr(A[i, j], 22) for j in range(4)])
    print("bhat  =", [mp.nstr(x, 22) for x in bh])
    # order conditions of the advancing method, to 1e-35
    b = [A[3, j] for j in range(4)]
    assert abs(sum(b) - 1) < 1e-35 and abs(sum(b[i] * c[i] for i in range(4)) - mp.mpf(1) / 2) < 1e-35
    assert abs(sum(b[i] * c[i]**2 for i in range(4)) - mp.mpf(1) / 3) < 1e-35
    Ac = A * c
    assert abs(sum(b[i] * Ac[i] for i in range(4)) - mp.mpf(1) / 6) < 1e-35
