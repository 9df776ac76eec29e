#!/usr/bin/env python
"""Generate the polynomial coefficients of csc2_exp (csrc/cloudsc2_math.cuh).

exp(r) = 1 + r + r^2 * P(r) on |r| <= ln2/2 with P of degree DEG-2 obtained by a Remez exchange
(minimax in RELATIVE error of exp) carried out in 60-digit arithmetic with mpmath; coefficients
are then rounded to double and the achieved error re-measured with the rounded values.
With a table of NTAB entries 2^(j/NTAB) the reduced argument satisfies |r| <= ln2/(2 NTAB).
Run:  python tools/gen_exp_coeffs.py [DEG [NTAB]]      (prints C initialisers)
"""
import sys
import mpmath as mp

mp.mp.dps = 60
DEG = int(sys.argv[1]) if len(sys.argv) > 1 else 11
NTAB = int(sys.argv[2]) if len(sys.argv) > 2 else 1     # table size: |r| <= ln2 / (2 NTAB)
N = DEG - 2 + 1            # number of coefficients of P
B = mp.log(2) / (2 * NTAB) * mp.mpf("1.0001")


def f(r):                  # target for P: (exp(r) - 1 - r) / r^2
    if abs(r) < mp.mpf("1e-25"):
        return mp.mpf(1) / 2 + r / 6
    return (mp.exp(r) - 1 - r) / (r * r)


def weight(r):             # relative error of exp: err_exp = r^2 (P - f) / exp(r)
    return r * r / mp.exp(r)


def solve(nodes):
    # unknowns: N coefficients + E ;  w(x_i) (P(x_i) - f(x_i)) = (-1)^i E
    A = mp.matrix(N + 1, N + 1)
    b = mp.matrix(N + 1, 1)
    for i, x in enumerate(nodes):
        w = weight(x)
        for j in range(N):
            A[i, j] = w * x ** j
        A[i, N] = -(-1) ** i
        b[i] = w * f(x)
    sol = mp.lu_solve(A, b)
    return [sol[j] for j in range(N)], sol[N]


def err(c, x):
    p = sum(cj * x ** j for j, cj in enumerate(c))
    return weight(x) * (p - f(x))


# initial nodes: Chebyshev extrema, avoiding r = 0 where the weight vanishes
nodes = [B * mp.cos(mp.pi * (N - i) / N) for i in range(N + 1)]
nodes = [x if abs(x) > 1e-6 else mp.mpf("1e-3") for x in nodes]
for it in range(30):
    c, E = solve(nodes)
    # locate the extrema of the error on a fine grid, one per sign interval
    M = 4000
    xs = [-B + 2 * B * k / M for k in range(M + 1)]
    es = [err(c, x) for x in xs]
    ext = []
    k = 0
    cur = [xs[0], es[0]]
    for x, e in zip(xs[1:], es[1:]):
        if (e > 0) == (cur[1] > 0) or e == 0:
            if abs(e) > abs(cur[1]):
                cur = [x, e]
        else:
            ext.append(cur)
            cur = [x, e]
    ext.append(cur)
    if len(ext) != N + 1:
        # weight has a double zero at r = 0: merge the smallest extrema around it
        ext.sort(key=lambda t: -abs(t[1]))
        ext = sorted(ext[:N + 1], key=lambda t: t[0])
    new_nodes = [t[0] for t in ext]
    if max(abs(a - b) for a, b in zip(new_nodes, nodes)) < B / M:
        nodes = new_nodes
        break
    nodes = new_nodes

cd = [float(x) for x in c]
worst = max(abs(err([mp.mpf(v) for v in cd], -B + 2 * B * k / 2000)) for k in range(2001))
print(f"// degree {DEG}: minimax relative error {float(abs(E)):.3e}, with double coefficients {float(worst):.3e}")
print("{" + ", ".join(f"{v:.17e}" for v in cd) + "}")

if NTAB > 1:
    print(f"// 2^(j/{NTAB}), j = 0..{NTAB - 1}, correctly rounded")
    print("{" + ", ".join(f"{float(mp.mpf(2) ** (mp.mpf(j) / NTAB)):.17e}" for j in range(NTAB)) + "}")
    l2 = mp.log(2) / NTAB
    hi = float(l2)
    # ln2/NTAB split: hi with 21 trailing zero bits so that k*hi is exact for |k| < 2^21
    import struct
    bits = struct.unpack("<Q", struct.pack("<d", hi))[0] & ~((1 << 21) - 1)
    hi = struct.unpack("<d", struct.pack("<Q", bits))[0]
    lo = float(l2 - mp.mpf(hi))
    print(f"// ln2/{NTAB} = hi + lo : {hi:.20e} {lo:.20e} ; {NTAB}/ln2 = {float(NTAB / mp.log(2)):.17e}")
