"""Search the linear (non-swizzled) plane layouts: pads of the z-strides of T1 / T2 / quadrature planes and the row pitches, per
(P, Q, elements per group, lanes per group), minimising modelled shared-memory wavefronts; ties -> smaller planes.
usage: python scripts/model/search_linear.py  -> prints a C++ table for b200_opgen.cpp"""
import sys
sys.path.insert(0, "scripts/model")
from bank_conflicts import wavefronts


def cost_stage(ntasks, TS, fns):
    c = 0
    for base in range(0, ntasks, TS):
        for w0 in range(base, min(base + TS, ntasks), 32):
            lanes = [t if t < ntasks else None for t in range(w0, w0 + 32)]
            for fn, wgt in fns:
                c += wgt * wavefronts([fn(t) if t is not None else None for t in lanes])
    return c


def parts(P, Q, Qs, Ps, Sz1, Sz2, SZ, collocated, TS):
    """(cost depending on Sz1, cost depending on (Ps, Sz2), cost depending on (Qs, SZ))"""
    c1 = c2 = c3 = 0
    if not collocated:
        dec = lambda t: (t % P, t // P)
        c1 += 2 * cost_stage(P * Q, TS, [(lambda t, j=j: dec(t)[1] * Sz1 + j * P + dec(t)[0], 1) for j in range(P)])
        c2 += 2 * cost_stage(P * Q, TS, [(lambda t, q=q: dec(t)[1] * Sz2 + q * Ps + dec(t)[0], 1) for q in range(Q)])
        dr = lambda t: (t % Q, t // Q)
        c2 += 2 * cost_stage(Q * Q, TS, [(lambda t, i=i: dr(t)[1] * Sz2 + dr(t)[0] * Ps + i, 1) for i in range(P)])
        c3 += 2 * cost_stage(Q * Q, TS, [(lambda t, q=q: dr(t)[1] * SZ + dr(t)[0] * Qs + q, 2) for q in range(Q)])
        c1 += cost_stage(P * P, TS, [(lambda t, q=q: q * Sz1 + t, 2) for q in range(Q)])
    else:
        c3 += cost_stage(P * P, TS, [(lambda t, k=k: k * SZ + (t // P) * Qs + t % P, 2) for k in range(P)])
        dr = lambda t: (t % Q, t // Q)  # d/dx on rows: read the values, write the derivative (and the transpose)
        c3 += 2 * cost_stage(Q * Q, TS, [(lambda t, q=q: dr(t)[1] * SZ + dr(t)[0] * Qs + q, 2) for q in range(Q)])
    d = lambda t: (t % Q, t // Q)
    c3 += 2 * cost_stage(Q * Q, TS, [(lambda t, m=m: d(t)[1] * SZ + m * Qs + d(t)[0], 2) for m in range(Q)])
    c3 += cost_stage(Q * Q, TS, [(lambda t, m=m: m * SZ + d(t)[1] * Qs + d(t)[0], 6) for m in range(Q)])
    return c1, c2, c3


def model(P, Q, Qs, Ps, Sz1, Sz2, SZ, collocated, TS):
    return sum(parts(P, Q, Qs, Ps, Sz1, Sz2, SZ, collocated, TS))


def search(P, Q, collocated, TS):
    Ps0, Qs0 = P + (1 - P % 2), Q + (1 - Q % 2)
    b3 = min(((parts(P, Q, Qs, Ps0, P * P, Q * Ps0, Q * Qs + pz, collocated, TS)[2], Q * (Q * Qs + pz)), (Qs, Q * Qs + pz)) for Qs in (Q, Q + 1, Q + 2) for pz in range(16))
    Qs, SZ = b3[1]
    if collocated:
        return (b3[0][0], Q * SZ), (Qs, P, P * P, Q * P, SZ)
    b1 = min(((parts(P, Q, Qs0, Ps0, P * P + p1, Q * Ps0, Q * Qs0, collocated, TS)[0], p1), P * P + p1) for p1 in range(16))
    b2 = min(((parts(P, Q, Qs0, Ps, P * P, Q * Ps + p2, Q * Qs0, collocated, TS)[1], Q * Ps + p2), (Ps, Q * Ps + p2)) for Ps in (P, P + 1, P + 2) for p2 in range(16))
    Sz1, (Ps, Sz2) = b1[1], b2[1]
    return (b3[0][0] + b1[0][0] + b2[0][0], max(Q * SZ, Q * Sz1, Q * Sz2)), (Qs, Ps, Sz1, Sz2, SZ)


if __name__ == "__main__":
    cases = [(P, P + 1, False) for P in range(2, 10)] + [(P, P, True) for P in range(4, 9)] + [(P, P + 2, False) for P in range(2, 8)]
    for P, Q, col in cases:
        TS = 32 if Q * Q <= 32 else (64 if Q * Q <= 64 else 128)
        cur = model(P, Q, Q + (1 - Q % 2), P + (1 - P % 2), P * P, Q * (P + (1 - P % 2)), Q * (Q + (1 - Q % 2)), col, TS)
        (c, S), lay = search(P, Q, col, TS)
        print(f"  {{{P}, {Q}, {int(col)}, {lay[0]}, {lay[1]}, {lay[2]}, {lay[3]}, {lay[4]}}},  // wavefronts {cur} -> {c}, plane {S} doubles", flush=True)
