"""Shared-memory bank-conflict model of the fused operator's line-owned stages (development aid, not product).

A 64-bit shared access of a warp is served per half-warp; a half-warp needs max-multiplicity(bank) wavefronts where
bank = (address in doubles) mod 16 and lanes reading the SAME address count once.  For every stage of b200_opgen.cpp the lane ->
address maps are restated here as functions of the plane strides, so that pads can be searched per (P, Q, elements per group).
"""
import itertools
import sys


def wavefronts(addrs):
    """addrs: list of per-lane addresses (None = inactive) for a full warp -> wavefronts (ideal: 2 for a full warp)."""
    total = 0
    for h in range(0, len(addrs), 16):
        half = [a for a in addrs[h:h + 16] if a is not None]
        if not half:
            continue
        banks = {}
        for a in set(half):
            banks[a % 16] = banks.get(a % 16, 0) + 1
        total += max(banks.values())
    return total


def stage_cost(ntasks, TS, addr_fns):
    """Sum of wavefronts over all task rounds and all accesses of a stage; addr_fns: list of (fn(t) -> addr, weight)."""
    cost = ideal = 0
    for base in range(0, ntasks, TS):
        for w0 in range(base, min(base + TS, ntasks), 32):
            lanes = [t if t < ntasks else None for t in range(w0, w0 + 32)]
            for fn, weight in addr_fns:
                addrs = [fn(t) if t is not None else None for t in lanes]
                cost += weight * wavefronts(addrs)
                n = sum(a is not None for a in addrs)
                ideal += weight * ((n + 15) // 16)
    return cost, ideal


def model(P, Q, E, TS, Sz1, Ps, Sz2, Qs, S, collocated=False, grad=True):
    """Stages of a scalar (nc = 1) operator; returns dict stage -> (wavefronts, ideal)."""
    out = {}
    QQs = Q * Qs
    pl = lambda k, le: (k * E + le) * S
    if not collocated:
        # S2 y: t -> i, qz, le ; reads T1 (A), writes T2 (B)
        def dec(t):
            return t % P, (t // P) % Q, t // (P * Q)
        fns = []
        for j in range(P):
            fns.append((lambda t, j=j: pl(0, dec(t)[2]) + dec(t)[1] * Sz1 + j * P + dec(t)[0], 1))
        for q in range(Q):
            fns.append((lambda t, q=q: pl(1, dec(t)[2]) + dec(t)[1] * Sz2 + q * Ps + dec(t)[0], 1))
        out["y / yT"] = tuple(2 * x for x in stage_cost(E * P * Q, TS, fns))  # the transpose stage mirrors it
        # S3 x: t -> row, le ; reads T2 rows, writes Uq (+gx)
        def decr(t):
            return t % (Q * Q), t // (Q * Q)
        fns = []
        for i in range(P):
            fns.append((lambda t, i=i: pl(1, decr(t)[1]) + (decr(t)[0] // Q) * Sz2 + (decr(t)[0] % Q) * Ps + i, 1))
        for q in range(Q):
            fns.append((lambda t, q=q: pl(0, decr(t)[1]) + (decr(t)[0] // Q) * QQs + (decr(t)[0] % Q) * Qs + q, 2 if grad else 1))
        out["x / xT"] = tuple(2 * x for x in stage_cost(E * Q * Q, TS, fns))
    if grad:
        # S4 grad_y: t -> qx, qz, le
        def decy(t):
            return t % Q, (t // Q) % Q, t // (Q * Q)
        fns = [(lambda t, m=m: pl(0, decy(t)[2]) + decy(t)[1] * QQs + m * Qs + decy(t)[0], 2) for m in range(Q)]
        out["dy / dyT"] = tuple(2 * x for x in stage_cost(E * Q * Q, TS, fns))
        # S5 z-line QF: t -> qx, qy, le
        def decz(t):
            return t % Q, (t // Q) % Q, t // (Q * Q)
        fns = [(lambda t, m=m: pl(0, decz(t)[2]) + m * QQs + decz(t)[1] * Qs + decz(t)[0], 2 + 4) for m in range(Q)]
        out["qf z-line"] = stage_cost(E * Q * Q, TS, fns)
    # S1/S9 gather/scatter z: t -> ij, le
    def decg(t):
        return t % (P * P), t // (P * P)
    if collocated:
        fns = [(lambda t, k=k: pl(0, decg(t)[1]) + k * QQs + (decg(t)[0] // P) * Qs + decg(t)[0] % P, 2) for k in range(P)]
    else:
        fns = [(lambda t, q=q: pl(0, decg(t)[1]) + q * Sz1 + decg(t)[0], 2) for q in range(Q)]
    out["gather/scatter z"] = stage_cost(E * P * P, TS, fns)
    return out


def odd(n):
    return n if n % 2 else n + 1


def total(res):
    return sum(v[0] for v in res.values()), sum(v[1] for v in res.values())


def current(P, Q, E, TS, collocated=False, grad=True):
    Qs, Ps = odd(Q), odd(P)
    S = max(Q * Q * Qs, Q * P * P, Q * Q * Ps)
    return model(P, Q, E, TS, P * P, Ps, Q * Ps, Qs, S, collocated, grad), S


def search(P, Q, E, TS, collocated=False, grad=True):
    best = None
    for Qs in (Q, Q + 1, Q + 2):
        for Ps in (P, P + 1, P + 2):
            for p1 in range(0, 16):
                for p2 in range(0, 16):
                    Sz1, Sz2 = P * P + p1, Q * Ps + p2
                    S0 = max(Q * Q * Qs, Q * Sz1, Q * Sz2)
                    for ps in range(0, 16 if E > 1 else 1):
                        S = S0 + ps
                        c, i = total(model(P, Q, E, TS, Sz1, Ps, Sz2, Qs, S, collocated, grad))
                        key = (c, S)
                        if best is None or key < best[0]:
                            best = (key, dict(Qs=Qs, Ps=Ps, Sz1=Sz1, Sz2=Sz2, S=S), c, i)
                    if collocated:
                        break
                if collocated:
                    break
            if collocated:
                break
    return best


if __name__ == "__main__":
    cases = [(4, 5, 2, 32, False, False), (2, 3, 14, 32, False, True), (3, 4, 6, 32, False, True), (4, 5, 5, 64, False, True), (5, 6, 7, 128, False, True),
             (6, 7, 1, 32, False, True), (7, 8, 1, 64, False, True), (8, 9, 1, 32, False, True), (9, 10, 1, 128, False, True),
             (5, 5, 5, 64, True, True), (6, 6, 3, 64, True, True), (7, 7, 1, 32, True, True), (8, 8, 1, 64, True, True)]
    for P, Q, E, TS, col, grad in cases:
        res, S = current(P, Q, E, TS, col, grad)
        c, i = total(res)
        print(f"P{P} Q{Q} E{E} TS{TS} {'c' if col else ' '}: current S={S} wavefronts {c} ideal {i} ({c / i:.2f}x)  " +
              "  ".join(f"{k}: {v[0] / max(v[1], 1):.2f}x" for k, v in res.items()))
        if "--search" in sys.argv:
            b = search(P, Q, E, TS, col, grad)
            print(f"      best {b[1]} -> {b[2]} ({b[2] / b[3]:.2f}x)")
