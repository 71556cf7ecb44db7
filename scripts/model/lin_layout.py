"""Shared-memory wavefront model of the quadrature-plane stages for the padded layout (odd row pitch) against the even-Q linear layout
(stage bit 512: unpadded rows, z-stride = Q (mod 16), x-lines as 16-byte accesses).  usage: python scripts/model/lin_layout.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bank_conflicts import wavefronts, odd


def wf128(addrs):
    """16-byte accesses: served per quarter-warp (8 lanes), 8 sixteen-byte units per wavefront"""
    total = 0
    for h in range(0, len(addrs), 8):
        q = [a for a in addrs[h:h + 8] if a is not None]
        if not q:
            continue
        banks = {}
        for a in set(q):
            banks[(a // 2) % 8] = banks.get((a // 2) % 8, 0) + 1
        total += max(banks.values())
    return total


def cost(ntasks, TS, fns, wide=False):
    c = i = 0
    for base in range(0, ntasks, TS):
        for w0 in range(base, min(base + TS, ntasks), 32):
            lanes = [t if t < ntasks else None for t in range(w0, w0 + 32)]
            for fn, wgt in fns:
                ad = [fn(t) if t is not None else None for t in lanes]
                n = sum(a is not None for a in ad)
                c += wgt * (wf128(ad) if wide else wavefronts(ad))
                i += wgt * ((n + 7) // 8 if wide else (n + 15) // 16)
    return c, i


def quad_part(Q, E, TS, Qs, SZ, S, x128):
    """x rows (values + derivative, forward and transpose), y-lines (d/dy and its transpose), z-line QFunction stage of a scalar gradient operator"""
    dr = lambda t: (t % Q, (t // Q) % Q, t // (Q * Q))
    out = {}
    fns = [(lambda t, q=q: dr(t)[2] * S + dr(t)[1] * SZ + dr(t)[0] * Qs + q, 2) for q in range(0, Q, 2 if x128 else 1)]
    out["x"] = tuple(2 * v for v in cost(E * Q * Q, TS, fns, x128))
    fns = [(lambda t, m=m: dr(t)[2] * S + dr(t)[1] * SZ + m * Qs + dr(t)[0], 2) for m in range(Q)]
    out["y"] = tuple(2 * v for v in cost(E * Q * Q, TS, fns))
    fns = [(lambda t, m=m: dr(t)[2] * S + m * SZ + dr(t)[1] * Qs + dr(t)[0], 6) for m in range(Q)]
    out["z"] = cost(E * Q * Q, TS, fns)
    return out


if __name__ == "__main__":
    for Q, E, TS in ((6, 7, 128), (6, 3, 64), (10, 1, 128), (8, 1, 32), (4, 6, 32)):
        for name, Qs, SZ, x128 in (("padded", odd(Q), Q * odd(Q), False), ("linear", Q, None, True)):
            if SZ is None:
                SZ = Q * Q
                while SZ % 16 != Q % 16:
                    SZ += 1
            S = Q * SZ
            r = quad_part(Q, E, TS, Qs, SZ, S, x128)
            tot, ideal = sum(v[0] for v in r.values()), sum(v[1] for v in r.values())
            print(f"Q{Q} E{E} lanes {TS} {name:7s} pitch {Qs} z-stride {SZ}: {tot} wavefronts, ideal {ideal} ({tot / ideal:.2f}x)  " +
                  " ".join(f"{k}:{v[0] / v[1]:.2f}x" for k, v in r.items()))
