"""Bank-conflict model of the XOR-swizzled quadrature-point planes (development aid).
Layout of a Uq-type plane for Q <= 8: index(qx, qy, qz) = qz * SZ + qy * 8 + (qx ^ qy), SZ = 8 (mod 16).
T1 [qz][j][i] -> qz * Sz1 + j * P + i ; T2 [qz][qy][i] -> qz * Sz2 + qy * Ps + i.
Lane enumerations use a padded fastest extent A8 = 8 where noted (idle lanes for i >= P / qx >= Q)."""
import sys
sys.path.insert(0, "scripts/model")
from bank_conflicts import wavefronts


def cost(lanes, fns):
    c = i = 0
    for w0 in range(0, len(lanes), 32):
        ws = lanes[w0:w0 + 32]
        for fn, wgt in fns:
            addrs = [fn(*t) if t is not None else None for t in ws]
            c += wgt * wavefronts(addrs)
            i += wgt * ((sum(a is not None for a in addrs) + 15) // 16)
    return c, i


def run(P, Q, collocated, Sz1, Sz2, Ps, SZ, TS):
    uq = lambda x, y, z: z * SZ + y * 8 + (x ^ y)
    out = {}
    pad = lambda n, A: [(t % 8, t // 8) if t % 8 < A else None for t in range(n * 8)]
    if not collocated:
        lanes = pad(Q, P)  # (i, qz), i padded to 8
        fns = [(lambda i, z, j=j: z * Sz1 + j * P + i, 1) for j in range(P)] + [(lambda i, z, q=q: z * Sz2 + q * Ps + i, 1) for q in range(Q)]
        out["y/yT"] = tuple(2 * v for v in cost(lanes, fns))
        lanes = pad(Q, Q)  # row = (qy, qz), qy padded to 8
        fns = [(lambda y, z, i=i: z * Sz2 + y * Ps + i, 1) for i in range(P)] + [(lambda y, z, q=q: uq(q, y, z), 2) for q in range(Q)]
        out["x/xT"] = tuple(2 * v for v in cost(lanes, fns))
    lanes = pad(Q, Q)  # (qx, qz) padded
    out["dy/dyT"] = tuple(2 * v for v in cost(lanes, [(lambda x, z, m=m: uq(x, m, z), 2) for m in range(Q)]))
    lanes = pad(Q, Q)  # (qx, qy) padded
    out["qf zline"] = cost(lanes, [(lambda x, y, m=m: uq(x, y, m), 6) for m in range(Q)])
    if collocated:
        lanes = pad(P, P)
        out["gather/scatter"] = cost(lanes, [(lambda i, j, k=k: uq(i, j, k), 2) for k in range(P)])
    else:
        lanes = [(t,) for t in range(P * P)]
        out["gather/scatter"] = cost(lanes, [(lambda ij, q=q: q * Sz1 + ij, 2) for q in range(Q)])
    tot = (sum(v[0] for v in out.values()), sum(v[1] for v in out.values()))
    print(f"P{P} Q{Q} {'c' if collocated else ' '} Sz1={Sz1} Sz2={Sz2} Ps={Ps} SZ={SZ}: total {tot[0]} ideal {tot[1]} ({tot[0]/tot[1]:.2f}x)", {k: f"{v[0]}/{v[1]}" for k, v in out.items()})
    return tot


if __name__ == "__main__":
    run(7, 8, False, 56, 56, 7, 72, 64)
    run(7, 8, False, 56, 72, 9, 72, 64)
    run(8, 8, True, 0, 0, 0, 72, 64)
    run(6, 7, False, 40, 56, 7, 56, 64)
    run(6, 7, False, 40, 56, 8, 56, 64)
    run(7, 7, True, 0, 0, 0, 56, 64)
    run(5, 6, False, 40, 56, 7, 56, 64)
    run(6, 6, True, 0, 0, 0, 56, 64)
