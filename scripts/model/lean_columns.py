"""Model of the lean kernel's gather / scatter column stages (b200_opgen_lean.cpp, stage bits 1 and 2): 128-byte lines touched per
warp request by the gather of u (the L1 data pipe handles one line per wavefront) and shared-memory wavefronts of the plane stores,
for the default (i, j, element) and the element-interleaved (i, element, j) lane maps.  No GPU needed.
usage: python scripts/model/lean_columns.py [p] [E ...]"""
import sys
import numpy as np

p = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Es = [int(a) for a in sys.argv[2:]] or [4, 6, 8]
P, Q = p + 1, p + 2
nx, ny, nz = 72, 71, 71  # the 10 M-DoF BP1 p=3 mesh is 72 x 71 x 71 elements; the model only needs a few rows of it
NX, NY = nx * p + 1, ny * p + 1


def offsets(e):
    ex, ey, ez = e % nx, (e // nx) % ny, e // (nx * ny)
    i, j, k = np.meshgrid(np.arange(P), np.arange(P), np.arange(P), indexing="ij")
    return ((ex * p + i) + NX * ((ey * p + j) + NY * (ez * p + k)))  # [i][j][k]


def decode(t, E, ilv):
    if ilv:
        return t % P, (t // (P * E)), (t // P) % E  # i, j, le
    ij = t % (P * P)
    return ij % P, ij // P, t // (P * P)


def smem_wavefronts(addr):
    """8-byte accesses: a warp request needs max over the two half-warps ... modelled as the sum over half-warps of the worst bank multiplicity"""
    w = 0
    for h in (addr[:16], addr[16:]):
        if len(h) == 0: continue
        banks = {}
        for a in set(h.tolist()): banks[a % 16] = banks.get(a % 16, 0) + 1
        w += max(banks.values())
    return w


print(f"p={p} P={P} Q={Q}")
for E in Es:
    for ilv in (0, 1):
        for pad in ((0, 1) if ilv else (0,)):
            ES = Q * Q * P
            if pad:
                while ES % 16 != P % 16: ES += 1
            T = E * P * P
            lines = reqs = sw = sreq = ilines = 0
            for e0 in range(nx * ny * 3 + 5, nx * ny * 3 + 5 + 40 * E, E):
                offs = [offsets(e0 + le) for le in range(E)]
                for r in range((T + 31) // 32):
                    t = np.arange(32 * r, min(T, 32 * r + 32))
                    i, j, le = decode(t, E, ilv)
                    for k in range(P):
                        a = np.array([offs[l][ii, jj, k] for ii, jj, l in zip(i, j, le)]) * 8
                        lines += len(set((a // 128).tolist()))
                        reqs += 1
                        # the int32 offsets (and, identically, the scatter targets) of the same request: element-major table
                        ia = ((e0 + le) * P ** 3 + k * P * P + j * P + i) * 4
                        ilines += len(set((ia // 128).tolist()))
                    sa = le * ES + j * P + i
                    sw += smem_wavefronts(sa) * Q
                    sreq += Q
            nelem = 40 * E
            print(f"  E={E} interleaved={ilv} pad={pad} ES={ES}: gather lines/request {lines/reqs:5.2f}, lines/element {lines/nelem:6.2f}, offset-table lines/element {ilines/nelem:5.2f}; "
                  f"plane-store wavefronts/request {sw/sreq:4.2f} (ideal 2), per element {sw/nelem:5.2f}")
