"""Synthetic structured hexahedral meshes for the BP benchmarks (host-side, numpy only).

Behavioural reference: the Cartesian mesh builder of examples/ceed/ex2-surface.c:314-433 (lexicographic node
numbering, offsets x-fastest, smooth coordinate perturbation) and the process-grid split of
examples/petsc/bpsraw.c:48-107.  Unlike ex2-surface the element counts are not rounded to powers of two: they are
chosen so that the number of DoFs is close to the request and the box is near-cubic.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def gll_nodes(P):
    """P Gauss-Legendre-Lobatto points on [-1, 1] (product host routine; no GPU needed)."""
    if P == 1:
        return np.zeros(1)
    x = np.zeros(P)
    L.lib().ceedb200_host_lobatto_quadrature(P, x.ctypes.data_as(L.c_scalar_p), None)
    return x


def choose_elements(num_dofs, p, ncomp=1):
    """(nx, ny, nz) with (nx p + 1)(ny p + 1)(nz p + 1) ncomp ~ num_dofs, near-cubic."""
    nodes = max(1.0, num_dofs / ncomp)
    n = max(1, int(round((nodes ** (1.0 / 3.0) - 1) / p)))
    best = None
    for nx in range(max(1, n - 1), n + 2):
        for ny in range(max(1, n - 1), n + 2):
            nz = max(1, int(round((nodes / ((nx * p + 1) * (ny * p + 1)) - 1) / p)))
            tot = (nx * p + 1) * (ny * p + 1) * (nz * p + 1)
            cost = abs(tot - nodes) / nodes + 0.01 * (abs(nx - ny) + abs(ny - nz))
            if best is None or cost < best[0]:
                best = (cost, (nx, ny, nz))
    return best[1]


def hex_offsets(nx, ny, nz, p):
    """int32 [nelem, P^3] element -> node map, x fastest inside an element, elements x fastest."""
    P = p + 1
    NX, NY = nx * p + 1, ny * p + 1
    i = np.arange(P)
    local = (i[None, None, :] + NX * i[None, :, None] + NX * NY * i[:, None, None]).reshape(-1)  # k, j, i
    # element order: ex fastest, e = (ez * ny + ey) * nx + ex
    ez = np.arange(nz, dtype=np.int64)[:, None, None]
    ey = np.arange(ny, dtype=np.int64)[None, :, None]
    ex = np.arange(nx, dtype=np.int64)[None, None, :]
    base = (ez * p * NX * NY + ey * p * NX + ex * p).reshape(-1)
    off = base[:, None] + local[None, :]
    assert off.max() < 2 ** 31
    return off.astype(np.int32)


def morton_permutation(nx, ny, nz):
    """Element permutation (new position -> lexicographic element number) that orders the elements of an nx x ny x nz box
    along the Z-order (Morton) curve: any contiguous range of elements is a compact blob, the locality-preserving ordering
    applications use for cache reuse.  The deterministic run scatter of the lean kernel (B200RunScatter) profits from it: more
    E-entries have all their earlier touchers inside the same warp's run."""
    def spread(v):
        v = v.astype(np.uint64)
        out = np.zeros_like(v)
        for b in range(21):
            out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        return out
    ez, ey, ex = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    key = spread(ex.reshape(-1)) | (spread(ey.reshape(-1)) << np.uint64(1)) | (spread(ez.reshape(-1)) << np.uint64(2))
    return np.argsort(key, kind="stable").astype(np.int64)


def hex_coords(nx, ny, nz, p, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), n_global=None, e0=(0, 0, 0), perturb=True):
    """[3, nnodes] node coordinates: GLL points in every element of a uniform grid on [lo, hi], then the smooth map
    x -> 0.5 + sin(2 pi / 3 (x - 0.5)) / sqrt(3) of ex2-surface.c:424-433 so that qdata is non-trivial.
    (n_global, e0) place a sub-box of a larger global grid (used by the multi-GPU partition)."""
    g = (gll_nodes(p + 1) + 1.0) / 2.0
    n_global = n_global or (nx, ny, nz)

    def axis(n, ng, start, a, b):
        h = (b - a) / ng
        pts = np.zeros(n * p + 1)
        for e in range(n):
            pts[e * p:(e + 1) * p + 1] = a + (start + e + g) * h
        return pts

    xs = axis(nx, n_global[0], e0[0], lo[0], hi[0])
    ys = axis(ny, n_global[1], e0[1], lo[1], hi[1])
    zs = axis(nz, n_global[2], e0[2], lo[2], hi[2])
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)])
    if perturb:
        coords = 0.5 + np.sin(2.0 * np.pi / 3.0 * (coords - 0.5)) / np.sqrt(3.0)
    return coords


def split3(n):
    """Process grid for n ranks: 1->1x1x1, 2->2x1x1, 4->2x2x1, 8->2x2x2 ... (examples/petsc/bpsraw.c:48-56 Split3)."""
    dims = [1, 1, 1]
    f, d = n, 0
    factors = []
    q = 2
    while f > 1:
        while f % q == 0:
            factors.append(q)
            f //= q
        q += 1
    for fac in sorted(factors, reverse=True):
        dims[d % 3] *= fac
        d += 1
    return tuple(dims)


def block_range(n, parts, idx):
    """[start, stop) of block idx when n items are split into `parts` nearly equal contiguous blocks."""
    base, rem = divmod(n, parts)
    start = idx * base + min(idx, rem)
    return start, start + base + (1 if idx < rem else 0)


class Partition:
    """Element-wise 3-D block partition of a global nx x ny x nz hex mesh over `size` ranks.

    Every rank owns a box of elements and a LOCAL L-vector that includes a copy of the interface nodes
    (SURVEY.md section 8(e)).  `neighbors` lists, for every rank sharing nodes with this one, the local node indices of
    the shared nodes in an order both sides agree on (global lexicographic), so that after a local operator apply the
    interface values can be exchanged and summed in a fixed rank order (deterministic).
    """

    def __init__(self, n_global, p, size, rank):
        self.n_global, self.p, self.size, self.rank = tuple(n_global), p, size, rank
        self.grid = split3(size)
        px, py, pz = self.grid
        self.coord = (rank % px, (rank // px) % py, rank // (px * py))
        self.ranges = [block_range(n_global[d], self.grid[d], self.coord[d]) for d in range(3)]
        self.n_local = tuple(r[1] - r[0] for r in self.ranges)
        self.e0 = tuple(r[0] for r in self.ranges)
        self.node_lo = tuple(r[0] * p for r in self.ranges)
        self.node_n = tuple((r[1] - r[0]) * p + 1 for r in self.ranges)
        self.neighbors = self._neighbors()

    def _rank_of(self, c):
        px, py, _ = self.grid
        return c[0] + px * (c[1] + py * c[2])

    def _neighbors(self):
        out = []
        NX, NY, _ = self.node_n
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if dx == dy == dz == 0:
                        continue
                    c = (self.coord[0] + dx, self.coord[1] + dy, self.coord[2] + dz)
                    if any(c[d] < 0 or c[d] >= self.grid[d] for d in range(3)):
                        continue
                    sel = []
                    for d, delta in enumerate((dx, dy, dz)):
                        n = self.node_n[d]
                        sel.append(np.array([0]) if delta < 0 else (np.array([n - 1]) if delta > 0 else np.arange(n)))
                    idx = (sel[2][:, None, None] * NY + sel[1][None, :, None]) * NX + sel[0][None, None, :]
                    out.append((self._rank_of(c), idx.reshape(-1).astype(np.int64)))
        out.sort(key=lambda t: t[0])
        return out

    def boundary_first_permutation(self):
        """(perm, n_boundary): order of the local elements with the ones touching a rank interface first (each class keeps its
        lexicographic order).  Used for the overlapped multi-GPU step: the boundary elements are applied first, their interface
        values travel while the interior elements are applied (the reference pattern is
        VecScatterBegin / local work / VecScatterEnd, examples/petsc/bpsraw.c:240-262)."""
        nx, ny, nz = self.n_local
        ez, ey, ex = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        touch = np.zeros((nz, ny, nx), dtype=bool)
        for d, e in enumerate((ex, ey, ez)):
            if self.coord[d] > 0:
                touch |= e == 0
            if self.coord[d] < self.grid[d] - 1:
                touch |= e == self.n_local[d] - 1
        touch = touch.reshape(-1)
        perm = np.concatenate([np.nonzero(touch)[0], np.nonzero(~touch)[0]]).astype(np.int64)
        return perm, int(touch.sum())

    @property
    def num_local_nodes(self):
        return int(np.prod(self.node_n))

    def owned_mask(self):
        """True for local nodes this rank owns (lowest rank touching a shared node owns it): used to count global DoFs
        and to compare against a single-rank run."""
        mask = np.ones(self.node_n[::-1], dtype=bool)  # [z, y, x]
        for d in range(3):
            if self.coord[d] > 0:
                sl = [slice(None)] * 3
                sl[2 - d] = 0
                mask[tuple(sl)] = False
        return mask.reshape(-1)

    def global_node_ids(self):
        """Global lexicographic node id of every local node."""
        GX, GY = self.n_global[0] * self.p + 1, self.n_global[1] * self.p + 1
        x = self.node_lo[0] + np.arange(self.node_n[0])
        y = self.node_lo[1] + np.arange(self.node_n[1])
        z = self.node_lo[2] + np.arange(self.node_n[2])
        return ((z[:, None, None] * GY + y[None, :, None]) * GX + x[None, None, :]).reshape(-1)
