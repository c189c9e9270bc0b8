"""TEST INFRASTRUCTURE -- numpy restatement of marching cubes as the reference uses it:
`vertices, triangles = mcubes.marching_cubes(u, threshold)` at lib/models/renderers/NeuS.py:35, followed by the
index -> bbox affine at NeuS.py:36-39.  Never imported by the product path (color_neus_b200/).

PARITY UNPINNED: PyMCubes (pymcubes==0.1.4, requirements.txt:13) is a third-party dependency that is neither vendored
under /root/reference nor installed in this image, and the reference holds no test or golden mesh for it.  What is
restated here is the published algorithm (Lorensen & Cline 1987: per-cell sign case -> table of triangles over the 12
cube edges, vertices by linear interpolation along the cut edges, shared between cells), with these documented choices:
  * corner numbering / edge numbering of the classic tables (corner i at (x,y,z) offsets below, edge e between
    EDGE_CORNERS[e]); a corner is "inside" when value < isovalue, exactly the classic `if (val < iso) cubeindex |= 1<<i`;
  * the per-case triangulation is DERIVED (build_tables) instead of copied from a printed table: on every cube face the
    cut edges are joined by segments, ambiguous faces (4 cut edges) always cut off the inside corners, segments are
    chained into closed loops and each loop is fan-triangulated.  The face rule depends on the face's corner signs only,
    so neighbouring cells agree and the mesh is watertight (the classic printed table can leave cracks on ambiguous
    faces); in unambiguous cells -- all cells of a smooth SDF away from thin features -- the surface is the classic one;
  * triangles are oriented with their normal towards the value < isovalue side (u = -sdf, NeuS.py:415: outwards);
  * vertex on the edge from grid point a (value f1) to a + e_axis (value f2): a + (iso - f1) / (f2 - f1) along the axis,
    in float64 index coordinates like PyMCubes' double vertices;
  * output order (ours): vertices by (owning grid point in x-major order, axis), triangles by (cell, table order).
Property tests (tests/test_marching_cubes.py) check what does not depend on these choices: closed 2-manifold, vertices on
the iso-surface of the trilinear interpolant's edges, Euler characteristic and enclosed volume of analytic shapes.
"""
import numpy as np

CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
EDGE_CORNERS = np.array([[0, 1], [1, 2], [3, 2], [0, 3], [4, 5], [5, 6], [7, 6], [4, 7], [0, 4], [1, 5], [2, 6], [3, 7]])
# owning grid point (offset from the cell's corner 0) and axis of every edge
EDGE_OWNER = np.array([[0, 0, 0, 0], [1, 0, 0, 1], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [1, 0, 1, 1], [0, 1, 1, 0],
                       [0, 0, 1, 1], [0, 0, 0, 2], [1, 0, 0, 2], [1, 1, 0, 2], [0, 1, 0, 2]])
# faces as corner cycles, counter-clockwise when seen from outside the cube
FACES = [[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [2, 3, 7, 6], [0, 4, 7, 3], [1, 2, 6, 5]]


def _edge_between(a, b):
    for e, (p, q) in enumerate(EDGE_CORNERS):
        if (p == a and q == b) or (p == b and q == a):
            return e
    raise KeyError((a, b))


def case_triangles(case):
    """Triangles (list of edge triples) of one sign case; bit i of `case` = corner i inside (value < iso)."""
    inside = [(case >> i) & 1 for i in range(8)]
    nxt = {}   # directed segments: cut edge -> next cut edge, inside region on the left seen from outside the cube
    for f in FACES:
        # walk the face boundary counter-clockwise; a transition inside -> outside on (c_i, c_i+1) starts a segment that
        # ends on the next outside -> inside transition found walking BACKWARDS round the inside run, i.e. the segment
        # leaves through the edge where the run of inside corners ends and keeps that run on its left.
        cuts = []   # (edge, kind) in ccw order; kind +1: inside -> outside, -1: outside -> inside
        for i in range(4):
            a, b = f[i], f[(i + 1) % 4]
            if inside[a] != inside[b]:
                cuts.append((_edge_between(a, b), +1 if inside[a] else -1))
        if not cuts:
            continue
        # every maximal run of inside corners is bounded by an "enter" cut (out -> in) before it and a "leave" cut
        # (in -> out) after it (ccw).  Cutting the run off with the inside on the left means travelling from the leave
        # cut back to the enter cut.  With 4 cuts there are two runs of one corner each: both are cut off separately.
        n = len(cuts)
        for i, (e, kind) in enumerate(cuts):
            if kind == -1:                      # enter cut of a run; its leave cut is the next cut ccw
                leave = cuts[(i + 1) % n]
                assert leave[1] == +1
                assert leave[0] not in nxt
                nxt[leave[0]] = e
    tris, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [], start
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == start and len(loop) >= 3
        tris += [tuple(loop[i] for i in t) for t in _triangulate(loop)]
    return tris


_EDGE_FACES = [{i for i, f in enumerate(FACES) if a in f and b in f} for a, b in EDGE_CORNERS]


def _polygon_triangulations(i, j):
    """All triangulations of the polygon with vertices i..j (index triples in increasing order keep the orientation)."""
    if j - i < 2:
        return [[]]
    return [l + [(i, k, j)] + r for k in range(i + 1, j) for l in _polygon_triangulations(i, k) for r in _polygon_triangulations(k, j)]


def _triangulate(loop):
    """Triangulation of one loop of cut edges without a diagonal that lies in a cube face: such a diagonal could coincide with
    a face segment of the neighbouring cell (a mesh edge used four times).  Every loop of every case has one (checked by the
    assert); among those the first in enumeration order is taken (a fan whenever a fan qualifies)."""
    n = len(loop)
    best = None
    for T in _polygon_triangulations(0, n - 1):
        diags = {(loop[p], loop[q]) for t in T for p, q in ((t[0], t[1]), (t[1], t[2]), (t[0], t[2])) if (q - p) % n not in (1, n - 1)}
        if any(_EDGE_FACES[a] & _EDGE_FACES[b] for a, b in diags):
            continue
        is_fan = any(all(a in t for t in T) for a in range(n))
        if best is None or (is_fan and not best[0]):
            best = (is_fan, T)
            if is_fan:
                break
    assert best is not None
    return best[1]


_TABLES = None


def build_tables():
    """(n_tri[256] uint8, tri_table[256, 15] int8 padded with -1, edge_mask[256] uint16)."""
    global _TABLES
    if _TABLES is None:
        n_tri = np.zeros(256, np.uint8)
        table = -np.ones((256, 15), np.int8)
        mask = np.zeros(256, np.uint16)
        for c in range(256):
            t = case_triangles(c)
            assert len(t) <= 5
            n_tri[c] = len(t)
            flat = [e for tri in t for e in tri]
            table[c, :len(flat)] = flat
            for e in flat:
                mask[c] |= 1 << e
        _TABLES = (n_tri, table, mask)
    return _TABLES


def marching_cubes(u, isovalue=0.0):
    """-> (vertices float64 [V,3] in index coordinates, triangles int64 [F,3]); see the module docstring for conventions."""
    u = np.asarray(u, dtype=np.float64)   # PyMCubes works on doubles; float32 -> float64 is exact
    nx, ny, nz = u.shape
    iso = float(isovalue)
    n_tri, table, _ = build_tables()
    ins = u < iso
    # --- vertices: owned edges of every grid point, order (point, axis)
    cut = np.zeros((nx, ny, nz, 3), bool)
    cut[:-1, :, :, 0] = ins[:-1] != ins[1:]
    cut[:, :-1, :, 1] = ins[:, :-1] != ins[:, 1:]
    cut[:, :, :-1, 2] = ins[:, :, :-1] != ins[:, :, 1:]
    flat_cut = cut.reshape(-1)
    vid = np.cumsum(flat_cut) - 1          # vertex id of (point, axis) where cut
    vid = vid.reshape(nx, ny, nz, 3)
    px, py, pz, ax = np.nonzero(cut)
    f1 = u[px, py, pz]
    q = np.stack([px, py, pz], 1)
    q2 = q.copy()
    q2[np.arange(len(ax)), ax] += 1
    f2 = u[q2[:, 0], q2[:, 1], q2[:, 2]]
    t = (iso - f1) / (f2 - f1)
    verts = q.astype(np.float64)
    verts[np.arange(len(ax)), ax] += t
    # --- triangles: cells in x-major order
    if min(nx, ny, nz) < 2:
        return verts, np.zeros((0, 3), np.int64)
    case = np.zeros((nx - 1, ny - 1, nz - 1), np.int32)
    for i, (dx, dy, dz) in enumerate(CORNERS):
        case |= ins[dx:nx - 1 + dx, dy:ny - 1 + dy, dz:nz - 1 + dz].astype(np.int32) << i
    cx, cy, cz = np.nonzero((case != 0) & (case != 255))
    cc = case[cx, cy, cz]
    out = []
    for k in range(5):
        sel = n_tri[cc] > k
        if not sel.any():
            break
        x, y, z, c = cx[sel], cy[sel], cz[sel], cc[sel]
        tri = np.empty((len(c), 3), np.int64)
        for j in range(3):
            e = table[c, 3 * k + j].astype(np.int64)
            o = EDGE_OWNER[e]
            tri[:, j] = vid[x + o[:, 0], y + o[:, 1], z + o[:, 2], o[:, 3]]
        cell_lin = (x * ny + y) * nz + z     # grid-point linear index of the cell's corner 0
        out.append((cell_lin, np.full(len(c), k), tri))
    if not out:
        return verts, np.zeros((0, 3), np.int64)
    lin = np.concatenate([o[0] for o in out])
    kk = np.concatenate([o[1] for o in out])
    tri = np.concatenate([o[2] for o in out])
    order = np.lexsort((kk, lin))
    return verts, tri[order]


def index_to_bbox(vertices, resolution, bound_min, bound_max):
    """NeuS.py:36-39."""
    b_min, b_max = np.asarray(bound_min), np.asarray(bound_max)
    return vertices / (resolution - 1.0) * (b_max - b_min)[None, :] + b_min[None, :]


# ---- mesh invariants used by the tests ---------------------------------------------------------------------------------
def edge_use_counts(triangles):
    """{undirected edge: count}, plus whether every directed edge is matched by its reverse exactly once."""
    t = np.asarray(triangles, np.int64)
    d = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    und = np.sort(d, axis=1)
    _, cnt = np.unique(und, axis=0, return_counts=True)
    dir_keys = d[:, 0] * (t.max() + 1) + d[:, 1]
    rev_keys = d[:, 1] * (t.max() + 1) + d[:, 0]
    oriented = len(np.unique(dir_keys)) == len(dir_keys) and np.array_equal(np.sort(dir_keys), np.sort(rev_keys))
    return cnt, oriented


def signed_volume(vertices, triangles):
    v = np.asarray(vertices)
    a, b, c = v[triangles[:, 0]], v[triangles[:, 1]], v[triangles[:, 2]]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)
