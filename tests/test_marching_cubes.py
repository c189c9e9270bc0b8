"""marching cubes + PLY.  CPU: the oracle's invariants on analytic / random fields, the library's derived case
table against the oracle's independent derivation (host code, no GPU), PLY round trip.  GPU: device mesh == oracle mesh
(indices bit-exact, float64 vertices bit-exact: same IEEE operations), invariants at a size the oracle does not run."""
import os

import numpy as np
import pytest
import torch

from oracle import mc_oracle as M


def sphere_field(n, r=0.6, c=(0.05, -0.02, 0.03)):
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return (r - np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2)).astype(np.float32)   # u = -sdf


def torus_field(n, R=0.55, r=0.2):
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return (r - np.sqrt((np.sqrt(X ** 2 + Y ** 2) - R) ** 2 + Z ** 2)).astype(np.float32)


def noise_field(shape, seed, closed=True):
    u = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    if closed:   # keep the surface away from the grid boundary so that it is closed
        u[0] = u[-1] = -5
        u[:, 0] = u[:, -1] = -5
        u[:, :, 0] = u[:, :, -1] = -5
    return u


def check_closed_manifold(v, f):
    cnt, oriented = M.edge_use_counts(f)
    assert set(np.unique(cnt)) == {2}, "every mesh edge must be shared by exactly two triangles"
    assert oriented, "neighbouring triangles must traverse their common edge in opposite directions"
    return len(v) - len(cnt) + len(f)   # Euler characteristic


def test_oracle_sphere_and_torus():
    n = 40
    v, f = M.marching_cubes(sphere_field(n), 0.0)
    assert check_closed_manifold(v, f) == 2
    vw = M.index_to_bbox(v, n, [-1, -1, -1], [1, 1, 1])
    r = np.linalg.norm(vw - np.array([0.05, -0.02, 0.03]), axis=1)
    assert np.abs(r - 0.6).max() < 2e-3          # linear interpolation of a distance field along grid edges
    vol = M.signed_volume(vw, f)
    assert 0 < vol and abs(vol - 4 / 3 * np.pi * 0.6 ** 3) < 0.01   # positive: normals towards u < iso, i.e. outwards
    v, f = M.marching_cubes(torus_field(n), 0.0)
    assert check_closed_manifold(v, f) == 0      # genus 1


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_noise_is_watertight(seed):
    """Noise exercises every ambiguous face / cell configuration: the derived table must still give a closed oriented
    2-manifold (the classic printed table does not)."""
    v, f = M.marching_cubes(noise_field((14, 17, 12), seed), 0.0)
    check_closed_manifold(v, f)


def test_oracle_vertices_interpolate_the_field():
    u = noise_field((9, 8, 10), 5, closed=False)
    iso = 0.137
    v, f = M.marching_cubes(u, iso)
    frac = v - np.floor(v)
    on_axis = (frac > 0).sum(1)
    assert on_axis.max() <= 1
    for p in v[:200]:
        a = np.floor(p).astype(int)
        ax = int(np.argmax(p - a)) if (p - a).max() > 0 else 0
        b = a.copy()
        b[ax] += 1
        f1, f2 = float(u[tuple(a)]), float(u[tuple(b)])
        t = p[ax] - a[ax]
        assert abs(f1 + t * (f2 - f1) - iso) < 1e-12 and (f1 < iso) != (f2 < iso)
    assert f.min() >= 0 and f.max() < len(v)


def test_isovalue_equal_to_grid_values_and_empty_grids():
    u = np.zeros((4, 4, 4), np.float32)
    v, f = M.marching_cubes(u, 0.0)       # nothing is < iso: empty mesh
    assert len(v) == 0 and len(f) == 0
    u[1:3, 1:3, 1:3] = -1.0               # values equal to iso count as outside; vertices land exactly on grid points
    v, f = M.marching_cubes(u, 0.0)
    check_closed_manifold(v, f)
    assert np.all(v == np.round(v))
    v, f = M.marching_cubes(np.ones((1, 5, 5), np.float32), 0.5)
    assert len(v) == 0 and len(f) == 0


def test_library_case_table_matches_oracle_derivation():
    import __graft_entry__ as g
    g.build()
    from color_neus_b200.marching_cubes import case_tables
    n_tri, tri = case_tables()
    on, ot, _ = M.build_tables()
    assert np.array_equal(n_tri, on) and np.array_equal(tri[:, :15], ot) and np.all(tri[:, 15] == -1)
    assert int(n_tri.max()) == 5 and int(n_tri.sum()) == 820
    # complementary cases cut the same edges
    for c in range(256):
        assert set(tri[c][tri[c] >= 0].tolist()) == set(tri[255 - c][tri[255 - c] >= 0].tolist())


def test_ply_round_trip(tmp_path):
    from color_neus_b200.marching_cubes import read_ply, write_ply
    v, f = M.marching_cubes(sphere_field(12), 0.0)
    col = np.random.default_rng(0).random((len(v), 3)).astype(np.float32)
    p = os.path.join(tmp_path, "m.ply")
    write_ply(p, v, f, col)
    head = open(p, "rb").read(400).decode("latin1")
    for line in ("format binary_little_endian 1.0", f"element vertex {len(v)}", "property float x", "property uchar red",
                 "property uchar alpha", f"element face {len(f)}", "property list uchar int vertex_indices"):
        assert line in head
    v2, f2, rgba = read_ply(p)
    assert np.array_equal(v2, v.astype(np.float32)) and np.array_equal(f2, f.astype(np.int32))
    assert np.array_equal(rgba[:, :3], np.round(col * 255).astype(np.uint8)) and np.all(rgba[:, 3] == 255)
    write_ply(p, v, f)
    v3, f3, none = read_ply(p)
    assert none is None and np.array_equal(f3, f2)


def test_no_cpu_path():
    from color_neus_b200._lib import CneusError
    from color_neus_b200.marching_cubes import marching_cubes_device
    with pytest.raises(CneusError):
        marching_cubes_device(torch.zeros(4, 4, 4), 0.0)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["sphere", "torus", "noise", "noise_open", "ragged", "iso"])
def test_gpu_mesh_equals_oracle(case):
    from color_neus_b200.marching_cubes import marching_cubes
    iso = 0.0
    if case == "sphere":
        u = sphere_field(33)
    elif case == "torus":
        u = torus_field(40)
    elif case == "noise":
        u = noise_field((21, 19, 23), 3)
    elif case == "noise_open":
        u = noise_field((16, 16, 16), 4, closed=False)
    elif case == "ragged":
        u = noise_field((5, 70, 3), 6, closed=False)       # chunks straddle rows and planes
    else:
        u, iso = noise_field((12, 12, 12), 7), 0.3173
    v, f = marching_cubes(u, iso)
    vo, fo = M.marching_cubes(u, iso)
    assert v.dtype == np.float64 and f.dtype == np.int64
    assert np.array_equal(f, fo)            # bit-exact indices, same order
    assert np.array_equal(v, vo)            # bit-exact float64 vertices (same IEEE operations)


@pytest.mark.gpu
def test_gpu_degenerate_grids():
    from color_neus_b200.marching_cubes import marching_cubes
    v, f = marching_cubes(np.zeros((6, 6, 6), np.float32), 0.0)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    v, f = marching_cubes(np.ones((1, 5, 5), np.float32), 0.5)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    u = np.zeros((4, 4, 4), np.float32)
    u[1:3, 1:3, 1:3] = -1.0
    v, f = marching_cubes(u, 0.0)
    vo, fo = M.marching_cubes(u, 0.0)
    assert np.array_equal(v, vo) and np.array_equal(f, fo)


@pytest.mark.gpu
def test_gpu_large_grid_invariants():
    """256^3 sphere + torus union: closed oriented manifold, Euler characteristic 2 + 0, volume (size-independent properties)."""
    from color_neus_b200.marching_cubes import marching_cubes_device
    n = 256
    g = torch.linspace(-1, 1, n, device="cuda")
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    sph = 0.25 - torch.sqrt((X - 0.6) ** 2 + (Y - 0.6) ** 2 + (Z - 0.6) ** 2)
    tor = 0.12 - torch.sqrt((torch.sqrt(X ** 2 + Y ** 2) - 0.45) ** 2 + Z ** 2)
    u = torch.maximum(sph, tor).float().contiguous()
    v, f = marching_cubes_device(u, 0.0)
    v, f = v.cpu().numpy(), f.cpu().numpy().astype(np.int64)
    assert check_closed_manifold(v, f) == 2 + 0
    vw = M.index_to_bbox(v, n, [-1, -1, -1], [1, 1, 1])
    want = 4 / 3 * np.pi * 0.25 ** 3 + 2 * np.pi ** 2 * 0.45 * 0.12 ** 2
    assert abs(M.signed_volume(vw, f) - want) < 0.01 * want


@pytest.mark.gpu
def test_gpu_extract_geometry_and_color_end_to_end(tmp_path):
    """NeuS_Trainer.validate_mesh's sequence (NeuS_Trainer.py:279-307) through the drop-in renderer: extract_geometry ->
    extract_color -> PLY; the geometric-init SDF is a sphere of radius ~ BIAS / SCALE."""
    import __graft_entry__ as g
    import color_neus_b200 as cn
    from color_neus_b200.marching_cubes import read_ply, write_ply
    from oracle import neus_oracle as O
    torch.manual_seed(1)
    ren = cn.Color_NeuS(g._Cfg(O.default_cfg())).cuda().eval()
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    v, f = ren.extract_geometry(bmin, bmax, "cuda", resolution=64, threshold=0.0)
    assert check_closed_manifold(v, f) == 2
    r = np.linalg.norm(v, axis=1)
    assert 0.05 < r.min() and r.max() < 0.4
    # the same grid through the oracle's marching cubes gives the same mesh
    u = ren.extract_fields(bmin, bmax, 64).reshape(64, 64, 64).cpu().numpy()
    vo, fo = M.marching_cubes(u, 0.0)
    assert np.array_equal(f, fo) and np.allclose(v, M.index_to_bbox(vo, 64, bmin.numpy(), bmax.numpy()), rtol=0, atol=1e-12)
    col = ren.extract_color(v, "cuda")
    assert col.shape == (len(v), 3) and col.dtype == np.float32 and 0 <= col.min() and col.max() <= 1
    p = os.path.join(tmp_path, "00000000_color.ply")
    write_ply(p, v, f, col)
    v2, f2, rgba = read_ply(p)
    assert len(v2) == len(v) and np.array_equal(f2, f.astype(np.int32)) and rgba is not None
