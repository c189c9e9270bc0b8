"""`marching_cubes(u, isovalue)` with PyMCubes' call shape (`mcubes.marching_cubes`, the call at
lib/models/renderers/NeuS.py:35), running on the device over the grid produced by `cneus_sdf_grid` -- the 512 MB grid of a
512^3 extraction never leaves the GPU -- plus the PLY writer the trainer's `validate_mesh` needs (NeuS_Trainer.py:279-307,
there through trimesh).  Conventions and the parity statement: csrc/marching_cubes.cu, oracle/mc_oracle.py."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_bound = False


def _bind():
    global _bound
    lib = L.lib()
    if not _bound:
        vp, i32, i64, sz, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_double
        lib.cneus_mc_workspace_bytes.restype, lib.cneus_mc_workspace_bytes.argtypes = sz, [i32, i32, i32]
        lib.cneus_mc_count.restype, lib.cneus_mc_count.argtypes = C.c_int, [vp, i32, i32, i32, f64, vp, sz, vp, vp]
        lib.cneus_mc_emit.restype, lib.cneus_mc_emit.argtypes = C.c_int, [vp, i32, i32, i32, f64, vp, sz, i64, i64, vp, vp, vp]
        lib.cneus_mc_tables.restype, lib.cneus_mc_tables.argtypes = C.c_int, [vp, vp]
        _bound = True
    return lib


def case_tables():
    """(n_tri uint8[256], tri int8[256,16]) as derived by the library (host code; needs no GPU)."""
    lib = _bind()
    n_tri, tri = np.zeros(256, np.uint8), np.zeros((256, 16), np.int8)
    L.check(lib.cneus_mc_tables(n_tri.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p)), "cneus_mc_tables")
    return n_tri, tri


def marching_cubes_device(u, isovalue=0.0):
    """u: float32 CUDA tensor [nx, ny, nz] -> (vertices float64 CUDA [V,3] in index coordinates, triangles int32 CUDA [F,3])."""
    if not (torch.is_tensor(u) and u.is_cuda and u.dtype == torch.float32 and u.dim() == 3):
        raise L.CneusError("marching_cubes_device expects a float32 CUDA tensor [nx, ny, nz] (no CPU path)")
    u = u.contiguous()
    lib = _bind()
    nx, ny, nz = u.shape
    dev = u.device
    nbytes = lib.cneus_mc_workspace_bytes(nx, ny, nz)
    ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.cneus_mc_count(L.ptr(u), nx, ny, nz, float(isovalue), C.c_void_p(ws.data_ptr()), ws.numel() * 8,
                                   C.c_void_p(counts.data_ptr()), L.stream_ptr()), "cneus_mc_count")
        n_v, n_t = (int(x) for x in counts.tolist())   # the one host sync: the mesh size decides the allocation
        vertices = torch.empty(n_v, 3, dtype=torch.float64, device=dev)
        triangles = torch.empty(n_t, 3, dtype=torch.int32, device=dev)
        L.check(lib.cneus_mc_emit(L.ptr(u), nx, ny, nz, float(isovalue), C.c_void_p(ws.data_ptr()), ws.numel() * 8, n_v, n_t,
                                  C.c_void_p(vertices.data_ptr()), C.c_void_p(triangles.data_ptr()), L.stream_ptr()),
                "cneus_mc_emit")
    return vertices, triangles


def marching_cubes(u, isovalue=0.0):
    """mcubes.marching_cubes(u, isovalue) -> (vertices float64 ndarray [V,3], triangles int64 ndarray [F,3]).  `u` may be a
    numpy array (copied to the current CUDA device) or a CUDA tensor."""
    if not torch.is_tensor(u):
        if not torch.cuda.is_available():
            raise L.CneusError("marching_cubes needs a CUDA device (no CPU path)")
        u = torch.as_tensor(np.ascontiguousarray(u, dtype=np.float32)).cuda()
    v, t = marching_cubes_device(u.float(), isovalue)
    return v.cpu().numpy(), t.cpu().numpy().astype(np.int64)


# ---- PLY ------------------------------------------------------------------------------------------------------------------
def _to_rgba_u8(colors, n):
    c = np.asarray(colors)
    if c.shape[0] != n or c.ndim != 2 or c.shape[1] not in (3, 4):
        raise ValueError("vertex_colors must be [V,3] or [V,4]")
    if c.dtype.kind == "f":   # trimesh.visual.color.to_rgba: floats in [0,1] -> uint8
        c = np.round(np.clip(c, 0.0, 1.0) * 255.0).astype(np.uint8)
    else:
        c = c.astype(np.uint8)
    if c.shape[1] == 3:
        c = np.concatenate([c, np.full((n, 1), 255, np.uint8)], axis=1)
    return c


def write_ply(path, vertices, triangles, vertex_colors=None):
    """Binary little-endian PLY with the element / property layout trimesh's `mesh.export('*.ply')` writes
    (NeuS_Trainer.py:306-307): float x y z [+ uchar red green blue alpha] per vertex, `list uchar int vertex_indices` faces."""
    v = np.asarray(vertices, dtype="<f4").reshape(-1, 3)
    f = np.asarray(triangles).reshape(-1, 3).astype("<i4")
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    header = ["ply", "format binary_little_endian 1.0", "comment color_neus_b200", f"element vertex {len(v)}",
              "property float x", "property float y", "property float z"]
    if vertex_colors is not None:
        rgba = _to_rgba_u8(vertex_colors, len(v))
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1"), ("alpha", "u1")]
        header += ["property uchar red", "property uchar green", "property uchar blue", "property uchar alpha"]
    header += [f"element face {len(f)}", "property list uchar int vertex_indices", "end_header"]
    vrec = np.empty(len(v), dtype=fields)
    vrec["x"], vrec["y"], vrec["z"] = v[:, 0], v[:, 1], v[:, 2]
    if vertex_colors is not None:
        vrec["red"], vrec["green"], vrec["blue"], vrec["alpha"] = rgba[:, 0], rgba[:, 1], rgba[:, 2], rgba[:, 3]
    frec = np.empty(len(f), dtype=[("n", "u1"), ("i", "<i4", (3,))])
    frec["n"], frec["i"] = 3, f
    with open(path, "wb") as fh:
        fh.write(("\n".join(header) + "\n").encode("ascii"))
        fh.write(vrec.tobytes())
        fh.write(frec.tobytes())


def read_ply(path):
    """Reader for the files `write_ply` (and trimesh's binary export with the same properties) produce ->
    (vertices float32 [V,3], triangles int32 [F,3], rgba uint8 [V,4] or None)."""
    with open(path, "rb") as fh:
        data = fh.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").split("\n")
    if "format binary_little_endian 1.0" not in lines:
        raise ValueError("only binary little-endian PLY is supported")
    n_v = int(next(x for x in lines if x.startswith("element vertex")).split()[-1])
    n_f = int(next(x for x in lines if x.startswith("element face")).split()[-1])
    has_color = "property uchar red" in lines
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")] + ([("red", "u1"), ("green", "u1"), ("blue", "u1"), ("alpha", "u1")] if has_color else [])
    vdt = np.dtype(fields)
    vrec = np.frombuffer(data, dtype=vdt, count=n_v, offset=end)
    fdt = np.dtype([("n", "u1"), ("i", "<i4", (3,))])
    frec = np.frombuffer(data, dtype=fdt, count=n_f, offset=end + n_v * vdt.itemsize)
    v = np.stack([vrec["x"], vrec["y"], vrec["z"]], 1)
    rgba = np.stack([vrec["red"], vrec["green"], vrec["blue"], vrec["alpha"]], 1) if has_color else None
    return v, frec["i"].copy(), rgba
