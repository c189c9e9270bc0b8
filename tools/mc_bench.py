"""Timing of the device marching cubes on a 512^3 grid (BASELINE config C5's extraction size) next to the numpy oracle on
128^3 (secondary measurement; prints one JSON line)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def field(n, device):
    ax = torch.linspace(-1, 1, n, device=device)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    u = 0.6 - torch.sqrt(X * X + Y * Y + Z * Z) + 0.05 * torch.sin(9 * X) * torch.sin(7 * Y) * torch.sin(8 * Z)
    return u.float().contiguous()


if __name__ == "__main__":
    g.build()
    from color_neus_b200.marching_cubes import marching_cubes_device
    from oracle import mc_oracle as M
    n = 512
    u = field(n, "cuda")
    marching_cubes_device(u, 0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        v, f = marching_cubes_device(u, 0.0)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    uc = field(128, "cpu").numpy()
    t0 = time.perf_counter()
    vo, fo = M.marching_cubes(uc, 0.0)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    grid_bytes = n ** 3 * 4
    print(json.dumps({"metric": "marching cubes 512^3 (count + emit, incl. the size read-back)", "ms": ms, "vertices": int(v.shape[0]),
                      "triangles": int(f.shape[0]), "grid_GBps_equiv": grid_bytes / ms / 1e6,
                      "oracle_numpy_128^3_ms": cpu_ms, "oracle_vertices": int(len(vo))}))
