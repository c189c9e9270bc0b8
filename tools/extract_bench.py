"""BASELINE config C5 on one GPU (secondary measurement; prints one JSON line): 512^3 SDF-grid query (cneus_sdf_grid,
SURVEY 8a a13), marching cubes on the device (8f #3), per-vertex colour (cneus_vertex_color, a14).  Geometric-init network
(a sphere); the reference does the same work in 512 blocks of 64^3 with a D2H copy each, marching cubes on one CPU core and
64 vertices per colour call (NeuS.py:14-64)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402

if __name__ == "__main__":
    g.build()
    import color_neus_b200 as cn
    from color_neus_b200.marching_cubes import marching_cubes_device
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    torch.manual_seed(1)
    ren = cn.Color_NeuS(bench.renderer_cfg()).cuda().eval()
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, (time.perf_counter() - t0) * 1e3

    ren.extract_fields(bmin, bmax, 32)   # warm-up (packing, workspace)
    marching_cubes_device(ren.extract_fields(bmin, bmax, 32).reshape(32, 32, 32), 0.0)   # warm-up (table derivation, module load)
    u, ms_grid = timed(lambda: ren.extract_fields(bmin, bmax, res))
    (v, f), ms_mc = timed(lambda: marching_cubes_device(u.reshape(res, res, res), 0.0))
    vw = (v / (res - 1.0) * (bmax - bmin).cuda().double() + bmin.cuda().double()).float().contiguous()
    ren.extract_color(vw[:256].cpu().numpy())
    col, ms_col = timed(lambda: ren.extract_color(vw.cpu().numpy()))
    n = res ** 3
    print(json.dumps({"metric": f"C5 extraction at {res}^3 on one B200", "grid_ms": ms_grid, "grid_points_per_s": n / ms_grid * 1e3,
                      "grid_algorithmic_tflops": n * 2 * bench.MAC_SDF_ONLY / ms_grid / 1e9, "marching_cubes_ms": ms_mc,
                      "vertices": int(v.shape[0]), "triangles": int(f.shape[0]), "vertex_color_ms_incl_host_copies": ms_col,
                      "vertices_per_s": v.shape[0] / ms_col * 1e3}))
