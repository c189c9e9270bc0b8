"""One forward pass of the C2 renderer over a few hundred rays (for `ncu -k regex:shade_tc ...` captures).
Launch order of shade_tc_kernel per forward: 4 SDF-only launches (coarse + 3 up-sampling rounds), then render_core."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, color_neus_b200 as cn
from color_neus_b200.rays import synthetic_camera_rays
torch.manual_seed(1)
ren = cn.Color_NeuS(bench.renderer_cfg()).cuda().eval()
ro, rd, near, far = synthetic_camera_rays(800, 800, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8
sl = slice(300 * 800, 300 * 800 + n)
with torch.no_grad():
    out = ren(ro[sl], rd[sl], near[sl], far[sl])
torch.cuda.synchronize()
print("ok", float(out["color_fine"].sum()))
