"""Small workload for compute-sanitizer (tools/run_sanitizers.sh): every kernel family of the hot path once, on batches that
leave the last 128-point tile partly empty -- stand-alone field calls, sampling (up-sample / merge), the fused render launch
of the tcgen05 kernel (in-place A operand, register hand-over between warpgroups, mbarrier protocol), compositing, the
training backward (fused recompute with dumps + tcgen05 GEMMs), loss / clip / Adam, grid query + marching cubes.
Results are checked against the CPU oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from oracle import neus_oracle as O  # noqa: E402


def main():
    g.build()
    import color_neus_b200 as cn
    from color_neus_b200 import train_ops as TR
    n_rays = int(os.environ.get("SAN_RAYS", "5"))
    cfg = O.default_cfg("Color_NeuS", 64, 64, 256, 8, 0.45)
    Pn = O.make_params(cfg, seed=3, trained_like=True)
    ren = cn.Color_NeuS(g._Cfg(cfg))
    ren.load_state_dict({k: torch.as_tensor(v).reshape(ren.state_dict()[k].shape) for k, v in Pn.items()}, strict=True)
    ren = ren.cuda().eval()
    c2w = O.pose_spherical(30.0, -30.0, 2.8)
    ro, rd = O.get_rays_at(c2w, torch.tensor([6.0 * 3, 6.0 * 3]), 3, 3)
    ro, rd = ro[:n_rays].contiguous(), rd[:n_rays].contiguous()
    near, far = O.near_far_from_sphere(ro, rd)
    t_rand = torch.rand([n_rays, 1], generator=torch.Generator().manual_seed(7))
    ref = O.render_forward(O.to_torch(Pn), cfg, ro, rd, near, far, t_rand=t_rand)
    with torch.no_grad():
        got = ren._forward_impl(ro.cuda(), rd.cuda(), near.cuda(), far.cuda(), t_rand=t_rand)
        pts = (ro[:, None, :] + rd[:, None, :] * torch.linspace(2.0, 3.5, 37)[None, :, None]).reshape(-1, 3).cuda()   # 185 points
        y = ren.sdf_network(pts)
        n = ren.sdf_network.gradient(pts).squeeze(1)
        cgl = ren.color_network(pts, n, None, y[:, 1:].contiguous())
        ren.relight_network(cgl, pts, torch.nn.functional.normalize(pts, dim=-1), n)
        u = ren.extract_fields(torch.tensor([-0.3] * 3), torch.tensor([0.3] * 3), 12)
        v, f = ren.extract_geometry(torch.tensor([-0.3] * 3), torch.tensor([0.3] * 3), "cuda", 12)
        ren.extract_color(v[:50])
    for k in ("color_fine", "weight_sum", "depth"):
        a, b = got[k].float().cpu().numpy().reshape(-1), ref[k].numpy().reshape(-1)
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
        print(f"[sanitize] forward {k}: rel err {err:.2e}")
        assert err < 1e-4, (k, err)
    # one training step: forward (training dumps), loss, backward, clip + Adam
    ren.train()
    opt = TR.FusedClipAdam(ren.parameters(), lr=5e-4, betas=(0.9, 0.99))
    loss_fn = TR.NeusLoss({"LAMBDA_MASK": 0.1}, include_mask=True)
    r = ren(ro.cuda(), rd.cuda(), near.cuda(), far.cuda())
    r["rgb_map_gt"] = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(1)).cuda()
    r["mask"] = (r["weight_sum"].detach().squeeze(-1) > 0.5).float()
    loss, _ = loss_fn(r)
    loss.backward()
    TR.clip_gradient(opt, 1.0, 2)
    opt.step()
    torch.cuda.synchronize()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in ren.parameters())
    print(f"[sanitize] training step ok, loss {float(loss):.5f}, grid {tuple(u.shape)}, mesh {v.shape[0]} vertices")


if __name__ == "__main__":
    main()
