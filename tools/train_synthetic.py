"""Optimise a Color_NeuS on an analytic synthetic scene with the repo's own training path (GPU only):

    python tools/train_synthetic.py [--steps 2000] [--out gpurun_out/optimised_state.npz]

Scene: the union of two spheres with a procedural albedo, one directional light (diffuse + a view-dependent specular lobe),
black background, exact masks -- rendered analytically per ray, so there is no dataset.  Training loop = train.py:63-77 with
this repo's drop-ins: `Color_NeuS.forward` (training mode: analytic backward kernels) -> `NeusLoss` (compute_loss) ->
`loss.backward()` -> `clip_gradient` -> `FusedClipAdam.step` -> `NeuS_lr_scheduler.step`.

Purpose (VERDICT r1, next #1b): produce a state that has REALLY been optimised (non-trivial SDF, sharpened variance,
trained colour / relight weights) for the parity fixtures -- tests/golden/make_golden.py turns the saved state_dict into a
golden case by running the UNMODIFIED reference on it -- and a loss / PSNR trace as evidence that the training path
converges (profiles/).  Prints one JSON line per logging interval and a final summary line."""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402

SPHERES = [((-0.15, 0.0, 0.0), 0.35), ((0.25, 0.1, 0.05), 0.25)]
LIGHT = (0.4, -0.5, -0.77)


def analytic_scene(ro, rd):
    """-> (rgb [n,3], mask [n]) of rays against the two-sphere scene (rd normalised)."""
    n = ro.shape[0]
    t_best = torch.full((n,), float("inf"), device=ro.device)
    nrm = torch.zeros(n, 3, device=ro.device)
    for c, r in SPHERES:
        c = torch.tensor(c, device=ro.device)
        oc = ro - c
        b = (oc * rd).sum(-1)
        disc = b * b - ((oc * oc).sum(-1) - r * r)
        t = -b - torch.sqrt(disc.clamp_min(0.0))
        hit = (disc > 0) & (t > 0) & (t < t_best)
        t_best = torch.where(hit, t, t_best)
        p = ro + rd * t[:, None]
        nrm = torch.where(hit[:, None], (p - c) / r, nrm)
    mask = torch.isfinite(t_best)
    p = ro + rd * torch.where(mask, t_best, torch.zeros_like(t_best))[:, None]
    light = torch.nn.functional.normalize(torch.tensor(LIGHT, device=ro.device), dim=0)
    albedo = 0.5 + 0.5 * torch.sin(7.0 * p + torch.tensor([0.0, 2.0, 4.0], device=ro.device))
    diff = (-(nrm * light).sum(-1)).clamp_min(0.0)
    refl = rd - 2.0 * (rd * nrm).sum(-1, keepdim=True) * nrm
    spec = (-(refl * light).sum(-1)).clamp_min(0.0) ** 16
    rgb = (albedo * (0.3 + 0.7 * diff[:, None]) + 0.3 * spec[:, None]).clamp(0.0, 1.0)
    return rgb * mask[:, None].float(), mask.float()


def random_rays(n_rays, gen, dev, side=256, focal_mul=2.5):
    from color_neus_b200.rays import get_rays_selected, pose_spherical
    theta = float(torch.rand((), generator=gen)) * 360.0
    phi = -10.0 - 50.0 * float(torch.rand((), generator=gen))
    c2w = pose_spherical(theta, phi, 2.8).to(dev)
    focal = torch.tensor([focal_mul * side, focal_mul * side], device=dev)
    idx = torch.randint(0, side * side, (n_rays,), generator=gen)
    # opengl=True: blender-style poses look down -z, so the rays point at the object and near / far are positive (with the
    # non-OpenGL convention of the DTU configs the same poses see the object "behind" the camera at negative depths)
    ro, rd, near, far, _, _ = get_rays_selected(c2w, focal, side, side, idx, normalize=True, opengl=True, with_near_far=True)
    return ro, rd, near, far


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--n-rays", type=int, default=1024)
    ap.add_argument("--n-importance", type=int, default=64)
    ap.add_argument("--log-every", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "optimised_state.npz"))
    args = ap.parse_args()
    g.build()
    import color_neus_b200 as cn
    from color_neus_b200 import train_ops as TR
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    ren = cn.Color_NeuS(bench.renderer_cfg(64, args.n_importance)).to(dev).train()
    opt_cfg = g._Cfg(dict(TYPE="adam", LR=5e-4, SCHEDULER_TYPE="NEUS", WARM_UP=100, LR_ALPHA=0.05))
    opt, sched = TR.build_optimizer_nerf(ren, opt_cfg, -1, iterations=args.steps)
    loss_fn = TR.NeusLoss({"LAMBDA_MASK": 0.1}, include_mask=True)
    gen = torch.Generator().manual_seed(5)
    trace = []
    t_ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    t_ev[0].record()
    for it in range(args.steps):
        ro, rd, near, far = random_rays(args.n_rays, gen, dev)
        gt, mask = analytic_scene(ro, rd)
        opt.zero_grad(set_to_none=True)
        r = ren(ro, rd, near, far)
        r["rgb_map_gt"], r["mask"] = gt, mask
        loss, terms = loss_fn(r)
        loss.backward()
        TR.clip_gradient(opt, 1.0, 2)
        opt.step()
        sched.step()
        if it % args.log_every == 0 or it == args.steps - 1:
            rec = {"step": it, "loss": float(loss), "rgb": float(terms["rgb_fine_loss"]), "eikonal": float(terms["eikonal_loss"]),
                   "mask": float(terms["mask_loss"]), "relight": float(terms["relight_loss"]), "psnr": float(loss_fn.psnr),
                   "variance": float(ren.deviation_network.variance), "lr": opt.param_groups[0]["lr"]}
            trace.append(rec)
            print(json.dumps(rec), flush=True)
    t_ev[1].record()
    torch.cuda.synchronize()
    # held-out view: PSNR + silhouette IoU of a full 128x128 image in eval mode
    from color_neus_b200.rays import get_rays_at, near_far_from_sphere, pose_spherical
    ren.eval()
    c2w = pose_spherical(77.0, -35.0, 2.8).to(dev)
    o, d = get_rays_at(c2w, torch.tensor([2.5 * 128, 2.5 * 128], device=dev), 128, 128, normalize=True, opengl=True)
    ro, rd = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
    near, far = near_far_from_sphere(ro, rd)
    with torch.no_grad():
        out = ren(ro, rd, near, far, perturb_overwrite=0)
    gt, mask = analytic_scene(ro, rd)
    mse = float(((out["color_fine"] - gt) ** 2).mean())
    sil = (out["weight_sum"].squeeze(-1) > 0.5).float()
    iou = float((sil * mask).sum() / ((sil + mask) > 0).float().sum().clamp_min(1))
    summary = {"summary": True, "steps": args.steps, "ms_per_step_incl_ray_generation": t_ev[0].elapsed_time(t_ev[1]) / args.steps,
               "heldout_hit_fraction": float(mask.mean()), "heldout_psnr": -10.0 * math.log10(max(mse, 1e-12)), "heldout_silhouette_iou": iou,
               "final_variance": float(ren.deviation_network.variance),
               "final_inv_s": float(torch.exp(ren.deviation_network.variance * 10.0)), "first": trace[0], "last": trace[-1]}
    print(json.dumps(summary), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.savez_compressed(args.out, **{k: v.detach().cpu().numpy() for k, v in ren.state_dict().items()})
    with open(os.path.splitext(args.out)[0] + "_trace.json", "w") as fh:
        json.dump({"trace": trace, "summary": summary}, fh, indent=1)


if __name__ == "__main__":
    main()
