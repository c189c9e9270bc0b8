"""Training-step timing (BASELINE config C3 shape: N_RAYS=1024, 64+128 samples; also 64+64): forward + loss + backward
through the drop-in renderer on one GPU.  Prints one JSON line per configuration (secondary measurement; bench.py is
the contract benchmark)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402


def run(n_imp, n_rays=1024, steps=5, warmup=2, fused=True, chunk_rays=0, fused_recompute=1):
    import color_neus_b200 as cn
    from color_neus_b200 import _lib as L
    from color_neus_b200.rays import synthetic_camera_rays
    L.lib().cneus_backward_chunk_rays(chunk_rays)
    L.lib().cneus_backward_fused_recompute(fused_recompute)
    cfg = bench.renderer_cfg()
    cfg["N_IMPORTANCE"] = n_imp
    torch.manual_seed(1)
    ren = cn.Color_NeuS(cfg).cuda().train()
    ro, rd, near, far = synthetic_camera_rays(768, 576, device="cuda")
    gen = torch.Generator().manual_seed(3)
    idx = torch.randint(0, ro.shape[0], (n_rays,), generator=gen).cuda()
    ro, rd, near, far = ro[idx].contiguous(), rd[idx].contiguous(), near[idx].contiguous(), far[idx].contiguous()
    gt = torch.rand(n_rays, 3, generator=gen).cuda()
    from color_neus_b200 import train_ops as TR
    if fused:   # loss (2 launches) + clip + Adam (2 launches), no host sync
        opt = TR.FusedClipAdam(ren.parameters(), lr=5e-4, betas=(0.9, 0.99))
        loss_fn = TR.NeusLoss({"LAMBDA_MASK": 0.1}, include_mask=True)
    else:       # what the reference's train.py does: torch losses, per-tensor clip_grad_norm_, torch.optim.Adam
        opt = torch.optim.Adam(ren.parameters(), lr=5e-4, betas=(0.9, 0.99))

    def step():
        opt.zero_grad(set_to_none=True)
        r = ren(ro, rd, near, far)
        mask = (r["weight_sum"].detach().squeeze(-1) > 0.5).float()
        if fused:
            r["rgb_map_gt"], r["mask"] = gt, mask
            loss, _ = loss_fn(r)
            loss.backward()
            TR.clip_gradient(opt, 1.0, 2)
            opt.step()
            return loss
        loss = torch.nn.functional.mse_loss(r["color_fine"], gt) + 0.1 * r["gradient_error"]
        loss = loss + 0.1 * torch.nn.functional.binary_cross_entropy(r["weight_sum"].squeeze(-1).clip(1e-3, 1 - 1e-3), mask)
        loss = loss + torch.mean(r["delta_relight"] * mask[:, None, None]) ** 2
        loss.backward()
        for group in opt.param_groups:   # net_utils.py:174-184
            for p in group["params"]:
                torch.nn.utils.clip_grad_norm_(p, 1.0, 2)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
    S = cfg["N_SAMPLES"] + n_imp
    flop = 2 * ((64 + 3 * n_imp // 4) * bench.MAC_SDF_ONLY + 3 * S * (bench.MAC_SDF_FULL + bench.MAC_GRAD + bench.MAC_COLOR + bench.MAC_RELIGHT))
    print(json.dumps({"metric": "training step (fwd + loss + bwd + clip + Adam)", "fused_loss_clip_adam": fused, "backward_chunk_rays": chunk_rays, "backward_fused_recompute": fused_recompute, "n_rays": n_rays, "samples": f"64+{n_imp}",
                      "ms_per_step": ms, "rays_per_s": n_rays / ms * 1e3, "algorithmic_tflops": n_rays * flop / ms / 1e9,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


if __name__ == "__main__":
    g.build()
    if "--sweep" in sys.argv:
        for c in (0, 512, 256, 128):
            run(64, chunk_rays=c)
        for c in (0, 256, 128):
            run(128, chunk_rays=c)
    else:
        run(64, fused=False, fused_recompute=0)
        run(64, fused_recompute=0)
        run(64)
        run(64, fused_recompute=3)
        run(128, fused_recompute=0)
        run(128)
