"""Multi-GPU training step over NCCL: `torchrun --nproc-per-node 2 tools/ddp_train_check.py`.

Every rank renders its contiguous slice of one 1024-ray batch through the drop-in Color_NeuS (same weights), forms the loss
of the UNION batch from all-reduced partial sums (the Eikonal ratio and the relight mean are batch-global: 3 scalars),
back-propagates, and joins the parameter gradients with ONE all-reduce of the flat gradient buffer
(`parallel.FlatGradBuffer.all_reduce`; loss shares from `parallel.union_batch_loss`); then clip + Adam (`FusedClipAdam`) run identically on every rank.  Rank 0 also
runs the whole batch alone and the two results are compared: loss, every parameter gradient, parameters after the step.
Prints one JSON line (rank 0) with the worst relative differences and the step time (max over ranks, CUDA events)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402


def main():
    dist.init_process_group("nccl")
    rank, ws = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    g.build()
    import color_neus_b200 as cn
    from color_neus_b200 import parallel as par
    from color_neus_b200 import train_ops as TR
    from color_neus_b200.rays import synthetic_camera_rays
    n_rays = 1024
    torch.manual_seed(1)
    ren = cn.Color_NeuS(bench.renderer_cfg()).cuda().train()
    ref = cn.Color_NeuS(bench.renderer_cfg()).cuda().train()
    ref.load_state_dict(ren.state_dict())
    ro, rd, near, far = synthetic_camera_rays(768, 576, device="cuda")
    gen = torch.Generator().manual_seed(3)
    idx = torch.randint(0, ro.shape[0], (n_rays,), generator=gen).cuda()
    ro, rd, near, far = ro[idx].contiguous(), rd[idx].contiguous(), near[idx].contiguous(), far[idx].contiguous()
    gt = torch.rand(n_rays, 3, generator=gen).cuda()
    b, e = par.shard_range(n_rays, rank, ws)

    buf = par.FlatGradBuffer(ren.parameters(), n_extra=1)   # p.grad = views of one flat buffer, reduced in place

    def sharded_step(opt):
        buf.zero()
        r = ren(ro[b:e], rd[b:e], near[b:e], far[b:e], perturb_overwrite=0)   # no jitter: the slices see the same samples as the union
        mask = (r["weight_sum"].detach().squeeze(-1) > 0.5).float()
        sums = par.loss_partial_sums(r, mask)
        dist.all_reduce(sums)
        loss = par.union_batch_loss(r, gt[b:e], mask, n_rays, sums)
        loss.backward()
        buf.extra.copy_(loss.detach().reshape(1))
        tot = float(buf.all_reduce()[0])
        TR.clip_gradient(opt, 1.0, 2)
        opt.step()
        return tot

    opt = TR.FusedClipAdam(ren.parameters(), lr=5e-4, betas=(0.9, 0.99))
    grads = {}
    loss_sharded = None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        if it == 2:
            dist.barrier(); torch.cuda.synchronize(); ev0.record()
        if it > 0:                                   # steps 1, 2 are timing repeats from the same start state
            ren.load_state_dict(ref.state_dict())
            opt = TR.FusedClipAdam(ren.parameters(), lr=5e-4, betas=(0.9, 0.99))
        loss_sharded = sharded_step(opt)
        if it == 2:
            ev1.record(); torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    grads = {k: p.grad.clone() for k, p in ren.named_parameters()}
    out = {"n_gpus": ws, "rays": n_rays, "ms_per_step_max_over_ranks": float(ms[0]), "loss_sharded": loss_sharded}
    if rank == 0:
        opt_ref = TR.FusedClipAdam(ref.parameters(), lr=5e-4, betas=(0.9, 0.99))
        opt_ref.zero_grad(set_to_none=True)
        r = ref(ro, rd, near, far, perturb_overwrite=0)
        mask = (r["weight_sum"].detach().squeeze(-1) > 0.5).float()
        loss = (torch.nn.functional.mse_loss(r["color_fine"], gt) + 0.1 * r["gradient_error"]
                + 0.1 * torch.nn.functional.binary_cross_entropy(r["weight_sum"].squeeze(-1).clip(1e-3, 1 - 1e-3), mask)
                + torch.mean(r["delta_relight"] * mask[:, None, None]) ** 2)
        loss.backward()
        worst_g = max(float((grads[k] - p.grad).abs().max() / (p.grad.abs().max() + 1e-12)) for k, p in ref.named_parameters())
        TR.clip_gradient(opt_ref, 1.0, 2)
        opt_ref.step()
        worst_p = max(float((dict(ren.named_parameters())[k] - p).abs().max() / (p.abs().max() + 1e-12)) for k, p in ref.named_parameters())
        out.update(loss_single=float(loss), loss_rel_diff=abs(loss_sharded - float(loss)) / abs(float(loss)),
                   worst_grad_rel_diff=worst_g, worst_param_rel_diff_after_step=worst_p)
        print(json.dumps(out), flush=True)
        # the first Adam step moves every element by ~lr * sign(g): elements whose tiny gradient changes sign between the
        # two summation orders differ by up to 2 lr, hence the looser bar on the parameters
        assert out["loss_rel_diff"] < 1e-5 and worst_g < 5e-3 and worst_p < 1e-3, out
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
