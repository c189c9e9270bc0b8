"""N>1 host logic on CPU: world_size-2 gloo processes exercise ray sharding, the re-derivation of the batch-global
Eikonal ratio from partial sums, and the single flat all-reduce of gradients + loss partial sums."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401  (puts the repo on sys.path)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(rays_o, rays_d, near, far, **kw):
    """Stand-in with the renderer's contract: per-ray outputs depend only on that ray; exports Eikonal partial sums."""
    n = rays_o.shape[0]
    color = torch.sin(rays_o * 3.0 + rays_d) * 0.5 + 0.5
    depth = (near + far) * 0.5 + rays_d[:, 0]
    e = (rays_d.norm(dim=1) - 1.0) ** 2
    relax = (rays_o.norm(dim=1) < 1.2).float()
    return {"color_fine": color, "depth": depth, "weights": torch.ones(n, 4), "eikonal_num": (relax * e).sum(),
            "eikonal_den": relax.sum(), "gradient_error": (relax * e).sum() / (relax.sum() + 1e-5)}


def _chk(ok, fails, tag, cond):
    if not cond:
        fails.append(tag)
    return ok and bool(cond)


def _worker(rank, world, port, n_rays, ret):
    fails = []
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from color_neus_b200 import parallel as par
        g = torch.Generator().manual_seed(3)
        ro, rd = torch.randn(n_rays, 3, generator=g), torch.randn(n_rays, 3, generator=g)
        near, far = torch.rand(n_rays, generator=g), torch.rand(n_rays, generator=g) + 1
        full = _fake_render(ro, rd, near, far)
        out, (b, e) = par.render_sharded(_fake_render, ro, rd, near, far, chunk=5)
        ok = _chk(True, fails, 41, torch.equal(out["color_fine"], full["color_fine"]) and torch.equal(out["depth"], full["depth"]))
        ok = _chk(ok, fails, 42, (("weights" not in out) if e == b else out["weights"].shape[0] == e - b))  # un-gathered keys stay local
        ok = _chk(ok, fails, 43, abs(float(out["gradient_error"]) - float(full["gradient_error"])) < 1e-6)
        bs, es = zip(*[par.shard_range(n_rays, r, world) for r in range(world)])
        ok = _chk(ok, fails, 45, bs[0] == 0 and es[-1] == n_rays and all(es[i] == bs[i + 1] for i in range(world - 1)))
        # gradients + loss partial sums in ONE all-reduce
        p1, p2 = torch.nn.Parameter(torch.zeros(7, 3)), torch.nn.Parameter(torch.zeros(5))
        p1.grad, p2.grad = torch.full((7, 3), float(rank + 1)), torch.arange(5.0) * (rank + 1)
        red = par.allreduce_grads_and_losses([p1, p2], torch.tensor([1.0 + rank, 10.0]))
        tot = sum(r + 1 for r in range(world))
        ok = _chk(ok, fails, 51, torch.equal(p1.grad, torch.full((7, 3), float(tot))) and torch.equal(p2.grad, torch.arange(5.0) * tot))
        ok = _chk(ok, fails, 52, torch.allclose(red, torch.tensor([sum(1.0 + r for r in range(world)), 10.0 * world])))
        # keep= drops the other keys per chunk; per_ray_kw tensors are sliced like the rays
        t_extra = torch.arange(float(n_rays))
        out2, _ = par.render_sharded(lambda o, d, n_, f, shift: dict(_fake_render(o, d, n_, f), depth=shift), ro, rd, near, far,
                                     chunk=4, keep=("color_fine", "depth"), per_ray_kw={"shift": t_extra})
        ok = _chk(ok, fails, 57, "weights" not in out2 and torch.equal(out2["depth"], t_extra) and torch.equal(out2["color_fine"], full["color_fine"]))
        ok = _chk(ok, fails, 58, abs(float(out2["gradient_error"]) - float(full["gradient_error"])) < 1e-6)
        # persistent flat gradient buffer: autograd accumulates into views of it, one all-reduce, extras ride along
        q1, q2 = torch.nn.Parameter(torch.ones(4, 2)), torch.nn.Parameter(torch.ones(3))
        buf = par.FlatGradBuffer([q1, q2], n_extra=2)
        for _ in range(2):   # second round: zero() keeps the views
            buf.zero()
            ((q1 * (rank + 1.0)).sum() + (q2 * torch.arange(3.0)).sum() * (rank + 2.0)).backward()
            buf.extra.copy_(torch.tensor([1.0 + rank, 5.0]))
            red = buf.all_reduce()
            ok = _chk(ok, fails, 67, q1.grad.untyped_storage().data_ptr() == buf.flat.untyped_storage().data_ptr())
            ok = _chk(ok, fails, 68, torch.equal(q1.grad, torch.full((4, 2), float(tot))))
            ok = _chk(ok, fails, 69, torch.equal(q2.grad, torch.arange(3.0) * sum(r + 2.0 for r in range(world))))
            ok = _chk(ok, fails, 70, torch.allclose(red, torch.tensor([sum(1.0 + r for r in range(world)), 5.0 * world])))
        # union-batch loss: the ranks' shares add up to the single-process loss, their gradients to its gradient
        if n_rays >= 2:
            w = torch.nn.Parameter(torch.tensor([0.3, -0.2, 0.5]))
            gt_c = torch.rand(n_rays, 3, generator=g)

            def diff_render(o, d):
                e_ = ((d * w).norm(dim=1) - 1.0) ** 2
                relax = (o.norm(dim=1) < 1.5).float()
                num, den = (relax * e_).sum(), relax.sum()
                return {"color_fine": torch.sigmoid(o * w), "weight_sum": torch.sigmoid((d * w).sum(1, keepdim=True)),
                        "delta_relight": (o * w)[:, None, :].expand(-1, 4, -1) * 0.1, "gradient_error": num / (den + 1e-5),
                        "eikonal_num": num, "eikonal_den": den}

            def ref_loss(r, gt_, m):   # NeuS_Trainer.compute_loss, MSE flavour, default lambdas of the shipped configs
                pw = r["weight_sum"].squeeze(-1).clip(1e-3, 1 - 1e-3)
                return (torch.nn.functional.mse_loss(r["color_fine"], gt_) + 0.1 * r["gradient_error"]
                        + 0.1 * torch.nn.functional.binary_cross_entropy(pw, m)
                        + torch.mean(r["delta_relight"] * m[:, None, None]) ** 2)
            r_full = diff_render(ro, rd)
            m_full = (r_full["weight_sum"].detach().squeeze(-1) > 0.5).float()
            l_full = ref_loss(r_full, gt_c, m_full)
            g_full, = torch.autograd.grad(l_full, w)
            r_loc = diff_render(ro[b:e], rd[b:e])
            sums = par.loss_partial_sums(r_loc, m_full[b:e])
            dist.all_reduce(sums)
            l_loc = par.union_batch_loss(r_loc, gt_c[b:e], m_full[b:e], n_rays, sums)
            gb = par.FlatGradBuffer([w], n_extra=1)
            l_loc.backward()
            gb.extra.copy_(l_loc.detach().reshape(1))
            l_tot = gb.all_reduce()
            ok = _chk(ok, fails, 101, abs(float(l_tot[0]) - float(l_full)) < 1e-6 * max(1.0, abs(float(l_full))))
            ok = _chk(ok, fails, 102, torch.allclose(w.grad, g_full, rtol=1e-5, atol=1e-7))
        # BASELINE C5: grid slabs and vertex ranges (x-major flattening, contiguous shards, one all-gather each)
        res = 5
        ax = torch.linspace(-1, 1, res)
        full_grid = (ax[:, None, None] * 100 + ax[None, :, None] * 10 + ax[None, None, :]).reshape(-1)
        grid = par.extract_fields_sharded(lambda b_, e_: full_grid[b_:e_].clone(), res)
        ok = _chk(ok, fails, 108, torch.equal(grid, full_grid))
        verts = torch.randn(n_rays, 3, generator=g)
        col = par.extract_color_sharded(lambda v: v * 2.0 + 1.0, verts)   # exactly rounded: slice == full bit for bit
        ok = _chk(ok, fails, 111, torch.equal(col, verts * 2.0 + 1.0))
        ret[rank] = bool(ok) if not fails else f"failed checks at lines {fails}"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [37, 2, 1])
def test_ray_sharding_world2(n_rays):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_rays, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_range_single_process():
    from color_neus_b200 import parallel as par
    assert par.shard_range(10, 0, 1) == (0, 10)
    assert [par.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert par.shard_range(2, 3, 4) == (2, 2)
