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


def _worker(rank, world, port, n_rays, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from color_neus_b200 import parallel as par
        g = torch.Generator().manual_seed(3)
        ro, rd = torch.randn(n_rays, 3, generator=g), torch.randn(n_rays, 3, generator=g)
        near, far = torch.rand(n_rays, generator=g), torch.rand(n_rays, generator=g) + 1
        full = _fake_render(ro, rd, near, far)
        out, (b, e) = par.render_sharded(_fake_render, ro, rd, near, far, chunk=5)
        ok = torch.equal(out["color_fine"], full["color_fine"]) and torch.equal(out["depth"], full["depth"])
        ok = ok and (("weights" not in out) if e == b else out["weights"].shape[0] == e - b)  # un-gathered keys stay local
        ok = ok and abs(float(out["gradient_error"]) - float(full["gradient_error"])) < 1e-6
        bs, es = zip(*[par.shard_range(n_rays, r, world) for r in range(world)])
        ok = ok and bs[0] == 0 and es[-1] == n_rays and all(es[i] == bs[i + 1] for i in range(world - 1))
        # gradients + loss partial sums in ONE all-reduce
        p1, p2 = torch.nn.Parameter(torch.zeros(7, 3)), torch.nn.Parameter(torch.zeros(5))
        p1.grad, p2.grad = torch.full((7, 3), float(rank + 1)), torch.arange(5.0) * (rank + 1)
        red = par.allreduce_grads_and_losses([p1, p2], torch.tensor([1.0 + rank, 10.0]))
        tot = sum(r + 1 for r in range(world))
        ok = ok and torch.equal(p1.grad, torch.full((7, 3), float(tot))) and torch.equal(p2.grad, torch.arange(5.0) * tot)
        ok = ok and torch.allclose(red, torch.tensor([sum(1.0 + r for r in range(world)), 10.0 * world]))
        # BASELINE C5: grid slabs and vertex ranges (x-major flattening, contiguous shards, one all-gather each)
        res = 5
        ax = torch.linspace(-1, 1, res)
        full_grid = (ax[:, None, None] * 100 + ax[None, :, None] * 10 + ax[None, None, :]).reshape(-1)
        grid = par.extract_fields_sharded(lambda b_, e_: full_grid[b_:e_].clone(), res)
        ok = ok and torch.equal(grid, full_grid)
        verts = torch.randn(n_rays, 3, generator=g)
        col = par.extract_color_sharded(lambda v: torch.sigmoid(v * 2.0), verts)
        ok = ok and torch.equal(col, torch.sigmoid(verts * 2.0))
        ret[rank] = bool(ok)
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
