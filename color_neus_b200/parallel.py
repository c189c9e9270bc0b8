"""Ray-sharded multi-GPU execution (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The reference is single-GPU (`train.py:111` asserts n_gpus == 1).  Rays are independent, so the path shards with no
data-path collective: rank g renders the contiguous index range [g*ceil(N/G), (g+1)*ceil(N/G)) of the flattened
y*W+x ray order and owns that slice of the image.  The only batch-global quantities are the Eikonal ratio
sum(relax*e)/(sum(relax)+1e-5) (NeuS.py:275-277) and the relight mean (NeuS_Trainer.py:153): they are re-derived
from all-reduced partial sums so that the sharded loss equals the single-GPU loss on the union batch.
Training needs exactly one all-reduce per step over a flat buffer = all parameter gradients + the loss partial sums.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank=None, world_size=None):
    """Contiguous slice of n items owned by `rank`: [rank*ceil(n/G), min(n, (rank+1)*ceil(n/G)))."""
    if rank is None or world_size is None:
        rank, world_size = world()
    per = -(-n // world_size)
    b = min(n, rank * per)
    return b, min(n, b + per)


def _all_gather_rows(t, n_total, world_size):
    """all_gather of row-sharded tensors produced with shard_range (last shards may be short or empty)."""
    per = -(-n_total // world_size)
    pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(bufs, pad)
    return torch.cat(bufs, dim=0)[:n_total]


def render_sharded(render_fn, rays_o, rays_d, near, far, gather=("color_fine", "depth"), chunk=None, keep=None,
                   per_ray_kw=None, **kw):
    """Render this rank's slice of the rays with `render_fn(rays_o, rays_d, near, far, **kw) -> dict` and
    (optionally) all-gather the listed per-ray outputs so every rank holds the full image.

    `keep`: keys to retain from every chunk's dict (None = all; a full-image render only needs colour + depth, the
    per-sample outputs of every chunk would be hundreds of MB).  `per_ray_kw`: dict of per-ray tensors (e.g. injected
    `t_rand`) sliced like the rays and passed to `render_fn` as keyword arguments.
    Returns (out, (begin, end)): `out[k]` is the full-length tensor for gathered keys, the local slice otherwise;
    `out['gradient_error']` is the batch-global Eikonal ratio rebuilt from the ranks' partial sums when the renderer
    exports them (`eikonal_num`, `eikonal_den`).  No host synchronisation: the partial sums stay on the device."""
    rank, ws = world()
    n = rays_o.shape[0]
    dev = rays_o.device
    b, e = shard_range(n, rank, ws)
    pieces, counts = [], []
    step = chunk or max(e - b, 1)
    sums = torch.zeros(2, dtype=torch.float64, device=dev)   # [eikonal numerator, denominator] of this rank's rays
    have_sums = False
    for s in range(b, e, step):
        t = min(e, s + step)
        extra = {k: v[s:t] for k, v in (per_ray_kw or {}).items()}
        r = render_fn(rays_o[s:t], rays_d[s:t], near[s:t], far[s:t], **extra, **kw)
        if "eikonal_num" in r and "eikonal_den" in r:
            sums += torch.stack([r["eikonal_num"].detach().reshape(()), r["eikonal_den"].detach().reshape(())]).double()
            have_sums = True
        pieces.append(r if keep is None else {k: r[k] for k in keep if k in r})
        counts.append(t - s)
    out = {}
    if pieces:
        for k in pieces[0]:
            vals = [p[k] for p in pieces]
            per_ray = all(torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == c for v, c in zip(vals, counts))
            out[k] = (vals[0] if len(vals) == 1 else torch.cat(vals, dim=0)) if per_ray else vals[-1]
    if ws > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)  # every rank takes part, also those with an empty shard
        for k in gather or ():
            local = out.get(k)
            if local is None:  # empty shard
                ref_shape = (0, 3) if k in ("color_fine", "global_color") else (0,)
                local = torch.zeros(ref_shape, dtype=torch.float32, device=dev)
            out[k] = _all_gather_rows(local, n, ws)
    if have_sums or ws > 1:
        # sum(relax e) / (sum(relax) + 1e-5) over the union batch (NeuS.py:275-277); 0 / 1e-5 = 0 when nothing was rendered
        out["gradient_error"] = (sums[0] / (sums[1] + 1e-5)).to(torch.float32)
        out["eikonal_num"], out["eikonal_den"] = sums[0].to(torch.float32), sums[1].to(torch.float32)
    return out, (b, e)


class FlatGradBuffer:
    """One persistent flat fp32 buffer holding every parameter gradient plus `n_extra` trailing scalars (loss partial
    sums).  Each `p.grad` is a view into it, so autograd accumulates straight into the buffer, ONE `all_reduce` joins the
    ranks' gradients and the extras (no per-step `torch.cat` / copy-back), and `FusedClipAdam` reads the reduced gradients
    in place.  Use `zero()` instead of `optimizer.zero_grad(set_to_none=True)` (which would drop the views)."""

    def __init__(self, params, n_extra=0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBuffer: no parameter requires grad")
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.n_grad = sum(sizes)
        self.flat = torch.zeros(self.n_grad + int(n_extra), dtype=torch.float32, device=dev)
        off = 0
        for p, k in zip(self.params, sizes):
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatGradBuffer needs float32 parameters on one device")
            p.grad = self.flat[off:off + k].view(p.shape)
            off += k
        self.extra = self.flat[self.n_grad:]

    def zero(self):
        self.flat.zero_()
        for p in self.params:   # a zero_grad(set_to_none=True) in between would have dropped the views: restore them
            if p.grad is None or p.grad.untyped_storage().data_ptr() != self.flat.untyped_storage().data_ptr():
                raise RuntimeError("FlatGradBuffer: a parameter's .grad no longer aliases the flat buffer "
                                   "(use buffer.zero() instead of zero_grad(set_to_none=True))")

    def all_reduce(self):
        """Sum gradients (+ extras) over the ranks in place; returns the reduced extras (a view)."""
        _, ws = world()
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.extra


def allreduce_grads_and_losses(params, partial_sums=None):
    """One all-reduce(sum) over [all parameter gradients | loss partial sums]; gradients are written back in place.
    Convenience form for parameters whose gradients are separate tensors (it has to gather and scatter them);
    `FlatGradBuffer` is the copy-free form the training loop uses.

    `params`: iterable of tensors with `.grad`; `partial_sums`: 1-D tensor of per-rank partial sums (numerators /
    denominators / counts) or None.  Returns the reduced partial sums (or None)."""
    rank, ws = world()
    params = [p for p in params if p.grad is not None]
    if not params and partial_sums is None:
        return None
    dev = params[0].grad.device if params else partial_sums.device
    parts = [p.grad.reshape(-1).to(torch.float32) for p in params]
    n_extra = 0
    if partial_sums is not None:
        parts.append(partial_sums.reshape(-1).to(device=dev, dtype=torch.float32))
        n_extra = parts[-1].numel()
    flat = torch.cat(parts) if parts else torch.zeros(0, device=dev)
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for p in params:
        k = p.grad.numel()
        p.grad.copy_(flat[off:off + k].reshape(p.grad.shape))
        off += k
    return flat[off:off + n_extra] if n_extra else None


def union_batch_loss(r, rgb_gt, mask, n_total, sums, lambda_eikonal=0.1, lambda_mask=0.1, lambda_relight=1.0):
    """This rank's additive share of NeuS_Trainer.compute_loss (NeuS_Trainer.py:129-171, MSE colour loss) of the UNION
    batch of `n_total` rays, from the render dict `r` of the rank's slice.  The shares of all ranks add up to the
    single-GPU loss and their gradients to its gradient, so that one all-reduce(sum) of the parameter gradients finishes
    the step.  The two batch-global terms need the all-reduced partial sums `sums` = [sum(relax e), sum(relax),
    sum(delta_relight * mask)] (3 scalars, one tiny all-reduce before backward):
      * Eikonal ratio: the local ratio re-weighted by the constant denominators;
      * (mean delta_relight)^2: linearised around the global sum S -- d/dx (S/n)^2 = 2 (S/n) / n -- with the constant
        chosen so that the shares add up to (S/n)^2."""
    _, ws = world()
    mse = ((r["color_fine"] - rgb_gt) ** 2).sum() / (n_total * 3)
    loss = mse + lambda_eikonal * r["gradient_error"] * (r["eikonal_den"].detach() + 1e-5) / (sums[1] + 1e-5)
    if lambda_mask:
        p = r["weight_sum"].squeeze(-1).clip(1e-3, 1 - 1e-3)
        loss = loss + lambda_mask * (-(mask * torch.log(p) + (1 - mask) * torch.log(1 - p)).sum() / n_total)
    if "delta_relight" in r and lambda_relight:
        n_rel = n_total * r["delta_relight"].shape[1] * 3
        rel_local = (r["delta_relight"] * mask[:, None, None]).sum()
        mean_g = sums[2] / n_rel
        loss = loss + lambda_relight * (2.0 * mean_g * rel_local / n_rel - mean_g ** 2 / ws)
    return loss


def loss_partial_sums(r, mask):
    """[sum(relax e), sum(relax), sum(delta_relight * mask)] of this rank's slice (device tensor, no sync)."""
    rel = (r["delta_relight"].detach() * mask[:, None, None]).sum() if "delta_relight" in r else r["eikonal_num"].new_zeros(())
    return torch.stack([r["eikonal_num"].detach().reshape(()), r["eikonal_den"].detach().reshape(()), rel.reshape(())])


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config C5: SDF-grid extraction + per-vertex colour, sharded over ranks
# ---------------------------------------------------------------------------------------------------------------
def extract_fields_sharded(fields_fn, resolution):
    """Every rank evaluates the contiguous slab [b, e) of the flattened (x-major, NeuS.py:14-28) grid with
    `fields_fn(lin_begin, lin_end) -> float32 [e-b]` (e.g. `lambda b, e: renderer.extract_fields(bmin, bmax, res, b, e)`),
    then one all-gather leaves the full res^3 grid on every rank (512^3 fp32 = 512 MB over NVLink: a few ms)."""
    rank, ws = world()
    total = int(resolution) ** 3
    b, e = shard_range(total, rank, ws)
    local = fields_fn(b, e)
    if ws == 1:
        return local
    return _all_gather_rows(local.reshape(-1, 1), total, ws).reshape(-1)


def extract_color_sharded(color_fn, vertices):
    """Per-vertex colour (NeuS.py:44-64) of a contiguous vertex range per rank, all-gathered: `color_fn(v [n,3]) -> [n,3]`
    (torch tensors on the rank's device).  Vertices are independent, so no other exchange is needed."""
    rank, ws = world()
    n = vertices.shape[0]
    b, e = shard_range(n, rank, ws)
    local = color_fn(vertices[b:e]) if e > b else torch.zeros(0, 3, dtype=torch.float32, device=vertices.device)
    if ws == 1:
        return local
    return _all_gather_rows(local, n, ws)
