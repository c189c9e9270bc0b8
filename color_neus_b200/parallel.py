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


def render_sharded(render_fn, rays_o, rays_d, near, far, gather=("color_fine", "depth"), chunk=None, **kw):
    """Render this rank's slice of the rays with `render_fn(rays_o, rays_d, near, far, **kw) -> dict` and
    (optionally) all-gather the listed per-ray outputs so every rank holds the full image.

    Returns (out, (begin, end)): `out[k]` is the full-length tensor for gathered keys, the local slice otherwise;
    `out['gradient_error']` is the batch-global Eikonal ratio rebuilt from the ranks' partial sums when the renderer
    exports them (`eikonal_num`, `eikonal_den`)."""
    rank, ws = world()
    n = rays_o.shape[0]
    b, e = shard_range(n, rank, ws)
    pieces, counts = [], []
    step = chunk or max(e - b, 1)
    for s in range(b, e, step):
        t = min(e, s + step)
        pieces.append(render_fn(rays_o[s:t], rays_d[s:t], near[s:t], far[s:t], **kw))
        counts.append(t - s)
    out = {}
    if pieces:
        for k in pieces[0]:
            vals = [p[k] for p in pieces]
            per_ray = all(torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == c for v, c in zip(vals, counts))
            out[k] = torch.cat(vals, dim=0) if per_ray else vals[-1]
    num = sum(float(p["eikonal_num"]) for p in pieces if "eikonal_num" in p)
    den = sum(float(p["eikonal_den"]) for p in pieces if "eikonal_den" in p)
    if ws > 1:
        dev = rays_o.device
        s = torch.tensor([num, den], dtype=torch.float64, device=dev)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)  # every rank takes part, also those with an empty shard
        num, den = float(s[0]), float(s[1])
        for k in gather or ():
            local = out.get(k)
            if local is None:  # empty shard
                ref_shape = (0, 3) if k in ("color_fine", "global_color") else (0,)
                local = torch.zeros(ref_shape, dtype=torch.float32, device=dev)
            out[k] = _all_gather_rows(local, n, ws)
    if den > 0 or num > 0:
        out["gradient_error"] = torch.tensor(num / (den + 1e-5), dtype=torch.float32, device=rays_o.device)
    return out, (b, e)


def allreduce_grads_and_losses(params, partial_sums=None):
    """One all-reduce(sum) over [all parameter gradients | loss partial sums]; gradients are written back in place.

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
