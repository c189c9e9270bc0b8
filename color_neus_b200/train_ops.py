"""the per-step work around the renderer -- loss, gradient clipping, Adam, learning-rate
schedule -- as a handful of launches behind the reference's own names:

    compute_loss / NeusLoss      NeuS_Trainer.compute_loss            (lib/models/NeuS_Trainer.py:129-171)
    clip_gradient                net_utils.clip_gradient              (lib/utils/net_utils.py:174-184)
    FusedClipAdam                torch.optim.Adam as train.py uses it (net_utils.py:88, train.py:72-76)
    NeuS_lr_scheduler            net_utils.NeuS_lr_scheduler          (net_utils.py:56-80)
    build_optimizer_nerf         net_utils.build_optimizer_nerf       (net_utils.py:83-112)

The reference clips every parameter tensor on its own in a Python loop (53 tensors x ~5 launches + one host sync each
inside clip_grad_norm_), steps Adam, and calls `.item()` on the loss terms.  Here `clip_gradient` only records the
request; `FusedClipAdam.step` then runs norm + clip + Adam for all tensors in two launches (csrc/optim.cu) with no host
synchronisation; the loss and the seeds of its backward are two launches (csrc/loss.cu).  No CPU / PyTorch fallback:
CPU tensors or a missing libcneus.so raise CneusError.
"""
import ctypes as C
import math

import torch
from torch.optim.lr_scheduler import _LRScheduler

from . import _lib as L


class AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64)]


_bound = False


def _bind():
    global _bound
    lib = L.lib()
    if not _bound:
        vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
        lib.cneus_clip_adam_workspace_bytes.restype, lib.cneus_clip_adam_workspace_bytes.argtypes = sz, [i32]
        lib.cneus_clip_adam_step.restype = C.c_int
        lib.cneus_clip_adam_step.argtypes = [C.POINTER(AdamTensor), i32, f32, f32, f32, f32, f32, f32, i64, i32, vp, vp, sz, vp]
        lib.cneus_loss_workspace_bytes.restype, lib.cneus_loss_workspace_bytes.argtypes = sz, []
        lib.cneus_neus_loss.restype = C.c_int
        lib.cneus_neus_loss.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, f32, f32, f32, f32, i32, vp, vp, vp, vp, vp, sz, vp]
        _bound = True
    return lib


# ---------------------------------------------------------------------------------------------------------------------
# clip + Adam
# ---------------------------------------------------------------------------------------------------------------------
class FusedClipAdam(torch.optim.Optimizer):
    """torch.optim.Adam (amsgrad=False, maximize=False) whose step also applies the per-tensor gradient-norm clipping
    requested through `clip_gradient`.  `state_dict()` has torch.optim.Adam's layout (`step`, `exp_avg`, `exp_avg_sq`
    per parameter; lr / betas / eps / weight_decay per group), so the reference's checkpoints load either way."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")   # same checks as torch.optim.Adam
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                                      decoupled_weight_decay=False))
        self._pending_clip = None      # (max_norm, write_grad) set by clip_gradient, consumed by the next step()
        self.last_grad_norms = None    # device tensor of the un-clipped per-tensor norms of the last step (no sync)
        self._ws = {}

    def request_clip(self, max_norm, norm_type=2):
        if float(norm_type) != 2.0:
            raise L.CneusError(f"FusedClipAdam clips the L2 norm only (GRAD_CLIP.TYPE: 2), got norm_type={norm_type}")
        self._pending_clip = float(max_norm)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _bind()
        max_norm = self._pending_clip if self._pending_clip is not None else 0.0
        self._pending_clip = None
        norms_all = []
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize"):
                raise L.CneusError("FusedClipAdam: amsgrad / maximize are not implemented")
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            tab = (AdamTensor * len(ps))()
            step_no = None
            for i, p in enumerate(ps):
                g = p.grad
                if g.is_sparse:
                    raise L.CneusError("FusedClipAdam does not support sparse gradients")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()
                        and g.dtype == torch.float32 and p.device == dev):
                    raise L.CneusError("FusedClipAdam needs contiguous float32 CUDA parameters on one device (no CPU path)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                s = int(st["step"].item())   # CPU scalar: no device sync
                if step_no is None:
                    step_no = s
                elif s != step_no:
                    raise L.CneusError("FusedClipAdam: parameters of one group must share the step count")
                tab[i] = AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
            nbytes = lib.cneus_clip_adam_workspace_bytes(len(ps))
            key = (dev, len(ps))
            if key not in self._ws:
                self._ws[key] = (torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev),
                                 torch.empty(len(ps), dtype=torch.float32, device=dev))
            ws, norms = self._ws[key]
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                L.check(lib.cneus_clip_adam_step(tab, len(ps), max_norm, float(group["lr"]), float(b1), float(b2),
                                                 float(group["eps"]), float(group["weight_decay"]), step_no, 1,
                                                 C.c_void_p(norms.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel() * 8,
                                                 L.stream_ptr()), "cneus_clip_adam_step")
            norms_all.append(norms)
            # the kernel wrote the parameters (and the clipped gradients) through raw pointers: bump the autograd version
            # counters like an in-place torch op would, so that everything keyed on them (NetHandle's packed device images
            # of the weights, autograd's saved-tensor checks) sees the update
            torch._C._increment_version(ps)
            torch._C._increment_version([p.grad for p in ps])
        self.last_grad_norms = norms_all[0] if len(norms_all) == 1 else (torch.cat(norms_all) if norms_all else None)
        return loss


def clip_gradient(optimizer, max_norm, norm_type):
    """net_utils.clip_gradient (lib/utils/net_utils.py:174-184): clip every parameter tensor's gradient to `max_norm` on its
    own.  With a FusedClipAdam the clipping is folded into the following `optimizer.step()` (train.py:72-75 calls them back
    to back and zeroes the gradients right after), which also writes the clipped gradients back like clip_grad_norm_."""
    if not isinstance(optimizer, FusedClipAdam):
        raise L.CneusError("clip_gradient expects the FusedClipAdam built by build_optimizer_nerf (no unfused path here)")
    optimizer.request_clip(max_norm, norm_type)


class NeuS_lr_scheduler(_LRScheduler):
    """net_utils.NeuS_lr_scheduler (lib/utils/net_utils.py:56-80): linear warm-up over `warm_up` steps, then a cosine from
    1 to `alpha` at `end_iter`.  Host scalar arithmetic, same state_dict keys."""

    def __init__(self, optimizer, warm_up, alpha, end_iter, last_epoch=-1, verbose=False):
        self.warm_up, self.alpha, self.end_iter = warm_up, alpha, end_iter
        super().__init__(optimizer, last_epoch)

    def _factor(self):
        if self.last_epoch < self.warm_up:
            return self.last_epoch / self.warm_up
        progress = (self.last_epoch - self.warm_up) / (self.end_iter - self.warm_up)
        return (math.cos(math.pi * progress) + 1.0) * 0.5 * (1 - self.alpha) + self.alpha

    def get_lr(self):
        if self.last_epoch == 0:
            return [group["lr"] for group in self.optimizer.param_groups]
        f = self._factor()
        return [base_lr * f for base_lr in self.base_lrs]

    def _get_closed_form_lr(self):
        f = self._factor()
        return [base_lr * f for base_lr in self.base_lrs]


def build_optimizer_nerf(model, cfg, it, **kwargs):
    """net_utils.build_optimizer_nerf (lib/utils/net_utils.py:83-112) for TYPE 'adam' + SCHEDULER_TYPE 'NEUS' (every shipped
    NeuS / Color_NeuS config); other types belong to the reference's NeRF baselines and raise."""
    if cfg.TYPE != "adam":
        raise NotImplementedError(f"optimizer TYPE {cfg.TYPE!r}: only 'adam' is on the Color-NeuS path")
    optimizer = FusedClipAdam(model.parameters(), lr=cfg.LR, betas=(0.9, 0.99), eps=1e-8)
    lr = optimizer.param_groups[0]["lr"]
    if cfg.SCHEDULER_TYPE != "NEUS":
        raise ValueError(f"get unexcepted scheduler type: {cfg.SCHEDULER_TYPE}")
    if it != -1:
        for group in optimizer.param_groups:   # what torch requires of a resumed scheduler (checkpoints carry it)
            group.setdefault("initial_lr", group["lr"])
    scheduler = NeuS_lr_scheduler(optimizer, cfg.WARM_UP, cfg.LR_ALPHA, kwargs["iterations"], it)
    optimizer.param_groups[0]["lr"] = lr   # "ensure lr is not decreased again" (net_utils.py:110)
    return optimizer, scheduler


# ---------------------------------------------------------------------------------------------------------------------
# loss
# ---------------------------------------------------------------------------------------------------------------------
class _LossFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, color_fine, weight_sum, gradient_error, delta_relight, rgb_gt, mask, opts):
        lib = _bind()
        dev = color_fine.device
        if not color_fine.is_cuda:
            raise L.CneusError("compute_loss needs CUDA tensors (no CPU path)")
        f = lambda t: None if t is None else t.detach().to(dev, torch.float32).contiguous()   # noqa: E731
        color, gt, eik = f(color_fine), f(rgb_gt), f(gradient_error).reshape(1)
        B = color.shape[0]
        use_mask_term = opts["lambda_mask"] != 0
        ws = f(weight_sum).reshape(B) if weight_sum is not None else None
        m = f(mask).reshape(B) if mask is not None else None
        dl = f(delta_relight) if (delta_relight is not None and opts["lambda_relight"] != 0) else None
        if dl is not None and dl.dim() != 3:
            raise L.CneusError("delta_relight must be [B, S, 3]")
        S = dl.shape[1] if dl is not None else 0
        terms = torch.empty(5, device=dev)
        g_color = torch.empty_like(color)
        g_ws = torch.empty(B, device=dev) if use_mask_term else None
        g_dl = torch.empty_like(dl) if dl is not None else None
        wsb = torch.empty((lib.cneus_loss_workspace_bytes() + 7) // 8, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.cneus_neus_loss(L.ptr(color), L.ptr(gt), L.ptr(ws), L.ptr(m), L.ptr(eik), L.ptr(dl), B, S,
                                        1 if opts["rgb_l1"] else 0, opts["lambda_fine"], opts["lambda_eikonal"],
                                        opts["lambda_mask"], opts["lambda_relight"], 1 if opts["mask_relight"] else 0,
                                        L.ptr(terms), L.ptr(g_color), L.ptr(g_ws), L.ptr(g_dl),
                                        C.c_void_p(wsb.data_ptr()), wsb.numel() * 8, L.stream_ptr()), "cneus_neus_loss")
        ctx.shapes = (color_fine.shape, None if weight_sum is None else weight_sum.shape, gradient_error.shape,
                      None if delta_relight is None else delta_relight.shape)
        ctx.lambda_eikonal = opts["lambda_eikonal"]
        ctx.save_for_backward(g_color, g_ws if g_ws is not None else terms.new_empty(0),
                              g_dl if g_dl is not None else terms.new_empty(0))
        ctx.mark_non_differentiable(terms)
        return terms[0].clone(), terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        g_color, g_ws, g_dl = ctx.saved_tensors
        s_color, s_ws, s_eik, s_dl = ctx.shapes
        out_c = (g_color * g_loss).reshape(s_color)
        out_ws = (g_ws * g_loss).reshape(s_ws) if (s_ws is not None and g_ws.numel()) else None
        out_e = (g_loss * ctx.lambda_eikonal).reshape(s_eik)
        out_dl = (g_dl * g_loss).reshape(s_dl) if (s_dl is not None and g_dl.numel()) else None
        return out_c, out_ws, out_e, out_dl, None, None, None


class NeusLoss:
    """NeuS_Trainer.compute_loss (lib/models/NeuS_Trainer.py:129-171) with the LOSS cfg keys read at :66-76."""

    def __init__(self, loss_cfg=None, include_mask=True, renderer_type="Color_NeuS"):
        get = (loss_cfg.get if loss_cfg is not None else (lambda k, d: d))
        self.lambda_fine = get("LAMBDA_FINE", 1.0)
        self.lambda_eikonal = get("LAMBDA_EIKONAL", 0.1)
        self.lambda_mask = get("LAMBDA_MASK", 0.0)
        self.lambda_relight = get("LAMBDA_RELIGHT", 1.0)
        rgb_loss_type = get("RGB_LOSS_TYPE", "mse")
        assert self.lambda_fine != 0 and self.lambda_eikonal != 0
        if renderer_type == "Color_NeuS":
            assert self.lambda_relight != 0
        if rgb_loss_type not in ("mse", "l1"):
            raise ValueError(f"no such rgb loss type: {rgb_loss_type}")
        self.rgb_l1 = rgb_loss_type == "l1"
        self.include_mask = include_mask
        self.psnr = None   # device scalar of the last call; the reference's `psnr_toshow` without the per-step .item()

    def __call__(self, render_dict, **kwargs):
        has_relight = "delta_relight" in render_dict and self.lambda_relight != 0
        opts = dict(lambda_fine=float(self.lambda_fine), lambda_eikonal=float(self.lambda_eikonal),
                    lambda_mask=float(self.lambda_mask), lambda_relight=float(self.lambda_relight) if has_relight else 0.0,
                    rgb_l1=self.rgb_l1, mask_relight=bool(self.include_mask))
        mask = render_dict.get("mask")
        loss, terms = _LossFn.apply(render_dict["color_fine"], render_dict["weight_sum"] if self.lambda_mask != 0 else None,
                                    render_dict["gradient_error"], render_dict["delta_relight"] if has_relight else None,
                                    render_dict["rgb_map_gt"], mask, opts)
        loss_dict = {"loss": loss, "rgb_fine_loss": terms[1], "eikonal_loss": terms[2]}
        if self.lambda_mask != 0:
            loss_dict["mask_loss"] = terms[3]
        if has_relight:
            loss_dict["relight_loss"] = terms[4]
        self.psnr = -10.0 * torch.log10(terms[1])   # metrics/similarity.py mse2psnr; stays on the device
        return loss, loss_dict


def compute_loss(render_dict, loss_cfg=None, include_mask=True, renderer_type="Color_NeuS"):
    return NeusLoss(loss_cfg, include_mask, renderer_type)(render_dict)
