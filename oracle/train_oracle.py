"""TEST INFRASTRUCTURE -- CPU restatement (plain torch fp32 formulas) of the per-step work around the renderer
(SURVEY.md section 8f #2).  Never imported by the product path (color_neus_b200/).

Pinned by tests/golden/train_ops.npz, which tests/golden/make_golden_train.py generates from the UNMODIFIED reference
(NeuS_Trainer.compute_loss, net_utils.clip_gradient, build_optimizer_nerf -> torch.optim.Adam, NeuS_lr_scheduler).
"""
import math

import torch


def compute_loss(color_fine, rgb_gt, gradient_error, weight_sum=None, mask=None, delta_relight=None, lambda_fine=1.0,
                 lambda_eikonal=0.1, lambda_mask=0.0, lambda_relight=1.0, rgb_l1=False, include_mask=True):
    """NeuS_Trainer.compute_loss (lib/models/NeuS_Trainer.py:129-171).  Returns (loss, dict of terms)."""
    d = color_fine - rgb_gt
    rgb = d.abs().mean() if rgb_l1 else (d * d).mean()                      # :133-135 (MSELoss / L1Loss, :71-74)
    loss = lambda_fine * rgb + lambda_eikonal * gradient_error             # :136-139
    terms = {"rgb_fine_loss": rgb, "eikonal_loss": gradient_error}
    if lambda_mask != 0:                                                    # :141-144
        p = weight_sum.reshape(-1).clamp(1e-3, 1.0 - 1e-3)
        m = mask.reshape(-1)
        bce = -(m * torch.log(p).clamp_min(-100.0) + (1.0 - m) * torch.log(1.0 - p).clamp_min(-100.0)).mean()
        loss = loss + lambda_mask * bce
        terms["mask_loss"] = bce
    if lambda_relight != 0 and delta_relight is not None:                   # :146-155
        dl = delta_relight * mask.reshape(-1, 1, 1) if include_mask else delta_relight
        rel = dl.mean() ** 2
        loss = loss + lambda_relight * rel
        terms["relight_loss"] = rel
    terms["loss"] = loss
    return loss, terms


def clip_coefficient(grad, max_norm):
    """clip_grad_norm_(p, max_norm, 2) for ONE tensor (net_utils.py:174-184 calls it per parameter):
    coefficient = min(1, max_norm / (||g||_2 + 1e-6)); returns (norm, coefficient)."""
    norm = torch.linalg.vector_norm(grad.double(), 2).float()
    return norm, torch.clamp(max_norm / (norm + 1e-6), max=1.0)


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.99, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam single-tensor update (amsgrad False), `step` 1-based.  Mutates and returns its arguments."""
    if weight_decay != 0:
        grad = grad + weight_decay * param
    exp_avg.lerp_(grad, 1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = (exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-(lr / bc1))
    return param, exp_avg, exp_avg_sq


def neus_lr_factor(it, warm_up, alpha, end_iter):
    """NeuS_lr_scheduler._get_lr_neus (net_utils.py:64-70) for last_epoch = it."""
    if it < warm_up:
        return it / warm_up
    progress = (it - warm_up) / (end_iter - warm_up)
    return (math.cos(math.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha
