"""Golden vectors for SURVEY 8f #2 from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_train.py     ->  tests/golden/train_ops.npz

* NeuS_Trainer.compute_loss is taken from the reference source by AST (the module itself imports trimesh / imageio /
  kornia, which are not installed) and executed unmodified on seeded inputs, for four settings (mse / l1, with / without
  mask and relight terms); its autograd gradients are stored too.
* net_utils.build_optimizer_nerf + clip_gradient + NeuS_lr_scheduler run train.py's step order (train.py:70-77) for five
  steps on a toy parameter set whose gradients straddle the clipping threshold.
"""
import ast
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_import import REF_ROOT, CfgDict, load_reference  # noqa: E402


def reference_compute_loss():
    src = open(os.path.join(REF_ROOT, "lib", "models", "NeuS_Trainer.py")).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "NeuS_Trainer")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "compute_loss")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "F": F, "mse2psnr": lambda x: -10. * torch.log(x) / torch.log(torch.Tensor([10.]))}
    exec(compile(mod, "NeuS_Trainer.compute_loss", "exec"), ns)
    return ns["compute_loss"]


LOSS_CASES = {
    # name: (rgb type, lambda_mask, lambda_relight, include_mask, has delta_relight, B, S)
    "mse_mask_relight": ("mse", 0.1, 1.0, True, True, 96, 24),
    "l1_mask_relight": ("l1", 0.1, 1.0, True, True, 64, 16),
    "mse_nomask_relight": ("mse", 0.0, 1.0, False, True, 64, 16),
    "mse_mask_norelight": ("mse", 0.1, 0.0, True, False, 80, 0),
}


def _accept_verbose():
    """torch >= 2.7 dropped LRScheduler's `verbose` argument, which the reference still passes (net_utils.py:62): accept and
    ignore it so the unmodified reference classes construct (an environment shim, not a change of the reference)."""
    from torch.optim import lr_scheduler
    orig = lr_scheduler.LRScheduler.__init__

    def init(self, optimizer, last_epoch=-1, verbose=False):
        orig(self, optimizer, last_epoch)

    lr_scheduler.LRScheduler.__init__ = init


def main():
    load_reference()
    _accept_verbose()
    net_utils = importlib.import_module("lib.utils.net_utils")
    compute_loss = reference_compute_loss()
    out = {}
    for name, (kind, lm, lr_, inc, has_dl, B, S) in LOSS_CASES.items():
        g = torch.Generator().manual_seed(hash(name) % 1000 + 11)
        color = torch.rand(B, 3, generator=g).requires_grad_(True)
        gt = torch.rand(B, 3, generator=g)
        ws = (torch.rand(B, 1, generator=g) * 1.2 - 0.1)      # some values outside the clip interval
        ws[0, 0], ws[1, 0] = 1e-3, 1.0 - 1e-3                  # and exactly on its ends
        ws.requires_grad_(True)
        mask = (torch.rand(B, generator=g) > 0.4).float()
        eik = torch.rand([], generator=g).requires_grad_(True)
        dl = (torch.randn(B, S, 3, generator=g) * 0.05 + 0.01).requires_grad_(True) if has_dl else None
        self = types.SimpleNamespace(rgb_loss=torch.nn.MSELoss() if kind == "mse" else torch.nn.L1Loss(), lambda_fine=1.0,
                                     lambda_eikonal=0.1, lambda_mask=lm, lambda_relight=lr_, include_mask=inc, psnr_toshow=None)
        rd = {"rgb_map_gt": gt, "color_fine": color, "gradient_error": eik, "weight_sum": ws, "mask": mask if inc or lm != 0 else None}
        if has_dl:
            rd["delta_relight"] = dl
        loss, ld = compute_loss(self, rd)
        loss.backward()
        pre = f"loss/{name}/"
        for k, v in (("color", color), ("gt", gt), ("ws", ws), ("mask", mask), ("eik", eik)):
            out[pre + k] = v.detach().numpy()
        if has_dl:
            out[pre + "dl"] = dl.detach().numpy()
            out[pre + "g_dl"] = dl.grad.numpy()
        for k, v in ld.items():
            out[pre + "term_" + k] = v.detach().numpy()
        out[pre + "g_color"] = color.grad.numpy()
        out[pre + "g_eik"] = eik.grad.numpy()
        if lm != 0:
            out[pre + "g_ws"] = ws.grad.numpy()
        out[pre + "psnr"] = np.float32(self.psnr_toshow)

    # ---- optimizer: train.py:70-77 order on a toy model ----------------------------------------------------------------
    torch.manual_seed(5)
    shapes = [(64, 39), (256,), (256, 1), (3, 256), (), (57, 64), (1,)]
    model = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s) * 0.3) for s in shapes])
    ocfg = CfgDict(TYPE="adam", LR=5e-4, SCHEDULER_TYPE="NEUS", WARM_UP=3, LR_ALPHA=0.05)
    optimizer, scheduler = net_utils.build_optimizer_nerf(model, ocfg, -1, iterations=40)
    n_steps = 5
    g = torch.Generator().manual_seed(17)
    for i, p in enumerate(model):
        out[f"opt/p0_{i}"] = p.detach().numpy().copy()
    for s in range(n_steps):
        optimizer.zero_grad()
        out[f"opt/lr_{s}"] = np.float64(optimizer.param_groups[0]["lr"])
        for i, p in enumerate(model):
            scale = [3.0, 0.01, 1.0, 0.2, 5.0, 1e-4, 0.9][i] * (1.0 + s)      # norms on both sides of max_norm = 1
            p.grad = torch.randn(p.shape, generator=g) * scale / max(1.0, float(p.numel()) ** 0.5)
            out[f"opt/g_{s}_{i}"] = p.grad.numpy().copy()
        net_utils.clip_gradient(optimizer, 1.0, 2)
        for i, p in enumerate(model):
            out[f"opt/gclip_{s}_{i}"] = p.grad.numpy().copy()
        optimizer.step()
        scheduler.step()
        for i, p in enumerate(model):
            out[f"opt/p_{s}_{i}"] = p.detach().numpy().copy()
    st = optimizer.state_dict()["state"]
    for i in range(len(shapes)):
        out[f"opt/m_{i}"] = st[i]["exp_avg"].numpy().copy()
        out[f"opt/v_{i}"] = st[i]["exp_avg_sq"].numpy().copy()
    out["opt/n_steps"], out["opt/n_tensors"] = np.int64(n_steps), np.int64(len(shapes))
    # learning-rate schedule over a whole run (NEUS: warm-up 5000, alpha 0.05 in the configs; short run here)
    sch_model = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(2))])
    o2, s2 = net_utils.build_optimizer_nerf(sch_model, CfgDict(TYPE="adam", LR=5e-4, SCHEDULER_TYPE="NEUS", WARM_UP=10, LR_ALPHA=0.05), -1,
                                            iterations=60)
    lrs = []
    for _ in range(60):
        lrs.append(o2.param_groups[0]["lr"])
        o2.step()
        s2.step()
    out["sched/lrs"] = np.asarray(lrs, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "train_ops.npz"), **out)
    print("wrote train_ops.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
