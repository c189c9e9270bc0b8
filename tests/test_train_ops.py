"""SURVEY 8f #2 (loss + per-tensor clip + Adam + NeuS learning-rate schedule).

CPU: the oracle restatement (oracle/train_oracle.py) and the host-side scheduler against tests/golden/train_ops.npz, which
was produced by the unmodified reference (tests/golden/make_golden_train.py).  GPU: the CUDA path (csrc/loss.cu,
csrc/optim.cu through color_neus_b200.train_ops) against the same fixtures; bars written at each assert.
"""
import os

import numpy as np
import pytest
import torch

from helpers import HERE, rel_err
from oracle import train_oracle as TO

G = dict(np.load(os.path.join(HERE, "golden", "train_ops.npz")))
LOSS_CASES = {
    "mse_mask_relight": dict(rgb_l1=False, lambda_mask=0.1, lambda_relight=1.0, include_mask=True),
    "l1_mask_relight": dict(rgb_l1=True, lambda_mask=0.1, lambda_relight=1.0, include_mask=True),
    "mse_nomask_relight": dict(rgb_l1=False, lambda_mask=0.0, lambda_relight=1.0, include_mask=False),
    "mse_mask_norelight": dict(rgb_l1=False, lambda_mask=0.1, lambda_relight=0.0, include_mask=True),
}
TERMS = ["loss", "rgb_fine_loss", "eikonal_loss", "mask_loss", "relight_loss"]


def _loss_inputs(name, device="cpu", grad=True):
    pre = f"loss/{name}/"
    t = lambda k: torch.as_tensor(G[pre + k]).to(device)   # noqa: E731
    d = dict(color=t("color"), gt=t("gt"), ws=t("ws"), mask=t("mask"), eik=t("eik"), dl=t("dl") if pre + "dl" in G else None)
    if grad:
        for k in ("color", "ws", "eik", "dl"):
            if d[k] is not None:
                d[k].requires_grad_(True)
    return d


@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_oracle_loss_matches_reference(name):
    o, d = LOSS_CASES[name], _loss_inputs(name)
    loss, terms = TO.compute_loss(d["color"], d["gt"], d["eik"], d["ws"], d["mask"], d["dl"], 1.0, 0.1, o["lambda_mask"],
                                  o["lambda_relight"], o["rgb_l1"], o["include_mask"])
    loss.backward()
    pre = f"loss/{name}/"
    for k in TERMS:
        if pre + "term_" + k in G:
            assert rel_err(terms[k].detach().numpy(), G[pre + "term_" + k]) < 1e-6, k
    assert rel_err(d["color"].grad.numpy(), G[pre + "g_color"]) < 1e-6
    if pre + "g_ws" in G:
        assert rel_err(d["ws"].grad.numpy(), G[pre + "g_ws"]) < 1e-6
    if pre + "g_dl" in G:
        assert rel_err(d["dl"].grad.numpy(), G[pre + "g_dl"]) < 1e-5


def _opt_fixture():
    n_steps, n_t = int(G["opt/n_steps"]), int(G["opt/n_tensors"])
    return n_steps, n_t


def test_oracle_clip_adam_matches_reference():
    n_steps, n_t = _opt_fixture()
    p = [torch.as_tensor(G[f"opt/p0_{i}"]).clone() for i in range(n_t)]
    m = [torch.zeros_like(x) for x in p]
    v = [torch.zeros_like(x) for x in p]
    for s in range(n_steps):
        lr = float(G[f"opt/lr_{s}"])
        for i in range(n_t):
            g = torch.as_tensor(G[f"opt/g_{s}_{i}"])
            _, coef = TO.clip_coefficient(g, 1.0)
            gc = g * coef
            assert rel_err(gc.numpy(), G[f"opt/gclip_{s}_{i}"]) < 1e-6
            TO.adam_step(p[i], gc, m[i], v[i], s + 1, lr)
            assert rel_err(p[i].numpy(), G[f"opt/p_{s}_{i}"]) < 1e-6, (s, i)
    for i in range(n_t):
        assert rel_err(m[i].numpy(), G[f"opt/m_{i}"]) < 1e-5 and rel_err(v[i].numpy(), G[f"opt/v_{i}"]) < 1e-5


def test_lr_schedule_matches_reference():
    from color_neus_b200.train_ops import NeuS_lr_scheduler
    lrs = G["sched/lrs"]
    for it, lr in enumerate(lrs):
        if it > 0:
            assert abs(5e-4 * TO.neus_lr_factor(it, 10, 0.05, 60) - lr) < 1e-15
    # the torch scheduler object (host logic; a plain SGD stands in for the optimizer so this runs without a GPU)
    w = torch.nn.Parameter(torch.zeros(2))
    opt = torch.optim.SGD([w], lr=5e-4)
    sch = NeuS_lr_scheduler(opt, 10, 0.05, 60, -1)
    opt.param_groups[0]["lr"] = 5e-4
    got = []
    for _ in range(len(lrs)):
        got.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    assert np.abs(np.asarray(got) - lrs).max() < 1e-15
    assert set(sch.state_dict()) >= {"warm_up", "alpha", "end_iter", "last_epoch", "base_lrs"}


def test_no_cpu_path():
    from color_neus_b200 import train_ops as TR
    from color_neus_b200._lib import CneusError
    w = torch.nn.Parameter(torch.zeros(4))
    w.grad = torch.ones(4)
    opt = TR.FusedClipAdam([w], lr=1e-3)
    with pytest.raises(CneusError):
        opt.step()
    with pytest.raises(CneusError):
        TR.clip_gradient(torch.optim.Adam([w]), 1.0, 2)
    with pytest.raises(CneusError):
        TR.clip_gradient(opt, 1.0, 1)
    d = _loss_inputs("mse_mask_relight")
    with pytest.raises(CneusError):
        TR.NeusLoss({"LAMBDA_MASK": 0.1})({"color_fine": d["color"], "rgb_map_gt": d["gt"], "gradient_error": d["eik"],
                                           "weight_sum": d["ws"], "mask": d["mask"], "delta_relight": d["dl"]})


def test_state_dict_has_adam_layout():
    from color_neus_b200 import train_ops as TR
    ws = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 3))]
    a, b = TR.FusedClipAdam(ws, lr=5e-4, betas=(0.9, 0.99)), torch.optim.Adam(ws, lr=5e-4, betas=(0.9, 0.99))
    ka, kb = a.state_dict()["param_groups"][0], b.state_dict()["param_groups"][0]
    assert set(ka) == set(kb)
    a.load_state_dict(b.state_dict())   # a torch.optim.Adam checkpoint loads


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_gpu_loss_matches_reference(name):
    from color_neus_b200 import train_ops as TR
    o, d = LOSS_CASES[name], _loss_inputs(name, "cuda")
    fn = TR.NeusLoss({"LAMBDA_MASK": o["lambda_mask"], "LAMBDA_RELIGHT": o["lambda_relight"] if d["dl"] is not None else 1.0,
                      "RGB_LOSS_TYPE": "l1" if o["rgb_l1"] else "mse"}, include_mask=o["include_mask"],
                     renderer_type="Color_NeuS" if d["dl"] is not None else "NeuS")
    rd = {"color_fine": d["color"], "rgb_map_gt": d["gt"], "gradient_error": d["eik"], "weight_sum": d["ws"], "mask": d["mask"]}
    if d["dl"] is not None:
        rd["delta_relight"] = d["dl"]
    loss, ld = fn(rd)
    loss.backward()
    pre = f"loss/{name}/"
    for k in TERMS:
        if pre + "term_" + k in G:
            assert rel_err(ld[k].detach().cpu().numpy(), G[pre + "term_" + k]) < 2e-6, k   # fp64 sums vs torch's fp32 trees
    assert rel_err(d["color"].grad.cpu().numpy(), G[pre + "g_color"]) < 2e-6
    assert abs(float(d["eik"].grad) - float(G[pre + "g_eik"])) < 1e-7
    if pre + "g_ws" in G:
        assert rel_err(d["ws"].grad.cpu().numpy(), G[pre + "g_ws"]) < 2e-6
    if pre + "g_dl" in G:
        assert rel_err(d["dl"].grad.cpu().numpy(), G[pre + "g_dl"]) < 2e-5
    assert abs(float(fn.psnr) - float(G[pre + "psnr"])) < 1e-4


@pytest.mark.gpu
def test_gpu_clip_adam_matches_reference():
    """train.py:70-77 order through build_optimizer_nerf / clip_gradient / step / scheduler.step on the fixture's gradients."""
    from color_neus_b200 import train_ops as TR

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    n_steps, n_t = _opt_fixture()
    model = torch.nn.ParameterList([torch.nn.Parameter(torch.as_tensor(G[f"opt/p0_{i}"]).cuda()) for i in range(n_t)])
    opt, sch = TR.build_optimizer_nerf(model, Cfg(TYPE="adam", LR=5e-4, SCHEDULER_TYPE="NEUS", WARM_UP=3, LR_ALPHA=0.05), -1,
                                       iterations=40)
    for s in range(n_steps):
        opt.zero_grad()
        assert abs(opt.param_groups[0]["lr"] - float(G[f"opt/lr_{s}"])) < 1e-15
        for i, p in enumerate(model):
            p.grad = torch.as_tensor(G[f"opt/g_{s}_{i}"]).cuda()
        TR.clip_gradient(opt, 1.0, 2)
        opt.step()
        sch.step()
        for i, p in enumerate(model):
            # bars: clipped gradient 1e-6 (fp64 norm vs torch's fp32 norm), parameters 1e-6 of their magnitude
            assert rel_err(p.grad.cpu().numpy(), G[f"opt/gclip_{s}_{i}"]) < 1e-6, (s, i)
            assert rel_err(p.detach().cpu().numpy(), G[f"opt/p_{s}_{i}"]) < 1e-6, (s, i)
        norms = opt.last_grad_norms.cpu().numpy()
        want = np.array([np.linalg.norm(G[f"opt/g_{s}_{i}"].astype(np.float64).ravel()) for i in range(n_t)])
        assert np.abs(norms - want).max() / want.max() < 1e-6
    st = opt.state_dict()["state"]
    for i in range(n_t):
        assert rel_err(st[i]["exp_avg"].cpu().numpy(), G[f"opt/m_{i}"]) < 1e-5
        assert rel_err(st[i]["exp_avg_sq"].cpu().numpy(), G[f"opt/v_{i}"]) < 1e-5
        assert float(st[i]["step"]) == n_steps


@pytest.mark.gpu
def test_gpu_adam_without_clip_equals_torch_adam():
    from color_neus_b200 import train_ops as TR
    torch.manual_seed(3)
    shapes = [(300, 70), (1000003,), ()]
    a = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    b = [torch.nn.Parameter(x.detach().clone()) for x in a]
    oa, ob = TR.FusedClipAdam(a, lr=1e-2, betas=(0.9, 0.99), weight_decay=0.01), torch.optim.Adam(b, lr=1e-2, betas=(0.9, 0.99),
                                                                                                 weight_decay=0.01)
    for _ in range(4):
        for x, y in zip(a, b):
            x.grad = torch.randn_like(x)
            y.grad = x.grad.clone()
        oa.step()
        ob.step()
    for x, y in zip(a, b):
        assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 1e-6


@pytest.mark.gpu
def test_renderer_sees_the_weights_the_fused_optimizer_wrote():
    """Three training steps of the drop-in renderer with FusedClipAdam (whose kernel updates the parameters through raw
    device pointers) against the same three steps with torch.optim.Adam + per-tensor clip_grad_norm_ (the reference's
    train.py:70-77): the renderer's packed device weights must follow the optimiser, i.e. colours and losses agree after
    EVERY step, not only the first (NetHandle re-packs when the parameters' version counters move; the fused step bumps
    them).  Also: a stale pack is detected by a direct raw write + `invalidate()`."""
    import __graft_entry__ as g
    import color_neus_b200 as cn
    from color_neus_b200 import train_ops as TR
    from helpers import O
    cfg = O.default_cfg("Color_NeuS", 64, 64, 256, 8, 0.3)
    torch.manual_seed(1)
    a = cn.Color_NeuS(g._Cfg(cfg)).cuda().train()
    b = cn.Color_NeuS(g._Cfg(cfg)).cuda().train()
    b.load_state_dict(a.state_dict())
    c2w = O.pose_spherical(30.0, -30.0, 2.8)
    ro, rd = O.get_rays_at(c2w, torch.tensor([6.0 * 6, 6.0 * 6]), 6, 6)
    near, far = O.near_far_from_sphere(ro, rd)
    ro, rd, near, far = ro.cuda(), rd.cuda(), near.cuda(), far.cuda()
    gt = torch.rand(36, 3, generator=torch.Generator().manual_seed(2)).cuda()
    opt_a = TR.FusedClipAdam(a.parameters(), lr=5e-3, betas=(0.9, 0.99))       # a large lr: three steps move the image visibly
    opt_b = torch.optim.Adam(b.parameters(), lr=5e-3, betas=(0.9, 0.99))
    first = None
    for it in range(3):
        outs = []
        for ren, opt, fused in ((a, opt_a, True), (b, opt_b, False)):
            opt.zero_grad(set_to_none=True)
            r = ren(ro, rd, near, far, perturb_overwrite=0)
            loss = torch.nn.functional.mse_loss(r["color_fine"], gt) + 0.1 * r["gradient_error"]
            loss.backward()
            if fused:
                TR.clip_gradient(opt, 1.0, 2)
            else:
                for p in ren.parameters():
                    torch.nn.utils.clip_grad_norm_(p, 1.0, 2)
            opt.step()
            outs.append((r["color_fine"].detach().clone(), float(loss)))
        (ca, la), (cb, lb) = outs
        if first is None:
            first = ca
        assert rel_err(ca.cpu(), cb.cpu()) < 2e-4, it
        assert abs(la - lb) < 2e-4 * max(1.0, abs(lb)), (it, la, lb)
    assert rel_err(ca.cpu(), first.cpu()) > 1e-3, "three optimiser steps did not change the rendered colours"
    # raw writes behind the handle's back (a second tensor object over the same storage has its own version counter): the
    # renderer keeps using the packed copy until invalidate() forces the re-pack
    a.eval()
    with torch.no_grad():
        before = a(ro, rd, near, far, perturb_overwrite=0)["color_fine"].clone()
        w = a.color_network.lin4.bias
        alias = torch.empty(0, device="cuda").set_(w.detach().untyped_storage(), 0, w.shape)
        alias.copy_(alias + 0.5)
        stale = a(ro, rd, near, far, perturb_overwrite=0)["color_fine"].clone()
        a.handle().invalidate()
        fresh = a(ro, rd, near, far, perturb_overwrite=0)["color_fine"]
    assert torch.equal(stale, before)
    assert rel_err(fresh.cpu(), before.cpu()) > 1e-2
