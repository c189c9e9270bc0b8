"""GPU parity of the training backward (cneus_render_backward) against the reference's own loss.backward():
parameter gradients, ray gradients and the loss value of one training step (tests/golden/*.npz, `bw_*` entries),
plus a comparison with the CPU oracle's autograd on the same batch."""
import numpy as np
import pytest
import torch

from helpers import CASE_NAMES, MG, O, T, load_case, record, rel_err
from test_gpu_parity import cu, make_renderer

pytestmark = pytest.mark.gpu


def training_loss(ret, rgb_gt, mask):
    loss = torch.nn.functional.mse_loss(ret["color_fine"], rgb_gt) + 0.1 * ret["gradient_error"]
    loss = loss + 0.1 * torch.nn.functional.binary_cross_entropy(ret["weight_sum"].squeeze(-1).clip(1e-3, 1 - 1e-3), mask)
    if "delta_relight" in ret:
        loss = loss + torch.mean(ret["delta_relight"] * mask[:, None, None]) ** 2
    return loss


# relative bar on every gradient (ray gradients, per-tensor norms, 512 sampled entries per tensor): 3x the worst error the
# tcgen05 path measured over all nine cases on a B200 -- 1.40e-3 (c3_64_128, a small-norm tensor; ray gradients <= 2.4e-5;
# profiles/r2_parity_errors.json) -- not a guess.  Round 1 used 5e-3 on fixtures that rendered empty space.
BACKWARD_BAR = 4e-3


@pytest.mark.parametrize("name", CASE_NAMES)
def test_training_step_gradients_match_reference(name):
    cfg, Pn, G = load_case(name)
    ren = make_renderer(cfg, Pn).train()
    nb = G["bw_rgb_gt"].shape[0]
    ro = cu(G["rays_o"][:nb]).requires_grad_(True)
    rd = cu(G["rays_d"][:nb]).requires_grad_(True)
    ret = ren(ro, rd, cu(G["near"][:nb]), cu(G["far"][:nb]), z_vals=cu(G["bw_z_vals"]))
    assert ret["color_fine"].requires_grad and ret["gradient_error"].requires_grad
    loss = training_loss(ret, cu(G["bw_rgb_gt"]), cu(G["bw_mask"]))
    assert abs(float(loss) - float(G["bw_loss"])) < 1e-4 * max(1.0, abs(float(G["bw_loss"])))
    loss.backward()
    e_o = record("backward", name, "d_rays_o", rel_err(ro.grad.cpu(), G["bw_d_rays_o"]))
    e_d = record("backward", name, "d_rays_d", rel_err(rd.grad.cpu(), G["bw_d_rays_d"]))
    worst = 0.0
    for k, p in ren.named_parameters():
        assert p.grad is not None, k
        g = p.grad.detach().cpu().numpy().reshape(-1)
        gn = float(G["bwgn_" + k])
        ref = G["bwg_" + k]
        e_norm = abs(np.linalg.norm(g.astype(np.float64)) - gn) / (gn + 1e-12) if gn > 1e-9 else np.abs(g).max()
        e_samp = np.abs(g[MG.grad_sample_index(g.size)] - ref).max() / max(np.abs(ref).max(), 1e-9) if np.abs(ref).max() > 1e-9 else 0.0
        worst = max(worst, e_norm, e_samp)
        record("backward", name, "worst_param_grad", worst)
        assert e_norm < BACKWARD_BAR and e_samp < BACKWARD_BAR, (k, e_norm, e_samp)
    assert e_o < BACKWARD_BAR and e_d < BACKWARD_BAR, (e_o, e_d)
    print(f"{name}: worst relative gradient error {worst:.2e}")


def test_backward_matches_oracle_autograd_on_fresh_batch():
    cfg = O.default_cfg("Color_NeuS", 64, 64, 256, 8, 0.4)
    Pn = O.make_params(cfg, seed=4, trained_like=True)
    ren = make_renderer(cfg, Pn).train()
    c2w = O.pose_spherical(10.0, -35.0, 2.7)
    ro, rd = O.get_rays_at(c2w, torch.tensor([6.0 * 4, 6.0 * 4]), 4, 4)
    near, far = O.near_far_from_sphere(ro, rd)
    P = O.to_torch(Pn, requires_grad=True)
    t_rand = torch.rand([16, 1], generator=torch.Generator().manual_seed(2))
    z = O.hierarchical_z(O.to_torch(Pn), cfg, ro, rd, near, far, t_rand)
    rs = np.random.RandomState(0)
    gt, mask = T(rs.uniform(0, 1, (16, 3))), T((rs.uniform(0, 1, 16) > 0.4).astype(np.float32))
    ref = O.render_forward(P, cfg, ro, rd, near, far, z_vals=z, grad_mode="autograd")
    # exercise every differentiable output, not only the ones the trainer uses
    extra = lambda r: (r["depth"].sum() * 0.3 + (r["weights"] ** 2).sum() + r["weight_max"].sum() * 0.2 + r["global_color"].sum() * 0.1 +
                       (r["gradients"] * 0.01).sum() + r["cdf_fine"].mean() + r["s_val"].sum())
    (training_loss(ref, gt, mask) + extra(ref)).backward()
    ret = ren(ro.cuda(), rd.cuda(), near.cuda(), far.cuda(), z_vals=z.cuda())
    (training_loss(ret, gt.cuda(), mask.cuda()) + extra(ret)).backward()
    for k, p in ren.named_parameters():
        a, b = p.grad.detach().cpu().numpy().reshape(-1), P[k].grad.numpy().reshape(-1)
        assert np.abs(a - b).max() <= 5e-3 * max(np.abs(b).max(), 1e-8) + 1e-9, k


@pytest.mark.parametrize("n_rays,n_imp", [(3, 128), (7, 64)])
def test_fused_recompute_equals_layerwise_backward(n_rays, n_imp):
    """A/B of the two recompute paths of cneus_render_backward (one fused kernel launch with training dumps vs GEMMs +
    element-wise kernels layer by layer) on point counts that do not fill the last 128-point tile (3 x 192 = 4.5 tiles,
    7 x 128 = 7 tiles): same gradients within the backward's bar, chunked passes included."""
    from color_neus_b200 import _lib as L
    cfg = O.default_cfg("Color_NeuS", 64, n_imp, 256, 8, 0.45)
    Pn = O.make_params(cfg, seed=6, trained_like=True)
    ren = make_renderer(cfg, Pn).train()
    c2w = O.pose_spherical(40.0, -25.0, 2.7)
    ro, rd = O.get_rays_at(c2w, torch.tensor([6.0 * 3, 6.0 * 3]), 3, 3)
    ro, rd = ro[:n_rays].cuda().contiguous(), rd[:n_rays].cuda().contiguous()
    near, far = O.near_far_from_sphere(ro, rd)
    rs = np.random.RandomState(1)
    gt, mask = cu(rs.uniform(0, 1, (n_rays, 3)).astype(np.float32)), cu((rs.uniform(0, 1, n_rays) > 0.3).astype(np.float32))
    lib = L.lib()
    grads = {}
    try:
        for mode, (fused, chunk) in {"layerwise": (0, 0), "fused_recompute": (1, 0), "fused": (3, 0), "fused_chunked": (3, 2)}.items():
            lib.cneus_backward_fused_recompute(fused)
            lib.cneus_backward_chunk_rays(chunk)
            ren.zero_grad(set_to_none=True)
            r_o, r_d = ro.clone().requires_grad_(True), rd.clone().requires_grad_(True)
            ret = ren(r_o, r_d, near, far, perturb_overwrite=0)
            training_loss(ret, gt, mask).backward()
            grads[mode] = {k: p.grad.detach().clone() for k, p in ren.named_parameters()}
            grads[mode]["rays_o"], grads[mode]["rays_d"] = r_o.grad.clone(), r_d.grad.clone()
    finally:
        lib.cneus_backward_fused_recompute(1)
        lib.cneus_backward_chunk_rays(0)
    for mode in ("fused_recompute", "fused", "fused_chunked"):
        for k, ref in grads["layerwise"].items():
            err = float((grads[mode][k] - ref).abs().max() / (ref.abs().max() + 1e-12))
            assert err < 5e-3, (mode, k, err)
