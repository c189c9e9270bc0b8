"""Pin the CPU oracle against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from helpers import CASE_NAMES, O, T, load_case, rel_err, MG

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_stagewise_mlps(name):
    cfg, Pn, G = load_case(name)
    P = O.to_torch(Pn)
    pts, dirs = T(G["st_pts"]), T(G["st_dirs"])
    assert rel_err(O.embed(pts * cfg["SDF"]["SCALE"], 6), G["st_embed6"]) < 1e-6
    y = O.sdf_forward(P, cfg, pts)
    assert rel_err(y, G["st_sdf_out"]) < 2e-6
    g = O.sdf_gradient(P, cfg, pts)
    assert rel_err(g, G["st_grad"]) < 5e-6
    cg = O.color_forward(P, cfg, pts, T(G["st_grad"]), dirs, T(G["st_sdf_out"])[:, 1:])
    assert rel_err(cg, G["st_color"]) < 2e-6
    if cfg["TYPE"] == "Color_NeuS":
        c, d = O.relight_forward(P, cfg, T(G["st_color"]), pts, dirs, T(G["st_grad"]))
        assert rel_err(c, G["st_relit"]) < 2e-6
        assert rel_err(d, G["st_drgb"]) < 5e-6


@pytest.mark.parametrize("name", [n for n in CASE_NAMES if MG.CASES[n][2] > 0])
def test_stagewise_upsample(name):
    cfg, Pn, G = load_case(name)
    P = O.to_torch(Pn)
    ro, rd = T(G["rays_o"]), T(G["rays_d"])
    z0 = O.coarse_z(cfg, T(G["near"]), T(G["far"]), T(G["t_rand"]))
    assert np.array_equal(z0.numpy(), G["us_z0"])
    m = cfg["N_IMPORTANCE"] // cfg["UP_SAMPLE_STEPS"]
    newz = O.up_sample(ro, rd, T(G["us_z0"]), T(G["us_sdf0"]), m, 64)
    assert np.abs(newz.numpy() - G["us_new_z"]).max() < 2e-6
    z1, sdf1 = O.cat_z_vals(P, cfg, ro, rd, T(G["us_z0"]), T(G["us_new_z"]), T(G["us_sdf0"]), last=False)
    assert np.array_equal(z1.numpy(), G["us_z1"])
    assert np.abs(sdf1.numpy() - G["us_sdf1"]).max() < 2e-6


@pytest.mark.parametrize("name", CASE_NAMES)
def test_render_core_given_z(name):
    """Per-sample tensors are only comparable with z_vals injected (SURVEY.md section 4 conditioning caveat)."""
    cfg, Pn, G = load_case(name)
    P = O.to_torch(Pn)
    r = O.render_forward(P, cfg, T(G["rays_o"]), T(G["rays_d"]), T(G["near"]), T(G["far"]), z_vals=T(G["z_vals"]))
    # per-ray outputs 3e-5; per-SAMPLE weights of a sharp surface (inv_s = 403, a real zero crossing) differ by up to
    # ~5e-5 of the largest weight between two fp32 evaluation orders of the same math (closed-form gradient here, autograd in
    # the reference): sdf * inv_s amplifies 1-ulp differences of the SDF (SURVEY section 4 conditioning caveat)
    sharp = float(np.exp(10.0 * float(Pn["deviation_network.variance"]))) > 200.0
    for k in ("color_fine", "weight_sum", "depth", "weights", "gradients", "cdf_fine", "weight_max", "s_val"):
        assert rel_err(r[k], G["fwd_" + k]) < (1e-4 if sharp and k in ("weights", "weight_max", "cdf_fine") else 3e-5), k
    assert np.array_equal(r["inside_sphere"].numpy(), G["fwd_inside_sphere"])
    assert abs(float(r["gradient_error"]) - float(G["fwd_gradient_error"])) < 1e-5 * max(1.0, float(G["fwd_gradient_error"]))
    if cfg["TYPE"] == "Color_NeuS":
        assert rel_err(r["global_color"], G["fwd_global_color"]) < 3e-5
        assert rel_err(r["delta_relight"], G["fwd_delta_relight"]) < 3e-5


@pytest.mark.parametrize("name", CASE_NAMES)
def test_full_forward(name):
    """End to end incl. hierarchical sampling from the recorded RNG draw: per-ray outputs <= 1e-4."""
    cfg, Pn, G = load_case(name)
    P = O.to_torch(Pn)
    r = O.render_forward(P, cfg, T(G["rays_o"]), T(G["rays_d"]), T(G["near"]), T(G["far"]), t_rand=T(G["t_rand"]))
    assert np.abs(r["z_vals"].numpy() - G["z_vals"]).max() < 5e-4
    for k in ("color_fine", "weight_sum", "depth"):
        assert rel_err(r[k], G["fwd_" + k]) < 1e-4, k


@pytest.mark.parametrize("name", ["c2_color_init", "c2_neus_idr", "c1_small_sdf"])
def test_training_gradients(name):
    """Oracle autograd (double-backward through the restated SDF) vs the reference's parameter gradients."""
    cfg, Pn, G = load_case(name)
    P = O.to_torch(Pn, requires_grad=True)
    nb = G["bw_rgb_gt"].shape[0]
    ro = T(G["rays_o"][:nb]).requires_grad_(True)
    rd = T(G["rays_d"][:nb]).requires_grad_(True)
    ret = O.render_forward(P, cfg, ro, rd, T(G["near"][:nb]), T(G["far"][:nb]), z_vals=T(G["bw_z_vals"]),
                           grad_mode="autograd")
    mask = T(G["bw_mask"])
    loss = torch.nn.functional.mse_loss(ret["color_fine"], T(G["bw_rgb_gt"])) + 0.1 * ret["gradient_error"]
    loss = loss + 0.1 * torch.nn.functional.binary_cross_entropy(ret["weight_sum"].squeeze(-1).clip(1e-3, 1 - 1e-3), mask)
    if "delta_relight" in ret:
        loss = loss + torch.mean(ret["delta_relight"] * mask[:, None, None]) ** 2
    loss.backward()
    assert abs(float(loss) - float(G["bw_loss"])) < 1e-5 * max(1.0, abs(float(G["bw_loss"])))
    assert rel_err(ro.grad, G["bw_d_rays_o"]) < 2e-3
    assert rel_err(rd.grad, G["bw_d_rays_d"]) < 2e-3
    for k, p in P.items():
        g = p.grad.detach().numpy().reshape(-1) if p.grad is not None else np.zeros(p.numel(), np.float32)
        gn = float(G["bwgn_" + k])
        assert abs(np.linalg.norm(g.astype(np.float64)) - gn) <= 2e-3 * gn + 1e-7, k
        ref = G["bwg_" + k]
        assert np.abs(g[MG.grad_sample_index(g.size)] - ref).max() <= 2e-3 * max(np.abs(ref).max(), 1e-7) + 1e-8, k


@pytest.mark.parametrize("name", ["c2_color_trained", "c1_small_sdf"])
def test_mesh_queries(name):
    cfg, Pn, G = load_case(name)
    P = O.to_torch(Pn)
    u = O.extract_fields(P, cfg, G["grid_bmin"], G["grid_bmax"], int(G["grid_res"]))
    assert np.abs(u - G["grid_u"]).max() < 2e-6
    c = O.extract_color(P, cfg, G["vc_vertices"])
    assert rel_err(c, G["vc_color"]) < 1e-5
