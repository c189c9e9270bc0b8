"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference through oracle/ref_import.py, loads the deterministic synthetic
parameters of oracle.neus_oracle.make_params into the reference modules, runs the reference's
own functions on seeded inputs and stores inputs + outputs as small .npz fixtures next to this
script.  The fixtures are what pins the oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py) on machines where /root/reference does not exist.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import neus_oracle as O  # noqa: E402
from oracle.ref_import import CfgDict, load_reference  # noqa: E402

CASES = {
    # name: (kind, n_samples, n_importance, sdf_hidden, sdf_layers, variance, trained_like, n_rays)
    "c2_color_init": ("Color_NeuS", 64, 64, 256, 8, 0.3, False, 48),
    "c2_color_trained": ("Color_NeuS", 64, 64, 256, 8, 0.5, True, 48),
    "c2_neus_idr": ("NeuS", 64, 64, 256, 8, 0.3, True, 32),
    "c1_small_sdf": ("Color_NeuS", 64, 0, 128, 4, 0.3, False, 64),
    "c3_64_128": ("Color_NeuS", 64, 128, 256, 8, 0.4, True, 16),
    "c4_128_128": ("Color_NeuS", 128, 128, 256, 8, 0.3, False, 16),
    # late-training sharpness: variance 0.6 -> inv_s = e^6 = 403 (SURVEY 8d "sharp" state, fields.py:284-286): the regime where
    # one-pass reduced-precision schemes were measured to fail the 1e-4 bar
    "c2_color_sharp": ("Color_NeuS", 64, 64, 256, 8, 0.6, True, 32),
    "c3_sharp_64_128": ("Color_NeuS", 64, 128, 256, 8, 0.6, True, 16),
    # a state that has really been optimised: 2000 steps of this repo's own training path on an analytic two-sphere scene
    # (tools/train_synthetic.py, run on a B200; the state_dict is stored next to this script), then the UNMODIFIED reference
    # evaluated on that state like on the synthetic ones
    "c2_color_optimised": ("Color_NeuS", 64, 64, 256, 8, None, "optimised_state.npz", 48),
}


def case_params(name):
    """Parameters of a case: the deterministic synthetic generator, or a stored state_dict (really-optimised cases)."""
    kind, n_s, n_i, hid, lay, var, trained, n_rays = CASES[name]
    if isinstance(trained, str):
        with np.load(os.path.join(HERE, trained)) as z:
            return {k: z[k] for k in z.files}
    cfg, _, _ = case_cfg(name)
    return O.make_params(cfg, seed=1, trained_like=trained)


def available_cases():
    """Cases whose inputs exist (a stored-state case needs its state file AND its fixture)."""
    out = []
    for name, c in CASES.items():
        if isinstance(c[6], str) and not (os.path.isfile(os.path.join(HERE, c[6])) and os.path.isfile(os.path.join(HERE, name + ".npz"))):
            continue
        out.append(name)
    return out


def case_cfg(name):
    kind, n_s, n_i, hid, lay, var, trained, n_rays = CASES[name]
    cfg = O.default_cfg(kind, n_s, n_i, hid, lay, 0.3 if var is None else var)   # var None: taken from the stored state
    return cfg, trained, n_rays


def params_digest(P):
    h = hashlib.sha256()
    for k in sorted(P):
        h.update(k.encode())
        h.update(np.ascontiguousarray(P[k]).tobytes())
    return h.hexdigest()


def synth_rays(n_rays, seed):
    """Half zoomed-in rays (most hit the object), half wide rays; reference get_rays_at ordering."""
    ns = load_reference()
    side = 8
    c2w = O.pose_spherical(30.0 + seed, -30.0, 2.8)
    outs_o, outs_d = [], []
    for focal_mul in (6.0, 1.2):
        f = torch.tensor([focal_mul * side, focal_mul * side])
        o, d = ns.ray_utils.get_rays_at(c2w, f, side, side, normalize=True)
        outs_o.append(o.reshape(-1, 3))
        outs_d.append(d.reshape(-1, 3))
    o = torch.cat(outs_o)[: max(n_rays, 2)]
    d = torch.cat(outs_d)[: max(n_rays, 2)]
    if n_rays <= 64:
        # interleave so that both kinds are present in small batches
        idx = torch.arange(128).reshape(2, 64).t().reshape(-1)[:n_rays]
        o, d = torch.cat(outs_o)[idx], torch.cat(outs_d)[idx]
    near, far = ns.ray_utils.near_far_from_sphere(o, d)
    return o.contiguous(), d.contiguous(), near.contiguous(), far.contiguous()


def build_reference(cfg, P):
    ns = load_reference()
    cls = ns.Color_NeuS if cfg["TYPE"] == "Color_NeuS" else ns.NeuS
    torch.manual_seed(1)
    r = cls(CfgDict(cfg))
    sd = {k: torch.as_tensor(v).reshape(r.state_dict()[k].shape) for k, v in P.items()}
    r.load_state_dict(sd, strict=True)
    return r


def grad_sample_index(numel):
    """Big gradient tensors are pinned by their L2 norm plus 512 fixed entries (keeps fixtures small)."""
    if numel <= 4096:
        return np.arange(numel)
    return np.sort(np.random.RandomState(11).choice(numel, 512, replace=False))


def npy(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def make_case(name):
    cfg, trained, n_rays = case_cfg(name)
    P = case_params(name)
    ren = build_reference(cfg, P)
    rays_o, rays_d, near, far = synth_rays(n_rays, seed=len(name))
    out = {"params_sha256": np.array(params_digest(P))}
    out.update(rays_o=npy(rays_o), rays_d=npy(rays_d), near=npy(near), far=npy(far))

    # ---- the one CPU RNG draw of NeuS.forward (NeuS.py:325)
    torch.manual_seed(7)
    t_rand = torch.rand([n_rays, 1])
    out["t_rand"] = npy(t_rand)

    # ---- full forward with captured z_vals
    captured = {}
    orig_core = ren.render_core

    def spy(rays_o, rays_d, z_vals, *a, **k):
        captured["z_vals"] = z_vals.detach().clone()
        r = orig_core(rays_o, rays_d, z_vals, *a, **k)
        captured["core"] = r
        return r

    ren.render_core = spy
    torch.manual_seed(7)
    ret = ren(rays_o, rays_d, near, far)
    ren.render_core = orig_core
    out["z_vals"] = npy(captured["z_vals"])
    for k, v in ret.items():
        out["fwd_" + k] = npy(v)
    for k in ("sdf", "dists", "mid_z_vals"):
        out["core_" + k] = npy(captured["core"][k])

    # ---- stage-wise: embedder, SDF, gradient, colour, relight on section mid-points of a few rays
    nr = min(2, n_rays)
    mid = captured["core"]["mid_z_vals"][:nr]
    pts = (rays_o[:nr, None, :] + rays_d[:nr, None, :] * mid[..., :, None]).reshape(-1, 3).detach()
    dirs = rays_d[:nr, None, :].expand(nr, mid.shape[1], 3).reshape(-1, 3)
    out["st_pts"], out["st_dirs"] = npy(pts), npy(dirs)
    out["st_embed6"] = npy(ren.sdf_network.embed_fn_fine(pts * ren.sdf_network.scale))
    y = ren.sdf_network(pts)
    out["st_sdf_out"] = npy(y)
    g = ren.sdf_network.gradient(pts.clone()).squeeze(1)
    out["st_grad"] = npy(g)
    cg = ren.color_network(pts, g, dirs, y[:, 1:])
    out["st_color"] = npy(cg)
    if cfg["TYPE"] == "Color_NeuS":
        c, d = ren.relight_network(cg, pts, dirs, gradients=g)
        out["st_relit"], out["st_drgb"] = npy(c), npy(d)

    # ---- stage-wise: up_sample / cat_z_vals round 0 on the coarse samples
    if cfg["N_IMPORTANCE"] > 0:
        with torch.no_grad():
            z0 = O.coarse_z(cfg, near, far, t_rand)
            sdf0 = ren.sdf_network.sdf((rays_o[:, None, :] + rays_d[:, None, :] * z0[..., :, None]).reshape(-1, 3))
            sdf0 = sdf0.reshape(z0.shape)
            m = cfg["N_IMPORTANCE"] // cfg["UP_SAMPLE_STEPS"]
            newz = ren.up_sample(rays_o, rays_d, z0, sdf0, m, 64)
            z1, sdf1 = ren.cat_z_vals(rays_o, rays_d, z0, newz, sdf0, last=False)
        out.update(us_z0=npy(z0), us_sdf0=npy(sdf0), us_new_z=npy(newz), us_z1=npy(z1), us_sdf1=npy(sdf1))

    # ---- training step gradients (autograd ON, reference double-backward), loss like NeuS_Trainer.compute_loss
    rs = np.random.RandomState(3)
    nb = min(8, n_rays)
    rgb_gt = torch.as_tensor(rs.uniform(0, 1, size=(nb, 3)).astype(np.float32))
    ro = rays_o[:nb].clone().requires_grad_(True)
    rd = rays_d[:nb].clone().requires_grad_(True)
    ren.zero_grad()
    captured.clear()
    ren.render_core = spy
    torch.manual_seed(7)
    ret = ren(ro, rd, near[:nb], far[:nb])
    ren.render_core = orig_core
    mask = (ret["weight_sum"].detach().squeeze(-1) > 0.5).float()
    loss = torch.nn.functional.mse_loss(ret["color_fine"], rgb_gt) + 0.1 * ret["gradient_error"]
    loss = loss + 0.1 * torch.nn.functional.binary_cross_entropy(ret["weight_sum"].squeeze(-1).clip(1e-3, 1 - 1e-3), mask)
    if "delta_relight" in ret:
        loss = loss + torch.mean(ret["delta_relight"] * mask[:, None, None]) ** 2
    loss.backward()
    out["bw_rgb_gt"], out["bw_mask"], out["bw_loss"] = npy(rgb_gt), npy(mask), npy(loss)
    out["bw_z_vals"] = npy(captured["z_vals"])
    out["bw_d_rays_o"], out["bw_d_rays_d"] = npy(ro.grad), npy(rd.grad)
    for k, p in ren.named_parameters():
        g = npy(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
        out["bwgn_" + k] = np.array(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out["bwg_" + k] = g[grad_sample_index(g.size)]

    # ---- mesh-side queries (SDF grid through the reference's own extract_fields; per-vertex colour)
    ns = load_reference()
    bmin, bmax = torch.tensor([-0.4, -0.45, -0.5]), torch.tensor([0.5, 0.45, 0.4])
    res = 10
    u = ns.neus_mod.extract_fields(bmin, bmax, "cpu", res, lambda p: -ren.sdf_network.sdf(p), N=4)
    out["grid_bmin"], out["grid_bmax"], out["grid_res"], out["grid_u"] = npy(bmin), npy(bmax), np.array(res), u
    verts = rs.uniform(-0.4, 0.4, size=(70, 3)).astype(np.float32)
    out["vc_vertices"] = verts
    out["vc_color"] = ns.neus_mod.extract_color(verts, "cpu", ren.sdf_network, ren.color_network, N=64)

    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(out)} arrays")


RAY_CASES = {
    # name: (n_cam, H, W, n_rays, normalize, opengl, with_mask, mask_rate, seed)
    "nomask": (3, 6, 5, 16, True, False, False, 0.9, 11),
    "mask": (2, 7, 6, 24, True, True, True, 0.7, 12),
    "mask_few_valid": (2, 5, 4, 12, False, False, True, 0.9, 13),
}


def ray_case_inputs(name):
    n_cam, H, W, n_rays, normalize, opengl, with_mask, mask_rate, seed = RAY_CASES[name]
    g = torch.Generator().manual_seed(seed)
    c2w = torch.stack([O.pose_spherical(20.0 * k + seed, -30.0 + 5 * k, 2.5 + 0.1 * k) for k in range(n_cam)])
    focal = torch.tensor([1.3 * W, 1.1 * W])
    image = torch.rand(n_cam, H, W, 3, generator=g)
    mask = None
    if with_mask:
        thr = 0.85 if name == "mask_few_valid" else 0.5
        mask = (torch.rand(n_cam, H, W, generator=g) > thr).float()
    return c2w, focal, image, mask, n_rays, normalize, opengl, mask_rate, seed


def make_ray_cases():
    """get_rays_multicam of the unmodified reference (ray_utils.py:16-87) on tiny cameras, with the CPU RNG seeded;
    the value drawn AFTER the call pins the generator consumption."""
    ns = load_reference()
    out = {}
    for name in RAY_CASES:
        c2w, focal, image, mask, n_rays, normalize, opengl, mask_rate, seed = ray_case_inputs(name)
        torch.manual_seed(seed)
        ro, rd, rgb, msel = ns.ray_utils.get_rays_multicam(c2w=c2w, focal=focal, image=image, n_rays=n_rays, normalize=normalize,
                                                            mask=mask, mask_rate=mask_rate, return_mask=mask is not None, opengl=opengl)
        after = torch.rand(4)
        out[name + "_rays_o"], out[name + "_rays_d"], out[name + "_rgb"] = ro.numpy(), rd.numpy(), rgb.numpy()
        if msel is not None:
            out[name + "_mask_sel"] = msel.numpy()
        out[name + "_rng_after"] = after.numpy()
    np.savez_compressed(os.path.join(HERE, "rays_multicam.npz"), **out)
    print("rays_multicam.npz:", sorted(out))


if __name__ == "__main__":
    torch.set_num_threads(8)
    names = sys.argv[1:] or (list(CASES) + ["rays"])
    for n in names:
        if n == "rays":
            make_ray_cases()
        else:
            make_case(n)
