"""SURVEY.md section 8f #1 -- ray generation + selection: the oracle restatement and the drop-in get_rays_multicam /
get_rays_selected against golden outputs of the unmodified reference (tests/golden/rays_multicam.npz, made by
tests/golden/make_golden.py rays): same selected pixels (CPU RNG consumed identically), rays within 2 ulp-ish
(1e-6 relative); GPU: the cneus_gen_rays kernel against the same fixtures and against full-image get_rays_at."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import ROOT, O  # noqa: F401

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as MG  # noqa: E402
from color_neus_b200 import rays as R  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "rays_multicam.npz"))


def close(a, b, tol=1e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("name", list(MG.RAY_CASES))
def test_oracle_matches_reference_golden(name):
    c2w, focal, image, mask, n_rays, normalize, opengl, mask_rate, seed = MG.ray_case_inputs(name)
    torch.manual_seed(seed)
    ro, rd, rgb, msel, _ = O.get_rays_multicam(c2w, focal, image, n_rays, normalize, mask, mask_rate, mask is not None, opengl)
    after = torch.rand(4)
    assert np.array_equal(after.numpy(), GOLD[name + "_rng_after"])           # generator consumed identically
    assert np.array_equal(rgb.numpy(), GOLD[name + "_rgb"])                   # same pixels selected
    assert np.array_equal(ro.numpy(), GOLD[name + "_rays_o"]) and close(rd.numpy(), GOLD[name + "_rays_d"])
    if mask is not None:
        assert np.array_equal(msel.numpy(), GOLD[name + "_mask_sel"])


@pytest.mark.parametrize("name", list(MG.RAY_CASES))
def test_dropin_cpu_path_matches_reference_golden(name):
    """Torch path of the drop-in (what runs when the cameras require grad): only the selected pixels are generated."""
    c2w, focal, image, mask, n_rays, normalize, opengl, mask_rate, seed = MG.ray_case_inputs(name)
    torch.manual_seed(seed)
    ro, rd, rgb, msel = R.get_rays_multicam(c2w, focal, image, n_rays, normalize=normalize, mask=mask, mask_rate=mask_rate,
                                            return_mask=mask is not None, opengl=opengl)
    assert np.array_equal(torch.rand(4).numpy(), GOLD[name + "_rng_after"])
    assert np.array_equal(rgb.numpy(), GOLD[name + "_rgb"])
    assert close(ro.numpy(), GOLD[name + "_rays_o"]) and close(rd.numpy(), GOLD[name + "_rays_d"])
    if mask is not None:
        assert np.array_equal(msel.numpy(), GOLD[name + "_mask_sel"])


def test_dropin_is_differentiable_wrt_cameras():
    c2w, focal, image, mask, n_rays, normalize, opengl, mask_rate, seed = MG.ray_case_inputs("mask")
    c2w = c2w.clone().requires_grad_(True)
    focal = focal.clone().requires_grad_(True)
    ro, rd, _, _ = R.get_rays_multicam(c2w, focal, image, n_rays, normalize=normalize, mask=mask, mask_rate=mask_rate, opengl=opengl)
    (ro.sum() + (rd ** 2).sum()).backward()
    assert c2w.grad is not None and focal.grad is not None and float(c2w.grad.abs().sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MG.RAY_CASES))
def test_kernel_matches_reference_golden(name):
    c2w, focal, image, mask, n_rays, normalize, opengl, mask_rate, seed = MG.ray_case_inputs(name)
    torch.manual_seed(seed)
    ro, rd, rgb, msel = R.get_rays_multicam(c2w.cuda(), focal.cuda(), image.cuda(), n_rays, normalize=normalize,
                                            mask=mask.cuda() if mask is not None else None, mask_rate=mask_rate,
                                            return_mask=mask is not None, opengl=opengl)
    assert np.array_equal(torch.rand(4).numpy(), GOLD[name + "_rng_after"])
    assert np.array_equal(rgb.cpu().numpy(), GOLD[name + "_rgb"])
    assert close(ro.cpu().numpy(), GOLD[name + "_rays_o"]) and close(rd.cpu().numpy(), GOLD[name + "_rays_d"])
    if mask is not None:
        assert np.array_equal(msel.cpu().numpy(), GOLD[name + "_mask_sel"])


@pytest.mark.gpu
def test_kernel_full_image_order_and_near_far():
    """index == all pixels: ray t is pixel (y, x) = (t // W, t % W) of get_rays_at; fused normalisation + near/far."""
    H, W = 37, 53
    c2w = R.pose_spherical(25.0, -20.0, 3.1).cuda()
    focal = torch.tensor([1.2 * W, 1.25 * W]).cuda()
    o_ref, d_ref = R.get_rays_at(c2w, focal, H, W, normalize=True)
    origin, radius = torch.tensor([0.1, -0.05, 0.02]).cuda(), torch.tensor([1.3]).cuda()
    o_ref = ((o_ref.reshape(-1, 3) - origin) / radius).float()
    n_ref, f_ref = R.near_far_from_sphere(o_ref, d_ref.reshape(-1, 3))
    idx = torch.arange(H * W)
    ro, rd, near, far, _, _ = R.get_rays_selected(c2w, focal, H, W, idx, normalize=True, origin=origin, radius=radius,
                                                  with_near_far=True)
    assert close(ro.cpu(), o_ref.cpu()) and close(rd.cpu(), d_ref.reshape(-1, 3).cpu())
    assert close(near.cpu(), n_ref.cpu(), 1e-5) and close(far.cpu(), f_ref.cpu(), 1e-5)
    sub = torch.tensor([5 * W + 7, 0, H * W - 1, 11 * W])
    ro2, rd2, _, _, _, _ = R.get_rays_selected(c2w, focal, H, W, sub, normalize=True, origin=origin, radius=radius)
    assert torch.equal(rd2, rd[sub.cuda()]) and torch.equal(ro2, ro[sub.cuda()])
