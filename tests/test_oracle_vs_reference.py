"""Run the unmodified reference (imported from /root/reference) next to the oracle on fresh seeded inputs.
Skipped where the reference tree is absent (e.g. the GPU box); the golden fixtures cover that case."""
import numpy as np
import pytest
import torch

from helpers import O, MG, rel_err
from oracle.ref_import import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("kind,n_s,n_i,var", [("Color_NeuS", 64, 64, 0.3), ("NeuS", 32, 32, 0.45)])
def test_forward_matches_reference(kind, n_s, n_i, var):
    cfg = O.default_cfg(kind, n_s, n_i, 256, 8, var)
    Pn = O.make_params(cfg, seed=5, trained_like=True)
    ren = MG.build_reference(cfg, Pn)
    ro, rd, near, far = MG.synth_rays(24, seed=3)
    torch.manual_seed(11)
    t_rand = torch.rand([24, 1])
    torch.manual_seed(11)
    ref = ren(ro, rd, near, far)
    got = O.render_forward(O.to_torch(Pn), cfg, ro, rd, near, far, t_rand=t_rand)
    for k in ("color_fine", "weight_sum", "depth"):
        assert rel_err(got[k], ref[k].detach()) < 1e-4, k


def test_rays_and_near_far_match_reference():
    ns = MG.load_reference()
    c2w = O.pose_spherical(40.0, -30.0, 2.7)
    f = torch.tensor([1.2 * 40, 1.2 * 40])
    o1, d1 = ns.ray_utils.get_rays_at(c2w, f, 30, 40, normalize=True)
    o2, d2 = O.get_rays_at(c2w, f, 30, 40, normalize=True)
    assert np.array_equal(o1.reshape(-1, 3).numpy(), o2.numpy())
    assert np.array_equal(d1.reshape(-1, 3).numpy(), d2.numpy())
    n1, f1 = ns.ray_utils.near_far_from_sphere(o2, d2)
    n2, f2 = O.near_far_from_sphere(o2, d2)
    assert np.array_equal(n1.numpy(), n2.numpy()) and np.array_equal(f1.numpy(), f2.numpy())
