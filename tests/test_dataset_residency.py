"""uint8-resident training images + checkpoint-format compatibility.

CPU: the pixel-pipeline oracle against the torchvision-generated table (tests/golden/pixel_lut.npz, every uint8 value), the
drop-in modules' state_dict / optimizer / scheduler layout against what the UNMODIFIED reference produces
(tests/golden/checkpoint_layout.json), a `NeuS_Trainer.pth.tar`-style round trip through torch.save.
GPU: cneus_gather_pixels_u8 bit-exact against the oracle (exhaustive value table + random batches), the resident sampler
against the float path of get_rays_multicam on the same RNG stream."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

import __graft_entry__ as g
from helpers import HERE, O
from oracle import data_oracle as D

LUT = np.load(os.path.join(HERE, "golden", "pixel_lut.npz"))
LAYOUT = json.load(open(os.path.join(HERE, "golden", "checkpoint_layout.json")))


def test_oracle_pixel_pipeline_is_the_reference_one():
    v = np.arange(256, dtype=np.uint8)
    assert np.array_equal(D.image_value(v, 0.5), LUT["img_std0.5"])          # bit-exact, all 256 values
    assert np.array_equal(D.image_value(v, 0.25), LUT["img_std0.25"])
    assert np.array_equal(D.mask_value(v), LUT["mask"])
    m = ((v.astype(np.int32) * 37) % 256).astype(np.uint8)
    assert np.array_equal(D.image_value(v, 0.5) * D.mask_value(m), LUT["premul_ch0"])


def _random_set(n_img=5, H=12, W=10, seed=0):
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, (n_img, H, W, 3), dtype=np.uint8)
    masks = (rng.random((n_img, H, W)) > 0.5).astype(np.uint8) * 255
    masks[0, 0, :4] = [1, 127, 128, 254]      # non-binary mask values take the same path
    return images, masks


@pytest.mark.parametrize("kind", ["Color_NeuS", "NeuS"])
def test_state_dict_layout_equals_reference(kind):
    import color_neus_b200 as cn
    ren = getattr(cn, kind)(g._Cfg(O.default_cfg(kind)))
    sd = ren.state_dict()
    want = LAYOUT[kind]
    assert list(sd.keys()) == list(sorted(want.keys(), key=list(sd.keys()).index)) and set(sd.keys()) == set(want.keys())
    for k, (shape, dtype) in want.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k


def test_trainer_checkpoint_round_trip(tmp_path):
    """io_utils.save_states / net_utils.init_weights (lib/utils/io_utils.py:44-56, net_utils.py:252-283): the trainer's
    state_dict (renderer under the `renderer.` prefix) saved with torch.save, loaded back strictly -- plain OrderedDict,
    {"state_dict": ...} and DataParallel's `module.` prefix."""
    import color_neus_b200 as cn

    class Trainer(torch.nn.Module):          # stands in for NeuS_Trainer: the renderer is the attribute `renderer`
        def __init__(self):
            super().__init__()
            self.renderer = cn.Color_NeuS(g._Cfg(O.default_cfg("Color_NeuS", 64, 0, 128, 4, 0.3)))

    torch.manual_seed(3)
    a, b = Trainer(), Trainer()
    path = os.path.join(tmp_path, "NeuS_Trainer.pth.tar")
    torch.save(a.state_dict(), path)
    ck = torch.load(path, map_location="cpu")
    assert isinstance(ck, OrderedDict) and all(k.startswith("renderer.") for k in ck)
    b.load_state_dict(ck, strict=True)
    wrapped = {"state_dict": OrderedDict(("module." + k, v) for k, v in ck.items())}
    sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in wrapped["state_dict"].items())   # net_utils.py:266-275
    b.load_state_dict(sd, strict=True)
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k


def test_optimizer_and_scheduler_state_layout_equals_reference():
    from color_neus_b200 import train_ops as TR
    ws = [torch.nn.Parameter(torch.zeros(3)) for _ in range(LAYOUT["optimizer"]["n_params"])]

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    opt, sch = TR.build_optimizer_nerf(torch.nn.ParameterList(ws), Cfg(TYPE="adam", LR=5e-4, SCHEDULER_TYPE="NEUS", WARM_UP=5000,
                                                                     LR_ALPHA=0.05), -1, iterations=300000)
    osd = opt.state_dict()
    assert sorted(osd["param_groups"][0].keys()) == LAYOUT["optimizer"]["param_group_keys"]
    assert len(osd["param_groups"][0]["params"]) == LAYOUT["optimizer"]["n_params"]
    assert sorted(sch.state_dict().keys()) == LAYOUT["scheduler"]["keys"]
    # a reference-format optimizer state (torch.optim.Adam after one step) loads into the fused optimizer
    ref = torch.optim.Adam(ws, lr=5e-4, betas=(0.9, 0.99))
    for p in ws:
        p.grad = torch.ones_like(p)
    ref.step()
    rsd = ref.state_dict()
    assert sorted(rsd["state"][0].keys()) == LAYOUT["optimizer"]["state_keys"]
    rsd["param_groups"][0]["initial_lr"] = 5e-4
    opt.load_state_dict(rsd)
    assert float(opt.state_dict()["state"][0]["step"]) == 1.0


def test_resident_set_validates_and_has_no_cpu_path():
    from color_neus_b200._lib import CneusError
    from color_neus_b200.rays import ResidentImageSet
    images, masks = _random_set()
    with pytest.raises(ValueError):
        ResidentImageSet(images.astype(np.float32), masks, device="cpu")
    with pytest.raises(ValueError):
        ResidentImageSet(images, masks[:, :-1], device="cpu")
    rs = ResidentImageSet(images, masks, device="cpu")
    torch.manual_seed(4)
    batch = rs.get_rand_batch(3)
    torch.manual_seed(4)
    assert torch.equal(batch["use_index"], torch.randperm(5)[:3])      # the reference's draw (dtu.py:168-169)
    with pytest.raises(CneusError):
        rs.gather(batch["use_index"], torch.tensor([0, 1]))


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("std,premul", [(0.5, True), (0.5, False), (0.25, True)])
def test_gpu_gather_every_value_bit_exact(std, premul):
    from color_neus_b200.rays import ResidentImageSet
    v = np.arange(256, dtype=np.uint8)
    m = ((v.astype(np.int32) * 37) % 256).astype(np.uint8)
    images = np.stack([v, v[::-1], m], -1).reshape(1, 16, 16, 3)
    rs = ResidentImageSet(images, m.reshape(1, 16, 16), std=std, premultiply_mask=premul)
    rgb, msel = rs.gather(torch.tensor([0]), torch.arange(256), return_mask=True)
    want, wm = D.gather_pixels(images, m.reshape(1, 16, 16), [0], np.arange(256), std, premul)
    assert np.array_equal(rgb.cpu().numpy(), want) and np.array_equal(msel.cpu().numpy(), wm)
    if std == 0.5 and premul:
        assert np.array_equal(rgb.cpu().numpy()[:, 0], LUT["premul_ch0"])      # the torchvision-generated table itself


@pytest.mark.gpu
def test_gpu_resident_sampler_equals_float_path():
    """Same RNG stream -> same pixels, rays and colours as get_rays_multicam on the float images the reference would upload."""
    from color_neus_b200 import rays as R
    images, masks = _random_set(6, 20, 24, seed=2)
    rs = R.ResidentImageSet(images, masks)
    c2w = torch.stack([R.pose_spherical(20.0 * i, -30.0, 2.7) for i in range(6)]).cuda()
    focal = torch.tensor([30.0, 30.0]).cuda()
    torch.manual_seed(11)
    batch = rs.get_rand_batch(4)
    ro, rd, near, far, rgb, msel = rs.sample_rays(c2w[batch["use_index"].cuda()], focal, batch, 64, normalize=True, mask_rate=0.9,
                                                  return_mask=True, with_near_far=True)
    after = torch.rand(3)
    # the reference's data path on the same seed: float images / masks of the batch, get_rays_multicam
    torch.manual_seed(11)
    use = torch.randperm(6)[:4]
    img_f = torch.as_tensor(D.image_value(images[use.numpy()], 0.5) * D.mask_value(masks[use.numpy()])[..., None]).cuda()
    msk_f = torch.as_tensor(D.mask_value(masks[use.numpy()])).cuda()
    ro2, rd2, rgb2, msel2 = R.get_rays_multicam(c2w[use.cuda()], focal, img_f, 64, normalize=True, mask=msk_f, mask_rate=0.9,
                                                return_mask=True)
    assert torch.equal(after, torch.rand(3))
    assert torch.equal(rgb, rgb2) and torch.equal(msel, msel2) and torch.equal(ro, ro2) and torch.equal(rd, rd2)
    n2, f2 = R.near_far_from_sphere(ro2, rd2)
    assert torch.allclose(near, n2, rtol=1e-6, atol=1e-6) and torch.allclose(far, f2, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_gpu_resident_no_mask_dataset():
    from color_neus_b200 import rays as R
    images, _ = _random_set(3, 8, 8, seed=5)
    rs = R.ResidentImageSet(images, None, premultiply_mask=False)
    idx = torch.randint(0, 2 * 64, (50,))
    rgb, msel = rs.gather(torch.tensor([2, 0]), idx)
    want, _ = D.gather_pixels(images, None, [2, 0], idx.numpy(), 0.5, False)
    assert msel is None and np.array_equal(rgb.cpu().numpy(), want)
