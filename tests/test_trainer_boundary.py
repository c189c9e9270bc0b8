"""Trainer-level proof of the drop-in boundary (north_star: "lib/models and train.py call it unchanged").

The reference's own `NeuS_Trainer` (lib/models/NeuS_Trainer.py, imported UNMODIFIED from /root/reference or its staged copy
oracle/_ref through oracle/ref_import.load_trainer, with inert stand-ins only for packages that are not installed:
matplotlib / kornia / imageio / trimesh / pytorch3d) is built twice from the same shipped-config-shaped node:
once on the stock `RENDERER` registry (reference renderer, CPU) and once after `color_neus_b200.register()` (sm_100a
renderer, GPU).  Same seed, same synthetic batch ->
  * `training_step` (NeuS_Trainer.py:173-214: pose / focal nets -> get_rays_multicam -> renderer -> compute_loss): same
    selected rays, same loss terms, same CPU generator state afterwards, gradients of `loss.backward()` within the bar;
  * `validation_step` -> `validate_image` (:216-277, eval mode, autograd on, EVAL_RAY_SIZE chunks): same uint8 image;
  * `testing_step` -> `validate_mesh` (:279-307) runs on the drop-in (the reference needs PyMCubes for it).
"""
import numpy as np
import pytest
import torch

from helpers import O, record, rel_err
from oracle import ref_import as R

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="no reference tree (/root/reference or oracle/_ref)")

H = W = 16
N_IMG = 3


def trainer_cfg(n_rays=48):
    ren = O.default_cfg("Color_NeuS", 64, 64, 256, 8, 0.3)
    model = dict(TYPE="NeuS_Trainer", PRETRAINED=None, N_RAYS=n_rays, EVAL_RAY_SIZE=128, NORMALIZE_DIR=True, FOCAL_ORDER=2,
                 LEARN_FOCAL=False, LEARN_R=False, LEARN_T=False, MASK_RATE=[0.5, 0.8], POSE_MODE="6d", RENDERER=ren,
                 LOSS=dict(RGB_LOSS_TYPE="mse", LAMBDA_FINE=1.0, LAMBDA_EIKONAL=0.1, LAMBDA_MASK=0.1, LAMBDA_RELIGHT=1.0),
                 DATA_PRESET=dict(FX_ONLY=False, INCLUDE_MASK=True, OPENGL_SYS=False),
                 TRAIN=dict(ITERATIONS=1000, LOG_INTERVAL=10, VIZ_IMAGE_INTERVAL=1, VIZ_MESH_INTERVAL=10 ** 9))
    return R.CfgDict(model)


def synthetic_scene():
    g = torch.Generator().manual_seed(4)
    poses = torch.stack([O.pose_spherical(25.0 + 40.0 * k, -30.0 + 4.0 * k, 2.7) for k in range(N_IMG)])
    data = dict(origin=torch.zeros(3), radius=torch.ones(1), focal=torch.tensor([5.0 * W, 5.0 * W]), poses=poses, n_imgs=N_IMG, H=H, W=W,
                scale_mats_np=[np.eye(4, dtype=np.float32)], object_bbox_min=np.array([-0.4, -0.4, -0.4], dtype=np.float32),
                object_bbox_max=np.array([0.4, 0.4, 0.4], dtype=np.float32))
    images = torch.rand(N_IMG, 3, H, W, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    disc = (((yy - H / 2 + 0.5) ** 2 + (xx - W / 2 + 0.5) ** 2) < (0.33 * H) ** 2).float()
    masks = disc[None].repeat(N_IMG, 1, 1)
    batch = dict(images=images * masks[:, None], masks=masks, img_ids=torch.arange(N_IMG))
    return data, batch


def build_pair():
    import color_neus_b200 as cn
    t = R.load_trainer()
    registry = t.builder.RENDERER
    stock = {k: registry.get(k) for k in ("NeuS", "Color_NeuS")}
    data, batch = synthetic_scene()
    try:
        torch.manual_seed(1)
        ref = t.NeuS_Trainer(trainer_cfg(), data)
        assert type(ref.renderer) is t.Color_NeuS
        cn.register(registry)
        torch.manual_seed(1)
        ours = t.NeuS_Trainer(trainer_cfg(), data)      # the reference's trainer class, unmodified, on the drop-in renderer
        assert type(ours.renderer) is cn.Color_NeuS
    finally:
        for k, v in stock.items():
            registry.register_module(name=k, force=True, module=v)
    return t, ref, ours, batch


def test_trainer_builds_on_the_drop_in_with_identical_initial_state():
    """CPU part: same seed -> the trainer built on the drop-in holds bit-identical parameters under the same names."""
    t, ref, ours, _ = build_pair()
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert list(sd_r.keys()) == list(sd_o.keys())
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_o[k]), k
    ours.load_state_dict(sd_r, strict=True)


@pytest.mark.gpu
def test_training_step_validate_image_and_mesh_through_the_reference_trainer(tmp_path, monkeypatch):
    from test_gpu_backward import BACKWARD_BAR
    monkeypatch.chdir(tmp_path)   # validate_image / validate_mesh create ./tmp/NeuS_Trainer/... relative to the cwd
    t, ref, ours, batch = build_pair()
    ours = ours.cuda()
    cu_batch = {k: v.cuda() for k, v in batch.items()}
    # ---- training_step (train.py:62-70): model.train(), forward, loss.backward()
    ref.train(); ours.train()
    torch.manual_seed(11)
    rd_r, ld_r = ref(batch, 1, "train")
    rng_r = torch.get_rng_state()
    torch.manual_seed(11)
    rd_o, ld_o = ours(cu_batch, 1, "train")
    assert torch.equal(rng_r, torch.get_rng_state()), "the drop-in consumed the CPU generator differently"
    assert torch.equal(rd_r["rgb_map_gt"], rd_o["rgb_map_gt"].cpu()) and torch.equal(rd_r["mask"], rd_o["mask"].cpu())   # same rays
    assert list(rd_r.keys()) == [k for k in rd_o.keys() if k not in ("eikonal_num", "eikonal_den")]
    for k in ("color_fine", "weight_sum", "depth", "global_color"):
        assert record("trainer", "training_step", k, rel_err(rd_o[k].detach().cpu(), rd_r[k].detach())) < 1e-4, k
    assert set(ld_r) == set(ld_o)
    for k in ld_r:
        a, b = float(ld_o[k]), float(ld_r[k])
        assert record("trainer", "training_step", "loss_" + k, abs(a - b) / max(abs(b), 1e-6)) < 2e-4, (k, a, b)
    ld_r["loss"].backward()
    ld_o["loss"].backward()
    worst = 0.0
    for (k, p_r), (_, p_o) in zip(ref.named_parameters(), ours.named_parameters()):
        if p_r.grad is None:
            assert p_o.grad is None or float(p_o.grad.abs().max()) == 0.0, k
            continue
        gr, go = p_r.grad.double(), p_o.grad.detach().cpu().double()
        e = float((go.norm() - gr.norm()).abs() / (gr.norm() + 1e-12)) if float(gr.norm()) > 1e-9 else float(go.abs().max())
        worst = max(worst, e)
        assert e < BACKWARD_BAR, (k, e)
    record("trainer", "training_step", "worst_grad_norm_err", worst)
    # ---- validation_step -> validate_image (train.py:90-93: model.eval(), autograd left on)
    ref.eval(); ours.eval()
    torch.manual_seed(5)
    ref(batch, 0, "val")
    _, img_r = t.imageio.last_written
    torch.manual_seed(5)
    ours(cu_batch, 0, "val")
    _, img_o = t.imageio.last_written
    assert img_r.shape == img_o.shape == (H, 3 * W, 3) and img_r.dtype == np.uint8
    assert np.array_equal(img_r[:, :W], img_o[:, :W])                                           # ground-truth panel
    d_rgb = np.abs(img_r[:, W:2 * W].astype(int) - img_o[:, W:2 * W].astype(int)).max()          # rendered panel
    record("trainer", "validate_image", "max_uint8_diff_rgb", d_rgb)
    assert d_rgb <= 1
    assert abs(ref.PSNR.get_result() - ours.PSNR.get_result()) < 1e-3
    # ---- testing_step -> validate_mesh on the drop-in (extract_geometry + extract_color; the reference needs PyMCubes here)
    t.trimesh.Trimesh.exported.clear()
    ours(None, 7, "test", recon_res=48)
    (p_mesh, mesh), (p_col, mesh_col) = t.trimesh.Trimesh.exported
    assert p_mesh.endswith("00000007_mesh.ply") and p_col.endswith("00000007_color.ply")
    v, f, c = np.asarray(mesh_col.vertices), np.asarray(mesh_col.faces), np.asarray(mesh_col.vertex_colors)
    assert v.shape[0] > 100 and f.shape[1] == 3 and c.shape == (v.shape[0], 3) and np.isfinite(c).all()
    r = np.linalg.norm(v, axis=1)                      # geometric init: a sphere of radius ~ BIAS / SCALE = 1/6
    assert abs(float(r.mean()) - 1.0 / 6.0) < 0.03
