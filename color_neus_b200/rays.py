"""Caller-side ray helpers needed to drive the renderer (synthetic cameras for bench/tests).

Same arithmetic and ray ordering (index = y*W + x) as lib/models/tools/ray_utils.py:7-13 (near_far_from_sphere),
:90-119 (get_rays_at) and lib/utils/transform.py:322-337 (pose_spherical); torch ops on whatever device the
camera tensors live on.

SURVEY.md section 8f "next" #1 -- on-device ray generation + selection: `get_rays_multicam` / `get_rays_selected` below are
drop-ins for the reference's `get_rays_multicam` that draw the ray indices with the reference's own CPU RNG call
sequence (bit-identical selection) but generate ONLY the selected rays (`cneus_gen_rays`, csrc/raygen.cu) instead of
materialising all N*H*W rays of every camera each step.
"""
import math

import torch


def pose_spherical(theta_deg, phi_deg, radius):
    th, ph = theta_deg / 180.0 * math.pi, phi_deg / 180.0 * math.pi
    t = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=torch.float32)
    rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0], [0, math.sin(ph), math.cos(ph), 0],
                       [0, 0, 0, 1]], dtype=torch.float32)
    rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0],
                       [0, 0, 0, 1]], dtype=torch.float32)
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ (rt @ (rp @ t))


def get_rays_at(c2w, focal, H, W, normalize=False, opengl=False):
    """All rays of one pinhole camera -> (rays_o [H,W,3], rays_d [H,W,3])."""
    assert c2w.dim() == 2, "single camera"
    device = c2w.device
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='xy')
    i, j = i.to(device), j.to(device)
    ys, zs = (-1, -1) if opengl else (1, 1)
    dirs = torch.stack([(i - 0.5 * W) / focal[0], ys * (j - 0.5 * H) / focal[1], zs * torch.ones_like(i)], -1)
    if normalize:
        dirs = dirs / torch.norm(dirs, dim=-1).unsqueeze(-1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def near_far_from_sphere(rays_o, rays_d):
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return (mid - 1.0).squeeze(-1), (mid + 1.0).squeeze(-1)


def synthetic_camera_rays(H, W, theta_deg=30.0, phi_deg=-30.0, radius=2.8, focal_mul=1.2, device="cpu"):
    """SURVEY.md section 8d synthetic camera: focal = 1.2*W, camera on a sphere looking at the origin,
    NORMALIZE_DIR=True.  Returns flat (rays_o, rays_d, near, far) in y*W+x order."""
    c2w = pose_spherical(theta_deg, phi_deg, radius).to(device)
    focal = torch.tensor([focal_mul * W, focal_mul * W], dtype=torch.float32, device=device)
    o, d = get_rays_at(c2w, focal, H, W, normalize=True)
    o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
    near, far = near_far_from_sphere(o, d)
    return o, d, near.contiguous(), far.contiguous()


# ---------------------------------------------------------------------------------------------------------------
# selected-ray generation
# ---------------------------------------------------------------------------------------------------------------
def select_ray_indices(n_cam, H, W, n_rays, mask=None, mask_rate=0.9):
    """Flat pixel indices ((cam*H + y)*W + x) of one training batch, drawn exactly like the reference
    (ray_utils.py:57-75): same CPU-generator calls in the same order, including the no-mask quirk that indices are
    drawn in [0, H*W), i.e. from the first camera only.  `mask` may live on any device; the draws stay on the CPU."""
    if mask is None:
        return torch.randint(0, H * W, (n_rays,))
    mask_all = mask.reshape(-1)
    dev = mask_all.device
    valid = torch.where(mask_all > 0)[0]
    rand_valid = torch.randperm(valid.shape[0])
    n_in = int(mask_rate * n_rays)
    if n_in > valid.shape[0]:
        n_in = valid.shape[0]
    n_bkg = n_rays - n_in
    invalid = torch.where(mask_all == 0)[0]
    rand_invalid = torch.randperm(invalid.shape[0])
    idx = torch.cat([valid[rand_valid[:n_in].to(dev)], invalid[rand_invalid[:n_bkg].to(dev)]], dim=-1)
    return idx[torch.randperm(idx.shape[0]).to(dev)]


def _selected_rays_torch(c2w, focal, H, W, idx, normalize, opengl):
    """Differentiable torch path (pose / focal networks being trained): the reference's expressions on the selected pixels."""
    hw = H * W
    cam = torch.div(idx, hw, rounding_mode="floor")
    pix = idx - cam * hw
    j = torch.div(pix, W, rounding_mode="floor").to(torch.float32)
    i = (pix - torch.div(pix, W, rounding_mode="floor") * W).to(torch.float32)
    ys, zs = (-1, -1) if opengl else (1, 1)
    dirs = torch.stack([(i - W * 0.5) / focal[0], ys * (j - H * 0.5) / focal[1], zs * torch.ones_like(i)], -1)
    if normalize:
        dirs = dirs / torch.norm(dirs, dim=-1).unsqueeze(-1)
    R = c2w[cam]                                            # [n, 4, 4]
    rays_d = torch.sum(dirs[:, None, :] * R[:, :3, :3], -1)
    rays_o = R[:, :3, -1]
    return rays_o, rays_d


def get_rays_selected(c2w, focal, H, W, idx, normalize=False, opengl=False, origin=None, radius=None, with_near_far=False,
                      image=None, mask=None):
    """Rays of the flat pixel indices `idx` ([n] int64) of cameras `c2w` [N,4,4] (or [4,4]).

    Returns (rays_o, rays_d, near, far, rgb, mask_sel); near/far are None unless `with_near_far`, rgb / mask_sel are
    None unless `image` [N,H,W,3] / `mask` [N,H,W] are given.  On CUDA tensors without autograd this is one
    `cneus_gen_rays` launch; when the cameras require grad (pose / focal networks) the same expressions run as
    differentiable torch ops on the selected pixels only."""
    if c2w.dim() == 2:
        c2w = c2w[None]
    needs_grad = torch.is_grad_enabled() and (c2w.requires_grad or (torch.is_tensor(focal) and focal.requires_grad))
    if c2w.is_cuda and not needs_grad:
        import ctypes as C
        from . import _lib as L
        lib = L.lib()
        dev = c2w.device
        n = idx.numel()
        f32 = dict(dtype=torch.float32, device=dev)
        c2w_c = c2w.detach().to(torch.float32).contiguous()
        focal_c = torch.as_tensor(focal, **f32).detach().reshape(-1)[:2].contiguous()
        idx_c = idx.to(device=dev, dtype=torch.int64).contiguous()
        rays_o, rays_d = torch.empty(n, 3, **f32), torch.empty(n, 3, **f32)
        near = torch.empty(n, **f32) if with_near_far else None
        far = torch.empty(n, **f32) if with_near_far else None
        org = torch.as_tensor(origin, **f32).detach().reshape(-1)[:3].contiguous() if origin is not None else None
        rad = torch.as_tensor(radius, **f32).detach().reshape(-1)[:1].contiguous() if radius is not None else None
        # images / masks may live on the host (the reference's datasets keep them there) or on another GPU: everything the
        # kernel reads is brought to the cameras' device first, and L.ptr validates device / dtype / contiguity
        img = image.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3).contiguous() if image is not None else None
        msk = mask.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous() if mask is not None else None
        rgb = torch.empty(n, 3, **f32) if img is not None else None
        msel = torch.empty(n, **f32) if msk is not None else None
        p = L.ptr
        with torch.cuda.device(dev):
            L.check(lib.cneus_gen_rays(p(c2w_c), c2w_c.shape[0], p(focal_c), H, W, C.c_void_p(idx_c.data_ptr()), 0, n,
                                       int(normalize), int(opengl), p(org), p(rad), p(img), p(msk), p(rays_o), p(rays_d),
                                       p(near), p(far), p(rgb), p(msel), L.stream_ptr()), "cneus_gen_rays")
        return rays_o, rays_d, near, far, rgb, msel
    idx = idx.to(c2w.device)
    rays_o, rays_d = _selected_rays_torch(c2w, focal, H, W, idx, normalize, opengl)
    if origin is not None:
        rays_o = (rays_o - torch.as_tensor(origin, device=rays_o.device)).float()
    if radius is not None:
        rays_o = (rays_o / torch.as_tensor(radius, device=rays_o.device)).float()
    near = far = None
    if with_near_far:
        near, far = near_far_from_sphere(rays_o, rays_d)
    rgb = image.reshape(-1, 3)[idx] if image is not None else None
    msel = mask.reshape(-1)[idx] if mask is not None else None
    return rays_o, rays_d, near, far, rgb, msel


def get_rays_multicam(c2w, focal, image, n_rays, normalize=False, mask=None, mask_rate=0.9, return_mask=False, opengl=False):
    """Drop-in for lib/models/tools/ray_utils.py:16-87 (same signature, same return tuple, same CPU RNG consumption and
    therefore the same selected pixels), without building the rays of every pixel of every camera."""
    assert c2w.dim() == 3 and image.dim() == 4, "this is a multicam implementation"
    N, H, W = c2w.shape[0], image.shape[1], image.shape[2]
    idx = select_ray_indices(N, H, W, n_rays, mask=mask, mask_rate=mask_rate)
    if return_mask:
        assert mask is not None
    rays_o, rays_d, _, _, rgb, msel = get_rays_selected(c2w, focal, H, W, idx, normalize=normalize, opengl=opengl, image=image,
                                                        mask=mask if return_mask else None)
    return rays_o, rays_d, rgb, (msel if return_mask else None)


# ---------------------------------------------------------------------------------------------------------------
# dataset residency
# ---------------------------------------------------------------------------------------------------------------
class ResidentImageSet:
    """All training images (and masks) of a scene resident on the device as the uint8 they are stored as.

    The reference keeps float32 images on the host (`get_all_img`, lib/datasets/dtu.py:143-160: 23 MB per 1200x1600 image)
    and copies BATCH_SIZE full images to the device every step (`get_rand_batch_smaples`, dtu.py:166-177: 184 MB of H2D for
    8 images) to read N_RAYS = 1024 pixels of them.  Here the uint8 data (a quarter of the size) is uploaded once and the
    selected pixels are converted by `cneus_gather_pixels_u8` with get_image's float pipeline bit for bit
    (dtu.py:98-113), so a step moves 12 KB instead.  RNG contract: `get_rand_batch` makes the same
    `torch.randperm(n_imgs)` CPU draw as the reference, `sample_rays` the same draws as `get_rays_multicam`."""

    def __init__(self, images_u8, masks_u8=None, img_ids=None, std=0.5, premultiply_mask=True, device="cuda"):
        images_u8 = torch.as_tensor(images_u8)
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise ValueError("images_u8 must be uint8 [N, H, W, 3] (RGB)")
        self.images = images_u8.contiguous().to(device)
        self.masks = None
        if masks_u8 is not None:
            masks_u8 = torch.as_tensor(masks_u8)
            if masks_u8.dtype != torch.uint8 or tuple(masks_u8.shape) != tuple(images_u8.shape[:3]):
                raise ValueError("masks_u8 must be uint8 [N, H, W]")
            self.masks = masks_u8.contiguous().to(device)
        self.n_imgs, self.H, self.W = (int(x) for x in images_u8.shape[:3])
        self.img_ids = torch.arange(self.n_imgs) if img_ids is None else torch.as_tensor(img_ids)
        self.std, self.premultiply_mask = float(std), bool(premultiply_mask)

    def get_rand_batch(self, batch_size):
        """dtu.py:166-177 without the copies: the batch is the list of image indices."""
        use_index = torch.randperm(self.n_imgs)[:batch_size]
        return {"use_index": use_index, "img_ids": self.img_ids[use_index]}

    def batch_masks(self, use_index):
        """uint8 masks [B,H,W] of the batch (device gather); `> 0` / `== 0` select the same pixels as on mask / 255."""
        return None if self.masks is None else self.masks[use_index.to(self.masks.device)]

    def gather(self, use_index, idx, return_mask=False):
        """(rgb [n,3], mask [n] or None) of the flat batch indices `idx` ((cam*H + y)*W + x, cam = position in use_index)."""
        from . import _lib as L
        import ctypes as C
        dev = self.images.device
        if not self.images.is_cuda:
            raise L.CneusError("ResidentImageSet.gather needs the images on a CUDA device (no CPU path)")
        lib = L.lib()
        if not hasattr(lib, "_gp_bound"):
            vp = C.c_void_p
            lib.cneus_gather_pixels_u8.restype = C.c_int
            lib.cneus_gather_pixels_u8.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_int32, vp, vp, vp]
            lib._gp_bound = True
        idx_c = idx.to(device=dev, dtype=torch.int64).contiguous()
        cam_map = use_index.to(device=dev, dtype=torch.int64).contiguous()
        n = idx_c.numel()
        rgb = torch.empty(n, 3, dtype=torch.float32, device=dev)
        want_mask = return_mask and self.masks is not None
        msel = torch.empty(n, dtype=torch.float32, device=dev) if want_mask else None
        if self.masks is not None and (self.masks.device != dev or not self.masks.is_contiguous()):
            raise L.CneusError("ResidentImageSet: images and masks must be contiguous tensors on the same CUDA device")

        def p(t):   # uint8 / int64 / float32 tensors created above on `dev`
            if t is None:
                return None
            if not (t.is_cuda and t.device == dev and t.is_contiguous()):
                raise L.CneusError("ResidentImageSet.gather: expected a contiguous tensor on " + str(dev))
            return C.c_void_p(t.data_ptr())

        with torch.cuda.device(dev):
            L.check(lib.cneus_gather_pixels_u8(p(self.images), p(self.masks), p(cam_map), p(idx_c), n, self.H, self.W, self.std,
                                               int(self.premultiply_mask), p(rgb), p(msel),
                                               L.stream_ptr()), "cneus_gather_pixels_u8")
        return rgb, msel

    def sample_rays(self, c2w, focal, batch, n_rays, normalize=False, use_mask=True, mask_rate=0.9, return_mask=False, opengl=False,
                    origin=None, radius=None, with_near_far=False):
        """`get_rays_multicam` (ray_utils.py:16-87) on a resident batch: same index draws, rays of the selected pixels only
        (`cneus_gen_rays`), pixels gathered from the uint8 store.  -> (rays_o, rays_d, near, far, rgb, mask_sel)."""
        use_index = batch["use_index"]
        mask = self.batch_masks(use_index) if use_mask else None
        idx = select_ray_indices(len(use_index), self.H, self.W, n_rays, mask=mask, mask_rate=mask_rate)
        rays_o, rays_d, near, far, _, _ = get_rays_selected(c2w, focal, self.H, self.W, idx, normalize=normalize, opengl=opengl,
                                                            origin=origin, radius=radius, with_near_far=with_near_far)
        rgb, msel = self.gather(use_index, idx, return_mask=return_mask)
        return rays_o, rays_d, near, far, rgb, msel
