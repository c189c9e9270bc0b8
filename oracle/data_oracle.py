"""TEST INFRASTRUCTURE -- CPU restatement (numpy fp32, one IEEE operation per step) of the pixel pipeline of the reference
datasets' get_image.  Never imported by the product path.  Pinned by tests/golden/pixel_lut.npz,
which tests/golden/make_golden_data.py produces with torchvision's to_tensor / normalize exactly as
lib/datasets/dtu.py:98-113 calls them."""
import numpy as np


def image_value(u8, std=0.5):
    """dtu.py:104-107: to_tensor (u8 / 255), normalize(mean 0.5, std), * 0.5 + 0.5 -- float32 at every step."""
    x = np.asarray(u8, np.uint8).astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(std)
    return x * np.float32(0.5) + np.float32(0.5)


def mask_value(m8):
    """dtu.py:110-111: to_tensor of the uint8 grey mask."""
    return np.asarray(m8, np.uint8).astype(np.float32) / np.float32(255.0)


def gather_pixels(images_u8, masks_u8, use_index, idx, std=0.5, premultiply=True):
    """Pixels (and mask values) of flat batch indices idx = (cam * H + y) * W + x, cam = position in use_index."""
    images_u8, use_index, idx = np.asarray(images_u8), np.asarray(use_index), np.asarray(idx)
    n_img, H, W, _ = images_u8.shape
    cam, pix = idx // (H * W), idx % (H * W)
    src = use_index[cam] * (H * W) + pix
    rgb = image_value(images_u8.reshape(-1, 3)[src], std)
    m = None
    if masks_u8 is not None:
        m = mask_value(np.asarray(masks_u8).reshape(-1)[src])
        if premultiply:
            rgb = rgb * m[:, None]            # dtu.py:113
    return rgb.astype(np.float32), m
