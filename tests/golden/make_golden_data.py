"""Golden data for SURVEY 8f #4 (build container only):   python tests/golden/make_golden_data.py

* pixel_lut.npz -- every uint8 value through the float pipeline of the reference datasets' get_image, executed with the very
  calls lib/datasets/dtu.py:104-113 makes (torchvision to_tensor / normalize, * 0.5 + 0.5, to_tensor of the mask, product)
  for std 0.5 (every shipped dataset) and 0.25.
* checkpoint_layout.json -- names / shapes / dtypes of the state_dict of the UNMODIFIED reference renderers (what
  `NeuS_Trainer.pth.tar` holds under the `renderer.` prefix, lib/utils/io_utils.py:44-56) and the structure of the
  optimizer / scheduler entries of `train_param.pth.tar` (recorder.py:89-99) as the reference's own
  build_optimizer_nerf produces them.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torchvision.transforms.functional as tvF

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import neus_oracle as O  # noqa: E402
from oracle.ref_import import CfgDict, load_reference  # noqa: E402


def pixel_lut():
    out = {}
    vals = np.arange(256, dtype=np.uint8)
    img = np.stack([vals, vals[::-1], ((vals.astype(np.int32) * 7) % 256).astype(np.uint8)], -1).reshape(16, 16, 3).astype(np.uint8)   # RGB, H=W=16
    for std in (0.5, 0.25):
        image = tvF.to_tensor(img)                                        # dtu.py:104
        image = tvF.normalize(image, [0.5, 0.5, 0.5], [std, std, std])    # dtu.py:106
        image = image * 0.5 + 0.5                                         # dtu.py:107
        out[f"img_std{std}"] = image[0].reshape(-1).numpy()               # channel 0 holds value v at position v
    mask = tvF.to_tensor(vals.reshape(16, 16)).squeeze()                  # dtu.py:110-111
    out["mask"] = mask.reshape(-1).numpy()
    image = tvF.to_tensor(img)
    image = tvF.normalize(image, [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]) * 0.5 + 0.5
    m2 = tvF.to_tensor(((vals.astype(np.int32) * 37) % 256).astype(np.uint8).reshape(16, 16)).squeeze()
    out["premul_ch0"] = (image * m2.unsqueeze(0))[0].reshape(-1).numpy()  # dtu.py:113 with mask value (37 v) % 256
    np.savez_compressed(os.path.join(HERE, "pixel_lut.npz"), **out)


def checkpoint_layout():
    import make_golden_train as MT
    ns = load_reference()
    MT._accept_verbose()
    net_utils = importlib.import_module("lib.utils.net_utils")
    layout = {}
    for kind in ("Color_NeuS", "NeuS"):
        cfg = CfgDict(O.default_cfg(kind))
        torch.manual_seed(1)
        ren = getattr(ns, kind)(cfg)
        layout[kind] = {k: [list(v.shape), str(v.dtype)] for k, v in ren.state_dict().items()}
        if kind == "Color_NeuS":
            ocfg = CfgDict(TYPE="adam", LR=5e-4, SCHEDULER_TYPE="NEUS", WARM_UP=5000, LR_ALPHA=0.05)
            opt, sch = net_utils.build_optimizer_nerf(ren, ocfg, -1, iterations=300000)
            for p in ren.parameters():
                p.grad = torch.zeros_like(p)
            opt.step()
            sch.step()
            osd = opt.state_dict()
            layout["optimizer"] = {"param_group_keys": sorted(osd["param_groups"][0].keys()),
                                   "n_params": len(osd["param_groups"][0]["params"]),
                                   "state_keys": sorted(osd["state"][0].keys()),
                                   "step_dtype": str(osd["state"][0]["step"].dtype)}
            layout["scheduler"] = {"keys": sorted(sch.state_dict().keys())}
    with open(os.path.join(HERE, "checkpoint_layout.json"), "w") as f:
        json.dump(layout, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    pixel_lut()
    checkpoint_layout()
    print("wrote pixel_lut.npz, checkpoint_layout.json")
