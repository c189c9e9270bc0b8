"""color_neus_b200 -- B200-native (sm_100a) implementation of the Color-NeuS volume-rendering hot path.

Public surface = the reference's own renderer/field classes (same names, signatures, state_dict keys):
    NeuS, Color_NeuS                      (lib/models/renderers/NeuS.py, Color_NeuS.py)
    SDFNetwork, RenderingNetwork, RelightNetwork, SingleVarianceNetwork   (lib/models/renderers/fields.py)
plus `register()` to install them in the reference's RENDERER registry.  All arithmetic of the hot path runs in
hand-written CUDA kernels behind the C ABI of include/cneus.h (libcneus.so); there is no CPU fallback.
"""
from .embedder import Embedder, get_embedder  # noqa: F401
from .fields import RelightNetwork, RenderingNetwork, SDFNetwork, SingleVarianceNetwork  # noqa: F401
from .renderer import Color_NeuS, NeuS, register  # noqa: F401

ColorNetwork = RenderingNetwork  # BASELINE.json's name for the colour MLP

__all__ = ["NeuS", "Color_NeuS", "SDFNetwork", "RenderingNetwork", "ColorNetwork", "RelightNetwork",
           "SingleVarianceNetwork", "register", "get_embedder", "Embedder"]
