"""Import the UNMODIFIED reference hot-path modules from /root/reference (container only).

TEST INFRASTRUCTURE -- never imported by the product path (color_neus_b200/).

The reference tree pulls in packages that are not installed here (termcolor,
yacs, mcubes, pytorch3d); none of them does arithmetic on the hot path, so we
register inert stand-ins *before* importing `lib.models.renderers.*`, and we
pre-register empty package shells for `lib`, `lib.models`, ... so that
`lib/models/__init__.py` (which drags in trimesh/kornia/imageio through the
trainer) is bypassed.  See SURVEY.md section 8c.

Used by `tests/golden/make_golden.py` (fixture generation) and by
`tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent,
e.g. on the GPU box).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("CNEUS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "lib", "models", "renderers", "NeuS.py"))


class CfgDict(dict):
    """dict with attribute access standing in for yacs.config.CfgNode."""

    def __init__(self, init=None, new_allowed=True, **kw):
        super().__init__()
        init = dict(init or {})
        init.update(kw)
        for k, v in init.items():
            self[k] = CfgDict(v) if isinstance(v, dict) and not isinstance(v, CfgDict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return CfgDict({k: (v.clone() if isinstance(v, CfgDict) else v) for k, v in self.items()})

    def defrost(self):
        return None

    def freeze(self):
        return None

    def set_new_allowed(self, flag):
        return None

    def merge_from_other_cfg(self, other):
        for k, v in other.items():
            self[k] = v

    def merge_from_file(self, f):
        raise NotImplementedError

    def dump(self, *a, **k):
        return repr(dict(self))


def _install_stubs():
    if "termcolor" not in sys.modules:
        m = types.ModuleType("termcolor")
        m.colored = lambda s, *a, **k: s
        m.cprint = lambda *a, **k: None
        sys.modules["termcolor"] = m
    if "yacs" not in sys.modules:
        y = types.ModuleType("yacs")
        yc = types.ModuleType("yacs.config")
        yc.CfgNode = CfgDict
        y.config = yc
        sys.modules["yacs"] = y
        sys.modules["yacs.config"] = yc
    if "mcubes" not in sys.modules:
        m = types.ModuleType("mcubes")

        def _no_mc(*a, **k):
            raise RuntimeError("PyMCubes is not installed (third-party, out of scope)")

        m.marching_cubes = _no_mc
        sys.modules["mcubes"] = m
    if "pytorch3d" not in sys.modules:
        p = types.ModuleType("pytorch3d")
        pt = types.ModuleType("pytorch3d.transforms")
        for name in ("axis_angle_to_matrix", "axis_angle_to_quaternion", "euler_angles_to_matrix",
                     "matrix_to_euler_angles", "matrix_to_quaternion", "matrix_to_rotation_6d",
                     "quaternion_to_axis_angle", "quaternion_to_matrix", "rotation_6d_to_matrix"):
            setattr(pt, name, None)
        p.transforms = pt
        sys.modules["pytorch3d"] = p
        sys.modules["pytorch3d.transforms"] = pt


def _shell(name, path):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


_CACHE = {}


def load_reference():
    """Returns a namespace with the reference classes/functions of the hot path."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    _install_stubs()
    lib = os.path.join(REF_ROOT, "lib")
    _shell("lib", lib)
    _shell("lib.utils", os.path.join(lib, "utils"))
    _shell("lib.models", os.path.join(lib, "models"))
    _shell("lib.models.tools", os.path.join(lib, "models", "tools"))
    _shell("lib.models.renderers", os.path.join(lib, "models", "renderers"))
    import logging
    fields = importlib.import_module("lib.models.renderers.fields")
    neus = importlib.import_module("lib.models.renderers.NeuS")
    cneus = importlib.import_module("lib.models.renderers.Color_NeuS")
    ray_utils = importlib.import_module("lib.models.tools.ray_utils")
    pe = importlib.import_module("lib.models.tools.PositionEncoding")
    builder = importlib.import_module("lib.utils.builder")
    logging.getLogger().setLevel(logging.WARNING)
    try:
        importlib.import_module("lib.utils.logger").logger.setLevel(logging.ERROR)
    except Exception:
        pass
    ns = types.SimpleNamespace(fields=fields, NeuS=neus.NeuS, Color_NeuS=cneus.Color_NeuS, neus_mod=neus,
                               ray_utils=ray_utils, pe=pe, builder=builder, CfgDict=CfgDict)
    _CACHE["ns"] = ns
    return ns
