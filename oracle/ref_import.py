"""Import the UNMODIFIED reference hot-path modules from /root/reference (container only).

TEST INFRASTRUCTURE -- never imported by the product path (color_neus_b200/).

The reference tree pulls in packages that are not installed here (termcolor,
yacs, mcubes, pytorch3d); none of them does arithmetic on the hot path, so we
register inert stand-ins *before* importing `lib.models.renderers.*`, and we
pre-register empty package shells for `lib`, `lib.models`, ... so that
`lib/models/__init__.py` (which drags in trimesh/kornia/imageio through the
trainer) is bypassed.  See SURVEY.md section 8c.

Used by `tests/golden/make_golden.py` (fixture generation), by
`tests/test_oracle_vs_reference.py` and by the CPU legs of `bench.py`
(`--impl reference`, `cpu_baseline`).

The reference is pure Python, so there is nothing to compile; instead
`stage_reference()` (called by `__graft_entry__.build()` in the build container)
copies the reference's `lib/` package VERBATIM from /root/reference into
`oracle/_ref/lib/` -- git-ignored (never part of the history), not
gpurun-ignored, so like the built `.so` it travels to the GPU box, where
/root/reference does not exist.  Resolution order of the tree that gets
imported: $CNEUS_REFERENCE_ROOT, /root/reference, oracle/_ref.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")
SOURCE_ROOT = "/root/reference"


def _has_tree(root):
    return bool(root) and os.path.isfile(os.path.join(root, "lib", "models", "renderers", "NeuS.py"))


def _resolve_root():
    for cand in (os.environ.get("CNEUS_REFERENCE_ROOT"), SOURCE_ROOT, STAGED_ROOT):
        if _has_tree(cand):
            return cand
    return SOURCE_ROOT


REF_ROOT = _resolve_root()


def reference_available() -> bool:
    return _has_tree(REF_ROOT)


def reference_kind() -> str:
    """'source' = /root/reference itself, 'staged' = the verbatim copy under oracle/_ref, 'absent'."""
    if not reference_available():
        return "absent"
    return "staged" if os.path.abspath(REF_ROOT) == os.path.abspath(STAGED_ROOT) else "source"


def stage_reference(force=False) -> bool:
    """Copy the reference's `lib/**/*.py` unmodified into oracle/_ref/lib (build container only; a no-op where
    /root/reference is absent).  Returns True when oracle/_ref holds a usable tree afterwards."""
    import filecmp
    import shutil
    if not _has_tree(SOURCE_ROOT):
        return _has_tree(STAGED_ROOT)
    src_lib = os.path.join(SOURCE_ROOT, "lib")
    for d, _, files in os.walk(src_lib):
        rel = os.path.relpath(d, SOURCE_ROOT)
        for f in files:
            if not f.endswith(".py"):
                continue
            dst_dir = os.path.join(STAGED_ROOT, rel)
            os.makedirs(dst_dir, exist_ok=True)
            src, dst = os.path.join(d, f), os.path.join(dst_dir, f)
            if force or not os.path.isfile(dst) or not filecmp.cmp(src, dst, shallow=False):
                shutil.copyfile(src, dst)
    with open(os.path.join(STAGED_ROOT, "README"), "w") as fh:
        fh.write("Verbatim copy of /root/reference/lib/**/*.py made by oracle/ref_import.py:stage_reference().\n"
                 "Git-ignored build artefact (test infrastructure / CPU baseline only); do not edit.\n")
    return _has_tree(STAGED_ROOT)


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
            # only the two conversions Pose_Net calls get a working stand-in (caller side of the boundary); the rest is inert
            setattr(pt, name, {"axis_angle_to_matrix": _axis_angle_to_matrix,
                               "rotation_6d_to_matrix": _rotation_6d_to_matrix}.get(name))
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


# ---------------------------------------------------------------------------------------------------------------------
# trainer-level import (lib/models/NeuS_Trainer.py): the caller side of the drop-in boundary, run UNMODIFIED
# ---------------------------------------------------------------------------------------------------------------------
def _axis_angle_to_matrix(aa):
    """Rodrigues' formula (stand-in for pytorch3d.transforms.axis_angle_to_matrix, which is not installed)."""
    import torch
    theta = aa.norm(dim=-1, keepdim=True).clamp_min(1e-12)[..., None]
    k = aa / theta[..., 0]
    K = torch.zeros(aa.shape[:-1] + (3, 3), dtype=aa.dtype, device=aa.device)
    K[..., 0, 1], K[..., 0, 2], K[..., 1, 0] = -k[..., 2], k[..., 1], k[..., 2]
    K[..., 1, 2], K[..., 2, 0], K[..., 2, 1] = -k[..., 0], -k[..., 1], k[..., 0]
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device).expand_as(K)
    return eye + torch.sin(theta) * K + (1.0 - torch.cos(theta)) * (K @ K)


def _rotation_6d_to_matrix(d6):
    """Gram-Schmidt of the two 3-vectors (Zhou et al. 2019; stand-in for pytorch3d.transforms.rotation_6d_to_matrix)."""
    import torch
    import torch.nn.functional as F
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def _install_trainer_stubs():
    """Inert or minimal stand-ins for the packages NeuS_Trainer.py imports that are not installed here (matplotlib, kornia,
    imageio, trimesh) and working rotation conversions for Pose_Net.  None of them is on the renderer's side of the
    boundary.  `imageio.imwrite` keeps the last image in `imageio.last_written` so a test can look at what validate_image
    produced without touching the file system."""
    _install_stubs()

    def module(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = module("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = module("matplotlib.pyplot")
        tk = module("mpl_toolkits")
        tk.mplot3d = module("mpl_toolkits.mplot3d", Axes3D=object)
    try:
        import kornia  # noqa: F401
    except ImportError:
        import torch
        k = module("kornia")
        k.metrics = module("kornia.metrics", ssim=lambda a, b, w: torch.full((1,), float("nan")))
    try:
        import imageio  # noqa: F401
    except ImportError:
        io = module("imageio", last_written=None)

        def imwrite(path, image, *a, **k):
            io.last_written = (path, image)
        io.imwrite = imwrite
    try:
        import trimesh  # noqa: F401
    except ImportError:
        class Trimesh:
            def __init__(self, vertices=None, faces=None, vertex_colors=None, **kw):
                self.vertices, self.faces, self.vertex_colors = vertices, faces, vertex_colors

            def export(self, path, *a, **k):
                Trimesh.exported.append((path, self))
        Trimesh.exported = []
        module("trimesh", Trimesh=Trimesh)


def load_trainer():
    """-> namespace(NeuS_Trainer, trainer_mod, builder, net_utils, ...) with the reference's trainer module imported unmodified."""
    if "trainer" in _CACHE:
        return _CACHE["trainer"]
    ns = load_reference()
    _install_trainer_stubs()
    _shell("lib.metrics", os.path.join(REF_ROOT, "lib", "metrics"))
    importlib.import_module("lib.metrics.basic_metric")
    sys.modules["lib.metrics"].LossMetric = sys.modules["lib.metrics.basic_metric"].LossMetric
    sys.modules["lib.metrics"].Metric = sys.modules["lib.metrics.basic_metric"].Metric
    renderers = sys.modules["lib.models.renderers"]
    if not hasattr(renderers, "build_renderer"):   # run the package's own __init__.py (build_renderer, :4-5) inside the shell
        init = os.path.join(REF_ROOT, "lib", "models", "renderers", "__init__.py")
        exec(compile(open(init).read(), init, "exec"), renderers.__dict__)
    trainer_mod = importlib.import_module("lib.models.NeuS_Trainer")
    import logging
    importlib.import_module("lib.utils.logger").logger.setLevel(logging.ERROR)
    out = types.SimpleNamespace(NeuS_Trainer=trainer_mod.NeuS_Trainer, trainer_mod=trainer_mod, builder=ns.builder,
                                Color_NeuS=ns.Color_NeuS, NeuS=ns.NeuS, CfgDict=CfgDict, imageio=sys.modules["imageio"],
                                trimesh=sys.modules["trimesh"])
    _CACHE["trainer"] = out
    return out
