"""The four field modules of the hot path with the reference's constructor/forward signatures and
state_dict keys, evaluated by the sm_100a kernels of libcneus.so.

Mirrors (file:line in the reference tree):
  SDFNetwork            lib/models/renderers/fields.py:12-115
  RenderingNetwork      lib/models/renderers/fields.py:119-188   (BASELINE.json calls it ColorNetwork)
  SingleVarianceNetwork lib/models/renderers/fields.py:277-286
  RelightNetwork        lib/models/renderers/fields.py:289-368
Parameters live in ordinary nn.Parameters named exactly like the reference's (`lin{l}.weight_g`,
`lin{l}.weight_v`, `lin{l}.bias`, `in_layer.weight`, `rl_mlp.{i}.weight`, `variance`), so checkpoints load
with strict=True in both directions.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .net import NetHandle


class WNLinear(nn.Module):
    """Parameter container equivalent to nn.utils.weight_norm(nn.Linear(in, out)) (legacy API, dim=0):
    effective weight = weight_g * weight_v / ||weight_v||_row.  The matmul itself runs inside the CUDA kernels.

    Built from an already initialised nn.Linear exactly like the legacy hook does it (fields.py:72-73 ->
    torch.nn.utils.weight_norm): g = norm_except_dim(weight, 2, 0), v = weight; registration order bias, weight_g,
    weight_v, so `state_dict()` has the reference's key order too."""

    def __init__(self, linear):
        super().__init__()
        self.in_features, self.out_features = linear.in_features, linear.out_features
        w = linear.weight.detach()
        self.bias = nn.Parameter(linear.bias.detach().clone())
        self.weight_g = nn.Parameter(torch.norm_except_dim(w, 2, 0).detach().clone())
        self.weight_v = nn.Parameter(w.clone())

    def effective_weight(self):
        return self.weight_v * (self.weight_g / self.weight_v.norm(2, dim=1, keepdim=True))


def _geometric_init_(lin, l, n_lin, in0, multires, skip_in, bias, inside_outside):
    """Sphere initialisation of SDF layer `l` (fields.py:52-70), applied to a freshly constructed nn.Linear with the same
    torch.nn.init calls in the same order on the same (possibly strided) views, so that the CPU generator is consumed draw
    for draw like in the reference: `torch.manual_seed(s); SDFNetwork(cfg)` gives bit-identical parameters."""
    init = torch.nn.init
    out_dim, in_dim = lin.weight.shape
    std = math.sqrt(2) / math.sqrt(out_dim)
    if l == n_lin - 1:  # last layer: mean +-sqrt(pi)/sqrt(in), bias -+BIAS -> sdf ~ |x| - BIAS/SCALE
        sign = -1.0 if inside_outside else 1.0
        init.normal_(lin.weight, mean=sign * math.sqrt(math.pi) / math.sqrt(in_dim), std=0.0001)
        init.constant_(lin.bias, -sign * bias)
        return
    init.constant_(lin.bias, 0.0)
    if multires > 0 and l == 0:  # only the raw coordinates feed the first layer
        init.constant_(lin.weight[:, 3:], 0.0)
        init.normal_(lin.weight[:, :3], 0.0, std)
    elif multires > 0 and l in skip_in:  # the re-injected encoding's sin / cos columns start at zero
        init.normal_(lin.weight, 0.0, std)
        init.constant_(lin.weight[:, -(in0 - 3):], 0.0)
    else:
        init.normal_(lin.weight, 0.0, std)


def _as_f32c(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class _NoBackward(torch.autograd.Function):
    """Marks the result of a forward-only entry point: the value is what the kernels computed, and a backward pass that
    reaches it raises instead of silently delivering zero gradients (the reference's stand-alone sub-module calls are
    differentiable, fields.py:105-115 even with create_graph=True; here only `NeuS.forward` / `Color_NeuS.forward` in
    training mode carry an analytic backward, color_neus_b200/autograd.py)."""

    @staticmethod
    def forward(ctx, what, n_out, *outs_and_deps):
        ctx.what = what
        return tuple(o.view_as(o) for o in outs_and_deps[:n_out])

    @staticmethod
    def backward(ctx, *grads):
        raise L.CneusError(f"{ctx.what} is forward-only in color_neus_b200: there is no backward through this entry point "
                           "(differentiate through the renderer's forward() in training mode instead)")


def guard_no_backward(what, outs, deps):
    """outs: tensor or tuple of tensors computed by a forward-only kernel; deps: the tensors / parameters a caller could
    expect gradients for.  No-op unless autograd is recording and one of them requires grad."""
    if not torch.is_grad_enabled():
        return outs
    deps = [d for d in deps if torch.is_tensor(d) and d.requires_grad]
    if not deps:
        return outs
    if torch.is_tensor(outs):
        return _NoBackward.apply(what, 1, outs, *deps)[0]
    if isinstance(outs, dict):
        keys = [k for k, v in outs.items() if torch.is_tensor(v) and v.is_floating_point()]
        res = _NoBackward.apply(what, len(keys), *[outs[k] for k in keys], *deps)
        return {**outs, **dict(zip(keys, res))}
    return tuple(_NoBackward.apply(what, len(outs), *outs, *deps))


class SDFNetwork(nn.Module):

    def __init__(self, cfg):
        super().__init__()
        self.name = type(self).__name__
        self.cfg = cfg
        d_in = cfg.get('D_IN', 3)
        self.d_out = cfg.get('D_OUT', 257)
        self.d_hidden = cfg.get('D_HIDDEN', 256)
        n_layers = cfg.get('N_LAYERS', 8)
        self.skip_in = list(cfg.get('SKIP_IN', [4]))
        self.multires = cfg.get('MULTIRES', 6)
        bias = cfg.get('BIAS', 0.5)
        self.scale = cfg.get('SCALE', 3.0)
        geometric_init = cfg.get('GEOMETRIC_INIT', True)
        weight_norm = cfg.get('WEIGHT_NORM', True)
        inside_outside = cfg.get('INSIDE_OUTSIDE', False)
        if d_in != 3:
            raise L.CneusError("SDFNetwork: D_IN must be 3")
        in0 = d_in * (1 + 2 * self.multires) if self.multires > 0 else d_in
        dims = [in0] + [self.d_hidden] * n_layers + [self.d_out]
        self.num_layers = len(dims)
        if any(k >= self.num_layers - 1 for k in self.skip_in):
            raise L.CneusError("SDFNetwork: SKIP_IN beyond the last linear layer is not supported")
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)  # consumes the generator like the reference's constructor (fields.py:50)
            if geometric_init:
                with torch.no_grad():
                    _geometric_init_(lin, l, self.num_layers - 1, dims[0], self.multires, self.skip_in, bias, inside_outside)
            setattr(self, "lin" + str(l), WNLinear(lin) if weight_norm else lin)
        self._handle = None

    def handle(self):
        if self._handle is None:
            self._handle = NetHandle(self)
        return self._handle

    def _pts(self, x):
        dev = next(self.parameters()).device
        x = _as_f32c(x, dev).reshape(-1, 3)
        return x

    def forward(self, inputs):
        """[P,3] -> [P, D_OUT]; column 0 is the SDF (already divided by SCALE) -- fields.py:81-97."""
        h, x = self.handle(), self._pts(inputs)
        out = torch.empty(x.shape[0], self.d_out, device=x.device, dtype=torch.float32)
        ws, wsb = h.workspace(n_points=x.shape[0])
        with torch.cuda.device(x.device):
            L.check(L.lib().cneus_sdf_forward(h.dref(), h.packed(), L.ptr(x), x.shape[0], L.ptr(out), self.d_out, ws, wsb,
                                              L.stream_ptr()), "cneus_sdf_forward")
        return guard_no_backward("SDFNetwork.forward", out, [inputs, *self.parameters()])

    def sdf(self, x):
        """fields.py:99-100 -- only column 0 is computed."""
        x_in = x
        h, x = self.handle(), self._pts(x)
        out = torch.empty(x.shape[0], 1, device=x.device, dtype=torch.float32)
        ws, wsb = h.workspace(n_points=x.shape[0])
        with torch.cuda.device(x.device):
            L.check(L.lib().cneus_sdf_forward(h.dref(), h.packed(), L.ptr(x), x.shape[0], L.ptr(out), 1, ws, wsb,
                                              L.stream_ptr()), "cneus_sdf_forward")
        return guard_no_backward("SDFNetwork.sdf", out, [x_in, *self.parameters()])

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def gradient(self, x):
        """d sdf / d x as [P,1,3] -- fields.py:105-115 (closed-form reverse chain instead of autograd)."""
        x_in = x
        h, x = self.handle(), self._pts(x)
        out = torch.empty(x.shape[0], 3, device=x.device, dtype=torch.float32)
        ws, wsb = h.workspace(n_points=x.shape[0])
        with torch.cuda.device(x.device):
            L.check(L.lib().cneus_sdf_gradient(h.dref(), h.packed(), L.ptr(x), x.shape[0], L.ptr(out), ws, wsb,
                                               L.stream_ptr()), "cneus_sdf_gradient")
        return guard_no_backward("SDFNetwork.gradient", out.unsqueeze(1), [x_in, *self.parameters()])


class RenderingNetwork(nn.Module):

    def __init__(self, cfg):
        super().__init__()
        self.name = type(self).__name__
        self.cfg = cfg
        self.d_feature = cfg.get('D_FEATURE', 256)
        self.mode = cfg.get('MODE', 'idr')
        d_in = cfg.get('D_IN', 9)
        self.d_out = cfg.get('D_OUT', 3)
        self.d_hidden = cfg.get('D_HIDDEN', 256)
        n_layers = cfg.get('N_LAYERS', 4)
        weight_norm = cfg.get('WEIGHT_NORM', True)
        self.multires_view = cfg.get('MULTIRES_VIEW', 4)
        self.squeeze_out = cfg.get('SQUEEZE_OUT', True)
        if self.mode not in L.COLOR_MODES:
            raise ValueError(f'no such mode: {self.mode}')
        dims = [d_in + self.d_feature] + [self.d_hidden] * n_layers + [self.d_out]
        if self.multires_view > 0:
            dims[0] += 3 * (1 + 2 * self.multires_view) - 3
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])  # PyTorch's default init, as in fields.py:150
            setattr(self, "lin" + str(l), WNLinear(lin) if weight_norm else lin)
        self._handle = None

    def handle(self):
        if self._handle is None:
            self._handle = _color_only_handle(self)
        return self._handle

    def forward(self, points, normals, view_dirs, feature_vectors):
        """fields.py:161-188; `normals` is the raw (un-normalised) SDF gradient."""
        dev = next(self.parameters()).device
        p = _as_f32c(points, dev).reshape(-1, 3)
        n = _as_f32c(normals, dev).reshape(-1, 3) if normals is not None else None
        v = _as_f32c(view_dirs, dev).reshape(-1, 3) if view_dirs is not None else None
        f = _as_f32c(feature_vectors, dev).reshape(-1, self.d_feature)
        h = self.handle()
        out = torch.empty(p.shape[0], 3, device=dev, dtype=torch.float32)
        ws, wsb = h.workspace(n_points=p.shape[0])
        with torch.cuda.device(dev):
            L.check(L.lib().cneus_color_forward(h.dref(), h.packed(), L.ptr(p), L.ptr(n), L.ptr(v), L.ptr(f), p.shape[0],
                                                L.ptr(out), ws, wsb, L.stream_ptr()), "cneus_color_forward")
        return guard_no_backward("RenderingNetwork.forward", out,
                                 [points, normals, view_dirs, feature_vectors, *self.parameters()])


class SingleVarianceNetwork(nn.Module):

    def __init__(self, cfg):
        super().__init__()
        self.register_parameter('variance', nn.Parameter(torch.tensor(float(cfg.get('INIT_VAL', 0.3)))))

    def forward(self, x):
        """ones([len(x),1]) * exp(10 * variance) -- fields.py:284-286 (one scalar; plain torch)."""
        return torch.ones([len(x), 1], device=x.device) * torch.exp(self.variance * 10.0)


class RelightNetwork(nn.Module):

    def __init__(self, cfg):
        super().__init__()
        self.name = type(self).__name__
        self.cfg = cfg
        d_in = cfg.get('D_IN', 6)
        d_out = cfg.get('D_OUT', 3)
        self.d_hidden = cfg.get('D_HIDDEN', 256)
        self.n_layers = cfg.get('N_LAYERS', 4)
        self.y_in_layer = cfg.get('Y_IN_LAYER', 3)
        self.multires_view = cfg.get('MULTIRES_VIEW', 4)
        self.include_grad = cfg.get('INCLUDE_GRAD', True)
        self.inv_sigmoid = cfg.get('INV_SIGMOID', True)
        if d_out != 3 or d_in != 6:
            raise L.CneusError("RelightNetwork: D_IN must be 6 and D_OUT 3")
        if self.include_grad:
            d_in += 3
        if self.multires_view > 0:
            d_in += 3 * (1 + 2 * self.multires_view) - 3
        self.in_layer = nn.Linear(d_in, self.d_hidden)
        self.rl_mlp = nn.ModuleList()
        for i in range(self.n_layers):
            k_in = self.d_hidden + (3 if i == self.y_in_layer - 1 else 0)
            k_out = d_out if i == self.n_layers - 1 else self.d_hidden
            self.rl_mlp.append(nn.Linear(k_in, k_out))
        self._handle = None

    def handle(self):
        if self._handle is None:
            self._handle = _relight_only_handle(self)
        return self._handle

    def relight(self, rgb, pts, dirs, gradients):
        dev = next(self.parameters()).device
        c = _as_f32c(rgb, dev).reshape(-1, 3)
        p = _as_f32c(pts, dev).reshape(-1, 3)
        d = _as_f32c(dirs, dev).reshape(-1, 3)
        g = _as_f32c(gradients, dev).reshape(-1, 3) if gradients is not None else None
        h = self.handle()
        out = torch.empty(p.shape[0], 3, device=dev, dtype=torch.float32)
        drgb = torch.empty(p.shape[0], 3, device=dev, dtype=torch.float32)
        ws, wsb = h.workspace(n_points=p.shape[0])
        with torch.cuda.device(dev):
            L.check(L.lib().cneus_relight_forward(h.dref(), h.packed(), L.ptr(c), L.ptr(p), L.ptr(d), L.ptr(g), p.shape[0],
                                                  L.ptr(out), L.ptr(drgb), ws, wsb, L.stream_ptr()),
                    "cneus_relight_forward")
        return guard_no_backward("RelightNetwork.forward", (out, drgb), [rgb, pts, dirs, gradients, *self.parameters()])

    def forward(self, rgb, pts, dirs, gradients):
        """fields.py:361-368 -> (relit rgb, delta rgb)."""
        return self.relight(rgb, pts, dirs, gradients)


class _SdfShape(nn.Module):
    """Minimal SDF stand-in so that a colour- or relight-only NetHandle has a consistent descriptor: one tiny
    hidden layer feeding a D_OUT-wide output (never evaluated by the stand-alone entry points)."""

    def __init__(self, d_out, device):
        super().__init__()
        self.num_layers, self.d_hidden, self.d_out, self.multires, self.scale, self.skip_in = 3, 64, d_out, 0, 1.0, []
        self.lin0 = nn.Linear(3, 64).to(device)
        self.lin1 = nn.Linear(64, d_out).to(device)


def _color_only_handle(color):
    stub = _SdfShape(color.d_feature + 1, next(color.parameters()).device)  # kept out of color's state_dict
    return NetHandle(stub, color, None, primary=color)


def _relight_only_handle(rel):
    stub = _SdfShape(17, next(rel.parameters()).device)
    return NetHandle(stub, None, rel, primary=rel)
