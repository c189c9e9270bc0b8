"""Training path of the drop-in renderer: forward through the sm_100a kernels, backward through
`cneus_render_backward` (the explicit double-backward program of csrc/backward.cu) wrapped in a
torch.autograd.Function so that `loss.backward()` in the reference's train.py:70 works unchanged.

The Function's differentiable inputs are rays_o, rays_d, the variance and the EFFECTIVE weights / biases of every
layer; the effective weights are built with differentiable torch ops from (weight_g, weight_v), so autograd itself
finishes the weight-norm chain (fields.py:72-73) from dL/dW_eff.  z_vals are constants (NeuS.py:343-355).
"""
import ctypes as C

import torch

from . import _lib as L

_BWD_IN_FIELDS = ["rays_o", "rays_d", "z", "mid_z", "dists", "sdf", "gradients", "sampled_color", "global_sampled", "alpha",
                  "weights", "variance", "eikonal_den", "g_color_fine", "g_global_color", "g_weight_sum", "g_weight_max",
                  "g_depth", "g_weights", "g_cdf", "g_gradients", "g_delta_relight", "g_gradient_error", "g_s_val_sum"]


class BackwardIn(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _BWD_IN_FIELDS]


class LinearGrad(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p)]


class ParamGrads(C.Structure):
    _fields_ = [("sdf", LinearGrad * L.MAX_SDF_LIN), ("color", LinearGrad * L.MAX_COLOR_LIN), ("relight_in", LinearGrad),
                ("relight_mlp", LinearGrad * L.MAX_RELIGHT_LIN), ("variance", C.c_void_p)]


_bound = False


def _bind():
    global _bound
    lib = L.lib()
    if not _bound:
        lib.cneus_backward_workspace_bytes.restype = C.c_size_t
        lib.cneus_backward_workspace_bytes.argtypes = [C.POINTER(L.NetDesc), C.c_int64, C.c_int32]
        lib.cneus_render_backward.restype = C.c_int
        lib.cneus_render_backward.argtypes = [C.POINTER(L.NetDesc), C.POINTER(L.Params), C.POINTER(BackwardIn), C.c_int64, C.c_int32,
                                              C.c_float, C.POINTER(ParamGrads), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                              C.c_void_p]
        _bound = True
    return lib


def effective_linears(handle):
    """[(kind, idx, W_eff, bias)] with W_eff differentiable w.r.t. the module parameters."""
    out = []
    for kind, idx, m in handle._linears():
        if hasattr(m, "weight_g"):
            w = m.weight_v * (m.weight_g / m.weight_v.norm(2, dim=1, keepdim=True))
        else:
            w = m.weight
        out.append((kind, idx, w, m.bias))
    return out


DIFF_KEYS = ["color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max", "gradients", "weights", "gradient_error", "depth",
             "global_color", "delta_relight"]


class RenderFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, renderer, meta, rays_o, rays_d, variance, *eff):
        ret = renderer._forward_impl(rays_o, rays_d, meta["near"], meta["far"], meta["perturb_overwrite"], None,
                                     meta["cos_anneal_ratio"], z_vals=meta.get("z_vals"))
        core, z = renderer._last["core"], renderer._last["z_vals"]
        keys = [k for k in DIFF_KEYS if k in ret]
        ctx.renderer, ctx.keys, ctx.meta = renderer, keys, meta
        ctx.kinds = meta["kinds"]
        saved = [renderer._f32(rays_o, z.device), renderer._f32(rays_d, z.device), z, core["mid_z_vals"], core["dists"],
                 core["sdf"].reshape(z.shape), core["gradients"], core["sampled_color"],
                 core.get("global_sampled", core["sampled_color"]), core["alpha"], core["weights"],
                 variance.detach().reshape(1).float().contiguous(), core["eikonal_den"].reshape(1)]
        ctx.save_for_backward(*saved, *[e.detach() for e in eff])
        ctx.n_saved = len(saved)
        meta["extras"] = {k: ret[k] for k in ret if k not in keys}
        meta["keys"] = keys
        return tuple(ret[k].clone() if ret[k].dim() == 0 else ret[k] for k in keys)

    @staticmethod
    def backward(ctx, *grads):
        ren = ctx.renderer
        lib = _bind()
        saved = ctx.saved_tensors
        (ro, rd, z, mid, dists, sdf, nrm, sc, gs, alpha, weights, var, eden) = saved[:ctx.n_saved]
        eff = saved[ctx.n_saved:]
        dev = z.device
        B, S = z.shape
        h = ren.handle()
        g = {k: (gr.contiguous().float() if gr is not None else None) for k, gr in zip(ctx.keys, grads)}

        bi = BackwardIn()
        for name, t in (("rays_o", ro), ("rays_d", rd), ("z", z), ("mid_z", mid), ("dists", dists), ("sdf", sdf),
                        ("gradients", nrm), ("sampled_color", sc), ("global_sampled", gs), ("alpha", alpha),
                        ("weights", weights), ("variance", var), ("eikonal_den", eden)):
            setattr(bi, name, L.ptr(t.contiguous()).value)
        hold = []

        def gp(key, shape=None):
            t = g.get(key)
            if t is None:
                return None
            t = t.reshape(shape).contiguous() if shape is not None else t
            hold.append(t)
            return t.data_ptr()

        bi.g_color_fine, bi.g_global_color = gp("color_fine"), gp("global_color")
        bi.g_weight_sum, bi.g_weight_max, bi.g_depth = gp("weight_sum", (B,)), gp("weight_max", (B,)), gp("depth", (B,))
        bi.g_weights, bi.g_cdf, bi.g_gradients = gp("weights"), gp("cdf_fine"), gp("gradients")
        bi.g_delta_relight = gp("delta_relight")
        bi.g_gradient_error = gp("gradient_error", (1,))
        if g.get("s_val") is not None:
            s = g["s_val"].sum().reshape(1).contiguous()
            hold.append(s)
            bi.g_s_val_sum = s.data_ptr()

        # effective weights (plain row-major) and zero-initialised gradient buffers in the C structs
        P, G = L.Params(), ParamGrads()
        grads_eff = []
        for (kind, idx), w, b in zip(ctx.kinds, eff[0::2], eff[1::2]):
            w, b = w.contiguous().float(), b.contiguous().float()
            gw, gb = torch.zeros_like(w), torch.zeros_like(b)
            hold += [w, b]
            grads_eff += [gw, gb]
            dst = P.relight_in if kind == "relight_in" else getattr(P, kind)[idx]
            dst.weight_g, dst.weight_v, dst.bias, dst.out, dst.in_ = None, w.data_ptr(), b.data_ptr(), w.shape[0], w.shape[1]
            gd = G.relight_in if kind == "relight_in" else getattr(G, kind)[idx]
            gd.weight, gd.bias = gw.data_ptr(), gb.data_ptr()
        dvar = torch.zeros(1, device=dev)
        G.variance = dvar.data_ptr()
        need_o, need_d = ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        d_o = torch.zeros(B, 3, device=dev) if need_o else None
        d_d = torch.zeros(B, 3, device=dev) if need_d else None
        nbytes = lib.cneus_backward_workspace_bytes(h.dref(), B, S)
        ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.cneus_render_backward(h.dref(), C.byref(P), C.byref(bi), B, S, float(ctx.meta["cos_anneal_ratio"]),
                                              C.byref(G), L.ptr(d_o), L.ptr(d_d), L.ptr(ws), ws.numel() * 4, L.stream_ptr()),
                    "cneus_render_backward")
        return (None, None, d_o, d_d, dvar.reshape(()), *grads_eff)


def render_with_grad(renderer, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0,
                     z_vals=None):
    """forward() of the drop-in renderers when autograd is on and the module is in training mode."""
    h = renderer.handle()
    effs = effective_linears(h)
    meta = dict(near=near, far=far, perturb_overwrite=perturb_overwrite, cos_anneal_ratio=cos_anneal_ratio,
                kinds=[(k, i) for k, i, _, _ in effs], z_vals=z_vals)
    flat = []
    for _, _, w, b in effs:
        flat += [w, b]
    outs = RenderFn.apply(renderer, meta, rays_o, rays_d, renderer.deviation_network.variance, *flat)
    ret = dict(zip(meta["keys"], outs))
    ret.update(meta["extras"])
    if background_rgb is not None:  # NeuS.py:274-275
        ret["color_fine"] = ret["color_fine"] + background_rgb * (1.0 - ret["weight_sum"])
    # same key order as the reference's return dict (NeuS.py:388-408)
    order = ["color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max", "gradients", "weights", "gradient_error",
             "inside_sphere", "depth", "global_color", "delta_relight", "eikonal_num", "eikonal_den"]
    return {k: ret[k] for k in order if k in ret}
