"""Drop-in renderers: `NeuS` and `Color_NeuS` with the reference's constructor / forward /
extract_geometry / extract_color signatures, return-dict keys and state_dict layout.

Mirrors lib/models/renderers/NeuS.py:68-420 and lib/models/renderers/Color_NeuS.py:10-138.
The whole forward (coarse depths, 4 rounds of SDF-guided up-sampling, fused SDF/gradient/colour/relight
evaluation, alpha compositing, Eikonal term) is a short sequence of stream-ordered kernels from libcneus.so;
the only host work per call is the one CPU RNG draw the reference makes (NeuS.py:325), kept on the host so the
caller's random stream stays bit-identical.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .fields import RelightNetwork, RenderingNetwork, SDFNetwork, SingleVarianceNetwork
from .net import NetHandle


class NeuS(nn.Module):

    def __init__(self, cfg):
        super().__init__()
        self.name = type(self).__name__
        self.cfg = cfg
        self.sdf_network = SDFNetwork(cfg.SDF if hasattr(cfg, "SDF") else cfg["SDF"])
        self.deviation_network = SingleVarianceNetwork(cfg.DEVIATION if hasattr(cfg, "DEVIATION") else cfg["DEVIATION"])
        self.color_network = RenderingNetwork(cfg.COLOR if hasattr(cfg, "COLOR") else cfg["COLOR"])
        self.n_samples = cfg.get('N_SAMPLES', 64)
        self.n_importance = cfg.get('N_IMPORTANCE', 64)
        self.n_outside = cfg.get('N_OUTSIDE', 0)
        self.up_sample_steps = cfg.get('UP_SAMPLE_STEPS', 4)
        self.perturb = cfg.get('PERTURB', 1.0)
        self.N = cfg.get('N', 64)
        if self.n_outside > 0:
            # render_core_outside / NeRF background (NeuS.py:95-134) is outside the hot-path scope
            raise NotImplementedError("N_OUTSIDE > 0 (NeRF background model) is out of scope for the B200 hot path")
        self._handle = None
        self._const = {}

    # ------------------------------------------------------------------------------------------------------
    def _relight(self):
        return None

    def handle(self):
        if self._handle is None:
            self._handle = NetHandle(self.sdf_network, self.color_network, self._relight())
        return self._handle

    def _device_const(self, key, maker, device):
        k = (key, str(device))
        if k not in self._const:
            self._const[k] = maker().to(torch.float32).contiguous().to(device)  # computed on the CPU, bit-identical
        return self._const[k]

    @staticmethod
    def _f32(t, device):
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------------------------------------------
    def sample_z(self, rays_o, rays_d, near, far, t_rand=None):
        """The no_grad sampling block of forward (NeuS.py:311-357). t_rand: raw U[0,1) draws [B,1] or None."""
        dev = rays_o.device
        h = self.handle()
        B = rays_o.shape[0]
        S = self.n_samples + self.n_importance
        lin = self._device_const(("lin", self.n_samples), lambda: torch.linspace(0.0, 1.0, self.n_samples), dev)
        u = None
        if self.n_importance > 0:
            m = self.n_importance // self.up_sample_steps
            u = self._device_const(("u", m), lambda: torch.linspace(0.0 + 0.5 / m, 1.0 - 0.5 / m, steps=m), dev)
        z = torch.empty(B, S, device=dev, dtype=torch.float32)
        ws, wsb = h.workspace(n_rays=B, n_samples=S)
        with torch.cuda.device(dev):
            L.check(L.lib().cneus_sample_z(h.dref(), h.packed(), L.ptr(rays_o), L.ptr(rays_d), L.ptr(near), L.ptr(far),
                                           L.ptr(t_rand), L.ptr(lin), L.ptr(u), B, self.n_samples, self.n_importance,
                                           self.up_sample_steps, L.ptr(z), ws, wsb, L.stream_ptr()), "cneus_sample_z")
        return z

    def up_sample(self, rays_o, rays_d, z_vals, sdf, n_importance, inv_s):
        """NeuS.py:136-181."""
        dev = rays_o.device
        ro, rd, z, s = (self._f32(t, dev) for t in (rays_o, rays_d, z_vals, sdf.reshape(z_vals.shape)))
        B, n = z.shape
        m = int(n_importance)
        u = self._device_const(("u", m), lambda: torch.linspace(0.0 + 0.5 / m, 1.0 - 0.5 / m, steps=m), dev)
        out = torch.empty(B, m, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.lib().cneus_up_sample(L.ptr(ro), L.ptr(rd), L.ptr(z), L.ptr(s), B, n, m, float(inv_s), L.ptr(u),
                                            L.ptr(out), L.stream_ptr()), "cneus_up_sample")
        return out

    def cat_z_vals(self, rays_o, rays_d, z_vals, new_z_vals, sdf, last=False):
        """NeuS.py:183-197."""
        dev = rays_o.device
        ro, rd, z, nz = (self._f32(t, dev) for t in (rays_o, rays_d, z_vals, new_z_vals))
        s = self._f32(sdf.reshape(z.shape), dev) if sdf is not None else None
        B, n = z.shape
        m = nz.shape[1]
        h = self.handle()
        z_out = torch.empty(B, n + m, device=dev, dtype=torch.float32)
        s_out = None if last else torch.empty(B, n + m, device=dev, dtype=torch.float32)
        ws, wsb = h.workspace(n_rays=B, n_samples=n + m)
        with torch.cuda.device(dev):
            L.check(L.lib().cneus_cat_z_vals(h.dref(), h.packed(), L.ptr(ro), L.ptr(rd), L.ptr(z), L.ptr(nz), L.ptr(s), B, n,
                                             m, int(bool(last)), L.ptr(z_out), L.ptr(s_out), ws, wsb, L.stream_ptr()),
                    "cneus_cat_z_vals")
        return z_out, (sdf if last else s_out)

    def render_core(self, rays_o, rays_d, z_vals, sample_dist, sdf_network=None, deviation_network=None,
                    color_network=None, background_alpha=None, background_sampled_color=None, background_rgb=None,
                    cos_anneal_ratio=0.0, **kwargs):
        """NeuS.py:199-292 / Color_NeuS.py:24-138 (background model off). Sub-network arguments are accepted for
        signature compatibility; the renderer's own networks are the ones packed on the device."""
        if background_alpha is not None or background_sampled_color is not None:
            raise NotImplementedError("background model (N_OUTSIDE > 0) is out of scope")
        for given, own in ((sdf_network, self.sdf_network), (deviation_network, self.deviation_network),
                           (color_network, self.color_network), (kwargs.get("relight_network"), self._relight())):
            if given is not None and given is not own:
                raise L.CneusError("render_core evaluates the renderer's own networks (their weights are packed on the "
                                   "device); a different network object was passed")
        dev = rays_o.device
        ro, rd, z = (self._f32(t, dev) for t in (rays_o, rays_d, z_vals))
        B, S = z.shape
        h = self.handle()
        relit = self._relight() is not None
        f = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        t = dict(color_fine=f(B, 3), weight_sum=f(B), weight_max=f(B), depth=f(B), weights=f(B, S), cdf=f(B, S),
                 inside_sphere=f(B, S), gradients=f(B, S, 3), sdf=f(B, S), sampled_color=f(B, S, 3), alpha=f(B, S),
                 mid_z=f(B, S), dists=f(B, S), scalars=f(4))
        if relit:
            t.update(global_color=f(B, 3), delta_relight=f(B, S, 3), global_sampled=f(B, S, 3))
        ro_ = L.RenderOut()
        for k in L.RENDER_OUT_FIELDS:
            setattr(ro_, k, t[k].data_ptr() if k in t else None)
        ws, wsb = h.workspace(n_rays=B, n_samples=S)
        var = self.deviation_network.variance.detach().reshape(1).float().contiguous()
        with torch.cuda.device(dev):
            L.check(L.lib().cneus_render_core(h.dref(), h.packed(), L.ptr(var), L.ptr(ro), L.ptr(rd), L.ptr(z), B, S,
                                              float(sample_dist), float(cos_anneal_ratio), C.byref(ro_), ws, wsb,
                                              L.stream_ptr()), "cneus_render_core")
        color = t["color_fine"]
        weights_sum = t["weight_sum"].unsqueeze(-1)
        if background_rgb is not None:  # NeuS.py:274-275
            color = color + background_rgb * (1.0 - weights_sum)
        out = {
            'color': color,
            'sdf': t["sdf"].reshape(-1, 1),
            'dists': t["dists"],
            'gradients': t["gradients"],
            's_val': t["scalars"][3].reshape(1, 1).expand(B * S, 1),
            'mid_z_vals': t["mid_z"],
            'weights': t["weights"],
            'cdf': t["cdf"],
            'gradient_error': t["scalars"][0],
            'inside_sphere': t["inside_sphere"],
            # extras (not in the reference dict) used by forward() and the sharded loss reduction
            'weight_sum': weights_sum, 'weight_max': t["weight_max"].unsqueeze(-1), 'depth': t["depth"],
            'alpha': t["alpha"], 'sampled_color': t["sampled_color"],
            'eikonal_num': t["scalars"][1], 'eikonal_den': t["scalars"][2],
        }
        if relit:
            out['global_color'] = t["global_color"]
            out['delta_relight'] = t["delta_relight"]
            out['global_sampled'] = t["global_sampled"]
        return out

    def forward(self, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0,
                **kwargs):
        """NeuS.py:294-408.  rays_o, rays_d [n_rays,3]; near, far [n_rays] -> dict of fp32 CUDA tensors."""
        dev = rays_d.device
        if dev.type != "cuda":
            raise L.CneusError("color_neus_b200.NeuS.forward needs CUDA tensors (there is no CPU path)")
        if len(rays_o) == 0:
            return self._empty_result(dev)
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            from .autograd import render_with_grad  # training path: analytic backward kernels
            return render_with_grad(self, rays_o, rays_d, near, far, perturb_overwrite, background_rgb,
                                    cos_anneal_ratio, z_vals=kwargs.get("z_vals"))
        ret = self._forward_impl(rays_o, rays_d, near, far, perturb_overwrite, background_rgb, cos_anneal_ratio,
                                 z_vals=kwargs.get("z_vals"))
        if torch.is_grad_enabled():
            # eval mode with autograd recording (validate_image runs like that, NeuS_Trainer.py:238-245): forward-only here;
            # a backward that reaches these outputs raises instead of silently delivering zero gradients
            from .fields import guard_no_backward
            deps = [rays_o, rays_d, near, far, *self.parameters()]
            ret = guard_no_backward("NeuS.forward in eval mode", ret, deps)
        return ret

    def _empty_result(self, dev):
        """forward() on zero rays (e.g. an empty shard of a ray-sharded render): nothing is launched; every per-ray output is
        empty and the batch-global Eikonal ratio is 0 / (0 + 1e-5) = 0 (NeuS.py:275-277), attached to the graph so that a
        training loop's backward still runs."""
        S = self.n_samples + self.n_importance
        f = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)   # noqa: E731
        zero = f(())
        if torch.is_grad_enabled() and self.training:
            zero = zero + 0.0 * self.deviation_network.variance.sum()
        ret = {'color_fine': f(0, 3), 's_val': f(0, 1), 'cdf_fine': f(0, S), 'weight_sum': f(0, 1), 'weight_max': f(0, 1),
               'gradients': f(0, S, 3), 'weights': f(0, S), 'gradient_error': zero, 'inside_sphere': f(0, S), 'depth': f(0)}
        if self._relight() is not None:
            ret['global_color'], ret['delta_relight'] = f(0, 3), f(0, S, 3)
        ret['eikonal_num'], ret['eikonal_den'] = f(()), f(())
        return ret

    def _draw_t_rand(self, n_rays, perturb_overwrite, device):
        perturb = self.perturb
        if perturb_overwrite >= 0:
            perturb = perturb_overwrite
        if perturb > 0:
            # the reference draws on the CPU default generator (NeuS.py:325); keep that so the caller's RNG
            # stream (ray selection in get_rays_multicam) advances identically
            return torch.rand([n_rays, 1]).pin_memory().to(device, non_blocking=True)
        return None

    def _forward_impl(self, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0,
                      t_rand=None, z_vals=None):
        dev = rays_d.device
        n_rays = len(rays_o)
        ro, rd = self._f32(rays_o, dev), self._f32(rays_d, dev)
        nr, fr = self._f32(near, dev).reshape(-1), self._f32(far, dev).reshape(-1)
        sample_dist = 2.0 / self.n_samples
        if z_vals is None:
            if t_rand is None:
                t_rand = self._draw_t_rand(n_rays, perturb_overwrite, dev)
            else:
                t_rand = self._f32(t_rand, dev)
            z_vals = self.sample_z(ro, rd, nr, fr, t_rand)
        else:
            z_vals = self._f32(z_vals, dev)
        r = self.render_core(ro, rd, z_vals, sample_dist, background_rgb=background_rgb,
                             cos_anneal_ratio=cos_anneal_ratio)
        n_samples = z_vals.shape[1]
        ret = {
            'color_fine': r['color'],
            # mean over the samples of one and the same value 1 / inv_s (NeuS.py:382-383): the scalar itself, no reduction launch
            's_val': r['s_val'][:1].expand(n_rays, 1),
            'cdf_fine': r['cdf'],
            'weight_sum': r['weight_sum'],
            'weight_max': r['weight_max'],
            'gradients': r['gradients'],
            'weights': r['weights'],
            'gradient_error': r['gradient_error'],
            'inside_sphere': r['inside_sphere'],
            'depth': r['depth'],
        }
        for k in ('global_color', 'delta_relight'):
            if k in r:
                ret[k] = r[k]
        # extras for ray-sharded execution: partial sums of the batch-global Eikonal ratio (parallel.py)
        ret['eikonal_num'], ret['eikonal_den'] = r['eikonal_num'], r['eikonal_den']
        self._last = dict(z_vals=z_vals, core=r)
        return ret

    # ------------------------------------------------------------------------------------------------------
    def extract_fields(self, bound_min, bound_max, resolution, lin_begin=0, lin_end=None):
        """u[x,y,z] = -sdf on linspace(bound_min, bound_max, resolution)^3 ('ij' order), the slab
        [lin_begin, lin_end) of the flattened grid -- NeuS.py:14-28 without the 64^3 blocking / per-block D2H."""
        dev = next(self.parameters()).device
        res = int(resolution)
        total = res ** 3
        lin_end = total if lin_end is None else int(lin_end)
        bmin = [float(v) for v in torch.as_tensor(bound_min).detach().cpu().reshape(-1)]
        bmax = [float(v) for v in torch.as_tensor(bound_max).detach().cpu().reshape(-1)]
        axes = [torch.linspace(bmin[i], bmax[i], res).to(dev).contiguous() for i in range(3)]
        h = self.handle()
        u = torch.empty(lin_end - int(lin_begin), device=dev, dtype=torch.float32)
        ws, wsb = h.workspace(n_points=u.numel())
        with torch.cuda.device(dev):
            L.check(L.lib().cneus_sdf_grid(h.dref(), h.packed(), L.ptr(axes[0]), L.ptr(axes[1]), L.ptr(axes[2]), res,
                                           int(lin_begin), lin_end, L.ptr(u), ws, wsb, L.stream_ptr()), "cneus_sdf_grid")
        return u

    def extract_geometry(self, bound_min, bound_max, device, resolution, threshold=0.0):
        """NeuS.py:410-417 -> (vertices [V,3] in bbox coordinates, triangles [F,3])."""
        from .marching_cubes import marching_cubes
        u = self.extract_fields(bound_min, bound_max, resolution).reshape(resolution, resolution, resolution)
        vertices, triangles = marching_cubes(u, threshold)   # on the device; only the mesh comes back (NeuS.py:35)
        b_max_np = torch.as_tensor(bound_max).detach().cpu().numpy()
        b_min_np = torch.as_tensor(bound_min).detach().cpu().numpy()
        vertices = vertices / (resolution - 1.0) * (b_max_np - b_min_np)[None, :] + b_min_np[None, :]
        return vertices, triangles

    def extract_color(self, vertices, device=None):
        """NeuS.py:419-420 / :44-64: global colour color_network(p, n, -n, feat) per vertex -> np.float32 [V,3]."""
        dev = next(self.parameters()).device
        pts = torch.as_tensor(np.asarray(vertices)).float().to(dev).contiguous().reshape(-1, 3)
        h = self.handle()
        rgb = torch.empty(pts.shape[0], 3, device=dev, dtype=torch.float32)
        ws, wsb = h.workspace(n_points=pts.shape[0])
        with torch.cuda.device(dev):
            L.check(L.lib().cneus_vertex_color(h.dref(), h.packed(), L.ptr(pts), pts.shape[0], L.ptr(rgb), ws, wsb,
                                               L.stream_ptr()), "cneus_vertex_color")
        return rgb.cpu().numpy()


class Color_NeuS(NeuS):

    def __init__(self, cfg):
        color_cfg = cfg.COLOR if hasattr(cfg, "COLOR") else cfg["COLOR"]
        assert color_cfg.get('MODE', 'idr') == 'no_view_dir'
        super().__init__(cfg)
        self.relight_network = RelightNetwork(cfg.RELIGHT if hasattr(cfg, "RELIGHT") else cfg["RELIGHT"])

    def _relight(self):
        return self.relight_network


def register(registry=None, patch_ray_utils=False):
    """Plug the B200 renderers into the reference's RENDERER registry (lib/utils/builder.py:309) under the same
    TYPE names, overriding the stock classes: call after `import lib.models`.

    patch_ray_utils=True additionally rebinds the trainer's `get_rays_multicam` (NeuS_Trainer.py:110) to
    `color_neus_b200.rays.get_rays_multicam`, which selects the same pixels from the same CPU RNG stream but generates
    only those rays on the device."""
    if registry is None:
        from lib.utils.builder import RENDERER as registry  # the reference tree must be importable
    registry.register_module(name="NeuS", force=True, module=NeuS)
    registry.register_module(name="Color_NeuS", force=True, module=Color_NeuS)
    if patch_ray_utils:
        import lib.models.NeuS_Trainer as trainer_mod
        from . import rays
        trainer_mod.get_rays_multicam = rays.get_rays_multicam
    return registry
