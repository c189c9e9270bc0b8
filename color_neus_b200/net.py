"""Host-side handle tying nn.Module parameters to the packed device buffers of libcneus.so."""
import ctypes as C

import torch

from . import _lib as L


def _cfg_get(cfg, key, default):
    return cfg.get(key, default) if cfg is not None else default


class NetHandle:
    """Topology descriptor + packed effective weights + workspace for one (sdf[, colour[, relight]]) stack.

    Re-packs lazily when any parameter's autograd version counter (or storage) changed, so optimiser steps (torch's and
    `FusedClipAdam`'s, which bumps the counters of the tensors its kernel writes) and `load_state_dict` are picked up
    without the caller doing anything; `invalidate()` is the explicit form for writers that bypass the counters."""

    def __init__(self, sdf_network, color_network=None, relight_network=None, primary=None):
        self.sdf, self.color, self.relight = sdf_network, color_network, relight_network
        self.primary = primary if primary is not None else sdf_network  # module whose device we follow
        self.desc = L.NetDesc()
        d = self.desc
        s = sdf_network
        d.sdf_n_lin, d.sdf_d_hidden, d.sdf_d_out = s.num_layers - 1, s.d_hidden, s.d_out
        d.sdf_multires, d.sdf_scale = s.multires, float(s.scale)
        skips = [k for k in s.skip_in if 0 < k < s.num_layers - 1]
        if len(skips) > 1:
            raise L.CneusError("only one skip connection is supported by the sm_100a kernels")
        d.sdf_skip = skips[0] if skips else -1
        if color_network is not None:
            c = color_network
            d.color_mode = L.COLOR_MODES[c.mode]
            d.color_n_lin, d.color_d_hidden, d.color_d_feature = c.num_layers - 1, c.d_hidden, c.d_feature
            d.color_multires_view, d.color_squeeze_out = c.multires_view, int(bool(c.squeeze_out))
            if c.d_out != 3:
                raise L.CneusError("colour network D_OUT must be 3")
        if relight_network is not None:
            r = relight_network
            d.has_relight = 1
            d.relight_n_layers, d.relight_y_in_layer, d.relight_d_hidden = r.n_layers, r.y_in_layer, r.d_hidden
            d.relight_multires_view = r.multires_view
            d.relight_include_grad, d.relight_inv_sigmoid = int(bool(r.include_grad)), int(bool(r.inv_sigmoid))
        self._packed = None
        self._stamp = None
        self._ws = None
        self._hold = None

    # ------------------------------------------------------------------------------------------------------
    def _linears(self):
        out = []
        for l in range(self.sdf.num_layers - 1):
            out.append(("sdf", l, getattr(self.sdf, f"lin{l}")))
        if self.color is not None:
            for l in range(self.color.num_layers - 1):
                out.append(("color", l, getattr(self.color, f"lin{l}")))
        if self.relight is not None:
            out.append(("relight_in", 0, self.relight.in_layer))
            for i, m in enumerate(self.relight.rl_mlp):
                out.append(("relight_mlp", i, m))
        return out

    def _current_stamp(self):
        st = []
        for _, _, m in self._linears():
            for p in m.parameters(recurse=False):
                st.append((p.data_ptr(), p._version))
        return tuple(st)

    def invalidate(self):
        """Force a re-pack on the next use: for code that writes the parameters behind autograd's back (raw device pointers,
        custom kernels) without bumping their version counters.  `FusedClipAdam.step` bumps them itself."""
        self._stamp = None

    def device(self):
        dev = next(self.primary.parameters()).device
        if self.primary is not self.sdf and next(self.sdf.parameters()).device != dev:
            self.sdf.to(dev)  # shape-only SDF stand-in of a colour-/relight-only handle
        return dev

    def packed(self):
        """Device pointer of the packed weights, re-packing first if the parameters changed."""
        dev = self.device()
        if dev.type != "cuda":
            raise L.CneusError("color_neus_b200 runs on CUDA devices only (no CPU path); move the module to cuda")
        lib = L.lib()
        stamp = self._current_stamp()
        if self._packed is None or self._packed.device != dev:
            nbytes = lib.cneus_packed_bytes(C.byref(self.desc))
            if nbytes == 0:
                raise L.CneusError("unsupported network topology: " + lib.cneus_last_error().decode())
            self._packed = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
            self._stamp = None
        if stamp != self._stamp:
            P = L.Params()
            hold = []

            def fill(dst, m):
                if hasattr(m, "weight_g"):
                    g, v = m.weight_g.detach(), m.weight_v.detach()
                else:
                    g, v = None, m.weight.detach()
                b = m.bias.detach()
                ts = [t.contiguous().float() if t is not None else None for t in (g, v, b)]
                hold.extend(ts)
                dst.weight_g = ts[0].data_ptr() if ts[0] is not None else None
                dst.weight_v, dst.bias = ts[1].data_ptr(), ts[2].data_ptr()
                dst.out, dst.in_ = v.shape[0], v.shape[1]

            for kind, idx, m in self._linears():
                if kind == "relight_in":
                    fill(P.relight_in, m)
                else:
                    fill(getattr(P, kind)[idx], m)
            with torch.cuda.device(dev):
                L.check(lib.cneus_pack_weights(C.byref(self.desc), C.byref(P), L.ptr(self._packed),
                                               self._packed.numel() * 4, L.stream_ptr()), "cneus_pack_weights")
            self._hold = hold  # keep temporaries alive until the stream consumed them
            self._stamp = stamp
        return L.ptr(self._packed)

    def workspace(self, n_rays=0, n_samples=0, n_points=0):
        lib = L.lib()
        need = lib.cneus_workspace_bytes(C.byref(self.desc), int(n_rays), int(n_samples), int(n_points))
        dev = self.device()
        if self._ws is None or self._ws.device != dev or self._ws.numel() * 4 < need:
            self._ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=dev)
        return L.ptr(self._ws), self._ws.numel() * 4

    def dref(self):
        return C.byref(self.desc)
