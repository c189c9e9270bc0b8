"""Stand-alone positional encoding with the reference's interface:
`get_embedder(multires, input_dims=3) -> (embed_fn, out_dim)` and `Embedder` (lib/models/tools/PositionEncoding.py:45-94).
Inside the drop-in renderers the encoding is fused into the point-shading kernels and never materialised; this mirror is for
user code that calls the embedder itself.  CUDA tensors only (no CPU path)."""
import torch

from . import _lib as L


class Embedder:
    """Same kwargs as the reference class; only the configuration `get_embedder` builds is implemented on the device
    (include_input, log-sampled bands 2^0 .. 2^max_freq_log2, periodic_fns = [sin, cos])."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d, n = kwargs["input_dims"], kwargs["num_freqs"]
        fns = list(kwargs.get("periodic_fns", [torch.sin, torch.cos]))
        if not (kwargs.get("include_input", True) and kwargs.get("log_sampling", True) and kwargs["max_freq_log2"] == n - 1
                and fns == [torch.sin, torch.cos]):
            raise NotImplementedError("only get_embedder's configuration is implemented (include_input, log_sampling, sin / cos)")
        self.input_dims, self.num_freqs = int(d), int(n)
        self.out_dim = self.input_dims * (1 + 2 * self.num_freqs)

    def embed(self, inputs):
        if not (torch.is_tensor(inputs) and inputs.is_cuda):
            raise L.CneusError("Embedder.embed needs a CUDA tensor (there is no CPU path)")
        if inputs.shape[-1] != self.input_dims:
            raise ValueError(f"expected last dimension {self.input_dims}, got {inputs.shape[-1]}")
        x = inputs.detach().to(torch.float32).contiguous().reshape(-1, self.input_dims)
        out = torch.empty(x.shape[0], self.out_dim, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            L.check(L.lib().cneus_embed(L.ptr(x) if x.numel() else None, x.shape[0], self.input_dims, self.num_freqs,
                                        L.ptr(out) if out.numel() else None, L.stream_ptr()), "cneus_embed")
        return out.reshape(*inputs.shape[:-1], self.out_dim)


def get_embedder(multires, input_dims=3):
    embedder_obj = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1, num_freqs=multires,
                            log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def embed(x, eo=embedder_obj):
        return eo.embed(x)

    return embed, embedder_obj.out_dim
