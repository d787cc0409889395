"""Drop-in for the subset of the `tinycudann` PyTorch bindings that ngp_pl uses
(ngp_pl/models/networks.py:34-83): ``NetworkWithInputEncoding``, ``Encoding`` and ``Network``, each an
``nn.Module`` holding ONE flat fp32 ``params`` Parameter (the checkpoint contract) and
called as ``module(x)``; outputs are fp16, sliced to ``[:, :n_output_dims]``.

Parameter layout inside ``params`` (documented here because upstream's is not visible in the reference):
network weights first, layer by layer, each a row-major (out, in) matrix with the input width padded to a
multiple of 16 and the output layer padded to 16 rows; then, for NetworkWithInputEncoding, the hash table
(level-major, 2 features per entry).  Initialisation: Xavier-uniform weights, table U(-1e-4, 1e-4), from
``torch.Generator(seed)`` (upstream uses its own RNG with the same default seed 1337).

Mixed precision follows tiny-cuda-nn: fp16 weights/activations with fp32 accumulation, a fixed internal
loss scale of 128 on the way back, fp32 master parameters and fp32 parameter gradients.
"""
import math

import torch

from . import _lib as L

LOSS_SCALE = 128.0
_f16, _f32 = torch.float16, torch.float32


def _pad16(n):
    return (n + 15) // 16 * 16


def hashgrid_layout(n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale):
    lay = L.GridLayout()
    L.call_nostream("b2n_hashgrid_layout", int(n_levels), int(n_features), int(log2_hashmap_size),
                    int(base_resolution), float(per_level_scale), lay)
    return lay


# ----------------------------------------------------------------------------------------------- kernels
def hashgrid_fw(x, table16, layout, out=None, n_dev=None):
    n = x.shape[0]
    if out is None:
        out = torch.empty(n, 2 * layout.n_levels, dtype=_f16, device=x.device)
    L.call("b2n_hashgrid_fw", L.ptr(x), L.ptr(table16), layout, n, L.ptr(n_dev), L.ptr(out), out.stride(0))
    return out


def hashgrid_bw(x, dy16, layout, grad_table, grad_scale, n_dev=None):
    L.call("b2n_hashgrid_bw", L.ptr(x), L.ptr(dy16), dy16.stride(0), layout, x.shape[0], L.ptr(n_dev),
           float(grad_scale), L.ptr(grad_table), None)


def frequency_fw(x, n_frequencies, out=None, n_dev=None, x_min=0.0, x_extent=1.0):
    width = _pad16(3 * n_frequencies * 2)
    if out is None:
        out = torch.empty(x.shape[0], width, dtype=_f16, device=x.device)
    L.call("b2n_frequency_fw", L.ptr(x), int(n_frequencies), x.shape[0], L.ptr(n_dev), L.ptr(out), out.stride(0),
           float(x_min), float(x_extent))
    return out


def sh4_fw(d, normalize=False, out=None, n_dev=None):
    if out is None:
        out = torch.empty(d.shape[0], 16, dtype=_f16, device=d.device)
    L.call("b2n_sh4_fw", L.ptr(d), int(normalize), d.shape[0], L.ptr(n_dev), L.ptr(out), out.stride(0))
    return out


def mlp_fw(x16, w16, in_width, n_hidden, out_act, save_hidden=True, n_dev=None):
    n, dev = x16.shape[0], x16.device
    hidden = torch.empty(n_hidden, n, 64, dtype=_f16, device=dev) if save_hidden else None
    out = torch.empty(n, 16, dtype=_f16, device=dev)
    L.call("b2n_mlp_fw", L.ptr(x16), x16.stride(0), int(in_width), L.ptr(w16), int(n_hidden), int(out_act), n,
           L.ptr(n_dev), L.ptr(hidden), L.ptr(out))
    return out, hidden


def mlp_bw(dy16, x16, w16, in_width, n_hidden, out_act, hidden, out, grad_w, grad_scale, need_din, n_dev=None):
    n = x16.shape[0]
    din = torch.empty(n, x16.stride(0), dtype=_f16, device=x16.device) if need_din else None
    L.call("b2n_mlp_bw", L.ptr(dy16), L.ptr(x16), x16.stride(0), int(in_width), L.ptr(w16), int(n_hidden),
           int(out_act), n, L.ptr(n_dev), L.ptr(hidden), L.ptr(out), float(grad_scale), L.ptr(din), L.ptr(grad_w))
    return din


def cast_half(src, dst=None):
    if dst is None:
        dst = torch.empty(src.numel(), dtype=_f16, device=src.device)
    L.call("b2n_cast_half", L.ptr(src), L.ptr(dst), src.numel())
    return dst


# ----------------------------------------------------------------------------------------------- configs
_ACT = {"None": 0, "Sigmoid": 1}


class _MlpSpec:
    def __init__(self, n_in, n_out, cfg):
        if cfg.get("otype", "FullyFusedMLP") not in ("FullyFusedMLP", "CutlassMLP"):
            raise ValueError(f"unsupported network otype {cfg.get('otype')}")
        if cfg.get("activation", "ReLU") != "ReLU" or int(cfg.get("n_neurons", 64)) != 64:
            raise ValueError("only ReLU / 64 neurons is built (the configuration NGP uses, networks.py:54-60,76-82)")
        if cfg.get("output_activation", "None") not in _ACT:
            raise ValueError(f"unsupported output_activation {cfg.get('output_activation')}")
        self.n_in, self.n_out = n_in, n_out
        self.in_width = _pad16(n_in)
        if self.in_width > 80 or n_out > 16:
            raise ValueError("FullyFusedMLP here supports in <= 80, out <= 16")
        self.n_hidden = int(cfg.get("n_hidden_layers", 1))
        self.out_act = _ACT[cfg.get("output_activation", "None")]
        self.shapes = [(64, self.in_width)] + [(64, 64)] * (self.n_hidden - 1) + [(16, 64)]
        self.n_params = sum(o * i for o, i in self.shapes)

    def init(self, gen):
        parts = []
        for li, (o, i) in enumerate(self.shapes):
            bound = math.sqrt(6.0 / (o + i))
            w = (torch.rand(o, i, generator=gen) * 2 - 1) * bound
            if li == len(self.shapes) - 1:
                w[self.n_out:] = 0                         # padded output rows carry no signal
            if li == 0:
                w[:, self.n_in:] = 0                       # padded input columns
            parts.append(w.reshape(-1))
        return torch.cat(parts)


class _EncSpec:
    def __init__(self, n_in, cfg):
        self.otype = cfg["otype"]
        self.n_in = n_in
        if self.otype in ("HashGrid", "Grid"):
            assert n_in == 3
            self.layout = hashgrid_layout(cfg.get("n_levels", 16), cfg.get("n_features_per_level", 2),
                                          cfg.get("log2_hashmap_size", 19), cfg.get("base_resolution", 16),
                                          cfg.get("per_level_scale", 2.0))
            self.n_out = self.layout.n_levels * 2
            self.n_params = self.layout.n_params
        elif self.otype == "Frequency":
            assert n_in == 3
            self.n_freq = int(cfg.get("n_frequencies", 12))
            self.n_out = _pad16(n_in * self.n_freq * 2)
            self.n_params = 0
        elif self.otype == "SphericalHarmonics":
            assert n_in == 3 and int(cfg.get("degree", 4)) == 4
            self.n_out = 16
            self.n_params = 0
        else:
            raise ValueError(f"unsupported encoding otype {self.otype}")

    def init(self, gen):
        return (torch.rand(self.n_params, generator=gen) * 2 - 1) * 1e-4

    def forward(self, x, table16):
        if self.otype in ("HashGrid", "Grid"):
            return hashgrid_fw(x, table16, self.layout)
        if self.otype == "Frequency":
            return frequency_fw(x, self.n_freq)
        return sh4_fw(x, normalize=False)


# ----------------------------------------------------------------------------------------------- autograd
class _ModuleFn(torch.autograd.Function):
    """encoding (optional) -> MLP (optional), one autograd node like tcnn's _module_function."""

    @staticmethod
    def forward(ctx, x, params, mod):
        L.require_cuda(x, params)
        enc, mlp = mod.enc, mod.mlp
        p16 = mod.half_params()
        n_mlp = mlp.n_params if mlp else 0
        if enc is not None:
            x32 = x.detach().to(_f32).contiguous()
            feat = enc.forward(x32, p16[n_mlp:] if enc.n_params else None)
        else:
            x32 = None
            feat = x.detach().to(_f16).contiguous()
            if feat.shape[1] != mlp.in_width:
                feat = torch.nn.functional.pad(feat, (0, mlp.in_width - feat.shape[1]))
        if mlp is None:
            ctx.mod = mod
            ctx.save_for_backward(x32)
            return feat
        need = any(ctx.needs_input_grad[:2])      # (grad mode is off inside Function.forward; ask the ctx)
        out, hidden = mlp_fw(feat, p16, mlp.in_width, mlp.n_hidden, mlp.out_act, save_hidden=need)
        ctx.mod = mod
        ctx.x_dtype, ctx.x_cols = x.dtype, x.shape[1]
        ctx.save_for_backward(x32, feat, hidden, out, p16)
        return out[:, :mlp.n_out]

    @staticmethod
    def backward(ctx, dy):
        mod = ctx.mod
        enc, mlp = mod.enc, mod.mlp
        if mlp is None:                                       # parameter-free encodings: nothing to do
            if enc.n_params == 0:
                return None, None, None
            (x32,) = ctx.saved_tensors
            grad = torch.zeros(enc.n_params, dtype=_f32, device=dy.device)
            dy16 = (dy.to(_f32) * LOSS_SCALE).to(_f16).contiguous()
            hashgrid_bw(x32, dy16, enc.layout, grad, 1.0 / LOSS_SCALE)
            return None, grad, None
        x32, feat, hidden, out, p16 = ctx.saved_tensors
        n = feat.shape[0]
        dy16 = torch.zeros(n, 16, dtype=_f16, device=dy.device)
        dy16[:, :mlp.n_out] = (dy.to(_f32) * LOSS_SCALE).to(_f16)
        grad = torch.zeros(mod.params.numel(), dtype=_f32, device=dy.device)
        need_din = (enc is not None and enc.n_params > 0) or (enc is None and ctx.needs_input_grad[0])
        din = mlp_bw(dy16, feat, p16, mlp.in_width, mlp.n_hidden, mlp.out_act, hidden, out, grad,
                     1.0 / LOSS_SCALE, need_din)
        dx = None
        if enc is not None and enc.n_params > 0:
            hashgrid_bw(x32, din, enc.layout, grad[mlp.n_params:], 1.0 / LOSS_SCALE)
        elif enc is None and ctx.needs_input_grad[0]:
            dx = (din[:, :ctx.x_cols].to(_f32) / LOSS_SCALE).to(ctx.x_dtype)
        return dx, grad, None


class Module(torch.nn.Module):
    def __init__(self, enc, mlp, seed=1337):
        super().__init__()
        self.enc, self.mlp, self.seed = enc, mlp, seed
        gen = torch.Generator().manual_seed(seed)
        parts = []
        if mlp is not None:
            parts.append(mlp.init(gen))
        if enc is not None and enc.n_params:
            parts.append(enc.init(gen))
        init = torch.cat(parts) if parts else torch.zeros(0)
        self.params = torch.nn.Parameter(init.to(_f32), requires_grad=True)
        self._p16, self._p16_key = None, None
        self.loss_scale = LOSS_SCALE
        self.n_input_dims = enc.n_in if enc is not None else mlp.n_in
        self.n_output_dims = mlp.n_out if mlp is not None else enc.n_out

    def half_params(self):
        """fp16 copy of the master parameters, refreshed when they change (what tcnn does per forward)."""
        p = self.params
        key = (p.data_ptr(), p._version, p.device)
        if self._p16 is None or self._p16_key != key:
            if p.numel() == 0:
                self._p16 = torch.zeros(0, dtype=_f16, device=p.device)
            else:
                self._p16 = cast_half(p.detach(), self._p16 if self._p16 is not None and
                                      self._p16.device == p.device else None)
            self._p16_key = key
        return self._p16

    def set_half_params(self, p16):
        """Let a fused optimiser hand over the fp16 copy it already produced."""
        p = self.params
        self._p16, self._p16_key = p16, (p.data_ptr(), p._version, p.device)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("tinycudann shim: CUDA tensors only (no CPU fallback)")
        return _ModuleFn.apply(x, self.params, self)


class Encoding(Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__(_EncSpec(n_input_dims, encoding_config), None, seed)
        self.encoding_config = encoding_config


class Network(Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__(None, _MlpSpec(n_input_dims, n_output_dims, network_config), seed)
        self.network_config = network_config


class NetworkWithInputEncoding(Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        enc = _EncSpec(n_input_dims, encoding_config)
        super().__init__(enc, _MlpSpec(enc.n_out, n_output_dims, network_config), seed)
        self.encoding_config, self.network_config = encoding_config, network_config
