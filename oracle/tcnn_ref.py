"""torch-CPU restatement of the tiny-cuda-nn modules that ``NGP`` builds -- TEST INFRASTRUCTURE.

tiny-cuda-nn is a third-party dependency that is absent from /root/reference (pip-installed from
github.com/NVlabs/tiny-cuda-nn, unpinned, ngp_pl/README.md:32).  Its published algorithms
(Mueller et al. 2022, "Instant Neural Graphics Primitives", and the tiny-cuda-nn encodings
documentation) are restated here; parity is anchored on the reference call sites:
  NetworkWithInputEncoding(HashGrid | Frequency -> FullyFusedMLP 64x1 -> 16)  ngp_pl/models/networks.py:34-61
  Encoding(SphericalHarmonics degree 4)                                        ngp_pl/models/networks.py:63-70
  Network(FullyFusedMLP 32 -> 64 -> 64 -> 3, sigmoid)                          ngp_pl/models/networks.py:72-83
PARITY UNPINNED (oracle/__init__.py).  All functions are differentiable torch code so that autograd
provides the gradient oracle; pass dtype=torch.float64 tables/weights for an fp64 "truth" variant.

Numerics contract (DESIGN.md "Numerics"): tables and weights are rounded to fp16, products are
accumulated in fp32, every layer output / encoding output is rounded to fp16.  The roundings are
straight-through for autograd (`_round16`): the gradient oracle is the EXACT fp32 gradient of the fp16-rounded
forward function, independent of any loss scale.  (A plain `.to(float16)` would also round every gradient that
flows back through it to fp16 -- and flush unscaled ones to zero -- which models neither tiny-cuda-nn's
loss-scaled fp16 backward pass nor anything a test should compare against.)
"""
import math

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)


class _Round16(torch.autograd.Function):
    """fp32 -> nearest fp16 value (kept in fp32); identity gradient."""
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.float16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def _round16(t):
    return _Round16.apply(t.to(torch.float32))


def _to_half(t):
    """Module output: an fp16 tensor -- or, while autograd is recording, the same fp16 values kept in an fp32 tensor,
    so that no dtype conversion (whose backward would round the gradient to fp16) sits on the gradient path."""
    return _round16(t) if (t.requires_grad and torch.is_grad_enabled()) else t.to(torch.float16)


# --------------------------------------------------------------------------- hash grid
def hashgrid_layout(n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=2.0):
    """Per-level (scale, resolution, n_entries, offset) exactly as tiny-cuda-nn sizes its GridEncoding:
    scale_l = exp2(l*log2(b))*N_min - 1 (fp32), res = ceil(scale)+1, entries = min(align8(res^3), 2^log2_T)."""
    # b^l is evaluated in double and snapped to the nearest integer multiple when within 1e-6, so that the
    # intended resolutions (16 ... 2048*scale) do not depend on one ulp of exp2f/log2f (DESIGN.md "Hash grid").
    scales, ress, sizes, offsets = [], [], [], [0]
    for l in range(n_levels):
        v = float(per_level_scale) ** l * base_resolution
        if abs(v - round(v)) < 1e-6 * v:
            v = float(round(v))
        s = np.float32(v - 1.0)
        res = int(np.ceil(s)) + 1
        n = min(res ** 3, 2 ** 31 - 1)
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        scales.append(float(s)); ress.append(res); sizes.append(n); offsets.append(offsets[-1] + n)
    return dict(n_levels=n_levels, n_features=n_features, scales=scales, resolutions=ress, sizes=sizes,
                offsets=offsets, n_params=offsets[-1] * n_features)


def _grid_index(pg, res, size):
    """pg (N,3) int64 corner coordinates -> entry index (uint32 arithmetic, wraps mod 2^32)."""
    M = 0xFFFFFFFF
    stride, index = 1, torch.zeros(pg.shape[0], dtype=torch.int64)
    for d in range(3):
        if stride > size:
            break
        index = (index + pg[:, d] * stride) & M
        stride = (stride * res) & M
    if size < stride:                                              # does not fit densely: spatial hash
        index = torch.zeros(pg.shape[0], dtype=torch.int64)
        for d in range(3):
            index = index ^ ((pg[:, d] * PRIMES[d]) & M)
    return index % size


def hashgrid_indices_weights(x, layout):
    """x (N,3) in [0,1] -> list over levels of (indices (N,8) into the flat table, weights (N,8) fp32)."""
    out = []
    x64 = x.detach().to(torch.float64)
    for l in range(layout['n_levels']):
        s, res, size, off = layout['scales'][l], layout['resolutions'][l], layout['sizes'][l], layout['offsets'][l]
        pos = (x64 * float(np.float32(s)) + 0.5).to(torch.float32)    # fmaf(scale, x, 0.5) rounded once
        pg = torch.floor(pos)
        fr = pos - pg
        pg = pg.to(torch.int64)
        idxs, ws = [], []
        for corner in range(8):
            w = torch.ones(x.shape[0], dtype=torch.float32)
            c = pg.clone()
            for d in range(3):
                if corner & (1 << d):
                    w = w * fr[:, d]; c[:, d] += 1
                else:
                    w = w * (torch.tensor(1.0) - fr[:, d])
            idxs.append(_grid_index(c, res, size) + off); ws.append(w)
        out.append((torch.stack(idxs, 1), torch.stack(ws, 1)))
    return out


def hashgrid_forward(x, table, layout, out_dtype=torch.float16):
    """table (n_entries_total, F) master parameters (fp32 or fp64).  -> (N, L*F) encoded features.
    fp32 path: entries rounded to fp16, trilinear sum in fp32, result rounded to fp16."""
    F = layout['n_features']
    tab = table.view(-1, F)
    if table.dtype == torch.float32:
        tab = _round16(tab)
    feats = []
    for idx, w in hashgrid_indices_weights(x, layout):
        v = tab[idx]                                                # (N,8,F)
        feats.append((v * w[..., None].to(tab.dtype)).sum(1))
    out = torch.cat(feats, 1)
    if table.dtype != torch.float32:
        return out
    return _to_half(out) if out_dtype == torch.float16 else out.to(out_dtype)


# --------------------------------------------------------------------------- frequency / SH
def frequency_forward(x, n_frequencies=12, out_dtype=torch.float16, pad_to=16):
    """tiny-cuda-nn Frequency: per input dim, per frequency f: sin(2^f*pi*x), cos via +pi/2 phase; the
    output is padded with ones to a multiple of 16 (3*12*2 = 72 -> 80)."""
    N, D = x.shape
    f = torch.arange(n_frequencies, dtype=torch.int32)
    xs = torch.ldexp(x.to(torch.float32)[:, :, None].expand(N, D, n_frequencies), f[None, None].expand(N, D, n_frequencies))  # scalbnf
    phase = torch.tensor([0.0, math.pi / 2], dtype=torch.float32)
    arg = xs[..., None] * torch.tensor(math.pi, dtype=torch.float32) + phase     # (N,D,F,2)
    out = torch.sin(arg).reshape(N, D * n_frequencies * 2)
    width = (out.shape[1] + pad_to - 1) // pad_to * pad_to
    out = torch.cat([out, torch.ones(N, width - out.shape[1])], 1)
    return out.to(out_dtype)


def sh4_forward(d01, out_dtype=torch.float16):
    """tiny-cuda-nn SphericalHarmonics degree 4: input in [0,1]^3 is mapped to [-1,1]^3 (networks.py:114)."""
    v = d01.to(torch.float32) * 2 - 1
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [torch.full_like(x, 0.28209479177387814),
         -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
         1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
         -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
         0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
         0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
         0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
         0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(o, 1).to(out_dtype)


# --------------------------------------------------------------------------- fully fused MLP
def mlp_layout(n_in, n_out, n_neurons=64, n_hidden_layers=1):
    """FullyFusedMLP parameter layout: per layer a row-major (out,in) matrix, no biases; the input width
    is padded to a multiple of 16 and the output width to 16 (tiny-cuda-nn).  -> list of (out,in)."""
    pin = (n_in + 15) // 16 * 16
    pout = (n_out + 15) // 16 * 16
    shapes = [(n_neurons, pin)] + [(n_neurons, n_neurons)] * (n_hidden_layers - 1) + [(pout, n_neurons)]
    return shapes


def mlp_n_params(shapes):
    return sum(o * i for o, i in shapes)


def mlp_forward(x, params, shapes, n_out, output_activation="None", return_hidden=False):
    """x (N, n_in) fp16/fp32; params flat master weights.  fp16 weights & activations, fp32 accumulate.
    -> (N, n_out) fp16 (fp64 if params are fp64)."""
    truth = params.dtype == torch.float64
    cast = (lambda t: t) if truth else _round16
    h = cast(x.to(params.dtype))
    if h.shape[1] < shapes[0][1]:
        h = torch.cat([h, torch.zeros(h.shape[0], shapes[0][1] - h.shape[1], dtype=h.dtype)], 1)
    off, hidden = 0, []
    for li, (o, i) in enumerate(shapes):
        W = cast(params[off:off + o * i].view(o, i)); off += o * i
        h = h @ W.T
        if li < len(shapes) - 1:
            h = cast(torch.relu(h)); hidden.append(h)
    if output_activation == "Sigmoid":
        h = torch.sigmoid(h)
    out = h[:, :n_out]
    out = out if truth else _to_half(out)
    return (out, hidden) if return_hidden else out


def xavier_uniform_(params, shapes, generator):
    off = 0
    for o, i in shapes:
        bound = math.sqrt(6.0 / (o + i))
        params[off:off + o * i] = (torch.rand(o * i, generator=generator) * 2 - 1) * bound
        off += o * i
    return params
