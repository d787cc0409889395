"""CPU tests of the oracle itself (parity is unpinned by the reference -- no tests/golden vectors exist there --
so the oracle is pinned by: its two independent restatements agreeing bit for bit, analytic closed forms, an
fp64 autograd check of the compositing gradient, and the committed golden fixtures in tests/golden/)."""
import os

import numpy as np
import pytest
import torch

from conftest import make_scene
from oracle import clib as C
from oracle import ngp_ref as O
from oracle import tcnn_ref as T
from oracle import vren_ref as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def hits_of(s):
    _, h, _ = C.ray_aabb_intersect(s["rays_o"], s["rays_d"], s["center"], s["half_size"], 1)
    h[(h[:, 0, 0] >= 0) & (h[:, 0, 0] < 0.05), 0, 0] = 0.05
    return h[:, 0].contiguous()


@pytest.fixture(scope="module")
def s05():
    return make_scene(0.5, 512, seed=21)


def test_aabb_c_equals_torch_and_analytic(s05):
    a = C.ray_aabb_intersect(s05["rays_o"], s05["rays_d"], s05["center"], s05["half_size"], 1)
    b = R.ray_aabb_intersect(s05["rays_o"], s05["rays_d"], s05["center"], s05["half_size"], 1)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # analytic: ray along +x from (-2,0,0) hits the unit-half box at t = 1.5 .. 2.5; parallel offset ray misses
    o = torch.tensor([[-2.0, 0.0, 0.0], [-2.0, 0.0, 0.9], [0.0, 0.0, 0.0]])
    d = torch.tensor([[1.0, 1e-9, 1e-9], [1.0, 1e-9, 1e-9], [0.0, 0.0, 2.0]])
    cnt, t, idx = C.ray_aabb_intersect(o, d, torch.zeros(1, 3), torch.full((1, 3), 0.5), 1)
    assert cnt.tolist() == [1, 0, 1]
    assert t[0, 0].tolist() == [1.5, 2.5] and t[1, 0].tolist() == [-1.0, -1.0] and idx[1, 0] == -1
    assert t[2, 0].tolist() == [0.0, 0.25]                     # origin inside: near clipped to 0
    # several boxes, near-to-far order
    centers = torch.tensor([[3.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0]]); half = torch.full((3, 3), 0.25)
    cnt, t, idx = C.ray_aabb_intersect(torch.zeros(1, 3), torch.tensor([[1.0, 1e-9, 1e-9]]), centers, half, 2)
    assert cnt.item() == 3 and idx[0].tolist() == [1, 2]
    assert torch.equal(R.ray_aabb_intersect(torch.zeros(1, 3), torch.tensor([[1.0, 1e-9, 1e-9]]), centers, half, 2)[2], idx)


def test_morton_and_packbits():
    g = torch.Generator().manual_seed(0)
    coords = torch.randint(0, 128, (5000, 3), generator=g, dtype=torch.int32)
    m = C.morton3D(coords)
    assert torch.equal(m, R.morton3D(coords))
    assert torch.equal(C.morton3D_invert(m), coords) and torch.equal(R.morton3D_invert(m), coords)
    assert C.morton3D(torch.tensor([[1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 3, 3]], dtype=torch.int32)).tolist() == [1, 2, 4, 63]
    grid = torch.randn(4096, generator=g)
    a = torch.zeros(512, dtype=torch.uint8); b = torch.zeros(512, dtype=torch.uint8)
    C.packbits(grid, 0.1, a); R.packbits(grid, 0.1, b)
    ref = np.packbits(grid.numpy() > 0.1, bitorder="little")
    assert np.array_equal(a.numpy(), ref) and np.array_equal(b.numpy(), ref)


@pytest.mark.parametrize("scale,esf,max_samples", [(0.5, 0.0, 1024), (4.0, 1 / 256, 1024), (0.5, 0.0, 48)])
def test_marcher_ladder_form_equals_serial_loop(scale, esf, max_samples):
    """The vectorised 'ladder' restatement (which the warp-cooperative CUDA kernel parallelises) is bit-identical
    to the serial reference loop."""
    s = make_scene(scale, 384, seed=22)
    h = hits_of(s)
    a = C.raymarching_train(s["rays_o"], s["rays_d"], h, s["bitfield"], s["cascades"], scale, esf, s["noise"], 128, max_samples)
    b = R.raymarching_train(s["rays_o"], s["rays_d"], h, s["bitfield"], s["cascades"], scale, esf, s["noise"], 128, max_samples)
    assert int(a[5][0]) > 0
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert int(a[0][:, 2].max()) <= max_samples
    # rays_a is CSR-like and deterministic: start = exclusive prefix sum of N in ray order
    assert torch.equal(a[0][:, 1], torch.cumsum(a[0][:, 2], 0) - a[0][:, 2]) and torch.equal(a[0][:, 0], torch.arange(384))


def test_marcher_properties(s05):
    h = hits_of(s05)
    args = (s05["rays_o"], s05["rays_d"], h)
    empty = torch.zeros_like(s05["bitfield"]); full = torch.full_like(s05["bitfield"], 255)
    r = C.raymarching_train(*args, empty, 1, 0.5, 0.0, s05["noise"], 128, 1024)
    assert int(r[5][0]) == 0
    r = C.raymarching_train(*args, full, 1, 0.5, 0.0, s05["noise"], 128, 1024)
    rays_a, xyzs, dirs, deltas, ts, _ = r
    dt = np.float32(1.73205080757) / np.float32(1024)
    assert torch.all(deltas == float(dt))                           # constant step for exp_step_factor = 0
    hit = h[:, 0] >= 0
    assert torch.all(rays_a[~hit, 2] == 0)                          # a miss (-1) emits nothing
    k = int(torch.nonzero(rays_a[:, 2] > 3)[0])
    seg = ts[rays_a[k, 1]:rays_a[k, 1] + rays_a[k, 2]]
    assert torch.all(seg[1:] > seg[:-1]) and torch.allclose(seg[1:] - seg[:-1], torch.full_like(seg[1:], float(dt)), atol=1e-6)
    # samples lie inside the box and on the ray; occupied-only scene: every sample's cell is occupied
    r = C.raymarching_train(*args, s05["bitfield"], 1, 0.5, 0.0, s05["noise"], 128, 1024)
    rays_a, xyzs, dirs, deltas, ts, _ = r
    assert float(xyzs.abs().max()) <= 0.5 + 1e-5
    ray = torch.repeat_interleave(torch.arange(512), rays_a[:, 2])
    assert torch.equal(xyzs, s05["rays_o"][ray] + ts[:, None] * s05["rays_d"][ray])
    cell = ((xyzs / 0.5 + 1) * 0.5 * 128).clamp(0, 127).to(torch.int32)
    idx = R.morton3D(cell).long()
    assert torch.all(((s05["bitfield"][idx // 8].long() >> (idx % 8)) & 1) == 1)


def test_test_marcher_c_equals_torch(s05):
    hA, hB = hits_of(s05), hits_of(s05)
    alive = torch.arange(512)
    for ns in (1, 4, 16):
        a = C.raymarching_test(s05["rays_o"], s05["rays_d"], hA, alive, s05["bitfield"], 1, 0.5, 0.0, 128, 1024, ns)
        b = R.raymarching_test(s05["rays_o"], s05["rays_d"], hB, alive, s05["bitfield"], 1, 0.5, 0.0, 128, 1024, ns)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        assert torch.equal(hA, hB)
        assert torch.all((a[1].abs().sum(-1) == 0) == (torch.arange(ns)[None] >= a[4][:, None]))   # unused slots zero


def test_composite_closed_form_and_early_stop():
    n, sigma, dt = 100, 2.0, 0.02
    rays_a = torch.tensor([[0, 0, n]])
    ts = torch.arange(n) * dt + 1.0
    for impl in (C, R):
        op, dp, d2, rgb = impl.composite_train_fw(torch.full((n,), sigma), torch.full((n, 3), 0.5), torch.full((n,), dt), ts, rays_a, 0.0)
        assert abs(op.item() - (1 - np.exp(-sigma * n * dt))) < 1e-5
        assert abs(rgb[0, 0].item() - 0.5 * op.item()) < 1e-6
        # early stop: with threshold 0.5 the ray stops at the first sample where T <= 0.5 (that sample included)
        op2 = impl.composite_train_fw(torch.full((n,), sigma), torch.full((n, 3), 0.5), torch.full((n,), dt), ts, rays_a, 0.5)[0]
        k = int(np.ceil(np.log(2) / (sigma * dt)))
        assert abs(op2.item() - (1 - np.exp(-sigma * dt * k))) < 1e-5
        # zero-sample ray -> zeros
        z = impl.composite_train_fw(torch.zeros(0), torch.zeros(0, 3), torch.zeros(0), torch.zeros(0), torch.tensor([[0, 0, 0]]), 1e-4)
        assert z[0].item() == 0 and z[3].abs().sum().item() == 0


def test_composite_backward_formula_vs_fp64_autograd(s05):
    h = hits_of(s05)
    rays_a, xyzs, dirs, deltas, ts, _ = C.raymarching_train(s05["rays_o"], s05["rays_d"], h, s05["bitfield"], 1, 0.5, 0.0, s05["noise"], 128, 1024)
    N, n = xyzs.shape[0], 512
    g = torch.Generator().manual_seed(1)
    sig = torch.rand(N, generator=g) * 10; col = torch.rand(N, 3, generator=g)
    fw = C.composite_train_fw(sig, col, deltas, ts, rays_a, 0.0)            # no early stop: formula is the exact gradient
    grads = [torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, 3, generator=g)]
    ds, dc = C.composite_train_bw(*grads, sig, col, deltas, ts, rays_a, *fw, 0.0)
    ds2, dc2 = R.composite_train_bw(*grads, sig, col, deltas, ts, rays_a, *fw, 0.0)
    assert (ds - ds2).abs().max() <= 1e-4 * ds.abs().max() and (dc - dc2).abs().max() <= 1e-5
    # fp64 autograd through a straightforward differentiable compositing
    s64 = sig.double().requires_grad_(True); c64 = col.double().requires_grad_(True)
    ray = torch.repeat_interleave(torch.arange(n), rays_a[:, 2])
    a = 1 - torch.exp(-s64 * deltas.double())
    logT = torch.log1p(-a)
    cum = torch.cumsum(logT, 0)
    seg_start = torch.zeros(n, dtype=torch.float64)
    first = rays_a[:, 1][rays_a[:, 2] > 0]
    base = torch.cat([torch.zeros(1, dtype=torch.float64), cum])[first]
    seg_start[rays_a[:, 2] > 0] = base
    T_before = torch.exp(cum - logT - seg_start[ray])
    w = a * T_before
    O_ = torch.zeros(n, dtype=torch.float64).index_add(0, ray, w)
    D_ = torch.zeros(n, dtype=torch.float64).index_add(0, ray, w * ts.double())
    D2_ = torch.zeros(n, dtype=torch.float64).index_add(0, ray, w * ts.double() ** 2)
    RGB_ = torch.zeros(n, 3, dtype=torch.float64).index_add(0, ray, w[:, None] * c64)
    torch.testing.assert_close(O_.float(), fw[0], rtol=1e-5, atol=1e-6)
    loss = (O_ * grads[0].double()).sum() + (D_ * grads[1].double()).sum() + (D2_ * grads[2].double()).sum() + (RGB_ * grads[3].double()).sum()
    loss.backward()
    assert (ds.double() - s64.grad).abs().max() <= 2e-4 * s64.grad.abs().max()
    assert (dc.double() - c64.grad).abs().max() <= 1e-5


def test_hashgrid_level0_is_trilinear_interpolation():
    lay = T.hashgrid_layout(16, 2, 19, 16, np.exp(np.log(2048 * 0.5 / 16) / 15))
    assert lay["resolutions"][0] == 16 and lay["resolutions"][-1] == 1024 and lay["n_params"] == 11420064
    assert lay["sizes"][5] == 262144 and lay["sizes"][6] == 524288
    g = torch.Generator().manual_seed(2)
    table = torch.zeros(lay["n_params"], dtype=torch.float64)
    lvl0 = torch.rand(16, 16, 16, 2, generator=g, dtype=torch.float64)            # [z][y][x][f]
    table[:16 ** 3 * 2] = lvl0.reshape(-1)
    x = torch.rand(200, 3, generator=g) * 0.9
    enc = T.hashgrid_forward(x, table, lay)[:, :2]
    # independent trilinear interpolation on the 16^3 lattice with pos = x*15 + 0.5
    p = x.double() * 15 + 0.5
    i0 = torch.floor(p).long(); f = p - i0
    want = torch.zeros(200, 2, dtype=torch.float64)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                wgt = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                want += wgt[:, None] * lvl0[i0[:, 2] + dz, i0[:, 1] + dy, i0[:, 0] + dx]
    torch.testing.assert_close(enc, want, rtol=1e-5, atol=1e-6)


def test_hash_index_brute_force():
    lay = T.hashgrid_layout(16, 2, 14, 16, 1.5)
    l = 8
    res, size, off = lay["resolutions"][l], lay["sizes"][l], lay["offsets"][l]
    assert res ** 3 > size == 1 << 14
    pg = torch.tensor([[3, 7, 11], [100, 5, 77], [res, res, res]])
    got = T._grid_index(pg, res, size)
    for (x, y, z), gi in zip(pg.tolist(), got.tolist()):
        assert gi == ((x * 1) ^ ((y * 2654435761) & 0xFFFFFFFF) ^ ((z * 805459861) & 0xFFFFFFFF)) % size


def test_sh4_and_frequency_reference_values():
    d = torch.tensor([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
    sh = T.sh4_forward((d + 1) / 2, out_dtype=torch.float32)
    assert abs(sh[0, 0].item() - 0.28209479) < 1e-6 and abs(sh[0, 2].item() - 0.48860251) < 1e-6
    assert abs(sh[0, 6].item() - (0.94617470 - 0.31539157)) < 1e-6 and abs(sh[1, 3].item() + 0.48860251) < 1e-6
    # orthonormality of the 16 real SH over the sphere (Monte-Carlo)
    g = torch.Generator().manual_seed(3)
    v = torch.randn(200000, 3, generator=g); v = v / v.norm(dim=-1, keepdim=True)
    Y = T.sh4_forward((v + 1) / 2, out_dtype=torch.float32).double()
    gram = (Y.T @ Y) / Y.shape[0] * 4 * np.pi
    assert (gram - torch.eye(16, dtype=torch.float64)).abs().max() < 0.03
    f = T.frequency_forward(torch.tensor([[0.25, 0.5, 0.0]]), out_dtype=torch.float32)
    assert f.shape == (1, 80) and torch.all(f[0, 72:] == 1)
    assert abs(f[0, 0].item() - np.sin(np.pi * 0.25)) < 1e-6 and abs(f[0, 1].item() - np.cos(np.pi * 0.25)) < 1e-6
    assert abs(f[0, 2].item() - np.sin(2 * np.pi * 0.25)) < 1e-6


def test_mlp_fp16_path_close_to_fp64_truth():
    g = torch.Generator().manual_seed(4)
    shapes = T.mlp_layout(32, 3, 64, 2)
    assert shapes == [(64, 32), (64, 64), (16, 64)] and T.mlp_n_params(shapes) == 7168
    p = T.xavier_uniform_(torch.zeros(7168), shapes, g)
    x = torch.randn(500, 32, generator=g) * 0.5
    y16 = T.mlp_forward(x, p, shapes, 3, "Sigmoid")
    y64 = T.mlp_forward(x.double(), p.double(), shapes, 3, "Sigmoid")
    assert y16.dtype == torch.float16 and (y16.double() - y64).abs().max() < 5e-3


def test_oracle_training_reduces_loss():
    from google_nerf_b200 import synthetic as syn
    s = make_scene(0.5, 256, seed=23)
    ref = O.NGPRef(0.5, log2_T=12, seed=1)
    ref.density_bitfield = s["bitfield"]
    tgt = syn.shade(s["rays_o"], s["rays_d"], 0.5)
    opt = O.AdamRef([ref.xyz_params, ref.rgb_params])
    losses = [O.train_step(ref, opt, s["rays_o"], s["rays_d"], tgt, s["noise"])[0] for _ in range(8)]
    assert losses[-1] < losses[0]


def _golden_cases():
    s = make_scene(0.5, 96, seed=31)
    h = hits_of(s)
    train = C.raymarching_train(s["rays_o"], s["rays_d"], h, s["bitfield"], 1, 0.5, 0.0, s["noise"], 128, 1024)
    g = torch.Generator().manual_seed(32)
    N = train[1].shape[0]
    sig = torch.rand(N, generator=g) * 30; col = torch.rand(N, 3, generator=g)
    fw = C.composite_train_fw(sig, col, train[3], train[4], train[0], 1e-4)
    lay = T.hashgrid_layout(16, 2, 12, 16, np.exp(np.log(2048 * 0.5 / 16) / 15))
    table = (torch.rand(lay["n_params"], generator=g) * 2 - 1) * 0.5
    x = torch.rand(64, 3, generator=g)
    enc = T.hashgrid_forward(x, table, lay)
    return dict(hits_t=h, rays_a=train[0], ts=train[4], xyzs=train[1], opacity=fw[0], rgb=fw[3], depth=fw[1],
                enc=enc.float(), sh=T.sh4_forward(x, out_dtype=torch.float32), freq=T.frequency_forward(x).float())


def test_golden_fixture_matches():
    """tests/golden/vren_tcnn_small.npz was written by tests/golden/make_golden.py from this oracle; it freezes the
    oracle so that later edits cannot silently move the target."""
    path = os.path.join(GOLDEN, "vren_tcnn_small.npz")
    gold = np.load(path)
    cur = _golden_cases()
    for k in gold.files:
        a, b = gold[k], cur[k].numpy()
        if k in ("hits_t", "rays_a", "ts", "xyzs"):
            assert np.array_equal(a, b), k                     # bit-exact integer / geometry outputs
        else:
            np.testing.assert_allclose(b, a, rtol=1e-5, atol=1e-6, err_msg=k)
