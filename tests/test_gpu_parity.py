"""CUDA path vs oracle on the same seeded inputs, through the C ABI (ctypes -> libb2n.so).

Tolerances (BASELINE.json north_star): ray-AABB and sample indices bit-exact; composited rgb/depth/opacity
1e-5 relative in fp32; encodings and gradients 1e-3 at fp16."""
import numpy as np
import pytest
import torch

from conftest import make_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(d):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in d.items()}


@pytest.fixture(scope="module")
def mods(built_lib):
    from google_nerf_b200 import vren, tinycudann
    from oracle import clib, vren_ref, tcnn_ref
    return dict(vren=vren, tcnn=tinycudann, C=clib, R=vren_ref, T=tcnn_ref)


def hits_of(mods, s, near=True):
    _, hits_t, _ = mods["C"].ray_aabb_intersect(s["rays_o"], s["rays_d"], s["center"], s["half_size"], 1)
    if near:
        hits_t[(hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < 0.05), 0, 0] = 0.05
    return hits_t


# ------------------------------------------------------------------------------------------------ geometry
@pytest.mark.parametrize("max_hits,n_vox", [(1, 1), (3, 5), (2, 7)])
def test_ray_aabb_bit_exact(mods, scene05, max_hits, n_vox):
    s = scene05
    g = torch.Generator().manual_seed(3)
    centers = torch.zeros(1, 3) if n_vox == 1 else (torch.rand(n_vox, 3, generator=g) - 0.5)
    halfs = torch.ones(1, 3) * 0.5 if n_vox == 1 else torch.rand(n_vox, 3, generator=g) * 0.3 + 0.05
    ro, rd = s["rays_o"].clone(), s["rays_d"].clone()
    ro[:8] = 0.1                                     # origins inside the box
    rd[8:12, 0] = 0.0                                # axis-parallel rays (1/0 = inf)
    ref = mods["C"].ray_aabb_intersect(ro, rd, centers, halfs, max_hits)
    got = mods["vren"].ray_aabb_intersect(ro.to(DEV), rd.to(DEV), centers.to(DEV), halfs.to(DEV), max_hits)
    for a, b in zip(ref, got):
        assert a.dtype == b.dtype
        assert torch.equal(a, b.cpu())


def test_ray_sphere_bit_exact(mods, scene05):
    s = scene05
    centers = torch.tensor([[0.0, 0.0, 0.0], [0.2, 0.1, -0.1]]); radii = torch.tensor([0.4, 0.2])
    ref = mods["C"].ray_aabb_intersect(s["rays_o"], s["rays_d"], centers, radii, 2, sphere=True)
    got = mods["vren"].ray_sphere_intersect(s["rays_o"].to(DEV), s["rays_d"].to(DEV), centers.to(DEV), radii.to(DEV), 2)
    assert torch.equal(ref[0], got[0].cpu()) and torch.equal(ref[2], got[2].cpu())
    torch.testing.assert_close(got[1].cpu(), ref[1], rtol=1e-5, atol=1e-6)   # sqrt/div: IEEE on both, fma differs


def test_morton_exhaustive_roundtrip(mods):
    G = 128
    r = torch.arange(G, dtype=torch.int32)
    z, y, x = torch.meshgrid(r, r, r, indexing="ij")
    coords = torch.stack([x, y, z], -1).reshape(-1, 3)
    idx = mods["vren"].morton3D(coords.to(DEV))
    assert torch.equal(idx.cpu(), mods["C"].morton3D(coords))
    assert torch.equal(torch.sort(idx)[0].cpu(), torch.arange(G ** 3, dtype=torch.int32))   # a bijection
    assert torch.equal(mods["vren"].morton3D_invert(idx).cpu(), coords)


def test_packbits_bit_exact(mods):
    g = torch.Generator().manual_seed(1)
    grid = torch.randn(2, 128 ** 3, generator=g)
    grid[0, :64] = -1.0
    bf = torch.zeros(2 * 128 ** 3 // 8, dtype=torch.uint8, device=DEV)
    mods["vren"].packbits(grid.to(DEV), 0.25, bf)
    ref = np.packbits((grid.numpy().reshape(-1) > 0.25), bitorder="little")
    assert np.array_equal(bf.cpu().numpy(), ref)
    thr = torch.tensor([0.5, 0, 0], device=DEV)
    mods["vren"].packbits(grid.to(DEV), 123.0, bf, threshold_dev=thr)
    assert np.array_equal(bf.cpu().numpy(), np.packbits((grid.numpy().reshape(-1) > 0.5), bitorder="little"))


# ------------------------------------------------------------------------------------------------ marcher
@pytest.mark.parametrize("scale,esf", [(0.5, 0.0), (4.0, 1 / 256), (4.0, 0.0), (0.5, 1 / 256)])
def test_raymarching_train_bit_exact(mods, scale, esf):
    s = make_scene(scale, 1536, seed=5)
    hits_t = hits_of(mods, s)[:, 0].contiguous()
    ref = mods["C"].raymarching_train(s["rays_o"], s["rays_d"], hits_t, s["bitfield"], s["cascades"], scale, esf,
                                      s["noise"], 128, 1024)
    d = cu(s)
    got = mods["vren"].raymarching_train(d["rays_o"], d["rays_d"], hits_t.to(DEV), d["bitfield"], s["cascades"], scale,
                                         esf, d["noise"], 128, 1024)
    assert int(ref[5][0]) > 1000
    for name, a, b in zip(["rays_a", "xyzs", "dirs", "deltas", "ts"], ref, got):
        assert torch.equal(a, b.cpu()), name
    assert int(got[5][0]) == int(ref[5][0])


@pytest.mark.parametrize("scale,esf,max_samples", [(0.5, 0.0, 1024), (4.0, 1 / 256, 1024), (4.0, 0.0, 1024), (0.5, 1 / 256, 1024),
                                                   (0.5, 0.0, 48)])
def test_raymarching_train_serial_count_bit_exact(mods, scale, esf, max_samples):
    """Thread-per-ray count pass + mask replay (the trainer's prefetch path) against the serial C loop."""
    s = make_scene(scale, 1536, seed=6)
    hits_t = hits_of(mods, s)[:, 0].contiguous()
    ref = mods["C"].raymarching_train(s["rays_o"], s["rays_d"], hits_t, s["bitfield"], s["cascades"], scale, esf,
                                      s["noise"], 128, max_samples)
    d = cu(s)
    args = (d["rays_o"], d["rays_d"], hits_t.to(DEV), d["bitfield"], s["cascades"], scale, esf, d["noise"], 128, max_samples)
    rays_a, counter, ws = mods["vren"].raymarching_train_count(*args, serial=True)
    rays_w, counter_w, ws_w = mods["vren"].raymarching_train_count(*args)
    total = int(counter[0].item())
    assert total == int(ref[5][0]) > 1000 and torch.equal(rays_a.cpu(), ref[0]) and torch.equal(counter, counter_w)
    got = mods["vren"].raymarching_train_write(*args, rays_a, total, ws)
    for name, a, b in zip(["xyzs", "dirs", "deltas", "ts"], ref[1:5], got):
        assert torch.equal(a, b.cpu()), name
    # both count passes leave the same replay masks for the rays they marched
    n_chunks = (ws[:, 0] & 0x7fffffff).long()
    col = torch.arange(1, 64, device=DEV)[None, :]
    live = col <= n_chunks[:, None]
    assert torch.equal(ws[:, 1:].where(live, torch.zeros_like(ws[:, 1:])), ws_w[:, 1:].where(live, torch.zeros_like(ws_w[:, 1:])))


def test_raymarching_train_edges(mods, scene05):
    s = scene05
    hits_t = hits_of(mods, s)[:, 0].contiguous()
    d = cu(s)
    # empty grid -> no samples; full grid -> capped uniform ladder; misses -> zero samples
    for fill, name in ((0, "empty"), (255, "full")):
        bf = torch.full_like(s["bitfield"], fill)
        ref = mods["C"].raymarching_train(s["rays_o"], s["rays_d"], hits_t, bf, 1, 0.5, 0.0, s["noise"], 128, 1024)
        got = mods["vren"].raymarching_train(d["rays_o"], d["rays_d"], hits_t.to(DEV), bf.to(DEV), 1, 0.5, 0.0, d["noise"], 128, 1024)
        for a, b in zip(ref[:5], got[:5]):
            assert torch.equal(a, b.cpu()), name
        if fill == 0:
            assert int(got[5][0]) == 0
    # max_samples cap
    bf = torch.full_like(s["bitfield"], 255)
    ref = mods["C"].raymarching_train(s["rays_o"], s["rays_d"], hits_t, bf, 1, 0.5, 0.0, s["noise"], 128, 64)
    got = mods["vren"].raymarching_train(d["rays_o"], d["rays_d"], hits_t.to(DEV), bf.to(DEV), 1, 0.5, 0.0, d["noise"], 128, 64)
    assert int(ref[0][:, 2].max()) == 64
    for a, b in zip(ref[:5], got[:5]):
        assert torch.equal(a, b.cpu())
    # zero rays
    z = torch.zeros(0, 3, device=DEV)
    got = mods["vren"].raymarching_train(z, z, torch.zeros(0, 2, device=DEV), bf.to(DEV), 1, 0.5, 0.0,
                                         torch.zeros(0, device=DEV), 128, 1024)
    assert got[1].shape == (0, 3) and int(got[5][0]) == 0


def test_raymarching_train_capacity_clamp(mods, scene05):
    s = scene05
    hits_t = hits_of(mods, s)[:, 0].contiguous()
    d = cu(s)
    full = mods["C"].raymarching_train(s["rays_o"], s["rays_d"], hits_t, s["bitfield"], 1, 0.5, 0.0, s["noise"], 128, 1024)
    total = int(full[5][0]); cap = total // 2
    args = (d["rays_o"], d["rays_d"], hits_t.to(DEV), d["bitfield"], 1, 0.5, 0.0, d["noise"], 128, 1024)
    rays_a, counter, ws = mods["vren"].raymarching_train_count(*args, capacity=cap)
    c = counter.cpu()
    assert int(c[0]) == cap and int(c[2]) == 1 and int(c[3]) == total
    assert int((rays_a[:, 1] + rays_a[:, 2]).max()) <= cap
    for workspace in (ws, None):                                   # mask replay and full re-march must agree
        xyzs, dirs, deltas, ts = mods["vren"].raymarching_train_write(*args, rays_a, cap, workspace)
        assert torch.equal(xyzs.cpu(), full[1][:cap]) and torch.equal(ts.cpu(), full[4][:cap])
        assert torch.equal(deltas.cpu(), full[3][:cap]) and torch.equal(dirs.cpu(), full[2][:cap])


@pytest.mark.parametrize("scale,esf,n_rays", [(0.5, 0.0, 1024), (4.0, 1 / 256, 1024), (0.5, 0.0, 20000),
                                               (4.0, 1 / 256, 17000)])
def test_raymarching_test_and_composite_test(mods, scale, esf, n_rays):
    # >= 16384 live rays take the marcher's thread-per-ray path, fewer the warp-per-ray one (rounds of the same run
    # cross over as rays die)
    s = make_scene(scale, n_rays, seed=7)
    n = s["rays_o"].shape[0]
    hA = hits_of(mods, s)[:, 0].contiguous(); hB = hA.clone().to(DEV)
    d = cu(s)
    alive = torch.arange(n)
    opA, dpA, cA = torch.zeros(n), torch.zeros(n), torch.zeros(n, 3)
    opB, dpB, cB = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(n, 3, device=DEV)
    g = torch.Generator().manual_seed(0)
    for ns in (1, 2, 4, 8, 64):
        if alive.numel() == 0:
            break
        ref = mods["C"].raymarching_test(s["rays_o"], s["rays_d"], hA, alive, s["bitfield"], s["cascades"], scale, esf, 128, 1024, ns)
        aB = alive.to(DEV)
        got = mods["vren"].raymarching_test(d["rays_o"], d["rays_d"], hB, aB, d["bitfield"], s["cascades"], scale, esf, 128, 1024, ns)
        for name, a, b in zip(["xyzs", "dirs", "deltas", "ts", "n_eff"], ref, got):
            assert torch.equal(a, b.cpu()), (ns, name)
        assert torch.equal(hA, hB.cpu())
        sig = torch.rand(alive.numel(), ns, generator=g) * 60
        col = torch.rand(alive.numel(), ns, 3, generator=g)
        aA = alive.clone()
        mods["C"].composite_test_fw(sig, col, ref[2], ref[3], hA, aA, 1e-2, ref[4], opA, dpA, cA)
        mods["vren"].composite_test_fw(sig.to(DEV), col.to(DEV), got[2], got[3], hB, aB, 1e-2, got[4], opB, dpB, cB)
        assert torch.equal(aA, aB.cpu())
        torch.testing.assert_close(opB.cpu(), opA, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(cB.cpu(), cA, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(dpB.cpu(), dpA, rtol=1e-5, atol=1e-6)
        alive = aA[aA >= 0]


# ------------------------------------------------------------------------------------------------ compositing
def _packed(mods, s, scale=0.5, esf=0.0):
    hits_t = hits_of(mods, s)[:, 0].contiguous()
    return mods["C"].raymarching_train(s["rays_o"], s["rays_d"], hits_t, s["bitfield"], s["cascades"], scale, esf,
                                       s["noise"], 128, 1024)


@pytest.mark.parametrize("sigma_max,thr", [(20.0, 1e-4), (300.0, 1e-4), (300.0, 1e-2), (0.0, 1e-4)])
def test_composite_train_fw_bw(mods, scene05, sigma_max, thr):
    rays_a, xyzs, dirs, deltas, ts, _ = _packed(mods, scene05)
    N, n = xyzs.shape[0], rays_a.shape[0]
    g = torch.Generator().manual_seed(2)
    sig = torch.rand(N, generator=g) * sigma_max
    col = torch.rand(N, 3, generator=g)
    ref = mods["C"].composite_train_fw(sig, col, deltas, ts, rays_a, thr)
    dv = lambda t: t.to(DEV)
    got = mods["vren"].composite_train_fw(dv(sig), dv(col), dv(deltas), dv(ts), dv(rays_a), thr)
    for name, a, b in zip(["opacity", "depth", "depth_sq", "rgb"], ref, got):
        torch.testing.assert_close(b.cpu(), a, rtol=1e-5, atol=1e-6, msg=lambda m: f"{name}: {m}")
    grads = [torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g),
             torch.randn(n, 3, generator=g)]
    rb = mods["C"].composite_train_bw(*grads, sig, col, deltas, ts, rays_a, *ref, thr)
    gb = mods["vren"].composite_train_bw(*[dv(v) for v in grads], dv(sig), dv(col), dv(deltas), dv(ts), dv(rays_a),
                                         *got, thr)
    # the sigma gradient is a difference of large terms: compare against its own scale
    scale_s = rb[0].abs().max().item() + 1e-12
    assert (gb[0].cpu() - rb[0]).abs().max().item() <= 2e-5 * scale_s
    # the compacted list of gradient-carrying samples (the ones composited before each ray's early stop)
    from google_nerf_b200 import _lib as L
    alive = torch.full((N,), -1, dtype=torch.int32, device=DEV); cnt = torch.full((4,), 77, dtype=torch.int32, device=DEV)
    ds2 = torch.empty(N, device=DEV); dc2 = torch.empty(N, 3, device=DEV)
    keep = [dv(v).contiguous() for v in (*grads, sig, col, deltas, ts)] + [dv(rays_a)] + [v.contiguous() for v in got]
    L.call("b2n_composite_train_bw", *[L.ptr(v) for v in keep], thr, n, L.ptr(ds2), L.ptr(dc2), L.ptr(alive), L.ptr(cnt))
    mask, idx, incl, *_ = mods["R"]._composite_terms(sig, col, deltas, ts, rays_a, thr)
    want = torch.sort(idx[incl])[0]
    k = int(cnt[0].item())
    assert k == want.numel() and torch.equal(torch.sort(alive[:k].cpu().long())[0], want)
    assert torch.equal(ds2, gb[0]) and torch.equal(dc2, gb[1])
    # a = 1 - exp(-x) cancels for small x, so one ulp of expf shows up as ~1e-7 absolute in w = a*T
    torch.testing.assert_close(gb[1].cpu(), rb[1], rtol=1e-5, atol=1e-6)


def test_composite_loss_fused_equals_separate(mods, scene05):
    """b2n_composite_loss_fwbw (the trainer's one-launch path) against composite_train_fw -> nerf_loss -> composite_train_bw."""
    from google_nerf_b200 import _lib as L
    rays_a, xyzs, dirs, deltas, ts, _ = _packed(mods, scene05)
    N, n = xyzs.shape[0], rays_a.shape[0]
    g = torch.Generator().manual_seed(12)
    dv = lambda t: t.to(DEV).contiguous()
    sig, col = dv(torch.rand(N, generator=g) * 40.0), dv(torch.rand(N, 3, generator=g))
    target = dv(torch.rand(n, 3, generator=g))
    deltas, ts, rays_a = dv(deltas), dv(ts), dv(rays_a)
    thr, bg, lam, ls = 1e-4, 0.37, 1e-3, 128.0
    P = L.ptr
    e = lambda *sh: torch.empty(*sh, device=DEV)
    op, dp, dp2, rgb = mods["vren"].composite_train_fw(sig, col, deltas, ts, rays_a, thr)
    rgb_out, loss, d_rgb, d_op = e(n, 3), torch.zeros(1, device=DEV), e(n, 3), e(n)
    L.call("b2n_nerf_loss_fwbw", P(rgb), P(op), P(target), n, bg, lam, ls, P(rgb_out), P(loss), P(d_rgb), P(d_op), None)
    zeros = torch.zeros(n, device=DEV)
    ds, dc = mods["vren"].composite_train_bw(d_op, zeros, zeros, d_rgb, sig, col, deltas, ts, rays_a, op, dp, dp2, rgb, thr)
    op2, dp2_, rgb_out2, loss2, ds2, dc2 = e(n), e(n), e(n, 3), torch.full((1,), 5.0, device=DEV), e(N), e(N, 3)
    alive = torch.empty(N, dtype=torch.int32, device=DEV); cnt = torch.full((4,), 9, dtype=torch.int32, device=DEV)
    L.call("b2n_composite_loss_fwbw", P(sig), P(col), P(deltas), P(ts), P(rays_a), P(target), thr, n, bg, lam, ls,
           P(op2), P(dp2_), P(rgb_out2), P(loss2), P(ds2), P(dc2), P(alive), P(cnt), None)
    # the device-side loss scale (b2n_hyper.loss_scale) overrides the argument
    ls_dev = torch.tensor([ls], device=DEV); ds3, dc3 = e(N), e(N, 3)
    L.call("b2n_composite_loss_fwbw", P(sig), P(col), P(deltas), P(ts), P(rays_a), P(target), thr, n, bg, lam, 1.0,
           P(op2), P(dp2_), P(rgb_out2), P(loss2), P(ds3), P(dc3), P(alive), P(cnt), P(ls_dev))
    assert torch.equal(ds3, ds2) and torch.equal(dc3, dc2)
    torch.testing.assert_close(op2, op, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(dp2_, dp, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(rgb_out2, rgb_out, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(loss2, loss, rtol=1e-5, atol=0)
    scale_s = ds.abs().max().item() + 1e-12
    assert (ds2 - ds).abs().max().item() <= 1e-5 * scale_s
    torch.testing.assert_close(dc2, dc, rtol=1e-5, atol=1e-7 * ls)
    # the alive list is the set of samples composited before each ray's early stop
    mask, idx, incl, *_ = mods["R"]._composite_terms(sig.cpu(), col.cpu(), deltas.cpu(), ts.cpu(), rays_a.cpu(), thr)
    want = torch.sort(idx[incl])[0]
    k = int(cnt[0].item())
    assert k == want.numel() and torch.equal(torch.sort(alive[:k].cpu().long())[0], want)
    # and the loss against the torch oracle (losses.py:32-40)
    from oracle import ngp_ref
    lref = ngp_ref.nerf_loss({"rgb": rgb_out.cpu().double(), "opacity": op.cpu().double()}, target.cpu().double(), lam)
    assert abs(float(lref) - float(loss2.item())) <= 1e-5 * abs(float(lref))


def test_composite_uniform_slab_closed_form(mods):
    # one ray, constant sigma over n equal steps: opacity = 1 - exp(-sigma * L)
    n, sigma, dt = 200, 3.0, 0.01
    rays_a = torch.tensor([[0, 0, n]], dtype=torch.int64, device=DEV)
    ts = (torch.arange(n, device=DEV) * dt + 0.5).float()
    out = mods["vren"].composite_train_fw(torch.full((n,), sigma, device=DEV), torch.ones(n, 3, device=DEV),
                                          torch.full((n,), dt, device=DEV), ts, rays_a, 0.0)
    expect = 1 - np.exp(-sigma * n * dt)
    assert abs(out[0].item() - expect) < 1e-5 * expect
    assert abs(out[3][0, 1].item() - expect) < 1e-5 * expect


def test_rays_from_indices(mods):
    """b2n_rays_from_indices against get_rays (datasets/ray_utils.py:152-175) on the synthetic cameras."""
    from google_nerf_b200 import _lib as L, synthetic as syn
    K = syn.intrinsics(64, 48); dirs = syn.directions(64, 48, K).to(DEV).contiguous()
    poses = syn.hemisphere_poses(7).to(DEV).contiguous()
    g = torch.Generator().manual_seed(5)
    n = 3001
    ii = torch.randint(7, (n,), generator=g).to(DEV); pi = torch.randint(64 * 48, (n,), generator=g).to(DEV)
    ro, rd = torch.empty(n, 3, device=DEV), torch.empty(n, 3, device=DEV)
    L.call("b2n_rays_from_indices", L.ptr(dirs), L.ptr(poses), L.ptr(ii), L.ptr(pi), n, L.ptr(ro), L.ptr(rd))
    ro_ref, rd_ref = syn.get_rays(dirs[pi].cpu().double(), poses[ii].cpu().double())
    assert torch.equal(ro.cpu(), poses[ii][:, :, 3].cpu())
    torch.testing.assert_close(rd.cpu().double(), rd_ref, rtol=0, atol=2e-7)


# ------------------------------------------------------------------------------------------------ encodings
def _layout_pair(mods, scale=0.5, log2_T=19, L=16):
    b = np.exp(np.log(2048 * scale / 16) / (L - 1))
    lay = mods["tcnn"].hashgrid_layout(L, 2, log2_T, 16, b)
    ref = mods["T"].hashgrid_layout(L, 2, log2_T, 16, b)
    return lay, ref


@pytest.mark.parametrize("scale,log2_T", [(0.5, 19), (16.0, 22), (0.5, 14)])
def test_hashgrid_layout_matches_oracle(mods, scale, log2_T):
    lay, ref = _layout_pair(mods, scale, log2_T)
    assert list(lay.resolution[:16]) == ref["resolutions"]
    assert list(lay.offset[:17]) == ref["offsets"]
    assert lay.n_params == ref["n_params"]
    np.testing.assert_array_equal(np.array(lay.scale[:16], dtype=np.float32), np.array(ref["scales"], dtype=np.float32))


@pytest.mark.parametrize("log2_T", [19, 14])
def test_hashgrid_fw_bw(mods, log2_T):
    lay, ref = _layout_pair(mods, 0.5, log2_T)
    g = torch.Generator().manual_seed(4)
    n = 5000
    x = torch.rand(n, 3, generator=g)
    x[:4] = torch.tensor([[0., 0., 0.], [1., 1., 1.], [0.5, 0.5, 0.5], [1.0, 0.0, 0.3]])   # box corners / faces
    table = (torch.rand(ref["n_params"], generator=g) * 2 - 1) * 0.5
    tab = table.clone().requires_grad_(True)
    enc_ref = mods["T"].hashgrid_forward(x, tab, ref)
    t16 = table.to(DEV).half()
    enc = mods["tcnn"].hashgrid_fw(x.to(DEV), t16, lay)
    torch.testing.assert_close(enc.cpu().float(), enc_ref.float(), rtol=1e-3, atol=1e-3)
    # exact index/weight check against the oracle's integer index function via a one-hot style table
    dy = torch.randn(n, 32, generator=g).half()
    enc_ref.backward(dy.float().to(enc_ref.dtype))
    grad = torch.zeros(ref["n_params"], device=DEV)
    mods["tcnn"].hashgrid_bw(x.to(DEV), dy.to(DEV), lay, grad, 1.0)
    scale_g = tab.grad.abs().max().item()
    assert (grad.cpu() - tab.grad).abs().max().item() <= 1e-3 * scale_g
    # compacted form: rows of dy indexed through sample_idx give the same table gradient
    from google_nerf_b200 import _lib as L
    perm = torch.randperm(n, generator=g)
    xs, dys = x.to(DEV), dy[perm].contiguous().to(DEV)            # row i of dys belongs to position x[perm[i]]
    idx = perm.to(torch.int32).to(DEV)
    grad2 = torch.zeros_like(grad)
    L.call("b2n_hashgrid_bw", L.ptr(xs), L.ptr(dys), 32, lay, n, None, 1.0, L.ptr(grad2), L.ptr(idx))
    assert (grad2 - grad).abs().max().item() <= 1e-4 * scale_g


def test_frequency_and_sh(mods):
    g = torch.Generator().manual_seed(6)
    x = torch.rand(3000, 3, generator=g)
    ref = mods["T"].frequency_forward(x)
    got = mods["tcnn"].frequency_fw(x.to(DEV), 12)
    assert got.shape == (3000, 80)
    torch.testing.assert_close(got.cpu().float(), ref.float(), rtol=1e-3, atol=1e-3)
    d = torch.randn(3000, 3, generator=g); d = d / d.norm(dim=-1, keepdim=True)
    ref = mods["T"].sh4_forward((d + 1) / 2)
    got = mods["tcnn"].sh4_fw(((d + 1) / 2).to(DEV))
    torch.testing.assert_close(got.cpu().float(), ref.float(), rtol=1e-3, atol=1e-3)
    got2 = mods["tcnn"].sh4_fw((d * 3.7).to(DEV), normalize=True)
    torch.testing.assert_close(got2.cpu().float(), ref.float(), rtol=1e-3, atol=2e-3)


# ------------------------------------------------------------------------------------------------ MLP
@pytest.mark.parametrize("n_in,n_out,n_hidden,act", [(32, 16, 1, "None"), (80, 16, 1, "None"), (32, 3, 2, "Sigmoid")])
@pytest.mark.parametrize("n", [1, 127, 1000])
def test_mlp_fw_bw(mods, n_in, n_out, n_hidden, act, n):
    T, tc = mods["T"], mods["tcnn"]
    g = torch.Generator().manual_seed(8)
    shapes = T.mlp_layout(n_in, n_out, 64, n_hidden)
    params = T.xavier_uniform_(torch.zeros(T.mlp_n_params(shapes)), shapes, g)
    x = (torch.randn(n, n_in, generator=g) * 0.5).half()
    p = params.clone().requires_grad_(True)
    xr = x.float().requires_grad_(True)
    ref, hid = T.mlp_forward(xr, p, shapes, 16, act, return_hidden=True)
    w16 = params.to(DEV).half()
    out, hidden = tc.mlp_fw(x.to(DEV), w16, shapes[0][1], n_hidden, 1 if act == "Sigmoid" else 0)
    torch.testing.assert_close(out.cpu().float(), ref.float(), rtol=2e-3, atol=2e-3)
    for l in range(n_hidden):
        torch.testing.assert_close(hidden[l].cpu().float(), hid[l].float(), rtol=2e-3, atol=2e-3)
    dy = torch.zeros(n, 16)
    dy[:, :n_out] = torch.randn(n, n_out, generator=g)
    dy = dy.half()
    ref.float().backward(dy.float())
    grad = torch.zeros_like(params, device=DEV)
    din = tc.mlp_bw(dy.to(DEV), x.to(DEV), w16, shapes[0][1], n_hidden, 1 if act == "Sigmoid" else 0, hidden, out,
                    grad, 1.0, True)
    gs = p.grad.abs().max().item() + 1e-12
    assert (grad.cpu() - p.grad).abs().max().item() <= 3e-3 * gs
    ds = xr.grad.abs().max().item() + 1e-12
    assert (din.cpu().float()[:, :n_in] - xr.grad).abs().max().item() <= 3e-3 * ds


# ------------------------------------------------------------------------------------------------ optimiser etc.
def test_adam_matches_reference_formula(mods):
    from google_nerf_b200 import _lib as L
    from oracle.ngp_ref import AdamRef
    g = torch.Generator().manual_seed(9)
    n = 4096 + 8
    p0 = torch.randn(n, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = AdamRef([pr], lr=1e-2, eps=1e-15)
    p = p0.to(DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    h = torch.empty(n, dtype=torch.float16, device=DEV)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * 10 ** float(torch.randint(-6, 1, (1,), generator=g))
        pr.grad = gr.clone(); opt.step()
        gd = (gr * 128).to(DEV)
        L.call("b2n_adam_step", L.ptr(p), L.ptr(gd), L.ptr(m), L.ptr(v), L.ptr(h), n, 1e-2, 0.9, 0.999, 1e-15,
               1.0 / 128, step, None, 1)
        assert float(gd.abs().max()) == 0.0                      # gradient buffer is zeroed by the step
        torch.testing.assert_close(p.cpu(), pr.detach(), rtol=1e-5, atol=1e-6)
        assert torch.equal(h.cpu(), p.cpu().half())


def test_grid_threshold_and_loss(mods):
    from google_nerf_b200 import _lib as L
    g = torch.Generator().manual_seed(10)
    grid = torch.randn(3 * 128 ** 3 // 8, generator=g)
    ws = torch.empty(3, dtype=torch.float64, device=DEV); stats = torch.empty(3, device=DEV)
    gd = grid.to(DEV)
    L.call("b2n_grid_threshold", L.ptr(gd), gd.numel(), 0.5, L.ptr(ws), L.ptr(stats))
    mean = grid[grid > 0].double().mean().item()
    assert abs(stats[1].item() - mean) < 1e-6 and abs(stats[0].item() - min(mean, 0.5)) < 1e-6
    assert int(stats[2].item()) == int((grid > 0).sum())
    # loss + gradients vs autograd of the reference formula
    from oracle.ngp_ref import nerf_loss
    n = 1000
    rgb = torch.rand(n, 3, generator=g, requires_grad=True); op = torch.rand(n, generator=g, requires_grad=True)
    tgt = torch.rand(n, 3, generator=g)
    loss = nerf_loss({"rgb": rgb + 1.0 * (1 - op)[:, None], "opacity": op}, tgt)
    loss.backward()
    out = torch.empty(n, 3, device=DEV); lossd = torch.zeros(1, device=DEV)
    drgb = torch.empty(n, 3, device=DEV); dop = torch.empty(n, device=DEV)
    rgb_d, op_d, tgt_d = rgb.detach().to(DEV), op.detach().to(DEV), tgt.to(DEV)      # keep the buffers alive
    L.call("b2n_nerf_loss_fwbw", L.ptr(rgb_d), L.ptr(op_d), L.ptr(tgt_d), n, 1.0, 1e-3, 1.0, L.ptr(out), L.ptr(lossd),
           L.ptr(drgb), L.ptr(dop), None)
    assert abs(lossd.item() - loss.item()) < 1e-5 * abs(loss.item())
    torch.testing.assert_close(drgb.cpu(), rgb.grad, rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(dop.cpu(), op.grad, rtol=1e-4, atol=1e-9)
