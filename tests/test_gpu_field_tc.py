"""The fused tcgen05 field kernels (field_tc.cu) against the oracle's MLP chain and its autograd gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(n, seed):
    from oracle import tcnn_ref as T
    g = torch.Generator().manual_seed(seed)
    s_shapes, r_shapes = T.mlp_layout(32, 16, 64, 1), T.mlp_layout(32, 3, 64, 2)
    ps = T.xavier_uniform_(torch.zeros(T.mlp_n_params(s_shapes)), s_shapes, g)
    pr = T.xavier_uniform_(torch.zeros(T.mlp_n_params(r_shapes)), r_shapes, g)
    enc = (torch.randn(n, 32, generator=g) * 0.5).half()
    dirs = torch.randn(n, 3, generator=g) * 2.0
    return T, s_shapes, r_shapes, ps, pr, enc, dirs, g


def _oracle_forward(T, s_shapes, r_shapes, ps, pr, enc, dirs):
    h, hs = T.mlp_forward(enc, ps, s_shapes, 16, "None", return_hidden=True)
    sigma = torch.exp(h[:, 0].float())
    d = dirs / dirs.norm(dim=-1, keepdim=True)
    sh = T.sh4_forward((d + 1) / 2)
    rgb, hr = T.mlp_forward(torch.cat([sh, h], 1), pr, r_shapes, 3, "Sigmoid", return_hidden=True)
    return sigma, rgb, h, hs, hr


@pytest.mark.parametrize("n", [128, 1000, 40000])
def test_field_tc_forward(built_lib, n):
    L = built_lib
    T, s_shapes, r_shapes, ps, pr, enc, dirs, _ = _setup(n, 1)
    sigma, rgb, h, hs, hr = _oracle_forward(T, s_shapes, r_shapes, ps, pr, enc, dirs)
    ws, wr = ps.to(DEV).half(), pr.to(DEV).half()
    image = torch.empty(10240, dtype=torch.float16, device=DEV)
    L.call("b2n_field_pack_weights", L.ptr(ws), L.ptr(wr), L.ptr(image))
    enc_d, dirs_d = enc.to(DEV), dirs.to(DEV)
    sig_d = torch.empty(n, device=DEV); rgb_d = torch.empty(n, 3, device=DEV)
    hs_d = torch.empty(n, 64, dtype=torch.float16, device=DEV); h_d = torch.empty(n, 16, dtype=torch.float16, device=DEV)
    hr_d = torch.empty(2, n, 64, dtype=torch.float16, device=DEV)
    L.call("b2n_field_mlp_fw", L.ptr(enc_d), L.ptr(dirs_d), L.ptr(image), n, None, L.ptr(sig_d), L.ptr(rgb_d),
           L.ptr(hs_d), L.ptr(h_d), L.ptr(hr_d))
    torch.cuda.synchronize()
    torch.testing.assert_close(hs_d.cpu().float(), hs[0].float(), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(h_d.cpu().float(), h.float(), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(sig_d.cpu(), sigma, rtol=5e-3, atol=1e-4)
    torch.testing.assert_close(hr_d[0].cpu().float(), hr[0].float(), rtol=3e-3, atol=3e-3)
    torch.testing.assert_close(hr_d[1].cpu().float(), hr[1].float(), rtol=3e-3, atol=3e-3)
    torch.testing.assert_close(rgb_d.cpu(), rgb.float(), rtol=2e-3, atol=2e-3)
    # device-side count: only the first n_dev rows are touched
    sig2 = torch.full((n,), -7.0, device=DEV)
    nd = torch.tensor([n // 2, 0, 0, 0], dtype=torch.int32, device=DEV)
    L.call("b2n_field_mlp_fw", L.ptr(enc_d), L.ptr(dirs_d), L.ptr(image), n, L.ptr(nd), L.ptr(sig2), L.ptr(rgb_d),
           None, None, None)
    assert torch.equal(sig2[:n // 2], sig_d[:n // 2]) and float(sig2[n // 2:].max()) == -7.0


@pytest.mark.parametrize("n", [128, 5000])
def test_field_tc_backward(built_lib, n):
    L = built_lib
    T, s_shapes, r_shapes, ps, pr, enc, dirs, g = _setup(n, 2)
    ps_r, pr_r = ps.clone().requires_grad_(True), pr.clone().requires_grad_(True)
    enc_r = enc.float().requires_grad_(True)
    sigma, rgb, h, hs, hr = _oracle_forward(T, s_shapes, r_shapes, ps_r, pr_r, enc_r, dirs)
    dsig = torch.randn(n, generator=g) * 4.0
    drgb = torch.randn(n, 3, generator=g) * 4.0
    # the reference's TruncExp backward clamps the exponent; identical here because |h0| < 15
    (sigma * dsig).sum().backward(retain_graph=True)
    (rgb.float() * drgb).sum().backward()

    ws, wr = ps.to(DEV).half(), pr.to(DEV).half()
    image = torch.empty(10240, dtype=torch.float16, device=DEV)
    L.call("b2n_field_pack_weights", L.ptr(ws), L.ptr(wr), L.ptr(image))
    enc_d, dirs_d = enc.to(DEV), dirs.to(DEV)
    sig_d = torch.empty(n, device=DEV); rgb_d = torch.empty(n, 3, device=DEV)
    hs_d = torch.empty(n, 64, dtype=torch.float16, device=DEV); h_d = torch.empty(n, 16, dtype=torch.float16, device=DEV)
    hr_d = torch.empty(2, n, 64, dtype=torch.float16, device=DEV)
    L.call("b2n_field_mlp_fw", L.ptr(enc_d), L.ptr(dirs_d), L.ptr(image), n, None, L.ptr(sig_d), L.ptr(rgb_d),
           L.ptr(hs_d), L.ptr(h_d), L.ptr(hr_d))
    denc = torch.empty(n, 32, dtype=torch.float16, device=DEV)
    gs = torch.zeros(3072, device=DEV); gr = torch.zeros(7168, device=DEV)
    dsig_d, drgb_d = dsig.to(DEV), drgb.to(DEV)
    L.call("b2n_field_mlp_bw", L.ptr(dsig_d), L.ptr(drgb_d), L.ptr(enc_d), L.ptr(dirs_d), L.ptr(image), n, None,
           L.ptr(rgb_d), L.ptr(hs_d), L.ptr(h_d), L.ptr(hr_d), 1.0, L.ptr(denc), L.ptr(gs), L.ptr(gr), None, 0)
    torch.cuda.synchronize()
    for got, want, name in ((gr, pr_r.grad, "rgb weights"), (gs, ps_r.grad, "sigma weights"),
                            (denc.float(), enc_r.grad, "dL/denc")):
        sc = want.abs().max().item()
        err = (got.cpu() - want).abs().max().item()
        assert err <= 5e-3 * sc, (name, err, sc)
    # compacted form: a shuffled subset of the rows with non-zero upstream gradient gives the same parameter
    # gradients, and row i of dL/denc belongs to sample idx[i]
    keep = torch.rand(n, generator=g) < 0.6
    dsig_m, drgb_m = dsig * keep, drgb * keep[:, None]
    idx = torch.nonzero(keep)[:, 0]
    idx = idx[torch.randperm(idx.numel(), generator=g)].to(torch.int32).to(DEV)
    cnt = torch.tensor([idx.numel(), 0, 0, 0], dtype=torch.int32, device=DEV)
    denc2 = torch.zeros(n, 32, dtype=torch.float16, device=DEV); gs2 = torch.zeros_like(gs); gr2 = torch.zeros_like(gr)
    gs_f = torch.zeros_like(gs); gr_f = torch.zeros_like(gr); denc_f = torch.empty_like(denc)
    dsig_md, drgb_md = dsig_m.to(DEV), drgb_m.to(DEV)
    common = (L.ptr(dsig_md), L.ptr(drgb_md), L.ptr(enc_d), L.ptr(dirs_d), L.ptr(image), n)
    L.call("b2n_field_mlp_bw", *common, None, L.ptr(rgb_d), L.ptr(hs_d), L.ptr(h_d), L.ptr(hr_d), 1.0, L.ptr(denc_f),
           L.ptr(gs_f), L.ptr(gr_f), None, 0)
    L.call("b2n_field_mlp_bw", *common, L.ptr(cnt), L.ptr(rgb_d), L.ptr(hs_d), L.ptr(h_d), L.ptr(hr_d), 1.0, L.ptr(denc2),
           L.ptr(gs2), L.ptr(gr2), L.ptr(idx), n)
    torch.cuda.synchronize()
    assert (gr2 - gr_f).abs().max().item() <= 2e-3 * gr_f.abs().max().item()
    assert (gs2 - gs_f).abs().max().item() <= 2e-3 * gs_f.abs().max().item()
    torch.testing.assert_close(denc2[:idx.numel()].float(), denc_f[idx.long()].float(), rtol=1e-2, atol=1e-3 * denc_f.abs().max().item())
