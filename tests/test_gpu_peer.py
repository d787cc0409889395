"""Single-GPU coverage of the NVLink peer-memory entry points (csrc/peer.cu) with world = 1: the block allocator and its
IPC handle, the flag barrier, the fp16 gradient pack and the fused reduce + Adam + broadcast kernel against
b2n_adam_step.  The multi-rank behaviour is checked by tests/dist_nccl_check.py under torchrun."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


@pytest.fixture()
def block(built_lib):
    from google_nerf_b200 import _lib as L
    n = 40960
    nbytes = 256 + 4 * n + 2 * n + 2 * n
    ptr, handle = C.c_void_p(), C.create_string_buffer(64)
    L.call_nostream("b2n_peer_alloc", nbytes, C.byref(ptr), handle)
    raw = torch.as_tensor(_Raw(ptr.value, nbytes), device=DEV)
    yield L, n, ptr.value, raw, handle
    torch.cuda.synchronize()
    del raw
    L.call_nostream("b2n_peer_free", ptr.value)


def test_peer_block_is_zeroed_and_has_a_handle(block):
    L, n, base, raw, handle = block
    assert int(raw.sum().item()) == 0 and any(handle.raw)            # zero-filled, non-trivial IPC handle
    with pytest.raises(RuntimeError):                               # a handle cannot be opened in the exporting process
        q = C.c_void_p()
        L.call_nostream("b2n_peer_open", C.create_string_buffer(handle.raw, 64), C.byref(q))


def test_peer_barrier_world_1_counts_epochs(block):
    L, n, base, raw, _ = block
    state = torch.zeros(2, dtype=torch.int32, device=DEV)
    flags = (C.c_void_p * 1)(base)
    for _ in range(5):
        L.call("b2n_peer_barrier", flags, 0, 1, L.ptr(state), 1.0)
    torch.cuda.synchronize()
    assert state.tolist() == [5, 0]                                  # five epochs, no timeout
    assert int(raw[:4].view(torch.int32).item()) == 5                # own flag slot carries the last epoch
    with pytest.raises(RuntimeError):
        L.call("b2n_peer_barrier", flags, 1, 1, L.ptr(state), 1.0)  # rank outside the world


def test_peer_barrier_times_out_instead_of_hanging(block):
    """world = 2 with a peer that never arrives (its flag block is this same, silent, memory): the barrier gives up after
    the timeout and records who was missing."""
    L, n, base, raw, _ = block
    state = torch.zeros(2, dtype=torch.int32, device=DEV)
    other = torch.zeros(64, dtype=torch.int32, device=DEV)           # stands in for rank 1's flag block
    flags = (C.c_void_p * 2)(base, other.data_ptr())
    L.call("b2n_peer_barrier", flags, 0, 2, L.ptr(state), 0.05)
    torch.cuda.synchronize()
    assert state.tolist() == [1, 2]                                  # epoch advanced, error = 1 + rank 1
    assert int(other[0].item()) == 1                                 # our announcement reached the peer's slot 0


@pytest.mark.parametrize("half", [False, True])
def test_adam_step_peer_world_1_equals_adam_step(block, half):
    L, n, base, raw, _ = block
    g = torch.Generator().manual_seed(3)
    grad = raw[256:256 + 4 * n].view(torch.float32)
    h_out = raw[256 + 4 * n:256 + 6 * n].view(torch.float16)
    g16 = raw[256 + 6 * n:256 + 8 * n].view(torch.float16)
    grad.copy_((torch.randn(n, generator=g) * 3).to(DEV))
    grad[5] = 1e6                                                    # saturates in the fp16 wire format
    p0 = torch.randn(n, generator=g).to(DEV)
    m0, v0 = torch.rand(n, generator=g).to(DEV) * 0.1, torch.rand(n, generator=g).to(DEV) * 0.01
    lo, hi = 1024, n - 2048                                          # fp16 range: the "table"; outside stays fp32
    # reference: plain Adam on the gradient the owner will see
    g_seen = grad.clone()
    if half:
        g_seen[lo:hi] = g_seen[lo:hi].clamp(-65504, 65504).half().float()
    p1, m1, v1 = p0.clone(), m0.clone(), v0.clone()
    h1 = torch.empty(n, dtype=torch.float16, device=DEV)
    L.call("b2n_adam_step", L.ptr(p1), L.ptr(g_seen), L.ptr(m1), L.ptr(v1), L.ptr(h1), n, 1e-2, 0.9, 0.999, 1e-15,
           1.0 / 128, 7, None, 1)
    # fused path, shard = the middle half of the vector
    first, cnt = n // 4, n // 2
    p2, m2, v2 = p0[first:first + cnt].clone(), m0[first:first + cnt].clone(), v0[first:first + cnt].clone()
    if half:
        L.call("b2n_grad_pack_half", L.ptr(grad), L.ptr(g16), lo, hi)
        assert float(grad[lo:hi].abs().max()) == 0.0 and float(grad[:lo].abs().max()) > 0      # packed part cleared only
        assert float(g16[5 + 0].float()) != float("inf")
    gp = (C.c_void_p * 1)(grad.data_ptr()); g16p = (C.c_void_p * 1)(g16.data_ptr()); hp = (C.c_void_p * 1)(h_out.data_ptr())
    L.call("b2n_adam_step_peer", L.ptr(p2), L.ptr(m2), L.ptr(v2), gp, g16p if half else None, lo, hi if half else lo, hp, 1,
           first, cnt, 1e-2, 0.9, 0.999, 1e-15, 1.0 / 128, 7, None, None, None)
    torch.cuda.synchronize()
    sl = slice(first, first + cnt)
    assert torch.equal(p2, p1[sl]) and torch.equal(m2, m1[sl]) and torch.equal(v2, v1[sl])
    assert torch.equal(h_out[sl], h1[sl])
    assert float(h_out[:first].float().abs().max()) == 0.0           # nothing written outside the shard
    with pytest.raises(RuntimeError):
        L.call("b2n_adam_step_peer", L.ptr(p2), L.ptr(m2), L.ptr(v2), gp, None, 0, 0, hp, 1, first + 2, cnt, 1e-2, 0.9,
               0.999, 1e-15, 1.0, 1, None, None, None)               # shard bounds must be multiples of 4
    # device control block: found_inf skips the update on both kernels, the scaler halves the scale and counts the skip;
    # a sticky barrier error turns the peer kernel into a no-op
    import struct
    f2i = lambda f: struct.unpack("i", struct.pack("f", f))[0]
    hyper = torch.tensor([f2i(1e-2), 7, 1, 0, f2i(128.0), 5, 0, 0], dtype=torch.int32, device=DEV)
    hyp = (C.c_void_p * 1)(hyper.data_ptr())
    grad.copy_((torch.randn(n, generator=g) * 3).to(DEV))
    p3, m3, v3 = p2.clone(), m2.clone(), v2.clone()
    L.call("b2n_adam_step_peer", L.ptr(p3), L.ptr(m3), L.ptr(v3), gp, None, 0, 0, hp, 1, first, cnt, 0.0, 0.9, 0.999,
           1e-15, 1.0, 0, L.ptr(hyper), hyp, None)
    assert torch.equal(p3, p2) and torch.equal(m3, m2)               # skipped
    p4, m4, v4, g4 = p1.clone(), m1.clone(), v1.clone(), grad.clone()
    L.call("b2n_adam_step", L.ptr(p4), L.ptr(g4), L.ptr(m4), L.ptr(v4), None, n, 0.0, 0.9, 0.999, 1e-15, 1.0, 0, L.ptr(hyper), 1)
    assert torch.equal(p4, p1) and torch.equal(m4, m1) and float(g4.abs().max()) == 0.0   # skipped, gradient cleared
    L.call("b2n_scaler_update", L.ptr(hyper), hyp, 1)
    assert hyper.tolist()[2:6] == [1, 1, f2i(64.0), 0]               # found_inf left for the caller, skipped, scale / 2
    hyper[2] = 0
    # clean step: effective step = step - skipped = 6, gradients divided by the device loss scale (64)
    p5, m5, v5, g5 = p1.clone(), m1.clone(), v1.clone(), grad.clone()
    L.call("b2n_adam_step", L.ptr(p5), L.ptr(g5), L.ptr(m5), L.ptr(v5), None, n, 0.0, 0.9, 0.999, 1e-15, 1.0, 0, L.ptr(hyper), 1)
    p6, m6, v6, g6 = p1.clone(), m1.clone(), v1.clone(), grad.clone()
    L.call("b2n_adam_step", L.ptr(p6), L.ptr(g6), L.ptr(m6), L.ptr(v6), None, n, 1e-2, 0.9, 0.999, 1e-15, 1.0 / 64, 6, None, 1)
    assert torch.equal(p5, p6) and torch.equal(v5, v6)
    L.call("b2n_scaler_update", L.ptr(hyper), None, 1)
    assert hyper.tolist()[2:6] == [0, 1, f2i(64.0), 0]               # growth_interval 0: the scale never grows
    state = torch.tensor([3, 2], dtype=torch.int32, device=DEV)      # sticky error: rank 1 timed out
    L.call("b2n_adam_step_peer", L.ptr(p3), L.ptr(m3), L.ptr(v3), gp, None, 0, 0, hp, 1, first, cnt, 0.0, 0.9, 0.999,
           1e-15, 1.0, 0, L.ptr(hyper), hyp, L.ptr(state))
    assert torch.equal(p3, p2)
