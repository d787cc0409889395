"""Debug: phase durations inside field_mlp_bw_kernel (needs a build with B2N_BW_TRACE=1)."""
import ctypes, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2N_BW_TRACE"] = "1"
subprocess.check_call(["touch", os.path.join(ROOT, "google-nerf_b200", "csrc", "field_tc.cu")])
import __graft_entry__ as g; g.build()
import torch
from google_nerf_b200 import _lib as L
n = 400000
dev = "cuda"
torch.manual_seed(0)
enc = (torch.randn(n, 32, device=dev) * 0.5).half(); dirs = torch.randn(n, 3, device=dev)
ws = (torch.randn(3072, device=dev) * 0.2).half(); wr = (torch.randn(7168, device=dev) * 0.2).half()
image = torch.empty(10240, dtype=torch.float16, device=dev)
L.call("b2n_field_pack_weights", L.ptr(ws), L.ptr(wr), L.ptr(image))
sig = torch.empty(n, device=dev); rgb = torch.empty(n, 3, device=dev)
hs = torch.empty(n, 64, dtype=torch.float16, device=dev); h = torch.empty(n, 16, dtype=torch.float16, device=dev)
hr = torch.empty(2, n, 64, dtype=torch.float16, device=dev)
L.call("b2n_field_mlp_fw", L.ptr(enc), L.ptr(dirs), L.ptr(image), n, None, L.ptr(sig), L.ptr(rgb), L.ptr(hs), L.ptr(h), L.ptr(hr))
dsig = torch.randn(n, device=dev); drgb = torch.randn(n, 3, device=dev)
denc = torch.empty(n, 32, dtype=torch.float16, device=dev); gs = torch.zeros(3072, device=dev); gr = torch.zeros(7168, device=dev)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call("b2n_field_mlp_bw", L.ptr(dsig), L.ptr(drgb), L.ptr(enc), L.ptr(dirs), L.ptr(image), n, None, L.ptr(rgb), L.ptr(hs), L.ptr(h), L.ptr(hr), 1.0, L.ptr(denc), L.ptr(gs), L.ptr(gr))
    e1.record(); torch.cuda.synchronize()
    print("bw kernel %.1f us" % (e0.elapsed_time(e1) * 1e3))
out = (ctypes.c_longlong * (64 * 16))()
lib = L.lib(); lib.b2n_debug_bw_trace.argtypes = [ctypes.c_void_p]
assert lib.b2n_debug_bw_trace(out) == 0
import numpy as np
t = np.array(out[:]).reshape(64, 16)
names = ["0 tile start->prologue done", "1 ->sync", "2 ->A mma waited", "3 ->A epilogue", "4 ->sync", "5 ->B issue+prefetch", "6 ->B mma waited", "7 ->B epilogue", "8 ->cp wait", "9 ->sync", "10 -> end of tile (C,D,E)"]
valid = t[:, 11] > 0
tt = t[valid][1:6]
order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 14, 15, 11]
lab = {0: "tile start", 1: "prologue done", 2: "sync", 3: "A mma waited", 4: "A epilogue", 5: "sync", 6: "B issue+prefetch",
       7: "B mma waited", 8: "B epilogue", 9: "cp wait", 10: "sync", 12: "C issue+prefetch+mma waited", 13: "C epilogue+cp wait+sync",
       14: "D issue+prefetch+mma waited", 15: "D epilogue+cp wait+sync", 11: "E whole step + store + syncs"}
for a, b in zip(order[:-1], order[1:]):
    print("%-34s %8.0f cycles" % (lab[b], (tt[:, b] - tt[:, a]).mean()))
print("tile total %.0f cycles; tiles traced %d" % ((tt[:, 11] - tt[:, 0]).mean(), valid.sum()))
