"""NVLink peer memory for one node (one process per GPU): host side of csrc/peer.cu.

Every rank allocates ONE block through ``b2n_peer_alloc`` (plain cudaMalloc, so that a CUDA-IPC handle exists), the
64-byte handles travel through ``torch.distributed`` once at start-up, and every rank maps every other rank's block.
Named regions of the block are handed out as torch tensors (zero-copy, via ``__cuda_array_interface__``) and as
per-rank pointer tables for the kernels.  After the rendezvous there is no communicator on the data path: the barrier
and the fused reduce + Adam + broadcast kernel are ordinary launches on the caller's stream (graph-capturable).

Replaces the DistributedDataParallel gradient all-reduce of ngp_pl/train.py:197-208 for the flat parameter vector.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L

_ALIGN = 256
FLAG_BYTES = 256                      # >= 16 u32 epoch slots (csrc/peer.cu: PEER_MAX_WORLD)


class _Raw:
    """Exposes a raw device pointer to torch.as_tensor."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerBlock:
    """regions: {name: (dtype, numel)}.  Collective over `group` (all ranks call it with the same regions)."""

    def __init__(self, regions, device, group=None, barrier_timeout_s=20.0):
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise ValueError("peer exchange supports up to 16 ranks of one node")
        self.timeout = float(barrier_timeout_s)
        self.layout, off = {}, FLAG_BYTES
        for name, (dtype, numel) in regions.items():
            nbytes = numel * torch.empty(0, dtype=dtype).element_size()
            self.layout[name] = (off, dtype, numel)
            off += -(-nbytes // _ALIGN) * _ALIGN
        self.nbytes = off
        # Every step that can fail locally (allocation, IPC export, mapping a peer) is followed by a collective vote, so
        # that either all ranks end up with a working block or all of them raise -- never a rank left waiting.
        self.local, self.base, err = None, [], None
        with torch.cuda.device(device):
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            try:
                L.call_nostream("b2n_peer_alloc", self.nbytes, C.byref(ptr), handle)
                self.local = ptr.value
            except RuntimeError as e:
                err = e
            handles = [None] * self.world
            dist.all_gather_object(handles, (handle.raw, self.nbytes) if err is None else None, group=group)
            if any(h is None for h in handles):
                self._abort(None)
                raise RuntimeError(f"peer memory unavailable: allocation / IPC export failed on a rank ({err})")
            for r, (h, nb) in enumerate(handles):
                if err is not None:
                    break
                if nb != self.nbytes:
                    err = RuntimeError("peer blocks differ in size across ranks")
                elif r == self.rank:
                    self.base.append(self.local)
                else:
                    q = C.c_void_p()
                    try:
                        L.call_nostream("b2n_peer_open", C.create_string_buffer(h, 64), C.byref(q))
                        self.base.append(q.value)
                    except RuntimeError as e:
                        err = e
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self._abort(group)
                raise RuntimeError(f"peer memory unavailable: mapping a peer's block failed on a rank ({err})")
        self._raw = torch.as_tensor(_Raw(self.local, self.nbytes), device=device)      # uint8 view of the local block
        self.state = torch.zeros(2, dtype=torch.int32, device=device)                  # {epoch, sticky error}
        self._flags = (C.c_void_p * self.world)(*self.base)
        self._tables = {}
        self.closed = False
        dist.barrier(group=group)                                                      # every mapping exists

    def _abort(self, group):
        """Undo a partly built block (all ranks call it together when `group` is given)."""
        with torch.cuda.device(self.device):
            for r, b in enumerate(self.base):
                if r != self.rank:
                    try:
                        L.call_nostream("b2n_peer_close", b)
                    except RuntimeError:
                        pass
            if group is not None:
                dist.barrier(group=group)                     # nobody frees a block a peer still has mapped
            if self.local is not None:
                try:
                    L.call_nostream("b2n_peer_free", self.local)
                except RuntimeError:
                    pass
        self.base, self.local = [], None

    def tensor(self, name):
        """The local region as a torch tensor."""
        off, dtype, numel = self.layout[name]
        es = torch.empty(0, dtype=dtype).element_size()
        return self._raw[off:off + numel * es].view(dtype)

    def table(self, name):
        """Host array of `world` device pointers: the region `name` on every rank (argument of the peer kernels)."""
        if name not in self._tables:
            off = self.layout[name][0]
            self._tables[name] = (C.c_void_p * self.world)(*[b + off for b in self.base])
        return self._tables[name]

    def barrier(self):
        """Device-side barrier on the current stream (no host synchronisation)."""
        L.call("b2n_peer_barrier", self._flags, self.rank, self.world, L.ptr(self.state), self.timeout)

    def check(self):
        """Raise if a barrier timed out (host sync; call between steps, not inside them)."""
        err = int(self.state[1].item())
        if err:
            raise RuntimeError(f"peer barrier timed out waiting for rank {err - 1}")

    def close(self):
        """Collective: unmap the peers' blocks, then free the local one."""
        if self.closed:
            return
        self.closed = True
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for r, b in enumerate(self.base):
                if r != self.rank:
                    L.call_nostream("b2n_peer_close", b)
            dist.barrier(group=self.group)
            self._raw = None
            L.call_nostream("b2n_peer_free", self.local)
