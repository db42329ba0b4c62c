"""PeerComm — the small-message all-reduce of the multi-GPU PPO update over NVLink peer memory (include/agx.h, csrc/agx_comm.cuh).

One process per GPU (torchrun).  Every rank allocates a region in its own HBM through libagx, the CUDA-IPC handles are exchanged
once through `torch.distributed` (the plumbing), every process maps its peers' regions, and from then on a collective is one
libagx kernel per rank on the caller's stream — capturable in a CUDA graph, no NCCL call, no host sync.  Reference counterpart:
the dist.all_reduce calls of lib/agent/a2c_base.py:293-309 and lib/agent/a2c_continuous.py:112-123."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi


class PeerComm:
    def __init__(self, rank, world, max_bytes, device):
        if not 1 <= world <= _capi.AGX_COMM_MAX_RANKS:
            raise ValueError(f"PeerComm: world {world} outside 1..{_capi.AGX_COMM_MAX_RANKS} (one node, one process per GPU)")
        self._lib = _capi.load()
        self.rank, self.world, self.device = int(rank), int(world), torch.device(device)
        self.slot_bytes = (int(max_bytes) + 255) // 256 * 256
        nbytes = int(self._lib.agx_comm_region_bytes(self.world, self.slot_bytes))
        handle = (C.c_ubyte * _capi.AGX_IPC_HANDLE_BYTES)()
        own = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(self._lib.agx_comm_alloc(nbytes, C.byref(own), handle), "agx_comm_alloc")
        self._own, self._peers = own, {}
        comm = _capi.AgxComm()
        comm.rank, comm.world, comm.slot_bytes = self.rank, self.world, self.slot_bytes
        comm.region[self.rank] = own.value
        if self.world > 1:
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
            everyone = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(everyone, mine)
            with torch.cuda.device(self.device):
                for r, h in enumerate(everyone):
                    if r == self.rank:
                        continue
                    raw = (C.c_ubyte * _capi.AGX_IPC_HANDLE_BYTES)(*h.cpu().tolist())
                    ptr = C.c_void_p()
                    _capi.check(self._lib.agx_comm_open(raw, C.byref(ptr)), f"agx_comm_open(rank {r})")
                    self._peers[r] = ptr
                    comm.region[r] = ptr.value
            dist.barrier()  # nobody pushes before every mapping exists
        self.c = comm

    def all_reduce(self, t):
        """In-place SUM over the ranks of a contiguous float32 / float64 tensor on this rank's device."""
        if t.dtype not in (torch.float32, torch.float64) or not t.is_contiguous() or t.device != self.device:
            raise ValueError("PeerComm.all_reduce: contiguous float32/float64 tensor on the communicator's device expected")
        nbytes = t.numel() * t.element_size()
        if nbytes > self.slot_bytes:
            raise ValueError(f"PeerComm.all_reduce: {nbytes} B message exceeds the {self.slot_bytes} B slot")
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        dt = _capi.AGX_F64 if t.dtype == torch.float64 else _capi.AGX_F32
        _capi.check(self._lib.agx_comm_allreduce(C.byref(self.c), t.data_ptr(), t.numel(), dt, st), "agx_comm_allreduce")
        return t

    def status(self):
        """(collectives completed, error word) — synchronises the current stream."""
        seq, err = C.c_uint64(), C.c_uint64()
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(self._lib.agx_comm_status(C.byref(self.c), C.byref(seq), C.byref(err), st), "agx_comm_status")
        return int(seq.value), int(err.value)

    def check(self):
        seq, err = self.status()
        if err:
            raise RuntimeError(f"PeerComm: collective #{err & ((1 << 56) - 1)} timed out waiting for rank {(err >> 56) - 1} "
                               f"(this is rank {self.rank}); the ranks issued different sequences of collectives or a peer died")
        return seq

    def close(self):
        if getattr(self, "_own", None) is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier()  # every rank is done with every region before anything is unmapped
        for ptr in self._peers.values():
            self._lib.agx_comm_close(ptr)
        self._lib.agx_comm_free(self._own)
        self._own, self._peers = None, {}


def make_local_group(world, max_bytes, device):
    """`world` communicators inside ONE process on one GPU (regions are plain allocations, no IPC): what the single-GPU tests use to
    exercise the protocol — issue every rank's collective on its own stream."""
    lib = _capi.load()
    slot = (int(max_bytes) + 255) // 256 * 256
    nbytes = int(lib.agx_comm_region_bytes(world, slot))
    regions = []
    with torch.cuda.device(device):
        for _ in range(world):
            p = C.c_void_p()
            _capi.check(lib.agx_comm_alloc(nbytes, C.byref(p), None), "agx_comm_alloc")
            regions.append(p)
    comms = []
    for r in range(world):
        pc = PeerComm.__new__(PeerComm)
        pc._lib, pc.rank, pc.world, pc.device, pc.slot_bytes = lib, r, world, torch.device(device), slot
        c = _capi.AgxComm()
        c.rank, c.world, c.slot_bytes = r, world, slot
        for q in range(world):
            c.region[q] = regions[q].value
        pc.c, pc._own, pc._peers = c, regions[r], {}
        comms.append(pc)
    return comms
