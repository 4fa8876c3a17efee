"""Host side of the data-parallel exchange kernels (csrc/comm.cu): allocation and peer mapping of the symmetric
buffer, and thin wrappers of the C entry points.

The reference is single-GPU (main.py:1-3), so this layer is net-new; its contract is "N ranks on a batch sharded in
equal slices == one process on the global batch" (SURVEY.md section 8e).  PyTorch is plumbing only: its symmetric
memory allocator (cuMemCreate + peer / multicast mappings exchanged over the process group's store) hands out the
buffer and the pointers; every byte that crosses NVLink is moved by the kernels of comm.cu.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch

from . import _lib
from ._lib import STRUCTS


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class SymmComm:
    """one symmetric buffer per rank: [flat gradient | BN exchange slots | barrier flags] (layout: comm.cu)."""

    def __init__(self, n_grad: int, device: torch.device, peer_ptrs: List[int], rank: int, buf: torch.Tensor,
                 multicast_ptr: int = 0, handle=None):
        self.n, self.device, self.rank, self.world = int(n_grad), device, int(rank), len(peer_ptrs)
        self.buf, self.handle = buf, handle
        self.grad = buf[:self.n]
        self.multicast = bool(multicast_ptr)
        self.epoch = torch.zeros(2 + _lib.CONSTS["RL_COMM_MAXG"], device=device, dtype=torch.int32)
        c = STRUCTS["rl_comm"]()
        c.world, c.rank = self.world, self.rank
        for i, p in enumerate(peer_ptrs):
            c.peer[i] = p
        c.mc = multicast_ptr or None
        c.n = self.n
        c.epoch = self.epoch.data_ptr()
        self.c = c
        self.grid = int(os.environ.get("RALENET_COMM_GRID", "0"))

    @property
    def kind(self) -> str:
        return "symmetric memory, " + ("NVLS multimem.ld_reduce / multimem.st" if self.multicast
                                       else "peer loads / stores")

    # ---- construction -------------------------------------------------------------------------------------
    @staticmethod
    def nbytes(n_grad: int) -> int:
        return int(_lib.load().ralenet_comm_bytes(int(n_grad)))

    @classmethod
    def create(cls, n_grad: int, device: torch.device, group=None) -> "SymmComm":
        """collective over `group` (default: WORLD): allocate, rendezvous, zero, barrier."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        if dist.get_world_size(group) > _lib.CONSTS["RL_COMM_MAXW"]:
            raise _lib.RalenetError("symmetric-memory exchange supports at most 8 ranks (one NVSwitch domain)")
        floats = cls.nbytes(n_grad) // 4
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)      # needed by older torch, deprecated later
        except Exception:      # noqa: BLE001
            pass
        buf = symm_mem.empty(floats, dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(buf, group=group.group_name)
        buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)
        mc = 0
        if os.environ.get("RALENET_COMM_MULTIMEM", "1") != "0":
            try:
                mc = int(hdl.multicast_ptr or 0)
            except Exception:      # noqa: BLE001
                mc = 0
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if ptrs[hdl.rank] != buf.data_ptr():
            raise _lib.RalenetError("symmetric memory: local peer pointer does not match the buffer")
        return cls(n_grad, device, ptrs, hdl.rank, buf, mc, hdl)

    # ---- the three exchange steps ---------------------------------------------------------------------------
    def exchange(self, which: int, vals: torch.Tensor):
        """vals (<= 32 local floats) <- sum over ranks; which = 0 (forward statistics) / 1 (backward sums)."""
        _lib.check(_lib.load().ralenet_comm_exchange(ctypes.byref(self.c), int(which), vals.data_ptr(), vals.numel(),
                                                     _stream()))

    def allreduce_adam(self, p, m, v, step_dev, lr, betas, eps, gscale=1.0):
        _lib.check(_lib.load().ralenet_comm_allreduce_adam(ctypes.byref(self.c), p.data_ptr(), m.data_ptr(),
                                                           v.data_ptr(), lr, betas[0], betas[1], eps,
                                                           step_dev.data_ptr(), gscale, self.grid, _stream()))
