"""Row/batch sharding and halo exchange for the paths that shard — the host-side
logic that replaces tag_distribute_level + create_xfer/gen_communication in the reference
(src/tiramisu_core.cpp:599-628, 8778-8900, 9101-9295) with static owner-computes tables.

One process per GPU.  torch.distributed is plumbing only: NCCL send/recv (CUDA tensors) on the
GPU box, gloo (CPU tensors) in the CPU tests — the exchange code is the same.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from . import _lib


def row_shard(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition of [0, n_rows) (1-D block distribution of the outermost loop,
    src/tiramisu_core.cpp:8997-9044).  Remainder rows go to the lowest ranks."""
    base, rem = divmod(n_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class HaloPlan:
    """What rank `rank` must hold besides its own rows [r0, r1): `top` rows owned by rank-1 and
    `bottom` rows owned by rank+1 (0 at the ends of the grid)."""
    rank: int
    world: int
    r0: int
    r1: int
    top: int
    bottom: int

    @property
    def rows(self) -> int:
        return self.r1 - self.r0

    @property
    def local_rows(self) -> int:
        return self.top + self.rows + self.bottom


def stencil_plan(n_rows: int, world: int, rank: int, up: int, down: int) -> HaloPlan:
    """A stencil that reads rows [y-up, y+down] of its input: heat2d up=down=1
    (benchmarks/halide/heat2d_dist_tiramisu.cpp:133-153 exchanges 1 row each way), the blurs
    up=0, down=2 (benchmarks/automatic_comm/blurautodist_ref.cpp:42-49: 2 rows from rank+1 only).
    exchange sets = sender_owned ∩ (receiver_needed − receiver_owned)."""
    r0, r1 = row_shard(n_rows, world, rank)
    if r1 - r0 < max(up, down) and world > 1:
        raise ValueError("shard thinner than the stencil halo")
    return HaloPlan(rank, world, r0, r1, up if rank > 0 else 0, down if rank < world - 1 else 0)


def exchange_halos(local, plan: HaloPlan, up: int, down: int, group=None):
    """local: tensor [plan.local_rows, W] whose rows [top, top+rows) are owned.  Fills the halo rows
    from the neighbours: my first `down` owned rows go to rank-1 (its bottom halo), my last `up`
    owned rows go to rank+1 (its top halo).  One batched isend/irecv (a single NCCL group)."""
    import torch.distributed as dist
    if plan.world == 1:
        return
    ops = []
    own0 = plan.top
    own1 = plan.top + plan.rows
    if plan.rank > 0:
        if down:
            ops.append(dist.P2POp(dist.isend, local[own0:own0 + down], plan.rank - 1, group))
        if plan.top:
            ops.append(dist.P2POp(dist.irecv, local[0:plan.top], plan.rank - 1, group))
    if plan.rank < plan.world - 1:
        if up:
            ops.append(dist.P2POp(dist.isend, local[own1 - up:own1], plan.rank + 1, group))
        if plan.bottom:
            ops.append(dist.P2POp(dist.irecv, local[own1:own1 + plan.bottom], plan.rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def batch_shard(batch: int, world: int, rank: int) -> tuple[int, int]:
    """conv / LSTM shard over the batch dimension: independent units, no per-step communication."""
    return row_shard(batch, world, rank)


# ---- C-level communicator (tiramisu_MPI_* over NCCL), bootstrapped through torch.distributed ---------
def comm_init_from_torch(group=None) -> int:
    """Create the library's own NCCL communicator (used by boxblur's halo exchange and the
    tiramisu_MPI_* shims), shipping the unique id over an existing torch.distributed group."""
    import torch.distributed as dist
    L = _lib.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = (C.c_char * 128)()
    if rank == 0:
        _lib.check(L.tiramisu_b200_comm_unique_id(buf), "tiramisu_b200_comm_unique_id")
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0, group=group)
    idbuf = (C.c_char * 128).from_buffer_copy(box[0])
    _lib.check(L.tiramisu_b200_comm_init(rank, world, idbuf), "tiramisu_b200_comm_init")
    return rank


def comm_destroy() -> None:
    _lib.check(_lib.lib().tiramisu_b200_comm_destroy(), "tiramisu_b200_comm_destroy")


# ---- peer-memory row blocks: halo rows are read in place over NVLink ---------------------------------
class SymmetricRows:
    """Rank-local block of `rows_max` x `width` fp32 rows in symmetric device memory
    (tiramisu_b200_comm_alloc_symmetric): every rank's block is mapped into every process, so a
    stencil launch reads its neighbours' boundary rows directly (stencil.cu: Halo) — the fused form
    of heat2d_dist's / blurautodist's send/recv exchange.  `n_buffers` equal blocks are carved out
    of one allocation (even/odd Jacobi buffers); 4 flag words per rank follow them.

    Collective constructor: every rank must call it with the same arguments."""

    FLAG_BYTES = 256

    def __init__(self, rows_max: int, width: int, n_buffers: int = 1):
        L = _lib.lib()
        self.world = L.tiramisu_b200_comm_size()
        self.rank = L.tiramisu_b200_comm_rank()
        self.width = width
        self.block_bytes = ((rows_max * width * 4 + 255) // 256) * 256
        self.n_buffers = n_buffers
        self.bytes = self.block_bytes * n_buffers + self.FLAG_BYTES
        local = C.c_void_p()
        peers = (C.c_void_p * self.world)()
        _lib.check(L.tiramisu_b200_comm_alloc_symmetric(self.bytes, C.byref(local), peers), "comm_alloc_symmetric")
        self.local = local.value
        self.peers = [peers[r] for r in range(self.world)]

    def buffer(self, i: int = 0, rank: int | None = None) -> int:
        """Device address of buffer i of `rank`'s block (default: mine) as seen from this GPU."""
        base = self.local if rank is None else self.peers[rank]
        return base + i * self.block_bytes

    def flags(self, rank: int | None = None) -> int:
        base = self.local if rank is None else self.peers[rank]
        return base + self.n_buffers * self.block_bytes

    def halo(self, i: int, r0: int, r1: int, rows_of, sync: bool = False, wait_value: int = 0,
             signal_value: int = 0) -> "_lib.Halo":
        """Descriptor for a sweep over my rows [r0, r1) of buffer i.  rows_of(rank) -> (r0, r1) of a
        neighbour (to find its last row / first rows inside its block)."""
        h = _lib.Halo()
        row_b = self.width * 4
        h.lo, h.hi = r0, r1
        if self.rank > 0:
            u0, u1 = rows_of(self.rank - 1)
            h.up = self.buffer(i, self.rank - 1) + (u1 - 1 - u0) * row_b
        if self.rank < self.world - 1:
            h.dn = self.buffer(i, self.rank + 1)
        if sync:
            h.flags = self.flags()
            h.sig_up = self.flags(self.rank - 1) + 4 if self.rank > 0 else None
            h.sig_dn = self.flags(self.rank + 1) + 0 if self.rank < self.world - 1 else None
            h.wait_value, h.signal_value = wait_value, signal_value
        return h

    def free(self) -> None:
        if self.local:
            _lib.check(_lib.lib().tiramisu_b200_comm_free_symmetric(self.local), "comm_free_symmetric")
            self.local = None


# ---- B of the row-sharded sgemm: broadcast by the DMA engines, consumed by the GEMM as it arrives -------
class SymmetricB:
    """Staging area (2 x K * N fp32 in symmetric memory, used alternately, + 256 flag bytes) for tiramisu_b200_sgemm_bcast_b.
    Collective constructor / free.  `sgemm(B_ptr, A_ptr, C_ptr, rows, alpha, beta)` is the fused call: the root's
    row-major B is re-laid out panel-major into the staging area and travels in column panels (scatter + all-gather
    by peer copies, comm.cu) while every rank's GEMM consumes the panels as they arrive.  `broadcast()` moves the
    staging area's bytes as they are (plain buffer semantics; `tensor()` views them as a K x N matrix)."""

    FLAG_BYTES = 256

    def __init__(self, K: int, N: int):
        L = _lib.lib()
        self.world, self.rank = L.tiramisu_b200_comm_size(), L.tiramisu_b200_comm_rank()
        self.K, self.N = K, N
        self.data_bytes = ((K * N * 4 + 255) // 256) * 256
        local = C.c_void_p()
        self._peers = (C.c_void_p * self.world)()
        _lib.check(L.tiramisu_b200_comm_alloc_symmetric(2 * self.data_bytes + self.FLAG_BYTES, C.byref(local), self._peers),
                   "comm_alloc_symmetric")
        self.local = local.value
        self.epoch = 0

    def data_ptr(self) -> int:
        return self.local

    def tensor(self):
        """The staging area as a K x N torch view (no copy) — plumbing for tests and the bench."""
        import torch

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.K, self.N), "typestr": "<f4", "data": (self.local, False), "version": 2}
        return torch.as_tensor(raw, device="cuda")

    @staticmethod
    def default_panels(n: int, unit: int = 256, least: int = 1024) -> int:
        """Panels of >= `least` columns, multiples of `unit`, at most 16."""
        p = max(1, min(16, n // least))
        while p > 1 and (n % p or (n // p) % unit):
            p -= 1
        return p

    def sgemm_panels(self) -> int:
        """Column panels for the fused sgemm: as many as possible (the first panel's flight time is exposed) while one
        rank's slice of a panel stays >= 32 MiB — a DMA copy between IPC-mapped blocks costs ~11 us of set-up whatever its
        size (profiles/r02_p2p_patterns_8gpu.jsonl), 16 MiB copies already lose a third of the link."""
        by_size = (self.K * self.N * 4) // ((32 << 20) * max(1, self.world - 1))
        p = max(1, min(16, by_size, self.N // 1024 if self.N >= 1024 else 1))
        while p > 1 and (self.N % p or (self.N // p) % 256):
            p -= 1
        return p

    def broadcast(self, root: int = 0, panels: int | None = None) -> None:
        self.epoch += 1
        _lib.check(_lib.lib().tiramisu_b200_comm_broadcast_sym(self._peers, 0, self.K * self.N * 4, 2 * self.data_bytes, root, self.epoch,
                                                               panels or self.default_panels(self.K, 32, 2048)), "comm_broadcast_sym")

    def sgemm(self, B_ptr: int | None, A_ptr: int, C_ptr: int, rows: int, alpha: float, beta: float, root: int = 0,
              panels: int | None = None, ldA: int = 0, ldB: int = 0, ldC: int = 0) -> None:
        """B_ptr: the root's row-major K x N matrix (device address; ignored on the other ranks)."""
        if panels is None:
            panels = self.sgemm_panels()
        self.last_panels = panels
        self.epoch += 1
        _lib.check(_lib.lib().tiramisu_b200_sgemm_bcast_b(self._peers, 0, 2 * self.data_bytes, root, self.epoch, panels,
                                                          C.c_void_p(B_ptr if self.rank == root else None), ldB,
                                                          C.c_void_p(A_ptr), C.c_void_p(C_ptr), rows, self.N, self.K, alpha, beta, ldA, ldC),
                   "sgemm_bcast_b")

    def free(self) -> None:
        if self.local:
            _lib.check(_lib.lib().tiramisu_b200_comm_free_symmetric(self.local), "comm_free_symmetric")
            self.local = None
