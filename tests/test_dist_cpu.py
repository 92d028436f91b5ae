"""CPU (gloo, world_size 2 and 3) tests of the N>1 host logic: row sharding, halo plans and the
halo exchange.  The per-shard compute is the oracle here (no GPU in this container); the same
plan/exchange code drives the CUDA kernels on the GPU box (tests/test_gpu_dist.py)."""
import os
import socket

import numpy as np
import pytest

from tiramisu_b200 import dist as tdist


def test_row_shard_covers_and_balances():
    for n, w in ((16384, 8), (10, 3), (3842, 4), (7, 7), (100, 1)):
        spans = [tdist.row_shard(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_stencil_plan_halos():
    p0 = tdist.stencil_plan(100, 4, 0, 1, 1)
    p1 = tdist.stencil_plan(100, 4, 1, 1, 1)
    p3 = tdist.stencil_plan(100, 4, 3, 1, 1)
    assert (p0.top, p0.bottom) == (0, 1) and (p1.top, p1.bottom) == (1, 1) and (p3.top, p3.bottom) == (1, 0)
    b = tdist.stencil_plan(1600, 16, 5, 0, 2)          # blurautodist: 2 rows from rank+1 only
    assert (b.top, b.bottom, b.rows) == (0, 2, 100)
    assert tdist.stencil_plan(100, 1, 0, 1, 1).local_rows == 100
    with pytest.raises(ValueError):
        tdist.stencil_plan(4, 4, 1, 0, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, q):
    import torch
    import torch.distributed as dist
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        grid = rng.random((H, W), dtype=np.float32)          # every rank derives the same global grid
        res = {}
        # heat2d: 1 halo row each way
        plan = tdist.stencil_plan(H, world, rank, 1, 1)
        local = torch.zeros((plan.local_rows, W), dtype=torch.float32)
        local[plan.top:plan.top + plan.rows] = torch.from_numpy(grid[plan.r0:plan.r1])
        tdist.exchange_halos(local, plan, 1, 1)
        out = oracle.heat2d(local.numpy(), "cpu")[plan.top:plan.top + plan.rows]
        if rank == 0:
            out[0] = 0                                       # global borders: first/last grid rows
        if rank == world - 1:
            out[-1] = 0
        res["heat"] = (plan.r0, out)
        # blur: 2 rows from below
        bplan = tdist.stencil_plan(H, world, rank, 0, 2)
        blocal = torch.zeros((bplan.local_rows, W), dtype=torch.float32)
        blocal[:bplan.rows] = torch.from_numpy(grid[bplan.r0:bplan.r1])
        tdist.exchange_halos(blocal, bplan, 0, 2)
        res["blur"] = (bplan.r0, oracle.blur_f32(blocal.numpy()))
        # uint32 box blur, the reference's distributed test (2-row halo from rank+1)
        img = (np.arange(H)[:, None] + np.arange(W + 2)[None, :]).astype(np.uint32)
        ulocal = torch.zeros((bplan.rows + 2, W + 2), dtype=torch.int32)
        ulocal[:bplan.rows] = torch.from_numpy(img[bplan.r0:bplan.r1].view(np.int32))
        uplan = tdist.HaloPlan(rank, world, bplan.r0, bplan.r1, 0, 2 if rank < world - 1 else 0)
        tdist.exchange_halos(ulocal[:uplan.local_rows], uplan, 0, 2)
        if rank == world - 1:                                # last rank fills its own halo (wrapper_test_160.cpp:19-28)
            ulocal[bplan.rows:] = torch.from_numpy((np.arange(H, H + 2)[:, None] + np.arange(W + 2)[None, :]).astype(np.int32))
        res["box"] = (bplan.r0, oracle.boxblur_u32(ulocal.numpy().view(np.uint32)))
        q.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_stencils_match_single_process(world):
    import torch.multiprocessing as mp
    from oracle import oracle
    H, W = 50, 36
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    grid = rng.random((H, W), dtype=np.float32)
    heat = np.concatenate([got[r]["heat"][1] for r in range(world)])
    assert np.array_equal(heat, oracle.heat2d(grid, "cpu"))
    blur = np.concatenate([got[r]["blur"][1] for r in range(world)])
    assert np.array_equal(blur, oracle.blur_f32(grid))
    img = (np.arange(H + 2)[:, None] + np.arange(W + 2)[None, :]).astype(np.uint32)
    box = np.concatenate([got[r]["box"][1] for r in range(world)])
    assert np.array_equal(box, oracle.boxblur_u32(img))


def test_peer_halo_descriptors():
    """dist.SymmetricRows.halo(): where each rank finds its neighbours' boundary rows and flag words inside the
    symmetric blocks (pure address arithmetic; the allocation itself needs GPUs and is covered by tools/dist_check.py)."""
    world, width, rows_max, nbuf = 3, 16, 10, 2
    bases = [0x10000000, 0x20000000, 0x30000000]              # each rank's block as mapped into "this" process
    H = 28
    rows_of = lambda r: tdist.row_shard(H, world, r)          # noqa: E731   (10, 9, 9 rows)
    row_b = width * 4
    for rank in range(world):
        s = object.__new__(tdist.SymmetricRows)
        s.world, s.rank, s.width, s.n_buffers = world, rank, width, nbuf
        s.block_bytes = ((rows_max * width * 4 + 255) // 256) * 256
        s.bytes = s.block_bytes * nbuf + tdist.SymmetricRows.FLAG_BYTES
        s.local, s.peers = bases[rank], list(bases)
        r0, r1 = rows_of(rank)
        for buf in range(nbuf):
            h = s.halo(buf, r0, r1, rows_of)
            assert (h.lo, h.hi) == (r0, r1) and not h.flags and not h.sig_up and not h.sig_dn
            if rank == 0:
                assert not h.up
            else:
                u0, u1 = rows_of(rank - 1)                    # row r0-1 = the upper neighbour's last row
                assert h.up == bases[rank - 1] + buf * s.block_bytes + (u1 - 1 - u0) * row_b and u1 == r0
            if rank == world - 1:
                assert not h.dn
            else:
                assert h.dn == bases[rank + 1] + buf * s.block_bytes          # row r1 = the lower neighbour's first row
        h = s.halo(1, r0, r1, rows_of, sync=True, wait_value=4, signal_value=5)
        flags = lambda r: bases[r] + nbuf * s.block_bytes                     # noqa: E731
        assert h.flags == flags(rank) and (h.wait_value, h.signal_value) == (4, 5)
        # I am my upper neighbour's lower neighbour: I write its flags[1]; and my lower neighbour's flags[0]
        assert (h.sig_up or 0) == (flags(rank - 1) + 4 if rank > 0 else 0)
        assert (h.sig_dn or 0) == (flags(rank + 1) if rank < world - 1 else 0)


# ---- peer-halo protocol, emulated on the CPU: shared memory plays the symmetric (CUDA IPC) block ----------------
def _jacobi_worker(rank, world, port, H, W, iters, q):
    """What a rank does on the GPU box (tools/dist_check.py section 5b, heat2d_dist.cu), with numpy in place of
    the sweep kernel: blocks of all ranks live in shared mappings (`multiprocessing.shared_memory` stands in for
    cudaIpcOpenMemHandle), descriptors come from dist.SymmetricRows.halo(sync=True), a sweep waits on its own flag
    words, reads the neighbours' boundary rows IN PLACE and publishes its sweep count into the neighbours' flags."""
    import ctypes as C
    import time
    from multiprocessing import shared_memory

    import torch.distributed as dist
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shms = []
    try:
        rows_of = lambda r: tdist.row_shard(H, world, r)      # noqa: E731
        rows_max = max(rows_of(r)[1] - rows_of(r)[0] for r in range(world))
        s = object.__new__(tdist.SymmetricRows)
        s.world, s.rank, s.width, s.n_buffers = world, rank, W, 2
        s.block_bytes = ((rows_max * W * 4 + 255) // 256) * 256
        s.bytes = s.block_bytes * 2 + tdist.SymmetricRows.FLAG_BYTES
        mine = shared_memory.SharedMemory(create=True, size=s.bytes)
        shms.append(mine)
        mine.buf[:] = bytes(s.bytes)
        names = [None] * world
        dist.all_gather_object(names, mine.name)              # the "IPC handles"
        maps = []
        for r in range(world):
            m = mine if r == rank else shared_memory.SharedMemory(name=names[r])
            if r != rank:
                shms.append(m)
            maps.append(m)
        base = [C.addressof(C.c_char.from_buffer(m.buf)) for m in maps]
        s.local, s.peers = base[rank], base

        def rows_at(addr, n):                                 # float rows at an absolute address inside some rank's block
            return np.ctypeslib.as_array((C.c_float * (n * W)).from_address(addr)).reshape(n, W)

        def word(addr):
            return C.c_int32.from_address(addr)

        r0, r1 = rows_of(rank)
        grid = np.random.default_rng(11).random((H, W), dtype=np.float32)
        for b in range(2):
            rows_at(s.buffer(b), r1 - r0)[:] = grid[r0:r1]
        dist.barrier()
        for t in range(iters):
            a, b = t % 2, (t + 1) % 2
            h = s.halo(a, r0, r1, rows_of, sync=True, wait_value=t, signal_value=t + 1)
            deadline = time.time() + 60
            for has, off in ((h.up, 0), (h.dn, 4)):           # halo_wait (stencil.cu): flags[0] / flags[1] >= wait_value
                while has and word(h.flags + off).value < h.wait_value:
                    assert time.time() < deadline, "handshake timed out"
                    time.sleep(0.0005)
            src = np.zeros((r1 - r0 + 2, W), np.float32)
            src[1:-1] = rows_at(s.buffer(a), r1 - r0)
            if h.up:
                src[0] = rows_at(h.up, 1)[0]                  # row r0-1, read in place from the upper neighbour
            if h.dn:
                src[-1] = rows_at(h.dn, 1)[0]                 # row r1, read in place from the lower neighbour
            dst = np.zeros_like(src)
            dst[1:-1] = rows_at(s.buffer(b), r1 - r0)
            oracle.heat2d_step(src, dst)                      # interior of `dst` only (Jacobi half-step)
            own = dst[1:-1].copy()
            if rank == 0:
                own[0] = rows_at(s.buffer(b), 1)[0]           # global border rows are not computed
            if rank == world - 1:
                own[-1] = rows_at(s.buffer(b), r1 - r0)[-1]
            rows_at(s.buffer(b), r1 - r0)[:] = own
            if h.sig_up:                                      # halo_signal: publish this rank's sweep count
                word(h.sig_up).value = h.signal_value
            if h.sig_dn:
                word(h.sig_dn).value = h.signal_value
        dist.barrier()
        q.put((rank, r0, rows_at(s.buffer(iters % 2), r1 - r0).copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()
        del s
        for m in shms:
            try:
                m.close()
            except BufferError:
                pass
        try:
            shms[0].unlink()
        except Exception:
            pass


@pytest.mark.parametrize("world", [2, 3])
def test_peer_halo_handshake_protocol_emulated(world):
    import torch.multiprocessing as mp
    from oracle import oracle
    H, W, iters = 23, 16, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_jacobi_worker, args=(r, world, port, H, W, iters, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    grid = np.random.default_rng(11).random((H, W), dtype=np.float32)
    ev, od = grid.copy(), grid.copy()
    for t in range(iters):
        if t % 2 == 0:
            oracle.heat2d_step(ev, od)
        else:
            oracle.heat2d_step(od, ev)
    want = ev if iters % 2 == 0 else od
    assert np.array_equal(np.concatenate([g[2] for g in got]), want)
