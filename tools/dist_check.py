"""Multi-GPU parity of the sharded paths against the single-process oracle.  One process per GPU:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

Checks (bit-exact unless noted): distributed boxblur through the generated-function ABI with the
halo rows received over the library's NCCL communicator (tests/test_160.cpp semantics); row-sharded
heat2d with a grouped halo exchange; the scatter+all-gather broadcast; row-sharded sgemm with B
broadcast from rank 0 (integer operands, exact on the TF32 path)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402  (checker only)
from tiramisu_b200 import _lib, api, dist as tdist  # noqa: E402


def main():
    os.environ["NCCL_DEBUG"] = os.environ.get("TB200_NCCL_DEBUG", "WARN")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdist.comm_init_from_torch()
    L = _lib.lib()
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)
    ok = True

    # ---- 1. distributed box blur (uint32), halo rows come from rank+1 inside the call -------------------
    rows_per, cols = 37, 1000
    full = (np.arange((rows_per * world + 2) * (cols + 2), dtype=np.uint32).reshape(rows_per * world + 2, cols + 2) * 7) % 1013
    want = oracle.boxblur_u32(full)[: rows_per * world]
    loc = np.zeros((rows_per + 2, cols + 2), np.uint32)
    loc[:rows_per] = full[rank * rows_per:(rank + 1) * rows_per]
    if rank == world - 1:
        loc[rows_per:] = full[world * rows_per:]            # the global bottom rows belong to the last rank
    out = np.zeros((rows_per, cols), np.uint32)
    d_in, d_out = torch.from_numpy(loc.view(np.int32)).cuda(), torch.from_numpy(out.view(np.int32)).cuda()
    b_in, b_out = api.Buffer(d_in), api.Buffer(d_out)
    b_in.raw.type.code = b_out.raw.type.code = 1
    _lib.check(L.boxblur(b_in.ptr(), b_out.ptr()), "boxblur")
    got = d_out.cpu().numpy().view(np.uint32)
    good = np.array_equal(got, want[rank * rows_per:(rank + 1) * rows_per])
    ok &= good
    print(f"[rank {rank}] boxblur distributed: {'ok' if good else 'MISMATCH'}", flush=True)

    # ---- 2. row-sharded heat2d with a grouped halo exchange ------------------------------------------------
    H, W = 64 * world + 5, 1024
    grid = np.random.default_rng(1).random((H, W), dtype=np.float32)
    wantg = oracle.heat2d(grid, "cpu")
    r0, r1 = tdist.row_shard(H, world, rank)
    top, bot = (1 if rank > 0 else 0), (1 if rank < world - 1 else 0)
    rows = r1 - r0
    d_loc = torch.zeros((top + rows + bot, W), device="cuda")
    d_loc[top:top + rows] = torch.from_numpy(grid[r0:r1]).cuda()
    d_o = torch.full((rows, W), float("nan"), device="cuda")
    row_b = W * 4
    own = d_loc.data_ptr() + top * row_b
    sp = (C.c_int32 * 2)(rank - 1 if top else -1, rank + 1 if bot else -1)
    sb = (C.c_void_p * 2)(own, own + (rows - 1) * row_b)
    sn = (C.c_uint64 * 2)(row_b, row_b)
    rp = (C.c_int32 * 2)(rank - 1 if top else -1, rank + 1 if bot else -1)
    rb = (C.c_void_p * 2)(d_loc.data_ptr(), own + rows * row_b)
    rn = (C.c_uint64 * 2)(row_b, row_b)
    _lib.check(L.tiramisu_b200_comm_exchange(2, sp, sb, sn, rp, rb, rn))
    _lib.check(L.tiramisu_b200_heat2d(d_loc.data_ptr() - (r0 - top) * row_b, d_o.data_ptr() - r0 * row_b, H, W, 0, r0, r1))
    api.synchronize()
    good = np.array_equal(d_o.cpu().numpy(), wantg[r0:r1])
    ok &= good
    print(f"[rank {rank}] heat2d row-sharded: {'ok' if good else 'MISMATCH'}", flush=True)

    # ---- 3. scatter + all-gather broadcast ------------------------------------------------------------------
    nbytes = 64 * 1024 * 1024 + 12345 * 4
    src = torch.arange(nbytes // 4, dtype=torch.int32, device="cuda") * 3 + 1
    buf = src.clone() if rank == 0 else torch.zeros_like(src)
    _lib.check(L.tiramisu_b200_comm_broadcast_large(buf.data_ptr(), nbytes, 0))
    api.synchronize()
    good = bool(torch.equal(buf, src))
    ok &= good
    print(f"[rank {rank}] broadcast_large: {'ok' if good else 'MISMATCH'}", flush=True)

    # ---- 4. row-sharded sgemm, B broadcast from rank 0 -----------------------------------------------------
    n = 1024
    rng = np.random.default_rng(7)
    A = rng.integers(-4, 5, (n, n)).astype(np.float32)
    B = rng.integers(-4, 5, (n, n)).astype(np.float32)
    Cm = rng.integers(-4, 5, (n, n)).astype(np.float32)
    wantc = 3.0 * (A.astype(np.float64) @ B.astype(np.float64)) + 2.0 * Cm
    a0, a1 = tdist.row_shard(n, world, rank)
    dA, dC = torch.from_numpy(A[a0:a1]).cuda(), torch.from_numpy(Cm[a0:a1]).cuda()
    dB = torch.from_numpy(B).cuda() if rank == 0 else torch.zeros((n, n), device="cuda")
    _lib.check(L.tiramisu_b200_comm_broadcast(dB.data_ptr(), n * n * 4, 0))
    api.cublas_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), a1 - a0, n, n, 3.0, 2.0)
    api.synchronize()
    good = np.array_equal(dC.cpu().numpy().astype(np.float64), wantc[a0:a1])
    ok &= good
    print(f"[rank {rank}] sgemm row-sharded: {'ok' if good else 'MISMATCH'}", flush=True)

    # ---- 4b. the same with B's arrival fused in (tiramisu_b200_sgemm_bcast_b): panel-major staging in symmetric memory,
    # DMA-engine scatter + all-gather, GEMM gated per tile on the panels' arrival; several epochs on one block, shards
    # large enough for the persistent kernels (in-kernel gate) and small ones (stream-side gate, one launch per panel)
    for Mb, Nb, Kb, panels in ((1024, 1024, 1024, 4), (2048, 2048, 512, 1), (512 * world, 4096, 1024, 8), (1024 * world, 8192, 256, 2),
                               (1024 * world, 8192, 288, 32), (16 * world + 3, 1024, 64, 2)):
        rngb = np.random.default_rng(70 + Nb + Kb)
        symB = tdist.SymmetricB(Kb, Nb)
        a0, a1 = tdist.row_shard(Mb, world, rank)
        for ep in range(3):
            A = rngb.integers(-3, 4, (Mb, Kb)).astype(np.float32)
            B = rngb.integers(-3, 4, (Kb, Nb)).astype(np.float32)
            Cm = rngb.integers(-3, 4, (Mb, Nb)).astype(np.float32)
            wantc = 3.0 * (A[a0:a1] @ B) + 2.0 * Cm[a0:a1]           # small integers: exact in fp32 (and on the TF32 path)
            dA, dC = torch.from_numpy(A[a0:a1]).cuda(), torch.from_numpy(Cm[a0:a1]).cuda()
            dB = torch.from_numpy(B).cuda() if rank == 0 else None
            torch.cuda.synchronize()
            if ep == 2:                                              # data motion alone (plain bytes), then a plain GEMM on the local copy
                if rank == 0:
                    symB.tensor().copy_(dB)
                    torch.cuda.synchronize()
                symB.broadcast(0, min(panels, 16))
                api.cublas_sgemm(dA.data_ptr(), symB.data_ptr(), dC.data_ptr(), a1 - a0, Nb, Kb, 3.0, 2.0)
            else:
                symB.sgemm(dB.data_ptr() if rank == 0 else None, dA.data_ptr(), dC.data_ptr(), a1 - a0, 3.0, 2.0, 0, panels)
            api.synchronize()
            good = np.array_equal(dC.cpu().numpy(), wantc)
            ok &= good
            print(f"[rank {rank}] sgemm + fused B broadcast {Mb}x{Nb}x{Kb} panels={panels} epoch={ep + 1}: {'ok' if good else 'MISMATCH'}", flush=True)
            torch.distributed.barrier()
        symB.free()

    # ---- 5. peer-memory halos: no exchange, the sweep reads the neighbours' rows over NVLink ---------------
    for Hp in (48 * world + 7, 49 * world, 13 * world + 1):      # row counts that leave short last row blocks
        Wp = 2048
        gridp = np.random.default_rng(21).random((Hp, Wp), dtype=np.float32)
        rows_of = lambda r: tdist.row_shard(Hp, world, r)
        p0, p1 = rows_of(rank)
        sym = tdist.SymmetricRows(max(rows_of(r)[1] - rows_of(r)[0] for r in range(world)), Wp, n_buffers=2)
        rowb = Wp * 4
        mine = np.ascontiguousarray(gridp[p0:p1])

        def h2d(dst, arr):          # blocking copies of the runtime (tiramisu_cuda_memcpy_to_device / _to_host)
            _lib.check(L.tiramisu_cuda_memcpy_to_device(C.c_void_p(dst), arr.ctypes.data_as(C.c_void_p), arr.nbytes))

        def d2h(arr, src):
            _lib.check(L.tiramisu_cuda_memcpy_to_host(arr.ctypes.data_as(C.c_void_p), C.c_void_p(src), arr.nbytes))

        h2d(sym.buffer(0), mine)
        torch.cuda.synchronize()
        torch.distributed.barrier()
        # (a) single sweep, forms of configs[1]: heat2d + fp32 blur
        d_o = torch.full((p1 - p0, Wp), float("nan"), device="cuda")
        h = sym.halo(0, p0, p1, rows_of)
        _lib.check(L.tiramisu_b200_heat2d_peer(sym.buffer(0) - p0 * rowb, d_o.data_ptr() - p0 * rowb, Hp, Wp, 0, p0, p1, C.byref(h)))
        api.synchronize()
        good = np.array_equal(d_o.cpu().numpy(), oracle.heat2d(gridp, "cpu")[p0:p1])
        ok &= good
        print(f"[rank {rank}] heat2d peer halos H={Hp}: {'ok' if good else 'MISMATCH'}", flush=True)
        q0, q1 = min(p0, Hp - 2), min(p1, Hp - 2)
        d_b = torch.full((max(q1 - q0, 1), Wp - 2), float("nan"), device="cuda")
        _lib.check(L.tiramisu_b200_blur_f32_peer(sym.buffer(0) - p0 * rowb, d_b.data_ptr() - q0 * (Wp - 2) * 4, Hp, Wp, q0, q1, C.byref(h)))
        api.synchronize()
        good = np.array_equal(d_b.cpu().numpy()[: q1 - q0], oracle.blur_f32(gridp)[q0:q1])
        ok &= good
        print(f"[rank {rank}] blur_f32 peer halos H={Hp}: {'ok' if good else 'MISMATCH'}", flush=True)
        # (b) iterated even/odd Jacobi (heat2d_dist_tiramisu.cpp:88-153), ITERS sweeps back to back on the
        # stream, neighbours synchronised only through the flag words inside the sweep kernels
        iters = 12
        h2d(sym.buffer(1), mine)                                         # both buffers start as the input (borders persist)
        torch.cuda.synchronize()
        torch.distributed.barrier()
        for t in range(iters):
            a, b = t % 2, (t + 1) % 2
            h = sym.halo(a, p0, p1, rows_of, sync=True, wait_value=t, signal_value=t + 1)
            _lib.check(L.tiramisu_b200_heat2d_peer(sym.buffer(a) - p0 * rowb, sym.buffer(b) - p0 * rowb, Hp, Wp, 2, p0, p1, C.byref(h)))
        api.synchronize()
        ev, od = gridp.copy(), gridp.copy()
        for t in range(iters):
            if t % 2 == 0:
                oracle.heat2d_step(ev, od)
            else:
                oracle.heat2d_step(od, ev)
        want_j = ev if iters % 2 == 0 else od
        got_j = np.empty((p1 - p0, Wp), np.float32)
        d2h(got_j, sym.buffer(iters % 2))
        good = np.array_equal(got_j, want_j[p0:p1])
        ok &= good
        print(f"[rank {rank}] heat2d_dist H={Hp} {iters} Jacobi sweeps, flag handshake in the kernels: {'ok' if good else 'MISMATCH'}", flush=True)
        torch.distributed.barrier()
        sym.free()

    # ---- 6. heat2d_dist_tiramisu through the generated-function ABI (host buffers, one call per rank) -------
    # global grid [world*rows + 2][cols]; rank q passes local rows [q*rows, q*rows + rows + 2) of both buffers
    # (rows 1..rows owned, 0 and rows+1 halos), as wrapper_heat2d_dist.cpp:53-66 does
    for rows_d, cols_d, iters_d, protocol in ((37, 2048, 4, 0), (16, 1024, 6, 0), (37, 2048, 1, 1), (24, 512, 3, 1)):
        Hd = world * rows_d + 2
        ge = np.random.default_rng(31).random((Hd, cols_d), dtype=np.float32)
        go = np.random.default_rng(32).random((Hd, cols_d), dtype=np.float32)
        le = np.ascontiguousarray(ge[rank * rows_d: rank * rows_d + rows_d + 2])
        lo_ = np.ascontiguousarray(go[rank * rows_d: rank * rows_d + rows_d + 2])
        for rep in range(2):                                      # twice: the symmetric block and flag epochs are reused
            e_in, o_in = le.copy(), lo_.copy()
            api.heat2d_dist_tiramisu(e_in, o_in, rows_per_rank=rows_d, iters=iters_d, protocol=protocol)
        we, wo = ge.copy(), go.copy()
        for s_ in range(iters_d):
            src_is_odd = (s_ % 2 == 0) if protocol == 0 else (s_ % 2 == 1)
            if src_is_odd:
                oracle.heat2d_step(wo, we)
            else:
                oracle.heat2d_step(we, wo)
        sl = slice(rank * rows_d, rank * rows_d + rows_d + 2)
        good = np.array_equal(e_in, we[sl]) and np.array_equal(o_in, wo[sl])
        ok &= good
        print(f"[rank {rank}] heat2d_dist_tiramisu rows/rank={rows_d} cols={cols_d} iters={iters_d} protocol={protocol}: "
              f"{'ok' if good else 'MISMATCH'}", flush=True)

    # ---- 7. tiramisu_MPI_* shims: Irecv-first bidirectional neighbour exchange (one NCCL group at Wait), host
    # and device buffers — the pattern the reference's distributed tests use (tests/test_98.cpp ... 102)
    up, dn = rank - 1, rank + 1
    n_el = 1000
    for on_dev in (False, True):
        mine_np = (np.arange(n_el, dtype=np.float32) + 1000 * rank)
        from_up_np, from_dn_np = np.full(n_el, -1, np.float32), np.full(n_el, -1, np.float32)
        if on_dev:
            mine_t, fu_t, fd_t = torch.from_numpy(mine_np).cuda(), torch.from_numpy(from_up_np).cuda(), torch.from_numpy(from_dn_np).cuda()
            pm, pu, pd = mine_t.data_ptr(), fu_t.data_ptr(), fd_t.data_ptr()
        else:
            pm, pu, pd = mine_np.ctypes.data, from_up_np.ctypes.data, from_dn_np.ctypes.data
        req = (C.c_long * 4)()
        if up >= 0:
            L.tiramisu_MPI_Irecv_f32(n_el, up, 0, C.c_void_p(pu), req)
        if dn < world:
            L.tiramisu_MPI_Irecv_f32(n_el, dn, 0, C.c_void_p(pd), req)
        if up >= 0:
            L.tiramisu_MPI_Isend_f32(n_el, up, 0, C.c_void_p(pm), req)
        if dn < world:
            L.tiramisu_MPI_Isend_f32(n_el, dn, 0, C.c_void_p(pm), req)
        L.tiramisu_MPI_Wait(req)
        api.synchronize()
        if on_dev:
            from_up_np, from_dn_np = fu_t.cpu().numpy(), fd_t.cpu().numpy()
        good = True
        if up >= 0:
            good &= np.array_equal(from_up_np, np.arange(n_el, dtype=np.float32) + 1000 * up)
        if dn < world:
            good &= np.array_equal(from_dn_np, np.arange(n_el, dtype=np.float32) + 1000 * dn)
        ok &= good
        print(f"[rank {rank}] tiramisu_MPI Irecv/Isend/Wait neighbour exchange ({'device' if on_dev else 'host'} buffers): "
              f"{'ok' if good else 'MISMATCH'}", flush=True)

    flag = torch.tensor([1 if ok else 0], device="cuda")
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    tdist.comm_destroy()
    torch.distributed.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_OK" if int(flag.item()) == 1 else "DIST_CHECK_FAILED", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
