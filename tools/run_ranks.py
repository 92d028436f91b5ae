"""Start N copies of a program as ranks 0..N-1 of one node — what `mpirun -np N` does for the reference's MPI
wrappers — with the torchrun-style environment that tiramisu_MPI_init() reads (RANK, WORLD_SIZE, LOCAL_RANK,
MASTER_ADDR, MASTER_PORT).

    python tools/run_ranks.py N program [args...]

Exit status: 0 if every rank exits 0.  A rank that fails takes the others down (they would wait in a collective)."""
import os
import subprocess
import sys
import time


def main():
    n = int(sys.argv[1])
    cmd = sys.argv[2:]
    port = os.environ.get("MASTER_PORT", "29811")
    procs = []
    for r in range(n):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(n), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen(cmd, env=env))
    rc = 0
    deadline = time.time() + float(os.environ.get("RUN_RANKS_TIMEOUT", "540"))
    while procs:
        for p in list(procs):
            code = p.poll()
            if code is None:
                continue
            procs.remove(p)
            if code != 0:
                rc = rc or code
                for q in procs:
                    q.kill()
        if time.time() > deadline:
            for q in procs:
                q.kill()
            rc = rc or 124
            break
        time.sleep(0.05)
    sys.exit(rc)


if __name__ == "__main__":
    main()
