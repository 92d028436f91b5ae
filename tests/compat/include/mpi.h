// mpi.h stand-in (TEST INFRASTRUCTURE) so that the reference's MPI-style wrappers — e.g.
// benchmarks/halide/wrapper_heat2d_dist.cpp — compile unmodified and run one process per GPU under a
// torchrun-style environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT): the handful
// of MPI calls they make map onto the library's tiramisu_MPI_* runtime (NCCL, tiramisu_b200/csrc/comm.cu).
#pragma once
#include <cassert>
#include <cstdlib>

extern "C" {
int tiramisu_MPI_init(void);
void tiramisu_MPI_cleanup(void);
void tiramisu_MPI_global_barrier(void);
int tiramisu_MPI_Comm_rank(int offset);
}

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef long MPI_Request;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
#define MPI_COMM_WORLD 0
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_SUCCESS 0

static inline int MPI_Init_thread(int *, char ***, int required, int *provided)
{
    if (provided) *provided = required;
    return tiramisu_MPI_init() < 0 ? 1 : MPI_SUCCESS;
}
static inline int MPI_Init(int *, char ***) { return tiramisu_MPI_init() < 0 ? 1 : MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = tiramisu_MPI_Comm_rank(0); return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int *size)
{
    const char *e = getenv("WORLD_SIZE");
    *size = e ? atoi(e) : 1;
    return MPI_SUCCESS;
}
static inline int MPI_Barrier(MPI_Comm) { tiramisu_MPI_global_barrier(); return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { tiramisu_MPI_cleanup(); return MPI_SUCCESS; }
