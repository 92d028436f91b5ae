// tests/emulation/mpi_emu.cpp — TEST INFRASTRUCTURE: a file-mailbox stand-in for the tiramisu_MPI_* runtime so that
// the distributed lowering can be checked on the CPU with as many ranks as the reference's tests assume (10 for
// tests/test_98..102 and test_160 — more than the 8 GPUs of a box, and one process per GPU is what the NCCL-backed
// runtime in csrc/comm.cu supports).  One OS process per rank; RANK / WORLD_SIZE as under torchrun; messages are files
// in $TB200_EMU_MPI_DIR named <src>_<dst>_<sequence>: a send never blocks (write + rename), a receive polls for the
// next sequence number from that source.  Tags are not matched (the generated code issues matching operations in the
// same order on both sides, as MPI requires for a single tag anyway).  Nothing here is part of the product.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <sys/stat.h>
#include <unistd.h>

namespace {
int g_rank = -1, g_world = 1;
std::string g_dir;
std::map<int, long> g_sent, g_recvd;
long g_barriers = 0;

void init_once()
{
    if (g_rank >= 0) return;
    const char *r = getenv("RANK"), *w = getenv("WORLD_SIZE"), *d = getenv("TB200_EMU_MPI_DIR");
    g_rank = r ? atoi(r) : 0;
    g_world = w ? atoi(w) : 1;
    g_dir = d ? d : "/tmp";
}
std::string path(int src, int dst, long seq, const char *ext)
{
    return g_dir + "/" + std::to_string(src) + "_" + std::to_string(dst) + "_" + std::to_string(seq) + ext;
}
void put(const std::string &final_path, const void *data, size_t bytes)
{
    const std::string tmp = final_path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f || (bytes && fwrite(data, 1, bytes, f) != bytes)) { fprintf(stderr, "mpi_emu: cannot write %s\n", tmp.c_str()); _exit(3); }
    fclose(f);
    rename(tmp.c_str(), final_path.c_str());
}
void wait_for(const std::string &p)
{
    struct stat st;
    for (long spins = 0; stat(p.c_str(), &st) != 0; spins++) {
        if (spins > 300000) { fprintf(stderr, "mpi_emu: rank %d timed out waiting for %s\n", g_rank, p.c_str()); _exit(4); }
        usleep(200);
    }
}
void send_bytes(int dest, const void *data, size_t bytes)
{
    init_once();
    if (dest < 0 || dest >= g_world) { fprintf(stderr, "mpi_emu: rank %d sends to invalid rank %d\n", g_rank, dest); _exit(5); }
    put(path(g_rank, dest, g_sent[dest]++, ".msg"), data, bytes);
}
void recv_bytes(int src, void *data, size_t bytes)
{
    init_once();
    if (src < 0 || src >= g_world) { fprintf(stderr, "mpi_emu: rank %d receives from invalid rank %d\n", g_rank, src); _exit(5); }
    const std::string p = path(src, g_rank, g_recvd[src]++, ".msg");
    wait_for(p);
    FILE *f = fopen(p.c_str(), "rb");
    fseek(f, 0, SEEK_END);
    const long have = ftell(f);
    fseek(f, 0, SEEK_SET);
    if ((size_t) have != bytes) { fprintf(stderr, "mpi_emu: rank %d expected %zu bytes from %d, the message has %ld\n", g_rank, bytes, src, have); _exit(6); }
    if (bytes && fread(data, 1, bytes, f) != bytes) _exit(7);
    fclose(f);
    unlink(p.c_str());
}
}  // namespace

// (the reference declares these three with C++ linkage, include/tiramisu/mpi_comm.h:7-9; the stand-in <mpi.h> that its
// header includes first declares them extern "C", which is the linkage every later redeclaration keeps)
extern "C" int tiramisu_MPI_init() { init_once(); return g_rank; }
extern "C" void tiramisu_MPI_cleanup() {}
extern "C" void tiramisu_MPI_global_barrier()
{
    init_once();
    const long b = g_barriers++;
    put(g_dir + "/barrier_" + std::to_string(b) + "_" + std::to_string(g_rank), "", 0);
    for (int r = 0; r < g_world; r++) wait_for(g_dir + "/barrier_" + std::to_string(b) + "_" + std::to_string(r));
}

extern "C" {
int tiramisu_MPI_Comm_rank(int offset) { init_once(); return g_rank + offset; }
void tiramisu_MPI_Wait(void *) {}
#define TB_EMU_MPI(S, T)                                                                                                     \
    void tiramisu_MPI_Send_##S(int count, int dest, int, T *data) { send_bytes(dest, data, (size_t) count * sizeof(T)); }      \
    void tiramisu_MPI_Ssend_##S(int count, int dest, int, T *data) { send_bytes(dest, data, (size_t) count * sizeof(T)); }     \
    void tiramisu_MPI_Isend_##S(int count, int dest, int, T *data, long *) { send_bytes(dest, data, (size_t) count * sizeof(T)); } \
    void tiramisu_MPI_Issend_##S(int count, int dest, int, T *data, long *) { send_bytes(dest, data, (size_t) count * sizeof(T)); } \
    void tiramisu_MPI_Recv_##S(int count, int src, int, T *data) { recv_bytes(src, data, (size_t) count * sizeof(T)); }        \
    void tiramisu_MPI_Irecv_##S(int count, int src, int, T *data, long *) { recv_bytes(src, data, (size_t) count * sizeof(T)); }
TB_EMU_MPI(int8, char) TB_EMU_MPI(uint8, unsigned char) TB_EMU_MPI(int16, short) TB_EMU_MPI(uint16, unsigned short)
TB_EMU_MPI(int32, int) TB_EMU_MPI(uint32, unsigned int) TB_EMU_MPI(int64, long) TB_EMU_MPI(uint64, unsigned long)
TB_EMU_MPI(f32, float) TB_EMU_MPI(f64, double)
}
