// comm.cu — distributed runtime: replaces /root/reference/src/tiramisu_mpi.cpp (MPI point-to-point
// on MPI_COMM_WORLD) with NCCL over NVLink, one process per GPU.  Same entry points and argument
// meaning as include/tiramisu/mpi_comm.h so that Tiramisu-generated distributed code
// (tag_distribute_level + send/recv computations, src/tiramisu_core.cpp:8400-8469) links unchanged:
//   tiramisu_MPI_init / cleanup / global_barrier / Comm_rank / Wait
//   tiramisu_MPI_{Send,Ssend,Isend,Issend,Recv,Irecv}_{int8..uint64,f32,f64}
// Messages are enqueued on the calling thread's stream (stream order replaces MPI's request
// objects: the non-blocking variants are the same enqueue, tiramisu_MPI_Wait is a no-op for device
// buffers).  Host buffers — what the reference's CPU code passes — are staged through a pooled
// device bounce buffer.  NCCL is dlopen'ed lazily so the library has no link-time dependency on it.
//
// Bootstrap: tiramisu_b200_comm_init(rank, world, id) with an id from tiramisu_b200_comm_unique_id
// shipped by any side channel (tiramisu_b200/dist.py uses torch.distributed), or tiramisu_MPI_init()
// which reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT (torchrun's variables) and
// exchanges the id over a TCP socket on MASTER_PORT+1 (TIRAMISU_B200_PORT overrides).
#include "common.cuh"
#include "tiramisu_b200.h"

#include <cuda.h>

#include <arpa/inet.h>
#include <dlfcn.h>
#include <netdb.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

using namespace tb200;

namespace {

// ---- minimal NCCL surface (ABI-stable since NCCL 2.7) -------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6,
       ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct Nccl {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    bool ok = false;
};

Nccl &nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);      // reuses torch's copy when already loaded
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
#define TB200_SYM(field, name) *(void **)(&n.field) = dlsym(h, name)
        TB200_SYM(GetUniqueId, "ncclGetUniqueId");
        TB200_SYM(CommInitRank, "ncclCommInitRank");
        TB200_SYM(CommDestroy, "ncclCommDestroy");
        TB200_SYM(Send, "ncclSend");
        TB200_SYM(Recv, "ncclRecv");
        TB200_SYM(GroupStart, "ncclGroupStart");
        TB200_SYM(GroupEnd, "ncclGroupEnd");
        TB200_SYM(Broadcast, "ncclBroadcast");
        TB200_SYM(AllReduce, "ncclAllReduce");
        TB200_SYM(AllGather, "ncclAllGather");
        TB200_SYM(GetErrorString, "ncclGetErrorString");
#undef TB200_SYM
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.Send && n.Recv && n.GroupStart && n.GroupEnd &&
               n.Broadcast && n.AllReduce && n.AllGather && n.GetErrorString;
    });
    return n;
}

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    float *scratch = nullptr;        // 1 float for the barrier all-reduce
};
CommState g_comm;
std::mutex g_comm_mu;

// The reference exits with code 28 on any MPI failure (src/tiramisu_mpi.cpp:70-76); we go through
// tb200::fail so that the error mode (exit vs return) is honoured.
int nccl_check(ncclResult_t r, const char *where)
{
    if (r == 0) return 0;
    return fail(where, std::string("NCCL: ") + nccl().GetErrorString(r));
}

bool is_device_pointer(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Point-to-point messages.  NCCL needs sends and receives that must progress together to sit in ONE group, and
// its operations are stream-ordered kernels, not host calls, so MPI's semantics are rebuilt like this:
//   * the non-blocking calls (Isend / Issend / Irecv) only QUEUE the message; the queue is flushed as one
//     ncclGroupStart/End by tiramisu_MPI_Wait, by any blocking call, by the barrier and by cleanup — an
//     Irecv-first or bidirectional neighbour exchange therefore progresses as one group instead of deadlocking
//     on the order of two stream-ordered kernels;
//   * a blocking Send / Recv joins whatever is queued and flushes at once;
//   * host buffers are staged through pooled device memory: a send's payload is copied (and the copy waited for)
//     before the call returns, so the caller may reuse its buffer; a receive into host memory is copied back and
//     waited for when its group is flushed.
// Not supported (documented in INTEGRATION.md): tags are ignored — messages between one pair of ranks match in
// issue order; a closed ring of *blocking* Sends issued before any Recv (legal with MPI's eager protocol) blocks.
struct Pending {
    bool send;
    void *dev;          // device address handed to NCCL
    size_t bytes;
    int peer;
    void *host;         // non-null: staged host buffer (copy back after a receive)
    bool staged;        // dev comes from the pool
};
thread_local std::vector<Pending> t_pending;

int flush_pending()
{
    if (t_pending.empty()) return 0;
    Nccl &n = nccl();
    int rc = 0;
    if (g_comm.comm) {
        rc |= nccl_check(n.GroupStart(), "ncclGroupStart");
        for (const Pending &m : t_pending)
            rc |= m.send ? nccl_check(n.Send(m.dev, m.bytes, ncclInt8, m.peer, g_comm.comm, stream()), "ncclSend")
                         : nccl_check(n.Recv(m.dev, m.bytes, ncclInt8, m.peer, g_comm.comm, stream()), "ncclRecv");
        rc |= nccl_check(n.GroupEnd(), "ncclGroupEnd");
    } else {
        rc = fail("tiramisu_MPI", "communicator not initialised");
    }
    bool host_recv = false;
    for (const Pending &m : t_pending)
        if (!m.send && m.host && !rc) { rc |= tiramisu_cuda_memcpy_to_host_async(m.host, m.dev, m.bytes); host_recv = true; }
    if (host_recv && !rc) rc |= check(cudaStreamSynchronize(stream()), "tiramisu_MPI_Recv");   // data is in the caller's memory on return
    for (const Pending &m : t_pending)
        if (m.staged) pool_free(m.dev);
    t_pending.clear();
    return rc;
}

int p2p(bool send, void *data, size_t bytes, int peer, bool blocking)
{
    if (!g_comm.comm) return fail(send ? "tiramisu_MPI_Send" : "tiramisu_MPI_Recv", "communicator not initialised");
    if (peer < 0 || peer >= g_comm.world) return fail("tiramisu_MPI p2p", "peer rank out of range");
    if (peer == g_comm.rank) return fail("tiramisu_MPI p2p", "message to self");
    Pending m{send, data, bytes, peer, nullptr, false};
    if (!is_device_pointer(data)) {
        // host buffer (the reference's MPI usage): bounce through pooled device memory
        m.dev = pool_alloc(bytes);
        if (!m.dev) return 1;
        m.host = data;
        m.staged = true;
        if (send) {
            if (tiramisu_cuda_memcpy_to_device_async(m.dev, data, bytes)) return 1;
            // the caller may overwrite `data` as soon as we return (MPI_Send / MPI_Isend + buffered copy)
            if (check(cudaStreamSynchronize(stream()), "tiramisu_MPI_Send")) return 1;
        }
    }
    t_pending.push_back(m);
    return blocking ? flush_pending() : 0;
}

// ---- TCP rendezvous for the unique id (single node, torchrun environment) -------------------------------
int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

int exchange_id_tcp(int rank, int world, ncclUniqueId *id)
{
    const char *addr = getenv("MASTER_ADDR");
    if (!addr) addr = "127.0.0.1";
    const int port = env_int("TIRAMISU_B200_PORT", env_int("MASTER_PORT", 29500) + 1);
    if (rank == 0) {
        int srv = socket(AF_INET, SOCK_STREAM, 0);
        if (srv < 0) return fail("tiramisu_MPI_init", "socket() failed");
        int one = 1;
        setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
        sockaddr_in sa{};
        sa.sin_family = AF_INET;
        sa.sin_addr.s_addr = htonl(INADDR_ANY);
        sa.sin_port = htons((uint16_t)port);
        if (bind(srv, (sockaddr *)&sa, sizeof(sa)) != 0 || listen(srv, world) != 0) {
            close(srv);
            return fail("tiramisu_MPI_init", "cannot listen on the rendezvous port");
        }
        for (int i = 1; i < world; i++) {
            int c = accept(srv, nullptr, nullptr);
            if (c < 0) { close(srv); return fail("tiramisu_MPI_init", "accept() failed"); }
            size_t off = 0;
            while (off < sizeof(*id)) {
                ssize_t w = write(c, (const char *)id + off, sizeof(*id) - off);
                if (w <= 0) { close(c); close(srv); return fail("tiramisu_MPI_init", "short write of the NCCL id"); }
                off += (size_t)w;
            }
            close(c);
        }
        close(srv);
        return 0;
    }
    sockaddr_in sa{};
    sa.sin_family = AF_INET;
    sa.sin_port = htons((uint16_t)port);
    if (inet_pton(AF_INET, addr, &sa.sin_addr) != 1) {
        hostent *he = gethostbyname(addr);
        if (!he) return fail("tiramisu_MPI_init", "cannot resolve MASTER_ADDR");
        memcpy(&sa.sin_addr, he->h_addr_list[0], sizeof(sa.sin_addr));
    }
    for (int attempt = 0; attempt < 600; attempt++) {           // up to ~60 s for rank 0 to come up
        int c = socket(AF_INET, SOCK_STREAM, 0);
        if (c < 0) return fail("tiramisu_MPI_init", "socket() failed");
        if (connect(c, (sockaddr *)&sa, sizeof(sa)) == 0) {
            size_t off = 0;
            while (off < sizeof(*id)) {
                ssize_t r = read(c, (char *)id + off, sizeof(*id) - off);
                if (r <= 0) break;
                off += (size_t)r;
            }
            close(c);
            if (off == sizeof(*id)) return 0;
            return fail("tiramisu_MPI_init", "short read of the NCCL id");
        }
        close(c);
        usleep(100 * 1000);
    }
    return fail("tiramisu_MPI_init", "timed out connecting to rank 0");
}

}  // namespace

namespace tb200 {

// Halo rows of the distributed box blur (tests/test_160.cpp:53 gen_communication(); manual form
// benchmarks/automatic_comm/blurautodist_ref.cpp:42-49): rank p sends its first 2 rows to rank p-1,
// which stores them in rows [rows, rows+2) of its input.  One grouped exchange on the stream.
int boxblur_halo_exchange(uint32_t *d_in, int rows, int cols)
{
    if (!g_comm.comm || g_comm.world == 1) return 0;       // single process: the caller supplied the halo rows
    Nccl &n = nccl();
    const size_t row_bytes = (size_t)(cols + 2) * 4;
    if (nccl_check(n.GroupStart(), "ncclGroupStart")) return 1;
    int rc = 0;
    if (g_comm.rank > 0)
        rc |= nccl_check(n.Send(d_in, 2 * row_bytes, ncclInt8, g_comm.rank - 1, g_comm.comm, stream()), "ncclSend");
    if (g_comm.rank < g_comm.world - 1)
        rc |= nccl_check(n.Recv(d_in + (size_t)rows * (cols + 2), 2 * row_bytes, ncclInt8, g_comm.rank + 1,
                                g_comm.comm, stream()),
                         "ncclRecv");
    rc |= nccl_check(n.GroupEnd(), "ncclGroupEnd");
    return rc;
}

}  // namespace tb200

// =============================================================================== communicator control
TB200_API int tiramisu_b200_comm_unique_id(void *out128)
{
    if (!nccl().ok) return fail(__func__, "libnccl.so.2 not found");
    return nccl_check(nccl().GetUniqueId(reinterpret_cast<ncclUniqueId *>(out128)), "ncclGetUniqueId");
}

TB200_API int tiramisu_b200_comm_init(int32_t rank, int32_t world, const void *id128)
{
    if (!nccl().ok) return fail(__func__, "libnccl.so.2 not found");
    if (world < 1 || rank < 0 || rank >= world || !id128) return fail(__func__, "bad rank/world/id");
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (g_comm.comm) return fail(__func__, "communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    if (nccl_check(nccl().CommInitRank(&c, world, id, rank), "ncclCommInitRank")) return 1;
    g_comm.comm = c;
    g_comm.rank = rank;
    g_comm.world = world;
    TB200_CHECK(cudaMalloc(&g_comm.scratch, sizeof(float)));
    TB200_CHECK(cudaMemset(g_comm.scratch, 0, sizeof(float)));
    return 0;
}

namespace tb200 { int heat2d_dist_release(); }        // heat2d_dist.cu: cached symmetric block of heat2d_dist_tiramisu

TB200_API int tiramisu_b200_comm_destroy(void)
{
    if (g_comm.comm) {
        flush_pending();
        heat2d_dist_release();                             // collective: every rank destroys its communicator
    }
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (!g_comm.comm) return 0;
    cudaStreamSynchronize(stream());
    int rc = nccl_check(nccl().CommDestroy(g_comm.comm), "ncclCommDestroy");
    cudaFree(g_comm.scratch);
    g_comm = CommState();
    return rc;
}

TB200_API int tiramisu_b200_comm_rank(void) { return g_comm.rank; }
TB200_API int tiramisu_b200_comm_size(void) { return g_comm.world; }

// Replicate `bytes` of device memory from `root` to every rank (B panels of the row-sharded sgemm,
// conv / LSTM weights of the batch-sharded blocks).
TB200_API int tiramisu_b200_comm_broadcast(void *dev_buf, uint64_t bytes, int32_t root)
{
    if (!g_comm.comm) return fail(__func__, "communicator not initialised");
    return nccl_check(nccl().Broadcast(dev_buf, dev_buf, bytes, ncclInt8, root, g_comm.comm, stream()), "ncclBroadcast");
}

// Bandwidth-optimal broadcast of a large buffer: the root scatters 1/world of it to every rank
// (one grouped send/recv) and the ranks all-gather the pieces in place.  Every GPU then receives at
// its full NVLink ingress rate from world-1 peers instead of everybody pulling from the root
// (ncclBroadcast, measured 350 GB/s per rank on 8 B200s for 1 GiB).  Falls back to ncclBroadcast
// for small or odd sizes.  Measured on 8 B200s (r01): NOT faster than ncclBroadcast when it overlaps
// GEMMs (3.85 vs 3.07 ms for 1 GiB) — kept as an alternative, not the default of the sgemm sweep.
TB200_API int tiramisu_b200_comm_broadcast_large(void *dev_buf, uint64_t bytes, int32_t root)
{
    if (!g_comm.comm) return fail(__func__, "communicator not initialised");
    Nccl &nc = nccl();
    const uint64_t world = (uint64_t)g_comm.world;
    if (world == 1) return 0;
    const uint64_t chunk = (bytes / world) & ~uint64_t(255);
    if (chunk < (1u << 20))
        return nccl_check(nc.Broadcast(dev_buf, dev_buf, bytes, ncclInt8, root, g_comm.comm, stream()), "ncclBroadcast");
    char *base = reinterpret_cast<char *>(dev_buf);
    int rc = nccl_check(nc.GroupStart(), "ncclGroupStart");
    if (g_comm.rank == root) {
        for (int r = 0; r < g_comm.world; r++)
            if (r != root) rc |= nccl_check(nc.Send(base + (uint64_t)r * chunk, chunk, ncclInt8, r, g_comm.comm, stream()), "ncclSend");
    } else {
        rc |= nccl_check(nc.Recv(base + (uint64_t)g_comm.rank * chunk, chunk, ncclInt8, root, g_comm.comm, stream()), "ncclRecv");
    }
    rc |= nccl_check(nc.GroupEnd(), "ncclGroupEnd");
    if (rc) return rc;
    if (nccl_check(nc.AllGather(base + (uint64_t)g_comm.rank * chunk, base, chunk, ncclInt8, g_comm.comm, stream()), "ncclAllGather"))
        return 1;
    const uint64_t done = chunk * world;
    if (done < bytes)
        return nccl_check(nc.Broadcast(base + done, base + done, bytes - done, ncclInt8, root, g_comm.comm, stream()), "ncclBroadcast");
    return 0;
}

// Grouped send/recv: up to `n` messages enqueued as ONE NCCL group on the stream (no ordering
// deadlocks between neighbours, one launch).  peer < 0 entries are skipped.
TB200_API int tiramisu_b200_comm_exchange(int32_t n, const int32_t *send_peer, void *const *send_buf,
                                          const uint64_t *send_bytes, const int32_t *recv_peer, void *const *recv_buf,
                                          const uint64_t *recv_bytes)
{
    if (!g_comm.comm) return fail(__func__, "communicator not initialised");
    Nccl &nc = nccl();
    if (nccl_check(nc.GroupStart(), "ncclGroupStart")) return 1;
    int rc = 0;
    for (int i = 0; i < n; i++) {
        if (send_peer[i] >= 0)
            rc |= nccl_check(nc.Send(send_buf[i], send_bytes[i], ncclInt8, send_peer[i], g_comm.comm, stream()), "ncclSend");
        if (recv_peer[i] >= 0)
            rc |= nccl_check(nc.Recv(recv_buf[i], recv_bytes[i], ncclInt8, recv_peer[i], g_comm.comm, stream()), "ncclRecv");
    }
    rc |= nccl_check(nc.GroupEnd(), "ncclGroupEnd");
    return rc;
}

// ---- symmetric memory over CUDA IPC --------------------------------------------------------------
// The stencil kernels read halo rows straight out of the neighbours' HBM (stencil.cu: Halo); this is
// the mapping that makes a neighbour's rows addressable.  cudaMalloc (not the stream-ordered pool:
// pool memory cannot be exported through cudaIpcGetMemHandle), handles all-gathered through the
// communicator itself so that any host program — torchrun or the reference's MPI-style wrappers
// after tiramisu_MPI_init() — can use it.
namespace {
struct SymBlock { void *local; std::vector<void *> opened; };
std::vector<SymBlock> g_sym;
}

// Collective and failure-safe: a rank whose cudaMalloc / cudaIpcGetMemHandle / cudaIpcOpenMemHandle fails still
// takes part in both exchanges below (status travels with the handle, then a second status round after the
// mappings), so either every rank returns success or every rank returns failure with nothing left allocated —
// nobody is left waiting in a collective for a rank that gave up.
TB200_API int tiramisu_b200_comm_alloc_symmetric(uint64_t bytes, void **local, void **peers)
{
    if (!local || !peers || bytes == 0) return fail(__func__, "bad argument");
    if (!g_comm.comm) return fail(__func__, "communicator not initialised");
    const int world = g_comm.world, rank = g_comm.rank;
    struct Record { cudaIpcMemHandle_t handle; int32_t status; int32_t pad[3]; };
    static_assert(sizeof(cudaIpcMemHandle_t) == 64 && sizeof(Record) == 80, "IPC record layout");
    Record rec;
    memset(&rec, 0, sizeof(rec));
    void *mine = nullptr;
    std::string why;
    auto note = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rec.status == 0) { rec.status = 1; why = std::string(what) + ": " + cudaGetErrorString(e); cudaGetLastError(); }
        return e == cudaSuccess;
    };
    if (note(cudaMalloc(&mine, bytes), "cudaMalloc") && note(cudaMemset(mine, 0, bytes), "cudaMemset"))
        note(cudaIpcGetMemHandle(&rec.handle, mine), "cudaIpcGetMemHandle");
    // round 1: handles + local status
    char *d_all = nullptr;
    TB200_CHECK(cudaMalloc(&d_all, (size_t)world * sizeof(Record)));       // a failure here is not recoverable collectively
    TB200_CHECK(cudaMemcpyAsync(d_all + (size_t)rank * sizeof(Record), &rec, sizeof(Record), cudaMemcpyHostToDevice, stream()));
    if (nccl_check(nccl().AllGather(d_all + (size_t)rank * sizeof(Record), d_all, sizeof(Record), ncclInt8, g_comm.comm, stream()), "ncclAllGather")) return 1;
    std::vector<Record> all(world);
    TB200_CHECK(cudaMemcpyAsync(all.data(), d_all, (size_t)world * sizeof(Record), cudaMemcpyDeviceToHost, stream()));
    TB200_CHECK(cudaStreamSynchronize(stream()));
    int bad_rank = -1;
    for (int r = 0; r < world; r++) if (all[r].status && bad_rank < 0) bad_rank = r;
    SymBlock blk{mine, {}};
    int open_failed = 0;
    if (bad_rank < 0) {
        for (int r = 0; r < world; r++) {
            if (r == rank) { peers[r] = mine; continue; }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                why = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(cudaGetLastError());
                open_failed = 1;
                break;
            }
            peers[r] = p;
            blk.opened.push_back(p);
        }
    }
    // round 2: did every rank map every peer?  (sum of failure flags; also the rendezvous before anybody frees)
    float flag = (bad_rank >= 0 || open_failed) ? 1.0f : 0.0f, total = 0.0f;
    float *d_flag = reinterpret_cast<float *>(d_all);
    TB200_CHECK(cudaMemcpyAsync(d_flag, &flag, sizeof(float), cudaMemcpyHostToDevice, stream()));
    if (nccl_check(nccl().AllReduce(d_flag, d_flag, 1, ncclFloat32, ncclSum, g_comm.comm, stream()), "ncclAllReduce")) return 1;
    TB200_CHECK(cudaMemcpyAsync(&total, d_flag, sizeof(float), cudaMemcpyDeviceToHost, stream()));
    TB200_CHECK(cudaStreamSynchronize(stream()));
    cudaFree(d_all);
    if (total != 0.0f) {
        for (void *p : blk.opened) cudaIpcCloseMemHandle(p);
        if (mine) cudaFree(mine);
        for (int r = 0; r < world; r++) peers[r] = nullptr;
        *local = nullptr;
        if (why.empty())
            why = bad_rank >= 0 ? "rank " + std::to_string(bad_rank) + " could not allocate or export its block"
                                : "another rank could not map a peer's block";
        return fail(__func__, why);
    }
    *local = mine;
    std::lock_guard<std::mutex> lk(g_comm_mu);
    g_sym.push_back(blk);
    return 0;
}

// ---- copy-engine broadcast over symmetric memory, fused with the row-sharded sgemm ------------------------
// Row-sharded sgemm needs rank 0's B on every rank while the GEMM runs.  ncclBroadcast does that with
// SM kernels that compete with a persistent GEMM for the same SMs, along a ring (measured r01: ~350 GB/s per rank, 3 ms
// for the 1 GiB B of 16384^3 against 1.6 ms of GEMM on 8 GPUs).  Here the data travels through symmetric memory in
// `panels` contiguous panels, every panel cut into world-1 slices, one per non-root rank:
//   scatter    : rank r pulls ITS slice of panel p from the root          — every byte leaves the root exactly once
//   all-gather : rank r pulls slice q of panel p from rank q (q != root)  — every other link carries 1/(world-1) of it
// as peer cudaMemcpyAsync on two library-owned streams per thread: DMA engines only, no SM work at all, so the transfer
// overlaps the tcgen05 GEMM instead of fighting it for SMs.  Contiguous ranges are what the DMA engines move fastest
// between processes (r02 run 9: 626 GB/s; a pitched copy of 4 KB rows of IPC-mapped memory: 257 GB/s), so for the sgemm
// the root first re-lays B out PANEL-MAJOR into its staging block — column panel j, B[:, j w .. (j+1) w), becomes a
// contiguous row-major [K][w] matrix — with local pitched DMA copies; the GEMM reads that layout through a 3-D tensor
// map (gemm_tc.cu: GemmParams::b_panel_cols).  Cross-rank ordering uses uint32 words inside the symmetric block at
// `flags_offset` (zero at allocation):
//   [0, world)         ready[q]  last token for which rank q holds its own slice; ready[root]: last token for which the
//                                root's staging panel is in place                  (written by q: 4-byte peer copy)
//   [world, 2 world)   done[q]   last epoch for which rank q finished reading from me   (written by q)
//   [2 world]          have      last token for which the WHOLE panel is in my block    (cuStreamWriteValue32, local)
//   [2 world + 1, +2]  source words of the two kinds of peer flag copies
// waits are cuStreamWaitValue32 on LOCAL words (stream-side, no spinning kernel), token(epoch, p) = 64 (epoch-1) + p + 1.
// The consumer of `have` is the GEMM itself: ONE persistent tcgen05 launch walks its tiles column panel by column
// panel and its TMA producer warp polls `have` before the B loads of a tile whose panel it has not yet seen
// (GemmParams::gate) — no K-panel accumulation passes over C, no host or event synchronisation.  That launch is
// submitted BEFORE the copies: a hardware work queue is FIFO in submission order across the streams that share it, and
// an event / value wait at the head of a queue holds back every later entry — submitted first, the copies' waits sat in
// front of the GEMM launch and the whole broadcast ran before the first MMA (r02 run 9: fused = broadcast + GEMM).
namespace tb200 {
void gemm_tc_gate(const uint32_t *flag, uint32_t value, int cols, uint32_t last);
void gemm_tc_b_panels(int cols, uint64_t panel_stride);
int gemm_tc_kernel_choice(uint64_t M, uint64_t N, uint64_t K, bool *was_auto);
bool gemm_tc_addressable(const float *A, const float *B, const float *C, uint64_t M, uint64_t N, uint64_t K, uint64_t ldA,
                         uint64_t ldB, uint64_t ldC, bool tA, bool tB);
int gemm_f32_tc(const float *A, const float *B, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha, float beta,
                uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tB);
int gemm_f32(const float *A, const float *B, float *C, uint64_t M, uint64_t N, uint64_t K, float alpha, float beta,
             uint64_t ldA, uint64_t ldB, uint64_t ldC, bool tA, bool tB);

using StreamMemOpFn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamMemOpFn driver_fn(const char *name)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<StreamMemOpFn>(p);
}

int stream_wait_value32(cudaStream_t st, const uint32_t *flag, uint32_t value)
{
    static StreamMemOpFn fn = driver_fn("cuStreamWaitValue32");
    if (!fn) return fail("stream_wait_value32", "cuStreamWaitValue32 unavailable");
    CUresult r = fn(st, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WAIT_VALUE_GEQ);
    return r == CUDA_SUCCESS ? 0 : fail("cuStreamWaitValue32", "error " + std::to_string((int)r));
}

int stream_write_value32(cudaStream_t st, uint32_t *flag, uint32_t value)
{
    static StreamMemOpFn fn = driver_fn("cuStreamWriteValue32");
    if (!fn) return fail("stream_write_value32", "cuStreamWriteValue32 unavailable");
    CUresult r = fn(st, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WRITE_VALUE_DEFAULT);
    return r == CUDA_SUCCESS ? 0 : fail("cuStreamWriteValue32", "error " + std::to_string((int)r));
}
}  // namespace tb200

namespace {
// ---- the two tiny kernels of the broadcast: flag fan-out and the root's panel-major re-layout ----------------------
// A 4-byte peer copy costs as much DMA set-up time as a multi-megabyte one (r02 run 13: ~11 us per copy between IPC-mapped
// blocks), so one rank's announcement to all its peers is ONE one-warp kernel storing the token into every peer's flag
// word (st.release.sys after the stream-ordered data copies), not world-1 copies.
struct FlagTargets { uint32_t *p[32]; };
__global__ void tiramisu_sm100_sym_flag_fanout(FlagTargets t, int n, uint32_t value)
{
    const int q = threadIdx.x;
    if (q < n && t.p[q]) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.p[q]), "r"(value) : "memory");
}
// dst[k][c] = src[k][col0 + c], c < cols: column panel `col0 / cols` of the row-major B becomes a contiguous [K][cols]
// matrix in the staging block.  HBM-bound (2 x 4 bytes per element), float4 lanes, few registers so that it co-resides
// with the root's own persistent GEMM.
__global__ void __launch_bounds__(256) tiramisu_sm100_pack_panel(const float *__restrict__ src, float *__restrict__ dst, uint64_t K,
                                                                  uint64_t ldB, uint64_t col0, uint32_t cols)
{
    const uint32_t vec_per_row = cols / 4;
    const uint64_t total = K * vec_per_row;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = i / vec_per_row;
        const uint32_t c = (uint32_t)(i - k * vec_per_row) * 4;
        const float4 v = *reinterpret_cast<const float4 *>(src + k * ldB + col0 + c);
        *reinterpret_cast<float4 *>(dst + k * cols + c) = v;
    }
}

constexpr int kMaxPanels = 64;
struct BcastStreams {
    cudaStream_t scatter = nullptr, gather = nullptr, pack = nullptr;
    cudaEvent_t fork = nullptr, end_s = nullptr, end_g = nullptr, slice[kMaxPanels] = {}, packed[kMaxPanels] = {};
    int init()
    {
        if (scatter) return 0;
        int lo = 0, hi = 0;                                  // copy streams at the highest priority: their (tiny) stream
        cudaDeviceGetStreamPriorityRange(&lo, &hi);          // operations are never queued behind compute of equal priority
        TB200_CHECK(cudaStreamCreateWithPriority(&scatter, cudaStreamNonBlocking, hi));
        TB200_CHECK(cudaStreamCreateWithPriority(&gather, cudaStreamNonBlocking, hi));
        TB200_CHECK(cudaStreamCreateWithPriority(&pack, cudaStreamNonBlocking, hi));
        for (cudaEvent_t *e : {&fork, &end_s, &end_g}) TB200_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (int i = 0; i < kMaxPanels; i++) {
            TB200_CHECK(cudaEventCreateWithFlags(&slice[i], cudaEventDisableTiming));
            TB200_CHECK(cudaEventCreateWithFlags(&packed[i], cudaEventDisableTiming));
        }
        // Load the two helper kernels NOW: with lazy module loading the first launch of a kernel loads its module, and
        // that load waits for running kernels — a gated GEMM that spins until these very kernels have announced a panel
        // would never see it (r02 run 15: the GEMM's bounded waits trapped after 2 s on the first fused call).
        cudaFuncAttributes fa;
        TB200_CHECK(cudaFuncGetAttributes(&fa, tiramisu_sm100_sym_flag_fanout));
        TB200_CHECK(cudaFuncGetAttributes(&fa, tiramisu_sm100_pack_panel));
        FlagTargets none{};
        tiramisu_sm100_sym_flag_fanout<<<1, 32, 0, scatter>>>(none, 0, 0u);
        tiramisu_sm100_pack_panel<<<1, 256, 0, scatter>>>(nullptr, nullptr, 0, 0, 0, 4);
        TB200_CHECK(cudaStreamSynchronize(scatter));
        return 0;
    }
};
thread_local BcastStreams t_bs;

// TB200_BCAST_TRACE=1: timing events after every operation of one broadcast, printed (relative to the fork) by
// tiramisu_b200_comm_broadcast_trace_dump() — diagnosis of where a panel spends its time (r02 run 14)
struct BcastTrace {
    bool on = false;
    cudaEvent_t t0 = nullptr;
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    void begin(cudaStream_t st)
    {
        on = getenv("TB200_BCAST_TRACE") != nullptr;
        if (!on) return;
        for (auto &m : marks) cudaEventDestroy(m.second);
        marks.clear();
        if (!t0) cudaEventCreate(&t0);
        cudaEventRecord(t0, st);
    }
    void mark(const std::string &what, cudaStream_t st)
    {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.push_back({what, e});
    }
};
thread_local BcastTrace t_trace;

// Row-major source the root re-lays out panel-major into its staging block before a panel travels (sgemm), or nothing
// (plain broadcast: the root's block already holds the bytes).
struct RootPack {
    const float *B = nullptr;                     // [K][N] row-major on the root, pitch ldB elements
    uint64_t K = 0, ldB = 0, cols = 0;            // columns per panel
};

// Fork of the copy streams from the caller's stream: recorded BEFORE the consumer (GEMM) is launched, so that the copies
// are not ordered behind it.
int bcast_sym_fork(cudaStream_t st)
{
    if (t_bs.init()) return 1;
    TB200_CHECK(cudaEventRecord(t_bs.fork, st));
    t_trace.begin(st);
    return 0;
}

// Enqueues the whole panelised broadcast of [data_offset, data_offset + bytes) on the library's two copy streams (after
// bcast_sym_fork); leaves end_s / end_g recorded (the caller joins them when it wants "complete / reusable" in stream order).
int bcast_sym_enqueue(void *const *peers, uint64_t data_offset, uint64_t bytes, uint64_t panel_bytes, uint64_t flags_offset,
                      int root, uint32_t epoch, int panels, const RootPack &pack, bool double_buffered = false)
{
    const int world = g_comm.world, rank = g_comm.rank;
    auto data = [&](int r) { return static_cast<char *>(peers[r]) + data_offset; };
    auto flags = [&](int r) { return reinterpret_cast<uint32_t *>(static_cast<char *>(peers[r]) + flags_offset); };
    uint32_t *mine = flags(rank);
    const uint32_t base = 64u * (epoch - 1);
    TB200_CHECK(cudaStreamWaitEvent(t_bs.gather, t_bs.fork, 0));
    const int ns = world - 1;                                               // slices per panel: one per non-root rank
    auto slice_of = [&](int r) { return r < root ? r : r - 1; };
    auto slice_range = [&](int p, int s, uint64_t *lo, uint64_t *len) {    // byte range of slice s of panel p inside the data
        const uint64_t p_lo = std::min<uint64_t>(bytes, (uint64_t)p * panel_bytes);
        const uint64_t p_len = std::min<uint64_t>(bytes, p_lo + panel_bytes) - p_lo;
        const uint64_t sl = ((p_len + ns - 1) / ns + 255) & ~uint64_t(255);
        const uint64_t a = std::min<uint64_t>(p_len, (uint64_t)s * sl), b = std::min<uint64_t>(p_len, (uint64_t)(s + 1) * sl);
        *lo = p_lo + a; *len = b - a;
    };
    auto fanout = [&](cudaStream_t cs, int word, uint32_t value, bool to_root) -> int {   // flags(q)[word] = value at every peer q
        FlagTargets t;
        int n = 0;
        for (int q = 0; q < world; q++)
            if (q != rank && (to_root || q != root)) t.p[n++] = flags(q) + word;
        if (!n) return 0;
        tiramisu_sm100_sym_flag_fanout<<<1, 32, 0, cs>>>(t, n, value);
        return post_launch("tiramisu_sm100_sym_flag_fanout");
    };
    if (rank == root) {
        // The root PUSHES: its DMA engines write every rank's slice of a panel into that rank's block at full link speed
        // (r02 run 14: seven ranks pulling from one root share its egress, each pull crawls at 1/7 of the link and keeps the
        // puller's engine busy for the whole scatter).  Panel after panel: stage it (sgemm: panel-major re-layout), push the
        // slices, announce the panel — ordered after whatever filled the source on the caller's stream.  The source may be
        // rewritten once the pushes are done and every peer has finished reading the staged copy.
        TB200_CHECK(cudaStreamWaitEvent(t_bs.scatter, t_bs.fork, 0));
        if (pack.B) TB200_CHECK(cudaStreamWaitEvent(t_bs.pack, t_bs.fork, 0));
        // Double-buffered staging (sgemm): this epoch's half was last read in epoch - 2; the peers' `done` words of that
        // epoch are all this call has to wait for, so the root runs up to two epochs ahead of the slowest consumer.
        if (double_buffered && epoch > 2)
            for (int q = 0; q < world; q++)
                if (q != root && stream_wait_value32(pack.B ? t_bs.pack : t_bs.gather, mine + world + q, epoch - 2)) return 1;
        for (int p = 0; p < panels; p++) {
            if (pack.B) {                                             // re-layout on its own stream, one panel ahead of the pushes
                const uint64_t total_vec = pack.K * pack.cols / 4;
                const unsigned grid = (unsigned)std::min<uint64_t>((total_vec + 255) / 256, (uint64_t)sm_count() * 4);
                tiramisu_sm100_pack_panel<<<grid, 256, 0, t_bs.pack>>>(pack.B, reinterpret_cast<float *>(data(root) + (uint64_t)p * panel_bytes),
                                                                        pack.K, pack.ldB, (uint64_t)p * pack.cols, (uint32_t)pack.cols);
                if (post_launch("tiramisu_sm100_pack_panel")) return 1;
                TB200_CHECK(cudaEventRecord(t_bs.packed[p], t_bs.pack));
                TB200_CHECK(cudaStreamWaitEvent(t_bs.gather, t_bs.packed[p], 0));
                TB200_CHECK(cudaStreamWaitEvent(t_bs.scatter, t_bs.packed[p], 0));
                t_trace.mark("root staged panel " + std::to_string(p), t_bs.pack);
            }
            // the slices leave on two streams (the fixed set-up time of one copy overlaps the transfer of the other)
            for (int k = 1; k < world; k++) {
                const int q = (root + k) % world;
                uint64_t lo, len;
                slice_range(p, slice_of(q), &lo, &len);
                if (len) TB200_CHECK(cudaMemcpyAsync(data(q) + lo, data(root) + lo, len, cudaMemcpyDeviceToDevice, (k & 1) ? t_bs.gather : t_bs.scatter));
            }
            TB200_CHECK(cudaEventRecord(t_bs.slice[p], t_bs.scatter));
            TB200_CHECK(cudaStreamWaitEvent(t_bs.gather, t_bs.slice[p], 0));
            if (fanout(t_bs.gather, root, base + (uint32_t)p + 1, false)) return 1;
            t_trace.mark("root pushed and announced panel " + std::to_string(p), t_bs.gather);
        }
        if (stream_write_value32(t_bs.gather, mine + 2 * world, base + (uint32_t)panels)) return 1;
        for (int q = 0; q < world && !double_buffered; q++)          // in-place data: nobody may still read it when the call returns
            if (q != root && stream_wait_value32(t_bs.gather, mine + world + q, epoch)) return 1;
        t_trace.mark("root saw every peer done", t_bs.gather);
        TB200_CHECK(cudaEventRecord(t_bs.end_g, t_bs.gather));
        TB200_CHECK(cudaEventRecord(t_bs.end_s, t_bs.gather));
        return 0;
    }
    TB200_CHECK(cudaStreamWaitEvent(t_bs.scatter, t_bs.fork, 0));
    for (int p = 0; p < panels; p++) {
        const uint32_t token = base + (uint32_t)p + 1;
        // scatter: the root has pushed my slice of the panel; tell every other non-root rank that it can be pulled from me
        if (stream_wait_value32(t_bs.scatter, mine + root, token)) return 1;
        t_trace.mark("p" + std::to_string(p) + " own slice arrived from root", t_bs.scatter);
        if (ns > 1 && fanout(t_bs.scatter, rank, token, false)) return 1;
        TB200_CHECK(cudaEventRecord(t_bs.slice[p], t_bs.scatter));
        // all-gather: slice q of this panel from rank q, nearest ranks first so that concurrent pulls use different links
        TB200_CHECK(cudaStreamWaitEvent(t_bs.gather, t_bs.slice[p], 0));
        for (int k = 1; k < world; k++) {
            const int q = (rank + k) % world;
            if (q == root) continue;
            uint64_t lo, len;
            slice_range(p, slice_of(q), &lo, &len);
            if (stream_wait_value32(t_bs.gather, mine + q, token)) return 1;
            if (len) TB200_CHECK(cudaMemcpyAsync(data(rank) + lo, data(q) + lo, len, cudaMemcpyDeviceToDevice, t_bs.gather));
            t_trace.mark("p" + std::to_string(p) + " slice of rank " + std::to_string(q) + " pulled", t_bs.gather);
        }
        if (stream_write_value32(t_bs.gather, mine + 2 * world, token)) return 1;           // have: panel p complete here
    }
    TB200_CHECK(cudaEventRecord(t_bs.end_s, t_bs.scatter));
    TB200_CHECK(cudaEventRecord(t_bs.end_g, t_bs.gather));
    return 0;
}

// Joins the copy streams into the caller's stream and — on a non-root rank — tells the root, from THAT stream, that this
// epoch is over here: the pulls from the peers are done and so is whatever the caller queued before (the GEMM that reads
// the staged bytes).  The root's call returns (in stream order) only after every rank's word, and the root alone writes
// into the other ranks' blocks, so the next epoch cannot overwrite anything that is still being read.
int bcast_sym_join(cudaStream_t st, void *const *peers, uint64_t flags_offset, int root, uint32_t epoch)
{
    TB200_CHECK(cudaStreamWaitEvent(st, t_bs.end_s, 0));
    TB200_CHECK(cudaStreamWaitEvent(st, t_bs.end_g, 0));
    if (g_comm.rank != root) {
        FlagTargets t;
        t.p[0] = reinterpret_cast<uint32_t *>(static_cast<char *>(peers[root]) + flags_offset) + g_comm.world + g_comm.rank;
        tiramisu_sm100_sym_flag_fanout<<<1, 32, 0, st>>>(t, 1, epoch);
        return post_launch("tiramisu_sm100_sym_flag_fanout");
    }
    return 0;
}

int bcast_sym_check(const char *who, void *const *peers, int32_t root, uint32_t epoch, int32_t panels)
{
    if (!g_comm.comm) return fail(who, "communicator not initialised");
    if (!peers || root < 0 || root >= g_comm.world || epoch == 0 || epoch >= (1u << 25) || panels < 1 || panels > kMaxPanels ||
        g_comm.world > 30)
        return fail(who, "bad argument");
    return 0;
}
}  // namespace

// Broadcast of [data_offset, data_offset + bytes) of the root's symmetric block into every rank's block, in `panels`
// equal parts (see above).  Stream-ordered on the calling thread's stream: work queued after the call sees the data.
TB200_API int tiramisu_b200_comm_broadcast_sym(void *const *peers, uint64_t data_offset, uint64_t bytes, uint64_t flags_offset,
                                               int32_t root, uint32_t epoch, int32_t panels)
{
    if (bcast_sym_check(__func__, peers, root, epoch, panels)) return 1;
    if (g_comm.world == 1 || bytes == 0) return 0;
    cudaStream_t st = stream();
    const uint64_t panel_bytes = (((bytes + panels - 1) / panels) + 255) & ~uint64_t(255);
    if (bcast_sym_fork(st) || bcast_sym_enqueue(peers, data_offset, bytes, panel_bytes, flags_offset, root, epoch, panels, RootPack{}))
        return 1;
    return bcast_sym_join(st, peers, flags_offset, root, epoch);
}

// Row-sharded sgemm with B's arrival fused in:  C[rows][N] = alpha * A[rows][K] * B[K][N] + beta * C  on every rank,
// A and C this rank's row shard, B an ordinary row-major device matrix on the ROOT only (ignored elsewhere).  The
// symmetric block (`stage_offset`, 2 * K * N * 4 bytes rounded up to 256 each: two halves used alternately, so that the
// root can stage epoch e+1 while epoch e is still being consumed) is the library's staging area: B travels through it in `panels`
// column panels (N / panels a multiple of 256; the same value on every rank), stored panel-major.  The root computes from
// B itself; every other rank runs ONE persistent GEMM launch over the staged layout, gated per tile on the panels'
// arrival (ranks whose shard is too small for a persistent kernel run one launch per panel behind stream-side waits).
// On return (in stream order) the root's B may be rewritten.  Collective over the ranks.
TB200_API int tiramisu_b200_sgemm_bcast_b(void *const *peers, uint64_t stage_offset, uint64_t flags_offset, int32_t root, uint32_t epoch,
                                          int32_t panels, const float *B, uint64_t ldB, const float *A, float *C, uint64_t rows,
                                          uint64_t N, uint64_t K, float alpha, float beta, uint64_t ldA, uint64_t ldC)
{
    if (!A || !C || K == 0 || N == 0) return fail(__func__, "bad argument");
    if (ldA == 0) ldA = K;
    if (ldB == 0) ldB = N;
    if (ldC == 0) ldC = N;
    const int world = g_comm.comm ? g_comm.world : 1, rank = g_comm.comm ? g_comm.rank : 0;
    if (world == 1) return (!B || !rows) ? (rows ? fail(__func__, "B is null on the root") : 0)
                                         : tb200::gemm_f32(A, B, C, rows, N, K, alpha, beta, ldA, ldB, ldC, false, false);
    if (bcast_sym_check(__func__, peers, root, epoch, panels)) return 1;
    if (N % panels || (panels > 1 && (N / panels) % 256)) return fail(__func__, "N / panels must be a multiple of 256");
    if (rank == root && !B) return fail(__func__, "B is null on the root");
    cudaStream_t st = stream();
    const uint64_t w = N / panels, panel_bytes = K * w * 4;
    stage_offset += (epoch & 1u) * ((K * N * 4 + 255) & ~uint64_t(255));       // two staging halves, alternating by epoch
    const float *stage = reinterpret_cast<const float *>(static_cast<char *>(peers[rank]) + stage_offset);
    const uint32_t *have = reinterpret_cast<const uint32_t *>(static_cast<char *>(peers[rank]) + flags_offset) + 2 * world;
    const uint32_t base = 64u * (epoch - 1);
    RootPack pack;
    if (rank == root) { pack.B = B; pack.K = K; pack.ldB = ldB; pack.cols = w; }
    if (bcast_sym_fork(st)) return 1;
    if (rank == root) {
        if (rows && tb200::gemm_f32(A, B, C, rows, N, K, alpha, beta, ldA, ldB, ldC, false, false)) return 1;
        if (bcast_sym_enqueue(peers, stage_offset, K * N * 4, panel_bytes, flags_offset, root, epoch, panels, pack, true)) return 1;
        return bcast_sym_join(st, peers, flags_offset, root, epoch);
    }
    const int kernel = tb200::gemm_tc_kernel_choice(rows, N, K, nullptr);
    const bool in_kernel = rows && (kernel == 2 || kernel == 3) && !getenv("TB200_GEMM_MODE") && !getenv("TB200_SGEMM_GATE_ON_STREAM") &&
                           tb200::gemm_tc_addressable(A, stage, C, rows, N, K, ldA, w, ldC, false, false);
    if (in_kernel) {
        // one launch over the staged layout, submitted before the copies (see the note on work queues above)
        tb200::gemm_tc_gate(have, base + 1, panels > 1 ? (int)w : 0, base + (uint32_t)panels);
        tb200::gemm_tc_b_panels(panels > 1 ? (int)w : 0, K * w);
        const int rc = tb200::gemm_f32_tc(A, stage, C, rows, N, K, alpha, beta, ldA, w, ldC, false);
        tb200::gemm_tc_b_panels(0, 0);
        if (rc) return 1;
        if (bcast_sym_enqueue(peers, stage_offset, K * N * 4, panel_bytes, flags_offset, root, epoch, panels, pack, true)) return 1;
        return bcast_sym_join(st, peers, flags_offset, root, epoch);
    }
    if (bcast_sym_enqueue(peers, stage_offset, K * N * 4, panel_bytes, flags_offset, root, epoch, panels, pack, true)) return 1;
    for (int p = 0; p < panels && rows; p++) {                      // every staged panel is a plain row-major [K][w] matrix
        if (tb200::stream_wait_value32(st, have, base + (uint32_t)p + 1)) return 1;
        if (tb200::gemm_f32(A, stage + (uint64_t)p * K * w, C + (uint64_t)p * w, rows, w, K, alpha, beta, ldA, w, ldC, false, false)) return 1;
    }
    return bcast_sym_join(st, peers, flags_offset, root, epoch);
}

// Prints the timeline of the last traced broadcast of this thread (TB200_BCAST_TRACE=1); synchronises the device.
TB200_API int tiramisu_b200_comm_broadcast_trace_dump(void)
{
    if (!t_trace.on || !t_trace.t0) return 0;
    TB200_CHECK(cudaDeviceSynchronize());
    std::string out = "[rank " + std::to_string(g_comm.rank) + "] broadcast timeline (ms after the fork):\n";
    for (auto &m : t_trace.marks) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t_trace.t0, m.second);
        char buf[160];
        snprintf(buf, sizeof buf, "[rank %d]   %8.3f  %s\n", g_comm.rank, ms, m.first.c_str());
        out += buf;
    }
    fputs(out.c_str(), stderr);
    return 0;
}

TB200_API int tiramisu_b200_comm_free_symmetric(void *local)
{
    SymBlock blk{nullptr, {}};
    {
        std::lock_guard<std::mutex> lk(g_comm_mu);
        for (size_t i = 0; i < g_sym.size(); i++)
            if (g_sym[i].local == local) { blk = g_sym[i]; g_sym.erase(g_sym.begin() + i); break; }
    }
    if (!blk.local) return fail(__func__, "not a symmetric block");
    TB200_CHECK(cudaStreamSynchronize(stream()));
    for (void *p : blk.opened) cudaIpcCloseMemHandle(p);
    // every rank must have unmapped before the owner frees: rendezvous through the communicator
    if (g_comm.comm) {
        if (nccl_check(nccl().AllReduce(g_comm.scratch, g_comm.scratch, 1, ncclFloat32, ncclSum, g_comm.comm, stream()), "ncclAllReduce")) return 1;
        TB200_CHECK(cudaStreamSynchronize(stream()));
    }
    TB200_CHECK(cudaFree(blk.local));
    return 0;
}

// =============================================================================== src/tiramisu_mpi.cpp names
TB200_API int tiramisu_MPI_init(void)
{
    // src/tiramisu_mpi.cpp:8-17: MPI_Init_thread(FUNNELED) and return the rank.
    if (g_comm.comm) return g_comm.rank;
    const int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1), local = env_int("LOCAL_RANK", rank);
    if (check(cudaSetDevice(local), "tiramisu_MPI_init")) return -1;
    ncclUniqueId id;
    memset(&id, 0, sizeof(id));
    if (rank == 0 && tiramisu_b200_comm_unique_id(&id)) return -1;
    if (world > 1 && exchange_id_tcp(rank, world, &id)) return -1;
    if (tiramisu_b200_comm_init(rank, world, &id)) return -1;
    return rank;
}

TB200_API void tiramisu_MPI_cleanup(void) { tiramisu_b200_comm_destroy(); }      // :19-21 MPI_Finalize (flushes queued messages)

TB200_API void tiramisu_MPI_global_barrier(void)                                 // :23 MPI_Barrier
{
    if (!g_comm.comm) return;
    flush_pending();
    if (nccl_check(nccl().AllReduce(g_comm.scratch, g_comm.scratch, 1, ncclFloat32, ncclSum, g_comm.comm, stream()),
                   "ncclAllReduce"))
        return;
    check(cudaStreamSynchronize(stream()), "tiramisu_MPI_global_barrier");
}

TB200_API int tiramisu_MPI_Comm_rank(int offset) { return g_comm.rank + offset; }   // :78-83

TB200_API void tiramisu_MPI_Wait(void *)                                            // :85-89
{
    // completes every message queued by the non-blocking calls (one NCCL group); for device buffers completion is
    // in stream order, host-buffer receives are copied back and waited for
    flush_pending();
}

// src/tiramisu_mpi.cpp:96-192 — (count, peer, tag, data[, reqs]); tags are implicit in NCCL's
// per-peer FIFO order, request handles are not needed (see tiramisu_MPI_Wait).
#define TB200_MPI_TYPE(SUFFIX, T)                                                                                     \
    TB200_API void tiramisu_MPI_Send_##SUFFIX(int count, int dest, int, T *data) { p2p(true, data, (size_t)count * sizeof(T), dest, true); }   \
    TB200_API void tiramisu_MPI_Ssend_##SUFFIX(int count, int dest, int, T *data) { p2p(true, data, (size_t)count * sizeof(T), dest, true); }  \
    TB200_API void tiramisu_MPI_Isend_##SUFFIX(int count, int dest, int, T *data, long *) { p2p(true, data, (size_t)count * sizeof(T), dest, false); }  \
    TB200_API void tiramisu_MPI_Issend_##SUFFIX(int count, int dest, int, T *data, long *) { p2p(true, data, (size_t)count * sizeof(T), dest, false); } \
    TB200_API void tiramisu_MPI_Recv_##SUFFIX(int count, int src, int, T *data) { p2p(false, data, (size_t)count * sizeof(T), src, true); }    \
    TB200_API void tiramisu_MPI_Irecv_##SUFFIX(int count, int src, int, T *data, long *) { p2p(false, data, (size_t)count * sizeof(T), src, false); }

TB200_MPI_TYPE(int8, char)
TB200_MPI_TYPE(int16, short)
TB200_MPI_TYPE(int32, int)
TB200_MPI_TYPE(int64, long)
TB200_MPI_TYPE(uint8, unsigned char)
TB200_MPI_TYPE(uint16, unsigned short)
TB200_MPI_TYPE(uint32, unsigned int)
TB200_MPI_TYPE(uint64, unsigned long)
TB200_MPI_TYPE(f32, float)
TB200_MPI_TYPE(f64, double)

// include/tiramisu/mpi_comm.h:6-8 declares these three OUTSIDE its extern "C" block, so wrappers
// such as tests/wrapper_test_160.cpp:11 reference the C++-mangled names; export those spellings too.
extern "C" __attribute__((visibility("default"), alias("tiramisu_MPI_init"))) int _Z17tiramisu_MPI_initv();
extern "C" __attribute__((visibility("default"), alias("tiramisu_MPI_cleanup"))) void _Z20tiramisu_MPI_cleanupv();
extern "C" __attribute__((visibility("default"), alias("tiramisu_MPI_global_barrier"))) void _Z27tiramisu_MPI_global_barrierv();
