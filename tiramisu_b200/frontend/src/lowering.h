// lowering.h — internal declarations of the sm_100a lowering (not installed, not part of the generator API).
//   lowering.cpp           order, loop tree, gpu axes, kernel argument collection, driver, nvcc invocation
//   lowering_emit.cpp      index simplification, expression printing, statement / kernel / host emission
//   lowering_patterns.cpp  pattern lowerings: contractions -> runtime GEMM, sgemm statement triple, heat2d pair, conv block idiom
//   lowering_dist.cpp      distributed levels (rank scope), send / recv / wait -> tiramisu_MPI_* on device addresses, message coalescing
//   lowering_retarget.cpp  CPU schedules -> GPU: constants, device mirrors, mapping, warp-shuffle reduction, loop distribution, nest fusion
//   lowering_forward.cpp   CPU register blocking undone: load / accumulate / store through a private staging buffer -> accumulate in place
#ifndef TIRAMISU_B200_FRONTEND_LOWERING_H
#define TIRAMISU_B200_FRONTEND_LOWERING_H

#include <tiramisu/core.h>
#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>

namespace tiramisu {
namespace lowering_detail {

[[noreturn]] inline void die(const std::string &msg)
{
    std::cerr << "tiramisu_b200 lowering: " << msg << std::endl;
    std::exit(1);
}

inline std::string san(const std::string &s)
{
    std::string o;
    for (char c : s) o += (isalnum((unsigned char) c) || c == '_') ? c : '_';
    return o;
}

inline const char *ctype(primitive_t t)
{
    switch (t) {
    case p_uint8: return "uint8_t"; case p_int8: return "int8_t"; case p_uint16: return "uint16_t"; case p_int16: return "int16_t";
    case p_uint32: return "uint32_t"; case p_int32: return "int32_t"; case p_uint64: return "uint64_t"; case p_int64: return "int64_t";
    case p_float32: return "float"; case p_float64: return "double"; case p_boolean: return "bool";
    default: return "uint8_t";
    }
}

struct Node {
    bool is_loop = false;
    // loop
    int id = 0;
    std::string dim_name;
    expr lo, hi;
    loop_dim::tag_t tag = loop_dim::none;
    int axis = -1;               // gpu axis (0 = x)
    bool launch_guard = false;   // launch dimension may exceed this loop's extent
    bool warp_reduce = false;    // reduction loop strided over the lanes of a warp (see Lowering::warp_reduction_terms)
    std::vector<std::unique_ptr<Node>> kids;
    // statement
    computation *comp = nullptr;
    std::vector<Node *> path;    // enclosing loops, outermost first (one per scheduled dimension)
    std::string cname() const { return "v" + std::to_string(id) + "_" + san(dim_name); }
};

struct Kernel {
    Node *root = nullptr;
    std::string name;
    std::vector<Node *> host_loops;        // enclosing host loops (passed by value)
    std::set<std::string> constants;       // function constants referenced
    std::vector<buffer *> globals;         // global buffers referenced, first-use order
    std::set<buffer *> written;
    std::vector<buffer *> locals;          // shared / local / register buffers
    expr grid[3], block[3];
};

class Lowering {
public:
    Lowering(function *f, const std::vector<buffer *> &args, bool cuda) : f_(f), args_(args), cuda_(cuda) {}
    std::string run();

private:
    function *f_;
    std::vector<buffer *> args_;
    bool cuda_;
    Node root_;
    int next_id_ = 0;
    int next_kernel_ = 0;
    std::map<std::string, computation *> constants_;
    std::vector<computation *> constant_order_;
    std::ostringstream kernels_;            // device code
    std::ostringstream protos_;             // extern "C" prototypes of o_call targets
    std::set<std::string> proto_seen_;
    bool pattern_ = true;
    bool retargeted_ = false;                 // CPU schedule mapped onto the GPU; host arguments mirrored on the device
    bool distributed_ = false;                // the function has distributed levels / communicators (lowering_dist.cpp)
    bool prepare_distribution(const std::vector<computation *> &ord);
    void emit_communicator(Node *n, std::ostringstream &os, int ind);
    std::set<buffer *> mirrored_;             // argument buffers with a device mirror (b_X device, h_X host)
    std::set<buffer *> no_copy_in_;           // outputs that are completely overwritten before being read
    std::set<buffer *> written_args_;         // mirrored arguments some statement stores into (copied back whatever their declared type)
    std::set<buffer *> direct_;               // mirrored arguments handed to a runtime entry point as halide buffers
    size_t n_statements_ = 0;
    void propagate_constants();
    void expand_lets();
    void expand_inline_computations();
    void hoist_buffer_extents();
    struct ExtentRewrite { buffer *b; expr from, to; };
    std::vector<ExtentRewrite> extent_rewrites_;
    void retarget(const std::vector<computation *> &ord);
    void sink_reduction_loops(const std::vector<computation *> &ord);
    void strip_divided_iterators(const std::vector<computation *> &ord);
    bool changed_ = false;                                              // distribute_parallel_shared_loops lowered a boundary
    bool forward_private_temporaries();                                 // lowering_forward.cpp
    bool expand_private_temporaries();
    bool drop_dead_temporaries();
    bool staging_triple(const std::vector<computation *> &ord, size_t i, bool apply);
    void mirror_buffers(const std::vector<computation *> &ord, bool gpu_only);
    void auto_map(computation *c, int shared_depth);
    void fuse_pointwise_nests(const std::vector<computation *> &ord, const std::vector<int> &shared);
    void distribute_parallel_shared_loops(const std::vector<computation *> &ord);
    void map_fused_parallel_loops(const std::vector<computation *> &ord);
    bool dim_is_parallel(computation *c, const std::string &dim);
    std::set<std::string> parallel_iterators(computation *c);     // jointly injective in the stored element
    std::map<computation *, std::set<std::string>> parallel_iters_;
    struct Access { buffer *b; std::vector<expr> idx; bool write; };
    std::vector<Access> accesses_of(computation *c);
    std::set<computation *> grid_fused_;          // statements of a fused run whose shared data-parallel loops became the grid
    std::set<computation *> plainly_mapped_;      // statements auto_map gave the standard element-per-thread mapping
    bool warp_reduction_terms(computation *c, const std::string &r, std::vector<expr> *summands);
    void emit_warp_reduction(Node *loop, std::ostringstream &os, int ind);
    static int fuse_depth(computation *prev, computation *c);
    bool try_heat2d_pattern(Node *zero, Node *heat, std::ostringstream &host, int ind);
    bool try_blurxy_pattern(Node *kx, Node *ky, const std::vector<computation *> &ord, std::ostringstream &host, int ind);
    bool try_staged_contraction(Node *kernel, std::ostringstream &host, int ind);
    bool try_conv_block_idiom(const std::vector<computation *> &ord, std::ostringstream &host);
    // index simplification context of the statement being emitted
    using Env = std::map<std::string, std::pair<int64_t, int64_t>>;   // loop dimension -> inclusive [lo, hi]
    const computation *cur_ = nullptr;
    Env env_;
    expr inline_defs(const expr &e) const;
    static bool bounds(const expr &e, const Env &env, int64_t &lo, int64_t &hi);
    static expr simp(const expr &e, const Env &env);
    void enter_stmt(computation *c);

    // 1-2
    std::vector<computation *> order();
    void build_tree(const std::vector<computation *> &ord);
    // 3
    void assign_axes(Node *n);
    void coalesce_thread_axes(Node *kernel_root);
    int count_below(const Node *n, loop_dim::tag_t t) const;
    void collect_kernel(Node *n, Kernel &k, bool inside);
    void note_expr(const expr &e, Kernel &k, bool is_store_target);
    buffer *buffer_of_access(const expr &e, computation **who = nullptr);
    // printing
    std::string print(const expr &e, bool device);
    std::string print_access(const expr &e, bool device);
    std::string lin_index(buffer *b, const std::vector<expr> &idx, bool device);
    std::string literal(const expr &e);
    std::string bytes_of(buffer *b, bool device);
    // emission
    void emit_host(Node *n, std::ostringstream &os, int ind, std::vector<Node *> &host_loops);
    void emit_host_stmt(Node *n, std::ostringstream &os, int ind);
    void emit_kernel(Node *n, std::ostringstream &host, int ind, std::vector<Node *> &host_loops);
    void emit_device(Node *n, Kernel &k, std::ostringstream &os, int ind);
    bool try_emit_accumulation_tile(Node *n, std::ostringstream &os, int ind);
    void emit_assignment(Node *n, std::ostringstream &os, int ind, bool device);
    void emit_bindings(Node *n, std::ostringstream &os, int ind, std::vector<std::string> &conds, bool device);
    struct Contraction { buffer *C = nullptr, *A = nullptr, *B = nullptr; bool tA = false, tB = false; expr M, N, K, offA, offB, offC; };
    bool match_contraction(computation *c, Contraction &m);
    bool match_contractions(computation *c, std::vector<Contraction> &out, std::vector<std::string> *outer);
    bool match_zero_fill(computation *z, buffer *C, const expr &M, const expr &N);
    void emit_sgemm_call(std::ostringstream &host, int ind, const Contraction &m, buffer *C, const std::string &alpha, const std::string &beta);
    bool try_gemm_pattern(Kernel &k, std::ostringstream &host, int ind);
    bool try_sgemm_triple(Node *kz, Node *kr, Node *ks, const std::vector<computation *> &ord, std::ostringstream &host, int ind);
    void need_proto(const expr &call);
};

inline std::string pad(int n) { return std::string((size_t) n * 4, ' '); }
bool affine_form(const expr &e, std::map<std::string, int64_t> &coef, int64_t &cst, int64_t scale = 1);
std::vector<Node *> statements_of(Node *root);

}  // namespace lowering_detail
}  // namespace tiramisu

#endif
