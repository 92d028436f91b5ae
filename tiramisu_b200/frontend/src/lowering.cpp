// lowering.cpp — `tiramisu::codegen(args, "x.o", true)` for sm_100a.
//
// Replaces the reference's generator-time back end for GPU functions: the cuda_ast emitter
// (/root/reference/src/tiramisu_codegen_cuda.cpp: kernel extraction :548-640, kernel wrappers
// :855-866, nvcc driver :1100-1200) and the Halide host-code generator
// (/root/reference/src/tiramisu_codegen_halide.cpp:2488-2603 temporaries, :1840-1859 runtime calls).
//
// Pipeline (all of it this file):
//   1. order the scheduled computations by their then()/after() chain;
//   2. build ONE loop tree: consecutive computations share the loops down to their fuse level, a
//      shared loop takes the union of the ranges and every statement re-checks its own range;
//   3. every outermost gpu-tagged loop becomes a kernel; block/thread loops become blockIdx/threadIdx
//      (x,y,z from the innermost tagged loop outwards, like the reference) so the innermost — unit
//      stride — iterator is threadIdx.x and global accesses coalesce; shared/local/register buffers
//      become __shared__ / per-thread arrays; read-only global buffers are `const __restrict__`;
//   4. pattern lowering: a kernel that is exactly one fp32 contraction C(i,j) += A(i,k)*B(k,j) over
//      row-major global buffers is not emitted as SIMT code at all — it becomes a call of the
//      runtime's tcgen05 GEMM entry point (tiramisu_cublas_sgemm in libtiramisu_b200.so);
//   5. the host function keeps the reference ABI `extern "C" int f(halide_buffer_t*...)`; copies,
//      allocation, synchronisation and launches go through the tiramisu_cuda_* runtime;
//   6. nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (x86 Tiramisu code has no FMA
//      contraction, src/tiramisu_codegen_halide.cpp:3988-3995) compiles the result to the object.
#include "lowering.h"

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

// ------------------------------------------------------------------------------------------------ order
std::vector<computation *> Lowering::order()
{
    std::vector<computation *> ord;
    std::set<computation *> seen;
    for (computation *c : f_->get_computations()) {
        if (c->is_input() || (!c->is_constant() && !c->should_schedule_this_computation())) continue;
        if (c->pred() != nullptr) continue;
        // function constants may be links of the chain (reference tests/test_93.cpp orders its statements after its
        // constants): they are evaluated at function entry anyway, so they are walked through and left out
        if (c->is_constant() && !c->succ()) continue;
        for (computation *p = c; p; p = p->succ()) {
            if (!seen.insert(p).second) die("cycle in the then()/after() order at '" + p->get_name() + "'");
            if (!p->is_constant() && p->should_schedule_this_computation()) ord.push_back(p);
        }
    }
    for (computation *c : f_->get_computations())
        if (!c->is_input() && !c->is_constant() && c->should_schedule_this_computation() && !seen.count(c))
            die("computation '" + c->get_name() + "' is not reachable in the then()/after() order");
    return ord;
}

// number of loops `c` shares with its predecessor in the order
int Lowering::fuse_depth(computation *prev, computation *c)
{
    if (!prev || c->pred() != prev) return 0;
    int depth = 0;
    if (c->pred_level_index() >= 0) depth = c->pred_level_index() + 1;
    else if (c->pred_level_index() == -2) {
        int d = c->find_dim(c->pred_level());
        if (d < 0) d = prev->find_dim(c->pred_level());
        if (d < 0) die("'" + c->get_name() + "' is ordered after '" + prev->get_name() + "' at unknown level '" + c->pred_level() + "'");
        depth = d + 1;
    }
    return std::min(depth, std::min((int) prev->dims().size(), (int) c->dims().size()));
}

static expr union_lo(const expr &a, const expr &b) { return a.is_equal(b) ? a : expr(o_min, a, b); }
static expr union_hi(const expr &a, const expr &b) { return a.is_equal(b) ? a : expr(o_max, a, b); }

void Lowering::build_tree(const std::vector<computation *> &ord)
{
    std::vector<Node *> open;
    computation *prev = nullptr;
    for (computation *c : ord) {
        if (!c->get_expr().is_defined()) die("computation '" + c->get_name() + "' has no expression");
        auto &dims = c->dims();
        int depth = fuse_depth(prev, c);
        // a level deeper than the predecessor's own nest only orders the two (reference generators write
        // `copy.then(kernel_stmt, i1)` after a loop-less copy): share what both have
        depth = std::min(depth, std::min((int) open.size(), (int) dims.size()));
        open.resize(depth);
        for (int d = 0; d < depth; d++) {
            Node *n = open[d];
            n->lo = union_lo(n->lo, dims[d].lo);
            n->hi = union_hi(n->hi, dims[d].hi);
            if (n->tag == loop_dim::none) n->tag = dims[d].tag;
        }
        for (int d = depth; d < (int) dims.size(); d++) {
            auto n = std::make_unique<Node>();
            n->is_loop = true; n->id = next_id_++; n->dim_name = dims[d].name;
            n->lo = dims[d].lo; n->hi = dims[d].hi; n->tag = dims[d].tag; n->warp_reduce = dims[d].warp_reduce;
            Node *raw = n.get();
            (open.empty() ? &root_ : open.back())->kids.push_back(std::move(n));
            open.push_back(raw);
        }
        auto s = std::make_unique<Node>();
        s->comp = c; s->path = open; s->id = next_id_++;
        (open.empty() ? &root_ : open.back())->kids.push_back(std::move(s));
        prev = c;
    }
}

// ------------------------------------------------------------------------------------------------ gpu axes
int Lowering::count_below(const Node *n, loop_dim::tag_t t) const
{
    int best = 0;
    for (auto &k : n->kids) {
        if (!k->is_loop) continue;
        int c = count_below(k.get(), t) + (k->tag == t ? 1 : 0);
        best = std::max(best, c);
    }
    return best;
}

void Lowering::assign_axes(Node *n)
{
    if (n->is_loop && (n->tag == loop_dim::gpu_block || n->tag == loop_dim::gpu_thread)) {
        n->axis = count_below(n, n->tag);
        if (n->axis > 2) die("more than three nested gpu " + std::string(n->tag == loop_dim::gpu_block ? "block" : "thread") + " levels at '" + n->dim_name + "'");
    }
    for (auto &k : n->kids) assign_axes(k.get());
}

buffer *Lowering::buffer_of_access(const expr &e, computation **who)
{
    computation *c = f_->find_computation(e.get_name());
    if (!c) die("access to unknown computation '" + e.get_name() + "'");
    if (who) *who = c;
    if (c->is_constant()) return nullptr;
    return c->get_buffer();
}

void Lowering::note_expr(const expr &e, Kernel &k, bool)
{
    if (!e.is_defined()) return;
    if (e.get_expr_type() == e_var) {
        if (constants_.count(e.get_name())) k.constants.insert(e.get_name());
        return;
    }
    if (e.get_expr_type() != e_op) return;
    if (e.get_op_type() == o_access) {
        computation *c = nullptr;
        buffer *b = buffer_of_access(e, &c);
        if (!b) k.constants.insert(c->get_name());
        else {
            auto loc = b->get_location();
            auto &list = (loc == mem_loc::gpu_global || loc == mem_loc::gpu_constant || loc == mem_loc::host) ? k.globals : k.locals;
            if (std::find(list.begin(), list.end(), b) == list.end()) list.push_back(b);
            for (const expr &d : b->get_dim_sizes()) note_expr(d, k, false);
            for (const expr &a : c->access()) note_expr(a, k, false);
        }
    }
    if (e.get_op_type() == o_allocate || e.get_op_type() == o_buffer) {
        buffer *b = f_->find_buffer(e.get_name());
        if (b && (b->get_location() == mem_loc::gpu_shared || b->get_location() == mem_loc::gpu_local || b->get_location() == mem_loc::gpu_register))
            if (std::find(k.locals.begin(), k.locals.end(), b) == k.locals.end()) k.locals.push_back(b);
    }
    for (const expr &o : e.get_operands()) note_expr(o, k, false);
}

void Lowering::collect_kernel(Node *n, Kernel &k, bool)
{
    if (n->is_loop) {
        note_expr(n->lo, k, false);
        note_expr(n->hi, k, false);
        for (auto &c : n->kids) collect_kernel(c.get(), k, true);
        return;
    }
    computation *c = n->comp;
    note_expr(c->get_expr(), k, false);
    note_expr(c->get_predicate(), k, false);
    for (auto &d : c->dims()) { note_expr(d.lo, k, false); note_expr(d.hi, k, false); }
    for (auto &d : c->defs()) note_expr(d.second, k, false);
    for (auto &g : c->guards()) note_expr(g, k, false);
    const expr &e = c->get_expr();
    bool is_stmt_only = e.get_expr_type() == e_sync || (e.get_expr_type() == e_op && (e.get_op_type() == o_allocate || e.get_op_type() == o_free));
    if (!is_stmt_only) {
        buffer *b = c->get_buffer();
        auto loc = b->get_location();
        auto &list = (loc == mem_loc::gpu_shared || loc == mem_loc::gpu_local || loc == mem_loc::gpu_register) ? k.locals : k.globals;
        if (std::find(list.begin(), list.end(), b) == list.end()) list.push_back(b);
        k.written.insert(b);
        for (const expr &d : b->get_dim_sizes()) note_expr(d, k, false);
        for (const expr &a : c->access()) note_expr(a, k, false);
    }
}

// ------------------------------------------------------------------------------------------------ driver
std::string Lowering::run()
{
    if (const char *m = getenv("TIRAMISU_B200_LOWERING")) pattern_ = std::string(m) != "generic";
    for (computation *c : f_->get_computations())
        if (c->pending_bounds() && c->should_schedule_this_computation())
            die("computation '" + c->get_name() + "' was declared without iterator bounds and function::compute_bounds() was not called");
    {
        // distributed functions: the `rank` constant has to exist before the constants are collected; the order is only
        // needed to look at the scheduled communicators (it is recomputed below)
        // literal function constants first: the message coalescing of prepare_distribution needs literal loop extents
        // (blurautodist_ref.cpp sends rows of `COLS + 2` elements with COLS a constant: 2004 one-element messages per rank
        // and call before this, one message after)
        if (!getenv("TIRAMISU_B200_LATE_CONSTANTS")) propagate_constants();
        std::vector<computation *> pre = order();
        prepare_distribution(pre);
    }
    expand_lets();
    expand_inline_computations();
    hoist_buffer_extents();
    propagate_constants();
    for (computation *c : f_->get_computations())
        if (c->is_constant() && !c->is_let()) { constants_[c->get_name()] = c; constant_order_.push_back(c); }
    for (buffer *b : args_)
        if (b->get_location() != mem_loc::host && b->get_location() != mem_loc::gpu_global)
            die("argument buffer '" + b->get_name() + "' is tagged on-chip");

    if (pattern_ && !cuda_) { drop_dead_temporaries(); forward_private_temporaries(); expand_private_temporaries(); }
    std::vector<computation *> ord = order();
    n_statements_ = ord.size();
    bool any_gpu = false;
    for (computation *c : ord)
        for (auto &d : c->dims()) any_gpu |= d.tag == loop_dim::gpu_block || d.tag == loop_dim::gpu_thread;
    std::ostringstream idiom_body;
    // TIRAMISU_B200_NO_IDIOM=1: skip the whole-function idiom and let the general passes lower the block
    const bool whole_function_idiom = !cuda_ && !any_gpu && !getenv("TIRAMISU_B200_NO_IDIOM") && try_conv_block_idiom(ord, idiom_body);
    if (whole_function_idiom) {
        std::cerr << "tiramisu_b200 lowering: '" << f_->get_name() << "' is the conv + ReLU + max-pool block: lowered to the runtime's fused kernel\n";
    } else if (!cuda_ && !any_gpu) {
        std::cerr << "tiramisu_b200 lowering: '" << f_->get_name() << "' is a CPU schedule; there is no CPU code generator: "
                     "re-targeting its loop nests to sm_100a (parallel -> blocks, vector -> threads, device mirrors for host buffers)\n";
        retarget(ord);
    } else if (any_gpu) {
        mirror_buffers(ord, true);          // gpu-tagged statements over plain host buffers (no memcpy computations)
    }
    if (!whole_function_idiom) build_tree(ord);
    assign_axes(&root_);

    std::ostringstream body;
    body << idiom_body.str();
    std::vector<Node *> host_loops;
    for (size_t i = 0; i < root_.kids.size(); i++) {
        Node *k = root_.kids[i].get();
        auto is_kernel = [](Node *n) { return n->is_loop && (n->tag == loop_dim::gpu_block || n->tag == loop_dim::gpu_thread); };
        if (i + 1 < root_.kids.size() && is_kernel(k) && is_kernel(root_.kids[i + 1].get()) &&
            try_heat2d_pattern(k, root_.kids[i + 1].get(), body, 1)) {
            i++;
            continue;
        }
        if (i + 1 < root_.kids.size() && is_kernel(k) && is_kernel(root_.kids[i + 1].get()) &&
            try_blurxy_pattern(k, root_.kids[i + 1].get(), ord, body, 1)) {
            i++;
            continue;
        }
        if (i + 2 < root_.kids.size() && is_kernel(k) && is_kernel(root_.kids[i + 1].get()) && is_kernel(root_.kids[i + 2].get()) &&
            try_sgemm_triple(k, root_.kids[i + 1].get(), root_.kids[i + 2].get(), ord, body, 1)) {
            i += 2;
            continue;
        }
        if (is_kernel(k) && try_staged_contraction(k, body, 1)) continue;
        emit_host(k, body, 1, host_loops);
    }

    std::ostringstream out;
    out << "// Generated by the tiramisu_b200 sm_100a lowering for Tiramisu function '" << f_->get_name() << "'. Do not edit.\n"
        << "#include <cstdint>\n#include <cstdlib>\n#include <cstring>\n#include <cmath>\n#include <type_traits>\n#include <cuda_runtime.h>\n\n"
        << "#ifndef HALIDE_HALIDERUNTIME_H\n"
        << "struct halide_dimension_t { int32_t min, extent, stride; uint32_t flags; };\n"
        << "struct halide_type_t { uint8_t code, bits; uint16_t lanes; };\n"
        << "struct halide_buffer_t { uint64_t device; const void *device_interface; uint8_t *host; uint64_t flags; halide_type_t type;\n"
        << "                         int32_t dimensions; halide_dimension_t *dim; void *padding; };\n"
        << "#endif\n"
        << "extern \"C\" {\n"
        << "int tiramisu_cuda_memcpy_to_device(void *, void *, uint64_t);\nint tiramisu_cuda_memcpy_to_host(void *, void *, uint64_t);\n"
        << "void *tiramisu_cuda_malloc(uint64_t);\nint tiramisu_cuda_free(void *);\nint32_t tiramisu_cuda_stream_synchronize(int32_t);\n"
        << "void *tiramisu_b200_get_stream(void);\nint tiramisu_b200_post_launch(const char *);\n"
        << "int tiramisu_cuda_memcpy_to_device_async(void *, const void *, uint64_t);\n"
        << "int tiramisu_b200_heat2d(const float *, float *, int32_t, int32_t, int32_t, int32_t, int32_t);\n"
        << "int tiramisu_b200_blurxy_u8(const uint8_t *, uint8_t *, int32_t, int32_t, int32_t);\n"
        << "int tiramisu_b200_heat2d_buffers(halide_buffer_t *, halide_buffer_t *, int32_t, int32_t, int32_t);\n"
        << "int tiramisu_b200_conv_relu_maxpool(const float *, const float *, const float *, float *, int32_t, int32_t, int32_t, int32_t);\n"
        << "int tiramisu_cublas_sgemm(float *, float *, float *, uint64_t, uint64_t, uint64_t, float, float, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, bool, bool);\n"
        << "int tiramisu_cublas_dgemm(double *, double *, double *, uint64_t, uint64_t, uint64_t, double, double, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, bool, bool);\n"
        << (distributed_ ? "int tiramisu_MPI_Comm_rank(int);\nvoid tiramisu_MPI_Wait(void *);\nvoid tiramisu_MPI_global_barrier(void);\n"
                           "#define TB_MPI_DECL(S, T) void tiramisu_MPI_Send_##S(int, int, int, T *); void tiramisu_MPI_Ssend_##S(int, int, int, T *); \\\n"
                           "    void tiramisu_MPI_Isend_##S(int, int, int, T *, long *); void tiramisu_MPI_Issend_##S(int, int, int, T *, long *); \\\n"
                           "    void tiramisu_MPI_Recv_##S(int, int, int, T *); void tiramisu_MPI_Irecv_##S(int, int, int, T *, long *);\n"
                           "TB_MPI_DECL(int8, char) TB_MPI_DECL(uint8, unsigned char) TB_MPI_DECL(int16, short) TB_MPI_DECL(uint16, unsigned short)\n"
                           "TB_MPI_DECL(int32, int) TB_MPI_DECL(uint32, unsigned int) TB_MPI_DECL(int64, long) TB_MPI_DECL(uint64, unsigned long)\n"
                           "TB_MPI_DECL(f32, float) TB_MPI_DECL(f64, double)\n"
                         : "")
        << "}\n"
        << "template <typename T> __host__ __device__ __forceinline__ T tb_max(T a, T b) { return a > b ? a : b; }\n"
        << "template <typename T> __host__ __device__ __forceinline__ T tb_min(T a, T b) { return a < b ? a : b; }\n"
        << "template <typename T> __host__ __device__ __forceinline__ T tb_abs(T a) { return a < 0 ? -a : a; }\n"
        << "// floor division / Euclidean remainder (Halide Div / Mod semantics; 0 for a zero divisor, like Halide)\n"
        << "template <typename T> __host__ __device__ __forceinline__ T tb_floordiv(T a, T b) {\n"
        << "    if (b == 0) return 0;\n    T r = a % b;\n    if (r < 0) r += (b < 0 ? -b : b);\n    return (a - r) / b;\n}\n"
        << "template <typename T> __host__ __device__ __forceinline__ T tb_floormod(T a, T b) {\n"
        << "    if (b == 0) return 0;\n    T r = a % b;\n    return r < 0 ? r + (b < 0 ? -b : b) : r;\n}\n"
        << "// Halide's lerp: an unsigned integer weight is a fixed-point fraction of its type's maximum\n"
        << "template <typename T, typename W> __host__ __device__ __forceinline__ T tb_lerp(T a, T b, W w) {\n"
        << "    const double wf = std::is_floating_point<W>::value ? (double) w : (double) w / (double) (W) ~(W) 0;\n"
        << "    const double r = (double) a * (1.0 - wf) + (double) b * wf;\n"
        << "    return std::is_floating_point<T>::value ? (T) r : (T) (r + 0.5);\n}\n"
        << "static inline void *tb_arg(halide_buffer_t *b) { return b->host ? (void *) b->host : (void *) b->device; }\n"
        << "// bytes to mirror for a host argument: what the caller's halide_buffer_t describes when it is larger than the\n"
        << "// generator's declaration (reference tests declare buffer(10) and run on 100 elements) and its metadata is sane\n"
        << "static inline uint64_t tb_arg_bytes(const halide_buffer_t *b, uint64_t declared) {\n"
        << "    if (b->dimensions <= 0 || b->dimensions > 8 || !b->dim || b->type.bits == 0 || b->type.bits > 64) return declared;\n"
        << "    uint64_t n = (b->type.bits + 7) / 8;\n"
        << "    for (int i = 0; i < b->dimensions; i++) { if (b->dim[i].extent <= 0 || b->dim[i].extent > (1 << 30)) return declared; n *= (uint64_t) b->dim[i].extent; }\n"
        << "    return n > declared && n <= (declared << 10) + (1u << 20) ? n : declared;\n}\n\n";
    out << protos_.str() << "\n" << kernels_.str();

    out << "extern \"C\" int " << f_->get_name() << "(";
    for (size_t i = 0; i < args_.size(); i++) out << (i ? ", " : "") << "halide_buffer_t *arg" << i;
    out << ")\n{\n";
    std::set<buffer *> is_arg(args_.begin(), args_.end());
    for (size_t i = 0; i < args_.size(); i++) {
        const std::string T = ctype(args_[i]->get_elements_type()), N = san(args_[i]->get_name());
        if (mirrored_.count(args_[i])) {
            // host argument of a re-targeted CPU schedule: h_X is the caller's memory, b_X its device mirror
            // (a device-resident argument, ->host == NULL, is used in place)
            out << pad(1) << T << " *h_" << N << " = (" << T << " *) arg" << i << "->host;\n";
            out << pad(1) << T << " *b_" << N << " = h_" << N << " ? nullptr : (" << T << " *) arg" << i << "->device;\n";
        } else {
            out << pad(1) << T << " *b_" << N << " = (" << T << " *) tb_arg(arg" << i << ");\n";
        }
    }
    for (buffer *b : f_->get_buffers())
        if (!is_arg.count(b) && (b->get_location() == mem_loc::host || b->get_location() == mem_loc::gpu_global))
            out << pad(1) << ctype(b->get_elements_type()) << " *b_" << san(b->get_name()) << " = nullptr;\n";
    for (computation *c : constant_order_)
        out << pad(1) << "const " << ctype(c->get_data_type()) << " c_" << san(c->get_name()) << " = (" << ctype(c->get_data_type()) << ") "
            << print(c->get_expr(), false) << ";\n";
    // temporaries: allocated at entry, freed at exit (reference src/tiramisu_codegen_halide.cpp:2488-2530)
    std::vector<buffer *> autos;
    for (buffer *b : f_->get_buffers())
        if (!is_arg.count(b) && b->get_auto_allocate() && (b->get_location() == mem_loc::host || b->get_location() == mem_loc::gpu_global))
            autos.push_back(b);
    autos.erase(std::remove_if(autos.begin(), autos.end(), [](buffer *b) { return b->get_elements_type() == p_wait_ptr; }), autos.end());   // request slots: the runtime queues
    for (buffer *b : autos)
        out << pad(1) << "b_" << san(b->get_name()) << " = (" << ctype(b->get_elements_type()) << " *) "
            << (b->get_location() == mem_loc::gpu_global ? "tiramisu_cuda_malloc(" : "malloc(") << bytes_of(b, false) << ");\n";
    for (buffer *b : args_) {
        if (!mirrored_.count(b) || direct_.count(b)) continue;
        const std::string N = san(b->get_name());
        const int ai = (int) (std::find(args_.begin(), args_.end(), b) - args_.begin());
        out << pad(1) << "const uint64_t nb_" << N << " = tb_arg_bytes(arg" << ai << ", " << bytes_of(b, false) << ");\n";
        out << pad(1) << "if (h_" << N << ") {\n" << pad(2) << "b_" << N << " = (" << ctype(b->get_elements_type()) << " *) tiramisu_cuda_malloc(nb_" << N << ");\n";
        if (!no_copy_in_.count(b))
            out << pad(2) << "tiramisu_cuda_memcpy_to_device_async(b_" << N << ", h_" << N << ", nb_" << N << ");\n";
        out << pad(1) << "}\n";
    }
    // The statements run inside a lambda so that a failed launch (`return 1` in the middle of the body) still reaches
    // the common exit below: device mirrors and temporaries are released on every path, outputs are copied back
    // only when everything succeeded.
    // Host loops around kernels (the lowered LSTM: seq_length x layers x 4 launches = 1600 per call, r01: 19 ms of which most
    // is launch latency) are captured ONCE into a CUDA graph and replayed: the graph is kept per thread, keyed by every
    // device address and every function constant the statements can see, so a later call with the same operands costs one
    // graph launch.  Only when everything in the body is a kernel launch or a runtime GEMM call on the thread's stream.
    bool host_graph = false;
    {
        std::function<void(Node *, int64_t, int64_t &, bool &)> scan = [&](Node *n, int64_t trips, int64_t &launches, bool &pure) {
            if (!n->is_loop) { pure = false; return; }                                   // a host-level statement
            if (n->tag == loop_dim::gpu_block || n->tag == loop_dim::gpu_thread) { launches += trips; return; }
            if (n->tag == loop_dim::distributed) { pure = false; return; }
            expr ext = n->hi - n->lo;
            const int64_t t = ext.is_int_literal() ? std::max<int64_t>(1, ext.get_int_val()) : 8;   // symbolic extent: assume a few
            for (auto &k : n->kids) scan(k.get(), trips * t, launches, pure);
        };
        int64_t launches = 0;
        bool pure = !whole_function_idiom && !distributed_;
        for (auto &k : root_.kids) scan(k.get(), 1, launches, pure);
        host_graph = pure && launches >= 32;
        if (const char *e = getenv("TIRAMISU_B200_HOST_GRAPH")) host_graph = pure && e[0] == '1';
    }
    if (!host_graph) {
        out << pad(1) << "const int tb_rc = [&]() -> int {\n" << body.str() << pad(1) << "return 0;\n" << pad(1) << "}();\n";
    } else {
        out << pad(1) << "auto tb_body = [&]() -> int {\n" << body.str() << pad(1) << "return 0;\n" << pad(1) << "};\n";
        out << pad(1) << "int tb_rc = 0;\n" << pad(1) << "{   // host loops of launches -> one CUDA graph, captured once per operand set and replayed\n";
        out << pad(2) << "static thread_local cudaGraphExec_t tb_exec = nullptr;\n" << pad(2) << "static thread_local uint64_t tb_key = 0;\n";
        out << pad(2) << "uint64_t key = 1469598103934665603ull;\n" << pad(2) << "auto mix = [&](uint64_t v) { key = (key ^ v) * 1099511628211ull; };\n";
        for (buffer *b : f_->get_buffers())
            if (b->get_location() == mem_loc::host || b->get_location() == mem_loc::gpu_global)
                out << pad(2) << "mix((uint64_t) (uintptr_t) b_" << san(b->get_name()) << ");\n";
        for (computation *c : constant_order_)
            out << pad(2) << "{ uint64_t bits = 0; const auto v = c_" << san(c->get_name()) << "; memcpy(&bits, &v, sizeof(v) < 8 ? sizeof(v) : 8); mix(bits); }\n";
        out << pad(2) << "cudaStream_t st = (cudaStream_t) tiramisu_b200_get_stream();\n";
        out << pad(2) << "if (getenv(\"TB200_NO_HOST_GRAPH\")) tb_rc = tb_body();\n";
        out << pad(2) << "else if (tb_exec && tb_key == key) { if (cudaGraphLaunch(tb_exec, st) != cudaSuccess) tb_rc = 1; }\n";
        out << pad(2) << "else {\n";
        out << pad(3) << "if (tb_exec) { cudaGraphExecDestroy(tb_exec); tb_exec = nullptr; }\n";
        out << pad(3) << "cudaGraph_t g = nullptr;\n";
        out << pad(3) << "if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); tb_rc = tb_body(); }\n";
        out << pad(3) << "else {\n";
        out << pad(4) << "tb_rc = tb_body();\n";
        out << pad(4) << "const cudaError_t e = cudaStreamEndCapture(st, &g);\n";
        out << pad(4) << "if (tb_rc == 0 && e == cudaSuccess && g && cudaGraphInstantiate(&tb_exec, g, 0) == cudaSuccess) {\n";
        out << pad(5) << "tb_key = key;\n" << pad(5) << "if (cudaGraphLaunch(tb_exec, st) != cudaSuccess) tb_rc = 1;\n";
        out << pad(4) << "} else {          // could not be captured: run the statements directly\n";
        out << pad(5) << "tb_exec = nullptr;\n" << pad(5) << "cudaGetLastError();\n" << pad(5) << "if (tb_rc == 0) tb_rc = tb_body();\n" << pad(4) << "}\n";
        out << pad(4) << "if (g) cudaGraphDestroy(g);\n";
        out << pad(3) << "}\n" << pad(2) << "}\n" << pad(1) << "}\n";
    }
    for (buffer *b : args_) {
        if (!mirrored_.count(b) || direct_.count(b)) continue;
        const std::string N = san(b->get_name());
        out << pad(1) << "if (h_" << N << ") {\n";
        if (b->get_argument_type() == a_output || written_args_.count(b))
            out << pad(2) << "if (tb_rc == 0) tiramisu_cuda_memcpy_to_host(h_" << N << ", b_" << N << ", nb_" << N << ");\n";
        out << pad(2) << "tiramisu_cuda_free(b_" << N << ");\n" << pad(1) << "}\n";
    }
    for (auto it = autos.rbegin(); it != autos.rend(); ++it)
        out << pad(1) << ((*it)->get_location() == mem_loc::gpu_global ? "tiramisu_cuda_free(" : "free(") << "b_" << san((*it)->get_name()) << ");\n";
    // temporaries with an explicit allocate() that the failing path never reached the free() of
    for (buffer *b : f_->get_buffers())
        if (!is_arg.count(b) && !b->get_auto_allocate() && (b->get_location() == mem_loc::host || b->get_location() == mem_loc::gpu_global))
            out << pad(1) << "if (tb_rc && b_" << san(b->get_name()) << ") " << (b->get_location() == mem_loc::gpu_global ? "tiramisu_cuda_free(" : "free(")
                << "b_" << san(b->get_name()) << ");\n";
    out << pad(1) << "return tb_rc;\n}\n\n";
    out << "extern \"C\" int " << f_->get_name() << "_argv(void **args)\n{\n" << pad(1) << "return " << f_->get_name() << "(";
    for (size_t i = 0; i < args_.size(); i++) out << (i ? ", " : "") << "(halide_buffer_t *) args[" << i << "]";
    out << ");\n}\n";
    return out.str();
}

}  // namespace lowering_detail

using lowering_detail::Lowering;
using lowering_detail::die;

void function::codegen(const std::vector<buffer *> &arguments, const std::string obj_filename, bool gen_cuda_stmt)
{
    set_arguments(arguments);
    resolve_static_order();
    Lowering l(this, arguments, gen_cuda_stmt);
    std::string src = l.run();
    std::string cu = obj_filename + ".b200.cu";
    {
        std::ofstream o(cu);
        if (!o) die("cannot write " + cu);
        o << src;
    }
    const char *nvcc = getenv("NVCC");
    const char *extra = getenv("TIRAMISU_B200_NVCC_FLAGS");
    std::string cmd = std::string(nvcc ? nvcc : "/usr/local/cuda/bin/nvcc") +
                      " -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC " +
                      (extra ? extra : "") + " -c " + cu + " -o " + obj_filename;
    // the image's $CC/$CXX wrappers are not usable nvcc host compilers
    cmd = "env -u CC -u CXX " + cmd;
    int rc = std::system(cmd.c_str());
    if (rc != 0) die("nvcc failed (" + std::to_string(rc) + ") on " + cu);
}

}  // namespace tiramisu
