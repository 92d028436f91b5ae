// tiramisu/core.h — function / buffer / computation API of the tiramisu_b200 front-end.
//
// Same class names, constructor shapes and scheduling commands as the reference's modern API
// (/root/reference/include/tiramisu/core.h: function :189, buffer :1424, computation :1863,
// input :4952, constant :5212; init/codegen :62-75), so the reference's generator sources
// (tests/test_16x.cpp, test_197.cpp, tutorials/developers/tutorial_04gpu, sgemm/gpu/generator*.cpp)
// compile against it unmodified.  What is behind the API is different by design:
//   * iteration domains are rectangular boxes of named loop dimensions (no ISL); split / tile /
//     interchange / gpu_tile rewrite that list and record how the original iterators are rebuilt;
//   * `codegen(args, "x.o", true)` lowers the ordered loop tree to sm_100a CUDA + a host function
//     with the halide_buffer_t ABI and compiles it with nvcc (src/lowering*.cpp) — it replaces the
//     reference's cuda_ast emitter (src/tiramisu_codegen_cuda.cpp) and Halide host code generator.
// The ISL-string constructors (`computation("{S[i,j]: ...}", ...)`) are not supported.
#ifndef TIRAMISU_B200_FRONTEND_CORE_H
#define TIRAMISU_B200_FRONTEND_CORE_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include <tiramisu/expr.h>
#include <tiramisu/type.h>

namespace tiramisu {

class computation;
class buffer;
class function;
class input;
class constant;
class communicator;
class send;
class recv;
class send_recv;
class wait;

/** Properties of a transfer (reference core.h:88-99, :5575-5608). */
enum xfer_attr { SYNC, ASYNC, BLOCK, NONBLOCK, MPI, CUDA, CPU2CPU, CPU2GPU, GPU2CPU, GPU2GPU };

class xfer_prop {
public:
    xfer_prop() {}
    xfer_prop(primitive_t d_type, std::initializer_list<xfer_attr> attrs) : attrs_(attrs), dtype_(d_type) {}
    xfer_prop(primitive_t d_type, std::initializer_list<xfer_attr> attrs, int xfer_prop_id) : attrs_(attrs), dtype_(d_type), id_(xfer_prop_id) {}
    primitive_t get_dtype() const { return dtype_; }
    int get_xfer_prop_id() const { return id_; }
    bool contains_attr(xfer_attr a) const { return std::find(attrs_.begin(), attrs_.end(), a) != attrs_.end(); }
    bool contains_attrs(std::vector<xfer_attr> as) const
    {
        for (xfer_attr a : as) if (!contains_attr(a)) return false;
        return true;
    }
    void add_attr(xfer_attr a) { attrs_.push_back(a); }

private:
    std::vector<xfer_attr> attrs_;
    primitive_t dtype_ = p_none;
    int id_ = -1;
};

/** What computation::create_xfer returns (reference core.h:101-105). */
struct xfer {
    tiramisu::send *s = nullptr;
    tiramisu::recv *r = nullptr;
    tiramisu::send_recv *sr = nullptr;
};

/** Loop level, first iteration, last iteration (-1: the whole loop), number of elements (reference core.h:110). */
typedef std::tuple<int, expr, expr, expr> collapse_group;

void init(std::string name);
void init();
void codegen(const std::vector<buffer *> &arguments, const std::string obj_filename, bool gen_cuda_stmt = false);

/** Where a buffer lives (reference: cuda_ast::memory_location, include/tiramisu/cuda_ast.h). */
enum class mem_loc { host, gpu_global, gpu_shared, gpu_local, gpu_register, gpu_constant };

class buffer {
public:
    buffer();
    buffer(std::string name, std::vector<expr> dim_sizes, primitive_t type, argument_t argt,
           function *fct = global::get_implicit_function());

    const std::string &get_name() const { return name_; }
    const std::vector<expr> &get_dim_sizes() const { return dims_; }
    int get_n_dims() const { return (int) dims_.size(); }
    primitive_t get_elements_type() const { return dtype_; }
    argument_t get_argument_type() const { return argt_; }
    void set_argument_type(argument_t a) { argt_ = a; }
    mem_loc get_location() const { return loc_; }
    bool get_auto_allocate() const { return auto_allocate_; }
    void set_auto_allocate(bool b) { auto_allocate_ = b; }

    void tag_gpu_global() { loc_ = mem_loc::gpu_global; }
    void tag_gpu_shared() { loc_ = mem_loc::gpu_shared; }
    void tag_gpu_local() { loc_ = mem_loc::gpu_local; }
    void tag_gpu_register() { loc_ = mem_loc::gpu_register; }
    /** Reference: __constant__ memory filled with cudaMemcpyToSymbol (src/tiramisu_codegen_cuda.cpp:817-823).  Here a
      * small read-only table lives in global memory and reaches the kernel as `const T * __restrict__` (LDG.NC, served by
      * L1 / the uniform datapath after the first touch): allocation, copies and kernel arguments are those of a
      * gpu_global buffer. */
    void tag_gpu_constant() { loc_ = mem_loc::gpu_global; }

    /** Allocate / free this buffer just before / after \p C at loop level \p level (reference core.h:1700-1712). */
    computation *allocate_at(computation &C, var level);
    computation *allocate_at(computation &C, int level);
    computation *deallocate_at(computation &C, var level);
    void dump(bool) const;
    void substitute_dims(const std::vector<std::string> &names, const std::vector<expr> &values);
    void set_dim_size(int dim, const expr &size) { dims_.at((size_t) dim) = size; }
    void replace_dim(const expr &from, const expr &to)
    {
        for (expr &d : dims_) if (d.is_equal(from)) d = to;
    }

private:
    std::string name_;
    std::vector<expr> dims_;
    primitive_t dtype_ = p_none;
    argument_t argt_ = a_temporary;
    mem_loc loc_ = mem_loc::host;
    bool auto_allocate_ = true;
    function *fct_ = nullptr;
};

/** One dimension of a computation's scheduled loop nest. */
struct loop_dim {
    std::string name;
    expr lo, hi;                 // half-open [lo, hi)
    enum tag_t { none, gpu_block, gpu_thread, parallel, vector, unrolled, distributed } tag = none;
    int vec_len = 0;
    bool warp_reduce = false;    // lowering: reduction loop strided over the 32 lanes of a warp, combined by shuffles
};

class computation {
public:
    static const var root;
    static const int root_dimension = -1;

    computation();
    computation(std::string name, std::vector<var> iterator_variables, expr e, bool schedule_this_computation, primitive_t t);
    computation(std::string name, std::vector<var> iterator_variables, expr e, bool schedule_this_computation);
    computation(std::string name, std::vector<var> iterator_variables, expr e);
    computation(std::vector<var> iterator_variables, expr e);
    computation(std::vector<var> iterator_variables, expr e, bool schedule_this_computation);
    computation(std::string name, std::vector<var> iterator_variables, primitive_t t);
    computation(std::vector<var> iterator_variables, primitive_t t);
    computation(std::string name, std::vector<var> iterator_variables, expr predicate, expr e);
    computation(std::vector<var> iterator_variables, expr predicate, expr e);
    /** Old-style declaration with an ISL set string (reference core.h:3236).  Only boxes are accepted:
      * `[params]->{S[i,j]: lo <= i <= hi and lo2 <= j < hi2}`; anything else stops the generator. */
    computation(std::string iteration_domain, expr e, bool schedule_this_computation, primitive_t t, function *fct);
    virtual ~computation() {}

    // ---- Layer I
    const std::string &get_name() const { return name_; }
    primitive_t get_data_type() const { return dtype_; }
    const expr &get_expr() const { return expr_; }
    void set_expression(const expr &e);
    void replace_expression(const expr &e) { expr_ = e; }
    bool should_schedule_this_computation() const { return scheduled_; }
    const std::vector<var> &get_iteration_variables() const { return iters_; }
    function *get_function() const { return fct_; }
    void add_predicate(expr p) { predicate_ = p; }
    const expr &get_predicate() const { return predicate_; }

    template <typename... Args> expr operator()(Args... args)
    {
        std::vector<expr> idx{expr(args)...};
        return access_expr(idx);
    }

    // ---- Layer II
    virtual void split(var L0, int sizeX);
    virtual void split(var L0, int sizeX, var L0_outer, var L0_inner);
    virtual void tile(var L0, var L1, int sizeX, int sizeY);
    virtual void tile(var L0, var L1, int sizeX, int sizeY, var L0_outer, var L1_outer, var L0_inner, var L1_inner);
    virtual void tile(var L0, var L1, var L2, int sizeX, int sizeY, int sizeZ);
    virtual void tile(var L0, var L1, var L2, int sizeX, int sizeY, int sizeZ,
                      var L0_outer, var L1_outer, var L2_outer, var L0_inner, var L1_inner, var L2_inner);
    virtual void interchange(var L0, var L1);
    virtual void gpu_tile(var L0, int sizeX);
    virtual void gpu_tile(var L0, int sizeX, var L0_outer, var L0_inner);
    virtual void gpu_tile(var L0, var L1, int sizeX, int sizeY);
    virtual void gpu_tile(var L0, var L1, int sizeX, int sizeY, var L0_outer, var L1_outer, var L0_inner, var L1_inner);
    virtual void gpu_tile(var L0, var L1, var L2, int sizeX, int sizeY, int sizeZ);
    virtual void gpu_tile(var L0, var L1, var L2, int sizeX, int sizeY, int sizeZ,
                          var L0_outer, var L1_outer, var L2_outer, var L0_inner, var L1_inner, var L2_inner);
    virtual void tag_gpu_level(var L0, var L1);
    virtual void tag_gpu_level(var L0, var L1, var L2, var L3);
    virtual void tag_gpu_level(var L0, var L1, var L2, var L3, var L4, var L5);
    virtual void tag_gpu_block_level(var L0);
    virtual void tag_gpu_block_level(var L0, var L1);
    virtual void tag_gpu_block_level(var L0, var L1, var L2);
    virtual void tag_gpu_thread_level(var L0);
    virtual void tag_gpu_thread_level(var L0, var L1);
    virtual void tag_gpu_thread_level(var L0, var L1, var L2);
    virtual void parallelize(var L);
    virtual void tag_parallel_level(var L);
    virtual void vectorize(var L, int v);
    virtual void vectorize(var L, int v, var L_outer, var L_inner);
    virtual void tag_vector_level(var L, int v);
    virtual void unroll(var L, int fac);
    virtual void unroll(var L, int fac, var L_outer, var L_inner);
    virtual void tag_unroll_level(var L);
    virtual void tag_unroll_level(var L, int) { tag_unroll_level(L); }
    virtual void shift(var L0, int n);
    // ---- the same commands addressed by loop LEVEL (position in the current loop order, outermost = 0) instead of by
    // name: the form the reference's older generators use (reference core.h:3420 after_low_level, :4583 tag_*_level(int),
    // :4717 tile(int, ...), :4560 split(int, int); benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_generator.cpp:132-262)
    void split(int L0, int sizeX);
    void tile(int L0, int L1, int sizeX, int sizeY);
    void tile(int L0, int L1, int L2, int sizeX, int sizeY, int sizeZ);
    void interchange(int L0, int L1);
    void tag_parallel_level(int L);
    void tag_vector_level(int L, int v);
    void tag_unroll_level(int L);
    void tag_unroll_level(int L, int) { tag_unroll_level(L); }
    void tag_gpu_level(int L0, int L1);
    void tag_gpu_level(int L0, int L1, int L2, int L3);
    /** reference src/tiramisu_core.cpp:2126: the ordering command of the older generators.  Every statement carries one
      * ordering number per loop depth (the "static dimensions" of its schedule, all 0 to begin with); this call copies
      * \p before_this 's numbers for depths 0..\p level and makes the number below level \p level that of
      * \p before_this plus 10.  The numbers are read when the call is made, a statement may be re-ordered any number of
      * times, and several statements may follow the same one.  The order of the program is the lexicographic order of
      * the number vectors; two neighbours share the loops above the first depth where their numbers differ
      * (function::resolve_static_order, run by the lowering). */
    void after_low_level(computation &before_this, int level);
    const std::vector<int> &static_order() const { return static_; }
    /** Number of loop levels around the statement (reference core.h:2370; the generators compare it before / after a
      * split to learn whether the split made a new level — here it always does). */
    int compute_maximal_AST_depth() const { return (int) dims_.size(); }
    /** Run loop `old_var` backwards; the loop is called `new_var` afterwards (reference core.h:4461). */
    virtual void loop_reversal(var old_var, var new_var);

    // ---- distribution (reference core.h:4583 tag_distribute_level, :4012 drop_rank_iter, :4817 gen_communication,
    // :4944 create_xfer; src/tiramisu_core.cpp:599-628, :7565, :8778-8815, :9174-9295).  One rank = one process = one GPU:
    // the iterations of a distributed loop are the ranks, every rank runs the one whose index equals
    // tiramisu_MPI_Comm_rank(0); buffers are per rank.
    virtual void tag_distribute_level(var L);
    virtual void tag_distribute_level(int L);
    /** The rank-local buffers are indexed without the distributed iterator: in everything this computation indexes with,
      * the distributed loop counts as 0. */
    void drop_rank_iter(var level);
    bool should_drop_rank_iter() const { return !drop_rank_.empty(); }
    /** Derive the exchange this computation's reads need from the other ranks (rows of a producer that is distributed the
      * same way, read at a positive offset along the distributed iterator), create the send / receive pair in front of it
      * and enlarge the producer's rank-local buffer by the received rows. */
    void gen_communication();
    void gen_communication(var) { gen_communication(); }
    /** Request slots of non-blocking transfers: completion is tracked by the runtime's message queue, the mapping is only
      * checked for syntax. */
    void set_wait_access(std::string access_str);
    static xfer create_xfer(std::string send_iter_domain, std::string recv_iter_domain, expr send_dest, expr recv_src,
                            xfer_prop send_prop, xfer_prop recv_prop, expr send_expr, function *fct);
    static xfer create_xfer(std::string iter_domain, xfer_prop prop, expr e, function *fct);
    virtual bool is_send() const { return false; }
    virtual bool is_recv() const { return false; }
    virtual bool is_send_recv() const { return false; }
    virtual bool is_wait() const { return false; }
    bool is_communicator() const { return is_send() || is_recv() || is_send_recv() || is_wait(); }
    /** Index of the distributed loop level, or -1. */
    int get_distributed_dimension() const;

    /** Order: \p next runs after this computation, sharing the loops up to and including \p level. */
    computation &then(computation &next, var level = root);
    computation &then(computation &next, int level);
    void after(computation &before_this, var level = root);
    void after(computation &before_this, int level);
    void before(computation &after_this, var level = root);
    void between(computation &before_c, var before_l, computation &after_c, var after_l);
    /** reference core.h:3889 `compute_at(consumer, L)`: on the CPU the producer is re-computed inside the consumer's
     *  loop L, region by region, for locality — the values every consumer iteration reads are those of the producer's
     *  own definition.  A data-parallel device has no use for the redundant recomputation: the producer is
     *  materialised once over its whole domain and ordered before the consumer at root (two kernels, or one when the
     *  stencil patterns fuse them); what is stored and read is unchanged. */
    void compute_at(computation &consumer, var L);
    void compute_at(computation &consumer, int L);
    /** Further definitions of the same computation over other domains ("updates", reference core.h:
      * add_definitions / get_update): each is a computation of its own that shares the name, so uses of the
      * name resolve to the element the latest definition is mapped to. */
    void add_definitions(std::string iteration_domain, expr e, bool schedule_this_computation, primitive_t t, function *fct);
    computation &get_update(int index) { return index == 0 ? *this : *updates_.at(index - 1); }
    computation &get_first_definition() { return *this; }
    computation &get_last_update() { return updates_.empty() ? *this : *updates_.back(); }
    computation *get_predecessor() const { return pred_; }
    computation *get_successor() const { return succ_; }
    /** An inline computation is not scheduled or stored: every use is replaced by its expression
      * (reference core.h: set_inline; done by the lowering before anything else looks at the program). */
    void set_inline(bool is_inline = true) { inline_ = is_inline; if (is_inline) scheduled_ = false; }
    bool is_inline_computation() const { return inline_; }
    /** Fuse with \p comp in the same loops down to \p level and run after it (reference core.h: fuse_after). */
    void fuse_after(var level, computation &comp) { after(comp, level); }
    /** Allocation scope of the computation's buffer.  Temporaries are allocated once per call here
      * (pooled), which is what any scope is equivalent to functionally: recorded as a no-op. */
    void store_at(computation &, var) {}

    /** Performance hint in the reference (stage \p inp through shared memory below \p level);
      * the sm_100a lowering picks its own staging, so this only records the hint. */
    computation *cache_shared(computation &inp, const var &level, std::vector<int> buffer_shape,
                              std::vector<expr> copy_offsets, bool pad_buffer = false);

    // ---- Layer III
    void store_in(buffer *buff);
    void store_in(buffer *buff, std::vector<expr> iterators);
    void store_in(std::vector<expr> mapping, std::vector<expr> sizes);
    buffer *get_buffer();
    buffer *get_automatically_allocated_buffer() { return get_buffer(); }
    void allocate_and_map_buffer_automatically(argument_t type = a_temporary) { get_buffer()->set_argument_type(type); }
    void set_access_exprs(buffer *b, std::vector<expr> idx);
    /** Old-style data mapping, `{S[i,j]->buf[i+1,j]}` (reference core.h:4180). */
    void set_access(std::string access_str);

    // ---- used by the lowering
    std::vector<loop_dim> &dims() { return dims_; }
    const std::vector<std::pair<std::string, expr>> &defs() const { return defs_; }
    const std::vector<expr> &guards() const { return guards_; }
    const std::vector<expr> &access() const { return access_; }
    bool has_access() const { return buf_ != nullptr; }
    computation *pred() const { return pred_; }
    computation *succ() const { return succ_; }
    const std::string &pred_level() const { return pred_level_; }
    int pred_level_index() const { return pred_level_index_; }
    /** Lowering: take the statement out of the program (its work was folded into a neighbour); the successor follows the
      * predecessor, sharing loops 0..\p succ_level_index with it. */
    void remove_from_order(int succ_level_index)
    {
        if (pred_) pred_->succ_ = succ_;
        if (succ_) { succ_->pred_ = pred_; succ_->pred_level_.clear(); succ_->pred_level_index_ = pred_ ? succ_level_index : -1; }
        pred_ = succ_ = nullptr;
        scheduled_ = false;
    }
    /** Lowering: replace the statement's iterators (Layer I) by \p new_iters — every use of an old iterator in the
      * expression, the predicate and the store index becomes \p values[i] for \p names[i] — and give it the plain loop
      * nest over the new iterators (whatever schedule it had described loops that no longer exist). */
    void rewrite_iterators(const std::vector<var> &new_iters, const std::vector<std::string> &names, const std::vector<expr> &values,
                           const std::vector<expr> &new_access, const std::vector<expr> &new_guards)
    {
        iters_ = new_iters;
        if (expr_.is_defined()) expr_ = expr_.substitute(names, values);
        if (predicate_.is_defined()) predicate_ = predicate_.substitute(names, values);
        access_ = new_access;
        dims_.clear(); defs_.clear(); guards_ = new_guards;
        for (const var &v : iters_) { loop_dim ld; ld.name = v.get_name(); ld.lo = v.get_lower(); ld.hi = v.get_upper(); dims_.push_back(ld); }
    }
    /** Lowering: share loops 0..level_index with the predecessor in the order (nest fusion). */
    void set_pred_level_index(int level_index) { pred_level_index_ = level_index; }
    bool is_input() const { return is_input_; }
    bool is_constant() const { return is_constant_; }
    /** A constant declared at a loop level (not function wide): a let, substituted into the statements that name it. */
    bool is_let() const { return is_let_; }
    /** Declared without iterator bounds (`{S[i,j]}`): function::compute_bounds() has to supply them. */
    bool pending_bounds() const { return pending_bounds_; }
    void set_iterator_ranges(const std::vector<std::pair<expr, expr>> &ranges);
    int find_dim(const std::string &name) const;
    void dump() const;
    /** Replace named scalars by values in every expression the computation holds (constant propagation). */
    void substitute_everywhere(const std::vector<std::string> &names, const std::vector<expr> &values);
    /** True when every guard only cuts the excess of a non-dividing split (the original box is still covered). */
    bool guards_only_trim() const { return true; }

protected:
    void construct(std::string name, std::vector<var> iters, expr e, bool schedule, primitive_t t, function *fct = nullptr);
    expr access_expr(const std::vector<expr> &idx);
    void tag(const var &L, loop_dim::tag_t t);
    void link_after(computation &before_this, const std::string &level_name, int level_index);

    std::string name_;
    std::vector<var> iters_;
    expr expr_, predicate_;
    primitive_t dtype_ = p_none;
    bool scheduled_ = true;
    bool is_input_ = false, is_constant_ = false, inline_ = false, is_let_ = false, schedule_when_defined_ = false, pending_bounds_ = false;
    function *fct_ = nullptr;

    std::vector<loop_dim> dims_;
    std::vector<std::pair<std::string, expr>> defs_;   // name := expression in newer names (oldest first)
    std::vector<expr> guards_;

    buffer *buf_ = nullptr;
    std::vector<expr> access_;                           // index expressions in the original iterators
    std::vector<computation *> updates_;
    bool own_access_ = false;                            // an update that was mapped explicitly (not through its first definition)
    std::string drop_rank_;                              // distributed level whose iterator the rank-local indices ignore
    std::vector<int> static_;                            // ordering numbers per loop depth (after_low_level); empty = all 0
    computation *let_host_ = nullptr;                    // a let: the statement it was declared inside (ordering it orders that one)
    computation *pred_ = nullptr, *succ_ = nullptr;
    std::string pred_level_;                             // "" = root
    int pred_level_index_ = -2;                          // -2 = by name, -1 = root, >=0 explicit depth
    friend class function;
    friend class buffer;
    friend class constant;
};

class input : public computation {
public:
    input(std::string name, std::vector<var> iterator_variables, primitive_t t);
    input(std::vector<var> iterator_variables, primitive_t t);
    input(std::string name, std::vector<std::string> dimension_names, std::vector<expr> dimension_sizes, primitive_t t);
};

/** A named window onto a buffer: declared like an input, mapped with store_in (reference core.h:5159). */
class view : public input {
public:
    view(std::string name, std::vector<var> iterator_variables, primitive_t t) : input(name, iterator_variables, t) {}
    view(std::string name, std::vector<var> iterator_variables, buffer *b) : input(name, iterator_variables, b->get_elements_type()) { store_in(b); }
};

class Input : public input {
public:
    Input(std::string name, std::vector<expr> sizes, primitive_t t);
};

/** A scalar evaluated once at function entry (reference core.h:5212; only the function-wide form). */
class constant : public computation {
public:
    constant(std::string param_name, const expr &param_expr, primitive_t t, bool function_wide,
             computation *with_computation, int at_loop_level, function *func = global::get_implicit_function());
    constant(std::string param_name, const expr &param_expr, primitive_t t, function *func = global::get_implicit_function());
    constant(std::string param_name, const expr &param_expr);
    operator expr();
};

/** A computation that is a message (reference core.h:5610-5647).  Its loops stay on the host; every iteration is one
  * call of the tiramisu_MPI_* runtime on the (device-resident) element its expression / access names. */
class communicator : public computation {
public:
    communicator(std::string iteration_domain_str, expr rhs, bool schedule_this_computation, primitive_t data_type, function *fct)
        : computation(iteration_domain_str, rhs, schedule_this_computation, data_type, fct) {}
    communicator(std::string iteration_domain_str, expr rhs, bool schedule_this_computation, primitive_t data_type, xfer_prop prop,
                 function *fct)
        : computation(iteration_domain_str, rhs, schedule_this_computation, data_type, fct), prop(prop) {}
    xfer_prop get_xfer_props() const { return prop; }
    /** Elements per message: the product of what collapse() folded into it. */
    expr get_num_elements() const;
    void add_dim(expr size) { msg_dims_.push_back(size); }
    /** One message for the whole of loop `level` (collapse_until_iter must be -1, as in the reference). */
    std::vector<communicator *> collapse(int level, expr collapse_from_iter, expr collapse_until_iter, expr num_collapsed);
    void collapse_many(std::vector<collapse_group> collapse_each);
    /** Lowering: levels folded into the message by the lowering's own coalescing (same on both sides of a transfer). */
    void fold_level(int level, const expr &count);
    bool explicitly_collapsed() const { return !msg_dims_.empty(); }

protected:
    communicator() {}
    xfer_prop prop;
    std::vector<expr> msg_dims_;
};

class send : public communicator {
public:
    send(std::string iteration_domain_str, computation *producer, expr rhs, xfer_prop prop, bool schedule_this_computation,
         std::vector<expr> dims, function *fct);
    bool is_send() const override { return true; }
    computation *get_producer() const { return producer_; }
    recv *get_matching_recv() const { return matching_recv_; }
    expr get_msg_tag() const { return msg_tag_; }
    expr get_src() const { return src_; }
    expr get_dest() const { return dest_; }
    void set_matching_recv(recv *r) { matching_recv_ = r; }
    void set_src(expr e) { src_ = e; }
    void set_dest(expr e) { dest_ = e; }
    void override_msg_tag(expr t) { msg_tag_ = t; }

private:
    computation *producer_ = nullptr;
    recv *matching_recv_ = nullptr;
    expr msg_tag_, src_, dest_;
    static int next_msg_tag;
};

class recv : public communicator {
public:
    recv(std::string iteration_domain_str, bool schedule_this, xfer_prop prop, function *fct);
    bool is_recv() const override { return true; }
    send *get_matching_send() const { return matching_send_; }
    expr get_msg_tag() const { return msg_tag_; }
    expr get_src() const { return src_; }
    expr get_dest() const { return dest_; }
    void set_matching_send(send *s) { matching_send_ = s; }
    void set_src(expr e) { src_ = e; }
    void set_dest(expr e) { dest_ = e; }
    void override_msg_tag(expr t) { msg_tag_ = t; }

private:
    send *matching_send_ = nullptr;
    expr msg_tag_, src_, dest_;
};

/** One-sided / memcpy transfers (reference core.h:5735): declared for source compatibility, not lowered. */
class send_recv : public communicator {
public:
    send_recv(std::string iteration_domain_str, computation *producer, expr rhs, xfer_prop prop, bool schedule_this_computation,
              std::vector<expr> dims, function *fct);
    bool is_send_recv() const override { return true; }
};

class wait : public communicator {
public:
    wait(expr rhs, xfer_prop prop, function *fct);
    wait(std::string iteration_domain_str, expr rhs, xfer_prop prop, bool schedule_this, function *fct);
    bool is_wait() const override { return true; }
    std::vector<computation *> get_op_to_wait_on() const;
};

class function {
public:
    explicit function(std::string name);
    const std::string &get_name() const { return name_; }
    void add_computation(computation *c) { comps_.push_back(c); }
    void add_buffer(buffer *b) { bufs_.push_back(b); }
    const std::vector<computation *> &get_computations() const { return comps_; }
    const std::vector<buffer *> &get_buffers() const { return bufs_; }
    computation *find_computation(const std::string &name) const;
    buffer *find_buffer(const std::string &name) const;
    /** The function constant called `name`, or nullptr (reference core.h:1198). */
    constant *get_invariant_by_name(std::string name) const;
    std::string fresh_name(const std::string &prefix);
    /** Heap objects the front-end creates on the user's behalf (auto buffers, allocate_at computations). */
    template <typename T> T *own(T *p)
    {
        owned_.push_back(std::shared_ptr<void>(p, [](void *q) { delete static_cast<T *>(q); }));
        return p;
    }
    void set_arguments(const std::vector<buffer *> &args) { args_ = args; }
    const std::vector<buffer *> &get_arguments() const { return args_; }
    void codegen(const std::vector<buffer *> &arguments, const std::string obj_filename, bool gen_cuda_stmt = false);
    void dump(bool) const;
    // Old-style driver sequence (reference core.h:289-1166).  The intermediate stages have no
    // counterpart here (no ISL AST, no Halide statement): they only record what was asked for and
    // gen_halide_obj() runs the whole lowering.  A function without gpu levels is a CPU schedule; this
    // back end has no CPU code generator, so such a function is re-targeted (lowering_retarget.cpp).
    void add_context_constraints(const std::string &) {}
    /** Reference core.h:1080: pads every schedule to the same number of dimensions; loop lists need no padding. */
    void align_schedules() {}
    /** Turn the ordering numbers of after_low_level into the then()/after() chain the lowering walks (no-op when the
      * function never called after_low_level; mixing it with then()/after() is refused). */
    void resolve_static_order();
    void note_static_order() { static_order_ = true; }
    /** Reference core.h:865: turns communicators into library calls; the lowering does that on its own. */
    void lift_dist_comps() {}
    void gen_time_space_domain() {}
    void gen_isl_ast() {}
    void gen_halide_stmt() {}
    void gen_cuda_stmt() { cuda_requested_ = true; }
    void dump_halide_stmt() const {}
    void gen_halide_obj(const std::string &obj_file_name, bool gen_python = false);
    void gen_c_code() const {}
    /** Buffers are mapped lazily (computation::get_buffer): nothing to do ahead of time. */
    void allocate_and_map_buffers_automatically() {}
    /** Ranges for iterators declared without bounds, from what the consumers read (reference core.h: compute_bounds). */
    void compute_bounds();
    void dump_dep_graph() {}
    void dump_halide_IR() const {}
    void dump_time_processor_domain() const {}
    void dump_trimmed_time_processor_domain() const {}
    void dump_iteration_domain() const {}
    void dump_schedule() const {}
    void dump_sched_graph() {}

private:
    std::string name_;
    std::vector<computation *> comps_;
    std::vector<buffer *> bufs_;
    std::vector<buffer *> args_;
    std::vector<std::shared_ptr<void>> owned_;
    int counter_ = 0;
    bool cuda_requested_ = false;
    bool static_order_ = false;
};

}  // namespace tiramisu

#endif
