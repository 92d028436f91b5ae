// distributed.cpp — the distribution half of the generator API: tag_distribute_level / drop_rank_iter, transfers
// (xfer_prop, send, recv, wait, create_xfer, collapse_many) and gen_communication.
//
// Reference: include/tiramisu/core.h:88-110 (xfer_attr, xfer, collapse_group), :4583 tag_distribute_level, :4012
// drop_rank_iter, :4944 create_xfer, :5575-5760 (xfer_prop, communicator, send, recv, send_recv, wait);
// src/tiramisu_core.cpp:599-628, :7565, :8355-8400 (collapse), :8473-8760 (constructors), :8778-8815 (create_xfer),
// :9174-9295 (gen_communication).  There the exchange sets are ISL sets and the messages become Halide extern calls
// (src/tiramisu_function.cpp: lift_mpi_comp); here a communicator stays a box of host loops whose body is one
// tiramisu_MPI_* call (lowering_dist.cpp) and the halo a consumer needs is read off its affine accesses.
#include <tiramisu/core.h>

#include <cstdlib>
#include <iostream>

namespace tiramisu {

namespace {
[[noreturn]] void die(const std::string &msg)
{
    std::cerr << "tiramisu_b200 front-end: " << msg << std::endl;
    std::exit(1);
}
expr lit(int64_t v) { return expr((int32_t) v); }
}  // namespace

// ------------------------------------------------------------------------------------------------ distribution tags
void computation::tag_distribute_level(var L) { tag(L, loop_dim::distributed); }

void computation::tag_distribute_level(int L)
{
    if (L < 0 || L >= (int) dims_.size()) die("tag_distribute_level: '" + name_ + "' has no loop level " + std::to_string(L));
    dims_[(size_t) L].tag = loop_dim::distributed;
}

int computation::get_distributed_dimension() const
{
    for (size_t i = 0; i < dims_.size(); i++)
        if (dims_[i].tag == loop_dim::distributed) return (int) i;
    return -1;
}

void computation::drop_rank_iter(var level)
{
    if (find_dim(level.get_name()) < 0) die("drop_rank_iter: '" + name_ + "' has no loop level '" + level.get_name() + "'");
    drop_rank_ = level.get_name();
    // every original iterator rebuilt from this level is rebuilt as if the level were 0: rank-local indices
    const std::vector<std::string> names{drop_rank_};
    const std::vector<expr> zero{lit(0)};
    for (auto &d : defs_) d.second = d.second.substitute(names, zero);
    for (auto &a : access_) a = a.substitute(names, zero);
}

void computation::set_wait_access(std::string access_str)
{
    if (access_str.find("->") == std::string::npos) die("set_wait_access of '" + name_ + "': not an access map: " + access_str);
}

// ------------------------------------------------------------------------------------------------ communicators
int send::next_msg_tag = 0;

expr communicator::get_num_elements() const
{
    expr n = lit(1);
    for (const expr &d : msg_dims_) n = d * n;
    return n;
}

void communicator::fold_level(int level, const expr &count)
{
    if (level < 0 || level >= (int) dims_.size()) die("collapse of '" + name_ + "': no loop level " + std::to_string(level));
    loop_dim &d = dims_[(size_t) level];
    d.hi = d.lo + lit(1);                               // the loop keeps its first iteration: the message starts there
    msg_dims_.push_back(count);
}

std::vector<communicator *> communicator::collapse(int level, expr collapse_from_iter, expr collapse_until_iter, expr num_collapsed)
{
    if (!(collapse_until_iter.is_int_literal() && collapse_until_iter.get_int_val() == -1))
        die("collapse of '" + name_ + "': only whole loops can be collapsed (collapse_until_iter == -1), as in the reference");
    if (level < 0 || level >= (int) dims_.size()) die("collapse of '" + name_ + "': no loop level " + std::to_string(level));
    dims_[(size_t) level].lo = collapse_from_iter;
    fold_level(level, num_collapsed);
    return {};
}

void communicator::collapse_many(std::vector<collapse_group> collapse_each)
{
    for (auto &c : collapse_each) collapse(std::get<0>(c), std::get<1>(c), std::get<2>(c), std::get<3>(c));
}

send::send(std::string iteration_domain_str, computation *producer, expr rhs, xfer_prop p, bool schedule_this_computation,
           std::vector<expr>, function *fct)
    : communicator(iteration_domain_str, rhs, schedule_this_computation, p.get_dtype(), p, fct), producer_(producer),
      msg_tag_(lit(next_msg_tag++))
{
    if (rhs.get_expr_type() != e_op || rhs.get_op_type() != o_access) die("send '" + name_ + "': the payload must be an access");
}

recv::recv(std::string iteration_domain_str, bool schedule_this, xfer_prop p, function *fct)
    : communicator(iteration_domain_str, expr(), schedule_this, p.get_dtype(), p, fct)
{
    // a receive has no right-hand side (the message is its value); it is a scheduled statement all the same
    expr_ = cast(p.get_dtype(), lit(0));
    scheduled_ = schedule_this;
    is_input_ = false;
    schedule_when_defined_ = false;
}

send_recv::send_recv(std::string iteration_domain_str, computation *, expr rhs, xfer_prop p, bool schedule_this_computation,
                     std::vector<expr>, function *fct)
    : communicator(iteration_domain_str, rhs, schedule_this_computation, p.get_dtype(), p, fct)
{
}

wait::wait(expr rhs, xfer_prop p, function *fct)
{
    if (rhs.get_expr_type() != e_op || rhs.get_op_type() != o_access) die("wait: the operand must be an access to a send or a receive");
    function *f = fct ? fct : global::get_implicit_function();
    computation *op = f ? f->find_computation(rhs.get_name()) : nullptr;
    if (!op) die("wait on unknown operation '" + rhs.get_name() + "'");
    construct(op->get_name() + "_wait", op->get_iteration_variables(), rhs, true, p_async, f);
    prop = p;
}

wait::wait(std::string iteration_domain_str, expr rhs, xfer_prop p, bool schedule_this, function *fct)
    : communicator(iteration_domain_str, rhs, schedule_this, p_async, p, fct)
{
}

std::vector<computation *> wait::get_op_to_wait_on() const
{
    computation *op = fct_->find_computation(expr_.get_name());
    return op ? std::vector<computation *>{op} : std::vector<computation *>{};
}

xfer computation::create_xfer(std::string send_iter_domain, std::string recv_iter_domain, expr send_dest, expr recv_src,
                              xfer_prop send_prop, xfer_prop recv_prop, expr send_expr, function *fct)
{
    if (send_prop.contains_attr(MPI) != recv_prop.contains_attr(MPI)) die("create_xfer: both ends must use the same transport");
    if (!send_prop.contains_attr(MPI)) die("create_xfer: only MPI transfers are lowered (they run over NCCL / NVLink, tiramisu_MPI_*)");
    if (send_expr.get_expr_type() != e_op || send_expr.get_op_type() != o_access) die("create_xfer: the payload must be an access");
    function *f = fct ? fct : global::get_implicit_function();
    computation *producer = f->find_computation(send_expr.get_name());
    if (!producer) die("create_xfer: unknown producer '" + send_expr.get_name() + "'");
    xfer c;
    c.s = f->own(new send(send_iter_domain, producer, send_expr, send_prop, true, {lit(1)}, f));
    c.r = f->own(new recv(recv_iter_domain, true, recv_prop, f));
    c.s->set_dest(send_dest);
    c.r->set_src(recv_src);
    c.s->set_matching_recv(c.r);
    c.r->set_matching_send(c.s);
    return c;
}

xfer computation::create_xfer(std::string iter_domain, xfer_prop prop, expr e, function *fct)
{
    function *f = fct ? fct : global::get_implicit_function();
    computation *producer = f->find_computation(e.get_name());
    xfer c;
    c.sr = f->own(new send_recv(iter_domain, producer, e, prop, true, {lit(1)}, f));
    return c;
}

// ------------------------------------------------------------------------------------------------ gen_communication
// The consumer `this` is distributed over loop level D, the outer part of a split of its original iterator `it` by F
// (F rows per rank), and indexes rank-locally (drop_rank_iter).  Every producer it reads at it + k with k > 0, stored
// in a rank-local buffer whose first dimension follows `it`, needs its first kmax rows from the next rank: they are
// received behind the rank's own F rows (buffer enlarged by kmax rows, as the reference's gen_communication_code does)
// and sent from the first kmax local rows to the previous rank.  Reads at negative offsets stop the generator (the
// reference handles neither).
// function constants with literal values, substituted everywhere (what the lowering does first as well): the exchange
// is derived from literal extents
static void propagate_literal_constants(function *f)
{
    std::vector<std::string> names;
    std::vector<expr> values;
    for (computation *c : f->get_computations()) {
        if (!c->is_constant() || c->is_let()) continue;
        expr v = c->get_expr().substitute(names, values);
        if (v.is_literal()) { names.push_back(c->get_name()); values.push_back(cast(c->get_data_type(), v)); }
    }
    if (names.empty()) return;
    for (computation *c : f->get_computations()) c->substitute_everywhere(names, values);
    for (buffer *b : f->get_buffers()) b->substitute_dims(names, values);
}

void computation::gen_communication()
{
    propagate_literal_constants(fct_);
    const int D = get_distributed_dimension();
    if (D < 0) die("gen_communication: '" + name_ + "' has no distributed level");
    const std::string dname = dims_[(size_t) D].name;
    // the original iterator the distributed level was split from, and the rows per rank
    std::string it;
    int64_t F = -1;
    if (D + 1 < (int) dims_.size()) {
        // the split that made the distributed level defined `it := lo + (level * F + inner)` (the level itself reads 0 after
        // drop_rank_iter): the definition that mentions the inner level
        const std::string inner = dims_[(size_t) D + 1].name;
        for (auto &d : defs_)
            if (!d.second.substitute({inner}, {lit(12345)}).is_equal(d.second)) { it = d.first; break; }
    }
    const expr ranks = dims_[(size_t) D].hi - dims_[(size_t) D].lo;
    if (!ranks.is_int_literal() && !dims_[(size_t) D].hi.is_int_literal()) die("gen_communication: the number of ranks of '" + name_ + "' is not a literal");
    const int64_t R = dims_[(size_t) D].hi.get_int_val() - (dims_[(size_t) D].lo.is_int_literal() ? dims_[(size_t) D].lo.get_int_val() : 0);
    if (D + 1 >= (int) dims_.size() || !dims_[(size_t) D + 1].hi.is_int_literal()) die("gen_communication: '" + name_ + "' — the distributed level must be the outer part of a split");
    F = dims_[(size_t) D + 1].hi.get_int_val();
    if (it.empty()) die("gen_communication: cannot find the iterator '" + dname + "' was split from");

    struct Need { computation *p; int64_t kmax; };
    std::vector<Need> needs;
    std::vector<expr> stack{expr_};
    while (!stack.empty()) {
        expr x = stack.back(); stack.pop_back();
        if (!x.is_defined() || x.get_expr_type() != e_op) continue;
        for (const expr &o : x.get_operands()) stack.push_back(o);
        if (x.get_op_type() != o_access || x.get_n_arg() == 0) continue;
        computation *p = fct_->find_computation(x.get_name());
        if (!p || p->is_constant() || p->get_distributed_dimension() < 0 || !p->should_drop_rank_iter()) continue;
        // offset of the first index along `it`
        expr off = x.get_operand(0) - expr::make_var(global::get_loop_iterator_data_type(), it);
        expr probe0 = x.get_operand(0).substitute({it}, {lit(0)}), probe1 = x.get_operand(0).substitute({it}, {lit(1)});
        if (!probe0.is_int_literal() || !probe1.is_int_literal() || probe1.get_int_val() - probe0.get_int_val() != 1) continue;
        const int64_t k = probe0.get_int_val();
        if (k < 0) die("gen_communication: '" + name_ + "' reads '" + p->get_name() + "' at a negative offset along the distributed iterator");
        if (k == 0) continue;
        bool found = false;
        for (Need &n : needs) if (n.p == p) { n.kmax = std::max(n.kmax, k); found = true; }
        if (!found) needs.push_back({p, k});
    }
    int id = 0;
    for (const Need &n : needs) {
        buffer *b = n.p->get_buffer();
        const auto &pit = n.p->get_iteration_variables();
        if (pit.empty() || (int) n.p->access().size() != b->get_n_dims()) die("gen_communication: producer '" + n.p->get_name() + "' has no usable mapping");
        // iteration boxes: [rank q][row i' < kmax][the producer's other iterators]
        std::string s_dom = "{" + name_ + "_halo_s" + std::to_string(id) + "[q,hr", r_dom = "{" + name_ + "_halo_r" + std::to_string(id) + "[q,hr";
        std::string rest_names, rest_cons;
        std::vector<expr> payload_idx{expr::make_var(global::get_loop_iterator_data_type(), "hr")};
        for (size_t i = 1; i < pit.size(); i++) {
            if (!pit[i].get_lower().is_int_literal() || !pit[i].get_upper().is_int_literal())
                die("gen_communication: iterator '" + pit[i].get_name() + "' of '" + n.p->get_name() + "' has no literal range (function constants must be propagated by value)");
            const std::string nm = "h" + std::to_string(i);
            rest_names += "," + nm;
            rest_cons += " and " + std::to_string(pit[i].get_lower().get_int_val()) + "<=" + nm + "<" + std::to_string(pit[i].get_upper().get_int_val());
            payload_idx.push_back(expr::make_var(global::get_loop_iterator_data_type(), nm));
        }
        s_dom += rest_names + "]: 1<=q<" + std::to_string(R) + " and 0<=hr<" + std::to_string(n.kmax) + rest_cons + "}";
        r_dom += rest_names + "]: 0<=q<" + std::to_string(R - 1) + " and 0<=hr<" + std::to_string(n.kmax) + rest_cons + "}";
        const expr q = expr::make_var(global::get_loop_iterator_data_type(), "q");
        xfer t = create_xfer(s_dom, r_dom, q - 1, q + 1, xfer_prop(n.p->get_data_type(), {MPI, BLOCK, ASYNC}),
                             xfer_prop(n.p->get_data_type(), {MPI, BLOCK, ASYNC}), expr(o_access, n.p->get_name(), payload_idx, n.p->get_data_type()), fct_);
        t.s->tag_distribute_level(var("q"));
        t.r->tag_distribute_level(var("q"));
        // the payload is read through the producer's own (rank-local) mapping; the received rows land behind the F own rows
        std::vector<expr> ridx;
        {
            std::vector<std::string> pn;
            for (const var &v : pit) pn.push_back(v.get_name());
            std::vector<expr> at = payload_idx;
            at[0] = at[0] + lit(F);
            for (const expr &a : n.p->access()) ridx.push_back(a.substitute(pn, at));
        }
        t.r->set_access_exprs(b, ridx);
        if (!b->get_dim_sizes()[0].is_int_literal()) die("gen_communication: the first extent of '" + b->get_name() + "' is not a literal");
        if (b->get_dim_sizes()[0].get_int_val() < F + n.kmax) b->set_dim_size(0, lit(F + n.kmax));
        // order: ... -> send -> recv -> this
        if (pred_) {
            computation *before = pred_;
            const std::string lvl = pred_level_;
            const int lvl_i = pred_level_index_;
            pred_->succ_ = nullptr; pred_ = nullptr;
            t.s->link_after(*before, lvl, lvl_i);
            t.r->after(*t.s, lvl_i >= 0 ? lvl_i : -1);
            this->after(*t.r, lvl_i >= 0 ? lvl_i : -1);
            if (lvl_i < 0 && !lvl.empty()) { /* named levels do not exist in the communicators: ordered at root */ }
        } else {
            t.s->before(*t.r, computation::root);
            t.r->before(*this, computation::root);
        }
        id++;
    }
}

}  // namespace tiramisu
