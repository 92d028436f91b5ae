// lowering_dist.cpp — distributed functions: one rank = one process = one GPU.
//
// Reference: a distributed loop becomes `if (rank == i)` around its body with `rank = tiramisu_MPI_Comm_rank(offset)`
// bound at function entry (src/tiramisu_codegen_halide.cpp:2554-2562), send / recv / wait computations become extern
// calls of tiramisu_MPI_* with (count, peer, tag, address[, request]) (src/tiramisu_function.cpp: lift_mpi_comp,
// src/tiramisu_core.cpp:8400-8470 names).  Here:
//   * the distributed level is a host-side scope `v = rank; if (lo <= v < hi)`; what is below it is lowered as usual
//     (kernels over the rank-local device mirrors of the buffers);
//   * a communicator keeps its loops on the host and each iteration issues ONE tiramisu_MPI_* call on the DEVICE
//     address of the element — the runtime moves device buffers over NCCL / NVLink (csrc/comm.cu), nothing is staged
//     through the host;
//   * messages are coalesced: an inner loop whose consecutive iterations address consecutive elements on BOTH sides
//     of a transfer is folded into the message (count *= extent), the same way on the send and on the receive, so the
//     reference's element-per-message loops (tests/test_99.cpp: 10 000 one-int messages per rank) become a few large
//     transfers.  collapse_many() folds by hand and switches the automatic folding off for that transfer;
//   * wait() is one call per rank: the runtime's Wait flushes every queued message.
#include "lowering.h"

namespace tiramisu {
namespace lowering_detail {

static expr rank_var() { return expr::make_var(p_int32, "rank"); }

// coefficient (in elements) of every own loop level of `c` in the linear index of `b[idx]`, innermost last;
// false if the address is not affine in them or an extent is not a literal
static bool linear_coefficients(computation *c, buffer *b, std::vector<expr> idx, std::vector<int64_t> &coef, std::vector<int64_t> &extent)
{
    // rebuild the original iterators from the scheduled levels
    const auto &defs = c->defs();
    for (auto it = defs.rbegin(); it != defs.rend(); ++it)
        for (expr &e : idx) e = e.substitute({it->first}, {it->second});
    std::vector<int64_t> pitch(idx.size(), 1);
    for (int d = (int) idx.size() - 2; d >= 0; d--) {
        const expr &e = b->get_dim_sizes()[(size_t) d + 1];
        if (!e.is_int_literal()) return false;
        pitch[(size_t) d] = pitch[(size_t) d + 1] * e.get_int_val();
    }
    coef.assign(c->dims().size(), 0);
    extent.assign(c->dims().size(), -1);
    for (size_t d = 0; d < idx.size(); d++) {
        std::map<std::string, int64_t> cf;
        int64_t cst = 0;
        if (!affine_form(idx[d], cf, cst)) return false;
        for (auto &kv : cf) {
            int l = c->find_dim(kv.first);
            if (l < 0) continue;                                     // a function constant: the same on every iteration
            coef[(size_t) l] += kv.second * pitch[d];
        }
    }
    for (size_t l = 0; l < c->dims().size(); l++) {
        expr ext = c->dims()[l].hi - c->dims()[l].lo;
        if (ext.is_int_literal()) extent[l] = ext.get_int_val();
        else if (c->dims()[l].hi.is_int_literal() && c->dims()[l].lo.is_int_literal()) extent[l] = c->dims()[l].hi.get_int_val() - c->dims()[l].lo.get_int_val();
    }
    return true;
}

bool Lowering::prepare_distribution(const std::vector<computation *> &ord)
{
    bool any = false;
    for (computation *c : f_->get_computations()) any |= c->get_distributed_dimension() >= 0 || c->is_communicator();
    if (!any) return false;
    distributed_ = true;
    if (!f_->find_computation("rank"))
        f_->own(new constant("rank", expr(o_call, "tiramisu_MPI_Comm_rank", {expr((int32_t) 0)}, p_int32), p_int32, true, nullptr, 0, f_));
    for (computation *c : ord) {
        if (c->is_send_recv()) die("'" + c->get_name() + "': one-sided / memcpy transfers (create_xfer with one domain) are not lowered");
        if (c->is_wait()) {
            for (auto &d : c->dims())
                if (d.tag != loop_dim::distributed) d.hi = d.lo + expr((int32_t) 1);
        }
    }
    // joint coalescing of every send / recv pair
    for (computation *c : ord) {
        if (!c->is_send()) continue;
        send *s = static_cast<send *>(c);
        recv *r = s->get_matching_recv();
        if (!r || s->explicitly_collapsed() || r->explicitly_collapsed()) continue;
        if (!r->has_access()) die("receive '" + r->get_name() + "' has no access (set_access)");
        computation *prod = f_->find_computation(s->get_expr().get_name());
        if (!prod) die("send '" + s->get_name() + "': unknown payload '" + s->get_expr().get_name() + "'");
        std::vector<std::string> pn;
        for (const var &v : prod->get_iteration_variables()) pn.push_back(v.get_name());
        if ((int) pn.size() != s->get_expr().get_n_arg()) die("send '" + s->get_name() + "': payload arity");
        std::vector<expr> sidx;
        for (const expr &a : prod->access()) sidx.push_back(a.substitute(pn, s->get_expr().get_operands()));
        std::vector<int64_t> sc, se, rc, re;
        if (!linear_coefficients(s, prod->get_buffer(), sidx, sc, se) ||
            !linear_coefficients(r, r->get_buffer(), std::vector<expr>(r->access().begin(), r->access().end()), rc, re))
            continue;
        int64_t count = 1;
        int ls = (int) s->dims().size() - 1, lr = (int) r->dims().size() - 1;
        while (ls >= 0 && lr >= 0) {
            if (s->dims()[(size_t) ls].tag == loop_dim::distributed || r->dims()[(size_t) lr].tag == loop_dim::distributed) break;
            if (se[(size_t) ls] <= 0 || se[(size_t) ls] != re[(size_t) lr]) break;
            if (se[(size_t) ls] > 1 && (sc[(size_t) ls] != count || rc[(size_t) lr] != count)) break;
            const int64_t ext = se[(size_t) ls];
            if (ext > 1) {
                s->fold_level(ls, expr((int32_t) ext));
                r->fold_level(lr, expr((int32_t) ext));
            }
            count *= ext;
            ls--; lr--;
        }
    }
    return true;
}

static const char *mpi_suffix(primitive_t t)
{
    switch (t) {
    case p_int8: return "int8"; case p_uint8: return "uint8"; case p_int16: return "int16"; case p_uint16: return "uint16";
    case p_int32: return "int32"; case p_uint32: return "uint32"; case p_int64: return "int64"; case p_uint64: return "uint64";
    case p_float32: return "f32"; case p_float64: return "f64";
    default: return nullptr;
    }
}

void Lowering::emit_communicator(Node *n, std::ostringstream &os, int ind)
{
    computation *c = n->comp;
    os << pad(ind) << "{   // " << c->get_name() << (c->is_send() ? " (send)" : c->is_recv() ? " (receive)" : " (wait)") << "\n";
    std::vector<std::string> conds;
    emit_bindings(n, os, ind + 1, conds, false);
    std::string cond;
    for (size_t i = 0; i < conds.size(); i++) cond += (i ? " && " : "") + conds[i];
    const std::string guard = cond.empty() ? "" : "if (" + cond + ") ";
    if (c->is_wait()) {
        os << pad(ind + 1) << guard << "tiramisu_MPI_Wait(nullptr);\n" << pad(ind) << "}\n";
        return;
    }
    communicator *cm = static_cast<communicator *>(c);
    const xfer_prop prop = cm->get_xfer_props();
    const char *sfx = mpi_suffix(prop.get_dtype());
    if (!sfx) die("'" + c->get_name() + "': no tiramisu_MPI_* transfer for this element type");
    const bool nonblock = prop.contains_attr(NONBLOCK);
    std::string fn = "tiramisu_MPI_", peer, tag, addr;
    const std::string T = ctype(prop.get_dtype());
    if (c->is_send()) {
        send *s = static_cast<send *>(c);
        const bool sync = prop.contains_attr(SYNC);
        fn += sync ? (nonblock ? "Issend_" : "Ssend_") : (nonblock ? "Isend_" : "Send_");
        if (!s->get_dest().is_defined()) die("send '" + c->get_name() + "' has no destination");
        peer = print(s->get_dest(), false);
        tag = print(s->get_msg_tag(), false);
        addr = "&" + print_access(s->get_expr(), true);
    } else {
        recv *r = static_cast<recv *>(c);
        fn += nonblock ? "Irecv_" : "Recv_";
        if (!r->get_src().is_defined()) die("receive '" + c->get_name() + "' has no source");
        if (!r->has_access()) die("receive '" + c->get_name() + "' has no access (set_access)");
        peer = print(r->get_src(), false);
        const expr t = r->get_msg_tag().is_defined() ? r->get_msg_tag() : (r->get_matching_send() ? r->get_matching_send()->get_msg_tag() : expr((int32_t) 0));
        tag = print(t, false);
        buffer *b = r->get_buffer();
        enter_stmt(c);
        addr = "&b_" + san(b->get_name()) + "[" + lin_index(b, r->access(), true) + "]";
        enter_stmt(nullptr);
    }
    os << pad(ind + 1) << guard << fn << sfx << "((int) (" << print(cm->get_num_elements(), false) << "), (int) (" << peer << "), (int) (" << tag
       << "), (" << T << " *) " << addr << (nonblock ? ", nullptr" : "") << ");\n";
    os << pad(ind) << "}\n";
}

}  // namespace lowering_detail
}  // namespace tiramisu
