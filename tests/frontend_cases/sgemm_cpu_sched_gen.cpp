// C = a*(A*B) + b*C with the argument list of the reference's sgemm_tiramisu
// (benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_generator.cpp:90-110: acc = 0; acc += A*B over k;
// C = a*acc + b*C) and a CPU schedule (parallelize i, vectorize j) — TEST INFRASTRUCTURE written for this
// repo.  Re-targeted by the lowering: the reduction loop k stays sequential inside each thread, so the
// result is bit-identical to the x86 statement order.
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    tiramisu::init(argc > 1 ? argv[1] : "sgemm_cpu_sched_gen");

    input SIZES("SIZES", {"s_i"}, {3}, p_int32);
    input alpha("alpha", {"a_i"}, {1}, p_float32);
    input beta("beta", {"b_i"}, {1}, p_float32);
    constant N("N", SIZES(0)), M("M", SIZES(1)), K("K", SIZES(2));
    constant a("a", alpha(0)), b("b", beta(0));
    var i("i", 0, N), j("j", 0, M), k("k", 0, K);

    input A("A", {i, k}, p_float32);
    input B("B", {k, j}, p_float32);
    computation acc_init("acc_init", {i, j}, expr(0.f));
    computation acc("acc", {i, j, k}, p_float32);
    acc.set_expression(acc(i, j, k - 1) + A(i, k) * B(k, j));
    computation C("C", {i, j}, p_float32);
    C.set_expression(a * acc(i, j, 0) + b * C(i, j));

    buffer b_A("b_A", {N, K}, p_float32, a_input);
    buffer b_B("b_B", {K, M}, p_float32, a_input);
    buffer b_C("b_C", {N, M}, p_float32, a_output);
    buffer b_acc("b_acc", {N, M}, p_float32, a_temporary);
    A.store_in(&b_A);
    B.store_in(&b_B);
    acc_init.store_in(&b_acc);
    acc.store_in(&b_acc, {i, j});
    C.store_in(&b_C);

    for (computation *c : std::vector<computation *>{&acc_init, &acc, &C}) {
        c->parallelize(i);
        c->vectorize(j, 8);
    }
    acc_init.then(acc, computation::root).then(C, computation::root);

    tiramisu::codegen({SIZES.get_buffer(), alpha.get_buffer(), beta.get_buffer(), &b_A, &b_B, &b_C},
                      std::string(argc > 1 ? argv[1] : "sgemm_cpu_sched_gen") + ".o");
    return 0;
}
