// fused_reduce_cpu_gen — TEST INFRASTRUCTURE, written for this repo.
// A CPU schedule in the shape of reference tests/test_85.cpp: C = A x B where the zero-init is the slice k = -1 of
// the reduction's own domain and both definitions are fused inside all three loops.  The k loop carries the
// dependence, i and j do not: the re-targeting maps i, j onto the grid and keeps k (with both statements) in the
// thread — one launch instead of one per (i, j, k).
//   args: sizes{N, K, M} int32, A[N][K], B[K][M], C[N][M] (output);  fp32, k sequential, separate multiply and add
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    const std::string name = argc > 1 ? argv[1] : "fused_reduce_cpu_gen";
    tiramisu::init(name);
    function *fn = global::get_implicit_function();

    var s("s", 0, 3);
    input sizes("sizes", {s}, p_int32);
    constant N("N", sizes(0)), K("K", sizes(1)), M("M", sizes(2));
    var i("i", 0, N), j("j", 0, M), k("k", 0, K);
    input A("A", {i, k}, p_float32);
    input B("B", {k, j}, p_float32);

    computation C("[N, K, M] -> {C[i, j, -1]: 0 <= i < N and 0 <= j < M}", expr(0.f), true, p_float32, fn);
    C.add_definitions("[N, K, M] -> {C[i, j, k]: 0 <= i < N and 0 <= j < M and 0 <= k < K}",
                      C(i, j, k - 1) + A(i, k) * B(k, j), true, p_float32, fn);
    C.before(C.get_update(1), k);

    buffer b_A("b_A", {N, K}, p_float32, a_input);
    buffer b_B("b_B", {K, M}, p_float32, a_input);
    buffer b_C("b_C", {N, M}, p_float32, a_output);
    A.store_in(&b_A);
    B.store_in(&b_B);
    C.set_access("{C[i, j, k] -> b_C[i, j]}");

    tiramisu::codegen({sizes.get_buffer(), &b_A, &b_B, &b_C}, name + ".o");
    return 0;
}
