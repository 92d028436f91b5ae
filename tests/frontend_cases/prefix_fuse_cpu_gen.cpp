// prefix_fuse_cpu_gen — TEST INFRASTRUCTURE, written for this repo.
// A fused run whose fuse depths differ:
//     T0(i)    : t[i] = 0            U0(i)    : u[i] = 0
//     S1(i, k) : t[i] = t[i] + A(i, k)
//     S2(i, k) : u[i] = u[i] + t[i]          ordered  T0.then(U0, i).then(S1, i).then(S2, k)
// S2 is fused with S1 INSIDE k: it reads the running prefix sum t[i] after every k, so u[i] = sum_k prefix_k(A[i]).
// Un-fusing the run down to "each statement over its whole box" (loop distribution) would make S2 read the final
// t[i] in every iteration (u[i] = K * total): legal for the T0 | U0 | S1 boundaries, not for S1 | S2.  The level
// k is not data parallel and only levels below the run's minimum fuse depth are looked at, so the distribution
// must leave this run alone (lowering_retarget.cpp: distribute_parallel_shared_loops, uniform-depth rule).
//   args: sizes{N, K} int32, A[N][K], u[N] (output);  fp32 (small integers: exact)
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    const std::string name = argc > 1 ? argv[1] : "prefix_fuse_cpu_gen";
    tiramisu::init(name);

    var s("s", 0, 2);
    input sizes("sizes", {s}, p_int32);
    constant N("N", sizes(0)), K("K", sizes(1));
    var i("i", 0, N), k("k", 0, K);
    input A("A", {i, k}, p_float32);

    computation T0("T0", {i}, expr(0.f));
    computation U0("U0", {i}, expr(0.f));
    computation S1("S1", {i, k}, p_float32);
    S1.set_expression(T0(i) + A(i, k));
    computation S2("S2", {i, k}, p_float32);
    S2.set_expression(U0(i) + T0(i));

    T0.then(U0, i).then(S1, i).then(S2, k);

    buffer b_A("b_A", {N, K}, p_float32, a_input);
    buffer b_t("b_t", {N}, p_float32, a_temporary);
    buffer b_u("b_u", {N}, p_float32, a_output);
    A.store_in(&b_A);
    T0.store_in(&b_t);
    U0.store_in(&b_u);
    S1.store_in(&b_t, {i});
    S2.store_in(&b_u, {i});

    tiramisu::codegen({sizes.get_buffer(), &b_A, &b_u}, name + ".o");
    return 0;
}
