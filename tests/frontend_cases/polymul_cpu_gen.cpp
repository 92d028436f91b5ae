// polymul_cpu_gen — TEST INFRASTRUCTURE, written for this repo.
// Polynomial multiply / full convolution accumulate: S(i, j) = O(i + j) + A(i) * B(j), stored at b_O[i + j].
// Both iterators appear (affinely, non-zero coefficient) in the store index, yet the map (i, j) -> i + j is not
// injective: every anti-diagonal is a reduction.  The re-targeting must keep one of the two iterators sequential
// in the thread (jointly-injective rule, lowering_retarget.cpp: parallel_iterators) — mapping both onto the grid
// is a data race that loses sums on the GPU.
//   args: sizes{N, M} int32, A[N], B[M], O[N + M - 1] (in/out, zero on entry);  fp32
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    const std::string name = argc > 1 ? argv[1] : "polymul_cpu_gen";
    tiramisu::init(name);

    var s("s", 0, 2);
    input sizes("sizes", {s}, p_int32);
    constant N("N", sizes(0)), M("M", sizes(1));
    var i("i", 0, N), j("j", 0, M), o("o", 0, N + M - 1);
    input A("A", {i}, p_float32);
    input B("B", {j}, p_float32);
    input O("O", {o}, p_float32);

    computation S("S", {i, j}, O(i + j) + A(i) * B(j));

    buffer b_A("b_A", {N}, p_float32, a_input);
    buffer b_B("b_B", {M}, p_float32, a_input);
    buffer b_O("b_O", {N + M - 1}, p_float32, a_output);
    A.store_in(&b_A);
    B.store_in(&b_B);
    O.store_in(&b_O);
    S.store_in(&b_O, {i + j});

    tiramisu::codegen({sizes.get_buffer(), &b_A, &b_B, &b_O}, name + ".o");
    return 0;
}
