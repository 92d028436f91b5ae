// C += A * B with run-time sizes through the front-end (TEST INFRASTRUCTURE, written for this repo):
// the contraction pattern of the reference's tests/test_167.cpp:24-28, but with symbolic M, N, K
// and B stored transposed, so the pattern lowering has to hand leading dimensions and the
// transpose flag to the runtime GEMM (TIRAMISU_B200_LOWERING=generic emits the SIMT loop nest).
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    tiramisu::init(argc > 1 ? argv[1] : "gemm_sym_gen");

    input sizes("sizes", {"s_i"}, {3}, p_int32);
    constant M("M", sizes(0)), N("N", sizes(1)), K("K", sizes(2));
    var i("i", 0, M), j("j", 0, N), k("k", 0, K);
    var i0("i0"), j0("j0"), i1("i1"), j1("j1");

    buffer b_A("b_A", {M, K}, p_float32, a_input);
    buffer b_Bt("b_Bt", {N, K}, p_float32, a_input);
    buffer b_C("b_C", {M, N}, p_float32, a_output);
    buffer g_A("g_A", {M, K}, p_float32, a_temporary);
    buffer g_Bt("g_Bt", {N, K}, p_float32, a_temporary);
    buffer g_C("g_C", {M, N}, p_float32, a_temporary);
    g_A.tag_gpu_global(); g_Bt.tag_gpu_global(); g_C.tag_gpu_global();

    input A("A", {i, k}, p_float32);
    input B("B", {k, j}, p_float32);
    A.store_in(&g_A);
    B.store_in(&g_Bt, {j, k});
    computation C("C", {i, j, k}, p_float32);
    C.set_expression(C(i, j, k - 1) + A(i, k) * B(k, j));
    C.store_in(&g_C, {i, j});

    computation cA({}, memcpy(b_A, g_A)), cB({}, memcpy(b_Bt, g_Bt)), cC({}, memcpy(b_C, g_C)), cOut({}, memcpy(g_C, b_C));
    C.gpu_tile(i, j, 16, 16, i0, j0, i1, j1);
    cA.then(cB, computation::root).then(cC, computation::root).then(C, computation::root).then(cOut, computation::root);

    tiramisu::codegen({sizes.get_buffer(), &b_A, &b_Bt, &b_C}, std::string(argc > 1 ? argv[1] : "gemm_sym_gen") + ".o", true);
    return 0;
}
