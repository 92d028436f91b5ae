// pieces_cpu_gen — TEST INFRASTRUCTURE, written for this repo.
// The old-style API the reference's sgemm_generator.cpp uses, at a size a test can afford and without the reference checkout:
// ISL-string domains, a statement cut into a full-tile piece and a partial-tile piece (`8*floor(M/8)` bounds), scheduling by
// loop LEVEL (tile / tag_parallel_level / tag_vector_level(int)), the ordering command after_low_level, and a temporary
// nothing reads.  B = 2 A + 1 over [N][M]; expected lowering: the pieces are one statement again (one kernel, one launch),
// the dead copy is not computed.
//   args: SIZES{N, M} int32, A[N][M], B[N][M] (output)
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    const std::string name = argc > 1 ? argv[1] : "pieces_cpu_gen";
    tiramisu::function fn(name);
    tiramisu::var i("i"), j("j");
    computation SIZES("{SIZES[i]}", expr(), false, p_int32, &fn);
    computation A("{A[i,j]}", expr(), false, p_float32, &fn);
    constant N_cst("N", SIZES(0), p_int32, true, NULL, 0, &fn);
    constant M_cst("M", SIZES(1), p_int32, true, NULL, 0, &fn);

    computation copy("[N, M]->{copy[i,j]: 0<=i<N and 0<=j<M}", A(i, j), true, p_float32, &fn);                       // dead
    computation S("[N, M]->{S[i,j]: 0<=i<N and 0<=j<8*floor(M/8)}", A(i, j) * 2.f + 1.f, true, p_float32, &fn);
    computation S_p("[N, M]->{S_p[i,j]: 0<=i<N and 8*floor(M/8)<=j<M}", A(i, j) * 2.f + 1.f, true, p_float32, &fn);

    fn.add_context_constraints("[N, M]->{: N>0 and M>8}");
    fn.align_schedules();
    S.tile(0, 1, 4, 8);
    S_p.tile(0, 1, 4, 8);
    S.tag_parallel_level(0);
    S.tag_vector_level(3, 8);
    S.tag_unroll_level(7);                 // a level the statement does not have: inert, as in the reference
    S.after_low_level(copy, -1);
    S_p.after_low_level(S, 1);

    buffer b_SIZES("b_SIZES", {2}, p_int32, a_input, &fn);
    buffer b_A("b_A", {var(p_int32, "N"), var(p_int32, "M")}, p_float32, a_input, &fn);
    buffer b_T("b_T", {var(p_int32, "N"), var(p_int32, "M")}, p_float32, a_temporary, &fn);
    buffer b_B("b_B", {var(p_int32, "N"), var(p_int32, "M")}, p_float32, a_output, &fn);
    SIZES.set_access("{SIZES[i]->b_SIZES[i]}");
    A.set_access("{A[i,j]->b_A[i,j]}");
    copy.set_access("{copy[i,j]->b_T[i,j]}");
    S.set_access("{S[i,j]->b_B[i,j]}");
    S_p.set_access("{S_p[i,j]->b_B[i,j]}");

    fn.set_arguments({&b_SIZES, &b_A, &b_B});
    fn.gen_time_space_domain();
    fn.gen_isl_ast();
    fn.gen_halide_stmt();
    fn.gen_halide_obj(name + ".o");
    return 0;
}
