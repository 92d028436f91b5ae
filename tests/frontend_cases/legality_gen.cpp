// Legality decisions of the CPU-schedule re-targeting (TEST INFRASTRUCTURE, written for this repo): six small
// functions whose generated code must show the conservative choice —
//   fuse_pointwise   : T = 2*A ; U = T + 1             element-wise exchange        -> ONE fused kernel
//   no_fuse_stencil  : T = 2*A ; U = T(i,j) + T(i,j+1) reads a neighbour's T        -> two kernels
//   inplace_serial   : S(i,j) = S(i,j) + S(0,0)        reads another element of the buffer it writes -> one thread
//   no_distribute    : S0.then(S1, j), S1 reads T(i,j-1) across the fused loops     -> loops stay on the host
//   no_grid_fuse     : T += A over k ; U += T(i,j+1) over k, fused inside i, j, k  (a neighbour's T) -> host loops stay
//   grid_fuse        : the same with U += T(i,j)                                   -> i, j on the grid, k in the thread
// Only the generated text is checked (tests/test_frontend.py); nothing here runs on a GPU.
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

static void declare_io(buffer *&bA, buffer *&bT, buffer *&bU, int n)
{
    bA = new buffer("b_A", {n, n + 1}, p_float32, a_input);
    bT = new buffer("b_T", {n, n + 1}, p_float32, a_temporary);
    bU = new buffer("b_U", {n, n}, p_float32, a_output);
}

int main(int, char **)
{
    const int n = 8;
    {
        tiramisu::init("fuse_pointwise");
        var i("i", 0, n), j("j", 0, n);
        buffer *bA, *bT, *bU; declare_io(bA, bT, bU, n);
        input A("A", {i, j}, p_float32); A.store_in(bA, {i, j});
        computation T("T", {i, j}, A(i, j) * 2.f); T.store_in(bT, {i, j});
        computation U("U", {i, j}, T(i, j) + 1.f); U.store_in(bU);
        T.then(U, computation::root);
        tiramisu::codegen({bA, bU}, "fuse_pointwise.o");
    }
    {
        tiramisu::init("no_fuse_stencil");
        var i("i", 0, n), j("j", 0, n), jw("j", 0, n + 1);
        buffer *bA, *bT, *bU; declare_io(bA, bT, bU, n);
        input A("A", {i, jw}, p_float32); A.store_in(bA);
        computation T("T", {i, jw}, A(i, jw) * 2.f); T.store_in(bT);
        computation U("U", {i, j}, T(i, j) + T(i, j + 1)); U.store_in(bU);
        T.then(U, computation::root);
        tiramisu::codegen({bA, bU}, "no_fuse_stencil.o");
    }
    {
        tiramisu::init("inplace_serial");
        var i("i", 0, n), j("j", 0, n);
        buffer bS("b_S", {n, n}, p_float32, a_output);
        computation S("S", {i, j}, p_float32);
        S.set_expression(S(i, j) + S(0, 0));
        S.store_in(&bS);
        tiramisu::codegen({&bS}, "inplace_serial.o");
    }
    {
        tiramisu::init("no_distribute");
        var i("i", 0, n), j("j", 1, n);
        buffer *bA, *bT, *bU; declare_io(bA, bT, bU, n);
        input A("A", {i, j}, p_float32); A.store_in(bA, {i, j});
        computation T("T", {i, j}, A(i, j) * 2.f); T.store_in(bT, {i, j});
        computation U("U", {i, j}, T(i, j) + T(i, j - 1)); U.store_in(bU, {i, j});
        T.then(U, j);
        tiramisu::codegen({bA, bU}, "no_distribute.o");
    }
    for (int neighbour = 0; neighbour < 2; neighbour++) {
        tiramisu::init(neighbour ? "no_grid_fuse" : "grid_fuse");
        var i("i", 0, n), j("j", 0, n), k("k", 0, n);
        buffer bA("b_A", {n, n}, p_float32, a_input);
        buffer bT("b_T", {n, n + 1}, p_float32, a_output);
        buffer bU("b_U", {n, n}, p_float32, a_output);
        input A("A", {i, k}, p_float32); A.store_in(&bA);
        computation T("T", {i, j, k}, p_float32);
        T.set_expression(T(i, j, k - 1) + A(i, k));
        T.store_in(&bT, {i, j});
        computation U("U", {i, j, k}, p_float32);
        U.set_expression(U(i, j, k - 1) + T(i, j + neighbour, k));
        U.store_in(&bU, {i, j});
        T.then(U, k);
        tiramisu::codegen({&bA, &bT, &bU}, std::string(neighbour ? "no_grid_fuse" : "grid_fuse") + ".o");
    }
    return 0;
}
