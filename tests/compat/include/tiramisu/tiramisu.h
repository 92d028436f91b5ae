// tiramisu/tiramisu.h — TEST INFRASTRUCTURE: some wrappers (wrapper_nn_block.cpp) include the whole
// compiler front-end although they only call the generated function; the front-end (ISL, Halide IR)
// is out of scope, so this stand-in is empty.
#pragma once
