// tiramisu/tiramisu.h — TEST INFRASTRUCTURE (wrapper side): some wrappers (wrapper_nn_block.cpp) include the whole
// compiler front-end although they only call the generated function and a few helpers (rtclock, median).  The
// wrappers are not generators: this stand-in only provides what they use — the standard headers the reference's
// header pulls in, and the reference's own utils.h (found further down the include path).
#pragma once
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <tiramisu/utils.h>
