// tiramisu/tiramisu.h — umbrella header of the tiramisu_b200 front-end
// (same role as /root/reference/include/tiramisu/tiramisu.h:1-10).
#ifndef TIRAMISU_B200_FRONTEND_TIRAMISU_H
#define TIRAMISU_B200_FRONTEND_TIRAMISU_H
#include <tiramisu/type.h>
#include <tiramisu/expr.h>
#include <tiramisu/core.h>
#include <tiramisu/block.h>
#include <tiramisu/debug.h>
#include <tiramisu/utils.h>
#endif
