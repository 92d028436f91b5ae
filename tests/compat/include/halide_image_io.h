// halide_image_io.h — TEST INFRASTRUCTURE: the wrappers compiled by tests/compat include this header
// but the ones we build do not call into it (no libpng/libjpeg in the image).
#pragma once
