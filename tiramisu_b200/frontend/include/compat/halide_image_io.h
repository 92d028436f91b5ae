/* halide_image_io.h — see Halide.h in this directory.  load_image here reads the SHAPE of a PNG (its IHDR chunk:
 * width, height, colour type) and returns a shape-only Halide::Buffer — all a generator needs to size its loop
 * bounds; pixel data is the wrapper's business. */
#ifndef TIRAMISU_B200_FRONTEND_HALIDE_IMAGE_IO_STANDIN_H
#define TIRAMISU_B200_FRONTEND_HALIDE_IMAGE_IO_STANDIN_H
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "Halide.h"
namespace Halide {
namespace Tools {
struct ImageShape {
    int width = 0, height = 0, channels = 0;
    template <typename T> operator Buffer<T>() const { return channels > 1 ? Buffer<T>(width, height, channels) : Buffer<T>(width, height); }
};
inline ImageShape load_image(const std::string &path)
{
    unsigned char h[26];
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f || std::fread(h, 1, sizeof h, f) != sizeof h || std::memcmp(h, "\x89PNG\r\n\x1a\n", 8) != 0 || std::memcmp(h + 12, "IHDR", 4) != 0) {
        std::fprintf(stderr, "tiramisu_b200 front-end: load_image: cannot read a PNG header from '%s'\n", path.c_str());
        std::exit(1);
    }
    std::fclose(f);
    auto be32 = [&](int o) { return (int)((uint32_t)h[o] << 24 | (uint32_t)h[o + 1] << 16 | (uint32_t)h[o + 2] << 8 | (uint32_t)h[o + 3]); };
    ImageShape s;
    s.width = be32(16);
    s.height = be32(20);
    const int colour = h[25];                       // 0 grey, 2 RGB, 3 palette (expanded to RGB by libpng users), 4 grey+A, 6 RGBA
    s.channels = colour == 0 ? 1 : (colour == 4 ? 2 : (colour == 6 ? 4 : 3));
    return s;
}
}  // namespace Tools
}  // namespace Halide
#endif
