/* halide_image_io.h — see Halide.h in this directory. */
