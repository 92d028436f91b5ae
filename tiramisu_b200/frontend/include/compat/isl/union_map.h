/* isl/union_map.h — the tiramisu_b200 front-end does not use ISL; the reference's generators include these headers without calling into them (stand-in so that they compile unmodified). */
