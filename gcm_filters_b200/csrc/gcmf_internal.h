// gcmf_internal.h -- private definitions shared by the translation units of libgcmf.so.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/gcmf.h"
#include "gcmf_stencils.cuh"

struct gcmf_plan {
    gcmf_plan_desc desc;
    int ncomp;
    int n_planes;
    gcmf::PlaneRef plane[gcmf::MAX_PLANES];
    int32_t n_steps;
    std::vector<double> p;  // Chebyshev coefficients p[0..n_steps]
    double c;
    int sm_count;
    int steps_per_block;  // 0 = auto (fuse when eligible), 1 = never fuse, 2..4 = cap
    // tensor maps of the arrays the fused kernels stage through the TMA engine, encoded once per distinct array
    struct MapEntry {
        const void* p;
        int64_t pitch, bstride, nb;
        gcmf::TmaDesc d;
    };
    std::vector<MapEntry> state_maps;   // small LRU: the recurrence rotates through <= 5 arrays
    MapEntry coef_maps[3];              // ce, cn, ra of the FLUX family (p == nullptr: not encoded yet)
};

int gcmf_set_error(int code, const char* fmt, ...);
void gcmf_count_launch(int64_t n);
