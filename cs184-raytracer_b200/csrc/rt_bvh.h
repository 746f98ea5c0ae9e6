// rt_bvh.h — LBVH build entry point (see rt_bvh.cu).
#pragma once
#include "rt_device.cuh"

namespace rt {

// Builds one LBVH per primitive group over the n primitive codes in d_codes (device; h_codes
// is the host copy), groups being consecutive ranges of sizes group_sizes[0..ngroups), and
// joins them under super nodes so that node 0 is always the root.  n < 2 builds nothing
// (*out_nodes = nullptr).  extra_abs: largest |coordinate| of ray origins outside the
// primitives (the camera eye), folded into the box padding.  Returns RT_OK or RT_ERR_CUDA.
int build_lbvh(const DScene& S, const int* d_codes, const int* h_codes, int n, const int* group_sizes, int ngroups,
               float extra_abs, cudaStream_t stream, BvhNode** out_nodes, size_t* out_count, int* launches,
               char* err, int errlen);

}  // namespace rt
