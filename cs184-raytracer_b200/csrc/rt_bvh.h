// rt_bvh.h — LBVH build entry point (see rt_bvh.cu).
#pragma once
#include "rt_device.cuh"

namespace rt {

// Builds the LBVH over the n primitive codes in d_codes (device).  n < 2 builds nothing
// (*out_nodes = nullptr).  extra_abs: largest |coordinate| of ray origins outside the
// primitives (the camera eye), folded into the box padding.  Returns RT_OK or RT_ERR_CUDA.
int build_lbvh(const DScene& S, const int* d_codes, int n, float extra_abs, cudaStream_t stream,
               BvhNode** out_nodes, int* launches, char* err, int errlen);

}  // namespace rt
