// cuda_floor.cu — the smallest CUDA program there is: create a context on device 0, launch one empty kernel,
// synchronise, exit.  tools/walltime.py times it next to `as2` to show how much of a short run's wall time is
// the CUDA driver / context start-up that no single-shot GPU program can avoid (measurement aid, not product).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_nothing() {}
int main() {
    if (cudaFree(0) != cudaSuccess) { std::fprintf(stderr, "no CUDA device\n"); return 1; }
    k_nothing<<<1, 32>>>();
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
