// Measures what bounds a latency-sensitive FP64 kernel on one SM: cycles per dependent DFMA for W warps per block
// (one block on one SM) with ILP independent chains per thread.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/probe/fp64_probe tools/probe/fp64_probe.cu ; run on the GPU box.  Profiling aid, not part of the library.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 1e-3 + k;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* cyc) {
  const int iters = 4096;
  chain<ILP><<<1, warps * 32>>>(out, cyc, iters, 0.999999, 1e-9);
  chain<ILP><<<1, warps * 32>>>(out, cyc, iters, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double per_iter = (double)h / iters;
  printf("warps/SM %2d  ILP %d : %.2f cycles per iteration, %.2f cycles per DFMA per warp, SM rate %.3f warp-DFMA/cycle (pipe peak 2.0)\n",
         warps, ILP, per_iter, per_iter / ILP, warps * ILP / per_iter);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 2048);
  cudaMalloc(&cyc, sizeof(long long) * 8);
  for (int w : {1, 4, 8, 12, 14, 16, 20, 24, 32}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); }
  return 0;
}
