// fp64_issue_probe.cu - dev tool: DFMA dependent latency and per-warp / per-SMSP issue rate on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_issue_probe tools/fp64_issue_probe.cu
// Prints cycles per DFMA-round for ILP = 1..16 independent chains per thread and 1, 2, 4, 8 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void probe(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps_per_smsp, double* out, long long* cyc) {
  const int iters = 4096;
  const int threads = 32 * 4 * warps_per_smsp;
  probe<ILP><<<1, threads>>>(out, cyc, iters, 0.999999, 1e-7);
  probe<ILP><<<1, threads>>>(out, cyc, iters, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("ILP %2d warps/SMSP %d : %7.2f cycles per round, %6.2f cycles per DFMA per warp, pipe busy %5.1f %%\n", ILP,
         warps_per_smsp, (double)c / iters, (double)c / iters / ILP, 100.0 * 2.0 * ILP * warps_per_smsp / ((double)c / iters));
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * 8 * 64);
  cudaMalloc(&cyc, 8 * 64);
  for (int w : {1, 2, 3, 4, 6, 8}) {
    run<1>(w, out, cyc);
    run<2>(w, out, cyc);
    run<4>(w, out, cyc);
    run<8>(w, out, cyc);
    run<16>(w, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
