// FP64 tensor-core probe (developer tool): throughput of mma.sync.m8n8k4.f64 with register operands, to decide whether
// the k_states = 30 products should move from DFMA (shared-memory bound: one operand per multiply-add) to DMMA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_probe tools/dmma_probe.cu ; run: build/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_kernel(int iters, double* sink) {
  double c0[NACC], c1[NACC];
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
  for (int i = 0; i < NACC; ++i) c0[i] = c1[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a + k, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  if (s == 123.456) sink[0] = s;
}

template <int NACC>
void run(int warps_per_cta, const char* tag) {
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* sink;
  cudaMalloc(&sink, 8);
  const int iters = 4000, blocks = sms * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dmma_kernel<NACC><<<blocks, 32 * warps_per_cta>>>(10, sink);
  cudaEventRecord(e0);
  dmma_kernel<NACC><<<blocks, 32 * warps_per_cta>>>(iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = (double)blocks * warps_per_cta * iters * 8.0 * NACC * 512.0;
  printf("{\"probe\": \"dmma_m8n8k4\", \"acc_per_warp\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"%s\": true}\n", NACC,
         4 * warps_per_cta, ms, flops / ms * 1e-9, tag);
}

int main() {
  run<1>(4, "dependent_chain");
  run<4>(1, "w4");
  run<8>(1, "w4");
  run<8>(2, "w8");
  run<16>(2, "w8");
  run<16>(4, "w16");
  return 0;
}
