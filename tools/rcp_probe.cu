// rcp_probe.cu - dev tool: accuracy of rcp.approx.ftz.f64 (MUFU.RCP64H) and of 1 / 2 Newton refinements vs IEEE division.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tools/rcp_probe tools/rcp_probe.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__global__ void probe(double* maxerr, unsigned long long seed, int per_thread) {
  unsigned long long s = seed + (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
  double e0 = 0, e1 = 0, e2 = 0;
  for (int i = 0; i < per_thread; ++i) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    // random mantissa, exponent in [-1000, 1000]
    const unsigned long long mant = s >> 12;
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    const long long ex = (long long)((s >> 33) % 2001) - 1000 + 1023;
    const double x = __longlong_as_double((ex << 52) | mant);
    const double ref = 1.0 / x;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    e0 = fmax(e0, fabs(r / ref - 1.0));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    double r1 = fma(r, e, r);
    e1 = fmax(e1, fabs(r1 - ref) / fabs(ref));
    e = fma(-x, r1, 1.0);
    double r2 = fma(r1, e, r1);
    e2 = fmax(e2, fabs(r2 - ref) / fabs(ref));
  }
  for (int o = 16; o; o >>= 1) {
    e0 = fmax(e0, __shfl_xor_sync(0xffffffffu, e0, o));
    e1 = fmax(e1, __shfl_xor_sync(0xffffffffu, e1, o));
    e2 = fmax(e2, __shfl_xor_sync(0xffffffffu, e2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    maxerr[3 * w] = e0; maxerr[3 * w + 1] = e1; maxerr[3 * w + 2] = e2;
  }
}

int main() {
  const int blocks = 592, threads = 256, warps = blocks * threads / 32;
  double* d;
  cudaMalloc(&d, warps * 3 * sizeof(double));
  probe<<<blocks, threads>>>(d, 12345, 20000);
  double* h = new double[warps * 3];
  cudaMemcpy(h, d, warps * 3 * sizeof(double), cudaMemcpyDeviceToHost);
  double m[3] = {0, 0, 0};
  for (int w = 0; w < warps; ++w) for (int k = 0; k < 3; ++k) m[k] = fmax(m[k], h[3 * w + k]);
  printf("%.3g samples; max rel err: seed %.3e (2^%.1f), seed + cubic step %.3e (%.2f ulp), + Newton step %.3e (%.2f ulp); %s\n",
         (double)blocks * threads * 20000, m[0], log2(m[0]), m[1], m[1] / 1.11e-16, m[2], m[2] / 1.11e-16,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
