// Micro-benchmarks for the roofline denominators that MEASURED_PEAKS.json does not hold:
//   * write-only HBM stream (the interior-rule writer is a pure store stream),
//   * dependent-free fp64 FMA issue rate (root finding is fp64 issue bound).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peaks peaks.cu && ./peaks
#include <cstdio>
#include <cuda_runtime.h>

__global__ void write_kernel(double *out, size_t n, double v)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = v;
}
__global__ void write2_kernel(double2 *out, size_t n, double v)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = make_double2(v, v);
}
__global__ void copy_kernel(const double2 *in, double2 *out, size_t n)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void dfma_kernel(double *out, int iters)
{
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int main()
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t n = (size_t)1 << 31;  // 16 GiB of doubles
  double *buf, *buf2;
  cudaMalloc(&buf, n * 8);
  cudaMalloc(&buf2, n * 8);
  float ms;
  for (int variant = 0; variant < 3; ++variant) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      if (variant == 0) write_kernel<<<148 * 16, 512>>>(buf, n, 1.0);
      if (variant == 1) write2_kernel<<<148 * 16, 512>>>((double2 *)buf, n / 2, 1.0);
      if (variant == 2) copy_kernel<<<148 * 16, 512>>>((const double2 *)buf, (double2 *)buf2, n / 2);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double bytes = variant == 2 ? 2.0 * n * 8 : 1.0 * n * 8;
    printf("%s: %.3f ms  %.1f GB/s\n", variant == 0 ? "write f64" : variant == 1 ? "write f64x2" : "copy f64x2 (read+write)", best,
      bytes / best / 1e6);
  }
  {
    const int iters = 20000, blocks = 148 * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      dfma_kernel<<<blocks, threads>>>(buf, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double fma = (double)iters * 8 * blocks * threads;
    printf("dfma: %.3f ms  %.2f TFMA/s = %.2f TFLOP/s fp64\n", best, fma / best / 1e9, 2 * fma / best / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
