// Micro-benchmark: FP64 peak of this GPU through (a) DFMA on the CUDA cores, (b) DMMA m8n8k4 on the tensor
// cores.  Used only to state the roofline of the regression Gram kernel (profiles/README.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  double s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[16][2];
  for (int i = 0; i < 16; ++i) { c[i][0] = 0; c[i][1] = 0; }
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1e-3;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = 32 * (warps > 32 ? 32 : warps), blocks = sms * (warps > 32 ? warps / 32 : 1);
    const int iters = 20000;
    float ms;
    dfma_kernel<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double f1 = 2.0 * 16 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    dmma_kernel<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double f2 = 2.0 * 256 * 16 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    printf("warps/SM %2d: DFMA %.1f TFLOP/s   DMMA m8n8k4 %.1f TFLOP/s\n", warps, f1, f2);
  }
  return 0;
}
