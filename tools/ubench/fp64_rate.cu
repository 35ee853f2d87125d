// fp64_rate.cu — microbenchmark: FP64 FMA and DMMA (mma.sync m8n8k4 f64) throughput per SM on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, int iters) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = a[i] * b + c;
  }
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double *out, int iters) {
  double acc0[2] = {0, 0}, acc1[2] = {0, 0}, acc2[2] = {0, 0}, acc3[2] = {0, 0};
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc0[0]), "+d"(acc0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc1[0]), "+d"(acc1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc2[0]), "+d"(acc2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc3[0]), "+d"(acc3[1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0[0] + acc0[1] + acc1[0] + acc1[1] + acc2[0] + acc2[1] + acc3[0] + acc3[1];
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
  double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 2; warps <= 32; warps *= 2) {
    int threads = warps * 32, iters = 20000; float ms;
    k_dfma<<<sms, threads>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dfma<<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * threads * iters * 8;
    printf("DFMA  warps/SM=%2d : %.2f TFLOP/s  (%.1f FMA/clk/SM at %.2f GHz nominal)\n", warps, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / (ghz * 1e9) / sms, ghz);
    k_dmma<<<sms, threads>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dmma<<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double mm = (double)sms * warps * iters * 4 * (8 * 8 * 4);
    printf("DMMA  warps/SM=%2d : %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", warps, 2 * mm / ms / 1e9, mm / (ms * 1e-3) / (ghz * 1e9) / sms);
  }
  // dependent-chain latency of DFMA: one warp, one chain
  return 0;
}
