// fp64 pipe micro-benchmark for the roofline argument in DESIGN.md: dependent-issue latency and
// per-SM throughput of DADD / DMUL / DFMA on sm_100a.   nvcc -O3 -arch=sm_100a -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void chain(double* out, long long* cyc, int n, double a, double b)
{
  double x[ILP];
  #pragma unroll
  for (int j = 0; j < ILP; j++) x[j] = a + j + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    #pragma unroll
    for (int j = 0; j < ILP; j++) {
      if (OP == 0) x[j] = __dadd_rn(x[j], b);
      else if (OP == 1) x[j] = __dmul_rn(x[j], b);
      else x[j] = __fma_rn(x[j], b, a);
    }
  }
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < ILP; j++) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP> void run(const char* name, int threads, int blocks)
{
  double* out; long long* cyc; cudaMalloc(&out, 8 * threads * blocks); cudaMalloc(&cyc, 8);
  const int n = 4096;
  chain<OP, ILP><<<blocks, threads>>>(out, cyc, n, 1.0, 1.0000001);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); chain<OP, ILP><<<blocks, threads>>>(out, cyc, n, 1.0, 1.0000001); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-5s ILP=%d threads/CTA=%4d CTAs=%4d : %.2f cycles per dependent step, %.1f warp-instr/clk/SM\n", name, ILP, threads, blocks,
         (double)c / n, (double)n * ILP * (threads / 32) * (blocks / 148.0 > 1 ? 1 : 1) * 1.0 / c * (blocks >= 148 ? (double)blocks / 148 : 1));
  cudaFree(out); cudaFree(cyc);
}
int main()
{
  run<0, 1>("DADD", 32, 1); run<1, 1>("DMUL", 32, 1); run<2, 1>("DFMA", 32, 1);
  run<0, 2>("DADD", 32, 1); run<0, 4>("DADD", 32, 1); run<0, 8>("DADD", 32, 1);
  run<0, 1>("DADD", 128, 148); run<0, 1>("DADD", 384, 148); run<0, 2>("DADD", 384, 148); run<0, 4>("DADD", 384, 148);
  run<0, 1>("DADD", 1024, 148); run<0, 4>("DADD", 1024, 148); run<2, 4>("DFMA", 1024, 148);
  return 0;
}
