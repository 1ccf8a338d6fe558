// Micro-benchmark: FP64 pipe throughput / latency on the box (DFMA, DADD, DMUL, MUFU.RSQ64H, F2I/I2F.F64), per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP> __global__ void k(double *out, int iters, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = a + threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (OP == 0) v[i] = fma(v[i], a, b);
      else if (OP == 1) v[i] = __dadd_rn(v[i], b);
      else if (OP == 2) v[i] = __dmul_rn(v[i], a);
      else if (OP == 3) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v[i])); v[i] = y; }
      else if (OP == 4) { int q = __double2int_rz(v[i]); v[i] = (double)(q + it); }
      else if (OP == 5) { v[i] = fmax(v[i], b); }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int OP> void run(const char *name, int warps_per_sm, int opsper) {
  int nsm = 148; int threads = 128; int blocks = nsm * (warps_per_sm * 32 / threads > 0 ? warps_per_sm * 32 / threads : 1);
  if (warps_per_sm * 32 < threads) threads = warps_per_sm * 32;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, OP><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<ILP, OP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = (double)blocks * threads / 32 * iters * ILP * opsper;
  double per_smsp_per_clk = warp_instr / (nsm * 4) / (ms * 1e-3 * 1.965e9);
  printf("%-10s ILP %2d warps/SM %2d : %.3f ms  %.4f warp-instr/clk/SMSP  (%.1f cycles per instr per warp)\n", name, ILP, warps_per_sm, ms, per_smsp_per_clk,
         (ms * 1e-3 * 1.965e9) / (iters * ILP * opsper));
  cudaFree(out);
}
int main() {
  run<1, 0>("DFMA", 4, 1); run<2, 0>("DFMA", 4, 1); run<4, 0>("DFMA", 4, 1); run<8, 0>("DFMA", 4, 1);
  run<1, 0>("DFMA", 8, 1); run<2, 0>("DFMA", 8, 1); run<4, 0>("DFMA", 8, 1); run<8, 0>("DFMA", 8, 1);
  run<4, 0>("DFMA", 16, 1); run<8, 0>("DFMA", 16, 1); run<8, 0>("DFMA", 32, 1);
  run<8, 1>("DADD", 8, 1); run<8, 1>("DADD", 16, 1); run<8, 2>("DMUL", 8, 1); run<8, 2>("DMUL", 16, 1); run<8, 5>("DMNMX", 16, 1);
  run<1, 3>("RSQ64H", 4, 1); run<8, 3>("RSQ64H", 8, 1); run<8, 3>("RSQ64H", 16, 1);
  run<1, 4>("F2I+I2F", 4, 2); run<8, 4>("F2I+I2F", 8, 2); run<8, 4>("F2I+I2F", 16, 2);
  return 0;
}
