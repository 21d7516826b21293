// Microbenchmark behind DESIGN.md §4 "what bounds the step": issue rates of the pipes the propose kernels live on.
// Each thread runs FP independent DFMA chains and INT independent integer chains per loop iteration (all register-
// resident, no memory traffic). Throughput cases: 4 blocks x 256 threads per SM (8 warps per scheduler, like the leaf
// kernel); latency cases: one warp per SM with one or two dependent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_model issue_model.cu && ./issue_model
#include <cstdio>
#include <cuda_runtime.h>

// IOP 0: add.u32 (ALU pipe), 1: mad.lo.u32 (IMAD, the FMA-heavy pipe)
template <int FP, int INT, int IOP>
__global__ void __launch_bounds__(256, 4) k_mix(double* out, unsigned* iout, int iters, double a, unsigned m) {
  double x[FP > 0 ? FP : 1];
  unsigned y[INT > 0 ? INT : 1];
#pragma unroll
  for (int k = 0; k < (FP > 0 ? FP : 1); ++k) x[k] = 1.0 + threadIdx.x * 1e-9 + k;
#pragma unroll
  for (int k = 0; k < (INT > 0 ? INT : 1); ++k) y[k] = threadIdx.x + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < FP; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(x[k]) : "d"(a));
#pragma unroll
      for (int k = 0; k < INT; ++k) {
        if (IOP == 0)
          asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(m));
        else
          asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(y[k]) : "r"(m));
      }
    }
  }
  double s = 0;
  unsigned t = 0;
#pragma unroll
  for (int k = 0; k < (FP > 0 ? FP : 1); ++k) s += x[k];
#pragma unroll
  for (int k = 0; k < (INT > 0 ? INT : 1); ++k) t ^= y[k];
  if (s == 123.456) out[0] = s;
  if (t == 0x12345u) iout[0] = t;
}

static double g_mhz;
static int g_sms;
static double* g_d;
static unsigned* g_di;

template <int FP, int INT, int IOP>
void run(const char* name, int blocksPerSm, int threads) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_mix<FP, INT, IOP><<<g_sms * blocksPerSm, threads>>>(g_d, g_di, 100, 1e-30, 3u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_mix<FP, INT, IOP><<<g_sms * blocksPerSm, threads>>>(g_d, g_di, iters, 1e-30, 3u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warpsPerSched = blocksPerSm * threads / 32 / 4.0;
  const double clocks = ms * 1e-3 * g_mhz * 1e6;
  const double perWarp = (double)iters * 8 * (FP + INT);
  if (warpsPerSched >= 1.0)
    printf("%-34s fp64 share %.2f  %.3f ms  warp instructions per clock per scheduler %.3f\n", name,
           (double)FP / (FP + INT), ms, perWarp * warpsPerSched / clocks);
  else
    printf("%-34s one warp per SM: %.2f clocks per instruction (dependent-issue latency / chains)\n", name,
           clocks / perWarp);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  g_mhz = khz / 1e3;
  g_sms = p.multiProcessorCount;
  printf("%s, %d SMs, %.0f MHz (max SM clock)\n", p.name, g_sms, g_mhz);
  cudaMalloc(&g_d, 8);
  cudaMalloc(&g_di, 4);
  run<4, 0, 0>("fp64 only (DFMA x4)", 4, 256);
  run<0, 4, 0>("ALU only (IADD x4)", 4, 256);
  run<0, 4, 1>("IMAD only (x4)", 4, 256);
  run<4, 4, 0>("1 DFMA : 1 IADD", 4, 256);
  run<2, 3, 0>("2 DFMA : 3 IADD (leaf-kernel mix)", 4, 256);
  run<2, 6, 0>("1 DFMA : 3 IADD", 4, 256);
  run<4, 4, 1>("1 DFMA : 1 IMAD", 4, 256);
  run<2, 2, 1>("2 DFMA : 2 IMAD (+ILP 2)", 4, 256);
  run<1, 0, 0>("DFMA, 1 dependent chain", 1, 32);
  run<2, 0, 0>("DFMA, 2 chains", 1, 32);
  run<0, 1, 0>("IADD, 1 dependent chain", 1, 32);
  run<0, 1, 1>("IMAD, 1 dependent chain", 1, 32);
  run<1, 0, 0>("fp64 only, 1 chain x 8 warps/sched", 4, 256);
  run<2, 0, 0>("fp64 only, 2 chains x 8 warps/sched", 4, 256);
  return 0;
}
