// Pipe-throughput micro-benchmark for sm_100a (B200): FFMA / FFMA2 / FMUL2 / FADD2 / MUFU warp-instructions per
// cycle and SM partition, measured with independent register chains.  nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 tools/ubench_pipes.cu -o tools/build/ubench_pipes ; run on the GPU box.  Not part of the library.
#include <cstdio>
#include <cuda_runtime.h>

#define NCH 8
template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
  float a[NCH], b = seed, c = seed * 0.5f;
  unsigned long long p[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    a[i] = seed + i + threadIdx.x * 1e-3f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 0.5f));
  }
  unsigned long long pb, pc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b + 0.25f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c + 0.25f));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
        if (OP == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
        if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
        if (OP == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (OP == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (OP == 7) {  // 4 packed : 1 MUFU, the mix of the NB epilogue
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
          if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
        if (OP == 8) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        if (OP == 9) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
    s += a[i] + lo + hi;
  }
  if (s == 12345.678f) out[0] = s;
}

template <int OP>
void run(const char* name, int threads, double per_iter_instr) {
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 4096, grid = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<OP><<<grid, threads>>>(out, 64, 1.0001f);
  cudaEventRecord(e0);
  k<OP><<<grid, threads>>>(out, iters, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double warps = threads / 32.0, instr = per_iter_instr * 4 * iters * warps;  // warp-instructions per SM
  const double cycles = ms * 1e-3 * clk_khz * 1e3;
  printf("%-34s threads %4d  %.3f ms  %.3f warp-instr/clk/SM  (%.3f per SM partition; clock %d MHz assumed)\n", name, threads, ms,
         instr / cycles, instr / cycles / 4, clk_khz / 1000);
  cudaFree(out);
}

int main() {
  for (int threads : {128, 512}) {
    run<0>("FFMA (3 distinct regs)", threads, NCH);
    run<9>("FFMA (2 distinct regs)", threads, NCH);
    run<8>("FADD", threads, NCH);
    run<1>("FFMA2", threads, NCH);
    run<2>("FMUL2", threads, NCH);
    run<3>("FADD2", threads, NCH);
    run<4>("MUFU.EX2", threads, NCH);
    run<5>("MUFU.LG2", threads, NCH);
    run<6>("MUFU.RCP", threads, NCH);
    run<7>("4 FFMA2 : 1 MUFU.EX2", threads, NCH + NCH / 4);
  }
  return 0;
}
