// exp_pipe_rates.cu — micro-benchmark: per-SM throughput (results per clock) of the special-function and packed-fp32 instructions
// the epilogues lean on (tanh.approx, ex2.approx, rcp.approx, fma.rn.f32x2, fma.rn.f32), measured with every SM full of warps
// running independent register-only chains.  Round-2 question: is the tanh-form GELU bound by the MUFU.TANH rate?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/exp_pipe_rates profiles/exp_pipe_rates.cu && /tmp/exp_pipe_rates
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f, a4 = a0 + 0.4f, a5 = a0 + 0.5f, a6 = a0 + 0.6f, a7 = a0 + 0.7f;
  unsigned long long p0, p1, p2, p3;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (OP == 0) {
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a4)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a5));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a6)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a7));
      } else if (OP == 1) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
      } else if (OP == 2) {
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a1));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a3));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a5));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a7));
      } else if (OP == 3) {   // packed fma: 2 results per instruction
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(p1), "l"(p2)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(p2), "l"(p3));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(p3), "l"(p0)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(p0), "l"(p1));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(p1), "l"(p2)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(p2), "l"(p3));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(p3), "l"(p0)); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(p0), "l"(p1));
      } else if (OP == 4) {   // scalar fma, 3 register operands
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(a1), "f"(a2)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(a2), "f"(a3));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(a3), "f"(a4)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(a4), "f"(a5));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a4) : "f"(a5), "f"(a6)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a5) : "f"(a6), "f"(a7));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a6) : "f"(a7), "f"(a0)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a7) : "f"(a0), "f"(a1));
      } else if (OP == 5) {   // ex2 on packed halves: 2 results per instruction
        unsigned h0 = __float_as_uint(a0), h1 = __float_as_uint(a1), h2 = __float_as_uint(a2), h3 = __float_as_uint(a3);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
        a0 = __uint_as_float(h0); a1 = __uint_as_float(h1); a2 = __uint_as_float(h2); a3 = __uint_as_float(h3);
      } else if (OP == 6) {   // tanh on packed halves
        unsigned h0 = __float_as_uint(a0), h1 = __float_as_uint(a1), h2 = __float_as_uint(a2), h3 = __float_as_uint(a3);
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h1));
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h3));
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h1));
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h3));
        a0 = __uint_as_float(h0); a1 = __uint_as_float(h1); a2 = __uint_as_float(h2); a3 = __uint_as_float(h3);
      } else if (OP == 7) {   // tanh on packed bf16
        unsigned h0 = __float_as_uint(a0), h1 = __float_as_uint(a1), h2 = __float_as_uint(a2), h3 = __float_as_uint(a3);
        asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h0)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h1));
        asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h2)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h3));
        asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h0)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h1));
        asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h2)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h3));
        a0 = __uint_as_float(h0); a1 = __uint_as_float(h1); a2 = __uint_as_float(h2); a3 = __uint_as_float(h3);
      }
    }
  }
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0 ^ p1 ^ p2 ^ p3));
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + lo + hi;
}

template <int OP>
void run(const char* name, int results_per_instr, float seed) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 4 * 512);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<sms * 4, 512>>>(out, 100, seed);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP><<<sms * 4, 512>>>(out, iters, seed);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = (double)sms * 4 * 512 * iters * 32.0;     // thread-level instructions
  const double per_s = instr * results_per_instr / (ms * 1e-3);
  printf("%-22s %8.3f ms  %8.2f results/ns/SM  = %6.2f results/clk/SM at the max clock %d MHz (%s)\n", name, ms, per_s / 1e9 / sms,
         per_s / 1e3 / khz / sms, khz / 1000, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("tanh.approx.f32", 1, 0.3f);
  run<1>("ex2.approx.ftz.f32", 1, -0.5f);
  run<2>("rcp.approx.ftz.f32", 1, 1.5f);
  run<3>("fma.rn.f32x2", 2, 0.9f);
  run<4>("fma.rn.f32", 1, 0.9f);
  run<5>("ex2.approx.f16x2", 2, 0.0f);
  run<6>("tanh.approx.f16x2", 2, 0.0f);
  run<7>("tanh.approx.bf16x2", 2, 0.0f);
  return 0;
}
