// Micro-benchmark: FP32 FMA throughput of one B200 SM for scalar FFMA, packed FFMA2, and mixes of the two
// (does the packed form reach the full 128 FMA/clk/SM, and can scalar FMAs run beside it on the other half
// of the FMA pipe?).  8 warps per SM (2 per scheduler, like the TXP-CNN kernel) and 16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_pipes fma_pipes.cu && ./fma_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<unsigned long long&>(d))
               : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}
__device__ __forceinline__ void ffma1(float& d, const float a, const float b) {
  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d) : "f"(a), "f"(b));
}

// per loop trip: NP packed FFMA2 and NS scalar FFMA, all independent accumulators
template <int NP, int NS>
__global__ void k(float* out, int iters, float x, float y) {
  float2 ap[NP > 0 ? NP : 1];
  float as[NS > 0 ? NS : 1];
  for (int i = 0; i < NP; ++i) ap[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  for (int i = 0; i < NS; ++i) as[i] = threadIdx.x * 2e-3f + i;
  const float2 a2 = make_float2(x, y), b2 = make_float2(y, x);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < (NP > NS ? NP : NS); ++i) {
        if (i < NP) ffma2(ap[i], a2, b2);
        if (i < NS) ffma1(as[i], x, y);
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < NP; ++i) s += ap[i].x + ap[i].y;
  for (int i = 0; i < NS; ++i) s += as[i];
  if (s == 123.456f) out[0] = s;
}

template <int NP, int NS>
void run(const char* name, int warps, float* out) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<NP, NS><<<sms, warps * 32>>>(out, 100, 1.0001f, 0.9999f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<NP, NS><<<sms, warps * 32>>>(out, iters, 1.0001f, 0.9999f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int mhz = 0;
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  const double fma = double(iters) * 4 * (2.0 * NP + NS) * warps * 32 * sms;
  const double instr = double(iters) * 4 * (NP + NS) * warps * sms;
  printf("%-28s warps/SM %2d : %7.2f TFLOP/s  (%6.1f FMA/clk/SM at %d MHz nominal, %5.2f warp-instr/clk/SM)\n", name, warps,
         2 * fma / ms / 1e9, fma / (ms * 1e-3) / (mhz * 1e3) / sms, mhz / 1000, instr / (ms * 1e-3) / (mhz * 1e3) / sms);
}

int main() {
  float* out;
  cudaMalloc(&out, 16);
  for (int warps : {8, 16, 32}) {
    run<0, 24>("scalar FFMA only", warps, out);
    run<12, 0>("packed FFMA2 only", warps, out);
    run<12, 6>("FFMA2 : FFMA = 2 : 1", warps, out);
    run<12, 12>("FFMA2 : FFMA = 1 : 1", warps, out);
    run<8, 16>("FFMA2 : FFMA = 1 : 2", warps, out);
  }
  return 0;
}
