// Micro-benchmark: issue rate of packed fp32 (FADD2 / FMUL2) against scalar FADD / FMUL on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, ILP = 8;

template <bool MUL>
__global__ void k_scalar(float *out, float a, float b)
{
  float x[2 * ILP];
  for (int i = 0; i < 2 * ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) x[i] = MUL ? __fmul_rn(__fmul_rn(x[i], a), b) : __fadd_rn(__fadd_rn(x[i], a), b);
  }
  float s = 0;
  for (int i = 0; i < 2 * ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool MUL>
__global__ void k_packed(float *out, float a, float b)
{
  unsigned long long x[ILP], aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  for (int i = 0; i < ILP; i++) {
    float lo = threadIdx.x + 2 * i, hi = threadIdx.x + 2 * i + 1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
  }
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      unsigned long long t;
      if (MUL) {
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x[i]), "l"(aa));
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(x[i]) : "l"(t), "l"(bb));
      } else {
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x[i]), "l"(aa));
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(x[i]) : "l"(t), "l"(bb));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < ILP; i++) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  float *out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 4; mode++) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k_scalar<false><<<148 * 8, 1024>>>(out, 1.0001f, 0.9999f);
      else if (mode == 1) k_packed<false><<<148 * 8, 1024>>>(out, 1.0001f, 0.9999f);
      else if (mode == 2) k_scalar<true><<<148 * 8, 1024>>>(out, 1.0001f, 0.9999f);
      else k_packed<true><<<148 * 8, 1024>>>(out, 1.0001f, 0.9999f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    double flops = 148.0 * 8 * 1024 * (double)ITER * 2 * ILP * 2;
    const char *names[4] = {"FADD", "FADD2", "FMUL", "FMUL2"};
    printf("%s: %.3f ms, %.2f Tflop/s\n", names[mode], best, flops / best / 1e9);
  }
  float h[4];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("check %g\n", h[1]);
  return 0;
}
