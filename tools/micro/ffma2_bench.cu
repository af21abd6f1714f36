// Issue rate / latency of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a: one warp per scheduler (4 warps) and 16 warps per SM,
// N independent accumulator chains per thread.   nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, bool PACKED>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[CHAINS * 2];
  for (int i = 0; i < CHAINS * 2; ++i) a[i] = threadIdx.x * 0.001f + i;
  const float x = 1.0001f, y = 0.0001f;
  unsigned long long xx, yy;
  asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
  asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (PACKED) {
        unsigned long long v;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[2 * c]), "f"(a[2 * c + 1]));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(xx), "l"(yy));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * c]), "=f"(a[2 * c + 1]) : "l"(v));
      } else {
        a[2 * c] = fmaf(a[2 * c], x, y);
        a[2 * c + 1] = fmaf(a[2 * c + 1], x, y);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < CHAINS * 2; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS, bool PACKED>
void run(int threads, const char* name) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<CHAINS, PACKED><<<148, threads>>>(out, cyc, iters);
  k<CHAINS, PACKED><<<148, threads>>>(out, cyc, iters);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double fma_per_warp = (double)iters * CHAINS * 2;
  printf("%-8s chains=%2d threads=%4d: %9lld cycles, %.2f cycles per scalar-FMA-equivalent per warp, %.1f FMA lanes/clk/SM\n", name,
         CHAINS, threads, h, h / fma_per_warp, fma_per_warp * threads / (double)h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1, false>(128, "FFMA"); run<1, true>(128, "FFMA2");
  run<4, false>(128, "FFMA"); run<4, true>(128, "FFMA2");
  run<8, false>(128, "FFMA"); run<8, true>(128, "FFMA2");
  run<8, false>(512, "FFMA"); run<8, true>(512, "FFMA2");
  run<4, false>(1024, "FFMA"); run<4, true>(1024, "FFMA2");
  return 0;
}
