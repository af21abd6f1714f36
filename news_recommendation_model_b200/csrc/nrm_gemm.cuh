// Generic fp32 tiled GEMM with fused epilogues (CUDA-core FFMA path).
//
//   C(m,n) = epilogue( sum_k A(m,k) * B(k,n) )
//   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn],  C(m,n) = C[m*scm + n*scn]
//
// All three operands are addressed through (row, col) strides so that the forward
// (x W^T), data-gradient (dy W) and weight-gradient (dy^T x, split over rows) products of
// the head MLPs, of w1 and of the attention blocks share one kernel.
//
// Every reduction here is short per CTA (K <= 264 for the layer products, a <= 256-row
// slice for the split weight gradients), so a CTA stages its whole operand slabs in shared
// memory 128 k-steps at a time: all global loads of a slab are in flight together and the
// FFMA loop runs uninterrupted; 2-4 CTAs per SM overlap each other's load phases.
//
// gridDim.z > 1 splits K; split z writes its partial to C + z*split_stride (summed later in
// fixed order by reduce_splits_kernel: run-to-run deterministic, no atomics).
// ones_row / ones_col append a virtual all-ones row to A (m == M) or column to B (n == N):
// the extra output row/column is the bias gradient and goes to Cb[n] / Cb[m].
#pragma once
#include "nrm_common.cuh"

namespace nrm {

enum Epi : int {
  EPI_NONE = 0,        // C = acc
  EPI_BIAS,            // C = acc + bias[n]
  EPI_BIAS_GELU2,      // C = acc + bias[n];  C2 = gelu(C)
  EPI_BIAS_MUL2,       // C = acc + bias[n];  C2 = C * aux1(m,n)
  EPI_MUL_GELUGRAD,    // C = acc * gelu'(aux1(m,n))
  EPI_DX2,             // C = acc * aux1(m,n);  C2 = acc * aux2(m,n)
};

struct GemmArgs {
  int M, N, K;
  const float* A; long long sam, sak;
  const float* B; long long sbk, sbn;
  float* C; long long scm, scn;
  float* C2;                 // second output, same strides as C
  const float* bias;         // [N]
  const float* aux1; const float* aux2; long long saux;   // aux(m,n) = aux[m*saux + n]
  int k_chunk;               // K range per split (== K when gridDim.z == 1)
  long long split_stride;    // floats between split partials (applies to C and Cb)
  int ones_row, ones_col;    // virtual ones row of A / ones column of B (bias gradients)
  float* Cb;                 // destination of the virtual row / column
};

constexpr int GEMM_BK = 128;

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int BM, int BN>
constexpr size_t gemm_smem_bytes() { return sizeof(float) * GEMM_BK * ((BM + 4) + (BN + 4)); }

template <int BM, int BN, int EPI>
__global__ void __launch_bounds__((BM / 4) * (BN / 4))
gemm_kernel(const GemmArgs g) {
  constexpr int TM = 4, TN = 4;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int LDA = BM + 4, LDB = BN + 4;
  extern __shared__ __align__(16) float gemm_smem[];
  float* As = gemm_smem;                 // [GEMM_BK][LDA]
  float* Bs = gemm_smem + GEMM_BK * LDA; // [GEMM_BK][LDB]

  const int tid = threadIdx.x;
  const int tn = tid % (BN / TN), tm = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_chunk;
  const int kend = min(g.K, kbeg + g.k_chunk);
  const int Mx = g.M + g.ones_row, Nx = g.N + g.ones_col;
  const bool a_kfast = (g.sak == 1);     // which index varies fastest across threads on load
  const bool b_nfast = (g.sbn == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += GEMM_BK) {
    const int kc = min(GEMM_BK, kend - k0);
    if (k0 != kbeg) __syncthreads();
    // Stage the slabs with 4-byte cp.async: every copy of the slab is in flight at once and
    // lands directly in its transposed (k-major) slot; one wait for the whole slab.
    // Index split uses compile-time divisors only (GEMM_BK, BM, BN); k >= kc lanes are skipped.
    for (int i = tid; i < BM * GEMM_BK; i += NT) {
      int m, k;
      if (a_kfast) { k = i % GEMM_BK; m = i / GEMM_BK; } else { m = i % BM; k = i / BM; }
      if (k >= kc) continue;
      const int gm = m0 + m;
      float* dst = As + k * LDA + m;
      if (gm < g.M) cp_async4(dst, g.A + gm * g.sam + (k0 + k) * g.sak);
      else *dst = (gm < Mx) ? 1.f : 0.f;
    }
    for (int i = tid; i < BN * GEMM_BK; i += NT) {
      int n, k;
      if (b_nfast) { n = i % BN; k = i / BN; } else { k = i % GEMM_BK; n = i / GEMM_BK; }
      if (k >= kc) continue;
      const int gn = n0 + n;
      float* dst = Bs + k * LDB + n;
      if (gn < g.N) cp_async4(dst, g.B + (k0 + k) * g.sbk + gn * g.sbn);
      else *dst = (gn < Nx) ? 1.f : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kc; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(As + kk * LDA + tm * TM);
      const float4 b = *reinterpret_cast<const float4*>(Bs + kk * LDB + tn * TN);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  const long long zoff = (long long)blockIdx.z * g.split_stride;
  float* C = g.C + zoff;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + tm * TM + i;
    if (m >= Mx) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn * TN + j;
      if (n >= Nx) continue;
      float v = acc[i][j];
      if (m >= g.M || n >= g.N) {          // bias-gradient row / column (EPI_NONE only)
        if (m >= g.M && n >= g.N) continue;
        g.Cb[zoff + (m >= g.M ? n : m)] = v;
        continue;
      }
      const long long ci = m * g.scm + n * g.scn;
      if (EPI == EPI_NONE) {
        C[ci] = v;
      } else if (EPI == EPI_BIAS) {
        C[ci] = v + __ldg(g.bias + n);
      } else if (EPI == EPI_BIAS_GELU2) {
        v += __ldg(g.bias + n);
        C[ci] = v;
        g.C2[ci] = gelu_f(v);
      } else if (EPI == EPI_BIAS_MUL2) {
        v += __ldg(g.bias + n);
        C[ci] = v;
        g.C2[ci] = v * __ldg(g.aux1 + m * g.saux + n);
      } else if (EPI == EPI_MUL_GELUGRAD) {
        C[ci] = v * gelu_grad_f(__ldg(g.aux1 + m * g.saux + n));
      } else if (EPI == EPI_DX2) {
        C[ci] = v * __ldg(g.aux1 + m * g.saux + n);
        g.C2[ci] = v * __ldg(g.aux2 + m * g.saux + n);
      }
    }
  }
}

// dst[i] = sum_{z < nsplit} src[z*stride + i], fixed order.
__global__ void reduce_splits_kernel(const float* __restrict__ src, int nsplit, long long stride,
                                     float* __restrict__ dst, long long count);

template <int BM, int BN, int EPI>
int launch_gemm_tile(const GemmArgs& g, int splits, cudaStream_t s) {
  constexpr size_t smem = gemm_smem_bytes<BM, BN>();
  static bool configured = false;
  if (!configured) {
    NRM_CUDA(cudaFuncSetAttribute(gemm_kernel<BM, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int Mx = g.M + g.ones_row, Nx = g.N + g.ones_col;
  dim3 grid((Nx + BN - 1) / BN, (Mx + BM - 1) / BM, splits);
  gemm_kernel<BM, BN, EPI><<<grid, (BM / 4) * (BN / 4), smem, s>>>(g);
  NRM_LAUNCH_CHECK("gemm_kernel");
  return NRM_OK;
}

// Returns the number of K splits actually written (>= 1) or a negative error code.
template <int EPI>
int launch_gemm(GemmArgs g, int splits, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 1;
  if (splits <= 1) {
    splits = 1; g.k_chunk = g.K; g.split_stride = 0;
  } else {
    g.k_chunk = (g.K + splits - 1) / splits;
    g.k_chunk = (g.k_chunk + 7) / 8 * 8;
    splits = (g.K + g.k_chunk - 1) / g.k_chunk;
  }
  const int Nx = g.N + g.ones_col, Mx = g.M + g.ones_row;
  // column tile: 72 covers the 66(+1)-wide hidden layers in one tile and 264(+1) in four;
  // row tile: 32 when that is what it takes to fill the machine
  const bool wide72 = (Nx > 64 && Nx <= 72) || (Nx > 128 && (Nx + 71) / 72 < (Nx + 63) / 64);
  const long long ctas64 = (long long)((Mx + 63) / 64) * ((Nx + (wide72 ? 71 : 63)) / (wide72 ? 72 : 64)) * splits;
  const bool small_rows = ctas64 < 2LL * sm_count();
  int rc;
  if (wide72) rc = small_rows ? launch_gemm_tile<32, 72, EPI>(g, splits, s) : launch_gemm_tile<64, 72, EPI>(g, splits, s);
  else rc = small_rows ? launch_gemm_tile<32, 64, EPI>(g, splits, s) : launch_gemm_tile<64, 64, EPI>(g, splits, s);
  return rc < 0 ? rc : splits;
}

}  // namespace nrm
