// Generic fp32 tiled GEMM with fused epilogues (CUDA-core FFMA path).
//
//   C(m,n) = epilogue( sum_k A(m,k) * B(k,n) )
//   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn],  C(m,n) = C[m*scm + n*scn]
//
// All three operands are addressed through (row, col) strides so that the forward
// (x W^T), data-gradient (dy W) and weight-gradient (dy^T x, split over rows) products of
// the head MLPs, of w1 and of the attention fc1 blocks share one kernel.  gridDim.z > 1
// splits K; split z writes its partial to C + z*split_stride (summed later in fixed order
// by reduce_splits_kernel, so results are run-to-run deterministic without atomics).
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
  long long split_stride;    // floats between split partials
};

template <int BM, int BN, int TM, int TN, int EPI>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(const GemmArgs g) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  static_assert(TM == 4 && TN == 4, "micro-tile is 4x4 (float4 shared loads)");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tn = tid % (BN / TN), tm = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_chunk;
  const int kend = min(g.K, kbeg + g.k_chunk);
  const bool a_kfast = (g.sak == 1);     // which index varies fastest across threads on load
  const bool b_nfast = (g.sbn == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    for (int i = tid; i < BM * BK; i += NT) {
      int m, k;
      if (a_kfast) { k = i % BK; m = i / BK; } else { m = i % BM; k = i / BM; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < g.M && gk < kend) ? __ldg(g.A + gm * g.sam + gk * g.sak) : 0.f;
    }
    for (int i = tid; i < BN * BK; i += NT) {
      int n, k;
      if (b_nfast) { n = i % BN; k = i / BN; } else { k = i % BK; n = i / BK; }
      const int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < g.N && gk < kend) ? __ldg(g.B + gk * g.sbk + gn * g.sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tm * TM]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tn * TN]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* C = g.C + (long long)blockIdx.z * g.split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + tm * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn * TN + j;
      if (n >= g.N) continue;
      const long long ci = m * g.scm + n * g.scn;
      float v = acc[i][j];
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

// Column sums of src[M,N] (row stride ld) split over STAT-style row chunks:
// part[blockIdx.y][n] = sum over this chunk's rows; follow with reduce_splits_kernel.
__global__ void colsum_partial_kernel(const float* __restrict__ src, long long ld, long long M, int N,
                                      int rows_per_chunk, float* __restrict__ part);

template <int EPI>
int launch_gemm(GemmArgs g, int splits, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return NRM_OK;
  if (splits <= 1) {
    splits = 1; g.k_chunk = g.K; g.split_stride = 0;
  } else {
    g.k_chunk = (g.K + splits - 1) / splits;
    g.k_chunk = (g.k_chunk + 15) / 16 * 16;
    splits = (g.K + g.k_chunk - 1) / g.k_chunk;
  }
  // 64x64 tiles when N is a multiple of 64 or large; a 64x72 tile covers the 66-wide
  // hidden layers of the head MLPs in one column tile.
  if (g.N > 64 && g.N <= 72) {
    dim3 grid(1, (g.M + 63) / 64, splits);
    gemm_kernel<64, 72, 4, 4, EPI><<<grid, 288, 0, s>>>(g);
  } else {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, splits);
    gemm_kernel<64, 64, 4, 4, EPI><<<grid, 256, 0, s>>>(g);
  }
  NRM_LAUNCH_CHECK("gemm_kernel");
  return splits;   // > 0: number of partials actually written
}

}  // namespace nrm
