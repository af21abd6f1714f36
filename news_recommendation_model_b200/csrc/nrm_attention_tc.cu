// Pairwise MLP attention on the 5th-generation tensor cores (precision = bf16).
//
// Same reduced algebra as nrm_attention.cu (hid[c,h,:] = W_c h + tp_c, W_c = Wd diag(t_c) + A):
// per (impression, candidate) ITEM the hidden tile is one 64(h) x 64(j) x 64(k) product
//     D[h][j] = sum_k H[h][k] * W_c[j][k]
// issued as four tcgen05.mma (M=64, N=64, K=16, bf16 operands, fp32 accumulation in TMEM).
//   A operand: the staged history tile in bf16 (one per impression, shared by its C candidates)
//   B operand: W_c, generated per item from the fp32 blocks Wd, A and the candidate vector
// An M=64 accumulator occupies only the lower 16 lanes of each 32-lane TMEM sub-partition, so
// two items are run as a PAIR: the second accumulator lives in the upper 16 lanes of the same
// columns, and in the epilogue every lane of the four warps owns one (item, history row):
// it reads its 64 hidden pre-activations with tcgen05.ld, applies +tp, GELU and the fc2 dot
// product entirely in registers (no cross-lane traffic) and emits the attention score.
// A CTA iteration handles TWO impressions (2C items, always an even count).
#include "nrm_kernels.cuh"
#include "nrm_umma.cuh"

namespace nrm {

constexpr int TC_THREADS = 128;
constexpr int TC_MAXC = 16;          // items per impression handled per chunk

// derived weights per branch in the workspace (att_prep_kernel): Wd | A | BmT | b1 | w2 | b2
constexpr int DER_WD = 0, DER_A = 4096, DER_BMT = 8192, DER_B1 = 12288, DER_W2 = 12352, DER_B2 = 12416, DER_SIZE = 12420;

__global__ void __launch_bounds__(256)
att_prep_kernel(const float* __restrict__ P, float* __restrict__ der) {
  const AttOffsets off = blockIdx.y == 0 ? ATT_LABEL : ATT_TI;
  float* d = der + (long long)blockIdx.y * DER_SIZE;
  const float* W = P + off.fc1_w;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < 4096; i += gridDim.x * 256) {
    const int j = i >> 6, k = i & 63;
    const float wa = W[j * 256 + k], wb = W[j * 256 + 64 + k], wc = W[j * 256 + 128 + k], wd = W[j * 256 + 192 + k];
    const int blk = (k >> 3) * 512 + j * 8 + (k & 7);      // [k/8][j][k%8]: coalesced for lane = j readers
    d[DER_WD + blk] = wd;
    d[DER_A + blk] = wa - wc;
    d[DER_BMT + k * 64 + j] = wb + wc;
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {
    d[DER_B1 + threadIdx.x] = P[off.fc1_b + threadIdx.x];
    d[DER_W2 + threadIdx.x] = P[off.fc2_w + threadIdx.x];
    if (threadIdx.x == 0) d[DER_B2] = P[off.fc2_b];
  }
}

struct TcSmemFwd {
  float w2[64];
  float t[2 * TC_MAXC * 64];               // candidate vectors of the 2 impressions' items
  float tp[2 * TC_MAXC * 64];              // (Wb + Wc) t + b1 per item
  float s[2 * TC_MAXC * 64];               // scores per item and history row
  __align__(128) unsigned char opA[2][umma::TILE64_BYTES];   // bf16 history tiles (one per impression)
  __align__(128) unsigned char opB[2][umma::TILE64_BYTES];   // bf16 W_c of the two items of a pair
  uint64_t mbar;
  uint32_t tmem_base;
};

// history rows [r0, r0+64) of impression b -> bf16 canonical tile; rows >= H are zero
template <int BRANCH>
__device__ __forceinline__ void stage_history_bf16(const double* __restrict__ xh, const float* __restrict__ xhp,
                                                   long long b, int H, int r0, unsigned char* tile) {
  for (int it = threadIdx.x; it < 64 * 8; it += TC_THREADS) {
    const int row = it & 63, kb = it >> 6;
    float v[8];
    if (r0 + row < H) {
      if (BRANCH == 0) {
        const float4* src = reinterpret_cast<const float4*>(xhp + (b * H + r0 + row) * 64 + kb * 8);
        const float4 a = __ldg(src), c = __ldg(src + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
      } else {
        const double2* src = reinterpret_cast<const double2*>(xh + (b * H + r0 + row) * HC + 4 + kb * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double2 d = __ldg(src + i); v[2 * i] = (float)d.x; v[2 * i + 1] = (float)d.y; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    umma::store_bf16x8(tile + umma::tile64_offset(row, kb), v);
  }
}

// fp32 history value (for the pooling sum): label branch from xh, text/img from the packed rows
template <int BRANCH>
__device__ __forceinline__ float history_value(const double* __restrict__ xh, const float* __restrict__ xhp, long long row, int k) {
  return BRANCH == 0 ? __ldg(xhp + row * 64 + k) : (float)__ldg(xh + row * HC + 4 + k);
}

// W_c[j][k] = Wd[j][k] * t[k] + A[j][k] -> bf16 canonical tile (rows = j).  Wd / A come from the
// derived-weight buffer in [k/8][j][8] order (32 KB, L1-resident, fully coalesced for lane = j).
__device__ __forceinline__ void build_Wc_bf16(const float* __restrict__ der, const float* t, unsigned char* tile) {
  for (int it = threadIdx.x; it < 64 * 8; it += TC_THREADS) {
    const int j = it & 63, kb = it >> 6;
    const float4* wd = reinterpret_cast<const float4*>(der + DER_WD + kb * 512 + j * 8);
    const float4* wa = reinterpret_cast<const float4*>(der + DER_A + kb * 512 + j * 8);
    const float4 d0 = __ldg(wd), d1 = __ldg(wd + 1), a0 = __ldg(wa), a1 = __ldg(wa + 1);
    const float4 t0 = *reinterpret_cast<const float4*>(t + kb * 8), t1 = *reinterpret_cast<const float4*>(t + kb * 8 + 4);
    float v[8];
    v[0] = fmaf(d0.x, t0.x, a0.x); v[1] = fmaf(d0.y, t0.y, a0.y); v[2] = fmaf(d0.z, t0.z, a0.z); v[3] = fmaf(d0.w, t0.w, a0.w);
    v[4] = fmaf(d1.x, t1.x, a1.x); v[5] = fmaf(d1.y, t1.y, a1.y); v[6] = fmaf(d1.z, t1.z, a1.z); v[7] = fmaf(d1.w, t1.w, a1.w);
    umma::store_bf16x8(tile + umma::tile64_offset(j, kb), v);
  }
}

template <int BRANCH>
__global__ void __launch_bounds__(TC_THREADS)
attention_forward_tc_kernel(const double* __restrict__ xh, const float* __restrict__ xhp, int B, int H, int C,
                            const float* __restrict__ der, float* __restrict__ e) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmemFwd& sm = *reinterpret_cast<TcSmemFwd*>(smem_raw);
  constexpr int TOFF = BRANCH == 0 ? E_XT : E_PCAT;
  constexpr int POFF = BRANCH == 0 ? E_LAB : E_TI;
  constexpr uint32_t IDESC = umma::make_idesc_bf16(64, 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* d = der + (long long)BRANCH * DER_SIZE;

  if (tid < 64) sm.w2[tid] = __ldg(d + DER_W2 + tid);
  const float b2 = __ldg(d + DER_B2);
  if (warp == 0) umma::tmem_alloc(&sm.tmem_base, 64);
  if (tid == 0) umma::mbar_init(&sm.mbar, 1);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  uint32_t phase = 0;

  // epilogue role of this thread: sub-partition = warp, lower / upper half-lanes = item 0 / 1 of the pair
  const int half = lane >> 4, row = warp * 16 + (lane & 15);
  const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);

  const int npairs_b = (B + 1) / 2;
  for (int pb = blockIdx.x; pb < npairs_b; pb += gridDim.x) {
    const long long b0 = 2LL * pb;
    const int nimp = (b0 + 1 < B) ? 2 : 1;
    for (int r0 = 0; r0 < H; r0 += 64) {
      for (int c0 = 0; c0 < C; c0 += TC_MAXC) {
        const int nc = min(TC_MAXC, C - c0);
        const int nitems = nimp * nc;                 // item = imp * nc + cl
        __syncthreads();                              // previous chunk fully consumed
        if (c0 == 0)
          for (int imp = 0; imp < nimp; ++imp) stage_history_bf16<BRANCH>(xh, xhp, b0 + imp, H, r0, sm.opA[imp]);
        for (int i = tid; i < nitems * 64; i += TC_THREADS) {
          const int item = i >> 6, k = i & 63;
          const long long rc = (b0 + item / nc) * C + c0 + item % nc;
          sm.t[i] = e[rc * E + TOFF + k];
        }
        __syncthreads();
        for (int i = tid; i < nitems * 64; i += TC_THREADS) {
          const int item = i >> 6, j = i & 63;
          float v = __ldg(d + DER_B1 + j);
#pragma unroll 8
          for (int k = 0; k < 64; ++k) v = fmaf(__ldg(d + DER_BMT + k * 64 + j), sm.t[item * 64 + k], v);
          sm.tp[i] = v;
        }
        for (int p0 = 0; p0 < nitems; p0 += 2) {
          const int np = min(2, nitems - p0);
          for (int q = 0; q < np; ++q) build_Wc_bf16(d, sm.t + (p0 + q) * 64, sm.opB[q]);
          umma::fence_async_smem();
          umma::fence_before_sync();
          __syncthreads();
          if (tid == 0) {
            umma::fence_after_sync();
            for (int q = 0; q < np; ++q) {
              const int imp = (p0 + q) / nc;
              umma::mma_tile64(tmem + ((uint32_t)(16 * q) << 16), umma::smem_u32(sm.opA[imp]), umma::smem_u32(sm.opB[q]),
                               IDESC, false);
            }
            umma::mma_commit(&sm.mbar);
          }
          umma::mbar_wait(&sm.mbar, phase);
          phase ^= 1;
          umma::fence_after_sync();
          {
            const int item = p0 + half;
            float acc = 0.f;
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
              float v[32];
              umma::tmem_ld32(my_tmem + cb * 32, v);          // all lanes take part (.sync.aligned)
              if (half < np) {
                const float* tp = sm.tp + item * 64 + cb * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) acc = fmaf(gelu_f(v[j] + tp[j]), sm.w2[cb * 32 + j], acc);
              }
            }
            if (half < np) sm.s[item * 64 + row] = acc + b2;
          }
          umma::fence_before_sync();
          __syncthreads();                                    // TMEM and opB free for the next pair
        }
        // pooled[item][k] (+)= sum_row s[item][row] * h[row][k]   (fp32 history re-read through L1/L2)
        for (int i = tid; i < nitems * 64; i += TC_THREADS) {
          const int item = i >> 6, k = i & 63;
          const long long b = b0 + item / nc;
          const int rows = min(64, H - r0);
          float v = 0.f;
          for (int r = 0; r < rows; ++r)
            v = fmaf(sm.s[item * 64 + r], history_value<BRANCH>(xh, xhp, b * H + r0 + r, k), v);
          float* dst = e + (b * C + c0 + item % nc) * E + POFF + k;
          if (r0 == 0) *dst = v; else *dst += v;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------------
// Self test of the tensor-core building blocks (tests/test_gpu_umma.py): two 64x64x64
// products with the interleaved half-lane accumulators; out[q][row][col] fp32.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS)
umma_selftest_kernel(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ b0,
                     const float* __restrict__ b1, float* __restrict__ out) {
  __shared__ __align__(128) unsigned char opA[2][umma::TILE64_BYTES];
  __shared__ __align__(128) unsigned char opB[2][umma::TILE64_BYTES];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* srcs[4] = {a0, a1, b0, b1};
  for (int m = 0; m < 4; ++m) {
    unsigned char* tile = m < 2 ? opA[m] : opB[m - 2];
    for (int it = tid; it < 64 * 8; it += TC_THREADS) {
      const int r = it & 63, kb = it >> 6;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = srcs[m][r * 64 + kb * 8 + i];
      umma::store_bf16x8(tile + umma::tile64_offset(r, kb), v);
    }
  }
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 64);
  if (tid == 0) umma::mbar_init(&mbar, 1);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    constexpr uint32_t IDESC = umma::make_idesc_bf16(64, 64);
    umma::mma_tile64(tmem, umma::smem_u32(opA[0]), umma::smem_u32(opB[0]), IDESC, false);
    umma::mma_tile64(tmem + (16u << 16), umma::smem_u32(opA[1]), umma::smem_u32(opB[1]), IDESC, false);
    umma::mma_commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();
  const int half = lane >> 4, row = warp * 16 + (lane & 15);
#pragma unroll
  for (int cb = 0; cb < 2; ++cb) {
    float v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(half * 64 + row) * 64 + cb * 32 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
int launch_attention_prep(const float* P, Workspace& w, cudaStream_t s) {
  att_prep_kernel<<<dim3(4, 2), 256, 0, s>>>(P, w.att_derived);
  NRM_LAUNCH_CHECK("att_prep_kernel");
  return NRM_OK;
}

int launch_attention_forward_tc(const BatchPtrs& in, Workspace& w, int branch, cudaStream_t s) {
  const size_t smem = sizeof(TcSmemFwd);
  const int grid = min((w.B + 1) / 2, 4 * sm_count());
  if (branch == 0) {
    NRM_CUDA(cudaFuncSetAttribute(attention_forward_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attention_forward_tc_kernel<0><<<grid, TC_THREADS, smem, s>>>(in.xh, w.xh, w.B, w.H, w.C, w.att_derived, w.e);
  } else {
    NRM_CUDA(cudaFuncSetAttribute(attention_forward_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attention_forward_tc_kernel<1><<<grid, TC_THREADS, smem, s>>>(in.xh, w.xh, w.B, w.H, w.C, w.att_derived, w.e);
  }
  NRM_LAUNCH_CHECK("attention_forward_tc_kernel");
  return NRM_OK;
}

}  // namespace nrm

using namespace nrm;

// a0, a1, b0, b1: [64,64] fp32 (device); out: [2,64,64] fp32 = bf16(a_q) bf16(b_q)^T accumulated in fp32.
extern "C" int nrm_debug_umma_selftest(const float* a0, const float* a1, const float* b0, const float* b1, float* out,
                                       void* stream) {
  if (!a0 || !a1 || !b0 || !b1 || !out) { set_error("nrm_debug_umma_selftest: null pointer"); return NRM_EINVAL; }
  umma_selftest_kernel<<<1, TC_THREADS, 0, (cudaStream_t)stream>>>(a0, a1, b0, b1, out);
  NRM_LAUNCH_CHECK("umma_selftest_kernel");
  return NRM_OK;
}
