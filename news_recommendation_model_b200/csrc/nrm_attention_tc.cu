// Pairwise MLP attention + sum pooling on the 5th-generation tensor cores, forward AND backward.
// Reference: PointwiseAttentionExpanded.forward (models/attention_model.py:52-97) and the pooling at
// models/user_invariant_interest_model.py:83-87.
//
// Reduced algebra (nrm_attention.cu): per ITEM = (impression b, candidate c)
//     hid[h][j] = sum_k H[h][k] W_c[j][k] + tp[j],   W_c = Wd diag(t_c) + A,   tp = Bm t_c + b1
//     s[h] = w2 . gelu(hid[h][:]) + b2,               pooled[k] = sum_h s[h] H[h][k]
// and, given dP = dL/dpooled,
//     ds[h] = sum_k dP[k] H[h][k];     dhid[h][j] = ds[h] w2[j] gelu'(hid[h][j])
//     S^T[k][j] = sum_h H[h][k] dhid[h][j]  (= dL/dW_c^T);   dA += S;  dWd[j][k] += S[j][k] t[k]
//     dt[k] = sum_j S[j][k] Wd[j][k] (+ the tp path, nrm_attention_post);   Gt[j] = sum_h dhid[h][j] = dL/dtp[j]
//     dH[h][k] += sum_j dhid[h][j] W_c[j][k] + s[h] dP[k]
//
// Every contraction above is a tcgen05.mma (M = 64, bf16 operands in shared memory, fp32 accumulators in
// tensor memory); the CUDA cores only generate operands and run the per-row epilogues:
//
//   product            M   N   K    A operand                     B operand
//   hid    = H W_c^T   h   j   k    H tile, K-major               W_c tile, K-major      (per item)
//   pooled = H^T s     k   c   h    H tile read MN-major          score tile [c][h]      (per impression, forward)
//   ds     = H dP^T    h   c   k    H tile, K-major               dP tile [c][k]         (per impression)
//   S^T    = H^T dhid  k   j   h    H tile read MN-major          dhid tile read MN-major
//   dA^T  += H^T dhid  (same product again, accumulated over every item of the CTA; never read until the end)
//   dH    += dhid W_c  h   k   j    dhid tile, K-major            W_c tile read MN-major
//   Gt     = dhid^T 1  j   8   h    dhid tile read MN-major       constant ones tile
//
// "read MN-major" = the same K-major tile consumed through a descriptor with the major bit set
// (nrm_umma.cuh), so no transposed copy is ever written.
//
// An M = 64 accumulator fills only lanes 0-15 of each 32-lane TMEM sub-partition, so the CTA (128
// threads) always works on a PAIR of impressions: impression 0's accumulators live in the lower
// half-lanes, impression 1's in the upper half-lanes of the same columns, and thread (warp w, lane l)
// owns row 16 w + (l & 15) of impression l >> 4 in every epilogue: it reads its 64 accumulator columns
// with tcgen05.ld and does +tp, GELU, the fc2 dot product, dhid, ... entirely in registers.
//
// SPLIT = 1: plain bf16 operands ("bf16").  SPLIT = 3: every operand is written as hi + lo bf16 parts and every
// product is issued as A_hi B_hi + A_hi B_lo + A_lo B_hi ("bf16x3", relative operand error 2^-16; this is
// torch's float32 matmul precision "high").  Accumulation is fp32 in both.
#include "nrm_kernels.cuh"
#include "nrm_umma.cuh"

namespace nrm {

constexpr int TC_THREADS = 128;
constexpr int TC_MAXC = 8;           // candidates per chunk (N of the ds / pooling products)

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

// ---- small operand tiles -------------------------------------------------------------------------
// [8 rows][64 k] K-major: byte offset of (r, 8*kb) = kb*128 + r*16   (LBO = 128, one 8-row group)
constexpr uint32_t T8_LBO = 128, T8_SBO = 128, T8_BYTES = 1024;
__device__ __forceinline__ umma::Operand op_tile8_k(uint32_t addr) { return umma::Operand{addr, T8_LBO, T8_SBO, 2 * T8_LBO, T8_BYTES}; }

template <int NP> struct TileBytes { static constexpr uint32_t T64 = NP * umma::TILE64_BYTES, T8 = NP * T8_BYTES; };

// history rows [r0, r0+64) of impression b -> canonical K-major tile(s); rows >= H are zero
template <int BRANCH, int NP>
__device__ __forceinline__ void stage_history(const double* __restrict__ xh, const float* __restrict__ xhp,
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
    umma::store_operand8<NP>(tile, umma::tile64_offset(row, kb), umma::TILE64_BYTES, v);
  }
}

// W_c[j][k] = Wd[j][k] * t[k] + A[j][k] -> K-major tile(s) (rows = j).  Wd / A come from the derived-weight
// buffer in [k/8][j][8] order (32 KB per branch, L1-resident, coalesced for lane = j); t is the candidate vector.
template <int NP>
__device__ __forceinline__ void build_Wc(const float* __restrict__ der, const float* __restrict__ t, unsigned char* tile) {
  for (int it = threadIdx.x; it < 64 * 8; it += TC_THREADS) {
    const int j = it & 63, kb = it >> 6;
    const float4* wd = reinterpret_cast<const float4*>(der + DER_WD + kb * 512 + j * 8);
    const float4* wa = reinterpret_cast<const float4*>(der + DER_A + kb * 512 + j * 8);
    const float4 d0 = __ldg(wd), d1 = __ldg(wd + 1), a0 = __ldg(wa), a1 = __ldg(wa + 1);
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(t + kb * 8)), t1 = __ldg(reinterpret_cast<const float4*>(t + kb * 8 + 4));
    float v[8];
    v[0] = fmaf(d0.x, t0.x, a0.x); v[1] = fmaf(d0.y, t0.y, a0.y); v[2] = fmaf(d0.z, t0.z, a0.z); v[3] = fmaf(d0.w, t0.w, a0.w);
    v[4] = fmaf(d1.x, t1.x, a1.x); v[5] = fmaf(d1.y, t1.y, a1.y); v[6] = fmaf(d1.z, t1.z, a1.z); v[7] = fmaf(d1.w, t1.w, a1.w);
    umma::store_operand8<NP>(tile, umma::tile64_offset(j, kb), umma::TILE64_BYTES, v);
  }
}

// tp[q][j] = b1[j] + sum_k Bm[j][k] t_q[k] for the (up to) two items of a pair; thread = (q, j)
__device__ __forceinline__ void compute_tp(const float* __restrict__ der, const float* __restrict__ t0,
                                           const float* __restrict__ t1, int nimp, float* tp) {
  const int q = threadIdx.x >> 6, j = threadIdx.x & 63;
  if (q >= nimp) return;
  const float* t = q == 0 ? t0 : t1;
  float v = __ldg(der + DER_B1 + j);
#pragma unroll 8
  for (int k4 = 0; k4 < 16; ++k4) {
    const float4 tv = __ldg(reinterpret_cast<const float4*>(t) + k4);
    v = fmaf(__ldg(der + DER_BMT + (4 * k4 + 0) * 64 + j), tv.x, v);
    v = fmaf(__ldg(der + DER_BMT + (4 * k4 + 1) * 64 + j), tv.y, v);
    v = fmaf(__ldg(der + DER_BMT + (4 * k4 + 2) * 64 + j), tv.z, v);
    v = fmaf(__ldg(der + DER_BMT + (4 * k4 + 3) * 64 + j), tv.w, v);
  }
  tp[q * 64 + j] = v;
}

__device__ __forceinline__ void st_bf16(unsigned char* p, float v) {
  *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
}
// one element of an operand tile (hi, and lo when NP == 2)
template <int NP>
__device__ __forceinline__ void store_operand1(unsigned char* tile, uint32_t off, uint32_t part_bytes, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(tile + off) = h;
  if (NP == 2) *reinterpret_cast<__nv_bfloat16*>(tile + part_bytes + off) = __float2bfloat16_rn(v - __bfloat162float(h));
}

// =====================================================================================================
// forward
// =====================================================================================================
template <int NP>
struct TcSmemFwd {
  __align__(128) unsigned char opA[2][TileBytes<NP>::T64];   // history tiles, one per impression of the pair
  __align__(128) unsigned char opB[2][TileBytes<NP>::T64];   // W_c of the two items of a candidate pair
  __align__(128) unsigned char opS[2][TileBytes<NP>::T8];    // scores [c][h] per impression (B operand of the pooling product)
  float tp[2 * 64];
  float w2[64];
  uint64_t mbar;
  uint32_t tmem_base;
};

// TMEM columns: [0,64) hid of the current pair, [64,72) pooled^T (N = 8)
constexpr uint32_t FWD_TMEM_COLS = 128, FWD_COL_HID = 0, FWD_COL_POOL = 64;

template <int BRANCH, int SPLIT>
__global__ void __launch_bounds__(TC_THREADS, SPLIT == 3 ? 3 : 4)
attention_forward_tc_kernel(const double* __restrict__ xh, const float* __restrict__ xhp, int B, int H, int C,
                            const float* __restrict__ der_all, float* __restrict__ e) {
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmemFwd<NP>& sm = *reinterpret_cast<TcSmemFwd<NP>*>(smem_raw);
  constexpr int TOFF = BRANCH == 0 ? E_XT : E_PCAT;
  constexpr int POFF = BRANCH == 0 ? E_LAB : E_TI;
  constexpr uint32_t IDESC_HID = umma::make_idesc_bf16(64, 64);
  constexpr uint32_t IDESC_POOL = umma::make_idesc_bf16(64, 8, true, false);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* der = der_all + (long long)BRANCH * DER_SIZE;

  if (tid < 64) sm.w2[tid] = __ldg(der + DER_W2 + tid);
  const float b2 = __ldg(der + DER_B2);
  if (warp == 0) umma::tmem_alloc(&sm.tmem_base, FWD_TMEM_COLS);
  if (tid == 0) umma::mbar_init(&sm.mbar, 1);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  uint32_t phase = 0;

  // epilogue role: sub-partition = warp, half-lanes = impression 0 / 1 of the pair
  const int half = lane >> 4, row = warp * 16 + (lane & 15);
  const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);

  const int npairs_b = (B + 1) / 2;
  for (int pb = blockIdx.x; pb < npairs_b; pb += gridDim.x) {
    const long long b0 = 2LL * pb;
    const int nimp = (b0 + 1 < B) ? 2 : 1;
    for (int r0 = 0; r0 < H; r0 += 64) {
      for (int c0 = 0; c0 < C; c0 += TC_MAXC) {
        const int nc = min(TC_MAXC, C - c0);
        if (c0 == 0) {
          __syncthreads();                              // previous tile's products have completed (waited below)
          for (int imp = 0; imp < nimp; ++imp) stage_history<BRANCH, NP>(xh, xhp, b0 + imp, H, r0, sm.opA[imp]);
        }
        for (int c = 0; c < nc; ++c) {
          const float* t0 = e + ((b0 * C + c0 + c) * E) + TOFF;
          const float* t1 = e + (((b0 + 1) * C + c0 + c) * E) + TOFF;
          compute_tp(der, t0, t1, nimp, sm.tp);
          build_Wc<NP>(der, t0, sm.opB[0]);
          if (nimp == 2) build_Wc<NP>(der, t1, sm.opB[1]);
          umma::fence_async_smem();
          umma::fence_before_sync();
          __syncthreads();
          if (tid == 0) {
            umma::fence_after_sync();
            for (int q = 0; q < nimp; ++q)
              umma::mma_product<SPLIT, 4>(tmem + FWD_COL_HID + ((uint32_t)(16 * q) << 16), umma::op_tile64_k(umma::smem_u32(sm.opA[q])),
                                          umma::op_tile64_k(umma::smem_u32(sm.opB[q])), IDESC_HID, false);
            umma::mma_commit(&sm.mbar);
          }
          umma::mbar_wait(&sm.mbar, phase);
          phase ^= 1;
          umma::fence_after_sync();
          {
            float acc = 0.f;
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
              float v[32];
              umma::tmem_ld32(my_tmem + FWD_COL_HID + cb * 32, v);       // all lanes take part (.sync.aligned)
              if (half < nimp) {
                const float* tp = sm.tp + half * 64 + cb * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) acc = fmaf(gelu_f(v[j] + tp[j]), sm.w2[cb * 32 + j], acc);
              }
            }
            // score of (impression half, candidate c, history row) -> B operand of the pooling product, [c][h] K-major
            if (half < nimp)
              store_operand1<NP>(sm.opS[half], (uint32_t)(row >> 3) * T8_LBO + (uint32_t)c * 16 + (uint32_t)(row & 7) * 2, T8_BYTES, acc + b2);
          }
          umma::fence_async_smem();
          umma::fence_before_sync();
          __syncthreads();                                    // TMEM hid, opB and tp free for the next pair; opS visible
        }
        // pooled^T[k][c] = sum_h H[h][k] s[c][h]  for both impressions (columns c >= nc are never read)
        if (tid == 0) {
          umma::fence_after_sync();
          for (int q = 0; q < nimp; ++q)
            umma::mma_product<SPLIT, 4>(tmem + FWD_COL_POOL + ((uint32_t)(16 * q) << 16), umma::op_tile64_mn(umma::smem_u32(sm.opA[q])),
                                        op_tile8_k(umma::smem_u32(sm.opS[q])), IDESC_POOL, false);
          umma::mma_commit(&sm.mbar);
        }
        umma::mbar_wait(&sm.mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
        {
          float v[8];
          umma::tmem_ld8(my_tmem + FWD_COL_POOL, v);
          if (half < nimp) {
            float* dst = e + ((b0 + half) * C + c0) * E + POFF + row;     // here `row` is the feature index k
#pragma unroll
            for (int c = 0; c < TC_MAXC; ++c)
              if (c < nc) { if (r0 == 0) dst[(long long)c * E] = v[c]; else dst[(long long)c * E] += v[c]; }
          }
        }
        umma::fence_before_sync();
      }
    }
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, FWD_TMEM_COLS);
}

// =====================================================================================================
// backward
// =====================================================================================================
template <int NP>
struct TcSmemBwd {
  __align__(128) unsigned char opA[2][TileBytes<NP>::T64];   // history tiles [h][k]
  __align__(128) unsigned char opBD[2][2][TileBytes<NP>::T64];   // [0][q]: W_c [j][k] of item q;  [1][q]: dhid [h][j] of item q
  __align__(128) unsigned char opP[2][TileBytes<NP>::T8];    // dP [c][k] per impression (B operand of the ds product)
  __align__(128) unsigned char ones[T8_BYTES];               // [8][64] ones (B operand of the Gt product)
  float ds[2][TC_MAXC][64];                                  // [impression][candidate][history row]
  float sc[2][TC_MAXC][64];                                  // attention scores, same indexing (label branch)
  float tp[2 * 64];
  float w2[64];
  uint64_t mbar;
  uint32_t tmem_base;
};

// TMEM columns: [0,64) hid, then Gt in [0,8);  [64,128) ds in [64,72), then S^T;  [128,192) dH;  [192,256) dA^T
constexpr uint32_t BWD_TMEM_COLS = 256, BWD_COL_HID = 0, BWD_COL_S = 64, BWD_COL_DH = 128, BWD_COL_DA = 192;
// per-CTA partial sums (floats): dA^T [2 halves][64 k][64 j] | dWd^T [2][64][64] | dw2 [64] | db2 [1] (+3 pad)
constexpr int TCP_DA = 0, TCP_DWD = 2 * 4096, TCP_DW2 = 4 * 4096, TCP_DB2 = 4 * 4096 + 64, TC_PARTIAL = ATT_TC_PARTIAL;

template <int BRANCH, int SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 2)
attention_backward_tc_kernel(const double* __restrict__ xh, const float* __restrict__ xhp, int B, int H, int C,
                             const float* __restrict__ der_all, const float* __restrict__ P, const float* __restrict__ e,
                             const float* __restrict__ de, float* __restrict__ dxh, float* __restrict__ dxt,
                             float* __restrict__ dtp, float* __restrict__ part) {
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  constexpr bool INPUT_GRADS = (BRANCH == 0);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmemBwd<NP>& sm = *reinterpret_cast<TcSmemBwd<NP>*>(smem_raw);
  constexpr AttOffsets off = BRANCH == 0 ? ATT_LABEL : ATT_TI;
  constexpr int TOFF = BRANCH == 0 ? E_XT : E_PCAT;
  constexpr int POFF = BRANCH == 0 ? E_LAB : E_TI;
  constexpr uint32_t IDESC_HID = umma::make_idesc_bf16(64, 64);               // H W_c^T
  constexpr uint32_t IDESC_DS = umma::make_idesc_bf16(64, 8);                 // H dP^T
  constexpr uint32_t IDESC_ST = umma::make_idesc_bf16(64, 64, true, true);    // H^T dhid
  constexpr uint32_t IDESC_DH = umma::make_idesc_bf16(64, 64, false, true);   // dhid W_c
  constexpr uint32_t IDESC_GT = umma::make_idesc_bf16(64, 8, true, false);    // dhid^T ones
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* der = der_all + (long long)BRANCH * DER_SIZE;
  const float* Wd_rm = P + off.fc1_w + 192;            // Wd[j][k] = fc1.weight[j][192 + k]

  if (tid < 64) sm.w2[tid] = __ldg(der + DER_W2 + tid);
  for (int i = tid; i < (int)T8_BYTES / 2; i += TC_THREADS) reinterpret_cast<__nv_bfloat16*>(sm.ones)[i] = __float2bfloat16_rn(1.0f);
  const float b2 = __ldg(der + DER_B2);
  if (warp == 0) umma::tmem_alloc(&sm.tmem_base, BWD_TMEM_COLS);
  if (tid == 0) umma::mbar_init(&sm.mbar, 1);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  uint32_t phase = 0;

  const int half = lane >> 4, row = warp * 16 + (lane & 15);
  const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t half_off[2] = {0u, 16u << 16};

  // persistent per-thread accumulators
  float dw2_acc[64];       // thread (half, history row): sum over items of ds * gelu(hid[row][j])
  float dwd_acc[64];       // thread (half, k):           sum over items of t[k] * S^T[k][j]
#pragma unroll
  for (int j = 0; j < 64; ++j) { dw2_acc[j] = 0.f; dwd_acc[j] = 0.f; }
  float db2_acc = 0.f;
  bool da_started[2] = {false, false};

  const int npairs_b = (B + 1) / 2;
  for (int pb = blockIdx.x; pb < npairs_b; pb += gridDim.x) {
    const long long b0 = 2LL * pb;
    const int nimp = (b0 + 1 < B) ? 2 : 1;
    const long long bmine = b0 + half;                 // this thread's impression (valid when half < nimp)
    for (int r0 = 0; r0 < H; r0 += 64) {
      __syncthreads();                                 // every product of the previous tile has completed
      for (int imp = 0; imp < nimp; ++imp) stage_history<BRANCH, NP>(xh, xhp, b0 + imp, H, r0, sm.opA[imp]);
      bool dh_started = false;
      for (int c0 = 0; c0 < C; c0 += TC_MAXC) {
        const int nc = min(TC_MAXC, C - c0);
        // ---- ds[c][h] = sum_k dP_c[k] H[h][k] for the chunk's candidates of both impressions
        for (int i = tid; i < nimp * TC_MAXC * 8; i += TC_THREADS) {
          const int imp = i / (TC_MAXC * 8), c = (i / 8) % TC_MAXC, kb = i & 7;
          float v[8];
          if (c < nc) {
            const float4* src = reinterpret_cast<const float4*>(de + ((b0 + imp) * C + c0 + c) * E + POFF + kb * 8);
            const float4 a = __ldg(src), bq = __ldg(src + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bq.x; v[5] = bq.y; v[6] = bq.z; v[7] = bq.w;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = 0.f;
          }
          umma::store_operand8<NP>(sm.opP[imp], (uint32_t)kb * T8_LBO + (uint32_t)c * 16, T8_BYTES, v);
        }
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
          umma::fence_after_sync();
          for (int q = 0; q < nimp; ++q)
            umma::mma_product<SPLIT, 4>(tmem + BWD_COL_S + half_off[q], umma::op_tile64_k(umma::smem_u32(sm.opA[q])),
                                        op_tile8_k(umma::smem_u32(sm.opP[q])), IDESC_DS, false);
          umma::mma_commit(&sm.mbar);
        }
        umma::mbar_wait(&sm.mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
        {
          float v[8];
          umma::tmem_ld8(my_tmem + BWD_COL_S, v);
          if (half < nimp) {
#pragma unroll
            for (int c = 0; c < TC_MAXC; ++c) sm.ds[half][c][row] = v[c];
          }
        }
        umma::fence_before_sync();

        for (int c = 0; c < nc; ++c) {
          const long long rc0 = b0 * C + c0 + c, rc1 = (b0 + 1) * C + c0 + c;
          const long long rcm = bmine * C + c0 + c;
          const float* t0 = e + rc0 * E + TOFF;
          const float* t1 = e + rc1 * E + TOFF;
          compute_tp(der, t0, t1, nimp, sm.tp);
          build_Wc<NP>(der, t0, sm.opBD[0][0]);
          if (nimp == 2) build_Wc<NP>(der, t1, sm.opBD[0][1]);
          umma::fence_async_smem();
          umma::fence_before_sync();
          __syncthreads();                               // (1) operands visible; previous S^T / Gt reads done
          if (tid == 0) {
            umma::fence_after_sync();
            for (int q = 0; q < nimp; ++q)
              umma::mma_product<SPLIT, 4>(tmem + BWD_COL_HID + half_off[q], umma::op_tile64_k(umma::smem_u32(sm.opA[q])),
                                          umma::op_tile64_k(umma::smem_u32(sm.opBD[0][q])), IDESC_HID, false);
            umma::mma_commit(&sm.mbar);
          }
          umma::mbar_wait(&sm.mbar, phase);
          phase ^= 1;
          umma::fence_after_sync();
          // ---- epilogue 1: thread = (impression half, history row)
          {
            const bool act = half < nimp;
            const float dsr = act ? sm.ds[half][c][row] : 0.f;
            float sacc = 0.f;
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
              float v[32];
              umma::tmem_ld32(my_tmem + BWD_COL_HID + cb * 32, v);
              if (act) {
                const float* tp = sm.tp + half * 64 + cb * 32;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                  float dh[8];
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) {
                    const int j = j8 * 8 + jj;
                    float gp;
                    const float g = gelu_both(v[j] + tp[j], gp);
                    const float w = sm.w2[cb * 32 + j];
                    if (INPUT_GRADS) sacc = fmaf(g, w, sacc);
                    dw2_acc[cb * 32 + j] = fmaf(dsr, g, dw2_acc[cb * 32 + j]);
                    dh[jj] = dsr * w * gp;
                  }
                  umma::store_operand8<NP>(sm.opBD[1][half], umma::tile64_offset(row, cb * 4 + j8), umma::TILE64_BYTES, dh);
                }
              }
            }
            if (act) {
              db2_acc += dsr;
              if (INPUT_GRADS) sm.sc[half][c][row] = sacc + b2;
            }
          }
          umma::fence_async_smem();
          umma::fence_before_sync();
          __syncthreads();                               // (2) dhid tiles visible; hid columns free
          if (tid == 0) {
            umma::fence_after_sync();
            for (int q = 0; q < nimp; ++q) {
              const umma::Operand h_mn = umma::op_tile64_mn(umma::smem_u32(sm.opA[q]));
              const umma::Operand d_mn = umma::op_tile64_mn(umma::smem_u32(sm.opBD[1][q]));
              const umma::Operand d_k = umma::op_tile64_k(umma::smem_u32(sm.opBD[1][q]));
              const umma::Operand w_mn = umma::op_tile64_mn(umma::smem_u32(sm.opBD[0][q]));
              umma::mma_product<SPLIT, 4>(tmem + BWD_COL_S + half_off[q], h_mn, d_mn, IDESC_ST, false);
              umma::mma_product<SPLIT, 4>(tmem + BWD_COL_DA + half_off[q], h_mn, d_mn, IDESC_ST, da_started[q]);
              if (INPUT_GRADS) umma::mma_product<SPLIT, 4>(tmem + BWD_COL_DH + half_off[q], d_k, w_mn, IDESC_DH, dh_started);
              // Gt: ones is exact in bf16, so only the (hi, hi) and (lo, hi) terms exist
              {
                const umma::Operand one = op_tile8_k(umma::smem_u32(sm.ones));
#pragma unroll
                for (int t = 0; t < NP; ++t)
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma::mma_bf16(tmem + BWD_COL_HID + half_off[q],
                                   umma::make_desc(d_mn.addr + t * d_mn.part + ks * d_mn.kstep, d_mn.lbo, d_mn.sbo),
                                   umma::make_desc(one.addr + ks * one.kstep, one.lbo, one.sbo), IDESC_GT, (t > 0 || ks > 0) ? 1u : 0u);
              }
            }
            umma::mma_commit(&sm.mbar);
          }
          for (int q = 0; q < nimp; ++q) da_started[q] = true;
          dh_started = true;
          umma::mbar_wait(&sm.mbar, phase);
          phase ^= 1;
          umma::fence_after_sync();
          // ---- epilogue 2: thread = (impression half, feature k = row) for S^T, (half, j = row) for Gt
          {
            const bool act = half < nimp;
            const float tk = act ? __ldg(e + rcm * E + TOFF + row) : 0.f;
            float dt = 0.f;
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
              float v[32];
              umma::tmem_ld32(my_tmem + BWD_COL_S + cb * 32, v);
              if (act) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  dwd_acc[cb * 32 + j] = fmaf(tk, v[j], dwd_acc[cb * 32 + j]);
                  if (INPUT_GRADS) dt = fmaf(v[j], __ldg(Wd_rm + (cb * 32 + j) * 256 + row), dt);
                }
              }
            }
            const float gt = umma::tmem_ld1(my_tmem + BWD_COL_HID);
            if (act) {
              float* gdst = dtp + rcm * 64 + row;
              if (r0 == 0) *gdst = gt; else *gdst += gt;
              if (INPUT_GRADS) {
                float* dst = dxt + rcm * 64 + row;
                if (r0 == 0) *dst = dt + __ldg(de + rcm * E + E_XT + row);   // + the direct ec path (user_model.py:31)
                else *dst += dt;
              }
            }
          }
          umma::fence_before_sync();
        }
        __syncthreads();                                 // ds / opP of this chunk consumed before the next chunk rewrites them
        if (INPUT_GRADS) {
          // pooling path of this chunk: dxh[row][k] (+)= sum_c s_c[row] dP_c[k]; kept in global (each thread owns its row)
          // so that candidate chunks compose; the W_c path is added from tensor memory at the end of the tile
          if (half < nimp && r0 + row < H) {
            float* dst = dxh + (bmine * H + r0 + row) * 64;
            for (int k4 = 0; k4 < 16; ++k4) {
              float4 acc = (c0 == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(dst + 4 * k4);
              for (int c = 0; c < nc; ++c) {
                const float s = sm.sc[half][c][row];
                const float4 dp = __ldg(reinterpret_cast<const float4*>(de + (bmine * C + c0 + c) * E + POFF) + k4);
                acc.x = fmaf(s, dp.x, acc.x); acc.y = fmaf(s, dp.y, acc.y); acc.z = fmaf(s, dp.z, acc.z); acc.w = fmaf(s, dp.w, acc.w);
              }
              *reinterpret_cast<float4*>(dst + 4 * k4) = acc;
            }
          }
        }
      }
      if (INPUT_GRADS) {
        // dH of this tile (all candidates) from TMEM + the pooling-path partial already in global
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          float v[32];
          umma::tmem_ld32(my_tmem + BWD_COL_DH + cb * 32, v);
          if (half < nimp && r0 + row < H) {
            float4* dst = reinterpret_cast<float4*>(dxh + (bmine * H + r0 + row) * 64 + cb * 32);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              float4 a = dst[k4];
              a.x += v[4 * k4]; a.y += v[4 * k4 + 1]; a.z += v[4 * k4 + 2]; a.w += v[4 * k4 + 3];
              dst[k4] = a;
            }
          }
        }
        umma::fence_before_sync();
      }
    }
  }

  // ---- per-CTA partial sums -> part[blockIdx.x]
  __syncthreads();
  float* out = part + (long long)blockIdx.x * TC_PARTIAL;
  {
    // dA^T from TMEM: thread (half, k) -> out[TCP_DA + half*4096 + k*64 + j]
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      float v[32];
      umma::tmem_ld32(my_tmem + BWD_COL_DA + cb * 32, v);
      float4* dst = reinterpret_cast<float4*>(out + TCP_DA + half * 4096 + row * 64 + cb * 32);
      float4* dst2 = reinterpret_cast<float4*>(out + TCP_DWD + half * 4096 + row * 64 + cb * 32);
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        dst[k4] = da_started[half] ? make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst2[k4] = make_float4(dwd_acc[cb * 32 + 4 * k4], dwd_acc[cb * 32 + 4 * k4 + 1], dwd_acc[cb * 32 + 4 * k4 + 2], dwd_acc[cb * 32 + 4 * k4 + 3]);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  {
    // dw2[j] = sum over the 128 threads of dw2_acc[j]: transpose through shared memory (operand buffers are free now)
    float* red = reinterpret_cast<float*>(sm.opBD);          // [128][65] floats = 33 KB <= sizeof(opBD)
    static_assert(sizeof(sm.opBD) >= 128 * 65 * sizeof(float) || NP == 1, "reduce scratch");
    float* red2 = reinterpret_cast<float*>(sm.opA);          // db2 partials
    if (NP == 2) {
#pragma unroll
      for (int j = 0; j < 64; ++j) red[tid * 65 + j] = dw2_acc[j];
      red2[tid] = db2_acc;
      __syncthreads();
      if (tid < 64) {
        float s = 0.f;
        for (int t = 0; t < 128; ++t) s += red[t * 65 + tid];
        out[TCP_DW2 + tid] = s;
      }
    } else {
      // NP == 1: opBD is 32 KB = 128 x 64 floats exactly; use an unpadded layout in two passes of 64 threads
      for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
        if ((tid >> 6) == pass) {
#pragma unroll
          for (int j = 0; j < 64; ++j) red[(tid & 63) * 65 + j] = dw2_acc[j];
        }
        if (pass == 0) red2[tid] = db2_acc;
        __syncthreads();
        if (tid < 64) {
          float s = 0.f;
          for (int t = 0; t < 64; ++t) s += red[t * 65 + tid];
          if (pass == 0) out[TCP_DW2 + tid] = s; else out[TCP_DW2 + tid] += s;
        }
      }
    }
    if (tid == 0) {
      float d = 0.f;
      for (int t = 0; t < 128; ++t) d += red2[t];
      out[TCP_DB2] = d;
    }
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, BWD_TMEM_COLS);
}

// Sum the per-CTA partials (fixed order) and write the attention-MLP gradients that do not depend on tp:
//   fc1.weight grad blocks: [:, 0:64] = dA, [:, 192:256] = dWd   (the Bm-dependent blocks are completed by
//   attention_tp_grad_kernel); fc2.weight = dw2; fc2.bias = db2.   Partials are transposed ([k][j]).
__global__ void __launch_bounds__(256)
attention_tc_compose_kernel(const float* __restrict__ part, int nparts, AttOffsets off, float* __restrict__ grads,
                            float* __restrict__ dA_out) {
  __shared__ float red[4][2][64];
  const int lane = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + lane;              // k*64 + j, grid = 64 blocks
  float dA = 0.f, dWd = 0.f;
  for (int p = grp; p < nparts; p += 4) {
    const float* q = part + (long long)p * TC_PARTIAL;
    dA += q[TCP_DA + i] + q[TCP_DA + 4096 + i];
    dWd += q[TCP_DWD + i] + q[TCP_DWD + 4096 + i];
  }
  red[grp][0][lane] = dA; red[grp][1][lane] = dWd;
  __syncthreads();
  if (grp == 0) {
    dA = ((red[0][0][lane] + red[1][0][lane]) + red[2][0][lane]) + red[3][0][lane];
    dWd = ((red[0][1][lane] + red[1][1][lane]) + red[2][1][lane]) + red[3][1][lane];
    const int k = i >> 6, j = i & 63;
    float* rowp = grads + off.fc1_w + j * 256;
    rowp[k] = dA; rowp[192 + k] = dWd;
    dA_out[j * 64 + k] = dA;
  }
  if (blockIdx.x == 0 && grp == 1) {
    float w = 0.f;
    for (int p = 0; p < nparts; ++p) w += part[(long long)p * TC_PARTIAL + TCP_DW2 + lane];
    grads[off.fc2_w + lane] = w;
    if (lane == 0) {
      float d = 0.f;
      for (int p = 0; p < nparts; ++p) d += part[(long long)p * TC_PARTIAL + TCP_DB2];
      grads[off.fc2_b] = d;
    }
  }
}

// The tp = Bm t + b1 path, over the R candidate rows: dBm[j][k] = sum_r dtp[r][j] t[r][k], db1[j] = sum_r dtp[r][j],
// and (label branch) dxt[r][k] += sum_j dtp[r][j] Bm[j][k].  Row chunks per CTA -> partials, summed by the finish kernel.
constexpr int TPG_ROWS = 64;
__global__ void __launch_bounds__(256)
attention_tp_grad_kernel(const float* __restrict__ dtp, const float* __restrict__ e, int toff, long long R,
                         const float* __restrict__ der, int input_grads, float* __restrict__ dxt, float* __restrict__ part) {
  __shared__ float sd[TPG_ROWS][64];    // dtp rows
  __shared__ float st[TPG_ROWS][64];    // t rows
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * TPG_ROWS;
  const int nr = (int)min((long long)TPG_ROWS, R - r0);
  for (int i = tid; i < TPG_ROWS * 16; i += 256) {
    const int r = i >> 4, k4 = i & 15;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < nr) {
      a = __ldg(reinterpret_cast<const float4*>(dtp + (r0 + r) * 64) + k4);
      b = __ldg(reinterpret_cast<const float4*>(e + (r0 + r) * E + toff) + k4);
    }
    *reinterpret_cast<float4*>(&sd[r][4 * k4]) = a;
    *reinterpret_cast<float4*>(&st[r][4 * k4]) = b;
  }
  __syncthreads();
  // dBm partial: thread (tj, tk) owns a 4 x 4 block of [j][k]
  {
    const int tj = tid >> 4, tk = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    float b1[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < TPG_ROWS; ++r) {
      const float4 dj = *reinterpret_cast<const float4*>(&sd[r][4 * tj]);
      const float4 tk4 = *reinterpret_cast<const float4*>(&st[r][4 * tk]);
      const float dv[4] = {dj.x, dj.y, dj.z, dj.w}, tv[4] = {tk4.x, tk4.y, tk4.z, tk4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        b1[a] += dv[a];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(dv[a], tv[b], acc[a][b]);
      }
    }
    float* out = part + (long long)blockIdx.x * (4096 + 64);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      *reinterpret_cast<float4*>(out + (4 * tj + a) * 64 + 4 * tk) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      if (tk == 0) out[4096 + 4 * tj + a] = b1[a];
    }
  }
  if (input_grads) {
    // dxt[r][k] += sum_j dtp[r][j] BmT[j... ] : BmT is stored [k][j] -> Bm[j][k] = BmT[k*64 + j]
    for (int i = tid; i < nr * 64; i += 256) {
      const int r = i >> 6, k = i & 63;
      float v = 0.f;
#pragma unroll 8
      for (int j = 0; j < 64; ++j) v = fmaf(sd[r][j], __ldg(der + DER_BMT + k * 64 + j), v);
      dxt[(r0 + r) * 64 + k] += v;
    }
  }
}

// fc1.weight grad blocks [:, 64:128] = dBm and [:, 128:192] = dBm - dA; fc1.bias = db1
__global__ void __launch_bounds__(256)
attention_tp_finish_kernel(const float* __restrict__ part, int nparts, const float* __restrict__ dA, AttOffsets off,
                           float* __restrict__ grads) {
  const int i = blockIdx.x * 256 + threadIdx.x;      // 0 .. 4096+64
  if (i >= 4096 + 64) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(long long)p * (4096 + 64) + i];
  if (i < 4096) {
    const int j = i >> 6, k = i & 63;
    grads[off.fc1_w + j * 256 + 64 + k] = s;
    grads[off.fc1_w + j * 256 + 128 + k] = s - dA[i];
  } else {
    grads[off.fc1_b + (i - 4096)] = s;
  }
}

// ---------------------------------------------------------------------------------
// Self test of the tensor-core building blocks (tests/test_gpu_tensorcore.py): four 64x64x64 products with the
// interleaved half-lane accumulators:
//   out[0] = a0 b0^T (K-major x K-major)      out[1] = a1 b1^T (upper half-lanes)
//   out[2] = a0^T b0 (both tiles read MN-major)   out[3] = a1 b1^T... see below
// mode 0: K-major / K-major;  mode 1: A MN-major, B MN-major (out = a^T b);  mode 2: A K-major, B MN-major (out = a b)
// split = 1 or 3.
// ---------------------------------------------------------------------------------
template <int SPLIT>
__global__ void __launch_bounds__(TC_THREADS)
umma_selftest_kernel(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ b0,
                     const float* __restrict__ b1, float* __restrict__ out, int mode) {
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  extern __shared__ __align__(128) unsigned char st_raw[];
  unsigned char* opA[2] = {st_raw, st_raw + TileBytes<NP>::T64};
  unsigned char* opB[2] = {st_raw + 2 * TileBytes<NP>::T64, st_raw + 3 * TileBytes<NP>::T64};
  uint64_t* mbar = reinterpret_cast<uint64_t*>(st_raw + 4 * TileBytes<NP>::T64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* srcs[4] = {a0, a1, b0, b1};
  for (int m = 0; m < 4; ++m) {
    unsigned char* tile = m < 2 ? opA[m] : opB[m - 2];
    for (int it = tid; it < 64 * 8; it += TC_THREADS) {
      const int r = it & 63, kb = it >> 6;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = srcs[m][r * 64 + kb * 8 + i];
      umma::store_operand8<NP>(tile, umma::tile64_offset(r, kb), umma::TILE64_BYTES, v);
    }
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 64);
  if (tid == 0) umma::mbar_init(mbar, 1);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    for (int q = 0; q < 2; ++q) {
      const uint32_t d = tmem + ((uint32_t)(16 * q) << 16);
      const uint32_t aa = umma::smem_u32(opA[q]), bb = umma::smem_u32(opB[q]);
      if (mode == 0) umma::mma_product<SPLIT, 4>(d, umma::op_tile64_k(aa), umma::op_tile64_k(bb), umma::make_idesc_bf16(64, 64), false);
      else if (mode == 1) umma::mma_product<SPLIT, 4>(d, umma::op_tile64_mn(aa), umma::op_tile64_mn(bb), umma::make_idesc_bf16(64, 64, true, true), false);
      else umma::mma_product<SPLIT, 4>(d, umma::op_tile64_k(aa), umma::op_tile64_mn(bb), umma::make_idesc_bf16(64, 64, false, true), false);
    }
    umma::mma_commit(mbar);
  }
  umma::mbar_wait(mbar, 0);
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

template <int BRANCH, int SPLIT>
static int launch_fwd(const BatchPtrs& in, Workspace& w, cudaStream_t s) {
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  // pad the request so that no more CTAs become resident than tensor memory can serve (512 columns / 128)
  const size_t smem = sizeof(TcSmemFwd<NP>) < 57 * 1024 ? 57 * 1024 : sizeof(TcSmemFwd<NP>);
  const int per_sm = SPLIT == 3 ? 3 : 4;
  const int grid = min((w.B + 1) / 2, per_sm * sm_count());
  NRM_CUDA(cudaFuncSetAttribute(attention_forward_tc_kernel<BRANCH, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_forward_tc_kernel<BRANCH, SPLIT><<<grid, TC_THREADS, smem, s>>>(in.xh, w.xh, w.B, w.H, w.C, w.att_derived, w.e);
  NRM_LAUNCH_CHECK("attention_forward_tc_kernel");
  return NRM_OK;
}

int launch_attention_forward_tc(const BatchPtrs& in, Workspace& w, int branch, int precision, cudaStream_t s) {
  if (precision == NRM_PRECISION_BF16) return branch == 0 ? launch_fwd<0, 1>(in, w, s) : launch_fwd<1, 1>(in, w, s);
  return branch == 0 ? launch_fwd<0, 3>(in, w, s) : launch_fwd<1, 3>(in, w, s);
}

static int att_tc_bwd_grid(int B) { return min((B + 1) / 2, min(2 * sm_count(), ATT_TC_PARTS_MAX)); }

template <int BRANCH, int SPLIT>
static int launch_bwd(const BatchPtrs& in, const float* P, Workspace& w, cudaStream_t s) {
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  // two CTAs per SM at most (2 x 256 tensor-memory columns): pad small requests accordingly
  const size_t smem = sizeof(TcSmemBwd<NP>) < 80 * 1024 ? 80 * 1024 : sizeof(TcSmemBwd<NP>);
  const int grid = att_tc_bwd_grid(w.B);
  float* part = w.att_part + (long long)BRANCH * ATT_TC_PARTS_MAX * TC_PARTIAL;
  float* dtp = w.dtp + (long long)BRANCH * w.R * 64;
  NRM_CUDA(cudaFuncSetAttribute(attention_backward_tc_kernel<BRANCH, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_backward_tc_kernel<BRANCH, SPLIT><<<grid, TC_THREADS, smem, s>>>(in.xh, w.xh, w.B, w.H, w.C, w.att_derived, P, w.e, w.de,
                                                                            w.dxh, w.dxt, dtp, part);
  NRM_LAUNCH_CHECK("attention_backward_tc_kernel");
  return NRM_OK;
}

int launch_attention_backward_tc(const BatchPtrs& in, const float* P, Workspace& w, int branch, int precision, cudaStream_t s) {
  if (precision == NRM_PRECISION_BF16) return branch == 0 ? launch_bwd<0, 1>(in, P, w, s) : launch_bwd<1, 1>(in, P, w, s);
  return branch == 0 ? launch_bwd<0, 3>(in, P, w, s) : launch_bwd<1, 3>(in, P, w, s);
}

int launch_attention_finish_tc(const float* P, Workspace& w, int branch, float* grads, cudaStream_t s) {
  (void)P;
  const AttOffsets off = branch == 0 ? ATT_LABEL : ATT_TI;
  const float* part = w.att_part + (long long)branch * ATT_TC_PARTS_MAX * TC_PARTIAL;
  float* dA = w.att_dA + branch * 4096;
  attention_tc_compose_kernel<<<64, 256, 0, s>>>(part, att_tc_bwd_grid(w.B), off, grads, dA);
  NRM_LAUNCH_CHECK("attention_tc_compose_kernel");
  const int nparts = (int)((w.R + TPG_ROWS - 1) / TPG_ROWS);
  float* tpart = w.tp_part;
  const float* dtp = w.dtp + (long long)branch * w.R * 64;
  attention_tp_grad_kernel<<<nparts, 256, 0, s>>>(dtp, w.e, branch == 0 ? E_XT : E_PCAT, w.R, w.att_derived + (long long)branch * DER_SIZE,
                                                  branch == 0 ? 1 : 0, w.dxt, tpart);
  NRM_LAUNCH_CHECK("attention_tp_grad_kernel");
  attention_tp_finish_kernel<<<(4096 + 64 + 255) / 256, 256, 0, s>>>(tpart, nparts, dA, off, grads);
  NRM_LAUNCH_CHECK("attention_tp_finish_kernel");
  return NRM_OK;
}

}  // namespace nrm

using namespace nrm;

// a0, a1, b0, b1: [64,64] fp32 (device); out: [2,64,64] fp32.  mode 0: out_q = a_q b_q^T; 1: a_q^T b_q; 2: a_q b_q.
// split 1: operands rounded to bf16; 3: hi/lo split (fp32-grade).
extern "C" int nrm_debug_umma_selftest(const float* a0, const float* a1, const float* b0, const float* b1, float* out,
                                       int mode, int split, void* stream) {
  if (!a0 || !a1 || !b0 || !b1 || !out || mode < 0 || mode > 2 || (split != 1 && split != 3)) {
    set_error("nrm_debug_umma_selftest: bad argument"); return NRM_EINVAL;
  }
  if (split == 1) {
    const size_t smem = 4 * TileBytes<1>::T64 + 64;
    NRM_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<1><<<1, TC_THREADS, smem, (cudaStream_t)stream>>>(a0, a1, b0, b1, out, mode);
  } else {
    const size_t smem = 4 * TileBytes<2>::T64 + 64;
    NRM_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<3><<<1, TC_THREADS, smem, (cudaStream_t)stream>>>(a0, a1, b0, b1, out, mode);
  }
  NRM_LAUNCH_CHECK("umma_selftest_kernel");
  return NRM_OK;
}
