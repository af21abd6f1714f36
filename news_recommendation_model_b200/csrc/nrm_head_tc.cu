// Scoring head of UserModel (models/user_model.py:31-35) on the tensor cores (precision = bf16 / bf16x3):
//     z = BatchNorm(e);  gate = G2 gelu(G1 z);  x = gate * e;  y = M2 gelu(M1 x);  r = o2 . gelu(O1 y)
// Five 264 <-> 66 products per candidate row.  A CTA takes a tile of 64 candidate rows through the whole chain: every product is
// a sequence of tcgen05.mma (M = 64) with BOTH operands in shared memory -- the activations as an un-swizzled K-major bf16 tile the
// epilogue warps write, the weights as pre-built operand images a loader warp streams with bulk async copies (TMA engine) through
// a three-slot ring -- and the accumulator in tensor memory, read back by 16 epilogue warps that add the bias, apply GELU / the
// gating product, keep what the backward needs and write the next layer's operand.
//
// Precision.  BatchNorm multiplies channels whose batch variance is ~0 by up to 316, so the head needs fp32-grade products to
// keep the 1e-4 absolute logit tolerance: with bf16x3 every operand is split into THREE bf16 parts (hi + mid + lo = 24 mantissa
// bits) and a product is issued as the six part products of order <= 2^-16 (hh, hm, mh, mm, hl, lh), accumulated in fp32 by the
// tensor core.  precision = bf16 uses two parts / three products (2^-16 relative).
//
// M = 64: an accumulator occupies lanes 0-15 of every 32-lane tensor-memory sub-partition.  The 264-wide layers are issued as two
// column halves (136 + 128), the second one at lane offset 16 of the SAME columns, so a 32x32b tcgen05.ld gives every lane of an
// epilogue warp useful data: lane = (column half, row).  The 66-wide layers (N = 72) leave lanes 16-31 idle.
//
// The data-gradient chain of the backward (head_backward_tc_kernel) is the same skeleton walked in the other direction.
#include "nrm_kernels.cuh"
#include <cstddef>

#include "nrm_umma.cuh"

namespace nrm {
namespace htc {

// -DNRM_RS_PROFILE: CTA 0 adds the clock64 cycles of each role's waits / phases to g_hprof[role * 8 + kind] (slot 7 = role total);
// roles: 0 loader, 1 MMA issuer, 2 epilogue warp 0.  Read and cleared by nrm_debug_headprof.
#ifdef NRM_RS_PROFILE
__device__ long long g_hprof[32];
#define HPROF(role, kind, stmt) do { const long long t__ = clock64(); stmt; if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_hprof[(role) * 8 + (kind)] += clock64() - t__; } while (0)
#else
#define HPROF(role, kind, stmt) do { stmt; } while (0)
#endif

constexpr int ROWS = 64;                       // candidate rows per tile
constexpr int N_EPI = 16;                      // epilogue warps: sub-partition = warp & 3, column quarter = warp >> 2
constexpr int W_MMA = 16, W_LOAD = 17;
constexpr int THREADS = 32 * 20;                // warps 18, 19 idle: setmaxnreg moves registers between whole warpgroups
constexpr int EPI_THREADS = 32 * N_EPI;
constexpr int KE = 272, KH = 80;               // contraction lengths padded to multiples of 16 (264, 66)
constexpr int NN = 72;                         // 66 outputs padded to a multiple of 8
constexpr int NW0 = 136, NW1 = 128;            // the two column halves of a 264-wide layer
constexpr uint32_t A_LBO = ROWS * 16;          // activation tiles: (r, 8 kb) at kb * 1024 + (r / 8) * 128 + (r % 8) * 16
constexpr uint32_t AW_PART = (KE / 8) * A_LBO; // 34 816 B per part (264-wide activations)
constexpr uint32_t AN_PART = (KH / 8) * A_LBO; // 10 240 B per part (66-wide activations)
constexpr uint32_t BN_LBO = NN * 16;           // weight image of a 66-output layer: 72 rows
constexpr uint32_t BW_LBO = E * 16;            // weight image of a 264-output layer: 264 rows
constexpr int NCHUNK = 5;                      // ring chunks per layer: 4 + 4 + 4 + 4 + 1 K steps (264 in) or 1 K step each (66 in)
constexpr uint32_t SLOT = 3 * 8 * BN_LBO;      // 27 648 B: largest chunk (three parts of 8 K blocks x 72 rows)
constexpr int STAGES = 3;
constexpr uint32_t IMG_STRIDE = 3 * (KH / 8) * BW_LBO;   // 126 720 B per layer image (>= 3 * 34 * 1152 = 117 504)
constexpr uint32_t COL_W = 0, COL_N = 160, TMEM_COLS = 256;
// per-CTA copy of the small parameter vectors the forward epilogues read (shared memory instead of an L2 round trip per use)
constexpr int PAR_GAMMA = 0, PAR_BETA = E, PAR_BG2 = 2 * E, PAR_BM2 = 3 * E, PAR_BG1 = 4 * E, PAR_BM1 = 4 * E + NN, PAR_BO1 = 4 * E + 2 * NN,
              PAR_WO2 = 4 * E + 3 * NN, PAR_FLOATS = 8 * E;        // the backward uses it as [4 sub-partitions][2][264] partial sums

__host__ __device__ constexpr int narrow_chunk_kb(int c) { return c < 4 ? 8 : 2; }
template <int NP> __host__ __device__ constexpr uint32_t narrow_chunk_off(int c) { return (uint32_t)c * NP * 8 * BN_LBO; }
template <int NP> __host__ __device__ constexpr uint32_t wide_chunk_off(int c) { return (uint32_t)c * NP * 2 * BW_LBO; }

// part products of one algebraic product: (A part, B part), order of magnitude 2^-8 (pa + pb)
__host__ __device__ constexpr int n_products(int NP) { return NP == 1 ? 1 : NP == 2 ? 3 : 6; }
__host__ __device__ constexpr int prod_a(int NP, int p) { return NP == 2 ? (p == 2 ? 1 : 0) : (p == 2 || p == 3) ? 1 : p == 5 ? 2 : 0; }
__host__ __device__ constexpr int prod_b(int NP, int p) { return NP == 2 ? (p == 1 ? 1 : 0) : (p == 1 || p == 3) ? 1 : p == 4 ? 2 : 0; }

// 8 fp32 -> NP bf16 parts (hi, mid = bf16(v - hi), lo = bf16(v - hi - mid)), one 16-byte store per part
template <int NP>
__device__ __forceinline__ void split_store8(unsigned char* dst, uint32_t part_bytes, const float* v) {
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = v[2 * i], b = v[2 * i + 1];
    const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    if (NP >= 2) {
      const float ra = a - __uint_as_float(h[i] << 16), rb = b - __uint_as_float(h[i] & 0xffff0000u);
      const __nv_bfloat162 mm = __floats2bfloat162_rn(ra, rb);
      m[i] = *reinterpret_cast<const uint32_t*>(&mm);
      if (NP >= 3) {
        const __nv_bfloat162 ll = __floats2bfloat162_rn(ra - __uint_as_float(m[i] << 16), rb - __uint_as_float(m[i] & 0xffff0000u));
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
      }
    }
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
  if (NP >= 2) *reinterpret_cast<uint4*>(dst + part_bytes) = make_uint4(m[0], m[1], m[2], m[3]);
  if (NP >= 3) *reinterpret_cast<uint4*>(dst + 2 * part_bytes) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------------------------------------------------
// weight images.  Layer image l (0-4 forward, 5-9 backward) = B operand [out][in] of that layer's product, B[o][i] =
// src[o * so + i * si], zero-padded, split into NP parts, laid out chunk by chunk exactly as the ring slots are consumed.
//   66-output layers (in = 264): chunk c = K blocks [8c, 8c + 8) (last: 2 blocks); part p of a chunk follows part p - 1
//   264-output layers (in = 66): chunk c = K blocks [2c, 2c + 2)
// ---------------------------------------------------------------------------------------------------------------------
struct ImgDesc { long long src; int so, si, narrow; };
__constant__ ImgDesc c_img[10] = {
    {P_GATE_FC1_W, E, 1, 1},   {P_GATE_FC2_W, HID, 1, 0}, {P_MLP_FC1_W, E, 1, 1},  {P_MLP_FC2_W, HID, 1, 0}, {P_OUT_FC1_W, E, 1, 1},
    // backward: dy = da3 O1 | da2 = dy M2 | dx = da2 M1 | da1 = dgate G2 | dz = da1 G1
    {P_OUT_FC1_W, 1, E, 0},    {P_MLP_FC2_W, 1, HID, 1},  {P_MLP_FC1_W, 1, E, 0},  {P_GATE_FC2_W, 1, HID, 1}, {P_GATE_FC1_W, 1, E, 0}};

template <int NP>
__global__ void __launch_bounds__(256)
head_image_kernel(const float* __restrict__ P, unsigned char* __restrict__ img, int first_layer, unsigned int* __restrict__ arrive) {
  pdl_wait();
  pdl_trigger();
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *arrive = 0u;
  const int layer = first_layer + blockIdx.y;
  const ImgDesc d = c_img[layer];
  const float* src = P + d.src;
  unsigned char* out = img + (size_t)layer * IMG_STRIDE;
  const int nrow = d.narrow ? NN : E, nkb = d.narrow ? KE / 8 : KH / 8;
  const int nout = d.narrow ? HID : E, nin = d.narrow ? E : HID;
  for (int it = blockIdx.x * 256 + threadIdx.x; it < nrow * nkb; it += gridDim.x * 256) {
    const int n = it % nrow, kb = it / nrow;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = 8 * kb + i;
      v[i] = (n < nout && k < nin) ? __ldg(src + (long long)n * d.so + (long long)k * d.si) : 0.f;
    }
    uint32_t off, part;
    if (d.narrow) {
      const int c = kb < 32 ? kb >> 3 : 4, kbl = kb - 8 * c;
      part = (uint32_t)narrow_chunk_kb(c) * BN_LBO;
      off = narrow_chunk_off<NP>(c) + (uint32_t)kbl * BN_LBO;
    } else {
      const int c = kb >> 1, kbl = kb & 1;
      part = 2 * BW_LBO;
      off = wide_chunk_off<NP>(c) + (uint32_t)kbl * BW_LBO;
    }
    split_store8<NP>(out + off + (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u, part, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// shared skeleton
// ---------------------------------------------------------------------------------------------------------------------
struct Smem {
  __align__(128) unsigned char aw[3 * AW_PART];          // 264-wide activations (z, x, y | dy, dgate): A operand of the 66-output layers
  __align__(128) unsigned char an[3 * AN_PART];          // 66-wide activations
  __align__(128) unsigned char ring[STAGES][SLOT];       // weight chunks
  float mean[E], rstd[E];
  float red[4][ROWS + 4];                                // cross-warp partial sums per row
  float par[PAR_FLOATS];                                 // forward: BatchNorm affine + biases + out_mlp.fc2 (see PAR_*); backward: reduction scratch
  uint64_t full[STAGES], empty[STAGES];                  // ring: chunk landed / chunk consumed (tcgen05.commit)
  uint64_t a_ready, d_ready;                             // operand of the next layer written / accumulator of this layer complete
  uint32_t tmem_base;
};

// bar.sync is the ALIGNED barrier: a warp must arrive converged (a diverged warp arriving in two groups releases it early)
__device__ __forceinline__ void epi_sync() { __syncwarp(); asm volatile("bar.sync 1, %0;\n" ::"n"(EPI_THREADS) : "memory"); }
// the calling epilogue warp has written its share of the next operand: make it visible to the tensor core, one arrival per warp
__device__ __forceinline__ void operand_written(uint64_t* bar) {
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bar);
}

// loader warp (one lane): the 25 chunks of every tile of this CTA, in consumption order
template <int NP>
__device__ __forceinline__ void loader(Smem& sm, const unsigned char* __restrict__ img, int first_layer, bool first_narrow, long long my_tiles) {
  uint32_t g = 0;
  for (long long t = 0; t < my_tiles; ++t)
    for (int l = 0; l < 5; ++l) {
      const bool narrow = ((l & 1) == 0) == first_narrow;
      const unsigned char* base = img + (size_t)(first_layer + l) * IMG_STRIDE;
      for (int c = 0; c < NCHUNK; ++c, ++g) {
        const int slot = g % STAGES, use = g / STAGES;
        if (use > 0) HPROF(0, 0, umma::mbar_wait(&sm.empty[slot], (uint32_t)((use - 1) & 1)));
        const uint32_t off = narrow ? narrow_chunk_off<NP>(c) : wide_chunk_off<NP>(c);
        const uint32_t bytes = narrow ? (uint32_t)(NP * narrow_chunk_kb(c)) * BN_LBO : (uint32_t)NP * 2 * BW_LBO;
        umma::bulk_load(sm.ring[slot], base + off, bytes, &sm.full[slot]);
      }
    }
}

// one ring chunk of a 66-output layer: KS K steps x the part products, D[64 x 72] += A[64 x 16 KS] B^T
template <int NP, int KS>
__device__ __forceinline__ void narrow_chunk(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool accumulate) {
  constexpr uint32_t ID_N = umma::make_idesc_bf16(64, NN);
  constexpr uint32_t bpart = (uint32_t)(2 * KS) * BN_LBO;
#pragma unroll
  for (int p = 0; p < n_products(NP); ++p) {
    const uint64_t da = umma::make_desc(a_addr + prod_a(NP, p) * AW_PART, A_LBO, 128);
    const uint64_t db = umma::make_desc(b_addr + prod_b(NP, p) * bpart, BN_LBO, 128);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      umma::mma_bf16(tmem_d, da + (uint64_t)(ks * ((2 * A_LBO) >> 4)), db + (uint64_t)(ks * ((2 * BN_LBO) >> 4)), ID_N,
                     (accumulate || p > 0 || ks > 0) ? 1u : 0u);
  }
}

// MMA warp (one elected lane): five layers per tile, alternating 66-output (A = aw, N = 72) and 264-output (A = an, two column halves)
template <int NP>
__device__ __forceinline__ void issuer(Smem& sm, uint32_t tmem, bool first_narrow, long long my_tiles) {
  constexpr uint32_t ID_W0 = umma::make_idesc_bf16(64, NW0), ID_W1 = umma::make_idesc_bf16(64, NW1);
  const uint32_t aw = umma::smem_u32(sm.aw), an = umma::smem_u32(sm.an);
  uint32_t g = 0, ause = 0;
  for (long long t = 0; t < my_tiles; ++t)
    for (int l = 0; l < 5; ++l) {
      const bool narrow = ((l & 1) == 0) == first_narrow;
      HPROF(1, 0, umma::mbar_wait(&sm.a_ready, ause & 1));
      ++ause;
      umma::fence_after_sync();
      for (int c = 0; c < NCHUNK; ++c, ++g) {
        const int slot = g % STAGES;
        HPROF(1, 1, umma::mbar_wait(&sm.full[slot], (uint32_t)((g / STAGES) & 1)));
        umma::fence_after_sync();
        const uint32_t rb = umma::smem_u32(sm.ring[slot]);
        if (narrow) {
          // fully unrolled per chunk shape: the descriptors of consecutive MMAs differ by constants, so the issue loop is
          // back-to-back UTCHMMA (a run-time K-step loop costs ~12 uniform-datapath instructions per MMA: 65 instead of 45 cycles)
          const uint32_t a0 = aw + (uint32_t)(8 * c) * A_LBO;
          if (c < 4) narrow_chunk<NP, 4>(tmem + COL_N, a0, rb, c > 0);
          else narrow_chunk<NP, 1>(tmem + COL_N, a0, rb, true);
        } else {
#pragma unroll
          for (int p = 0; p < n_products(NP); ++p) {
            const uint64_t da = umma::make_desc(an + prod_a(NP, p) * AN_PART + (uint32_t)(2 * c) * A_LBO, A_LBO, 128);
            const uint64_t db = umma::make_desc(rb + prod_b(NP, p) * (2 * BW_LBO), BW_LBO, 128);
            const uint32_t acc = (c > 0 || p > 0) ? 1u : 0u;
            umma::mma_bf16(tmem + COL_W, da, db, ID_W0, acc);
            umma::mma_bf16(tmem + COL_W + (16u << 16), da, db + (uint64_t)(((NW0 / 8) * 128) >> 4), ID_W1, acc);
          }
        }
        umma::mma_commit(&sm.empty[slot]);
      }
      umma::mma_commit(&sm.d_ready);
    }
}

// column blocks (8 columns each) of the calling warp's quarter: 264-wide layers 17 blocks per half (5 | 4 | 4 | 4), 66-wide 9 (3 | 2 | 2 | 2)
__device__ __forceinline__ int wide_kb0(int q) { return q == 0 ? 0 : 4 * q + 1; }
__device__ __forceinline__ int wide_nkb(int q) { return q == 0 ? 5 : 4; }
__device__ __forceinline__ int narrow_kb0(int q) { return q == 0 ? 0 : 2 * q + 1; }
__device__ __forceinline__ int narrow_nkb(int q) { return q == 0 ? 3 : 2; }

// accumulator columns of this thread's lane for its quarter.  Warp-collective.
__device__ __forceinline__ void load_wide(uint32_t tmem_lane, int q, float* v /*[40]*/) {
  const uint32_t a = tmem_lane + COL_W + 8 * wide_kb0(q);
  umma::tmem_ld32(a, v);
  if (q == 0) umma::tmem_ld8(a + 32, v + 32);
}
__device__ __forceinline__ void load_narrow(uint32_t tmem_lane, int q, float* v /*[24]*/) {
  const uint32_t a = tmem_lane + COL_N + 8 * narrow_kb0(q);
  umma::tmem_ld16(a, v);
  if (q == 0) umma::tmem_ld8(a + 16, v + 16);
}

__device__ __forceinline__ uint32_t a_off(int row, int kb) { return (uint32_t)kb * A_LBO + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u; }

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(THREADS, 1)
head_forward_tc_kernel(const float* __restrict__ e, const double* __restrict__ bn_sums, long long bn_rows, int training,
                       float* __restrict__ run_mean, float* __restrict__ run_var, long long* __restrict__ nbt,
                       float* __restrict__ mean_out, float* __restrict__ rstd_out, const float* __restrict__ P,
                       const unsigned char* __restrict__ img, long long R, int keep, float* __restrict__ a1g, float* __restrict__ gateg,
                       float* __restrict__ a2g, float* __restrict__ yg, float* __restrict__ a3g, float* __restrict__ logits) {
#ifdef NRM_RS_PROFILE
  const long long t_entry = clock64();
#endif
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (R + ROWS - 1) / ROWS;
  const long long my_tiles = (long long)blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  if (tid == 0) {
    for (int j = 0; j < STAGES; ++j) { umma::mbar_init(&sm.full[j], 1); umma::mbar_init(&sm.empty[j], 1); }
    umma::mbar_init(&sm.a_ready, N_EPI);
    umma::mbar_init(&sm.d_ready, 1);
  }
  if (warp == W_MMA) umma::tmem_alloc(&sm.tmem_base, TMEM_COLS);
  // padding K blocks of the two activation tiles: never written again, must be finite zeros
  if (tid < ROWS) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      *reinterpret_cast<uint4*>(sm.aw + p * AW_PART + a_off(tid, KE / 8 - 1)) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sm.an + p * AN_PART + a_off(tid, KH / 8 - 1)) = make_uint4(0, 0, 0, 0);
    }
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp >= N_EPI) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");       // the loader / issuer warpgroup hands its registers to the epilogue warps
    if (warp == W_LOAD) {
      if (umma::elect_one()) HPROF(0, 7, loader<NP>(sm, img, 0, true, my_tiles));
    } else if (warp == W_MMA) {
      if (umma::elect_one()) HPROF(1, 7, issuer<NP>(sm, tmem, true, my_tiles));
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
    // ---- epilogue warps --------------------------------------------------------------------------------------------
    const int sp = warp & 3, q = warp >> 2, l16 = lane & 15, half = lane >> 4;
    const int row = 16 * sp + l16;
    const uint32_t tmem_lane = tmem + ((uint32_t)(32 * sp) << 16);
    // small parameter vectors -> shared memory (zero beyond the 66 real entries)
    for (int i = tid; i < E; i += EPI_THREADS) {
      sm.par[PAR_GAMMA + i] = __ldg(P + P_BN_W + i); sm.par[PAR_BETA + i] = __ldg(P + P_BN_B + i);
      sm.par[PAR_BG2 + i] = __ldg(P + P_GATE_FC2_B + i); sm.par[PAR_BM2 + i] = __ldg(P + P_MLP_FC2_B + i);
    }
    if (tid < NN) {
      const bool in = tid < HID;
      sm.par[PAR_BG1 + tid] = in ? __ldg(P + P_GATE_FC1_B + tid) : 0.f; sm.par[PAR_BM1 + tid] = in ? __ldg(P + P_MLP_FC1_B + tid) : 0.f;
      sm.par[PAR_BO1 + tid] = in ? __ldg(P + P_OUT_FC1_B + tid) : 0.f;  sm.par[PAR_WO2 + tid] = in ? __ldg(P + P_OUT_FC2_W + tid) : 0.f;
    }
    // BatchNorm1d statistics (models/user_model.py:18,32), every CTA for itself: training = batch mean / biased variance from the
    // (global) column sums; eval = running statistics.  CTA 0 publishes them for the backward and updates the running statistics
    if (tid < E) {
      const int n = tid;
      float m, v;
      if (training) {
        const double dm = bn_sums[n] / (double)bn_rows;
        double dv = bn_sums[E + n] / (double)bn_rows - dm * dm;
        if (dv < 0.0) dv = 0.0;
        m = (float)dm; v = (float)dv;
        if (blockIdx.x == 0) {
          const double unbiased = bn_rows > 1 ? dv * (double)bn_rows / (double)(bn_rows - 1) : dv;
          run_mean[n] = (1.f - BN_MOMENTUM) * run_mean[n] + BN_MOMENTUM * m;
          run_var[n] = (1.f - BN_MOMENTUM) * run_var[n] + BN_MOMENTUM * (float)unbiased;
          if (n == 0) *nbt += 1;
        }
      } else {
        m = run_mean[n]; v = run_var[n];
      }
      const float rs = 1.0f / sqrtf(v + BN_EPS);
      sm.mean[n] = m; sm.rstd[n] = rs;
      if (blockIdx.x == 0) { mean_out[n] = m; rstd_out[n] = rs; }
    }
    constexpr int PRO_ITEMS = ROWS * (E / 8), PRO_IT = (PRO_ITEMS + EPI_THREADS - 1) / EPI_THREADS;
    // the e rows of a tile for the prologue: items (row, K block), consecutive lanes take consecutive rows (conflict-free 16-byte
    // tile stores); every load is in flight before the first use
    auto load_e_items = [&](long long r0, int nr, float4 (&ev)[PRO_IT][2]) {
#pragma unroll
      for (int u = 0; u < PRO_IT; ++u) {
        const int it = tid + u * EPI_THREADS, r = it & (ROWS - 1), kb = it >> 6;
        ev[u][0] = ev[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < PRO_ITEMS && r < nr) {
          const float4* ep = reinterpret_cast<const float4*>(e + (r0 + r) * E + 8 * kb);
          ev[u][0] = __ldg(ep); ev[u][1] = __ldg(ep + 1);
        }
      }
    };
    epi_sync();
#ifdef NRM_RS_PROFILE
    if (blockIdx.x == 0 && tid == 0) g_hprof[24] += clock64() - t_entry;
#endif
    uint32_t duse = 0;
    for (long long t = 0; t < my_tiles; ++t) {
      const long long r0 = ((long long)blockIdx.x + t * gridDim.x) * ROWS;
      const int nr = (int)min((long long)ROWS, R - r0);
      const bool live = row < nr;
#ifdef NRM_RS_PROFILE
      const long long tp0 = clock64();
#endif
      // z = BatchNorm(e) -> aw
      float4 ev0[PRO_IT][2];
      load_e_items(r0, nr, ev0);
#pragma unroll
      for (int u = 0; u < PRO_IT; ++u) {
        const int it = tid + u * EPI_THREADS, r = it & (ROWS - 1), kb = it >> 6;
        if (it >= PRO_ITEMS) break;
        (void)nr;
        const float xv[8] = {ev0[u][0].x, ev0[u][0].y, ev0[u][0].z, ev0[u][0].w, ev0[u][1].x, ev0[u][1].y, ev0[u][1].z, ev0[u][1].w};
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = 8 * kb + i;
          z[i] = (xv[i] - sm.mean[c]) * sm.rstd[c] * sm.par[PAR_GAMMA + c] + sm.par[PAR_BETA + c];   // rows >= nr: finite, never stored
        }
        split_store8<NP>(sm.aw + a_off(r, kb), AW_PART, z);
      }
      operand_written(&sm.a_ready);
#ifdef NRM_RS_PROFILE
      if (blockIdx.x == 0 && tid == 0) g_hprof[2 * 8 + 2] += clock64() - tp0;
#endif

      // 66-output layer: pre-activation (+ bias) kept for the backward, gelu -> an (or the dot product with out_mlp.fc2)
      auto narrow_epilogue = [&](const float* bias, float* keepg, bool last, float& dot) {
        HPROF(2, 0, umma::mbar_wait(&sm.d_ready, duse & 1));
        ++duse;
        umma::fence_after_sync();
        float v[24];
        HPROF(2, 4, load_narrow(tmem_lane, q, v));
        umma::fence_before_sync();
        if (half == 0) {
          const int kb0 = narrow_kb0(q), nkb = narrow_nkb(q);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j >= nkb) break;
            const int kb = kb0 + j;
            float a[8], gl[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              a[i] = v[8 * j + i] + bias[8 * kb + i];            // columns 66-71: zero weights, zero bias -> gelu(0) = 0
              gl[i] = gelu_f(a[i]);
            }
            if (keepg != nullptr && live) {
              float* kp = keepg + (r0 + row) * HID + 8 * kb;
              if (kb < 8) {
#pragma unroll
                for (int i = 0; i < 4; ++i) reinterpret_cast<float2*>(kp)[i] = make_float2(a[2 * i], a[2 * i + 1]);
              } else {
                reinterpret_cast<float2*>(kp)[0] = make_float2(a[0], a[1]);
              }
            }
            if (last) {
#pragma unroll
              for (int i = 0; i < 8; ++i) dot = fmaf(gl[i], sm.par[PAR_WO2 + 8 * kb + i], dot);
            } else {
              split_store8<NP>(sm.an + a_off(row, kb), AN_PART, gl);
            }
          }
        }
      };
      // 264-output layer: + bias, kept, (optionally times e) -> aw
      auto wide_epilogue = [&](const float* bias, float* keepg, bool gating) {
        const int kb0 = wide_kb0(q);
        const int nkb = (half == 1 && q == 3) ? 3 : wide_nkb(q);          // the second half has 16 blocks
        const int c00 = NW0 * half + 8 * kb0;
        float4 ge[5][2];
        if (gating) {                                                     // e travels while the products run
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            ge[j][0] = ge[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < nkb && live) {
              const float4* ep = reinterpret_cast<const float4*>(e + (r0 + row) * E + c00 + 8 * j);
              ge[j][0] = __ldg(ep); ge[j][1] = __ldg(ep + 1);
            }
          }
        }
        HPROF(2, 1, umma::mbar_wait(&sm.d_ready, duse & 1));
        ++duse;
        umma::fence_after_sync();
        float v[40];
        HPROF(2, 4, load_wide(tmem_lane, q, v));
        umma::fence_before_sync();
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          if (j >= nkb) break;
          const int c0 = c00 + 8 * j;
          float o[8];
          const float4 b0 = *reinterpret_cast<const float4*>(bias + c0), b1 = *reinterpret_cast<const float4*>(bias + c0 + 4);
          o[0] = v[8 * j] + b0.x; o[1] = v[8 * j + 1] + b0.y; o[2] = v[8 * j + 2] + b0.z; o[3] = v[8 * j + 3] + b0.w;
          o[4] = v[8 * j + 4] + b1.x; o[5] = v[8 * j + 5] + b1.y; o[6] = v[8 * j + 6] + b1.z; o[7] = v[8 * j + 7] + b1.w;
          if (live) {
            if (keepg != nullptr) {
              float4* kp = reinterpret_cast<float4*>(keepg + (r0 + row) * E + c0);
              kp[0] = make_float4(o[0], o[1], o[2], o[3]);
              kp[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
            if (gating) {
              o[0] *= ge[j][0].x; o[1] *= ge[j][0].y; o[2] *= ge[j][0].z; o[3] *= ge[j][0].w;
              o[4] *= ge[j][1].x; o[5] *= ge[j][1].y; o[6] *= ge[j][1].z; o[7] *= ge[j][1].w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.f;
          }
          split_store8<NP>(sm.aw + a_off(row, (NW0 / 8) * half + kb0 + j), AW_PART, o);
        }
      };
      float dot = 0.f;
      narrow_epilogue(sm.par + PAR_BG1, keep ? a1g : nullptr, false, dot);
      HPROF(2, 3, operand_written(&sm.a_ready));
      wide_epilogue(sm.par + PAR_BG2, keep ? gateg : nullptr, true);
      HPROF(2, 3, operand_written(&sm.a_ready));
      narrow_epilogue(sm.par + PAR_BM1, keep ? a2g : nullptr, false, dot);
      HPROF(2, 3, operand_written(&sm.a_ready));
      wide_epilogue(sm.par + PAR_BM2, keep ? yg : nullptr, false);
      HPROF(2, 3, operand_written(&sm.a_ready));
      narrow_epilogue(sm.par + PAR_BO1, keep ? a3g : nullptr, true, dot);
      // out_mlp.fc2: the four column quarters of a row, added in quarter order
      if (half == 0) sm.red[q][row] = dot;
      epi_sync();
#ifdef NRM_RS_PROFILE
      if (blockIdx.x == 0 && tid == 0) g_hprof[2 * 8 + 7] += clock64() - tp0;
#endif
      if (tid < ROWS && tid < nr) logits[r0 + tid] = ((sm.red[0][tid] + sm.red[1][tid]) + sm.red[2][tid]) + sm.red[3][tid] + __ldg(P + P_OUT_FC2_B);
      epi_sync();
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == W_MMA) umma::tmem_dealloc(tmem, TMEM_COLS);
#ifdef NRM_RS_PROFILE
  if (blockIdx.x == 0 && tid == 0) g_hprof[25] += clock64() - t_entry;
#endif
}


// ---------------------------------------------------------------------------------------------------------------------
// backward: the data-gradient chain of a 64-row tile (same outputs as head_backward_kernel in nrm_head_fused.cu)
//   da3 = dr o2 gelu'(a3) | dy = da3 O1 | da2 = (dy M2) gelu'(a2) | dx = da2 M1, dgate = dx e, de = dx gate |
//   da1 = (dgate G2) gelu'(a1) | dz = da1 G1, BatchNorm partial sums of dz and dz xhat over the tile's rows
// tile partials: part_f [tile][68] = dO2[66] | do2 ; part_bn [tile][2][264] (double)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HB_F = 68;

template <int NP>
__global__ void __launch_bounds__(THREADS, 1)
head_backward_tc_kernel(const float* __restrict__ e, const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ P, const unsigned char* __restrict__ img, long long R, const float* __restrict__ dr,
                        const float* __restrict__ a1g, const float* __restrict__ gateg, const float* __restrict__ a2g,
                        const float* __restrict__ a3g, float* __restrict__ da3g, float* __restrict__ dyg, float* __restrict__ da2g,
                        float* __restrict__ dgateg, float* __restrict__ da1g, float* __restrict__ dzg, float* __restrict__ deg,
                        float* __restrict__ part_f, double* __restrict__ part_bn, unsigned int* __restrict__ arrive,
                        float* __restrict__ grads, double* __restrict__ bn_bwd_sums) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (R + ROWS - 1) / ROWS;
  const long long my_tiles = (long long)blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  if (tid == 0) {
    for (int j = 0; j < STAGES; ++j) { umma::mbar_init(&sm.full[j], 1); umma::mbar_init(&sm.empty[j], 1); }
    umma::mbar_init(&sm.a_ready, N_EPI);
    umma::mbar_init(&sm.d_ready, 1);
  }
  if (warp == W_MMA) umma::tmem_alloc(&sm.tmem_base, TMEM_COLS);
  if (tid < ROWS) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      *reinterpret_cast<uint4*>(sm.aw + p * AW_PART + a_off(tid, KE / 8 - 1)) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sm.an + p * AN_PART + a_off(tid, KH / 8 - 1)) = make_uint4(0, 0, 0, 0);
    }
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp >= N_EPI) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
    if (warp == W_LOAD) {
      if (umma::elect_one()) loader<NP>(sm, img, 5, false, my_tiles);
    } else if (warp == W_MMA) {
      if (umma::elect_one()) issuer<NP>(sm, tmem, false, my_tiles);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
    const int sp = warp & 3, q = warp >> 2, l16 = lane & 15, half = lane >> 4;
    const int row = 16 * sp + l16;
    const uint32_t tmem_lane = tmem + ((uint32_t)(32 * sp) << 16);
    for (int i = tid; i < E; i += EPI_THREADS) { sm.mean[i] = __ldg(mean + i); sm.rstd[i] = __ldg(rstd + i); }
    uint32_t duse = 0;
    for (long long t = 0; t < my_tiles; ++t) {
      const long long tile = (long long)blockIdx.x + t * gridDim.x;
      const long long r0 = tile * ROWS;
      const int nr = (int)min((long long)ROWS, R - r0);
      const bool live = row < nr;
      float* colsum = sm.par;                     // [2 row halves][72]: dO2 partial sums of the prologue
      // ---- da3 = dr o2 gelu'(a3) -> an; dO2 / do2 partial sums.  Items (row, K block), a warp = 32 consecutive rows of one block
      constexpr int PRO_ITEMS = ROWS * (NN / 8);
#pragma unroll
      for (int u = 0; u < (PRO_ITEMS + EPI_THREADS - 1) / EPI_THREADS; ++u) {
        const int it = tid + u * EPI_THREADS;
        if (it >= PRO_ITEMS) break;                 // warp-uniform (items per warp = 32)
        const int r = it & (ROWS - 1), kb = it >> 6;
        float d[8], tg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { d[i] = 0.f; tg[i] = 0.f; }
        if (r < nr) {
          const float drr = __ldg(dr + r0 + r);
          const float* ap = a3g + (r0 + r) * HID + 8 * kb;
          float av[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 x = make_float2(0.f, 0.f);
            if (8 * kb + 2 * i < HID) x = __ldg(reinterpret_cast<const float2*>(ap) + i);
            av[2 * i] = x.x; av[2 * i + 1] = x.y;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = 8 * kb + i;
            if (c < HID) {
              float gp;
              const float g = gelu_both(av[i], gp);
              d[i] = drr * __ldg(P + P_OUT_FC2_W + c) * gp;
              tg[i] = drr * g;
            }
          }
          float* dp = da3g + (r0 + r) * HID + 8 * kb;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (8 * kb + 2 * i < HID) reinterpret_cast<float2*>(dp)[i] = make_float2(d[2 * i], d[2 * i + 1]);
        }
        split_store8<NP>(sm.an + a_off(r, kb), AN_PART, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) tg[i] += __shfl_xor_sync(0xffffffffu, tg[i], o);
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) colsum[(r >> 5) * NN + 8 * kb + i] = tg[i];
        }
      }
      operand_written(&sm.a_ready);
      if (warp == 0) {                              // do2 = sum of dr over the tile's rows
        float s = (lane < nr ? __ldg(dr + r0 + lane) : 0.f) + (lane + 32 < nr ? __ldg(dr + r0 + lane + 32) : 0.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) part_f[tile * HB_F + HID] = s;
      }
      epi_sync();
      if (tid < HID) part_f[tile * HB_F + tid] = colsum[tid] + colsum[NN + tid];

      // 264-output layer: v = accumulator columns of this thread (row, column half, quarter); f(j, c0, o[8]) finishes a block of 8
      // columns starting at feature c0 and leaves in o[] what goes into aw
      auto wide_layer = [&](auto&& prefetch, auto&& finish, bool to_operand) {
        const int kb0 = wide_kb0(q);
        const int nkb = (half == 1 && q == 3) ? 3 : wide_nkb(q);
        const int c00 = NW0 * half + 8 * kb0;
        prefetch(c00, nkb);
        umma::mbar_wait(&sm.d_ready, duse & 1);
        ++duse;
        umma::fence_after_sync();
        float v[40];
        load_wide(tmem_lane, q, v);
        umma::fence_before_sync();
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          if (j >= nkb) break;
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = live ? v[8 * j + i] : 0.f;
          finish(j, c00 + 8 * j, o);
          if (to_operand) split_store8<NP>(sm.aw + a_off(row, (NW0 / 8) * half + kb0 + j), AW_PART, o);
        }
      };
      // 66-output layer: da = accumulator * gelu'(a) -> global, an
      auto narrow_layer = [&](const float* ag, float* dag) {
        const int kb0 = narrow_kb0(q), nkb = narrow_nkb(q);
        float av[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) av[i] = 0.f;
        if (half == 0 && live) {                    // the saved pre-activations travel while the products run
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j >= nkb) break;
            const float* ap = ag + (r0 + row) * HID + 8 * (kb0 + j);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (8 * (kb0 + j) + 2 * i < HID) {
                const float2 x = __ldg(reinterpret_cast<const float2*>(ap) + i);
                av[8 * j + 2 * i] = x.x; av[8 * j + 2 * i + 1] = x.y;
              }
          }
        }
        umma::mbar_wait(&sm.d_ready, duse & 1);
        ++duse;
        umma::fence_after_sync();
        float v[24];
        load_narrow(tmem_lane, q, v);
        umma::fence_before_sync();
        if (half == 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j >= nkb) break;
            const int kb = kb0 + j;
            float d[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = (live && 8 * kb + i < HID) ? v[8 * j + i] * gelu_grad_f(av[8 * j + i]) : 0.f;
            if (live) {
              float* dp = dag + (r0 + row) * HID + 8 * kb;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (8 * kb + 2 * i < HID) reinterpret_cast<float2*>(dp)[i] = make_float2(d[2 * i], d[2 * i + 1]);
            }
            split_store8<NP>(sm.an + a_off(row, kb), AN_PART, d);
          }
        }
      };
      auto store8 = [&](float* base, int c0, const float* o) {
        float4* p4 = reinterpret_cast<float4*>(base + (r0 + row) * E + c0);
        p4[0] = make_float4(o[0], o[1], o[2], o[3]);
        p4[1] = make_float4(o[4], o[5], o[6], o[7]);
      };
      auto load8 = [&](const float* base, int c0, float4 (&x)[2]) {
        const float4* p4 = reinterpret_cast<const float4*>(base + (r0 + row) * E + c0);
        x[0] = __ldg(p4); x[1] = __ldg(p4 + 1);
      };

      // dy = da3 O1 -> global, aw
      wide_layer([](int, int) {}, [&](int, int c0, float* o) { if (live) store8(dyg, c0, o); }, true);
      operand_written(&sm.a_ready);
      // da2 = (dy M2) gelu'(a2)
      narrow_layer(a2g, da2g);
      operand_written(&sm.a_ready);
      // dx = da2 M1 -> global (the BatchNorm combine kernel forms the direct path de = dx gate);  dgate = dx e -> global, aw
      {
        float4 ev[5][2];
        wide_layer(
            [&](int c00, int nkb) {
#pragma unroll
              for (int j = 0; j < 5; ++j) {
                ev[j][0] = ev[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < nkb && live) load8(e, c00 + 8 * j, ev[j]);
              }
            },
            [&](int j, int c0, float* o) {
              if (live) store8(deg, c0, o);
              o[0] *= ev[j][0].x; o[1] *= ev[j][0].y; o[2] *= ev[j][0].z; o[3] *= ev[j][0].w;
              o[4] *= ev[j][1].x; o[5] *= ev[j][1].y; o[6] *= ev[j][1].z; o[7] *= ev[j][1].w;
              if (live) store8(dgateg, c0, o);
            },
            true);
      }
      operand_written(&sm.a_ready);
      // da1 = (dgate G2) gelu'(a1)
      narrow_layer(a1g, da1g);
      operand_written(&sm.a_ready);
      // dz = da1 G1 -> global; BatchNorm partial sums over the tile's rows: 16 rows by shuffles, the four sub-partitions in shared memory
      {
        float4 ev[5][2];
        float* bnp = sm.par;                        // [4 sub-partitions][2][264]
        const uint32_t hmask = half ? 0xffff0000u : 0x0000ffffu;
        wide_layer(
            [&](int c00, int nkb) {
#pragma unroll
              for (int j = 0; j < 5; ++j) {
                ev[j][0] = ev[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < nkb && live) load8(e, c00 + 8 * j, ev[j]);
              }
              epi_sync();                           // colsum (the same scratch) has been read by everyone
            },
            [&](int j, int c0, float* o) {
              if (live) store8(dzg, c0, o);
              const float xe[8] = {ev[j][0].x, ev[j][0].y, ev[j][0].z, ev[j][0].w, ev[j][1].x, ev[j][1].y, ev[j][1].z, ev[j][1].w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float s = o[i];                                                    // rows >= nr hold zeros
                float sx = o[i] * ((xe[i] - sm.mean[c0 + i]) * sm.rstd[c0 + i]);
#pragma unroll
                for (int of = 8; of > 0; of >>= 1) {             // the two 16-lane halves run different block counts: half-warp masks
                  s += __shfl_xor_sync(hmask, s, of);
                  sx += __shfl_xor_sync(hmask, sx, of);
                }
                if (l16 == 0) { bnp[sp * 2 * E + c0 + i] = s; bnp[sp * 2 * E + E + c0 + i] = sx; }
              }
            },
            false);
      }
      epi_sync();
      for (int i = tid; i < 2 * E; i += EPI_THREADS)
        part_bn[tile * 2 * E + i] = ((double)sm.par[i] + (double)sm.par[2 * E + i]) + ((double)sm.par[4 * E + i] + (double)sm.par[6 * E + i]);
      epi_sync();
    }
    // the last CTA to arrive adds the tiles' BatchNorm partials in tile order: column sums of dz and dz * xhat over this rank's
    // rows (-> bn_bwd_sums for the encoder backward) = the gradients of bn.bias / bn.weight
    __threadfence();
    epi_sync();
    if (tid == 0) {
      const unsigned int prev = atomicAdd(arrive, 1u);
      sm.tmem_base = (prev == gridDim.x - 1) ? 1u : 0u;         // (the allocation address has been read into `tmem` by everyone)
    }
    epi_sync();
    if (sm.tmem_base != 0u) {
      __threadfence();
      for (int i = tid; i < 2 * E; i += EPI_THREADS) {
        double acc = 0.0;
#pragma unroll 8
        for (long long tl = 0; tl < ntiles; ++tl) acc += __ldcg(part_bn + tl * 2 * E + i);
        bn_bwd_sums[i] = acc;
        if (i < E) grads[P_BN_B + i] = (float)acc; else grads[P_BN_W + (i - E)] = (float)acc;
      }
      if (tid == 0) *arrive = 0u;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == W_MMA) umma::tmem_dealloc(tmem, TMEM_COLS);
}


// ---------------------------------------------------------------------------------------------------------------------
// weight gradients of the five Linear layers: dW_l [66][264] = Pn^T Qw over the candidate rows (same operands, partial layout and
// finishing kernel as head_wgrad_kernel in nrm_head_fused.cu).  blockIdx.y = layer, blockIdx.x = row chunk.
//   layer 0 out_mlp.fc1: Pn = da3,      Qw = y         | 1 mlp.fc2: Pn = gelu(a2), Qw = dy     | 2 mlp.fc1: Pn = da2, Qw = gate * e
//   layer 3 gate.fc2   : Pn = gelu(a1), Qw = dgate     | 4 gate.fc1: Pn = da1,     Qw = BN(e)
// The contraction runs over the ROWS: a tile of 64 rows is written once as two K-major bf16 tiles (Qw: 64 x 272, Pn: 64 x 80, hi | lo)
// and both are read MN-major through the descriptor, D[feature][n] += Qw^T Pn, three M blocks (features 0-127, 128-255, 208-271),
// accumulated in tensor memory over all tiles of the CTA.  Column 264 of Qw and column 66 of Pn are ones: row 264 / column 66 of D are
// the column sums, i.e. the bias gradients.  hi / lo split (three products) as in the attention kernels.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 256;
constexpr uint32_t WG_QPART = (KE / 8) * A_LBO, WG_PPART = (KH / 8) * A_LBO;
constexpr uint32_t WG_COLS = 256;                     // three accumulator blocks of 72 columns
constexpr int WG_PART = HID * E + E;                  // partial per (layer, chunk): [66][264] in (n, feature) order | bias [264]
struct WgSmem {
  __align__(128) unsigned char q[2 * WG_QPART];
  __align__(128) unsigned char p[2 * WG_PPART];
  uint64_t mbar;
  uint32_t tmem_base;
};

template <int NP>
__global__ void __launch_bounds__(WG_THREADS, 1)
head_wgrad_tc_kernel(const float* __restrict__ e, const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ P,
                     long long R, int rows_per_chunk, const float* __restrict__ a1g, const float* __restrict__ gateg,
                     const float* __restrict__ a2g, const float* __restrict__ yg, const float* __restrict__ da3g, const float* __restrict__ dyg,
                     const float* __restrict__ da2g, const float* __restrict__ dgateg, const float* __restrict__ da1g, float* __restrict__ part) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WgSmem& sm = *reinterpret_cast<WgSmem*>(smem_raw);
  const int layer = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long rbeg = (long long)blockIdx.x * rows_per_chunk, rend = min(R, rbeg + rows_per_chunk);
  const float* Psrc = layer == 0 ? da3g : layer == 1 ? a2g : layer == 2 ? da2g : layer == 3 ? a1g : da1g;
  const float* Qsrc = layer == 0 ? yg : layer == 1 ? dyg : layer == 2 ? gateg : layer == 3 ? dgateg : e;
  const bool p_gelu = layer == 1 || layer == 3, wide_bias = p_gelu;
  if (tid == 0) umma::mbar_init(&sm.mbar, 1);
  if (warp == 0) umma::tmem_alloc(&sm.tmem_base, WG_COLS);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  constexpr uint32_t ID128 = umma::make_idesc_bf16(128, NN, true, true), ID64 = umma::make_idesc_bf16(64, NN, true, true);
  uint32_t phase = 0;
  bool started = false;
  for (long long t0 = rbeg; t0 < rend; t0 += ROWS) {
    const int nr = (int)min((long long)ROWS, rend - t0);
    // ---- operand tiles of rows [t0, t0 + 64): items (row, 8-column block), consecutive lanes take consecutive rows.  The loads of
    // a batch of items are all issued before the first one is used (a load -> convert -> store loop pays one memory round trip per item)
    constexpr int QITEMS = ROWS * (KE / 8), QIT = (QITEMS + WG_THREADS - 1) / WG_THREADS;      // 9 items per thread
    constexpr int QBATCH = 5;
#pragma unroll 1
    for (int u0 = 0; u0 < QIT; u0 += QBATCH) {
      float4 qv[QBATCH][2], ev[QBATCH][2];
#pragma unroll
      for (int u = 0; u < QBATCH; ++u) {
        const int it = tid + (u0 + u) * WG_THREADS, r = it & (ROWS - 1), kb = it >> 6;
        qv[u][0] = qv[u][1] = ev[u][0] = ev[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u0 + u < QIT && it < QITEMS && r < nr && kb < E / 8) {
          const long long gi = (t0 + r) * E + 8 * kb;
          qv[u][0] = __ldg(reinterpret_cast<const float4*>(Qsrc + gi)); qv[u][1] = __ldg(reinterpret_cast<const float4*>(Qsrc + gi) + 1);
          if (layer == 2) { ev[u][0] = __ldg(reinterpret_cast<const float4*>(e + gi)); ev[u][1] = __ldg(reinterpret_cast<const float4*>(e + gi) + 1); }
        }
      }
#pragma unroll
      for (int u = 0; u < QBATCH; ++u) {
        const int it = tid + (u0 + u) * WG_THREADS, r = it & (ROWS - 1), kb = it >> 6;
        if (u0 + u >= QIT || it >= QITEMS) break;
        float v[8] = {qv[u][0].x, qv[u][0].y, qv[u][0].z, qv[u][0].w, qv[u][1].x, qv[u][1].y, qv[u][1].z, qv[u][1].w};
        if (r < nr) {
          if (kb < E / 8) {
            if (layer == 2) {
              v[0] *= ev[u][0].x; v[1] *= ev[u][0].y; v[2] *= ev[u][0].z; v[3] *= ev[u][0].w;
              v[4] *= ev[u][1].x; v[5] *= ev[u][1].y; v[6] *= ev[u][1].z; v[7] *= ev[u][1].w;
            } else if (layer == 4) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = 8 * kb + i;
                v[i] = (v[i] - __ldg(mean + c)) * __ldg(rstd + c) * __ldg(P + P_BN_W + c) + __ldg(P + P_BN_B + c);
              }
            }
          } else {
            v[0] = 1.0f;                               // column 264: ones -> row 264 of D = column sums of Pn
          }
        }
        split_store8<NP>(sm.q + a_off(r, kb), WG_QPART, v);
      }
    }
    for (int it = tid; it < ROWS * (KH / 8); it += WG_THREADS) {
      const int r = it & (ROWS - 1), kb = it >> 6;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      if (r < nr) {
        const float* ap = Psrc + (t0 + r) * HID + 8 * kb;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (8 * kb + 2 * i < HID) {
            const float2 x = __ldg(reinterpret_cast<const float2*>(ap) + i);
            v[2 * i] = p_gelu ? gelu_f(x.x) : x.x; v[2 * i + 1] = p_gelu ? gelu_f(x.y) : x.y;
          }
        if (kb == HID / 8) v[HID & 7] = 1.0f;          // column 66: ones -> column 66 of D = column sums of Qw
      }
      split_store8<NP>(sm.p + a_off(r, kb), WG_PPART, v);
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0 && umma::elect_one()) {
      umma::fence_after_sync();
      const uint32_t qa = umma::smem_u32(sm.q), pa = umma::smem_u32(sm.p);
#pragma unroll
      for (int pr = 0; pr < n_products(NP); ++pr) {
        const uint32_t qo = prod_a(NP, pr) * WG_QPART, po = prod_b(NP, pr) * WG_PPART;
        // MN-major views of the K-major tiles: LBO (between K = row groups of 8) = 128, SBO (between 8-feature groups) = A_LBO
        const uint64_t db = umma::make_desc(pa + po, 128, A_LBO);
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
          const uint32_t f0 = blk == 0 ? 0 : blk == 1 ? 128 : 208;
          const uint64_t da = umma::make_desc(qa + qo + (f0 / 8) * A_LBO, 128, A_LBO);
#pragma unroll
          for (int ks = 0; ks < ROWS / 16; ++ks)       // 16 rows per K step: two row groups of 8 = 256 bytes
            umma::mma_bf16(tmem + NN * blk, da + (uint64_t)(ks * (256 >> 4)), db + (uint64_t)(ks * (256 >> 4)), blk == 2 ? ID64 : ID128,
                           (started || pr > 0 || ks > 0) ? 1u : 0u);
        }
      }
      umma::mma_commit(&sm.mbar);
    }
    started = true;
    umma::mbar_wait(&sm.mbar, phase);                  // the tiles are rewritten by the next iteration
    phase ^= 1;
  }
  // ---- per-CTA partial: out[n][feature] = D[feature][n]; bias = column sums
  umma::fence_after_sync();
  float* out = part + ((long long)layer * gridDim.x + blockIdx.x) * WG_PART;
  const int sp = warp & 3, half = warp >> 2;           // 8 warps: sub-partition, half of the 72 columns
  float v[40];
#pragma unroll 1
  for (int blk = 0; blk < 3; ++blk) {
    const uint32_t a = tmem + NN * blk + 36 * half + ((uint32_t)(32 * sp) << 16);
    umma::tmem_ld32(a, v);
    umma::tmem_ld8(a + 32, v + 32);                    // (4 columns more than this half needs)
    int f;
    bool ok;
    if (blk < 2) { f = 128 * blk + 32 * sp + lane; ok = true; }
    else { f = 208 + 16 * sp + lane; ok = lane < 16 && f >= 256; }       // M = 64: lanes 0-15 of each sub-partition; features 256-264 are new
    if (ok) {
#pragma unroll
      for (int j = 0; j < 36; ++j) {
        const int n = 36 * half + j;
        const float x = started ? v[j] : 0.f;
        if (f < E) {
          if (n < HID) out[n * E + f] = x;
          else if (n == HID && wide_bias) out[HID * E + f] = x;
        } else if (f == E && !wide_bias && n < HID) {
          out[HID * E + n] = x;
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, WG_COLS);
}

}  // namespace htc

int headprof_read(long long* host_out32) {
#ifdef NRM_RS_PROFILE
  long long zero[32] = {0};
  NRM_CUDA(cudaDeviceSynchronize());
  NRM_CUDA(cudaMemcpyFromSymbol(host_out32, htc::g_hprof, sizeof(zero)));
  NRM_CUDA(cudaMemcpyToSymbol(htc::g_hprof, zero, sizeof(zero)));
  return NRM_OK;
#else
  (void)host_out32;
  set_error("nrm_debug_headprof: library built without -DNRM_RS_PROFILE");
  return NRM_EINVAL;
#endif
}

size_t head_tc_image_bytes() { return (size_t)10 * htc::IMG_STRIDE; }

static int head_np(int precision) { return precision == NRM_PRECISION_BF16 ? 2 : 3; }

// weight images of the forward (and, when training, the backward) chain: weights only, enqueued ahead of the encoder
int launch_head_images_tc(const float* P, Workspace& w, int precision, bool with_backward, cudaStream_t s) {
  unsigned char* img = reinterpret_cast<unsigned char*>(w.head_img);
  const dim3 grid(10, with_backward ? 10 : 5);
  if (head_np(precision) == 2) launch_pdl(htc::head_image_kernel<2>, grid, dim3(256), 0, s, P, img, 0, w.head_arrive);
  else launch_pdl(htc::head_image_kernel<3>, grid, dim3(256), 0, s, P, img, 0, w.head_arrive);
  NRM_LAUNCH_CHECK("head_image_kernel");
  return NRM_OK;
}

template <int NP>
static int launch_fwd(const float* P, Workspace& w, float* run_mean, float* run_var, long long* nbt, int training, int keep,
                      const double* bn_sums, long long bn_rows, float* logits, cudaStream_t s) {
  static DeviceOnce configured;
  if (configured.first_time())
    NRM_CUDA(cudaFuncSetAttribute(htc::head_forward_tc_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(htc::Smem)));
  const long long ntiles = (w.R + htc::ROWS - 1) / htc::ROWS;
  const int grid = (int)min(ntiles, (long long)sm_count());
  launch_pdl(htc::head_forward_tc_kernel<NP>, dim3(grid), dim3(htc::THREADS), sizeof(htc::Smem), s, w.e, bn_sums, bn_rows, training, run_mean,
             run_var, nbt, w.mean, w.rstd, P, reinterpret_cast<const unsigned char*>(w.head_img), w.R, keep, w.a1, w.gate, w.a2, w.y, w.a3, logits);
  NRM_LAUNCH_CHECK("head_forward_tc_kernel");
  return NRM_OK;
}

int launch_head_forward_tc(const float* P, Workspace& w, int precision, float* run_mean, float* run_var, long long* nbt, int training, int keep,
                           const double* bn_sums, long long bn_rows, float* logits, cudaStream_t s) {
  return head_np(precision) == 2 ? launch_fwd<2>(P, w, run_mean, run_var, nbt, training, keep, bn_sums, bn_rows, logits, s)
                                 : launch_fwd<3>(P, w, run_mean, run_var, nbt, training, keep, bn_sums, bn_rows, logits, s);
}


template <int NP>
static int launch_bwd(const float* P, Workspace& w, const float* dlogits, float* G, cudaStream_t s) {
  static DeviceOnce configured;
  if (configured.first_time())
    NRM_CUDA(cudaFuncSetAttribute(htc::head_backward_tc_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(htc::Smem)));
  const long long ntiles = (w.R + htc::ROWS - 1) / htc::ROWS;
  const int grid = (int)min(ntiles, (long long)sm_count());
  launch_pdl(htc::head_backward_tc_kernel<NP>, dim3(grid), dim3(htc::THREADS), sizeof(htc::Smem), s, w.e, w.mean, w.rstd, P,
             reinterpret_cast<const unsigned char*>(w.head_img), w.R, dlogits, w.a1, w.gate, w.a2, w.a3, w.da3, w.dy, w.da2, w.dgate, w.da1, w.dz,
             w.de, w.head_part_f, w.head_part_bn, w.head_arrive, G, w.bn_bwd_sums);
  NRM_LAUNCH_CHECK("head_backward_tc_kernel");
  return NRM_OK;
}


// weight gradients on the tensor cores: partials in head_wgrad_kernel's layout; returns the number of row chunks
int launch_head_wgrad_tc(const float* P, Workspace& w, int precision, int rows_per_chunk, int nchunks, cudaStream_t s) {
  static DeviceOnce configured;
  if (configured.first_time()) {
    NRM_CUDA(cudaFuncSetAttribute(htc::head_wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(htc::WgSmem)));
    NRM_CUDA(cudaFuncSetAttribute(htc::head_wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(htc::WgSmem)));
  }
  if (precision == NRM_PRECISION_BF16)
    launch_pdl(htc::head_wgrad_tc_kernel<1>, dim3(nchunks, 5), dim3(htc::WG_THREADS), sizeof(htc::WgSmem), s, w.e, w.mean, w.rstd, P, w.R, rows_per_chunk,
               w.a1, w.gate, w.a2, w.y, w.da3, w.dy, w.da2, w.dgate, w.da1, w.head_part_w);
  else
    launch_pdl(htc::head_wgrad_tc_kernel<2>, dim3(nchunks, 5), dim3(htc::WG_THREADS), sizeof(htc::WgSmem), s, w.e, w.mean, w.rstd, P, w.R, rows_per_chunk,
               w.a1, w.gate, w.a2, w.y, w.da3, w.dy, w.da2, w.dgate, w.da1, w.head_part_w);
  NRM_LAUNCH_CHECK("head_wgrad_tc_kernel");
  return NRM_OK;
}

// data-gradient chain; per-tile partials for head_tc_tiles(R) tiles (head_grad_finish_kernel sums them)
int head_tc_tiles(long long R) { return (int)((R + htc::ROWS - 1) / htc::ROWS); }
int launch_head_backward_dgrad_tc(const float* P, Workspace& w, int precision, const float* dlogits, float* G, cudaStream_t s) {
  return head_np(precision) == 2 ? launch_bwd<2>(P, w, dlogits, G, s) : launch_bwd<3>(P, w, dlogits, G, s);
}

}  // namespace nrm
