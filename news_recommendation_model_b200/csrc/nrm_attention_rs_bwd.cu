// Backward of the pairwise MLP attention in the row-stacked tensor-core formulation (forward: nrm_attention_rs.cu).
// Reference: autograd of PointwiseAttentionExpanded.forward (models/attention_model.py:52-97) and of the pooling at
// models/user_invariant_interest_model.py:83-87.
//
// Operand row of pair (c, h):  a[(c,h)] = [t_c (.) h | h]  (K = 128),  W = [Wd | A]  (64 x 128),  hid = a W^T + tp_c.
// Given dP_c = dL/dpooled_c:
//     ds[(c,h)]   = dP_c . h                                   dhid[(c,h)][j] = ds w2[j] gelu'(hid[(c,h)][j])
//     dW^T[k'][j] = sum_rows a[r][k'] dhid[r][j]               (rows 0-63 of the result: dWd^T, rows 64-127: dA^T)
//     dtp_c[j]    = sum_h dhid[(c,h)][j]                       dw2[j] = sum_rows ds gelu(hid[r][j]),   db2 = sum_rows ds
// and, for the label branch only (its inputs are activations), with  da = dhid W  (X = da[:, :64], Y = da[:, 64:]):
//     dh[h] += sum_c ( t_c (.) X[(c,h)] + Y[(c,h)] + s[c][h] dP_c ),        dt_c = sum_h h (.) X[(c,h)]
//
// Two kernels:
//   attention_backward_rs_kernel   everything that is a sum over rows (both branches).  The rows of a unit (impression, <= 8
//       candidates) are stacked into 128-row tiles exactly as in the forward; per tile: operand rows -> tensor memory (hid product,
//       A operand in TMEM), a copy of them and dhid as K-major bf16 tiles in shared memory, which the weight-gradient product
//       reads MN-major (contraction over the rows) and ACCUMULATES IN TENSOR MEMORY over all tiles of the CTA (one partial per CTA at the
//       end, no per-item epilogue); dtp is the product of dhid with a 0/1 row-to-candidate selection tile.  For the label branch
//       the dhid tiles and the scores are also written to global memory (tile images) for the second kernel.
//   attention_input_grad_rs_kernel  (label branch) da = dhid W on the tensor cores from those tile images, then dh / dt with two
//       fixed-order reductions through shared memory.
// Warp roles, barriers and the work split follow the forward kernel (see there).
#include "nrm_attention_rs.cuh"

namespace nrm {
namespace rsb {

using namespace rs;

constexpr int NST = 2;                        // history-chunk stages
constexpr uint32_t LBO128 = 2048;             // K-major tile of 128 rows: (r, 8 kb) at kb * 2048 + (r / 8) * 128 + (r % 8) * 16
constexpr uint32_t AROW_PART = 16 * LBO128;   // 32 KB per part: [128 r][128 k']
constexpr uint32_t DHID_PART = 8 * LBO128;    // 16 KB per part: [128 r][64 j]
constexpr uint32_t SEL_BYTES = 16 * 128;      // [8 slots][128 r] K-major: (slot, 8 rb) at rb * 128 + slot * 16

// per-CTA partial sums (floats), the layout attention_finish_kernel (nrm_attention_tc.cu) sums: dA^T [64 k][64 j] | dWd^T | dw2 | db2
constexpr int TCP_DA = 0, TCP_DWD = 4096, TCP_DW2 = 2 * 4096, TCP_DB2 = 2 * 4096 + 64;

__device__ __forceinline__ uint32_t tile_off(int r, int kb) { return (uint32_t)kb * LBO128 + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u; }
__device__ __forceinline__ void tmem_ld8u(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

struct StageB {
  float hf[HCH * HF_STRIDE];                       // fp32 history rows of the chunk
  __align__(16) float tv[CG][64];                  // candidate vectors t_c
  __align__(16) float tpv[CG][64];                 // tp_c
  __align__(16) float dpv[CG][64];                 // dP_c = dL / dpooled_c
};

template <int NP>
struct SmemB {
  __align__(128) unsigned char W[2][W_TILE];       // this branch's weights, hi | lo
  __align__(128) unsigned char arow[2][AROW_PART]; // operand rows of the tile, hi | lo (A of the weight-gradient product, read MN-major)
  __align__(128) unsigned char dhid[2][DHID_PART]; // dhid of the tile, hi | lo
  __align__(128) unsigned char sel[SEL_BYTES];     // row -> candidate slot selection (0 / 1)
  StageB st[NST];
  float ds_part[4][2][128];                        // [tile % 4][K half][row]  partial dP . h
  float red[16][17];                               // final reduction of dw2 / db2
  float w2[64];
  float b2;
  __align__(16) float zrow[64];
  uint64_t stage_full[NST], stage_empty[NST], a_full[2], a_empty[2], d_full[2], d_empty[2], as_full, e1_done, dw_done, dtp_full[2], dtp_empty[2], wbar, fin;
  uint32_t tmem_base;
};

// TMEM columns
constexpr uint32_t C_D = 0;            // 2 x 64   hid accumulators
constexpr uint32_t C_A = 128;          // 2 x (64 hi + 64 lo)   operand rows (K = 128 bf16 = 64 columns per part), double-buffered
constexpr uint32_t C_DW = 384;         // 64   dW^T accumulator [128 k'][64 j], whole CTA
constexpr uint32_t C_DTP = 448;        // 2 x 8   dtp^T [64 j][8 slots] per unit
constexpr uint32_t B_TMEM_COLS = 512;

struct GeoB {
  int B, H, C, G, nchunks;
  __device__ __forceinline__ void unit(int u, int& b, int& c0, int& ncg) const {
    b = u / G;
    c0 = (u - b * G) * CG;
    ncg = min(CG, C - c0);
  }
  __device__ __forceinline__ void chunk(int ci, int ncg, int& h0, int& hl, int& rows, int& ntiles) const {
    h0 = ci * HCH;
    hl = min(HCH, H - h0);
    rows = ncg * hl;
    ntiles = (rows + 127) >> 7;
  }
  // tiles of all units in front of unit u (every impression has the same tile count; the groups differ only in their candidate count)
  __device__ __forceinline__ long long tile_base(int u) const {
    const int b = u / G, g = u - b * G;
    long long per_imp = 0, before = 0;
    for (int gg = 0; gg < G; ++gg) {
      const int ncg = min(CG, C - gg * CG);
      int t = 0;
      for (int ci = 0; ci < nchunks; ++ci) t += (ncg * min(HCH, H - ci * HCH) + 127) >> 7;
      if (gg < g) before += t;
      per_imp += t;
    }
    return (long long)b * per_imp + before;
  }
};

struct ChunkIterB {
  int u, ci, b, c0, ncg, h0, hl, rows, ntiles;
  __device__ __forceinline__ void set(const GeoB& g, int u_, int ci_) {
    u = u_; ci = ci_;
    g.unit(u, b, c0, ncg);
    g.chunk(ci, ncg, h0, hl, rows, ntiles);
  }
  __device__ __forceinline__ void next(const GeoB& g) {
    if (ci + 1 < g.nchunks) set(g, u, ci + 1); else set(g, u + 1, 0);
  }
};

// flattened (unit, chunk, tile) sequence of a CTA
struct TileIterB {
  int u, ci, ti, b, c0, ncg, h0, hl, rows, ntiles;
  uint32_t chunk_seq, tile_seq;
  bool live;
  __device__ __forceinline__ void load(const GeoB& g) { g.unit(u, b, c0, ncg); g.chunk(ci, ncg, h0, hl, rows, ntiles); }
  __device__ __forceinline__ void init(const GeoB& g, int u0, int u1) {
    u = u0; ci = 0; ti = 0; chunk_seq = 0; tile_seq = 0;
    live = u0 < u1;
    if (live) load(g);
  }
  __device__ __forceinline__ void next(const GeoB& g, int u1) {
    ++tile_seq;
    if (++ti < ntiles) return;
    ti = 0; ++chunk_seq;
    if (++ci == g.nchunks) { ci = 0; ++u; }
    live = u < u1;
    if (live) load(g);
  }
};

__device__ __forceinline__ void issue_stage_b(StageB& st, const float* __restrict__ rows, const float* __restrict__ e, const float* __restrict__ de,
                                              int toff, int poff, const float* __restrict__ tpg, int H, int C, int b, int c0, int ncg, int h0, int hl) {
  const int lane = threadIdx.x & 31;
  const float* src = rows + ((long long)b * H + h0) * 64;
  for (int i = lane; i < hl * 16; i += 32) cp_async16(&st.hf[(i >> 4) * HF_STRIDE + 4 * (i & 15)], src + 4 * i);
  for (int i = lane; i < ncg * 16; i += 32) {
    const int c = i >> 4, q = i & 15;
    const long long rc = (long long)b * C + c0 + c;
    cp_async16(&st.tv[c][4 * q], e + rc * E + toff + 4 * q);
    cp_async16(&st.tpv[c][4 * q], tpg + rc * 64 + 4 * q);
    cp_async16(&st.dpv[c][4 * q], de + rc * E + poff + 4 * q);
  }
  cp_async_commit();
}

// EXPORT: also write every tile's dhid image (hi | lo, DHID_PART bytes each) and the four partial scores per row to global memory
template <int SPLIT, bool EXPORT>
__global__ void __launch_bounds__(THREADS, 1)
attention_backward_rs_kernel(const float* __restrict__ rows_g, int branch, int B, int H, int C, const unsigned char* __restrict__ img,
                             const float* __restrict__ tpg, const float* __restrict__ e, const float* __restrict__ de,
                             float* __restrict__ dtp, float* __restrict__ part, unsigned char* __restrict__ dhid_g, float* __restrict__ sc_g) {
  pdl_wait();
  pdl_trigger();
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  constexpr int NT = SPLIT == 3 ? 3 : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemB<NP>& sm = *reinterpret_cast<SmemB<NP>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int toff = branch == 0 ? E_XT : E_PCAT, poff = branch == 0 ? E_LAB : E_TI;
  const unsigned char* img_b = img + (size_t)branch * IMG_BRANCH_BYTES;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { umma::mbar_init(&sm.stage_full[i], 1); umma::mbar_init(&sm.stage_empty[i], N_PROD + N_EPI); }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&sm.d_full[i], 1); umma::mbar_init(&sm.d_empty[i], N_EPI);
      umma::mbar_init(&sm.dtp_full[i], 1); umma::mbar_init(&sm.dtp_empty[i], 4);
      umma::mbar_init(&sm.a_full[i], N_PROD); umma::mbar_init(&sm.a_empty[i], 1);
    }
    umma::mbar_init(&sm.as_full, N_PROD); umma::mbar_init(&sm.e1_done, N_EPI); umma::mbar_init(&sm.dw_done, 1);
    umma::mbar_init(&sm.wbar, 1); umma::mbar_init(&sm.fin, 1);
    const uint32_t bar = umma::smem_u32(&sm.wbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2u * W_TILE) : "memory");
    for (int p = 0; p < 2; ++p)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(umma::smem_u32(sm.W[p])),
                   "l"(img_b + (size_t)p * W_TILE), "r"(W_TILE), "r"(bar) : "memory");
  }
  if (tid < 64) {
    const float* tail = reinterpret_cast<const float*>(img_b + 2 * W_TILE);
    sm.w2[tid] = __ldg(tail + tid);
    if (tid == 0) sm.b2 = __ldg(tail + 64);
    sm.zrow[tid] = 0.f;
  }
  if (warp == 0) umma::tmem_alloc(&sm.tmem_base, B_TMEM_COLS);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  GeoB g;
  g.B = B; g.H = H; g.C = C; g.G = (C + CG - 1) / CG; g.nchunks = (H + HCH - 1) / HCH;
  const long long U = (long long)B * g.G;
  const int u0 = (int)(U * blockIdx.x / gridDim.x), u1 = (int)(U * (blockIdx.x + 1) / gridDim.x);

  if (warp == W_LOAD) {
    // =========================================== loader ===========================================
    const int total = (u1 - u0) * g.nchunks;
    ChunkIterB it_issue, it_fin;
    int n_issued = 0;
    auto issue = [&]() {
      const int s = n_issued % NST;
      umma::mbar_wait(&sm.stage_empty[s], ((n_issued / NST) & 1) ^ 1);
      issue_stage_b(sm.st[s], rows_g, e, de, toff, poff, tpg, H, C, it_issue.b, it_issue.c0, it_issue.ncg, it_issue.h0, it_issue.hl);
      ++n_issued;
      if (n_issued < total) it_issue.next(g);
    };
    if (total > 0) { it_issue.set(g, u0, 0); it_fin.set(g, u0, 0); }
    while (n_issued < total && n_issued < NST - 1) issue();
    for (int k = 0; k < total; ++k) {
      if (n_issued - k - 1 >= 1) cp_async_wait<1>(); else cp_async_wait<0>();
      arrive_warp(&sm.stage_full[k % NST]);
      if (k + 1 < total) it_fin.next(g);
      if (n_issued < total) issue();
    }
  } else if (warp == W_MMA) {
    // =========================================== MMA issuer ===========================================
    if (umma::elect_one()) {
      umma::mbar_wait(&sm.wbar, 0);
      constexpr uint32_t IDESC_HID = umma::make_idesc_bf16(128, 64);
      constexpr uint32_t IDESC_DW = umma::make_idesc_bf16(128, 64, true, true);    // arow^T x dhid^T: both tiles read MN-major
      constexpr uint32_t IDESC_DTP = umma::make_idesc_bf16(64, 8, true, false);    // dhid^T x sel
      const uint64_t wdesc0 = umma::make_desc(umma::smem_u32(sm.W[0]), 1024, 128);
      const umma::Operand arow_mn = umma::make_operand(umma::smem_u32(sm.arow[0]), 128, LBO128, 256, AROW_PART);
      const umma::Operand dhid_mn = umma::make_operand(umma::smem_u32(sm.dhid[0]), 128, LBO128, 256, DHID_PART);
      const uint64_t sel_desc = umma::make_desc(umma::smem_u32(sm.sel), 128, 128);
      uint32_t tile_seq = 0, unit_seq = 0;
      bool dw_started = false;
      // second product group of the previous tile (weight gradient, dtp), issued after the next tile's hid product
      bool pend = false; uint32_t p_seq = 0, p_ps = 0, p_pph = 0; bool p_first = false, p_last = false;
      auto second_group = [&]() {
        umma::mbar_wait(&sm.e1_done, p_seq & 1);
        umma::mbar_wait(&sm.as_full, p_seq & 1);
        if (p_first) umma::mbar_wait(&sm.dtp_empty[p_ps], p_pph ^ 1);
        umma::fence_after_sync();
        umma::mma_product<SPLIT, 8>(tmem + C_DW, arow_mn, dhid_mn, IDESC_DW, dw_started);
        dw_started = true;
#pragma unroll
        for (int t = 0; t < NP; ++t)                         // the selection tile is exact in bf16: (hi, sel) and (lo, sel) only
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma::mma_bf16(tmem + C_DTP + 8 * p_ps, dhid_mn.desc + (uint64_t)(t * dhid_mn.part16 + ks * dhid_mn.kstep16), sel_desc + (uint64_t)(ks * 16),
                           IDESC_DTP, (!p_first || t > 0 || ks > 0) ? 1u : 0u);
        umma::mma_commit(&sm.dw_done);
        if (p_last) umma::mma_commit(&sm.dtp_full[p_ps]);
        pend = false;
      };
      for (int u = u0; u < u1; ++u, ++unit_seq) {
        int b, c0, ncg; g.unit(u, b, c0, ncg);
        bool first_of_unit = true;
        for (int ci = 0; ci < g.nchunks; ++ci) {
          int h0, hl, rows, ntiles; g.chunk(ci, ncg, h0, hl, rows, ntiles);
          for (int ti = 0; ti < ntiles; ++ti, ++tile_seq) {
            const uint32_t ds = tile_seq & 1, dph = (tile_seq >> 1) & 1;
            umma::mbar_wait(&sm.a_full[ds], dph);
            umma::mbar_wait(&sm.d_empty[ds], dph ^ 1);
            umma::fence_after_sync();
            const uint32_t d = tmem + C_D + 64 * ds, a = tmem + C_A + 128 * ds;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
              const uint32_t ap = (t == 2) ? 64u : 0u;
              const uint64_t wd = wdesc0 + (uint64_t)((t == 1 ? 1 : 0) * (W_TILE / 16));
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) mma_bf16_ts(d, a + ap + 8 * ks, wd + (uint64_t)(ks * 128), IDESC_HID, (t > 0 || ks > 0) ? 1u : 0u);
            }
            umma::mma_commit(&sm.a_empty[ds]);
            umma::mma_commit(&sm.d_full[ds]);
            if (pend) second_group();
            pend = true; p_seq = tile_seq; p_ps = unit_seq & 1; p_pph = (unit_seq >> 1) & 1;
            p_first = first_of_unit; p_last = (ci == g.nchunks - 1 && ti == ntiles - 1);
            first_of_unit = false;
          }
        }
      }
      if (pend) second_group();
      umma::mma_commit(&sm.fin);
    }
  } else if (warp < W_EPI) {
    // =========================================== producers ===========================================
    // Software-pipelined over the flattened tile sequence: phase 1 (operand row -> tensor memory, partial ds) of tile t + 1 runs BEFORE
    // phase 2 (shared-memory copies for the weight-gradient product) of tile t, which has to wait for the previous tile's second
    // product group; the operand rows are double-buffered in tensor memory, so the next hid product never waits for that chain.
    const int sp = warp & 3, khalf = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(32 * sp) << 16;
    const int r = 32 * sp + lane;
    TileIterB p1, p2;
    p1.init(g, u0, u1);
    p2 = p1;
    auto phase1 = [&](const TileIterB& it) {
      const uint32_t s = it.chunk_seq % NST;
      const StageB& st = sm.st[s];
      if (it.ti == 0) umma::mbar_wait(&sm.stage_full[s], (it.chunk_seq / NST) & 1);
      const uint32_t as = it.tile_seq & 1;
      umma::mbar_wait(&sm.a_empty[as], ((it.tile_seq >> 1) & 1) ^ 1);
      umma::fence_after_sync();
      const int rg = it.ti * 128 + r;
      const bool valid = rg < it.rows;
      const int cl = valid ? __float2int_rz(((float)rg + 0.5f) * (1.0f / (float)it.hl)) : 0, hloc = valid ? rg - cl * it.hl : 0;
      const uint32_t a_hi = tmem + C_A + 128 * as + lane_sel, a_lo = a_hi + 64;
      const float* hrow = valid ? &st.hf[hloc * HF_STRIDE] : sm.zrow;
      const float* trow = &st.tv[cl][0];
      const float* dprow = &st.dpv[cl][0];
      float dsp = 0.f;
#pragma unroll
      for (int kq = 0; kq < 2; ++kq) {
        const int kp = 2 * khalf + kq;
        uint32_t ph_hi[8], ph_lo[8], hh[8], hl_[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int kb = 2 * kp + q;
          const float4 h0v = *reinterpret_cast<const float4*>(hrow + 8 * kb), h1v = *reinterpret_cast<const float4*>(hrow + 8 * kb + 4);
          const float4 t0v = *reinterpret_cast<const float4*>(trow + 8 * kb), t1v = *reinterpret_cast<const float4*>(trow + 8 * kb + 4);
          const float4 p0v = *reinterpret_cast<const float4*>(dprow + 8 * kb), p1v = *reinterpret_cast<const float4*>(dprow + 8 * kb + 4);
          const float hv[8] = {h0v.x, h0v.y, h0v.z, h0v.w, h1v.x, h1v.y, h1v.z, h1v.w};
          const float tv8[8] = {t0v.x, t0v.y, t0v.z, t0v.w, t1v.x, t1v.y, t1v.z, t1v.w};
          const float pv8[8] = {p0v.x, p0v.y, p0v.z, p0v.w, p1v.x, p1v.y, p1v.z, p1v.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) dsp = fmaf(hv[i], pv8[i], dsp);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (NP == 2) {
              split2(hv[2 * i] * tv8[2 * i], hv[2 * i + 1] * tv8[2 * i + 1], ph_hi[4 * q + i], ph_lo[4 * q + i]);
              split2(hv[2 * i], hv[2 * i + 1], hh[4 * q + i], hl_[4 * q + i]);
            } else {
              ph_hi[4 * q + i] = pack_bf16(hv[2 * i] * tv8[2 * i], hv[2 * i + 1] * tv8[2 * i + 1]);
              hh[4 * q + i] = pack_bf16(hv[2 * i], hv[2 * i + 1]);
            }
          }
        }
        tmem_st8(a_hi + 8 * kp, ph_hi);
        tmem_st8(a_hi + 32 + 8 * kp, hh);
        if (NP == 2) { tmem_st8(a_lo + 8 * kp, ph_lo); tmem_st8(a_lo + 32 + 8 * kp, hl_); }
      }
      sm.ds_part[it.tile_seq & 3][khalf][r] = dsp;          // rows past the end read the zero row: 0
      tmem_st_wait();
      umma::fence_before_sync();
      arrive_warp(&sm.a_full[as]);
      if (it.ti == it.ntiles - 1) arrive_warp(&sm.stage_empty[s]);      // last read of this chunk's stage by this warp
    };
    auto phase2 = [&](const TileIterB& it) {
      // once the previous tile's weight-gradient product has read them, refill the shared-memory copies: this thread's operand row
      // (read back from tensor memory: no recomputation) and the row -> candidate selection entries
      const uint32_t as = it.tile_seq & 1;
      const int rg = it.ti * 128 + r;
      const bool valid = rg < it.rows;
      const int cl = valid ? __float2int_rz(((float)rg + 0.5f) * (1.0f / (float)it.hl)) : 0;
      const uint32_t a_hi = tmem + C_A + 128 * as + lane_sel;
      umma::mbar_wait(&sm.dw_done, (it.tile_seq & 1) ^ 1);
      umma::fence_after_sync();
#pragma unroll
      for (int kq = 0; kq < 2; ++kq) {
        const int kp = 2 * khalf + kq;
#pragma unroll
        for (int pt = 0; pt < NP; ++pt) {
          uint32_t w8[8];
          tmem_ld8u(a_hi + 64 * pt + 8 * kp, w8);                          // t (.) h half: k' blocks 2 kp, 2 kp + 1
          *reinterpret_cast<uint4*>(sm.arow[pt] + tile_off(r, 2 * kp)) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
          *reinterpret_cast<uint4*>(sm.arow[pt] + tile_off(r, 2 * kp + 1)) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
          tmem_ld8u(a_hi + 64 * pt + 32 + 8 * kp, w8);                     // h half: k' blocks 8 + 2 kp, 9 + 2 kp
          *reinterpret_cast<uint4*>(sm.arow[pt] + tile_off(r, 8 + 2 * kp)) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
          *reinterpret_cast<uint4*>(sm.arow[pt] + tile_off(r, 9 + 2 * kp)) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
        }
      }
      if (khalf == 0) {
        unsigned char* sp_ = sm.sel + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 2u;
#pragma unroll
        for (int c = 0; c < CG; ++c) *reinterpret_cast<unsigned short*>(sp_ + c * 16) = (valid && cl == c) ? (unsigned short)0x3f80 : (unsigned short)0;
      }
      umma::fence_before_sync();
      umma::fence_async_smem();
      arrive_warp(&sm.as_full);
    };
    if (p1.live) { phase1(p1); p1.next(g, u1); }
    while (p2.live) {
      if (p1.live) { phase1(p1); p1.next(g, u1); }
      phase2(p2);
      p2.next(g, u1);
    }
  } else {
    // =========================================== epilogue ===========================================
    const int sp = warp & 3, cq = (warp - W_EPI) >> 2;
    const uint32_t lane_sel = (uint32_t)(32 * sp) << 16;
    uint32_t chunk_seq = 0, tile_seq = 0, unit_seq = 0;
    f32x2 dw2a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dw2a[i] = pk(0.f, 0.f);
    float db2a = 0.f;
    f32x2 w2p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w2p[i] = pk(sm.w2[16 * cq + 2 * i], sm.w2[16 * cq + 2 * i + 1]);
    const float b2 = sm.b2;
    bool pend = false; uint32_t p_ps = 0, p_pph = 0; int p_b = 0, p_c0 = 0, p_ncg = 0;
    auto write_dtp = [&]() {
      if (cq == 0) {
        umma::mbar_wait(&sm.dtp_full[p_ps], p_pph);
        umma::fence_after_sync();
        float v[8];
        umma::tmem_ld8(tmem + C_DTP + 8 * p_ps + lane_sel, v);
        umma::fence_before_sync();
        arrive_warp(&sm.dtp_empty[p_ps]);
        if (lane < 16) {                                   // M = 64 accumulator: j = 16 sp + lane
          float* dst = dtp + ((long long)p_b * C + p_c0) * 64 + 16 * sp + lane;
#pragma unroll
          for (int c = 0; c < CG; ++c)
            if (c < p_ncg) dst[c * 64] = v[c];
        }
      }
      pend = false;
    };
    const long long tile0 = EXPORT ? g.tile_base(u0) : 0;
    for (int u = u0; u < u1; ++u, ++unit_seq) {
      int b, c0, ncg; g.unit(u, b, c0, ncg);
      for (int ci = 0; ci < g.nchunks; ++ci, ++chunk_seq) {
        int h0, hl, rows, ntiles; g.chunk(ci, ncg, h0, hl, rows, ntiles);
        const uint32_t s = chunk_seq % NST, ph = (chunk_seq / NST) & 1;
        StageB& st = sm.st[s];
        const float inv_hl = 1.0f / (float)hl;
        umma::mbar_wait(&sm.stage_full[s], ph);
        for (int ti = 0; ti < ntiles; ++ti, ++tile_seq) {
          const uint32_t ds = tile_seq & 1, dph = (tile_seq >> 1) & 1;
          umma::mbar_wait(&sm.d_full[ds], dph);
          umma::fence_after_sync();
          float x[16];
          umma::tmem_ld16(tmem + C_D + 64 * ds + 16 * cq + lane_sel, x);
          umma::fence_before_sync();
          arrive_warp(&sm.d_empty[ds]);
          const int r = 32 * sp + lane, rg = ti * 128 + r;
          const bool valid = rg < rows;
          const int cl = valid ? __float2int_rz(((float)rg + 0.5f) * inv_hl) : 0;
          const float* tp = &st.tpv[cl][16 * cq];
          const float dsr = valid ? sm.ds_part[tile_seq & 3][0][r] + sm.ds_part[tile_seq & 3][1][r] : 0.f;
          const f32x2 ds2 = pk(dsr, dsr);
          uint32_t dh_hi[8], dh_lo[8];
          f32x2 sacc2 = pk(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 t2 = *reinterpret_cast<const float2*>(tp + 2 * i);
            f32x2 gg, gp;
            gelu_both2(add2(pk(x[2 * i], x[2 * i + 1]), pk(t2.x, t2.y)), gg, gp);
            dw2a[i] = fma2(ds2, gg, dw2a[i]);
            if (EXPORT) sacc2 = fma2(gg, w2p[i], sacc2);
            float d0, d1;
            upk(mul2(mul2(ds2, w2p[i]), gp), d0, d1);
            if (NP == 2) split2(d0, d1, dh_hi[i], dh_lo[i]); else dh_hi[i] = pack_bf16(d0, d1);
          }
          if (cq == 0) db2a += dsr;
          // the shared-memory dhid tile is free once the previous tile's weight-gradient product has completed
          umma::mbar_wait(&sm.dw_done, (tile_seq & 1) ^ 1);
          const uint32_t o0 = tile_off(r, 2 * cq), o1 = tile_off(r, 2 * cq + 1);
          *reinterpret_cast<uint4*>(sm.dhid[0] + o0) = make_uint4(dh_hi[0], dh_hi[1], dh_hi[2], dh_hi[3]);
          *reinterpret_cast<uint4*>(sm.dhid[0] + o1) = make_uint4(dh_hi[4], dh_hi[5], dh_hi[6], dh_hi[7]);
          if (NP == 2) {
            *reinterpret_cast<uint4*>(sm.dhid[1] + o0) = make_uint4(dh_lo[0], dh_lo[1], dh_lo[2], dh_lo[3]);
            *reinterpret_cast<uint4*>(sm.dhid[1] + o1) = make_uint4(dh_lo[4], dh_lo[5], dh_lo[6], dh_lo[7]);
          }
          if (EXPORT) {
            unsigned char* gt = dhid_g + (size_t)(tile0 + tile_seq) * (NP * DHID_PART);
            *reinterpret_cast<uint4*>(gt + o0) = make_uint4(dh_hi[0], dh_hi[1], dh_hi[2], dh_hi[3]);
            *reinterpret_cast<uint4*>(gt + o1) = make_uint4(dh_hi[4], dh_hi[5], dh_hi[6], dh_hi[7]);
            if (NP == 2) {
              *reinterpret_cast<uint4*>(gt + DHID_PART + o0) = make_uint4(dh_lo[0], dh_lo[1], dh_lo[2], dh_lo[3]);
              *reinterpret_cast<uint4*>(gt + DHID_PART + o1) = make_uint4(dh_lo[4], dh_lo[5], dh_lo[6], dh_lo[7]);
            }
            float s0, s1;
            upk(sacc2, s0, s1);
            sc_g[(size_t)(tile0 + tile_seq) * 512 + cq * 128 + r] = s0 + s1 + (cq == 0 ? b2 : 0.f);
          }
          umma::fence_async_smem();
          arrive_warp(&sm.e1_done);
          if (pend) write_dtp();
        }
        arrive_warp(&sm.stage_empty[s]);
      }
      pend = true; p_ps = unit_seq & 1; p_pph = (unit_seq >> 1) & 1; p_b = b; p_c0 = c0; p_ncg = ncg;
    }
    if (pend) write_dtp();
    // ---- per-CTA partial sums: dW^T from tensor memory, dw2 / db2 from registers
    umma::mbar_wait(&sm.fin, 0);
    umma::fence_after_sync();
    float* out = part + (long long)blockIdx.x * ATT_TC_PARTIAL;
    {
      float v[16];
      umma::tmem_ld16(tmem + C_DW + 16 * cq + lane_sel, v);
      const int kk = 32 * sp + lane;                       // k' of this lane: 0-63 dWd^T rows, 64-127 dA^T rows
      float* dst = out + (kk < 64 ? TCP_DWD + kk * 64 : TCP_DA + (kk - 64) * 64) + 16 * cq;
      const bool any = u1 > u0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(dst + 4 * q) = any ? make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {
      float a[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) upk(dw2a[i], a[2 * i], a[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) db2a += __shfl_xor_sync(0xffffffffu, db2a, o);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sm.red[warp - W_EPI][i] = a[i];
        sm.red[warp - W_EPI][16] = db2a;
      }
      __syncwarp();
      asm volatile("bar.sync 1, 512;\n" ::: "memory");     // the 16 epilogue warps
      const int et = tid - W_EPI * 32;
      if (et < 64) {                                       // column j = 16 cq' + i: the four sub-partition warps of quarter cq', in order
        const int cqq = et >> 4, i = et & 15;
        out[TCP_DW2 + et] = ((sm.red[cqq * 4 + 0][i] + sm.red[cqq * 4 + 1][i]) + sm.red[cqq * 4 + 2][i]) + sm.red[cqq * 4 + 3][i];
      } else if (et == 64) {
        out[TCP_DB2] = ((sm.red[0][16] + sm.red[1][16]) + sm.red[2][16]) + sm.red[3][16];
      }
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, B_TMEM_COLS);
}

// =====================================================================================================
// input gradients of the label branch from the exported dhid tiles
// =====================================================================================================
constexpr int IG_THREADS = 608;                  // 19 warps: 16 epilogue (warp % 4 = TMEM sub-partition, warp / 4 = column quarter), MMA, 2 loaders
constexpr int IG_W_MMA = 16, IG_W_LOAD = 17, IG_W_TILES = 18;
constexpr int ZS = 68;                           // floats per row of the exchange tiles

struct StageI {
  float hf[HCH * HF_STRIDE];
  __align__(16) float tv[CG][64];
  __align__(16) float dpv[CG][64];
};
template <int NP>
struct SmemI {
  __align__(128) unsigned char W[2][W_TILE];
  __align__(128) unsigned char dhid[2][NP * DHID_PART];   // two tile stages, each hi | lo (the exported image, copied by the TMA engine)
  __align__(16) float sc[2][4][128];                      // partial scores of the tile
  StageI st[NST];
  __align__(16) float Z[128 * ZS];                        // per row: t (.) X + Y + s dP   (what the row adds to dh of its history row)
  __align__(16) float Uu[128 * ZS];                       // per row: h (.) X              (what the row adds to dt of its candidate)
  float b2;
  __align__(16) float zrow[64];
  uint64_t stage_full[NST], stage_empty[NST], tile_full[2], tile_empty[2], xy_full[2], xy_empty[2], wbar;
  uint32_t tmem_base;
};

template <int SPLIT>
__global__ void __launch_bounds__(IG_THREADS, 1)
attention_input_grad_rs_kernel(const float* __restrict__ rows_g, int B, int H, int C, const unsigned char* __restrict__ img,
                               const float* __restrict__ e, const float* __restrict__ de, const unsigned char* __restrict__ dhid_g,
                               const float* __restrict__ sc_g, float* __restrict__ dxh, float* __restrict__ dxt) {
  pdl_wait();
  pdl_trigger();
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemI<NP>& sm = *reinterpret_cast<SmemI<NP>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TILE_BYTES = NP * DHID_PART;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { umma::mbar_init(&sm.stage_full[i], 1); umma::mbar_init(&sm.stage_empty[i], 16); }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&sm.tile_full[i], 1); umma::mbar_init(&sm.tile_empty[i], 1 + 16);     // product's commit + 16 epilogue warps (scores)
      umma::mbar_init(&sm.xy_full[i], 1); umma::mbar_init(&sm.xy_empty[i], 16);
    }
    umma::mbar_init(&sm.wbar, 1);
    const uint32_t bar = umma::smem_u32(&sm.wbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2u * W_TILE) : "memory");
    for (int p = 0; p < 2; ++p)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(umma::smem_u32(sm.W[p])),
                   "l"(img + (size_t)p * W_TILE), "r"(W_TILE), "r"(bar) : "memory");
  }
  if (tid < 64) {
    sm.zrow[tid] = 0.f;
    if (tid == 0) sm.b2 = 0.f;
  }
  if (warp == 0) umma::tmem_alloc(&sm.tmem_base, 256);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  GeoB g;
  g.B = B; g.H = H; g.C = C; g.G = (C + CG - 1) / CG; g.nchunks = (H + HCH - 1) / HCH;
  // more than two candidate groups per impression: whole impressions per CTA, so that dxh is accumulated group after group by ONE
  // CTA (plain read-modify-write, fixed order); exactly two groups: each adds once into a zeroed buffer (commutative); one: plain store
  int u0, u1;
  if (g.G > 2) {
    u0 = (int)((long long)B * blockIdx.x / gridDim.x) * g.G;
    u1 = (int)((long long)B * (blockIdx.x + 1) / gridDim.x) * g.G;
  } else {
    const long long U = (long long)B * g.G;
    u0 = (int)(U * blockIdx.x / gridDim.x); u1 = (int)(U * (blockIdx.x + 1) / gridDim.x);
  }
  const long long tile0 = g.tile_base(u0);

  if (warp == IG_W_LOAD) {
    // =========================================== loader ===========================================
    const int total = (u1 - u0) * g.nchunks;
    ChunkIterB it_issue, it_fin;
    int n_issued = 0;
    auto issue = [&]() {
      const int s = n_issued % NST;
      umma::mbar_wait(&sm.stage_empty[s], ((n_issued / NST) & 1) ^ 1);
      StageI& st = sm.st[s];
      const float* src = rows_g + ((long long)it_issue.b * H + it_issue.h0) * 64;
      for (int i = lane; i < it_issue.hl * 16; i += 32) cp_async16(&st.hf[(i >> 4) * HF_STRIDE + 4 * (i & 15)], src + 4 * i);
      for (int i = lane; i < it_issue.ncg * 16; i += 32) {
        const int c = i >> 4, q = i & 15;
        const long long rc = (long long)it_issue.b * C + it_issue.c0 + c;
        cp_async16(&st.tv[c][4 * q], e + rc * E + E_XT + 4 * q);
        cp_async16(&st.dpv[c][4 * q], de + rc * E + E_LAB + 4 * q);
      }
      cp_async_commit();
      ++n_issued;
      if (n_issued < total) it_issue.next(g);
    };
    if (total > 0) { it_issue.set(g, u0, 0); it_fin.set(g, u0, 0); }
    while (n_issued < total && n_issued < NST - 1) issue();
    for (int k = 0; k < total; ++k) {
      if (n_issued - k - 1 >= 1) cp_async_wait<1>(); else cp_async_wait<0>();
      arrive_warp(&sm.stage_full[k % NST]);
      if (k + 1 < total) it_fin.next(g);
      if (n_issued < total) issue();
    }
  } else if (warp == IG_W_TILES) {
    // =========================================== tile loader: dhid images + scores, bulk copies by the TMA engine ===========================================
    if (lane == 0) {
      uint32_t tile_seq = 0;
      for (int u = u0; u < u1; ++u) {
        int b, c0, ncg; g.unit(u, b, c0, ncg);
        for (int ci = 0; ci < g.nchunks; ++ci) {
          int h0, hl, rows, ntiles; g.chunk(ci, ncg, h0, hl, rows, ntiles);
          for (int ti = 0; ti < ntiles; ++ti, ++tile_seq) {
            const uint32_t ts = tile_seq & 1;
            umma::mbar_wait(&sm.tile_empty[ts], ((tile_seq >> 1) & 1) ^ 1);
            const uint32_t bar = umma::smem_u32(&sm.tile_full[ts]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(TILE_BYTES + 2048u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(umma::smem_u32(sm.dhid[ts])),
                         "l"(dhid_g + (size_t)(tile0 + tile_seq) * TILE_BYTES), "r"(TILE_BYTES), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(umma::smem_u32(sm.sc[ts])),
                         "l"(sc_g + (size_t)(tile0 + tile_seq) * 512), "r"(2048u), "r"(bar) : "memory");
          }
        }
      }
    }
  } else if (warp == IG_W_MMA) {
    // =========================================== MMA issuer ===========================================
    if (umma::elect_one()) {
      umma::mbar_wait(&sm.wbar, 0);
      constexpr uint32_t IDESC_XY = umma::make_idesc_bf16(128, 128, false, true);     // dhid (K-major) x W read MN-major: da[r][k'] = sum_j dhid[r][j] W[j][k']
      const umma::Operand w_mn = umma::make_operand(umma::smem_u32(sm.W[0]), 128, 1024, 256, W_TILE);
      uint32_t tile_seq = 0;
      for (int u = u0; u < u1; ++u) {
        int b, c0, ncg; g.unit(u, b, c0, ncg);
        for (int ci = 0; ci < g.nchunks; ++ci) {
          int h0, hl, rows, ntiles; g.chunk(ci, ncg, h0, hl, rows, ntiles);
          for (int ti = 0; ti < ntiles; ++ti, ++tile_seq) {
            const uint32_t ts = tile_seq & 1, ph = (tile_seq >> 1) & 1;
            umma::mbar_wait(&sm.tile_full[ts], ph);
            umma::mbar_wait(&sm.xy_empty[ts], ph ^ 1);
            umma::fence_after_sync();
            umma::mma_product<SPLIT, 4>(tmem + 128 * ts, umma::make_operand(umma::smem_u32(sm.dhid[ts]), LBO128, 128, 2 * LBO128, DHID_PART), w_mn,
                                        IDESC_XY, false);
            umma::mma_commit(&sm.tile_empty[ts]);
            umma::mma_commit(&sm.xy_full[ts]);
          }
        }
      }
    }
  } else if (warp < 16) {
    // =========================================== epilogue ===========================================
    const int sp = warp & 3, cq = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(32 * sp) << 16;
    const int et = tid;                                     // 0 .. 511
    const int rh = et >> 3, rk8 = et & 7;                   // dh reduction: history row, 8-column block
    // dt reduction: (candidate slot, 4-column block, quarter of the slot's rows); the four quarters (adjacent lanes) are combined with
    // shuffles in a fixed order, so that no thread walks more than 16 rows
    const int tslot = et >> 6, tk4 = (et >> 2) & 15, tq = et & 3;
    uint32_t chunk_seq = 0, tile_seq = 0;
    for (int u = u0; u < u1; ++u) {
      int b, c0, ncg; g.unit(u, b, c0, ncg);
      const int gidx = u - (u / g.G) * g.G;
      float4 dt_acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ci = 0; ci < g.nchunks; ++ci, ++chunk_seq) {
        int h0, hl, rows, ntiles; g.chunk(ci, ncg, h0, hl, rows, ntiles);
        const uint32_t s = chunk_seq % NST, ph = (chunk_seq / NST) & 1;
        const StageI& st = sm.st[s];
        const float inv_hl = 1.0f / (float)hl;
        umma::mbar_wait(&sm.stage_full[s], ph);
        float dh_acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dh_acc[i] = 0.f;
        for (int ti = 0; ti < ntiles; ++ti, ++tile_seq) {
          const uint32_t ts = tile_seq & 1, tph = (tile_seq >> 1) & 1;
          umma::mbar_wait(&sm.xy_full[ts], tph);
          umma::fence_after_sync();
          float X[16], Y[16];
          umma::tmem_ld16(tmem + 128 * ts + 16 * cq + lane_sel, X);
          umma::tmem_ld16(tmem + 128 * ts + 64 + 16 * cq + lane_sel, Y);
          umma::fence_before_sync();
          arrive_warp(&sm.xy_empty[ts]);
          const int r = 32 * sp + lane, rg = ti * 128 + r;
          const bool valid = rg < rows;
          const int cl = valid ? __float2int_rz(((float)rg + 0.5f) * inv_hl) : 0, hloc = valid ? rg - cl * hl : 0;
          // the scores were written by the TMA engine (async proxy): every reading thread acquires the tile's own barrier -- its phase
          // completed long ago (the product waited for it), but observing xy_full alone does not order these reads behind the copy
          umma::mbar_wait(&sm.tile_full[ts], tph);
          const float srow = valid ? ((sm.sc[ts][0][r] + sm.sc[ts][1][r]) + sm.sc[ts][2][r]) + sm.sc[ts][3][r] : 0.f;
          arrive_warp(&sm.tile_empty[ts]);
          const float* hrow = valid ? &st.hf[hloc * HF_STRIDE + 16 * cq] : sm.zrow;
          const float* trow = &st.tv[cl][16 * cq];
          const float* dprow = &st.dpv[cl][16 * cq];
          float* zr = &sm.Z[r * ZS + 16 * cq];
          float* ur = &sm.Uu[r * ZS + 16 * cq];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(trow + 4 * q), p4 = *reinterpret_cast<const float4*>(dprow + 4 * q);
            const float4 h4 = *reinterpret_cast<const float4*>(hrow + 4 * q);
            float4 z, uu;
            z.x = fmaf(t4.x, X[4 * q], fmaf(srow, p4.x, Y[4 * q]));         uu.x = h4.x * X[4 * q];
            z.y = fmaf(t4.y, X[4 * q + 1], fmaf(srow, p4.y, Y[4 * q + 1])); uu.y = h4.y * X[4 * q + 1];
            z.z = fmaf(t4.z, X[4 * q + 2], fmaf(srow, p4.z, Y[4 * q + 2])); uu.z = h4.z * X[4 * q + 2];
            z.w = fmaf(t4.w, X[4 * q + 3], fmaf(srow, p4.w, Y[4 * q + 3])); uu.w = h4.w * X[4 * q + 3];
            if (!valid) { z = make_float4(0.f, 0.f, 0.f, 0.f); uu = z; }
            *reinterpret_cast<float4*>(zr + 4 * q) = z;
            *reinterpret_cast<float4*>(ur + 4 * q) = uu;
          }
          __syncwarp();                                                  // bar.sync is the ALIGNED barrier: the warp must arrive converged
          asm volatile("bar.sync 1, 512;\n" ::: "memory");              // Z / U of the tile complete
          // dh[h][8 k8 ..]: the rows (c, h) of this tile, candidates in order
          if (rh < hl) {
#pragma unroll 1
            for (int c = 0; c < ncg; ++c) {
              const int rr = c * hl + rh - ti * 128;
              if (rr >= 0 && rr < 128) {
                const float4 a = *reinterpret_cast<const float4*>(&sm.Z[rr * ZS + 8 * rk8]), bq = *reinterpret_cast<const float4*>(&sm.Z[rr * ZS + 8 * rk8 + 4]);
                dh_acc[0] += a.x; dh_acc[1] += a.y; dh_acc[2] += a.z; dh_acc[3] += a.w;
                dh_acc[4] += bq.x; dh_acc[5] += bq.y; dh_acc[6] += bq.z; dh_acc[7] += bq.w;
              }
            }
          }
          // dt[slot][k]: the rows of candidate `slot` inside this tile, in row order
          if (tslot < ncg) {
            const int lo = max(tslot * hl - ti * 128, 0), hi = min(min((tslot + 1) * hl - ti * 128, 128), rows - ti * 128);
#pragma unroll 4
            for (int rr = lo + tq; rr < hi; rr += 4) {
              const float4 v = *reinterpret_cast<const float4*>(&sm.Uu[rr * ZS + 4 * tk4]);
              dt_acc.x += v.x; dt_acc.y += v.y; dt_acc.z += v.z; dt_acc.w += v.w;
            }
          }
          __syncwarp();                                                  // (the dt loop above has lane-dependent trip counts)
          asm volatile("bar.sync 2, 512;\n" ::: "memory");              // exchange tiles free for the next tile
        }
        // dh of this chunk's history rows
        if (rh < hl) {
          float* dst = dxh + ((long long)b * H + h0 + rh) * 64 + 8 * rk8;
          if (g.G == 1 || (g.G > 2 && gidx == 0)) {
            *reinterpret_cast<float4*>(dst) = make_float4(dh_acc[0], dh_acc[1], dh_acc[2], dh_acc[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(dh_acc[4], dh_acc[5], dh_acc[6], dh_acc[7]);
          } else if (g.G == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(dst + i, dh_acc[i]);
          } else {
            float4 a = *reinterpret_cast<float4*>(dst), bq = *reinterpret_cast<float4*>(dst + 4);
            a.x += dh_acc[0]; a.y += dh_acc[1]; a.z += dh_acc[2]; a.w += dh_acc[3];
            bq.x += dh_acc[4]; bq.y += dh_acc[5]; bq.z += dh_acc[6]; bq.w += dh_acc[7];
            *reinterpret_cast<float4*>(dst) = a; *reinterpret_cast<float4*>(dst + 4) = bq;
          }
        }
        arrive_warp(&sm.stage_empty[s]);
      }
      // dt of the unit's candidates + the direct ec path (user_model.py:31: e_concat holds the candidate's own features)
      {
        float4 t = dt_acc;                                  // quarters 0..3 of the rows: (q0 + q1) + (q2 + q3)
        t.x += __shfl_xor_sync(0xffffffffu, t.x, 1); t.y += __shfl_xor_sync(0xffffffffu, t.y, 1);
        t.z += __shfl_xor_sync(0xffffffffu, t.z, 1); t.w += __shfl_xor_sync(0xffffffffu, t.w, 1);
        t.x += __shfl_xor_sync(0xffffffffu, t.x, 2); t.y += __shfl_xor_sync(0xffffffffu, t.y, 2);
        t.z += __shfl_xor_sync(0xffffffffu, t.z, 2); t.w += __shfl_xor_sync(0xffffffffu, t.w, 2);
        if (tslot < ncg && tq == 0) {
          const long long rc = (long long)b * C + c0 + tslot;
          const float4 d = __ldg(reinterpret_cast<const float4*>(de + rc * E + E_XT) + tk4);
          *reinterpret_cast<float4*>(dxt + rc * 64 + 4 * tk4) = make_float4(t.x + d.x, t.y + d.y, t.z + d.z, t.w + d.w);
        }
      }
      if (g.G > 2) __threadfence();                          // the next group of this impression (same CTA) reads dxh back
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace rsb

static int rs_bwd_grid(const Workspace& w) {
  const long long U = (long long)w.B * ((w.C + rs::CG - 1) / rs::CG);
  return (int)std::min<long long>(U, std::min(sm_count(), ATT_TC_PARTS_MAX));
}
long long attention_rs_tiles(int B, int H, int C) {
  long long per_imp = 0;
  for (int c0 = 0; c0 < C; c0 += rs::CG) {
    const int ncg = std::min(rs::CG, C - c0);
    for (int h0 = 0; h0 < H; h0 += rs::HCH) per_imp += (ncg * std::min(rs::HCH, H - h0) + 127) / 128;
  }
  return per_imp * B;
}

template <int SPLIT, bool EXPORT>
static int launch_bwd_rs(Workspace& w, int branch, cudaStream_t s) {
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  const size_t smem = sizeof(rsb::SmemB<NP>);
  const int grid = rs_bwd_grid(w);
  w.att_tc_parts[branch] = grid;
  float* part = w.att_part + (long long)branch * ATT_TC_PARTS_MAX * ATT_TC_PARTIAL;
  float* dtp = w.dtp + (long long)branch * w.R * 64;
  const float* tpg = w.tp + (long long)branch * w.R * 64;
  NRM_CUDA(cudaFuncSetAttribute(rsb::attention_backward_rs_kernel<SPLIT, EXPORT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_pdl(rsb::attention_backward_rs_kernel<SPLIT, EXPORT>, dim3(grid), dim3(rs::THREADS), smem, s, branch == 0 ? w.xh : w.pca_h, branch, w.B, w.H,
             w.C, reinterpret_cast<const unsigned char*>(w.att_rs_img), tpg, w.e, w.de, dtp, part, reinterpret_cast<unsigned char*>(w.att_dhid),
             w.att_sc);
  NRM_LAUNCH_CHECK("attention_backward_rs_kernel");
  return NRM_OK;
}

template <int SPLIT>
static int launch_ig_rs(Workspace& w, cudaStream_t s) {
  constexpr int NP = SPLIT == 3 ? 2 : 1;
  const size_t smem = sizeof(rsb::SmemI<NP>);
  const int G = (w.C + rs::CG - 1) / rs::CG;
  const int grid = (int)std::min<long long>(G > 2 ? (long long)w.B : (long long)w.B * G, (long long)sm_count());
  if (G == 2) NRM_CUDA(cudaMemsetAsync(w.dxh, 0, sizeof(float) * (size_t)w.NH * 64, s));     // the two groups of an impression add into it
  NRM_CUDA(cudaFuncSetAttribute(rsb::attention_input_grad_rs_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_pdl(rsb::attention_input_grad_rs_kernel<SPLIT>, dim3(grid), dim3(rsb::IG_THREADS), smem, s, w.xh, w.B, w.H, w.C,
             reinterpret_cast<const unsigned char*>(w.att_rs_img), w.e, w.de, reinterpret_cast<const unsigned char*>(w.att_dhid), w.att_sc, w.dxh, w.dxt);
  NRM_LAUNCH_CHECK("attention_input_grad_rs_kernel");
  return NRM_OK;
}
// label branch: dh (-> w.dxh) and dt (-> w.dxt) from the tiles exported by launch_attention_backward_rs(..., export_dhid = true)
int launch_attention_input_grad_rs(Workspace& w, int precision, cudaStream_t s) {
  return precision == NRM_PRECISION_BF16 ? launch_ig_rs<1>(w, s) : launch_ig_rs<3>(w, s);
}

// weights-side backward of one branch (0 = label, 1 = text/img).  The label branch exports its dhid tiles for
// launch_attention_input_grad_rs.
int launch_attention_backward_rs(Workspace& w, int branch, int precision, bool export_dhid, cudaStream_t s) {
  if (precision == NRM_PRECISION_BF16) return export_dhid ? launch_bwd_rs<1, true>(w, branch, s) : launch_bwd_rs<1, false>(w, branch, s);
  return export_dhid ? launch_bwd_rs<3, true>(w, branch, s) : launch_bwd_rs<3, false>(w, branch, s);
}

}  // namespace nrm
