// Feature-row kernels: decode the packed float64 rows, gather the embedding tables
// (reference: user_invariant_interest_model.py:58-79, user_instant_interest_model.py:20-23)
// and, for the backward, build sorted segments of table ids and reduce the row gradients
// into the tables without atomics on floating-point data.
#include "nrm_kernels.cuh"

namespace nrm {

// ---------------------------------------------------------------------------------
// embed_rows_kernel: 8 threads per row, 32 rows per CTA.
//   history row  -> xin_h[row, 0:66] = [cat+mean(sub) 32 | relu(sent) 16 | type 8 | time 8 | rt | scroll]
//   target row   -> e[row, 136:200] = same first 64 columns, e[row, 200:264] = pca (fp32),
//                   e[row, 128:136] = relu(Linear(3,8)(x_global row))
// Rows are addressed as one list: [0, NH) history, [NH, NH+R) targets.
// When keys32/keys8 are non-null the decoded table ids are also written as sort keys.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_rows_kernel(const double* __restrict__ xh, const double* __restrict__ xt, long long xt_bs,
                  const double* __restrict__ xg, long long xg_bs, int H, int C, long long NH, long long N,
                  const float* __restrict__ P, float* __restrict__ xin_h, float* __restrict__ e,
                  int* __restrict__ keys32, int* __restrict__ keys8) {
  const long long row = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  const int q = threadIdx.x & 7;
  if (row >= N) return;
  const bool is_hist = row < NH;
  const double* src;
  if (is_hist) {
    src = xh + row * HC;
  } else {
    const long long r = row - NH;
    src = xt + (r / C) * xt_bs + (r % C) * TC;
  }
  // ids are exact integers stored as doubles; float32 -> int64 truncation in the reference
  int tix[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tix[i] = (int)(float)src[i];
  const int cat = clampi((int)(float)src[68], 0, NCAT - 1);
  int sub[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) sub[i] = clampi((int)(float)src[69 + i], 0, NCAT - 1);
  const float s0 = (float)src[74], s1 = (float)src[75], s2 = (float)src[76];
  const int typ = clampi((int)(float)src[77], 0, NTYPE - 1);
  const int iy = clampi(tix[0], 0, NYEAR - 1), im = clampi(tix[1], 0, NMONTH - 1);
  const int id = clampi(tix[2], 0, NDAY - 1), ih = clampi(tix[3], 0, NHOUR - 1);

  // category + mean of the 5 sub-categories, columns 4q..4q+3
  const float4* tab = reinterpret_cast<const float4*>(P + P_CAT);
  const float4 c = __ldg(tab + cat * 8 + q);
  float4 s = __ldg(tab + sub[0] * 8 + q);
#pragma unroll
  for (int i = 1; i < 5; ++i) {
    const float4 v = __ldg(tab + sub[i] * 8 + q);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float4 both;
  both.x = c.x + s.x / 5.0f; both.y = c.y + s.y / 5.0f; both.z = c.z + s.z / 5.0f; both.w = c.w + s.w / 5.0f;

  // sentiment Linear(3,16)+ReLU, outputs 2q, 2q+1
  float se[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int o = 2 * q + i;
    const float* w = P + P_SENT_W + o * 3;
    float v = __ldg(P + P_SENT_B + o);
    v = fmaf(s0, __ldg(w + 0), v); v = fmaf(s1, __ldg(w + 1), v); v = fmaf(s2, __ldg(w + 2), v);
    se[i] = fmaxf(v, 0.f);
  }
  const float ty = __ldg(P + P_TYPE + typ * 8 + q);
  float tm = __ldg(P + P_YEAR + iy * 8 + q);
  tm += __ldg(P + P_MONTH + im * 8 + q);
  tm += __ldg(P + P_DAY + id * 8 + q);
  tm += __ldg(P + P_HOUR + ih * 8 + q);

  float* dst = is_hist ? (xin_h + row * XIN) : (e + (row - NH) * E + E_XT);
  // 66-float rows are only 8-byte aligned: use float2 stores for the 4-wide group
  reinterpret_cast<float2*>(dst + 4 * q)[0] = make_float2(both.x, both.y);
  reinterpret_cast<float2*>(dst + 4 * q)[1] = make_float2(both.z, both.w);
  reinterpret_cast<float2*>(dst + 32 + 2 * q)[0] = make_float2(se[0], se[1]);
  dst[48 + q] = ty;
  dst[56 + q] = tm;
  if (is_hist) {
    if (q < 2) dst[64 + q] = (float)src[78 + q];
  } else {
    const long long r = row - NH;
    float* er = e + r * E;
    // pca columns 4..67 -> e[200:264]; 8 threads x 8 values
#pragma unroll
    for (int i = 0; i < 8; ++i) er[E_PCAT + q * 8 + i] = (float)src[4 + q * 8 + i];
    // instant-interest: relu(W g + b), output q
    const double* gsrc = xg + (r / C) * xg_bs + (r % C) * GC;
    const float g0 = (float)gsrc[0], g1 = (float)gsrc[1], g2 = (float)gsrc[2];
    const float* w = P + P_INST_W + q * 3;
    float v = __ldg(P + P_INST_B + q);
    v = fmaf(g0, __ldg(w + 0), v); v = fmaf(g1, __ldg(w + 1), v); v = fmaf(g2, __ldg(w + 2), v);
    er[E_INST + q] = fmaxf(v, 0.f);
  }
  if (keys32 != nullptr && q == 0) {
    int* k32 = keys32 + row * 6;
    k32[0] = cat;
#pragma unroll
    for (int i = 0; i < 5; ++i) k32[1 + i] = sub[i];
    int* k8 = keys8 + row * 5;
    k8[0] = K8_TYPE + typ; k8[1] = K8_YEAR + iy; k8[2] = K8_MONTH + im; k8[3] = K8_DAY + id; k8[4] = K8_HOUR + ih;
  }
}

// ---------------------------------------------------------------------------------
// Deterministic counting sort of table ids (keys < nkeys <= 3000):
//   sort_hist:    per 2048-entry chunk, histogram of keys           -> chunk_hist[chunk][key]
//   sort_colscan: per key (a warp each), exclusive scan over chunks (in place) + key totals
//   sort_keyscan: exclusive scan over keys of the totals -> seg[0..nkeys] (segment starts),
//                 the same for the number of SEG_GROUP-sized groups -> seg[nkeys+1 .. 2nkeys+1],
//                 and the group -> key map
//   sort_scatter: stable rank of each entry inside its chunk + the two offsets -> perm
// (four kernels.)  Integer atomics are used only for counts (order independent).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sort_hist_kernel(const int* __restrict__ keys, long long n, int nkeys, int* __restrict__ chunk_hist) {
  extern __shared__ int hist[];
  for (int i = threadIdx.x; i < nkeys; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * SORT_CHUNK;
  for (int i = threadIdx.x; i < SORT_CHUNK; i += blockDim.x)
    if (base + i < n) atomicAdd(&hist[keys[base + i]], 1);
  __syncthreads();
  int* out = chunk_hist + (long long)blockIdx.x * nkeys;
  for (int i = threadIdx.x; i < nkeys; i += blockDim.x) out[i] = hist[i];
}

// Per key (one warp each): exclusive scan of the chunk histograms over chunks, in place, and
// the key's total count.
__global__ void __launch_bounds__(256)
sort_colscan_kernel(int* __restrict__ chunk_hist, int nchunks, int nkeys, int* __restrict__ totals) {
  const int key = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (key >= nkeys) return;
  int carry = 0;
  for (int c0 = 0; c0 < nchunks; c0 += 32) {
    const int c = c0 + lane;
    int* p = chunk_hist + (long long)c * nkeys + key;
    const int own = (c < nchunks) ? *p : 0;
    int v = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (c < nchunks) *p = carry + v - own;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) totals[key] = carry;
}

// One CTA: exclusive scan over keys of the totals -> segment starts, and of the number of
// SEG_GROUP-sized groups -> group starts; also the group -> key map.
__global__ void __launch_bounds__(1024)
sort_keyscan_kernel(const int* __restrict__ totals, int nkeys, int* __restrict__ seg, int* __restrict__ group_key) {
  __shared__ int warp_off[2][32];
  __shared__ int block_tot[2];
  __shared__ int carry[2];
  int* seg_start = seg;
  int* grp_start = seg + nkeys + 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
  __syncthreads();
  for (int k0 = 0; k0 < nkeys; k0 += 1024) {
    const int key = k0 + threadIdx.x;
    const int total = key < nkeys ? totals[key] : 0;
    const int vals[2] = {total, (total + SEG_GROUP - 1) / SEG_GROUP};
    int excl[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int v = vals[s];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
      if (lane == 31) warp_off[s][wid] = v;      // inclusive warp total
      excl[s] = v - vals[s];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int own = warp_off[s][lane];
        int v = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        warp_off[s][lane] = v - own;               // exclusive offset of each warp
        if (lane == 31) block_tot[s] = v;
      }
    }
    __syncthreads();
    const int off0 = carry[0] + warp_off[0][wid] + excl[0];
    const int off1 = carry[1] + warp_off[1][wid] + excl[1];
    if (key < nkeys) {
      seg_start[key] = off0;
      grp_start[key] = off1;
      for (int gi = 0; gi < vals[1]; ++gi) group_key[off1 + gi] = key;
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry[0] += block_tot[0]; carry[1] += block_tot[1]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { seg_start[nkeys] = carry[0]; grp_start[nkeys] = carry[1]; }
}

// One warp per chunk walks its entries in order, 32 at a time: entries with equal keys
// inside a round are ranked with match_any, earlier rounds through a shared counter.
__global__ void __launch_bounds__(32)
sort_scatter_kernel(const int* __restrict__ keys, long long n, int nkeys, const int* __restrict__ chunk_hist,
                    const int* __restrict__ seg, int* __restrict__ perm) {
  extern __shared__ int cnt[];
  const int lane = threadIdx.x;
  for (int i = lane; i < nkeys; i += 32) cnt[i] = 0;
  __syncwarp();
  const long long base = (long long)blockIdx.x * SORT_CHUNK;
  const int total = (int)min((long long)SORT_CHUNK, n - base);
  const int* cb = chunk_hist + (long long)blockIdx.x * nkeys;    // exclusive per-chunk offsets
  for (int r0 = 0; r0 < total; r0 += 32) {
    const int i = r0 + lane;
    const bool valid = i < total;
    const int key = valid ? keys[base + i] : (nkeys + lane);     // distinct dummy keys never match
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int before = 0;
    if (valid && lane == leader) { before = cnt[key]; cnt[key] = before + __popc(peers); }
    before = __shfl_sync(0xffffffffu, before, leader);
    if (valid) perm[seg[key] + cb[key] + before + rank] = (int)(base + i);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------
// Segment reduction of row gradients into table rows.
//   entry id -> (row, slot); source vector = dsrc(row)[col0 + col], weight by slot.
//   level 1: one warp per group of <= SEG_GROUP sorted entries (all of one key)
//   level 2: one warp per key sums its groups in order and writes the table gradient row
// W = 32: category table (6 entries per row: cat weight 1, five sub-categories weight 1/5)
// W = 8 : type / year / month / day / hour tables (5 entries per row)
// ---------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ float seg_load(int entry, const float* __restrict__ dxin_h, const float* __restrict__ dxt,
                                          long long NH, int col) {
  constexpr int SLOTS = (W == 32) ? 6 : 5;
  const int row = entry / SLOTS, slot = entry - row * SLOTS;
  const int c = (W == 32) ? col : ((slot == 0 ? 48 : 56) + col);
  const float v = (row < NH) ? __ldg(dxin_h + (long long)row * XIN + c) : __ldg(dxt + ((long long)row - NH) * D + c);
  return (W == 32 && slot != 0) ? v / 5.0f : v;     // d mean(sub)/d sub_i = 1/5
}

template <int W>
__global__ void __launch_bounds__(256)
table_grad_l1_kernel(const int* __restrict__ perm, const int* __restrict__ seg, const int* __restrict__ group_key,
                     int nkeys, const float* __restrict__ dxin_h, const float* __restrict__ dxt, long long NH,
                     float* __restrict__ gpart) {
  const int* seg_start = seg;
  const int* grp_start = seg + nkeys + 1;
  const int ngroups = grp_start[nkeys];
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (g >= ngroups) return;
  const int lane = threadIdx.x & 31;
  const int key = group_key[g];
  const int beg = seg_start[key] + (g - grp_start[key]) * SEG_GROUP;
  const int end = min(beg + SEG_GROUP, seg_start[key + 1]);
  if (W == 32) {
    float acc = 0.f;
    int i = beg;
    for (; i + 4 <= end; i += 4) {
      const int e0 = perm[i], e1 = perm[i + 1], e2 = perm[i + 2], e3 = perm[i + 3];
      const float v0 = seg_load<32>(e0, dxin_h, dxt, NH, lane), v1 = seg_load<32>(e1, dxin_h, dxt, NH, lane);
      const float v2 = seg_load<32>(e2, dxin_h, dxt, NH, lane), v3 = seg_load<32>(e3, dxin_h, dxt, NH, lane);
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; i < end; ++i) acc += seg_load<32>(perm[i], dxin_h, dxt, NH, lane);
    gpart[(long long)g * 32 + lane] = acc;
  } else {
    const int sub = lane >> 3, col = lane & 7;
    float acc = 0.f;
    for (int i = beg + sub; i < end; i += 4) acc += seg_load<8>(perm[i], dxin_h, dxt, NH, col);
    // combine the four interleaved partial sums in a fixed order
    const float a1 = __shfl_down_sync(0xffffffffu, acc, 8);
    const float a2 = __shfl_down_sync(0xffffffffu, acc, 16);
    const float a3 = __shfl_down_sync(0xffffffffu, acc, 24);
    if (sub == 0) gpart[(long long)g * 8 + col] = ((acc + a1) + a2) + a3;
  }
}

// out_row(key) points into the flat gradient buffer.
template <int W>
__global__ void __launch_bounds__(256)
table_grad_l2_kernel(const int* __restrict__ seg, int nkeys, const float* __restrict__ gpart, float* __restrict__ grads) {
  const int* grp_start = seg + nkeys + 1;
  constexpr int KPW = 32 / W;                      // keys per warp
  const int warp = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int key = warp * KPW + lane / W;
  const int col = lane % W;
  if (key >= nkeys) return;
  const int g0 = grp_start[key], g1 = grp_start[key + 1];
  float acc = 0.f;
  int g = g0;
  for (; g + 4 <= g1; g += 4) {
    const float v0 = gpart[(long long)g * W + col], v1 = gpart[(long long)(g + 1) * W + col];
    const float v2 = gpart[(long long)(g + 2) * W + col], v3 = gpart[(long long)(g + 3) * W + col];
    acc += v0; acc += v1; acc += v2; acc += v3;
  }
  for (; g < g1; ++g) acc += gpart[(long long)g * W + col];
  float* dst;
  if (W == 32) {
    dst = grads + P_CAT + (long long)key * 32;
  } else {
    if (key < K8_YEAR) dst = grads + P_TYPE + (long long)(key - K8_TYPE) * 8;
    else if (key < K8_MONTH) dst = grads + P_YEAR + (long long)(key - K8_YEAR) * 8;
    else if (key < K8_DAY) dst = grads + P_MONTH + (long long)(key - K8_MONTH) * 8;
    else if (key < K8_HOUR) dst = grads + P_DAY + (long long)(key - K8_DAY) * 8;
    else dst = grads + P_HOUR + (long long)(key - K8_HOUR) * 8;
  }
  dst[col] = acc;
}

// ---------------------------------------------------------------------------------
// Sentiment Linear(3,16)+ReLU and instant Linear(3,8)+ReLU gradients.
// Thread (o, i) of a 4-row-group x 64 layout walks its rows in order; i == 3 is the bias.
// part[blockIdx.x][0:64]  = sentiment (o*4 + i), part[..][64:96] = instant (o*4 + i)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
small_linear_grad_kernel(const double* __restrict__ xh, const double* __restrict__ xt, long long xt_bs,
                         const double* __restrict__ xg, long long xg_bs, int C, long long NH, long long N,
                         const float* __restrict__ xin_h, const float* __restrict__ e,
                         const float* __restrict__ dxin_h, const float* __restrict__ dxt, const float* __restrict__ de,
                         int rows_per_cta, float* __restrict__ part) {
  __shared__ float red[4][96];
  const int grp = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int o = t >> 2, i = t & 3;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(N, r0 + rows_per_cta);
  float acc_s = 0.f, acc_i = 0.f;
  for (long long row = r0 + grp; row < r1; row += 4) {
    const bool is_hist = row < NH;
    const long long r = row - NH;
    const double* src = is_hist ? (xh + row * HC) : (xt + (r / C) * xt_bs + (r % C) * TC);
    const float act = is_hist ? xin_h[row * XIN + 32 + o] : e[r * E + E_XT + 32 + o];
    const float dv = is_hist ? dxin_h[row * XIN + 32 + o] : dxt[r * D + 32 + o];
    const float dpre = act > 0.f ? dv : 0.f;
    const float in = (i < 3) ? (float)src[74 + i] : 1.0f;
    acc_s = fmaf(dpre, in, acc_s);
    if (!is_hist && o < 8) {
      const double* gsrc = xg + (r / C) * xg_bs + (r % C) * GC;
      const float ia = e[r * E + E_INST + o];
      const float dpi = ia > 0.f ? de[r * E + E_INST + o] : 0.f;
      const float gi = (i < 3) ? (float)gsrc[i] : 1.0f;
      acc_i = fmaf(dpi, gi, acc_i);
    }
  }
  red[grp][t] = acc_s;
  if (o < 8) red[grp][64 + o * 4 + i] = acc_i;
  __syncthreads();
  if (threadIdx.x < 96) {
    const int k = threadIdx.x;
    part[(long long)blockIdx.x * 96 + k] = ((red[0][k] + red[1][k]) + red[2][k]) + red[3][k];
  }
}

// Reduce the per-CTA partials (8 interleaved groups, combined in group order) and scatter the
// [96] vector into the flat gradient entries.
__global__ void __launch_bounds__(768)
small_linear_grad_finish_kernel(const float* __restrict__ part, int nparts, float* __restrict__ grads) {
  __shared__ float red[8][96];
  const int k = threadIdx.x % 96, grp = threadIdx.x / 96;
  float acc = 0.f;
#pragma unroll 4
  for (int p = grp; p < nparts; p += 8) acc += part[(long long)p * 96 + k];
  red[grp][k] = acc;
  __syncthreads();
  if (grp != 0) return;
  acc = red[0][k];
#pragma unroll
  for (int q = 1; q < 8; ++q) acc += red[q][k];
  if (k < 64) {
    const int o = k >> 2, i = k & 3;
    if (i < 3) grads[P_SENT_W + o * 3 + i] = acc; else grads[P_SENT_B + o] = acc;
  } else {
    const int o = (k - 64) >> 2, i = k & 3;
    if (i < 3) grads[P_INST_W + o * 3 + i] = acc; else grads[P_INST_B + o] = acc;
  }
}

// ---------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------
int launch_embed_rows(const BatchPtrs& in, const float* P, Workspace& w, bool with_keys, cudaStream_t s) {
  const int grid = (int)((w.N + 31) / 32);
  embed_rows_kernel<<<grid, 256, 0, s>>>(in.xh, in.xt, in.xt_bs, in.xg, in.xg_bs, w.H, w.C, w.NH, w.N, P, w.xin_h, w.e,
                                         with_keys ? w.keys32 : nullptr, with_keys ? w.keys8 : nullptr);
  NRM_LAUNCH_CHECK("embed_rows_kernel");
  return NRM_OK;
}

static int sort_stream(const int* keys, long long n, int nkeys, int* chunk_hist, int* seg, int* gkey, int* perm,
                       cudaStream_t s) {
  const int nchunks = (int)((n + SORT_CHUNK - 1) / SORT_CHUNK);
  sort_hist_kernel<<<nchunks, 1024, nkeys * sizeof(int), s>>>(keys, n, nkeys, chunk_hist);
  NRM_LAUNCH_CHECK("sort_hist_kernel");
  int* totals = seg + 2 * (nkeys + 1);
  sort_colscan_kernel<<<(nkeys + 7) / 8, 256, 0, s>>>(chunk_hist, nchunks, nkeys, totals);
  NRM_LAUNCH_CHECK("sort_colscan_kernel");
  sort_keyscan_kernel<<<1, 1024, 0, s>>>(totals, nkeys, seg, gkey);
  NRM_LAUNCH_CHECK("sort_keyscan_kernel");
  sort_scatter_kernel<<<nchunks, 32, nkeys * sizeof(int), s>>>(keys, n, nkeys, chunk_hist, seg, perm);
  NRM_LAUNCH_CHECK("sort_scatter_kernel");
  return NRM_OK;
}

int launch_table_grads(Workspace& w, float* grads, cudaStream_t s) {
  const long long n32 = w.N * 6, n8 = w.N * 5;
  NRM_TRY(sort_stream(w.keys32, n32, NKEY32, w.chunk_hist32, w.seg32, w.gkey32, w.perm32, s));
  NRM_TRY(sort_stream(w.keys8, n8, NKEY8, w.chunk_hist8, w.seg8, w.gkey8, w.perm8, s));
  const long long g32 = n32 / SEG_GROUP + NKEY32 + 1, g8 = n8 / SEG_GROUP + NKEY8 + 1;   // upper bounds
  table_grad_l1_kernel<32><<<(int)((g32 + 7) / 8), 256, 0, s>>>(w.perm32, w.seg32, w.gkey32, NKEY32, w.dxin_h, w.dxt, w.NH, w.gpart32);
  NRM_LAUNCH_CHECK("table_grad_l1_kernel<32>");
  table_grad_l1_kernel<8><<<(int)((g8 + 7) / 8), 256, 0, s>>>(w.perm8, w.seg8, w.gkey8, NKEY8, w.dxin_h, w.dxt, w.NH, w.gpart8);
  NRM_LAUNCH_CHECK("table_grad_l1_kernel<8>");
  table_grad_l2_kernel<32><<<(NKEY32 + 7) / 8, 256, 0, s>>>(w.seg32, NKEY32, w.gpart32, grads);
  NRM_LAUNCH_CHECK("table_grad_l2_kernel<32>");
  table_grad_l2_kernel<8><<<((NKEY8 + 3) / 4 + 7) / 8, 256, 0, s>>>(w.seg8, NKEY8, w.gpart8, grads);
  NRM_LAUNCH_CHECK("table_grad_l2_kernel<8>");
  return NRM_OK;
}

int launch_small_linear_grads(const BatchPtrs& in, Workspace& w, float* grads, cudaStream_t s) {
  const int nparts = 1024;
  const int rows_per_cta = (int)((w.N + nparts - 1) / nparts);
  small_linear_grad_kernel<<<nparts, 256, 0, s>>>(in.xh, in.xt, in.xt_bs, in.xg, in.xg_bs, w.C, w.NH, w.N, w.xin_h, w.e,
                                                  w.dxin_h, w.dxt, w.de, rows_per_cta, w.small_part);
  NRM_LAUNCH_CHECK("small_linear_grad_kernel");
  small_linear_grad_finish_kernel<<<1, 768, 0, s>>>(w.small_part, nparts, grads);
  NRM_LAUNCH_CHECK("small_linear_grad_finish_kernel");
  return NRM_OK;
}

}  // namespace nrm
