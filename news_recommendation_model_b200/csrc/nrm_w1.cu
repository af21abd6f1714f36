// w1 = nn.Linear(66, 64) over the embedded history rows (models/user_invariant_interest_model.py:33,78):
//   forward   xh[r][j]   = sum_i xin[r][i] W1[j][i] + b1[j]
//   backward  dxin[r][i] = sum_j dxh[r][j] W1[j][i];   dW1[j][i] = sum_r dxh[r][j] xin[r][i];   db1[j] = sum_r dxh[r][j]
// NH = B*H rows (51 200 at B=1024, H=50), i.e. a tall-skinny product: tiles of 128 rows staged in shared memory with
// straight 16-byte copies (a tile of rows is one contiguous block of global memory), 8 x 4 / 8 x 5 register tiles.
// The backward runs on a persistent grid: every CTA keeps its share of dW1 / db1 in registers across its tiles and
// writes ONE partial at the end; the finish kernel adds the partials in CTA order (deterministic).
#include "nrm_kernels.cuh"

namespace nrm {

constexpr int W1_ROWS = 128, W1_THREADS = 256;     // backward tiles
constexpr int W1F_ROWS = 64, W1F_THREADS = 128;     // forward tiles: 34 KB of shared memory, six CTAs per SM, the whole grid resident
constexpr int W1_PART = 64 * XIN + 64;          // dW1 [64][66] | db1 [64]

// Asynchronous tile copy (cp.async, 16 bytes per request): every request of the tile is in flight at once; the caller
// waits with copy_wait() before the __syncthreads() that publishes the tile.
template <int THREADS = W1_THREADS>
__device__ __forceinline__ void copy_tile(float* dst, const float* __restrict__ src, int nfloats) {
  // src is 16-byte aligned (tile starts are multiples of 64 rows); nfloats is a multiple of 2
  const int n4 = nfloats >> 2;
  for (int i = threadIdx.x; i < n4; i += THREADS) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 4 * i);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + 4 * i) : "memory");
  }
  for (int i = (n4 << 2) + threadIdx.x; i < nfloats; i += THREADS) dst[i] = __ldg(src + i);
}
__device__ __forceinline__ void copy_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

struct W1SmemFwd { __align__(16) float xs[W1F_ROWS * XIN]; __align__(16) float wt[XIN * 64]; };

__global__ void __launch_bounds__(W1F_THREADS, 6)
w1_forward_kernel(const float* __restrict__ xin, const float* __restrict__ P, const float* __restrict__ w1t, float* __restrict__ xh,
                  long long NH) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char w1_raw[];
  W1SmemFwd& sm = *reinterpret_cast<W1SmemFwd*>(w1_raw);
  const int tid = threadIdx.x, rg = tid >> 4, cg = tid & 15;
  const long long ntiles = (NH + W1F_ROWS - 1) / W1F_ROWS;
  long long tile = blockIdx.x;
  if (tile < ntiles) copy_tile<W1F_THREADS>(sm.xs, xin + tile * W1F_ROWS * XIN, (int)min((long long)W1F_ROWS, NH - tile * W1F_ROWS) * XIN);
  // W1^T [66][64] (transposed once per step by head_transpose_kernel): a straight copy, in flight with the first tile
  copy_tile<W1F_THREADS>(sm.wt, w1t, XIN * 64);
  const float4 bias = __ldg(reinterpret_cast<const float4*>(P + P_W1_B) + cg);
  for (; tile < ntiles; tile += gridDim.x) {
    const long long r0 = tile * W1F_ROWS;
    const int nr = (int)min((long long)W1F_ROWS, NH - r0);
    copy_wait();
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = bias.x; acc[i][1] = bias.y; acc[i][2] = bias.z; acc[i][3] = bias.w; }
    const float* xr = sm.xs + rg * 8 * XIN;
#pragma unroll 2
    for (int k = 0; k < XIN; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(sm.wt + k * 64 + 4 * cg);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = xr[i * XIN + k];
        acc[i][0] = fmaf(a, w.x, acc[i][0]); acc[i][1] = fmaf(a, w.y, acc[i][1]);
        acc[i][2] = fmaf(a, w.z, acc[i][2]); acc[i][3] = fmaf(a, w.w, acc[i][3]);
      }
    }
    __syncthreads();                                  // tile consumed
    const long long nt = tile + gridDim.x;
    if (nt < ntiles) copy_tile<W1F_THREADS>(sm.xs, xin + nt * W1F_ROWS * XIN, (int)min((long long)W1F_ROWS, NH - nt * W1F_ROWS) * XIN);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (rg * 8 + i < nr)
        *reinterpret_cast<float4*>(xh + (r0 + rg * 8 + i) * 64 + 4 * cg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

struct W1SmemBwd { __align__(16) float ds[W1_ROWS * 64]; __align__(16) float xs[W1_ROWS * XIN]; __align__(16) float w[64 * XIN]; };

__global__ void __launch_bounds__(W1_THREADS, 2)
w1_backward_kernel(const float* __restrict__ xin, const float* __restrict__ dxh, const float* __restrict__ P,
                   float* __restrict__ dxin, long long NH, float* __restrict__ part) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char w1_raw[];
  W1SmemBwd& sm = *reinterpret_cast<W1SmemBwd*>(w1_raw);
  const int tid = threadIdx.x, rg = tid >> 4, cg = tid & 15;
  for (int i = tid; i < 64 * XIN; i += W1_THREADS) sm.w[i] = __ldg(P + P_W1_W + i);
  // dW1 share of this thread: rows j = 4 rg .. 4 rg + 3, columns i = cg + 16 t (t < 5; i >= 66 unused)
  float dw[4][5];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int t = 0; t < 5; ++t) dw[a][t] = 0.f;
  float db[4] = {0.f, 0.f, 0.f, 0.f};                 // only cg == 0 threads keep db1
  const int c4 = min(cg + 64, XIN - 1);               // clamped fifth column
  const long long ntiles = (NH + W1_ROWS - 1) / W1_ROWS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r0 = tile * W1_ROWS;
    const int nr = (int)min((long long)W1_ROWS, NH - r0);
    __syncthreads();
    copy_tile(sm.ds, dxh + r0 * 64, nr * 64);
    copy_tile(sm.xs, xin + r0 * XIN, nr * XIN);
    if (nr < W1_ROWS) {                               // ragged last tile: zero rows contribute nothing
      for (int i = nr * 64 + tid; i < W1_ROWS * 64; i += W1_THREADS) sm.ds[i] = 0.f;
      for (int i = nr * XIN + tid; i < W1_ROWS * XIN; i += W1_THREADS) sm.xs[i] = 0.f;
    }
    copy_wait();
    __syncthreads();
    {  // dxin tile: rows rg*8 .. +7, columns cg + 16 t
      float acc[8][5];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int t = 0; t < 5; ++t) acc[i][t] = 0.f;
      const float* dr = sm.ds + rg * 8 * 64;
#pragma unroll 2
      for (int j = 0; j < 64; ++j) {
        float w[5];
#pragma unroll
        for (int t = 0; t < 4; ++t) w[t] = sm.w[j * XIN + cg + 16 * t];
        w[4] = sm.w[j * XIN + c4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = dr[i * 64 + j];
#pragma unroll
          for (int t = 0; t < 5; ++t) acc[i][t] = fmaf(d, w[t], acc[i][t]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (rg * 8 + i < nr) {
          float* dst = dxin + (r0 + rg * 8 + i) * XIN;
#pragma unroll
          for (int t = 0; t < 4; ++t) dst[cg + 16 * t] = acc[i][t];
          if (cg + 64 < XIN) dst[cg + 64] = acc[i][4];
        }
      }
    }
    {  // dW1 / db1 contributions of this tile
#pragma unroll 2
      for (int r = 0; r < W1_ROWS; ++r) {
        const float4 d = *reinterpret_cast<const float4*>(sm.ds + r * 64 + 4 * rg);
        float x[5];
#pragma unroll
        for (int t = 0; t < 4; ++t) x[t] = sm.xs[r * XIN + cg + 16 * t];
        x[4] = sm.xs[r * XIN + c4];
        const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          db[a] += dv[a];
#pragma unroll
          for (int t = 0; t < 5; ++t) dw[a][t] = fmaf(dv[a], x[t], dw[a][t]);
        }
      }
    }
  }
  float* out = part + (long long)blockIdx.x * W1_PART;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int t = 0; t < 4; ++t) out[(4 * rg + a) * XIN + cg + 16 * t] = dw[a][t];
    if (cg + 64 < XIN) out[(4 * rg + a) * XIN + cg + 64] = dw[a][4];
    if (cg == 0) out[64 * XIN + 4 * rg + a] = db[a];
  }
}

// w1.weight / w1.bias are adjacent in the flat layout: one pass over W1_PART entries
__global__ void __launch_bounds__(256)
w1_finish_kernel(const float* __restrict__ part, int nparts, float* __restrict__ grads) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  __shared__ float red[4][64];
  const int lane = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + lane;
  float s = 0.f;
  if (i < W1_PART)
    for (int p = grp; p < nparts; p += 4) s += part[(long long)p * W1_PART + i];
  red[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && i < W1_PART) grads[P_W1_W + i] = ((red[0][lane] + red[1][lane]) + red[2][lane]) + red[3][lane];
}

static_assert(P_W1_B == P_W1_W + 64 * XIN, "w1.weight / w1.bias must be contiguous in the flat layout");

int launch_w1_forward(const float* P, Workspace& w, cudaStream_t s) {
  static DeviceOnce configured;                          // function attributes are per device
  if (configured.first_time()) {
    NRM_CUDA(cudaFuncSetAttribute(w1_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(W1SmemFwd)));
    NRM_CUDA(cudaFuncSetAttribute(w1_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(W1SmemBwd)));
  }
  const long long ntiles = (w.NH + W1F_ROWS - 1) / W1F_ROWS;
  const int grid = (int)min(ntiles, (long long)6 * sm_count());
  launch_pdl(w1_forward_kernel, dim3(grid), dim3(W1F_THREADS), sizeof(W1SmemFwd), s, w.xin_h, P, w.head_wt + HEAD_WT_W1T, w.xh, w.NH);
  NRM_LAUNCH_CHECK("w1_forward_kernel");
  return NRM_OK;
}

int launch_w1_finish(Workspace& w, int nparts, float* G, cudaStream_t s) {
  launch_pdl(w1_finish_kernel, dim3((W1_PART + 63) / 64), dim3(256), 0, s, w.splitk, nparts, G);
  NRM_LAUNCH_CHECK("w1_finish_kernel");
  return NRM_OK;
}

int launch_w1_backward(const float* P, Workspace& w, float* G, cudaStream_t s) {
  static DeviceOnce configured;                          // function attributes are per device
  if (configured.first_time()) {
    NRM_CUDA(cudaFuncSetAttribute(w1_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(W1SmemBwd)));
  }
  const long long ntiles = (w.NH + W1_ROWS - 1) / W1_ROWS;
  const int grid = (int)min(ntiles, (long long)min(2 * sm_count(), W1_SPLITS));
  launch_pdl(w1_backward_kernel, dim3(grid), dim3(W1_THREADS), sizeof(W1SmemBwd), s, w.xin_h, w.dxh, P, w.dxin_h, w.NH, w.splitk);
  NRM_LAUNCH_CHECK("w1_backward_kernel");
  launch_pdl(w1_finish_kernel, dim3((W1_PART + 63) / 64), dim3(256), 0, s, w.splitk, grid, G);
  NRM_LAUNCH_CHECK("w1_finish_kernel");
  return NRM_OK;
}

}  // namespace nrm
