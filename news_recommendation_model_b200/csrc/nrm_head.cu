// Scoring head of UserModel (models/user_model.py:31-35): BatchNorm1d(264) -> gate MLP ->
// (gate * e_concat) -> mlp -> out_mlp, forward and backward, plus the generic reduction
// kernels.  The five 264<->66 layers run on the shared fp32 GEMM (nrm_gemm.cuh) with the
// bias / GELU / gating fused into its epilogues.
#include "nrm_kernels.cuh"
#include "nrm_gemm.cuh"

namespace nrm {

__global__ void reduce_splits_kernel(const float* __restrict__ src, int nsplit, long long stride,
                                     float* __restrict__ dst, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float acc = 0.f;
  for (int z = 0; z < nsplit; ++z) acc += src[z * stride + i];
  dst[i] = acc;
}

// Sum the split partials of the head weight/bias gradients (region [P_GATE_FC1_W, P_DELTA) of
// the flat layout, mirrored in every partial) into the flat gradient.  Alignment padding
// between entries is never written by the GEMMs, so it is skipped here (stays zero).
__global__ void __launch_bounds__(256)
reduce_head_splits_kernel(const float* __restrict__ part, int nsplit, float* __restrict__ grads) {
  constexpr long long BEG = P_GATE_FC1_W, LEN = P_DELTA - P_GATE_FC1_W;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= LEN) return;
  const long long o = BEG + i;
  // entries with a padded tail: 66-wide biases / out_mlp.fc2.weight (2 pad floats) and the scalar bias (3)
  const bool pad = (o >= P_GATE_FC1_B + HID && o < P_GATE_FC2_W) || (o >= P_MLP_FC1_B + HID && o < P_MLP_FC2_W) ||
                   (o >= P_OUT_FC1_B + HID && o < P_OUT_FC2_W) || (o >= P_OUT_FC2_W + HID && o < P_OUT_FC2_B) ||
                   (o >= P_OUT_FC2_B + 1);
  if (pad) return;
  float acc = 0.f;
  for (int z = 0; z < nsplit; ++z) acc += part[z * LEN + i];
  grads[o] = acc;
}

// ---------------------------------------------------------------------------------
// BatchNorm statistics.  Column sums and sums of squares in double so that
// var = E[x^2] - mean^2 is safe (|mean| reaches 25 with var < 1 on some channels).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(E)
bn_partial_kernel(const float* __restrict__ e, long long R, int rows_per_chunk, double* __restrict__ part) {
  const int n = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(R, r0 + rows_per_chunk);
  double s = 0.0, q = 0.0;
#pragma unroll 8
  for (long long r = r0; r < r1; ++r) { const double v = (double)__ldg(e + r * E + n); s += v; q += v * v; }
  part[((long long)blockIdx.x * 2 + 0) * E + n] = s;
  part[((long long)blockIdx.x * 2 + 1) * E + n] = q;
}

// sums[i] = sum over parts of part[p][i], i < 2*264; block = 66 entries x 4 interleaved groups of partials combined in
// group order (8 blocks)
__global__ void __launch_bounds__(264)
bn_partial_reduce_kernel(const double* __restrict__ part, int nparts, double* __restrict__ sums) {
  __shared__ double red[4][66];
  const int lane = threadIdx.x % 66, grp = threadIdx.x / 66;
  const int i = blockIdx.x * 66 + lane;
  double s = 0.0;
#pragma unroll 4
  for (int p = grp; p < nparts; p += 4) s += part[(long long)p * 2 * E + i];
  red[grp][lane] = s;
  __syncthreads();
  if (grp == 0) sums[i] = ((red[0][lane] + red[1][lane]) + red[2][lane]) + red[3][lane];
}

// training: batch mean / biased variance from (global) sums; running stats with momentum
// 0.1 and the unbiased variance (nn.BatchNorm1d);  eval: running statistics.
__global__ void __launch_bounds__(E)
bn_finalize_kernel(const double* __restrict__ sums, long long rows, int training, float* __restrict__ run_mean,
                   float* __restrict__ run_var, long long* __restrict__ nbt, float* __restrict__ mean,
                   float* __restrict__ rstd) {
  const int n = threadIdx.x;
  float m, v;
  if (training) {
    const double dm = sums[n] / (double)rows;
    double dv = sums[E + n] / (double)rows - dm * dm;
    if (dv < 0.0) dv = 0.0;
    m = (float)dm; v = (float)dv;
    const double unbiased = rows > 1 ? dv * (double)rows / (double)(rows - 1) : dv;
    run_mean[n] = (1.f - BN_MOMENTUM) * run_mean[n] + BN_MOMENTUM * m;
    run_var[n] = (1.f - BN_MOMENTUM) * run_var[n] + BN_MOMENTUM * (float)unbiased;
    if (n == 0) *nbt += 1;
  } else {
    m = run_mean[n]; v = run_var[n];
  }
  mean[n] = m;
  rstd[n] = 1.0f / sqrtf(v + BN_EPS);
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ e, const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, long long total, float* __restrict__ z) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int n = (int)(i % E);
  z[i] = (e[i] - mean[n]) * rstd[n] * gamma[n] + beta[n];
}

// r[m] = u3[m,:] . w + b   (out_mlp.fc2), one warp per row
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ b, long long M,
              float* __restrict__ out) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc = 0.f;
  for (int n = lane; n < HID; n += 32) acc = fmaf(u[m * HID + n], __ldg(w + n), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[m] = acc + __ldg(b);
}

// da3[m,n] = dr[m] * O2[n] * gelu'(a3[m,n])
__global__ void __launch_bounds__(256)
out_fc2_backward_kernel(const float* __restrict__ dr, const float* __restrict__ w, const float* __restrict__ a3,
                        long long total, float* __restrict__ da3) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const long long m = i / HID; const int n = (int)(i % HID);
  da3[i] = dr[m] * __ldg(w + n) * gelu_grad_f(a3[i]);
}

// column sums of dz and dz * xhat over a row chunk (double), for the BatchNorm backward
__global__ void __launch_bounds__(E)
bn_bwd_partial_kernel(const float* __restrict__ dz, const float* __restrict__ e, const float* __restrict__ mean,
                      const float* __restrict__ rstd, long long R, int rows_per_chunk, double* __restrict__ part) {
  const int n = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(R, r0 + rows_per_chunk);
  const float mu = mean[n], rs = rstd[n];
  double s = 0.0, q = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const float d = dz[r * E + n];
    const float xh = (e[r * E + n] - mu) * rs;
    s += (double)d; q += (double)(d * xh);
  }
  part[((long long)blockIdx.x * 2 + 0) * E + n] = s;
  part[((long long)blockIdx.x * 2 + 1) * E + n] = q;
}

// bn.weight / bn.bias gradients from this rank's sums
__global__ void __launch_bounds__(E)
bn_param_grad_kernel(const double* __restrict__ sums, float* __restrict__ grads) {
  const int n = threadIdx.x;
  grads[P_BN_B + n] = (float)sums[n];
  grads[P_BN_W + n] = (float)sums[E + n];
}

// de += rstd * gamma * (dz - mean_r(dz) - xhat * mean_r(dz * xhat))      (training)
// de += rstd * gamma * dz                                                 (eval)
__global__ void __launch_bounds__(256)
bn_bwd_combine_kernel(const float* __restrict__ dz, const float* __restrict__ e, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma, const double* __restrict__ sums,
                      long long rows, int training, long long total, float* __restrict__ de) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int n = (int)(i % E);
  const float rs = rstd[n], g = gamma[n];
  float v = dz[i];
  if (training) {
    const float m1 = (float)(sums[n] / (double)rows), m2 = (float)(sums[E + n] / (double)rows);
    const float xh = (e[i] - mean[n]) * rs;
    v = v - m1 - xh * m2;
  }
  de[i] += rs * g * v;
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
static inline int stat_rows(long long R) { return (int)((R + STAT_BLOCKS - 1) / STAT_BLOCKS); }
static inline int stat_chunks(long long R) { const int rp = stat_rows(R); return (int)((R + rp - 1) / rp); }

int launch_bn_partial_sums(Workspace& w, cudaStream_t s) {
  const int rp = stat_rows(w.R), nch = stat_chunks(w.R);
  bn_partial_kernel<<<nch, E, 0, s>>>(w.e, w.R, rp, w.stat_part);
  NRM_LAUNCH_CHECK("bn_partial_kernel");
  bn_partial_reduce_kernel<<<2 * E / 66, 264, 0, s>>>(w.stat_part, nch, w.bn_sums);
  NRM_LAUNCH_CHECK("bn_partial_reduce_kernel");
  return NRM_OK;
}

static GemmArgs linear_fwd(const float* X, int K, const float* W, const float* bias, float* Y, float* Y2, int N, long long M) {
  GemmArgs g{};
  g.M = (int)M; g.N = N; g.K = K;
  g.A = X; g.sam = K; g.sak = 1;
  g.B = W; g.sbk = 1; g.sbn = K;          // W is [N,K] row-major: B(k,n) = W[n*K + k]
  g.C = Y; g.scm = N; g.scn = 1; g.C2 = Y2;
  g.bias = bias;
  return g;
}
// dX[M,K] = dY[M,N] W[N,K]
static GemmArgs linear_bwd_data(const float* dY, int N, const float* W, float* dX, float* dX2, int K, long long M) {
  GemmArgs g{};
  g.M = (int)M; g.N = K; g.K = N;
  g.A = dY; g.sam = N; g.sak = 1;
  g.B = W; g.sbk = K; g.sbn = 1;          // B(k=n', n=k') = W[n'*K + k']
  g.C = dX; g.scm = K; g.scn = 1; g.C2 = dX2;
  return g;
}

int launch_head_forward(const float* P, Workspace& w, float* run_mean, float* run_var, long long* nbt, int training, int keep,
                        const double* bn_sums, long long global_rows, float* logits, cudaStream_t s) {
  bn_finalize_kernel<<<1, E, 0, s>>>(bn_sums, global_rows, training, run_mean, run_var, nbt, w.mean, w.rstd);
  NRM_LAUNCH_CHECK("bn_finalize_kernel");
  return launch_head_forward_fused(P, w, keep, logits, s);
}

// dW (layout of the nn.Linear weight, [out,in]) = dY^T X and db = colsum(dY), summed over rows by
// a split GEMM whose partials mirror the flat layout of the head region (one reduce at the end).
// X: [R,in], dY: [R,out].  Returns the number of splits (same for every head layer) or < 0.
static int weight_grad(const float* dY, int out, const float* X, int in, long long R, long long w_off, long long b_off,
                       Workspace& w, cudaStream_t s) {
  constexpr long long BEG = P_GATE_FC1_W, LEN = P_DELTA - P_GATE_FC1_W;
  GemmArgs g{};
  g.K = (int)R;
  if (out == HID || out == 1) {        // dW[o][i]: put the narrow dimension on N; bias = virtual ones row of A
    g.M = in; g.N = out;
    g.A = X; g.sam = 1; g.sak = in;
    g.B = dY; g.sbk = out; g.sbn = 1;
    g.scm = 1; g.scn = in;
    g.ones_row = 1;
  } else {                             // bias = virtual ones column of B
    g.M = out; g.N = in;
    g.A = dY; g.sam = 1; g.sak = out;
    g.B = X; g.sbk = in; g.sbn = 1;
    g.scm = in; g.scn = 1;
    g.ones_col = 1;
  }
  g.C = w.splitk + (w_off - BEG);
  g.Cb = w.splitk + (b_off - BEG);
  g.split_stride = LEN;
  return launch_gemm<EPI_NONE>(g, WGRAD_SPLITS, s);
}

int launch_head_backward(const float* P, Workspace& w, const float* dlogits, float* G, cudaStream_t s) {
  return launch_head_backward_fused(P, w, dlogits, G, s);
}

int launch_bn_backward_combine(const float* P, Workspace& w, int training, const double* bn_bwd_sums,
                               long long global_rows, cudaStream_t s) {
  const long long total = w.R * E;
  bn_bwd_combine_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(w.dz, w.e, w.mean, w.rstd, P + P_BN_W, bn_bwd_sums,
                                                                  global_rows, training, total, w.de);
  NRM_LAUNCH_CHECK("bn_bwd_combine_kernel");
  return NRM_OK;
}

}  // namespace nrm
