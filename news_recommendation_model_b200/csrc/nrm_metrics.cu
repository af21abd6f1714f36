// Batched per-impression ranking metrics on the GPU ("next" row N1 of SURVEY.md section 8f).
// The reference evaluates one impression at a time on the host: train.py:77-80 calls sklearn's roc_auc_score once per
// sample of every training batch (B host calls + a device sync per step), verify.py:25-37 once per validation
// impression (tool/evaluation.py:3-5).  Here one warp scores one impression:
//   auc   = ( #{(p,n): s_p > s_n} + 0.5 #{(p,n): s_p == s_n} ) / (n_pos n_neg)   -- sklearn's tie-averaged ROC AUC
//   hit   = argmax(score) == argmax(label)        (first maximum, as numpy.argmax; verify.py:32)
//   rr    = 1 / rank of the best-ranked positive  (descending score, ties keep index order like test.py:124-127)
//   ndcg  = DCG@k / IDCG@k with binary gains      (MRR / nDCG do not exist in the reference: parity unpinned)
// Impressions with no positive or no negative candidate get auc = NaN (sklearn raises there).
#include "nrm_kernels.cuh"

namespace nrm {

__global__ void __launch_bounds__(256)
batch_metrics_kernel(const float* __restrict__ scores, long long score_stride, const double* __restrict__ labels,
                     long long label_stride, const int* __restrict__ n_valid, int B, int C, int k,
                     float* __restrict__ auc, float* __restrict__ hit, float* __restrict__ rr, float* __restrict__ ndcg) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const int n = n_valid ? min(max(n_valid[b], 0), C) : C;
  const float* s = scores + (long long)b * score_stride;
  const double* y = labels + (long long)b * label_stride;
  float gt = 0.f, eq = 0.f;                  // pair counts
  int npos = 0, nneg = 0;
  float best_rr = 0.f, dcg = 0.f;
  // argmax of score / label (first maximum)
  float smax = -INFINITY; int sarg = 0x7fffffff; double ymax = -1e300; int yarg = 0x7fffffff;
  for (int i = lane; i < n; i += 32) {
    const float si = s[i];
    const bool pos = y[i] > 0.5;
    if (si > smax) { smax = si; sarg = i; }
    if (y[i] > ymax) { ymax = y[i]; yarg = i; }
    npos += pos; nneg += !pos;
    if (pos) {
      int above = 0;                          // candidates ranked before i: higher score, or equal score and lower index
      float g = 0.f, e = 0.f;
      for (int j = 0; j < n; ++j) {
        const float sj = s[j];
        const bool negj = !(y[j] > 0.5);
        above += (sj > si) || (sj == si && j < i);
        if (negj) { g += (si > sj) ? 1.f : 0.f; e += (si == sj) ? 1.f : 0.f; }
      }
      gt += g; eq += e;
      const int rank = above + 1;
      best_rr = fmaxf(best_rr, 1.0f / (float)rank);
      if (rank <= k) dcg += 1.0f / log2f(1.0f + (float)rank);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gt += __shfl_xor_sync(0xffffffffu, gt, o); eq += __shfl_xor_sync(0xffffffffu, eq, o);
    npos += __shfl_xor_sync(0xffffffffu, npos, o); nneg += __shfl_xor_sync(0xffffffffu, nneg, o);
    best_rr = fmaxf(best_rr, __shfl_xor_sync(0xffffffffu, best_rr, o));
    dcg += __shfl_xor_sync(0xffffffffu, dcg, o);
    const float os = __shfl_xor_sync(0xffffffffu, smax, o); const int oa = __shfl_xor_sync(0xffffffffu, sarg, o);
    if (os > smax || (os == smax && oa < sarg)) { smax = os; sarg = oa; }
    const double oy = __shfl_xor_sync(0xffffffffu, ymax, o); const int ob = __shfl_xor_sync(0xffffffffu, yarg, o);
    if (oy > ymax || (oy == ymax && ob < yarg)) { ymax = oy; yarg = ob; }
  }
  if (lane != 0) return;
  if (auc) auc[b] = (npos > 0 && nneg > 0) ? (gt + 0.5f * eq) / ((float)npos * (float)nneg) : __int_as_float(0x7fc00000);
  if (hit) hit[b] = (n > 0 && sarg == yarg) ? 1.f : 0.f;
  if (rr) rr[b] = best_rr;
  if (ndcg) {
    float idcg = 0.f;
    for (int i = 1; i <= min(npos, k); ++i) idcg += 1.0f / log2f(1.0f + (float)i);
    ndcg[b] = idcg > 0.f ? dcg / idcg : 0.f;
  }
}

}  // namespace nrm

using namespace nrm;

extern "C" int nrm_batch_metrics(const float* scores, long long score_stride, const double* labels, long long label_stride,
                                 const int* n_valid, int B, int C, int k, float* auc, float* hit, float* rr, float* ndcg,
                                 void* stream) {
  if (!scores || !labels || B <= 0 || C <= 0 || k <= 0 || score_stride < C || label_stride < C) {
    set_error("nrm_batch_metrics: bad argument"); return NRM_EINVAL;
  }
  batch_metrics_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(scores, score_stride, labels, label_stride, n_valid, B, C, k,
                                                                       auc, hit, rr, ndcg);
  NRM_LAUNCH_CHECK("batch_metrics_kernel");
  return NRM_OK;
}
