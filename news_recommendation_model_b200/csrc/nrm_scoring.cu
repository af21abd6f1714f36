// Scoring epilogue and submission formatting on the GPU ("next" row N2 of SURVEY.md section 8f).
// The reference finishes every scoring batch on the host, one impression at a time:
//   test.py:58-63    out = mean over the ensemble of softmax(model(x), dim=1)      (over ALL columns of the batch, pads included)
//   test.py:64-70    rows that still hold z pad candidates: softmax(out[0:-z]) AGAIN (a softmax of probabilities: kept, it is
//                    what the shipped submission was ranked with); one .cpu().numpy() per row
//   test.py:118-132  rank string: sorted(enumerate(scores), key=score, reverse=True) -> 1-based rank per candidate, Python's sort
//                    is stable and reverse=True keeps the original order of equal scores, then
//                    "{impression_id} [{r0},{r1},...]\n", formatted by `thread_num` worker processes
// Here: one warp per impression produces the scores and the ranks; a second pass formats the text of a whole batch into one
// byte buffer (lengths -> exclusive scan -> write), so that the host does one D2H copy and one file.write per batch.
#include "nrm_kernels.cuh"

namespace nrm {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// scores[b][i] (i < n_b = C - empty_num[b]) as test.py:58-70; scores of pad columns are 0, their ranks -1.
__global__ void __launch_bounds__(256)
score_epilogue_kernel(const float* __restrict__ logits, int n_models, long long model_stride, long long row_stride, int B, int C,
                      const long long* __restrict__ empty_num, float* __restrict__ scores, int* __restrict__ ranks) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  int z = empty_num ? (int)empty_num[b] : 0;
  z = min(max(z, 0), C);
  const int n = C - z;
  float* s = scores + (long long)b * C;
  // ensemble mean of the per-model softmax over all C columns: out = softmax(m0); out += softmax(m1) ...; out = out / M
  for (int m = 0; m < n_models; ++m) {
    const float* x = logits + (long long)m * model_stride + (long long)b * row_stride;
    float mx = -INFINITY;
    for (int i = lane; i < C; i += 32) mx = fmaxf(mx, x[i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < C; i += 32) sum += expf(x[i] - mx);
    sum = warp_sum(sum);
    for (int i = lane; i < C; i += 32) {
      const float p = expf(x[i] - mx) / sum;
      s[i] = (m == 0) ? p : s[i] + p;
    }
  }
  const float mcount = (float)n_models;
  for (int i = lane; i < C; i += 32) s[i] = s[i] / mcount;
  if (z > 0) {                                   // test.py:65-68: softmax of the averaged probabilities of the real candidates
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, s[i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) sum += expf(s[i] - mx);
    sum = warp_sum(sum);
    for (int i = lane; i < C; i += 32) s[i] = (i < n) ? expf(s[i] - mx) / sum : 0.f;
  }
  __syncwarp();
  // stable descending rank (test.py:124-127): candidates ranked before i have a higher score, or the same score and a lower index
  if (ranks) {
    int* r = ranks + (long long)b * C;
    for (int i = lane; i < C; i += 32) {
      int rank = -1;
      if (i < n) {
        const float si = s[i];
        int above = 0;
        for (int j = 0; j < n; ++j) { const float sj = s[j]; above += (sj > si) || (sj == si && j < i); }
        rank = above + 1;
      }
      r[i] = rank;
    }
  }
}

__device__ __forceinline__ int dec_digits(unsigned long long v) {
  int d = 1;
  while (v >= 10ull) { v /= 10ull; ++d; }
  return d;
}
__device__ __forceinline__ void write_dec(char* dst, unsigned long long v, int digits) {
  for (int k = digits - 1; k >= 0; --k) { dst[k] = (char)('0' + (int)(v % 10ull)); v /= 10ull; }
}

// length of "{id} [{r0},{r1},...]\n" per impression
__global__ void __launch_bounds__(256)
rank_string_length_kernel(const long long* __restrict__ impression_id, const int* __restrict__ ranks,
                          const long long* __restrict__ empty_num, int B, int C, long long* __restrict__ lengths) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  int z = empty_num ? (int)empty_num[b] : 0;
  z = min(max(z, 0), C);
  const int n = C - z;
  int len = 0;
  for (int i = lane; i < n; i += 32) len += dec_digits((unsigned long long)max(ranks[(long long)b * C + i], 0)) + (i + 1 < n ? 1 : 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
  if (lane == 0) {
    const long long id = impression_id[b];
    const unsigned long long mag = id < 0 ? (unsigned long long)(-(id + 1)) + 1ull : (unsigned long long)id;
    lengths[b] = len + dec_digits(mag) + (id < 0 ? 1 : 0) + 4;        // " [" and "]\n"
  }
}

// offsets[0] = 0, offsets[b + 1] = lengths[0] + ... + lengths[b]; one CTA, chunks of 1024 impressions (in place is allowed)
__global__ void __launch_bounds__(1024)
rank_string_scan_kernel(const long long* __restrict__ lengths, int B, long long* __restrict__ offsets) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry = 0; offsets[0] = 0; }
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + tid;
    long long v = i < B ? lengths[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      long long w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const long long incl = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + v;
    if (i < B) offsets[i + 1] = incl;
    __syncthreads();
    if (tid == 1023) carry = incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
rank_string_write_kernel(const long long* __restrict__ impression_id, const int* __restrict__ ranks,
                         const long long* __restrict__ empty_num, int B, int C, const long long* __restrict__ offsets,
                         char* __restrict__ out, long long capacity) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  if (offsets[b + 1] > capacity) return;                 // the caller sees offsets[B] > capacity and reports it
  int z = empty_num ? (int)empty_num[b] : 0;
  z = min(max(z, 0), C);
  const int n = C - z;
  char* dst = out + offsets[b];
  const long long id = impression_id[b];
  const unsigned long long mag = id < 0 ? (unsigned long long)(-(id + 1)) + 1ull : (unsigned long long)id;
  const int idd = dec_digits(mag), neg = id < 0 ? 1 : 0;
  if (lane == 0) {
    if (neg) dst[0] = '-';
    write_dec(dst + neg, mag, idd);
    dst[neg + idd] = ' ';
    dst[neg + idd + 1] = '[';
  }
  int pos = neg + idd + 2;                               // running offset of the next token (same in every lane)
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    int rk = 0, tok = 0;
    if (i < n) { rk = max(ranks[(long long)b * C + i], 0); tok = dec_digits((unsigned long long)rk) + (i + 1 < n ? 1 : 0); }
    int incl = tok;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (i < n) {
      const int d = dec_digits((unsigned long long)rk);
      char* t = dst + pos + incl - tok;
      write_dec(t, (unsigned long long)rk, d);
      if (i + 1 < n) t[d] = ',';
    }
    pos += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) { dst[pos] = ']'; dst[pos + 1] = '\n'; }
}

}  // namespace nrm

using namespace nrm;

extern "C" int nrm_score_epilogue(const float* logits, int n_models, long long model_stride, long long row_stride, int B, int C,
                                  const long long* empty_num, float* scores, int* ranks, void* stream) {
  if (!logits || !scores || n_models <= 0 || B <= 0 || C <= 0 || row_stride < C || (n_models > 1 && model_stride < (long long)B * C)) {
    set_error("nrm_score_epilogue: bad argument"); return NRM_EINVAL;
  }
  score_epilogue_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(logits, n_models, model_stride, row_stride, B, C, empty_num,
                                                                        scores, ranks);
  NRM_LAUNCH_CHECK("score_epilogue_kernel");
  return NRM_OK;
}

// worst case of one line: 20-character id with sign, " [", "]\n", and C tokens of up to 10 digits plus a comma
extern "C" size_t nrm_rank_strings_capacity(int B, int C) {
  if (B <= 0 || C <= 0) return 0;
  return (size_t)B * (size_t)(24 + 11 * (size_t)C);
}

extern "C" int nrm_rank_strings(const long long* impression_id, const int* ranks, const long long* empty_num, int B, int C,
                                long long* offsets, char* out, long long capacity, void* stream) {
  if (!impression_id || !ranks || !offsets || !out || B <= 0 || C <= 0 || capacity <= 0) {
    set_error("nrm_rank_strings: bad argument"); return NRM_EINVAL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  rank_string_length_kernel<<<(B + 7) / 8, 256, 0, s>>>(impression_id, ranks, empty_num, B, C, offsets + 1);
  NRM_LAUNCH_CHECK("rank_string_length_kernel");
  rank_string_scan_kernel<<<1, 1024, 0, s>>>(offsets + 1, B, offsets);
  NRM_LAUNCH_CHECK("rank_string_scan_kernel");
  rank_string_write_kernel<<<(B + 7) / 8, 256, 0, s>>>(impression_id, ranks, empty_num, B, C, offsets, out, capacity);
  NRM_LAUNCH_CHECK("rank_string_write_kernel");
  return NRM_OK;
}
