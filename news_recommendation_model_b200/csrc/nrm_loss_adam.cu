// UserModel.loss (models/user_model.py:37-43) forward + backward, and the fused Adam step
// (torch.optim.Adam with coupled weight decay, train.py:48,74).
#include "nrm_kernels.cuh"

namespace nrm {

constexpr int LOSS_BLOCKS = 64;

struct LossScratch {
  float* dlog;      // [B*C]  d loss / d logits for grad_loss == 1
  float* drow;      // [B]    d loss / d delta[user_id[b]] contribution of impression b
  double* lpart;    // [LOSS_BLOCKS][2]
  unsigned* ticket; // arrival counter of the forward kernel's blocks (zero between calls)
};
static size_t carve_loss(LossScratch& ls, void* base, int B, int C) {
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
  ls.dlog = (float*)take(sizeof(float) * (size_t)B * C);
  ls.drow = (float*)take(sizeof(float) * (size_t)B);
  ls.lpart = (double*)take(sizeof(double) * LOSS_BLOCKS * 2);
  ls.ticket = (unsigned*)take(sizeof(unsigned));
  return off;
}

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

// One softmax + BCE term of one impression, handled by one warp.
// Returns this lane's share of sum_c bce(c); adds coef * dlogit to dl[c]; returns row sum of
// dlogit through *rowsum (all lanes).
__device__ __forceinline__ float bce_softmax_term(const float* __restrict__ out, const double* __restrict__ label, int C,
                                                  float shift, float invN, float coef, float* __restrict__ dl,
                                                  bool accumulate, float* rowsum) {
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, out[c] + shift);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(out[c] + shift - mx);
  sum = warp_sum(sum);
  float loss = 0.f, dot = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float p = expf(out[c] + shift - mx) / sum;
    const float y = (float)label[c];
    loss -= y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f);
    const float dp = (p - y) / fmaxf((1.f - p) * p, 1e-12f) * invN;
    dot = fmaf(dp, p, dot);
  }
  dot = warp_sum(dot);
  float rs = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float p = expf(out[c] + shift - mx) / sum;
    const float y = (float)label[c];
    const float dp = (p - y) / fmaxf((1.f - p) * p, 1e-12f) * invN;
    const float d = coef * p * (dp - dot);
    dl[c] = accumulate ? dl[c] + d : d;
    rs += d;
  }
  *rowsum = warp_sum(rs);
  return loss;
}

__global__ void __launch_bounds__(256)
loss_forward_kernel(const float* __restrict__ logits, const float* __restrict__ delta, const long long* __restrict__ uid,
                    const double* __restrict__ label, int B, int C, long long delta_numel, float alpha, float* __restrict__ dlog,
                    float* __restrict__ drow, double* __restrict__ lpart, unsigned* __restrict__ ticket, float* __restrict__ loss) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  __shared__ double red[8][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invN = 1.0f / ((float)B * (float)C);
  double l1 = 0.0, l2 = 0.0;
  for (int b = blockIdx.x * 8 + warp; b < B; b += gridDim.x * 8) {
    const float* o = logits + (long long)b * C;
    const double* y = label + (long long)b * C;
    float* dl = dlog + (long long)b * C;
    // user id outside delta[0 .. user_num]: the reference raises IndexError (user_model.py:38).  A kernel cannot raise, so the
    // step's loss becomes NaN (loud, no host sync) and the row's shift is 0; the backward skips such ids.
    const long long id = uid[b];
    const bool id_ok = id >= 0 && id < delta_numel;
    const float sh = id_ok ? delta[id] : 0.f;
    if (!id_ok) l2 = __longlong_as_double(0x7ff8000000000000LL);
    float rs1, rs2;
    const float a = bce_softmax_term(o, y, C, 0.f, invN, 1.f - alpha, dl, false, &rs1);
    __syncwarp();
    const float c = bce_softmax_term(o, y, C, sh, invN, alpha, dl, true, &rs2);
    l1 += (double)warp_sum(a);
    l2 += (double)warp_sum(c);
    if (lane == 0) drow[b] = rs2;
  }
  if (lane == 0) { red[warp][0] = l1; red[warp][1] = l2; }
  __syncthreads();
  __shared__ unsigned last_block;
  if (threadIdx.x == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < 8; ++i) { s1 += red[i][0]; s2 += red[i][1]; }
    lpart[blockIdx.x * 2 + 0] = s1; lpart[blockIdx.x * 2 + 1] = s2;
    __threadfence();
    last_block = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  // the block that arrives last adds the block sums (deterministic: lane l takes blocks l, l + 32, ... in order, the lanes are
  // combined in lane order; every load is in flight at once) and leaves the counter at zero
  if (last_block != 0u && warp == 0) {
    __threadfence();
    double s1 = 0.0, s2 = 0.0;
    for (unsigned i = lane; i < gridDim.x; i += 32) { s1 += __ldcg(lpart + i * 2); s2 += __ldcg(lpart + i * 2 + 1); }
    double t1 = 0.0, t2 = 0.0;
    for (int l = 0; l < 32; ++l) { t1 += __shfl_sync(0xffffffffu, s1, l); t2 += __shfl_sync(0xffffffffu, s2, l); }
    if (lane == 0) {
      const double n = (double)B * (double)C;
      const float a = (float)(t1 / n), c = (float)(t2 / n);
      *loss = (1.f - alpha) * a + alpha * c;
      *ticket = 0u;
    }
  }
}

// One launch for the loss backward: blocks [0, nscale) scale the unit gradients by the upstream gradient, the others give
// ddelta[user] = grad_loss * sum of drow over the impressions of that user, in batch order.  One warp per impression:
// it owns the sum iff no earlier impression has the same user id (no atomics, deterministic).
constexpr int LOSS_UID_STAGE = 4096;                  // batches up to this size stage their user ids in shared memory (32 KB)
__global__ void __launch_bounds__(256)
loss_backward_kernel(const float* __restrict__ dlog, const float* __restrict__ grad_loss, long long n, int nscale, float* __restrict__ out,
                     const long long* __restrict__ uid, const float* __restrict__ drow, int B, float* __restrict__ ddelta,
                     long long delta_numel) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  if ((int)blockIdx.x < nscale) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = dlog[i] * __ldg(grad_loss);
    return;
  }
  // the ids of the whole batch go to shared memory first (one coalesced round trip): the duplicate search below is a chain of
  // dependent loads otherwise (12 us for B = 1024 when it ran on global memory)
  extern __shared__ long long s_uid[];
  const bool staged = B <= LOSS_UID_STAGE;
  if (staged) {
    for (int j = threadIdx.x; j < B; j += 256) s_uid[j] = uid[j];
    __syncthreads();
  }
  const long long* ids = staged ? s_uid : uid;
  const int b = (blockIdx.x - nscale) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const long long id = ids[b];
  if (id < 0 || id >= delta_numel) return;
  bool dup = false;
  for (int j0 = 0; j0 < b && !dup; j0 += 32) {
    const int j = j0 + lane;
    dup = __any_sync(0xffffffffu, j < b && ids[j] == id);
  }
  if (dup) return;
  float acc = 0.f;                                   // lane-strided partial sums, combined in lane order
  for (int j = b + lane; j < B; j += 32) if (ids[j] == id) acc += drow[j];
  float tot = 0.f;
  for (int l = 0; l < 32; ++l) tot += __shfl_sync(0xffffffffu, acc, l);
  if (lane == 0) ddelta[id] = tot * __ldg(grad_loss);
}

// Sparse variant for large user tables: dlogits as above; instead of the dense [user_num + 1] gradient, impression b publishes
// (uid[b], d loss / d delta[uid[b]] of that impression) -- duplicates are summed by the consumer (adam_allreduce_kernel).
__global__ void __launch_bounds__(256)
loss_backward_sparse_kernel(const float* __restrict__ dlog, const float* __restrict__ grad_loss, long long n, float* __restrict__ out,
                            const long long* __restrict__ uid, const float* __restrict__ drow, int B, long long delta_numel,
                            long long* __restrict__ suid, float* __restrict__ sval) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const float gl = __ldg(grad_loss);
  if (i < n) out[i] = dlog[i] * gl;
  if (i < B) {
    const long long id = uid[i];
    suid[i] = (id >= 0 && id < delta_numel) ? id : -1;
    sval[i] = drow[i] * gl;
  }
}

// ---------------------------------------------------------------------------------
// Adam
// ---------------------------------------------------------------------------------
struct AdamArgs { float step_size, bc2_sqrt, beta1, beta2, eps, wd, grad_scale; };

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
  g = fmaf(a.wd, p, g * a.grad_scale);                     // grad.add(param, alpha=wd)
  m = m + (g - m) * (1.f - a.beta1);                       // exp_avg.lerp_(grad, 1-beta1)
  v = fmaf((1.f - a.beta2) * g, g, v * a.beta2);           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - a.step_size * (m / denom);                       // param.addcdiv_(m, denom, value=-step_size)
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, AdamArgs a) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, a); adam_one(pp.y, gg.y, mm.y, vv.y, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, a); adam_one(pp.w, gg.w, mm.w, vv.w, a);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
    adam_one(p[i], g[i], m[i], v[i], a);
}

__global__ void __launch_bounds__(256)
adam_scalar_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   long long n, AdamArgs a) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) adam_one(p[i], g[i], m[i], v[i], a);
}

// ---- graph-capturable variant: hyper-parameters and the step counter live on the device ----
struct AdamDeviceState {          // mirrored by FusedTrainStep (python): 1 x int64 + 8 x float32
  long long step;
  float lr, beta1, beta2, eps, wd, grad_scale;
  float step_size, bc2_sqrt;      // derived by adam_prepare_kernel every step
};
static_assert(sizeof(AdamDeviceState) == 40, "AdamDeviceState layout is part of the C ABI");

__global__ void adam_prepare_kernel(AdamDeviceState* st) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long t = ++st->step;
  const double bc1 = 1.0 - pow((double)st->beta1, (double)t);
  const double bc2 = 1.0 - pow((double)st->beta2, (double)t);
  st->step_size = (float)((double)st->lr / bc1);
  st->bc2_sqrt = (float)sqrt(bc2);
}

__global__ void __launch_bounds__(256)
adam_device_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   long long n, const AdamDeviceState* __restrict__ st) {
  pdl_wait();                                          // programmatic dependent launch: see nrm_common.cuh
  pdl_trigger();
  AdamArgs a;
  a.step_size = st->step_size; a.bc2_sqrt = st->bc2_sqrt; a.beta1 = st->beta1; a.beta2 = st->beta2;
  a.eps = st->eps; a.wd = st->wd; a.grad_scale = st->grad_scale;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, a); adam_one(pp.y, gg.y, mm.y, vv.y, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, a); adam_one(pp.w, gg.w, mm.w, vv.w, a);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
    adam_one(p[i], g[i], m[i], v[i], a);
}

// ---------------------------------------------------------------------------------
// Data parallelism over peer memory (NVLink / NVSwitch): the gradient all-reduce fused into the optimizer.
//
// Every rank's flat gradient buffer, a few flag words and its BatchNorm statistics live in SYMMETRIC memory (one allocation per
// rank, mapped into every peer's address space; torch.distributed._symmetric_memory does the mapping, dp.PeerContext fills the
// PeerCtx below).  One step of rank r:
//     backward kernels write grads_r                                     (stream order)
//     adam_allreduce_kernel:  flag "ready(step)" -> every peer;  wait for every peer's flag;
//                             g = (g_0 + g_1 + ... + g_{N-1}) / N  read straight from the N buffers in RANK ORDER (every rank computes
//                             the same bits: replicas cannot drift), Adam on the local weights / moments;
//                             the last block flags "consumed(step)" -> every peer
//     ... next forward ...
//     peer_wait_consumed_kernel: before the next backward overwrites grads_r, every peer has flagged consumed(step)
// No NCCL launch, no second graph, no collective channels competing with the backward kernels for SMs; the transfer (N-1 remote reads
// of 0.9 MB per rank) overlaps the update element by element.  Flags are monotonic step numbers, so CUDA-graph replay is safe.
// ---------------------------------------------------------------------------------
constexpr int PEER_MAX = 8;
constexpr int PEER_FLAG_BASE = 64;           // first u32 word of the signal pad this library uses (the words below belong to torch)
enum { PEER_SLOT_READY = 0, PEER_SLOT_CONSUMED = 1, PEER_SLOT_STATS_FWD = 2, PEER_SLOT_STATS_BWD = 3 };
constexpr int BN_STATS = 2 * E;              // doubles per BatchNorm statistics exchange
struct PeerCtx {                             // mirrored by dp.PeerContext (python): part of the C ABI
  int rank, world;
  long long n;                               // floats of the flat buffers
  const float* grad[PEER_MAX];               // flat gradient buffers, peer-mapped, indexed by rank
  unsigned* pad[PEER_MAX];                   // signal pads (u32 words), peer-mapped
  double* stats[PEER_MAX];                   // [2][BN_STATS]: forward | backward BatchNorm sums, peer-mapped
  // sparse exchange of the per-user bias gradient (delta, user_model.py:23: one float per user, <= B non-zeros per rank and step)
  const long long* suid[PEER_MAX];           // [rows] user id of each impression of the rank (-1: out of range), peer-mapped
  const float* sval[PEER_MAX];               // [rows] d loss / d delta[uid] contributed by that impression, peer-mapped
  long long* acc;                            // [delta_n] local fixed-point accumulator, all zero between steps
  unsigned* gridbar;                         // local: [0] arrival counter, [1] generation of the in-kernel grid barrier
  long long delta_off, delta_n;              // delta's range in the flat buffers; delta_n = 0: no sparse part (delta is averaged densely)
  long long rows;                            // impressions per rank
};
static_assert(sizeof(PeerCtx) == 16 + 5 * 8 * PEER_MAX + 5 * 8, "PeerCtx layout is part of the C ABI");
constexpr double DELTA_FIXED_SCALE = 17592186044416.0;      // 2^44: |sum| < 2^19 fits 63 bits, resolution 5.7e-14

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data changes every step and is cached by the local L1 only (peer addresses bypass the local L2): read it uncached
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f(const float* p) { float v; asm volatile("ld.volatile.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double ld_peer_d(const double* p) { double v; asm volatile("ld.volatile.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory"); return v; }

// threads [0, world) of the calling block: flag `epoch` into slot `slot` of every peer's pad (word index = this rank)
__device__ __forceinline__ void peer_signal(const PeerCtx* c, int slot, unsigned epoch) {
  if ((int)threadIdx.x < c->world) {
    __threadfence_system();
    st_release_sys(c->pad[threadIdx.x] + PEER_FLAG_BASE + slot * PEER_MAX + c->rank, epoch);
  }
}
// threads [0, world): wait until every peer has flagged at least `epoch` in this rank's own pad (local memory).  Bounded: a rank
// that never arrives (a crashed peer) traps after ~10 s instead of wedging the GPU.
__device__ __forceinline__ void peer_wait(const PeerCtx* c, int slot, unsigned epoch) {
  if ((int)threadIdx.x < c->world) {
    const unsigned* p = c->pad[c->rank] + PEER_FLAG_BASE + slot * PEER_MAX + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - epoch) < 0) {
      if (clock64() - t0 > 20000000000LL) asm volatile("trap;\n");
    }
  }
}

__global__ void __launch_bounds__(256)
adam_allreduce_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const AdamDeviceState* __restrict__ st,
                      const PeerCtx* __restrict__ c, unsigned* __restrict__ ticket) {
  pdl_wait();
  pdl_trigger();
  __shared__ int is_last;
  const unsigned epoch = (unsigned)st->step;                 // adam_prepare_kernel has already counted this step
  if (blockIdx.x == 0) peer_signal(c, PEER_SLOT_READY, epoch);      // my gradients are complete (stream order)
  peer_wait(c, PEER_SLOT_READY, epoch);
  __syncthreads();
  AdamArgs a;
  a.step_size = st->step_size; a.bc2_sqrt = st->bc2_sqrt; a.beta1 = st->beta1; a.beta2 = st->beta2;
  a.eps = st->eps; a.wd = st->wd; a.grad_scale = st->grad_scale;
  const int world = c->world;
  const float inv = 1.0f / (float)world;
  const bool sparse = c->delta_n > 0;
  const long long n = sparse ? c->delta_off : c->n, n4 = n >> 2;     // densely averaged range (delta_off is a multiple of 4)
  const long long stride = (long long)gridDim.x * 256;
  if (sparse) {
    // every rank's (user id, value) list -> local fixed-point accumulator.  Integer atomics: the sum does not depend on the order,
    // so every rank ends up with the same bits.
    const long long rows = c->rows, total = rows * world;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += stride) {
      const int r = (int)(e / rows);
      const long long b = e - r * rows;
      long long id; float val;
      asm volatile("ld.volatile.global.s64 %0, [%1];\n" : "=l"(id) : "l"(c->suid[r] + b) : "memory");
      val = ld_peer_f(c->sval[r] + b);
      if (id >= 0 && id < c->delta_n) atomicAdd(reinterpret_cast<unsigned long long*>(c->acc + id), (unsigned long long)__double2ll_rn((double)val * DELTA_FIXED_SCALE));
    }
  }
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
    // all ranks' loads in flight at once (a load-add loop serialises one NVLink round trip per rank), summed in rank order
    float4 x[PEER_MAX];
#pragma unroll
    for (int r = 0; r < PEER_MAX; ++r)
      if (r < world) x[r] = ld_peer_f4(c->grad[r] + 4 * i);
    float4 gg = x[0];
#pragma unroll
    for (int r = 1; r < PEER_MAX; ++r)
      if (r < world) { gg.x += x[r].x; gg.y += x[r].y; gg.z += x[r].z; gg.w += x[r].w; }
    gg.x *= inv; gg.y *= inv; gg.z *= inv; gg.w *= inv;
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, a); adam_one(pp.y, gg.y, mm.y, vv.y, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, a); adam_one(pp.w, gg.w, mm.w, vv.w, a);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    float g = ld_peer_f(c->grad[0] + i);
    for (int r = 1; r < world; ++r) g += ld_peer_f(c->grad[r] + i);
    adam_one(p[i], g * inv, m[i], v[i], a);
  }
  if (sparse) {
    // grid barrier (every block is resident: the launcher caps the grid), then the dense Adam pass over delta with the accumulated
    // gradient; touched accumulator entries go back to zero
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned gen = *reinterpret_cast<volatile unsigned*>(c->gridbar + 1);
      if (atomicAdd(c->gridbar, 1u) == gridDim.x - 1) {
        *reinterpret_cast<volatile unsigned*>(c->gridbar) = 0u;
        __threadfence();
        atomicAdd(c->gridbar + 1, 1u);
      } else {
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned*>(c->gridbar + 1) == gen) {
          if (clock64() - t0 > 20000000000LL) asm volatile("trap;\n");
        }
      }
      __threadfence();
    }
    __syncthreads();
    const double unscale = 1.0 / DELTA_FIXED_SCALE;
    float* dp = p + c->delta_off; float* dm = m + c->delta_off; float* dv = v + c->delta_off;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < c->delta_n; i += stride) {
      const long long acc = c->acc[i];
      if (acc != 0) c->acc[i] = 0;
      adam_one(dp[i], (float)((double)acc * unscale) * inv, dm[i], dv[i], a);
    }
  }
  // the block that finishes last tells every peer that this rank no longer reads their gradients of this step
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) *ticket = 0u;
    peer_signal(c, PEER_SLOT_CONSUMED, epoch);
  }
}

// one small block, in front of the next backward: every peer has consumed this rank's gradients of the previous step
__global__ void peer_wait_consumed_kernel(const AdamDeviceState* __restrict__ st, const PeerCtx* __restrict__ c) {
  pdl_wait();
  pdl_trigger();
  peer_wait(c, PEER_SLOT_CONSUMED, (unsigned)st->step);
}

// BatchNorm statistics (2 x 264 doubles) summed over the ranks in rank order: local -> this rank's symmetric slot -> flags -> sum.
// which = 0: forward (column sums of e_concat), 1: backward.  One block of 544 threads.
__global__ void __launch_bounds__(544)
peer_allsum_stats_kernel(const double* __restrict__ local, int which, double* __restrict__ out, const AdamDeviceState* __restrict__ st,
                         const PeerCtx* __restrict__ c) {
  pdl_wait();
  pdl_trigger();
  const unsigned epoch = (unsigned)st->step + 1u;              // the step in flight (adam_prepare_kernel counts it at its end)
  const int t = threadIdx.x;
  double* mine = c->stats[c->rank] + which * BN_STATS;
  if (t < BN_STATS) mine[t] = local[t];
  __threadfence_system();
  __syncthreads();
  peer_signal(c, PEER_SLOT_STATS_FWD + which, epoch);
  peer_wait(c, PEER_SLOT_STATS_FWD + which, epoch);
  __syncthreads();
  if (t < BN_STATS) {
    double x[PEER_MAX];
#pragma unroll
    for (int r = 0; r < PEER_MAX; ++r)
      if (r < c->world) x[r] = ld_peer_d(c->stats[r] + which * BN_STATS + t);      // every rank's value in flight at once
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < PEER_MAX; ++r)
      if (r < c->world) s += x[r];
    out[t] = s;
  }
}

}  // namespace nrm

using namespace nrm;

extern "C" size_t nrm_peer_ctx_bytes(void) { return sizeof(PeerCtx); }
extern "C" size_t nrm_peer_stats_bytes(void) { return sizeof(double) * 2 * BN_STATS; }
extern "C" int nrm_peer_flag_words(void) { return PEER_FLAG_BASE + 4 * PEER_MAX; }
// Loads the peer kernels' code (CUDA loads kernels lazily on first use; a first use inside a stream capture is best avoided).
extern "C" int nrm_peer_preload(void) {
  cudaFuncAttributes a;
  NRM_CUDA(cudaFuncGetAttributes(&a, adam_allreduce_kernel));
  NRM_CUDA(cudaFuncGetAttributes(&a, peer_wait_consumed_kernel));
  NRM_CUDA(cudaFuncGetAttributes(&a, peer_allsum_stats_kernel));
  NRM_CUDA(cudaFuncGetAttributes(&a, loss_backward_sparse_kernel));
  return NRM_OK;
}

// Adam with the gradient average over the ranks fused in (see the block comment above).  `peer_ctx`: device PeerCtx;
// `ticket`: one zeroed u32 of device scratch.  The caller's own gradient buffer is peer_ctx->grad[rank].
extern "C" int nrm_adam_step_allreduce(float* param, float* exp_avg, float* exp_avg_sq, long long n, void* adam_state,
                                       const void* peer_ctx, void* ticket, void* stream) {
  if (!param || !exp_avg || !exp_avg_sq || !adam_state || !peer_ctx || !ticket || n <= 0) { set_error("nrm_adam_step_allreduce: bad argument"); return NRM_EINVAL; }
  if ((((uintptr_t)param | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) { set_error("nrm_adam_step_allreduce: buffers must be 16-byte aligned"); return NRM_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  AdamDeviceState* st = (AdamDeviceState*)adam_state;
  launch_pdl(adam_prepare_kernel, dim3(1), dim3(32), 0, s, st);
  NRM_LAUNCH_CHECK("adam_prepare_kernel");
  long long blocks = ((n + 3) / 4 + 255) / 256;
  // every block must be resident (each one waits for the peers' flags, and the sparse path has a grid barrier): the cap comes from
  // the occupancy the driver reports for this kernel, not from an assumed register count
  static DeviceOnce configured;
  static int per_sm[64];
  const int dev = current_device();
  if (configured.first_time() || per_sm[dev & 63] == 0) {
    int nb = 0;
    NRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, adam_allreduce_kernel, 256, 0));
    per_sm[dev & 63] = nb > 0 ? nb : 1;
  }
  const long long cap = (long long)sm_count() * per_sm[dev & 63];
  if (blocks > cap) blocks = cap;
  KernelTimer t("adam", s);
  launch_pdl(adam_allreduce_kernel, dim3((int)blocks), dim3(256), 0, s, param, exp_avg, exp_avg_sq, st, (const PeerCtx*)peer_ctx, (unsigned*)ticket);
  NRM_LAUNCH_CHECK("adam_allreduce_kernel");
  return NRM_OK;
}
extern "C" int nrm_peer_wait_consumed(const void* adam_state, const void* peer_ctx, void* stream) {
  if (!adam_state || !peer_ctx) { set_error("nrm_peer_wait_consumed: bad argument"); return NRM_EINVAL; }
  launch_pdl(peer_wait_consumed_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (const AdamDeviceState*)adam_state, (const PeerCtx*)peer_ctx);
  NRM_LAUNCH_CHECK("peer_wait_consumed_kernel");
  return NRM_OK;
}
extern "C" int nrm_peer_allsum_stats(const double* local, int which, double* out, const void* adam_state, const void* peer_ctx, void* stream) {
  if (!local || !out || !adam_state || !peer_ctx || which < 0 || which > 1) { set_error("nrm_peer_allsum_stats: bad argument"); return NRM_EINVAL; }
  launch_pdl(peer_allsum_stats_kernel, dim3(1), dim3(544), 0, (cudaStream_t)stream, local, which, out, (const AdamDeviceState*)adam_state,
             (const PeerCtx*)peer_ctx);
  NRM_LAUNCH_CHECK("peer_allsum_stats_kernel");
  return NRM_OK;
}

extern "C" int nrm_adam_step_device(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                    void* adam_state, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !adam_state || n < 0) { set_error("nrm_adam_step_device: bad argument"); return NRM_EINVAL; }
  if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) {
    set_error("nrm_adam_step_device: buffers must be 16-byte aligned"); return NRM_EINVAL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  AdamDeviceState* st = (AdamDeviceState*)adam_state;
  launch_pdl(adam_prepare_kernel, dim3(1), dim3(32), 0, s, st);
  NRM_LAUNCH_CHECK("adam_prepare_kernel");
  if (n == 0) return NRM_OK;
  long long blocks = ((n + 3) / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  KernelTimer t("adam", s);
  launch_pdl(adam_device_kernel, dim3((int)blocks), dim3(256), 0, s, param, grad, exp_avg, exp_avg_sq, n, st);
  NRM_LAUNCH_CHECK("adam_device_kernel");
  return NRM_OK;
}

extern "C" size_t nrm_loss_scratch_bytes(int B, int C) {
  LossScratch ls;
  return carve_loss(ls, nullptr, B > 0 ? B : 1, C > 0 ? C : 1);
}

extern "C" int nrm_loss_forward(const float* logits, const float* delta, long long delta_numel, const long long* user_id,
                                const double* label, int B, int C, float alpha, float* loss, void* scratch, size_t scratch_bytes,
                                void* stream) {
  if (!logits || !delta || delta_numel <= 0 || !user_id || !label || !loss || !scratch || B <= 0 || C <= 0) {
    set_error("nrm_loss_forward: bad argument"); return NRM_EINVAL;
  }
  LossScratch ls;
  if (carve_loss(ls, scratch, B, C) > scratch_bytes) { set_error("nrm_loss_forward: scratch too small"); return NRM_EWORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = min(LOSS_BLOCKS, (B + 7) / 8);
  launch_pdl(loss_forward_kernel, dim3(blocks), dim3(256), 0, s, logits, delta, user_id, label, B, C, delta_numel, alpha, ls.dlog, ls.drow, ls.lpart,
             ls.ticket, loss);
  NRM_LAUNCH_CHECK("loss_forward_kernel");
  return NRM_OK;
}

extern "C" int nrm_loss_backward(const long long* user_id, int B, int C, const float* grad_loss, float* dlogits,
                                 float* ddelta, long long delta_numel, const void* scratch, size_t scratch_bytes, void* stream) {
  if (!user_id || !grad_loss || !dlogits || !scratch || B <= 0 || C <= 0) {
    set_error("nrm_loss_backward: bad argument"); return NRM_EINVAL;
  }
  LossScratch ls;
  if (carve_loss(ls, const_cast<void*>(scratch), B, C) > scratch_bytes) { set_error("nrm_loss_backward: scratch too small"); return NRM_EWORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)B * C;
  const int nscale = (int)((n + 255) / 256);
  const bool want_delta = ddelta != nullptr && delta_numel > 0;
  if (want_delta) NRM_CUDA(cudaMemsetAsync(ddelta, 0, sizeof(float) * (size_t)delta_numel, s));
  launch_pdl(loss_backward_kernel, dim3(nscale + (want_delta ? (B + 7) / 8 : 0)), dim3(256),
             (want_delta && B <= LOSS_UID_STAGE) ? sizeof(long long) * (size_t)B : 0, s, ls.dlog, grad_loss, n, nscale, dlogits,
             user_id, ls.drow, B, ddelta, delta_numel);
  NRM_LAUNCH_CHECK("loss_backward_kernel");
  return NRM_OK;
}

extern "C" int nrm_loss_backward_sparse(const long long* user_id, int B, int C, const float* grad_loss, float* dlogits, long long delta_numel,
                                        long long* sparse_uid, float* sparse_val, const void* scratch, size_t scratch_bytes, void* stream) {
  if (!user_id || !grad_loss || !dlogits || !sparse_uid || !sparse_val || !scratch || B <= 0 || C <= 0) {
    set_error("nrm_loss_backward_sparse: bad argument"); return NRM_EINVAL;
  }
  LossScratch ls;
  if (carve_loss(ls, const_cast<void*>(scratch), B, C) > scratch_bytes) { set_error("nrm_loss_backward_sparse: scratch too small"); return NRM_EWORKSPACE; }
  const long long n = (long long)B * C;
  launch_pdl(loss_backward_sparse_kernel, dim3((int)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, ls.dlog, grad_loss, n, dlogits, user_id,
             ls.drow, B, delta_numel, sparse_uid, sparse_val);
  NRM_LAUNCH_CHECK("loss_backward_sparse_kernel");
  return NRM_OK;
}

extern "C" int nrm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, long long step, float grad_scale,
                             void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1) { set_error("nrm_adam_step: bad argument"); return NRM_EINVAL; }
  if (n == 0) return NRM_OK;
  AdamArgs a;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.grad_scale = grad_scale;
  cudaStream_t s = (cudaStream_t)stream;
  const bool aligned = (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0;
  const long long work = aligned ? (n + 3) / 4 : n;
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  KernelTimer t("adam", s);
  if (aligned) launch_pdl(adam_kernel, dim3((int)blocks), dim3(256), 0, s, param, grad, exp_avg, exp_avg_sq, n, a);
  else launch_pdl(adam_scalar_kernel, dim3((int)blocks), dim3(256), 0, s, param, grad, exp_avg, exp_avg_sq, n, a);
  NRM_LAUNCH_CHECK("adam_kernel");
  return NRM_OK;
}
