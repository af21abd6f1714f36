// Shared pieces of the row-stacked attention kernels (nrm_attention_rs.cu: forward, nrm_attention_rs_bwd.cu: backward): warp roles,
// work-unit geometry, the weight image, and the inline PTX that nrm_umma.cuh does not have (tensor-memory stores, MMA with the A
// operand in tensor memory, cp.async).
#pragma once
#include "nrm_kernels.cuh"
#include <cstddef>

#include "nrm_umma.cuh"

namespace nrm {
namespace rs {

// Optional per-role wait accounting (make EXTRA=-DNRM_RS_PROFILE; tools/rs_roleprof.py): in CTA 0, lane 0 of one warp per role
// adds the clock64 cycles it spends inside each kind of mbarrier wait to g_rsprof[role * 8 + kind]; slot 7 = the role's total.
#ifdef NRM_RS_PROFILE
static __device__ long long g_rsprof[64];      // one copy per translation unit; rsprof_read returns the forward kernels' copy
#define RSPROF_WAIT(role, kind, stmt) do { const long long t__ = clock64(); stmt; if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_rsprof[(role) * 8 + (kind)] += clock64() - t__; } while (0)
#define RSPROF_TOTAL_BEGIN const long long rsprof_t0 = clock64();
#define RSPROF_TOTAL_END(role) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_rsprof[(role) * 8 + 7] += clock64() - rsprof_t0; } while (0)
#else
#define RSPROF_WAIT(role, kind, stmt) do { stmt; } while (0)
#define RSPROF_TOTAL_BEGIN
#define RSPROF_TOTAL_END(role) do { } while (0)
#endif

constexpr int THREADS = 832;                 // 26 warps
constexpr int N_PROD = 8, N_EPI = 16;        // producer warps (sub-partition x K half), epilogue warps (sub-partition x column quarter)
constexpr int W_PROD = 0, W_EPI = 8, W_MMA = 24, W_LOAD = 25;   // first warp of each role; warp % 4 = TMEM sub-partition for producers and epilogue
constexpr int CG = 8;                        // candidates per unit
constexpr int HCH = 64;                      // history rows per chunk
constexpr int NSTAGE = 3;                    // history-chunk stages (the loader runs two chunks ahead)
constexpr int HF_STRIDE = 68;                // floats per staged fp32 history row (272 B: conflict-free 16-byte reads across rows)
constexpr uint32_t W_TILE = 16384;           // [64 j][128 k'] bf16, K-major, un-swizzled: (j, 8 kb) at kb*1024 + (j/8)*128 + (j%8)*16
// weight image in global memory (att_prep_rs_kernel): per branch  W hi | W lo | w2[64] | b2 (+3 pad)
constexpr int IMG_BRANCH_BYTES = 2 * (int)W_TILE + 64 * 4 + 16;
__host__ __device__ constexpr int img_bytes() { return 2 * IMG_BRANCH_BYTES; }

// ---- inline PTX not in nrm_umma.cuh ------------------------------------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 16 slice; A: 128 lanes x 8 columns (two bf16 per 32-bit column, even k in the low half)
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// this thread's lane, 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {       // a -> low half
  const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&p);
}
// hi / lo split of two floats -> packed bf16 pairs: hi = bf16_rn(v), lo = bf16_rn(v - hi).  The two hi values come back as floats
// with one shift and one mask of the packed word (6 instructions per pair).
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

struct Geo {
  int B, H, C, G, nchunks;
  __device__ __forceinline__ void unit(int u, int& branch, int& b, int& c0, int& ncg) const {
    const int per_branch = B * G;
    branch = u >= per_branch ? 1 : 0;
    const int r = u - branch * per_branch;
    b = r / G;
    c0 = (r - b * G) * CG;
    ncg = min(CG, C - c0);
  }
  __device__ __forceinline__ void chunk(int ci, int ncg, int& h0, int& hl, int& rows, int& ntiles) const {
    h0 = ci * HCH;
    hl = min(HCH, H - h0);
    rows = ncg * hl;
    ntiles = (rows + 127) >> 7;
  }
};

__device__ __forceinline__ void arrive_warp(uint64_t* bar) {       // one arrival per warp, after all its lanes are done
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bar);
}

// ---- loader: one history chunk + the unit's candidate vectors -> stage ---------------------------------------------------------
// Both branches read fp32 rows [NH, 64] (label: the w1 projection xh; text/img: the PCA slice embed_rows_kernel wrote as fp32).
// The copies are 16-byte cp.async (no registers, any number in flight): the loader issues chunk n + 2 before it finishes chunk n.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(umma::smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// flattened (unit, chunk) sequence of a CTA
struct ChunkIter {
  int u, ci, branch, b, c0, ncg, h0, hl, rows, ntiles;
  __device__ __forceinline__ void set(const Geo& g, int u_, int ci_) {
    u = u_; ci = ci_;
    g.unit(u, branch, b, c0, ncg);
    g.chunk(ci, ncg, h0, hl, rows, ntiles);
  }
  __device__ __forceinline__ void next(const Geo& g) {
    if (ci + 1 < g.nchunks) set(g, u, ci + 1); else set(g, u + 1, 0);
  }
};

}  // namespace rs
}  // namespace nrm
