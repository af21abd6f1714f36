// Thin inline-PTX layer over the sm_100a tensor-core path: tcgen05.mma (operands in shared
// memory, accumulator in tensor memory), tcgen05.ld, TMEM allocation, mbarrier completion.
//
// Operand tiles are written by the CTA's own threads (they are generated on the fly: a
// bf16 copy of a history tile, a per-candidate weight matrix), so they use the UN-swizzled
// K-major canonical layout, which is the simplest one to produce with 16-byte stores:
//
//   core matrix = 8 rows x 8 bf16 (16 B per row, 128 contiguous bytes)
//   byte offset of element (r, k) = (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2
//
// with SBO the distance between 8-row groups and LBO the distance between consecutive
// 8-element K blocks.  One tcgen05.mma (kind::f16) consumes K = 16, i.e. two K blocks.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace nrm {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (SWIZZLE_NONE, version 1 = Blackwell) ----------------
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // bits [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // bits [16,30) leading (K) byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // bits [32,46) stride (M/N) byte offset >> 4
  d |= (uint64_t)1 << 46;                                // bits [46,48) descriptor version
  return d;                                              // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}

// ---- instruction descriptor: kind::f16, bf16 x bf16 -> fp32, both operands K-major -------------
// a_mn / b_mn: the operand is read MN-major (bit 15 / 16), i.e. the tile in shared memory holds the
// TRANSPOSE of what a K-major descriptor would describe.  The canonical un-swizzled MN-major layout
//   byte offset of element (mn, k) = (mn/8)*SBO + (k/8)*LBO + (k%8)*16 + (mn%8)*2
// is the K-major layout above with the roles of the two indices exchanged, so ONE tile written as
// K-major X[r][c] (SBO_k between 8-row groups, LBO_k between 8-column blocks) doubles as the
// MN-major operand X^T with SBO = LBO_k and LBO = SBO_k -- no transposed copy is ever built.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4)                     // D format  : f32
         | (1u << 7)                   // A format  : bf16
         | (1u << 10)                  // B format  : bf16
         | ((a_mn ? 1u : 0u) << 15)    // A major   : 0 = K, 1 = MN
         | ((b_mn ? 1u : 0u) << 16)    // B major   : 0 = K, 1 = MN
         | ((uint32_t)(N >> 3) << 17)  // N / 8
         | ((uint32_t)(M >> 4) << 24); // M / 16
}

// ---- tensor memory -----------------------------------------------------------------------
// Called by one full warp.  `ncols` power of two >= 32.  The base address lands in *slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
// One thread: arm `bar` with the byte count and let the bulk-copy (TMA) engine move `bytes` (multiple of 16, both addresses
// 16-byte aligned) from global to shared memory; waiters use mbar_wait(bar, parity).
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
// try_wait is a hardware-assisted sleep: the waiting warp is suspended until the phase completes or the time hint (in ns) runs out,
// so a waiting role does not take issue slots from the working warps of its scheduler (with the default hint a waiter returns every
// ~100 ns: in the warp-specialised kernels half of all issued instructions were wait loops).  A protocol error must never hang the
// GPU: after ~4 s of waiting the thread traps, which surfaces as a launch failure on the host instead of a wedged device.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 1000000u) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
#ifdef NRM_WAIT_NANOSLEEP
    asm volatile("nanosleep.u32 %0;\n" ::"n"(NRM_WAIT_NANOSLEEP));      // experiment: fewer wake-ups of a waiting warp
#endif
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) asm volatile("trap;\n");     // 4 s: a protocol error must not wedge the GPU
  }
}

// ---- MMA ---------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T for one K = 16 slice.  One thread issues.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on `bar` once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- accumulator read-back: this thread's lane, 32 consecutive fp32 columns ------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 / 1 consecutive fp32 columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  return __uint_as_float(r);
}

// ---- operand tile helpers ------------------------------------------------------------------
// 8 fp32 -> 8 bf16 (round to nearest even) packed in one 16-byte store at `dst`.
__device__ __forceinline__ void store_bf16x8(void* dst, const float* v) {
  const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
  const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<const uint32_t*>(&p0); u.y = *reinterpret_cast<const uint32_t*>(&p1);
  u.z = *reinterpret_cast<const uint32_t*>(&p2); u.w = *reinterpret_cast<const uint32_t*>(&p3);
  *reinterpret_cast<uint4*>(dst) = u;
}

// hi / lo split of 8 fp32 values: hi = bf16(v), lo = bf16(v - hi).  a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi keeps
// ~16 mantissa bits per operand (relative error <= 2^-16), which is what torch calls float32 matmul precision
// "high" (3 x bf16); the products are accumulated in fp32 by the tensor core.
__device__ __forceinline__ void store_bf16x8_split(void* dst_hi, void* dst_lo, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // the two hi values come back as floats with one shift and one mask of the packed word (6 instructions per pair)
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * i] - __uint_as_float(hi[i] << 16), v[2 * i + 1] - __uint_as_float(hi[i] & 0xffff0000u));
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
  *reinterpret_cast<uint4*>(dst_hi) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst_lo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// NP = 1: one bf16 tile; NP = 2: hi tile followed by lo tile `part_bytes` later.
template <int NP>
__device__ __forceinline__ void store_operand8(unsigned char* tile, uint32_t off, uint32_t part_bytes, const float* v) {
  if (NP == 1) store_bf16x8(tile + off, v);
  else store_bf16x8_split(tile + off, tile + part_bytes + off, v);
}

// Canonical K-major tile of 64 rows x 64 k (bf16): 8 KB, SBO = 128, LBO = 1024.
constexpr uint32_t TILE64_SBO = 128, TILE64_LBO = 1024, TILE64_BYTES = 8192;
__device__ __forceinline__ uint32_t tile64_offset(int row, int kblock) {   // byte offset of (row, 8*kblock)
  return (uint32_t)kblock * TILE64_LBO + (uint32_t)(row >> 3) * TILE64_SBO + (uint32_t)(row & 7) * 16;
}

// Issue the four K = 16 MMAs of a 64 x 64 x 64 product: D[tmem_d] (+)= A_tile * B_tile^T.
__device__ __forceinline__ void mma_tile64(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t da = make_desc(a_smem + ks * 2 * TILE64_LBO, TILE64_LBO, TILE64_SBO);
    const uint64_t db = make_desc(b_smem + ks * 2 * TILE64_LBO, TILE64_LBO, TILE64_SBO);
    mma_bf16(tmem_d, da, db, idesc, (accumulate || ks > 0) ? 1u : 0u);
  }
}

// One elected lane of a converged warp (elect.sync): MMAs are issued inside `if (warp == w) { if (elect_one()) ... }`
// so that the tcgen05 instructions sit in warp-uniform control flow -- under a per-thread condition such as
// `threadIdx.x == 0` the compiler wraps every MMA in a lane-election loop (~120 cycles of issue per MMA, measured).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n" : "=r"(pred));
  return pred != 0;
}

// Generic product over a K range of KSTEPS * 16.  An operand = descriptor of its first K step plus the (>> 4)
// distances between K = 16 steps and between its hi and lo parts: every further descriptor is one 64-bit add.
// SPLIT = 1: A_hi B_hi.  SPLIT = 3: A_hi B_hi + A_hi B_lo + A_lo B_hi.
struct Operand { uint64_t desc; uint32_t kstep16, part16; };
__device__ __forceinline__ Operand make_operand(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t kstep, uint32_t part) {
  return Operand{make_desc(addr, lbo, sbo), kstep >> 4, part >> 4};
}
template <int SPLIT, int KSTEPS>
__device__ __forceinline__ void mma_product(uint32_t tmem_d, const Operand a, const Operand b, uint32_t idesc, bool accumulate) {
  constexpr int NT = SPLIT == 3 ? 3 : 1;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const uint32_t ao = (t == 2) ? a.part16 : 0u, bo = (t == 1) ? b.part16 : 0u;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
      mma_bf16(tmem_d, a.desc + (uint64_t)(ao + ks * a.kstep16), b.desc + (uint64_t)(bo + ks * b.kstep16), idesc,
               (accumulate || t > 0 || ks > 0) ? 1u : 0u);
  }
}
// 64 x 64 bf16 tile written K-major (tile64_offset): as a K-major operand, and as the MN-major operand of its transpose
__device__ __forceinline__ Operand op_tile64_k(uint32_t addr) { return make_operand(addr, TILE64_LBO, TILE64_SBO, 2 * TILE64_LBO, TILE64_BYTES); }
__device__ __forceinline__ Operand op_tile64_mn(uint32_t addr) { return make_operand(addr, TILE64_SBO, TILE64_LBO, 2 * TILE64_SBO, TILE64_BYTES); }

}  // namespace umma
}  // namespace nrm
