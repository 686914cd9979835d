// tic_tc.cuh — tcgen05 / TMEM / mbarrier plumbing of the tensor-core FDCT (sm_100a only).
//
// The 8x8 FDCT (utils.py:32-37), the division by the quantisation table (utils.py:48-53) and the zigzag
// permutation (constants.py:23-34) of the 128 blocks of a tile are ONE GEMM on the 5th-generation tensor cores:
//
//     D[128 blocks][64 zigzag positions] = A[128][64 pixels - 128, f16, exact] x (Bhi + Blo)[64][64]
//
// B[k = 8y + x][n] = c(u) c(v) cos((2y+1) u pi/16) cos((2x+1) v pi/16) / (qt[u][v] * hthr[n]) * 2^E, (u, v) = zigzag[n],
// split into two f16 terms (hi + lo carry ~22 bits), stacked along K: K = 128 = 8 MMAs of K = 16, the four lo
// chunks first.  The f32 accumulator lives in tensor memory; a block's 64 values are read back by the thread that
// owns the block (TMEM lane = block, column = zigzag position).  Measured against float64 (tools/tc_probe.cu,
// profiles/r2a_tc_probe.jsonl): max error 1.5e-4 coefficient units on adversarial blocks, 5e-5 on natural content.
// It is the FAST path only — the guard band + float64 exact path of tic_kernels.cuh decide every near-tie.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tic {
namespace tc {

constexpr int kABytes = 128 * 64 * 2;   // A operand of one tile: 128 blocks x 64 f16 (aliases TileShared::coef)
// N = 80 output columns: 64 scaled zigzag coefficients, then 8 column sums S_x = sum_y (p - 128)[y][x] and 8 signed
// column sums I_x = sum_y s_y (p - 128)[y][x], s = + for y in {0,3,4,7}, - for y in {1,2,5,6} — exact small integers.
// They are everything the reference's float64 column pass needs for u = 0 and u = 4 (every
// operation before the last multiply acts on integers), i.e. for the four coefficients whose quotients can sit on
// exact .5 ties: (0,0), (4,0), (0,4), (4,4).  A lane with such a tie reads them from tensor memory.
#ifndef TIC_RATIONAL
#define TIC_RATIONAL 0   // 1: the 16 column-sum outputs exist and ties at the four rational positions are settled from them
                         // (N = 80: at most 6 groups per CTA share the 512 TMEM columns; 0: N = 64, 8 groups — faster, tic_kernels.cuh)
#endif
constexpr int kN = TIC_RATIONAL ? 80 : 64;
constexpr int kBBytes = kN * 128 * 2;   // B operand: 80 columns x (64 hi + 64 lo) f16
constexpr int kColsPerGroup = kN;       // TMEM columns per tile accumulator (f32)
constexpr int kColSums = 64;            // first of the 16 sum columns

// Shared-memory layouts (canonical K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous 128 bytes)
//   A[m][k = 8y + x]: (m / 8) * 128 + y * 2048 + (m % 8) * 16 + x * 2     SBO = 128 (next 8 blocks), LBO = 2048 (next 8 k)
//   B[n][k]         : (n / 8) * 128 + (k / 8) * 1280 + (n % 8) * 16 + (k % 8) * 2     SBO = 128, LBO = 1280 (10 row groups)
constexpr uint32_t kLboA = 2048, kSboA = 128, kLboB = (kN / 8) * 128, kSboB = 128;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// cute::UMMA::SmemDescriptor: start address [0,14), leading byte offset [16,30), stride byte offset [32,46) — all
// >> 4 —, version 1 at [46,48), layout type 0 (no swizzle) at [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: D = f32 (1 at [4,6)), A = B = f16 (0), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdescF16M128 = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// One try: the thread sleeps in hardware until the phase completes or the time hint (ns) runs out.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(100000u)
        : "memory");
    return ok;
}
// Bounded (the MMA of a tile completes within microseconds): a fault in the asynchronous pipe must surface as a
// status, never as a hung GPU.  Returns false on timeout.
#ifndef TIC_POLL_SLEEP_NS
#define TIC_POLL_SLEEP_NS 0   // > 0: sleep between two polls of the MMA's mbarrier (fewer issue slots spent polling)
#endif
#ifndef TIC_POLL_FIRST_SLEEP_NS
#define TIC_POLL_FIRST_SLEEP_NS 0   // > 0: one sleep in front of the first poll (the MMAs of a tile take that long anyway)
#endif
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    if (TIC_POLL_FIRST_SLEEP_NS > 0) __nanosleep(TIC_POLL_FIRST_SLEEP_NS);
#pragma unroll 1
    for (int i = 0; i < (1 << 16); i++) {
        if (mbar_try_wait(bar, parity)) return true;
        if (TIC_POLL_SLEEP_NS > 0) __nanosleep(TIC_POLL_SLEEP_NS);
    }
    return false;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // one full warp; ncols: power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {   // the warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {   // arrives on `bar` when every MMA issued so far is complete
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id; the lane quarter is in taddr)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
// The registers of every tcgen05.ld issued so far are valid after this.  They are passed through the statement
// ("+r") so that the compiler cannot move a use of them above the wait.
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// The 8 MMAs of one tile, issued by ONE thread: lo halves first (small terms first), then hi.  desc_a0 / desc_b0:
// descriptors of the first K chunk; a later chunk only moves the start-address field (low word, 16-byte units).
__device__ __forceinline__ void issue_tile_mma(uint64_t desc_a0, uint64_t desc_b0, uint32_t tmem_d, uint32_t bar) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int half = 1 - (i >> 2);   // 1: lo (k = 64..127 of B), 0: hi
        const int j = i & 3;             // pixel rows 2j, 2j+1 of every block
        const uint64_t da = desc_a0 + (uint64_t)((uint32_t)j * 2u * kLboA >> 4);
        const uint64_t db = desc_b0 + (uint64_t)((uint32_t)(half * 4 + j) * 2u * kLboB >> 4);
        mma_f16(tmem_d, da, db, kIdescF16M128, i > 0 ? 1u : 0u);
    }
    commit(bar);
}

// One pixel row (8 bytes) of a block -> 8 x f16(p - 128), exact: 0x64pp is 1024 + p, minus 1152.
__device__ __forceinline__ uint4 row_to_f16(uint2 row) {
    uint4 v;
    v.x = __byte_perm(row.x, 0x64646464u, 0x4140); v.y = __byte_perm(row.x, 0x64646464u, 0x4342);
    v.z = __byte_perm(row.y, 0x64646464u, 0x4140); v.w = __byte_perm(row.y, 0x64646464u, 0x4342);
    const __half2 c = __floats2half2_rn(-1152.0f, -1152.0f);
    __half2 h;
    h = __hadd2(*reinterpret_cast<__half2*>(&v.x), c); v.x = *reinterpret_cast<uint32_t*>(&h);
    h = __hadd2(*reinterpret_cast<__half2*>(&v.y), c); v.y = *reinterpret_cast<uint32_t*>(&h);
    h = __hadd2(*reinterpret_cast<__half2*>(&v.z), c); v.z = *reinterpret_cast<uint32_t*>(&h);
    h = __hadd2(*reinterpret_cast<__half2*>(&v.w), c); v.w = *reinterpret_cast<uint32_t*>(&h);
    return v;
}

}  // namespace tc
}  // namespace tic
