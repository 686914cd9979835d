// tic_kernels.cuh — sm_100a device code of the tinyimgcodec encode path.
//
// A TILE is kTile consecutive 8x8 blocks of one image in the stream's block-raster order
// (tinyimgcodec/codec.py:34-36), one thread per block.  The encode kernel turns a tile into its
// piece of the bit stream without talking to any other tile:
//
//   load      8 x LDG.64 per thread; a warp reads 256 contiguous bytes per pixel row
//   transform level shift + 8x8 FDCT in registers, FP32 (AAN butterflies), scaled into the
//             quantiser: the fast path for utils.py:32-37,48-53
//   quantise  groups of 8 zigzag coefficients; a group the whole warp quantises to zero is skipped
//             after one multiply + max per coefficient; otherwise round-to-nearest by magic number,
//             and a coefficient whose fraction is within a guard band of a .5 tie goes to the EXACT path
//   exact     the reference's float64 FDCT (SciPy/ducc0 op order), 8 lanes per
//             flagged coefficient, per warp: reproduces the reference's rounding of ties.  The DC of
//             the block in front of each warp's 32 blocks is one more exact item (codec.py:34-35).
//   symbols   one walk over the non-zero mask per block: zigzag (constants.py:23-34), DC difference,
//             run lengths (huffman.py:12-33), Huffman code + value bits (huffman.py:41-63) packed
//             MSB-first (bitbuffer.py:17-40) into a few private 32-bit words
//   scan      bit lengths -> warp shuffle scan -> CTA scan: tile-relative bit offset of every block
//   place     private words funnel-shifted into the tile's staging window in shared memory
//   copy      staging window -> the tile's slot in the ARENA (bump-allocated, 16-byte granules)
//
// Three small kernels then scan the tile bit counts (scan_*), and compact_kernel moves every tile
// from the arena to its final bit position in the dense output (funnel shift + byte swap).
// Quantised coefficients never touch HBM.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "tic_tables.h"
#include "tic_tc.cuh"

namespace tic {

// 1 (default): the FDCT + quantiser scaling run on the tensor cores (tcgen05, tic_tc.cuh);
// 0: the round-1 FP32 CUDA-core transform (kept for A/B measurements: tools/build_variants.sh)
#ifndef TIC_FDCT_TC
#define TIC_FDCT_TC 1
#endif
// 128-thread GROUPS per CTA of the persistent encode kernel: one CTA per SM, every group works on its own tile
// with its own named barrier, TMEM columns and mbarrier; the B operand and the TMEM allocation are shared.
#ifndef TIC_GROUPS
#define TIC_GROUPS 8
#endif
#ifndef TIC_GROUPS_AUTO
#define TIC_GROUPS_AUTO 7   // per-image tables: every group keeps its own image's tables (TabShared) in shared memory
#endif
// 8 groups (32 warps per SM: everything an SM holds) need 64 TMEM columns per group (the 16 column-sum outputs of
// TIC_RATIONAL go: 7 x 80 > 512), 64 registers per thread and a group state of at most 26 KB: the fixed Huffman tables
// live once per CTA (TabShared), TIC_PRIV 8, TIC_WIN 832.  Measured: 4.10 ms against 4.29 ms for 7 groups at 72
// registers and 4.50 ms for 6 groups with the column sums (6 groups without them: 4.60 ms).
// (Prefetching a group's next tile was measured and dropped: held in registers it spills — with 227 KB of shared
// memory there is no L1 to catch a spill — and through cp.async + LDS it costs more than the latency it hides:
// 5.52 ms against 5.01 ms for plain loads at the top of the tile, profiles/r2_variants.md.)

#ifndef TIC_TILE
#define TIC_TILE 128
#endif
#ifndef TIC_CTAS
#define TIC_CTAS 7   // CTAs per SM of the single-group kernels (the C-variant encoder: 72 registers, 28 warps per SM; 6: 1.49 vs 1.45 ms per 1024 images)
#endif
#ifndef TIC_PRIV
#define TIC_PRIV 8
#endif
// 1: only warp 0 of a group polls the MMA's mbarrier, the other warps sleep at the group barrier
#ifndef TIC_POLL_WARP0
#define TIC_POLL_WARP0 0
#endif
#ifndef TIC_WIN
#define TIC_WIN 832
#endif
#ifndef TIC_QUANT_F32X2
#define TIC_QUANT_F32X2 1  // the tensor-core quantiser rounds coefficient pairs with packed FP32 instructions
#endif
#ifndef TIC_PREFETCH_L2
#define TIC_PREFETCH_L2 0  // 1 / 2: the pixel rows of a group's next tile are pulled into L2 one tile ahead (bulk prefetch per row /
                           // one PREFETCH per line).  Measured: long-scoreboard stalls 8.5 -> 6.5 %, kernel time unchanged or worse
                           // (4.62 vs 4.50 ms next to the FP32 predictor, 4.56 vs 4.58 ms without): off.
#endif
#ifndef TIC_SLOW_WALK_CALL
#define TIC_SLOW_WALK_CALL 1 // 1: the second walk of long blocks is a call, not inline code (4.27 vs 4.30 ms: instruction cache)
#endif
#ifndef TIC_STATS_GROUPS
#define TIC_STATS_GROUPS 1 // per-image tables: symbol statistics by the persistent multi-group kernel (0: single-group CTAs)
#endif
#ifndef TIC_EXACT_INT_COLS
#define TIC_EXACT_INT_COLS 1 // exact path: the column pass for u = 0 / 4 from one integer sum per column (exact by the same argument as settle_rational)
#endif
#ifndef TIC_EXACT_ROW_INLINE
#define TIC_EXACT_ROW_INLINE 1   // 1: the exact path's row pass for v = 0 / 4 inline, no call (3.84 against 3.87 ms)
#endif
#ifndef TIC_STATS_DIRECT
#define TIC_STATS_DIRECT 1 // per-image tables: AC symbol counts by one shared-memory atomic per lane (1) or combined per warp and step with match.any (0)
#endif
#ifndef TIC_HALO_F32
#define TIC_HALO_F32 1     // tensor-core path: the DC predictor in front of a warp from its pixel sum in FP32 (ties: exact path)
#endif
#ifndef TIC_RAT_FAST
#define TIC_RAT_FAST 1     // ties at the four rational positions: the float64 sequence with its power-of-two factors moved to the end
#endif
#ifndef TIC_WALK_PIPE
#define TIC_WALK_PIPE 1   // the walk fetches the next coefficient before it codes the current one
#endif
constexpr int kTile = TIC_TILE;                             // blocks (= threads) per tile
constexpr int kWarps = kTile / 32;
constexpr int kCtasPerSm = TIC_CTAS;                        // CTAs per SM of the single-group kernels
constexpr int kGroups = TIC_GROUPS;             // groups per CTA: fixed tables (one table copy per CTA), statistics
constexpr int kGroupsAuto = TIC_GROUPS_AUTO;    // groups per CTA: per-image tables
constexpr bool kFdctTc = TIC_FDCT_TC != 0;
static_assert(!kFdctTc || kTile == 128, "the tensor-core transform is M = 128: one tile = 128 blocks");
static_assert((kTile & (kTile - 1)) == 0, "thread-in-group = threadIdx.x & (kTile - 1)");
constexpr int kPrivWords = TIC_PRIV;                           // private words per block on the fast path (256 bits)
// The bits of a tile are assembled in a WINDOW of kWinWords 32-bit words of shared memory: a tile
// whose stream is longer (worst case 128 x 1662 bits + a table header) is emitted in several rounds.
constexpr int kWinWords = TIC_WIN;
static_assert(kTile % 32 == 0 && kTile >= 64 && kTile <= 512, "a tile is 2..16 warps of blocks");
static_assert(kWinWords % 4 == 0 && kWinWords >= 256, "window: 16-byte copies");
static_assert(kPrivWords >= 2 && kPrivWords <= 52, "a block has at most 1662 bits");
constexpr int kWarpWork = 32;                               // exact-path worklist entries per warp and round

// Quality-dependent constants, passed BY VALUE so that every entry is a constant-bank operand.
struct QuantParams {
    // the three fast-path arrays are indexed by ZIGZAG position k (coefficient u*8+v = zigzag[k]), so that the
    // constants of a coefficient pair / group are adjacent and arrive with one wide uniform load
    float qmul[64];   // 1 / (8 * aan[u] * aan[v] * qt[u][v]):   t = d * qmul is coefficient / qt
    float zmul[64];   // qmul / hthr: |d * zmul| < 1  =>  rounds to 0 and is nowhere near a tie
    float hthr[64];   // 0.5 - w: |t - round(t)| > hthr  =>  within the guard band of a .5 tie
    // tensor-core path: the accumulator holds D = t / hthr * 2^tc_exp (tic_tc.cuh), zigzag order
    float cn[64];     // t = D * cn
    float tc_live;    // 2^tc_exp: |D| < tc_live  =>  |t| < hthr: rounds to 0 and is nowhere near a tie
    int tc_exp;
    double qt[64];    // [u*8+v]  the reference's float64 divisor (utils.py:50-53)
    double dcinv;     // 1 / (8 * qt[0]): quantised DC = (sum of pixels - 8192) * dcinv
    float dcinv_f;    // the same in FP32 (tensor-core path: the predictor in front of a warp, transform_tile_tc)
};

struct ImageDesc {
    const uint8_t* px;
    int h, w;          // unpadded
    int bw;            // blocks per row = ceil(w/8)
    int nblk;          // ceil(h/8)*ceil(w/8)
    long long tile0;   // index of this image's first tile in the batch
    int bw_shift;      // log2(bw) when bw is a power of two (block row = block >> bw_shift), else -1
    int pad;
};

// What the encode kernel leaves behind per tile.
struct TileRec {
    uint32_t bits;     // [29:0] bits of the tile (header included for a first tile), [30] first, [31] closing
    uint32_t off16;    // arena offset of the tile's words, in 16-byte units
    int32_t img;
    uint32_t pad;
};
constexpr uint32_t kRecFirst = 1u << 30, kRecClosing = 1u << 31, kRecBitsMask = (1u << 30) - 1u;

// counters[] layout in the workspace
enum { kCtrArena = 0, kCtrOverflow = 1, kCtrExactItems = 2, kCtrExactChanged = 3, kCtrTotalBits = 4,
       kCtrAnyStatus = 5, kCtrUnflagged = 6 /* debug: exact != fast outside the guard band */,
       kCtrTcTimeout = 7 /* an MMA completion never arrived */,
       kCtrTicketA = 8, kCtrTicketB = 9 /* last-CTA-done tickets of the two scan kernels */, kCtrCount = 10 };

__device__ __constant__ uint8_t c_zigzag[64] = {TIC_ZIGZAG_LIST};
__device__ __constant__ HuffTables c_default_tables;
// C-variant stream (c/img.c of the reference): scaledQuant[qfactor][i] = 65536 / (QUANT[i] << qfactor)
// (c/img.c:157-181), raster order
__device__ __constant__ uint16_t c_cvar_scaled_quant[4][64];

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
// thread within its group (= within the CTA for the single-group kernels)
__device__ __forceinline__ int tid() { return (int)threadIdx.x & (kTile - 1); }
// The same, re-read from the special register every time: costs an S2R where a cached copy would cost a register
// that lives across the whole tile loop (and, with none to spare, a reload from local memory = an L2 round trip).
__device__ __forceinline__ int tid_now() {
    unsigned x;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(x));
    return (int)x & (kTile - 1);
}
template <int G>
__device__ __forceinline__ void group_sync(int g) {   // barrier over the kTile threads of group g
    if constexpr (G == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kTile) : "memory");
}

__device__ __forceinline__ int reflect_idx(int i, int n) {   // numpy "reflect", utils.py:56-61
    if (i < n) return i;
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    // i >= 0 and i < n + 7 (padding to the next multiple of 8): i % p as a loop that runs zero times for n >= 8 — the
    // division routine behind `%`, inlined 32 times, was 700 of the kernel's 4800 instructions
    int m = i;
    while (m >= p) m -= p;
    return m < n ? m : p - m;
}

__device__ __forceinline__ int bitlen(int v) {               // bits_required, utils.py:9-10
    return 32 - __clz(v < 0 ? -v : v);
}

__device__ __forceinline__ long long round_up128(long long v) { return (v + 127) & ~127ll; }

// ---------------------------------------------------------------------------------------------
// fast path: FP32 AAN 8-point DCT (5 multiplies, 29 adds; the output scale is folded into
// QuantParams::qmul), two transforms at a time on Blackwell's packed FP32 pipe (FADD2 / FMUL2 /
// FFMA2: one issue slot for two lanes of arithmetic).  Not bit-exact with the reference — the guard
// band + exact path is.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;   // two floats in an aligned register pair
__device__ __forceinline__ f32x2 pk2(float x, float y) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& x, float& y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void aan8x2(f32x2& d0, f32x2& d1, f32x2& d2, f32x2& d3, f32x2& d4, f32x2& d5,
                                       f32x2& d6, f32x2& d7) {
    const f32x2 c707 = pk2(0.707106781f, 0.707106781f), c382 = pk2(0.382683433f, 0.382683433f);
    const f32x2 c541 = pk2(0.541196100f, 0.541196100f), c1306 = pk2(1.306562965f, 1.306562965f);
    f32x2 t0 = add2(d0, d7), t7 = sub2(d0, d7), t1 = add2(d1, d6), t6 = sub2(d1, d6);
    f32x2 t2 = add2(d2, d5), t5 = sub2(d2, d5), t3 = add2(d3, d4), t4 = sub2(d3, d4);
    f32x2 t10 = add2(t0, t3), t13 = sub2(t0, t3), t11 = add2(t1, t2), t12 = sub2(t1, t2);
    d0 = add2(t10, t11);
    d4 = sub2(t10, t11);
    f32x2 z1 = mul2(add2(t12, t13), c707);
    d2 = add2(t13, z1);
    d6 = sub2(t13, z1);
    t10 = add2(t4, t5);
    t11 = add2(t5, t6);
    t12 = add2(t6, t7);
    f32x2 z5 = mul2(sub2(t10, t12), c382);
    f32x2 z2 = fma2(c541, t10, z5);
    f32x2 z4 = fma2(c1306, t12, z5);
    f32x2 z3 = mul2(t11, c707);
    f32x2 z11 = add2(t7, z3), z13 = sub2(t7, z3);
    d5 = add2(z13, z2);
    d3 = sub2(z13, z2);
    d1 = add2(z11, z4);
    d7 = sub2(z11, z4);
}

// Column pass for a pair of adjacent columns with a SCALAR final stage: the 16 results are written as single
// floats, so that they can be re-paired by rows for the second pass without a single register move.
__device__ __forceinline__ void aan8x2_split(f32x2 d0, f32x2 d1, f32x2 d2, f32x2 d3, f32x2 d4, f32x2 d5, f32x2 d6,
                                             f32x2 d7, float (&lo)[8], float (&hi)[8]) {
    const f32x2 c707 = pk2(0.707106781f, 0.707106781f), c382 = pk2(0.382683433f, 0.382683433f);
    const f32x2 c541 = pk2(0.541196100f, 0.541196100f), c1306 = pk2(1.306562965f, 1.306562965f);
    f32x2 t0 = add2(d0, d7), t7 = sub2(d0, d7), t1 = add2(d1, d6), t6 = sub2(d1, d6);
    f32x2 t2 = add2(d2, d5), t5 = sub2(d2, d5), t3 = add2(d3, d4), t4 = sub2(d3, d4);
    f32x2 t10 = add2(t0, t3), t13 = sub2(t0, t3), t11 = add2(t1, t2), t12 = sub2(t1, t2);
    f32x2 z1 = mul2(add2(t12, t13), c707);
    f32x2 u10 = add2(t4, t5), u11 = add2(t5, t6), u12 = add2(t6, t7);
    f32x2 z5 = mul2(sub2(u10, u12), c382);
    f32x2 z2 = fma2(c541, u10, z5);
    f32x2 z4 = fma2(c1306, u12, z5);
    f32x2 z3 = mul2(u11, c707);
    f32x2 z11 = add2(t7, z3), z13 = sub2(t7, z3);
    float a0, a1, b0, b1;
    upk2(t10, a0, a1); upk2(t11, b0, b1); lo[0] = a0 + b0; hi[0] = a1 + b1; lo[4] = a0 - b0; hi[4] = a1 - b1;
    upk2(t13, a0, a1); upk2(z1, b0, b1);  lo[2] = a0 + b0; hi[2] = a1 + b1; lo[6] = a0 - b0; hi[6] = a1 - b1;
    upk2(z13, a0, a1); upk2(z2, b0, b1);  lo[5] = a0 + b0; hi[5] = a1 + b1; lo[3] = a0 - b0; hi[3] = a1 - b1;
    upk2(z11, a0, a1); upk2(z4, b0, b1);  lo[1] = a0 + b0; hi[1] = a1 + b1; lo[7] = a0 - b0; hi[7] = a1 - b1;
}

// 2-D transform of d[y*8+x] in place: columns two at a time, then rows two at a time.
__device__ __forceinline__ void fdct8x8(float (&d)[64]) {
    float z[8][8];   // after the column pass
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float lo[8], hi[8];
        aan8x2_split(pk2(d[0 * 8 + 2 * j], d[0 * 8 + 2 * j + 1]), pk2(d[1 * 8 + 2 * j], d[1 * 8 + 2 * j + 1]),
                     pk2(d[2 * 8 + 2 * j], d[2 * 8 + 2 * j + 1]), pk2(d[3 * 8 + 2 * j], d[3 * 8 + 2 * j + 1]),
                     pk2(d[4 * 8 + 2 * j], d[4 * 8 + 2 * j + 1]), pk2(d[5 * 8 + 2 * j], d[5 * 8 + 2 * j + 1]),
                     pk2(d[6 * 8 + 2 * j], d[6 * 8 + 2 * j + 1]), pk2(d[7 * 8 + 2 * j], d[7 * 8 + 2 * j + 1]), lo, hi);
#pragma unroll
        for (int u = 0; u < 8; u++) { z[u][2 * j] = lo[u]; z[u][2 * j + 1] = hi[u]; }
    }
#pragma unroll
    for (int u = 0; u < 8; u += 2) {   // rows u and u+1
        f32x2 q[8];
#pragma unroll
        for (int c = 0; c < 8; c++) q[c] = pk2(z[u][c], z[u + 1][c]);
        aan8x2(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);
#pragma unroll
        for (int c = 0; c < 8; c++) upk2(q[c], d[u * 8 + c], d[(u + 1) * 8 + c]);
    }
}

// ---------------------------------------------------------------------------------------------
// exact path: scipy.fftpack.dct(x, norm="ortho"), N=8, float64, as executed by ducc0.
// One IEEE operation per intrinsic, never contracted.  Returns y[sel].
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ double dct8_exact(const double x0, const double x1, const double x2, const double x3,
                                          const double x4, const double x5, const double x6, const double x7,
                                          int sel) {
    const double TW0 = 0x1.f6297cff75cb0p-1, TW1 = 0x1.d906bcf328d46p-1, TW2 = 0x1.a9b66290ea1a3p-1,
                 TW3 = 0x1.6a09e667f3bccp-1, TW4 = 0x1.1c73b39ae68c8p-1, TW5 = 0x1.87de2a6aea963p-2,
                 TW6 = 0x1.8f8b83c69a60ap-3;
    const double WR = 0x1.6a09e667f3bccp-1, WI = 0x1.6a09e667f3bcdp-1, HSQ = 0x1.6a09e667f3bcdp-1;
    // (every value keeps its own operation sequence; what outputs 0 and 4 do not need is computed behind their return)
    double c0 = __dmul_rn(2.0, x0), c7 = __dmul_rn(2.0, x7);
    double c1 = __dadd_rn(x1, x2);
    double c3 = __dadd_rn(x3, x4);
    double c5 = __dadd_rn(x5, x6);
    double h0 = __dadd_rn(c0, c7);
    double h3 = __dmul_rn(2.0, c3);
    double h1 = __dadd_rn(c1, c5);
    double s;  // 0.25 * r_k for the FFT output(s) this selection needs
    if (sel == 0 || sel == 4) {
        double a = __dadd_rn(h0, h3);
        double e1 = __dmul_rn(2.0, h1);
        if (sel == 0) {
            s = __dmul_rn(0.25, __dadd_rn(a, e1));
            return __dmul_rn(s, HSQ);
        }
        s = __dmul_rn(0.25, __dsub_rn(a, e1));
        return __dmul_rn(s, TW3);
    }
    double c2 = __dsub_rn(x2, x1), c4 = __dsub_rn(x4, x3), c6 = __dsub_rn(x6, x5);
    double h4 = __dsub_rn(c0, c7), h7 = __dmul_rn(-2.0, c4);
    double tr = __dsub_rn(c1, c5);
    double ti = __dadd_rn(c2, c6), h2 = __dsub_rn(c2, c6);
    double h6 = __dadd_rn(__dmul_rn(WR, ti), __dmul_rn(WI, tr));
    double h5 = __dsub_rn(__dmul_rn(WR, tr), __dmul_rn(WI, ti));
    double b = __dsub_rn(h0, h3);
    double e2 = __dmul_rn(2.0, h2);
    double a2 = __dadd_rn(h4, h7), b2 = __dsub_rn(h4, h7);
    double e5 = __dmul_rn(2.0, h5), e6 = __dmul_rn(2.0, h6);
    double sk, skc, twk, twkc;  // pair (k, kc=8-k): t1 = tw[k-1]*s_kc + tw[kc-1]*s_k
    int k = sel < 4 ? sel : 8 - sel;
    if (k == 1) {
        sk = __dmul_rn(0.25, __dadd_rn(a2, e5));   // r1
        skc = __dmul_rn(0.25, __dadd_rn(b2, e6));  // r7
        twk = TW0; twkc = TW6;
    } else if (k == 2) {
        sk = __dmul_rn(0.25, __dsub_rn(b, e2));    // r2
        skc = __dmul_rn(0.25, __dadd_rn(b, e2));   // r6
        twk = TW1; twkc = TW5;
    } else {
        sk = __dmul_rn(0.25, __dsub_rn(b2, e6));   // r3
        skc = __dmul_rn(0.25, __dsub_rn(a2, e5));  // r5
        twk = TW2; twkc = TW4;
    }
    double t1 = __dadd_rn(__dmul_rn(twk, skc), __dmul_rn(twkc, sk));
    double t2 = __dsub_rn(__dmul_rn(twk, sk), __dmul_rn(twkc, skc));
    return sel < 4 ? __dmul_rn(0.5, __dadd_rn(t1, t2)) : __dmul_rn(0.5, __dsub_rn(t1, t2));
}

// The reference's quantised DC from the 8 column sums of the LEVEL-SHIFTED block alone.  In its float64 column
// pass every operation before the last multiply acts on small integers and is exact, so column x contributes
// RN(0.5 * colsum_x * HSQ); the row pass over those eight doubles is dct8_exact(.., 0).  Used where the fast
// path finds the DC within the guard band of a tie (every block of a flat area whose DC lands on one).
__device__ __noinline__ int dc_exact_from_colsums(double c0, double c1, double c2, double c3, double c4, double c5,
                                                  double c6, double c7, double qt0) {
    const double HSQ = 0x1.6a09e667f3bcdp-1;
    const double y = dct8_exact(__dmul_rn(__dmul_rn(0.5, c0), HSQ), __dmul_rn(__dmul_rn(0.5, c1), HSQ),
                                __dmul_rn(__dmul_rn(0.5, c2), HSQ), __dmul_rn(__dmul_rn(0.5, c3), HSQ),
                                __dmul_rn(__dmul_rn(0.5, c4), HSQ), __dmul_rn(__dmul_rn(0.5, c5), HSQ),
                                __dmul_rn(__dmul_rn(0.5, c6), HSQ), __dmul_rn(__dmul_rn(0.5, c7), HSQ), 0);
    return __double2int_rn(__ddiv_rn(y, qt0));   // np.round(coeffs / qt), utils.py:53
}

struct TileInfo {
    const uint8_t* px;
    int h, w, bw, bw_shift;
    int img;          // image index
    int blk0;         // first block of the tile within the image
    int nb;           // blocks in this tile (0..kTile)
    bool first;       // first tile of its image (carries the header)
    bool closing;     // last tile of its image (the stream is padded to a byte after it)
};

// ---------------------------------------------------------------------------------------------
// shared memory of one CTA
// ---------------------------------------------------------------------------------------------
// Huffman tables in the form the walk wants, per (run << 4 | size).  Fixed tables: ONE 32-bit word per entry,
// (len + size) << 27 | code << size (0: not in the table), in the first half of the array — a random-index 64-bit
// lookup cost 6.6 shared-memory wavefronts, a third of the kernel's LSU traffic (profiles/r2d) — and one copy per CTA
// (they are the same for every group: 2 KB per group back, which is what lets 8 groups fit 227 KB).  Per-image
// tables: {code, kHuffPresent | len}, .y == 0: absent; one copy per group, behind its TileShared.
struct TabShared {
    uint2 ac_tab[256];
    uint2 dc_tab[16];
};

struct TileShared {
    uint32_t coef[32][kTile];        // zigzag pairs (2i, 2i+1) packed lo/hi int16, one column per thread
    uint32_t priv[kPrivWords + 1][kTile];   // the block's bits, MSB-first from block bit 0, one column per
                                            // thread; the last row only catches the overflow of long blocks
    uint32_t nz_lo[kTile];           // bit 31-k set: zigzag coefficient k != 0 (k = 1..31; k = 0 unused)
    uint32_t nz_hi[kTile];           // bit 63-k set, k = 32..63
    int dcq[kTile];                  // quantised DC of thread t's block
    int dc_halo[kWarps];             // quantised DC of the block in front of the warp's first block
    int blocksum[kWarps];            // tensor-core path: pixel sum of the warp's last block (the next warp's predictor)
    uint32_t work[kWarps][kWarpWork];// exact-path worklist: lane << 6 | zigzag index; bit 31: halo DC
    int work_count[kWarps];
    int pending[kWarps];             // flagged coefficients that did not fit the worklist this round
    int warp_bits[kWarps];
    int warp_err[kWarps];
    unsigned int arena_off;          // 16-byte units; 0xffffffff: arena exhausted
    // The description of the tile in work and of the group's next tile: written ONE TILE AHEAD by warp 0 while the
    // tensor core runs, read from shared memory where needed — a tile's geometry costs no registers.
    TileInfo tinfo[2];
    int u_img, u_lt;                 // uniform batches: (image, tile in image) of the tile after tinfo's newest (lane 0 of warp 0)
    unsigned int stat_items, stat_changed, stat_unflagged, tc_timeout;   // flushed to the batch counters at the end
    alignas(16) uint32_t stage[kWinWords];   // window of the tile-relative MSB-first bit buffer (kept zeroed)
    // LAST member: the kernels that keep ONE table copy per CTA (fixed tables) or none (statistics) pack their groups
    // at kGroupStrideNoTab and never touch this member — it overlaps the next group there.
    TabShared tab;
};

// Two tenants of the private-word area while no walk is using it:
//   rows 0-1, the WARP'S OWN 32 columns   exact path: the column pass results of the warp's 4 entries in flight
//                                         (4 x 8 doubles = 2 x 128 bytes).  A warp's exact path runs in front of its own
//                                         walk, and no other warp touches its columns — the warps of a group are not
//                                         synchronised between the two, so a group-wide array here would be overwritten
//                                         by a faster warp's private words.
//   bytes [1024, 4352)                    per-image tables: 272 symbol counters (room for 288) + 272 first-occurrence
//                                         keys (the statistics kernels never walk)
__device__ __forceinline__ double* exact_colres(TileShared& sm, int warp, int grp) {
    return reinterpret_cast<double*>(&sm.priv[grp >> 1][32 * warp + (grp & 1) * 16]);
}
__device__ __forceinline__ uint32_t* stats_hist(TileShared& sm) { return &sm.priv[0][0] + 256; }
__device__ __forceinline__ unsigned long long* stats_first(TileShared& sm) {
    return reinterpret_cast<unsigned long long*>(&sm.priv[0][0] + 256 + 288);
}
static_assert((kPrivWords + 1) * kTile * 4 >= 1024 + (288 + 2 * 272) * 4 && kTile * 4 * 2 == 1024,
              "the private-word area also holds the exact path's column results and the statistics bins");

// Tile number lt of image `img`.
__device__ __forceinline__ TileInfo tile_info(const ImageDesc& d, int img, long long lt) {
    TileInfo ti;
    ti.px = d.px; ti.h = d.h; ti.w = d.w; ti.bw = d.bw; ti.bw_shift = d.bw_shift; ti.img = img;
    ti.blk0 = (int)lt * kTile;
    const int rem = d.nblk - ti.blk0;
    ti.nb = rem < 0 ? 0 : (rem > kTile ? kTile : rem);
    ti.first = (lt == 0);
    ti.closing = (ti.blk0 + kTile >= d.nblk);
    return ti;
}

// Warp-cooperative (all 32 lanes of one warp call): the image a tile belongs to and its place in it.
// uniform_tpi > 0: every image has that many tiles (the common batch), no search; otherwise a 32-ary
// search for the last image whose first tile is <= tile.
__device__ __forceinline__ TileInfo locate_tile(const ImageDesc* __restrict__ descs, int n_images,
                                                long long tile, int uniform_tpi) {
    const int lane = threadIdx.x & 31;
    int lo = 0;
    if (uniform_tpi > 0) {
        lo = (int)(tile / uniform_tpi);
    } else {
        int hi = n_images;   // answer in [lo, hi)
        while (hi - lo > 1) {
            const int span = hi - lo;
            const int step = (span + 31) >> 5;
            const int probe = lo + (lane + 1) * step;   // candidates lo+step, lo+2*step, ...
            const bool le = probe < hi && __ldg(&descs[probe].tile0) <= tile;
            const int cnt = __popc(__ballot_sync(0xffffffffu, le));   // monotone: first cnt probes are <= tile
            const int new_lo = lo + cnt * step;
            hi = new_lo + step < hi ? new_lo + step : hi;
            lo = new_lo;
        }
    }
    return tile_info(descs[lo], lo, tile - descs[lo].tile0);
}

__device__ __forceinline__ double load_px_exact(const TileInfo& ti, int y, int x) {
    int yy = reflect_idx(y, ti.h), xx = reflect_idx(x, ti.w);
    return (double)((int)__ldg(ti.px + (size_t)yy * ti.w + xx) - 128);   // codec.py:29
}

// ---------------------------------------------------------------------------------------------
// phase 1: pixels -> quantised zigzag coefficients in shared memory (fast path) + flags
// ---------------------------------------------------------------------------------------------
template <int K>
struct ZZ {  // compile-time zigzag -> raster
    static constexpr int tab[64] = {TIC_ZIGZAG_LIST};
    static constexpr int r = tab[K];
};

// Group G = zigzag coefficients 8G .. 8G+7 of every block of the warp: max |d * zmul| over the group.
__device__ __forceinline__ float fmax3(float a, float b, float c) {   // FMNMX3
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
template <int G>
__device__ __forceinline__ float group_peak(const float (&d)[64], const QuantParams& qp) {
#define TIC_Z(J) fabsf(d[ZZ<8 * G + J>::r] * qp.zmul[8 * G + J])
    float m = fmax3(TIC_Z(0), TIC_Z(1), TIC_Z(2));
    m = fmax3(m, TIC_Z(3), TIC_Z(4));
    m = fmax3(m, TIC_Z(5), TIC_Z(6));
    return fmaxf(m, TIC_Z(7));
#undef TIC_Z
}

template <int G, int P>
__device__ __forceinline__ void quantise_pairs(const float (&d)[64], const QuantParams& qp, TileShared& sm,
                                               int t, uint32_t& nz, uint32_t& fl, int& dc) {
    if constexpr (P < 4) {
        constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23: (x + kMagic) - kMagic = round-to-nearest-even(x)
        constexpr int k0 = 8 * G + 2 * P, k1 = k0 + 1;
        constexpr int r0 = ZZ<k0>::r, r1 = ZZ<k1>::r;
        const float b0 = fmaf(d[r0], qp.qmul[k0], kMagic), b1 = fmaf(d[r1], qp.qmul[k1], kMagic);
        const float q0 = b0 - kMagic, q1 = b1 - kMagic;
        const float e0 = fmaf(d[r0], qp.qmul[k0], -q0), e1 = fmaf(d[r1], qp.qmul[k1], -q1);   // t - round(t)
        if (k0 != 0 && q0 != 0.0f) nz |= 0x80000000u >> (k0 & 31);
        if (q1 != 0.0f) nz |= 0x80000000u >> (k1 & 31);
        if (fabsf(e0) > qp.hthr[k0]) fl |= 0x80000000u >> (k0 & 31);   // the exact path decides
        if (fabsf(e1) > qp.hthr[k1]) fl |= 0x80000000u >> (k1 & 31);
        if constexpr (k0 == 0) dc = __float_as_int(b0) - 0x4B400000;
        // the low 16 bits of the magic sums are the two's-complement quantised values
        sm.coef[k0 >> 1][t] = __byte_perm(__float_as_uint(b0), __float_as_uint(b1), 0x5410);
        quantise_pairs<G, P + 1>(d, qp, sm, t, nz, fl, dc);
    }
}

template <int G>
__device__ __forceinline__ void quantise_groups(const float (&d)[64], const QuantParams& qp, TileShared& sm,
                                                int t, uint32_t& nz_lo, uint32_t& nz_hi, uint32_t& fl_lo,
                                                uint32_t& fl_hi, int& dc) {
    if constexpr (G < 8) {
        bool live = true;
        if constexpr (G > 0) live = __any_sync(0xffffffffu, group_peak<G>(d, qp) >= 1.0f);
        if (live) {   // warp-uniform
            if constexpr (G < 4) quantise_pairs<G, 0>(d, qp, sm, t, nz_lo, fl_lo, dc);
            else quantise_pairs<G, 0>(d, qp, sm, t, nz_hi, fl_hi, dc);
        }
        quantise_groups<G + 1>(d, qp, sm, t, nz_lo, nz_hi, fl_lo, fl_hi, dc);
    }
}

// Coefficients sit in shared memory as two's-complement 16-bit values.
__device__ __forceinline__ int coef_get(const TileShared& sm, int t, int k) {
    const uint32_t w = sm.coef[k >> 1][t];
    return (int)(short)((k & 1) ? (w >> 16) : w);
}

// Pixel origin of thread t's block.  A lane without a block of its own (last tile of an image) gets the
// tile's last block: it transforms a copy whose results land in its own shared-memory column, never read.
__device__ __forceinline__ void block_origin(const TileInfo& ti, int t, int& y0, int& x0) {
    const int b = ti.blk0 + (t < ti.nb ? t : ti.nb - 1);
    int br, bc;
    if (ti.bw_shift >= 0) {   // blocks per row is a power of two (uniform branch)
        br = b >> ti.bw_shift;
        bc = b & (ti.bw - 1);
    } else {
        br = b / ti.bw;
        bc = b - br * ti.bw;
    }
    y0 = br * 8;
    x0 = bc * 8;
}
__device__ __forceinline__ bool block_is_fast(const TileInfo& ti, int y0) {   // aligned interior block: 8 x LDG.64
    return ((ti.w & 7) == 0) && ((reinterpret_cast<uintptr_t>(ti.px) & 7) == 0) && (y0 + 8 <= ti.h);
}

// One block, all 32 lanes of a warp that owns at least one block call (the group test votes); `active` =
// this lane's block exists.
__device__ __forceinline__ void transform_block(const TileInfo& ti, const QuantParams& qp, TileShared& sm,
                                                int t, bool active, uint32_t& fl_lo, uint32_t& fl_hi) {
    int y0, x0;
    block_origin(ti, t, y0, x0);
    float d[64];
    const bool fast = block_is_fast(ti, y0);
    if (fast) {
        const uint8_t* p = ti.px + (size_t)y0 * ti.w + x0;
        uint2 rows[8];
#pragma unroll
        for (int i = 0; i < 8; i++) rows[i] = __ldg(reinterpret_cast<const uint2*>(p + (size_t)i * ti.w));
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                d[i * 8 + j] = (float)((rows[i].x >> (8 * j)) & 255u);
                d[i * 8 + 4 + j] = (float)((rows[i].y >> (8 * j)) & 255u);
            }
        }
    } else {
        int cx[8];
#pragma unroll
        for (int j = 0; j < 8; j++) cx[j] = reflect_idx(x0 + j, ti.w);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint8_t* row = ti.px + (size_t)reflect_idx(y0 + i, ti.h) * ti.w;
#pragma unroll
            for (int j = 0; j < 8; j++) d[i * 8 + j] = (float)__ldg(row + cx[j]);
        }
    }
    fdct8x8(d);
    d[0] -= 8192.0f;   // level shift (codec.py:29) only moves the DC term: 64 * 128, exact in FP32
    uint32_t nz_lo = 0, nz_hi = 0;
    int dc = 0;
    quantise_groups<0>(d, qp, sm, t, nz_lo, nz_hi, fl_lo, fl_hi, dc);
    sm.nz_lo[t] = nz_lo;
    sm.nz_hi[t] = nz_hi;
    sm.dcq[t] = dc;
    if (!active) fl_lo = fl_hi = 0;
}

// ---------------------------------------------------------------------------------------------
// C-variant stream (flag bit 30): the reference's embedded encoder c/img.c — integer AAN FDCT with 8-bit
// fixed-point constants, rows first, every stage result truncated to int16 (c/img.c:47-125), quantiser
// sign(d) * (((QUANT/2 + |d|) * scaledQuant) >> 16) (c/img.c:194-205).  Integer only: bit-exact by
// construction, no guard band, no exact path.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int as_i16(int v) { return (int)(short)v; }

__device__ __forceinline__ void aan8_c(int& d0, int& d1, int& d2, int& d3, int& d4, int& d5, int& d6, int& d7) {
    int tmp0 = d0 + d7, tmp7 = d0 - d7, tmp1 = d1 + d6, tmp6 = d1 - d6;
    int tmp2 = d2 + d5, tmp5 = d2 - d5, tmp3 = d3 + d4, tmp4 = d3 - d4;
    int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    d0 = as_i16(tmp10 + tmp11);
    d4 = as_i16(tmp10 - tmp11);
    const int z1 = ((tmp12 + tmp13) * 181) >> 8;
    d2 = as_i16(tmp13 + z1);
    d6 = as_i16(tmp13 - z1);
    tmp10 = tmp4 + tmp5;
    tmp11 = tmp5 + tmp6;
    tmp12 = tmp6 + tmp7;
    const int z5 = (tmp10 - tmp12) * 98;
    const int z2 = (z5 + tmp10 * 139) >> 8;
    const int z4 = (z5 + tmp12 * 334) >> 8;
    const int z3 = (tmp11 * 181) >> 8;
    const int z11 = tmp7 + z3, z13 = tmp7 - z3;
    d5 = as_i16(z13 + z2);
    d3 = as_i16(z13 - z2);
    d1 = as_i16(z11 + z4);
    d7 = as_i16(z11 - z4);
}

template <int R>
__device__ __forceinline__ int quantise_c(int d, int qfactor) {   // c/img.c:194-205
    constexpr int q = kQuantBaseDev(R) >> 1;
    const int sq = (int)c_cvar_scaled_quant[qfactor][R];
    const int a = d < 0 ? -d : d;
    const int m = ((q + a) * sq) >> 16;
    return as_i16(d < 0 ? -m : m);
}

template <int P>
__device__ __forceinline__ void quantise_pairs_c(const int (&d)[64], int qfactor, TileShared& sm, int t,
                                                 uint32_t& nz_lo, uint32_t& nz_hi, int& dc) {
    if constexpr (P < 32) {
        constexpr int k0 = 2 * P, k1 = k0 + 1;
        const int v0 = quantise_c<ZZ<k0>::r>(d[ZZ<k0>::r], qfactor), v1 = quantise_c<ZZ<k1>::r>(d[ZZ<k1>::r], qfactor);
        if constexpr (k0 < 32) {
            if (k0 != 0 && v0 != 0) nz_lo |= 0x80000000u >> k0;
            if (v1 != 0) nz_lo |= 0x80000000u >> k1;
        } else {
            if (v0 != 0) nz_hi |= 0x80000000u >> (k0 - 32);
            if (v1 != 0) nz_hi |= 0x80000000u >> (k1 - 32);
        }
        if constexpr (k0 == 0) dc = v0;
        sm.coef[P][t] = __byte_perm((uint32_t)v0, (uint32_t)v1, 0x5410);
        quantise_pairs_c<P + 1>(d, qfactor, sm, t, nz_lo, nz_hi, dc);
    }
}

// Block coordinates in the C variant.  (c/encode.c:47 loops `while (!feof(in))` and therefore codes one more
// block row after the image, from a stripe buffer whose tail the C library's stack frames have overwritten:
// its bits differ from run to run of the reference binary itself.  That row is not part of the image and is
// not produced here; everything before it is byte-identical.)
__device__ __forceinline__ void block_origin_c(const TileInfo& ti, int b, int& y0, int& x0) {
    int br, bc;
    if (ti.bw_shift >= 0) {
        br = b >> ti.bw_shift;
        bc = b & (ti.bw - 1);
    } else {
        br = b / ti.bw;
        bc = b - br * ti.bw;
    }
    y0 = br * 8;
    x0 = bc * 8;
}

__device__ __forceinline__ void transform_block_c(const TileInfo& ti, int qfactor, TileShared& sm, int t, bool active) {
    int y0, x0;
    block_origin_c(ti, ti.blk0 + (active ? t : ti.nb - 1), y0, x0);
    const uint8_t* p = ti.px + (size_t)y0 * ti.w + x0;
    int d[64];
    if ((reinterpret_cast<uintptr_t>(ti.px) & 7) == 0) {   // width is a multiple of 8 in this mode
        uint2 rows[8];
#pragma unroll
        for (int i = 0; i < 8; i++) rows[i] = __ldg(reinterpret_cast<const uint2*>(p + (size_t)i * ti.w));
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {   // block[i] = (int8_t)(data[i] ^ 0x80), c/img.c:213
                d[i * 8 + j] = (int)((rows[i].x >> (8 * j)) & 255u) - 128;
                d[i * 8 + 4 + j] = (int)((rows[i].y >> (8 * j)) & 255u) - 128;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) d[i * 8 + j] = (int)__ldg(p + (size_t)i * ti.w + j) - 128;
    }
#pragma unroll
    for (int r = 0; r < 8; r++)   // rows first (c/img.c:54)
        aan8_c(d[r * 8], d[r * 8 + 1], d[r * 8 + 2], d[r * 8 + 3], d[r * 8 + 4], d[r * 8 + 5], d[r * 8 + 6], d[r * 8 + 7]);
#pragma unroll
    for (int c = 0; c < 8; c++)
        aan8_c(d[c], d[8 + c], d[16 + c], d[24 + c], d[32 + c], d[40 + c], d[48 + c], d[56 + c]);
    uint32_t nz_lo = 0, nz_hi = 0;
    int dc = 0;
    quantise_pairs_c<0>(d, qfactor, sm, t, nz_lo, nz_hi, dc);
    sm.nz_lo[t] = nz_lo;
    sm.nz_hi[t] = nz_hi;
    sm.dcq[t] = dc;
}

// Phase 1 of the C variant for the 32 blocks of one warp.  The DC of the block in front of the warp is the
// quantised (sum of pixels - 8192): the DC path of c/img.c has no rounding before the quantiser.
__device__ __forceinline__ void transform_warp_c(const TileInfo& ti, int qfactor, TileShared& sm) {
    const int t = tid(), lane = t & 31, warp = t >> 5;
    if (warp * 32 >= ti.nb) return;   // warp-uniform
    transform_block_c(ti, qfactor, sm, t, t < ti.nb);
    int halo_dc = 0;
    const int hb = ti.blk0 + warp * 32 - 1;
    if (hb >= 0) {
        int y0, x0;
        block_origin_c(ti, hb, y0, x0);
        int sum = 0;
        if (lane < 8) {
            const uint8_t* p = ti.px + (size_t)(y0 + lane) * ti.w + x0;
#pragma unroll
            for (int j = 0; j < 8; j++) sum += (int)__ldg(p + j);
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        halo_dc = quantise_c<0>(as_i16(sum - 8192), qfactor);
    }
    if (lane == 0) sm.dc_halo[warp] = halo_dc;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// phase 2: exact recomputation of every flagged coefficient of the warp's 32 blocks (and of the DC
// of the block in front of them, which the DC difference of codec.py:34-35 needs).  8 lanes per
// entry: lane c transforms column c of the block (axis -2 first, utils.py:33-34), lane 0 the row.
// ---------------------------------------------------------------------------------------------
// Exact-path statistics live in shared memory (TileShared::stat_*): rare events, no registers.
// stat_unflagged (TIC_FLAG_DEBUG_ALL_EXACT): the exact value differs although the guard did not flag it.
struct ExactStats {   // handle passed around instead of three counters
    TileShared* sm;
    __device__ __forceinline__ void items(unsigned n) const;
    __device__ __forceinline__ void changed(unsigned n) const;
    __device__ __forceinline__ void unflagged(unsigned n) const;
};
__device__ __forceinline__ void ExactStats::items(unsigned n) const { atomicAdd(&sm->stat_items, n); }
__device__ __forceinline__ void ExactStats::changed(unsigned n) const { atomicAdd(&sm->stat_changed, n); }
__device__ __forceinline__ void ExactStats::unflagged(unsigned n) const { atomicAdd(&sm->stat_unflagged, n); }
// worklist entry: lane << 6 | zigzag index; bit 31: the DC of the block in front of the tile / warp;
// bit 30: flagged by the guard band (always set outside TIC_FLAG_DEBUG_ALL_EXACT)
constexpr uint32_t kWorkHalo = 0x80000000u, kWorkGuard = 0x40000000u;

// One exact-path item per 8-lane group of the warp (grp = lane >> 3, c = lane & 7; `act`, `halo`, `guard`, `owner`
// (thread in tile) and `k` (zigzag index) are uniform over the group): 8 lanes recompute the column pass from the pixels,
// lane c == 0 the row pass, the division and the update.  All 32 lanes call (warp barriers inside).
__device__ __forceinline__ void exact_item(const TileInfo& ti, const QuantParams& qp, TileShared& sm, int warp, int grp,
                                           int c, bool act, bool halo, bool guard, int owner, int k, const ExactStats& st) {
    {
        {
        const int wt0 = warp * 32;   // first thread of the warp
        const int r = c_zigzag[k];
        const int u = r >> 3, v = r & 7;
        const int b = halo ? ti.blk0 + wt0 - 1 : ti.blk0 + owner;
        if (act) {
            int br, bc;
            if (ti.bw_shift >= 0) { br = b >> ti.bw_shift; bc = b & (ti.bw - 1); }   // uniform branch: no division routine
            else { br = b / ti.bw; bc = b - br * ti.bw; }
            const int y0 = br * 8, x = bc * 8 + c;
            double x0, x1, x2, x3, x4, x5, x6, x7;
            const bool interior = y0 + 8 <= ti.h && bc * 8 + 8 <= ti.w;   // no reflection (uniform per group)
#if TIC_EXACT_INT_COLS
            if (interior && (u & 3) == 0) {
                // u = 0 or 4 (9 in 10 of all items: the true .5 ties sit at (0,0), (4,0), (0,4), (4,4)): every operation
                // of dct8_exact's column pass in front of its last multiplication acts on small integers and is exact —
                // s = 0.5 * (x0 + x7 + x3 + x4 +- (x1 + x2 + x5 + x6)) — so the result is RN(s * HSQ) resp. RN(s * TW3),
                // from one integer sum (the same identity as dc_exact_from_colsums / settle_rational)
                const uint8_t* p = ti.px + (size_t)y0 * ti.w + x;
                const size_t w = (size_t)ti.w;
                const int p0 = __ldg(p), p1 = __ldg(p + w), p2 = __ldg(p + 2 * w), p3 = __ldg(p + 3 * w);
                const int p4 = __ldg(p + 4 * w), p5 = __ldg(p + 5 * w), p6 = __ldg(p + 6 * w), p7 = __ldg(p + 7 * w);
                const int outer = p0 + p7 + p3 + p4, inner = p1 + p2 + p5 + p6;
                const int sum = u == 0 ? outer + inner - 1024 : outer - inner;   // level shift: 8 x 128 resp. 0
                const double m = u == 0 ? 0x1.6a09e667f3bcdp-1 : 0x1.6a09e667f3bccp-1;   // HSQ : TW3
                exact_colres(sm, warp, grp)[c] = __dmul_rn(__dmul_rn(0.5, (double)sum), m);
            } else
#endif
            if (interior) {
                const uint8_t* p = ti.px + (size_t)y0 * ti.w + x;
                const size_t w = (size_t)ti.w;
                x0 = (double)((int)__ldg(p) - 128);         x1 = (double)((int)__ldg(p + w) - 128);
                x2 = (double)((int)__ldg(p + 2 * w) - 128); x3 = (double)((int)__ldg(p + 3 * w) - 128);
                x4 = (double)((int)__ldg(p + 4 * w) - 128); x5 = (double)((int)__ldg(p + 5 * w) - 128);
                x6 = (double)((int)__ldg(p + 6 * w) - 128); x7 = (double)((int)__ldg(p + 7 * w) - 128);
                exact_colres(sm, warp, grp)[c] = dct8_exact(x0, x1, x2, x3, x4, x5, x6, x7, u);
            } else {
                x0 = load_px_exact(ti, y0 + 0, x); x1 = load_px_exact(ti, y0 + 1, x);
                x2 = load_px_exact(ti, y0 + 2, x); x3 = load_px_exact(ti, y0 + 3, x);
                x4 = load_px_exact(ti, y0 + 4, x); x5 = load_px_exact(ti, y0 + 5, x);
                x6 = load_px_exact(ti, y0 + 6, x); x7 = load_px_exact(ti, y0 + 7, x);
                exact_colres(sm, warp, grp)[c] = dct8_exact(x0, x1, x2, x3, x4, x5, x6, x7, u);
            }
        }
        __syncwarp();
        if (act && c == 0) {
            const double* cr = exact_colres(sm, warp, grp);
            double y;
#if TIC_EXACT_ROW_INLINE
            if ((v & 3) == 0) {   // outputs 0 and 4: dct8_exact's sequence for them, without the call
                const double c0 = __dmul_rn(2.0, cr[0]), c7 = __dmul_rn(2.0, cr[7]);
                const double c1 = __dadd_rn(cr[1], cr[2]), c3 = __dadd_rn(cr[3], cr[4]), c5 = __dadd_rn(cr[5], cr[6]);
                const double h0 = __dadd_rn(c0, c7), h3 = __dmul_rn(2.0, c3), h1 = __dadd_rn(c1, c5);
                const double a = __dadd_rn(h0, h3), e1 = __dmul_rn(2.0, h1);
                y = v == 0 ? __dmul_rn(__dmul_rn(0.25, __dadd_rn(a, e1)), 0x1.6a09e667f3bcdp-1)
                           : __dmul_rn(__dmul_rn(0.25, __dsub_rn(a, e1)), 0x1.6a09e667f3bccp-1);
            } else
#endif
            y = dct8_exact(cr[0], cr[1], cr[2], cr[3], cr[4], cr[5], cr[6], cr[7], v);
            int q = __double2int_rn(__ddiv_rn(y, qp.qt[r]));   // np.round(coeffs / qt), utils.py:53
            if (halo) {
                sm.dc_halo[warp] = q;
            } else {
                int old = k == 0 ? sm.dcq[owner] : coef_get(sm, owner, k);
                if (old != q) {
                    // two flagged coefficients of one block may share a packed word: a 16-bit store touches its own half only
                    reinterpret_cast<short*>(&sm.coef[k >> 1][owner])[k & 1] = (short)q;
                    if (k == 0) {
                        sm.dcq[owner] = q;
                    } else {
                        uint32_t* m = k < 32 ? &sm.nz_lo[owner] : &sm.nz_hi[owner];
                        const uint32_t bit = 0x80000000u >> (k & 31);
                        if (q) atomicOr(m, bit); else atomicAnd(m, ~bit);
                    }
                    st.changed(1u);
                    if (!guard) st.unflagged(1u);
                }
            }
        }
        __syncwarp();
        }
    }
}

// The warp's worklist in shared memory (sm.work[warp][0 .. count)), four entries at a time.
__device__ __forceinline__ void exact_round(const TileInfo& ti, const QuantParams& qp, TileShared& sm,
                                            int warp, int count, const ExactStats& st) {
    const int lane = threadIdx.x & 31;
    const int grp = lane >> 3, c = lane & 7;
    for (int base = 0; base < count; base += 4) {
        const int idx = base + grp;
        const bool act = idx < count;   // uniform per 8-lane group
        const uint32_t item = act ? sm.work[warp][idx] : 0;
        exact_item(ti, qp, sm, warp, grp, c, act, (item >> 31) != 0, (item & kWorkGuard) != 0,
                   warp * 32 + (int)((item >> 6) & 31), (int)(item & 63), st);
    }
}

// Out of line (rare): the thread's block has its DC on a rounding tie.  The 8 column sums of the block are
// all the reference's DC depends on (dc_exact_from_colsums); the pixels are re-read — they are in L1/L2.
// Returns false for blocks that need reflection padding: those stay with the worklist.
// (Scalars by value: taking the TileInfo by reference would pin it in local memory for the whole kernel.)
__device__ __noinline__ bool settle_dc_ties(const uint8_t* px, int w, int h, int bw, int blk0, int nb, double qt0,
                                            uint32_t* coef0 /* &sm.coef[0][0] */, int* dcq, uint32_t fl_lo) {
    const int t = tid();
    if (!(fl_lo & 0x80000000u)) return false;
    const int b = blk0 + (t < nb ? t : nb - 1);
    const int br = b / bw, bc = b - br * bw;
    const int y0 = br * 8, x0 = bc * 8;
    if ((w & 7) != 0 || (reinterpret_cast<uintptr_t>(px) & 7) != 0 || y0 + 8 > h) return false;
    const uint8_t* p = px + (size_t)y0 * w + x0;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;   // column sums, two 16-bit lanes per word
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p + (size_t)i * w));
        w0 += __byte_perm(v.x, 0u, 0x4140); w1 += __byte_perm(v.x, 0u, 0x4342);
        w2 += __byte_perm(v.y, 0u, 0x4140); w3 += __byte_perm(v.y, 0u, 0x4342);
    }
    const int dc = dc_exact_from_colsums(
        (double)((int)(w0 & 0xffffu) - 1024), (double)((int)(w0 >> 16) - 1024), (double)((int)(w1 & 0xffffu) - 1024),
        (double)((int)(w1 >> 16) - 1024), (double)((int)(w2 & 0xffffu) - 1024), (double)((int)(w2 >> 16) - 1024),
        (double)((int)(w3 & 0xffffu) - 1024), (double)((int)(w3 >> 16) - 1024), qt0);
    reinterpret_cast<unsigned short*>(&coef0[t])[0] = (unsigned short)dc;
    dcq[t] = dc;
    return true;
}

// Phases 1 and 2 for the 32 blocks of one warp; only warp-level synchronisation.  On return
// sm.coef / nz / dcq / dc_halo hold the reference's quantised coefficients for those blocks.
__device__ __forceinline__ void transform_warp(const TileInfo& ti, const QuantParams& qp, TileShared& sm,
                                               const ExactStats& st) {
    const int t = tid(), lane = t & 31, warp = t >> 5;
    if (lane == 0) {
        sm.pending[warp] = 0;
        sm.dc_halo[warp] = 0;
        sm.work_count[warp] = 0;
    }
    __syncwarp();
    uint32_t fl_lo = 0, fl_hi = 0;
    if (warp * 32 < ti.nb) transform_block(ti, qp, sm, t, t < ti.nb, fl_lo, fl_hi);   // warp-uniform
    // A DC in the guard band of a tie: 1 block in 128 on ordinary content — that one rides along in the
    // exact-path worklist — but EVERY block of a flat area whose level lands on a tie (at quality 50: any odd
    // pixel value, 255 included).  When several lanes of the warp are affected, each settles its own DC in
    // float64 (settle_dc_ties, kept out of line): one pass for all of them instead of one entry per block.
    if (__popc(__ballot_sync(0xffffffffu, (fl_lo & 0x80000000u) != 0)) >= 4) {   // warp-uniform
        if (settle_dc_ties(ti.px, ti.w, ti.h, ti.bw, ti.blk0, ti.nb, qp.qt[0], &sm.coef[0][0], sm.dcq, fl_lo)) {
            fl_lo &= 0x7fffffffu;
            st.items(1u);
        }
    }
    // halo: quantised DC of the block in front of the warp's first block.  The DC coefficient is
    // (sum of pixels - 8192) / 8 exactly, so unless its quotient by qt lands within 1e-9 of a .5 tie
    // (where the reference's float64 rounding errors decide) one pixel sum settles it; otherwise, and
    // for blocks that need reflection padding, it becomes an exact-path item.
    int halo_item = 0, halo_dc = 0;
    const int hb = ti.blk0 + warp * 32 - 1;
    if (hb >= 0 && warp * 32 < ti.nb) {   // warp-uniform
        const int br = hb / ti.bw, bc = hb - br * ti.bw;
        const int y0 = br * 8;
        halo_item = 1;
        if (((ti.w & 7) == 0) && ((reinterpret_cast<uintptr_t>(ti.px) & 7) == 0) && (y0 + 8 <= ti.h)) {
            uint2 v = make_uint2(0u, 0u);   // lanes 0..7: one pixel row of that block each
            if (lane < 8) v = __ldg(reinterpret_cast<const uint2*>(ti.px + (size_t)(y0 + lane) * ti.w + bc * 8));
            int sum = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, 0u));
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            sum = __shfl_sync(0xffffffffu, sum, 0);   // lanes 8..31 summed zeros: every lane must take the same branch below
            const double tq = __dmul_rn((double)(sum - 8192), qp.dcinv);
            halo_dc = __double2int_rn(tq);
            halo_item = 0;
            if (!(fabs(tq - (double)halo_dc) < 0.5 - 1.0e-9)) {   // on a tie: the 8 column sums decide (warp-uniform)
                uint32_t w0 = __byte_perm(v.x, 0u, 0x4140), w1 = __byte_perm(v.x, 0u, 0x4342);   // 16-bit lanes
                uint32_t w2 = __byte_perm(v.y, 0u, 0x4140), w3 = __byte_perm(v.y, 0u, 0x4342);
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    w0 += __shfl_xor_sync(0xffffffffu, w0, o); w1 += __shfl_xor_sync(0xffffffffu, w1, o);
                    w2 += __shfl_xor_sync(0xffffffffu, w2, o); w3 += __shfl_xor_sync(0xffffffffu, w3, o);
                }
                if (lane == 0)
                    halo_dc = dc_exact_from_colsums((double)((int)(w0 & 0xffffu) - 1024), (double)((int)(w0 >> 16) - 1024),
                                                    (double)((int)(w1 & 0xffffu) - 1024), (double)((int)(w1 >> 16) - 1024),
                                                    (double)((int)(w2 & 0xffffu) - 1024), (double)((int)(w2 >> 16) - 1024),
                                                    (double)((int)(w3 & 0xffffu) - 1024), (double)((int)(w3 >> 16) - 1024),
                                                    qp.qt[0]);
            }
        }
    }
    if (lane == 0) {
        if (halo_item) sm.work[warp][atomicAdd(&sm.work_count[warp], 1)] = 0x80000000u;   // slot 0: nothing pushed yet
        else sm.dc_halo[warp] = halo_dc;
    }
    __syncwarp();
    // push flagged coefficients; loop in rounds if the worklist overflows (high quality only)
    while (true) {
        while (fl_lo | fl_hi) {
            const int k = fl_lo ? __clz(fl_lo) : 32 + __clz(fl_hi);
            const int slot = atomicAdd(&sm.work_count[warp], 1);
            if (slot >= kWarpWork) { atomicAdd(&sm.pending[warp], 1); break; }
            sm.work[warp][slot] = kWorkGuard | ((uint32_t)lane << 6) | (uint32_t)k;
            if (k < 32) fl_lo ^= 0x80000000u >> k; else fl_hi ^= 0x80000000u >> (k - 32);
        }
        __syncwarp();
        const int raw = sm.work_count[warp];
        const int count = raw < kWarpWork ? raw : kWarpWork;
        const int pending = sm.pending[warp];
        __syncwarp();
        if (count) {   // warp-uniform
            if (lane == 0) st.items((unsigned)count);   // (DC ties settled above count too)
            exact_round(ti, qp, sm, warp, count, st);
        }
        if (pending == 0) break;
        if (lane == 0) { sm.work_count[warp] = 0; sm.pending[warp] = 0; }
        __syncwarp();
    }
    __syncwarp();
}


// ---------------------------------------------------------------------------------------------
// phases 1 + 2 on the tensor cores (tic_tc.cuh), for the kTile = 128 blocks of one tile; all threads of the
// group call.  Pixels -> f16 A operand in shared memory (it aliases sm.coef: the quantised coefficients only
// arrive after the MMA is complete) -> 8 MMAs issued by thread 0 -> the thread that owns a block reads its 64
// scaled coefficients t / hthr * 2^E back from tensor memory, 16 columns at a time, and quantises them in
// zigzag groups of 8 exactly like the FP32 path (group vote, magic-number rounding, residual test against the
// guard band).  The exact path is unchanged.  On return (after a group barrier) sm.coef / nz / dcq hold the
// reference's quantised coefficients of the whole tile and sm.dc_halo[0] the DC in front of its first block.
// ---------------------------------------------------------------------------------------------
// Everything but the mbarrier phase is recomputed where it is used (a handful of instructions per tile) instead
// of living in registers across the whole tile loop.
struct TcGroup {
    uint32_t ctl;      // shared address of the CTA's control block: mbarrier g at +8g, the TMEM base address at +64;
                       // the B operand sits tc::kBBytes in front of it
    uint32_t phase;
    __device__ __forceinline__ uint32_t bar(int g) const { return ctl + 8u * (uint32_t)g; }
    __device__ __forceinline__ uint64_t desc_b0() const { return tc::smem_desc(ctl - (uint32_t)tc::kBBytes, tc::kLboB, tc::kSboB); }
    // the group's accumulator columns, lane field = the calling warp's quarter of the 128 lanes
    __device__ __forceinline__ uint32_t tmem(int g) const {
        uint32_t base;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(base) : "r"(ctl + 64u));
        return base + (uint32_t)(g * tc::kColsPerGroup) + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16);
    }
};

template <int G, int P>
__device__ __forceinline__ void quantise_pairs_tc(const uint32_t* r /* 8 columns of group G */, const QuantParams& qp,
                                                  TileShared& sm, int t, uint32_t& nz, uint32_t& fl, int& dc) {
    if constexpr (P < 4) {
        constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23
        constexpr int k0 = 8 * G + 2 * P, k1 = k0 + 1;
        const float d0 = __uint_as_float(r[2 * P]), d1 = __uint_as_float(r[2 * P + 1]);
#if TIC_QUANT_F32X2
        // packed FP32 (FFMA2 / FADD2): the pair's three operations in three issue slots instead of six
        const f32x2 d01 = pk2(d0, d1), cn01 = pk2(qp.cn[k0], qp.cn[k1]), mg = pk2(kMagic, kMagic);
        const f32x2 b01 = fma2(d01, cn01, mg);
        const f32x2 nq01 = sub2(mg, b01);             // -(round(t))
        const f32x2 e01 = fma2(d01, cn01, nq01);      // t - round(t)
        float b0, b1, q0, q1, e0, e1;
        upk2(b01, b0, b1); upk2(nq01, q0, q1); upk2(e01, e0, e1);
#else
        const float b0 = fmaf(d0, qp.cn[k0], kMagic), b1 = fmaf(d1, qp.cn[k1], kMagic);
        const float q0 = b0 - kMagic, q1 = b1 - kMagic;
        const float e0 = fmaf(d0, qp.cn[k0], -q0), e1 = fmaf(d1, qp.cn[k1], -q1);   // t - round(t)
#endif
        if (k0 != 0 && q0 != 0.0f) nz |= 0x80000000u >> (k0 & 31);
        if (q1 != 0.0f) nz |= 0x80000000u >> (k1 & 31);
        if (fabsf(e0) > qp.hthr[k0]) fl |= 0x80000000u >> (k0 & 31);   // the exact path decides
        if (fabsf(e1) > qp.hthr[k1]) fl |= 0x80000000u >> (k1 & 31);
        if constexpr (k0 == 0) dc = __float_as_int(b0) - 0x4B400000;
        sm.coef[k0 >> 1][t] = __byte_perm(__float_as_uint(b0), __float_as_uint(b1), 0x5410);
        quantise_pairs_tc<G, P + 1>(r, qp, sm, t, nz, fl, dc);
    }
}

template <int G>
__device__ __forceinline__ void quantise_group_tc(const uint32_t* r, const QuantParams& qp, TileShared& sm, int t,
                                                  bool all_live, uint32_t& nz_lo, uint32_t& nz_hi, uint32_t& fl_lo,
                                                  uint32_t& fl_hi, int& dc) {
    bool live = true;
    if constexpr (G > 0) {
        float m = fmax3(fabsf(__uint_as_float(r[0])), fabsf(__uint_as_float(r[1])), fabsf(__uint_as_float(r[2])));
        m = fmax3(m, fabsf(__uint_as_float(r[3])), fabsf(__uint_as_float(r[4])));
        m = fmax3(m, fabsf(__uint_as_float(r[5])), fabsf(__uint_as_float(r[6])));
        m = fmaxf(m, fabsf(__uint_as_float(r[7])));
        live = __any_sync(0xffffffffu, m >= qp.tc_live) || all_live;
    }
    if (live) {   // warp-uniform
        if constexpr (G < 4) quantise_pairs_tc<G, 0>(r, qp, sm, t, nz_lo, fl_lo, dc);
        else quantise_pairs_tc<G, 0>(r, qp, sm, t, nz_hi, fl_hi, dc);
    }
}

// Pixels of thread t's block as 8 rows of 8 bytes (reflection padding where the block leaves the image).
__device__ __forceinline__ void load_block_rows(const TileInfo& ti, int t, uint2 (&rows)[8]) {
    int y0, x0;
    block_origin(ti, t, y0, x0);
    if (block_is_fast(ti, y0)) {
        const uint8_t* p = ti.px + (size_t)y0 * ti.w + x0;
#pragma unroll
        for (int i = 0; i < 8; i++) rows[i] = __ldg(reinterpret_cast<const uint2*>(p + (size_t)i * ti.w));
    } else {
        int cx[8];
#pragma unroll
        for (int j = 0; j < 8; j++) cx[j] = reflect_idx(x0 + j, ti.w);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint8_t* row = ti.px + (size_t)reflect_idx(y0 + i, ti.h) * ti.w;
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                lo |= (uint32_t)__ldg(row + cx[j]) << (8 * j);
                hi |= (uint32_t)__ldg(row + cx[4 + j]) << (8 * j);
            }
            rows[i] = make_uint2(lo, hi);
        }
    }
}

// Exact value of the coefficients of thread t's block that can sit on true .5 ties — (0,0), (4,0), (0,4), (4,4) =
// zigzag 0, 10, 14, 39 — from the 16 column sums the tensor core left in the accumulator (tic_tc.cuh): the
// reference's float64 column pass (axis -2 first, utils.py:33-34) for u = 0 is RN(0.5 * S_x * HSQ) and for u = 4
// RN(0.5 * I_x * TW3), every earlier operation being exact on integers; the row pass is
// dct8_exact over those eight doubles.  rat: bit 0 / 1 / 2 / 3 = zigzag 0 / 10 / 14 / 39 flagged.  Out of line, one
// lane at a time in practice; only this thread touches its own column of sm.coef here.
__device__ __noinline__ unsigned settle_rational(uint32_t taddr /* the lane's accumulator, column 0 */, uint32_t rat,
                                                 const double* __restrict__ qt /* QuantParams::qt */, TileShared* smp, int t) {
    uint32_t r[16];   // every lane of the warp calls (tcgen05.ld is warp-wide); lanes without a flag return below
    tc::tmem_ld16(taddr + tc::kColSums, r);
    tc::tmem_wait_ld(r);
    if (!rat) return 0;
    const float s0 = __uint_as_float(r[0]), s1 = __uint_as_float(r[1]), s2 = __uint_as_float(r[2]), s3 = __uint_as_float(r[3]);
    const float s4 = __uint_as_float(r[4]), s5 = __uint_as_float(r[5]), s6 = __uint_as_float(r[6]), s7 = __uint_as_float(r[7]);
    const float i0 = __uint_as_float(r[8]), i1 = __uint_as_float(r[9]), i2 = __uint_as_float(r[10]), i3 = __uint_as_float(r[11]);
    const float i4 = __uint_as_float(r[12]), i5 = __uint_as_float(r[13]), i6 = __uint_as_float(r[14]), i7 = __uint_as_float(r[15]);
    const double qt00 = qt[0], qt40 = qt[32], qt04 = qt[4], qt44 = qt[36];
    uint32_t* coef_col = &smp->coef[0][t];
    uint32_t* nz_lo = &smp->nz_lo[t];
    uint32_t* nz_hi = &smp->nz_hi[t];
    int* dcq = &smp->dcq[t];
    const double HSQ = 0x1.6a09e667f3bcdp-1, TW3 = 0x1.6a09e667f3bccp-1;
    unsigned changed = 0;
#pragma unroll 1
    for (int which = 0; which < 4; which++) {
        if (!((rat >> which) & 1u)) continue;
        const bool u4 = (which == 1 || which == 3), v4 = (which >= 2);
        const double m = u4 ? TW3 : HSQ;
#if TIC_RAT_FAST
        // dct8_exact(.., 0 or 4) over c_x = RN(RN(0.5 s_x) m), with every multiplication by a power of two moved to
        // the end (exact: they commute with rounding, nothing here is near overflow or underflow):
        // y = 0.25 RN(RN(RN(RN(X0 + X7) + RN(X3 + X4)) +- RN(RN(X1 + X2) + RN(X5 + X6))) K), X_x = RN(s_x m).
        // Checked bit for bit against the long form on 8e7 random column-sum vectors (tools/rat_check.c).
        const double X0 = __dmul_rn((double)(u4 ? i0 : s0), m), X1 = __dmul_rn((double)(u4 ? i1 : s1), m);
        const double X2 = __dmul_rn((double)(u4 ? i2 : s2), m), X3 = __dmul_rn((double)(u4 ? i3 : s3), m);
        const double X4 = __dmul_rn((double)(u4 ? i4 : s4), m), X5 = __dmul_rn((double)(u4 ? i5 : s5), m);
        const double X6 = __dmul_rn((double)(u4 ? i6 : s6), m), X7 = __dmul_rn((double)(u4 ? i7 : s7), m);
        const double A = __dadd_rn(__dadd_rn(X0, X7), __dadd_rn(X3, X4)), H = __dadd_rn(__dadd_rn(X1, X2), __dadd_rn(X5, X6));
        const double y = __dmul_rn(0.25, __dmul_rn(v4 ? __dsub_rn(A, H) : __dadd_rn(A, H), v4 ? TW3 : HSQ));
#else
        const double c0 = __dmul_rn(__dmul_rn(0.5, (double)(u4 ? i0 : s0)), m), c1 = __dmul_rn(__dmul_rn(0.5, (double)(u4 ? i1 : s1)), m);
        const double c2 = __dmul_rn(__dmul_rn(0.5, (double)(u4 ? i2 : s2)), m), c3 = __dmul_rn(__dmul_rn(0.5, (double)(u4 ? i3 : s3)), m);
        const double c4 = __dmul_rn(__dmul_rn(0.5, (double)(u4 ? i4 : s4)), m), c5 = __dmul_rn(__dmul_rn(0.5, (double)(u4 ? i5 : s5)), m);
        const double c6 = __dmul_rn(__dmul_rn(0.5, (double)(u4 ? i6 : s6)), m), c7 = __dmul_rn(__dmul_rn(0.5, (double)(u4 ? i7 : s7)), m);
        const double y = dct8_exact(c0, c1, c2, c3, c4, c5, c6, c7, v4 ? 4 : 0);
#endif
        const double qt = which == 0 ? qt00 : (which == 1 ? qt40 : (which == 2 ? qt04 : qt44));
        const int q = __double2int_rn(__ddiv_rn(y, qt));   // np.round(coeffs / qt), utils.py:53
        const int k = which == 0 ? 0 : (which == 1 ? 10 : (which == 2 ? 14 : 39));
        uint32_t* wp = coef_col + (k >> 1) * kTile;
        const uint32_t w = *wp;
        const int old = (int)(short)((k & 1) ? (w >> 16) : w);
        if (old != q) {
            const uint32_t qb = (uint32_t)q & 0xffffu;
            *wp = (k & 1) ? ((w & 0x0000ffffu) | (qb << 16)) : ((w & 0xffff0000u) | qb);
            if (k == 0) {
                *dcq = q;
            } else {
                uint32_t* mp = k < 32 ? nz_lo : nz_hi;
                const uint32_t bit = 0x80000000u >> (k & 31);
                *mp = q ? (*mp | bit) : (*mp & ~bit);
            }
            changed++;
        }
    }
    return changed;
}

// rows: the pixel rows of this thread's block (load_block_rows), loaded by the caller — one tile ahead in the
// persistent kernel; `next` (may be null): the tile whose rows are fetched into `rows` while the tensor core works.
// The block in front of the tile (it belongs to another tile, i.e. another CTA): lanes 0..7 of warp 0 fetch one pixel
// row of it each, TOGETHER with the tile's own rows, so that its DRAM latency is paid once, not again behind the MMA.
__device__ __forceinline__ bool tile_halo_is_fast(const TileInfo& ti, int& y0, int& x0) {
    const int hb = ti.blk0 - 1;
    if (hb < 0) return false;
    int br, bc;
    if (ti.bw_shift >= 0) { br = hb >> ti.bw_shift; bc = hb & (ti.bw - 1); }
    else { br = hb / ti.bw; bc = hb - br * ti.bw; }
    y0 = br * 8; x0 = bc * 8;
    return ((ti.w & 7) == 0) && ((reinterpret_cast<uintptr_t>(ti.px) & 7) == 0) && (y0 + 8 <= ti.h);
}
__device__ __forceinline__ uint2 load_tile_halo(const TileInfo& ti, int t) {
    uint2 v = make_uint2(0u, 0u);
    int y0, x0;
    if (t < 8 && tile_halo_is_fast(ti, y0, x0)) v = __ldg(reinterpret_cast<const uint2*>(ti.px + (size_t)(y0 + t) * ti.w + x0));
    return v;
}

// The pixel rows of tile `ti` -> L2, one cp.async.bulk.prefetch per pixel row and block row of the tile (lane = block
// row of the tile * 8 + pixel row): called by one warp for the group's NEXT tile while the tensor core works on the
// current one, so that the 8 x LDG.64 of load_block_rows find their lines in L2 instead of in HBM (the kernel keeps
// 6 warps per scheduler: a DRAM round trip at the top of every tile is not hidden, profiles/r2e: 8.5 % of all stall
// samples).  Rows that are not 16-byte aligned, and tiles that span more than four block rows, are left alone.
__device__ __forceinline__ void prefetch_tile_l2(const TileInfo& ti, int lane) {
    if (ti.nb <= 0 || (ti.w & 15) != 0 || (reinterpret_cast<uintptr_t>(ti.px) & 15) != 0) return;
    const int last = ti.blk0 + ti.nb - 1;
    int br0, br1;
    if (ti.bw_shift >= 0) { br0 = ti.blk0 >> ti.bw_shift; br1 = last >> ti.bw_shift; }
    else { br0 = ti.blk0 / ti.bw; br1 = last / ti.bw; }
    const int nrows = br1 - br0 + 1;
    const int r = lane >> 3, y = lane & 7;
    if (nrows > 4 || r >= nrows) return;
    const int Y = (br0 + r) * 8 + y;
    if (Y >= ti.h) return;
    int xa = r == 0 ? (ti.blk0 - br0 * ti.bw) * 8 : 0;
    int xb = r == nrows - 1 ? (last - br1 * ti.bw + 1) * 8 : ti.w;
    xa &= ~15;
    xb = (xb + 15) & ~15;
    if (xb > ti.w) xb = ti.w;
    if (xb <= xa) return;
    const uint8_t* p = ti.px + (size_t)Y * ti.w + xa;
#if TIC_PREFETCH_L2 == 2
    // one 128-byte line per instruction and lane (the bulk form takes its operands from uniform registers: the
    // compiler serialises the lanes)
    const uint8_t* pe = p + (xb - xa);
    for (p = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)127); p < pe; p += 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
#else
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)(xb - xa)) : "memory");
#endif
}

// after_issue(): called by every thread of the group right after the MMA has been issued (the kernel's tile loop
// uses it to prepare the description of the group's next tile in the tensor core's shadow).
template <int G, class AfterIssue>
__device__ __forceinline__ void transform_tile_tc(const TileInfo& ti, const QuantParams& qp, TileShared& sm,
                                                  TcGroup& tg, int g, bool debug_all, const ExactStats& st,
                                                  uint2 (&rows)[8], uint2 halo_v, AfterIssue after_issue) {
    const int t = tid(), lane = t & 31, warp = t >> 5;
    if (lane == 0) {
        sm.pending[warp] = 0;
        sm.dc_halo[warp] = 0;
        sm.work_count[warp] = 0;
    }
    // ---- pixels -> A operand ----------------------------------------------------------------------------
    {
        unsigned char* dst = reinterpret_cast<unsigned char*>(&sm.coef[0][0]) + (t >> 3) * 128 + (t & 7) * 16;
#pragma unroll
        for (int y = 0; y < 8; y++) *reinterpret_cast<uint4*>(dst + y * (int)tc::kLboA) = tc::row_to_f16(rows[y]);
        if (lane == 31) {   // the next warp's DC predictor needs this block's DC: its pixel sum settles it (below)
            uint32_t sum = 0;
#pragma unroll
            for (int y = 0; y < 8; y++) sum = __dp4a(rows[y].x, 0x01010101u, __dp4a(rows[y].y, 0x01010101u, sum));
            sm.blocksum[warp] = (int)sum;
        }
    }
    tc::fence_proxy_async();    // generic-proxy stores -> visible to the tensor core's async proxy
    tc::fence_before_sync();    // this thread's TMEM reads of the previous tile are ordered before the barrier
    group_sync<G>(g);
    if (t == 0) {
        tc::fence_after_sync();
        tc::issue_tile_mma(tc::smem_desc(tc::smem_addr(&sm.coef[0][0]), tc::kLboA, tc::kSboA), tg.desc_b0(),
                           tg.tmem(g) & 0x0000ffffu, tg.bar(g));
    }
    after_issue();
    // ---- while the tensor core works: quantised DC of the block in front of the warp's first block ------------
    // (codec.py:34-35).  The DC coefficient is (sum of pixels - 8192) / 8 exactly, so unless its quotient by qt
    // lands within 1e-9 of a .5 tie (where the reference's float64 rounding errors decide) one pixel sum settles
    // it: the previous warp's last lane left it in shared memory; the block in front of the TILE belongs to
    // another CTA and is summed here.  A tie, or a tile predecessor that needs reflection padding, becomes an
    // exact-path item.
    int halo_item = 0, halo_dc = 0;
    const int hb = ti.blk0 + warp * 32 - 1;
    if (hb >= 0 && warp * 32 < ti.nb) {   // warp-uniform
        int sum = 0;
        bool have_sum = true;
        if (warp > 0) {
            sum = sm.blocksum[warp - 1];
        } else {
            int y0, x0;
            have_sum = tile_halo_is_fast(ti, y0, x0);
            if (have_sum) {
                sum = __dp4a(halo_v.x, 0x01010101u, __dp4a(halo_v.y, 0x01010101u, 0u));   // lanes 8..31 hold zeros
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                sum = __shfl_sync(0xffffffffu, sum, 0);
            }
        }
        halo_item = 1;
        if (have_sum) {
#if TIC_HALO_F32
            // FP32 is enough to tell "nowhere near a tie": |sum - 8192| <= 8192 is exact, the product is off by at
            // most |tq| * 2^-22 <= 3200 * 2^-22 = 7.7e-4 (the rounded reciprocal and the multiplication; quality 99 has
            // the smallest divisor, 8 * 0.32), and everything within 2e-3 of a tie goes to the exact path.  (The float64 form spent 3 % of the kernel's stall samples
            // waiting for the FP64 pipe, profiles/r2f.)
            const float tq = (float)(sum - 8192) * qp.dcinv_f;
            const float rq = rintf(tq);
            halo_dc = (int)rq;
            halo_item = (fabsf(tq - rq) < 0.5f - 2.0e-3f) ? 0 : 1;
#else
            const double tq = __dmul_rn((double)(sum - 8192), qp.dcinv);
            halo_dc = __double2int_rn(tq);
            halo_item = (fabs(tq - (double)halo_dc) < 0.5 - 1.0e-9) ? 0 : 1;   // on a tie the exact path decides
#endif
        }
    }
    // ---- accumulator -> quantised coefficients ------------------------------------------------------------
    // warp 0 waits for the MMA; the other warps sleep at the group barrier instead of polling
    if (TIC_POLL_WARP0) {
        if (warp == 0) { if (!tc::mbar_wait(tg.bar(g), tg.phase)) sm.tc_timeout = 1u; }
        group_sync<G>(g);
    } else {
        if (!tc::mbar_wait(tg.bar(g), tg.phase)) sm.tc_timeout = 1u;
    }
    tg.phase ^= 1u;
    tc::fence_after_sync();
    uint32_t fl_lo = 0, fl_hi = 0;
    if (warp * 32 < ti.nb) {   // warp-uniform: the warp owns at least one block
        uint32_t nz_lo = 0, nz_hi = 0;
        int dc = 0;
        uint32_t ra[8], rb[8];   // 8 columns = one zigzag group; the next group's load is in flight while this one is quantised
        const uint32_t tmem = tg.tmem(g);
        tc::tmem_ld8(tmem, ra);
        tc::tmem_wait_ld8(ra);
        tc::tmem_ld8(tmem + 8, rb);
        quantise_group_tc<0>(ra, qp, sm, t, debug_all, nz_lo, nz_hi, fl_lo, fl_hi, dc);
        tc::tmem_wait_ld8(rb);
        tc::tmem_ld8(tmem + 16, ra);
        quantise_group_tc<1>(rb, qp, sm, t, debug_all, nz_lo, nz_hi, fl_lo, fl_hi, dc);
        tc::tmem_wait_ld8(ra);
        tc::tmem_ld8(tmem + 24, rb);
        quantise_group_tc<2>(ra, qp, sm, t, debug_all, nz_lo, nz_hi, fl_lo, fl_hi, dc);
        tc::tmem_wait_ld8(rb);
        tc::tmem_ld8(tmem + 32, ra);
        quantise_group_tc<3>(rb, qp, sm, t, debug_all, nz_lo, nz_hi, fl_lo, fl_hi, dc);
        tc::tmem_wait_ld8(ra);
        tc::tmem_ld8(tmem + 40, rb);
        quantise_group_tc<4>(ra, qp, sm, t, debug_all, nz_lo, nz_hi, fl_lo, fl_hi, dc);
        tc::tmem_wait_ld8(rb);
        tc::tmem_ld8(tmem + 48, ra);
        quantise_group_tc<5>(rb, qp, sm, t, debug_all, nz_lo, nz_hi, fl_lo, fl_hi, dc);
        tc::tmem_wait_ld8(ra);
        tc::tmem_ld8(tmem + 56, rb);
        quantise_group_tc<6>(ra, qp, sm, t, debug_all, nz_lo, nz_hi, fl_lo, fl_hi, dc);
        tc::tmem_wait_ld8(rb);
        quantise_group_tc<7>(rb, qp, sm, t, debug_all, nz_lo, nz_hi, fl_lo, fl_hi, dc);
        sm.nz_lo[t] = nz_lo;
        sm.nz_hi[t] = nz_hi;
        sm.dcq[t] = dc;
        if (!(t < ti.nb)) fl_lo = fl_hi = 0;
        // Flagged coefficients at the four positions where exact .5 ties occur (9 in 10 of all flagged ones; every
        // block of a flat area whose level lands on a tie): settled by the lane itself from the accumulator's
        // column sums.  Everything else that is flagged goes to the warp's worklist below.
        constexpr uint32_t kRatLo = 0x80000000u | (0x80000000u >> 10) | (0x80000000u >> 14), kRatHi = 0x80000000u >> (39 - 32);
        if (!TIC_RATIONAL) {
            if (__popc(__ballot_sync(0xffffffffu, (fl_lo & 0x80000000u) != 0)) >= 4) {   // warp-uniform
                if (settle_dc_ties(ti.px, ti.w, ti.h, ti.bw, ti.blk0, ti.nb, qp.qt[0], &sm.coef[0][0], sm.dcq, fl_lo)) {
                    fl_lo &= 0x7fffffffu;
                    st.items(1u);
                }
            }
        } else if (!debug_all && __any_sync(0xffffffffu, ((fl_lo & kRatLo) | (fl_hi & kRatHi)) != 0)) {   // warp-uniform
            const uint32_t rat = (fl_lo >> 31) | (((fl_lo >> (31 - 10)) & 1u) << 1) | (((fl_lo >> (31 - 14)) & 1u) << 2) |
                                 (((fl_hi >> (31 - 7)) & 1u) << 3);
            const unsigned ch = settle_rational(tmem, rat, qp.qt, &sm, t);
            if (rat) {
                st.items((unsigned)__popc(rat));
                if (ch) st.changed(ch);
                fl_lo &= ~kRatLo;
                fl_hi &= ~kRatHi;
            }
        }
    }
    if (lane == 0) {
        if (halo_item) sm.work[warp][atomicAdd(&sm.work_count[warp], 1)] = kWorkHalo | kWorkGuard;   // slot 0: nothing pushed yet
        else sm.dc_halo[warp] = halo_dc;
    }
    __syncwarp();
    // ---- exact path: flagged coefficients of the warp's 32 blocks (all of them under TIC_FLAG_DEBUG_ALL_EXACT) ----
    uint32_t dbg_lo = 0, dbg_hi = 0;   // coefficients that go to the exact path although the guard did not flag them
    if (debug_all && t < ti.nb) { dbg_lo = ~fl_lo; dbg_hi = ~fl_hi; }
    while (true) {
        while (fl_lo | fl_hi | dbg_lo | dbg_hi) {
            const bool guard = (fl_lo | fl_hi) != 0;
            uint32_t& lo = guard ? fl_lo : dbg_lo;
            uint32_t& hi = guard ? fl_hi : dbg_hi;
            const int k = lo ? __clz(lo) : 32 + __clz(hi);
            const int slot = atomicAdd(&sm.work_count[warp], 1);
            if (slot >= kWarpWork) { atomicAdd(&sm.pending[warp], 1); break; }
            sm.work[warp][slot] = (guard ? kWorkGuard : 0u) | ((uint32_t)lane << 6) | (uint32_t)k;
            if (k < 32) lo ^= 0x80000000u >> k; else hi ^= 0x80000000u >> (k - 32);
        }
        __syncwarp();
        const int raw = sm.work_count[warp];
        const int count = raw < kWarpWork ? raw : kWarpWork;
        const int pending = sm.pending[warp];
        __syncwarp();
        if (count) {   // warp-uniform
            if (lane == 0) st.items((unsigned)count);
            exact_round(ti, qp, sm, warp, count, st);
        }
        if (pending == 0) break;
        if (lane == 0) { sm.work_count[warp] = 0; sm.pending[warp] = 0; }
        __syncwarp();
    }
    __syncwarp();
}

// DC of the block before thread t's block, after transform_warp (codec.py:34-35: plain difference
// over the whole image, no reset per block row; 0 in front of the first block).
__device__ __forceinline__ int dc_before(const TileShared& sm, int t) {
    return (t & 31) ? sm.dcq[t - 1] : sm.dc_halo[t >> 5];
}

// ---------------------------------------------------------------------------------------------
// phase 3: symbols -> bits.  One walk per block; the words go to the thread's private column
// (fast path) or, for the rare block longer than kPrivWords words, straight into the staging window.
// The loop is the hottest scalar code of the kernel: explicit 32-bit shared addresses, a sign-extending
// 16-bit load per coefficient, and a branch-free word flush.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_s16(uint32_t a) {
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
// Same load, out of line: used on error paths only, so that the compiler branches around them instead of
// predicating their instructions into the hot loop.
__device__ __noinline__ uint2 lds_v2_cold(uint32_t a) { return lds_v2(a); }
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, int s) {   // 0 for s >= 32 (PTX semantics)
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}

__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int msb_index(uint32_t v) {   // position of the highest set bit; -1 for 0
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

template <bool kToStage>
struct BitSink;

// fast path: thread-private column sm.priv[.][t]; row kPrivWords is a dump row for longer blocks.
// The pending bits sit right-aligned in a 64-bit accumulator; a word leaves when bit 5 of the bit
// count flips.
template <>
struct BitSink<false> {
    uint32_t lo = 0, hi = 0;
    int tot = 0;             // bits so far
    uint32_t ptr, ptr_end;   // shared addresses: next word, dump row
    __device__ __forceinline__ void put(uint32_t bits, int n) {   // 0 <= n <= 32, bits < 2^n
        hi = __funnelshift_lc(lo, hi, n);
        lo = shl_clamp(lo, n) | bits;
        const int nt = tot + n;
        const bool full = ((nt ^ tot) & 32) != 0;
        if (full) sts_u32(ptr, __funnelshift_r(lo, hi, nt));      // the 32 bits above the (nt & 31) left over
        ptr = full ? ptr + kTile * 4 : ptr;
        ptr = ptr < ptr_end ? ptr : ptr_end;
        tot = nt;
    }
    __device__ __forceinline__ int finish() {   // returns the block's bit count
        if (tot & 31) sts_u32(ptr, lo << (32 - (tot & 31)));
        return tot;
    }
};

// slow path: the block starts at window-relative bit (w0 * 32 + sh); w0 may be outside the window
template <>
struct BitSink<true> {
    uint32_t cur = 0;
    int nb = 0, cnt = 0;
    uint32_t* stage;
    int w0, sh;
    __device__ __forceinline__ void word(uint32_t w) {
        const int W = w0 + cnt;
        if ((unsigned)W < (unsigned)kWinWords) atomicOr(&stage[W], w >> sh);
        if (sh && (unsigned)(W + 1) < (unsigned)kWinWords) atomicOr(&stage[W + 1], w << (32 - sh));
        cnt++;
    }
    __device__ __forceinline__ void put(uint32_t bits, int n) {
        const uint32_t x = bits << ((32 - n) & 31);
        cur |= x >> nb;
        const uint32_t nxt = __funnelshift_r(0u, x, nb);
        nb += n;
        if (nb >= 32) {
            word(cur);
            cur = nxt;
            nb -= 32;
        }
    }
    __device__ __forceinline__ int finish() {
        const int bits = cnt * 32 + nb;
        if (nb > 0) word(cur);
        return bits;
    }
};

__device__ __forceinline__ uint32_t value_bits(int v, int sz) {   // huffman.py:59-63
    uint32_t mask;   // sz low bits set: one BMSK instead of a shifted -1
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(mask) : "r"(0), "r"(sz));
    return (uint32_t)(v + (v >> 31)) & mask;
}

template <bool kAuto, bool kToStage>
__device__ __forceinline__ void put_symbol(BitSink<kToStage>& s, uint2 e, int v, int sz) {
    if constexpr (kAuto) {   // e = {code, kHuffPresent | length}: code up to 32 bits + value up to 15
        s.put(e.x, (int)(e.y & kHuffLenMask));
        s.put(value_bits(v, sz), sz);
    } else {                 // e.x = (code length + size) << 27 | code << size;  length + size <= 26, code << size < 2^26
        s.put((e.x & 0x07ffffffu) | value_bits(v, sz), (int)(e.x >> 27));
    }
}
// one table entry: fixed tables 32 bits (returned in .x, .y = .x so that "absent" is .y == 0 in both forms)
template <bool kAuto>
__device__ __forceinline__ uint2 tab_entry(uint32_t tab, uint32_t idx) {
    if constexpr (kAuto) return lds_v2(tab + idx * 8u);
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(tab + idx * 4u) : "memory");
    return make_uint2(w, w);
}
template <bool kAuto>
__device__ __noinline__ uint2 tab_entry_cold(uint32_t tab, uint32_t idx) { return tab_entry<kAuto>(tab, idx); }
template <bool kAuto>
__device__ __forceinline__ int tab_len(uint2 e) { return kAuto ? (int)(e.y & kHuffLenMask) : (int)(e.x >> 27); }
template <bool kAuto>
__device__ __forceinline__ uint32_t tab_code(uint2 e) { return kAuto ? e.x : (e.x & 0x07ffffffu); }
template <bool kAuto>
__device__ __forceinline__ uint32_t tab_present(uint2 e) { return e.y; }   // 0: the symbol is not in the table

// Emits the bits of thread t's block (DC difference `diff`, AC from sm.coef / nz masks) and returns
// their number.  |quantised value| <= 1024 / 0.2 (quality 99), so a size never exceeds 14 and the
// table index stays inside the 16 x 16 table; sizes missing from the table have length word 0.
// sbase: shared address of `sm`; tabbase: shared address of the TabShared in force.
template <bool kAuto, bool kToStage>
__device__ __forceinline__ int walk_block(const TileShared& sm, uint32_t sbase, uint32_t tabbase, int t, int diff,
                                          BitSink<kToStage>& s, int& err) {
    const uint32_t dc_tab = tabbase + (uint32_t)offsetof(TabShared, dc_tab);
    const uint32_t ac_tab = tabbase + (uint32_t)offsetof(TabShared, ac_tab);
    const uint32_t col = sbase + (uint32_t)offsetof(TileShared, coef) + (uint32_t)t * 4u;
    int sz = msb_index((uint32_t)(diff < 0 ? -diff : diff)) + 1;  // bits_required, utils.py:9-10
    uint2 e = tab_entry<kAuto>(dc_tab, (uint32_t)sz);
    if (e.y == 0) { err = 1; sz = 0; e = tab_entry_cold<kAuto>(dc_tab, 0u); }  // KeyError, huffman.py:62
    put_symbol<kAuto, kToStage>(s, e, diff, sz);
    // AC coefficients.  The walk works on BIT indices of the non-zero masks (bit b of a mask word = zigzag coefficient
    // 31 - b resp. 63 - b = b ^ 31 resp. b ^ 63): the run is the distance to the previous symbol's bit, and the byte
    // offset of coefficient k in the thread's column, (k >> 1) * P + (k & 1) * 2 with P = kTile * 4, places the bits of
    // k one by one, so offset(b ^ c) = offset(b) ^ offset(c): one multiply, one AND-XOR, no subtraction.  A run of
    // 16 or more emits ZRL (huffman.py:25-29) as an iteration of its own — same code path, coefficient not consumed —
    // so that the loop body has no inner branch (the two convergence barriers around the ZRL loop and the
    // KeyError call cost 9 of 45 instructions per symbol, profiles/r2e).
    constexpr uint32_t kOffMul = kTile * 2u + 2u, kOffMask = 63u * kTile * 4u | 2u;
    int prev = 31;   // bit index of the last coded coefficient in the current word's numbering (the DC is bit 31 of word 0)
    uint32_t present = 0xffffffffu;   // AND-like minimum over the table entries used: 0 = a symbol outside the table
#pragma unroll 1
    for (int word = 0; word < 2; word++) {
        uint32_t m = word ? sm.nz_hi[t] : (sm.nz_lo[t] & 0x7fffffffu);   // k = 0 is the DC
        const uint32_t flip = word ? ((63u * kOffMul) & kOffMask) : ((31u * kOffMul) & kOffMask);
#if TIC_WALK_PIPE
        // software-pipelined: the next coefficient is fetched before this one is coded, so that one shared-memory
        // latency per symbol (the table entry) is exposed instead of two
        int b = msb_index(m) & 31;
        int v = lds_s16(col + ((((uint32_t)b * kOffMul) & kOffMask) ^ flip));
        while (m) {
            const int run = prev - b - 1;
            const bool zrl = run >= 16;
            uint32_t below;   // bits below b
            asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(below) : "r"(0), "r"(b));
            const uint32_t m2 = zrl ? m : (m & below);
            const int b2 = msb_index(m2) & 31;
            const int v2 = lds_s16(col + ((((uint32_t)b2 * kOffMul) & kOffMask) ^ flip));
            const int szv = msb_index((uint32_t)(v < 0 ? -v : v)) + 1;
            sz = zrl ? 0 : szv;
            e = tab_entry<kAuto>(ac_tab, (uint32_t)((zrl ? 15 : run) * 16 + sz));
            present = present < tab_present<kAuto>(e) ? present : tab_present<kAuto>(e);
            put_symbol<kAuto, kToStage>(s, e, v, sz);
            prev = zrl ? prev - 16 : b;
            m = m2; b = b2; v = v2;
        }
#else
        while (m) {
            const int b = msb_index(m);
            const int run = prev - b - 1;
            const bool zrl = run >= 16;
            int v = lds_s16(col + ((((uint32_t)b * kOffMul) & kOffMask) ^ flip));
            const int szv = msb_index((uint32_t)(v < 0 ? -v : v)) + 1;
            sz = zrl ? 0 : szv;
            e = tab_entry<kAuto>(ac_tab, (uint32_t)((zrl ? 15 : run) * 16 + sz));
            present = present < tab_present<kAuto>(e) ? present : tab_present<kAuto>(e);
            put_symbol<kAuto, kToStage>(s, e, v, sz);
            prev = zrl ? prev - 16 : b;
            uint32_t below;   // bits below b
            asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(below) : "r"(0), "r"(b));
            m = zrl ? m : (m & below);
        }
#endif
        prev += 32;
    }
    if (present == 0) err = 1;   // KeyError, huffman.py:62 (the stream is void; the status word says so)
    e = tab_entry<kAuto>(ac_tab, 0u);
    s.put(tab_code<kAuto>(e), tab_len<kAuto>(e));                // EOB always, huffman.py:33
    return s.finish();
}

// Symbol statistics of the warp's 32 blocks for the per-image tables (huffman.py:101-109, 187-194): counts and, for
// the dict-insertion order the reference's tree depends on, the first occurrence of every symbol as key = block
// index * 256 + ordinal of the symbol inside the block's list.  hist/first: 272 entries in shared memory, [0,256) AC
// (run*16+size), [256,272) DC size.  All 32 lanes call.  The lanes step through their symbol lists together; equal
// symbols of one step are combined (match.any) so that one lane adds their number and offers their smallest key —
// one thread per symbol with a shared-memory atomicAdd + 64-bit atomicMin each made this kernel four times as slow
// as the encode kernel (the few frequent symbols serialise).
// rel: (thread in tile) << 8 | ordinal — the key relative to the tile's first block; key0 = first block << 8
__device__ __forceinline__ void stats_step(int sym, unsigned rel, unsigned long long key0, uint32_t* hist, unsigned long long* first) {
    const unsigned act = __ballot_sync(0xffffffffu, sym >= 0);
    if (sym < 0) return;
    const unsigned peers = __match_any_sync(act, sym);
    if ((peers & (0u - peers)) == (1u << (threadIdx.x & 31))) {          // the lowest lane of the group: rel's high bits are
        atomicAdd(&hist[sym], (uint32_t)__popc(peers));                  // the thread index, so its key is the group's smallest
        const unsigned long long k = key0 + rel;
        if (k < first[sym]) atomicMin(&first[sym], k);                   // only until the symbol's first occurrence is settled
    }
}
// TIC_STATS_DIRECT: the AC symbols without the vote — every lane adds 1 to its symbol's counter (shared-memory
// atomic, no return value) and offers its key only while it is smaller than the one on record.
__device__ __forceinline__ void stats_step_direct(int sym, unsigned rel, unsigned long long key0, uint32_t* hist, unsigned long long* first) {
    if (sym < 0) return;
    atomicAdd(&hist[sym], 1u);
    const unsigned long long k = key0 + rel;
    if (k < *reinterpret_cast<volatile unsigned long long*>(&first[sym])) atomicMin(&first[sym], k);
}
__device__ __forceinline__ void warp_block_stats(const TileShared& sm, int t, bool active, unsigned long long blk0,
                                                 uint32_t* hist, unsigned long long* first, int& err) {
    const unsigned long long key0 = blk0 << 8;
    const unsigned rel0 = (unsigned)t << 8;
    // DC symbol: ordinal 0
    int s = 0;
    if (active) {
        s = bitlen(sm.dcq[t] - dc_before(sm, t));
        if (s > 15) { err = 1; s = 15; }   // int2ba(category, 4) overflows in the reference (codec.py:76)
    }
    stats_step(active ? 256 + s : -1, rel0, key0, hist, first);
    uint32_t lo = active ? sm.nz_lo[t] : 0u, hi = active ? sm.nz_hi[t] : 0u;
    int prev = 0, ord = 0;
    int zrl_left = 0;   // ZRL symbols still to be counted in front of the pending coefficient symbol
    int pend_sym = -1;
    while (__any_sync(0xffffffffu, (lo | hi) != 0u || zrl_left > 0 || pend_sym >= 0)) {
        if (zrl_left == 0 && pend_sym < 0 && (lo | hi)) {   // next non-zero coefficient of this lane's block
            int k;
            if (lo) { k = __clz(lo); lo ^= 0x80000000u >> k; } else { k = __clz(hi); hi ^= 0x80000000u >> k; k += 32; }
            const int run = k - prev - 1;
            prev = k;
            int sz = bitlen(coef_get(sm, t, k));
            if (sz > 15) { err = 1; sz = 15; }
            zrl_left = run >> 4;
            pend_sym = ((run & 15) << 4) | sz;
        }
        int sym = -1;
        const unsigned rel = rel0 | (unsigned)ord;
        if (zrl_left > 0) {   // every ZRL of a run has the same symbol; the first one carries the smallest ordinal
            sym = 0xF0;
            ord += 1;
            zrl_left -= 1;
        } else if (pend_sym >= 0) {
            sym = pend_sym;
            pend_sym = -1;
            ord += 1;
        }
#if TIC_STATS_DIRECT
        stats_step_direct(sym, rel, key0, hist, first);
#else
        stats_step(sym, rel, key0, hist, first);
#endif
    }
    stats_step(active ? 0 : -1, rel0 | (unsigned)ord, key0, hist, first);   // EOB (huffman.py:33)
}

// ---------------------------------------------------------------------------------------------
// the scan over tiles.  A tile moves the write position by its bit count; a tile that closes an image
// then rounds the position up to 128 bits (streams start 16-byte aligned).  Any run of tiles acts as
//     g(P) = closed ? round_up128(P + a) + b : P + a
// and these compose associatively.
// ---------------------------------------------------------------------------------------------
struct Span {
    long long a, b;
    int closed;
};
__device__ __forceinline__ Span span_of(uint32_t rec_bits) {
    Span s;
    s.a = (long long)(rec_bits & kRecBitsMask);
    s.b = 0;
    s.closed = (rec_bits & kRecClosing) ? 1 : 0;
    return s;
}
__device__ __forceinline__ Span span_then(const Span& x, const Span& y) {   // x first, then y
    Span r;
    if (!x.closed) { r.a = x.a + y.a; r.b = y.b; r.closed = y.closed; }
    else if (!y.closed) { r.a = x.a; r.b = x.b + y.a; r.closed = 1; }
    else { r.a = x.a; r.b = round_up128(x.b + y.a) + y.b; r.closed = 1; }
    return r;
}
__device__ __forceinline__ long long span_apply(long long p, const Span& s) {
    return s.closed ? round_up128(p + s.a) + s.b : p + s.a;
}
__device__ __forceinline__ Span span_shfl_up(const Span& s, int o) {
    Span r;
    r.a = __shfl_up_sync(0xffffffffu, s.a, o);
    r.b = __shfl_up_sync(0xffffffffu, s.b, o);
    r.closed = __shfl_up_sync(0xffffffffu, s.closed, o);
    return r;
}

}  // namespace tic
