// tic_kernels.cuh — sm_100a device code of the tinyimgcodec encode path.
//
// One CTA encodes one TILE: kTile consecutive 8x8 blocks of one image in the stream's
// block-raster order (tinyimgcodec/codec.py:34-36), one thread per block.  Everything the
// reference does between `image` and `bytes` happens inside that CTA:
//
//   load      8 x LDG.64 per thread; a warp reads 256 contiguous bytes per pixel row
//   transform level shift + 8x8 FDCT in registers, FP32 (AAN butterflies), scaled into the
//             quantiser: the fast path for utils.py:32-37,48-53
//   quantise  one FFMA per coefficient into 2^-F fixed point; a coefficient whose fraction
//             is within a guard band of a .5 rounding tie is sent to the EXACT path
//   exact     the reference's float64 FDCT (SciPy/ducc0 op order),
//             8 lanes per flagged coefficient: reproduces the reference's rounding of ties
//   symbols   zigzag (constants.py:23-34), DC difference (codec.py:34-35), run lengths
//             (huffman.py:12-33), Huffman code + value bits (huffman.py:41-63)
//   scan      bit lengths -> CTA scan -> decoupled look-back across tiles and images
//   pack      MSB-first bit packing (bitbuffer.py:17-40) into shared memory, then a
//             funnel-shifted, byte-swapped copy to the global stream; the word shared
//             with the previous tile arrives through a 64-bit mailbox
//
// Quantised coefficients never touch HBM: algorithmic traffic is pixels in + stream out.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tic_tables.h"

namespace tic {

constexpr int kTile = 128;                                  // blocks (= threads) per tile
constexpr int kWarps = kTile / 32;
// The bits of a tile are packed through a WINDOW of kWinWords 32-bit words of shared memory: a tile
// whose stream is longer (worst case 128 x 1662 bits + a table header) is emitted in several
// rounds.  Typical tiles (<= 320 bits per block on average) need one.
constexpr int kWinWords = 1280;
constexpr int kStageWords = kWinWords + 2;
constexpr int kWorkCap = 64;                                // exact-path worklist entries per round
constexpr int kExactPerRound = kTile / 8;                   // 8 lanes per worklist entry

// look-back status word: [63:62] flag, [61] closing, [60:0] value
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kFlagMask = 3ull << 62;
constexpr unsigned long long kClosingBit = 1ull << 61;
constexpr unsigned long long kValueMask = (1ull << 61) - 1;

// Quality-dependent constants, passed BY VALUE so that every entry is a constant-bank
// operand of the FFMA / compare that uses it.
struct QuantParams {
    float qmul[64];   // [u*8+v]  2^F / (8 * aan[u] * aan[v] * qt[u][v])
    float magic[64];  // [u*8+v]  1.5*2^23 + 2^(F-1) + 2^(k-1): rounding offset + tie-window offset
    int gmask[64];    // [u*8+v]  (2^F - 1) & ~(2^k - 1): inside the tie window iff (bits & gmask) == 0
    double qt[64];    // [u*8+v]  the reference's float64 divisor (utils.py:50-53)
    int fbits;        // F: fixed-point fraction bits of the fast quantiser
    int qbias;        // 0x4B400000 >> F: what (bits >> F) reads for a zero coefficient
    int pad[2];
};

struct ImageDesc {
    const uint8_t* px;
    int h, w;          // unpadded
    int bw;            // blocks per row = ceil(w/8)
    int nblk;          // ceil(h/8)*ceil(w/8)
    long long tile0;   // index of this image's first tile in the batch
};

// counters[] layout in the workspace
enum { kCtrTicket = 0, kCtrOverflow = 1, kCtrExactItems = 2, kCtrExactChanged = 3, kCtrTotalBits = 4,
       kCtrAnyStatus = 5, kCtrCount = 8 };

__device__ __constant__ uint8_t c_zigzag[64] = {TIC_ZIGZAG_LIST};
__device__ __constant__ HuffTables c_default_tables;

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int i, int n) {   // numpy "reflect", utils.py:56-61
    if (i < n) return i;
    if (n == 1) return 0;
    int p = 2 * (n - 1);
    int m = i % p;
    return m < n ? m : p - m;
}

__device__ __forceinline__ int bitlen(int v) {               // bits_required, utils.py:9-10
    return 32 - __clz(v < 0 ? -v : v);
}

// Look-back words are self-contained 64-bit messages (flag + value in one word), so relaxed
// gpu-scope accesses are enough: nothing else has to become visible with them.  (acquire loads
// compile to LDG + CCTL.IVALL, an L1 invalidate per poll.)
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ long long round_up128(long long v) { return (v + 127) & ~127ll; }

// ---------------------------------------------------------------------------------------------
// fast path: FP32 AAN 8-point DCT (5 multiplies, 29 adds; the output scale is folded into
// QuantParams::qmul).  Not bit-exact with the reference — the guard band + exact path is.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void aan8(float& d0, float& d1, float& d2, float& d3, float& d4, float& d5,
                                     float& d6, float& d7) {
    float t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6;
    float t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    float t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    d0 = t10 + t11;
    d4 = t10 - t11;
    float z1 = (t12 + t13) * 0.707106781f;
    d2 = t13 + z1;
    d6 = t13 - z1;
    t10 = t4 + t5;
    t11 = t5 + t6;
    t12 = t6 + t7;
    float z5 = (t10 - t12) * 0.382683433f;
    float z2 = 0.541196100f * t10 + z5;
    float z4 = 1.306562965f * t12 + z5;
    float z3 = t11 * 0.707106781f;
    float z11 = t7 + z3, z13 = t7 - z3;
    d5 = z13 + z2;
    d3 = z13 - z2;
    d1 = z11 + z4;
    d7 = z11 - z4;
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
    double c0 = __dmul_rn(2.0, x0), c7 = __dmul_rn(2.0, x7);
    double c1 = __dadd_rn(x1, x2), c2 = __dsub_rn(x2, x1);
    double c3 = __dadd_rn(x3, x4), c4 = __dsub_rn(x4, x3);
    double c5 = __dadd_rn(x5, x6), c6 = __dsub_rn(x6, x5);
    double h0 = __dadd_rn(c0, c7), h4 = __dsub_rn(c0, c7);
    double h3 = __dmul_rn(2.0, c3), h7 = __dmul_rn(-2.0, c4);
    double h1 = __dadd_rn(c1, c5), tr = __dsub_rn(c1, c5);
    double ti = __dadd_rn(c2, c6), h2 = __dsub_rn(c2, c6);
    double h6 = __dadd_rn(__dmul_rn(WR, ti), __dmul_rn(WI, tr));
    double h5 = __dsub_rn(__dmul_rn(WR, tr), __dmul_rn(WI, ti));
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
    double a = __dadd_rn(h0, h3), b = __dsub_rn(h0, h3);
    double e1 = __dmul_rn(2.0, h1), e2 = __dmul_rn(2.0, h2);
    double a2 = __dadd_rn(h4, h7), b2 = __dsub_rn(h4, h7);
    double e5 = __dmul_rn(2.0, h5), e6 = __dmul_rn(2.0, h6);
    (void)a; (void)e1;
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

struct TileInfo {
    const uint8_t* px;
    int h, w, bw;
    int img;          // image index
    int blk0;         // first block of the tile within the image
    int nb;           // blocks in this tile (0..kTile)
    bool first;       // first tile of its image (writes the header)
    bool closing;     // last tile of its image (pads to a byte, closes the stream)
};

// ---------------------------------------------------------------------------------------------
// shared memory of one tile
// ---------------------------------------------------------------------------------------------
struct TileShared {
    uint32_t coef[32][kTile];        // zigzag pairs (2i, 2i+1) packed lo/hi int16, one column per thread
    uint32_t nz_lo[kTile];           // bit k set: zigzag coefficient k != 0 (k = 1..31; bit 0 unused)
    uint32_t nz_hi[kTile];           // k = 32..63
    int dcq[kTile + 1];              // quantised DC: [0] = block before the tile, [t+1] = thread t
    uint32_t work[kWorkCap];         // exact-path worklist: thread << 6 | zigzag index; bit 31: halo DC
    double colres[kExactPerRound][8];
    uint2 ac_tab[256];               // {code, length} per (run << 4 | size); length 0 = not in table
    uint2 dc_tab[16];                // {code, length} per size category
    int work_count;
    int pending;                     // flagged coefficients that did not fit the worklist this round
    int warp_bits[kWarps];
    int err;
    long long s_bits;                // absolute bit position where this tile's block data starts
    unsigned int tail_prev;
    TileInfo ti;                     // the tile being encoded (written by warp 0)
    long long tile;                  // its index, or >= ntiles when the work is exhausted
    unsigned int carry;              // last word of the previous window (multi-round tiles)
    alignas(16) uint32_t stage[kStageWords];   // window of the tile-relative MSB-first bit buffer
};


// Warp-cooperative (all 32 lanes of one warp call): 32-ary search for the last image whose first
// tile is <= tile.  uniform_tpi > 0: every image has that many tiles (the common batch), no search.
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
    const ImageDesc d = descs[lo];
    TileInfo ti;
    ti.px = d.px; ti.h = d.h; ti.w = d.w; ti.bw = d.bw; ti.img = lo;
    const long long lt = tile - d.tile0;
    ti.blk0 = (int)(lt * kTile);
    const int rem = d.nblk - ti.blk0;
    ti.nb = rem < 0 ? 0 : (rem > kTile ? kTile : rem);
    ti.first = (lt == 0);
    ti.closing = (ti.blk0 + kTile >= d.nblk);
    return ti;
}

__device__ __forceinline__ double load_px_exact(const TileInfo& ti, int y, int x) {
    int yy = reflect_idx(y, ti.h), xx = reflect_idx(x, ti.w);
    return (double)((int)__ldg(ti.px + (size_t)yy * ti.w + xx) - 128);   // codec.py:29
}

// ---------------------------------------------------------------------------------------------
// phase 1: pixels -> quantised zigzag coefficients in shared memory (fast path) + worklist
// ---------------------------------------------------------------------------------------------
template <int K>
struct ZZ {  // compile-time zigzag -> raster
    static constexpr int tab[64] = {TIC_ZIGZAG_LIST};
    static constexpr int r = tab[K];
};

template <int I>
__device__ __forceinline__ void quantise_pairs(const float (&d)[64], const QuantParams& qp, TileShared& sm,
                                               int t, uint32_t& nz_lo, uint32_t& nz_hi, uint32_t& fl_lo,
                                               uint32_t& fl_hi) {
    if constexpr (I < 32) {
        constexpr int k0 = 2 * I, k1 = 2 * I + 1;
        constexpr int r0 = ZZ<k0>::r, r1 = ZZ<k1>::r;
        // bits = 0x4B400000 + round(t * 2^F) + 2^(F-1) + 2^(k-1)   (one FFMA, magic-number rounding)
        const int b0 = __float_as_int(fmaf(d[r0], qp.qmul[r0], qp.magic[r0]));
        const int b1 = __float_as_int(fmaf(d[r1], qp.qmul[r1], qp.magic[r1]));
        const int s0 = b0 >> qp.fbits, s1 = b1 >> qp.fbits;   // qbias + floor(t + 0.5) outside the window
        const bool f0 = (b0 & qp.gmask[r0]) == 0;             // within the tie window: exact path decides
        const bool f1 = (b1 & qp.gmask[r1]) == 0;
        if constexpr (k0 < 32) {
            if (k0 != 0 && s0 != qp.qbias) nz_lo |= 1u << k0;
            if (s1 != qp.qbias) nz_lo |= 1u << k1;
            if (f0) fl_lo |= 1u << k0;
            if (f1) fl_lo |= 1u << k1;
        } else {
            if (s0 != qp.qbias) nz_hi |= 1u << (k0 - 32);
            if (s1 != qp.qbias) nz_hi |= 1u << (k1 - 32);
            if (f0) fl_hi |= 1u << (k0 - 32);
            if (f1) fl_hi |= 1u << (k1 - 32);
        }
        if constexpr (I == 0) sm.dcq[t + 1] = s0 - qp.qbias;
        sm.coef[I][t] = __byte_perm((uint32_t)s0, (uint32_t)s1, 0x5410);   // biased int16 pair
        quantise_pairs<I + 1>(d, qp, sm, t, nz_lo, nz_hi, fl_lo, fl_hi);
    }
}

// Coefficients sit in shared memory as 16-bit values biased by (qbias & 0xffff).
__device__ __forceinline__ int coef_get(const TileShared& sm, int t, int k, int bias) {
    uint32_t w = sm.coef[k >> 1][t];
    return (int)(short)(((k & 1) ? (w >> 16) : w) - (uint32_t)bias);
}
__device__ __forceinline__ void transform_block(const TileInfo& ti, const QuantParams& qp, TileShared& sm,
                                                int t, uint32_t& fl_lo, uint32_t& fl_hi) {
    const int b = ti.blk0 + t;
    const int br = b / ti.bw, bc = b - br * ti.bw;
    const int y0 = br * 8, x0 = bc * 8;
    float d[64];
    const bool fast = ((ti.w & 7) == 0) && ((reinterpret_cast<uintptr_t>(ti.px) & 7) == 0) && (y0 + 8 <= ti.h);
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
#pragma unroll
    for (int c = 0; c < 8; c++)
        aan8(d[c], d[8 + c], d[16 + c], d[24 + c], d[32 + c], d[40 + c], d[48 + c], d[56 + c]);
#pragma unroll
    for (int r = 0; r < 8; r++)
        aan8(d[r * 8], d[r * 8 + 1], d[r * 8 + 2], d[r * 8 + 3], d[r * 8 + 4], d[r * 8 + 5], d[r * 8 + 6],
             d[r * 8 + 7]);
    d[0] -= 8192.0f;   // level shift (codec.py:29) only moves the DC term: 64 * 128, exact in FP32
    uint32_t nz_lo = 0, nz_hi = 0;
    quantise_pairs<0>(d, qp, sm, t, nz_lo, nz_hi, fl_lo, fl_hi);
    sm.nz_lo[t] = nz_lo;
    sm.nz_hi[t] = nz_hi;
}

// ---------------------------------------------------------------------------------------------
// phase 2: exact recomputation of every flagged coefficient (and of the DC of the block that
// precedes the tile, which the DC difference of codec.py:34-35 needs).  8 lanes per entry:
// lane c transforms column c of the block (axis -2 first, utils.py:33-34), lane 0 the row.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void exact_round(const TileInfo& ti, const QuantParams& qp, TileShared& sm,
                                            int count, unsigned long long* counters) {
    const int tid = threadIdx.x;
    const int grp = tid >> 3, c = tid & 7;
    for (int base = 0; base < count; base += kExactPerRound) {
        const int idx = base + grp;
        const bool act = idx < count;   // uniform per 8-lane group
        uint32_t item = act ? sm.work[idx] : 0;
        const bool halo = (item >> 31) != 0;
        const int owner = (int)((item >> 6) & 0xffff);
        const int k = (int)(item & 63);
        const int r = c_zigzag[k];
        const int u = r >> 3, v = r & 7;
        const int b = halo ? ti.blk0 - 1 : ti.blk0 + owner;
        if (act) {
            const int br = b / ti.bw, bc = b - br * ti.bw;
            const int y0 = br * 8, x = bc * 8 + c;
            double x0 = load_px_exact(ti, y0 + 0, x), x1 = load_px_exact(ti, y0 + 1, x);
            double x2 = load_px_exact(ti, y0 + 2, x), x3 = load_px_exact(ti, y0 + 3, x);
            double x4 = load_px_exact(ti, y0 + 4, x), x5 = load_px_exact(ti, y0 + 5, x);
            double x6 = load_px_exact(ti, y0 + 6, x), x7 = load_px_exact(ti, y0 + 7, x);
            sm.colres[grp][c] = dct8_exact(x0, x1, x2, x3, x4, x5, x6, x7, u);
        }
        __syncwarp();
        if (act && c == 0) {
            const double* cr = sm.colres[grp];
            double y = dct8_exact(cr[0], cr[1], cr[2], cr[3], cr[4], cr[5], cr[6], cr[7], v);
            int q = __double2int_rn(__ddiv_rn(y, qp.qt[r]));   // np.round(coeffs / qt), utils.py:53
            if (halo) {
                sm.dcq[0] = q;
            } else {
                int old = coef_get(sm, owner, k, qp.qbias);
                if (old != q) {
                    // two flagged coefficients of one block may share a packed word: serialise
                    // through a 32-bit CAS on that word
                    uint32_t* wp = &sm.coef[k >> 1][owner];
                    uint32_t seen = *wp, want;
                    do {
                        const uint32_t qb = (uint32_t)(q + qp.qbias) & 0xffffu;
                        want = (k & 1) ? ((seen & 0x0000ffffu) | (qb << 16)) : ((seen & 0xffff0000u) | qb);
                        uint32_t prev = atomicCAS(wp, seen, want);
                        if (prev == seen) break;
                        seen = prev;
                    } while (true);
                    if (k == 0) {
                        sm.dcq[owner + 1] = q;
                    } else if (k < 32) {
                        if (q) atomicOr(&sm.nz_lo[owner], 1u << k); else atomicAnd(&sm.nz_lo[owner], ~(1u << k));
                    } else {
                        if (q) atomicOr(&sm.nz_hi[owner], 1u << (k - 32)); else atomicAnd(&sm.nz_hi[owner], ~(1u << (k - 32)));
                    }
                    atomicAdd(&counters[kCtrExactChanged], 1ull);
                }
            }
        }
        __syncwarp();
    }
}

// Runs phases 1 and 2 for the tile.  On return (after a __syncthreads) sm.coef / nz / dcq hold
// the reference's quantised coefficients for every block of the tile.
__device__ __forceinline__ void transform_tile(const TileInfo& ti, const QuantParams& qp, TileShared& sm,
                                               unsigned long long* counters) {
    const int t = threadIdx.x;
    if (t == 0) {
        sm.work_count = 0;
        sm.pending = 0;
        sm.dcq[0] = 0;
        if (ti.blk0 > 0 && ti.nb > 0) {   // halo: DC of the previous block of the same image
            sm.work[0] = 0x80000000u;
            sm.work_count = 1;
        }
    }
    __syncthreads();
    uint32_t fl_lo = 0, fl_hi = 0;
    if (t < ti.nb) transform_block(ti, qp, sm, t, fl_lo, fl_hi);
    // push flagged coefficients; loop in rounds if the worklist overflows (very high quality only)
    while (true) {
        while (fl_lo | fl_hi) {
            int k = fl_lo ? (__ffs(fl_lo) - 1) : (31 + __ffs(fl_hi));
            int slot = atomicAdd(&sm.work_count, 1);
            if (slot >= kWorkCap) { atomicAdd(&sm.pending, 1); break; }
            sm.work[slot] = ((uint32_t)t << 6) | (uint32_t)k;
            if (k < 32) fl_lo &= fl_lo - 1; else fl_hi &= fl_hi - 1;
        }
        __syncthreads();
        int count = sm.work_count < kWorkCap ? sm.work_count : kWorkCap;
        int pending = sm.pending;
        if (t == 0 && count) atomicAdd(&counters[kCtrExactItems], (unsigned long long)count);
        exact_round(ti, qp, sm, count, counters);
        __syncthreads();
        if (pending == 0) break;
        if (t == 0) { sm.work_count = 0; sm.pending = 0; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// phase 3: symbols.  Bit length of one block, then its bits.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_bits(const TileShared& sm, int t, int bias, int& err) {
    int diff = sm.dcq[t + 1] - sm.dcq[t];                      // codec.py:34-35
    int s = bitlen(diff);
    if (s > 15 || sm.dc_tab[s].y == 0) { err = 1; s = 0; }     // KeyError, huffman.py:62
    int bits = (int)(sm.dc_tab[s].y & kHuffLenMask) + s;
    uint32_t lo = sm.nz_lo[t], hi = sm.nz_hi[t];
    int prev = 0;
    const int zrl = (int)(sm.ac_tab[0xF0].y & kHuffLenMask);
    while (lo | hi) {
        int k;
        if (lo) { k = __ffs(lo) - 1; lo &= lo - 1; } else { k = 31 + __ffs(hi); hi &= hi - 1; }
        int run = k - prev - 1;
        prev = k;
        int sz = bitlen(coef_get(sm, t, k, bias));
        int sym = ((run & 15) << 4) | sz;
        if (sz > 15 || sm.ac_tab[sym & 255].y == 0) { err = 1; sz = 1; sym = ((run & 15) << 4) | 1; }
        bits += (run >> 4) * zrl + (int)(sm.ac_tab[sym].y & kHuffLenMask) + sz; // huffman.py:25-29
    }
    return bits + (int)(sm.ac_tab[0].y & kHuffLenMask);        // EOB always, huffman.py:33
}

// Symbol statistics of one block for the per-image tables (huffman.py:101-109, 187-194): counts
// and, for the dict-insertion order the reference's tree depends on, the first occurrence of every
// symbol as key = block index * 256 + ordinal of the symbol inside the block's list.
// hist/first: 272 entries, [0,256) AC (run*16+size), [256,272) DC size.
__device__ __forceinline__ void block_stats(const TileShared& sm, int t, int bias, unsigned long long blk,
                                            uint32_t* hist, unsigned long long* first, int& err) {
    int diff = sm.dcq[t + 1] - sm.dcq[t];
    int s = bitlen(diff);
    if (s > 15) { err = 1; s = 15; }   // int2ba(category, 4) overflows in the reference (codec.py:76)
    atomicAdd(&hist[256 + s], 1u);
    atomicMin(&first[256 + s], blk << 8);
    uint32_t lo = sm.nz_lo[t], hi = sm.nz_hi[t];
    int prev = 0, ord = 0;
    while (lo | hi) {
        int k;
        if (lo) { k = __ffs(lo) - 1; lo &= lo - 1; } else { k = 31 + __ffs(hi); hi &= hi - 1; }
        int run = k - prev - 1;
        prev = k;
        int sz = bitlen(coef_get(sm, t, k, bias));
        if (sz > 15) { err = 1; sz = 15; }
        if (run >> 4) {
            atomicAdd(&hist[0xF0], (uint32_t)(run >> 4));
            atomicMin(&first[0xF0], (blk << 8) | (unsigned)ord);
            ord += run >> 4;
        }
        int sym = ((run & 15) << 4) | sz;
        atomicAdd(&hist[sym], 1u);
        atomicMin(&first[sym], (blk << 8) | (unsigned)ord);
        ord++;
    }
    atomicMin(&first[0], (blk << 8) | (unsigned)ord);   // EOB; its count is the number of blocks
}

struct BitSink {
    uint32_t* stage;          // window of the tile's bit buffer
    unsigned long long acc;   // left-aligned pending bits
    int nb;                   // number of pending bits (< 32 between calls)
    int word;                 // window-relative index of the word being filled (may be outside)
    bool first;
    __device__ __forceinline__ void put(uint32_t code, int len) {   // 0 <= len <= 32
        if (len == 0) return;
        acc |= (unsigned long long)code << (64 - nb - len);
        nb += len;
        if (nb >= 32) {
            uint32_t w = (uint32_t)(acc >> 32);
            if ((unsigned)word < (unsigned)kWinWords) {   // words outside the window belong to another round
                if (first) atomicOr(&stage[word], w); else stage[word] = w;
            }
            first = false;
            word++;
            acc <<= 32;
            nb -= 32;
        }
    }
    __device__ __forceinline__ void flush() {
        if (nb > 0 && (unsigned)word < (unsigned)kWinWords) atomicOr(&stage[word], (uint32_t)(acc >> 32));
    }
};

__device__ __forceinline__ uint32_t value_bits(int v, int sz) {   // huffman.py:59-63
    return (uint32_t)(v + (v >> 31)) & ((1u << sz) - 1u);
}

__device__ __forceinline__ void block_emit(TileShared& sm, int t, int bias, int bitpos, int wbase) {
    BitSink s;
    s.stage = sm.stage;
    s.acc = 0;
    s.nb = bitpos & 31;
    s.word = (bitpos >> 5) - wbase;
    s.first = true;
    int diff = sm.dcq[t + 1] - sm.dcq[t];
    int sz = bitlen(diff);
    if (sz > 15 || sm.dc_tab[sz].y == 0) sz = 0;
    uint2 e = sm.dc_tab[sz];
    e.y &= kHuffLenMask;
    if ((int)e.y + sz <= 32) {
        s.put((e.x << sz) | value_bits(diff, sz), (int)e.y + sz);
    } else {
        s.put(e.x, (int)e.y);
        s.put(value_bits(diff, sz), sz);
    }
    uint32_t lo = sm.nz_lo[t], hi = sm.nz_hi[t];
    int prev = 0;
    uint2 zrl = sm.ac_tab[0xF0];
    zrl.y &= kHuffLenMask;
    while (lo | hi) {
        int k;
        if (lo) { k = __ffs(lo) - 1; lo &= lo - 1; } else { k = 31 + __ffs(hi); hi &= hi - 1; }
        int run = k - prev - 1;
        prev = k;
        int v = coef_get(sm, t, k, bias);
        sz = bitlen(v);
        int sym = ((run & 15) << 4) | sz;
        if (sz > 15 || sm.ac_tab[sym & 255].y == 0) { sz = 1; sym = ((run & 15) << 4) | 1; v = 1; }
        for (int z = run >> 4; z > 0; z--) s.put(zrl.x, (int)zrl.y);
        e = sm.ac_tab[sym];
        e.y &= kHuffLenMask;
        if ((int)e.y + sz <= 32) {
            s.put((e.x << sz) | value_bits(v, sz), (int)e.y + sz);   // code + value bits in one go
        } else {
            s.put(e.x, (int)e.y);
            s.put(value_bits(v, sz), sz);
        }
    }
    s.put(sm.ac_tab[0].x, (int)(sm.ac_tab[0].y & kHuffLenMask));
    s.flush();
}

}  // namespace tic
