// tic_encode.cu — kernels and C ABI of libtinyimgcodec_cuda.so (see include/tinyimgcodec_cuda.h).
//
// Build (sm_100a only, no other architecture is supported):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -shared ...
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/tinyimgcodec_cuda.h"
#include "tic_kernels.cuh"

namespace tic {

// The second walk of a block longer than the private words, out of line (TIC_SLOW_WALK_CALL): the loop is a tenth of the
// kernel's code and runs for one block in thousands on ordinary content.
template <bool kAuto>
__device__ __noinline__ void walk_block_to_stage(TileShared* sm, uint32_t sbase, uint32_t tabbase, int t, int diff, int w0, int sh) {
    BitSink<true> s;
    s.stage = sm->stage; s.w0 = w0; s.sh = sh;
    int e2 = 0;
    walk_block<kAuto, true>(*sm, sbase, tabbase, t, diff, s, e2);
}

// Huffman tables -> shared memory in the form the walk wants (see TabShared).  All threads of ONE group call.
template <bool kAuto>
__device__ __forceinline__ void load_tables(TabShared& sm, const HuffTables& g) {
    uint32_t* ac32 = reinterpret_cast<uint32_t*>(sm.ac_tab);   // fixed tables: 32-bit entries (TabShared)
    uint32_t* dc32 = reinterpret_cast<uint32_t*>(sm.dc_tab);
    for (int i = tid(); i < 256; i += kTile) {
        const uint32_t len = g.ac[i].len, code = g.ac[i].code;
        if constexpr (kAuto) sm.ac_tab[i] = make_uint2(code, len);
        else ac32[i] = len ? (((len & kHuffLenMask) + (uint32_t)(i & 15)) << 27) | (code << (i & 15)) : 0u;
    }
    if (tid() < 16) {
        const int i = tid();
        const uint32_t len = g.dc[i].len, code = g.dc[i].code;
        if constexpr (kAuto) sm.dc_tab[i] = make_uint2(code, len);
        else dc32[i] = len ? (((len & kHuffLenMask) + (uint32_t)i) << 27) | (code << i) : 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// Shared memory of a CTA: [B operand | mbarriers + TMEM base] (tensor-core kernels only), then either one TileShared
// per group, each with its own tables behind it (per-image tables, C variant, single-group kernels), or — kShareTab —
// ONE TabShared for the CTA and the groups packed without theirs (fixed tables; the statistics kernel, which needs
// none, keeps the same layout).  tc_cta_setup: B -> shared memory, one mbarrier per group, one TMEM allocation per CTA.
// ---------------------------------------------------------------------------------------------
constexpr size_t kTcCtlBytes = 128;   // G mbarriers (8 bytes each, G <= 8) + the TMEM base address at byte 64
constexpr size_t kGroupStride = (sizeof(TileShared) + 127) & ~(size_t)127;
constexpr size_t kGroupStrideNoTab = (offsetof(TileShared, tab) + 127) & ~(size_t)127;
constexpr size_t kTabBytes = (sizeof(TabShared) + 127) & ~(size_t)127;
template <int G, bool kTc, bool kShareTab = false>
constexpr size_t cta_smem_bytes() {
    return (kTc ? (size_t)tc::kBBytes + kTcCtlBytes : 0) + (kShareTab ? kTabBytes + (size_t)G * kGroupStrideNoTab : (size_t)G * kGroupStride);
}
template <int G>
struct TmemCols {   // TMEM allocations are powers of two >= 32 columns
    static constexpr uint32_t value = G * tc::kColsPerGroup <= 64 ? 64u : (G * tc::kColsPerGroup <= 128 ? 128u : (G * tc::kColsPerGroup <= 256 ? 256u : 512u));
};
static_assert(kGroups >= 1 && kGroups <= 8 && kGroups * tc::kColsPerGroup <= 512, "TMEM has 512 columns");
static_assert(kGroupsAuto >= 1 && kGroupsAuto <= 8 && kGroupsAuto * tc::kColsPerGroup <= 512, "TMEM has 512 columns");

template <int G>
__device__ __forceinline__ unsigned char* tc_cta_setup(unsigned char* smem_raw, const uint4* __restrict__ bmat, int g,
                                                       TcGroup& tg, uint32_t& tmem_base) {
    uint4* sB = reinterpret_cast<uint4*>(smem_raw);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem_raw + tc::kBBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + tc::kBBytes + 64);
    for (int i = threadIdx.x; i < tc::kBBytes / 16; i += kTile * G) sB[i] = __ldg(bmat + i);
    if (threadIdx.x == 0) {
        for (int i = 0; i < G; i++) tc::mbar_init(tc::smem_addr(mbar + i), 1);
        tc::fence_mbar_init();
    }
    if (threadIdx.x < 32) tc::tmem_alloc(tc::smem_addr(tmem_slot), TmemCols<G>::value);
    tc::fence_proxy_async();   // the B operand is read through the async proxy
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    tmem_base = *tmem_slot;
    tg.ctl = tc::smem_addr(mbar);
    tg.phase = 0;
    (void)g;
    return smem_raw + tc::kBBytes + kTcCtlBytes;
}
template <int G>
__device__ __forceinline__ void tc_cta_teardown(uint32_t tmem_base) {
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base, TmemCols<G>::value);
}

// ---------------------------------------------------------------------------------------------
// compress(), stage 1: every tile -> its bits in the arena + a TileRec.  Persistent CTAs of G groups, tiles
// dealt round-robin to the groups; no tile waits for another.
// ---------------------------------------------------------------------------------------------
// kMode: 0 = fixed Huffman tables, 1 = per-image tables (auto_generate_huffman_table), 2 = C-variant stream
// (flag bit 30; `quality` is then the reference's IMG_Q_BEST .. IMG_Q_LOW = 0 .. 3)
template <int kMode, int G>
__global__ void __launch_bounds__(kTile * G, G == 1 ? (kMode == 2 || !kFdctTc ? kCtasPerSm : 4) : 1)
encode_tiles_kernel(const __grid_constant__ QuantParams qp, const ImageDesc* __restrict__ descs, int n_images,
                    int uniform_tpi, long long ntiles, TileRec* __restrict__ recs, uint4* __restrict__ arena,
                    unsigned long long arena_cap16, unsigned long long* __restrict__ counters,
                    int* __restrict__ status, int quality, const AutoTables* __restrict__ auto_tabs,
                    const uint4* __restrict__ bmat, uint32_t flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool kAuto = kMode == 1, kCVar = kMode == 2, kTc = kFdctTc && !kCVar;
    constexpr bool kShareTab = kTc && !kAuto;   // fixed tables: one copy per CTA in front of the groups
    const int t = tid(), g = (int)threadIdx.x / kTile;
    const int lane = t & 31, warp = t >> 5;
    unsigned char* gbase = smem_raw;
    TcGroup tg{};
    uint32_t tmem_base = 0;
    if constexpr (kTc) gbase = tc_cta_setup<G>(smem_raw, bmat, g, tg, tmem_base);
    TileShared& sm = *reinterpret_cast<TileShared*>(gbase + (kShareTab ? kTabBytes + (size_t)g * kGroupStrideNoTab : (size_t)g * kGroupStride));
    TabShared& tabs = kShareTab ? *reinterpret_cast<TabShared*>(gbase) : sm.tab;
    uint32_t sbase = smem_u32(&sm);   // kept in a register: the walk addresses shared memory directly
    asm volatile("mov.u32 %0, %0;" : "+r"(sbase));
    const uint32_t tabbase = kShareTab ? smem_u32(gbase) : sbase + (uint32_t)offsetof(TileShared, tab);
    const bool debug_all = (flags & TIC_FLAG_DEBUG_ALL_EXACT) != 0;

    if constexpr (!kAuto) {   // constants.py:53-242
        if (!kShareTab || g == 0) load_tables<false>(tabs, c_default_tables);
        if constexpr (kShareTab && G > 1) __syncthreads();   // group 0's copy serves every group
    }
    for (int i = t; i < kWinWords; i += kTile) sm.stage[i] = 0;
    group_sync<G>(g);

    const ExactStats st{&sm};
    if (t == 0) sm.stat_items = sm.stat_changed = sm.stat_unflagged = sm.tc_timeout = 0u;
    int tab_img = -1;   // auto mode: image whose tables are in shared memory
    // tiles of this group: first, first + stride, ... (ntiles < 2^31: unsigned arithmetic cannot wrap)
    const unsigned first = blockIdx.x * G + g, stride = gridDim.x * G, ntiles_u = (unsigned)ntiles;
    // The description of tile `tl` into sm.tinfo[slot]: all lanes of warp 0 call (the search of a ragged batch is
    // warp-cooperative); uniform batches step (image, tile in image) by a fixed amount, no division in the loop.
    auto describe = [&](unsigned tl, int slot) {
        if (uniform_tpi > 0) {
            if (lane == 0) {
                int img = sm.u_img, lt = sm.u_lt;
                sm.tinfo[slot] = tile_info(descs[img], img, lt);
                const int dq = (int)(stride / (unsigned)uniform_tpi), dr = (int)(stride - (unsigned)dq * (unsigned)uniform_tpi);
                img += dq; lt += dr;
                if (lt >= uniform_tpi) { lt -= uniform_tpi; img++; }
                sm.u_img = img; sm.u_lt = lt;
            }
        } else {
            const TileInfo r = locate_tile(descs, n_images, (long long)tl, 0);
            if (lane == 0) sm.tinfo[slot] = r;
        }
    };
    if (warp == 0 && first < ntiles_u) {
        if (lane == 0 && uniform_tpi > 0) { sm.u_img = (int)(first / (unsigned)uniform_tpi); sm.u_lt = (int)(first - (unsigned)sm.u_img * (unsigned)uniform_tpi); }
        describe(first, 0);
    }
    group_sync<G>(g);

    constexpr int kDescWarp = (kTc && kWarps > 1) ? 1 : 0;   // warp 0 already issues the MMA and owns the tile's halo
    // (Requesting a tile's pixel rows one step earlier — behind barrier B1 of the tile before it, or in front of B2 — and
    // holding them in registers across the placement and the copy-out was measured: 4.47 / 4.63 ms against 4.10, and
    // 4.50 against 4.29 at 7 groups without a spill; commit 5d47ac4)
    // one loop register for two values: bits 0-30 the tile, bit 31 the slot of its description in sm.tinfo (the tile
    // loop is where the kernel's register pressure peaks; ntiles < 2^31)
    for (unsigned tcur = first; (tcur & 0x7fffffffu) < ntiles_u; tcur = (tcur + stride) ^ 0x80000000u) {
        const unsigned tile = tcur & 0x7fffffffu;
        const int slot = (int)(tcur >> 31);
        const TileInfo& ti = sm.tinfo[slot];
        // one warp describes the group's next tile one tile ahead: in the tensor core's shadow (tensor-core path), or
        // here; barrier B1 of this tile orders the write before every later read
        auto describe_next = [&]() {
            if (warp == kDescWarp && tile + stride < ntiles_u) {
                describe(tile + stride, slot ^ 1);
                if constexpr (TIC_PREFETCH_L2 != 0 && kTc) {
                    __syncwarp();   // lane 0 wrote the description
                    prefetch_tile_l2(sm.tinfo[slot ^ 1], lane);
                }
            }
        };
        if constexpr (kAuto) {
            if (tab_img != ti.img) {   // per-image tables (codec.py:146-148); group-uniform branch
                group_sync<G>(g);      // everyone is done with the previous image's tables
                load_tables<true>(tabs, auto_tabs[ti.img].tab);
                tab_img = ti.img;
                group_sync<G>(g);
            }
        }

        // ---- coefficients of the tile, then the bits of every block into its private words ---------
        if constexpr (kCVar) { describe_next(); transform_warp_c(ti, quality, sm); }
        else if constexpr (kTc) {
            if (ti.nb > 0) {
                uint2 rows[8];
                load_block_rows(ti, t, rows);
                const uint2 halo_v = load_tile_halo(ti, t);   // with the tile's own rows: one DRAM latency, not two
                transform_tile_tc<G>(ti, qp, sm, tg, g, debug_all, st, rows, halo_v, describe_next);
            } else {
                describe_next();
            }
        }
        else { describe_next(); transform_warp(ti, qp, sm, st); }
        int err = 0, bits = 0, nwords = 0, diff = 0;
        if (const int tw = tid_now(); tw < ti.nb) {
            diff = sm.dcq[tw] - dc_before(sm, tw);                    // codec.py:34-35
            BitSink<false> s;
            s.ptr = sbase + (uint32_t)offsetof(TileShared, priv) + (uint32_t)tw * 4u;
            s.ptr_end = s.ptr + (uint32_t)kPrivWords * kTile * 4u;
            bits = walk_block<kAuto, false>(sm, sbase, tabbase, tw, diff, s, err);
            nwords = (bits + 31) >> 5;
        }
        int incl = bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const bool warp_err = __any_sync(0xffffffffu, err != 0);
        if (lane == 31) { sm.warp_bits[warp] = incl; sm.warp_err[warp] = warp_err ? 1 : 0; }
        group_sync<G>(g);   // B1: warp totals visible; the staging window is zero (previous copy-out done)

        int warp_base = 0, tile_bits = 0, tile_err = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const int wb = sm.warp_bits[w];
            if (w < warp) warp_base += wb;
            tile_bits += wb;
            tile_err |= sm.warp_err[w];
        }
        // the first tile of an image carries the header in front of its blocks: 128 bits with the fixed
        // tables (codec.py:102-114), 128 + the serialised tables in auto mode (codec.py:110-112)
        const int hdr_bits = !ti.first ? 0 : (kAuto ? (int)auto_tabs[ti.img].hdr_bits : 128);
        const int bitpos = hdr_bits + warp_base + incl - bits;   // tile-relative bit offset of this block
        tile_bits += hdr_bits;
        // C variant: IMG_encodeComplete always writes one more byte (c/img.h BB_flushBits) = one pad bit here
        if constexpr (kCVar) tile_bits += ti.closing ? 1 : 0;
        const int nw_tile = (tile_bits + 31) >> 5;               // tile-relative words holding data
        const int n16_tile = (nw_tile + 3) >> 2;
        if (t == 0) {   // reserve the tile's slot in the arena; the latency hides behind the placement
            const unsigned long long off = atomicAdd(&counters[kCtrArena], (unsigned long long)n16_tile);
            const bool ok = off + (unsigned long long)n16_tile <= arena_cap16;
            if (!ok) atomicExch(&counters[kCtrOverflow], 1ull);
            sm.arena_off = ok ? (unsigned int)off : 0xffffffffu;
            TileRec r;
            r.bits = (uint32_t)tile_bits | (ti.first ? kRecFirst : 0u) | (ti.closing ? kRecClosing : 0u);
            r.off16 = ok ? (unsigned int)off : 0u;
            r.img = ti.img;
            r.pad = ok ? 0u : 1u;
            *reinterpret_cast<uint4*>(&recs[tile]) = *reinterpret_cast<uint4*>(&r);
            if (tile_err) atomicOr(&status[ti.img], TIC_STATUS_CATEGORY);
        }

        const int hdr_words = (hdr_bits + 31) >> 5;
        for (int wbase = 0;;) {   // one round per window; a second one only for tiles longer than the window
            // ---- header words (they OR into the zeroed window like everything else) --------------
            if (ti.first) {
                for (int i = wbase + t; i < hdr_words && i < wbase + kWinWords; i += kTile) {
                    uint32_t w;
                    if constexpr (kAuto) {
                        w = auto_tabs[ti.img].hdr_words[i];
                    } else {   // struct.pack("III") is little-endian, the stream is MSB-first; flag word 0
                        // C variant: all four words are little-endian structs, flag = 1 << 30 (c/img.c:183-192)
                        const uint32_t v = i == 0 ? (uint32_t)ti.h : (i == 1 ? (uint32_t)ti.w : (i == 2 ? (uint32_t)quality : (kCVar ? 0x40000000u : 0u)));
                        w = __byte_perm(v, 0, 0x0123);
                    }
                    if (w) atomicOr(&sm.stage[i - wbase], w);
                }
            }
            // ---- private words -> window, shifted to the block's bit offset ------------------------
            if (t < ti.nb && bitpos + bits > wbase * 32 && bitpos < (wbase + kWinWords) * 32) {
                const int sh = bitpos & 31;
                const int w0 = (bitpos >> 5) - wbase;
                if (nwords <= kPrivWords) {
                    const int nout = (sh + bits + 31) >> 5;
                    uint32_t prev = 0;
                    for (int j = 0; j < nout; j++) {
                        const uint32_t x = j < nwords ? sm.priv[j][t] : 0u;
                        const uint32_t v = __funnelshift_r(x, prev, sh);   // (prev << (32 - sh)) | (x >> sh)
                        prev = x;
                        const int W = w0 + j;
                        if ((unsigned)W < (unsigned)kWinWords) {
                            // a word entirely inside this block needs no atomic
                            if (j > 0 && (j + 1) * 32 - sh <= bits) sm.stage[W] = v; else atomicOr(&sm.stage[W], v);
                        }
                    }
                } else {   // long block: walk again, straight into the window
#if TIC_SLOW_WALK_CALL
                    walk_block_to_stage<kAuto>(&sm, sbase, tabbase, t, diff, w0, sh);
#else
                    BitSink<true> s;
                    s.stage = sm.stage; s.w0 = w0; s.sh = sh;
                    int e2 = 0;
                    walk_block<kAuto, true>(sm, sbase, tabbase, t, diff, s, e2);
#endif
                }
            }
            group_sync<G>(g);   // B2: window complete, arena offset visible

            // ---- window -> arena (16-byte stores), and the window is zero again -----------------
            const unsigned int aoff = sm.arena_off;
            const int i_end = (n16_tile - wbase / 4) < kWinWords / 4 ? (n16_tile - wbase / 4) : kWinWords / 4;
            uint4* st4 = reinterpret_cast<uint4*>(sm.stage);
            for (int i = t; i < i_end; i += kTile) {
                const uint4 v = st4[i];
                st4[i] = make_uint4(0u, 0u, 0u, 0u);
                if (aoff != 0xffffffffu) arena[(size_t)aoff + (size_t)(wbase / 4) + i] = v;
            }
            wbase += kWinWords;
            if (wbase >= nw_tile) break;
            group_sync<G>(g);
        }
    }
    if constexpr (kTc) tc_cta_teardown<G>(tmem_base);
    // exact-path statistics of the group: one thread, at the very end
    group_sync<G>(g);
    if (t == 0) {
        if (sm.stat_items) atomicAdd(&counters[kCtrExactItems], (unsigned long long)sm.stat_items);
        if (sm.stat_changed) atomicAdd(&counters[kCtrExactChanged], (unsigned long long)sm.stat_changed);
        if (sm.stat_unflagged) atomicAdd(&counters[kCtrUnflagged], (unsigned long long)sm.stat_unflagged);
        if (sm.tc_timeout) atomicExch(&counters[kCtrTcTimeout], 1ull);
    }
}

// ---------------------------------------------------------------------------------------------
// compress(), stage 2: absolute bit position of every tile.  Chunks of kScanChunk tiles:
//   scan_chunks_kernel   what each chunk does to the position (a Span); its last CTA: position at the start of every chunk
//   scan_apply_kernel    position of every tile; stream offset / end of every image; its last CTA: sizes and the
//                        batch summary (what used to be two more launches: scan_spine_kernel, finalize_kernel)
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256, kScanItems = 8, kScanChunk = kScanThreads * kScanItems;

// What a batch leaves for tic_encode_finish: `sticky` accumulates over every batch enqueued since the last finish
// (errors are ORed, so a later batch cannot hide an earlier batch's overflow; the byte total is the last batch's).
// One thread, after every kernel that writes `counters` has completed (stream order) or behind a CTA barrier.
__device__ __forceinline__ void fold_counters(unsigned long long* counters, unsigned long long* sticky) {
    const unsigned long long overflow = atomicAdd(&counters[kCtrOverflow], 0ull), timeout = atomicAdd(&counters[kCtrTcTimeout], 0ull);
    if (overflow) atomicOr(&sticky[kCtrOverflow], 1ull);
    if (timeout) atomicOr(&sticky[kCtrTcTimeout], 1ull);
    atomicExch(&sticky[kCtrExactItems], atomicAdd(&counters[kCtrExactItems], 0ull));     // statistics: the last batch's
    atomicExch(&sticky[kCtrExactChanged], atomicAdd(&counters[kCtrExactChanged], 0ull));
    atomicAdd(&sticky[kCtrUnflagged], atomicAdd(&counters[kCtrUnflagged], 0ull));        // guard misses: never lose one
    atomicExch(&sticky[kCtrTotalBits], atomicAdd(&counters[kCtrTotalBits], 0ull));
}



// Exclusive scan of one Span per thread over the CTA (in thread order); `total` = all of them.
template <int kThreads>
__device__ __forceinline__ Span cta_exclusive_span(const Span& mine, Span* warp_tot /* smem [kThreads/32] */, Span& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Span incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Span n = span_shfl_up(incl, o);
        if (lane >= o) incl = span_then(n, incl);
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    Span before{0, 0, 0};
    total = Span{0, 0, 0};
    for (int w = 0; w < kThreads / 32; w++) {
        if (w == warp) before = total;
        total = span_then(total, warp_tot[w]);
    }
    Span up = span_shfl_up(incl, 1);
    if (lane == 0) up = Span{0, 0, 0};
    __syncthreads();   // warp_tot may be reused by the caller
    return span_then(before, up);
}

__device__ __forceinline__ Span span_ldcg(const Span* p) {   // written by another CTA of the same launch: read at L2
    Span r;
    r.a = __ldcg(&p->a); r.b = __ldcg(&p->b); r.closed = __ldcg(&p->closed);
    return r;
}
// True in exactly one CTA of the launch: the last one to get here.  Everything the other CTAs wrote to global memory
// before their call is visible to it (fence + ticket), provided it reads with __ldcg / atomics.
__device__ __forceinline__ bool last_cta_done(unsigned long long* ticket) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1ull;
    __syncthreads();
    return is_last != 0;
}

// What each chunk of kScanChunk tiles does to the position; the last CTA to finish then scans the chunk spans
// (the "spine": position at the start of every chunk) — one launch instead of two.
__global__ void __launch_bounds__(kScanThreads)
scan_chunks_kernel(const TileRec* __restrict__ recs, long long ntiles, Span* __restrict__ chunk_span, long long nchunks,
                   long long* __restrict__ chunk_pos, unsigned long long* __restrict__ counters) {
    __shared__ Span warp_tot[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanItems;
    Span mine{0, 0, 0};
#pragma unroll
    for (int i = 0; i < kScanItems; i++)
        if (base + i < ntiles) mine = span_then(mine, span_of(recs[base + i].bits));
    Span total;
    cta_exclusive_span<kScanThreads>(mine, warp_tot, total);
    if (threadIdx.x == 0) chunk_span[blockIdx.x] = total;
    if (!last_cta_done(&counters[kCtrTicketA])) return;
    const long long per = (nchunks + kScanThreads - 1) / kScanThreads;
    const long long lo = (long long)threadIdx.x * per;
    const long long hi = lo + per < nchunks ? lo + per : nchunks;
    mine = Span{0, 0, 0};
    for (long long c = lo; c < hi; c++) mine = span_then(mine, span_ldcg(&chunk_span[c]));
    const Span before = cta_exclusive_span<kScanThreads>(mine, warp_tot, total);
    long long p = span_apply(0, before);
    for (long long c = lo; c < hi; c++) {
        chunk_pos[c] = p;
        p = span_apply(p, span_ldcg(&chunk_span[c]));
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const TileRec* __restrict__ recs, long long ntiles, const long long* __restrict__ chunk_pos,
                  long long* __restrict__ tile_pos, long long out_cap, long long* __restrict__ out_off,
                  long long* __restrict__ out_end, unsigned long long* __restrict__ counters, int n_images,
                  long long* __restrict__ out_sizes, const int* __restrict__ status, unsigned long long* __restrict__ sticky) {
    __shared__ Span warp_tot[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanItems;
    uint32_t rb[kScanItems];
    int img[kScanItems];
    Span mine{0, 0, 0};
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        rb[i] = 0; img[i] = 0;
        if (base + i < ntiles) {
            const TileRec r = recs[base + i];
            rb[i] = r.bits; img[i] = r.img;
            mine = span_then(mine, span_of(r.bits));
        }
    }
    Span total;
    const Span before = cta_exclusive_span<kScanThreads>(mine, warp_tot, total);
    long long p = span_apply(chunk_pos[blockIdx.x], before);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i >= ntiles) break;
        tile_pos[base + i] = p;
        const long long eb = p + (long long)(rb[i] & kRecBitsMask);   // end of this tile's data bits
        if (rb[i] & kRecFirst) out_off[img[i]] = p >> 3;
        if (rb[i] & kRecClosing) {
            const long long end_byte = (eb + 7) >> 3;
            out_end[img[i]] = end_byte;
            if (((end_byte + 3) & ~3ll) > out_cap) atomicExch(&counters[kCtrOverflow], 1ull);
            atomicMax(&counters[kCtrTotalBits], (unsigned long long)(end_byte << 3));
        }
        p = span_apply(p, span_of(rb[i]));
    }
    // the last CTA to finish: sizes = end - offset, and the batch summary tic_encode_finish reads
    if (!last_cta_done(&counters[kCtrTicketB])) return;
    for (int i = threadIdx.x; i < n_images; i += kScanThreads) {
        out_sizes[i] = __ldcg(&out_end[i]) - __ldcg(&out_off[i]);
        const int st = status[i];
        if (st) atomicOr(&sticky[kCtrAnyStatus], (unsigned long long)st);
    }
    if (threadIdx.x == 0) fold_counters(counters, sticky);
}

// Small batches (at most kSmallScan tiles: one 8K image, 50 x 512^2, ...): the three scan kernels and
// the size / summary pass in ONE launch of one CTA — what counts there is launch latency, not throughput.
constexpr int kSmallThreads = 1024, kSmallScan = kSmallThreads * kScanItems;
__global__ void __launch_bounds__(kSmallThreads)
scan_small_kernel(const TileRec* __restrict__ recs, int ntiles, long long* __restrict__ tile_pos, long long out_cap,
                  long long* __restrict__ out_off, long long* __restrict__ out_end, long long* __restrict__ out_sizes,
                  int n_images, const int* __restrict__ status, unsigned long long* __restrict__ counters,
                  unsigned long long* __restrict__ sticky) {
    __shared__ Span warp_tot[kSmallThreads / 32];
    const int base = (int)threadIdx.x * kScanItems;
    uint32_t rb[kScanItems];
    int img[kScanItems];
    Span mine{0, 0, 0};
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        rb[i] = 0; img[i] = 0;
        if (base + i < ntiles) {
            const TileRec r = recs[base + i];
            rb[i] = r.bits; img[i] = r.img;
            mine = span_then(mine, span_of(r.bits));
        }
    }
    Span total;
    const Span before = cta_exclusive_span<kSmallThreads>(mine, warp_tot, total);
    long long p = span_apply(0, before);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i >= ntiles) break;
        tile_pos[base + i] = p;
        const long long eb = p + (long long)(rb[i] & kRecBitsMask);
        if (rb[i] & kRecFirst) out_off[img[i]] = p >> 3;
        if (rb[i] & kRecClosing) {
            const long long end_byte = (eb + 7) >> 3;
            out_end[img[i]] = end_byte;
            if (((end_byte + 3) & ~3ll) > out_cap) atomicExch(&counters[kCtrOverflow], 1ull);
            atomicMax(&counters[kCtrTotalBits], (unsigned long long)(end_byte << 3));
        }
        p = span_apply(p, span_of(rb[i]));
    }
    __syncthreads();   // out_off / out_end of every image are in place (same CTA: visible after the barrier)
    for (int i = threadIdx.x; i < n_images; i += kSmallThreads) {
        out_sizes[i] = out_end[i] - out_off[i];
        if (status[i]) atomicOr(&sticky[kCtrAnyStatus], (unsigned long long)status[i]);
    }
    if (threadIdx.x == 0) fold_counters(counters, sticky);
}

// ---------------------------------------------------------------------------------------------
// compress(), stage 3: arena -> dense output.  One warp per tile; every 32-bit output word is written
// by exactly one lane: the tile that holds the word's LAST bit owns it (a closing tile also owns the
// final partial word) and gathers the leading bits from the tiles before it.  Funnel shift to the
// final bit position, byte swap to the stream's MSB-first byte order (bitbuffer.py:36-40).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tile_bits_at(const uint32_t* __restrict__ src, int nw, long long rb) {
    // 32 bits of the tile starting at its bit rb >= 0, left-aligned; words past the tile read as 0
    const int i = (int)(rb >> 5), s = (int)(rb & 31);
    const uint32_t hi = i < nw ? __ldg(src + i) : 0u;
    const uint32_t lo = (s && i + 1 < nw) ? __ldg(src + i + 1) : 0u;
    return __funnelshift_l(lo, hi, s);
}

constexpr int kCompactThreads = 256;
// Output word j of a tile (j = 0 is the word that holds the tile's first bit, `off` bits into it) is
// (T[j-1] : T[j]) >> off with T = the tile's words in the arena and T[-1] = the last bits in front of the tile: every
// lane loads ONE word per round and takes its left neighbour's from a shuffle (the first version loaded two words per
// output word and did its index arithmetic in 64 bits: 25 instructions per word, one warp two latencies deep per
// tile).  The next tile's record and position are fetched before the current tile's words.
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const TileRec* __restrict__ recs, const long long* __restrict__ tile_pos, long long ntiles,
               const uint32_t* __restrict__ arena_words, uint32_t* __restrict__ out_words, long long out_cap) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = ((long long)gridDim.x * kCompactThreads) >> 5;
    long long tile = ((long long)blockIdx.x * kCompactThreads + threadIdx.x) >> 5;
    if (tile >= ntiles) return;
    TileRec r = recs[tile];
    long long P = tile_pos[tile];
    for (; tile < ntiles; tile += nwarps) {
        const TileRec cur = r;
        const long long Pc = P;
        if (tile + nwarps < ntiles) { r = recs[tile + nwarps]; P = tile_pos[tile + nwarps]; }
        if (cur.pad) continue;   // the arena was exhausted: TIC_E_CAPACITY is already flagged
        const long long bits = (long long)(cur.bits & kRecBitsMask);
        const long long E = Pc + bits;
        if (((((E + 7) >> 3) + 3) & ~3ll) > out_cap) continue;   // nothing past out_capacity is written
        const uint32_t* src = arena_words + (size_t)cur.off16 * 4;
        const int nw = (int)((bits + 31) >> 5);
        const int off = (int)(Pc & 31);
        const long long w_lo = Pc >> 5;
        const int nout = (int)(((cur.bits & kRecClosing) ? ((E + 31) >> 5) : (E >> 5)) - w_lo);   // words this tile owns
        if (nout <= 0) continue;   // warp-uniform
        // T[-1]: the `off` bits in front of the tile, right-aligned (never from another image: streams start 128-bit
        // aligned, so off > 0 means the image has earlier tiles)
        uint32_t lead = 0;
        if (off > 0 && lane == 0) {
            int need = off;
            for (long long pt = tile - 1; need > 0 && pt >= 0; pt--) {
                const TileRec pr = recs[pt];
                const long long pb = (long long)(pr.bits & kRecBitsMask);
                if (pb == 0) continue;
                const int take = pb < need ? (int)pb : need;
                const uint32_t* psrc = arena_words + (size_t)pr.off16 * 4;
                const uint32_t v = tile_bits_at(psrc, (int)((pb + 31) >> 5), pb - take) >> (32 - take);   // the tile's last `take` bits
                lead |= v << (off - need);
                need -= take;
            }
        }
        uint32_t* dst = out_words + w_lo;
        uint32_t carry = lead;   // lane 0: the word in front of this round's first word
        for (int j0 = 0; j0 < nout; j0 += 128) {   // four rounds of 32 words, their loads in flight together
            uint32_t x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) x[u] = j0 + 32 * u + lane < nw ? __ldg(src + j0 + 32 * u + lane) : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + 32 * u + lane;
                uint32_t prev = __shfl_up_sync(0xffffffffu, x[u], 1);
                if (lane == 0) prev = carry;
                carry = __shfl_sync(0xffffffffu, x[u], 31);
                const uint32_t val = __funnelshift_r(x[u], prev, off);   // off == 0: x itself
                if (j < nout) dst[j] = __byte_perm(val, 0, 0x0123);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// auto_generate_huffman_table=True (codec.py:146-148): symbol statistics, then the tables
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile, kFdctTc ? 4 : kCtasPerSm)
symbol_stats_kernel(const __grid_constant__ QuantParams qp, const ImageDesc* __restrict__ descs, int n_images,
                    int uniform_tpi, long long ntiles, unsigned long long* __restrict__ counters,
                    uint32_t* __restrict__ g_hist, unsigned long long* __restrict__ g_first,
                    int* __restrict__ status, const uint4* __restrict__ bmat) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* gbase = smem_raw;
    TcGroup tg{};
    uint32_t tmem_base = 0;
    if constexpr (kFdctTc) gbase = tc_cta_setup<1>(smem_raw, bmat, 0, tg, tmem_base);
    TileShared& sm = *reinterpret_cast<TileShared*>(gbase);
    const int t = threadIdx.x;
    uint32_t* hist = stats_hist(sm);                     // 272 counters
    unsigned long long* first = stats_first(sm);         // 272 keys
    const ExactStats st{&sm};
    if (t == 0) sm.stat_items = sm.stat_changed = sm.stat_unflagged = sm.tc_timeout = 0u;
    // Every CTA takes a CONTIGUOUS range of tiles and keeps one image's statistics in shared memory until the image
    // changes: the per-image bins in global memory see a few atomics per CTA and image, not 2 x 272 per tile from
    // hundreds of CTAs at once (which serialised in L2: this kernel took 4x the encode kernel).
    const long long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long long t_lo = (long long)blockIdx.x * per, t_hi = t_lo + per < ntiles ? t_lo + per : ntiles;
    int cur_img = -1;
    auto flush = [&]() {   // all threads; barriers inside
        __syncthreads();
        if (cur_img >= 0) {
            for (int i = t; i < 272; i += kTile) {
                const uint32_t c = hist[i];
                if (c) {
                    atomicAdd(&g_hist[(size_t)cur_img * 272 + i], c);
                    atomicMin(&g_first[(size_t)cur_img * 272 + i], first[i]);
                }
            }
            if (t == 0 && sm.warp_err[0]) atomicOr(&status[cur_img], TIC_STATUS_TABLE);
        }
        __syncthreads();
        for (int i = t; i < 272; i += kTile) { hist[i] = 0; first[i] = ~0ull; }
        if (t == 0) sm.warp_err[0] = 0;
        __syncthreads();
    };
    for (long long tile = t_lo; tile < t_hi; tile++) {
        const TileInfo ti = locate_tile(descs, n_images, tile, uniform_tpi);
        if (ti.img != cur_img) {   // CTA-uniform
            flush();
            cur_img = ti.img;
        }
        if constexpr (kFdctTc) {
            if (ti.nb > 0) {
                uint2 rows[8];
                load_block_rows(ti, t, rows);
                transform_tile_tc<1>(ti, qp, sm, tg, 0, false, st, rows, load_tile_halo(ti, t), [] {});
            }
        } else transform_warp(ti, qp, sm, st);
        int err = 0;
        if ((t & ~31) < ti.nb) warp_block_stats(sm, t, t < ti.nb, (unsigned long long)ti.blk0, hist, first, err);   // warp-uniform
        if (err) sm.warp_err[0] = 1;
        __syncthreads();   // the next tile's staging overwrites the coefficients this tile's statistics read
    }
    flush();
    if constexpr (kFdctTc) tc_cta_teardown<1>(tmem_base);
    __syncthreads();
    if (t == 0 && sm.tc_timeout) atomicExch(&counters[kCtrTcTimeout], 1ull);
}

// The same statistics from the persistent multi-group structure of encode_tiles_kernel (tensor-core path only): G groups
// per CTA share one TMEM allocation and one B operand, 32 warps per SM instead of the 16 that four single-group CTAs
// give (TMEM: 4 x 128 columns).  Every GROUP takes a contiguous range of tiles and keeps one image's statistics in its
// own staging window until the image changes.
template <int G>
__global__ void __launch_bounds__(kTile * G, 1)
symbol_stats_groups_kernel(const __grid_constant__ QuantParams qp, const ImageDesc* __restrict__ descs, int n_images,
                           int uniform_tpi, long long ntiles, unsigned long long* __restrict__ counters,
                           uint32_t* __restrict__ g_hist, unsigned long long* __restrict__ g_first,
                           int* __restrict__ status, const uint4* __restrict__ bmat) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int t = tid(), g = (int)threadIdx.x / kTile, lane = t & 31, warp = t >> 5;
    TcGroup tg{};
    uint32_t tmem_base = 0;
    unsigned char* gbase = tc_cta_setup<G>(smem_raw, bmat, g, tg, tmem_base);
    TileShared& sm = *reinterpret_cast<TileShared*>(gbase + kTabBytes + (size_t)g * kGroupStrideNoTab);   // encode_tiles_kernel<0>'s layout
    uint32_t* hist = stats_hist(sm);                     // 272 counters
    unsigned long long* first = stats_first(sm);         // 272 keys
    const ExactStats st{&sm};
    if (t == 0) { sm.stat_items = sm.stat_changed = sm.stat_unflagged = sm.tc_timeout = 0u; sm.warp_err[0] = 0; }
    for (int i = t; i < 272; i += kTile) { hist[i] = 0; first[i] = ~0ull; }
    const long long ngroups = (long long)gridDim.x * G, gi = (long long)blockIdx.x * G + g;
    const long long per = (ntiles + ngroups - 1) / ngroups;
    const long long t_lo = gi * per, t_hi = t_lo + per < ntiles ? t_lo + per : ntiles;
    group_sync<G>(g);
    int cur_img = -1;
    auto flush = [&]() {   // all threads of the group; barriers inside
        group_sync<G>(g);
        if (cur_img >= 0) {
            for (int i = t; i < 272; i += kTile) {
                const uint32_t c = hist[i];
                if (c) {
                    atomicAdd(&g_hist[(size_t)cur_img * 272 + i], c);
                    atomicMin(&g_first[(size_t)cur_img * 272 + i], first[i]);
                }
            }
            if (t == 0 && sm.warp_err[0]) atomicOr(&status[cur_img], TIC_STATUS_TABLE);
        }
        group_sync<G>(g);
        for (int i = t; i < 272; i += kTile) { hist[i] = 0; first[i] = ~0ull; }
        if (t == 0) sm.warp_err[0] = 0;
        group_sync<G>(g);
    };
    for (long long tile = t_lo; tile < t_hi; tile++) {
        if (warp == 0) {
            const TileInfo r = locate_tile(descs, n_images, tile, uniform_tpi);
            if (lane == 0) sm.tinfo[0] = r;
        }
        group_sync<G>(g);
        const TileInfo& ti = sm.tinfo[0];
        if (ti.img != cur_img) {   // group-uniform
            flush();
            cur_img = ti.img;
        }
        if (ti.nb > 0) {
            uint2 rows[8];
            load_block_rows(ti, t, rows);
            transform_tile_tc<G>(ti, qp, sm, tg, g, false, st, rows, load_tile_halo(ti, t), [] {});
        }
        int err = 0;
        if ((t & ~31) < ti.nb) warp_block_stats(sm, t, t < ti.nb, (unsigned long long)ti.blk0, hist, first, err);   // warp-uniform
        if (err) sm.warp_err[0] = 1;
        group_sync<G>(g);   // the next tile's description and staging overwrite what this tile's statistics read
    }
    flush();
    tc_cta_teardown<G>(tmem_base);
    group_sync<G>(g);
    if (t == 0 && sm.tc_timeout) atomicExch(&counters[kCtrTcTimeout], 1ull);
}

// One warp per image, everything in shared memory: HuffmanTree (huffman.py:112-194) on CPython's heapq, which
// queue.PriorityQueue uses — leaves pushed in first-occurrence order, nodes compared by frequency
// only, DFS with left = "0" — then write_huffman_table (codec.py:73-84) into the header words.  The heap and the
// DFS are sequential by nature (lane 0); loading, ordering the symbols by first occurrence (a rank sort over
// the 32 lanes) and writing the tables back are warp-wide.  (Round 1 ran one THREAD per image on scratch in
// global memory: 18 ms for 4096 images, three times the encode kernel.)
struct TreeScratch {
    unsigned long long freq[544];
    short left[544], right[544], sym[544];
    short heap[272];
    short stack_node[272];
    unsigned int stack_code[272];
    unsigned char stack_len[272];
};

__device__ void heap_siftdown(TreeScratch& s, int startpos, int pos) {   // heapq._siftdown
    short newitem = s.heap[pos];
    while (pos > startpos) {
        int parentpos = (pos - 1) >> 1;
        short parent = s.heap[parentpos];
        if (s.freq[newitem] < s.freq[parent]) { s.heap[pos] = parent; pos = parentpos; continue; }
        break;
    }
    s.heap[pos] = newitem;
}
__device__ void heap_siftup(TreeScratch& s, int hlen, int pos) {   // heapq._siftup
    int startpos = pos;
    short newitem = s.heap[pos];
    int childpos = 2 * pos + 1;
    while (childpos < hlen) {
        int rightpos = childpos + 1;
        if (rightpos < hlen && !(s.freq[s.heap[childpos]] < s.freq[s.heap[rightpos]])) childpos = rightpos;
        s.heap[pos] = s.heap[childpos];
        pos = childpos;
        childpos = 2 * pos + 1;
    }
    s.heap[pos] = newitem;
    heap_siftdown(s, startpos, pos);
}

struct HdrWriter {
    uint32_t* words;
    int nbits;
    __device__ void put(uint32_t v, int n) {   // n <= 32, MSB-first; bits past the buffer are counted, not stored
        if (n <= 0) return;
        if (n < 32) v &= (1u << n) - 1u;
        const int w = nbits >> 5, sh = nbits & 31;                   // the value occupies bits [sh, sh + n) of words w, w + 1
        const unsigned long long x = ((unsigned long long)v << (64 - n)) >> sh;   // left-aligned at bit sh of a 64-bit window
        if (w < kMaxHdrWords) words[w] |= (uint32_t)(x >> 32);
        if (w + 1 < kMaxHdrWords && (uint32_t)x) words[w + 1] |= (uint32_t)x;
        nbits += n;
    }
};

// Builds one alphabet: symbols base..base+count-1 of hist/first.  Returns status bits.
// order[0..n): the present symbols in first-occurrence order (dict insertion order, huffman.py:187-194)
__device__ int build_alphabet(TreeScratch& s, const uint32_t* hist, int base, const short* order, int n, bool is_dc,
                              HuffEntry* table, HdrWriter& hw) {
    int status = 0;
    hw.put((uint32_t)n, 16);                                     // codec.py:74,79
    if (n == 0) return status;
    int nnodes = 0, hlen = 0;
    for (int i = 0; i < n; i++) {                                // huffman.py:156-157
        s.freq[nnodes] = hist[base + order[i]];
        s.sym[nnodes] = order[i];
        s.left[nnodes] = s.right[nnodes] = -1;
        s.heap[hlen++] = (short)nnodes++;
        heap_siftdown(s, 0, hlen - 1);
    }
    while (hlen >= 2) {                                          // huffman.py:159-163
        short uv[2];
        for (int r = 0; r < 2; r++) {                            // heapq.heappop
            short last = s.heap[--hlen];
            if (hlen) { uv[r] = s.heap[0]; s.heap[0] = last; heap_siftup(s, hlen, 0); } else uv[r] = last;
        }
        s.freq[nnodes] = s.freq[uv[0]] + s.freq[uv[1]];
        s.sym[nnodes] = -1;
        s.left[nnodes] = uv[0];
        s.right[nnodes] = uv[1];
        s.heap[hlen++] = (short)nnodes++;
        heap_siftdown(s, 0, hlen - 1);
    }
    // DFS, left = "0" first (huffman.py:175-185); table order = DFS order
    int sp = 0;
    s.stack_node[0] = s.heap[0]; s.stack_code[0] = 0; s.stack_len[0] = 0; sp = 1;
    while (sp) {
        sp--;
        const int node = s.stack_node[sp];
        const unsigned int code = s.stack_code[sp];
        const int len = s.stack_len[sp];
        if (s.sym[node] >= 0) {
            const int sym = s.sym[node];
            if (len > 32) status |= TIC_STATUS_LONGCODE;         // device limit: codes up to 32 bits
            table[sym].code = code;
            table[sym].len = kHuffPresent | (uint32_t)(len > 32 ? 32 : len);
            if (is_dc) {                                         // codec.py:75-78
                if (len >= 16) status |= TIC_STATUS_TABLE;       // int2ba(len, 4) raises OverflowError
                hw.put((uint32_t)sym, 4);
                hw.put((uint32_t)len & 15u, 4);
            } else {                                             // codec.py:80-84
                hw.put((uint32_t)sym >> 4, 4);
                hw.put((uint32_t)sym & 15u, 4);
                hw.put((uint32_t)len & 255u, 8);
            }
            hw.put(code, len > 32 ? 32 : len);
            continue;
        }
        // push right first so that the left subtree is visited first
        s.stack_node[sp] = s.right[node]; s.stack_code[sp] = (code << 1) | 1u; s.stack_len[sp] = (unsigned char)(len + 1); sp++;
        s.stack_node[sp] = s.left[node];  s.stack_code[sp] = (code << 1);      s.stack_len[sp] = (unsigned char)(len + 1); sp++;
    }
    return status;
}

struct TableShared {
    TreeScratch s;
    AutoTables at;
    uint32_t hist[272];
    unsigned long long first[272];
    short order[2][256];
    int n[2];
};

// Warp-wide rank sort of the present symbols of one alphabet by first occurrence (the keys are distinct).
__device__ void order_alphabet(const uint32_t* hist, const unsigned long long* first, int base, int count, short* order, int* n_out) {
    const int lane = threadIdx.x;
    int n = 0;
    for (int i0 = 0; i0 < count; i0 += 32) {
        const int i = i0 + lane;
        const bool present = i < count && hist[base + i] != 0;
        if (present) {
            const unsigned long long key = first[base + i];
            int rank = 0;
            for (int j = 0; j < count; j++) rank += (hist[base + j] != 0 && first[base + j] < key) ? 1 : 0;
            order[rank] = (short)i;
        }
        n += __popc(__ballot_sync(0xffffffffu, present));
    }
    if (lane == 0) *n_out = n;
}

__global__ void __launch_bounds__(32)
build_tables_kernel(const ImageDesc* __restrict__ descs, int n_images, int quality, int le_flag,
                    const uint32_t* __restrict__ g_hist, const unsigned long long* __restrict__ g_first,
                    AutoTables* __restrict__ tabs, int* __restrict__ status) {
    __shared__ TableShared ts;
    const int img = blockIdx.x, lane = threadIdx.x;
    if (img >= n_images) return;
    for (int i = lane; i < 272; i += 32) { ts.hist[i] = g_hist[(size_t)img * 272 + i]; ts.first[i] = g_first[(size_t)img * 272 + i]; }
    uint32_t* atw = reinterpret_cast<uint32_t*>(&ts.at);
    for (int i = lane; i < (int)(sizeof(AutoTables) / 4); i += 32) atw[i] = 0u;
    __syncwarp();
    order_alphabet(ts.hist, ts.first, 256, 16, ts.order[0], &ts.n[0]);    // DC sizes
    order_alphabet(ts.hist, ts.first, 0, 256, ts.order[1], &ts.n[1]);     // AC (run, size)
    __syncwarp();
    if (lane == 0) {
        HdrWriter hw{ts.at.hdr_words, 0};
        const ImageDesc d = descs[img];
        hw.put(__byte_perm((uint32_t)d.h, 0, 0x0123), 32);           // struct.pack("III"), codec.py:103-109
        hw.put(__byte_perm((uint32_t)d.w, 0, 0x0123), 32);
        hw.put(__byte_perm((uint32_t)quality, 0, 0x0123), 32);
        // codec.py:111 writes the flag MSB-first (80 00 00 00); le_flag: as the little-endian word the
        // reference's parse_header (codec.py:119) can read
        hw.put(le_flag ? 0x00000080u : 0x80000000u, 32);
        int st = build_alphabet(ts.s, ts.hist, 256, ts.order[0], ts.n[0], true, ts.at.tab.dc, hw);    // table[DC] first, codec.py:74-78
        st |= build_alphabet(ts.s, ts.hist, 0, ts.order[1], ts.n[1], false, ts.at.tab.ac, hw);         // then table[AC], codec.py:79-84
        if (hw.nbits > kMaxHdrWords * 32) st |= TIC_STATUS_LONGCODE;
        ts.at.hdr_bits = (uint32_t)hw.nbits;
        ts.at.status = (uint32_t)st;
        if (st) atomicOr(&status[img], st);
    }
    __syncwarp();
    uint32_t* dst = reinterpret_cast<uint32_t*>(&tabs[img]);
    for (int i = lane; i < (int)(sizeof(AutoTables) / 4); i += 32) dst[i] = atw[i];
}

// Uniform batch (every image H x W, equally spaced in memory): the descriptors are written on the device and the
// per-batch status / counters are cleared by the same launch — instead of one H2D copy and two memsets.
__global__ void prep_uniform_kernel(ImageDesc* __restrict__ descs, int n_images, const uint8_t* base, long long img_stride,
                                    int h, int w, int bw, int bw_shift, int nblk, long long tiles_per_image,
                                    int* __restrict__ status, unsigned long long* __restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kCtrCount) counters[i] = 0ull;
    if (i >= n_images) return;
    ImageDesc d;
    d.px = base + (long long)i * img_stride;
    d.h = h; d.w = w; d.bw = bw; d.nblk = nblk; d.tile0 = (long long)i * tiles_per_image; d.bw_shift = bw_shift; d.pad = 0;
    descs[i] = d;
    status[i] = 0;
}

// ---------------------------------------------------------------------------------------------
// encode(): the same transform, coefficients written out (parity checkpoint for codec.py:26-43)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile, kFdctTc ? 4 : kCtasPerSm)
coeffs_kernel(const __grid_constant__ QuantParams qp, const ImageDesc* __restrict__ descs,
              unsigned long long* __restrict__ counters, int* __restrict__ dc, int* __restrict__ ac,
              const uint4* __restrict__ bmat, uint32_t flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* gbase = smem_raw;
    TcGroup tg{};
    uint32_t tmem_base = 0;
    if constexpr (kFdctTc) gbase = tc_cta_setup<1>(smem_raw, bmat, 0, tg, tmem_base);
    TileShared& sm = *reinterpret_cast<TileShared*>(gbase);
    const int t = threadIdx.x;
    const TileInfo ti = locate_tile(descs, 1, blockIdx.x, 0);
    const ExactStats st{&sm};
    if (t == 0) sm.stat_items = sm.stat_changed = sm.stat_unflagged = sm.tc_timeout = 0u;
    __syncthreads();
    if constexpr (kFdctTc) {
        if (ti.nb > 0) {
            uint2 rows[8];
            load_block_rows(ti, t, rows);
            transform_tile_tc<1>(ti, qp, sm, tg, 0, (flags & TIC_FLAG_DEBUG_ALL_EXACT) != 0, st, rows, load_tile_halo(ti, t), [] {});
        }
    } else transform_warp(ti, qp, sm, st);
    if (t < ti.nb) {
        const size_t b = (size_t)ti.blk0 + t;
        dc[b] = sm.dcq[t] - dc_before(sm, t);
        int* row = ac + b * 63;
        const uint32_t lo = sm.nz_lo[t], hi = sm.nz_hi[t];
        for (int k = 1; k < 64; k++) {   // a coefficient without its mask bit was never stored: it is 0
            const bool nz = ((k < 32 ? lo : hi) >> (31 - (k & 31))) & 1u;
            row[k - 1] = nz ? coef_get(sm, t, k) : 0;
        }
    }
    if constexpr (kFdctTc) tc_cta_teardown<1>(tmem_base);
    __syncthreads();
    if (t == 0) {
        if (sm.stat_items) atomicAdd(&counters[kCtrExactItems], (unsigned long long)sm.stat_items);
        if (sm.stat_changed) atomicAdd(&counters[kCtrExactChanged], (unsigned long long)sm.stat_changed);
        if (sm.stat_unflagged) atomicAdd(&counters[kCtrUnflagged], (unsigned long long)sm.stat_unflagged);
        if (sm.tc_timeout) atomicExch(&counters[kCtrTcTimeout], 1ull);
    }
}

}  // namespace tic

// =============================================================================================
// host side
// =============================================================================================
using namespace tic;

struct tic_handle_s {
    int device = 0;
    std::string err;
    // workspace (grown on demand)
    ImageDesc* d_descs = nullptr;       size_t descs_cap = 0;
    // pinned descriptor staging: one slot per batch in flight, guarded by an event recorded behind its H2D copy
    static constexpr int kDescSlots = 4;
    ImageDesc* h_descs_ring[kDescSlots] = {};
    cudaEvent_t desc_ev[kDescSlots] = {};
    long long desc_seq = 0;
    ImageDesc* h_descs = nullptr;       // the slot of the batch being prepared
    unsigned long long* d_sticky = nullptr;      // what tic_encode_finish reads: accumulated over batches (fold_counters)
    uint4* d_bmat = nullptr;            // tensor-core B operands, one 16 KB matrix per quality (built on first use)
    bool bmat_ready[100] = {};
    TileRec* d_recs = nullptr; long long* d_tile_pos = nullptr; size_t tiles_cap = 0;
    Span* d_chunk_span = nullptr; long long* d_chunk_pos = nullptr; size_t chunks_cap = 0;
    uint4* d_arena = nullptr;           size_t arena_cap16 = 0;          // 16-byte units
    unsigned long long* d_counters = nullptr;
    unsigned long long* h_counters = nullptr;    // pinned
    long long* d_out_end = nullptr;     size_t end_cap = 0;
    // auto-table mode: per-image statistics, tables, tree scratch
    uint32_t* d_hist = nullptr; unsigned long long* d_first = nullptr; AutoTables* d_tabs = nullptr;
    size_t auto_cap = 0;
    // single-image host path
    uint8_t* d_px = nullptr;            size_t px_cap = 0;
    uint8_t* d_out = nullptr;           size_t out_cap = 0;
    uint8_t* h_stage = nullptr;         size_t h_stage_cap = 0;          // pinned
    long long* d_meta = nullptr;        // off, size, status for the single-image path
    long long* h_meta = nullptr;        // pinned
    cudaStream_t own_stream = nullptr;
    // device timing of the two big kernels: a ring of event quadruples, one per batch since the last finish
    static constexpr int kEvRing = 64;
    cudaEvent_t ev[kEvRing][4] = {};
    long long ev_batches = 0;            // batches enqueued since the last tic_encode_finish
    double sum_encode_ms = 0.0, sum_compact_ms = 0.0;
    long long timed_batches = 0;
    long long last_tiles = 0, last_blocks = 0, last_launches = 0;
    bool tables_ready = false;
    int sm_count = 148, ctas_per_sm = 1, ctas_per_sm_c = kCtasPerSm, ctas_per_sm_single = 4;
    void* dec_ws = nullptr;              // decode-side workspace, owned by tic_decode.cu
};

// hooks for tic_decode.cu (the decode side lives in its own translation unit)
void tic_internal_dec_release(void* ws);
void** tic_internal_dec_slot(tic_handle h) { return &h->dec_ws; }
void tic_internal_set_error(tic_handle h, const std::string& msg) { h->err = msg; }
int tic_internal_device(tic_handle h) { return h->device; }
cudaStream_t tic_internal_own_stream(tic_handle h) { return h->own_stream; }

#define TIC_CUDA(h, call)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return TIC_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

static int bw_shift_of(int bw) {   // log2 for a power of two, else -1
    if (bw <= 0 || (bw & (bw - 1))) return -1;
    int s = 0;
    while ((1 << s) < bw) s++;
    return s;
}

static void build_default_tables(HuffTables& t) {
    memset(&t, 0, sizeof t);
    uint32_t code = 0;
    int k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < kDcBits[l]; i++) { t.dc[kDcVals[k]].code = code++; t.dc[kDcVals[k]].len = kHuffPresent | (uint32_t)l; k++; }
        code <<= 1;
    }
    code = 0;
    k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < kAcBits[l]; i++) { t.ac[kAcVals[k]].code = code++; t.ac[kAcVals[k]].len = kHuffPresent | (uint32_t)l; k++; }
        code <<= 1;
    }
}

// Quality -> quantiser constants.  qt follows tinyimgcodec/utils.py:50-53 operation by operation.
static int make_quant_params(int quality, QuantParams& qp) {
    if (quality < 1 || quality > 99) return TIC_E_QUALITY;
    for (int i = 0; i < 64; i++) {
        double q;
        if (quality < 50) {
            double factor = 5000.0 / (double)quality;
            q = ((double)kQuantBase[i] * factor) / 100.0;
        } else {
            long factor = 200 - 2 * (long)quality;
            q = (double)((long)kQuantBase[i] * factor) / 100.0;
        }
        qp.qt[i] = q;
    }
    double aan[8];
    aan[0] = 1.0;
    for (int k = 1; k < 8; k++) aan[k] = cos(k * 3.14159265358979323846 / 16.0) * sqrt(2.0);
    // kFastErr: bound on |FP32 AAN coefficient - float64 reference coefficient| in coefficient
    // units ("tie guard").  A coefficient can only be rounded differently from the
    // reference if a .5 tie lies within w = kFastErr/qt (+ the FP32 multiplier's relative error at
    // the largest possible |t| = 1024/qt, + the rounding of the residual itself) of the fast value t;
    // every such coefficient has |t - round(t)| > 0.5 - w and is recomputed exactly.
    const double kFastErr = 6.0e-4;
    static const int zigzag[64] = {TIC_ZIGZAG_LIST};
    for (int k = 0; k < 64; k++) {
        const int i = zigzag[k], u = i >> 3, v = i & 7;
        double m = 1.0 / (8.0 * aan[u] * aan[v] * qp.qt[i]);
        double w = kFastErr / qp.qt[i] + 2.4e-7 * (1024.0 / qp.qt[i]) + 1.0e-6;
        double hthr = 0.5 - w;
        if (hthr < 0.25) return TIC_E_QUALITY;   // unreachable for quality <= 99 (w <= 4.2e-3); the tensor-core scaling divides by hthr
        qp.qmul[k] = (float)m;
        qp.hthr[k] = (float)(hthr * (1.0 - 1.0e-6));
        // |d * zmul| < 1  =>  |t| < hthr: rounds to zero, and the residual test cannot fire
        qp.zmul[k] = hthr > 0.0 ? (float)(m / hthr * (1.0 + 1.0e-6)) : 3.0e38f;
    }
    qp.dcinv = 1.0 / (8.0 * qp.qt[0]);
    qp.dcinv_f = (float)qp.dcinv;
    // tensor-core path: the accumulator holds t / hthr * 2^E with ONE power of two per quality, chosen so that the
    // largest entry of B = basis / (qt * hthr) * 2^E is in [2^13, 2^14): its f16 hi + lo split then carries >= 21 bits.
    double mx = 0.0;
    for (int k = 0; k < 64; k++) {
        const int i = zigzag[k], u = i >> 3, v = i & 7;
        const double cu = u ? 0.5 : sqrt(0.125), cv = v ? 0.5 : sqrt(0.125);
        const double m = cu * cv / (qp.qt[i] * (double)qp.hthr[k]);   // |cos| <= 1
        if (m > mx) mx = m;
    }
    int e = 0;
    frexp(mx, &e);
    qp.tc_exp = 14 - e;
    qp.tc_live = (float)ldexp(1.0, qp.tc_exp);
    for (int k = 0; k < 64; k++) qp.cn[k] = (float)ldexp((double)qp.hthr[k], -qp.tc_exp);   // exact: a power-of-two scaling
    return TIC_OK;
}

// B operand of the tensor-core FDCT for one quality, in the shared-memory layout of tic_tc.cuh:
// B[n][k] = basis(zigzag[n], pixel k) / (qt * hthr) * 2^E as f16 hi (k) and lo (k + 64).
static void build_bmat(const QuantParams& qp, __half* blob /* tc::kN x 128, zero-filled */) {
    static const int zigzag[64] = {TIC_ZIGZAG_LIST};
    const double pi = 3.14159265358979323846;
    auto at = [&](int n, int k) -> __half& {   // shared-memory layout of tic_tc.cuh, in f16 units
        return blob[(size_t)(n / 8) * 64 + (size_t)(k / 8) * (tc::kLboB / 2) + (size_t)(n % 8) * 8 + (size_t)(k % 8)];
    };
    for (int n = 0; n < 64; n++) {
        const int i = zigzag[n], u = i >> 3, v = i & 7;
        const double cu = u ? 0.5 : sqrt(0.125), cv = v ? 0.5 : sqrt(0.125);
        const double scale = ldexp(cu * cv / (qp.qt[i] * (double)qp.hthr[n]), qp.tc_exp);
        for (int k = 0; k < 64; k++) {
            const int y = k >> 3, x = k & 7;
            const double b = scale * cos((2 * y + 1) * u * pi / 16.0) * cos((2 * x + 1) * v * pi / 16.0);
            const __half hi = __double2half(b);
            at(n, k) = hi;
            at(n, k + 64) = __double2half(b - (double)__half2float(hi));
        }
    }
    // columns 64..71: S_x, columns 72..79: I_x (tic_tc.cuh); entries +-1 in the hi half, the lo half stays 0
    static const int sgn[8] = {1, -1, -1, 1, 1, -1, -1, 1};
    for (int x = 0; x < 8 && tc::kN >= tc::kColSums + 16; x++)
        for (int y = 0; y < 8; y++) {
            at(tc::kColSums + x, 8 * y + x) = __double2half(1.0);
            at(tc::kColSums + 8 + x, 8 * y + x) = __double2half((double)sgn[y]);
        }
}

constexpr size_t kSmemEncode = cta_smem_bytes<kGroups, kFdctTc, kFdctTc>();   // persistent encode kernel, fixed tables; statistics
constexpr size_t kSmemEncodeAuto = cta_smem_bytes<kGroupsAuto, kFdctTc>();    // persistent encode kernel, per-image tables
constexpr size_t kSmemEncodeC = cta_smem_bytes<1, false>();               // C variant: CUDA-core integer transform
constexpr size_t kSmemSingle = cta_smem_bytes<1, kFdctTc>();              // symbol_stats_kernel, coeffs_kernel
#ifndef TIC_SKIP_SMEM_ASSERT
static_assert(kSmemEncode <= 232448, "shared memory of the persistent encode kernel exceeds 227 KB: lower TIC_GROUPS");
static_assert(kSmemEncodeAuto <= 232448, "shared memory of the persistent encode kernel exceeds 227 KB: lower TIC_GROUPS_AUTO");
#endif

static int ensure_tables(tic_handle h) {
    if (h->tables_ready) return TIC_OK;
    HuffTables t;
    build_default_tables(t);
    TIC_CUDA(h, cudaMemcpyToSymbol(c_default_tables, &t, sizeof t));
    TIC_CUDA(h, cudaFuncSetAttribute(encode_tiles_kernel<0, kGroups>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemEncode));
    TIC_CUDA(h, cudaFuncSetAttribute(encode_tiles_kernel<1, kGroupsAuto>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemEncodeAuto));
    TIC_CUDA(h, cudaFuncSetAttribute(encode_tiles_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemEncodeC));
    uint16_t sq[4][64];   // c/img.c:157-181
    for (int f = 0; f < 4; f++)
        for (int i = 0; i < 64; i++) sq[f][i] = (uint16_t)(65536 / (kQuantBase[i] << f));
    TIC_CUDA(h, cudaMemcpyToSymbol(c_cvar_scaled_quant, sq, sizeof sq));
    TIC_CUDA(h, cudaFuncSetAttribute(coeffs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSingle));
    TIC_CUDA(h, cudaFuncSetAttribute(symbol_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSingle));
    if (kFdctTc) TIC_CUDA(h, cudaFuncSetAttribute(symbol_stats_groups_kernel<kGroups>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemEncode));
    int per_sm = 0, per_sm_c = 0, per_sm_single = 0;
    TIC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_tiles_kernel<0, kGroups>, kTile * kGroups, kSmemEncode));
    TIC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, encode_tiles_kernel<2, 1>, kTile, kSmemEncodeC));
    TIC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_single, symbol_stats_kernel, kTile, kSmemSingle));
    cudaDeviceProp prop;
    TIC_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    h->ctas_per_sm = per_sm > 0 ? per_sm : 1;
    h->ctas_per_sm_c = per_sm_c > 0 ? per_sm_c : 1;
    h->ctas_per_sm_single = per_sm_single > 4 ? per_sm_single : 4;   // launch bounds: 4 CTAs of 128 threads per SM (54 KB, 128 registers)
    h->tables_ready = true;
    return TIC_OK;
}

// The B operand of `quality` in device memory (built once per handle and quality).
static int ensure_bmat(tic_handle h, int quality, const QuantParams& qp, const uint4** out) {
    *out = nullptr;
    if (!kFdctTc) return TIC_OK;
    if (!h->d_bmat) TIC_CUDA(h, cudaMalloc(&h->d_bmat, (size_t)100 * tc::kBBytes));
    uint4* dst = h->d_bmat + (size_t)quality * (tc::kBBytes / 16);
    if (!h->bmat_ready[quality]) {
        std::vector<__half> blob((size_t)tc::kN * 128, __double2half(0.0));
        build_bmat(qp, blob.data());
        // synchronous and device-wide: whatever stream a later batch of this quality runs on, the matrix is there
        TIC_CUDA(h, cudaMemcpy(dst, blob.data(), tc::kBBytes, cudaMemcpyHostToDevice));
        TIC_CUDA(h, cudaDeviceSynchronize());
        h->bmat_ready[quality] = true;
    }
    *out = dst;
    return TIC_OK;
}

extern "C" {

const char* tic_version(void) { return "tinyimgcodec_cuda 0.1 sm_100a"; }

int64_t tic_num_blocks(int32_t height, int32_t width) {
    if (height <= 0 || width <= 0) return 0;
    return (int64_t)((height + 7) / 8) * (int64_t)((width + 7) / 8);
}

int64_t tic_max_out_bytes(int32_t height, int32_t width) {
    return 16 + (tic_num_blocks(height, width) * 1662 + 7) / 8 + 16;
}

int tic_create(int device, tic_handle* out) {
    if (!out) return TIC_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return TIC_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TIC_E_CUDA;
    if (prop.major != 10) return TIC_E_CUDA;   // sm_100a binary only: no fallback, fail loudly
    tic_handle h = new tic_handle_s();
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&h->d_counters, kCtrCount * 8) != cudaSuccess ||
        cudaMalloc(&h->d_sticky, kCtrCount * 8) != cudaSuccess || cudaMemset(h->d_sticky, 0, kCtrCount * 8) != cudaSuccess ||
        cudaMallocHost(&h->h_counters, kCtrCount * 8) != cudaSuccess ||
        cudaMalloc(&h->d_meta, 64) != cudaSuccess || cudaMallocHost(&h->h_meta, 64) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return TIC_E_CUDA;
    }
    for (int i = 0; i < tic_handle_s::kEvRing; i++)
        for (int j = 0; j < 4; j++)
            if (cudaEventCreate(&h->ev[i][j]) != cudaSuccess) { tic_destroy(h); return TIC_E_CUDA; }
    for (int i = 0; i < tic_handle_s::kDescSlots; i++)
        if (cudaEventCreateWithFlags(&h->desc_ev[i], cudaEventDisableTiming) != cudaSuccess) { tic_destroy(h); return TIC_E_CUDA; }
    *out = h;
    return TIC_OK;
}

int tic_destroy(tic_handle h) {
    if (!h) return TIC_E_INVALID;
    cudaSetDevice(h->device);
    tic_internal_dec_release(h->dec_ws);
    cudaFree(h->d_descs); cudaFree(h->d_sticky); cudaFree(h->d_bmat); cudaFree(h->d_recs);
    for (int i = 0; i < tic_handle_s::kDescSlots; i++) { cudaFreeHost(h->h_descs_ring[i]); if (h->desc_ev[i]) cudaEventDestroy(h->desc_ev[i]); } cudaFree(h->d_tile_pos); cudaFree(h->d_chunk_span); cudaFree(h->d_chunk_pos);
    cudaFree(h->d_arena); cudaFree(h->d_counters);
    cudaFree(h->d_hist); cudaFree(h->d_first); cudaFree(h->d_tabs);
    cudaFreeHost(h->h_counters); cudaFree(h->d_out_end); cudaFree(h->d_px); cudaFree(h->d_out);
    cudaFreeHost(h->h_stage); cudaFree(h->d_meta); cudaFreeHost(h->h_meta);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (int i = 0; i < tic_handle_s::kEvRing; i++)
        for (int j = 0; j < 4; j++) if (h->ev[i][j]) cudaEventDestroy(h->ev[i][j]);
    delete h;
    return TIC_OK;
}

const char* tic_last_error(tic_handle h) { return h ? h->err.c_str() : "null handle"; }

// Room for n descriptors, and the pinned slot this batch stages them in.  A slot is reused only after the H2D copy
// of the batch that used it last has completed (an event recorded behind that copy), so batches may be enqueued
// back to back without tic_encode_finish in between.
static int grow_descs(tic_handle h, size_t n) {
    if (n > h->descs_cap) {
        TIC_CUDA(h, cudaDeviceSynchronize());   // nothing in flight reads the old buffers
        size_t cap = n < 64 ? 64 : n * 2;
        cudaFree(h->d_descs);
        h->d_descs = nullptr; h->descs_cap = 0;
        for (int i = 0; i < tic_handle_s::kDescSlots; i++) { cudaFreeHost(h->h_descs_ring[i]); h->h_descs_ring[i] = nullptr; }
        TIC_CUDA(h, cudaMalloc(&h->d_descs, cap * sizeof(ImageDesc)));
        for (int i = 0; i < tic_handle_s::kDescSlots; i++) TIC_CUDA(h, cudaMallocHost(&h->h_descs_ring[i], cap * sizeof(ImageDesc)));
        h->descs_cap = cap;
    }
    const int slot = (int)(h->desc_seq % tic_handle_s::kDescSlots);
    TIC_CUDA(h, cudaEventSynchronize(h->desc_ev[slot]));   // returns at once for an event never recorded
    h->h_descs = h->h_descs_ring[slot];
    return TIC_OK;
}
// after the H2D copy of the staged descriptors has been enqueued on `stream`
static int descs_staged(tic_handle h, cudaStream_t stream) {
    const int slot = (int)(h->desc_seq % tic_handle_s::kDescSlots);
    TIC_CUDA(h, cudaEventRecord(h->desc_ev[slot], stream));
    h->desc_seq++;
    return TIC_OK;
}

int tic_encode_batch(tic_handle h, const void* const* d_pixels, const int32_t* heights, const int32_t* widths,
                     int32_t n_images, int32_t quality, uint32_t flags, void* d_out, int64_t out_capacity,
                     int64_t* d_out_offsets, int64_t* d_out_sizes, int32_t* d_status, void* stream_v) {
    if (!h) return TIC_E_INVALID;
    h->err.clear();
    if (n_images < 0 || out_capacity < 0 || (n_images > 0 && (!d_pixels || !heights || !widths || !d_out ||
                                                              !d_out_offsets || !d_out_sizes || !d_status))) {
        h->err = "invalid argument";
        return TIC_E_INVALID;
    }
    const bool auto_mode = (flags & TIC_FLAG_AUTO_HUFFMAN) != 0;
    const bool c_variant = (flags & TIC_FLAG_C_VARIANT) != 0;
    if (auto_mode && c_variant) { h->err = "the C-variant stream has fixed tables"; return TIC_E_INVALID; }
    if ((flags & TIC_FLAG_AUTO_LE_FLAG) && !auto_mode) { h->err = "TIC_FLAG_AUTO_LE_FLAG needs TIC_FLAG_AUTO_HUFFMAN"; return TIC_E_INVALID; }
    if ((reinterpret_cast<uintptr_t>(d_out) & 15) != 0) {
        h->err = "d_out must be 16-byte aligned";
        return TIC_E_INVALID;
    }
    QuantParams qp;
    int rc = TIC_OK;
    if (c_variant) {   // quality is IMG_Q_BEST .. IMG_Q_LOW (c/img.h:22); the float quantiser is unused
        if (quality < 0 || quality > 3) { h->err = "C variant: quality must be 0..3 (IMG_Q_BEST..IMG_Q_LOW)"; return TIC_E_QUALITY; }
        memset(&qp, 0, sizeof qp);
    } else {
        rc = make_quant_params(quality, qp);
        if (rc) { h->err = "quality must be in 1..99"; return rc; }
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    TIC_CUDA(h, cudaSetDevice(h->device));
    rc = ensure_tables(h);
    if (rc) return rc;
    h->last_tiles = h->last_blocks = h->last_launches = 0;
    cudaEvent_t* evq = h->ev[h->ev_batches % tic_handle_s::kEvRing];
    if (n_images == 0) {   // nothing to encode: the byte total of "the last batch" is 0
        TIC_CUDA(h, cudaMemsetAsync(h->d_sticky + kCtrTotalBits, 0, 8, stream));
        return TIC_OK;
    }
    rc = grow_descs(h, (size_t)n_images);
    if (rc) return rc;
    long long ntiles = 0, nblocks = 0, worst = 16;
    int uniform_tpi = 0;
    // same shape and equally spaced pointers (an (N, H, W) tensor): the descriptors can be built on the device
    bool regular = n_images >= 1;
    const long long img_stride = n_images >= 2 ? (long long)((const uint8_t*)d_pixels[1] - (const uint8_t*)d_pixels[0]) : 0;
    for (int i = 0; i < n_images; i++) {
        if (heights[i] != heights[0] || widths[i] != widths[0] ||
            (const uint8_t*)d_pixels[i] != (const uint8_t*)d_pixels[0] + (long long)i * img_stride) regular = false;
        if (heights[i] < 0 || widths[i] < 0) { h->err = "negative image dimension"; return TIC_E_INVALID; }
        long long nblk = tic_num_blocks(heights[i], widths[i]);
        if (c_variant) {
            if ((heights[i] & 7) || (widths[i] & 7)) {   // c/encode.c:38-41
                h->err = "C variant: width and height must be multiples of 8";
                return TIC_E_INVALID;
            }
        }
        if (nblk > 0x7fffffffll - kTile) { h->err = "image too large"; return TIC_E_INVALID; }
        ImageDesc& d = h->h_descs[i];
        d.px = (const uint8_t*)d_pixels[i];
        d.h = heights[i];
        d.w = widths[i];
        d.bw = (widths[i] + 7) / 8;
        d.bw_shift = bw_shift_of(d.bw);
        d.pad = 0;
        d.nblk = (int)nblk;
        d.tile0 = ntiles;
        long long nt = (nblk + kTile - 1) / kTile;
        if (nt < 1) nt = 1;          // an empty image still owns one tile: it writes the header
        if (i == 0) uniform_tpi = (int)nt; else if (nt != uniform_tpi) uniform_tpi = -1;
        ntiles += nt;
        nblocks += nblk;
        worst += tic_max_out_bytes(heights[i], widths[i]) + (auto_mode ? 1664 : 0);
        if (nblk > 0 && !d.px) { h->err = "null pixel pointer"; return TIC_E_INVALID; }
    }
    if (uniform_tpi < 0) uniform_tpi = 0;
    // the tile loop keeps (tile + stride) in 31 bits
    if (ntiles > 0x7fffffffll - (1ll << 20)) { h->err = "batch too large for one launch"; return TIC_E_INVALID; }
    const long long nchunks = (ntiles + kScanChunk - 1) / kScanChunk;
    if ((size_t)ntiles > h->tiles_cap) {
        cudaFree(h->d_recs); cudaFree(h->d_tile_pos);
        h->d_recs = nullptr; h->d_tile_pos = nullptr; h->tiles_cap = 0;
        size_t cap = (size_t)ntiles + (size_t)ntiles / 4 + 1024;
        TIC_CUDA(h, cudaMalloc(&h->d_recs, cap * sizeof(TileRec)));
        TIC_CUDA(h, cudaMalloc(&h->d_tile_pos, cap * sizeof(long long)));
        h->tiles_cap = cap;
    }
    if ((size_t)nchunks > h->chunks_cap) {
        cudaFree(h->d_chunk_span); cudaFree(h->d_chunk_pos);
        h->d_chunk_span = nullptr; h->d_chunk_pos = nullptr; h->chunks_cap = 0;
        size_t cap = (size_t)nchunks * 2 + 64;
        TIC_CUDA(h, cudaMalloc(&h->d_chunk_span, cap * sizeof(Span)));
        TIC_CUDA(h, cudaMalloc(&h->d_chunk_pos, cap * sizeof(long long)));
        h->chunks_cap = cap;
    }
    // The arena holds every tile's bits at a 16-byte granule before they move to their final place:
    // never more than the streams themselves (<= out_capacity, or the caller gets TIC_E_CAPACITY
    // anyway) plus one granule of slack per tile.
    unsigned long long arena_need = (unsigned long long)(worst < out_capacity ? worst : out_capacity);
    arena_need = (arena_need + 15) / 16 + (unsigned long long)ntiles + 4096;
    if (arena_need > 0xfffffff0ull) arena_need = 0xfffffff0ull;   // 32-bit granule offsets: 64 GiB
    if (arena_need > h->arena_cap16) {
        cudaFree(h->d_arena);
        h->d_arena = nullptr; h->arena_cap16 = 0;
        TIC_CUDA(h, cudaMalloc(&h->d_arena, (size_t)arena_need * 16));
        h->arena_cap16 = (size_t)arena_need;
    }
    if ((size_t)n_images > h->end_cap) {
        cudaFree(h->d_out_end);
        h->d_out_end = nullptr; h->end_cap = 0;
        TIC_CUDA(h, cudaMalloc(&h->d_out_end, (size_t)n_images * 2 * 8));
        h->end_cap = (size_t)n_images * 2;
    }
    const bool prep_on_device = regular && uniform_tpi > 0;
    if (prep_on_device) {
        const ImageDesc& d0 = h->h_descs[0];
        prep_uniform_kernel<<<(n_images + 255) / 256, 256, 0, stream>>>(h->d_descs, n_images, d0.px, img_stride, d0.h, d0.w, d0.bw,
                                                                      d0.bw_shift, d0.nblk, (long long)uniform_tpi, d_status,
                                                                      h->d_counters);
        TIC_CUDA(h, cudaGetLastError());
    } else {
        TIC_CUDA(h, cudaMemcpyAsync(h->d_descs, h->h_descs, (size_t)n_images * sizeof(ImageDesc),
                                    cudaMemcpyHostToDevice, stream));
        rc = descs_staged(h, stream);
        if (rc) return rc;
        TIC_CUDA(h, cudaMemsetAsync(h->d_counters, 0, kCtrCount * 8, stream));
        TIC_CUDA(h, cudaMemsetAsync(d_status, 0, (size_t)n_images * 4, stream));
    }
    const uint4* d_bmat = nullptr;
    if (!c_variant) {
        rc = ensure_bmat(h, quality, qp, &d_bmat);
        if (rc) return rc;
    }
    // persistent kernel: one CTA of kGroups groups per SM; every group takes tiles round-robin
    long long grid = (long long)h->sm_count * h->ctas_per_sm;
    long long grid_auto = grid;
    if (grid > (ntiles + kGroups - 1) / kGroups) grid = (ntiles + kGroups - 1) / kGroups;
    if (grid_auto > (ntiles + kGroupsAuto - 1) / kGroupsAuto) grid_auto = (ntiles + kGroupsAuto - 1) / kGroupsAuto;
    long long grid_c = (long long)h->sm_count * h->ctas_per_sm_c;
    if (grid_c > ntiles) grid_c = ntiles;
    long long grid_single = (long long)h->sm_count * h->ctas_per_sm_single;
    if (grid_single > ntiles) grid_single = ntiles;
    const uint32_t kflags = flags & TIC_FLAG_DEBUG_ALL_EXACT;
    const AutoTables* d_tabs = nullptr;
    h->last_launches = 6;
    if (auto_mode) {   // calc_huffman_table (huffman.py:101-109) + write_huffman_table (codec.py:73-84)
        if ((size_t)n_images > h->auto_cap) {
            cudaFree(h->d_hist); cudaFree(h->d_first); cudaFree(h->d_tabs);
            h->d_hist = nullptr; h->d_first = nullptr; h->d_tabs = nullptr; h->auto_cap = 0;
            size_t cap = (size_t)n_images + (size_t)n_images / 4 + 8;
            TIC_CUDA(h, cudaMalloc(&h->d_hist, cap * 272 * sizeof(uint32_t)));
            TIC_CUDA(h, cudaMalloc(&h->d_first, cap * 272 * sizeof(unsigned long long)));
            TIC_CUDA(h, cudaMalloc(&h->d_tabs, cap * sizeof(AutoTables)));
            h->auto_cap = cap;
        }
        TIC_CUDA(h, cudaMemsetAsync(h->d_hist, 0, (size_t)n_images * 272 * sizeof(uint32_t), stream));
        TIC_CUDA(h, cudaMemsetAsync(h->d_first, 0xff, (size_t)n_images * 272 * sizeof(unsigned long long), stream));
        if (kFdctTc && TIC_STATS_GROUPS)
            symbol_stats_groups_kernel<kGroups><<<(unsigned)grid, kTile * kGroups, kSmemEncode, stream>>>(
                qp, h->d_descs, n_images, uniform_tpi, ntiles, h->d_counters, h->d_hist, h->d_first, d_status, d_bmat);
        else
            symbol_stats_kernel<<<(unsigned)grid_single, kTile, kSmemSingle, stream>>>(
                qp, h->d_descs, n_images, uniform_tpi, ntiles, h->d_counters, h->d_hist, h->d_first, d_status, d_bmat);
        TIC_CUDA(h, cudaGetLastError());
        build_tables_kernel<<<n_images, 32, 0, stream>>>(h->d_descs, n_images, quality, (flags & TIC_FLAG_AUTO_LE_FLAG) ? 1 : 0,
                                                         h->d_hist, h->d_first, h->d_tabs, d_status);
        TIC_CUDA(h, cudaGetLastError());
        d_tabs = h->d_tabs;
        h->last_launches = 8;
        TIC_CUDA(h, cudaEventRecord(evq[0], stream));
        encode_tiles_kernel<1, kGroupsAuto><<<(unsigned)grid_auto, kTile * kGroupsAuto, kSmemEncodeAuto, stream>>>(
            qp, h->d_descs, n_images, uniform_tpi, ntiles, h->d_recs, h->d_arena, (unsigned long long)h->arena_cap16,
            h->d_counters, d_status, quality, d_tabs, d_bmat, kflags);
    } else if (c_variant) {
        TIC_CUDA(h, cudaEventRecord(evq[0], stream));
        encode_tiles_kernel<2, 1><<<(unsigned)grid_c, kTile, kSmemEncodeC, stream>>>(
            qp, h->d_descs, n_images, uniform_tpi, ntiles, h->d_recs, h->d_arena, (unsigned long long)h->arena_cap16,
            h->d_counters, d_status, quality, d_tabs, d_bmat, kflags);
    } else {
        TIC_CUDA(h, cudaEventRecord(evq[0], stream));
        encode_tiles_kernel<0, kGroups><<<(unsigned)grid, kTile * kGroups, kSmemEncode, stream>>>(
            qp, h->d_descs, n_images, uniform_tpi, ntiles, h->d_recs, h->d_arena, (unsigned long long)h->arena_cap16,
            h->d_counters, d_status, quality, d_tabs, d_bmat, kflags);
    }
    TIC_CUDA(h, cudaGetLastError());
    TIC_CUDA(h, cudaEventRecord(evq[1], stream));
    const bool small = ntiles <= kSmallScan;
    if (small) {   // one launch instead of four
        scan_small_kernel<<<1, kSmallThreads, 0, stream>>>(h->d_recs, (int)ntiles, h->d_tile_pos, (long long)out_capacity,
                                                          (long long*)d_out_offsets, h->d_out_end, (long long*)d_out_sizes,
                                                          n_images, d_status, h->d_counters, h->d_sticky);
        TIC_CUDA(h, cudaGetLastError());
        h->last_launches -= 3;
    } else {
        scan_chunks_kernel<<<(unsigned)nchunks, kScanThreads, 0, stream>>>(h->d_recs, ntiles, h->d_chunk_span, nchunks,
                                                                           h->d_chunk_pos, h->d_counters);
        TIC_CUDA(h, cudaGetLastError());
        scan_apply_kernel<<<(unsigned)nchunks, kScanThreads, 0, stream>>>(h->d_recs, ntiles, h->d_chunk_pos, h->d_tile_pos,
                                                                          (long long)out_capacity, (long long*)d_out_offsets,
                                                                          h->d_out_end, h->d_counters, n_images,
                                                                          (long long*)d_out_sizes, d_status, h->d_sticky);
        TIC_CUDA(h, cudaGetLastError());
        h->last_launches -= 2;
    }
    long long cgrid = (ntiles * 32 + kCompactThreads - 1) / kCompactThreads;
    const long long cmax = (long long)h->sm_count * 8 * 4;
    if (cgrid > cmax) cgrid = cmax;
    TIC_CUDA(h, cudaEventRecord(evq[2], stream));
    compact_kernel<<<(unsigned)cgrid, kCompactThreads, 0, stream>>>(h->d_recs, h->d_tile_pos, ntiles,
                                                                    (const uint32_t*)h->d_arena, (uint32_t*)d_out,
                                                                    (long long)out_capacity);
    TIC_CUDA(h, cudaGetLastError());
    TIC_CUDA(h, cudaEventRecord(evq[3], stream));
    h->ev_batches++;
    if (prep_on_device) h->last_launches += 1;
    h->last_tiles = ntiles;
    h->last_blocks = nblocks;
    return TIC_OK;
}

int tic_encode_finish(tic_handle h, void* stream_v, int64_t* total_bytes) {
    if (!h) return TIC_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_v;
    TIC_CUDA(h, cudaSetDevice(h->device));
    TIC_CUDA(h, cudaMemcpyAsync(h->h_counters, h->d_sticky, kCtrCount * 8, cudaMemcpyDeviceToHost, stream));
    TIC_CUDA(h, cudaMemsetAsync(h->d_sticky, 0, kCtrCount * 8, stream));   // the next finish reports the batches after this one
    TIC_CUDA(h, cudaStreamSynchronize(stream));
    // device time of the two big kernels, summed over the batches enqueued since the last finish
    // (CUDA events recorded on the batches' stream; at most the newest kEvRing batches)
    h->sum_encode_ms = h->sum_compact_ms = 0.0;
    h->timed_batches = h->ev_batches < tic_handle_s::kEvRing ? h->ev_batches : tic_handle_s::kEvRing;
    for (long long b = h->ev_batches - h->timed_batches; b < h->ev_batches; b++) {
        cudaEvent_t* q = h->ev[b % tic_handle_s::kEvRing];
        float a = 0.f, c = 0.f;
        if (cudaEventElapsedTime(&a, q[0], q[1]) == cudaSuccess) h->sum_encode_ms += a;
        if (cudaEventElapsedTime(&c, q[2], q[3]) == cudaSuccess) h->sum_compact_ms += c;
    }
    h->ev_batches = 0;
    (void)cudaGetLastError();
    if (total_bytes) *total_bytes = (int64_t)(h->h_counters[kCtrTotalBits] >> 3);
    if (h->h_counters[kCtrTcTimeout]) { h->err = "tensor-core pipeline: an MMA completion never arrived"; return TIC_E_CUDA; }
    if (h->h_counters[kCtrOverflow]) { h->err = "output buffer too small"; return TIC_E_CAPACITY; }
    if (h->h_counters[kCtrAnyStatus] & TIC_STATUS_TABLE) {
        h->err = "auto-generated Huffman table cannot be serialised (reference: OverflowError)";
        return TIC_E_TABLE;
    }
    if (h->h_counters[kCtrAnyStatus] & TIC_STATUS_LONGCODE) {
        h->err = "auto-generated Huffman code longer than 32 bits is not supported on the device";
        return TIC_E_UNSUPPORTED;
    }
    if (h->h_counters[kCtrAnyStatus] & TIC_STATUS_CATEGORY) {
        h->err = "coefficient category outside the fixed Huffman tables (reference: KeyError)";
        return TIC_E_CATEGORY;
    }
    return TIC_OK;
}

int tic_last_stats(tic_handle h, int64_t stats[8]) {
    if (!h || !stats) return TIC_E_INVALID;
    memset(stats, 0, 8 * sizeof(int64_t));
    stats[0] = h->last_launches;
    stats[1] = h->last_tiles;
    stats[2] = (int64_t)h->h_counters[kCtrExactItems];
    stats[3] = (int64_t)h->h_counters[kCtrExactChanged];
    stats[4] = h->last_blocks;
    stats[5] = (int64_t)(h->sum_encode_ms * 1.0e6);
    stats[6] = (int64_t)(h->sum_compact_ms * 1.0e6);
    stats[7] = h->timed_batches;
    return TIC_OK;
}

int64_t tic_last_guard_misses(tic_handle h) { return h ? (int64_t)h->h_counters[kCtrUnflagged] : -1; }

int tic_encode_coeffs(tic_handle h, const void* d_pixels, int32_t height, int32_t width, int32_t quality,
                      int32_t* d_dc, int32_t* d_ac, void* stream_v) {
    if (!h) return TIC_E_INVALID;
    h->err.clear();
    if (height < 0 || width < 0) { h->err = "negative image dimension"; return TIC_E_INVALID; }
    QuantParams qp;
    int rc = make_quant_params(quality, qp);
    if (rc) { h->err = "quality must be in 1..99"; return rc; }
    long long nblk = tic_num_blocks(height, width);
    if (nblk == 0) return TIC_OK;
    if (!d_pixels || !d_dc || !d_ac) { h->err = "invalid argument"; return TIC_E_INVALID; }
    if (nblk > 0x7fffffffll - kTile) { h->err = "image too large"; return TIC_E_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_v;
    TIC_CUDA(h, cudaSetDevice(h->device));
    rc = ensure_tables(h);
    if (rc) return rc;
    rc = grow_descs(h, 1);
    if (rc) return rc;
    ImageDesc& d = h->h_descs[0];
    d.px = (const uint8_t*)d_pixels;
    d.h = height; d.w = width; d.bw = (width + 7) / 8; d.nblk = (int)nblk; d.tile0 = 0;
    d.bw_shift = bw_shift_of(d.bw); d.pad = 0;
    TIC_CUDA(h, cudaMemcpyAsync(h->d_descs, h->h_descs, sizeof(ImageDesc), cudaMemcpyHostToDevice, stream));
    rc = descs_staged(h, stream);
    if (rc) return rc;
    const uint4* d_bmat = nullptr;
    rc = ensure_bmat(h, quality, qp, &d_bmat);
    if (rc) return rc;
    TIC_CUDA(h, cudaMemsetAsync(h->d_counters, 0, kCtrCount * 8, stream));
    long long ntiles = (nblk + kTile - 1) / kTile;
    coeffs_kernel<<<(unsigned)ntiles, kTile, kSmemSingle, stream>>>(qp, h->d_descs, h->d_counters, d_dc, d_ac, d_bmat, 0u);
    TIC_CUDA(h, cudaGetLastError());
    h->last_launches = 1;
    h->last_tiles = ntiles;
    h->last_blocks = nblk;
    return TIC_OK;
}

int tic_compress_host(tic_handle h, const uint8_t* pixels, int32_t height, int32_t width, int32_t quality,
                      uint32_t flags, uint8_t* out, int64_t out_capacity, int64_t* out_size, int32_t* status) {
    if (!h) return TIC_E_INVALID;
    h->err.clear();
    if (height < 0 || width < 0 || !out || !out_size || out_capacity < 16) { h->err = "invalid argument"; return TIC_E_INVALID; }
    TIC_CUDA(h, cudaSetDevice(h->device));
    const size_t npx = (size_t)height * (size_t)width;
    if (npx && !pixels) { h->err = "null pixels"; return TIC_E_INVALID; }
    const size_t need_out = (size_t)tic_max_out_bytes(height, width) + ((flags & TIC_FLAG_AUTO_HUFFMAN) ? 1664 : 0);
    if (npx > h->px_cap) {
        cudaFree(h->d_px); h->d_px = nullptr; h->px_cap = 0;
        TIC_CUDA(h, cudaMalloc(&h->d_px, npx + 16));
        h->px_cap = npx;
    }
    if (need_out > h->out_cap) {
        cudaFree(h->d_out); h->d_out = nullptr; h->out_cap = 0;
        TIC_CUDA(h, cudaMalloc(&h->d_out, need_out));
        h->out_cap = need_out;
    }
    const size_t need_stage = npx > need_out ? npx : need_out;
    if (need_stage > h->h_stage_cap) {
        cudaFreeHost(h->h_stage); h->h_stage = nullptr; h->h_stage_cap = 0;
        TIC_CUDA(h, cudaMallocHost(&h->h_stage, need_stage));
        h->h_stage_cap = need_stage;
    }
    cudaStream_t s = h->own_stream;
    if (npx) {
        memcpy(h->h_stage, pixels, npx);
        TIC_CUDA(h, cudaMemcpyAsync(h->d_px, h->h_stage, npx, cudaMemcpyHostToDevice, s));
    }
    const void* ptrs[1] = {h->d_px};
    int32_t hs[1] = {height}, ws[1] = {width};
    long long* d_off = h->d_meta;
    long long* d_size = h->d_meta + 1;
    int32_t* d_stat = (int32_t*)(h->d_meta + 2);
    int rc = tic_encode_batch(h, ptrs, hs, ws, 1, quality, flags, h->d_out, (int64_t)h->out_cap, (int64_t*)d_off,
                              (int64_t*)d_size, d_stat, s);
    if (rc) return rc;
    TIC_CUDA(h, cudaMemcpyAsync(h->h_meta, h->d_meta, 24, cudaMemcpyDeviceToHost, s));
    int64_t total = 0;
    rc = tic_encode_finish(h, s, &total);
    if (status) *status = *(int32_t*)(h->h_meta + 2);
    if (rc) return rc;
    const int64_t size = h->h_meta[1];
    *out_size = size;
    if (size > out_capacity) { h->err = "host output buffer too small"; return TIC_E_CAPACITY; }
    TIC_CUDA(h, cudaMemcpyAsync(h->h_stage, h->d_out + h->h_meta[0], (size_t)size, cudaMemcpyDeviceToHost, s));
    TIC_CUDA(h, cudaStreamSynchronize(s));
    memcpy(out, h->h_stage, (size_t)size);
    return TIC_OK;
}

}  // extern "C"
