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

// ---------------------------------------------------------------------------------------------
// compress(): one CTA per tile, tiles taken in stream order through a ticket so that the
// decoupled look-back can never wait on a tile that has not started.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile, 8)
encode_tiles_kernel(const __grid_constant__ QuantParams qp, const ImageDesc* __restrict__ descs, int n_images,
                    int uniform_tpi, long long ntiles, unsigned long long* __restrict__ tile_status,
                    unsigned long long* __restrict__ tile_tail, unsigned long long* __restrict__ counters,
                    uint8_t* __restrict__ out, long long out_cap, long long* __restrict__ out_off,
                    long long* __restrict__ out_end, int* __restrict__ status, int quality,
                    const AutoTables* __restrict__ auto_tabs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared& sm = *reinterpret_cast<TileShared*>(smem_raw);
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int bias = qp.qbias;

    // default tables -> shared memory, once per (persistent) CTA (constants.py:53-242)
    for (int i = t; i < 256; i += kTile)
        sm.ac_tab[i] = make_uint2(c_default_tables.ac[i].code, c_default_tables.ac[i].len);
    if (t < 16) sm.dc_tab[t] = make_uint2(c_default_tables.dc[t].code, c_default_tables.dc[t].len);

    int tab_img = -1;   // auto mode: image whose tables are in shared memory
  for (;;) {   // persistent: tiles are claimed in stream order through the ticket
    __syncthreads();   // previous tile fully copied out; tables visible
    if (warp == 0) {   // claim the next tile and find its image
        long long tk = 0;
        if (lane == 0) tk = (long long)atomicAdd(&counters[kCtrTicket], 1ull);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk < ntiles) {
            const TileInfo f = locate_tile(descs, n_images, tk, uniform_tpi);
            if (lane == 0) sm.ti = f;
        }
        if (lane == 0) { sm.tile = tk; sm.err = 0; }
    }
    __syncthreads();
    const long long tile = sm.tile;
    if (tile >= ntiles) break;
    const TileInfo ti = sm.ti;
    if (auto_tabs != nullptr && tab_img != ti.img) {   // per-image tables (codec.py:146-148); uniform branch
        const HuffTables& g = auto_tabs[ti.img].tab;
        for (int i = t; i < 256; i += kTile) sm.ac_tab[i] = make_uint2(g.ac[i].code, g.ac[i].len);
        if (t < 16) sm.dc_tab[t] = make_uint2(g.dc[t].code, g.dc[t].len);
        tab_img = ti.img;   // visible to everyone after the barriers inside transform_tile
    }

    transform_tile(ti, qp, sm, counters);

    // ---- bit lengths and the CTA scan -------------------------------------------------------
    int err = 0;
    int bits = (t < ti.nb) ? block_bits(sm, t, bias, err) : 0;
    int incl = bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) sm.warp_bits[warp] = incl;
    if (err) sm.err = 1;
    __syncthreads();
    int warp_base = 0, tile_bits = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        int wb = sm.warp_bits[w];
        if (w < warp) warp_base += wb;
        tile_bits += wb;
    }
    // the first tile of an image carries the header in front of its blocks: 128 bits with the fixed
    // tables (codec.py:102-114), 128 + the serialised tables in auto mode (codec.py:110-112)
    const int hdr_bits = !ti.first ? 0 : (auto_tabs ? (int)auto_tabs[ti.img].hdr_bits : 128);
    const int bitpos = hdr_bits + warp_base + incl - bits;   // tile-relative bit offset of this block
    tile_bits += hdr_bits;
    const long long agg = (long long)tile_bits;

    // publish the aggregate as early as possible: successors only need it for their offset
    if (t == 0) {
        st_relaxed_u64(&tile_status[tile], kFlagAgg | (ti.closing ? kClosingBit : 0ull) | (unsigned long long)agg);
        if (sm.err) atomicOr(&status[ti.img], TIC_STATUS_CATEGORY);
    }
    const int nwords = (tile_bits + 31) >> 5;       // tile-relative words holding data
    const int hdr_words = (hdr_bits + 31) >> 5;
    const int rounds = nwords > kWinWords ? (nwords + kWinWords - 1) / kWinWords : 1;
    uint32_t* out_words = reinterpret_cast<uint32_t*>(out);
    long long s_bits = 0, e_bits = 0, g0 = 0, g_end = 0;
    int sh = 0;
    bool fits = true;
    unsigned int v_first = 0, tail = 0;             // thread 0: word g0 without its head; trailing partial word

    for (int r = 0; r < rounds; r++) {
        const int wbase = r * kWinWords;
        const int nw = nwords - wbase < kWinWords ? nwords - wbase : kWinWords;   // data words in this window
        for (int i = t; i <= nw; i += kTile) {       // clear the window (+1 word read by the funnel shift)
            uint32_t w = 0;
            if (wbase + i < hdr_words) {
                if (auto_tabs) {
                    w = auto_tabs[ti.img].hdr_words[wbase + i];
                } else {   // struct.pack("III") is little-endian, the stream is MSB-first; flag word 0
                    const uint32_t v = i == 0 ? (uint32_t)ti.h : (i == 1 ? (uint32_t)ti.w : (i == 2 ? (uint32_t)quality : 0u));
                    w = __byte_perm(v, 0, 0x0123);
                }
            }
            sm.stage[i] = w;
        }
        __syncthreads();

        // ---- bits into the window of the tile-relative staging buffer ----------------------------
        if (t < ti.nb && bitpos + bits > wbase * 32 && bitpos < (wbase + kWinWords) * 32)
            block_emit(sm, t, bias, bitpos, wbase);

        // ---- decoupled look-back (warp 0, 32 predecessors per round): absolute bit position -------
        if (r == 0 && warp == 0) {
            // composite of the tiles between the look-back cursor and this tile:
            //   g(P) = closed ? round_up128(P + a) + b : P + a
            long long a = 0, b = 0;
            bool closed = false;
            long long p_in = 0;
            long long j = tile - 1;   // nearest predecessor not folded in yet
            while (true) {
                const long long idx = j - lane;
                unsigned long long sw = kFlagPrefix;   // before the first tile: prefix 0
                if (idx >= 0) sw = ld_relaxed_u64(&tile_status[idx]);
                while (__any_sync(0xffffffffu, (sw & kFlagMask) == 0)) {
                    if ((sw & kFlagMask) == 0) sw = ld_relaxed_u64(&tile_status[idx]);
                }
                const unsigned prefix_mask = __ballot_sync(0xffffffffu, (sw & kFlagMask) == kFlagPrefix);
                const int p = prefix_mask ? (__ffs(prefix_mask) - 1) : 32;   // lanes < p hold aggregates
                const unsigned below = p >= 32 ? 0xffffffffu : ((1u << p) - 1u);
                const unsigned closing_mask = __ballot_sync(0xffffffffu, (sw & kClosingBit) != 0) & below;
                const long long val = (long long)(sw & kValueMask);
                if (closing_mask == 0) {
                    long long v = (lane < p) ? val : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    a += v;
                } else {
                    for (int l = 0; l < p; l++) {   // nearest first: prepend tile j-l to the composite
                        long long v = __shfl_sync(0xffffffffu, val, l);
                        if ((closing_mask >> l) & 1) {   // it closes an image: what follows is 128-bit aligned
                            b = closed ? round_up128(a) + b : a;
                            a = v;
                            closed = true;
                        } else {
                            a += v;
                        }
                    }
                }
                if (p < 32) {
                    long long base = __shfl_sync(0xffffffffu, val, p);
                    p_in = closed ? round_up128(base + a) + b : base + a;
                    break;
                }
                j -= 32;
            }
            if (lane == 0) {
                const long long eb = p_in + agg;   // end of this tile's data bits
                const long long p_out = ti.closing ? round_up128(eb) : eb;
                st_relaxed_u64(&tile_status[tile], kFlagPrefix | (unsigned long long)p_out);
                sm.s_bits = p_in;   // the header (first tile) sits in the staging buffer too
                const long long end_byte = (eb + 7) >> 3;
                if (((end_byte + 3) & ~3ll) > out_cap) atomicExch(&counters[kCtrOverflow], 1ull);
                if (ti.first) out_off[ti.img] = p_in >> 3;
                if (ti.closing) {
                    out_end[ti.img] = end_byte;
                    atomicMax(&counters[kCtrTotalBits], (unsigned long long)(end_byte << 3));
                }
            }
        }
        __syncthreads();

        // ---- copy-out: funnel shift to the global alignment, byte-swap to MSB-first byte order ----
        if (r == 0) {
            s_bits = sm.s_bits;
            e_bits = s_bits + tile_bits;
            sh = (int)(s_bits & 31);
            g0 = s_bits >> 5;
            // words [g0, g_end): full words, plus the final partial word when this tile closes the image
            g_end = ti.closing ? ((e_bits + 31) >> 5) : (e_bits >> 5);
            fits = ((((e_bits + 7) >> 3) + 3) & ~3ll) <= out_cap;
        }
        const int nout = (int)(g_end - g0);                      // output words of this tile
        const int j_hi = wbase + kWinWords < nout ? wbase + kWinWords : nout;
        const unsigned int carry = r ? sm.carry : 0u;            // tile-relative word wbase-1
        for (int jdx = wbase + t; jdx < j_hi; jdx += kTile) {
            const int i = jdx - wbase;
            const uint32_t prev = i ? sm.stage[i - 1] : carry;
            const uint32_t v = __funnelshift_r(sm.stage[i], prev, sh);
            if (jdx == 0) v_first = v;                            // word g0: written last, with its head
            else if (fits) out_words[g0 + jdx] = __byte_perm(v, 0, 0x0123);
        }
        if (t == 0 && r == rounds - 1 && (e_bits & 31)) {         // trailing partial word for the next tile
            const int i = (int)((e_bits >> 5) - g0) - wbase;
            const uint32_t prev = i ? sm.stage[i - 1] : carry;
            tail = __funnelshift_r(sm.stage[i], prev, sh);
        }
        if (r + 1 < rounds) {                                     // multi-round tiles only
            __syncthreads();
            if (t == 0) sm.carry = sm.stage[kWinWords - 1];
            __syncthreads();
        }
    }
    if (t == 0) {
        // Hand the trailing partial word to the next tile FIRST (it only depends on the previous
        // tile's tail when this whole tile sits inside one word), then wait for our own head.
        const bool need_prev = !ti.first && sh != 0;   // word g0 starts with the previous tile's last bits
        const bool chained = need_prev && (e_bits >> 5) == g0;
        if (!ti.closing && !chained) st_relaxed_u64(&tile_tail[tile], (1ull << 63) | tail);
        unsigned int tail_prev = 0;
        if (need_prev) {
            unsigned long long tw;
            do { tw = ld_relaxed_u64(&tile_tail[tile - 1]); } while ((tw >> 63) == 0);
            tail_prev = (unsigned int)tw;
        }
        if (!ti.closing && chained) st_relaxed_u64(&tile_tail[tile], (1ull << 63) | tail | tail_prev);
        if (g0 < g_end && fits) out_words[g0] = __byte_perm(v_first | tail_prev, 0, 0x0123);
    }
  }   // persistent loop
}

// ---------------------------------------------------------------------------------------------
// auto_generate_huffman_table=True (codec.py:146-148): symbol statistics, then the tables
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile, 8)
symbol_stats_kernel(const __grid_constant__ QuantParams qp, const ImageDesc* __restrict__ descs, int n_images,
                    int uniform_tpi, long long ntiles, unsigned long long* __restrict__ counters,
                    uint32_t* __restrict__ g_hist, unsigned long long* __restrict__ g_first,
                    int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared& sm = *reinterpret_cast<TileShared*>(smem_raw);
    const int t = threadIdx.x;
    uint32_t* hist = sm.stage;                                                       // 272 counters
    unsigned long long* first = reinterpret_cast<unsigned long long*>(sm.stage + 512); // 272 keys
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        if (t < 32) {
            const TileInfo f = locate_tile(descs, n_images, tile, uniform_tpi);
            if (t == 0) { sm.ti = f; sm.err = 0; }
        }
        for (int i = t; i < 272; i += kTile) { hist[i] = 0; first[i] = ~0ull; }
        __syncthreads();
        const TileInfo ti = sm.ti;
        transform_tile(ti, qp, sm, counters);
        int err = 0;
        if (t < ti.nb) block_stats(sm, t, qp.qbias, (unsigned long long)(ti.blk0 + t), hist, first, err);
        if (err) sm.err = 1;
        __syncthreads();
        for (int i = t; i < 272; i += kTile) {
            uint32_t c = hist[i] + (i == 0 ? (uint32_t)ti.nb : 0u);   // one EOB per block (huffman.py:33)
            if (c) {
                atomicAdd(&g_hist[(size_t)ti.img * 272 + i], c);
                atomicMin(&g_first[(size_t)ti.img * 272 + i], first[i]);
            }
        }
        if (t == 0 && sm.err) atomicOr(&status[ti.img], TIC_STATUS_TABLE);
    }
}

// One thread per image: HuffmanTree (huffman.py:112-194) on CPython's heapq, which
// queue.PriorityQueue uses — leaves pushed in first-occurrence order, nodes compared by frequency
// only, DFS with left = "0" — then write_huffman_table (codec.py:73-84) into the header words.
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
    __device__ void put(uint32_t v, int n) {   // n <= 32, MSB-first
        for (int i = n - 1; i >= 0; i--) {
            if (nbits < kMaxHdrWords * 32 && ((v >> i) & 1)) words[nbits >> 5] |= 0x80000000u >> (nbits & 31);
            nbits++;
        }
    }
};

// Builds one alphabet: symbols base..base+count-1 of hist/first.  Returns status bits.
__device__ int build_alphabet(TreeScratch& s, const uint32_t* hist, const unsigned long long* first, int base,
                              int count, bool is_dc, HuffEntry* table, HdrWriter& hw) {
    // present symbols in first-occurrence order (dict insertion order, huffman.py:187-194)
    short order[256];
    int n = 0;
    for (int i = 0; i < count; i++) {
        if (hist[base + i] == 0) continue;
        int j = n++;
        while (j > 0 && first[base + order[j - 1]] > first[base + i]) { order[j] = order[j - 1]; j--; }
        order[j] = (short)i;
    }
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

__global__ void build_tables_kernel(const ImageDesc* __restrict__ descs, int n_images, int quality,
                                    const uint32_t* __restrict__ g_hist,
                                    const unsigned long long* __restrict__ g_first,
                                    AutoTables* __restrict__ tabs, TreeScratch* __restrict__ scratch,
                                    int* __restrict__ status) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= n_images) return;
    AutoTables& at = tabs[img];
    TreeScratch& s = scratch[img];
    for (int i = 0; i < 256; i++) at.tab.ac[i].code = at.tab.ac[i].len = 0;
    for (int i = 0; i < 16; i++) at.tab.dc[i].code = at.tab.dc[i].len = 0;
    for (int i = 0; i < kMaxHdrWords; i++) at.hdr_words[i] = 0;
    HdrWriter hw{at.hdr_words, 0};
    const ImageDesc d = descs[img];
    hw.put(__byte_perm((uint32_t)d.h, 0, 0x0123), 32);           // struct.pack("III"), codec.py:103-109
    hw.put(__byte_perm((uint32_t)d.w, 0, 0x0123), 32);
    hw.put(__byte_perm((uint32_t)quality, 0, 0x0123), 32);
    hw.put(0x80000000u, 32);                                     // codec.py:111
    const uint32_t* hist = g_hist + (size_t)img * 272;
    const unsigned long long* first = g_first + (size_t)img * 272;
    int st = build_alphabet(s, hist, first, 256, 16, true, at.tab.dc, hw);    // table[DC] first, codec.py:74-78
    st |= build_alphabet(s, hist, first, 0, 256, false, at.tab.ac, hw);        // then table[AC], codec.py:79-84
    if (hw.nbits > kMaxHdrWords * 32) st |= TIC_STATUS_LONGCODE;
    at.hdr_bits = (uint32_t)hw.nbits;
    at.status = (uint32_t)st;
    if (st) atomicOr(&status[img], st);
}

// sizes = end - offset, and the batch summary the host reads back in tic_encode_finish
__global__ void finalize_kernel(int n_images, const long long* __restrict__ out_off,
                                const long long* __restrict__ out_end, long long* __restrict__ out_sizes,
                                const int* __restrict__ status, unsigned long long* __restrict__ counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_images) {
        out_sizes[i] = out_end[i] - out_off[i];
        if (status[i]) atomicOr(&counters[kCtrAnyStatus], (unsigned long long)status[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// encode(): the same transform, coefficients written out (parity checkpoint for codec.py:26-43)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile, 4)
coeffs_kernel(const __grid_constant__ QuantParams qp, const ImageDesc* __restrict__ descs,
              unsigned long long* __restrict__ counters, int* __restrict__ dc, int* __restrict__ ac) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared& sm = *reinterpret_cast<TileShared*>(smem_raw);
    const int t = threadIdx.x;
    if (t < 32) {
        const TileInfo f = locate_tile(descs, 1, blockIdx.x, 0);
        if (t == 0) sm.ti = f;
    }
    __syncthreads();
    const TileInfo ti = sm.ti;
    transform_tile(ti, qp, sm, counters);
    if (t < ti.nb) {
        const size_t b = (size_t)ti.blk0 + t;
        dc[b] = sm.dcq[t + 1] - sm.dcq[t];
        int* row = ac + b * 63;
        for (int k = 1; k < 64; k++) row[k - 1] = coef_get(sm, t, k, qp.qbias);
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
    ImageDesc* h_descs = nullptr;       // pinned
    unsigned long long* d_tile_status = nullptr; size_t tiles_cap = 0;   // status + tail, 2 * tiles_cap
    unsigned long long* d_counters = nullptr;
    unsigned long long* h_counters = nullptr;    // pinned
    long long* d_out_end = nullptr;     size_t end_cap = 0;
    // auto-table mode: per-image statistics, tables, tree scratch
    uint32_t* d_hist = nullptr; unsigned long long* d_first = nullptr; AutoTables* d_tabs = nullptr;
    TreeScratch* d_tree = nullptr;      size_t auto_cap = 0;
    // single-image host path
    uint8_t* d_px = nullptr;            size_t px_cap = 0;
    uint8_t* d_out = nullptr;           size_t out_cap = 0;
    uint8_t* h_stage = nullptr;         size_t h_stage_cap = 0;          // pinned
    long long* d_meta = nullptr;        // off, size, status for the single-image path
    long long* h_meta = nullptr;        // pinned
    cudaStream_t own_stream = nullptr;
    long long last_tiles = 0, last_blocks = 0, last_launches = 0;
    bool tables_ready = false;
    int sm_count = 148, ctas_per_sm = 4;
};

#define TIC_CUDA(h, call)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return TIC_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

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
    double qt_min = 1e30;
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
        if (q < qt_min) qt_min = q;
    }
    // |coefficient| <= 1024 (orthonormal transform of 64 values in [-128,127]).  The fixed-point
    // value round(t*2^F) + 2^(F-1) + 2^(k-1) must stay inside the +-2^22 window of the
    // magic-number rounding; F is chosen with a factor 2 to spare.
    const double t_max = 1024.0 / qt_min;
    int F = 0;
    while (F < 15 && (t_max + 2.0) * (double)(1 << (F + 1)) < 4194304.0 * 0.98) F++;
    qp.fbits = F;
    qp.qbias = 0x4B400000 >> F;
    qp.pad[0] = qp.pad[1] = 0;
    const int fmask = (1 << F) - 1;
    double aan[8];
    aan[0] = 1.0;
    for (int k = 1; k < 8; k++) aan[k] = cos(k * 3.14159265358979323846 / 16.0) * sqrt(2.0);
    // kFastErr: bound on |FP32 AAN coefficient - float64 reference coefficient| in coefficient
    // units ("tie guard").  A coefficient can only be rounded differently from the
    // reference if a .5 tie lies within tol = kFastErr/qt (+ the FP32 multiplier's relative error)
    // of the fast value; the window [-2^(k-1), 2^(k-1)) in 2^-F units around every tie covers
    // tol*2^F + 1 (the +1: round-to-integer of the FFMA and the one-sided window).
    const double kFastErr = 6.0e-4;
    for (int u = 0; u < 8; u++)
        for (int v = 0; v < 8; v++) {
            int i = u * 8 + v;
            double m = (double)(1 << F) / (8.0 * aan[u] * aan[v] * qp.qt[i]);
            qp.qmul[i] = (float)m;
            double tol = kFastErr / qp.qt[i] + 2.4e-7 * (1024.0 / qp.qt[i]);
            int g = (int)ceil(tol * (double)(1 << F) + 1.0);
            int k = 1;
            while ((1 << (k - 1)) < g + 1 && k <= F) k++;
            if (k > F) k = F;   // whole range: every coefficient goes to the exact path
            qp.gmask[i] = F == 0 ? 0 : (fmask & ~((1 << k) - 1));
            qp.magic[i] = 12582912.0f + (F > 0 ? (float)(1 << (F - 1)) : 0.0f) + (F > 0 ? (float)(1 << (k - 1)) : 0.0f);
        }
    return TIC_OK;
}

static int ensure_tables(tic_handle h) {
    if (h->tables_ready) return TIC_OK;
    HuffTables t;
    build_default_tables(t);
    TIC_CUDA(h, cudaMemcpyToSymbol(c_default_tables, &t, sizeof t));
    TIC_CUDA(h, cudaFuncSetAttribute(encode_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(TileShared)));
    TIC_CUDA(h, cudaFuncSetAttribute(coeffs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(TileShared)));
    TIC_CUDA(h, cudaFuncSetAttribute(symbol_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(TileShared)));
    int per_sm = 0;
    TIC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_tiles_kernel, kTile, sizeof(TileShared)));
    cudaDeviceProp prop;
    TIC_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    h->ctas_per_sm = per_sm > 0 ? per_sm : 1;
    h->tables_ready = true;
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
        cudaMallocHost(&h->h_counters, kCtrCount * 8) != cudaSuccess ||
        cudaMalloc(&h->d_meta, 64) != cudaSuccess || cudaMallocHost(&h->h_meta, 64) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return TIC_E_CUDA;
    }
    *out = h;
    return TIC_OK;
}

int tic_destroy(tic_handle h) {
    if (!h) return TIC_E_INVALID;
    cudaSetDevice(h->device);
    cudaFree(h->d_descs); cudaFreeHost(h->h_descs); cudaFree(h->d_tile_status); cudaFree(h->d_counters);
    cudaFree(h->d_hist); cudaFree(h->d_first); cudaFree(h->d_tabs); cudaFree(h->d_tree);
    cudaFreeHost(h->h_counters); cudaFree(h->d_out_end); cudaFree(h->d_px); cudaFree(h->d_out);
    cudaFreeHost(h->h_stage); cudaFree(h->d_meta); cudaFreeHost(h->h_meta);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return TIC_OK;
}

const char* tic_last_error(tic_handle h) { return h ? h->err.c_str() : "null handle"; }

static int grow_descs(tic_handle h, size_t n) {
    if (n <= h->descs_cap) return TIC_OK;
    size_t cap = n < 64 ? 64 : n * 2;
    cudaFree(h->d_descs); cudaFreeHost(h->h_descs);
    h->d_descs = nullptr; h->h_descs = nullptr; h->descs_cap = 0;
    TIC_CUDA(h, cudaMalloc(&h->d_descs, cap * sizeof(ImageDesc)));
    TIC_CUDA(h, cudaMallocHost(&h->h_descs, cap * sizeof(ImageDesc)));
    h->descs_cap = cap;
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
    if ((reinterpret_cast<uintptr_t>(d_out) & 15) != 0) {
        h->err = "d_out must be 16-byte aligned";
        return TIC_E_INVALID;
    }
    QuantParams qp;
    int rc = make_quant_params(quality, qp);
    if (rc) { h->err = "quality must be in 1..99"; return rc; }
    cudaStream_t stream = (cudaStream_t)stream_v;
    TIC_CUDA(h, cudaSetDevice(h->device));
    rc = ensure_tables(h);
    if (rc) return rc;
    h->last_tiles = h->last_blocks = h->last_launches = 0;
    if (n_images == 0) {
        TIC_CUDA(h, cudaMemsetAsync(h->d_counters, 0, kCtrCount * 8, stream));
        return TIC_OK;
    }
    rc = grow_descs(h, (size_t)n_images);
    if (rc) return rc;
    long long ntiles = 0, nblocks = 0;
    int uniform_tpi = 0;
    for (int i = 0; i < n_images; i++) {
        if (heights[i] < 0 || widths[i] < 0) { h->err = "negative image dimension"; return TIC_E_INVALID; }
        long long nblk = tic_num_blocks(heights[i], widths[i]);
        if (nblk > 0x7fffffffll - kTile) { h->err = "image too large"; return TIC_E_INVALID; }
        ImageDesc& d = h->h_descs[i];
        d.px = (const uint8_t*)d_pixels[i];
        d.h = heights[i];
        d.w = widths[i];
        d.bw = (widths[i] + 7) / 8;
        d.nblk = (int)nblk;
        d.tile0 = ntiles;
        long long nt = (nblk + kTile - 1) / kTile;
        if (nt < 1) nt = 1;          // an empty image still owns one tile: it writes the header
        if (i == 0) uniform_tpi = (int)nt; else if (nt != uniform_tpi) uniform_tpi = -1;
        ntiles += nt;
        nblocks += nblk;
        if (nblk > 0 && !d.px) { h->err = "null pixel pointer"; return TIC_E_INVALID; }
    }
    if (uniform_tpi < 0) uniform_tpi = 0;
    if (ntiles > 0x7fffffffll) { h->err = "batch too large for one launch"; return TIC_E_INVALID; }
    if ((size_t)ntiles > h->tiles_cap) {
        cudaFree(h->d_tile_status);
        h->d_tile_status = nullptr; h->tiles_cap = 0;
        size_t cap = (size_t)ntiles + (size_t)ntiles / 4 + 1024;
        TIC_CUDA(h, cudaMalloc(&h->d_tile_status, cap * 16));
        h->tiles_cap = cap;
    }
    if ((size_t)n_images > h->end_cap) {
        cudaFree(h->d_out_end);
        h->d_out_end = nullptr; h->end_cap = 0;
        TIC_CUDA(h, cudaMalloc(&h->d_out_end, (size_t)n_images * 2 * 8));
        h->end_cap = (size_t)n_images * 2;
    }
    unsigned long long* d_status_words = h->d_tile_status;
    unsigned long long* d_tail_words = h->d_tile_status + h->tiles_cap;
    TIC_CUDA(h, cudaMemcpyAsync(h->d_descs, h->h_descs, (size_t)n_images * sizeof(ImageDesc),
                                cudaMemcpyHostToDevice, stream));
    TIC_CUDA(h, cudaMemsetAsync(d_status_words, 0, (size_t)ntiles * 8, stream));
    TIC_CUDA(h, cudaMemsetAsync(d_tail_words, 0, (size_t)ntiles * 8, stream));
    TIC_CUDA(h, cudaMemsetAsync(h->d_counters, 0, kCtrCount * 8, stream));
    TIC_CUDA(h, cudaMemsetAsync(d_status, 0, (size_t)n_images * 4, stream));
    long long grid = (long long)h->sm_count * h->ctas_per_sm;
    if (grid > ntiles) grid = ntiles;
    const AutoTables* d_tabs = nullptr;
    h->last_launches = 2;
    if (auto_mode) {   // calc_huffman_table (huffman.py:101-109) + write_huffman_table (codec.py:73-84)
        if ((size_t)n_images > h->auto_cap) {
            cudaFree(h->d_hist); cudaFree(h->d_first); cudaFree(h->d_tabs); cudaFree(h->d_tree);
            h->d_hist = nullptr; h->d_first = nullptr; h->d_tabs = nullptr; h->d_tree = nullptr; h->auto_cap = 0;
            size_t cap = (size_t)n_images + (size_t)n_images / 4 + 8;
            TIC_CUDA(h, cudaMalloc(&h->d_hist, cap * 272 * sizeof(uint32_t)));
            TIC_CUDA(h, cudaMalloc(&h->d_first, cap * 272 * sizeof(unsigned long long)));
            TIC_CUDA(h, cudaMalloc(&h->d_tabs, cap * sizeof(AutoTables)));
            TIC_CUDA(h, cudaMalloc(&h->d_tree, cap * sizeof(TreeScratch)));
            h->auto_cap = cap;
        }
        TIC_CUDA(h, cudaMemsetAsync(h->d_hist, 0, (size_t)n_images * 272 * sizeof(uint32_t), stream));
        TIC_CUDA(h, cudaMemsetAsync(h->d_first, 0xff, (size_t)n_images * 272 * sizeof(unsigned long long), stream));
        symbol_stats_kernel<<<(unsigned)grid, kTile, sizeof(TileShared), stream>>>(
            qp, h->d_descs, n_images, uniform_tpi, ntiles, h->d_counters, h->d_hist, h->d_first, d_status);
        TIC_CUDA(h, cudaGetLastError());
        build_tables_kernel<<<(n_images + 31) / 32, 32, 0, stream>>>(h->d_descs, n_images, quality, h->d_hist,
                                                                    h->d_first, h->d_tabs, h->d_tree, d_status);
        TIC_CUDA(h, cudaGetLastError());
        d_tabs = h->d_tabs;
        h->last_launches = 4;
    }
    encode_tiles_kernel<<<(unsigned)grid, kTile, sizeof(TileShared), stream>>>(
        qp, h->d_descs, n_images, uniform_tpi, ntiles, d_status_words, d_tail_words, h->d_counters, (uint8_t*)d_out,
        (long long)out_capacity, (long long*)d_out_offsets, h->d_out_end, d_status, quality, d_tabs);
    TIC_CUDA(h, cudaGetLastError());
    finalize_kernel<<<(n_images + 255) / 256, 256, 0, stream>>>(n_images, (const long long*)d_out_offsets,
                                                               h->d_out_end, (long long*)d_out_sizes, d_status,
                                                               h->d_counters);
    TIC_CUDA(h, cudaGetLastError());
    h->last_tiles = ntiles;
    h->last_blocks = nblocks;
    return TIC_OK;
}

int tic_encode_finish(tic_handle h, void* stream_v, int64_t* total_bytes) {
    if (!h) return TIC_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_v;
    TIC_CUDA(h, cudaSetDevice(h->device));
    TIC_CUDA(h, cudaMemcpyAsync(h->h_counters, h->d_counters, kCtrCount * 8, cudaMemcpyDeviceToHost, stream));
    TIC_CUDA(h, cudaStreamSynchronize(stream));
    if (total_bytes) *total_bytes = (int64_t)(h->h_counters[kCtrTotalBits] >> 3);
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
    return TIC_OK;
}

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
    TIC_CUDA(h, cudaMemcpyAsync(h->d_descs, h->h_descs, sizeof(ImageDesc), cudaMemcpyHostToDevice, stream));
    TIC_CUDA(h, cudaMemsetAsync(h->d_counters, 0, kCtrCount * 8, stream));
    long long ntiles = (nblk + kTile - 1) / kTile;
    coeffs_kernel<<<(unsigned)ntiles, kTile, sizeof(TileShared), stream>>>(qp, h->d_descs, h->d_counters, d_dc, d_ac);
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
