// tic_decode.cu — the decode side of libtinyimgcodec_cuda.so for B200 (sm_100a): SURVEY.md §8(f)3.
//
// Replaces, for a batch of .img streams resident in device memory,
//   tinyimgcodec.codec.decompress   tinyimgcodec/codec.py:167-189   (header, per-block Huffman decode)
//   tinyimgcodec.codec.decode       tinyimgcodec/codec.py:46-70     (DC cumsum, de-zigzag, dequantise,
//                                                                    IDCT, +128, clip, crop, uint8)
//   parse_header / read_huffman_table   codec.py:117-130 / :87-99
//   decode_huffman / read_huffman_code / decode_run_length   tinyimgcodec/huffman.py:77-98 / :66-74 / :36-38
//   BitBuffer.read / read_uint / read_int   tinyimgcodec/bitbuffer.py:20-23 / :42-45 / :56-66
//   block_idct, block_quantize(inverse=True), block_combine   tinyimgcodec/utils.py:40-45, :51-52, :23-29
//
// The stream has no markers, no restart intervals and no per-block lengths: block i can only be found by
// decoding blocks 0..i-1.  The reference therefore decodes serially.  Here every stream is cut into
// subsequences of kSubBits bits, one thread per subsequence, and the threads find their true entry
// state by SELF-SYNCHRONISATION: a Huffman decoder started at a wrong bit falls back into step with the
// true symbol sequence after a few symbols, so
//   1. every thread decodes its subsequence from a guessed entry state (bit 0, "a DC symbol is next") and
//      publishes the state in which it leaves (overshoot into the next subsequence, zigzag index);
//   2. a thread whose published entry differs from the one it used decodes again; this repeats (inside a
//      CTA through __syncthreads_or, across CTAs through relaunches) until nothing changes.  The first
//      subsequence of a stream has a KNOWN entry, so by induction the fixed point is the serial parse;
//   3. an exclusive scan over the subsequences' block counts and DC-difference sums gives each thread
//      the index of its first block and the running DC predictor (np.cumsum, codec.py:53);
//   4. the threads walk the parse once more, warp-synchronously, and the warp stores every completed block's
//      coefficients (int16, raster order) with coalesced 128-byte stores;
//   5. one thread per 8x8 block dequantises and runs the inverse DCT in FP32 with a proven error bound; the
//      blocks with a pixel too close to an integer for FP32 to decide are transformed again in float64,
//      operation for operation what scipy.fftpack.idct (ducc0) executes — so pixels are bit-identical to the
//      reference decoder's, including values that land on an integer boundary.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/tinyimgcodec_cuda.h"
#include "tic_tables.h"

// hooks into the handle (defined in tic_encode.cu)
void** tic_internal_dec_slot(tic_handle h);
void tic_internal_set_error(tic_handle h, const std::string& msg);
int tic_internal_device(tic_handle h);
cudaStream_t tic_internal_own_stream(tic_handle h);

namespace ticd {

#ifndef TICD_SUB_BITS
#define TICD_SUB_BITS 1024
#endif
#ifndef TICD_FAST_MIN_CTAS
#define TICD_FAST_MIN_CTAS 6
#endif
constexpr int kSubBits = TICD_SUB_BITS;   // bits per subsequence (one thread each)
constexpr int kMaxNodes = 1024;     // trie nodes per table (a Huffman tree over <= 256 symbols has <= 255)
constexpr int kMaxSymbols = 4096;   // symbols decoded per subsequence at most (zero-length codes)
#ifndef TICD_SYNC_THREADS
#define TICD_SYNC_THREADS 128
#endif
constexpr int kSyncThreads = TICD_SYNC_THREADS;   // subsequences (= threads) per CTA of the symbol passes
constexpr int kSyncIters = 64;      // in-CTA repair rounds per launch
constexpr uint32_t kLeaf = 0x8000u;
constexpr int kMulfStride = 128;    // FP32 multipliers per image: [u*8+v] then the transposed copy [v*8+u]

// Decode table of one alphabet: the reference's `table.inverse` dict (huffman.py:83) as a binary trie
// (child[n][bit]: 0 = no code, kLeaf | symbol, else node index) plus a first-level table over the next
// 8 bits (0 = no code, kLeaf | len << 8 | symbol for codes of <= 8 bits, else the node reached).
struct DecTable {
    uint16_t lut[256];
    uint16_t child[kMaxNodes][2];
};
struct DecTables {
    DecTable dc, ac;
};

struct DecImage {
    // filled by the host
    const uint32_t* words;   // stream, 4-byte aligned
    long long nbits;         // 8 * size
    long long sub_first;     // index of the stream's first subsequence in the batch
    long long blk_first;     // index of the image's first block in the batch
    uint8_t* pixels;
    int height, width;       // as the caller expects them (checked against the header)
    int nsubs, bw, nblk;
    // filled by dec_setup_kernel
    uint32_t quality, flag;
    int mode;                // 0 fixed tables, 1 per-image tables, 2 scaled integer DCT (flag bit 30)
    int skip_entropy;        // nothing to decode (error, no blocks, or every block is zero bits long)
    int skip_pixels;         // header unusable: pixels are not written
    int anchor_sub;          // subsequence in which the first block starts
    int has_tables;
    double two_q;            // 2**quality for mode 2 (codec.py:61)
};

enum { MODE_FIXED = 0, MODE_TABLES = 1, MODE_SCALED = 2 };

__constant__ uint8_t c_zigzag[64] = {TIC_ZIGZAG_LIST};
// tinyimgcodec/constants.py:37-51: ANNSCALES = these / 2048
__device__ const double d_ann[64] = {
    16384 / 2048.0, 22725 / 2048.0, 21407 / 2048.0, 19266 / 2048.0, 16384 / 2048.0, 12873 / 2048.0, 8867 / 2048.0,  4520 / 2048.0,
    22725 / 2048.0, 31521 / 2048.0, 29692 / 2048.0, 26722 / 2048.0, 22725 / 2048.0, 17855 / 2048.0, 12299 / 2048.0, 6270 / 2048.0,
    21407 / 2048.0, 29692 / 2048.0, 27969 / 2048.0, 25172 / 2048.0, 21407 / 2048.0, 16819 / 2048.0, 11585 / 2048.0, 5906 / 2048.0,
    19266 / 2048.0, 26722 / 2048.0, 25172 / 2048.0, 22654 / 2048.0, 19266 / 2048.0, 15137 / 2048.0, 10426 / 2048.0, 5315 / 2048.0,
    16384 / 2048.0, 22725 / 2048.0, 21407 / 2048.0, 19266 / 2048.0, 16384 / 2048.0, 12873 / 2048.0, 8867 / 2048.0,  4520 / 2048.0,
    12873 / 2048.0, 17855 / 2048.0, 16819 / 2048.0, 15137 / 2048.0, 12873 / 2048.0, 10114 / 2048.0, 6967 / 2048.0,  3552 / 2048.0,
    8867 / 2048.0,  12299 / 2048.0, 11585 / 2048.0, 10426 / 2048.0, 8867 / 2048.0,  6967 / 2048.0,  4799 / 2048.0,  2446 / 2048.0,
    4520 / 2048.0,  6270 / 2048.0,  5906 / 2048.0,  5315 / 2048.0,  4520 / 2048.0,  3552 / 2048.0,  2446 / 2048.0,  1247 / 2048.0};
__constant__ int c_qbase[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};

// ---------------------------------------------------------------------------------------------------
// Trie construction (host and device): insert one codeword; the shortest matching prefix wins, like
// `while prefix not in table` (huffman.py:69).  Returns false when the node pool is exhausted.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline void table_clear(DecTable& t, int& nodes) {
    for (int i = 0; i < 256; i++) t.lut[i] = 0;
    t.child[0][0] = t.child[0][1] = 0;
    nodes = 1;
}

// `root_leaf`: a zero-length codeword (one-symbol alphabet, huffman.py:175-180) makes the root a leaf.
__host__ __device__ inline bool table_insert(DecTable& t, int& nodes, int& root_leaf, uint32_t code, int len, int sym) {
    if (len == 0) { root_leaf = sym; return true; }
    int node = 0;
    for (int i = 0; i < len; i++) {
        int bit = (code >> (len - 1 - i)) & 1;
        uint16_t c = t.child[node][bit];
        if (i == len - 1) { t.child[node][bit] = (uint16_t)(kLeaf | (uint32_t)sym); return true; }
        if (c & kLeaf) return true;   // a shorter codeword is a prefix of this one: unreachable
        if (c == 0) {
            if (nodes >= kMaxNodes) return false;
            c = (uint16_t)nodes++;
            t.child[c][0] = t.child[c][1] = 0;
            t.child[node][bit] = c;
        }
        node = c;
    }
    return true;
}

__host__ __device__ inline void table_finish(DecTable& t, int root_leaf) {
    for (int x = 0; x < 256; x++) {
        if (root_leaf >= 0) { t.lut[x] = (uint16_t)(kLeaf | (uint32_t)root_leaf); continue; }
        int node = 0;
        uint16_t e = 0;
        for (int d = 0; d < 8; d++) {
            uint16_t c = t.child[node][(x >> (7 - d)) & 1];
            if (c == 0) { e = 0; break; }
            if (c & kLeaf) { e = (uint16_t)(kLeaf | ((uint32_t)(d + 1) << 8) | (c & 0xffu)); break; }
            node = c;
            e = c;   // after 8 steps: the node to continue from
        }
        t.lut[x] = e;
    }
}

// Returns the larger trie node count of the two tables (the kernels keep kShNodes = 256 of them in shared memory).
static int build_default_tables(DecTables& t) {
    int nodes, root = -1, most = 0;
    table_clear(t.dc, nodes);
    uint32_t code = 0;
    int k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < tic::kDcBits[l]; i++) table_insert(t.dc, nodes, root, code++, l, tic::kDcVals[k++]);
        code <<= 1;
    }
    table_finish(t.dc, -1);
    most = nodes;
    table_clear(t.ac, nodes);
    code = 0;
    k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < tic::kAcBits[l]; i++) table_insert(t.ac, nodes, root, code++, l, tic::kAcVals[k++]);
        code <<= 1;
    }
    table_finish(t.ac, -1);
    return nodes > most ? nodes : most;
}

struct ShTables;
static bool build_sh_tables(const DecTables& t, ShTables& sh);   // defined next to ShTables

// ---------------------------------------------------------------------------------------------------
// Bit access: the stream is MSB-first bytes (bitarray(endian="big"), bitbuffer.py:7); 32 bits starting at
// bit p, zeros past the end (a slice past the end of a bitarray is empty).
// ---------------------------------------------------------------------------------------------------
struct BitSrc {
    const uint32_t* words;
    long long nwords;
    uint32_t tailmask;   // valid bits of the last word (stream sizes need not be multiples of 4)
};

__device__ __forceinline__ BitSrc make_src(const DecImage& im) {
    BitSrc s;
    s.words = im.words;
    s.nwords = (im.nbits + 31) >> 5;
    int tail = (int)(im.nbits & 31);
    s.tailmask = tail ? (0xffffffffu << (32 - tail)) : 0xffffffffu;
    return s;
}

__device__ __forceinline__ uint32_t load_be(const BitSrc& s, long long i) {
    if (i >= s.nwords) return 0u;
    uint32_t w = __byte_perm(__ldg(s.words + i), 0, 0x0123);
    return i == s.nwords - 1 ? (w & s.tailmask) : w;
}

__device__ __forceinline__ uint32_t peek32(const BitSrc& s, long long p) {
    long long i = p >> 5;
    return __funnelshift_l(load_be(s, i + 1), load_be(s, i), (uint32_t)(p & 31));
}

// One Huffman symbol from the 32 bits `v`: read_huffman_code (huffman.py:66-74).  len < 0: no codeword of
// at most 16 bits matches (the reference raises ValueError).
__device__ __forceinline__ void lookup(const DecTable* t, uint32_t v, int& sym, int& len) {
    uint32_t e = t->lut[v >> 24];
    if (e & kLeaf) { sym = (int)(e & 0xffu); len = (int)((e >> 8) & 0x7fu); return; }
    len = 8;
    uint32_t node = e;
    while (node != 0 && len < 16) {
        uint32_t c = t->child[node][(v >> (31 - len)) & 1u];
        len++;
        if (c & kLeaf) { sym = (int)(c & 0xffu); return; }
        node = c;
    }
    sym = 0;
    len = -1;
}

// The same lookup for a DecTable that lives in SHARED memory, addressed through its 32-bit shared address: the fixed
// tables of every CTA.  (Through a `const DecTable*` that may also point at a per-image table in global memory the
// compiler has to emit generic loads with 64-bit address arithmetic: LD.E + IMAD.WIDE + IADD3 + IMAD.X per access,
// two accesses per symbol in the hottest loop of the decoder.)
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t shared_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// The shared-memory form of the FIXED tables: two levels, no bit-by-bit walk.  Level 1 = DecTable::lut of each
// alphabet over the next 8 bits (a leaf entry: kLeaf | length << 8 | symbol); for a code longer than 8 bits the entry
// is 1 + the number of a level-2 block (no kLeaf bit), and that block resolves the following 8 bits the same way
// (length = the whole code's).  The reference's tables need 1 block for the DC and 5 for the AC alphabet; the host
// builds the image once per handle (build_sh_tables) and every CTA copies its 5 KB.  In the benchmark's streams one
// symbol in eight has a code of more than 8 bits, so some lane of a warp took the bit-by-bit branch in 98 % of all symbol
// steps — 15 instructions with ONE lane active (profiles/r2g_dec_sync_kernel).
constexpr int kL2Blocks = 8;
struct ShTables {
    uint16_t lut_dc[256], lut_ac[256];
    uint16_t lut2[kL2Blocks][256];
};
// all threads of the CTA; the caller synchronises before the first lookup
__device__ __forceinline__ void load_sh_tables(ShTables& sh, const ShTables* __restrict__ image, int nthreads) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(image);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh);
    for (int i = threadIdx.x; i < (int)(sizeof(ShTables) / 4); i += nthreads) dst[i] = __ldg(src + i);
}
// tabs: shared address of the ShTables.  No branch: the second load is always issued (entry 0 of block 0 for a leaf).
__device__ __forceinline__ void lookup_sh(uint32_t tabs, bool is_dc, uint32_t v, int& sym, int& len) {
    const uint32_t e1 = lds_u16(tabs + (is_dc ? 0u : 512u) + ((v >> 24) << 1));
    const bool leaf = (e1 & kLeaf) != 0;
    const uint32_t e2 = lds_u16(tabs + 1024u + (leaf || e1 == 0 ? 0u : ((e1 - 1u) << 9) + ((v >> 15) & 0x1feu)));
    const uint32_t e = leaf ? e1 : (e1 == 0 ? 0u : e2);
    sym = (int)(e & 0xffu);
    len = (e & kLeaf) ? (int)((e >> 8) & 0x7fu) : -1;   // no codeword of at most 16 bits matches
}
static_assert(offsetof(ShTables, lut_ac) == 512 && offsetof(ShTables, lut2) == 1024, "lookup_sh addresses ShTables by hand");
// Host: the two-level image of the fixed tables from their tries.  False if they need more than kL2Blocks blocks.
static bool build_sh_tables(const DecTables& t, ShTables& sh) {
    memset(&sh, 0, sizeof sh);
    int blocks = 0;
    const DecTable* src[2] = {&t.dc, &t.ac};
    uint16_t* lut[2] = {sh.lut_dc, sh.lut_ac};
    for (int a = 0; a < 2; a++) {
        int node_block[kMaxNodes];
        for (int i = 0; i < kMaxNodes; i++) node_block[i] = -1;
        for (int x = 0; x < 256; x++) {
            const uint16_t e = src[a]->lut[x];
            if ((e & kLeaf) || e == 0) { lut[a][x] = e; continue; }
            if (node_block[e] < 0) {   // a new level-2 block: the 8 bits behind trie node e
                if (blocks >= kL2Blocks) return false;
                const int b = blocks++;
                node_block[e] = b;
                for (int y = 0; y < 256; y++) {
                    int node = e;
                    uint16_t out = 0;
                    for (int d = 0; d < 8; d++) {
                        const uint16_t c = src[a]->child[node][(y >> (7 - d)) & 1];
                        if (c == 0) break;
                        if (c & kLeaf) { out = (uint16_t)(kLeaf | ((uint32_t)(9 + d) << 8) | (c & 0xffu)); break; }
                        node = c;
                    }
                    sh.lut2[b][y] = out;
                }
            }
            lut[a][x] = (uint16_t)(node_block[e] + 1);
        }
    }
    return true;
}

// read_int (bitbuffer.py:56-66) on the window t = v << len: `size` bits, a leading 0 = negative (one's complement)
__device__ __forceinline__ int read_value(uint32_t t, int size) {
    uint32_t vb, ones;
    asm("shr.b32 %0, %1, %2;" : "=r"(vb) : "r"(t), "r"(32 - size));    // size == 0: shift by 32 = 0 (PTX clamps)
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(ones) : "r"(0), "r"(size));   // (1 << size) - 1
    return (int)vb - ((int)t >= 0 ? (int)ones : 0);
}

// A subsequence's bits staged in shared memory: kSubBits/32 words plus the two words a symbol that starts in
// the last bit can reach (16 code bits + 15 value bits), MSB-first, zeros past the end of the stream.
constexpr int kSubWords = kSubBits / 32;
constexpr int kRowWords = kSubWords + 3;   // 34 used; odd stride: threads walk their rows at different speeds

// Warp-cooperative, coalesced: for each of the warp's 32 subsequences the 32 lanes fetch its words together.
__device__ __forceinline__ void stage_rows(uint32_t (*rows)[kRowWords], bool active, const BitSrc& src, long long word0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = 0; j < 32; j++) {
        if (!__shfl_sync(0xffffffffu, (int)active, j)) continue;
        const uint32_t* words = reinterpret_cast<const uint32_t*>(
            __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(src.words), j));
        long long nw = __shfl_sync(0xffffffffu, src.nwords, j);
        uint32_t tm = __shfl_sync(0xffffffffu, src.tailmask, j);
        long long first = __shfl_sync(0xffffffffu, word0, j);
        for (int c = lane; c < kSubWords + 2; c += 32) {
            long long i = first + c;
            uint32_t w = 0;
            if (i < nw) {
                w = __byte_perm(__ldg(words + i), 0, 0x0123);
                if (i == nw - 1) w &= tm;
            }
            rows[warp * 32 + j][c] = w;
        }
    }
    __syncwarp();
}

struct SubResult {
    uint32_t exit;   // entry state of the next subsequence: overshoot bits | zigzag index << 16
    int n;           // DC symbols (= blocks) that START in this subsequence
    int dsum;        // sum of their DC differences
    uint32_t err;
};

// Decode one subsequence from `entry`, reading its bits from the staged row `sw`; nothing is stored (phases 1
// and 2; dec_write_kernel walks the same parse once more with the stores).  `end_rel`: bits of the subsequence
// that belong to the stream (kSubBits except at the end).
// kSh: the tables are the CTA's fixed tables in shared memory (sh_tabs = shared address of the DecTables); the row and the
// tables are then read with 32-bit shared addresses, and the iteration cap is not needed (every symbol of the fixed
// tables is at least one bit long, so a subsequence has at most kSubBits + 31 symbols).
// What a thread remembers of its last complete decode of its subsequence (fixed tables only): where the first
// blocks start, and the sums in front of them.  A later decode from a different entry that reaches one of these
// positions expecting a DC symbol is, from there on, the remembered decode — decode() is a function of (bit, zigzag
// index) — so it stops there and takes the rest of its result from the record.  Measured on a CPU model
// (tests/test_sync_model.py): two parses of one subsequence meet after 39 bits on average at quality 50 (p90 70,
// max 159 of 1024), 108 bits at quality 90: the repeat decode of the synchronisation rounds shrinks from ~210
// symbols to ~15.  valid: the record describes a complete decode and has not been used yet.
struct SubRec {
    uint32_t pos01 = 0xffffffffu, pos23 = 0xffffffffu;   // bit positions of the first block starts, 16 bits each (0xffff: none)
    int pre1 = 0, pre2 = 0, pre3 = 0;                    // sum of the DC differences of the blocks in front of start 1, 2, 3
    int n = 0, dsum = 0;
    uint32_t exit = 0;
    bool valid = false;
};

template <bool kSh>
__device__ SubResult decode_sub(const uint32_t* sw, int end_rel, const DecTable* tdc, const DecTable* tac, uint32_t sh_tabs,
                                uint32_t entry, SubRec* rec = nullptr) {
    int p = (int)(entry & 0xffffu);
    int z = (int)((entry >> 16) & 0xffu);
    SubResult r;
    r.n = 0; r.dsum = 0; r.err = 0;
    const uint32_t row = shared_addr(sw);
    const bool match = kSh && rec != nullptr && rec->valid, record = kSh && rec != nullptr && !rec->valid;
    if (record) { rec->pos01 = rec->pos23 = 0xffffffffu; rec->pre1 = rec->pre2 = rec->pre3 = 0; }
    // the bookkeeping below only runs while it can still matter: the first four block starts of a recording
    // decode, the stretch up to the last remembered start of a comparing one (one branch per symbol step otherwise)
    int watch_end = record ? end_rel : (match ? (int)(rec->pos23 >> 16 != 0xffffu ? rec->pos23 >> 16 : (rec->pos23 & 0xffffu) != 0xffffu ? rec->pos23 & 0xffffu : rec->pos01 >> 16 != 0xffffu ? rec->pos01 >> 16 : rec->pos01 & 0xffffu) : -1);
    if (match && watch_end == 0xffff) watch_end = -1;   // nothing remembered
    int it = 0;
    for (; p < end_rel && (kSh || it < kMaxSymbols); it++) {
        uint32_t v;
        int sym, len;
        if constexpr (kSh) {
            if (z == 0 && p <= watch_end) {   // a block starts here
                if (match) {
                    const uint32_t pp = (uint32_t)p;
                    const int j = pp == (rec->pos01 & 0xffffu) ? 0 : (pp == (rec->pos01 >> 16) ? 1 : (pp == (rec->pos23 & 0xffffu) ? 2 : (pp == (rec->pos23 >> 16) ? 3 : -1)));
                    if (j >= 0) {   // from here on this decode IS the remembered one
                        r.n += rec->n - j;
                        r.dsum += rec->dsum - (j == 0 ? 0 : (j == 1 ? rec->pre1 : (j == 2 ? rec->pre2 : rec->pre3)));
                        r.exit = rec->exit;
                        rec->valid = false;
                        return r;
                    }
                } else {   // (a position that fails to decode is overwritten: n has not moved)
                    const uint32_t pp = (uint32_t)p;
                    if (r.n == 0) rec->pos01 = (rec->pos01 & 0xffff0000u) | pp;
                    else if (r.n == 1) { rec->pos01 = (rec->pos01 & 0x0000ffffu) | (pp << 16); rec->pre1 = r.dsum; }
                    else if (r.n == 2) { rec->pos23 = (rec->pos23 & 0xffff0000u) | pp; rec->pre2 = r.dsum; }
                    else if (r.n == 3) { rec->pos23 = (rec->pos23 & 0x0000ffffu) | (pp << 16); rec->pre3 = r.dsum; }
                    else watch_end = -1;   // four starts are on record
                }
            }
            const uint32_t a = row + ((uint32_t)(p >> 5) << 2);
            v = __funnelshift_l(lds_u32(a + 4u), lds_u32(a), (uint32_t)(p & 31));
            lookup_sh(sh_tabs, z == 0, v, sym, len);
        } else {
            const int wi = p >> 5;
            v = __funnelshift_l(sw[wi + 1], sw[wi], (uint32_t)(p & 31));
            lookup(z == 0 ? tdc : tac, v, sym, len);
        }
        if (len < 0) {   // no codeword: a deterministic rule so that decode(entry) stays a function
            r.err |= TIC_DSTATUS_CODE;
            p += 1;
            continue;
        }
        const int size = sym & 15;   // DC: the category is the size; AC: (run, size), huffman.py:89-93
        const int val = read_value(v << len, size);
        const int p0 = p;
        p += len + size;
        if (z == 0) {
            r.n++;
            r.dsum += val;
            z = 1;
        } else if (sym == 0) {   // EOB (huffman.py:94-95)
            z = 0;
        } else {
            z += sym >> 4;       // decode_run_length (huffman.py:36-38): `run` zeros, then the value
            if (z > 63) {
                // More than 63 coefficients without an EOB: never on the true parse of a valid stream (the
                // reference's ac[i, :len] assignment raises).  A guessed entry can fall into such a parse and
                // stay in it for ever when the bits are periodic (a flat area is `00 1010` repeated, and from
                // phase 4 that reads as the 6-bit AC symbol `100 010` again and again), so the rule that
                // keeps decode(entry) a function also has to break the cycle: start over one bit further
                // on, expecting a DC symbol.
                r.err |= TIC_DSTATUS_CODE;
                z = 0;
                p = p0 + 1;
                continue;
            }
            z += 1;
        }
    }
    int over = p - kSubBits;
    r.exit = (uint32_t)(over > 0 ? over : 0) | ((uint32_t)z << 16);
    if (kSh && rec != nullptr) {
        rec->valid = record;   // a complete decode that recorded its block starts; a record that was compared is spent
        if (record) { rec->n = r.n; rec->dsum = r.dsum; rec->exit = r.exit; }
    }
    return r;
}

__device__ __forceinline__ int find_owner(const long long* __restrict__ first, int n, long long g) {
    int lo = 0, hi = n;   // last i with first[i] <= g
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(first + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// The same for a whole CTA whose threads hold consecutive indices g0 + threadIdx.x: one binary search (12
// dependent loads for 4096 images) by thread 0, then a short walk forward for the threads behind an image
// boundary.  Contains a barrier: every thread of the CTA must call it.
__device__ __forceinline__ int find_owner_cta(const long long* __restrict__ first, int n, long long g0, long long g,
                                              long long total) {
    __shared__ int cta_owner;
    if (threadIdx.x == 0) cta_owner = find_owner(first, n, g0 < total ? g0 : total - 1);
    __syncthreads();
    int idx = cta_owner;
    if (g < total)
        while (idx + 1 < n && __ldg(first + idx + 1) <= g) idx++;
    return idx;
}

// ---------------------------------------------------------------------------------------------------
// dec_setup_kernel: one warp per image.  parse_header (codec.py:117-130), read_huffman_table
// (codec.py:87-99), the multiplier table of block_quantize(inverse=True) (utils.py:48-52).
// ---------------------------------------------------------------------------------------------------
// The multiplier table of block_quantize(inverse=True) (utils.py:48-52) — for mode 2 the table of quality 50
// (codec.py:62) and 2**quality (codec.py:61).  Returns status bits.
__device__ inline uint32_t fill_mul(DecImage& im, double* __restrict__ m, float* __restrict__ mf) {
    uint32_t q = im.quality;
    if (im.mode == MODE_SCALED) {
        if (q > 1000) return TIC_DSTATUS_QUALITY;
        im.two_q = scalbn(1.0, (int)q);
        for (int k = 0; k < 64; k++) m[k] = __ddiv_rn((double)((long long)c_qbase[k] * 100), 100.0);
    } else if (q == 0) {
        return TIC_DSTATUS_QUALITY;   // ZeroDivisionError, utils.py:50
    } else if (q < 50) {
        double factor = __ddiv_rn(5000.0, (double)q);
        for (int k = 0; k < 64; k++) m[k] = __ddiv_rn(__dmul_rn((double)c_qbase[k], factor), 100.0);
    } else {
        long long factor = 200 - 2 * (long long)q;
        for (int k = 0; k < 64; k++) m[k] = __ddiv_rn((double)((long long)c_qbase[k] * factor), 100.0);
    }
    // the fast pass's single-precision multiplier: one rounding of the whole factor (dec_idct_fast_kernel); a second
    // copy with the indices transposed (v*8+u) for the fused coefficient pass, whose lanes hold COLUMNS (kMulfStride)
    for (int k = 0; k < 64; k++) {
        mf[k] = (float)(im.mode == MODE_SCALED ? m[k] * im.two_q / d_ann[k] : m[k]);
        mf[64 + (k & 7) * 8 + (k >> 3)] = mf[k];
    }
    return 0;
}

__device__ inline uint32_t get_bits(const BitSrc& s, long long& p, int n) {   // read_uint, n <= 16
    uint32_t v = n ? (peek32(s, p) >> (32 - n)) : 0u;
    p += n;
    return v;
}

__global__ void __launch_bounds__(32) dec_setup_kernel(DecImage* __restrict__ imgs, int n_images, uint32_t flags,
                                                       DecTables* __restrict__ tabs, double* __restrict__ mul,
                                                       float* __restrict__ mulf, uint32_t* __restrict__ E, int* __restrict__ status,
                                                       int* __restrict__ summary) {
    __shared__ DecTables sh;
    __shared__ int sh_build;
    int i = blockIdx.x;
    if (i >= n_images) return;
    DecImage& im = imgs[i];
    int lane = threadIdx.x;
    if (lane == 0) {
        sh_build = 0;
        uint32_t st = 0;
        im.mode = MODE_FIXED; im.skip_entropy = 0; im.skip_pixels = 0; im.anchor_sub = 0; im.has_tables = 0;
        im.quality = 0; im.flag = 0; im.two_q = 1.0;
        BitSrc src = make_src(im);
        long long data_start = 128;
        if (im.nbits < 128) {
            st |= TIC_DSTATUS_HEADER;
        } else {
            // struct.unpack("IIII"): native little-endian words
            uint32_t hh = __ldg(im.words + 0), ww = __ldg(im.words + 1);
            im.quality = __ldg(im.words + 2);
            im.flag = __ldg(im.words + 3);
            if (hh != (uint32_t)im.height || ww != (uint32_t)im.width) st |= TIC_DSTATUS_HEADER;
        }
        if (st == 0) {
            uint32_t f = im.flag;
            if ((flags & TIC_DFLAG_ACCEPT_BE_FLAG) && f == 0x00000080u) f = 0x80000000u;
            if (f & 0x80000000u) im.mode = MODE_TABLES;
            else if (f & 0x40000000u) im.mode = MODE_SCALED;
            st |= fill_mul(im, mul + (size_t)i * 64, mulf + (size_t)i * kMulfStride);
        }
        if (st == 0 && im.mode == MODE_TABLES) {
            long long p = 128;
            bool ok = true;
            for (int pass = 0; pass < 2 && ok; pass++) {
                DecTable& t = pass ? sh.ac : sh.dc;
                int nodes, root = -1;
                table_clear(t, nodes);
                uint32_t cnt = get_bits(src, p, 16);
                for (uint32_t e = 0; e < cnt && p < im.nbits; e++) {
                    uint32_t a = get_bits(src, p, 4);
                    uint32_t b = pass ? get_bits(src, p, 4) : 0u;
                    uint32_t len = get_bits(src, p, pass ? 8 : 4);
                    int sym = pass ? (int)(a * 16 + b) : (int)a;
                    if (len <= 16) {
                        uint32_t code = get_bits(src, p, (int)len);
                        if (!table_insert(t, nodes, root, code, (int)len, sym)) ok = false;
                    } else {
                        p += len;   // can never match: read_huffman_code gives up after 17 bits
                    }
                }
                table_finish(t, root);
            }
            if (!ok) st |= TIC_DSTATUS_TABLE;
            data_start = p;
            sh_build = 1;
            im.has_tables = 1;
            // every block is zero bits long: one-symbol alphabets with empty codewords on both sides
            if (ok && sh.dc.lut[0] == (uint16_t)(kLeaf | 0u) && sh.ac.lut[0] == (uint16_t)(kLeaf | 0u)) im.skip_entropy = 1;
        }
        if (st & (TIC_DSTATUS_HEADER)) im.skip_pixels = 1;
        if (st) im.skip_entropy = 1;
        if (st & TIC_DSTATUS_QUALITY) im.skip_pixels = 1;
        if (st == 0 && im.nblk > 0 && im.nsubs == 0) st |= TIC_DSTATUS_TRUNCATED;   // a header and nothing else
        if (st || im.nblk == 0 || im.nsubs == 0) im.skip_entropy = 1;
        if (!im.skip_entropy) {
            long long rel = data_start - 128;
            if (rel >= (long long)im.nsubs * kSubBits) {
                im.skip_entropy = 1;   // tables run to the end of the stream: no block data at all
                st |= TIC_DSTATUS_TRUNCATED;
            } else {
                im.anchor_sub = (int)(rel / kSubBits);
                E[im.sub_first + im.anchor_sub] = (uint32_t)(rel % kSubBits);   // zigzag index 0: a DC symbol is next
            }
        }
        if (st) { atomicOr(&status[i], (int)st); atomicOr(summary, (int)st); }
    }
    __syncwarp();
    if (sh_build) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(&sh);
        uint32_t* d = reinterpret_cast<uint32_t*>(tabs + i);
        for (int k = lane; k < (int)(sizeof(DecTables) / 4); k += 32) d[k] = s[k];
    }
}

// ---------------------------------------------------------------------------------------------------
// dec_sync_kernel: phases 1 and 2.  E[g] = entry state of subsequence g (written by the thread of g-1),
// U[g] = the entry the thread of g last decoded from, ND[g] = (blocks started, DC-difference sum) of
// that decode.  Invariant: E[g+1] == decode(g, U[g]).exit; fixed point <=> U == E everywhere.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSyncThreads) dec_sync_kernel(const DecImage* __restrict__ imgs,
                                                                const long long* __restrict__ sub_first, int n_images,
                                                                long long total_subs, const ShTables* __restrict__ deftab,
                                                                const DecTables* __restrict__ tabs, uint32_t* E,
                                                                uint32_t* __restrict__ U, int2* __restrict__ ND,
                                                                int* __restrict__ changed, int early_stop) {
    __shared__ uint32_t rows[kSyncThreads][kRowWords];
    __shared__ ShTables sh_def;   // the fixed tables; per-image tables stay in global memory
    load_sh_tables(sh_def, deftab, kSyncThreads);
    long long g = (long long)blockIdx.x * kSyncThreads + threadIdx.x;
    bool active = g < total_subs;
    int idx = 0, k = 0, end_rel = 0;
    bool has_next = false;
    uint32_t used = 0xffffffffu;
    const DecTables* tb = nullptr;   // per-image tables (global memory); null: the fixed tables in shared memory
    BitSrc src = {};
    idx = find_owner_cta(sub_first, n_images, (long long)blockIdx.x * kSyncThreads, g, total_subs);   // barrier inside
    if (active) {
        const DecImage& im = imgs[idx];
        k = (int)(g - im.sub_first);
        active = !im.skip_entropy && k >= im.anchor_sub;
        has_next = k + 1 < im.nsubs;
        if (im.has_tables) tb = tabs + idx;
        src = make_src(im);
        long long left = im.nbits - (128 + (long long)k * kSubBits);
        end_rel = left < kSubBits ? (int)left : kSubBits;
        used = U[g];
    }
    volatile uint32_t* Ev = E;
    // later rounds: most CTAs have nothing left to repair — leave before fetching the bits again
    if (!__syncthreads_or(active && Ev[g] != used)) return;
    stage_rows(rows, active, src, 4 + (long long)k * kSubWords);
    const uint32_t* sw = rows[threadIdx.x];
    bool any = false;
    SubRec rec;
    for (int it = 0; it < kSyncIters; it++) {
        bool wrote = false;
        if (active) {
            uint32_t e = Ev[g];
            if (e != used) {
                used = e;
                SubResult r = tb == nullptr ? decode_sub<true>(sw, end_rel, nullptr, nullptr, shared_addr(&sh_def), e, early_stop ? &rec : nullptr)
                                            : decode_sub<false>(sw, end_rel, &tb->dc, &tb->ac, 0u, e);
                ND[g] = make_int2(r.n, r.dsum);
                if (has_next && Ev[g + 1] != r.exit) { Ev[g + 1] = r.exit; wrote = true; }
            }
        }
        if (!__syncthreads_or(wrote)) break;
        any = true;
    }
    if (active) U[g] = used;
    if (any && threadIdx.x == 0) *changed = 1;
}

// ---------------------------------------------------------------------------------------------------
// Phase 3: exclusive sums of ND = (blocks started, sum of DC differences) over each stream's subsequences
// (np.cumsum of the DC differences, codec.py:53, and the block index of every thread).  A stream is cut into
// slices of at most kSliceSubs subsequences, one CTA each (the host builds the slice list); a stream that
// needs several slices — a single large image — first gets its slice totals (dec_scan_totals_kernel), then
// every slice starts from the sum of the totals in front of it.
// ---------------------------------------------------------------------------------------------------
// A block goes to the exact pass (dec_idct_kernel over `list`): its index, once.  Entries beyond the capacity are
// dropped and reported (TIC_DSTATUS_CODE on the batch: only streams that are not valid .img data can get there).
__device__ __forceinline__ void list_push(long long* __restrict__ list, int* __restrict__ list_count, int list_cap,
                                          long long block, int* __restrict__ summary) {
    const int slot = atomicAdd(list_count, 1);
    if (slot < list_cap) list[slot] = block;
    else atomicOr(summary, TIC_DSTATUS_CODE);
}

constexpr int kSliceSubs = 32768;
struct ScanSlice {
    int img, first, count, index_in_img;
};

__device__ __forceinline__ int2 block_sum_1024(int2 v, int2* warp_buf) {   // sum over the CTA, result in every thread
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 16; d > 0; d >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, d);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, d);
    }
    if (lane == 0) warp_buf[warp] = v;
    __syncthreads();
    int2 t = warp_buf[lane];
    for (int d = 16; d > 0; d >>= 1) {
        t.x += __shfl_xor_sync(0xffffffffu, t.x, d);
        t.y += __shfl_xor_sync(0xffffffffu, t.y, d);
    }
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(1024) dec_scan_totals_kernel(const DecImage* __restrict__ imgs,
                                                               const ScanSlice* __restrict__ slices,
                                                               const int2* __restrict__ ND, int2* __restrict__ slice_tot) {
    __shared__ int2 warp_buf[32];
    const ScanSlice sl = slices[blockIdx.x];
    const DecImage& im = imgs[sl.img];
    int2 acc = make_int2(0, 0);
    if (!im.skip_entropy)
        for (int k = threadIdx.x; k < sl.count; k += 1024) {
            int2 v = ND[im.sub_first + sl.first + k];
            acc.x += v.x; acc.y += v.y;
        }
    acc = block_sum_1024(acc, warp_buf);
    if (threadIdx.x == 0) slice_tot[blockIdx.x] = acc;
}

// E, coef: when given, the block a subsequence is entered in the middle of (zigzag index > 0 in its entry state)
// is zeroed here, 128 bytes: those are the only blocks dec_write_kernel fills with partial stores.
__global__ void __launch_bounds__(1024) dec_scan_kernel(const DecImage* __restrict__ imgs, const ScanSlice* __restrict__ slices,
                                                        const int2* __restrict__ slice_tot, const int2* __restrict__ ND,
                                                        int2* __restrict__ NB, int* __restrict__ status,
                                                        int* __restrict__ summary, const uint32_t* __restrict__ E,
                                                        int16_t* __restrict__ coef, int* __restrict__ ndec,
                                                        long long* __restrict__ list, int* __restrict__ list_count,
                                                        int list_cap) {
    const ScanSlice sl = slices[blockIdx.x];
    const DecImage& im = imgs[sl.img];
    if (im.skip_entropy) return;
    __shared__ int2 warp_tot[32], warp_excl[32];
    __shared__ int2 carry_sh, chunk_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        int2 acc = make_int2(0, 0);   // the slices of this stream in front of this one
        for (int j = threadIdx.x; j < sl.index_in_img; j += 1024) {
            int2 v = slice_tot[blockIdx.x - sl.index_in_img + j];
            acc.x += v.x; acc.y += v.y;
        }
        acc = block_sum_1024(acc, warp_tot);
        if (threadIdx.x == 0) carry_sh = acc;
    }
    __syncthreads();
    const long long base_g = im.sub_first + sl.first;
    for (int base = 0; base < sl.count; base += 1024) {
        int k = base + threadIdx.x;
        int2 v = k < sl.count ? ND[base_g + k] : make_int2(0, 0);
        int2 inc = v;
        for (int d = 1; d < 32; d <<= 1) {
            int a = __shfl_up_sync(0xffffffffu, inc.x, d), b = __shfl_up_sync(0xffffffffu, inc.y, d);
            if (lane >= d) { inc.x += a; inc.y += b; }
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int2 w = warp_tot[lane], wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                int a = __shfl_up_sync(0xffffffffu, wi.x, d), b = __shfl_up_sync(0xffffffffu, wi.y, d);
                if (lane >= d) { wi.x += a; wi.y += b; }
            }
            warp_excl[lane] = make_int2(wi.x - w.x, wi.y - w.y);
            if (lane == 31) chunk_tot = wi;
        }
        __syncthreads();
        int2 c = carry_sh, wo = warp_excl[warp];
        if (k < sl.count) {
            const int first_blk = c.x + wo.x + inc.x - v.x;
            NB[base_g + k] = make_int2(first_blk, c.y + wo.y + inc.y - v.y);
            if (E != nullptr && ((E[base_g + k] >> 16) & 0xffu) != 0 && first_blk >= 1 && first_blk <= im.nblk) {
                uint4* dst = reinterpret_cast<uint4*>(coef + (im.blk_first + first_blk - 1) * 64);
#pragma unroll
                for (int i = 0; i < 8; i++) dst[i] = make_uint4(0u, 0u, 0u, 0u);
                // fused coefficient pass: these blocks are transformed by the exact pass.  A block entered in the middle
                // by two subsequences (longer than one of them) would be listed twice: only the first entry counts
                if (list != nullptr && (sl.first + k == 0 || ND[base_g + k - 1].x > 0))
                    list_push(list, list_count, list_cap, im.blk_first + first_blk - 1, summary);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = make_int2(c.x + chunk_tot.x, c.y + chunk_tot.y);
        __syncthreads();
    }
    if (threadIdx.x == 0 && sl.first + sl.count == im.nsubs) {
        ndec[sl.img] = carry_sh.x < im.nblk ? carry_sh.x : im.nblk;
        if (carry_sh.x < im.nblk) {
            atomicOr(&status[sl.img], TIC_DSTATUS_TRUNCATED);
            atomicOr(summary, TIC_DSTATUS_TRUNCATED);
        }
    }
}

// FP32 8-point inverse transform and the pixel store shared by the fused coefficient pass (dec_write_kernel) and
// dec_idct_fast_kernel; the error bound that makes it usable is derived at dec_idct_fast_kernel.
__device__ __forceinline__ void idct8_fast(float& x0, float& x1, float& x2, float& x3, float& x4, float& x5, float& x6,
                                           float& x7) {
    // c(u) cos((2y+1) u pi / 16), c(0) = sqrt(1/8), c(u > 0) = 1/2: even columns u = 0,2,4,6 and odd u = 1,3,5,7
    const float A = 0.35355339059327379f, B = 0.46193976625564337f, C = 0.19134171618254489f;
    const float P = 0.49039264020161522f, Q = 0.41573480615127262f, R = 0.27778511650980111f, T = 0.09754516100806413f;
    const float e0 = fmaf(C, x6, fmaf(A, x4, fmaf(B, x2, A * x0)));
    const float e1 = fmaf(-B, x6, fmaf(-A, x4, fmaf(C, x2, A * x0)));
    const float e2 = fmaf(B, x6, fmaf(-A, x4, fmaf(-C, x2, A * x0)));
    const float e3 = fmaf(-C, x6, fmaf(A, x4, fmaf(-B, x2, A * x0)));
    const float o0 = fmaf(T, x7, fmaf(R, x5, fmaf(Q, x3, P * x1)));
    const float o1 = fmaf(-R, x7, fmaf(-P, x5, fmaf(-T, x3, Q * x1)));
    const float o2 = fmaf(Q, x7, fmaf(T, x5, fmaf(-P, x3, R * x1)));
    const float o3 = fmaf(-P, x7, fmaf(Q, x5, fmaf(-R, x3, T * x1)));
    x0 = e0 + o0; x7 = e0 - o0;
    x1 = e1 + o1; x6 = e1 - o1;
    x2 = e2 + o2; x5 = e2 - o2;
    x3 = e3 + o3; x4 = e3 - o3;
}

__device__ __forceinline__ void store_pixel_row(const DecImage& im, int y, int x0, uint32_t lo, uint32_t hi) {
    if (y >= im.height) return;
    uint8_t* row = im.pixels + (size_t)y * (size_t)im.width + x0;
    if (x0 + 8 <= im.width && (((uintptr_t)im.pixels | (uintptr_t)im.width) & 7u) == 0) {
        *reinterpret_cast<uint2*>(row) = make_uint2(lo, hi);
    } else {
#pragma unroll
        for (int v = 0; v < 8; v++)
            if (x0 + v < im.width) row[v] = (uint8_t)((v < 4 ? lo >> (8 * v) : hi >> (8 * (v - 4))) & 0xffu);
    }
}

// 8 x 8 transpose across the 8 lanes of a group (lane j = lane & 7): in: a[u] = element (u, j), out: a[v] = element (j, v).
__device__ __forceinline__ void transpose8(float (&a)[8], int j) {
#pragma unroll
    for (int d = 4; d >= 1; d >>= 1) {
        const bool up = (j & d) != 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if ((i & d) == 0) {   // element (i + d, j) of a lane with bit d clear <-> element (i, j ^ d) of its partner
                const float got = __shfl_xor_sync(0xffffffffu, up ? a[i] : a[i + d], d);
                if (up) a[i] = got; else a[i + d] = got;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// dec_write_kernel: phase 4.  coef: int16[total_blocks][64] in RASTER order (u*8+v).
// !kFused (the default): every block's coefficients into the buffer; the inverse transform is two kernels of its own.
// kFused (TIC_DFLAG_FUSED, opt-in): a block whose symbols all lie in one subsequence never reaches the coefficient
// buffer — the warp that has just completed it (four blocks per pass, 8 lanes each) dequantises it, runs the FP32
// inverse transform with the guard band of dec_idct_fast_kernel and stores the pixels.  Only three kinds of block still
// go through the buffer and the work list of the exact pass: blocks that span subsequences (listed by
// dec_scan_kernel), blocks the stream ends in, and blocks with a pixel inside the guard band.  It removes 8.6 GB
// written + 8.6 GB read for the benchmark batch (five times the pixels) and is bit-identical — and it is SLOWER:
// 15.8 ms + 1.7 ms (exact pass) against 5.3 + 5.1 ms for the two separate kernels.  A warp completes about 4 blocks per
// symbol step, so a pass of ~330 warp instructions serves 2.9 blocks on average (113 per block; the thread-per-block
// kernel needs 49), inside a loop that was latency-bound, not bandwidth-bound, to begin with.  Kept for the record.
// ---------------------------------------------------------------------------------------------------
template <bool kFused>
__global__ void __launch_bounds__(kSyncThreads) dec_write_kernel(const DecImage* __restrict__ imgs,
                                                                 const long long* __restrict__ sub_first, int n_images,
                                                                 long long total_subs, const ShTables* __restrict__ deftab,
                                                                 const DecTables* __restrict__ tabs,
                                                                 const uint32_t* __restrict__ E, const int2* __restrict__ NB,
                                                                 int16_t* __restrict__ coef, int* __restrict__ status,
                                                                 int* __restrict__ summary, const double* __restrict__ mul,
                                                                 const float* __restrict__ mulf, long long* __restrict__ list,
                                                                 int* __restrict__ list_count, int list_cap) {
    __shared__ uint32_t rows[kSyncThreads][kRowWords];
    __shared__ ShTables sh_def;
    __shared__ uint8_t zz[64];    // zigzag index -> raster index u*8+v (direct stores into the coefficient buffer)
    __shared__ uint8_t zzs[64];   // zigzag index -> index in a slot: raster, or (kFused) transposed v*8+u: a slot row is a COLUMN
    __shared__ __align__(16) int16_t slots[kSyncThreads][64];   // one block under construction per thread
    for (int i = threadIdx.x; i < kSyncThreads * 8; i += kSyncThreads)
        reinterpret_cast<uint4*>(&slots[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    load_sh_tables(sh_def, deftab, kSyncThreads);
    if (threadIdx.x < 64) {
        const int r = c_zigzag[threadIdx.x];
        zz[threadIdx.x] = (uint8_t)r;
        zzs[threadIdx.x] = (uint8_t)(kFused ? (r & 7) * 8 + (r >> 3) : r);
    }
    long long g = (long long)blockIdx.x * kSyncThreads + threadIdx.x;
    int idx = find_owner_cta(sub_first, n_images, (long long)blockIdx.x * kSyncThreads, g, total_subs);   // barrier inside
    bool active = g < total_subs;
    int k = 0, end_rel = 0, nblk = 0;
    bool last_sub = false;
    int2 nb = make_int2(0, 0);
    const DecTables* tb = nullptr;   // per-image tables (global memory); null: the fixed tables in shared memory
    BitSrc src = {};
    long long blk0 = 0;   // index of the image's first block in the batch
    if (active) {
        const DecImage& im = imgs[idx];
        k = (int)(g - im.sub_first);
        active = !im.skip_entropy && k >= im.anchor_sub;
        if (active) {
            nb = NB[g];
            // past the last block: padding bits (to_bytes, bitbuffer.py:17-18) or trailing bytes
            if (nb.x > im.nblk) active = false;
        }
        if (im.has_tables) tb = tabs + idx;
        src = make_src(im);
        long long left = im.nbits - (128 + (long long)k * kSubBits);
        end_rel = left < kSubBits ? (int)left : kSubBits;
        nblk = im.nblk;
        last_sub = k + 1 >= im.nsubs;
        blk0 = im.blk_first;
    }
    stage_rows(rows, active, src, 4 + (long long)k * kSubWords);

    // Phase 4 proper, warp-synchronous: every step each running lane decodes ONE symbol; the lanes whose symbol
    // was the EOB of a block they started (the block sits in their slot) are then served by the whole warp,
    // four blocks per pass, 8 lanes x 16 bytes per block: 128-byte coalesced stores, zeros included, so the
    // coefficient buffer needs no zero fill and sees no partial-sector writes.  A block that spans subsequences is
    // written with 2-byte stores by every thread that decodes a piece of it (the thread that entered in the middle
    // of it directly, the thread that started it from its slot when its subsequence ends) into a block
    // dec_scan_kernel zeroed beforehand; the pieces are disjoint coefficients.
    const int lane = threadIdx.x & 31;
    const uint32_t* sw = rows[threadIdx.x];
    int16_t* slot = slots[threadIdx.x];
    const bool fixed_tabs = tb == nullptr;
    const DecTable* tdc = fixed_tabs ? nullptr : &tb->dc;
    const DecTable* tac = fixed_tabs ? nullptr : &tb->ac;
    const uint32_t sh_tabs = shared_addr(&sh_def), row = shared_addr(sw);
    uint32_t entry = active ? E[g] : 0u;
    int p = (int)(entry & 0xffffu), z = (int)((entry >> 16) & 0xffu);
    int blk = nb.x, dc_run = nb.y, cur = nb.x - 1, it = 0;
    bool staged = false;
    uint32_t err = 0;
    for (;;) {
        const bool running = active && p < end_rel && it < kMaxSymbols && !(z == 0 && blk >= nblk);
        if (!__any_sync(0xffffffffu, running)) break;
        bool ready = false;
        if (running) {
            it++;
            const uint32_t a = row + ((uint32_t)(p >> 5) << 2);
            const uint32_t v = __funnelshift_l(lds_u32(a + 4u), lds_u32(a), (uint32_t)(p & 31));
            int sym, len;
            if (fixed_tabs) lookup_sh(sh_tabs, z == 0, v, sym, len);   // warp-uniform but at image boundaries
            else lookup(z == 0 ? tdc : tac, v, sym, len);
            if (len < 0) {   // the same rules as decode_sub: this pass must follow the parse the entries belong to
                err |= TIC_DSTATUS_CODE;
                p += 1;
            } else {
                const int size = sym & 15;
                const int val = read_value(v << len, size);
                const int p0 = p;
                p += len + size;
                if (z == 0) {
                    cur = blk++;
                    dc_run += val;
                    int s = dc_run < -32768 ? -32768 : (dc_run > 32767 ? 32767 : dc_run);
                    if (s != dc_run) err |= TIC_DSTATUS_RANGE;
                    slot[0] = (int16_t)s;
                    staged = true;
                    z = 1;
                } else if (sym == 0) {
                    z = 0;
                    ready = staged;
                    staged = false;
                } else {
                    z += sym >> 4;
                    if (z > 63) {
                        err |= TIC_DSTATUS_CODE;
                        z = 0;
                        p = p0 + 1;
                        if (staged) {   // damaged stream: the block reads as zero
#pragma unroll
                            for (int i = 0; i < 8; i++) reinterpret_cast<uint4*>(slot)[i] = make_uint4(0u, 0u, 0u, 0u);
                            ready = true;
                            staged = false;
                        }
                    } else {
                        if (val != 0) {
                            if (staged) slot[zzs[z]] = (int16_t)val;
                            else if (cur >= 0 && cur < nblk) coef[(blk0 + cur) * 64 + zz[z]] = (int16_t)val;
                        }
                        z += 1;
                    }
                }
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, ready);
        __syncwarp();   // the lanes' stores into their slots are ordered before the other lanes' reads below (vote and
                        // shuffle do not order memory)
        const long long my_block = blk0 + cur;
        while (m) {
            // the (lane >> 3)-th of the lowest four ready lanes
            unsigned mm = m;
            int src_lane = -1;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int b = mm ? __ffs(mm) - 1 : -1;
                if (j == (lane >> 3)) src_lane = b;
                mm &= mm - 1;
            }
            m = mm;
            const long long dst_block = __shfl_sync(0xffffffffu, my_block, src_lane & 31);
            if constexpr (!kFused) {
                if (src_lane >= 0) {
                    uint4* sp = reinterpret_cast<uint4*>(slots[(threadIdx.x & ~31) + src_lane]) + (lane & 7);
                    reinterpret_cast<uint4*>(coef + dst_block * 64)[lane & 7] = *sp;
                    *sp = make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
                // ---- the group's block: column j of its coefficients -> row j of its pixels ----------------------
                const int j = lane & 7, gshift = lane & ~7;
                const bool valid = src_lane >= 0;
                const int sidx = __shfl_sync(0xffffffffu, idx, valid ? src_lane : lane);   // the block's image
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (valid) {
                    uint4* sp = reinterpret_cast<uint4*>(slots[(threadIdx.x & ~31) + src_lane]) + j;
                    q = *sp;
                    *sp = make_uint4(0u, 0u, 0u, 0u);
                }
                const DecImage& sim = imgs[sidx];
                const float4* mT = reinterpret_cast<const float4*>(mulf + (size_t)sidx * kMulfStride + 64 + j * 8);
                const float4 m0 = __ldg(mT), m1 = __ldg(mT + 1);
                float a[8];
                a[0] = (float)(int)(short)(q.x & 0xffffu) * m0.x; a[1] = (float)((int)q.x >> 16) * m0.y;
                a[2] = (float)(int)(short)(q.y & 0xffffu) * m0.z; a[3] = (float)((int)q.y >> 16) * m0.w;
                a[4] = (float)(int)(short)(q.z & 0xffffu) * m1.x; a[5] = (float)((int)q.z >> 16) * m1.y;
                a[6] = (float)(int)(short)(q.w & 0xffffu) * m1.z; a[7] = (float)((int)q.w >> 16) * m1.w;
                float S = ((fabsf(a[0]) + fabsf(a[1])) + (fabsf(a[2]) + fabsf(a[3]))) + ((fabsf(a[4]) + fabsf(a[5])) + (fabsf(a[6]) + fabsf(a[7])));
                S += __shfl_xor_sync(0xffffffffu, S, 1);
                S += __shfl_xor_sync(0xffffffffu, S, 2);
                S += __shfl_xor_sync(0xffffffffu, S, 4);
                const bool has_ac = ((j == 0 ? (q.x & 0xffff0000u) : q.x) | q.y | q.z | q.w) != 0u;
                const unsigned acm = (__ballot_sync(0xffffffffu, has_ac) >> gshift) & 0xffu;
                idct8_fast(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);   // down column j
                transpose8(a, j);
                idct8_fast(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);   // along row j
                // the same decision as dec_idct_fast_kernel: a pixel further than delta from every integer is final
                const float delta = fmaf(6e-7f, S, 4e-5f);
                float rmin = 1.0f;
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int v = 0; v < 8; v++) {
                    const float W = a[v] + 128.0f;
                    const float d = W - ((W + 12582912.0f) - 12582912.0f);   // W - rint(W), exact below 2^22
                    rmin = fminf(rmin, fabsf(d));
                    const uint32_t px = (uint32_t)min(max(__float2int_rz(W), 0), 255);
                    if (v < 4) lo |= px << (8 * v); else hi |= px << (8 * (v - 4));
                }
                const bool flag = !(delta < 0.25f) || rmin <= delta;
                unsigned flm = (__ballot_sync(0xffffffffu, flag) >> gshift) & 0xffu;
                const int dc_q = __shfl_sync(0xffffffffu, (int)(short)(q.x & 0xffffu), gshift);   // the group's lane 0 holds (0, 0)
                if (acm == 0) {   // DC only: idct8_exact on (T, 0, ..., 0) gives 0.25 RN(T sqrt2) everywhere; twice
                    const double SQ2 = 0x1.6a09e667f3bcdp+0;
                    double t = (double)dc_q;
                    if (sim.mode == MODE_SCALED) t = __dmul_rn(__ddiv_rn(t, d_ann[0]), sim.two_q);
                    t = __dmul_rn(t, __ldg(mul + (size_t)sidx * 64));
                    t = __dmul_rn(0.25, __dmul_rn(t, SQ2));
                    t = __dmul_rn(0.25, __dmul_rn(t, SQ2));
                    lo = hi = (uint32_t)min(max(__double2int_rz(__dadd_rn(t, 128.0)), 0), 255) * 0x01010101u;
                    flm = 0;
                }
                if (valid && !sim.skip_pixels) {
                    if (flm) {   // the exact pass decides: the block's coefficients (this lane: column j) and its index
                        int16_t* dst = coef + dst_block * 64 + j;
                        dst[0] = (int16_t)(q.x & 0xffffu); dst[8] = (int16_t)(q.x >> 16);
                        dst[16] = (int16_t)(q.y & 0xffffu); dst[24] = (int16_t)(q.y >> 16);
                        dst[32] = (int16_t)(q.z & 0xffffu); dst[40] = (int16_t)(q.z >> 16);
                        dst[48] = (int16_t)(q.w & 0xffffu); dst[56] = (int16_t)(q.w >> 16);
                        if (j == 0) list_push(list, list_count, list_cap, dst_block, summary);
                    } else {
                        const int b = (int)(dst_block - sim.blk_first);
                        const int by = b / sim.bw, bx = b - by * sim.bw;
                        store_pixel_row(sim, by * 8 + j, bx * 8, lo, hi);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (staged) {
        // slot index i -> raster index (kFused: the slot is transposed)
        if (last_sub) {   // the stream ends inside the block (truncated): nobody else writes it
            for (int i = 0; i < 64; i++) coef[(blk0 + cur) * 64 + (kFused ? (i & 7) * 8 + (i >> 3) : i)] = slot[i];
            if constexpr (kFused) list_push(list, list_count, list_cap, blk0 + cur, summary);
        } else {          // the block goes on in the next subsequence (dec_scan_kernel zeroed and, kFused, listed it)
            for (int i = 0; i < 64; i++) {
                int16_t c = slot[i];
                if (c != 0) coef[(blk0 + cur) * 64 + (kFused ? (i & 7) * 8 + (i >> 3) : i)] = c;
            }
        }
    }
    if (active && it >= kMaxSymbols) err |= TIC_DSTATUS_CODE;
    if (err) { atomicOr(&status[idx], (int)err); atomicOr(summary, (int)err); }
}

// ---------------------------------------------------------------------------------------------------
// Phase 5: scipy.fftpack.idct(x, norm="ortho") for N = 8 in float64 = ducc0's DCT-III (T_dcst23, type 3):
// c0 *= sqrt2; twiddle pre-pass; c4 *= 2 tw3; forward real FFT of length 8 (radf4, then radf2); scale
// 1/sqrt(2N) = 0.25 (exact); minus/plus post-pass.  One IEEE operation per intrinsic, never contracted;
// the constants are ducc0's, not the correctly rounded cosines.  The test suite
// checks the same operation sequence against the installed SciPy bit for bit on the CPU.
// ---------------------------------------------------------------------------------------------------
#define DA(a, b) __dadd_rn((a), (b))
#define DS(a, b) __dsub_rn((a), (b))
#define DM(a, b) __dmul_rn((a), (b))

__device__ __forceinline__ void idct8_exact(double& x0, double& x1, double& x2, double& x3, double& x4, double& x5,
                                            double& x6, double& x7) {
    const double TW0 = 0x1.f6297cff75cb0p-1, TW1 = 0x1.d906bcf328d46p-1, TW2 = 0x1.a9b66290ea1a3p-1,
                 TW3 = 0x1.6a09e667f3bccp-1, TW4 = 0x1.1c73b39ae68c8p-1, TW5 = 0x1.87de2a6aea963p-2,
                 TW6 = 0x1.8f8b83c69a60ap-3;
    const double WR = 0x1.6a09e667f3bccp-1, WI = 0x1.6a09e667f3bcdp-1, SQ2 = 0x1.6a09e667f3bcdp+0;
    double c0 = DM(x0, SQ2);
    double t1 = DA(x1, x7), t2 = DS(x1, x7);
    double c1 = DA(DM(TW0, t2), DM(TW6, t1)), c7 = DS(DM(TW0, t1), DM(TW6, t2));
    t1 = DA(x2, x6); t2 = DS(x2, x6);
    double c2 = DA(DM(TW1, t2), DM(TW5, t1)), c6 = DS(DM(TW1, t1), DM(TW5, t2));
    t1 = DA(x3, x5); t2 = DS(x3, x5);
    double c3 = DA(DM(TW2, t2), DM(TW4, t1)), c5 = DS(DM(TW2, t1), DM(TW4, t2));
    double c4 = DM(x4, 2.0 * TW3);
    // radf4
    double atr1 = DA(c6, c2), h2 = DS(c6, c2), atr2 = DA(c0, c4), h1 = DS(c0, c4);
    double h0 = DA(atr2, atr1), h3 = DS(atr2, atr1);
    double btr1 = DA(c7, c3), h6 = DS(c7, c3), btr2 = DA(c1, c5), h5 = DS(c1, c5);
    double h4 = DA(btr2, btr1), h7 = DS(btr2, btr1);
    // radf2
    double r0 = DA(h0, h4), r7 = DS(h0, h4);
    double r4 = -h7, r3 = h3;
    double tr2 = DA(DM(WR, h5), DM(WI, h6)), ti2 = DS(DM(WR, h6), DM(WI, h5));
    double r1 = DA(h1, tr2), r5 = DS(h1, tr2);
    double r2 = DA(ti2, h2), r6 = DS(ti2, h2);
    // scale (exact) and post-pass
    double s0 = DM(0.25, r0), s1 = DM(0.25, r1), s2 = DM(0.25, r2), s3 = DM(0.25, r3);
    double s4 = DM(0.25, r4), s5 = DM(0.25, r5), s6 = DM(0.25, r6), s7 = DM(0.25, r7);
    x0 = s0;
    x1 = DS(s1, s2); x2 = DA(s1, s2);
    x3 = DS(s3, s4); x4 = DA(s3, s4);
    x5 = DS(s5, s6); x6 = DA(s5, s6);
    x7 = s7;
}

// ---------------------------------------------------------------------------------------------------
// Phase 5a, the fast pass: one thread per block, FP32.  Pixels are trunc(clip(X + 128)) of the reference's float64
// result X, so an FP32 approximation Y gives the same pixel whenever no integer lies between Y + 128 and X + 128.
// Y is the separable transform written as 4-term FMA dot products (even / odd halves), for which the classical
// bound holds: with u = 2^-24, S = sum |dequantised coefficients| and all |cosine factors| <= 0.4904,
//     |Y - X*| <= u S (9 * 0.2405 + 11.5 * 0.2405) = 4.93 u S = 2.94e-7 S          (X* = the exact real transform)
// (per pass: gamma_8 for the accumulation, u for the rounded constant, 2.5 u for the FP32 dequantised input),
// |X - X*| < 1e-11 for the float64 chain, and forming Y + 128 in FP32 adds at most 3.1e-5 below 512.  A pixel whose
// FP32 value is further than delta = 6e-7 S + 4e-5 (twice the bound) from every integer is therefore final; a block
// with any pixel inside the band goes to the exact float64 kernel through a work list (dec_idct_kernel below, which
// with TIC_DFLAG_EXACT_ONLY transforms every block — tests compare the two paths bit for bit on whole batches).
// Blocks with no AC coefficient at all (flat areas — where DC * 16 / 8 lands on an integer every time) are settled
// here exactly: ducc0's operation sequence on (T, 0, ..., 0) collapses to two multiplications by sqrt2 and two
// exact scalings by 0.25 per pass.
// ---------------------------------------------------------------------------------------------------
// Two floats in an aligned register pair: the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2) do two
// independent IEEE operations per issue slot — the same roundings as the scalar forms, so the error bound above
// is unchanged.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float x, float y) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {   // FMNMX3
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// np.clip(x, 0, 255).astype(np.uint8) of four already truncated values, p0 in the low byte
__device__ __forceinline__ uint32_t pack_px4(int p0, int p1, int p2, int p3) {
    uint32_t t, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(p3), "r"(p2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(p1), "r"(p0), "r"(t));
    return r;
}

// idct8_fast (below) on two independent 8-point inputs at once
__device__ __forceinline__ void idct8_fast2(f32x2& x0, f32x2& x1, f32x2& x2, f32x2& x3, f32x2& x4, f32x2& x5, f32x2& x6,
                                            f32x2& x7) {
    const f32x2 A = pk2(0.35355339059327379f, 0.35355339059327379f), B = pk2(0.46193976625564337f, 0.46193976625564337f),
                C = pk2(0.19134171618254489f, 0.19134171618254489f), P = pk2(0.49039264020161522f, 0.49039264020161522f),
                Q = pk2(0.41573480615127262f, 0.41573480615127262f), R = pk2(0.27778511650980111f, 0.27778511650980111f),
                T = pk2(0.09754516100806413f, 0.09754516100806413f);
    const f32x2 nA = pk2(-0.35355339059327379f, -0.35355339059327379f), nB = pk2(-0.46193976625564337f, -0.46193976625564337f),
                nC = pk2(-0.19134171618254489f, -0.19134171618254489f), nP = pk2(-0.49039264020161522f, -0.49039264020161522f),
                nR = pk2(-0.27778511650980111f, -0.27778511650980111f), nT = pk2(-0.09754516100806413f, -0.09754516100806413f);
    const f32x2 a0 = mul2(A, x0);
    const f32x2 e0 = fma2(C, x6, fma2(A, x4, fma2(B, x2, a0)));
    const f32x2 e1 = fma2(nB, x6, fma2(nA, x4, fma2(C, x2, a0)));
    const f32x2 e2 = fma2(B, x6, fma2(nA, x4, fma2(nC, x2, a0)));
    const f32x2 e3 = fma2(nC, x6, fma2(A, x4, fma2(nB, x2, a0)));
    const f32x2 o0 = fma2(T, x7, fma2(R, x5, fma2(Q, x3, mul2(P, x1))));
    const f32x2 o1 = fma2(nR, x7, fma2(nP, x5, fma2(nT, x3, mul2(Q, x1))));
    const f32x2 o2 = fma2(Q, x7, fma2(T, x5, fma2(nP, x3, mul2(R, x1))));
    const f32x2 o3 = fma2(nP, x7, fma2(Q, x5, fma2(nR, x3, mul2(T, x1))));
    x0 = add2(e0, o0); x7 = sub2(e0, o0);
    x1 = add2(e1, o1); x6 = sub2(e1, o1);
    x2 = add2(e2, o2); x5 = sub2(e2, o2);
    x3 = add2(e3, o3); x4 = sub2(e3, o3);
}

__global__ void __launch_bounds__(128, TICD_FAST_MIN_CTAS) dec_idct_fast_kernel(const DecImage* __restrict__ imgs,
                                                            const long long* __restrict__ blk_first, int n_images,
                                                            long long total_blocks, const int16_t* __restrict__ coef,
                                                            const double* __restrict__ mul, const float* __restrict__ mulf,
                                                            const int* __restrict__ ndec, long long* __restrict__ list,
                                                            int* __restrict__ list_count) {
    // the warp's 32 blocks are 4 KB of consecutive coefficients: fetched with coalesced 16-byte loads into shared
    // memory (rows padded to 9 x 16 bytes: conflict-free both ways), then every thread takes its own block
    __shared__ uint4 stage[4][32][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gb = (long long)blockIdx.x * 128 + threadIdx.x;
    {
        const long long warp_blk0 = (long long)blockIdx.x * 128 + warp * 32;
        const uint4* src = reinterpret_cast<const uint4*>(coef) + warp_blk0 * 8;
        const long long limit = (total_blocks - warp_blk0) * 8;   // 16-byte pieces that exist
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int i = j * 32 + lane;
            stage[warp][i >> 3][i & 7] = i < limit ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
    int idx = find_owner_cta(blk_first, n_images, (long long)blockIdx.x * 128, gb, total_blocks);   // barrier inside
    if (gb >= total_blocks) return;
    const DecImage& im = imgs[idx];
    if (im.skip_pixels) return;
    const int b = (int)(gb - im.blk_first);
    const int by = b / im.bw, bx = b - by * im.bw;
    const int y0 = by * 8, x0 = bx * 8;

    uint32_t w[32];
    {
        const bool have = b < __ldg(ndec + idx);   // a block the stream never reached reads as zero
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint4 q = have ? stage[warp][lane][u] : make_uint4(0u, 0u, 0u, 0u);
            w[u * 4] = q.x; w[u * 4 + 1] = q.y; w[u * 4 + 2] = q.z; w[u * 4 + 3] = q.w;
        }
    }
    uint32_t ac = w[0] & 0xffff0000u;
#pragma unroll
    for (int i = 1; i < 32; i++) ac |= w[i];
    if (ac == 0) {
        // DC only: idct8_exact on (T, 0, ..., 0) gives 0.25 RN(T sqrt2) everywhere; twice
        const double SQ2 = 0x1.6a09e667f3bcdp+0;
        double t = (double)(int)(short)(w[0] & 0xffffu);
        if (im.mode == MODE_SCALED) t = DM(__ddiv_rn(t, d_ann[0]), im.two_q);
        t = DM(t, __ldg(mul + (size_t)idx * 64));
        t = DM(0.25, DM(t, SQ2));
        t = DM(0.25, DM(t, SQ2));
        const uint32_t px = (uint32_t)min(max(__double2int_rz(DA(t, 128.0)), 0), 255) * 0x01010101u;
#pragma unroll
        for (int u = 0; u < 8; u++) store_pixel_row(im, y0 + u, x0, px, px);
        return;
    }
    // pairs of horizontally adjacent coefficients — the two int16 of one coefficient word — through the column pass
    const float2* __restrict__ fm = reinterpret_cast<const float2*>(mulf + (size_t)idx * kMulfStride);
    f32x2 c2[8][4];
    float S = 0.0f;
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int vp = 0; vp < 4; vp++) {
            const uint32_t word = w[u * 4 + vp];
            const float2 m2 = __ldg(fm + u * 4 + vp);
            const float a = (float)(int)(short)(word & 0xffffu) * m2.x, bb = (float)((int)word >> 16) * m2.y;
            S += fabsf(a);
            S += fabsf(bb);
            c2[u][vp] = pk2(a, bb);
        }
    }
#pragma unroll
    for (int vp = 0; vp < 4; vp++)   // down the columns, two columns per instruction
        idct8_fast2(c2[0][vp], c2[1][vp], c2[2][vp], c2[3][vp], c2[4][vp], c2[5][vp], c2[6][vp], c2[7][vp]);
    const float delta = fmaf(6e-7f, S, 4e-5f);
    bool flag = !(delta < 0.25f);   // also catches magnitudes where the rounding trick below stops being valid
    float rmin = 1.0f;
    uint32_t lo[8], hi[8];
    const f32x2 k128 = pk2(128.0f, 128.0f), kmag = pk2(12582912.0f, 12582912.0f);
#pragma unroll
    for (int up = 0; up < 4; up++) {   // along the rows, two rows per instruction: re-pair (row 2up, row 2up+1)
        f32x2 r2[8];
#pragma unroll
        for (int vp = 0; vp < 4; vp++) {
            float a0, a1, b0, b1;
            upk2(c2[2 * up][vp], a0, a1);
            upk2(c2[2 * up + 1][vp], b0, b1);
            r2[2 * vp] = pk2(a0, b0);
            r2[2 * vp + 1] = pk2(a1, b1);
        }
        idct8_fast2(r2[0], r2[1], r2[2], r2[3], r2[4], r2[5], r2[6], r2[7]);
        // truncate (F2I), then clamp and pack four pixels with two saturating pack instructions (the first version
        // spent 9 instructions per pixel here: a third of the kernel)
#pragma unroll
        for (int half = 0; half < 2; half++) {
            int i0[4], i1[4];
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const f32x2 W2 = add2(r2[4 * half + v], k128);
                const f32x2 d2 = sub2(W2, sub2(add2(W2, kmag), kmag));   // W - rint(W), exact below 2^22
                float W0, W1, d0, d1;
                upk2(W2, W0, W1);
                upk2(d2, d0, d1);
                rmin = fmin3(rmin, fabsf(d0), fabsf(d1));
                i0[v] = __float2int_rz(W0);
                i1[v] = __float2int_rz(W1);
            }
            const uint32_t a = pack_px4(i0[0], i0[1], i0[2], i0[3]), b = pack_px4(i1[0], i1[1], i1[2], i1[3]);
            if (half == 0) { lo[2 * up] = a; lo[2 * up + 1] = b; } else { hi[2 * up] = a; hi[2 * up + 1] = b; }
        }
    }
    flag |= rmin <= delta;
    if (flag) {
        list[atomicAdd(list_count, 1)] = gb;
        return;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) store_pixel_row(im, y0 + u, x0, lo[u], hi[u]);
}

// ---------------------------------------------------------------------------------------------------
// Phase 5b, the exact pass.  8 lanes per 8x8 block (4 blocks per warp, kIdctBlocks per CTA): lane t dequantises
// row t of the coefficients, the block goes through a padded shared-memory tile, lane t transforms column t, back
// through the tile, lane t transforms row t and stores its 8 pixels.  `list` given: only the blocks the fast pass
// listed (their number is read from the device); otherwise every block.
// ---------------------------------------------------------------------------------------------------
constexpr int kIdctBlocks = 32;   // blocks in flight per CTA (8 lanes each)
constexpr int kIdctIters = 16;    // consecutive groups of kIdctBlocks per CTA: one owner search per 512 blocks

__global__ void __launch_bounds__(kIdctBlocks * 8) dec_idct_kernel(const DecImage* __restrict__ imgs,
                                                                   const long long* __restrict__ blk_first, int n_images,
                                                                   long long total_blocks, const int16_t* __restrict__ coef,
                                                                   const double* __restrict__ mul, const int* __restrict__ ndec,
                                                                   const long long* __restrict__ list,
                                                                   const int* __restrict__ list_count, int list_cap) {
    __shared__ double tile[kIdctBlocks][8][9];   // 9: column and row accesses both conflict-free
    const bool listed = list != nullptr;
    const long long total = listed ? (long long)min(*list_count, list_cap) : total_blocks;
    const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
    const long long cta0 = (long long)blockIdx.x * (kIdctBlocks * kIdctIters);
    if (cta0 >= total) return;
    int idx = listed ? 0 : find_owner_cta(blk_first, n_images, cta0, cta0 + lb, total);   // barrier inside
    const unsigned group = 0xffu << ((threadIdx.x & 31) & ~7);   // the 8 lanes of this block
#pragma unroll 1
    for (int iter = 0; iter < kIdctIters; iter++) {
        const long long li = cta0 + (long long)iter * kIdctBlocks + lb;
        if (li >= total) return;   // whole 8-lane groups leave together; only __syncwarp(group) follows
        const long long gb = listed ? list[li] : li;
        if (listed) idx = find_owner(blk_first, n_images, gb);
        else while (idx + 1 < n_images && __ldg(blk_first + idx + 1) <= gb) idx++;
        const DecImage& im = imgs[idx];
        if (im.skip_pixels) continue;
        const int b = (int)(gb - im.blk_first);
        const int by = b / im.bw, bx = b - by * im.bw;
        const double* __restrict__ m = mul + (size_t)idx * 64 + t * 8;
        const bool scaled = im.mode == MODE_SCALED;
        const double two_q = im.two_q;

        double x[8];
        double (*tl)[9] = tile[lb];
        {
            // row t: 8 int16; a block the stream never reached (truncated, or every block zero bits long) is zero
            uint4 q = b < __ldg(ndec + idx) ? __ldg(reinterpret_cast<const uint4*>(coef + gb * 64) + t)
                                            : make_uint4(0u, 0u, 0u, 0u);
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int v = 0; v < 8; v++) {
                int c = (int)(int16_t)((w[v >> 1] >> ((v & 1) * 16)) & 0xffffu);
                double d = (double)c;
                if (scaled) d = DM(__ddiv_rn(d, d_ann[t * 8 + v]), two_q);   // codec.py:59-61
                tl[t][v] = DM(d, __ldg(m + v));                              // utils.py:51-52
            }
        }
        __syncwarp(group);
        // utils.py:40-45: axis -2 (down the columns) first, then axis -1 (along the rows)
#pragma unroll
        for (int u = 0; u < 8; u++) x[u] = tl[u][t];
        idct8_exact(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
        __syncwarp(group);
#pragma unroll
        for (int u = 0; u < 8; u++) tl[u][t] = x[u];
        __syncwarp(group);
#pragma unroll
        for (int v = 0; v < 8; v++) x[v] = tl[t][v];
        __syncwarp(group);   // the tile is rewritten by the next iteration
        idct8_exact(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int v = 0; v < 8; v++) {
            // np.clip(coeffs + 128, 0, 255) then astype(np.uint8) (codec.py:68-70) = truncate, then clamp: the two
            // commute on [-inf, inf] (trunc is monotonic and fixes 0 and 255), and the clamp is integer work
            int pi = __double2int_rz(DA(x[v], 128.0));
            uint32_t px = (uint32_t)min(max(pi, 0), 255);
            if (v < 4) lo |= px << (8 * v); else hi |= px << (8 * (v - 4));
        }
        store_pixel_row(im, by * 8 + t, bx * 8, lo, hi);
    }
}

// Fused coefficient pass: the blocks a stream never reaches (truncated streams; images whose every block is zero bits
// long) have all-zero coefficients, i.e. pixel 128 everywhere (idct8_exact of zeros is exactly zero in both modes).
// grid (images, chunks); returns at once for every image that was decoded completely.
__global__ void __launch_bounds__(256) dec_fill_kernel(const DecImage* __restrict__ imgs, const int* __restrict__ ndec) {
    const DecImage& im = imgs[blockIdx.x];
    const int first = ndec[blockIdx.x];
    if (im.skip_pixels || first >= im.nblk) return;
    for (long long b = (long long)first + blockIdx.y * 256 + threadIdx.x; b < im.nblk; b += (long long)gridDim.y * 256) {
        const int by = (int)(b / im.bw), bx = (int)(b - (long long)by * im.bw);
#pragma unroll
        for (int u = 0; u < 8; u++) store_pixel_row(im, by * 8 + u, bx * 8, 0x80808080u, 0x80808080u);
    }
}

// ---------------------------------------------------------------------------------------------------
// decode() from coefficient arrays (codec.py:46-70): dc[nblk] differences, ac[nblk][63] in zigzag order.
// The DC cumsum (codec.py:53) reuses dec_scan_kernel with one "subsequence" per block.
// ---------------------------------------------------------------------------------------------------
__global__ void dec_coeffs_prep_kernel(DecImage* __restrict__ im, double* __restrict__ mul, float* __restrict__ mulf,
                                       const int32_t* __restrict__ dc, int2* __restrict__ ND, int nblk,
                                       int* __restrict__ summary) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) {
        uint32_t st = fill_mul(*im, mul, mulf);
        if (st) { im->skip_pixels = 1; atomicOr(summary, (int)st); }
    }
    if (b < nblk) ND[b] = make_int2(1, dc[b]);
}

__global__ void dec_coeffs_pack_kernel(const int32_t* __restrict__ dc, const int32_t* __restrict__ ac,
                                       const int2* __restrict__ NB, int16_t* __restrict__ coef, int nblk,
                                       int* __restrict__ summary) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int b = (int)(t >> 6), k = (int)(t & 63);
    if (b >= nblk) return;
    int v = k == 0 ? NB[b].y + dc[b] : ac[(long long)b * 63 + k - 1];
    int s = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
    if (s != v) atomicOr(summary, TIC_DSTATUS_RANGE);
    coef[(long long)b * 64 + c_zigzag[k]] = (int16_t)s;   // coeffs[:, ZIGZAG_ORDER] = coeffs.copy(), codec.py:57
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct DecWs {
    DecImage* d_imgs = nullptr; DecImage* h_imgs = nullptr; size_t imgs_cap = 0;
    long long* d_first = nullptr; long long* h_first = nullptr;   // sub_first[n+1] then blk_first[n+1]
    DecTables* d_tabs = nullptr; double* d_mul = nullptr; float* d_mulf = nullptr;
    ShTables* d_deftab = nullptr;   // the fixed tables in the two-level form the kernels keep in shared memory
    uint32_t* d_E = nullptr; uint32_t* d_U = nullptr; int2* d_ND = nullptr; int2* d_NB = nullptr; size_t subs_cap = 0;
    int16_t* d_coef = nullptr; long long* d_list = nullptr; size_t blocks_cap = 0;   // coefficients; exact-pass work list
    int list_cap = 0;         // entries of d_list (2 x blocks_cap: spanning + flagged blocks, see list_push)
    int* d_flags = nullptr;   // [0] changed, [1] summary, [2] blocks listed for the exact IDCT pass
    int* h_flags = nullptr;   // pinned
    int* d_status_own = nullptr; size_t status_cap = 0;
    int* d_ndec = nullptr;    // per image: blocks whose coefficients were written (the rest read as zero)
    ScanSlice* d_slices = nullptr; int2* d_slice_tot = nullptr; size_t slices_cap = 0;
    std::vector<ScanSlice> h_slices;
    // single-stream host path
    uint8_t* d_stream = nullptr; size_t stream_cap = 0;
    uint8_t* d_px = nullptr; size_t px_cap = 0;
    cudaEvent_t ev[6] = {};
    long long stats[12] = {};
    bool pending_times = false;
    // the batch tic_decode_batch enqueued last (tic_decode_finish may have to run its rounds again, see run_batch)
    struct {
        bool active = false, speculative = false;
        int n_images = 0;
        long long subs = 0, blocks = 0;
        uint32_t flags = 0;
        int* status = nullptr;
    } run;
    int spec_rounds = 2;   // synchronisation launches enqueued without looking at their outcome; raised when a batch needed more
};

}  // namespace ticd

using namespace ticd;

#define TICD_CUDA(h, call)                                                                           \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            tic_internal_set_error((h), std::string(#call) + ": " + cudaGetErrorString(e_));        \
            return TIC_E_CUDA;                                                                       \
        }                                                                                            \
    } while (0)

static DecWs* get_ws(tic_handle h) {
    void** slot = tic_internal_dec_slot(h);
    if (!*slot) *slot = new DecWs();
    return static_cast<DecWs*>(*slot);
}

void tic_internal_dec_release(void* p) {
    DecWs* w = static_cast<DecWs*>(p);
    if (!w) return;
    cudaFree(w->d_imgs); cudaFreeHost(w->h_imgs); cudaFree(w->d_first); cudaFreeHost(w->h_first);
    cudaFree(w->d_tabs); cudaFree(w->d_mul); cudaFree(w->d_mulf); cudaFree(w->d_list); cudaFree(w->d_deftab); cudaFree(w->d_E); cudaFree(w->d_U);
    cudaFree(w->d_ND); cudaFree(w->d_NB); cudaFree(w->d_coef); cudaFree(w->d_flags); cudaFreeHost(w->h_flags);
    cudaFree(w->d_status_own); cudaFree(w->d_stream); cudaFree(w->d_px); cudaFree(w->d_slices); cudaFree(w->d_slice_tot); cudaFree(w->d_ndec);
    for (auto& e : w->ev) if (e) cudaEventDestroy(e);
    delete w;
}

int tic_parse_header(const uint8_t* data, int64_t nbytes, int32_t* height, int32_t* width, int32_t* quality,
                     uint32_t* flag) {
    if (!data || nbytes < 16) return TIC_E_INVALID;   // struct.error in the reference (codec.py:119)
    uint32_t f[4];
    for (int i = 0; i < 4; i++)
        f[i] = (uint32_t)data[4 * i] | ((uint32_t)data[4 * i + 1] << 8) | ((uint32_t)data[4 * i + 2] << 16) |
               ((uint32_t)data[4 * i + 3] << 24);
    if (f[0] > 0x7fffffffu || f[1] > 0x7fffffffu) return TIC_E_INVALID;
    if (height) *height = (int32_t)f[0];
    if (width) *width = (int32_t)f[1];
    if (quality) *quality = (int32_t)f[2];
    if (flag) *flag = f[3];
    return TIC_OK;
}

static int ensure_base(tic_handle h, DecWs* w) {
    if (w->d_flags) return TIC_OK;
    TICD_CUDA(h, cudaMalloc(&w->d_flags, 4 * sizeof(int)));
    TICD_CUDA(h, cudaMallocHost(&w->h_flags, 4 * sizeof(int)));
    for (auto& e : w->ev) TICD_CUDA(h, cudaEventCreate(&e));
    DecTables* def = new DecTables();
    ShTables* sh = new ShTables();
    build_default_tables(*def);
    const bool fits = build_sh_tables(*def, *sh);
    cudaError_t e1 = cudaSuccess;
    if (fits) {
        e1 = cudaMalloc(&w->d_deftab, sizeof(ShTables));
        if (e1 == cudaSuccess) e1 = cudaMemcpy(w->d_deftab, sh, sizeof(ShTables), cudaMemcpyHostToDevice);
    }
    delete def;
    delete sh;
    if (!fits) {
        tic_internal_set_error(h, "the fixed Huffman tables need more than 8 second-level blocks");
        return TIC_E_INVALID;
    }
    TICD_CUDA(h, e1);
    return TIC_OK;
}

template <typename T>
static int grow(tic_handle h, T*& p, size_t n) {
    cudaFree(p);
    p = nullptr;
    TICD_CUDA(h, cudaMalloc(&p, n * sizeof(T)));
    return TIC_OK;
}

// Phase 3 on `stream`: h_slices must list the slices of the images in image order.
static int launch_scan(tic_handle h, DecWs* w, int* status, cudaStream_t stream, long long& launches,
                       const uint32_t* E, int16_t* coef, bool list_spanning = false) {
    const size_t ns = w->h_slices.size();
    if (ns == 0) return TIC_OK;
    if (ns > w->slices_cap) {
        size_t cap = ns * 2 + 64;
        w->slices_cap = 0;
        if (int rc = grow(h, w->d_slices, cap)) return rc;
        if (int rc = grow(h, w->d_slice_tot, cap)) return rc;
        w->slices_cap = cap;
    }
    bool multi = false;
    for (const ScanSlice& sl : w->h_slices) multi |= sl.index_in_img > 0;
    TICD_CUDA(h, cudaMemcpyAsync(w->d_slices, w->h_slices.data(), ns * sizeof(ScanSlice), cudaMemcpyHostToDevice, stream));
    if (multi) {
        dec_scan_totals_kernel<<<(unsigned)ns, 1024, 0, stream>>>(w->d_imgs, w->d_slices, w->d_ND, w->d_slice_tot);
        launches++;
    }
    dec_scan_kernel<<<(unsigned)ns, 1024, 0, stream>>>(w->d_imgs, w->d_slices, w->d_slice_tot, w->d_ND, w->d_NB, status,
                                                       w->d_flags + 1, E, coef, w->d_ndec,
                                                       list_spanning ? w->d_list : nullptr, w->d_flags + 2, w->list_cap);
    launches++;
    TICD_CUDA(h, cudaGetLastError());
    return TIC_OK;
}

// Phase 5 on `stream`: the FP32 pass with its work list for the exact pass, or (TIC_DFLAG_EXACT_ONLY) the exact
// float64 pass over every block.
static int launch_idct(tic_handle h, DecWs* w, const long long* d_blk_first, int n_images, long long blocks, uint32_t flags,
                       cudaStream_t stream, long long& launches) {
    if (blocks == 0) return TIC_OK;
    const unsigned exact_grid = (unsigned)((blocks + kIdctBlocks * kIdctIters - 1) / (kIdctBlocks * kIdctIters));
    if (flags & TIC_DFLAG_EXACT_ONLY) {
        dec_idct_kernel<<<exact_grid, kIdctBlocks * 8, 0, stream>>>(w->d_imgs, d_blk_first, n_images, blocks, w->d_coef, w->d_mul,
                                                                    w->d_ndec, nullptr, nullptr, 0);
        launches++;
    } else {
        dec_idct_fast_kernel<<<(unsigned)((blocks + 127) / 128), 128, 0, stream>>>(w->d_imgs, d_blk_first, n_images, blocks,
                                                                                 w->d_coef, w->d_mul, w->d_mulf, w->d_ndec,
                                                                                 w->d_list, w->d_flags + 2);
        dec_idct_kernel<<<exact_grid, kIdctBlocks * 8, 0, stream>>>(w->d_imgs, d_blk_first, n_images, blocks, w->d_coef, w->d_mul,
                                                                    w->d_ndec, w->d_list, w->d_flags + 2, w->list_cap);
        launches += 2;
    }
    TICD_CUDA(h, cudaGetLastError());
    return TIC_OK;
}

static void push_slices(std::vector<ScanSlice>& v, int img, long long nsubs) {
    int j = 0;
    for (long long f = 0; f < nsubs; f += kSliceSubs, j++) {
        long long c = nsubs - f < kSliceSubs ? nsubs - f : kSliceSubs;
        v.push_back(ScanSlice{img, (int)f, (int)c, j});
    }
}

// Everything of a batch that runs on the device once descriptors, E / U / ND and the status words are in place:
// header + tables, synchronisation rounds, scan, coefficient pass, inverse transform.
// speculative: `spec_rounds` synchronisation launches are enqueued without reading their outcome, so the call is
// asynchronous (no host synchronisation; it could be captured in a graph); the flag of the LAST of them — did it still
// repair an entry state? — travels to the host with the batch's summary, and tic_decode_finish runs the batch again
// from the state it reached, this time reading the flag after every round (speculative = false), if it did.  Two
// launches settle everything measured except uniform noise at q >= 50 (three); a handle remembers what it needed.
static int run_batch(tic_handle h, DecWs* w, cudaStream_t stream, bool speculative) {
    const int n_images = w->run.n_images;
    const size_t n = (size_t)n_images;
    const long long subs = w->run.subs, blocks = w->run.blocks;
    const uint32_t flags = w->run.flags;
    int* status = w->run.status;
    const long long* d_sub_first = w->d_first;
    const long long* d_blk_first = w->d_first + (n + 1);
    TICD_CUDA(h, cudaEventRecord(w->ev[0], stream));
    TICD_CUDA(h, cudaMemsetAsync(w->d_ndec, 0, n * sizeof(int), stream));   // the coefficient buffer itself needs no fill
    long long launches = 0;
    dec_setup_kernel<<<n_images, 32, 0, stream>>>(w->d_imgs, n_images, flags, w->d_tabs, w->d_mul, w->d_mulf, w->d_E, status,
                                                  w->d_flags + 1);
    launches++;
    TICD_CUDA(h, cudaGetLastError());
    TICD_CUDA(h, cudaEventRecord(w->ev[1], stream));
    long long rounds = 0;
    const bool fused = (flags & TIC_DFLAG_FUSED) != 0 && (flags & TIC_DFLAG_EXACT_ONLY) == 0;
    if (subs) {
        unsigned grid = (unsigned)((subs + kSyncThreads - 1) / kSyncThreads);
        for (;;) {
            if (speculative && rounds == w->spec_rounds - 1) TICD_CUDA(h, cudaMemsetAsync(w->d_flags, 0, sizeof(int), stream));
            dec_sync_kernel<<<grid, kSyncThreads, 0, stream>>>(w->d_imgs, d_sub_first, n_images, subs, w->d_deftab,
                                                               w->d_tabs, w->d_E, w->d_U, w->d_ND, w->d_flags,
                                                               (flags & TIC_DFLAG_NO_EARLY_STOP) ? 0 : 1);
            launches++;
            rounds++;
            TICD_CUDA(h, cudaGetLastError());
            if (speculative) {
                if (rounds >= w->spec_rounds) break;
                continue;
            }
            TICD_CUDA(h, cudaMemcpyAsync(w->h_flags, w->d_flags, sizeof(int), cudaMemcpyDeviceToHost, stream));
            TICD_CUDA(h, cudaMemsetAsync(w->d_flags, 0, sizeof(int), stream));
            TICD_CUDA(h, cudaStreamSynchronize(stream));
            if (!w->h_flags[0]) break;
            if (rounds > subs + 2) {
                tic_internal_set_error(h, "tic_decode_batch: synchronisation did not converge");
                return TIC_E_CUDA;
            }
        }
        TICD_CUDA(h, cudaEventRecord(w->ev[2], stream));
        if (int rc = launch_scan(h, w, status, stream, launches, w->d_E, w->d_coef, fused)) return rc;
        TICD_CUDA(h, cudaEventRecord(w->ev[3], stream));
        if (fused)
            dec_write_kernel<true><<<grid, kSyncThreads, 0, stream>>>(w->d_imgs, d_sub_first, n_images, subs, w->d_deftab,
                                                                      w->d_tabs, w->d_E, w->d_NB, w->d_coef, status,
                                                                      w->d_flags + 1, w->d_mul, w->d_mulf, w->d_list,
                                                                      w->d_flags + 2, w->list_cap);
        else
            dec_write_kernel<false><<<grid, kSyncThreads, 0, stream>>>(w->d_imgs, d_sub_first, n_images, subs, w->d_deftab,
                                                                       w->d_tabs, w->d_E, w->d_NB, w->d_coef, status,
                                                                       w->d_flags + 1, nullptr, nullptr, nullptr, nullptr, 0);
        launches++;
        TICD_CUDA(h, cudaGetLastError());
    } else {
        TICD_CUDA(h, cudaEventRecord(w->ev[2], stream));
        TICD_CUDA(h, cudaEventRecord(w->ev[3], stream));
    }
    TICD_CUDA(h, cudaEventRecord(w->ev[4], stream));
    if (fused) {
        // what the coefficient pass did not finish itself: unreached blocks (pixel 128), then the exact pass over the
        // listed blocks (spanning subsequences, guard band, end of a truncated stream)
        if (blocks) {
            dec_fill_kernel<<<dim3((unsigned)n_images, 32), 256, 0, stream>>>(w->d_imgs, w->d_ndec);
            const unsigned exact_grid = (unsigned)((blocks + kIdctBlocks * kIdctIters - 1) / (kIdctBlocks * kIdctIters));
            dec_idct_kernel<<<exact_grid, kIdctBlocks * 8, 0, stream>>>(w->d_imgs, d_blk_first, n_images, blocks, w->d_coef,
                                                                        w->d_mul, w->d_ndec, w->d_list, w->d_flags + 2,
                                                                        w->list_cap);
            launches += 2;
            TICD_CUDA(h, cudaGetLastError());
        }
    } else if (int rc = launch_idct(h, w, d_blk_first, n_images, blocks, flags, stream, launches)) return rc;
    TICD_CUDA(h, cudaEventRecord(w->ev[5], stream));
    // [0] the last synchronisation launch still repaired something (speculative runs), [1] summary, [2] exact-pass blocks
    TICD_CUDA(h, cudaMemcpyAsync(w->h_flags + (speculative && subs ? 0 : 1), w->d_flags + (speculative && subs ? 0 : 1),
                                 (speculative && subs ? 3 : 2) * sizeof(int), cudaMemcpyDeviceToHost, stream));
    w->stats[0] += launches;
    w->stats[2] += rounds;
    w->pending_times = true;
    return TIC_OK;
}

int tic_decode_batch(tic_handle h, const void* const* d_streams, const int64_t* sizes, const int32_t* heights,
                     const int32_t* widths, int32_t n_images, uint32_t flags, void* const* d_pixels,
                     int32_t* d_status, void* stream_v) {
    if (!h) return TIC_E_INVALID;
    if (n_images < 0 || (n_images > 0 && (!d_streams || !sizes || !heights || !widths || !d_pixels))) {
        tic_internal_set_error(h, "tic_decode_batch: bad argument");
        return TIC_E_INVALID;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    TICD_CUDA(h, cudaSetDevice(tic_internal_device(h)));
    DecWs* w = get_ws(h);
    memset(w->stats, 0, sizeof w->stats);
    w->pending_times = false;
    w->run.active = false;
    if (int rc = ensure_base(h, w)) return rc;
    w->h_flags[0] = 0; w->h_flags[1] = 0; w->h_flags[2] = 0;
    TICD_CUDA(h, cudaMemsetAsync(w->d_flags, 0, 4 * sizeof(int), stream));
    if (n_images == 0) return TIC_OK;
    const size_t n = (size_t)n_images;
    if (n > w->imgs_cap) {
        size_t cap = n < 64 ? 64 : n * 2;
        cudaFreeHost(w->h_imgs); cudaFreeHost(w->h_first);
        w->h_imgs = nullptr; w->h_first = nullptr; w->imgs_cap = 0;
        if (int rc = grow(h, w->d_imgs, cap)) return rc;
        if (int rc = grow(h, w->d_first, 2 * (cap + 1))) return rc;
        if (int rc = grow(h, w->d_tabs, cap)) return rc;
        if (int rc = grow(h, w->d_mul, cap * 64)) return rc;
        if (int rc = grow(h, w->d_mulf, cap * kMulfStride)) return rc;
        if (int rc = grow(h, w->d_status_own, cap)) return rc;
        if (int rc = grow(h, w->d_ndec, cap)) return rc;
        TICD_CUDA(h, cudaMallocHost(&w->h_imgs, cap * sizeof(DecImage)));
        TICD_CUDA(h, cudaMallocHost(&w->h_first, 2 * (cap + 1) * sizeof(long long)));
        w->imgs_cap = cap;
    }
    long long* sub_first = w->h_first;
    long long* blk_first = w->h_first + (n + 1);
    long long subs = 0, blocks = 0;
    w->h_slices.clear();
    for (size_t i = 0; i < n; i++) {
        if (sizes[i] < 0 || heights[i] < 0 || widths[i] < 0 || (sizes[i] > 0 && !d_streams[i]) ||
            ((uintptr_t)d_streams[i] & 3u)) {
            tic_internal_set_error(h, "tic_decode_batch: stream " + std::to_string(i) +
                                          ": negative size / dimension, NULL or not 4-byte aligned");
            return TIC_E_INVALID;
        }
        DecImage& im = w->h_imgs[i];
        memset(&im, 0, sizeof im);
        im.words = static_cast<const uint32_t*>(d_streams[i]);
        im.nbits = sizes[i] * 8;
        im.height = heights[i];
        im.width = widths[i];
        im.bw = (widths[i] + 7) / 8;
        long long nblk = (heights[i] == 0 || widths[i] == 0) ? 0 : (long long)((heights[i] + 7) / 8) * im.bw;
        long long nsub = im.nbits > 128 ? (im.nbits - 128 + kSubBits - 1) / kSubBits : 0;
        if (nblk > 0x7fffffffLL || nsub > 0x7fffffffLL) {
            tic_internal_set_error(h, "tic_decode_batch: image too large");
            return TIC_E_INVALID;
        }
        if (nblk > 0 && !d_pixels[i]) {
            tic_internal_set_error(h, "tic_decode_batch: NULL pixel pointer");
            return TIC_E_INVALID;
        }
        im.nblk = (int)nblk;
        im.nsubs = (int)nsub;
        im.pixels = static_cast<uint8_t*>(d_pixels[i]);
        im.sub_first = subs;
        im.blk_first = blocks;
        sub_first[i] = subs;
        blk_first[i] = blocks;
        subs += nsub;
        blocks += nblk;
        push_slices(w->h_slices, (int)i, nsub);
    }
    sub_first[n] = subs;
    blk_first[n] = blocks;
    if ((size_t)subs > w->subs_cap) {
        size_t cap = (size_t)subs + (size_t)subs / 4 + 1024;
        w->subs_cap = 0;
        if (int rc = grow(h, w->d_E, cap + 1)) return rc;
        if (int rc = grow(h, w->d_U, cap + 1)) return rc;
        if (int rc = grow(h, w->d_ND, cap + 1)) return rc;
        if (int rc = grow(h, w->d_NB, cap + 1)) return rc;
        w->subs_cap = cap;
    }
    if ((size_t)blocks > w->blocks_cap) {
        size_t cap = (size_t)blocks + (size_t)blocks / 8 + 1024;
        w->blocks_cap = 0;
        if (int rc = grow(h, w->d_coef, cap * 64)) return rc;
        if (int rc = grow(h, w->d_list, 2 * cap)) return rc;
        w->list_cap = (int)std::min<size_t>(2 * cap, 0x7fffffffu);
        w->blocks_cap = cap;
    }
    int* status = d_status ? d_status : w->d_status_own;
    TICD_CUDA(h, cudaMemcpyAsync(w->d_imgs, w->h_imgs, n * sizeof(DecImage), cudaMemcpyHostToDevice, stream));
    TICD_CUDA(h, cudaMemcpyAsync(w->d_first, w->h_first, 2 * (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, stream));
    TICD_CUDA(h, cudaMemsetAsync(status, 0, n * sizeof(int), stream));
    if (subs) {
        TICD_CUDA(h, cudaMemsetAsync(w->d_E, 0, (size_t)subs * 4, stream));
        TICD_CUDA(h, cudaMemsetAsync(w->d_U, 0xff, (size_t)subs * 4, stream));
        TICD_CUDA(h, cudaMemsetAsync(w->d_ND, 0, (size_t)subs * sizeof(int2), stream));
    }
    w->run.active = true;
    w->run.speculative = (flags & TIC_DFLAG_SYNC_ROUNDS) == 0;
    w->run.n_images = n_images;
    w->run.subs = subs;
    w->run.blocks = blocks;
    w->run.flags = flags;
    w->run.status = status;
    w->stats[1] = subs;
    w->stats[3] = blocks;
    return run_batch(h, w, stream, w->run.speculative);
}

int tic_decode_coeffs(tic_handle h, const int32_t* d_dc, const int32_t* d_ac, int32_t height, int32_t width,
                      int32_t quality, int32_t scaled_dct, void* d_pixels, void* stream_v) {
    if (!h || height < 0 || width < 0 || quality < 0) return TIC_E_INVALID;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    TICD_CUDA(h, cudaSetDevice(tic_internal_device(h)));
    DecWs* w = get_ws(h);
    if (int rc = ensure_base(h, w)) return rc;
    memset(w->stats, 0, sizeof w->stats);
    w->pending_times = false;
    w->run.active = false;
    w->h_flags[0] = 0; w->h_flags[1] = 0; w->h_flags[2] = 0;
    TICD_CUDA(h, cudaMemsetAsync(w->d_flags, 0, 4 * sizeof(int), stream));
    long long nblk = (height == 0 || width == 0) ? 0 : (long long)((height + 7) / 8) * ((width + 7) / 8);
    if (nblk == 0) return TIC_OK;
    if (nblk > 0x7fffffffLL || !d_dc || !d_ac || !d_pixels) return TIC_E_INVALID;
    if (w->imgs_cap == 0) {   // the same per-image arrays tic_decode_batch uses, sized for one image
        const size_t cap = 64;
        if (int rc = grow(h, w->d_imgs, cap)) return rc;
        if (int rc = grow(h, w->d_first, 2 * (cap + 1))) return rc;
        if (int rc = grow(h, w->d_tabs, cap)) return rc;
        if (int rc = grow(h, w->d_mul, cap * 64)) return rc;
        if (int rc = grow(h, w->d_mulf, cap * kMulfStride)) return rc;
        if (int rc = grow(h, w->d_status_own, cap)) return rc;
        if (int rc = grow(h, w->d_ndec, cap)) return rc;
        TICD_CUDA(h, cudaMallocHost(&w->h_imgs, cap * sizeof(DecImage)));
        TICD_CUDA(h, cudaMallocHost(&w->h_first, 2 * (cap + 1) * sizeof(long long)));
        w->imgs_cap = cap;
    }
    if ((size_t)nblk > w->subs_cap) {
        size_t cap = (size_t)nblk + 1024;
        w->subs_cap = 0;
        if (int rc = grow(h, w->d_E, cap + 1)) return rc;
        if (int rc = grow(h, w->d_U, cap + 1)) return rc;
        if (int rc = grow(h, w->d_ND, cap + 1)) return rc;
        if (int rc = grow(h, w->d_NB, cap + 1)) return rc;
        w->subs_cap = cap;
    }
    if ((size_t)nblk > w->blocks_cap) {
        size_t cap = (size_t)nblk + 1024;
        w->blocks_cap = 0;
        if (int rc = grow(h, w->d_coef, cap * 64)) return rc;
        if (int rc = grow(h, w->d_list, 2 * cap)) return rc;
        w->list_cap = (int)std::min<size_t>(2 * cap, 0x7fffffffu);
        w->blocks_cap = cap;
    }
    DecImage& im = w->h_imgs[0];
    memset(&im, 0, sizeof im);
    im.pixels = static_cast<uint8_t*>(d_pixels);
    im.height = height; im.width = width; im.bw = (width + 7) / 8;
    im.nblk = (int)nblk; im.nsubs = (int)nblk;   // one scan element per block
    im.quality = (uint32_t)quality;
    im.mode = scaled_dct ? MODE_SCALED : MODE_FIXED;
    im.two_q = 1.0;
    w->h_first[0] = 0; w->h_first[1] = nblk;
    TICD_CUDA(h, cudaMemcpyAsync(w->d_imgs, w->h_imgs, sizeof(DecImage), cudaMemcpyHostToDevice, stream));
    TICD_CUDA(h, cudaMemcpyAsync(w->d_first, w->h_first, 2 * sizeof(long long), cudaMemcpyHostToDevice, stream));
    TICD_CUDA(h, cudaMemsetAsync(w->d_status_own, 0, sizeof(int), stream));
    TICD_CUDA(h, cudaMemsetAsync(w->d_ndec, 0, sizeof(int), stream));
    unsigned g1 = (unsigned)((nblk + 255) / 256), g2 = (unsigned)((nblk * 64 + 255) / 256);
    dec_coeffs_prep_kernel<<<g1, 256, 0, stream>>>(w->d_imgs, w->d_mul, w->d_mulf, d_dc, w->d_ND, (int)nblk, w->d_flags + 1);
    long long launches = 3;
    w->h_slices.clear();
    push_slices(w->h_slices, 0, nblk);
    if (int rc = launch_scan(h, w, w->d_status_own, stream, launches, nullptr, nullptr)) return rc;
    dec_coeffs_pack_kernel<<<g2, 256, 0, stream>>>(d_dc, d_ac, w->d_NB, w->d_coef, (int)nblk, w->d_flags + 1);
    if (int rc = launch_idct(h, w, w->d_first, 1, nblk, 0, stream, launches)) return rc;
    TICD_CUDA(h, cudaMemcpyAsync(w->h_flags + 1, w->d_flags + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    w->stats[0] = launches;
    w->stats[3] = nblk;
    return TIC_OK;
}

int tic_decode_finish(tic_handle h, void* stream_v) {
    if (!h) return TIC_E_INVALID;
    DecWs* w = get_ws(h);
    TICD_CUDA(h, cudaSetDevice(tic_internal_device(h)));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    TICD_CUDA(h, cudaStreamSynchronize(stream));
    if (w->run.active && w->run.speculative && w->run.subs && w->h_flags[0]) {
        // The last of the synchronisation launches enqueued blindly still repaired an entry state: what followed worked on
        // an unsettled parse.  Again from the state reached (E / U / ND are intact), this time looking at every round;
        // status words, summary and work-list count start over (dec_setup_kernel reports the header errors again).
        TICD_CUDA(h, cudaMemsetAsync(w->run.status, 0, (size_t)w->run.n_images * sizeof(int), stream));
        TICD_CUDA(h, cudaMemsetAsync(w->d_flags, 0, 4 * sizeof(int), stream));
        w->h_flags[0] = w->h_flags[1] = w->h_flags[2] = 0;
        if (int rc = run_batch(h, w, stream, false)) { w->run.active = false; return rc; }
        TICD_CUDA(h, cudaStreamSynchronize(stream));
        const long long needed = w->stats[2];   // launches in all, the last one without a repair
        if (needed > w->spec_rounds) w->spec_rounds = (int)(needed < 6 ? needed : 6);
    }
    w->run.active = false;
    if (w->pending_times) {
        float ms = 0.f;
        // [4] synchronisation rounds, [5] scan, [6] coefficient scatter (+ its zero fill), [7] IDCT: device ns
        if (cudaEventElapsedTime(&ms, w->ev[1], w->ev[2]) == cudaSuccess) w->stats[4] = (long long)(ms * 1e6);
        if (cudaEventElapsedTime(&ms, w->ev[2], w->ev[3]) == cudaSuccess) w->stats[5] = (long long)(ms * 1e6);
        if (cudaEventElapsedTime(&ms, w->ev[3], w->ev[4]) == cudaSuccess) w->stats[6] = (long long)(ms * 1e6);
        if (cudaEventElapsedTime(&ms, w->ev[4], w->ev[5]) == cudaSuccess) w->stats[7] = (long long)(ms * 1e6);
        w->pending_times = false;
    }
    if (w->h_flags && w->h_flags[1]) {
        tic_internal_set_error(h, "tic_decode: stream errors, OR of status bits = " + std::to_string(w->h_flags[1]));
        return TIC_E_STREAM;
    }
    return TIC_OK;
}

int tic_decode_stats(tic_handle h, int64_t stats[12]) {
    if (!h || !stats) return TIC_E_INVALID;
    DecWs* w = get_ws(h);
    w->stats[8] = w->h_flags ? w->h_flags[2] : 0;
    for (int i = 0; i < 12; i++) stats[i] = w->stats[i];
    return TIC_OK;
}

int tic_decompress_host(tic_handle h, const uint8_t* data, int64_t nbytes, uint32_t flags, uint8_t* out,
                        int64_t out_capacity, int32_t* status) {
    if (!h || !data || nbytes < 0) return TIC_E_INVALID;
    int32_t H, W, q;
    uint32_t flag;
    if (tic_parse_header(data, nbytes, &H, &W, &q, &flag)) {
        tic_internal_set_error(h, "tic_decompress_host: header needs 16 bytes");
        return TIC_E_INVALID;
    }
    long long npx = (long long)H * W;
    if (npx > out_capacity || (npx > 0 && !out)) return TIC_E_CAPACITY;
    TICD_CUDA(h, cudaSetDevice(tic_internal_device(h)));
    DecWs* w = get_ws(h);
    cudaStream_t s = tic_internal_own_stream(h);
    size_t padded = ((size_t)nbytes + 3) & ~(size_t)3;
    if (padded > w->stream_cap) {
        size_t cap = padded * 2 + 1024;
        w->stream_cap = 0;
        if (int rc = grow(h, w->d_stream, cap)) return rc;
        w->stream_cap = cap;
    }
    if ((size_t)npx > w->px_cap) {
        size_t cap = (size_t)npx * 2 + 1024;
        w->px_cap = 0;
        if (int rc = grow(h, w->d_px, cap)) return rc;
        w->px_cap = cap;
    }
    TICD_CUDA(h, cudaMemcpyAsync(w->d_stream, data, (size_t)nbytes, cudaMemcpyHostToDevice, s));
    const void* sp = w->d_stream;
    void* pp = w->d_px;
    int64_t sz = nbytes;
    int rc = tic_decode_batch(h, &sp, &sz, &H, &W, 1, flags, &pp, nullptr, s);
    if (rc) return rc;
    // finish first: it repeats the batch when its blindly enqueued synchronisation rounds did not settle, and only then
    // are status and pixels final
    rc = tic_decode_finish(h, s);
    if (rc != TIC_OK && rc != TIC_E_STREAM) return rc;
    int st = 0;
    TICD_CUDA(h, cudaMemcpyAsync(&st, w->d_status_own, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (npx) TICD_CUDA(h, cudaMemcpyAsync(out, w->d_px, (size_t)npx, cudaMemcpyDeviceToHost, s));
    TICD_CUDA(h, cudaStreamSynchronize(s));
    // A stream the device refuses as a whole (header mismatch, quality 0) writes no pixel: the reused workspace
    // still holds the previous decode.  The caller gets zeros, not somebody else's image.
    if (npx && (st & (TIC_DSTATUS_HEADER | TIC_DSTATUS_QUALITY))) memset(out, 0, (size_t)npx);
    if (status) *status = st;
    return rc;
}
