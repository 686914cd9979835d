// tic_tables.h — constants of the .img format, shared by host and device code.
//
// The numbers restate tinyimgcodec/constants.py of the reference:
//   LUMINANCE_QUANTIZATION_TABLE  constants.py:9-20
//   ZIGZAG_ORDER                  constants.py:23-34
//   HUFFMAN_CATEGORY_CODEWORD     constants.py:53-242 (JPEG Annex K.3 / K.5 luminance tables,
//                                 kept here as the canonical BITS/HUFFVAL lists)
#pragma once
#include <stdint.h>

namespace tic {

static const int kQuantBase[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};

#ifdef __CUDACC__
__host__ __device__
#endif
constexpr int kQuantBaseDev(int r) {   // same table, usable in device code with a compile-time index
    constexpr int q[64] = {
        16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
        14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
        18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
        49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    return q[r];
}

// zigzag index k -> raster index u*8+v (u vertical frequency)
#define TIC_ZIGZAG_LIST                                                                     \
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34,  \
    27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37,    \
    44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63

static const uint8_t kDcBits[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcBits[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61,
    0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52,
    0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25,
    0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45,
    0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64,
    0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
    0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8,
    0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

// Huffman table as the device sees it: {code (right-aligned), kHuffPresent | length} per symbol.
// DC: symbol = size category (0..15).  AC: symbol = run*16 + size.  len == 0: not in the table
// (a present symbol can have a zero-length code: a one-symbol alphabet, huffman.py:175-180).
constexpr uint32_t kHuffPresent = 0x100;
constexpr uint32_t kHuffLenMask = 0xff;
struct HuffEntry {
    uint32_t code;
    uint32_t len;
};
struct HuffTables {
    HuffEntry ac[256];
    HuffEntry dc[16];
};

// Per-image tables of the auto_generate_huffman_table mode (codec.py:146-148), built on the device:
// the (code,len) entries plus the serialised header (codec.py:102-112, 73-84) as MSB-first words.
constexpr int kMaxHdrWords = 416;   // 160 + 16*(8+32) + 16 + 256*(16+32) bits
struct AutoTables {
    HuffTables tab;
    uint32_t hdr_bits;
    uint32_t status;
    uint32_t hdr_words[kMaxHdrWords];
};

}  // namespace tic
