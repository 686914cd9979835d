// tc_probe.cu — sm_100a probe for the tensor-core FDCT front end.
//
// Question: can the 8x8 FDCT + 1/qt scaling of 128 blocks run as ONE tcgen05 GEMM
//     D[128 blocks][64 zigzag coefficients] = A[128][64 pixels as f16] x (Bhi + Blo)[64][64]
// (K = 128: the f16 hi and lo halves of the coefficient matrix stacked along K, the A descriptor reused),
// accurately enough for the tie guard, and how fast are the MMA and the TMEM read-back?
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_probe tools/tc_probe.cu
//   run  : ./tc_probe            (prints one JSON object per experiment)
//
// Not part of the product; nothing here is linked into libtinyimgcodec_cuda.so.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"cuda_error\": \"%s\", \"at\": \"%s\"}\n", cudaGetErrorString(e_), #x); exit(2); } } while (0)

static const int kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
static const int kQuantBase[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57,
                                   69, 56, 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55,
                                   64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// bounded: a wrong descriptor must not hang the box
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    for (int i = 0; i < (1 << 22); i++)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start [0,14), LBO [16,30), SBO [32,46),
// version = 1 at [46,48), layout type 0 at [61,64); all byte quantities >> 4)
__host__ __device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = F16 (0), K-major both,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// ---------------------------------------------------------------------------------------------
// layouts: A[m][k] (m = block in tile, k = 8 * y + x), 16 KB: (m / 8) * 128 + y * 2048 + (m % 8) * 16 + x * 2
//          B[n][k] (n = zigzag index, k = 0..127: 64 hi then 64 lo), 16 KB: (n / 8) * 128 + (k / 8) * 1024 + (n % 8) * 16 + (k % 8) * 2
// ---------------------------------------------------------------------------------------------
constexpr int kABytes = 128 * 64 * 2, kBBytes = 64 * 128 * 2;
struct Params {
    uint32_t lbo_a, sbo_a, lbo_b, sbo_b;   // descriptor fields (bytes)
    int pixmode;                            // 0: f16(p - 128) via PRMT + HADD2; 1: f16 subnormal p * 2^-24; 2: f16(1024 + p)
    int lo_first;                           // accumulate the lo halves before the hi halves
    int width;                              // image width (tile = 128 consecutive blocks of one block row)
    int ntiles;
    int read_cols;                          // timing: TMEM columns read back per tile (0..64)
};

template <bool kStore>
__global__ void __launch_bounds__(128) fdct_tc_kernel(const uint8_t* __restrict__ px, const uint4* __restrict__ bmat,
                                                     float* __restrict__ out, unsigned int* __restrict__ result,
                                                     const Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + kABytes;
    __shared__ __align__(8) unsigned long long bar_storage;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t bar = smem_u32(&bar_storage);
    for (int i = t; i < kBBytes / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = bmat[i];
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;
    unsigned int live = 0, failed = 0;
    const int tiles_per_row = p.width / 1024;   // a tile is 1024 pixels wide
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int brow = tile / tiles_per_row, bcol0 = (tile - brow * tiles_per_row) * 128;
        const uint8_t* src = px + (size_t)brow * 8 * p.width + (size_t)(bcol0 + t) * 8;
        uint2 rows[8];
#pragma unroll
        for (int y = 0; y < 8; y++) rows[y] = __ldg(reinterpret_cast<const uint2*>(src + (size_t)y * p.width));
        unsigned char* dst = sA + (t >> 3) * 128 + (t & 7) * 16;
#pragma unroll
        for (int y = 0; y < 8; y++) {
            uint4 v;
            if (p.pixmode == 1) {   // subnormal halves: p * 2^-24
                v.x = __byte_perm(rows[y].x, 0u, 0x4140); v.y = __byte_perm(rows[y].x, 0u, 0x4342);
                v.z = __byte_perm(rows[y].y, 0u, 0x4140); v.w = __byte_perm(rows[y].y, 0u, 0x4342);
            } else {                // 0x64pp = 1024 + p
                v.x = __byte_perm(rows[y].x, 0x64646464u, 0x4140); v.y = __byte_perm(rows[y].x, 0x64646464u, 0x4342);
                v.z = __byte_perm(rows[y].y, 0x64646464u, 0x4140); v.w = __byte_perm(rows[y].y, 0x64646464u, 0x4342);
                if (p.pixmode == 0) {
                    const __half2 c = __floats2half2_rn(-1152.0f, -1152.0f);
                    __half2 h;
                    h = __hadd2(*reinterpret_cast<__half2*>(&v.x), c); v.x = *reinterpret_cast<uint32_t*>(&h);
                    h = __hadd2(*reinterpret_cast<__half2*>(&v.y), c); v.y = *reinterpret_cast<uint32_t*>(&h);
                    h = __hadd2(*reinterpret_cast<__half2*>(&v.z), c); v.z = *reinterpret_cast<uint32_t*>(&h);
                    h = __hadd2(*reinterpret_cast<__half2*>(&v.w), c); v.w = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            *reinterpret_cast<uint4*>(dst + y * 2048) = v;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int half = p.lo_first ? 1 - (i >> 2) : (i >> 2);   // 0: hi, 1: lo
                const int j = i & 3;
                // chunk j of A: pixel rows 2j, 2j+1 (2 core matrices along K, 2048 bytes apart in this layout)
                const uint64_t da = make_desc(a0 + (uint32_t)j * 2u * 2048u, p.lbo_a, p.sbo_a);
                const uint64_t db = make_desc(b0 + (uint32_t)(half * 4 + j) * 2u * 1024u, p.lbo_b, p.sbo_b);
                tc_mma_f16(tmem, da, db, kIdesc, i > 0 ? 1u : 0u);
            }
            tc_commit(bar);
        }
        if (!mbar_wait_bounded(bar, phase)) failed = 1;
        phase ^= 1;
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        if (kStore) {
            float* o = out + ((size_t)tile * 128 + t) * 64;
#pragma unroll
            for (int g = 0; g < 8; g++) {
                uint32_t r[8];
                tmem_ld8(taddr + g * 8, r);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 8; i++) o[g * 8 + i] = __uint_as_float(r[i]);
            }
        } else {
            for (int c = 0; c < p.read_cols; c += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + c, r);
                tmem_wait_ld();
                float m0 = 0.f, m1 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) { m0 = fmaxf(m0, fabsf(__uint_as_float(r[i]))); m1 = fmaxf(m1, fabsf(__uint_as_float(r[8 + i]))); }
                live += __any_sync(0xffffffffu, m0 >= 1.0f) ? 1u : 0u;
                live += __any_sync(0xffffffffu, m1 >= 1.0f) ? 1u : 0u;
            }
        }
    }
    if (failed) atomicOr(&result[0], 1u);
    if (!kStore && (t & 31) == 0) atomicAdd(&result[1], live);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// TMEM read-back rate: every warp reads its 32 lanes x 64 columns `reps` times
__global__ void __launch_bounds__(1024) tmem_read_kernel(unsigned int* __restrict__ sink, int reps, int ncols, long long* __restrict__ cycles) {
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64 % 512);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        for (int c = 0; c < ncols; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; i++) acc ^= v[i];
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (t == 0) cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static double basis(int n, int k) {   // orthonormal 2-D DCT-II, zigzag column n, pixel k = 8y + x
    const int r = kZigzag[n], u = r >> 3, v = r & 7, y = k >> 3, x = k & 7;
    const double cu = u ? 0.5 : sqrt(0.125), cv = v ? 0.5 : sqrt(0.125);
    return cu * cv * cos((2 * y + 1) * u * M_PI / 16.0) * cos((2 * x + 1) * v * M_PI / 16.0);
}

struct Quant {
    double qt[64];      // zigzag order
    double colscale[64];// B column = basis * colscale; t = D / colscale / qt... (see build_b)
    float cn[64];       // t = D * cn
    int E;
};

static void build_b(int quality, int pixmode, std::vector<__half>& blob, Quant& q, double* abs_err_bound) {
    for (int n = 0; n < 64; n++) {
        const int i = kZigzag[n];
        if (quality < 50) q.qt[n] = ((double)kQuantBase[i] * (5000.0 / quality)) / 100.0;
        else q.qt[n] = (double)((long)kQuantBase[i] * (200 - 2 * quality)) / 100.0;
    }
    // one power of two for the whole matrix so that the largest entry is in [2^13, 2^14)
    double mx = 0;
    for (int n = 0; n < 64; n++)
        for (int k = 0; k < 64; k++) mx = fmax(mx, fabs(basis(n, k)) / q.qt[n]);
    int e;
    frexp(mx, &e);   // mx = f * 2^e, f in [0.5, 1)
    q.E = 14 - e;
    blob.assign(64 * 128, __float2half(0.f));
    for (int n = 0; n < 64; n++) {
        double sum_abs_err = 0;
        for (int k = 0; k < 64; k++) {
            const double b = ldexp(basis(n, k) / q.qt[n], q.E);
            const __half hi = __double2half(b);
            const __half lo = __double2half(b - (double)__half2float(hi));
            const double err = b - (double)__half2float(hi) - (double)__half2float(lo);
            sum_abs_err += fabs(err);
            const size_t ih = (size_t)(n / 8) * 64 + (size_t)(k / 8) * 512 + (size_t)(n % 8) * 8 + (size_t)(k % 8);
            const int kl = k + 64;
            const size_t il = (size_t)(n / 8) * 64 + (size_t)(kl / 8) * 512 + (size_t)(n % 8) * 8 + (size_t)(kl % 8);
            blob[ih] = hi;
            blob[il] = lo;
        }
        q.cn[n] = (float)ldexp(1.0, -q.E + (pixmode == 1 ? 24 : 0));
        abs_err_bound[n] = ldexp(sum_abs_err * 255.0, -q.E);   // in t units: matrix rounding only
    }
}

static void gen_image(std::vector<uint8_t>& img, int W, int H, int kind, unsigned seed) {
    img.resize((size_t)W * H);
    srand(seed);
    if (kind == 0) {   // smooth + noise (the benchmark's kind of content)
        const int gw = W / 32 + 2, gh = H / 32 + 2;
        std::vector<int> g((size_t)gw * gh);
        for (auto& v : g) v = rand() % 256;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const int gx = x / 32, gy = y / 32, fx = x % 32, fy = y % 32;
                const int a = g[gy * gw + gx], b = g[gy * gw + gx + 1], c = g[(gy + 1) * gw + gx], d = g[(gy + 1) * gw + gx + 1];
                int v = ((a * (32 - fx) + b * fx) * (32 - fy) + (c * (32 - fx) + d * fx) * fy) / 1024 + (rand() % 21) - 10;
                img[(size_t)y * W + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
    } else if (kind == 1) {
        for (auto& v : img) v = (uint8_t)(rand() % 256);
    } else if (kind == 2) {
        for (auto& v : img) v = (rand() & 1) ? 255 : 0;
    } else if (kind == 3) {   // sign-matched to basis function (block index mod 64): maximises sum |p||b|
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const int n = ((y / 8) * (W / 8) + x / 8) % 64;
                img[(size_t)y * W + x] = basis(n, (y % 8) * 8 + (x % 8)) > 0 ? 255 : 0;
            }
    } else {   // flat blocks of every level
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) img[(size_t)y * W + x] = (uint8_t)(((y / 8) * (W / 8) + x / 8) % 256);
    }
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("{\"device\": \"%s\", \"sm\": %d, \"cc\": \"%d.%d\"}\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
    const int smem_bytes = kABytes + kBBytes + 1024;
    CK(cudaFuncSetAttribute(fdct_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CK(cudaFuncSetAttribute(fdct_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    unsigned int* d_result;
    CK(cudaMalloc(&d_result, 16));

    // ---------------- accuracy: descriptor variants x pixel modes x qualities x image kinds ----------------
    const int W = 1024, H = 512;   // 64 block rows = 64 tiles, 8192 blocks
    const int ntiles = (H / 8) * (W / 1024);
    uint8_t* d_px;
    uint4* d_b;
    float* d_out;
    CK(cudaMalloc(&d_px, (size_t)W * H));
    CK(cudaMalloc(&d_b, kBBytes));
    CK(cudaMalloc(&d_out, (size_t)ntiles * 128 * 64 * 4));
    std::vector<float> h_out((size_t)ntiles * 128 * 64);
    struct DescVar { const char* name; uint32_t lbo_a, sbo_a, lbo_b, sbo_b; };
    const DescVar dvars[2] = {{"lbo=K-step,sbo=M-step", 2048, 128, 1024, 128}, {"swapped", 128, 2048, 128, 1024}};
    int good_var = -1;
    for (int dv = 0; dv < 2 && good_var < 0; dv++) {
        for (int pixmode = 0; pixmode < 3; pixmode++) {
            for (int qi = 0; qi < 3; qi++) {
                const int quality = qi == 0 ? 50 : (qi == 1 ? 90 : 10);
                std::vector<__half> blob;
                Quant q;
                double bound[64];
                build_b(quality, pixmode, blob, q, bound);
                CK(cudaMemcpy(d_b, blob.data(), kBBytes, cudaMemcpyHostToDevice));
                for (int lo_first = 0; lo_first < 2; lo_first++) {
                    for (int kind = 0; kind < 5; kind++) {
                        if (dv == 1 && (kind || qi || pixmode || lo_first)) continue;
                        std::vector<uint8_t> img;
                        gen_image(img, W, H, kind, 1234 + kind);
                        CK(cudaMemcpy(d_px, img.data(), img.size(), cudaMemcpyHostToDevice));
                        CK(cudaMemset(d_result, 0, 16));
                        CK(cudaMemset(d_out, 0xff, (size_t)ntiles * 128 * 64 * 4));
                        Params p{dvars[dv].lbo_a, dvars[dv].sbo_a, dvars[dv].lbo_b, dvars[dv].sbo_b, pixmode, lo_first, W, ntiles, 64};
                        fdct_tc_kernel<true><<<ntiles, 128, smem_bytes>>>(d_px, d_b, d_out, d_result, p);
                        CK(cudaGetLastError());
                        CK(cudaDeviceSynchronize());
                        unsigned int res[4];
                        CK(cudaMemcpy(res, d_result, 16, cudaMemcpyDeviceToHost));
                        CK(cudaMemcpy(h_out.data(), d_out, h_out.size() * 4, cudaMemcpyDeviceToHost));
                        // compare with the float64 transform
                        double max_err_t = 0, max_err_c = 0, sum_err_c = 0, max_rel = 0;
                        long cnt = 0;
                        int worst_n = -1;
                        for (int tile = 0; tile < ntiles; tile++)
                            for (int m = 0; m < 128; m++) {
                                double pxv[64];
                                for (int k = 0; k < 64; k++)
                                    pxv[k] = (double)img[(size_t)(tile * 8 + (k >> 3)) * W + m * 8 + (k & 7)] - 128.0;
                                const float* o = &h_out[((size_t)tile * 128 + m) * 64];
                                for (int n = 0; n < 64; n++) {
                                    double c = 0, mag = 0;
                                    for (int k = 0; k < 64; k++) { c += pxv[k] * basis(n, k); mag += fabs(pxv[k] * basis(n, k)); }
                                    double got = (double)o[n] * (double)q.cn[n];
                                    if (pixmode != 0 && n == 0) got -= (pixmode == 1 ? 128.0 : 1152.0) * 8.0 / q.qt[0];   // un-centred pixels move the DC only
                                    if (pixmode != 0) mag += (pixmode == 1 ? 128.0 : 1152.0) * 6.5;
                                    const double et = fabs(got - c / q.qt[n]);
                                    const double ec = et * q.qt[n];
                                    if (ec > max_err_c) { max_err_c = ec; worst_n = n; }
                                    if (et > max_err_t) max_err_t = et;
                                    if (mag > 0 && ec / mag > max_rel) max_rel = ec / mag;
                                    sum_err_c += ec;
                                    cnt++;
                                }
                            }
                        printf("{\"exp\": \"accuracy\", \"desc\": \"%s\", \"pixmode\": %d, \"quality\": %d, \"lo_first\": %d, \"kind\": %d, "
                               "\"wait_failed\": %u, \"max_err_coef\": %.3e, \"mean_err_coef\": %.3e, \"max_err_t\": %.3e, "
                               "\"max_err_over_sum_abs_terms\": %.3e, \"worst_n\": %d, \"E\": %d}\n",
                               dvars[dv].name, pixmode, quality, lo_first, kind, res[0], max_err_c, sum_err_c / cnt, max_err_t,
                               max_rel, worst_n, q.E);
                        fflush(stdout);
                        if (dv == 0 && kind == 0 && qi == 0 && pixmode == 0 && lo_first == 0 && max_err_c < 0.05) good_var = 0;
                        if (dv == 1 && max_err_c < 0.05) good_var = 1;
                    }
                }
            }
        }
        if (dv == 0 && good_var == 0) break;
    }
    printf("{\"exp\": \"descriptor\", \"good_variant\": %d}\n", good_var);
    if (good_var < 0) return 1;

    // ---------------- throughput: load + convert + MMA (+ TMEM read-back) over 1 GiB of pixels ----------------
    {
        const int TW = 1024, TH = 1024, NI = 1024;
        uint8_t* d_big;
        CK(cudaMalloc(&d_big, (size_t)TW * TH * NI));
        std::vector<uint8_t> img;
        gen_image(img, TW, TH, 0, 77);
        for (int i = 0; i < NI; i++) CK(cudaMemcpy(d_big + (size_t)i * TW * TH, img.data(), img.size(), cudaMemcpyHostToDevice));
        std::vector<__half> blob;
        Quant q;
        double bound[64];
        build_b(50, 0, blob, q, bound);
        CK(cudaMemcpy(d_b, blob.data(), kBBytes, cudaMemcpyHostToDevice));
        const int nt = NI * (TH / 8);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int pixmode = 0; pixmode < 2; pixmode++)
            for (int ctas = 2; ctas <= 6; ctas += 2)
                for (int rc = 0; rc <= 64; rc += 32) {
                    Params p{dvars[good_var].lbo_a, dvars[good_var].sbo_a, dvars[good_var].lbo_b, dvars[good_var].sbo_b, pixmode, 1, TW, nt, rc};
                    float best = 1e9f;
                    for (int rep = 0; rep < 4; rep++) {
                        CK(cudaMemset(d_result, 0, 16));
                        CK(cudaEventRecord(e0));
                        fdct_tc_kernel<false><<<prop.multiProcessorCount * ctas, 128, smem_bytes>>>(d_big, d_b, nullptr, d_result, p);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        CK(cudaGetLastError());
                        float ms;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (rep && ms < best) best = ms;
                    }
                    unsigned int res[4];
                    CK(cudaMemcpy(res, d_result, 16, cudaMemcpyDeviceToHost));
                    printf("{\"exp\": \"throughput\", \"pixmode\": %d, \"ctas_per_sm\": %d, \"read_cols\": %d, \"ms_per_GiB\": %.4f, "
                           "\"GBps\": %.1f, \"ms_for_4096_images\": %.3f, \"wait_failed\": %u, \"live_groups\": %u}\n",
                           pixmode, ctas, rc, best, (double)TW * TH * NI / best / 1e6, best * 4.0, res[0], res[1]);
                    fflush(stdout);
                }
        CK(cudaFree(d_big));
    }
    // ---------------- TMEM read-back rate ----------------
    {
        long long* d_cyc;
        unsigned int* d_sink;
        CK(cudaMalloc(&d_cyc, 8 * 1024));
        CK(cudaMalloc(&d_sink, 16));
        for (int warps = 4; warps <= 32; warps *= 2) {
            const int reps = 2000;
            tmem_read_kernel<<<prop.multiProcessorCount, warps * 32>>>(d_sink, reps, 64, d_cyc);
            CK(cudaDeviceSynchronize());
            long long cyc[8];
            CK(cudaMemcpy(cyc, d_cyc, 64, cudaMemcpyDeviceToHost));
            const double bytes = (double)warps * 32 * 64 * 4 * reps;
            printf("{\"exp\": \"tmem_read\", \"warps_per_sm\": %d, \"cycles\": %lld, \"bytes_per_clk_per_sm\": %.1f}\n", warps, cyc[0],
                   bytes / (double)cyc[0]);
        }
    }
    return 0;
}
