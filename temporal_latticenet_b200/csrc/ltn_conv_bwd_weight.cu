// Weight gradient of the lattice convolution on the tensor cores (training, `ltn_conv_bwd_weight` of SURVEY.md 8b):
//
//   dW[s*C + c, f] += sum_{v < Vq} act[nbr[v,s], c] * dy[v, f]          (absent neighbours contribute nothing; S = 1: nbr[v] = v)
//
// As a GEMM this contracts over the VERTICES: D[m = channel][n = filter] = sum_k A[m][k = vertex] * B[n][k = vertex].  Both
// operands therefore arrive "the wrong way round" for a K-major MMA -- a gathered activation row is [k][m contiguous], a
// dy row is [k][n contiguous] -- and that is exactly the MN-major shared-memory operand form of tcgen05 (instruction
// descriptor bits 15 / 16): a tile is stored as rows of 128 bytes (32 tf32 values of consecutive channels / filters of ONE
// vertex).  For 32-bit operands the only MN-major form is the 128-byte swizzle with a 32-BYTE base (layout type 1,
// "SWIZZLE_128B_BASE32B"): four rows per swizzle atom, the 32-byte chunk c of row r stored at chunk c ^ (r & 3); the
// descriptor's leading byte offset is the distance between two 32-channel column blocks (4 KB here), the stride byte offset
// the distance between two 4-vertex groups (512 B); one K = 8 MMA reads two such groups.
//
// fp32 parity as in the forward kernel's tf32 form: x = hi + lo (cvt.rna.tf32 twice), products hi*hi + lo*hi + hi*lo, fp32
// accumulation in tensor memory.  tf32 rather than fp16 operands: gradients span more than fp16's exponent range.
//
// Launch shape: one CTA per (slot s, 128-channel tile, vertex range); the vertex range (split-K) is chosen so that the
// grid is one wave; partial sums leave through fp32 red.global.add (dW is zeroed by the caller).  Per k-block of 32
// vertices all 8 warps gather / load the rows (next block's loads in flight in registers while the current one is
// converted), write hi and lo tiles into a 2-stage shared-memory ring, and one thread issues the 12 MMAs of the stage.
// The kernel is bound by the L2 -> SM traffic of its operands (dy is re-read per slot and channel tile), not by the
// tensor pipe.
#include "ltn_common.cuh"
#include "ltn_conv_common.cuh"

namespace {

constexpr int kWThreads = 256;
constexpr int kWM = 128;          // channels per CTA tile (MMA M)
constexpr int kWKB = 32;          // vertices per k-block = 4 tf32 MMAs of K = 8
constexpr int kWMaxF = 256;

struct BwdWParams {
    const float* act;   // [Vx, C]
    const float* dy;    // [Vq, F]
    const int* nbr;     // [Vq, S] or null
    float* dW;          // [S*C, F]
    int Vx, Vq, C, S, F;
    int m_tiles, k_splits, nkb;   // nkb = ceil(Vq / 32)
    int tmem_cols;
};

// MN-major shared-memory matrix descriptor for 32-bit operands: start address >> 4 [0,14), leading byte offset (between 128-byte
// column blocks along M / N) >> 4 [16,30), stride byte offset (between 4-row groups along K) >> 4 [32,46), version 1 [46,48),
// layout SWIZZLE_128B_BASE32B = 1 [61,64)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of the 16-byte chunk q16 (0..7) of row v inside a [32 rows][128 B] column block
__device__ __forceinline__ uint32_t sw32(int v, int q16) {
    return (uint32_t)v * 128u + (uint32_t)((((q16 >> 1) ^ (v & 3)) << 5) | ((q16 & 1) << 4));
}

__device__ __forceinline__ void split_tf32(const float4 v, float4& hi, float4& lo) {
    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
    lo.x = tf32_hi(v.x - hi.x); lo.y = tf32_hi(v.y - hi.y); lo.z = tf32_hi(v.z - hi.z); lo.w = tf32_hi(v.w - hi.w);
}

__global__ void __launch_bounds__(kWThreads, 1)
k_conv_bwd_weight(const __grid_constant__ BwdWParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[3];   // stage_free[2], all_done
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // task of this CTA
    const int split = blockIdx.x % p.k_splits;
    const int task = blockIdx.x / p.k_splits;
    const int mt = task % p.m_tiles, s = task / p.m_tiles;
    const int c0 = mt * kWM;
    const int per = (p.nkb + p.k_splits - 1) / p.k_splits;
    const int kb0 = split * per, kb1 = min(p.nkb, kb0 + per);
    if (kb0 >= kb1) return;   // uniform per CTA, before any barrier / tensor-memory allocation

    const int F = p.F, C = p.C;
    const int f4 = F >> 2;                          // float4 per dy row
    const int nblk = (F + 31) >> 5;                 // 32-filter column blocks of the B tile
    const uint32_t a_tile = 4u * 4096u;             // [4 column blocks][32 rows][128 B]
    const uint32_t b_tile = (uint32_t)nblk * 4096u;
    const uint32_t stage = 2u * a_tile + 2u * b_tile;   // A hi | A lo | B hi | B lo
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t bar_free = smem_u32(&bars[0]), bar_done = smem_u32(&bars[2]);

    if (tid == 0) {
        mbar_init(bar_free, 1);
        mbar_init(bar_free + 8, 1);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)p.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // thread -> (rows warp + 8j of the k-block, float4 column `lane` (+32) of a row)
    const bool a_col_ok = c0 + 4 * lane < C;                       // channels of this tile beyond C are zeros
    const bool b_col_ok[2] = {lane < f4, lane + 32 < f4};
    const int nb = f4 > 32 ? 2 : 1;
    float4 ra[4], rb[4][2];
    auto load_block = [&](int kb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = kb * kWKB + warp + 8 * j;
            ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            rb[j][0] = ra[j]; rb[j][1] = ra[j];
            if (v < p.Vq) {
                const int u = p.nbr ? __ldg(p.nbr + (size_t)v * p.S + s) : v;
                if (u >= 0 && u < p.Vx && a_col_ok) ra[j] = __ldg(reinterpret_cast<const float4*>(p.act + (size_t)u * C + c0) + lane);
                const float4* dyr = reinterpret_cast<const float4*>(p.dy + (size_t)v * F);
                if (b_col_ok[0]) rb[j][0] = __ldg(dyr + lane);
                if (nb > 1 && b_col_ok[1]) rb[j][1] = __ldg(dyr + lane + 32);
            }
        }
    };
    auto store_block = [&](int st) {
        uint8_t* sa_hi = smem + (size_t)st * stage;
        uint8_t* sa_lo = sa_hi + a_tile;
        uint8_t* sb_hi = sa_lo + a_tile;
        uint8_t* sb_lo = sb_hi + b_tile;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = warp + 8 * j;
            float4 hi, lo;
            split_tf32(ra[j], hi, lo);
            const uint32_t oa = (uint32_t)(lane >> 3) * 4096u + sw32(v, lane & 7);
            *reinterpret_cast<float4*>(sa_hi + oa) = hi;
            *reinterpret_cast<float4*>(sa_lo + oa) = lo;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h < nb && (lane + 32 * h) < 8 * nblk) {   // column blocks are written whole (zeros beyond F)
                    const int q = lane + 32 * h;
                    split_tf32(rb[j][h], hi, lo);
                    const uint32_t ob = (uint32_t)(q >> 3) * 4096u + sw32(v, q & 7);
                    *reinterpret_cast<float4*>(sb_hi + ob) = hi;
                    *reinterpret_cast<float4*>(sb_lo + ob) = lo;
                }
            }
        }
    };

    // instruction descriptor: D = F32 [4,6), A / B = TF32 (2) [7,10) [10,13), A and B MN-major (bits 15, 16), N >> 3 [17,23), M >> 4 [24,29)
    const int Nmma = (F + 15) & ~15;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(Nmma >> 3) << 17) | ((uint32_t)(kWM >> 4) << 24);

    load_block(kb0);
    for (int i = 0; kb0 + i < kb1; ++i) {
        const int st = i & 1;
        if (i >= 2) {   // the MMAs that read this stage two blocks ago have completed
            mbar_wait(bar_free + 8 * st, (uint32_t)((i >> 1) - 1) & 1u);
            tc_fence_after();
        }
        store_block(st);
        if (kb0 + i + 1 < kb1) load_block(kb0 + i + 1);   // in flight while this stage's MMAs run
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t sa_hi = smem_u32(smem + (size_t)st * stage), sa_lo = sa_hi + a_tile, sb_hi = sa_lo + a_tile, sb_lo = sb_hi + b_tile;
#pragma unroll
            for (int k = 0; k < kWKB / 8; ++k) {
                const uint64_t a_hi = make_desc_mn(sa_hi + k * 1024u, 4096u, 512u), a_lo = make_desc_mn(sa_lo + k * 1024u, 4096u, 512u);
                const uint64_t b_hi = make_desc_mn(sb_hi + k * 1024u, 4096u, 512u), b_lo = make_desc_mn(sb_lo + k * 1024u, 4096u, 512u);
                umma_tf32(tmem_base, a_hi, b_hi, idesc, (i | k) != 0);
                umma_tf32(tmem_base, a_lo, b_hi, idesc, 1);
                umma_tf32(tmem_base, a_hi, b_lo, idesc, 1);
            }
            umma_commit(bar_free + 8 * st);
            if (kb0 + i + 1 >= kb1) umma_commit(bar_done);
        }
    }
    mbar_wait(bar_done, 0);
    tc_fence_after();

    // epilogue: warp w reads lanes 32 (w % 4) .. + 31 (= channels), column chunks alternate between the two warpgroups
    const int row = c0 + 32 * (warp & 3) + lane;
    float* dst = p.dW + ((size_t)s * C + row) * F;
    for (int cb = 32 * (warp >> 2); cb < F; cb += 64) {
        float acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)cb, acc);
        if (row < C) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                if (cb + j < F)   // F % 4 == 0
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + cb + j), "f"(acc[j]), "f"(acc[j + 1]), "f"(acc[j + 2]),
                                 "f"(acc[j + 3]) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
    }
}

}  // namespace

extern "C" {

// dW [S*C, F] += gathered-act^T . dy  (see the top of this file); dW is ZEROED by the caller.  act [Vx, C] = the layer's
// input activations (after GroupNorm / ReLU), dy [Vq, F] the gradient of its output, nbr [Vq, S] the layer's neighbour
// table (null: S = 1, row v of act).  C % 4 == 0, F % 16 == 0, 16 <= F <= 256.
int ltn_conv_bwd_weight(const float* act, int Vx, const int* nbr, int Vq, int C, int S, const float* dy, int F, float* dW,
                        void* stream) {
    if (Vq <= 0 || Vx <= 0) return 0;
    if (C <= 0 || C % 4 || F < 16 || F > kWMaxF || F % 16 || S < 1) return -2;
    BwdWParams p;
    p.act = act; p.dy = dy; p.nbr = nbr; p.dW = dW;
    p.Vx = Vx; p.Vq = Vq; p.C = C; p.S = nbr ? S : 1; p.F = F;
    p.m_tiles = (C + kWM - 1) / kWM;
    p.nkb = (Vq + kWKB - 1) / kWKB;
    const int tasks = p.S * p.m_tiles;
    int ks = 148 / tasks;
    if (ks < 1) ks = 1;
    if (ks > p.nkb) ks = p.nkb;
    p.k_splits = ks;
    int cols = 32;
    while (cols < ((F + 15) & ~15)) cols <<= 1;
    p.tmem_cols = cols;
    const int nblk = (F + 31) / 32;
    const size_t stage = 2 * (size_t)4 * 4096 + 2 * (size_t)nblk * 4096;
    const size_t smem = 2 * stage + 1024;
    cudaError_t e = cudaFuncSetAttribute(k_conv_bwd_weight, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_conv_bwd_weight<<<tasks * ks, kWThreads, smem, (cudaStream_t)stream>>>(p);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
