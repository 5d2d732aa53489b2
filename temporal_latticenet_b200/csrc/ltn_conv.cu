// Fused lattice convolution on the Blackwell tensor cores: neighbour gather (+ folded GroupNorm/ReLU)
// + GEMM + bias + residual in ONE kernel, no [V,9C] im2row buffer.
//
//   out[v, f] = sum_{s<S} sum_{c<C} act( x[nbr[v,s], c] ) * W[s*C + c, f]  (+ bias[f]) (+ res[v, f])
//   act(t)    = relu?( t * a_scale[c] + a_shift[c] )     -- GroupNorm folded to a per-channel affine,
//                                                           applied to PRESENT rows only (absent -> 0,
//                                                           exactly what im2row of the normalised values gives)
//
// serves ConvLatticeModule / CoarsenLattice / FinefyLattice (S = 9, seq_lattice/lattice_modules.py:440,573;
// models.py:353,398), GnRelu1x1 / Conv1x1 / nn.Linear and the GRU/LSTM gate GEMMs (S = 1, nbr = NULL).
//
// Blackwell mapping
//  * tcgen05.mma.cta_group::1.kind::tf32, M = 128 vertices per CTA, N = up to 256 output channels so the
//    A rows are gathered exactly once; fp32 accumulator lives in TMEM (N columns x 128 lanes).
//  * fp32 parity: every fp32 operand is split x = hi + lo (hi = round-to-tf32, lo = x - hi exactly) and
//    the product is issued as hi*hi + lo*hi + hi*lo -- three tensor-core passes, ~2^-21 relative error per
//    product, i.e. fp32-class results (the reference ran these GEMMs in fp32 / TF32 cuBLAS).
//    PASSES = 1 is the single-pass TF32 variant (reported separately, never the default).
//  * operands are staged K-major in 128-byte-swizzled shared memory (the UMMA canonical layout) by
//    4 producer warps: one thread per tile row gathers 128 contiguous bytes of one vertex row (a whole
//    cache line, L2-resident), applies the folded GroupNorm+ReLU, splits hi/lo and writes 16-byte chunks
//    at chunk ^ (row & 7).  A multi-stage mbarrier ring decouples them from the single MMA-issuing
//    thread; tcgen05.commit releases stages and finally hands the accumulator to the epilogue.
//  * epilogue: tcgen05.ld 32x32b (thread = row, 32 columns at a time) -> + bias, + residual -> 128-byte
//    row-segment stores.
//  * sizes that only the device knows (vertex counts after hash insertion) are read from device
//    memory (vq_dev / vx_dev), so the launch needs no host synchronisation.
#include "ltn_common.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;        // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;          // tf32: 32 bytes per MMA
constexpr int kProducerThreads = 128;
constexpr int kThreads = 160;      // 4 producer/epilogue warps + 1 MMA warp
constexpr int kMaxStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major, 1) [16,30), stride byte offset = 1024 B
// between 8-row groups [32,46), descriptor version 1 [46,48), layout type SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct ConvParams {
    const float* x;        // [Vx, C]
    const int* nbr;        // [Vq, S] or null (identity, S = 1)
    const float* wt;       // [F, S*C]  K-major weights
    const float* a_scale;  // [C] or null
    const float* a_shift;  // [C] or null
    const float* bias;     // [F] or null
    const float* res;      // [Vq, F] or null
    float* out;            // [Vq, ldo]
    const int* vq_dev;     // device-side row counts (nullable)
    const int* vx_dev;
    int Vq, Vx, C, S, F, ldo, relu;
    int n_tile;            // output channels per CTA (<= 256, multiple of 16)
    int stages;
};

// store 8 fp32 of one operand row as hi (and lo) parts at swizzled 16-byte chunks 2j, 2j+1
template <int PASSES>
__device__ __forceinline__ void stage_chunk(uint8_t* hi_row, uint8_t* lo_row, int row, int chunk, float4 v) {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    int off = ((chunk ^ (row & 7)) << 4);
    *reinterpret_cast<float4*>(hi_row + off) = h;
    if (PASSES == 3) {
        float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        *reinterpret_cast<float4*>(lo_row + off) = l;
    }
}

template <int PASSES>
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_affine[2 * 256];

    const int Vq = p.vq_dev ? min(p.Vq, __ldg(p.vq_dev)) : p.Vq;
    const int Vx = p.vx_dev ? min(p.Vx, __ldg(p.vx_dev)) : p.Vx;
    const int row0 = blockIdx.x * kBlockM;
    if (row0 >= Vq) return;  // uniform per CTA, before any barrier / TMEM allocation
    const int n0 = blockIdx.y * p.n_tile;
    const int N = min(p.n_tile, p.F - n0);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.S * p.C;
    const int num_kb = K / kBlockK;
    const int kb_per_slot = p.C / kBlockK;
    const int stages = p.stages;

    // shared-memory carve-up: per stage [A_hi | A_lo | B_hi | B_lo], every block 1024-byte aligned
    const uint32_t a_bytes = kBlockM * 128;
    const uint32_t b_bytes = (uint32_t)p.n_tile * 128;
    const uint32_t stage_bytes = (PASSES == 3 ? 2 : 1) * (a_bytes + b_bytes);
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kMaxStages]), bar_acc = smem_u32(&bars[2 * kMaxStages]);
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)N) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_full + 8 * s, kProducerThreads);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (p.a_scale)
        for (int c = tid; c < p.C; c += kThreads) {
            s_affine[c] = __ldg(p.a_scale + c);
            s_affine[256 + c] = __ldg(p.a_shift + c);
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp < 4) {
        // ===================== producers: gather A rows, stream B rows =====================
        const int r = tid;                 // tile row
        const int v = row0 + r;            // query vertex
        const bool row_ok = v < Vq;
        const bool affine = p.a_scale != nullptr;
        int stage = 0;
        uint32_t phase = 0;
        int src = -1, cur_slot = -1;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int slot = kb / kb_per_slot;
            const int c0 = (kb - slot * kb_per_slot) * kBlockK;
            if (slot != cur_slot) {
                cur_slot = slot;
                src = -1;
                if (row_ok) {
                    src = p.nbr ? __ldg(p.nbr + (size_t)v * p.S + slot) : v;
                    if (src >= Vx) src = -1;
                }
            }
            // issue the global loads before waiting for the stage to drain
            float4 a[8];
            if (src >= 0) {
                const float4* g = reinterpret_cast<const float4*>(p.x + (size_t)src * p.C + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = __ldg(g + j);
            }
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            uint8_t* st = smem + (size_t)stage * stage_bytes;
            uint8_t* a_hi = st + r * 128;
            uint8_t* a_lo = a_hi + a_bytes;
            if (src >= 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 t = a[j];
                    if (affine) {
                        const float* sc = s_affine + c0 + 4 * j;
                        const float* sh = sc + 256;
                        t.x = fmaf(t.x, sc[0], sh[0]); t.y = fmaf(t.y, sc[1], sh[1]);
                        t.z = fmaf(t.z, sc[2], sh[2]); t.w = fmaf(t.w, sc[3], sh[3]);
                    }
                    if (p.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
                    stage_chunk<PASSES>(a_hi, a_lo, r, j, t);
                }
            } else {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    *reinterpret_cast<float4*>(a_hi + (j << 4)) = z;
                    if (PASSES == 3) *reinterpret_cast<float4*>(a_lo + (j << 4)) = z;
                }
            }
            // B: weight rows n0 + r (and + 128) of this k-block
            uint8_t* b_base = st + (PASSES == 3 ? 2 : 1) * a_bytes;
            for (int n = r; n < p.n_tile; n += kProducerThreads) {
                uint8_t* b_hi = b_base + n * 128;
                uint8_t* b_lo = b_hi + b_bytes;
                if (n < N) {
                    const float4* g = reinterpret_cast<const float4*>(p.wt + (size_t)(n0 + n) * K + (size_t)kb * kBlockK);
#pragma unroll
                    for (int j = 0; j < 8; ++j) stage_chunk<PASSES>(b_hi, b_lo, n, j, __ldg(g + j));
                }
            }
            fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(bar_full + 8 * stage);
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }

        // ===================== epilogue: TMEM -> registers -> global =====================
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int cb = 0; cb < N; cb += 32) {
            float acc[32];
            tmem_ld32(t_lane + (uint32_t)cb, acc);   // warp-collective: executed by every lane
            if (row_ok) {
                const int ncol = min(32, N - cb);
                float* o = p.out + (size_t)v * p.ldo + n0 + cb;
                const float* rs = p.res ? p.res + (size_t)v * p.F + n0 + cb : nullptr;
                const float* bs = p.bias ? p.bias + n0 + cb : nullptr;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (j < ncol) {   // N is a multiple of 16, so whole float4 groups
                        float4 t = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                        if (bs) { float4 b = __ldg(reinterpret_cast<const float4*>(bs + j)); t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w; }
                        if (rs) { float4 q = __ldg(reinterpret_cast<const float4*>(rs + j)); t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
                        *reinterpret_cast<float4*>(o + j) = t;
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ===================== MMA issuer: one thread =====================
        if (lane == 0) {
            // instruction descriptor: D = F32 [4,6), A = B = TF32 [7,10) [10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + a_bytes);
                const uint32_t bb = st + (PASSES == 3 ? 2 : 1) * a_bytes;
                const uint64_t b_hi = make_desc(bb), b_lo = make_desc(bb + b_bytes);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);  // 32 bytes per K step inside the swizzle row
                    umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
                    if (PASSES == 3) {
                        umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, 1);
                        umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
                    }
                }
                umma_commit(bar_empty + 8 * stage);   // stage reusable once these MMAs have read it
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(bar_acc);                      // accumulator complete -> epilogue
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// [K, F] row-major -> [F, K] row-major (K-major operand for the tensor core)
__global__ void k_transpose(const float* __restrict__ in, int K, int F, float* __restrict__ out) {
    __shared__ float tile[32][33];
    int k0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int k = k0 + i, f = f0 + threadIdx.x;
        tile[i][threadIdx.x] = (k < K && f < F) ? in[(size_t)k * F + f] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int f = f0 + i, k = k0 + threadIdx.x;
        if (f < F && k < K) out[(size_t)f * K + k] = tile[threadIdx.x][i];
    }
}

// GroupNorm statistics -> per-channel affine: scale = rstd*gamma, shift = beta - mean*scale
__global__ void k_gn_affine(const double* __restrict__ sums, int V, const int* __restrict__ v_dev, int C, int cpg,
                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                            float* __restrict__ scale, float* __restrict__ shift) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (v_dev) V = min(V, *v_dev);
    int g = c / cpg;
    double n = (double)V * cpg;
    double mean = sums[2 * g] / n;
    double var = sums[2 * g + 1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float a = rstd * (gamma ? gamma[c] : 1.0f);
    scale[c] = a;
    shift[c] = (beta ? beta[c] : 0.0f) - (float)mean * a;
}

}  // namespace

extern "C" {

int ltn_transpose(const float* in, int K, int F, float* out, void* stream) {
    if (K <= 0 || F <= 0) return 0;
    dim3 grid((K + 31) / 32, (F + 31) / 32), block(32, 8);
    k_transpose<<<grid, block, 0, (cudaStream_t)stream>>>(in, K, F, out);
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_gn_affine(const double* sums, int V, const int* v_dev, int C, int G, const float* gamma, const float* beta,
                  float eps, float* scale, float* shift, void* stream) {
    if (G <= 0 || C % G) return -2;
    k_gn_affine<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, V, v_dev, C, C / G, gamma, beta, eps, scale, shift);
    LTN_CHECK_LAUNCH();
    return 0;
}

// Fused gather + GEMM on tcgen05.  See the header of this file; the C ABI is documented in
// include/latticenet_b200.h.  passes: 3 = fp32-parity split (default), 1 = single-pass TF32.
int ltn_conv_tc(const float* x, int Vx, const int* vx_dev, const int* nbr, int Vq, const int* vq_dev, int C, int S,
                const float* wt, int F, const float* a_scale, const float* a_shift, int relu, const float* bias,
                const float* res, float* out, int ldo, int passes, void* stream) {
    if (Vq <= 0) return 0;
    if (C <= 0 || C % kBlockK || C > 256 || F <= 0 || F % 16 || S < 1 || (passes != 1 && passes != 3) || ldo % 4) return -2;
    if ((a_scale == nullptr) != (a_shift == nullptr)) return -2;
    ConvParams p;
    p.x = x; p.nbr = nbr; p.wt = wt; p.a_scale = a_scale; p.a_shift = a_shift; p.bias = bias; p.res = res; p.out = out;
    p.vq_dev = vq_dev; p.vx_dev = vx_dev; p.Vq = Vq; p.Vx = Vx; p.C = C; p.S = nbr ? S : 1; p.F = F; p.ldo = ldo; p.relu = relu;
    // output channels per CTA: all of them when they fit one accumulator, else the fewest equal tiles
    int ny = (F + 255) / 256;
    int n_tile = ((F + ny - 1) / ny + 15) / 16 * 16;
    p.n_tile = n_tile;
    size_t stage_bytes = (size_t)(passes == 3 ? 2 : 1) * (kBlockM * 128 + (size_t)n_tile * 128);
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return -3;
    p.stages = stages;
    size_t smem = stage_bytes * stages + 1024;
    dim3 grid((Vq + kBlockM - 1) / kBlockM, ny);
    cudaError_t e;
    if (passes == 3) {
        e = cudaFuncSetAttribute(k_conv_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k_conv_tc<3><<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k_conv_tc<1><<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    }
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
