// PointNetSeqModule front end (seq_lattice/lattice_modules.py:448-530) as two kernels instead of three
// materialised [4N, 16/32/64] activations + scatter_max + index_select + cat + masked_fill:
//
//   k_pointnet_mlp_max : per distributed row  MLP 4 -> 16 -> 32 -> 64 (ReLU between, none after the last,
//                        lattice_modules.py:460-473) entirely in registers, weights broadcast from shared
//                        memory, then the segmented max onto the row's vertex with one 64-bit atomicMax per
//                        channel (value in the high word, ~row in the low word: largest value wins, smallest
//                        row breaks ties -- torch_scatter's result up to its own race on ties).
//   k_pointnet_decode  : per (vertex, channel) unpack max / arg-max, apply quirk Q3 literally
//                        (`argmax_clone[argmax > argmax.shape[0]] = 0`, lattice_modules.py:513-514), gather the
//                        barycentric weight of the winning row, concatenate [max(64) | bary(64)] and zero the
//                        vertices with fewer than 4 contributing rows (lattice_modules.py:519-530).
//
// HBM-bound by construction: reads 4N x 24 B of rows + ids once, writes V x 512 B; the 4N x 64 atomics
// resolve in L2 (V x 64 x 8 B = 7 MB at V = 14k).
#include "ltn_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int D0 = 4, D1 = 16, D2 = 32, D3 = 64;

__device__ __forceinline__ uint32_t ord_enc(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_dec(uint32_t e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}

struct MlpWeights {
    const float *w1, *b1, *w2, *b2, *w3, *b3;  // nn.Linear layout [out, in]
};

__global__ void __launch_bounds__(kThreads, 2)
k_pointnet_mlp_max(const float* __restrict__ rows, int width, const int* __restrict__ idx, int R,
                   const int* __restrict__ r_dev, MlpWeights w, int V, const int* __restrict__ v_dev,
                   unsigned long long* packed) {
    // weights transposed to [in][out] so a thread's inner loop reads consecutive (broadcast) words
    __shared__ __align__(16) float s1[D0 * D1], s2[D1 * D2], s3[D2 * D3], sb1[D1], sb2[D2], sb3[D3];
    for (int i = threadIdx.x; i < D0 * D1; i += kThreads) s1[(i % D0) * D1 + i / D0] = __ldg(w.w1 + i);
    for (int i = threadIdx.x; i < D1 * D2; i += kThreads) s2[(i % D1) * D2 + i / D1] = __ldg(w.w2 + i);
    for (int i = threadIdx.x; i < D2 * D3; i += kThreads) s3[(i % D2) * D3 + i / D2] = __ldg(w.w3 + i);
    for (int i = threadIdx.x; i < D1; i += kThreads) sb1[i] = __ldg(w.b1 + i);
    for (int i = threadIdx.x; i < D2; i += kThreads) sb2[i] = __ldg(w.b2 + i);
    for (int i = threadIdx.x; i < D3; i += kThreads) sb3[i] = __ldg(w.b3 + i);
    __syncthreads();
    if (r_dev) R = min(R, *r_dev);
    if (v_dev) V = min(V, *v_dev);
    for (int row = blockIdx.x * kThreads + threadIdx.x; row < R; row += gridDim.x * kThreads) {
        int id = __ldg(idx + row);
        id = id < 0 ? 0 : id;                       // lattice_modules.py:479-480
        if (id >= V) continue;
        const float* in = rows + (size_t)row * width;
        float x[D0];
#pragma unroll
        for (int i = 0; i < D0; ++i) x[i] = __ldg(in + i);
        float h1[D1];
#pragma unroll
        for (int o = 0; o < D1; ++o) h1[o] = sb1[o];
#pragma unroll
        for (int i = 0; i < D0; ++i)
#pragma unroll
            for (int o = 0; o < D1; ++o) h1[o] = fmaf(x[i], s1[i * D1 + o], h1[o]);
        float h2[D2];
#pragma unroll
        for (int o = 0; o < D2; ++o) h2[o] = sb2[o];
#pragma unroll
        for (int i = 0; i < D1; ++i) {
            const float a = fmaxf(h1[i], 0.f);
#pragma unroll
            for (int o = 0; o < D2; o += 4) {
                const float4 ww = *reinterpret_cast<const float4*>(s2 + i * D2 + o);
                h2[o] = fmaf(a, ww.x, h2[o]); h2[o + 1] = fmaf(a, ww.y, h2[o + 1]);
                h2[o + 2] = fmaf(a, ww.z, h2[o + 2]); h2[o + 3] = fmaf(a, ww.w, h2[o + 3]);
            }
        }
#pragma unroll
        for (int i = 0; i < D2; ++i) h2[i] = fmaxf(h2[i], 0.f);
        unsigned long long* dst = packed + (size_t)id * D3;
        const unsigned long long low = (unsigned long long)(0xFFFFFFFFu - (uint32_t)row);
        // last layer in four quarters of 16 outputs to bound the live registers (two blocks per SM).  The 16
        // current maxima of a quarter are fetched with 8 independent 128-bit loads issued BEFORE its 512 FMAs
        // (one overlapped L2 round trip instead of 16 dependent ones); afterwards an atomic is issued only
        // where this row actually raises the maximum.
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            unsigned long long cur[16];
#pragma unroll
            for (int o = 0; o < 16; o += 2) {
                const ulonglong2 c2 = __ldcg(reinterpret_cast<const ulonglong2*>(dst + q * 16 + o));
                cur[o] = c2.x;
                cur[o + 1] = c2.y;
            }
            float y[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) y[o] = sb3[q * 16 + o];
#pragma unroll
            for (int i = 0; i < D2; ++i) {
                const float a = h2[i];
#pragma unroll
                for (int o = 0; o < 16; o += 4) {
                    const float4 ww = *reinterpret_cast<const float4*>(s3 + i * D3 + q * 16 + o);
                    y[o] = fmaf(a, ww.x, y[o]); y[o + 1] = fmaf(a, ww.y, y[o + 1]);
                    y[o + 2] = fmaf(a, ww.z, y[o + 2]); y[o + 3] = fmaf(a, ww.w, y[o + 3]);
                }
            }
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const unsigned long long key = ((unsigned long long)ord_enc(y[o]) << 32) | low;
                if (cur[o] < key) atomicMax(dst + q * 16 + o, key);
            }
        }
    }
}

// out [V, 128] = [ max | bary of the arg-max row (Q3) ], rows with fewer than min_rows contributors zeroed
__global__ void __launch_bounds__(kThreads)
k_pointnet_decode(const unsigned long long* __restrict__ packed, int V, const int* __restrict__ v_dev, int R,
                  const int* __restrict__ r_dev, const float* __restrict__ rows, int width, const double* __restrict__ vert_acc,
                  int min_rows, float* __restrict__ out) {
    if (r_dev) R = min(R, *r_dev);
    if (v_dev) V = min(V, *v_dev);
    long long total = (long long)V * D3;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        const int v = (int)(t / D3), c = (int)(t - (long long)v * D3);
        const unsigned long long k = packed[t];
        float val = 0.f;
        long long arg = R;                                  // torch_scatter's empty-segment sentinel
        if (k != 0ull) {
            val = ord_dec((uint32_t)(k >> 32));
            arg = (long long)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
        }
        if (arg > V) arg = 0;                               // quirk Q3, literally (row index compared with V)
        if (arg > R - 1) arg = R - 1;
        float bary = (R > 0) ? __ldg(rows + (size_t)arg * width + (width - 1)) : 0.f;
        if (min_rows > 0 && vert_acc[(size_t)v * 4 + 3] < (double)min_rows) { val = 0.f; bary = 0.f; }
        out[(size_t)v * (2 * D3) + c] = val;
        out[(size_t)v * (2 * D3) + D3 + c] = bary;
    }
}

}  // namespace

extern "C" {

// rows [R, width] (first 4 columns feed the MLP, last column = barycentric weight), idx [R];
// packed [V,64] u64 scratch (zeroed here); vert_acc: the accumulator ltn_distribute left ([cap,4] double,
// [:,3] = rows per vertex); min_rows = 4 (0 disables the mask: early max-pool fusion); out [V,128].
int ltn_pointnet(const float* rows, int width, const int* idx, int R, const int* r_dev, const float* w1, const float* b1,
                 const float* w2, const float* b2, const float* w3, const float* b3, int V, const int* v_dev,
                 unsigned long long* packed, const double* vert_acc, int min_rows, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (V <= 0) return 0;
    if (width != D0 + 1) return -2;
    cudaError_t e = cudaMemsetAsync(packed, 0, sizeof(unsigned long long) * (size_t)V * D3, st);
    if (e != cudaSuccess) return (int)e;
    if (R > 0) {
        MlpWeights w{w1, b1, w2, b2, w3, b3};
        int blocks = (R + kThreads - 1) / kThreads;
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_pointnet_mlp_max<<<blocks, kThreads, 0, st>>>(rows, width, idx, R, r_dev, w, V, v_dev, packed);
        LTN_CHECK_LAUNCH();
    }
    long long total = (long long)V * D3;
    int blocks = (int)((total + kThreads - 1) / kThreads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_pointnet_decode<<<blocks, kThreads, 0, st>>>(packed, V, v_dev, R, r_dev, rows, width, vert_acc, min_rows, out);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
