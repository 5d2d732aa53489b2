// Segmented reductions and normalisation over lattice vertices: scatter_max (+argmax), scatter_add,
// GroupNorm statistics / apply (+ReLU).  sm_100a.
#include "ltn_common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t ord_enc(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_dec(uint32_t e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}

// packed[v,c] = max over rows of (ordered(value) << 32 | ~row): the largest value wins and, between
// equal values, the SMALLEST source row (deterministic; torch_scatter's CUDA path is a race there).
__global__ void __launch_bounds__(kThreads)
k_scatter_max(const float* __restrict__ src, const int* __restrict__ idx, int R, int C,
              unsigned long long* packed, int V) {
    long long total = (long long)R * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int row = (int)(t / C);
        int c = (int)(t - (long long)row * C);
        int id = __ldg(idx + row);
        id = id < 0 ? 0 : id;  // lattice_modules.py:479-480
        if (id >= V) continue;
        unsigned long long key = ((unsigned long long)ord_enc(__ldg(src + t)) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)row);
        atomicMax(packed + (size_t)id * C + c, key);
    }
}

// empty segment: value 0, argmax = R (torch_scatter 2.0.4 sentinel, quirk Q3)
__global__ void __launch_bounds__(kThreads)
k_scatter_max_decode(const unsigned long long* __restrict__ packed, long long n, int R, float* __restrict__ out,
                     long long* __restrict__ arg) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        unsigned long long k = packed[t];
        if (k == 0ull) { out[t] = 0.f; arg[t] = R; }
        else { out[t] = ord_dec((uint32_t)(k >> 32)); arg[t] = (long long)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull)); }
    }
}

__global__ void __launch_bounds__(kThreads)
k_scatter_add(const float* __restrict__ src, const int* __restrict__ idx, int R, int C, float* out, int V) {
    long long total = (long long)R * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int row = (int)(t / C);
        int c = (int)(t - (long long)row * C);
        int id = __ldg(idx + row);
        id = id < 0 ? 0 : id;
        if (id < V) atomicAdd(out + (size_t)id * C + c, __ldg(src + t));
    }
}

// GroupNorm statistics over x [V,C]: per group sum and sum of squares in double.
// One warp walks rows; each lane owns channels lane, lane+32, ... (C <= 512).
constexpr int kMaxCPL = 16;
__global__ void __launch_bounds__(kThreads)
k_gn_stats(const float* __restrict__ x, int V, const int* __restrict__ v_dev, int C, int cpg, double* sums /*[G,2]*/) {
    extern __shared__ double sh[];  // [G*2]
    if (v_dev) V = min(V, *v_dev);
    const int G = C / cpg;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double s1[kMaxCPL], s2[kMaxCPL];  // double accumulators: var = E[x^2]-mean^2 must not cancel in fp32
#pragma unroll
    for (int j = 0; j < kMaxCPL; ++j) { s1[j] = 0.0; s2[j] = 0.0; }
    int row = blockIdx.x * wpb + (threadIdx.x >> 5);
    int stride = gridDim.x * wpb;
    for (int v = row; v < V; v += stride) {
#pragma unroll
        for (int j = 0; j < kMaxCPL; ++j) {
            int c = lane + 32 * j;
            if (c < C) {
                double a = (double)__ldg(x + (size_t)v * C + c);
                s1[j] += a;
                s2[j] = fma(a, a, s2[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxCPL; ++j) {
        int c = lane + 32 * j;
        if (c < C) { atomicAdd(&sh[2 * (c / cpg)], s1[j]); atomicAdd(&sh[2 * (c / cpg) + 1], s2[j]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(sums + i, sh[i]);
}

// y = relu?( (x - mean_g) * rstd_g * gamma_c + beta_c ), folded into per-channel a_c, b_c in smem
__global__ void __launch_bounds__(kThreads)
k_gn_apply(const float* __restrict__ x, int V, const int* __restrict__ v_dev, int C, int cpg,
           const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
           float eps, int relu, float* __restrict__ y) {
    extern __shared__ float ab[];  // a[C], b[C]
    if (v_dev) V = min(V, *v_dev);
    double n = (double)V * cpg;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int g = c / cpg;
        double mean = sums[2 * g] / n;
        double var = sums[2 * g + 1] / n - mean * mean;
        if (var < 0.0) var = 0.0;
        float rstd = (float)(1.0 / sqrt(var + (double)eps));
        float a = rstd * (gamma ? __ldg(gamma + c) : 1.0f);
        ab[c] = a;
        ab[C + c] = (beta ? __ldg(beta + c) : 0.0f) - (float)mean * a;
    }
    __syncthreads();
    long long total = (long long)V * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int c = (int)(t % C);
        float r = fmaf(__ldg(x + t), ab[c], ab[C + c]);
        y[t] = (relu && r < 0.f) ? 0.f : r;
    }
}

inline int grid_for(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = 148LL * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

// packed: [V,C] u64 scratch (zeroed here); out [V,C] f32; arg [V,C] i64
int ltn_scatter_max(const float* src, const int* idx, int R, int C, int V, unsigned long long* packed, float* out,
                    long long* arg, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (V <= 0 || C <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(packed, 0, sizeof(unsigned long long) * (size_t)V * C, st);
    if (e != cudaSuccess) return (int)e;
    if (R > 0) {
        k_scatter_max<<<grid_for((long long)R * C, kThreads), kThreads, 0, st>>>(src, idx, R, C, packed, V);
        LTN_CHECK_LAUNCH();
    }
    k_scatter_max_decode<<<grid_for((long long)V * C, kThreads), kThreads, 0, st>>>(packed, (long long)V * C, R, out, arg);
    LTN_CHECK_LAUNCH();
    return 0;
}

// out [V,C] must be zeroed (or hold the accumulator) on entry
int ltn_scatter_add(const float* src, const int* idx, int R, int C, float* out, int V, void* stream) {
    if (R <= 0 || C <= 0) return 0;
    k_scatter_add<<<grid_for((long long)R * C, kThreads), kThreads, 0, (cudaStream_t)stream>>>(src, idx, R, C, out, V);
    LTN_CHECK_LAUNCH();
    return 0;
}

// sums: [G,2] double, zeroed here
int ltn_gn_stats(const float* x, int V, const int* v_dev, int C, int G, double* sums, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (G <= 0 || C % G || C > 32 * kMaxCPL) return -2;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)G, st);
    if (e != cudaSuccess) return (int)e;
    if (V <= 0) return 0;
    int wpb = kThreads / 32;
    int blocks = (V + wpb * 16 - 1) / (wpb * 16);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    k_gn_stats<<<blocks, kThreads, sizeof(double) * 2 * G, st>>>(x, V, v_dev, C, C / G, sums);
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_gn_apply(const float* x, int V, const int* v_dev, int C, int G, const double* sums, const float* gamma,
                 const float* beta, float eps, int relu, float* y, void* stream) {
    if (V <= 0) return 0;
    if (G <= 0 || C % G) return -2;
    k_gn_apply<<<grid_for((long long)V * C, kThreads), kThreads, sizeof(float) * 2 * C, (cudaStream_t)stream>>>(
        x, V, v_dev, C, C / G, sums, gamma, beta, eps, relu, y);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
