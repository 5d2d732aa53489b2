// Per-vertex temporal fusion kernels (seq_lattice/lattice_modules.py:17-339): GRU / LSTM gate
// pointwise stages with the index-prefix zero padding folded in, and the AFlow core (neighbour
// gather in h^{t-1}, feature distance, weights, weighted sum) as ONE bandwidth-bound kernel
// instead of the reference's two materialised [V,9C] im2row buffers and ~15 elementwise launches.
#include "ltn_common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// GRUCell pointwise (PyTorch gate order r,z,n; lattice_modules.py:58-62):
//   gi = x W_ih^T + b_ih            [V, 3C]   (precomputed)
//   gh = h W_hh^T + b_hh            [Vh,3C]   (precomputed for the Vh rows that have history)
//   rows v >= Vh were zero-padded (pad_sequence, lattice_modules.py:59-60): h = 0, gh = b_hh
//   r = s(gi_r+gh_r) ; z = s(gi_z+gh_z) ; n = tanh(gi_n + r*gh_n) ; h' = (1-z)*n + z*h
__global__ void __launch_bounds__(kThreads)
k_gru_pointwise(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h,
                const float* __restrict__ b_hh, int V, int Vh, const int* __restrict__ v_dev,
                const int* __restrict__ vh_dev, int C, float* __restrict__ out) {
    if (v_dev) V = min(V, *v_dev);
    if (vh_dev) Vh = min(Vh, *vh_dev);
    long long total = (long long)V * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int v = (int)(t / C);
        int c = (int)(t - (long long)v * C);
        const float* gi_v = gi + (size_t)v * 3 * C;
        float hr, hz, hn, hp;
        if (v < Vh) {
            const float* gh_v = gh + (size_t)v * 3 * C;
            hr = __ldg(gh_v + c); hz = __ldg(gh_v + C + c); hn = __ldg(gh_v + 2 * C + c);
            hp = __ldg(h + t);
        } else {
            hr = __ldg(b_hh + c); hz = __ldg(b_hh + C + c); hn = __ldg(b_hh + 2 * C + c);
            hp = 0.f;
        }
        float r = sigmoidf_(__ldg(gi_v + c) + hr);
        float z = sigmoidf_(__ldg(gi_v + C + c) + hz);
        float n = tanhf(__ldg(gi_v + 2 * C + c) + r * hn);
        out[t] = (1.0f - z) * n + z * hp;
    }
}

// the same, four channels per thread (C % 4 == 0): seven 128-bit loads and one 128-bit store instead of 28 + 4 scalar ones
__device__ __forceinline__ float gru_gate(float gir, float giz, float gin, float hr, float hz, float hn, float hp) {
    const float r = sigmoidf_(gir + hr);
    const float z = sigmoidf_(giz + hz);
    const float n = tanhf(gin + r * hn);
    return (1.0f - z) * n + z * hp;
}
__global__ void __launch_bounds__(kThreads)
k_gru_pointwise4(const float4* __restrict__ gi, const float4* __restrict__ gh, const float4* __restrict__ h,
                 const float4* __restrict__ b_hh, int V, int Vh, const int* __restrict__ v_dev,
                 const int* __restrict__ vh_dev, int C4, float4* __restrict__ out) {
    if (v_dev) V = min(V, *v_dev);
    if (vh_dev) Vh = min(Vh, *vh_dev);
    const long long total = (long long)V * C4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(t / C4);
        const int c = (int)(t - (long long)v * C4);
        const float4* gi_v = gi + (size_t)v * 3 * C4;
        const float4 a = __ldg(gi_v + c), b = __ldg(gi_v + C4 + c), d = __ldg(gi_v + 2 * C4 + c);
        float4 hr, hz, hn, hp;
        if (v < Vh) {
            const float4* gh_v = gh + (size_t)v * 3 * C4;
            hr = __ldg(gh_v + c); hz = __ldg(gh_v + C4 + c); hn = __ldg(gh_v + 2 * C4 + c);
            hp = __ldg(h + t);
        } else {
            hr = __ldg(b_hh + c); hz = __ldg(b_hh + C4 + c); hn = __ldg(b_hh + 2 * C4 + c);
            hp = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 o;
        o.x = gru_gate(a.x, b.x, d.x, hr.x, hz.x, hn.x, hp.x);
        o.y = gru_gate(a.y, b.y, d.y, hr.y, hz.y, hn.y, hp.y);
        o.z = gru_gate(a.z, b.z, d.z, hr.z, hz.z, hn.z, hp.z);
        o.w = gru_gate(a.w, b.w, d.w, hr.w, hz.w, hn.w, hp.w);
        out[t] = o;
    }
}

// k_gru_pointwise4 that ALSO leaves the GroupNorm statistics of its output behind ([G,2] double: per-group sum and sum of
// squares), so the layer that normalises h' next does not read it once more (k_gn_stats).  The launch makes the grid's
// thread count a multiple of C4, so a thread owns ONE channel quad for its whole grid-stride walk and keeps the eight
// partial sums in registers; they meet per block in shared memory and leave as one double atomic per group and block.
__global__ void __launch_bounds__(kThreads)
k_gru_pointwise4_stats(const float4* __restrict__ gi, const float4* __restrict__ gh, const float4* __restrict__ h,
                       const float4* __restrict__ b_hh, int V, int Vh, const int* __restrict__ v_dev,
                       const int* __restrict__ vh_dev, int C4, float4* __restrict__ out, double* __restrict__ sums, int cpg) {
    __shared__ float s_acc[2 * 256];   // per channel: sum | sum of squares (C <= 256)
    if (v_dev) V = min(V, *v_dev);
    if (vh_dev) Vh = min(Vh, *vh_dev);
    const int C = 4 * C4;
    for (int i = threadIdx.x; i < 2 * 256; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const long long total = (long long)V * C4;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(t0 % C4);
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sq = sa;
    for (long long t = t0; t < total; t += (long long)gridDim.x * blockDim.x) {   // stride % C4 == 0: c stays
        const int v = (int)(t / C4);
        const float4* gi_v = gi + (size_t)v * 3 * C4;
        const float4 a = __ldg(gi_v + c), b = __ldg(gi_v + C4 + c), d = __ldg(gi_v + 2 * C4 + c);
        float4 hr, hz, hn, hp;
        if (v < Vh) {
            const float4* gh_v = gh + (size_t)v * 3 * C4;
            hr = __ldg(gh_v + c); hz = __ldg(gh_v + C4 + c); hn = __ldg(gh_v + 2 * C4 + c);
            hp = __ldg(h + t);
        } else {
            hr = __ldg(b_hh + c); hz = __ldg(b_hh + C4 + c); hn = __ldg(b_hh + 2 * C4 + c);
            hp = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 o;
        o.x = gru_gate(a.x, b.x, d.x, hr.x, hz.x, hn.x, hp.x);
        o.y = gru_gate(a.y, b.y, d.y, hr.y, hz.y, hn.y, hp.y);
        o.z = gru_gate(a.z, b.z, d.z, hr.z, hz.z, hn.z, hp.z);
        o.w = gru_gate(a.w, b.w, d.w, hr.w, hz.w, hn.w, hp.w);
        out[t] = o;
        sa.x += o.x; sa.y += o.y; sa.z += o.z; sa.w += o.w;
        sq.x = fmaf(o.x, o.x, sq.x); sq.y = fmaf(o.y, o.y, sq.y); sq.z = fmaf(o.z, o.z, sq.z); sq.w = fmaf(o.w, o.w, sq.w);
    }
    atomicAdd(&s_acc[4 * c], sa.x); atomicAdd(&s_acc[4 * c + 1], sa.y); atomicAdd(&s_acc[4 * c + 2], sa.z); atomicAdd(&s_acc[4 * c + 3], sa.w);
    atomicAdd(&s_acc[256 + 4 * c], sq.x); atomicAdd(&s_acc[256 + 4 * c + 1], sq.y);
    atomicAdd(&s_acc[256 + 4 * c + 2], sq.z); atomicAdd(&s_acc[256 + 4 * c + 3], sq.w);
    __syncthreads();
    const int G = C / cpg;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int k = g * cpg; k < (g + 1) * cpg; ++k) { a += s_acc[k]; b += s_acc[256 + k]; }
        atomicAdd(sums + 2 * g, (double)a);
        atomicAdd(sums + 2 * g + 1, (double)b);
    }
}

// LSTMCell pointwise with c_prev = 0 (lattice_modules.py:36; gate order i,f,g,o):
//   c' = s(i)*tanh(g) ; h' = s(o)*tanh(c')
__global__ void __launch_bounds__(kThreads)
k_lstm_pointwise(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ b_hh, int V,
                 int Vh, const int* __restrict__ v_dev, const int* __restrict__ vh_dev, int C, float* __restrict__ out) {
    if (v_dev) V = min(V, *v_dev);
    if (vh_dev) Vh = min(Vh, *vh_dev);
    long long total = (long long)V * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int v = (int)(t / C);
        int c = (int)(t - (long long)v * C);
        const float* gi_v = gi + (size_t)v * 4 * C;
        const float* gh_v = (v < Vh) ? gh + (size_t)v * 4 * C : b_hh;
        float i = sigmoidf_(__ldg(gi_v + c) + __ldg(gh_v + c));
        float g = tanhf(__ldg(gi_v + 2 * C + c) + __ldg(gh_v + 2 * C + c));
        float o = sigmoidf_(__ldg(gi_v + 3 * C + c) + __ldg(gh_v + 3 * C + c));
        float cc = i * g;
        out[t] = o * tanhf(cc);
    }
}

// AFlow core, one warp per vertex (lattice_modules.py:298-339).  h has Vh rows; vertices >= Vh were
// padded with `pad_value` (-999999, lattice_modules.py:215); absent neighbours gather zeros and are
// masked.  NaN behaviour of the reference (0/0 when every masked distance is 0, quirk Q5) is kept.
constexpr int kAflowMaxCPL = 8;  // C <= 256
__global__ void __launch_bounds__(kThreads)
k_aflow(const float* __restrict__ lv, const float* __restrict__ h, int V, int Vh, const int* __restrict__ v_dev,
        const int* __restrict__ vh_dev, int C, const int* __restrict__ nbr, const float* __restrict__ alpha_p, const float* __restrict__ beta_p,
        const float* __restrict__ bias, float pad_value, int use_center, float* __restrict__ out,
        float* __restrict__ weights_out) {
    int lane = threadIdx.x & 31;
    int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (v_dev) V = min(V, *v_dev);
    if (vh_dev) Vh = min(Vh, *vh_dev);
    if (v >= V) return;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    float x[kAflowMaxCPL];
#pragma unroll
    for (int j = 0; j < kAflowMaxCPL; ++j) {
        int c = lane + 32 * j;
        x[j] = (c < C) ? __ldg(lv + (size_t)v * C + c) : 0.f;
    }
    float nb[LTN_FEXT][kAflowMaxCPL];
    float dist[LTN_FEXT], mask[LTN_FEXT];
#pragma unroll
    for (int k = 0; k < LTN_FEXT; ++k) {
        int id = __ldg(nbr + (size_t)v * LTN_FEXT + k);
        mask[k] = (id >= 0) ? 1.f : 0.f;
        float d2 = 0.f;
#pragma unroll
        for (int j = 0; j < kAflowMaxCPL; ++j) {
            int c = lane + 32 * j;
            float a = 0.f;
            if (c < C) {
                if (id >= Vh) a = pad_value;
                else if (id >= 0) a = __ldg(h + (size_t)id * C + c);
                float d = a - x[j];
                d2 = fmaf(d, d, d2);
            }
            nb[k][j] = a;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        dist[k] = sqrtf(d2) * mask[k];
    }
    if (!use_center) dist[LTN_FEXT - 1] = dist[LTN_FEXT - 1] * 0.f;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < LTN_FEXT; ++k) sum += dist[k];
    float w[LTN_FEXT];
#pragma unroll
    for (int k = 0; k < LTN_FEXT; ++k) {
        float d = dist[k] / sum;                       // 0/0 -> NaN on purpose (Q5)
        float m = (d != d) ? d : fminf(d, alpha);      // torch.min propagates NaN, fminf does not
        w[k] = (alpha - m) * beta * mask[k];
    }
    if (!use_center) w[LTN_FEXT - 1] = w[LTN_FEXT - 1] * 0.f;
#pragma unroll
    for (int j = 0; j < kAflowMaxCPL; ++j) {
        int c = lane + 32 * j;
        if (c < C) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < LTN_FEXT; ++k) acc = fmaf(nb[k][j], w[k], acc);
            out[(size_t)v * C + c] = acc + (bias ? __ldg(bias + c) : 0.f);
        }
    }
    if (weights_out && lane < LTN_FEXT) {
        float wl = w[0];
#pragma unroll
        for (int k = 1; k < LTN_FEXT; ++k) if (lane == k) wl = w[k];
        weights_out[(size_t)v * LTN_FEXT + lane] = wl;
    }
}

inline int grid_for(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = 148LL * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int ltn_gru_pointwise(const float* gi, const float* gh, const float* h, const float* b_hh, int V, int Vh, const int* v_dev,
                      const int* vh_dev, int C, float* out, void* stream) {
    if (V <= 0) return 0;
    if (C % 4 == 0) {
        k_gru_pointwise4<<<grid_for((long long)V * (C / 4), kThreads), kThreads, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(gi), reinterpret_cast<const float4*>(gh), reinterpret_cast<const float4*>(h),
            reinterpret_cast<const float4*>(b_hh), V, Vh, v_dev, vh_dev, C / 4, reinterpret_cast<float4*>(out));
        LTN_CHECK_LAUNCH();
        return 0;
    }
    k_gru_pointwise<<<grid_for((long long)V * C, kThreads), kThreads, 0, (cudaStream_t)stream>>>(gi, gh, h, b_hh, V, Vh, v_dev,
                                                                                                  vh_dev, C, out);
    LTN_CHECK_LAUNCH();
    return 0;
}

// ltn_gru_pointwise + the GroupNorm statistics of h' (sums [groups, 2] double, ZEROED by the caller; C % 4 == 0, C <= 256,
// C % groups == 0): see k_gru_pointwise4_stats
int ltn_gru_pointwise_stats(const float* gi, const float* gh, const float* h, const float* b_hh, int V, int Vh, const int* v_dev,
                            const int* vh_dev, int C, float* out, double* sums, int groups, void* stream) {
    if (V <= 0) return 0;
    if (C % 4 || C > 256 || groups <= 0 || C % groups || !sums) return -2;
    const int C4 = C / 4;
    long long want = ((long long)V * C4 + kThreads - 1) / kThreads;
    // every block ends with one double atomic per group on the SAME 2 x groups addresses: with 148 x 8 blocks those serialised
    // reductions, not the streaming pass, set the kernel's duration (29.5 us against the plain kernel's 14.0 in the ncu launch
    // list); two blocks per SM keep the memory system just as busy (each thread has seven 128-bit loads in flight per step)
    if (want > 148 * 2) want = 148 * 2;
    // grid * kThreads must be a multiple of C4 (a thread keeps its channel quad): C4 = 2^a * 3^b here, so a multiple of C4 / gcd(C4, kThreads)
    int g = C4, th = kThreads;
    while (th) { int r = g % th; g = th; th = r; }
    const int unit = C4 / g;
    int grid = (int)((want + unit - 1) / unit) * unit;
    if (grid < unit) grid = unit;
    k_gru_pointwise4_stats<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(gi), reinterpret_cast<const float4*>(gh), reinterpret_cast<const float4*>(h),
        reinterpret_cast<const float4*>(b_hh), V, Vh, v_dev, vh_dev, C4, reinterpret_cast<float4*>(out), sums, C / groups);
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_lstm_pointwise(const float* gi, const float* gh, const float* b_hh, int V, int Vh, const int* v_dev,
                       const int* vh_dev, int C, float* out, void* stream) {
    if (V <= 0) return 0;
    k_lstm_pointwise<<<grid_for((long long)V * C, kThreads), kThreads, 0, (cudaStream_t)stream>>>(gi, gh, b_hh, V, Vh, v_dev,
                                                                                                   vh_dev, C, out);
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_aflow(const float* lv, const float* h, int V, int Vh, const int* v_dev, const int* vh_dev, int C, const int* nbr,
              const float* alpha, const float* beta, const float* bias, float pad_value, int use_center, float* out,
              float* weights_out, void* stream) {
    if (V <= 0) return 0;
    if (C > 32 * kAflowMaxCPL) return -2;
    k_aflow<<<ltn_blocks((long long)V * 32, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        lv, h, V, Vh, v_dev, vh_dev, C, nbr, alpha, beta, bias, pad_value, use_center, out, weights_out);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
