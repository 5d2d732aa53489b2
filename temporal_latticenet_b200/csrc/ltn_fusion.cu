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
    k_gru_pointwise<<<grid_for((long long)V * C, kThreads), kThreads, 0, (cudaStream_t)stream>>>(gi, gh, h, b_hh, V, Vh, v_dev,
                                                                                                  vh_dev, C, out);
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
