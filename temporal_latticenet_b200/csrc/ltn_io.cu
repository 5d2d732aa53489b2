// Device-side window assembly: a SemanticKITTI scan arrives as the raw bytes of its .bin file
// ([N,4] float32: x, y, z, reflectance) and leaves as the model's inputs -- positions in the first frame's coordinate
// system of the window, y-up, and the reflectance column -- without a numpy pass on the host.
// Reference: dataloader/kitti_dataloader.py:129-132 (wire format), :160-171 (velo -> world -> first scan of the window ->
// -90 degrees about x, divide by w), DataTransformer.py:88-91 (float64 -> float32).
#include "ltn_common.cuh"

namespace {

struct Mats { double m[3][16]; };   // applied in order: m[0] (velo -> world), m[1] (world -> first scan), m[2] (axis change)

__global__ void __launch_bounds__(256)
k_assemble_scan(const float4* __restrict__ raw, int N, Mats M, int nr_mats, float* __restrict__ pos, float* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float4 p = __ldg(raw + i);
    double v[4] = {(double)p.x, (double)p.y, (double)p.z, 1.0};
    // the reference multiplies the three 4x4 matrices onto the homogeneous points one after the other in float64
    // (np.matmul); the same sequence here, each dot product accumulated in k order
    for (int t = 0; t < nr_mats; ++t) {
        double w[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(M.m[t][4 * r + k], v[k], acc);
            w[r] = acc;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = w[r];
    }
    pos[(size_t)i * 3 + 0] = (float)(v[0] / v[3]);
    pos[(size_t)i * 3 + 1] = (float)(v[1] / v[3]);
    pos[(size_t)i * 3 + 2] = (float)(v[2] / v[3]);
    val[i] = p.w;
}

}  // namespace

extern "C" {

// raw [N,4] float32 (device): the .bin payload.  mats (HOST pointer): nr_mats (1..3) row-major 4x4 float64 matrices applied
// in order to [x,y,z,1]; pos [N,3], val [N,1] (device).
int ltn_assemble_scan(const float* raw, int N, const double* mats, int nr_mats, float* pos, float* val, void* stream) {
    if (N <= 0) return 0;
    if (nr_mats < 1 || nr_mats > 3 || !mats) return -2;
    Mats M;
    for (int t = 0; t < 3; ++t)
        for (int e = 0; e < 16; ++e) M.m[t][e] = t < nr_mats ? mats[16 * t + e] : (e % 5 == 0 ? 1.0 : 0.0);
    k_assemble_scan<<<ltn_blocks(N, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(raw), N, M, nr_mats, pos, val);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
