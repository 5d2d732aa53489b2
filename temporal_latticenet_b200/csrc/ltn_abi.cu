// Version / bookkeeping entry points of the C ABI (include/latticenet_b200.h).
unsigned long long g_ltn_launches = 0;

extern "C" int ltn_version(void) { return 100; }

// number of kernels this library has launched since load, modulo 2^31 (bench.py's gpu_launches)
extern "C" int ltn_launch_count(void) { return (int)(g_ltn_launches & 0x7FFFFFFFull); }

// Writes two ints to device memory with the values carried BY VALUE in the launch (no host buffer whose
// lifetime could race with an asynchronous copy): the per-frame point counts [n, 4n] of the graph engine.
#include "ltn_common.cuh"
namespace {
__global__ void k_set_int2(int* dst, int a, int b) { dst[0] = a; dst[1] = b; }
}  // namespace
extern "C" int ltn_set_int2(int* dst, int a, int b, void* stream) {
    k_set_int2<<<1, 1, 0, (cudaStream_t)stream>>>(dst, a, b);
    LTN_CHECK_LAUNCH();
    return 0;
}
