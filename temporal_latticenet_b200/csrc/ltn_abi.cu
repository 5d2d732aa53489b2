// Version / bookkeeping entry points of the C ABI (include/latticenet_b200.h).
unsigned long long g_ltn_launches = 0;

extern "C" int ltn_version(void) { return 100; }

// number of kernels this library has launched since load, modulo 2^31 (bench.py's gpu_launches)
extern "C" int ltn_launch_count(void) { return (int)(g_ltn_launches & 0x7FFFFFFFull); }
