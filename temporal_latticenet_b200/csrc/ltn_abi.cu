// Version entry point of the C ABI (include/latticenet_b200.h).
extern "C" int ltn_version(void) { return 100; }
