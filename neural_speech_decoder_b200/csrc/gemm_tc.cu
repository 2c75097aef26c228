// K2 tensor-core path (tcgen05 + TMA).  Placeholder entry point until the kernel lands: fails loudly.
#include "common.cuh"

extern "C" int nsd_gemm_bf16(int, int, int, int, int, const void*, int, const void*, int, void*, int, int,
                             const float*, float, void*) {
    nsd::set_error("gemm_bf16: tcgen05 path not built yet");
    return NSD_ERR_INVALID;
}
