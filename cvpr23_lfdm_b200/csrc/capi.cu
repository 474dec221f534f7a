// capi.cu — C-ABI glue of liblfdm_b200.so (see include/lfdm_b200.h): engine dispatch + library info.
#include "common.cuh"

int lfdm_conv_simt(const lfdm_conv_desc* d, cudaStream_t stream);
int lfdm_conv_tc(const lfdm_conv_desc* d, cudaStream_t stream);

extern "C" int lfdm_conv(const lfdm_conv_desc* d, int engine, void* stream) {
    if (!d) return LFDM_E_BADARG;
    if (engine == LFDM_ENGINE_TC) return lfdm_conv_tc(d, (cudaStream_t)stream);
    if (engine == LFDM_ENGINE_SIMT) return d->rot_cos ? LFDM_E_UNSUPP : lfdm_conv_simt(d, (cudaStream_t)stream);
    return LFDM_E_BADARG;
}

extern "C" int lfdm_version(int* arch, int* has_tc) {
    if (arch) *arch = 100;
    if (has_tc) *has_tc = 1;
    return 1;
}
