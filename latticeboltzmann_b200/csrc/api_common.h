// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif
int lbm_fail(int code, const char *fmt, ...);
#ifdef __cplusplus
}
#endif

#define LBM_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t lbm_e_ = (call);                                                                \
        if (lbm_e_ != cudaSuccess) {                                                                \
            cudaGetLastError();                                                                     \
            return lbm_fail(-2, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(lbm_e_)); \
        }                                                                                           \
    } while (0)
