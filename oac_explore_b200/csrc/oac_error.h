// Error convention of the C ABI: int return codes + a thread-local message.
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace oac {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* what);
}  // namespace oac

#define OAC_CUDA(expr)                                                    \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) return ::oac::set_cuda_error(_e, #expr);   \
    } while (0)
