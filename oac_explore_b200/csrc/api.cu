// Error plumbing and version query of the C ABI (include/oac_b200.h).
#include <cuda_runtime.h>
#include <string>

#include "oac_error.h"
#include "../../include/oac_b200.h"

namespace oac {
static thread_local std::string g_last_error = "";

int set_error(int code, const char* msg) {
    g_last_error = msg ? msg : "";
    return code;
}
int set_cuda_error(cudaError_t e, const char* what) {
    g_last_error = std::string(what ? what : "") + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    cudaGetLastError();   // clear the sticky-less error state
    return (int)e;
}
}  // namespace oac

extern "C" const char* oac_last_error_string(void) { return oac::g_last_error.c_str(); }
extern "C" int oac_abi_version(void) { return OAC_ABI_VERSION; }
