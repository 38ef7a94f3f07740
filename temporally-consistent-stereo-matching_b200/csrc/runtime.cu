// Library-wide plumbing of libtcs_b200.so: ABI version, thread-local error text, cached device facts.
#include "tcs_common.cuh"

#include <cstdarg>
#include <cstdlib>
#include <cstring>

namespace tcs {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& slot = cached[dev & 63];
    if (slot == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        slot = n;
    }
    return slot;
}

int carveout_percent(const char* env, int tuned_default) {
    const char* e = getenv(env);
    if (e == nullptr || *e == '\0') return tuned_default;
    const int v = atoi(e);
    return v < 0 ? -1 : (v > 100 ? 100 : v);
}

}  // namespace tcs

extern "C" int tcs_abi_version(void) { return TCS_ABI_VERSION; }

extern "C" const char* tcs_last_error(void) { return tcs::g_error; }
