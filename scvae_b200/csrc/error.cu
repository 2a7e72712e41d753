// Thread-local error string + ABI version for libscvae_b200.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace scvae {
static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
static long long g_launches = 0;
int pdl_mask() {
    static int mask = -1;
    if (mask < 0) {
        const char *e = getenv("SCVAE_PDL");
        mask = e ? atoi(e) : 0;
    }
    return mask;
}
void count_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }
}  // namespace scvae

extern "C" int scvae_abi_version(void) { return SCVAE_B200_ABI_VERSION; }
extern "C" const char *scvae_last_error(void) { return scvae::g_error; }
extern "C" long long scvae_launch_count(void) { return __atomic_load_n(&scvae::g_launches, __ATOMIC_RELAXED); }
