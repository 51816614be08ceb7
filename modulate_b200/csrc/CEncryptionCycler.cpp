#include "CEncryptionCycler.h"

#include <cstdio>
#include <cstdlib>

#include "../../include/modulate_b200.h"

void CEncryptionCycler::Cycle(unsigned char* lpData, unsigned int liDataSize, int liInitialKey)
{
    // Host buffers of 256 MiB and more are split by offset range over every visible GPU (MOD_DEVICES =
    // bit mask restricts them): the reference's single serial loop (CEncryptionCycler.cpp:9-13) becomes
    // one jump-ahead + one stream set per device.  Everything else runs on the calling thread's device.
    const unsigned int kuShardThreshold = 256u << 20;
    int rc;
    if (liDataSize >= kuShardThreshold && mod_device_count() > 1 && !mod_is_device_pointer(lpData)) {
        const char* lpMask = std::getenv("MOD_DEVICES");
        rc = mod_cycle_sharded(lpData, (uint64_t)liDataSize, (int32_t)liInitialKey, lpMask && *lpMask ? std::strtoull(lpMask, nullptr, 0) : 0);
    } else {
        rc = mod_cycle(lpData, (uint64_t)liDataSize, (int32_t)liInitialKey);
    }
    if (rc != MOD_OK) {
        std::fprintf(stderr, "CEncryptionCycler::Cycle: CUDA path failed (%d): %s\n", rc, mod_last_error());
        std::abort();  // no CPU fallback by design
    }
}
