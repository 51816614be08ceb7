#include "CEncryptionCycler.h"

#include <cstdio>
#include <cstdlib>

#include "../../include/modulate_b200.h"

void CEncryptionCycler::Cycle(unsigned char* lpData, unsigned int liDataSize, int liInitialKey)
{
    const int rc = mod_cycle(lpData, (uint64_t)liDataSize, (int32_t)liInitialKey);
    if (rc != MOD_OK) {
        std::fprintf(stderr, "CEncryptionCycler::Cycle: CUDA path failed (%d): %s\n", rc, mod_last_error());
        std::abort();  // no CPU fallback by design
    }
}
