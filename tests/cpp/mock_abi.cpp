// tests/cpp/mock_abi.cpp -- TEST INFRASTRUCTURE.  A stand-in for the CUDA library's C ABI that
// answers the facade's calls with the CPU ORACLE (oracle/cycle_oracle.c), so the host-side C++
// (CArk, ArkHeader, the CLI) can be exercised end to end on a box without a GPU.  It is linked only
// into tests/_build/modulate_mock; the product library never contains it.
#include <cstdlib>
#include <cstring>

#include "../../include/modulate_b200.h"

extern "C" {
void oracle_cycle_at(unsigned char* data, uint64_t size, int32_t key, uint64_t pos);
int32_t oracle_key_jump(int32_t key, uint64_t pos);
struct oracle_desc {
    uint64_t src_off, dst_off;
    uint32_t len;
    int32_t key;
};
void oracle_cycle_batch(const oracle_desc* d, uint64_t n, const unsigned char* src, unsigned char* dst);

int mod_abi_version(void) { return MOD_ABI_VERSION; }
int mod_device_count(void) { return 0; }
int mod_init(int) { return MOD_OK; }
void mod_shutdown(void) {}
const char* mod_last_error(void) { return "mock ABI (oracle-backed, tests only)"; }
uint64_t mod_launch_count(void) { return 0; }
void* mod_host_alloc(uint64_t bytes) { return std::malloc(bytes ? bytes : 1); }
int mod_host_free(void* p)
{
    std::free(p);
    return MOD_OK;
}
int mod_cycle(void* data, uint64_t len, int32_t key)
{
    oracle_cycle_at((unsigned char*)data, len, key, 0);
    return MOD_OK;
}
int32_t mod_key_jump(int32_t key, uint64_t pos) { return oracle_key_jump(key, pos); }
int mod_cycle_batch(const mod_desc* descs, uint64_t n, const void* src, uint64_t src_bytes, void* dst, uint64_t dst_bytes)
{
    for (uint64_t i = 0; i < n; ++i)
        if (descs[i].src_off + descs[i].len > src_bytes || descs[i].dst_off + descs[i].len > dst_bytes)
            return MOD_ERR_ARG;
    static_assert(sizeof(mod_desc) == sizeof(oracle_desc), "descriptor layouts must match");
    oracle_cycle_batch((const oracle_desc*)descs, n, (const unsigned char*)src, (unsigned char*)dst);
    return MOD_OK;
}
}
