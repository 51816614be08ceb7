// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" handles onto the UNMODIFIED reference cipher.  oracle/Makefile compiles this
// file together with /root/reference/Modulate/CEncryptionCycler.cpp (where it lies; no
// reference source is copied into this repo) into oracle/_ref/libcycle_ref.so.  The shim
// contains no arithmetic of its own: every byte is produced by the reference's
// CEncryptionCycler::Cycle (CEncryptionCycler.cpp:4-14).
//
// ref_cycle_parts fans independent (offset, length, key) parts out over host threads, one
// reference Cycle() call per part -- the "all cores" CPU baseline of SURVEY.md section 8(d);
// a single part cannot be split because the reference has no jump-ahead.
#include <cstdint>
#include <thread>
#include <vector>
#include <atomic>

#include "CEncryptionCycler.h"   // -I/root/reference/Modulate

extern "C" {

void ref_cycle(unsigned char* data, unsigned int size, int key)
{
    CEncryptionCycler lCycler;
    lCycler.Cycle(data, size, key);
}

struct ref_part {
    uint64_t off;
    uint32_t len;
    int32_t key;
};

void ref_cycle_parts(unsigned char* base, const ref_part* parts, uint64_t n, int threads)
{
    if (threads < 1)
        threads = 1;
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        CEncryptionCycler lCycler;
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= n)
                return;
            lCycler.Cycle(base + parts[i].off, parts[i].len, parts[i].key);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t)
        pool.emplace_back(work);
    work();
    for (auto& th : pool)
        th.join();
}

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
