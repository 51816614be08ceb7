// CEncryptionCycler.h -- drop-in for the reference class of the same name
// (/root/reference/Modulate/CEncryptionCycler.h:3-10): same public signature, stateless, callers
// construct it on the stack per use (CArk.cpp:338, CArk.cpp:1135, Modulate.cpp:485).
//
// The work is done on the GPU through the C ABI (include/modulate_b200.h, mod_cycle): a host
// buffer is staged through HBM in pipelined slices, a device buffer is cycled where it lies.
// There is no CPU implementation behind this class; if the CUDA path fails the call aborts the
// process with the library's message, because the reference signature returns void.
#pragma once

class CEncryptionCycler
{
public:
    void Cycle(unsigned char* lpData, unsigned int liDataSize, int liInitialKey);
};
