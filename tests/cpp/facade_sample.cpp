// tests/cpp/facade_sample.cpp -- a caller written the way the reference's own command handlers are
// (Modulate.cpp:291-317 Unpack, :452-502 Decode): it includes only the facade headers and uses the
// classes by their reference names.  Built by tests/test_facade_host.py against the mock ABI (CPU)
// to check that the headers are self-contained and the signatures are the drop-in ones.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "CArk.h"
#include "CEncryptionCycler.h"
#include "Error.h"
#include "Settings.h"

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::printf("usage: facade_sample <main_xxx.hdr> <out_dir/>\n");
        return 2;
    }
    // 1. the cipher on a caller-owned buffer, stack-constructed like CArk.cpp:338-339
    unsigned char lacBuffer[64];
    std::memset(lacBuffer, 0, sizeof(lacBuffer));
    CEncryptionCycler lCycler;
    lCycler.Cycle(lacBuffer, sizeof(lacBuffer), (int)CSettings::kuEncryptedPS4Key);
    std::printf("keystream:");
    for (int ii = 0; ii < 8; ++ii)
        std::printf(" %02x", lacBuffer[ii]);
    std::printf("\n");

    // 2. the archive: Load + ExtractFiles with the reference's signatures and error convention
    CArk lArk;
    eError leError = lArk.Load(argv[1]);
    SHOW_ERROR_AND_RETURN;
    std::printf("files: %d, has first: %d\n", lArk.GetNumFiles(),
                lArk.GetNumFiles() ? (int)lArk.FileExists(lArk.Header().maFiles[0].mName.c_str()) : 0);
    if (lArk.Load(argv[1]) != eError_AlreadyLoaded)  // reference CArk.cpp:303-306
        return 3;
    leError = lArk.ExtractFiles(0, lArk.GetNumFiles(), argv[2]);
    SHOW_ERROR_AND_RETURN;
    return eError_NoError;
}
