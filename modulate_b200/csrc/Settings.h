// Settings.h -- global flags and platform constants, same names as the reference's CSettings
// (Settings.h:5-23, Settings.cpp:4-9) because callers toggle them directly (Modulate.cpp:45-70).
#pragma once

#include <iostream>
#include <string>

class CSettings
{
public:
    static bool mbPS4;
    static const char* msPlatform;

    static bool mbVerbose;
    static bool mbOverwriteOutputFiles;
    static bool mbIgnoreNewFiles;
    static bool mbPackAllFiles;

    static const unsigned int kuEncryptedVersionPS3 = 0xc64eed30;
    static const unsigned int kuEncryptedVersionPS4 = 0x6f303f55;

    static const unsigned int kuEncryptedPS3Key = 0xc64eed30;
    static const unsigned int kuEncryptedPS4Key = 0x90cfc0ab;
};

#define VERBOSE_OUT(out) \
    if (CSettings::mbVerbose) std::cout << out
