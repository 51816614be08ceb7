// Settings.h -- process-wide switches and per-platform constants.
//
// Callers of the archive classes flip these directly (the reference's command handlers do,
// Modulate.cpp:45-70, :570-586), so the class name and member names are part of the drop-in
// boundary and match the reference's CSettings (Settings.h:5-23).  Header-only here: C++17
// inline statics carry the reference's defaults (Settings.cpp:4-9), so there is no Settings.cpp.
#pragma once

#include <iostream>

class CSettings
{
public:
    // ---- platform (-ps3 switches all three) ---------------------------------------------------
    inline static bool mbPS4 = true;                 // false: PS3 magic / key / entry marker / entry order
    inline static const char* msPlatform = "ps4";    // used in file names: main_<platform>.hdr

    // ---- behaviour flags ------------------------------------------------------------------------
    inline static bool mbVerbose = false;                // -verbose
    inline static bool mbOverwriteOutputFiles = true;    // -force (already the default, as in the reference)
    inline static bool mbIgnoreNewFiles = true;          // cleared by -pack_add
    inline static bool mbPackAllFiles = false;           // -packall

    // ---- header magic (first four bytes of the .hdr, never ciphered) ---------------------------
    static constexpr unsigned int kuEncryptedVersionPS3 = 0xc64eed30u;
    static constexpr unsigned int kuEncryptedVersionPS4 = 0x6f303f55u;

    // ---- stream-cipher keys for everything after the magic (both negative as `int`) ------------
    static constexpr unsigned int kuEncryptedPS3Key = 0xc64eed30u;
    static constexpr unsigned int kuEncryptedPS4Key = 0x90cfc0abu;
};

// Same macro name as the reference (Settings.h:23): stream to stdout only in verbose mode.
#define VERBOSE_OUT(out)             \
    do {                             \
        if (CSettings::mbVerbose)    \
            std::cout << out;        \
    } while (0)
