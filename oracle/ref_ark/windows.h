// oracle/ref_ark/windows.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Stand-in for <windows.h> (and, through direct.h, <direct.h>) so that the reference's Win32/MSVC
// sources CArk.cpp, CDtaFile.cpp and Utils.cpp compile with g++ for the parity oracle
// (oracle/Makefile target `ark_ref`).  It supplies only what those files use (CArk.cpp:2,8,124,140,262,
// 312,470,553; CDtaFile.cpp:60,605,829; Utils.cpp:12-67) with the documented CRT / Win32 semantics.
// The one thing here that is a MODEL rather than a definition is the order FindFirstFileA /
// FindNextFileA enumerate a directory in: NTFS returns names in its index order (upper-cased
// code-point order); the emulation sorts the same way.
#pragma once

#include <dirent.h>
#include <strings.h>
#include <sys/stat.h>
#include <sys/types.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define __int64 long long
#define _int64 long long  // CDtaFile.cpp:843

inline int fopen_s(FILE** lpFile, const char* lpName, const char* lpMode)
{
    *lpFile = std::fopen(lpName, lpMode);
    return *lpFile ? 0 : errno;
}

inline int _mkdir(const char* lpPath) { return mkdir(lpPath, 0777); }

inline int memcpy_s(void* lpDst, size_t liDstSize, const void* lpSrc, size_t liCount)
{
    if (liCount > liDstSize)  // the CRT zeroes the destination and raises the invalid-parameter handler
        std::abort();
    std::memcpy(lpDst, lpSrc, liCount);
    return 0;
}

inline int _stricmp(const char* lpA, const char* lpB) { return strcasecmp(lpA, lpB); }

template <size_t N>
inline int _itoa_s(int liValue, char (&lacBuffer)[N], int liRadix)
{
    std::snprintf(lacBuffer, N, liRadix == 16 ? "%x" : "%d", liValue);
    return 0;
}

// ---- FindFirstFileA / FindNextFileA over opendir --------------------------------------------------

#define FILE_ATTRIBUTE_DIRECTORY 0x10u
#define MAX_PATH 260

struct WIN32_FIND_DATAA {
    unsigned int dwFileAttributes;
    char cFileName[MAX_PATH];
};

struct ModRefFindState {
    std::vector<std::pair<std::string, bool>> maEntries;  // name, is directory
    size_t miNext = 0;
};
typedef ModRefFindState* HANDLE;
#define INVALID_HANDLE_VALUE ((HANDLE) nullptr)

inline int FindNextFileA(HANDLE lpFind, WIN32_FIND_DATAA* lpData)
{
    if (!lpFind || lpFind->miNext >= lpFind->maEntries.size())
        return 0;
    const auto& lEntry = lpFind->maEntries[lpFind->miNext++];
    lpData->dwFileAttributes = lEntry.second ? FILE_ATTRIBUTE_DIRECTORY : 0x80u /* FILE_ATTRIBUTE_NORMAL */;
    std::snprintf(lpData->cFileName, MAX_PATH, "%s", lEntry.first.c_str());
    return 1;
}

// lpPattern is "<directory>*.*" (Utils.cpp:9-10): every entry of <directory>, "." and ".." included.
inline HANDLE FindFirstFileA(const char* lpPattern, WIN32_FIND_DATAA* lpData)
{
    std::string lDirectory = lpPattern;
    if (lDirectory.size() >= 3 && lDirectory.compare(lDirectory.size() - 3, 3, "*.*") == 0)
        lDirectory.resize(lDirectory.size() - 3);
    // failure leaves something both loops of GenerateFileList skip
    lpData->dwFileAttributes = FILE_ATTRIBUTE_DIRECTORY;
    std::snprintf(lpData->cFileName, MAX_PATH, ".");
    DIR* lpDir = opendir(lDirectory.empty() ? "." : lDirectory.c_str());
    if (!lpDir)
        return INVALID_HANDLE_VALUE;
    ModRefFindState* lpFind = new ModRefFindState();
    while (dirent* lpEntry = readdir(lpDir)) {
        struct stat lInfo;
        const std::string lFull = lDirectory + lpEntry->d_name;
        const bool lbDirectory = stat(lFull.c_str(), &lInfo) == 0 && S_ISDIR(lInfo.st_mode);
        lpFind->maEntries.emplace_back(lpEntry->d_name, lbDirectory);
    }
    closedir(lpDir);
    auto lUpper = [](const std::string& lName) {
        std::string lOut = lName;
        for (char& c : lOut)
            c = (char)std::toupper((unsigned char)c);
        return lOut;
    };
    std::sort(lpFind->maEntries.begin(), lpFind->maEntries.end(),
              [&](const auto& lA, const auto& lB) { return lUpper(lA.first) < lUpper(lB.first); });
    if (!FindNextFileA(lpFind, lpData)) {
        delete lpFind;
        return INVALID_HANDLE_VALUE;
    }
    return lpFind;
}
