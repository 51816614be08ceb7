// ArkHeader.h -- codec for the decrypted `.hdr` image that sits either side of the cipher.
//
// Layout (little-endian), as read by the reference's CArk::Load (CArk.cpp:341-416) and written by
// SaveArk's lSaveHeader (CArk.cpp:911-1133):
//   u32 magic (never ciphered)                      PS3 0xc64eed30 / PS4 0x6f303f55 (Settings.h:16-17)
//   u32 version = 9, u32 numChecksums = 1, char[16] checksum, i32 numArks        (CArk.h:27-32)
//   intlist  ark sizes   {i32 n, i32[n]}
//   strlist  ark paths   {i32 n, {i32 len, bytes}[n]}
//   intlist  checksums   {i32 n, i32[n]}   then  intlist string counts {i32 n, i32[n]}
//     (Load skips the first list, then n more words, then requires a zero word: CArk.cpp:379-390)
//   i32 numFiles, entries {i64 offset, i32 nameLen, name, i32 flags1, u32 size, u32 marker}
//   intlist  flags2      {i32 n, i32[n]}   (hash bucket -> last entry index of its chain)
// This is host-side, KB-sized, branchy work and stays on the CPU by design; only the Cycle() over
// the image and the bulk payload movement go to the GPU.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "Error.h"

namespace modark {

struct PartDef {
    std::string mPath;
    unsigned int muSize = 0;
};

struct FileDef {
    std::string mName;
    int64_t mi64Offset = 0;  // into the concatenation of all parts
    int miSize = 0;
    int miFlags1 = -1;
    int miFlags2 = -1;
    unsigned int muHash = 0;
};

struct HeaderImage {
    bool mbPS4 = true;
    std::vector<PartDef> maParts;
    std::vector<FileDef> maFiles;
};

constexpr int kMaxArks = 100;      // CArk.cpp:345
constexpr int kMaxFiles = 25000;   // CArk.cpp:395

// Parse a DECRYPTED header image (magic included).  Every read is bounds-checked: a truncated or
// corrupt image yields eError_InvalidData instead of the reference's out-of-bounds read.
eError ParseHeader(const unsigned char* lpData, size_t liSize, HeaderImage& lOut);

// Name-hash bucket used to order entries and thread the flags1 / flags2 chains on save.
int NameBucket(const std::string& lName, int liNumFiles);

// Plaintext header image for `lHeader` (entry offsets / sizes / names as given; flags1 and the
// flags2 table are recomputed).  The 16 checksum bytes are written as zeros.
std::vector<unsigned char> SerialiseHeader(const HeaderImage& lHeader);

}  // namespace modark
