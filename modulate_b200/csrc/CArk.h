// CArk.h -- drop-in for the archive-container class of the reference (/root/reference/Modulate/
// CArk.h:10-90): same public method names, argument meaning and eError convention, portable
// C++17 (the reference is Win32/MSVC-only), with the data-parallel half moved to the GPU:
//
//   Load                  reads the .hdr, deciphers it on the GPU (mod_cycle; reference call site
//                         CArk.cpp:338-339) and parses it on the host (ArkHeader.cpp)
//   LoadArkData           concatenates the .ark parts into one PINNED host image (CArk.cpp:723-758)
//   ExtractFiles          turns the file table into mod_desc descriptors and gathers the entries
//                         through the batched kernel (mod_cycle_batch; CArk.cpp:494) in a
//                         reader -> GPU -> writer pipeline over a ring of pinned slots: part files
//                         are still being read while earlier groups are on the GPU / being written
//   BuildArk              byte-packs the input files into the image, assigns offsets and part sizes
//                         (CArk.cpp:760-828) and, when entries carry keys, ciphers them in one batch
//   SaveArk               serialises + enciphers the header (CArk.cpp:1135-1136) and writes the parts
//
// The reference never ciphers ARK bodies (SURVEY.md Finding 2): all entry keys default to 0, the
// identity keystream, so the bytes produced are the reference's.  SetEntryKeys / SetUniformEntryKey
// are extensions used by the synthetic per-entry-key configurations of BASELINE.json.
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "ArkHeader.h"
#include "CDtaFile.h"  // SSongConfig, like the reference (CArk.h:6)
#include "Error.h"

class CArk
{
public:
    CArk();
    ~CArk();
    CArk(const CArk&) = delete;
    CArk& operator=(const CArk&) = delete;

    eError ConstructFromDirectory(const char* lpInputDirectory, const CArk& lReferenceHeader,
                                  std::vector<SSongConfig> laSongs);
    eError BuildArk(const char* lpInputDirectory, std::vector<SSongConfig> laSongs);
    eError SaveArk(const char* lpOutputDirectory, const char* lpHeaderFilename) const;

    eError Load(const char* lpHeaderFilename);
    eError ExtractFiles(int liFirstFileIndex, int liNumFiles, const char* lpTargetDirectory);
    bool FileExists(const char* lpFilename) const;

    int GetNumFiles() const;

    eError LoadArkData();

    // ---- extensions (not in the reference) ----
    void SetUniformEntryKey(int liKey);                  // every entry ciphered with liKey, stream restarting per entry
    void SetEntryKeys(const std::vector<int>& laKeys);   // one key per entry, in table order
    void SetPartDirectory(const char* lpDirectory);      // where LoadArkData looks for the part files ("" = cwd)
    const modark::HeaderImage& Header() const { return mHeader; }

private:
    int EntryKey(size_t liIndex) const;
    eError AllocateArkData();
    eError ReadParts();
    bool ReadImageRange(std::vector<FILE*>& lFiles, uint64_t luOffset, uint64_t luSize, unsigned char* lpDst) const;
    bool ShouldPackFile(const std::vector<SSongConfig>& laSongs, const char* lpFilename) const;
    void ReleaseArkData();

    modark::HeaderImage mHeader;
    bool mbLoaded = false;

    unsigned char* mpArkData = nullptr;  // pinned host memory (mod_host_alloc)
    uint64_t muArkDataSize = 0;

    std::vector<int> maEntryKeys;
    int miUniformKey = 0;
    std::string mPartDirectory;
};
