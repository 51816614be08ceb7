// oracle/ref_ark/ark_ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" handles onto the reference's OWN CArk / CDtaFile implementation
// (/root/reference/Modulate/CArk.cpp, CDtaFile.cpp, Utils.cpp, Settings.cpp, CEncryptionCycler.cpp),
// staged and compiled by `make -C oracle ark_ref` into oracle/_ref/libark_ref.so.  The shim holds no
// archive logic of its own: it constructs the reference's objects, calls their public methods
// (Load, ExtractFiles, ConstructFromDirectory, BuildArk, SaveArk; CDtaFile::Load, Save) and reads
// the tables they built.  tests/golden/make_ark_golden.py drives it to produce the reference-written
// .hdr / .ark / .dtb fixtures the product's header codec, part split and DTB codec are tested against.
#define private public  // read-only access to the file / part tables (CArk.h:76-89); this TU only
#include "CArk.h"
#undef private

#include <unistd.h>

#include "CDtaFile.h"
#include "Error.h"
#include "Settings.h"

extern "C" {

int ark_ref_chdir(const char* lpDirectory) { return chdir(lpDirectory); }

// -ps3 / default ps4 (Modulate.cpp:570-578); -force, -pack_add, -packall, -verbose (Modulate.cpp:45-70, 580-586)
void ark_ref_settings(int lbPS4, int lbOverwrite, int lbIgnoreNewFiles, int lbPackAllFiles, int lbVerbose)
{
    CSettings::mbPS4 = lbPS4 != 0;
    CSettings::msPlatform = lbPS4 ? "ps4" : "ps3";
    CSettings::mbOverwriteOutputFiles = lbOverwrite != 0;
    CSettings::mbIgnoreNewFiles = lbIgnoreNewFiles != 0;
    CSettings::mbPackAllFiles = lbPackAllFiles != 0;
    CSettings::mbVerbose = lbVerbose != 0;
}

void* ark_ref_new() { return new CArk(); }
void ark_ref_delete(void* lpArk) { delete (CArk*)lpArk; }

int ark_ref_load(void* lpArk, const char* lpHeaderFilename) { return (int)((CArk*)lpArk)->Load(lpHeaderFilename); }
int ark_ref_extract(void* lpArk, const char* lpTargetDirectory)
{
    CArk* lpRef = (CArk*)lpArk;
    return (int)lpRef->ExtractFiles(0, lpRef->GetNumFiles(), lpTargetDirectory);
}
int ark_ref_construct(void* lpArk, const char* lpInputDirectory, void* lpReferenceHeader)
{
    return (int)((CArk*)lpArk)->ConstructFromDirectory(lpInputDirectory, *(const CArk*)lpReferenceHeader, {});
}
// Modulate.cpp:410-432 (Pack without -packall): the song list comes from the two DTA configs.
static eError LoadSongs(const char* lpAmpConfig, const char* lpAmpSongsConfig, std::vector<SSongConfig>& laSongs)
{
    CDtaFile lAmpConfig;
    eError leError = lAmpConfig.Load(lpAmpConfig);
    if (leError != eError_NoError)
        return leError;
    laSongs = lAmpConfig.GetSongs();
    CDtaFile lSongsConfig;
    leError = lSongsConfig.Load(lpAmpSongsConfig);
    if (leError != eError_NoError)
        return leError;
    lSongsConfig.GetSongData(laSongs);
    return eError_NoError;
}

int ark_ref_construct_with_songs(void* lpArk, const char* lpInputDirectory, void* lpReferenceHeader,
                                 const char* lpAmpConfig, const char* lpAmpSongsConfig)
{
    std::vector<SSongConfig> laSongs;
    eError leError = LoadSongs(lpAmpConfig, lpAmpSongsConfig, laSongs);
    if (leError != eError_NoError)
        return (int)leError;
    return (int)((CArk*)lpArk)->ConstructFromDirectory(lpInputDirectory, *(const CArk*)lpReferenceHeader, laSongs);
}

// One line per song: id, name, unlock method, unlock count, path, arena, type (tab separated).
int dta_ref_songs(const char* lpAmpConfig, const char* lpAmpSongsConfig, char* lpOut, int liCapacity)
{
    std::vector<SSongConfig> laSongs;
    eError leError = LoadSongs(lpAmpConfig, lpAmpSongsConfig, laSongs);
    if (leError != eError_NoError)
        return -(int)leError;
    std::string lText;
    for (const SSongConfig& lSong : laSongs)
        lText += lSong.mId + "\t" + lSong.mName + "\t" + lSong.mUnlockMethod + "\t" + std::to_string(lSong.miUnlockCount) + "\t" +
                 lSong.mPath + "\t" + lSong.mArena + "\t" + lSong.mType + "\n";
    snprintf(lpOut, (size_t)liCapacity, "%s", lText.c_str());
    return (int)laSongs.size();
}

int ark_ref_build(void* lpArk, const char* lpInputDirectory) { return (int)((CArk*)lpArk)->BuildArk(lpInputDirectory, {}); }
int ark_ref_save(void* lpArk, const char* lpOutputDirectory, const char* lpHeaderFilename)
{
    return (int)((const CArk*)lpArk)->SaveArk(lpOutputDirectory, lpHeaderFilename);
}
int ark_ref_file_exists(void* lpArk, const char* lpName) { return ((const CArk*)lpArk)->FileExists(lpName) ? 1 : 0; }

int ark_ref_num_files(void* lpArk) { return ((const CArk*)lpArk)->miNumFiles; }
int ark_ref_num_arks(void* lpArk) { return ((const CArk*)lpArk)->miNumArks; }

int ark_ref_file(void* lpArk, int liIndex, char* lpName, int liNameCapacity, long long* lpOffset, int* lpSize,
                 int* lpFlags1, int* lpFlags2, int* lpHash)
{
    const CArk* lpRef = (const CArk*)lpArk;
    if (liIndex < 0 || liIndex >= lpRef->miNumFiles)
        return -1;
    const CArk::sFileDefinition& lFile = lpRef->mpFiles[liIndex];
    snprintf(lpName, (size_t)liNameCapacity, "%s", lFile.mName.c_str());
    *lpOffset = lFile.mi64Offset;
    *lpSize = lFile.miSize;
    *lpFlags1 = lFile.miFlags1;
    *lpFlags2 = lFile.miFlags2;
    *lpHash = lFile.miHash;
    return 0;
}

int ark_ref_part(void* lpArk, int liIndex, char* lpPath, int liPathCapacity, unsigned int* lpSize)
{
    const CArk* lpRef = (const CArk*)lpArk;
    if (liIndex < 0 || liIndex >= lpRef->miNumArks)
        return -1;
    snprintf(lpPath, (size_t)liPathCapacity, "%s", lpRef->mpArks[liIndex].mPath.c_str());
    *lpSize = lpRef->mpArks[liIndex].muSize;
    return 0;
}

// CDtaFile::Load then CDtaFile::Save (CDtaFile.cpp:57-100, :362-391)
int dta_ref_roundtrip(const char* lpInputFilename, const char* lpOutputFilename)
{
    CDtaFile lFile;
    eError leError = lFile.Load(lpInputFilename);
    if (leError != eError_NoError)
        return (int)leError;
    return (int)lFile.Save(lpOutputFilename);
}

}  // extern "C"
