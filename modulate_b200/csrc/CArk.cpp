#include "CArk.h"

#include <dirent.h>
#include <sys/stat.h>
#include <sys/types.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "../../include/modulate_b200.h"
#include "CEncryptionCycler.h"
#include "Settings.h"

namespace {

double NowSeconds()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool TraceEnabled()
{
    const char* lpValue = std::getenv("MOD_TRACE");
    return lpValue && *lpValue && *lpValue != '0';
}

bool ReadWholeFile(const std::string& lPath, std::vector<unsigned char>& lOut)
{
    FILE* lpFile = std::fopen(lPath.c_str(), "rb");
    if (!lpFile)
        return false;
    std::fseek(lpFile, 0, SEEK_END);
    const long liSize = std::ftell(lpFile);
    std::fseek(lpFile, 0, SEEK_SET);
    lOut.resize(liSize > 0 ? (size_t)liSize : 0);
    const size_t liRead = lOut.empty() ? 0 : std::fread(lOut.data(), 1, lOut.size(), lpFile);
    std::fclose(lpFile);
    return liRead == lOut.size();
}

long FileSize(const std::string& lPath)
{
    struct stat lInfo;
    if (stat(lPath.c_str(), &lInfo) != 0 || !S_ISREG(lInfo.st_mode))
        return -1;
    return (long)lInfo.st_size;
}

bool IsDirectory(const std::string& lPath)
{
    struct stat lInfo;
    return stat(lPath.c_str(), &lInfo) == 0 && S_ISDIR(lInfo.st_mode);
}

// mkdir -p for every directory component of lFilePath (the part before the last '/').
bool MakeParentDirectories(const std::string& lFilePath)
{
    for (size_t liSlash = lFilePath.find('/'); liSlash != std::string::npos; liSlash = lFilePath.find('/', liSlash + 1)) {
        if (liSlash == 0)
            continue;
        const std::string lDir = lFilePath.substr(0, liSlash);
        mkdir(lDir.c_str(), 0777);
        if (!IsDirectory(lDir))
            return false;
    }
    return true;
}

// An entry name may not climb out of the target directory ("..": the reference would follow it).
bool EscapesTarget(const std::string& lName)
{
    size_t liStart = 0;
    for (;;) {
        const size_t liSlash = lName.find('/', liStart);
        const std::string lPart = lName.substr(liStart, liSlash == std::string::npos ? std::string::npos : liSlash - liStart);
        if (lPart == "..")
            return true;
        if (liSlash == std::string::npos)
            return false;
        liStart = liSlash + 1;
    }
}

// Existing non-empty output is kept only when overwriting is disabled (reference CArk.cpp:439-454).
bool KeepExistingOutput(const std::string& lPath)
{
    if (CSettings::mbOverwriteOutputFiles)
        return false;
    return FileSize(lPath) > 0;
}

// Relative names of every regular file under lRoot (which ends in '/'), files of a directory first,
// then its sub-directories, each level in the order NTFS keeps its directory index in: by UPPER-CASED
// name (the reference walks with FindFirstFileA, Utils.cpp:5-70, and takes whatever order the file
// system returns; '_' therefore sorts after the letters, not before them).
void ListFiles(const std::string& lRoot, const std::string& lRelative, std::vector<std::string>& laOut)
{
    DIR* lpDir = opendir((lRoot + lRelative).c_str());
    if (!lpDir)
        return;
    std::vector<std::string> laFiles, laDirs;
    while (dirent* lpEntry = readdir(lpDir)) {
        const std::string lName = lpEntry->d_name;
        if (lName == "." || lName == "..")
            continue;
        const std::string lFull = lRoot + lRelative + lName;
        if (IsDirectory(lFull))
            laDirs.push_back(lName);
        else if (FileSize(lFull) >= 0)
            laFiles.push_back(lName);
    }
    closedir(lpDir);
    auto lLess = [](const std::string& lA, const std::string& lB) {
        return std::lexicographical_compare(lA.begin(), lA.end(), lB.begin(), lB.end(), [](char a, char b) {
            return std::toupper((unsigned char)a) < std::toupper((unsigned char)b);
        });
    };
    std::sort(laFiles.begin(), laFiles.end(), lLess);
    std::sort(laDirs.begin(), laDirs.end(), lLess);
    for (const std::string& lName : laFiles)
        laOut.push_back(lRelative + lName);
    for (const std::string& lName : laDirs)
        ListFiles(lRoot, lRelative + lName + "/", laOut);
}

}  // namespace

CArk::CArk() {}

CArk::~CArk() { ReleaseArkData(); }

void CArk::ReleaseArkData()
{
    if (mpArkData)
        mod_host_free(mpArkData);
    mpArkData = nullptr;
    muArkDataSize = 0;
}

void CArk::SetUniformEntryKey(int liKey)
{
    miUniformKey = liKey;
    maEntryKeys.clear();
}

void CArk::SetEntryKeys(const std::vector<int>& laKeys) { maEntryKeys = laKeys; }

void CArk::SetPartDirectory(const char* lpDirectory)
{
    mPartDirectory = lpDirectory ? lpDirectory : "";
    if (!mPartDirectory.empty() && mPartDirectory.back() != '/')
        mPartDirectory += '/';
}

int CArk::EntryKey(size_t liIndex) const { return liIndex < maEntryKeys.size() ? maEntryKeys[liIndex] : miUniformKey; }

int CArk::GetNumFiles() const { return (int)mHeader.maFiles.size(); }

bool CArk::FileExists(const char* lpFilename) const
{
    for (const modark::FileDef& lFile : mHeader.maFiles)
        if (lFile.mName == lpFilename)
            return true;
    return false;
}

// ---- read side ---------------------------------------------------------------------------------------

eError CArk::Load(const char* lpHeaderFilename)
{
    if (mbLoaded)
        return eError_AlreadyLoaded;  // reference CArk.cpp:303-306

    VERBOSE_OUT("Loading header file " << lpHeaderFilename);
    std::vector<unsigned char> lData;
    if (!ReadWholeFile(lpHeaderFilename, lData))
        return eError_FailedToOpenFile;
    VERBOSE_OUT("\nLoaded header (" << lData.size() << ") bytes\n");
    if (lData.size() < sizeof(uint32_t))
        return eError_UnknownVersionNumber;

    uint32_t luVersion = 0;
    std::memcpy(&luVersion, lData.data(), sizeof(luVersion));
    if (luVersion != CSettings::kuEncryptedVersionPS3 && luVersion != CSettings::kuEncryptedVersionPS4)
        return eError_UnknownVersionNumber;
    const unsigned int kuInitialKey =
        (luVersion == CSettings::kuEncryptedVersionPS3) ? CSettings::kuEncryptedPS3Key : CSettings::kuEncryptedPS4Key;

    // the magic stays in clear; everything after it is one Cycle() -- on the GPU
    CEncryptionCycler lDecrypt;
    lDecrypt.Cycle(lData.data() + sizeof(uint32_t), (unsigned int)(lData.size() - sizeof(uint32_t)), (int)kuInitialKey);

    eError leError = modark::ParseHeader(lData.data(), lData.size(), mHeader);
    SHOW_ERROR_AND_RETURN;
    mbLoaded = true;
    return eError_NoError;
}

// Read the parts back to back into the pinned image (reference CArk.cpp:741-755).
eError CArk::ReadParts()
{
    std::vector<FILE*> lFiles(mHeader.maParts.size(), nullptr);
    const bool lbOk = ReadImageRange(lFiles, 0, muArkDataSize, mpArkData);
    for (FILE* lpFile : lFiles)
        if (lpFile)
            std::fclose(lpFile);
    return lbOk ? eError_NoError : eError_FailedToOpenFile;
}

eError CArk::AllocateArkData()
{
    ReleaseArkData();
    uint64_t luTotalArkSize = 0;
    for (const modark::PartDef& lPart : mHeader.maParts)
        luTotalArkSize += lPart.muSize;
    mpArkData = (unsigned char*)mod_host_alloc(luTotalArkSize ? luTotalArkSize : 1);
    if (!mpArkData) {
        std::cout << "Failed to allocate pinned memory for the archive: " << mod_last_error() << "\n";
        return eError_NoData;
    }
    muArkDataSize = luTotalArkSize;
    return eError_NoError;
}

eError CArk::LoadArkData()
{
    eError leError = AllocateArkData();
    ERROR_RETURN;
    leError = ReadParts();
    SHOW_ERROR_AND_RETURN;
    return eError_NoError;
}

// Copy bytes [luOffset, luOffset + luSize) of the concatenated part files into lpDst, opening each
// part on demand (lFiles caches the handles).  Bytes a short part file does not have read as zero.
bool CArk::ReadImageRange(std::vector<FILE*>& lFiles, uint64_t luOffset, uint64_t luSize, unsigned char* lpDst) const
{
    uint64_t luPartStart = 0;
    for (size_t ii = 0; ii < mHeader.maParts.size() && luSize; ++ii) {
        const uint64_t luPartSize = mHeader.maParts[ii].muSize;
        const uint64_t luPartEnd = luPartStart + luPartSize;
        if (luOffset < luPartEnd) {
            if (!lFiles[ii]) {
                lFiles[ii] = std::fopen((mPartDirectory + mHeader.maParts[ii].mPath).c_str(), "rb");
                if (!lFiles[ii])
                    return false;
            }
            const uint64_t luTake = std::min(luSize, luPartEnd - luOffset);
            if (fseeko(lFiles[ii], (off_t)(luOffset - luPartStart), SEEK_SET) != 0)
                return false;
            const size_t liRead = std::fread(lpDst, 1, (size_t)luTake, lFiles[ii]);
            if (liRead != luTake)
                std::memset(lpDst + liRead, 0, (size_t)luTake - liRead);
            lpDst += luTake;
            luOffset += luTake;
            luSize -= luTake;
        }
        luPartStart = luPartEnd;
    }
    return luSize == 0;
}

// Extraction is a three-stage host pipeline around the GPU batch, on a small ring of pinned slots
// instead of the reference's one archive-sized buffer (CArk.cpp:431, :738) and its serial
// one-file-at-a-time writes (:435-501):
//   reader thread   the image range a group of entries spans: part files -> the slot's source buffer
//   this thread     ONE mod_cycle_batch per group (itself an overlapped upload / kernel / download)
//                   from the slot's source buffer into its byte-packed staging buffer
//   writer threads  create directories and write the group's files, then hand the slot back
// so disk reads, PCIe, the kernel and disk writes overlap, and pinned memory is O(slots), not
// O(archive).
eError CArk::ExtractFiles(int liFirstFileIndex, int liNumFiles, const char* lpTargetDirectory)
{
    (void)liFirstFileIndex;  // the reference ignores both and walks the whole table (CArk.cpp:435)
    (void)liNumFiles;
    if (mHeader.maFiles.empty())
        return eError_NoData;
    const double ldStart = NowSeconds();

    uint64_t luImageSize = 0;
    for (const modark::PartDef& lPart : mHeader.maParts)
        luImageSize += lPart.muSize;

    // validate, then order the entries by where they sit in the image
    const size_t liCount = mHeader.maFiles.size();
    std::vector<uint32_t> laOrder(liCount);
    for (size_t ii = 0; ii < liCount; ++ii) {
        const modark::FileDef& lFile = mHeader.maFiles[ii];
        if ((uint64_t)lFile.mi64Offset > luImageSize || (uint64_t)lFile.miSize > luImageSize - (uint64_t)lFile.mi64Offset) {
            std::cout << "Entry " << lFile.mName.c_str() << " lies outside the archive data\n";
            eError leError = eError_InvalidData;
            SHOW_ERROR_AND_RETURN;
        }
        laOrder[ii] = (uint32_t)ii;
    }
    std::stable_sort(laOrder.begin(), laOrder.end(), [&](uint32_t a, uint32_t b) {
        return mHeader.maFiles[a].mi64Offset < mHeader.maFiles[b].mi64Offset;
    });

    // groups of consecutive entries (~32 MiB of payload each) and the image range each one spans
    struct Group {
        size_t first, last;           // positions in laOrder
        uint64_t srcLo, srcHi;        // image range
        uint64_t payload;
    };
    const uint64_t kuGroupBytes = 32ull << 20;
    std::vector<Group> laGroups;
    uint64_t luMaxRange = 1, luMaxPayload = 1;
    for (size_t liFirst = 0; liFirst < liCount;) {
        Group lGroup{liFirst, liFirst, UINT64_MAX, 0, 0};
        while (lGroup.last < liCount && (lGroup.payload < kuGroupBytes || lGroup.last == liFirst)) {
            const modark::FileDef& lFile = mHeader.maFiles[laOrder[lGroup.last]];
            if (lFile.miSize) {
                lGroup.srcLo = std::min(lGroup.srcLo, (uint64_t)lFile.mi64Offset);
                lGroup.srcHi = std::max(lGroup.srcHi, (uint64_t)lFile.mi64Offset + (uint64_t)lFile.miSize);
            }
            lGroup.payload += (uint64_t)lFile.miSize;
            ++lGroup.last;
        }
        if (lGroup.srcLo == UINT64_MAX)
            lGroup.srcLo = lGroup.srcHi = 0;
        luMaxRange = std::max(luMaxRange, lGroup.srcHi - lGroup.srcLo);
        luMaxPayload = std::max(luMaxPayload, lGroup.payload);
        laGroups.push_back(lGroup);
        liFirst = lGroup.last;
    }

    // the ring of pinned slots
    constexpr int kiSlots = 3;
    enum eSlotState { eSlot_Free, eSlot_Filled, eSlot_Busy };
    struct Slot {
        unsigned char* mpSource = nullptr;
        unsigned char* mpStaging = nullptr;
        eSlotState meState = eSlot_Free;
    };
    Slot laSlots[kiSlots];
    const int liSlotsUsed = (int)std::min<size_t>(kiSlots, laGroups.size());
    for (int ii = 0; ii < liSlotsUsed; ++ii) {
        laSlots[ii].mpSource = (unsigned char*)mod_host_alloc(luMaxRange);
        laSlots[ii].mpStaging = (unsigned char*)mod_host_alloc(luMaxPayload);
    }
    auto lFreeSlots = [&]() {
        for (Slot& lSlot : laSlots) {
            mod_host_free(lSlot.mpSource);
            mod_host_free(lSlot.mpStaging);
        }
    };
    for (int ii = 0; ii < liSlotsUsed; ++ii) {
        if (!laSlots[ii].mpSource || !laSlots[ii].mpStaging) {
            std::cout << "Failed to allocate pinned staging memory: " << mod_last_error() << "\n";
            lFreeSlots();
            return eError_NoData;
        }
    }
    const double ldAllocated = NowSeconds();

    std::mutex lMutex;
    std::condition_variable lSignal;
    std::atomic<int> liError{(int)eError_NoError};
    auto lFail = [&](eError leWhat) {
        int liExpected = (int)eError_NoError;
        liError.compare_exchange_strong(liExpected, (int)leWhat);
        lSignal.notify_all();
    };
    auto lWaitFor = [&](Slot& lSlot, eSlotState leWanted) {
        std::unique_lock<std::mutex> lLock(lMutex);
        lSignal.wait(lLock, [&]() { return lSlot.meState == leWanted || liError.load() != (int)eError_NoError; });
        return liError.load() == (int)eError_NoError;
    };
    auto lSetState = [&](Slot& lSlot, eSlotState leState) {
        {
            std::lock_guard<std::mutex> lLock(lMutex);
            lSlot.meState = leState;
        }
        lSignal.notify_all();
    };

    // stage 1: reader
    std::thread lReader([&]() {
        std::vector<FILE*> lFiles(mHeader.maParts.size(), nullptr);
        for (size_t gg = 0; gg < laGroups.size(); ++gg) {
            Slot& lSlot = laSlots[gg % kiSlots];
            if (!lWaitFor(lSlot, eSlot_Free))
                break;
            const Group& lGroup = laGroups[gg];
            if (!ReadImageRange(lFiles, lGroup.srcLo, lGroup.srcHi - lGroup.srcLo, lSlot.mpSource)) {
                lFail(eError_FailedToOpenFile);
                break;
            }
            lSetState(lSlot, eSlot_Filled);
        }
        for (FILE* lpFile : lFiles)
            if (lpFile)
                std::fclose(lpFile);
    });

    // stage 3: writers
    struct Job {
        size_t group;
        std::vector<uint64_t> offsets;  // staging offset of every entry of the group
    };
    std::deque<Job> lJobs;
    bool lbNoMoreJobs = false;
    const std::string lTarget = lpTargetDirectory;
    auto lWriteGroup = [&](const Job& lJob) {
        const Group& lGroup = laGroups[lJob.group];
        Slot& lSlot = laSlots[lJob.group % kiSlots];
        for (size_t ii = lGroup.first; ii < lGroup.last; ++ii) {
            const modark::FileDef& lFile = mHeader.maFiles[laOrder[ii]];
            const std::string lOutputPath = lTarget + lFile.mName;
            if (EscapesTarget(lFile.mName)) {
                std::cout << "Refusing to write outside the target directory: " << lFile.mName.c_str() << "\n";
                continue;
            }
            if (KeepExistingOutput(lOutputPath)) {
                VERBOSE_OUT("Output file already exists, skipping: " << lOutputPath.c_str() << "\n");
                continue;
            }
            if (!MakeParentDirectories(lOutputPath)) {
                lFail(eError_FailedToCreateDirectory);
                return;
            }
            FILE* lpOutputFile = std::fopen(lOutputPath.c_str(), "wb");
            if (!lpOutputFile) {
                std::cout << "Failed to create " << lOutputPath.c_str() << "\n";  // the reference carries on too (CArk.cpp:488-491)
                continue;
            }
            const size_t liWritten =
                lFile.miSize ? std::fwrite(lSlot.mpStaging + lJob.offsets[ii - lGroup.first], 1, (size_t)lFile.miSize, lpOutputFile) : 0;
            std::fclose(lpOutputFile);
            if (liWritten != (size_t)lFile.miSize) {
                lFail(eError_FailedToWriteData);
                return;
            }
        }
    };
    auto lWriterLoop = [&]() {
        for (;;) {
            Job lJob;
            {
                std::unique_lock<std::mutex> lLock(lMutex);
                lSignal.wait(lLock, [&]() { return !lJobs.empty() || lbNoMoreJobs; });
                if (lJobs.empty())
                    return;
                lJob = std::move(lJobs.front());
                lJobs.pop_front();
            }
            if (liError.load() == (int)eError_NoError)
                lWriteGroup(lJob);
            lSetState(laSlots[lJob.group % kiSlots], eSlot_Free);
        }
    };
    std::vector<std::thread> laWriters;
    for (int ii = 0; ii < 3; ++ii)
        laWriters.emplace_back(lWriterLoop);

    // stage 2: one GPU batch per group, file table -> device descriptors
    std::vector<mod_desc> laDescs;
    for (size_t gg = 0; gg < laGroups.size(); ++gg) {
        const Group& lGroup = laGroups[gg];
        Slot& lSlot = laSlots[gg % kiSlots];
        if (!lWaitFor(lSlot, eSlot_Filled))
            break;
        Job lJob;
        lJob.group = gg;
        laDescs.clear();
        uint64_t luOut = 0;
        for (size_t ii = lGroup.first; ii < lGroup.last; ++ii) {
            const modark::FileDef& lFile = mHeader.maFiles[laOrder[ii]];
            mod_desc lDesc;
            lDesc.src_off = lFile.miSize ? (uint64_t)lFile.mi64Offset - lGroup.srcLo : 0;
            lDesc.dst_off = luOut;
            lDesc.len = (uint32_t)lFile.miSize;
            lDesc.key = EntryKey(laOrder[ii]);
            laDescs.push_back(lDesc);
            lJob.offsets.push_back(luOut);
            luOut += (uint64_t)lFile.miSize;
        }
        if (luOut && mod_cycle_batch(laDescs.data(), laDescs.size(), lSlot.mpSource, lGroup.srcHi - lGroup.srcLo,
                                     lSlot.mpStaging, luOut) != MOD_OK) {
            std::cout << "GPU extract failed: " << mod_last_error() << "\n";
            lFail(eError_InvalidData);
            break;
        }
        {
            std::lock_guard<std::mutex> lLock(lMutex);
            lSlot.meState = eSlot_Busy;
            lJobs.push_back(std::move(lJob));
        }
        lSignal.notify_all();
    }
    {
        std::lock_guard<std::mutex> lLock(lMutex);
        lbNoMoreJobs = true;
    }
    lSignal.notify_all();
    const double ldGpuDone = NowSeconds();
    lReader.join();
    for (std::thread& lWriter : laWriters)
        lWriter.join();
    const double ldWritten = NowSeconds();
    lFreeSlots();
    if (TraceEnabled())
        std::fprintf(stderr, "[mod] ExtractFiles: %zu groups, pinned slots %.3f s, read+GPU %.3f s, writers drain %.3f s, free %.3f s\n",
                     laGroups.size(), ldAllocated - ldStart, ldGpuDone - ldAllocated, ldWritten - ldGpuDone,
                     NowSeconds() - ldWritten);

    eError leError = (eError)liError.load();
    SHOW_ERROR_AND_RETURN;
    return eError_NoError;
}

// ---- write side --------------------------------------------------------------------------------------

bool CArk::ShouldPackFile(const std::vector<SSongConfig>& laSongs, const char* lpFilename) const
{
    // Files under ".../songs/<name>/" are packed only when <name> occurs in some configured song
    // path; everything else always is (policy of the reference, CArk.cpp:58-92).
    const char* lpSongs = std::strstr(lpFilename, "/songs/");
    if (!lpSongs)
        return true;
    const char* lpEnd = lpSongs + std::strlen("/songs/");
    while (*lpEnd && *lpEnd != '/')
        ++lpEnd;
    if (!*lpEnd)
        return false;
    std::string lSongName(lpSongs, lpEnd - lpSongs);
    std::transform(lSongName.begin(), lSongName.end(), lSongName.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    for (const SSongConfig& lSong : laSongs)
        if (lSong.mPath.find(lSongName) != std::string::npos)
            return true;
    return false;
}

eError CArk::ConstructFromDirectory(const char* lpInputDirectory, const CArk& lReferenceHeader, std::vector<SSongConfig> laSongs)
{
    for (const char* lpAlways : {"/songs/credits", "/songs/tut0", "/songs/tut1", "/songs/tutc"}) {
        SSongConfig lSong;
        lSong.mPath = lpAlways;
        laSongs.push_back(lSong);
    }

    std::string lRoot = lpInputDirectory;
    if (!lRoot.empty() && lRoot.back() != '/')
        lRoot += '/';
    std::vector<std::string> laFilenames;
    ListFiles(lRoot, "", laFilenames);
    if (laFilenames.empty()) {
        eError leError = eError_NoData;
        SHOW_ERROR_AND_RETURN;
    }
    VERBOSE_OUT("Found " << laFilenames.size() << " files\n");

    mHeader = modark::HeaderImage();
    mHeader.mbPS4 = CSettings::mbPS4;
    uint64_t luTotalFileSize = 0;
    for (const std::string& lFilename : laFilenames) {
        const modark::FileDef* lpReference = nullptr;
        for (const modark::FileDef& lCandidate : lReferenceHeader.mHeader.maFiles) {
            if (lCandidate.mName == lFilename) {
                lpReference = &lCandidate;
                break;
            }
        }
        if (!lpReference && CSettings::mbIgnoreNewFiles)
            continue;  // only files the reference header knows are repacked (-pack); -pack_add lifts this
        if (!CSettings::mbPackAllFiles && !ShouldPackFile(laSongs, lFilename.c_str()))
            continue;
        const long liSize = FileSize(lRoot + lFilename);
        if (liSize < 0) {
            std::cout << "Unable to open file: " << lFilename.c_str() << "\n";
            continue;
        }
        modark::FileDef lFile = lpReference ? *lpReference : modark::FileDef();
        lFile.mName = lFilename;
        lFile.miSize = (int)liSize;
        luTotalFileSize += (uint64_t)liSize;
        mHeader.maFiles.push_back(lFile);
    }
    if (mHeader.maFiles.empty()) {
        eError leError = eError_NoData;
        SHOW_ERROR_AND_RETURN;
    }

    // part plan: the reference header's part list, each part allowed an equal share of what is left
    mHeader.maParts = lReferenceHeader.mHeader.maParts;
    uint64_t luSizeRemaining = luTotalFileSize;
    const size_t liNumArks = mHeader.maParts.size();
    for (size_t ii = 0; ii < liNumArks; ++ii) {
        mHeader.maParts[ii].muSize = (unsigned int)(luSizeRemaining / (liNumArks - ii));
        luSizeRemaining -= mHeader.maParts[ii].muSize;
    }
    mbLoaded = true;
    return eError_NoError;
}

eError CArk::BuildArk(const char* lpInputDirectory, std::vector<SSongConfig> laSongs)
{
    (void)laSongs;  // the reference appends its four defaults and never reads them here (CArk.cpp:762-765)
    VERBOSE_OUT("Building ark\n");
    if (mHeader.maParts.empty())
        return eError_NoData;

    uint64_t luTotalArkSize = 0;
    for (const modark::FileDef& lFile : mHeader.maFiles)
        luTotalArkSize += (uint64_t)lFile.miSize;
    ReleaseArkData();
    mpArkData = (unsigned char*)mod_host_alloc(luTotalArkSize + 1);
    if (!mpArkData) {
        std::cout << "Failed to allocate pinned memory for the archive: " << mod_last_error() << "\n";
        return eError_NoData;
    }
    muArkDataSize = luTotalArkSize;

    // pass 1 (serial, sizes only): byte-packed running offsets and part sizes
    std::string lRoot = lpInputDirectory;
    size_t liArkIndex = 0;
    int64_t li64Allowed = mHeader.maParts[0].muSize;
    uint64_t luPtr = 0, luPartStart = 0;
    std::vector<mod_desc> laDescs;
    std::vector<size_t> laToRead;
    bool lbAnyKey = false;
    for (size_t ii = 0; ii < mHeader.maFiles.size(); ++ii) {
        modark::FileDef& lFile = mHeader.maFiles[ii];
        if (lFile.miSize == 0) {
            lFile.mi64Offset = 0;
            continue;
        }
        // scatter: the file lands at the running, byte-packed offset
        lFile.mi64Offset = (int64_t)luPtr;
        laToRead.push_back(ii);
        const int liKey = EntryKey(ii);
        lbAnyKey = lbAnyKey || (liKey % 0x7FFFFFFF) != 0;
        laDescs.push_back(mod_desc{luPtr, luPtr, (uint32_t)lFile.miSize, liKey});
        luPtr += (uint64_t)lFile.miSize;

        // a part closes once it EXCEEDS its allowance; the overshoot shortens the next allowance
        if ((int64_t)(luPtr - luPartStart) > li64Allowed) {
            const unsigned int liArkSize = (unsigned int)(luPtr - luPartStart);
            mHeader.maParts[liArkIndex].muSize = liArkSize;
            if (liArkIndex + 1 >= mHeader.maParts.size()) {
                // the reference would index past mpArks here (CArk.cpp:817-818); grow a part instead
                modark::PartDef lExtra = mHeader.maParts[liArkIndex];
                lExtra.mPath += ".extra";
                lExtra.muSize = 0;
                mHeader.maParts.push_back(lExtra);
            }
            ++liArkIndex;
            li64Allowed += (int64_t)mHeader.maParts[liArkIndex].muSize - (int64_t)liArkSize;
            luPartStart = luPtr;
        }
    }
    mHeader.maParts[liArkIndex].muSize = (unsigned int)(luPtr - luPartStart);

    // pass 2: the payload reads (the reference's one fread per file, CArk.cpp:796-811) fanned out over
    // a few threads -- thousands of small files are latency-bound on any real file system
    std::atomic<size_t> liNext{0};
    std::atomic<bool> lbOpenFailed{false};
    auto lReadFiles = [&]() {
        for (;;) {
            const size_t liSlot = liNext.fetch_add(1);
            if (liSlot >= laToRead.size() || lbOpenFailed.load())
                return;
            const modark::FileDef& lFile = mHeader.maFiles[laToRead[liSlot]];
            FILE* lpInputFile = std::fopen((lRoot + lFile.mName).c_str(), "rb");
            if (!lpInputFile) {
                lbOpenFailed.store(true);
                return;
            }
            unsigned char* lpDst = mpArkData + lFile.mi64Offset;
            const size_t liRead = std::fread(lpDst, 1, (size_t)lFile.miSize, lpInputFile);
            std::fclose(lpInputFile);
            if (liRead != (size_t)lFile.miSize)
                std::memset(lpDst + liRead, 0, (size_t)lFile.miSize - liRead);
        }
    };
    {
        std::vector<std::thread> laReaders;
        for (int ii = 0; ii < 3; ++ii)
            laReaders.emplace_back(lReadFiles);
        lReadFiles();
        for (std::thread& lThread : laReaders)
            lThread.join();
    }
    if (lbOpenFailed.load()) {
        eError leError = eError_FailedToOpenFile;
        SHOW_ERROR_AND_RETURN;
    }

    // entries that carry a key are ciphered where they lie, all in one batched launch
    if (lbAnyKey && !laDescs.empty()) {
        if (mod_cycle_batch(laDescs.data(), laDescs.size(), mpArkData, muArkDataSize, mpArkData, muArkDataSize) != MOD_OK) {
            std::cout << "GPU build failed: " << mod_last_error() << "\n";
            return eError_InvalidData;
        }
    }
    VERBOSE_OUT("Ark built\n");
    return eError_NoError;
}

eError CArk::SaveArk(const char* lpOutputDirectory, const char* lpHeaderFilename) const
{
    // header: serialise on the host, encipher everything after the magic on the GPU, write
    std::vector<unsigned char> lImage = modark::SerialiseHeader(mHeader);
    CEncryptionCycler lEncrypt;
    lEncrypt.Cycle(lImage.data() + sizeof(uint32_t), (unsigned int)(lImage.size() - sizeof(uint32_t)),
                   (int)(mHeader.mbPS4 ? CSettings::kuEncryptedPS4Key : CSettings::kuEncryptedPS3Key));

    const std::string lHeaderPath = std::string(lpOutputDirectory) + lpHeaderFilename;
    if (KeepExistingOutput(lHeaderPath)) {
        VERBOSE_OUT("Output file already exists, skipping: " << lHeaderPath.c_str() << "\n");
    } else {
        MakeParentDirectories(lHeaderPath);
        VERBOSE_OUT("Writing " << lHeaderPath.c_str() << "\n");
        FILE* lpOutputFile = std::fopen(lHeaderPath.c_str(), "wb");
        if (lpOutputFile) {
            const size_t liWritten = std::fwrite(lImage.data(), 1, lImage.size(), lpOutputFile);
            std::fclose(lpOutputFile);
            if (liWritten != lImage.size()) {
                eError leError = eError_FailedToWriteData;
                SHOW_ERROR_AND_RETURN;
            }
        } else {
            std::cout << "Failed to open file for writing: " << lHeaderPath.c_str() << "\n";
        }
    }

    // parts: consecutive slices of the image
    const unsigned char* lpArkPtr = mpArkData;
    for (const modark::PartDef& lPart : mHeader.maParts) {
        const std::string lFilename = std::string(lpOutputDirectory) + lPart.mPath;
        if (KeepExistingOutput(lFilename)) {
            std::cout << "Output file already exists: " << lFilename.c_str() << "\n";
            continue;  // like the reference, the cursor does not advance past a skipped part (CArk.cpp:863-868)
        }
        std::cout << "Writing " << lFilename.c_str() << "\n";
        MakeParentDirectories(lFilename);
        FILE* lpOutputFile = std::fopen(lFilename.c_str(), "wb");
        if (!lpOutputFile) {
            std::cout << "Failed to open file for writing: " << lFilename.c_str() << "\n";
            continue;
        }
        const size_t liWritten = (lPart.muSize && lpArkPtr) ? std::fwrite(lpArkPtr, 1, lPart.muSize, lpOutputFile) : 0;
        std::fclose(lpOutputFile);
        if (liWritten != lPart.muSize) {
            eError leError = eError_FailedToWriteData;
            SHOW_ERROR_AND_RETURN;
        }
        lpArkPtr += lPart.muSize;
    }
    return eError_NoError;
}
