#include "CArk.h"

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>
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

// Tracing aid: reports a pipeline operation that took longer than 50 ms (MOD_TRACE=1).
struct SlowOp {
    const char* mpWhat;
    double mdStart;
    explicit SlowOp(const char* lpWhat) : mpWhat(lpWhat), mdStart(TraceEnabled() ? NowSeconds() : 0.0) {}
    ~SlowOp()
    {
        if (mdStart > 0.0 && NowSeconds() - mdStart > 0.05)
            std::fprintf(stderr, "[mod] slow: %s took %.3f s\n", mpWhat, NowSeconds() - mdStart);
    }
};

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
        // the directory entry usually says what it is; stat only when it does not (or for symlinks)
        bool lbDirectory = lpEntry->d_type == DT_DIR, lbFile = lpEntry->d_type == DT_REG;
        if (!lbDirectory && !lbFile) {
            const std::string lFull = lRoot + lRelative + lName;
            lbDirectory = IsDirectory(lFull);
            lbFile = !lbDirectory && FileSize(lFull) >= 0;
        }
        if (lbDirectory)
            laDirs.push_back(lName);
        else if (lbFile)
            laFiles.push_back(lName);
    }
    closedir(lpDir);
    // names that differ only in case cannot coexist on NTFS; on a case-sensitive file system they are
    // ordered by their bytes so that the walk is deterministic
    auto lLess = [](const std::string& lA, const std::string& lB) {
        const auto lUpperLess = [](char a, char b) { return std::toupper((unsigned char)a) < std::toupper((unsigned char)b); };
        if (std::lexicographical_compare(lA.begin(), lA.end(), lB.begin(), lB.end(), lUpperLess))
            return true;
        if (std::lexicographical_compare(lB.begin(), lB.end(), lA.begin(), lA.end(), lUpperLess))
            return false;
        return lA < lB;
    };
    std::sort(laFiles.begin(), laFiles.end(), lLess);
    std::sort(laDirs.begin(), laDirs.end(), lLess);
    for (const std::string& lName : laFiles)
        laOut.push_back(lRelative + lName);
    for (const std::string& lName : laDirs)
        ListFiles(lRoot, lRelative + lName + "/", laOut);
}

// mkdir -p with a per-thread memo of directories already known to exist (thousands of entries share
// a handful of folders; the reference re-runs _mkdir + stat for every component of every file,
// CArk.cpp:463-483).
bool MakeParentDirectoriesCached(const std::string& lFilePath, std::unordered_set<std::string>& lKnown)
{
    const size_t liLast = lFilePath.rfind('/');
    if (liLast == std::string::npos || liLast == 0)
        return true;
    if (lKnown.count(lFilePath.substr(0, liLast)))
        return true;
    if (!MakeParentDirectories(lFilePath))
        return false;
    lKnown.insert(lFilePath.substr(0, liLast));
    return true;
}

// A few worker threads draining a queue of closures: the reader and writer stages of the pipelines.
class TaskPool
{
public:
    explicit TaskPool(int liThreads)
    {
        for (int ii = 0; ii < liThreads; ++ii)
            maThreads.emplace_back([this]() { Loop(); });
    }
    ~TaskPool() { Finish(); }
    void Push(std::function<void()> lTask)
    {
        {
            std::lock_guard<std::mutex> lLock(mMutex);
            maTasks.push_back(std::move(lTask));
        }
        mSignal.notify_one();
    }
    void Finish()  // run what is queued, then stop
    {
        {
            std::lock_guard<std::mutex> lLock(mMutex);
            mbClosed = true;
        }
        mSignal.notify_all();
        for (std::thread& lThread : maThreads)
            if (lThread.joinable())
                lThread.join();
    }

private:
    void Loop()
    {
        for (;;) {
            std::function<void()> lTask;
            {
                std::unique_lock<std::mutex> lLock(mMutex);
                mSignal.wait(lLock, [this]() { return !maTasks.empty() || mbClosed; });
                if (maTasks.empty())
                    return;
                lTask = std::move(maTasks.front());
                maTasks.pop_front();
            }
            lTask();
        }
    }
    std::mutex mMutex;
    std::condition_variable mSignal;
    std::deque<std::function<void()>> maTasks;
    std::vector<std::thread> maThreads;
    bool mbClosed = false;
};

int HostThreads()
{
    const unsigned int luCores = std::thread::hardware_concurrency();
    return luCores ? (int)luCores : 4;
}

// Tuning aid: MOD_IO_<NAME>=<n> overrides a pipeline's thread / slot count (tools/unpack_timing.py sweeps them).
int Tunable(const char* lpName, int liDefault)
{
    const char* lpValue = std::getenv(lpName);
    const int liValue = lpValue && *lpValue ? std::atoi(lpValue) : 0;
    return liValue > 0 ? liValue : liDefault;
}

// GPUs the facade spreads an archive over: every visible device (MOD_DEVICES = bit mask restricts
// them) once the archive is large enough to be worth it, else only the calling thread's device.
std::vector<int> FacadeDevices(uint64_t luBytes)
{
    std::vector<int> laDevices;
    const int liCurrent = std::max(0, mod_current_device());
    const int liCount = mod_device_count();
    const char* lpMask = std::getenv("MOD_DEVICES");
    const uint64_t luMask = lpMask && *lpMask ? std::strtoull(lpMask, nullptr, 0) : 0;
    const uint64_t kuMultiGpuThreshold = 256ull << 20;
    if (liCount > 1 && luBytes >= kuMultiGpuThreshold)
        for (int ii = 0; ii < liCount && ii < 64; ++ii)
            if (luMask == 0 || ((luMask >> ii) & 1ull))
                laDevices.push_back(ii);
    if (laDevices.empty())
        laDevices.push_back(liCurrent);
    return laDevices;
}

// One stage buffer set of the streaming pipelines: pinned host memory either side of the GPU, the
// two HBM windows the batched kernel works on, and the stream that orders upload -> kernel -> download.
struct Slot {
    enum eState { eFree, eFilling, eFilled, eDraining };
    int miDevice = 0;
    unsigned char* mpHostIn = nullptr;   // what the readers fill (image range / input files)
    unsigned char* mpHostOut = nullptr;  // what the writers drain (extracted files / ciphered image range)
    void* mpDevIn = nullptr;
    void* mpDevOut = nullptr;
    void* mpStream = nullptr;
    eState meState = eFree;
    int miFillLeft = 0;
    int miDrainLeft = 0;
    bool mbUsedGpu = false;
    uint64_t muInCapacity = 0, muOutCapacity = 0;
};

// Slots handed back by a finished pipeline, kept for the next one: page-locking memory costs more than
// moving the data does (0.3-1 ms per MiB), and a repack runs an extract and a build back to back.  At
// most kiMaxCached slots are kept; the cache is never destroyed (the CUDA context may be gone by the
// time static destructors run), the OS reclaims it with the process.
struct SlotCache {
    std::mutex mMutex;
    std::vector<Slot> maSlots;
    static constexpr size_t kiMaxCached = 24;
    static SlotCache& Get()
    {
        static SlotCache* lpCache = new SlotCache();
        return *lpCache;
    }
    bool Take(int liDevice, uint64_t luIn, uint64_t luOut, bool lbNeedGpu, Slot& lOut)
    {
        std::lock_guard<std::mutex> lLock(mMutex);
        for (size_t ii = 0; ii < maSlots.size(); ++ii) {
            Slot& lSlot = maSlots[ii];
            if (lSlot.muInCapacity >= luIn && lSlot.muOutCapacity >= luOut && (!lbNeedGpu || (lSlot.mpStream && lSlot.miDevice == liDevice))) {
                lOut = lSlot;
                maSlots.erase(maSlots.begin() + (long)ii);
                return true;
            }
        }
        return false;
    }
    bool Give(const Slot& lSlot)
    {
        std::lock_guard<std::mutex> lLock(mMutex);
        if (maSlots.size() >= kiMaxCached)
            return false;
        maSlots.push_back(lSlot);
        return true;
    }
};

struct SlotRing {
    std::vector<Slot> maSlots;
    std::mutex mMutex;
    std::condition_variable mSignal;
    std::atomic<int> miError{(int)eError_NoError};

    void Fail(eError leWhat)
    {
        int liExpected = (int)eError_NoError;
        miError.compare_exchange_strong(liExpected, (int)leWhat);
        mSignal.notify_all();
    }
    bool Failed() const { return miError.load() != (int)eError_NoError; }
    void Set(Slot& lSlot, Slot::eState leState)
    {
        {
            std::lock_guard<std::mutex> lLock(mMutex);
            lSlot.meState = leState;
        }
        mSignal.notify_all();
    }
    // one unit of fill / drain work done; the last one flips the slot's state
    void FillDone(Slot& lSlot)
    {
        bool lbLast = false;
        {
            std::lock_guard<std::mutex> lLock(mMutex);
            lbLast = --lSlot.miFillLeft == 0;
            if (lbLast)
                lSlot.meState = Slot::eFilled;
        }
        if (lbLast)
            mSignal.notify_all();
    }
    void DrainDone(Slot& lSlot)
    {
        bool lbLast = false;
        {
            std::lock_guard<std::mutex> lLock(mMutex);
            lbLast = --lSlot.miDrainLeft == 0;
            if (lbLast)
                lSlot.meState = Slot::eFree;
        }
        if (lbLast)
            mSignal.notify_all();
    }
    // Allocate liPerDevice slots on each device: pinned buffers of luInBytes / luOutBytes, HBM windows
    // of the same sizes (+ slack for the 16-byte phase), one stream each.
    bool Allocate(const std::vector<int>& laDevices, int liPerDevice, size_t liGroups, uint64_t luInBytes, uint64_t luOutBytes,
                  bool lbNeedGpu)
    {
        const size_t liWanted = std::min(liGroups, (size_t)liPerDevice * laDevices.size());
        maSlots.resize(std::max<size_t>(1, liWanted));
        std::vector<size_t> laFresh;  // slots the cache could not supply
        for (size_t ii = 0; ii < maSlots.size(); ++ii) {
            Slot& lSlot = maSlots[ii];
            const int liDevice = laDevices[ii % laDevices.size()];
            if (SlotCache::Get().Take(liDevice, luInBytes + 32, luOutBytes + 32, lbNeedGpu, lSlot)) {
                lSlot.meState = Slot::eFree;
                lSlot.mbUsedGpu = false;
                continue;
            }
            lSlot.miDevice = liDevice;
            lSlot.muInCapacity = luInBytes + 32;
            lSlot.muOutCapacity = luOutBytes + 32;
            laFresh.push_back(ii);
        }
        // Page-locking dominates a cold start (0.3-1 ms per MiB, i.e. more than the whole pipeline of a 1 GiB
        // archive): the fresh slots are therefore set up side by side, one thread each.
        std::atomic<bool> lbOk{true};
        auto lSetUp = [&](size_t ii) {
            Slot& lSlot = maSlots[ii];
            lSlot.mpHostIn = (unsigned char*)mod_host_alloc(lSlot.muInCapacity);
            lSlot.mpHostOut = (unsigned char*)mod_host_alloc(lSlot.muOutCapacity);
            bool lbGood = lSlot.mpHostIn && lSlot.mpHostOut;
            if (lbGood && lbNeedGpu) {
                lbGood = mod_init(lSlot.miDevice) == MOD_OK;
                if (lbGood) {
                    lSlot.mpDevIn = mod_device_alloc(lSlot.muInCapacity);
                    lSlot.mpDevOut = mod_device_alloc(lSlot.muOutCapacity);
                    lSlot.mpStream = mod_stream_create();
                    lbGood = lSlot.mpDevIn && lSlot.mpDevOut && lSlot.mpStream;
                }
            }
            if (!lbGood)
                lbOk.store(false);
        };
        if (laFresh.size() <= 1 || Tunable("MOD_IO_SERIAL_SETUP", 0)) {
            for (size_t ii : laFresh)
                lSetUp(ii);
        } else {
            std::vector<std::thread> laThreads;
            for (size_t ii : laFresh)
                laThreads.emplace_back(lSetUp, ii);
            for (std::thread& lThread : laThreads)
                lThread.join();
        }
        return lbOk.load();
    }
    void Release()
    {
        for (Slot& lSlot : maSlots) {
            if (lSlot.mpStream)
                mod_stream_sync(lSlot.mpStream);
            if (lSlot.mpHostIn && lSlot.mpHostOut && SlotCache::Get().Give(lSlot))
                continue;
            if (lSlot.mpStream)
                mod_stream_destroy(lSlot.mpStream);
            mod_device_free(lSlot.mpDevIn);
            mod_device_free(lSlot.mpDevOut);
            mod_host_free(lSlot.mpHostIn);
            mod_host_free(lSlot.mpHostOut);
        }
        maSlots.clear();
    }
};

// pread / pwrite the whole range (short transfers are retried; a short FILE reads as zeros).
bool ReadFully(int liFd, unsigned char* lpDst, uint64_t luSize, uint64_t luOffset)
{
    while (luSize) {
        const ssize_t liGot = pread(liFd, lpDst, (size_t)std::min<uint64_t>(luSize, 1u << 30), (off_t)luOffset);
        if (liGot < 0)
            return false;
        if (liGot == 0) {
            std::memset(lpDst, 0, (size_t)luSize);
            return true;
        }
        lpDst += liGot;
        luOffset += (uint64_t)liGot;
        luSize -= (uint64_t)liGot;
    }
    return true;
}

bool WriteFully(int liFd, const unsigned char* lpSrc, uint64_t luSize, uint64_t luOffset)
{
    while (luSize) {
        const ssize_t liPut = pwrite(liFd, lpSrc, (size_t)std::min<uint64_t>(luSize, 1u << 30), (off_t)luOffset);
        if (liPut <= 0)
            return false;
        lpSrc += liPut;
        luOffset += (uint64_t)liPut;
        luSize -= (uint64_t)liPut;
    }
    return true;
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
    std::vector<int> laFds;
    const bool lbOk = ReadImageRange(laFds, 0, muArkDataSize, mpArkData);
    for (int liFd : laFds)
        if (liFd >= 0)
            close(liFd);
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

// The reference's whole-image load (CArk.cpp:723-758), kept for callers that want the flat image;
// ExtractFiles does not need it any more (it streams ranges through a slot ring).
eError CArk::LoadArkData()
{
    eError leError = AllocateArkData();
    ERROR_RETURN;
    leError = ReadParts();
    SHOW_ERROR_AND_RETURN;
    return eError_NoError;
}

// Copy bytes [luOffset, luOffset + luSize) of the concatenated part files into lpDst with pread (safe
// from several threads at once), opening each part on first use (laFds caches the descriptors; it
// must be sized and guarded by the caller when shared).  Bytes a short part file lacks read as zero.
bool CArk::ReadImageRange(std::vector<int>& laFds, uint64_t luOffset, uint64_t luSize, unsigned char* lpDst) const
{
    static std::mutex lOpenMutex;
    {
        std::lock_guard<std::mutex> lLock(lOpenMutex);
        if (laFds.size() < mHeader.maParts.size())
            laFds.resize(mHeader.maParts.size(), -1);
    }
    uint64_t luPartStart = 0;
    for (size_t ii = 0; ii < mHeader.maParts.size() && luSize; ++ii) {
        const uint64_t luPartSize = mHeader.maParts[ii].muSize;
        const uint64_t luPartEnd = luPartStart + luPartSize;
        if (luOffset < luPartEnd) {
            int liFd;
            {
                std::lock_guard<std::mutex> lLock(lOpenMutex);
                if (laFds[ii] < 0)
                    laFds[ii] = open((mPartDirectory + mHeader.maParts[ii].mPath).c_str(), O_RDONLY);
                liFd = laFds[ii];
            }
            if (liFd < 0)
                return false;
            const uint64_t luTake = std::min(luSize, luPartEnd - luOffset);
            if (!ReadFully(liFd, lpDst, luTake, luOffset - luPartStart))
                return false;
            lpDst += luTake;
            luOffset += luTake;
            luSize -= luTake;
        }
        luPartStart = luPartEnd;
    }
    return luSize == 0;
}

// Extraction streams the archive through a ring of slots instead of the reference's one archive-sized
// buffer (CArk.cpp:431, :738) and its serial one-file-at-a-time writes (:435-501):
//   * the file table becomes ONE descriptor plan (mod_plan_create: one tile map for the whole archive,
//     on every GPU in use), the entries ordered by image offset and cut into ~32 MiB groups;
//   * reader threads pread the image range of a group from the part files into the slot's pinned buffer;
//   * this thread only ENQUEUES, per group and never waiting: upload of the range, the batched kernel over
//     the group's tile sub-range (mod_plan_run_window: gather + per-entry Cycle into a byte-packed
//     staging window), download -- on the slot's own stream, slots (and GPUs) taking groups in turn;
//   * writer threads wait for the slot's stream, create directories and write the group's files, then
//     hand the slot back.
// Disk reads, PCIe both ways, the kernel and disk writes all overlap; pinned memory and HBM are
// O(slots), not O(archive).  Entries that share an output name are resolved up front to the one whose
// bytes the reference's in-order loop would leave on disk.
eError CArk::ExtractFiles(int liFirstFileIndex, int liNumFiles, const char* lpTargetDirectory)
{
    (void)liFirstFileIndex;  // the reference ignores both and walks the whole table (CArk.cpp:435)
    (void)liNumFiles;
    if (mHeader.maFiles.empty())
        return eError_NoData;
    const double ldStart = NowSeconds();
    const std::string lTarget = lpTargetDirectory;

    uint64_t luImageSize = 0;
    for (const modark::PartDef& lPart : mHeader.maParts)
        luImageSize += lPart.muSize;

    const size_t liCount = mHeader.maFiles.size();
    for (size_t ii = 0; ii < liCount; ++ii) {
        const modark::FileDef& lFile = mHeader.maFiles[ii];
        if ((uint64_t)lFile.mi64Offset > luImageSize || (uint64_t)lFile.miSize > luImageSize - (uint64_t)lFile.mi64Offset) {
            std::cout << "Entry " << lFile.mName.c_str() << " lies outside the archive data\n";
            eError leError = eError_InvalidData;
            SHOW_ERROR_AND_RETURN;
        }
    }

    // Duplicate output names.  The reference writes in table order, one file at a time (CArk.cpp:435-501):
    // with overwriting on (the default) the LAST entry of a name is what stays on disk; with it off a
    // file is only rewritten while it is still empty, so the FIRST non-empty entry stays.  Resolve that
    // here so that exactly one entry per name is extracted, whatever order the pipeline writes in.
    std::vector<uint32_t> laWinners;
    {
        std::unordered_map<std::string, uint32_t> lByName;  // name -> position in laWinners
        laWinners.reserve(liCount);
        for (size_t ii = 0; ii < liCount; ++ii) {
            const modark::FileDef& lFile = mHeader.maFiles[ii];
            const auto lFound = lByName.find(lFile.mName);
            if (lFound == lByName.end()) {
                lByName.emplace(lFile.mName, (uint32_t)laWinners.size());
                laWinners.push_back((uint32_t)ii);
                continue;
            }
            uint32_t& luWinner = laWinners[lFound->second];
            if (CSettings::mbOverwriteOutputFiles || mHeader.maFiles[luWinner].miSize == 0)
                luWinner = (uint32_t)ii;
            VERBOSE_OUT("Duplicate entry " << lFile.mName.c_str() << "\n");
        }
    }

    // order by where the entries sit in the image, cut entries larger than a group into PIECES (a piece that
    // starts p bytes into its entry continues the entry's keystream from the jumped key), then form groups of
    // consecutive pieces (~32 MiB of payload each): a slot never has to hold more than about one group
    std::vector<uint32_t> laOrder = laWinners;
    std::stable_sort(laOrder.begin(), laOrder.end(), [&](uint32_t a, uint32_t b) {
        return mHeader.maFiles[a].mi64Offset < mHeader.maFiles[b].mi64Offset;
    });
    struct Piece {
        uint32_t file;        // index into mHeader.maFiles
        uint64_t fileOffset;  // where in the output file the piece goes
        uint32_t size;
        bool whole;           // the piece is the whole file
    };
    struct Group {
        size_t first, last;     // positions in laPieces
        uint64_t srcLo, srcHi;  // image range
        uint64_t dstLo, dstHi;  // range of the byte-packed staging space
    };
    const uint64_t kuGroupBytes = (uint64_t)Tunable("MOD_IO_GROUP_MIB", 32) << 20;
    std::vector<Piece> laPieces;
    laPieces.reserve(laOrder.size());
    for (uint32_t luFile : laOrder) {
        const uint64_t luSize = (uint64_t)mHeader.maFiles[luFile].miSize;
        if (luSize <= kuGroupBytes + kuGroupBytes / 2) {
            laPieces.push_back(Piece{luFile, 0, (uint32_t)luSize, true});
            continue;
        }
        for (uint64_t luAt = 0; luAt < luSize;) {
            uint64_t luTake = std::min(kuGroupBytes, luSize - luAt);
            if (luSize - luAt - luTake < kuGroupBytes / 2)
                luTake = luSize - luAt;  // no short tail piece
            laPieces.push_back(Piece{luFile, luAt, (uint32_t)luTake, false});
            luAt += luTake;
        }
    }
    std::vector<Group> laGroups;
    std::vector<mod_desc> laDescs(laPieces.size());
    uint64_t luMaxRange = 1, luMaxPayload = 1, luStaged = 0;
    for (size_t liFirst = 0; liFirst < laPieces.size();) {
        Group lGroup{liFirst, liFirst, UINT64_MAX, 0, luStaged, luStaged};
        while (lGroup.last < laPieces.size() && (lGroup.dstHi - lGroup.dstLo < kuGroupBytes || lGroup.last == liFirst)) {
            const Piece& lPiece = laPieces[lGroup.last];
            const uint64_t luSource = (uint64_t)mHeader.maFiles[lPiece.file].mi64Offset + lPiece.fileOffset;
            if (lPiece.size) {
                lGroup.srcLo = std::min(lGroup.srcLo, luSource);
                lGroup.srcHi = std::max(lGroup.srcHi, luSource + lPiece.size);
            }
            const int liKey = EntryKey(lPiece.file);
            laDescs[lGroup.last] = mod_desc{luSource, lGroup.dstHi, lPiece.size,
                                            lPiece.fileOffset ? mod_key_jump(liKey, lPiece.fileOffset) : liKey};
            lGroup.dstHi += lPiece.size;
            ++lGroup.last;
        }
        if (lGroup.srcLo == UINT64_MAX)
            lGroup.srcLo = lGroup.srcHi = 0;
        luStaged = lGroup.dstHi;
        luMaxRange = std::max(luMaxRange, lGroup.srcHi - lGroup.srcLo);
        luMaxPayload = std::max(luMaxPayload, lGroup.dstHi - lGroup.dstLo);
        laGroups.push_back(lGroup);
        liFirst = lGroup.last;
    }

    // files that arrive in several pieces are created (and emptied) up front, so that the pieces -- written by
    // whichever threads get them, in any order -- only ever pwrite into an existing file
    std::vector<uint8_t> laSkipFile(liCount, 0);
    for (const Piece& lPiece : laPieces) {
        if (lPiece.whole || lPiece.fileOffset != 0)
            continue;
        const modark::FileDef& lFile = mHeader.maFiles[lPiece.file];
        const std::string lOutputPath = lTarget + lFile.mName;
        laSkipFile[lPiece.file] = 1;
        if (EscapesTarget(lFile.mName)) {
            std::cout << "Refusing to write outside the target directory: " << lFile.mName.c_str() << "\n";
        } else if (KeepExistingOutput(lOutputPath)) {
            VERBOSE_OUT("Output file already exists, skipping: " << lOutputPath.c_str() << "\n");
        } else if (!MakeParentDirectories(lOutputPath)) {
            eError leError = eError_FailedToCreateDirectory;
            SHOW_ERROR_AND_RETURN;
        } else {
            const int liFd = open(lOutputPath.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
            if (liFd < 0) {
                std::cout << "Failed to create " << lOutputPath.c_str() << "\n";
            } else {
                close(liFd);
                laSkipFile[lPiece.file] = 0;
            }
        }
    }

    // slots on every GPU in use, and the archive's plan on each of them
    const int liOriginalDevice = mod_current_device();
    const std::vector<int> laDevices = FacadeDevices(luStaged);
    SlotRing lRing;
    std::vector<mod_plan*> laPlans(laDevices.size(), nullptr);
    auto lCleanup = [&]() {
        const double ldT0 = NowSeconds();
        lRing.Release();
        const double ldT1 = NowSeconds();
        for (mod_plan* lpPlan : laPlans)
            mod_plan_destroy(lpPlan);
        if (liOriginalDevice >= 0)
            mod_init(liOriginalDevice);
        if (TraceEnabled())
            std::fprintf(stderr, "[mod] ExtractFiles cleanup: slots %.3f s, plans %.3f s\n", ldT1 - ldT0, NowSeconds() - ldT1);
    };
    const double ldPlanned = NowSeconds();
    const int liSlotsPerDevice = Tunable("MOD_IO_SLOTS", std::max(3, 6 / (int)laDevices.size()));
    bool lbReady = lRing.Allocate(laDevices, liSlotsPerDevice, laGroups.size(), luMaxRange, luMaxPayload, true);
    const double ldSlots = NowSeconds();
    for (size_t dd = 0; dd < laDevices.size() && lbReady; ++dd)
        lbReady = mod_init(laDevices[dd]) == MOD_OK &&
                  mod_plan_create(laDescs.data(), laDescs.size(), luImageSize, luStaged, 0, &laPlans[dd]) == MOD_OK;
    if (!lbReady) {
        std::cout << "GPU extract could not be set up: " << mod_last_error() << "\n";
        lCleanup();
        return eError_NoData;
    }
    const double ldAllocated = NowSeconds();
    if (TraceEnabled())
        std::fprintf(stderr, "[mod] ExtractFiles setup: table -> groups %.3f s, slots %.3f s, plans %.3f s\n", ldPlanned - ldStart,
                     ldSlots - ldPlanned, ldAllocated - ldSlots);

    const int liCores = HostThreads();
    TaskPool lReaders(Tunable("MOD_IO_READERS", std::max(2, std::min(8, liCores / 4))));
    TaskPool lWriters(Tunable("MOD_IO_WRITERS", std::max(4, std::min(16, liCores / 2))));
    std::vector<int> laPartFds(mHeader.maParts.size(), -1);

    // stage 1 (reader threads): the image range of a group, in pieces
    auto lStartFill = [&](size_t gg) {
        const Group& lGroup = laGroups[gg];
        Slot& lSlot = lRing.maSlots[gg % lRing.maSlots.size()];
        const uint64_t kuPiece = 8ull << 20;
        const uint64_t luRange = lGroup.srcHi - lGroup.srcLo;
        const int liPieces = (int)std::max<uint64_t>(1, (luRange + kuPiece - 1) / kuPiece);
        {
            std::lock_guard<std::mutex> lLock(lRing.mMutex);
            lSlot.meState = Slot::eFilling;
            lSlot.miFillLeft = liPieces;
        }
        for (int pp = 0; pp < liPieces; ++pp) {
            lReaders.Push([&, pp, luRange, kuPiece]() {
                SlowOp lTimer("ExtractFiles read of an 8 MiB image piece");
                const uint64_t luLo = (uint64_t)pp * kuPiece, luHi = std::min(luRange, luLo + kuPiece);
                // the range keeps its (offset & 15) phase inside the pinned buffer, like on the device
                if (!lRing.Failed() && luHi > luLo &&
                    !ReadImageRange(laPartFds, lGroup.srcLo + luLo, luHi - luLo, lSlot.mpHostIn + (lGroup.srcLo & 15u) + luLo))
                    lRing.Fail(eError_FailedToOpenFile);
                lRing.FillDone(lSlot);
            });
        }
    };

    // stage 3 (writer threads): the files of a group, a few dozen per task
    auto lWriteFiles = [&](size_t gg, size_t liFrom, size_t liTo) {
        const Group& lGroup = laGroups[gg];
        Slot& lSlot = lRing.maSlots[gg % lRing.maSlots.size()];
        thread_local std::unordered_set<std::string> lKnownDirectories;
        SlowOp lTimer("ExtractFiles writer task (stream sync + up to 48 files)");
        if (!lRing.Failed() && lSlot.mbUsedGpu && [&]() { SlowOp lSync("ExtractFiles stream sync"); return mod_stream_sync(lSlot.mpStream); }() != MOD_OK) {
            std::cout << "GPU extract failed: " << mod_last_error() << "\n";
            lRing.Fail(eError_InvalidData);
        }
        for (size_t ii = liFrom; ii < liTo && !lRing.Failed(); ++ii) {
            const Piece& lPiece = laPieces[ii];
            const modark::FileDef& lFile = mHeader.maFiles[lPiece.file];
            const std::string lOutputPath = lTarget + lFile.mName;
            const unsigned char* lpBytes = lSlot.mpHostOut + (lGroup.dstLo & 15u) + (laDescs[ii].dst_off - lGroup.dstLo);
            if (!lPiece.whole) {  // one piece of a big file that was created up front
                if (laSkipFile[lPiece.file])
                    continue;
                const int liFd = open(lOutputPath.c_str(), O_WRONLY);
                const bool lbOk = liFd >= 0 && WriteFully(liFd, lpBytes, lPiece.size, lPiece.fileOffset);
                if (liFd >= 0)
                    close(liFd);
                if (!lbOk)
                    lRing.Fail(eError_FailedToWriteData);
                continue;
            }
            if (EscapesTarget(lFile.mName)) {
                std::cout << "Refusing to write outside the target directory: " << lFile.mName.c_str() << "\n";
                continue;
            }
            if (KeepExistingOutput(lOutputPath)) {
                VERBOSE_OUT("Output file already exists, skipping: " << lOutputPath.c_str() << "\n");
                continue;
            }
            if (!MakeParentDirectoriesCached(lOutputPath, lKnownDirectories)) {
                lRing.Fail(eError_FailedToCreateDirectory);
                break;
            }
            const int liFd = open(lOutputPath.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
            if (liFd < 0) {
                std::cout << "Failed to create " << lOutputPath.c_str() << "\n";  // the reference carries on too (CArk.cpp:488-491)
                continue;
            }
            const bool lbOk = lFile.miSize == 0 || WriteFully(liFd, lpBytes, (uint64_t)lFile.miSize, 0);
            close(liFd);
            if (!lbOk)
                lRing.Fail(eError_FailedToWriteData);
        }
        lRing.DrainDone(lSlot);
    };

    // stage 2 (this thread): enqueue upload -> kernel -> download per group, never waiting for the GPU
    size_t liNextFill = 0, liNextGpu = 0;
    while (liNextGpu < laGroups.size() && !lRing.Failed()) {
        {
            std::unique_lock<std::mutex> lLock(lRing.mMutex);
            lRing.mSignal.wait(lLock, [&]() {
                const bool lbCanFill = liNextFill < laGroups.size() && liNextFill < liNextGpu + lRing.maSlots.size() &&
                                       lRing.maSlots[liNextFill % lRing.maSlots.size()].meState == Slot::eFree;
                const bool lbCanRun = lRing.maSlots[liNextGpu % lRing.maSlots.size()].meState == Slot::eFilled && liNextGpu < liNextFill;
                return lbCanFill || lbCanRun || lRing.Failed();
            });
        }
        if (lRing.Failed())
            break;
        while (liNextFill < laGroups.size() && liNextFill < liNextGpu + lRing.maSlots.size() &&
               lRing.maSlots[liNextFill % lRing.maSlots.size()].meState == Slot::eFree)
            lStartFill(liNextFill++);
        while (liNextGpu < liNextFill && lRing.maSlots[liNextGpu % lRing.maSlots.size()].meState == Slot::eFilled) {
            const size_t gg = liNextGpu++;
            const Group& lGroup = laGroups[gg];
            Slot& lSlot = lRing.maSlots[gg % lRing.maSlots.size()];
            const uint64_t luRange = lGroup.srcHi - lGroup.srcLo, luPayload = lGroup.dstHi - lGroup.dstLo;
            lSlot.mbUsedGpu = luPayload != 0;
            if (luPayload) {
                size_t liDeviceIndex = 0;
                while (laDevices[liDeviceIndex] != lSlot.miDevice)
                    ++liDeviceIndex;
                uint64_t luTile0 = 0, luTile1 = 0;
                SlowOp lTimer("ExtractFiles GPU enqueue of a group");
                unsigned char* lpDevIn = (unsigned char*)lSlot.mpDevIn + (lGroup.srcLo & 15u);
                unsigned char* lpDevOut = (unsigned char*)lSlot.mpDevOut + (lGroup.dstLo & 15u);
                const bool lbOk =
                    mod_init(lSlot.miDevice) == MOD_OK &&
                    mod_plan_tile_range(laPlans[liDeviceIndex], lGroup.first, lGroup.last, &luTile0, &luTile1) == MOD_OK &&
                    mod_memcpy_h2d(lpDevIn, lSlot.mpHostIn + (lGroup.srcLo & 15u), luRange, lSlot.mpStream) == MOD_OK &&
                    mod_plan_run_window(laPlans[liDeviceIndex], luTile0, luTile1, lpDevIn, lGroup.srcLo, luRange, lpDevOut,
                                        lGroup.dstLo, luPayload, lSlot.mpStream) == MOD_OK &&
                    mod_memcpy_d2h(lSlot.mpHostOut + (lGroup.dstLo & 15u), lpDevOut, luPayload, lSlot.mpStream) == MOD_OK;
                if (!lbOk) {
                    std::cout << "GPU extract failed: " << mod_last_error() << "\n";
                    lRing.Fail(eError_InvalidData);
                    break;
                }
            }
            const size_t kiFilesPerTask = 48;
            const size_t liFiles = lGroup.last - lGroup.first;
            const int liTasks = (int)std::max<size_t>(1, (liFiles + kiFilesPerTask - 1) / kiFilesPerTask);
            {
                std::lock_guard<std::mutex> lLock(lRing.mMutex);
                lSlot.meState = Slot::eDraining;
                lSlot.miDrainLeft = liTasks;
            }
            for (int tt = 0; tt < liTasks; ++tt) {
                const size_t liFrom = lGroup.first + (size_t)tt * kiFilesPerTask;
                lWriters.Push([&, gg, liFrom, liTo = std::min(lGroup.last, liFrom + kiFilesPerTask)]() { lWriteFiles(gg, liFrom, liTo); });
            }
        }
    }
    const double ldEnqueued = NowSeconds();
    lReaders.Finish();
    lWriters.Finish();
    const double ldWritten = NowSeconds();
    for (int liFd : laPartFds)
        if (liFd >= 0)
            close(liFd);
    const size_t liSlotsUsed = lRing.maSlots.size();
    lCleanup();
    if (TraceEnabled())
        std::fprintf(stderr, "[mod] ExtractFiles: %zu groups on %zu GPU(s), %zu slots; setup %.3f s, pipeline %.3f s, drain %.3f s, free %.3f s\n",
                     laGroups.size(), laDevices.size(), liSlotsUsed,
                     ldAllocated - ldStart, ldEnqueued - ldAllocated, ldWritten - ldEnqueued, NowSeconds() - ldWritten);

    eError leError = (eError)lRing.miError.load();
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
    // name -> FIRST entry of that name in the reference header (the reference scans linearly for the first
    // match of the name hash, CArk.cpp:1194-1209)
    std::unordered_map<std::string, const modark::FileDef*> lReferenceByName;
    lReferenceByName.reserve(lReferenceHeader.mHeader.maFiles.size());
    for (const modark::FileDef& lCandidate : lReferenceHeader.mHeader.maFiles)
        lReferenceByName.emplace(lCandidate.mName, &lCandidate);
    for (const std::string& lFilename : laFilenames) {
        const auto lFound = lReferenceByName.find(lFilename);
        const modark::FileDef* lpReference = lFound == lReferenceByName.end() ? nullptr : lFound->second;
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

// BuildArk assigns every entry its byte-packed offset and closes the parts (reference CArk.cpp:760-828)
// from the SIZES alone; the payload itself is never gathered into one archive-sized buffer
// (CArk.cpp:769-811 does): SaveArk streams it from the input files straight into the part files.
eError CArk::BuildArk(const char* lpInputDirectory, std::vector<SSongConfig> laSongs)
{
    (void)laSongs;  // the reference appends its four defaults and never reads them here (CArk.cpp:762-765)
    VERBOSE_OUT("Building ark\n");
    if (mHeader.maParts.empty())
        return eError_NoData;
    ReleaseArkData();
    mBuildInputDirectory = lpInputDirectory;

    size_t liArkIndex = 0;
    int64_t li64Allowed = mHeader.maParts[0].muSize;
    uint64_t luPtr = 0, luPartStart = 0;
    for (size_t ii = 0; ii < mHeader.maFiles.size(); ++ii) {
        modark::FileDef& lFile = mHeader.maFiles[ii];
        if (lFile.miSize == 0) {
            lFile.mi64Offset = 0;
            continue;
        }
        // the reference opens (and reads) the file here and gives up on the first one it cannot open
        if (access((mBuildInputDirectory + lFile.mName).c_str(), R_OK) != 0) {
            eError leError = eError_FailedToOpenFile;
            SHOW_ERROR_AND_RETURN;
        }
        // scatter: the file lands at the running, byte-packed offset
        lFile.mi64Offset = (int64_t)luPtr;
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
    muBuiltImageSize = luPtr;
    mbBuilt = true;
    VERBOSE_OUT("Ark built\n");
    return eError_NoError;
}

// Stream the image BuildArk laid out into the part files: reader threads load the input files of a
// ~32 MiB group into a pinned slot, entries that carry a key are ciphered on the GPU (upload, one
// mod_plan_run_window over the group's tiles of the archive-wide plan, download -- enqueued on the
// slot's stream, never waited for here), writer threads pwrite the group's byte range into the part
// file(s) it falls in.  With all keys 0 (the reference never ciphers ARK bodies) the GPU is not
// involved and the bytes go from the read buffer straight to the part files.
eError CArk::StreamBuiltImage(const std::vector<PartTarget>& laTargets) const
{
    const double ldStart = NowSeconds();
    // non-empty entries in table order == image order; files larger than a group are read (and ciphered) in
    // PIECES, each continuing the entry's keystream from the jumped key, so a slot stays about one group big
    struct Piece {
        uint32_t file;        // index into mHeader.maFiles
        uint64_t fileOffset;  // where in the input file the piece starts
        uint32_t size;
    };
    const uint64_t kuGroupBytes = (uint64_t)Tunable("MOD_IO_GROUP_MIB", 32) << 20;
    std::vector<Piece> laPieces;
    std::vector<mod_desc> laDescs;
    bool lbAnyKey = false;
    for (size_t ii = 0; ii < mHeader.maFiles.size(); ++ii) {
        const modark::FileDef& lFile = mHeader.maFiles[ii];
        if (lFile.miSize == 0)
            continue;
        const int liKey = EntryKey(ii);
        lbAnyKey = lbAnyKey || (liKey % 0x7FFFFFFF) != 0;
        const uint64_t luSize = (uint64_t)lFile.miSize;
        for (uint64_t luAt = 0; luAt < luSize;) {
            uint64_t luTake = luSize <= kuGroupBytes + kuGroupBytes / 2 ? luSize : std::min(kuGroupBytes, luSize - luAt);
            if (luSize - luAt - luTake < kuGroupBytes / 2)
                luTake = luSize - luAt;  // no short tail piece
            const uint64_t luImage = (uint64_t)lFile.mi64Offset + luAt;
            laPieces.push_back(Piece{(uint32_t)ii, luAt, (uint32_t)luTake});
            laDescs.push_back(mod_desc{luImage, luImage, (uint32_t)luTake, luAt ? mod_key_jump(liKey, luAt) : liKey});
            luAt += luTake;
        }
    }
    if (laPieces.empty())
        return eError_NoError;

    struct Group {
        size_t first, last;  // positions in laPieces
        uint64_t lo, hi;     // image range
    };
    std::vector<Group> laGroups;
    uint64_t luMaxRange = 1;
    for (size_t liFirst = 0; liFirst < laPieces.size();) {
        Group lGroup{liFirst, liFirst, laDescs[liFirst].src_off, laDescs[liFirst].src_off};
        while (lGroup.last < laPieces.size() && (lGroup.hi - lGroup.lo < kuGroupBytes || lGroup.last == liFirst)) {
            lGroup.hi = laDescs[lGroup.last].src_off + laDescs[lGroup.last].len;
            ++lGroup.last;
        }
        luMaxRange = std::max(luMaxRange, lGroup.hi - lGroup.lo);
        laGroups.push_back(lGroup);
        liFirst = lGroup.last;
    }

    const int liOriginalDevice = lbAnyKey ? mod_current_device() : -1;
    const std::vector<int> laDevices = lbAnyKey ? FacadeDevices(muBuiltImageSize) : std::vector<int>{0};
    SlotRing lRing;
    std::vector<mod_plan*> laPlans(laDevices.size(), nullptr);
    auto lCleanup = [&]() {
        lRing.Release();
        for (mod_plan* lpPlan : laPlans)
            mod_plan_destroy(lpPlan);
        if (liOriginalDevice >= 0)
            mod_init(liOriginalDevice);
    };
    const int liSlotsPerDevice = Tunable("MOD_IO_SLOTS", std::max(3, 6 / (int)laDevices.size()));
    bool lbReady = lRing.Allocate(laDevices, liSlotsPerDevice, laGroups.size(), luMaxRange, lbAnyKey ? luMaxRange : 0, lbAnyKey);
    for (size_t dd = 0; dd < laDevices.size() && lbReady && lbAnyKey; ++dd)
        lbReady = mod_init(laDevices[dd]) == MOD_OK &&
                  mod_plan_create(laDescs.data(), laDescs.size(), muBuiltImageSize, muBuiltImageSize, 0, &laPlans[dd]) == MOD_OK;
    if (!lbReady) {
        std::cout << "GPU build could not be set up: " << mod_last_error() << "\n";
        lCleanup();
        return eError_NoData;
    }

    const double ldReady = NowSeconds();
    const int liCores = HostThreads();
    TaskPool lReaders(Tunable("MOD_IO_READERS", std::max(2, std::min(12, liCores / 2))));
    TaskPool lWriters(Tunable("MOD_IO_WRITERS", std::max(2, std::min(12, liCores / 2))));

    // stage 1 (reader threads): the input files of a group, a few dozen per task (the reference's one
    // fread per file, CArk.cpp:796-811 -- thousands of small files are latency-bound on any file system)
    auto lReadFiles = [&](size_t gg, size_t liFrom, size_t liTo) {
        const Group& lGroup = laGroups[gg];
        Slot& lSlot = lRing.maSlots[gg % lRing.maSlots.size()];
        SlowOp lTimer("SaveArk reader task (up to 48 input files)");
        for (size_t ii = liFrom; ii < liTo && !lRing.Failed(); ++ii) {
            const Piece& lPiece = laPieces[ii];
            const modark::FileDef& lFile = mHeader.maFiles[lPiece.file];
            const int liFd = open((mBuildInputDirectory + lFile.mName).c_str(), O_RDONLY);
            if (liFd < 0) {
                lRing.Fail(eError_FailedToOpenFile);
                break;
            }
            unsigned char* lpDst = lSlot.mpHostIn + (lGroup.lo & 15u) + (laDescs[ii].src_off - lGroup.lo);
            const bool lbOk = ReadFully(liFd, lpDst, lPiece.size, lPiece.fileOffset);  // a file that shrank reads as zeros
            close(liFd);
            if (!lbOk)
                lRing.Fail(eError_FailedToOpenFile);
        }
        lRing.FillDone(lSlot);
    };
    auto lStartFill = [&](size_t gg) {
        const Group& lGroup = laGroups[gg];
        Slot& lSlot = lRing.maSlots[gg % lRing.maSlots.size()];
        const size_t kiFilesPerTask = 48;
        const int liTasks = (int)((lGroup.last - lGroup.first + kiFilesPerTask - 1) / kiFilesPerTask);
        {
            std::lock_guard<std::mutex> lLock(lRing.mMutex);
            lSlot.meState = Slot::eFilling;
            lSlot.miFillLeft = liTasks;
        }
        for (int tt = 0; tt < liTasks; ++tt) {
            const size_t liFrom = lGroup.first + (size_t)tt * kiFilesPerTask;
            lReaders.Push([&, gg, liFrom, liTo = std::min(lGroup.last, liFrom + kiFilesPerTask)]() { lReadFiles(gg, liFrom, liTo); });
        }
    };

    // stage 3 (writer threads): the group's byte range, cut where it crosses into another part file
    auto lWriteRange = [&](size_t gg, size_t liTarget, uint64_t luLo, uint64_t luHi) {
        const Group& lGroup = laGroups[gg];
        Slot& lSlot = lRing.maSlots[gg % lRing.maSlots.size()];
        SlowOp lTimer("SaveArk writer task (stream sync + a 4 MiB range)");
        if (!lRing.Failed() && lSlot.mbUsedGpu && [&]() { SlowOp lSync("SaveArk stream sync"); return mod_stream_sync(lSlot.mpStream); }() != MOD_OK) {
            std::cout << "GPU build failed: " << mod_last_error() << "\n";
            lRing.Fail(eError_InvalidData);
        }
        const unsigned char* lpBytes = (lSlot.mbUsedGpu ? lSlot.mpHostOut : lSlot.mpHostIn) + (lGroup.lo & 15u) + (luLo - lGroup.lo);
        const PartTarget& lPart = laTargets[liTarget];
        if (!lRing.Failed()) {
            if (lPart.mpMap)  // pages of a mapped file are faulted in and filled by many threads at once
                std::memcpy(lPart.mpMap + (luLo - lPart.muImageStart), lpBytes, (size_t)(luHi - luLo));
            else if (!WriteFully(lPart.miFd, lpBytes, luHi - luLo, luLo - lPart.muImageStart))
                lRing.Fail(eError_FailedToWriteData);
        }
        lRing.DrainDone(lSlot);
    };

    // stage 2 (this thread)
    size_t liNextFill = 0, liNextGpu = 0;
    while (liNextGpu < laGroups.size() && !lRing.Failed()) {
        {
            std::unique_lock<std::mutex> lLock(lRing.mMutex);
            lRing.mSignal.wait(lLock, [&]() {
                const bool lbCanFill = liNextFill < laGroups.size() && liNextFill < liNextGpu + lRing.maSlots.size() &&
                                       lRing.maSlots[liNextFill % lRing.maSlots.size()].meState == Slot::eFree;
                const bool lbCanRun = liNextGpu < liNextFill && lRing.maSlots[liNextGpu % lRing.maSlots.size()].meState == Slot::eFilled;
                return lbCanFill || lbCanRun || lRing.Failed();
            });
        }
        if (lRing.Failed())
            break;
        while (liNextFill < laGroups.size() && liNextFill < liNextGpu + lRing.maSlots.size() &&
               lRing.maSlots[liNextFill % lRing.maSlots.size()].meState == Slot::eFree)
            lStartFill(liNextFill++);
        while (liNextGpu < liNextFill && lRing.maSlots[liNextGpu % lRing.maSlots.size()].meState == Slot::eFilled) {
            const size_t gg = liNextGpu++;
            const Group& lGroup = laGroups[gg];
            Slot& lSlot = lRing.maSlots[gg % lRing.maSlots.size()];
            const uint64_t luRange = lGroup.hi - lGroup.lo;
            lSlot.mbUsedGpu = lbAnyKey;
            if (lbAnyKey) {
                size_t liDeviceIndex = 0;
                while (laDevices[liDeviceIndex] != lSlot.miDevice)
                    ++liDeviceIndex;
                uint64_t luTile0 = 0, luTile1 = 0;
                SlowOp lTimer("SaveArk GPU enqueue of a group");
                unsigned char* lpDevIn = (unsigned char*)lSlot.mpDevIn + (lGroup.lo & 15u);
                unsigned char* lpDevOut = (unsigned char*)lSlot.mpDevOut + (lGroup.lo & 15u);
                const bool lbOk =
                    mod_init(lSlot.miDevice) == MOD_OK &&
                    mod_plan_tile_range(laPlans[liDeviceIndex], lGroup.first, lGroup.last, &luTile0, &luTile1) == MOD_OK &&
                    mod_memcpy_h2d(lpDevIn, lSlot.mpHostIn + (lGroup.lo & 15u), luRange, lSlot.mpStream) == MOD_OK &&
                    mod_plan_run_window(laPlans[liDeviceIndex], luTile0, luTile1, lpDevIn, lGroup.lo, luRange, lpDevOut, lGroup.lo,
                                        luRange, lSlot.mpStream) == MOD_OK &&
                    mod_memcpy_d2h(lSlot.mpHostOut + (lGroup.lo & 15u), lpDevOut, luRange, lSlot.mpStream) == MOD_OK;
                if (!lbOk) {
                    std::cout << "GPU build failed: " << mod_last_error() << "\n";
                    lRing.Fail(eError_InvalidData);
                    break;
                }
            }
            // pieces of [lo, hi) per part file that receives them
            struct Piece {
                size_t target;
                uint64_t lo, hi;
            };
            std::vector<Piece> laPieces;
            const uint64_t kuWritePiece = 4ull << 20;  // several writer threads share a group
            for (size_t tt = 0; tt < laTargets.size(); ++tt) {
                const uint64_t luLo = std::max(lGroup.lo, laTargets[tt].muImageStart);
                const uint64_t luHi = std::min(lGroup.hi, laTargets[tt].muImageStart + laTargets[tt].muSize);
                if (laTargets[tt].miFd < 0)
                    continue;
                for (uint64_t luAt = luLo; luAt < luHi; luAt += kuWritePiece)
                    laPieces.push_back(Piece{tt, luAt, std::min(luHi, luAt + kuWritePiece)});
            }
            if (laPieces.empty()) {  // every part this group falls in was skipped (existing output kept)
                lRing.Set(lSlot, Slot::eFree);
                continue;
            }
            {
                std::lock_guard<std::mutex> lLock(lRing.mMutex);
                lSlot.meState = Slot::eDraining;
                lSlot.miDrainLeft = (int)laPieces.size();
            }
            for (const Piece& lPiece : laPieces)
                lWriters.Push([&, gg, lPiece]() { lWriteRange(gg, lPiece.target, lPiece.lo, lPiece.hi); });
        }
    }
    lReaders.Finish();
    lWriters.Finish();
    const double ldDone = NowSeconds();
    lCleanup();
    if (TraceEnabled())
        std::fprintf(stderr, "[mod] SaveArk: streamed %zu groups (%s): setup %.3f s, pipeline %.3f s, free %.3f s\n", laGroups.size(),
                     lbAnyKey ? "ciphered on the GPU" : "plain copy, no GPU", ldReady - ldStart, ldDone - ldReady, NowSeconds() - ldDone);
    return (eError)lRing.miError.load();
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

    // parts: consecutive slices of the image.  Like the reference, the image cursor does not advance
    // past a part that is skipped because its output already exists (CArk.cpp:863-868) or cannot be
    // created (:880-897).
    std::vector<PartTarget> laTargets;
    uint64_t luCursor = 0;
    for (const modark::PartDef& lPart : mHeader.maParts) {
        const std::string lFilename = std::string(lpOutputDirectory) + lPart.mPath;
        if (KeepExistingOutput(lFilename)) {
            std::cout << "Output file already exists: " << lFilename.c_str() << "\n";
            continue;
        }
        std::cout << "Writing " << lFilename.c_str() << "\n";
        MakeParentDirectories(lFilename);
        PartTarget lTarget;
        lTarget.miFd = open(lFilename.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);
        if (lTarget.miFd < 0) {
            std::cout << "Failed to open file for writing: " << lFilename.c_str() << "\n";
            continue;
        }
        lTarget.muImageStart = luCursor;
        lTarget.muSize = lPart.muSize;
        // size the file up front and map it: writer threads then fill disjoint ranges concurrently (write()
        // calls on one file serialise on its inode lock); if the mapping fails they fall back to pwrite
        if (mbBuilt && lPart.muSize && ftruncate(lTarget.miFd, (off_t)lPart.muSize) == 0) {
            void* lpMap = mmap(nullptr, lPart.muSize, PROT_READ | PROT_WRITE, MAP_SHARED, lTarget.miFd, 0);
            if (lpMap != MAP_FAILED)
                lTarget.mpMap = (unsigned char*)lpMap;
        }
        laTargets.push_back(lTarget);
        luCursor += lPart.muSize;
    }

    eError leError = eError_NoError;
    if (mbBuilt) {
        leError = StreamBuiltImage(laTargets);
    } else if (mpArkData) {  // a loaded image (Load + LoadArkData) saved again
        for (const PartTarget& lTarget : laTargets) {
            const uint64_t luTake = lTarget.muImageStart < muArkDataSize ? std::min<uint64_t>(lTarget.muSize, muArkDataSize - lTarget.muImageStart) : 0;
            if (luTake != lTarget.muSize || !WriteFully(lTarget.miFd, mpArkData + lTarget.muImageStart, luTake, 0))
                leError = eError_FailedToWriteData;
        }
    } else {
        for (const PartTarget& lTarget : laTargets)
            if (lTarget.muSize)
                leError = eError_FailedToWriteData;  // nothing to write the parts from
    }
    for (const PartTarget& lTarget : laTargets) {
        if (lTarget.mpMap)
            munmap(lTarget.mpMap, lTarget.muSize);
        close(lTarget.miFd);
    }
    SHOW_ERROR_AND_RETURN;
    return eError_NoError;
}
