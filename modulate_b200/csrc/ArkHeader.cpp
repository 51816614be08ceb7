#include "ArkHeader.h"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <numeric>
#include <unordered_map>

#include "Settings.h"

namespace modark {

namespace {

// Bounds-checked little-endian cursor over the decrypted image.
class Cursor
{
public:
    Cursor(const unsigned char* lpData, size_t liSize, size_t liPos) : mpData(lpData), miSize(liSize), miPos(liPos) {}

    bool Ok() const { return mbOk; }

    template <typename T>
    T Read()
    {
        T lValue{};
        if (!mbOk || miSize - miPos < sizeof(T) || miPos > miSize) {
            mbOk = false;
            return lValue;
        }
        std::memcpy(&lValue, mpData + miPos, sizeof(T));
        miPos += sizeof(T);
        return lValue;
    }

    void Skip(uint64_t liBytes)
    {
        if (!mbOk || liBytes > miSize - miPos) {
            mbOk = false;
            return;
        }
        miPos += (size_t)liBytes;
    }

    // i32 length + bytes.  Like the reference's readers (CArk.cpp:533-551, :614-632) the value keeps
    // at most 255 characters and stops at an embedded NUL, while the cursor moves by the full length.
    std::string ReadString()
    {
        const int32_t liLength = Read<int32_t>();
        if (!mbOk || liLength < 0 || (uint64_t)liLength > miSize - miPos) {
            mbOk = false;
            return std::string();
        }
        const size_t liKeep = std::min<size_t>((size_t)liLength, 255);
        const char* lpChars = reinterpret_cast<const char*>(mpData + miPos);
        const size_t liVisible = strnlen(lpChars, liKeep);
        miPos += (size_t)liLength;
        return std::string(lpChars, liVisible);
    }

private:
    const unsigned char* mpData;
    size_t miSize;
    size_t miPos;
    bool mbOk = true;
};

template <typename T>
void Append(std::vector<unsigned char>& lOut, T lValue)
{
    const unsigned char* lpBytes = reinterpret_cast<const unsigned char*>(&lValue);
    lOut.insert(lOut.end(), lpBytes, lpBytes + sizeof(T));
}

void AppendString(std::vector<unsigned char>& lOut, const std::string& lValue)
{
    Append<int32_t>(lOut, (int32_t)lValue.size());
    lOut.insert(lOut.end(), lValue.begin(), lValue.end());
}

int CompareNoCase(const std::string& lA, const std::string& lB)
{
    const size_t liCommon = std::min(lA.size(), lB.size());
    for (size_t ii = 0; ii < liCommon; ++ii) {
        const int liA = std::tolower((unsigned char)lA[ii]);
        const int liB = std::tolower((unsigned char)lB[ii]);
        if (liA != liB)
            return liA < liB ? -1 : 1;
    }
    if (lA.size() == lB.size())
        return 0;
    return lA.size() < lB.size() ? -1 : 1;
}

std::vector<std::string> SplitPath(const std::string& lName)
{
    std::vector<std::string> lParts;
    size_t liStart = 0;
    for (;;) {
        const size_t liSlash = lName.find('/', liStart);
        if (liSlash == std::string::npos) {
            lParts.push_back(lName.substr(liStart));
            return lParts;
        }
        lParts.push_back(lName.substr(liStart, liSlash - liStart));
        liStart = liSlash + 1;
    }
}

}  // namespace

eError ParseHeader(const unsigned char* lpData, size_t liSize, HeaderImage& lOut)
{
    if (liSize < sizeof(uint32_t))
        return eError_InvalidData;
    uint32_t luMagic = 0;
    std::memcpy(&luMagic, lpData, sizeof(luMagic));
    if (luMagic != CSettings::kuEncryptedVersionPS3 && luMagic != CSettings::kuEncryptedVersionPS4)
        return eError_UnknownVersionNumber;
    lOut = HeaderImage();
    lOut.mbPS4 = (luMagic == CSettings::kuEncryptedVersionPS4);

    Cursor lIn(lpData, liSize, sizeof(uint32_t));
    lIn.Read<uint32_t>();  // version (the reference does not check it either)
    lIn.Read<uint32_t>();  // number of checksums
    lIn.Skip(16);          // checksum bytes
    const int32_t liNumArks = lIn.Read<int32_t>();
    if (!lIn.Ok())
        return eError_InvalidData;
    if (liNumArks < 0 || liNumArks > kMaxArks)
        return eError_ValueOutOfBounds;

    const int32_t liNumSizes = lIn.Read<int32_t>();
    if (!lIn.Ok() || liNumSizes < 0)
        return eError_InvalidData;
    if (liNumSizes < liNumArks)
        return eError_ValueOutOfBounds;  // sIntList::GetValue past miNum
    lOut.maParts.resize((size_t)liNumArks);
    for (int32_t ii = 0; ii < liNumSizes; ++ii) {
        const uint32_t luSize = lIn.Read<uint32_t>();
        if (ii < liNumArks)
            lOut.maParts[(size_t)ii].muSize = luSize;
    }
    const int32_t liNumPaths = lIn.Read<int32_t>();
    if (!lIn.Ok() || liNumPaths < 0)
        return eError_InvalidData;
    if (liNumPaths < liNumArks)
        return eError_ValueOutOfBounds;
    for (int32_t ii = 0; ii < liNumPaths; ++ii) {
        std::string lPath = lIn.ReadString();
        if (ii < liNumArks)
            lOut.maParts[(size_t)ii].mPath = lPath;
    }
    const int32_t liNumChecksums = lIn.Read<int32_t>();
    if (!lIn.Ok() || liNumChecksums < 0)
        return eError_InvalidData;
    lIn.Skip(4ull * (uint64_t)liNumChecksums);  // the checksum list itself
    lIn.Skip(4ull * (uint64_t)liNumChecksums);  // the word count + all but the last word of the next list
    const int32_t liMustBeZero = lIn.Read<int32_t>();
    if (!lIn.Ok() || liMustBeZero != 0)
        return eError_InvalidData;

    const int32_t liNumFiles = lIn.Read<int32_t>();
    if (!lIn.Ok())
        return eError_InvalidData;
    if (liNumFiles < 0 || liNumFiles > kMaxFiles)
        return eError_ValueOutOfBounds;
    lOut.maFiles.resize((size_t)liNumFiles);
    for (FileDef& lFile : lOut.maFiles) {
        lFile.mi64Offset = lIn.Read<int64_t>();
        lFile.mName = lIn.ReadString();
        lFile.miFlags1 = lIn.Read<int32_t>();
        lFile.miSize = (int)lIn.Read<uint32_t>();
        lFile.muHash = lIn.Read<uint32_t>();
        if (!lIn.Ok() || lFile.mName.empty() || lFile.mi64Offset < 0 || lFile.miSize < 0)
            return eError_InvalidData;
    }
    const int32_t liNumFlags2 = lIn.Read<int32_t>();
    if (!lIn.Ok())
        return eError_InvalidData;
    if (liNumFlags2 < liNumFiles)
        return liNumFiles ? eError_ValueOutOfBounds : eError_NoError;
    for (FileDef& lFile : lOut.maFiles)
        lFile.miFlags2 = lIn.Read<int32_t>();
    return lIn.Ok() ? eError_NoError : eError_InvalidData;
}

int NameBucket(const std::string& lName, int liNumFiles)
{
    // Rolling hash in 32-bit signed arithmetic, reduced modulo the file count after every character
    // (the bucket table the game walks has one slot per file).  An empty name still folds in its
    // terminating NUL once.
    int32_t liHash = 0;
    const char* lpChar = lName.c_str();
    do {
        liHash = (int32_t)((uint32_t)liHash * 0x7Fu + (uint32_t)(int32_t)(signed char)*lpChar);
        liHash %= liNumFiles;
    } while (*(++lpChar));
    return liHash;
}

std::vector<unsigned char> SerialiseHeader(const HeaderImage& lHeader)
{
    const int liNumArks = (int)lHeader.maParts.size();
    const int liNumFiles = (int)lHeader.maFiles.size();
    std::vector<unsigned char> lOut;
    lOut.reserve(64 + (size_t)liNumFiles * 48);

    Append<uint32_t>(lOut, lHeader.mbPS4 ? CSettings::kuEncryptedVersionPS4 : CSettings::kuEncryptedVersionPS3);
    Append<uint32_t>(lOut, 9u);
    Append<uint32_t>(lOut, 1u);
    lOut.insert(lOut.end(), 16, 0);
    Append<int32_t>(lOut, liNumArks);

    Append<int32_t>(lOut, liNumArks);
    for (const PartDef& lPart : lHeader.maParts)
        Append<uint32_t>(lOut, lPart.muSize);
    Append<int32_t>(lOut, liNumArks);
    for (const PartDef& lPart : lHeader.maParts)
        AppendString(lOut, lPart.mPath);
    for (int liList = 0; liList < 2; ++liList) {  // checksums, then string counts: all zero
        Append<int32_t>(lOut, liNumArks);
        lOut.insert(lOut.end(), 4 * (size_t)liNumArks, 0);
    }
    Append<int32_t>(lOut, liNumFiles);

    // Entry order.  PS3: by name-hash bucket, ties by position.  PS4: path order -- at every depth
    // files come before sub-directories, names compare case-insensitively, ties by flags then
    // position.  (The reference's PS4 comparator, CArk.cpp:969-1045, is not a strict weak ordering,
    // so its exact permutation is whatever MSVC's std::sort does with it; this one is the total
    // order that agrees with it wherever it is consistent.)
    std::vector<int> laBucket((size_t)liNumFiles);
    for (int ii = 0; ii < liNumFiles; ++ii)
        laBucket[(size_t)ii] = NameBucket(lHeader.maFiles[(size_t)ii].mName, liNumFiles);
    std::vector<int> laOrder((size_t)liNumFiles);
    std::iota(laOrder.begin(), laOrder.end(), 0);
    if (lHeader.mbPS4) {
        std::vector<std::vector<std::string>> laPaths((size_t)liNumFiles);
        for (int ii = 0; ii < liNumFiles; ++ii)
            laPaths[(size_t)ii] = SplitPath(lHeader.maFiles[(size_t)ii].mName);
        std::stable_sort(laOrder.begin(), laOrder.end(), [&](int liA, int liB) {
            const std::vector<std::string>& lA = laPaths[(size_t)liA];
            const std::vector<std::string>& lB = laPaths[(size_t)liB];
            for (size_t liDepth = 0;; ++liDepth) {
                const bool lbALeaf = liDepth + 1 == lA.size();
                const bool lbBLeaf = liDepth + 1 == lB.size();
                if (lbALeaf != lbBLeaf)
                    return lbALeaf;
                const int liCmp = CompareNoCase(lA[liDepth], lB[liDepth]);
                if (liCmp != 0)
                    return liCmp < 0;
                if (lbALeaf) {
                    const FileDef& lFA = lHeader.maFiles[(size_t)liA];
                    const FileDef& lFB = lHeader.maFiles[(size_t)liB];
                    if (lFA.miFlags1 != lFB.miFlags1)
                        return lFA.miFlags1 < lFB.miFlags1;
                    if (lFA.miFlags2 != lFB.miFlags2)
                        return lFA.miFlags2 < lFB.miFlags2;
                    return false;
                }
            }
        });
    } else {
        std::stable_sort(laOrder.begin(), laOrder.end(),
                         [&](int liA, int liB) { return laBucket[(size_t)liA] < laBucket[(size_t)liB]; });
    }

    // Thread the bucket chains while emitting: flags1 = position of the previous entry of the same
    // bucket (or -1), and remember per bucket the position a chain walk starts from.  A bucket whose
    // entries are not contiguous in the order (possible on PS4) is re-entered through the FIRST
    // record made for it, exactly as the reference does (CArk.cpp:1064-1110).
    std::vector<std::pair<int, int>> laChains;  // (bucket, position)
    std::unordered_map<int, size_t> lFirstRecord;  // bucket -> index of its first record in laChains
    int liPrevious = -1;
    for (int liPos = 0; liPos < liNumFiles; ++liPos) {
        const FileDef& lFile = lHeader.maFiles[(size_t)laOrder[(size_t)liPos]];
        const int liBucket = laBucket[(size_t)laOrder[(size_t)liPos]];
        int liFlags = -1;
        const auto lFound = lFirstRecord.find(liBucket);
        if (lFound != lFirstRecord.end()) {
            liFlags = laChains[lFound->second].second;
            laChains[lFound->second].second = liPos;
        }
        if (liPrevious != -1)
            liFlags = liPrevious;
        Append<int64_t>(lOut, lFile.mi64Offset);
        AppendString(lOut, lFile.mName);
        Append<int32_t>(lOut, liFlags);
        Append<uint32_t>(lOut, (uint32_t)lFile.miSize);
        Append<uint32_t>(lOut, lFile.miSize ? (lHeader.mbPS4 ? 0xDDB682F0u : 0x7D401F60u) : 0u);
        liPrevious = liPos;
        const bool lbRunEnds = (liPos + 1 == liNumFiles) || laBucket[(size_t)laOrder[(size_t)liPos + 1]] != liBucket;
        if (lbRunEnds) {
            lFirstRecord.emplace(liBucket, laChains.size());  // no-op if the bucket already has a record
            laChains.emplace_back(liBucket, liPos);
            liPrevious = -1;
        }
    }

    Append<int32_t>(lOut, liNumFiles);
    for (int ii = 0; ii < liNumFiles; ++ii) {  // first record per bucket wins
        const auto lFound = lFirstRecord.find(ii);
        Append<int32_t>(lOut, lFound == lFirstRecord.end() ? -1 : laChains[lFound->second].second);
    }
    return lOut;
}

}  // namespace modark
