#include "CDtaFile.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>

namespace {

class Reader
{
public:
    Reader(const unsigned char* lpData, size_t liSize) : mpData(lpData), miSize(liSize) {}
    bool Ok() const { return mbOk; }
    bool AtEnd() const { return miPos >= miSize; }

    template <typename T>
    T Read()
    {
        T lValue{};
        if (!mbOk || miPos > miSize || miSize - miPos < sizeof(T)) {
            mbOk = false;
            return lValue;
        }
        std::memcpy(&lValue, mpData + miPos, sizeof(T));
        miPos += sizeof(T);
        return lValue;
    }

    std::string ReadBytes(size_t liCount)
    {
        if (!mbOk || miSize - miPos < liCount) {
            mbOk = false;
            return std::string();
        }
        std::string lOut(reinterpret_cast<const char*>(mpData + miPos), liCount);
        miPos += liCount;
        return lOut;
    }

    void Skip(size_t liCount)
    {
        if (!mbOk || miSize - miPos < liCount)
            mbOk = false;
        else
            miPos += liCount;
    }

private:
    const unsigned char* mpData;
    size_t miSize;
    size_t miPos = 0;
    bool mbOk = true;
};

template <typename T>
void Put(std::vector<unsigned char>& lOut, T lValue)
{
    const unsigned char* lpBytes = reinterpret_cast<const unsigned char*>(&lValue);
    lOut.insert(lOut.end(), lpBytes, lpBytes + sizeof(T));
}

constexpr int kMaxDepth = 256;  // the reference recurses without a limit; a corrupt file must not smash the stack

eError ReadTree(Reader& lIn, SDtaNode& lTree, int liDepth)
{
    if (liDepth > kMaxDepth)
        return eError_InvalidData;
    const int16_t lsNumChildren = lIn.Read<int16_t>();
    if (!lIn.Ok() || lsNumChildren <= 0)
        return eError_InvalidData;
    lTree.msNodeId = lIn.Read<int16_t>();
    for (int ii = 0; ii < lsNumChildren; ++ii) {
        std::unique_ptr<SDtaNode> lpChild(new SDtaNode());
        lpChild->mpParent = &lTree;
        lpChild->miType = lIn.Read<int32_t>();
        if (!lIn.Ok())
            return eError_InvalidData;
        if (lpChild->IsString()) {
            const int32_t liLength = lIn.Read<int32_t>();
            if (!lIn.Ok() || liLength < 0)
                return eError_InvalidData;
            lpChild->mString = lIn.ReadBytes((size_t)liLength);
            // the reference builds a std::string from a NUL-terminated copy: an embedded NUL ends the value
            const size_t liNul = lpChild->mString.find('\0');
            if (liNul != std::string::npos)
                lpChild->mString.resize(liNul);
        } else if (lpChild->IsTree()) {
            lIn.Skip(sizeof(int32_t));
            const eError leError = ReadTree(lIn, *lpChild, liDepth + 1);
            if (leError != eError_NoError)
                return leError;
        } else if (lpChild->IsInteger()) {
            lpChild->miValue = lIn.Read<int32_t>();
        } else if (lpChild->miType == ENodeType_Float) {
            lpChild->mfValue = lIn.Read<float>();
        } else {
            return eError_InvalidData;
        }
        if (!lIn.Ok())
            return eError_InvalidData;
        lTree.maChildren.push_back(std::move(lpChild));
    }
    return eError_NoError;
}

void WriteTree(std::vector<unsigned char>& lOut, const SDtaNode& lTree)
{
    Put<int16_t>(lOut, (int16_t)lTree.maChildren.size());
    Put<int16_t>(lOut, lTree.msNodeId);
    for (const auto& lpChild : lTree.maChildren) {
        Put<int32_t>(lOut, lpChild->miType);
        if (lpChild->IsTree()) {
            Put<int32_t>(lOut, 1);
            WriteTree(lOut, *lpChild);
        } else if (lpChild->IsString()) {
            Put<int32_t>(lOut, (int32_t)lpChild->mString.size());
            lOut.insert(lOut.end(), lpChild->mString.begin(), lpChild->mString.end());
        } else if (lpChild->miType == ENodeType_Float) {
            Put<float>(lOut, lpChild->mfValue);
        } else {
            Put<int32_t>(lOut, lpChild->miValue);
        }
    }
}

SDtaNode* NextSibling(SDtaNode* lpNode)
{
    SDtaNode* lpParent = lpNode ? lpNode->mpParent : nullptr;
    if (!lpParent)
        return nullptr;
    for (size_t ii = 0; ii + 1 < lpParent->maChildren.size(); ++ii)
        if (lpParent->maChildren[ii].get() == lpNode)
            return lpParent->maChildren[ii + 1].get();
    return nullptr;
}

}  // namespace

SDtaNode* SDtaNode::FindNode(const std::string& lName)
{
    for (auto& lpChild : maChildren) {
        if (!lpChild->IsTree()) {
            if (lpChild->IsString() && lpChild->mString == lName)
                return lpChild.get();
            continue;
        }
        if (SDtaNode* lpFound = lpChild->FindNode(lName))
            return lpFound;
    }
    return nullptr;
}

eError CDtaFile::LoadFromMemory(const unsigned char* lpData, size_t liSize)
{
    maTrees.clear();
    Reader lIn(lpData, liSize);
    lIn.Skip(5);  // 0x01, i32 1
    if (!lIn.Ok())
        return eError_InvalidData;
    int liType = ENodeType_Tree1;
    while (!lIn.AtEnd()) {
        if (liType != ENodeType_Tree1 && liType != ENodeType_Tree2)
            return eError_InvalidData;
        std::unique_ptr<SDtaNode> lpTree(new SDtaNode());
        lpTree->miType = liType;
        const eError leError = ReadTree(lIn, *lpTree, 0);
        if (leError != eError_NoError)
            return leError;
        maTrees.push_back(std::move(lpTree));
        if (lIn.AtEnd())
            break;
        liType = lIn.Read<int32_t>();
        lIn.Skip(sizeof(int32_t));
        if (!lIn.Ok())
            return eError_InvalidData;
    }
    return eError_NoError;
}

std::vector<unsigned char> CDtaFile::SaveToMemory() const
{
    std::vector<unsigned char> lOut;
    lOut.push_back(1);
    Put<int32_t>(lOut, 1);
    // Like the reference's Save (CDtaFile.cpp:371-374) the top-level trees are streamed back to back,
    // WITHOUT the (type, 1) words Load consumes between them (:93-94): a file with several top-level
    // trees does not round-trip through the reference, and does not here either -- byte parity with
    // what the reference writes (tests/golden/dtb/two_trees) wins over repairing its format.
    for (const auto& lpTree : maTrees)
        WriteTree(lOut, *lpTree);
    return lOut;
}

eError CDtaFile::Load(const char* lpFilename)
{
    FILE* lpInputFile = std::fopen(lpFilename, "rb");
    if (!lpInputFile) {
        eError leError = eError_FailedToOpenFile;
        SHOW_ERROR_AND_RETURN;
    }
    std::fseek(lpInputFile, 0, SEEK_END);
    const long liFileSize = std::ftell(lpInputFile);
    std::fseek(lpInputFile, 0, SEEK_SET);
    std::vector<unsigned char> lData(liFileSize > 0 ? (size_t)liFileSize : 0);
    const size_t liRead = lData.empty() ? 0 : std::fread(lData.data(), 1, lData.size(), lpInputFile);
    std::fclose(lpInputFile);
    eError leError = liRead == lData.size() ? LoadFromMemory(lData.data(), lData.size()) : eError_InvalidData;
    SHOW_ERROR_AND_RETURN;
    return eError_NoError;
}

eError CDtaFile::Save(const char* lpFilename) const
{
    const std::vector<unsigned char> lData = SaveToMemory();
    FILE* lpOutputFile = std::fopen(lpFilename, "wb");
    if (!lpOutputFile) {
        eError leError = eError_FailedToCreateFile;
        SHOW_ERROR_AND_RETURN;
    }
    const size_t liWritten = std::fwrite(lData.data(), 1, lData.size(), lpOutputFile);
    std::fclose(lpOutputFile);
    return liWritten == lData.size() ? eError_NoError : eError_FailedToWriteData;
}

SDtaNode* CDtaFile::FindNode(const std::string& lName)
{
    for (auto& lpTree : maTrees)
        if (SDtaNode* lpFound = lpTree->FindNode(lName))
            return lpFound;
    return nullptr;
}

std::vector<SSongConfig> CDtaFile::GetSongs() const
{
    std::vector<SSongConfig> laSongs;
    const SDtaNode* lpTokens = FindNode("unlock_tokens");
    if (!lpTokens || !lpTokens->mpParent)
        return laSongs;
    for (const auto& lpSongNode : lpTokens->mpParent->maChildren) {
        // {id, name, unlock title, unlock description, icon, type}
        if (!lpSongNode->IsTree() || lpSongNode->maChildren.size() != 6)
            continue;
        const SDtaNode& lId = *lpSongNode->maChildren[0];
        const SDtaNode& lName = *lpSongNode->maChildren[1];
        if (!lId.IsString() || !lName.IsString())
            continue;
        // songs are the records whose id has no lower-case letter
        if (std::any_of(lId.mString.begin(), lId.mString.end(), [](char c) { return c != (char)std::toupper(c); }))
            continue;
        SSongConfig lConfig;
        lConfig.mId = lId.mString;
        lConfig.mName = lName.mString;
        laSongs.push_back(lConfig);
    }
    const SDtaNode* lpCampaign = FindNode("campaign");
    if (!lpCampaign || !lpCampaign->mpParent)
        return laSongs;
    for (const auto& lpUnlockNode : lpCampaign->mpParent->maChildren) {
        // {method, count, type, unlocked item id}
        if (!lpUnlockNode->IsTree() || lpUnlockNode->maChildren.size() != 4)
            continue;
        const SDtaNode& lMethod = *lpUnlockNode->maChildren[0];
        const SDtaNode& lCount = *lpUnlockNode->maChildren[1];
        const SDtaNode& lItem = *lpUnlockNode->maChildren[3];
        if (!lMethod.IsString() || !lCount.IsInteger() || !lItem.IsString())
            continue;
        for (SSongConfig& lSong : laSongs) {
            if (lSong.mId == lItem.mString) {
                lSong.mUnlockMethod = lMethod.mString;
                lSong.miUnlockCount = lCount.miValue;
            }
        }
    }
    return laSongs;
}

void CDtaFile::GetSongData(std::vector<SSongConfig>& laSongs) const
{
    for (SSongConfig& lSong : laSongs) {
        const SDtaNode* lpIdNode = FindNode(lSong.mId);
        if (!lpIdNode)
            continue;
        const SDtaNode* lpSongNode = lpIdNode->mpParent;  // {id, path, (type <type>)}
        if (!lpSongNode || lpSongNode->maChildren.size() < 3)
            continue;
        const SDtaNode* lpArenaNode = lpSongNode->mpParent;
        if (lpArenaNode && !lpArenaNode->maChildren.empty() && lpArenaNode->maChildren[0]->IsString())
            lSong.mArena = lpArenaNode->maChildren[0]->mString;
        const SDtaNode& lPath = *lpSongNode->maChildren[1];
        if (!lPath.IsString())
            continue;
        lSong.mPath = lPath.mString;
        std::transform(lSong.mPath.begin(), lSong.mPath.end(), lSong.mPath.begin(),
                       [](unsigned char c) { return (char)std::tolower(c); });
        const SDtaNode& lTypeTree = *lpSongNode->maChildren[2];
        if (lTypeTree.maChildren.size() < 2 || !lTypeTree.maChildren.back()->IsString())
            continue;
        lSong.mType = lTypeTree.maChildren.back()->mString;
    }
}

bool CDtaFile::SetIntAfter(const std::string& lKey, int32_t liValue)
{
    SDtaNode* lpValue = NextSibling(FindNode(lKey));
    if (!lpValue || !lpValue->IsInteger())
        return false;
    lpValue->miValue = liValue;
    return true;
}

bool CDtaFile::SetStringAfter(const std::string& lKey, const std::string& lValue)
{
    SDtaNode* lpValue = NextSibling(FindNode(lKey));
    if (!lpValue || !lpValue->IsString())
        return false;
    lpValue->mString = lValue;
    return true;
}
