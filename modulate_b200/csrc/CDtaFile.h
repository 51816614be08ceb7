// CDtaFile.h -- host-side codec for binary DTA ("DTB") script data: the step next to the hot path
// that BASELINE config 5 needs ("patch DTA on host" between extract and repack).  SURVEY.md 8(f)
// row 3: branchy recursive tree work on KB-sized inputs -- it stays on the CPU by design.
//
// On-disk layout, as read by the reference's CDtaFile::Load / AddTreeNode (CDtaFile.cpp:57-100,
// :393-509) and written by Save / SaveToStream (CDtaFile.cpp:362-391, :1302-1326, CDtaFile.h:262-284):
//   u8 1, i32 1                                     5-byte prefix (skipped on load)
//   tree := i16 nChildren (> 0), i16 nodeId, child[nChildren]
//   child := i32 type, payload
//     0 / 6 / 8 / 9   i32 value                     (four integer flavours, the type is preserved)
//     1               f32 value
//     5 / 18 / 33 / 35  i32 length, bytes           (string, id, include-file, define)
//     16 / 17         i32 (written as 1), tree      (two sub-tree flavours)
//   further top-level trees follow as: i32 type (16 | 17), i32, tree
//
// Same class name and Load/Save signatures as the reference (CDtaFile.h:169-188), plus the two
// read-only song queries `-pack` depends on (GetSongs / GetSongData, CDtaFile.cpp:102-181, :248-294:
// they decide which /songs/ folders a repack keeps).  The song-list EDITING commands (SetSongs,
// RemoveSong, UpdateSongData) are tool policy outside the scope of this repo.  The node model here
// is a tagged value tree, not the reference's class hierarchy.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "Error.h"

enum eNodeType {
    ENodeType_Integer0 = 0,
    ENodeType_Float = 1,
    ENodeType_String = 5,
    ENodeType_Integer6 = 6,
    ENodeType_Integer8 = 8,
    ENodeType_Integer9 = 9,
    ENodeType_Tree1 = 16,
    ENodeType_Tree2 = 17,
    ENodeType_Id = 18,
    ENodeType_IncludeFile = 33,
    ENodeType_Define = 35,
    ENodeType_Invalid
};

// Same fields as the reference's SSongConfig (CDtaFile.h:25-34).
struct SSongConfig {
    std::string mId = "";
    std::string mName = "";
    std::string mUnlockMethod = "";
    std::string mType = "";
    std::string mPath = "";
    std::string mArena = "";
    int miUnlockCount = -1;
};

struct SDtaNode {
    int miType = ENodeType_Tree1;  // eNodeType value as stored in the file
    int32_t miValue = 0;           // integer flavours
    float mfValue = 0.0f;          // ENodeType_Float
    std::string mString;           // string flavours
    int16_t msNodeId = 0;          // tree flavours
    std::vector<std::unique_ptr<SDtaNode>> maChildren;
    SDtaNode* mpParent = nullptr;

    bool IsTree() const { return miType == ENodeType_Tree1 || miType == ENodeType_Tree2; }
    bool IsString() const
    {
        return miType == ENodeType_String || miType == ENodeType_Id || miType == ENodeType_IncludeFile ||
               miType == ENodeType_Define;
    }
    bool IsInteger() const
    {
        return miType == ENodeType_Integer0 || miType == ENodeType_Integer6 || miType == ENodeType_Integer8 ||
               miType == ENodeType_Integer9;
    }
    // Depth-first search for the first string-flavoured node equal to lName (the lookup the
    // reference's tools are built on, CDtaFile.cpp:32-55).
    SDtaNode* FindNode(const std::string& lName);
};

class CDtaFile
{
public:
    eError Load(const char* lpFilename);
    eError Save(const char* lpFilename) const;

    eError LoadFromMemory(const unsigned char* lpData, size_t liSize);
    std::vector<unsigned char> SaveToMemory() const;

    // top-level trees of the file, in order
    std::vector<std::unique_ptr<SDtaNode>>& Trees() { return maTrees; }
    const std::vector<std::unique_ptr<SDtaNode>>& Trees() const { return maTrees; }
    SDtaNode* FindNode(const std::string& lName);
    const SDtaNode* FindNode(const std::string& lName) const { return const_cast<CDtaFile*>(this)->FindNode(lName); }

    // The songs amp_config lists (ids in upper case, six-field records next to "unlock_tokens") with
    // their unlock data (four-field records next to "campaign"): reference CDtaFile.cpp:102-181.
    // Where the reference dereferences a null pointer on a config without those nodes, this returns
    // an empty list.
    std::vector<SSongConfig> GetSongs() const;
    // Fill mPath (lower-cased), mArena and mType of each song from amp_songs_config: CDtaFile.cpp:248-294.
    void GetSongData(std::vector<SSongConfig>& laSongs) const;

    // Convenience for host-side patches: the value that FOLLOWS the string node `lKey` inside its
    // parent tree, i.e. the `(key value)` idiom of DTA.  Returns false if there is no such pair.
    bool SetIntAfter(const std::string& lKey, int32_t liValue);
    bool SetStringAfter(const std::string& lKey, const std::string& lValue);

private:
    std::vector<std::unique_ptr<SDtaNode>> maTrees;
};
