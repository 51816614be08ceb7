// modulate_main.cpp -- portable command line over the facade: the reference's archive commands
// (-ps3 -verbose -force -packall -unpack -pack -pack_add -decode; Modulate.cpp:45-70, :291-317,
// :380-502, :895-972) with the same left-to-right command deque and exit convention.  -pack reads
// the song list from the DTA configs like the reference; the song-list EDITING commands of the
// reference are host-side tooling outside the hot path and are not provided.
//
// Extensions: -bodykey K (cipher every entry body with key K, stream restarting per entry -- the
// synthetic per-entry-key configurations), -device N (bind GPU N).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <iostream>
#include <string>
#include <strings.h>
#include <vector>

#include "../../../include/modulate_b200.h"
#include "../CArk.h"
#include "../CDtaFile.h"
#include "../CEncryptionCycler.h"
#include "../Error.h"
#include "../Settings.h"

namespace {

int giBodyKey = 0;

std::string HeaderName() { return std::string("main_") + CSettings::msPlatform + ".hdr"; }

void WithSlash(std::string& lPath)
{
    if (lPath.empty() || (lPath.back() != '/' && lPath.back() != '\\'))
        lPath += "/";
}

eError PS3(std::deque<std::string>&)
{
    CSettings::mbPS4 = false;
    CSettings::msPlatform = "ps3";
    return eError_NoError;
}

eError EnableVerbose(std::deque<std::string>&)
{
    CSettings::mbVerbose = true;
    return eError_NoError;
}

eError EnableForceWrite(std::deque<std::string>&)
{
    CSettings::mbOverwriteOutputFiles = true;
    return eError_NoError;
}

eError EnablePackAll(std::deque<std::string>&)
{
    CSettings::mbPackAllFiles = true;
    return eError_NoError;
}

eError BodyKey(std::deque<std::string>& laParams)
{
    if (laParams.empty())
        return eError_InvalidParameter;
    giBodyKey = (int)std::strtoll(laParams.front().c_str(), nullptr, 0);
    laParams.pop_front();
    return eError_NoError;
}

eError Device(std::deque<std::string>& laParams)
{
    if (laParams.empty())
        return eError_InvalidParameter;
    const int liDevice = std::atoi(laParams.front().c_str());
    laParams.pop_front();
    if (mod_init(liDevice) != MOD_OK) {
        std::cout << mod_last_error() << "\n";
        return eError_InvalidParameter;
    }
    return eError_NoError;
}

eError Unpack(std::deque<std::string>& laParams)
{
    std::cout << "Unpacking " << HeaderName() << " to ";
    if (laParams.empty())
        return eError_InvalidParameter;
    std::string lOutputDirectory = laParams.front();
    laParams.pop_front();
    std::cout << lOutputDirectory << "\n";
    WithSlash(lOutputDirectory);

    const bool lbTrace = std::getenv("MOD_TRACE") != nullptr;
    const auto lNow = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double ldStart = lNow();
    CArk lArkHeader;
    lArkHeader.SetUniformEntryKey(giBodyKey);
    eError leError = lArkHeader.Load(HeaderName().c_str());
    SHOW_ERROR_AND_RETURN;
    const double ldLoaded = lNow();
    leError = lArkHeader.ExtractFiles(0, lArkHeader.GetNumFiles(), lOutputDirectory.c_str());
    SHOW_ERROR_AND_RETURN;
    if (lbTrace)
        std::fprintf(stderr, "[mod] Unpack: Load (CUDA init + header Cycle + parse) %.3f s, ExtractFiles %.3f s\n",
                     ldLoaded - ldStart, lNow() - ldLoaded);
    return eError_NoError;
}

eError PackImpl(std::deque<std::string>& laParams)
{
    std::cout << "Packing " << HeaderName() << " from ";
    if (laParams.empty())
        return eError_InvalidParameter;
    std::string lInputPath = laParams.front();
    laParams.pop_front();
    std::cout << lInputPath << " to ";
    if (laParams.empty())
        return eError_InvalidParameter;
    std::string lOutputPath = laParams.front();
    laParams.pop_front();
    std::cout << lOutputPath << "\n";
    WithSlash(lInputPath);
    WithSlash(lOutputPath);

    // Unless -packall is given the song list comes from the two DTA configs inside the input tree,
    // like the reference (Modulate.cpp:410-432): it decides which /songs/<name>/ folders are packed.
    eError leError = eError_NoError;
    std::vector<SSongConfig> lSongs;
    if (!CSettings::mbPackAllFiles) {
        const std::string lPlatform = CSettings::msPlatform;
        const std::string lAmpConfigPath = lInputPath + lPlatform + "/config/amp_config.dta_dta_" + lPlatform;
        std::cout << "Loading " << lAmpConfigPath << "\n";
        CDtaFile lAmpConfig;
        leError = lAmpConfig.Load(lAmpConfigPath.c_str());
        SHOW_ERROR_AND_RETURN;
        lSongs = lAmpConfig.GetSongs();
        CDtaFile lSongsConfig;
        const std::string lAmpSongsConfigPath = lInputPath + lPlatform + "/config/amp_songs_config.dta_dta_" + lPlatform;
        leError = lSongsConfig.Load(lAmpSongsConfigPath.c_str());
        SHOW_ERROR_AND_RETURN;
        lSongsConfig.GetSongData(lSongs);
    }

    CArk lReferenceArkHeader;
    leError = lReferenceArkHeader.Load(HeaderName().c_str());
    SHOW_ERROR_AND_RETURN;

    CArk lArkHeader;
    lArkHeader.SetUniformEntryKey(giBodyKey);
    leError = lArkHeader.ConstructFromDirectory(lInputPath.c_str(), lReferenceArkHeader, lSongs);
    SHOW_ERROR_AND_RETURN;
    leError = lArkHeader.BuildArk(lInputPath.c_str(), lSongs);
    SHOW_ERROR_AND_RETURN;
    leError = lArkHeader.SaveArk(lOutputPath.c_str(), HeaderName().c_str());
    SHOW_ERROR_AND_RETURN;
    return eError_NoError;
}

eError Pack(std::deque<std::string>& laParams) { return PackImpl(laParams); }

eError AddPack(std::deque<std::string>& laParams)
{
    CSettings::mbIgnoreNewFiles = false;  // reference Modulate.cpp:580-586
    return PackImpl(laParams);
}

eError Decode(std::deque<std::string>&)
{
    const std::string lHeaderFilename = HeaderName();
    FILE* lpHeaderFile = std::fopen(lHeaderFilename.c_str(), "rb");
    if (!lpHeaderFile)
        return eError_FailedToOpenFile;
    std::fseek(lpHeaderFile, 0, SEEK_END);
    const long liHeaderSize = std::ftell(lpHeaderFile);
    std::fseek(lpHeaderFile, 0, SEEK_SET);
    std::vector<unsigned char> lData((size_t)std::max(0l, liHeaderSize));
    const size_t liRead = lData.empty() ? 0 : std::fread(lData.data(), 1, lData.size(), lpHeaderFile);
    std::fclose(lpHeaderFile);
    if (liRead != lData.size() || lData.size() < 4)
        return eError_UnknownVersionNumber;

    unsigned int luVersion = 0;
    std::memcpy(&luVersion, lData.data(), 4);
    if (luVersion != CSettings::kuEncryptedVersionPS3 && luVersion != CSettings::kuEncryptedVersionPS4)
        return eError_UnknownVersionNumber;
    const unsigned int kuInitialKey =
        (luVersion == CSettings::kuEncryptedVersionPS3) ? CSettings::kuEncryptedPS3Key : CSettings::kuEncryptedPS4Key;
    CEncryptionCycler lDecrypt;
    lDecrypt.Cycle(lData.data() + 4, (unsigned int)(lData.size() - 4), (int)kuInitialKey);

    const std::string lOut = lHeaderFilename + ".dec";
    FILE* lpOut = std::fopen(lOut.c_str(), "wb");
    if (!lpOut)
        return eError_FailedToCreateFile;
    const size_t liWritten = std::fwrite(lData.data(), 1, lData.size(), lpOut);
    std::fclose(lpOut);
    return liWritten == lData.size() ? eError_NoError : eError_FailedToWriteData;
}

// -listsongs <dir>: the reference's read-only song listing (Modulate.cpp:504-547), same output format.
eError ListSongs(std::deque<std::string>& laParams)
{
    std::cout << "Loading ";
    if (laParams.empty())
        return eError_InvalidParameter;
    std::string lBasePath = laParams.front();
    laParams.pop_front();
    WithSlash(lBasePath);
    const std::string lPlatform = CSettings::msPlatform;
    const std::string lAmpConfigPath = lBasePath + lPlatform + "/config/amp_config.dta_dta_" + lPlatform;
    std::cout << lAmpConfigPath << "\n";
    CDtaFile lAmpConfig;
    eError leError = lAmpConfig.Load(lAmpConfigPath.c_str());
    SHOW_ERROR_AND_RETURN;
    const std::string lAmpSongsConfigPath = lBasePath + lPlatform + "/config/amp_songs_config.dta_dta_" + lPlatform;
    std::cout << "Loading " << lAmpSongsConfigPath << "\n";
    CDtaFile lSongsConfig;
    leError = lSongsConfig.Load(lAmpSongsConfigPath.c_str());
    SHOW_ERROR_AND_RETURN;
    std::cout << "\n";
    int ii = 1;
    std::vector<SSongConfig> lSongs = lAmpConfig.GetSongs();
    lSongsConfig.GetSongData(lSongs);
    for (const SSongConfig& lSong : lSongs) {
        std::cout << "Song " << ii << "\t  " << lSong.mId << " - " << lSong.mName << " - " << lSong.mType << "\n\t  "
                  << lSong.mPath << "\n\t  Unlocked by " << lSong.mUnlockMethod << " " << lSong.miUnlockCount << "\n"
                  << "\t  Arena: " << lSong.mArena << "\n\n";
        ++ii;
    }
    return eError_NoError;
}

// -dtaset <file> <key> <value>: load a binary DTA file, replace the value that follows the symbol
// <key> (integer or string, whichever is there) and save it back -- the host-side patch step of a
// repack.  -dtacopy <in> <out>: load + save (codec round trip).
eError DtaSet(std::deque<std::string>& laParams)
{
    if (laParams.size() < 3)
        return eError_InvalidParameter;
    const std::string lFile = laParams[0], lKey = laParams[1], lValue = laParams[2];
    laParams.erase(laParams.begin(), laParams.begin() + 3);
    CDtaFile lDta;
    eError leError = lDta.Load(lFile.c_str());
    ERROR_RETURN;
    char* lpEnd = nullptr;
    const long liValue = std::strtol(lValue.c_str(), &lpEnd, 0);
    const bool lbNumeric = lpEnd && *lpEnd == 0 && !lValue.empty();
    if (!((lbNumeric && lDta.SetIntAfter(lKey, (int32_t)liValue)) || lDta.SetStringAfter(lKey, lValue))) {
        std::cout << "No value of a matching type follows " << lKey << "\n";
        return eError_InvalidParameter;
    }
    return lDta.Save(lFile.c_str());
}

eError DtaCopy(std::deque<std::string>& laParams)
{
    if (laParams.size() < 2)
        return eError_InvalidParameter;
    const std::string lIn = laParams[0], lOut = laParams[1];
    laParams.erase(laParams.begin(), laParams.begin() + 2);
    CDtaFile lDta;
    eError leError = lDta.Load(lIn.c_str());
    ERROR_RETURN;
    return lDta.Save(lOut.c_str());
}

void PrintUsage()
{
    std::cout << "Usage: modulate <options> <command>\n\n"
              << "Options:\n"
              << "  -verbose        Show additional information during operation\n"
              << "  -ps3            Switch to PS3 mode (default PS4)\n"
              << "  -force          Overwrite existing output files\n"
              << "  -packall        Pack every file found, bypassing the /songs/ filter\n"
              << "  -bodykey <k>    Cipher entry bodies with key k (extension; default 0 = plain, like the game)\n"
              << "  -device <n>     Run on GPU n\n\n"
              << "Commands:\n"
              << "  -unpack <out_dir>           Unpack main_<platform>.hdr (+ .ark parts) from the current directory\n"
              << "  -pack <in_dir> <out_dir>    Repack the files the reference header knows\n"
              << "  -pack_add <in_dir> <out_dir> Repack, also adding new files\n"
              << "  -decode                     Write the deciphered header to main_<platform>.hdr.dec\n"
              << "  -listsongs <dir>            List the songs the DTA configs under <dir> define\n"
              << "  -dtaset <file> <key> <val>  Patch the value following symbol <key> in a binary DTA file\n"
              << "  -dtacopy <in> <out>         Load and re-save a binary DTA file\n";
}

}  // namespace

// MOD_TRACE=1: wall-clock stamps of the process phases (a cold start is dominated by CUDA context creation and
// page-locking, not by the archive work; tools/cli_timing.sh reads these).
static void TraceStamp(const char* lpWhat)
{
    const char* lpTrace = std::getenv("MOD_TRACE");
    if (!lpTrace || !*lpTrace || *lpTrace == '0')
        return;
    const double ldNow = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
    std::fprintf(stderr, "[mod] cli %s at %.3f\n", lpWhat, ldNow);
}

int main(int argc, char* argv[])
{
    TraceStamp("main entered");
    struct sCommandPair {
        const char* mpCommandName;
        std::function<eError(std::deque<std::string>&)> mFunction;
    };
    const sCommandPair kaCommands[] = {
        {"-ps3", PS3},       {"-verbose", EnableVerbose}, {"-force", EnableForceWrite}, {"-packall", EnablePackAll},
        {"-bodykey", BodyKey}, {"-device", Device},       {"-unpack", Unpack},          {"-pack", Pack},
        {"-pack_add", AddPack}, {"-decode", Decode},   {"-dtaset", DtaSet},          {"-dtacopy", DtaCopy},
        {"-listsongs", ListSongs},
    };

    std::deque<std::string> laParams;
    for (int ii = 1; ii < argc; ++ii)
        laParams.push_back(argv[ii]);
    if (laParams.empty()) {
        PrintUsage();
        return 0;
    }
    while (!laParams.empty()) {
        bool lbMatched = false;
        for (const sCommandPair& lCommand : kaCommands) {
            if (strcasecmp(laParams.front().c_str(), lCommand.mpCommandName) != 0)
                continue;
            lbMatched = true;
            laParams.pop_front();
            TraceStamp(lCommand.mpCommandName);
            const eError leError = lCommand.mFunction(laParams);
            if (leError != eError_NoError) {
                ShowError(leError);
                return -1;
            }
            std::cout << "\n";
            break;
        }
        if (!lbMatched) {
            std::cout << "Unkown parameter: " << laParams.front() << "\nAborting\n\n";
            PrintUsage();
            return -1;
        }
    }
    std::cout << "Complete!\n";
    TraceStamp("main returns");
    return 0;
}
