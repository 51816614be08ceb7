// ark_abi.cpp -- the C binding declared in include/modulate_ark.h: the reference's Unpack / Pack
// command bodies (Modulate.cpp:291-317, :380-450) on the facade classes.
#include "../../include/modulate_ark.h"

#include <exception>
#include <iostream>
#include <string>
#include <vector>

#include "CArk.h"
#include "CDtaFile.h"
#include "Error.h"
#include "Settings.h"

namespace {

// No C++ exception may cross the extern "C" boundary.
template <class F>
int Guarded(F lBody)
{
    try {
        return (int)lBody();
    } catch (const std::exception& lError) {
        std::cout << "ERROR: " << lError.what() << "\n";
        return (int)eError_InvalidData;
    } catch (...) {
        return (int)eError_InvalidData;
    }
}

}  // namespace

extern "C" {

int mod_ark_unpack(const char* header_path, const char* part_dir, const char* target_dir, int32_t body_key)
{
    return Guarded([&]() -> eError {
        if (!header_path || !target_dir)
            return eError_InvalidParameter;
        CArk lArkHeader;
        lArkHeader.SetUniformEntryKey(body_key);
        lArkHeader.SetPartDirectory(part_dir ? part_dir : "");
        eError leError = lArkHeader.Load(header_path);
        SHOW_ERROR_AND_RETURN;
        leError = lArkHeader.ExtractFiles(0, lArkHeader.GetNumFiles(), target_dir);
        SHOW_ERROR_AND_RETURN;
        return eError_NoError;
    });
}

int mod_ark_pack(const char* reference_header_path, const char* input_dir, const char* output_dir, const char* header_name,
                 int ps4, int pack_all, int ignore_new_files, int32_t body_key)
{
    return Guarded([&]() -> eError {
        if (!reference_header_path || !input_dir || !output_dir || !header_name)
            return eError_InvalidParameter;
        CSettings::mbPS4 = ps4 != 0;
        CSettings::msPlatform = ps4 ? "ps4" : "ps3";
        CSettings::mbPackAllFiles = pack_all != 0;
        CSettings::mbIgnoreNewFiles = ignore_new_files != 0;
        const std::string lInputPath = input_dir, lPlatform = CSettings::msPlatform;

        eError leError = eError_NoError;
        std::vector<SSongConfig> lSongs;
        if (!CSettings::mbPackAllFiles) {
            CDtaFile lAmpConfig;
            leError = lAmpConfig.Load((lInputPath + lPlatform + "/config/amp_config.dta_dta_" + lPlatform).c_str());
            SHOW_ERROR_AND_RETURN;
            lSongs = lAmpConfig.GetSongs();
            CDtaFile lSongsConfig;
            leError = lSongsConfig.Load((lInputPath + lPlatform + "/config/amp_songs_config.dta_dta_" + lPlatform).c_str());
            SHOW_ERROR_AND_RETURN;
            lSongsConfig.GetSongData(lSongs);
        }
        CArk lReferenceArkHeader;
        leError = lReferenceArkHeader.Load(reference_header_path);
        SHOW_ERROR_AND_RETURN;
        CArk lArkHeader;
        lArkHeader.SetUniformEntryKey(body_key);
        leError = lArkHeader.ConstructFromDirectory(input_dir, lReferenceArkHeader, lSongs);
        SHOW_ERROR_AND_RETURN;
        leError = lArkHeader.BuildArk(input_dir, lSongs);
        SHOW_ERROR_AND_RETURN;
        leError = lArkHeader.SaveArk(output_dir, header_name);
        SHOW_ERROR_AND_RETURN;
        return eError_NoError;
    });
}

int mod_dta_set_int(const char* dta_path, const char* key, int32_t value)
{
    return Guarded([&]() -> eError {
        if (!dta_path || !key)
            return eError_InvalidParameter;
        CDtaFile lDta;
        eError leError = lDta.Load(dta_path);
        ERROR_RETURN;
        if (!lDta.SetIntAfter(key, value))
            return eError_InvalidParameter;
        return lDta.Save(dta_path);
    });
}

}  // extern "C"
