#include "Settings.h"

// Defaults as in the reference (Settings.cpp:4-9).
bool CSettings::mbPS4 = true;
const char* CSettings::msPlatform = "ps4";
bool CSettings::mbVerbose = false;
bool CSettings::mbOverwriteOutputFiles = true;
bool CSettings::mbIgnoreNewFiles = true;
bool CSettings::mbPackAllFiles = false;
