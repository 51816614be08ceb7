// oracle/ref_ark/direct.h -- TEST INFRASTRUCTURE.  Stand-in for MSVC's <direct.h> (_mkdir): see windows.h.
#pragma once
#include "windows.h"
