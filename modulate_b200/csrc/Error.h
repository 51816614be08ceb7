// Error.h -- error convention of the drop-in boundary.
//
// The enumerator names, their order (hence their integer values) and the "print one line to
// stdout, then return the code" convention mirror the reference's Error.h:5-59, because CArk's
// public methods return eError and callers switch on it (Modulate.cpp:948-956).  The code below
// is this repo's own.
#pragma once

#include <cstdio>

enum eError {
    eError_NoError,
    eError_FailedToOpenFile,
    eError_FailedToCreateDirectory,
    eError_UnknownVersionNumber,
    eError_ValueOutOfBounds,
    eError_AlreadyLoaded,
    eError_InvalidData,
    eError_NoData,
    eError_FailedToCreateFile,
    eError_FailedToDeleteFile,
    eError_FailedToCopyFile,
    eError_InvalidParameter,
    eError_FailedToWriteData,
    eError_NumTypes
};

inline const char* ErrorName(eError leError)
{
    switch (leError) {
    case eError_NoError: return "No Error";
    case eError_FailedToOpenFile: return "Failed to open file";
    case eError_FailedToCreateDirectory: return "Failed to create directory";
    case eError_UnknownVersionNumber: return "Unknown version number";
    case eError_ValueOutOfBounds: return "Value of out bounds";  // sic: the reference's message text
    case eError_AlreadyLoaded: return "Already loaded";
    case eError_InvalidData: return "Bad data";
    case eError_NoData: return "Missing data";
    case eError_FailedToCreateFile: return "Failed to create file";
    case eError_FailedToDeleteFile: return "Failed to delete file";
    case eError_FailedToCopyFile: return "Failed to copy file";
    case eError_InvalidParameter: return "Invalid parameter";
    case eError_FailedToWriteData: return "Failed to write data";
    default: return "Unknown error";
    }
}

inline void ShowError(eError leError)
{
    std::printf("ERROR: %s\n", ErrorName(leError));
    std::fflush(stdout);
}

// Same spelling and behaviour as the reference's early-return macros (Error.h:43-59): they expect a
// local `eError leError` in scope.
#define ERROR_RETURN                       \
    if (leError != eError_NoError) {       \
        return leError;                    \
    }

#define SHOW_ERROR_AND_RETURN              \
    if (leError != eError_NoError) {       \
        ShowError(leError);                \
        return leError;                    \
    }

#define SHOW_ERROR_AND_RETURN_W(lTodo)     \
    if (leError != eError_NoError) {       \
        ShowError(leError);                \
        lTodo;                             \
        return leError;                    \
    }
