# oracle/ref_ark/patch.sed -- the only edits applied to the STAGED copies of the reference sources
# (oracle/_ref/ark_src/, git-ignored, deleted after the build) so that g++ accepts them.  None
# changes behaviour:
#   CDtaFile.h:262-284  four explicit member specialisations written without `template<>` (an MSVC
#                       extension); also `inline`, because the header is included by several TUs
#   CArk.h:8            `enum eError;` -- forward declaration of an unscoped enum without an underlying
#                       type (MSVC extension): replaced by the header that defines it
s/^void CDtaNode< \(int\|unsigned int\|float\|std::string\) >::SaveToStream/template<> inline void CDtaNode< \1 >::SaveToStream/
s/^enum eError;$/#include "Error.h"/
