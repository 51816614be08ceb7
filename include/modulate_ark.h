/*
 * modulate_ark.h -- C binding of the archive-level facade (CArk / CDtaFile), for callers that cannot
 * link C++ classes (scripting languages, bench.py's end-to-end repack workload).
 *
 * Each function is one of the reference's command bodies run on the facade classes of
 * modulate_b200/csrc/ (which forward the data-parallel half to the kernels behind
 * modulate_b200.h):
 *   mod_ark_unpack   Unpack   (reference Modulate.cpp:291-317: CArk::Load + CArk::ExtractFiles)
 *   mod_ark_pack     Pack     (reference Modulate.cpp:380-450: song list from the DTA configs unless
 *                              pack_all, CArk::Load of the reference header, ConstructFromDirectory,
 *                              BuildArk, SaveArk)
 *   mod_dta_set_int  the host-side DTA patch of a repack: CDtaFile::Load, replace the integer that
 *                    follows the symbol `key`, CDtaFile::Save
 * They return the reference's eError codes (0 = eError_NoError, reference Error.h:5-20) and print
 * the reference's messages to stdout.  Like the reference they go through the process-wide CSettings
 * switches, so they are not re-entrant.  `body_key` != 0 ciphers every entry body with that key,
 * stream restarting per entry (extension: the reference never ciphers ARK bodies).
 */
#ifndef MODULATE_ARK_H
#define MODULATE_ARK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* `header_path`: the .hdr file; `part_dir`: directory of the .ark parts ("" = current directory);
 * `target_dir`: where the files go (created as needed; must end in '/'). */
int mod_ark_unpack(const char* header_path, const char* part_dir, const char* target_dir, int32_t body_key);

/* `input_dir` and `output_dir` must end in '/'; the header is written to output_dir + header_name. */
int mod_ark_pack(const char* reference_header_path, const char* input_dir, const char* output_dir,
                 const char* header_name, int ps4, int pack_all, int ignore_new_files, int32_t body_key);

int mod_dta_set_int(const char* dta_path, const char* key, int32_t value);

#ifdef __cplusplus
}
#endif

#endif /* MODULATE_ARK_H */
