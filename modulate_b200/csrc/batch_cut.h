// batch_cut.h -- how the host-pointer batch path cuts a descriptor list into pipeline groups (pure host logic).
//
// Pinned <-> HBM copies only run at full speed between 128-byte aligned addresses (tools/copy_align_probe.py),
// and the entries of an archive are byte-packed: a group that simply ends with whichever entry fills it starts
// and ends anywhere.  So the group boundaries are put INSIDE entries instead: the entry that crosses the
// group size is split at the nearest destination address with the wanted alignment, and the second piece
// continues the entry's keystream from the jumped key (k0 a^pos) -- exactly what already happens to entries
// larger than a group.  The uploads and downloads of consecutive groups then meet at aligned addresses.
#pragma once

#include <stdint.h>

#include <vector>

#include "../../include/modulate_b200.h"
#include "lcg.h"

namespace modcut {

struct Cut {
    std::vector<mod_desc> pieces;  // what the plan is built from (descriptor order preserved)
    std::vector<uint8_t> closes;   // closes[i] != 0: a group ends after pieces[i]
};

// `dst_phase` + dst_off is the address whose low bits decide the alignment of a cut (`modulus` a power of two,
// a multiple of 16: the kernels need pieces to start on 16-byte boundaries of the destination).
inline void cut_into_groups(const mod_desc* descs, uint64_t n, uint64_t group_bytes, uint64_t dst_phase,
                            uint64_t modulus, Cut& out)
{
    out.pieces.clear();
    out.closes.clear();
    out.pieces.reserve(n + n / 64 + 16);
    out.closes.reserve(n + n / 64 + 16);
    uint64_t acc = 0;  // payload bytes of the group being filled
    for (uint64_t i = 0; i < n; ++i) {
        const mod_desc& d = descs[i];
        uint64_t pos = 0;
        do {
            const uint64_t rem = (uint64_t)d.len - pos;
            const uint64_t room = group_bytes - acc;  // acc < group_bytes always holds here
            uint64_t take = rem;
            bool close = false;
            if (rem >= room) {
                // the group fills up inside (or exactly at the end of) this entry: cut at the aligned
                // destination address just below the fill point, or just above it if that would be empty
                const uint64_t phase = (dst_phase + d.dst_off + pos + room) & (modulus - 1);
                take = room > phase ? room - phase : room + (modulus - phase);
                if (take >= rem)
                    take = rem;
                close = true;
            }
            const int32_t key = pos == 0 ? d.key
                                         : (int32_t)modlcg::mulmod(modlcg::key_residue(d.key), modlcg::pow_a(pos));
            out.pieces.push_back(mod_desc{d.src_off + pos, d.dst_off + pos, (uint32_t)take, key});
            out.closes.push_back(close ? 1 : 0);
            acc = close ? 0 : acc + take;
            pos += take;
        } while (pos < d.len);
    }
    if (!out.closes.empty())
        out.closes.back() = 1;
}

}  // namespace modcut
