// cycle_kernels.cuh -- launch interface of the sm_100a keystream kernels (see cycle_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace modk {

// Geometry of the work decomposition.  A "chunk" is one 16-byte, 16-byte-aligned piece of the
// DESTINATION address space; a "tile" is what one CTA processes before it retires: kUnroll rounds
// of kThreadsPerCta consecutive chunks, i.e. one contiguous 8 KiB span of one entry with the defaults.
#ifndef MODK_THREADS
#define MODK_THREADS 128
#endif
#ifndef MODK_UNROLL
#define MODK_UNROLL 4
#endif
constexpr int kThreadsPerCta = MODK_THREADS;
constexpr int kUnroll = MODK_UNROLL;                       // chunks in flight per thread
constexpr int kChunksPerTile = kThreadsPerCta * kUnroll;   // 512 chunks
constexpr uint32_t kTileBytes = 16u * kChunksPerTile;      // 8 KiB of destination per tile
constexpr int kMaxInlineDescs = 64;  // descriptors that travel in the kernel parameter block
static_assert(kChunksPerTile <= 2048, "TileRec::geom holds the chunk count in 12 bits");

// Device-side descriptor: mod_desc plus the index of the entry's first tile (32 bytes).
struct __align__(16) DevDesc {
    uint64_t src_off;
    uint64_t dst_off;
    uint32_t len;
    int32_t key;
    uint32_t first_tile;
    uint32_t neg_state;  // modlcg::key_to_neg_state(key), filled on the host (no division on the device)
};

// One record per tile, written once by the plan kernel and read (one 32-byte broadcast load) by
// the CTA that owns the tile: everything it needs to start streaming, so the hot kernel does no
// search, no division and no table walk.
struct __align__(16) TileRec {
    int64_t src_rel;   // source byte (relative to the src base) that pairs with byte 0 of the tile's chunk 0
    int64_t dst_rel;   // destination byte (relative to the dst base) of chunk 0; base + dst_rel is 16-byte aligned
    uint32_t state;    // negated LCG state just before byte 0 of chunk 0
    uint32_t geom;     // bits 0-11: chunks in the tile; 12-15: first valid byte of chunk 0; 16-20: valid bytes of the last chunk
    uint32_t entry;    // descriptor index (diagnostics)
    uint32_t pad;
};

__host__ __device__ inline uint32_t pack_geom(uint32_t n_valid, uint32_t head, uint32_t tail)
{
    return n_valid | (head << 12) | (tail << 16);
}

struct InlineDescs {
    DevDesc d[kMaxInlineDescs];
};

struct BatchArgs {
    const uint8_t* src;
    uint8_t* dst;
    const TileRec* tiles;  // HBM tile records (nullptr in inline mode)
    uint32_t n_tiles;
    uint32_t tiles_per_entry;  // inline mode: entry = tile / tiles_per_entry
    // 16-byte granules of the source may be loaded whole only inside [src_lo16, src_hi16).
    uint64_t src_lo16;
    uint64_t src_hi16;
    uint32_t two = 2;  // the literal 2, kept opaque to ptxas (see low8_canonical_fma)
    // (src_off - dst_off) & 15 if it is the same for every entry of the launch, else -1: together with
    // the base pointers it tells the host whether the co-aligned kernel may be used
    int32_t uniform_delta = -1;
};

// Number of tiles an entry of `len` bytes occupies when its first destination byte sits at
// (address & 15) == h0.
__host__ __device__ inline uint32_t tiles_for_entry(uint32_t h0, uint32_t len)
{
    if (len == 0)
        return 0;
    const uint64_t chunks = ((uint64_t)h0 + len + 15u) >> 4;
    return (uint32_t)((chunks + (uint32_t)kChunksPerTile - 1) / (uint32_t)kChunksPerTile);
}

cudaError_t upload_tables();  // jump tables -> __constant__ / global memory of the current device
cudaError_t launch_batch(const BatchArgs& args, cudaStream_t stream);
cudaError_t launch_batch_inline(const BatchArgs& args, const InlineDescs& descs, cudaStream_t stream);
// Plan kernel: expands descriptors into per-tile records (binary search + jump-ahead per tile).
cudaError_t launch_build_tiles(const DevDesc* descs, uint32_t n_descs, uint32_t dst_align, TileRec* tiles,
                               uint32_t n_tiles, cudaStream_t stream);

}  // namespace modk
