// cycle_kernels.cuh -- launch interface of the sm_100a keystream kernels (see cycle_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace modk {

// Geometry of the work decomposition.  A "chunk" is one 16-byte, 16-byte-aligned piece of the
// DESTINATION address space; a "tile" is what one warp processes at a time: kIters rounds of 32
// chunks (8 KiB of destination with the default kIters = 16).
#ifndef MODK_ITERS
#define MODK_ITERS 16
#endif
constexpr int kIters = MODK_ITERS;
constexpr int kChunksPerTile = 32 * kIters;           // 512 chunks
constexpr uint32_t kTileBytes = 16u * kChunksPerTile;  // 8 KiB of destination per tile
constexpr int kWarpsPerCta = 8;
constexpr int kThreadsPerCta = 32 * kWarpsPerCta;
constexpr int kMaxInlineDescs = 64;  // descriptors that travel in the kernel parameter block

// Device-side descriptor: mod_desc plus the index of the entry's first tile (32 bytes).
struct __align__(16) DevDesc {
    uint64_t src_off;
    uint64_t dst_off;
    uint32_t len;
    int32_t key;
    uint32_t first_tile;
    uint32_t pad;
};

// One record per tile, written once by the plan kernel and read (one 32-byte load, prefetched a
// tile ahead) by the batched kernel: everything a warp needs to start streaming, so the hot kernel
// does no search, no division and no table walk.
struct __align__(16) TileRec {
    uint64_t src_off;  // of the ENTRY this tile belongs to
    uint64_t dst_off;
    uint32_t len;      // entry length
    uint32_t state;    // negated LCG state just before byte (16 * c_begin - h0) of the entry
    uint32_t tin;      // tile index inside the entry (c_begin = tin * kChunksPerTile)
    uint32_t pad;
};

struct InlineDescs {
    DevDesc d[kMaxInlineDescs];
};

struct BatchArgs {
    const uint8_t* src;
    uint8_t* dst;
    const TileRec* tiles;  // HBM tile records (nullptr in inline mode)
    uint32_t n_tiles;
    uint32_t tiles_per_entry;  // inline mode: entry = tile / tiles_per_entry
    uint32_t rounds_per_tile = kIters;  // inline mode: tile length in rounds of 32 chunks (a divisor of kIters)
    // 16-byte granules of the source may be loaded whole only inside [src_lo16, src_hi16).
    uint64_t src_lo16;
    uint64_t src_hi16;
    uint32_t two = 2;  // the literal 2, kept opaque to ptxas (see low8_canonical_fma)
};

// Number of tiles an entry of `len` bytes occupies when its first destination byte sits at
// (address & 15) == h0.
__host__ __device__ inline uint32_t tiles_for_entry(uint32_t h0, uint32_t len,
                                                    uint32_t chunks_per_tile = (uint32_t)kChunksPerTile)
{
    if (len == 0)
        return 0;
    const uint64_t chunks = ((uint64_t)h0 + len + 15u) >> 4;
    return (uint32_t)((chunks + chunks_per_tile - 1) / chunks_per_tile);
}

cudaError_t upload_tables();  // jump tables -> __constant__ / global memory of the current device
// Persistent grid size for the current device (SM count x resident CTAs per SM), cached per device.
cudaError_t persistent_grid(int* grid_out, bool inline_kernel);
cudaError_t launch_batch(const BatchArgs& args, cudaStream_t stream);
cudaError_t launch_batch_inline(const BatchArgs& args, const InlineDescs& descs, cudaStream_t stream);
// Plan kernel: expands descriptors into per-tile records (binary search + jump-ahead per tile).
cudaError_t launch_build_tiles(const DevDesc* descs, uint32_t n_descs, uint32_t dst_align, TileRec* tiles,
                               uint32_t n_tiles, cudaStream_t stream);

}  // namespace modk
