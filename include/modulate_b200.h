/*
 * modulate_b200.h -- C ABI of the B200-native keystream / archive data-movement path.
 *
 * This is the drop-in boundary for the one data-parallel hot path of AdamClixby/Modulate:
 *   - CEncryptionCycler::Cycle            (reference CEncryptionCycler.h:6, CEncryptionCycler.cpp:4-25)
 *   - CArk::ExtractFiles' gather          (reference CArk.cpp:494, by mi64Offset / miSize)
 *   - CArk::BuildArk's scatter            (reference CArk.cpp:807-811)
 *   - the three Cycle call sites          (CArk.cpp:338-339, CArk.cpp:1135-1136, Modulate.cpp:485-486)
 *
 * The reference has no FFI layer (it is one statically linked C++ program), so the entry
 * points below are what a binding for this path would bind: plain pointers and sizes, no
 * C++ or torch types.  The C++ facade classes with the reference's own signatures
 * (modulate_b200/csrc/CEncryptionCycler.h, CArk.h) forward to these functions; see
 * INTEGRATION.md for the two-line change a maintainer of the reference would make.
 *
 * Conventions
 *   - Every function that can fail returns int: 0 (MOD_OK) on success, a negative MOD_ERR_*
 *     otherwise; mod_last_error() returns a thread-local human-readable message.
 *   - There is NO CPU fallback.  If no CUDA device / driver is usable, compute entry points
 *     return MOD_ERR_CUDA and the C++ facade aborts loudly (the reference's Cycle is void).
 *   - One process drives one GPU (mod_init binds it); multi-GPU runs are one process per GPU
 *     with the work split by mod_shard_descs / mod_shard_range -- shards are independent, so
 *     no collective is needed.
 *   - "stream" parameters are a cudaStream_t passed as void* (NULL = the legacy default stream).
 *     Functions taking a stream are asynchronous; the others return after the result is visible
 *     to the caller, which preserves Cycle's in-place contract.
 */
#ifndef MODULATE_B200_H
#define MODULATE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOD_OK 0
#define MOD_ERR_CUDA (-1)      /* CUDA runtime / driver error (message holds cudaGetErrorString) */
#define MOD_ERR_ARG (-2)       /* invalid argument (null pointer, descriptor outside its buffer, ...) */
#define MOD_ERR_ALIGN (-3)     /* plan was built for a different dst alignment (dst & 15) */
#define MOD_ERR_NOMEM (-4)     /* host or device allocation failed */

#define MOD_ABI_VERSION 1

/* One archive entry / stream piece: copy `len` bytes from src+src_off to dst+dst_off while XORing
 * them with the keystream of CEncryptionCycler::Cycle(., len, key), the stream restarting at the
 * entry's first byte.  Device form of CArk::sFileDefinition {mi64Offset, miSize} (CArk.h:55-69)
 * plus a per-entry key.  key == 0 (mod 2^31-1) is the identity keystream, i.e. the reference's
 * plain fwrite/fread copy.  `len` is 32-bit like the reference's `unsigned int liDataSize`. */
typedef struct mod_desc {
    uint64_t src_off;
    uint64_t dst_off;
    uint32_t len;
    int32_t key;
} mod_desc;

typedef struct mod_plan mod_plan; /* opaque: descriptors + tile map resident in HBM */

/* ---- library / device management --------------------------------------------------------- */

int mod_abi_version(void);
/* Number of CUDA devices visible, or a negative MOD_ERR_* . */
int mod_device_count(void);
/* Bind the calling process to `device` (-1 = keep the current device), create the internal
 * streams and upload the jump tables.  Idempotent.  Called implicitly by the compute entry points. */
int mod_init(int device);
void mod_shutdown(void);
const char* mod_last_error(void);
/* Kernels this library has launched in this process (all streams): bench.py's gpu_launches. */
uint64_t mod_launch_count(void);

/* ---- memory helpers ---------------------------------------------------------------------- */

void* mod_host_alloc(uint64_t bytes); /* pinned (page-locked) host memory; NULL on failure */
int mod_host_free(void* p);
void* mod_device_alloc(uint64_t bytes); /* HBM; NULL on failure */
int mod_device_free(void* p);
int mod_memcpy_h2d(void* d_dst, const void* h_src, uint64_t bytes, void* stream);
int mod_memcpy_d2h(void* h_dst, const void* d_src, uint64_t bytes, void* stream);
int mod_stream_sync(void* stream);

/* ---- CEncryptionCycler::Cycle ------------------------------------------------------------ */

/* Exact semantics of CEncryptionCycler::Cycle (CEncryptionCycler.cpp:4-14) on `len` bytes at `data`,
 * in place, synchronous.  `data` may be a host pointer (pageable or pinned: staged through HBM in
 * pipelined slices, H2D / kernel / D2H overlapped) or a device pointer (cycled in HBM).  `len` is
 * 64-bit: streams longer than the reference's 32-bit length continue the same keystream
 * (period 2^31-2). */
int mod_cycle(void* data, uint64_t len, int32_t key);

/* Device-resident, asynchronous on `stream`.  d_src == d_dst is the in-place form; otherwise the
 * ranges must not overlap.  No alignment requirement on either pointer. */
int mod_cycle_device(const void* d_src, void* d_dst, uint64_t len, int32_t key, void* stream);

/* Key k' such that Cycle(., ., k') emits the keystream of Cycle(., ., key) from byte `pos` on:
 * the O(log pos) modular jump-ahead used for offset-range sharding.  Pure host arithmetic. */
int32_t mod_key_jump(int32_t key, uint64_t pos);

/* ---- descriptor batches: CArk extract (gather) / build (scatter) / per-entry keys ----------- */

/* Validate `n` host descriptors against the two buffer sizes, upload them, and build the per-tile
 * records in HBM (the device form of the file table CArk::Load builds, CArk.cpp:401-408).  `dst_align` is (address of dst) & 15 for the buffer the plan will be
 * run on (0 for anything from mod_device_alloc / cudaMalloc).  Entries whose dst ranges overlap
 * give unspecified results (the reference never produces them). */
int mod_plan_create(const mod_desc* descs, uint64_t n, uint64_t src_bytes, uint64_t dst_bytes,
                    uint32_t dst_align, mod_plan** out);
int mod_plan_destroy(mod_plan* plan);
uint64_t mod_plan_payload_bytes(const mod_plan* plan); /* sum of len */
uint64_t mod_plan_num_tiles(const mod_plan* plan);

/* One launch of the variable-length batched kernel over every descriptor of the plan: replaces the
 * per-entry loops of CArk::ExtractFiles (CArk.cpp:435-501, payload move :494) and CArk::BuildArk
 * (CArk.cpp:784-822, payload move :807-811), with one CEncryptionCycler::Cycle per entry folded in.
 * Asynchronous on `stream`.  d_src == d_dst with src_off == dst_off is the in-place form. */
int mod_plan_run(const mod_plan* plan, const void* d_src, void* d_dst, void* stream);

/* Convenience: plan + run + destroy, synchronous.  `src` / `dst` are both host pointers (staged
 * through HBM in groups of entries, upload / kernel / download overlapped) or both device pointers.
 * In both forms bytes of dst that no descriptor covers keep their value. */
int mod_cycle_batch(const mod_desc* descs, uint64_t n, const void* src, uint64_t src_bytes,
                    void* dst, uint64_t dst_bytes);

/* ---- offset-range sharding (host logic, no GPU needed) -------------------------------------- */

/* (No reference counterpart: the reference is single-threaded)
 * Byte range [*begin, *end) of a `total`-byte stream owned by `rank` of `world`: equal shares
 * rounded to 16-byte boundaries (the last rank takes the remainder). */
int mod_shard_range(uint64_t total, int rank, int world, uint64_t* begin, uint64_t* end);

/* Split a descriptor list into `world` shards of (nearly) equal payload bytes and write rank's
 * shard to `out` (capacity out_cap; pass out == NULL to query the count).  An entry that straddles
 * a shard boundary is cut at a 16-byte-aligned position inside the entry and the second piece gets
 * the jumped key (mod_key_jump), so the union of all shards is byte-identical to the unsharded
 * batch.  Returns the number of descriptors in rank's shard, or a negative MOD_ERR_*. */
int64_t mod_shard_descs(const mod_desc* descs, uint64_t n, int rank, int world, mod_desc* out,
                        uint64_t out_cap);

#ifdef __cplusplus
}
#endif

#endif /* MODULATE_B200_H */
