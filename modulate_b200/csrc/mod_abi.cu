// mod_abi.cu -- the C ABI declared in include/modulate_b200.h: device binding, pinned / HBM
// memory helpers, CEncryptionCycler::Cycle on host or device buffers, descriptor plans for CArk's
// extract / build data movement, and the host-side shard planner.  No CPU compute path exists
// here: every byte of keystream is produced by the kernels in cycle_kernels.cu.
#include "../../include/modulate_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "cycle_kernels.cuh"
#include "lcg.h"

namespace {

#ifndef MOD_PIPE_SLOTS
#define MOD_PIPE_SLOTS 4
#endif
constexpr int kPipeSlots = MOD_PIPE_SLOTS;               // slices in flight on the host-pointer path
constexpr uint64_t kMaxPiece = 1ull << 30;  // a contiguous stream is cut into <= 1 GiB pieces
constexpr int kMaxDevices = 64;

thread_local std::string tl_error = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    tl_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(MOD_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct Context {
    std::mutex mu;
    bool tables_ready[kMaxDevices] = {};
    int device = -1;
    cudaStream_t pipe_stream[kPipeSlots] = {};
    bool streams_ready = false;
    int streams_device = -1;
    // grow-only HBM workspaces of the host-pointer paths (owned by streams_device)
    void* slice_buf[kPipeSlots] = {};
    uint64_t slice_bytes = 0;
    void* ws_src = nullptr;
    uint64_t ws_src_bytes = 0;
    void* ws_dst = nullptr;
    uint64_t ws_dst_bytes = 0;
    // plan scratch of the host-pointer batch path: pinned staging for descriptors, HBM descriptors + tiles
    void* h_descs = nullptr;
    uint64_t h_descs_bytes = 0;
    void* ws_descs = nullptr;
    uint64_t ws_descs_bytes = 0;
    void* ws_tiles = nullptr;
    uint64_t ws_tiles_bytes = 0;
    cudaEvent_t plan_ready = nullptr;
};

Context g_ctx;

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

uint64_t env_u64(const char* name, uint64_t dflt)
{
    const char* v = getenv(name);
    if (!v || !*v)
        return dflt;
    return strtoull(v, nullptr, 10);
}

void release_workspaces_locked()
{
    for (int i = 0; i < kPipeSlots; ++i) {
        if (g_ctx.slice_buf[i])
            cudaFree(g_ctx.slice_buf[i]);
        g_ctx.slice_buf[i] = nullptr;
    }
    g_ctx.slice_bytes = 0;
    if (g_ctx.ws_src)
        cudaFree(g_ctx.ws_src);
    if (g_ctx.ws_dst)
        cudaFree(g_ctx.ws_dst);
    g_ctx.ws_src = g_ctx.ws_dst = nullptr;
    g_ctx.ws_src_bytes = g_ctx.ws_dst_bytes = 0;
    if (g_ctx.ws_descs)
        cudaFree(g_ctx.ws_descs);
    if (g_ctx.ws_tiles)
        cudaFree(g_ctx.ws_tiles);
    if (g_ctx.h_descs)
        cudaFreeHost(g_ctx.h_descs);
    g_ctx.ws_descs = g_ctx.ws_tiles = g_ctx.h_descs = nullptr;
    g_ctx.ws_descs_bytes = g_ctx.ws_tiles_bytes = g_ctx.h_descs_bytes = 0;
    if (g_ctx.plan_ready) {
        cudaEventDestroy(g_ctx.plan_ready);
        g_ctx.plan_ready = nullptr;
    }
    if (g_ctx.streams_ready) {
        for (int i = 0; i < kPipeSlots; ++i)
            cudaStreamDestroy(g_ctx.pipe_stream[i]);
        g_ctx.streams_ready = false;
    }
}

// Bind + lazily prepare the current device.  Caller holds no lock.
int ensure_ready(int device)
{
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (count <= 0)
        return fail(MOD_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
    if (device >= 0) {
        if (device >= count || device >= kMaxDevices)
            return fail(MOD_ERR_ARG, "device %d out of range (%d visible)", device, count);
        CUDA_TRY(cudaSetDevice(device));
    }
    int cur = 0;
    CUDA_TRY(cudaGetDevice(&cur));
    if (cur >= kMaxDevices)
        return fail(MOD_ERR_ARG, "device %d beyond the supported %d", cur, kMaxDevices);
    if (!g_ctx.tables_ready[cur]) {
        CUDA_TRY(modk::upload_tables());
        g_ctx.tables_ready[cur] = true;
    }
    if (g_ctx.streams_ready && g_ctx.streams_device != cur) {
        cudaSetDevice(g_ctx.streams_device);
        release_workspaces_locked();
        cudaSetDevice(cur);
    }
    if (!g_ctx.streams_ready) {
        for (int i = 0; i < kPipeSlots; ++i)
            CUDA_TRY(cudaStreamCreateWithFlags(&g_ctx.pipe_stream[i], cudaStreamNonBlocking));
        g_ctx.streams_ready = true;
        g_ctx.streams_device = cur;
    }
    g_ctx.device = cur;
    return MOD_OK;
}

bool is_device_pointer(const void* p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int grow(void** buf, uint64_t* have, uint64_t need)
{
    if (*have >= need)
        return MOD_OK;
    if (*buf) {
        CUDA_TRY(cudaFree(*buf));
        *buf = nullptr;
        *have = 0;
    }
    need = (need + 255) & ~255ull;
    cudaError_t e = cudaMalloc(buf, need);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MOD_ERR_NOMEM, "cudaMalloc(%llu) failed: %s", (unsigned long long)need, cudaGetErrorString(e));
    }
    *have = need;
    return MOD_OK;
}

// One launch (or a few, beyond 64 pieces) of the batched kernel over a contiguous stream.
int launch_contiguous(const uint8_t* d_src, uint8_t* d_dst, uint64_t len, int32_t key, cudaStream_t stream)
{
    if (len == 0)
        return MOD_OK;
    const uint32_t h0 = (uint32_t)((uintptr_t)d_dst & 15u);
    // piece size: <= 1 GiB, and large enough that one launch covers up to 64 pieces
    uint64_t piece = kMaxPiece;
    const uint64_t src_lo16 = ((uint64_t)(uintptr_t)d_src + 15u) & ~15ull;
    const uint64_t src_hi16 = ((uint64_t)(uintptr_t)d_src + len) & ~15ull;
    // short buffers use shorter tiles so that they still spread over the whole GPU: halve the tile
    // until there is at least one tile per resident warp or a tile is a single round
    uint32_t rounds = (uint32_t)modk::kIters;
    {
        int grid_cap = 0;
        CUDA_TRY(modk::persistent_grid(&grid_cap, true));
        const uint64_t want_tiles = (uint64_t)grid_cap * modk::kWarpsPerCta;
        while (rounds > 1 && (len + 512ull * rounds - 1) / (512ull * rounds) < want_tiles)
            rounds >>= 1;
    }
    const uint32_t cpt = 32u * rounds;  // chunks per tile
    const uint32_t tpe = modk::tiles_for_entry(h0, (uint32_t)piece, cpt);
    uint64_t done = 0;
    uint32_t k0 = modlcg::key_residue(key);
    while (done < len) {
        modk::InlineDescs in;
        modk::BatchArgs args;
        uint32_t n = 0, tiles = 0;
        const uint64_t group_base = done;
        while (done < len && n < (uint32_t)modk::kMaxInlineDescs) {
            const uint64_t this_len = std::min(piece, len - done);
            modk::DevDesc& d = in.d[n];
            d.src_off = done - group_base;
            d.dst_off = done - group_base;
            d.len = (uint32_t)this_len;
            d.key = (int32_t)modlcg::mulmod(k0, modlcg::pow_a(done));
            d.first_tile = n * tpe;
            d.pad = 0;
            tiles = n * tpe + modk::tiles_for_entry(h0, (uint32_t)this_len, cpt);
            done += this_len;
            ++n;
        }
        args.src = d_src + group_base;
        args.dst = d_dst + group_base;
        args.tiles = nullptr;
        args.n_tiles = tiles;
        args.tiles_per_entry = tpe;
        args.rounds_per_tile = rounds;
        args.src_lo16 = src_lo16;
        args.src_hi16 = src_hi16;
        CUDA_TRY(modk::launch_batch_inline(args, in, stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return MOD_OK;
}

// Validate descriptors against the buffer sizes and expand them into device descriptors with their
// first-tile prefix.  Shared by mod_plan_create and the host-pointer batch path.
int expand_descs(const mod_desc* descs, uint64_t n, uint64_t src_bytes, uint64_t dst_bytes, uint32_t dst_align,
                 modk::DevDesc* out, uint64_t* tiles_out, uint64_t* payload_out)
{
    uint64_t tiles = 0, payload = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const mod_desc& d = descs[i];
        if (d.src_off > src_bytes || (uint64_t)d.len > src_bytes - d.src_off)
            return fail(MOD_ERR_ARG, "descriptor %llu: source range [%llu, +%u) leaves the %llu-byte buffer",
                        (unsigned long long)i, (unsigned long long)d.src_off, d.len, (unsigned long long)src_bytes);
        if (d.dst_off > dst_bytes || (uint64_t)d.len > dst_bytes - d.dst_off)
            return fail(MOD_ERR_ARG, "descriptor %llu: destination range [%llu, +%u) leaves the %llu-byte buffer",
                        (unsigned long long)i, (unsigned long long)d.dst_off, d.len, (unsigned long long)dst_bytes);
        modk::DevDesc& o = out[i];
        o.src_off = d.src_off;
        o.dst_off = d.dst_off;
        o.len = d.len;
        o.key = d.key;
        o.first_tile = (uint32_t)tiles;
        o.pad = 0;
        tiles += modk::tiles_for_entry((uint32_t)((dst_align + d.dst_off) & 15u), d.len);
        payload += d.len;
        if (tiles >= 0xFFFFFFFFull)
            return fail(MOD_ERR_ARG, "batch too large (tile count overflows 32 bits)");
    }
    *tiles_out = tiles;
    *payload_out = payload;
    return MOD_OK;
}

int grow_pinned(void** buf, uint64_t* have, uint64_t need)
{
    if (*have >= need)
        return MOD_OK;
    if (*buf) {
        CUDA_TRY(cudaFreeHost(*buf));
        *buf = nullptr;
        *have = 0;
    }
    need = (need + 4095) & ~4095ull;
    cudaError_t e = cudaHostAlloc(buf, need, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MOD_ERR_NOMEM, "cudaHostAlloc(%llu) failed: %s", (unsigned long long)need, cudaGetErrorString(e));
    }
    *have = need;
    return MOD_OK;
}

}  // namespace

struct mod_plan {
    int device = -1;
    uint64_t n = 0;
    uint64_t src_bytes = 0;
    uint64_t dst_bytes = 0;
    uint32_t dst_align = 0;
    uint64_t payload = 0;
    uint32_t n_tiles = 0;
    modk::TileRec* d_tiles = nullptr;
};

extern "C" {

int mod_abi_version(void) { return MOD_ABI_VERSION; }

const char* mod_last_error(void) { return tl_error.c_str(); }

uint64_t mod_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mod_device_count(void)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MOD_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return count;
}

int mod_init(int device) { return ensure_ready(device); }

void mod_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    if (g_ctx.streams_ready) {
        cudaSetDevice(g_ctx.streams_device);
        cudaDeviceSynchronize();
    }
    release_workspaces_locked();
}

/* ---- memory helpers ------------------------------------------------------------------------ */

void* mod_host_alloc(uint64_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(MOD_ERR_NOMEM, "cudaHostAlloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

int mod_host_free(void* p)
{
    if (!p)
        return MOD_OK;
    CUDA_TRY(cudaFreeHost(p));
    return MOD_OK;
}

void* mod_device_alloc(uint64_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(MOD_ERR_NOMEM, "cudaMalloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

int mod_device_free(void* p)
{
    if (!p)
        return MOD_OK;
    CUDA_TRY(cudaFree(p));
    return MOD_OK;
}

int mod_memcpy_h2d(void* d_dst, const void* h_src, uint64_t bytes, void* stream)
{
    CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return MOD_OK;
}

int mod_memcpy_d2h(void* h_dst, const void* d_src, uint64_t bytes, void* stream)
{
    CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return MOD_OK;
}

int mod_stream_sync(void* stream)
{
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return MOD_OK;
}

/* ---- CEncryptionCycler::Cycle ---------------------------------------------------------------- */

int32_t mod_key_jump(int32_t key, uint64_t pos)
{
    return (int32_t)modlcg::mulmod(modlcg::key_residue(key), modlcg::pow_a(pos));
}

int mod_cycle_device(const void* d_src, void* d_dst, uint64_t len, int32_t key, void* stream)
{
    if (len == 0)
        return MOD_OK;
    if (!d_src || !d_dst)
        return fail(MOD_ERR_ARG, "mod_cycle_device: null pointer");
    int rc = ensure_ready(-1);
    if (rc != MOD_OK)
        return rc;
    return launch_contiguous((const uint8_t*)d_src, (uint8_t*)d_dst, len, key, (cudaStream_t)stream);
}

int mod_cycle(void* data, uint64_t len, int32_t key)
{
    if (len == 0)
        return MOD_OK;
    if (!data)
        return fail(MOD_ERR_ARG, "mod_cycle: null pointer");
    int rc = ensure_ready(-1);
    if (rc != MOD_OK)
        return rc;

    if (is_device_pointer(data)) {
        rc = launch_contiguous((const uint8_t*)data, (uint8_t*)data, len, key, nullptr);
        if (rc != MOD_OK)
            return rc;
        CUDA_TRY(cudaStreamSynchronize(nullptr));
        return MOD_OK;
    }

    // Host buffer: slices travel H2D -> kernel -> D2H on kPipeSlots streams so that the upload of
    // slice i+1, the kernel of slice i and the download of slice i-1 overlap (two copy engines).
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    uint64_t slice = env_u64("MOD_SLICE_BYTES", 16ull << 20);
    slice = std::max<uint64_t>(4096, slice & ~4095ull);
    if (len < slice * 2)  // small buffers: still use several slots so both directions overlap
        slice = std::max<uint64_t>(4096, ((len / kPipeSlots) + 4095) & ~4095ull);
    if (g_ctx.slice_bytes < slice) {
        for (int i = 0; i < kPipeSlots; ++i) {
            uint64_t have = g_ctx.slice_bytes;
            rc = grow(&g_ctx.slice_buf[i], &have, slice);
            if (rc != MOD_OK) {
                g_ctx.slice_bytes = 0;
                return rc;
            }
        }
        g_ctx.slice_bytes = slice;
    }
    uint8_t* host = (uint8_t*)data;
    const uint32_t k0 = modlcg::key_residue(key);
    uint64_t pos = 0;
    for (uint64_t i = 0; pos < len; ++i) {
        const int slot = (int)(i % kPipeSlots);
        const uint64_t n = std::min(slice, len - pos);
        cudaStream_t s = g_ctx.pipe_stream[slot];
        uint8_t* d = (uint8_t*)g_ctx.slice_buf[slot];
        CUDA_TRY(cudaMemcpyAsync(d, host + pos, n, cudaMemcpyHostToDevice, s));
        rc = launch_contiguous(d, d, n, (int32_t)modlcg::mulmod(k0, modlcg::pow_a(pos)), s);
        if (rc != MOD_OK)
            return rc;
        CUDA_TRY(cudaMemcpyAsync(host + pos, d, n, cudaMemcpyDeviceToHost, s));
        pos += n;
    }
    for (int i = 0; i < kPipeSlots; ++i)
        CUDA_TRY(cudaStreamSynchronize(g_ctx.pipe_stream[i]));
    return MOD_OK;
}

/* ---- descriptor plans -------------------------------------------------------------------------- */

int mod_plan_create(const mod_desc* descs, uint64_t n, uint64_t src_bytes, uint64_t dst_bytes,
                    uint32_t dst_align, mod_plan** out)
{
    if (!out)
        return fail(MOD_ERR_ARG, "mod_plan_create: out is null");
    *out = nullptr;
    if (n && !descs)
        return fail(MOD_ERR_ARG, "mod_plan_create: descs is null");
    if (n >= 0xFFFFFFFFull)
        return fail(MOD_ERR_ARG, "mod_plan_create: too many descriptors (%llu)", (unsigned long long)n);
    if (dst_align > 15)
        return fail(MOD_ERR_ARG, "mod_plan_create: dst_align must be < 16");
    int rc = ensure_ready(-1);
    if (rc != MOD_OK)
        return rc;

    std::vector<modk::DevDesc> host;
    try {
        host.resize(n);
    } catch (const std::bad_alloc&) {
        return fail(MOD_ERR_NOMEM, "mod_plan_create: host allocation of %llu descriptors failed", (unsigned long long)n);
    }
    uint64_t tiles = 0, payload = 0;
    if ((rc = expand_descs(descs, n, src_bytes, dst_bytes, dst_align, host.data(), &tiles, &payload)) != MOD_OK)
        return rc;

    mod_plan* p = new (std::nothrow) mod_plan();
    if (!p)
        return fail(MOD_ERR_NOMEM, "mod_plan_create: out of host memory");
    p->device = g_ctx.device;
    p->n = n;
    p->src_bytes = src_bytes;
    p->dst_bytes = dst_bytes;
    p->dst_align = dst_align;
    p->payload = payload;
    p->n_tiles = (uint32_t)tiles;
    modk::DevDesc* d_descs = nullptr;  // only needed while the tile records are built
    auto cleanup = [&]() {
        if (d_descs) cudaFree(d_descs);
        if (p->d_tiles) cudaFree(p->d_tiles);
        delete p;
    };
    if (n && tiles) {
        cudaError_t e = cudaMalloc((void**)&d_descs, n * sizeof(modk::DevDesc));
        if (e == cudaSuccess)
            e = cudaMalloc((void**)&p->d_tiles, tiles * sizeof(modk::TileRec));
        if (e != cudaSuccess) {
            cudaGetLastError();
            cleanup();
            return fail(MOD_ERR_NOMEM, "mod_plan_create: cudaMalloc failed: %s", cudaGetErrorString(e));
        }
        e = cudaMemcpy(d_descs, host.data(), n * sizeof(modk::DevDesc), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            e = modk::launch_build_tiles(d_descs, (uint32_t)n, dst_align, p->d_tiles, p->n_tiles, nullptr);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(nullptr);
        if (e != cudaSuccess) {
            cleanup();
            return fail(MOD_ERR_CUDA, "mod_plan_create: %s", cudaGetErrorString(e));
        }
        cudaFree(d_descs);
        d_descs = nullptr;
    }
    *out = p;
    return MOD_OK;
}

int mod_plan_destroy(mod_plan* plan)
{
    if (!plan)
        return MOD_OK;
    if (plan->d_tiles)
        cudaFree(plan->d_tiles);
    delete plan;
    return MOD_OK;
}

uint64_t mod_plan_payload_bytes(const mod_plan* plan) { return plan ? plan->payload : 0; }
uint64_t mod_plan_num_tiles(const mod_plan* plan) { return plan ? plan->n_tiles : 0; }

int mod_plan_run(const mod_plan* plan, const void* d_src, void* d_dst, void* stream)
{
    if (!plan)
        return fail(MOD_ERR_ARG, "mod_plan_run: null plan");
    if (plan->n_tiles == 0)
        return MOD_OK;
    if (!d_src || !d_dst)
        return fail(MOD_ERR_ARG, "mod_plan_run: null buffer");
    if (((uintptr_t)d_dst & 15u) != plan->dst_align)
        return fail(MOD_ERR_ALIGN, "mod_plan_run: dst & 15 is %u but the plan was built for %u",
                    (unsigned)((uintptr_t)d_dst & 15u), plan->dst_align);
    modk::BatchArgs args;
    args.src = (const uint8_t*)d_src;
    args.dst = (uint8_t*)d_dst;
    args.tiles = plan->d_tiles;
    args.n_tiles = plan->n_tiles;
    args.tiles_per_entry = 0;
    args.src_lo16 = ((uint64_t)(uintptr_t)d_src + 15u) & ~15ull;
    args.src_hi16 = ((uint64_t)(uintptr_t)d_src + plan->src_bytes) & ~15ull;
    CUDA_TRY(modk::launch_batch(args, (cudaStream_t)stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return MOD_OK;
}

int mod_cycle_batch(const mod_desc* descs, uint64_t n, const void* src, uint64_t src_bytes, void* dst,
                    uint64_t dst_bytes)
{
    if (n == 0)
        return MOD_OK;
    if (!src || !dst)
        return fail(MOD_ERR_ARG, "mod_cycle_batch: null buffer");
    int rc = ensure_ready(-1);
    if (rc != MOD_OK)
        return rc;
    const bool src_dev = is_device_pointer(src), dst_dev = is_device_pointer(dst);
    if (src_dev != dst_dev)
        return fail(MOD_ERR_ARG, "mod_cycle_batch: src and dst must both be host or both be device pointers");

    mod_plan* plan = nullptr;
    if (src_dev) {
        rc = mod_plan_create(descs, n, src_bytes, dst_bytes, (uint32_t)((uintptr_t)dst & 15u), &plan);
        if (rc != MOD_OK)
            return rc;
        rc = mod_plan_run(plan, src, dst, nullptr);
        cudaError_t e = cudaStreamSynchronize(nullptr);
        mod_plan_destroy(plan);
        if (rc != MOD_OK)
            return rc;
        CUDA_TRY(e);
        return MOD_OK;
    }

    // Host buffers.  The HBM workspaces mirror the host buffers 1:1, the plan (descriptors -> tile
    // records) is built once, and the entries are then streamed in GROUPS of consecutive
    // descriptors (~MOD_GROUP_BYTES of payload each): group g uploads the source window its
    // entries span, runs the batched kernel over its tile sub-range and downloads the destination
    // runs it covers, on stream g % kPipeSlots -- so the upload of one group, the kernel of another
    // and the download of a third overlap.  Only bytes covered by descriptors are written back,
    // so untouched bytes of dst survive exactly like with the reference's per-entry fwrite/fread.
    // The plan lives in grow-only scratch (pinned staging + HBM) and is built asynchronously on pipe
    // stream 0; the other streams wait on an event before their first kernel, the host never does.
    const bool trace = env_u64("MOD_TRACE", 0) != 0;
    const double t_begin = now_ms();
    if (n >= 0xFFFFFFFFull)
        return fail(MOD_ERR_ARG, "mod_cycle_batch: too many descriptors (%llu)", (unsigned long long)n);
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    auto done = [&](int code) { return code; };
    if ((rc = grow_pinned(&g_ctx.h_descs, &g_ctx.h_descs_bytes, n * sizeof(modk::DevDesc))) != MOD_OK)
        return rc;
    uint64_t plan_tiles = 0, plan_payload = 0;
    if ((rc = expand_descs(descs, n, src_bytes, dst_bytes, 0, (modk::DevDesc*)g_ctx.h_descs, &plan_tiles, &plan_payload)) != MOD_OK)
        return rc;
    if (plan_tiles == 0)
        return MOD_OK;
    if ((rc = grow(&g_ctx.ws_descs, &g_ctx.ws_descs_bytes, n * sizeof(modk::DevDesc))) != MOD_OK)
        return rc;
    if ((rc = grow(&g_ctx.ws_tiles, &g_ctx.ws_tiles_bytes, plan_tiles * sizeof(modk::TileRec))) != MOD_OK)
        return rc;
    if ((rc = grow(&g_ctx.ws_src, &g_ctx.ws_src_bytes, src_bytes)) != MOD_OK)
        return rc;
    if ((rc = grow(&g_ctx.ws_dst, &g_ctx.ws_dst_bytes, dst_bytes)) != MOD_OK)
        return rc;
    if (!g_ctx.plan_ready)
        CUDA_TRY(cudaEventCreateWithFlags(&g_ctx.plan_ready, cudaEventDisableTiming));
    CUDA_TRY(cudaMemcpyAsync(g_ctx.ws_descs, g_ctx.h_descs, n * sizeof(modk::DevDesc), cudaMemcpyHostToDevice,
                             g_ctx.pipe_stream[0]));
    CUDA_TRY(modk::launch_build_tiles((const modk::DevDesc*)g_ctx.ws_descs, (uint32_t)n, 0, (modk::TileRec*)g_ctx.ws_tiles,
                                      (uint32_t)plan_tiles, g_ctx.pipe_stream[0]));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaEventRecord(g_ctx.plan_ready, g_ctx.pipe_stream[0]));
    const modk::TileRec* d_tiles = (const modk::TileRec*)g_ctx.ws_tiles;
    const double t_plan = now_ms();

    struct Group {
        uint64_t e0, e1;        // descriptor range
        uint32_t t0, t1;        // tile range
        uint64_t s_lo, s_hi;    // source window
    };
    const uint64_t group_bytes = std::max<uint64_t>(1 << 20, env_u64("MOD_GROUP_BYTES", 16ull << 20));
    std::vector<Group> groups;
    {
        Group g{0, 0, 0, 0, UINT64_MAX, 0};
        uint64_t acc = 0;
        uint32_t tile = 0;
        for (uint64_t i = 0; i < n; ++i) {
            const mod_desc& d = descs[i];
            if (d.len) {
                g.s_lo = std::min(g.s_lo, d.src_off);
                g.s_hi = std::max(g.s_hi, d.src_off + d.len);
            }
            acc += d.len;
            tile += modk::tiles_for_entry((uint32_t)(d.dst_off & 15u), d.len);
            if (acc >= group_bytes || i + 1 == n) {
                g.e1 = i + 1;
                g.t1 = tile;
                if (g.t1 > g.t0)
                    groups.push_back(g);
                g = Group{i + 1, 0, tile, 0, UINT64_MAX, 0};
                acc = 0;
            }
        }
    }
    // If the entries are not laid out in source order the windows overlap heavily and per-group
    // uploads would move the image many times: fall back to a single group then.
    uint64_t window_sum = 0;
    for (const Group& g : groups)
        window_sum += g.s_hi - g.s_lo;
    if (window_sum > src_bytes + src_bytes / 4 + (1 << 20)) {
        Group all{0, n, 0, (uint32_t)plan_tiles, 0, src_bytes};
        groups.assign(1, all);
    }

    // Destination runs (maximal intervals downloaded with one copy) per group.  Two entries that are
    // neighbours in the GLOBAL destination order and less than 16 bytes apart (alignment padding) are
    // bridged into one run; the host bytes of such a gap are saved first and put back after the
    // download, so every byte not covered by a descriptor keeps its value.  A batch whose destination
    // is full of larger holes (more than 64 runs in a group) is handled exactly but without
    // pipelining: the whole destination makes a round trip through HBM.
    std::vector<uint32_t> rank(n, 0);      // position of each non-empty entry in global dst order
    std::vector<uint8_t> bridge_after;     // by rank: the gap to the next entry may be bridged
    {
        std::vector<uint32_t> order;
        order.reserve(n);
        bool monotone = true;
        uint64_t last = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (!descs[i].len)
                continue;
            if (descs[i].dst_off < last)
                monotone = false;
            last = descs[i].dst_off;
            order.push_back((uint32_t)i);
        }
        if (!monotone)
            std::sort(order.begin(), order.end(),
                      [&](uint32_t a, uint32_t b) { return descs[a].dst_off < descs[b].dst_off; });
        bridge_after.assign(order.size(), 0);
        for (size_t r = 0; r < order.size(); ++r) {
            rank[order[r]] = (uint32_t)r;
            if (r + 1 < order.size()) {
                const uint64_t end = descs[order[r]].dst_off + descs[order[r]].len;
                const uint64_t nxt = descs[order[r + 1]].dst_off;
                bridge_after[r] = (nxt >= end && nxt - end < 16) ? 1 : 0;
            }
        }
    }
    struct Gap {
        uint64_t off;
        uint32_t len;
        uint8_t bytes[15];
    };
    std::vector<Gap> gaps;
    std::vector<uint32_t> members;
    auto merged_runs = [&](const Group& g, std::vector<std::pair<uint64_t, uint64_t>>& runs, bool save_gaps) {
        runs.clear();
        members.clear();
        for (uint64_t i = g.e0; i < g.e1; ++i)
            if (descs[i].len)
                members.push_back((uint32_t)i);
        std::sort(members.begin(), members.end(), [&](uint32_t a, uint32_t b) { return rank[a] < rank[b]; });
        uint32_t prev_rank = 0;
        for (uint32_t idx : members) {
            const uint64_t b0 = descs[idx].dst_off, b1 = b0 + descs[idx].len;
            const bool adjacent = !runs.empty() && rank[idx] == prev_rank + 1;
            if (adjacent && b0 <= runs.back().second) {
                runs.back().second = std::max(runs.back().second, b1);
            } else if (adjacent && bridge_after[prev_rank] && b0 - runs.back().second < 16) {
                if (save_gaps && b0 > runs.back().second) {
                    Gap gap;
                    gap.off = runs.back().second;
                    gap.len = (uint32_t)(b0 - runs.back().second);
                    std::memcpy(gap.bytes, (const uint8_t*)dst + gap.off, gap.len);
                    gaps.push_back(gap);
                }
                runs.back().second = b1;
            } else {
                runs.emplace_back(b0, b1);
            }
            prev_rank = rank[idx];
        }
    };
    std::vector<std::pair<uint64_t, uint64_t>> runs;
    bool holes = false;
    for (const Group& g : groups) {
        merged_runs(g, runs, false);
        if (runs.size() > 64) {
            holes = true;
            break;
        }
    }
    if (holes) {
        Group all{0, n, 0, (uint32_t)plan_tiles, 0, src_bytes};
        groups.assign(1, all);
    }

    const uint64_t src_lo16 = ((uint64_t)(uintptr_t)g_ctx.ws_src + 15u) & ~15ull;
    const uint64_t src_hi16 = ((uint64_t)(uintptr_t)g_ctx.ws_src + src_bytes) & ~15ull;
    cudaError_t e = cudaSuccess;
    for (size_t gi = 0; gi < groups.size() && e == cudaSuccess; ++gi) {
        const Group& g = groups[gi];
        cudaStream_t s = g_ctx.pipe_stream[gi % kPipeSlots];
        e = cudaMemcpyAsync((uint8_t*)g_ctx.ws_src + g.s_lo, (const uint8_t*)src + g.s_lo, g.s_hi - g.s_lo,
                            cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && holes)
            e = cudaMemcpyAsync(g_ctx.ws_dst, dst, dst_bytes, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess)
            break;
        if (!holes)
            merged_runs(g, runs, true);
        modk::BatchArgs args;
        args.src = (const uint8_t*)g_ctx.ws_src;
        args.dst = (uint8_t*)g_ctx.ws_dst;
        if (gi > 0 && gi < (size_t)kPipeSlots) {  // first use of this stream: the plan must be complete
            e = cudaStreamWaitEvent(s, g_ctx.plan_ready, 0);
            if (e != cudaSuccess)
                break;
        }
        args.tiles = d_tiles + g.t0;
        args.n_tiles = g.t1 - g.t0;
        args.tiles_per_entry = 0;
        args.src_lo16 = src_lo16;
        args.src_hi16 = src_hi16;
        e = modk::launch_batch(args, s);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (e != cudaSuccess)
            break;
        if (holes) {
            e = cudaMemcpyAsync(dst, g_ctx.ws_dst, dst_bytes, cudaMemcpyDeviceToHost, s);
        } else {
            for (const auto& r : runs) {
                e = cudaMemcpyAsync((uint8_t*)dst + r.first, (const uint8_t*)g_ctx.ws_dst + r.first,
                                    r.second - r.first, cudaMemcpyDeviceToHost, s);
                if (e != cudaSuccess)
                    break;
            }
        }
    }
    const double t_enq = now_ms();
    for (int k = 0; k < kPipeSlots; ++k) {
        cudaError_t es = cudaStreamSynchronize(g_ctx.pipe_stream[k]);
        if (e == cudaSuccess)
            e = es;
    }
    if (e == cudaSuccess)
        for (const Gap& gap : gaps)  // put the padding bytes the bridged downloads ran over back
            std::memcpy((uint8_t*)dst + gap.off, gap.bytes, gap.len);
    if (trace)
        fprintf(stderr, "[mod] cycle_batch: plan %.2f ms, enqueue %zu groups %.2f ms, drain %.2f ms\n", t_plan - t_begin,
                groups.size(), t_enq - t_plan, now_ms() - t_enq);
    if (e != cudaSuccess)
        return done(fail(MOD_ERR_CUDA, "mod_cycle_batch: %s", cudaGetErrorString(e)));
    return done(MOD_OK);
}

/* ---- offset-range sharding (pure host logic) ------------------------------------------------------ */

int mod_shard_range(uint64_t total, int rank, int world, uint64_t* begin, uint64_t* end)
{
    if (world <= 0 || rank < 0 || rank >= world || !begin || !end)
        return fail(MOD_ERR_ARG, "mod_shard_range: bad rank %d / world %d", rank, world);
    const uint64_t share = ((total / (uint64_t)world) + 15u) & ~15ull;
    uint64_t b = std::min(total, share * (uint64_t)rank);
    uint64_t e = (rank == world - 1) ? total : std::min(total, share * (uint64_t)(rank + 1));
    *begin = b;
    *end = e;
    return MOD_OK;
}

int64_t mod_shard_descs(const mod_desc* descs, uint64_t n, int rank, int world, mod_desc* out, uint64_t out_cap)
{
    if (world <= 0 || rank < 0 || rank >= world)
        return fail(MOD_ERR_ARG, "mod_shard_descs: bad rank %d / world %d", rank, world);
    if (n && !descs)
        return fail(MOD_ERR_ARG, "mod_shard_descs: descs is null");
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i)
        total += descs[i].len;
    // this rank owns cumulative payload bytes [lo, hi)
    const uint64_t lo = (uint64_t)(((unsigned __int128)total * (unsigned)rank) / (unsigned)world);
    const uint64_t hi = (rank == world - 1) ? total
                                            : (uint64_t)(((unsigned __int128)total * (unsigned)(rank + 1)) / (unsigned)world);
    // A boundary that falls inside an entry is moved down to the nearest position whose destination
    // offset is 16-byte aligned; the rule depends only on (entry, boundary) so all ranks agree.
    auto cut_at = [](const mod_desc& d, uint64_t raw) -> uint64_t {
        const uint64_t mis = (d.dst_off + raw) & 15u;
        return raw > mis ? raw - mis : 0;
    };
    uint64_t count = 0, cum = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const mod_desc& d = descs[i];
        const uint64_t c0 = cum, c1 = cum + d.len;
        cum = c1;
        if (d.len == 0 || c1 <= lo || c0 >= hi)
            continue;
        const uint64_t b = (lo <= c0) ? 0 : cut_at(d, lo - c0);
        const uint64_t e = (hi >= c1) ? d.len : cut_at(d, hi - c0);
        if (e <= b)
            continue;
        if (out) {
            if (count >= out_cap)
                return fail(MOD_ERR_ARG, "mod_shard_descs: output capacity %llu too small", (unsigned long long)out_cap);
            mod_desc& o = out[count];
            o.src_off = d.src_off + b;
            o.dst_off = d.dst_off + b;
            o.len = (uint32_t)(e - b);
            o.key = b ? mod_key_jump(d.key, b) : d.key;
        }
        ++count;
    }
    return (int64_t)count;
}

}  // extern "C"
