// cycle_kernels.cu -- hand-written sm_100a kernels for the Modulate keystream path.
//
// What the kernel computes (per descriptor {src_off, dst_off, len, key}):
//     dst[dst_off + i] = src[src_off + i] ^ low8(k0 * a^(i+1) mod m) ^ 0xFF,   0 <= i < len
// which is CEncryptionCycler::Cycle (reference CEncryptionCycler.cpp:4-14) applied to a copy of
// the entry -- the copy being CArk::ExtractFiles' gather (CArk.cpp:494) or CArk::BuildArk's
// scatter (CArk.cpp:807-811).  One variable-length batched kernel serves every caller; a
// contiguous Cycle() is the same kernel with a few descriptors passed in the parameter block.
//
// Decomposition (B200: HBM-bound byte work, co-limited by the integer pipes -- no tensor cores):
//   * destination space is cut into 16-byte aligned chunks; one thread owns one chunk per round
//     and moves it with one 128-bit load and one 128-bit store, a warp covering 512 contiguous
//     bytes per round (fully coalesced), kIters rounds per tile;
//   * the serial recurrence is broken by modular jump-ahead: per tile a handful of table
//     multiplies (a^(tile), a^(16*lane), a^(-head)) give each thread the state just before its
//     chunk; rounds advance by the constant a^512; inside the chunk the state is stepped 16
//     times with a lazily reduced Mersenne fold (IMAD.WIDE + one add); low bytes are packed
//     straight from the lazy states and a chunk is redone exactly only if one of its states needed
//     the canonical subtract (~2^-17 per byte); the chains of the 2-4 chunks a thread has in
//     flight are generated as one basic block so they interleave -- about 4 integer instructions
//     per payload byte in total;
//   * entries are byte-packed (BuildArk leaves no padding), so source and destination are
//     generally misaligned with respect to each other: the source is read as aligned 16-byte
//     granules (each lane takes the two granules its chunk straddles; the second is an L1 hit on
//     its neighbour's first) and a funnel shift re-aligns them, the word part of the shift being
//     a template parameter so no register-select network is needed; only the first and last
//     chunk of an entry take a byte path;
//   * the grid is persistent (SM count x resident CTAs): each warp strides over tiles and
//     prefetches the next tile's 32-byte record while it streams the current one, so no warp
//     ever waits on a chain of dependent metadata loads.  The warp count is kept LOW on purpose
//     (24-32 per SM): more bytes in flight cost DRAM efficiency (profiles/r01_tuning.md).
// Alternatives that were built, measured and dropped (register ping-pong, predicate-free group
// copies, cache hints, L2 bulk prefetch) are in the git history and profiles/r01_tuning.md; the
// bulk-async (cp.async.bulk + mbarrier) staging ring is kept behind -DMODK_BULK=1.
#include "cycle_kernels.cuh"
#include "lcg.h"

#include <cstdlib>

namespace modk {

using modlcg::mulmod;
using modlcg::step_lazy;
using modlcg::low8_canonical;

// ---- tuning knobs (defaults chosen from the measurements in profiles/) ------------------------------
#ifndef MODK_UNROLL
#define MODK_UNROLL 4            // batched kernel: independent 16-byte chunks in flight per thread
#endif
#ifndef MODK_UNROLL_INLINE
#define MODK_UNROLL_INLINE 2     // contiguous (inline-descriptor) kernel: chunks in flight per thread
#endif
#ifndef MODK_CANON_FMA_MASK
#define MODK_CANON_FMA_MASK 0x5  // which of every 4 bytes canonicalise on the FMA pipe (IMAD.HI) vs ALU (LEA.HI)
#endif
#ifndef MODK_MIN_CTAS
#define MODK_MIN_CTAS 3          // batched kernel: resident CTAs per SM requested through __launch_bounds__ (80 registers)
#endif
#ifndef MODK_MIN_CTAS_INLINE
#define MODK_MIN_CTAS_INLINE 4   // contiguous kernel: resident CTAs per SM (64 registers)
#endif
#ifndef MODK_GRID_MODE
#define MODK_GRID_MODE 0         // 0: persistent grid (SMs x resident CTAs), 1: one tile per warp, CTAs retire
#endif
#ifndef MODK_PIPELINE
#define MODK_PIPELINE 0          // batched kernel: 1 = two-stage register pipeline (next group's loads before this group's cipher)
#endif
#ifndef MODK_PIPELINE_INLINE
#define MODK_PIPELINE_INLINE 0   // contiguous kernel: same
#endif
#ifndef MODK_INTERLEAVE
#define MODK_INTERLEAVE 1        // generate the keystream of a whole load group as one basic block (ILP across chunks)
#endif
#ifndef MODK_SPECULATE
#define MODK_SPECULATE 1         // pack low bytes from lazy states, redo the ~1/8000 chunks that needed a canonical subtract
#endif
#ifndef MODK_BULK
#define MODK_BULK 0              // 1: stage the source through a per-warp shared-memory ring filled by bulk-async copies (TMA 1-D)
#endif
#ifndef MODK_STAGES
#define MODK_STAGES 2            // ring depth per warp (load groups in flight) when MODK_BULK
#endif

static_assert(kIters % (MODK_UNROLL * (MODK_PIPELINE ? 2 : 1)) == 0 && kIters % MODK_UNROLL_INLINE == 0,
              "rounds per tile must be a multiple of the unroll (twice the unroll when pipelined)");

// tile index within an entry < 2^32 / kTileBytes + 1; split 10 bits low / rest high
constexpr int kTw0Size = 1024;
constexpr int kTw1Size = (int)(((1ull << 32) / kTileBytes) / kTw0Size + 2);

__constant__ uint32_t c_tw0[kTw0Size];  // a^(kTileBytes * j)
__constant__ uint32_t c_tw1[kTw1Size];  // a^(kTileBytes * 1024 * j)
__constant__ uint32_t c_ainv[16];       // a^(-h): rewinds the stream to the chunk grid origin
__constant__ uint32_t c_round_pow[kIters];  // a^(512 * r): r rounds into a tile (inline kernel's short tiles)
__device__ uint32_t g_chunk_pow[kChunksPerTile];  // a^(16 * j): lane-divergent index, so HBM/L1 rather than the constant bank

constexpr uint32_t kRoundJump = modlcg::pow_a(512);  // one round = 32 lanes x 16 bytes further down the stream

// ---- 128-bit global accesses (explicit state space: the addresses are rebuilt from integers) ------

__device__ __forceinline__ uint4 ldg128(uint64_t addr)
{
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(addr));
    return r;
}

__device__ __forceinline__ void stg128(uint64_t addr, const uint4& v)
{
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// ---- per-chunk arithmetic -------------------------------------------------------------------

// low8_canonical on the FMA pipe: hi32(s * 2) + s == (s >> 31) + s.  The integer pipes are the
// co-limiter of this kernel and the ALU pipe carries the fold and the byte packing, so the
// canonicalisation is issued as an IMAD.HI instead of a second LEA.HI.
// `two` is the constant 2 delivered through the kernel parameter block: with a literal, ptxas
// strength-reduces the multiply back into an ALU-pipe LEA.HI.
__device__ __forceinline__ uint32_t low8_canonical_fma(uint32_t s, uint32_t two)
{
    uint32_t r;
    asm("mad.hi.u32 %0, %1, %2, %1;" : "=r"(r) : "r"(s), "r"(two));
    return r;
}

// XOR the 16 bytes of `d` with the keystream that follows (negated) state `s`, the state just
// before the chunk's first byte -- exact for every state.  Per byte: IMAD.WIDE + LEA.HI (step) and
// a canonical low byte (LEA.HI on the ALU pipe or IMAD.HI on the FMA pipe, per MODK_CANON_FMA_MASK);
// per word: three PRMTs pack four keystream bytes and one LOP3 applies them.
__device__ __forceinline__ uint4 cycle_chunk_exact(uint4 d, uint32_t s, const uint32_t two)
{
    uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s = step_lazy(s);
        const uint32_t b0 = (MODK_CANON_FMA_MASK & 1) ? low8_canonical_fma(s, two) : low8_canonical(s);
        s = step_lazy(s);
        const uint32_t b1 = (MODK_CANON_FMA_MASK & 2) ? low8_canonical_fma(s, two) : low8_canonical(s);
        s = step_lazy(s);
        const uint32_t b2 = (MODK_CANON_FMA_MASK & 4) ? low8_canonical_fma(s, two) : low8_canonical(s);
        s = step_lazy(s);
        const uint32_t b3 = (MODK_CANON_FMA_MASK & 8) ? low8_canonical_fma(s, two) : low8_canonical(s);
        const uint32_t lo = __byte_perm(b0, b1, 0x0040);
        const uint32_t hi = __byte_perm(b2, b3, 0x0040);
        w[j] ^= __byte_perm(lo, hi, 0x5410);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Out-of-line copy of the exact form for the rare chunks the speculative form below hands over.
__device__ __noinline__ uint4 cycle_chunk_rare(uint4 d, uint32_t s, uint32_t two)
{
    return cycle_chunk_exact(d, s, two);
}

// Speculative form used in the hot loop.  A lazily reduced state t = hi + lo31 (hi <= 16807) is
// already canonical unless bit 31 is set, which needs lo31 >= 2^31 - 16807: about 2^-17 per byte.
// So the 16 low bytes are packed straight from the lazy states, the states are OR-ed together
// (8 three-input LOP3 instead of 16 canonicalisations), and only if some state had bit 31 set is
// the chunk redone by the exact routine (about one chunk in 8 000).
__device__ __forceinline__ uint4 cycle_chunk(uint4 d, uint32_t s, const uint32_t two)
{
#if defined(MODK_EXPERIMENT_COPY_ONLY)
    // measurement aid (never shipped): no keystream at all -> the memory-system ceiling of this
    // kernel's access pattern
    d.x ^= s & two;
    return d;
#elif MODK_SPECULATE
    const uint32_t s0 = s;
    uint32_t w[4] = {d.x, d.y, d.z, d.w};
    uint32_t any = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t b0 = step_lazy(s);
        const uint32_t b1 = step_lazy(b0);
        const uint32_t b2 = step_lazy(b1);
        const uint32_t b3 = step_lazy(b2);
        s = b3;
        any |= b0 | b1;
        any |= b2 | b3;
        const uint32_t lo = __byte_perm(b0, b1, 0x0040);
        const uint32_t hi = __byte_perm(b2, b3, 0x0040);
        w[j] ^= __byte_perm(lo, hi, 0x5410);
    }
    if (__builtin_expect((int32_t)any < 0, 0))
        return cycle_chunk_rare(d, s0, two);
    return make_uint4(w[0], w[1], w[2], w[3]);
#else
    return cycle_chunk_exact(d, s, two);
#endif
}

// The 16 keystream bytes that follow state `s`, as four little-endian words (no data involved, so
// the bulk-staged path computes them while its source bytes are still in flight).
__device__ __forceinline__ uint4 keystream_chunk(uint32_t s, const uint32_t two)
{
    return cycle_chunk(make_uint4(0u, 0u, 0u, 0u), s, two);
}

#if MODK_BULK
// ---- bulk-async staging: mbarrier + cp.async.bulk (SASS: UBLKCP / SYNCS) --------------------------------
constexpr int kStages = MODK_STAGES;
constexpr uint32_t kStageBytes = 16u * (32u * (MODK_UNROLL > MODK_UNROLL_INLINE ? MODK_UNROLL : MODK_UNROLL_INLINE) + 1u);  // one load group + the straddled granule

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MODK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MODK_DONE;\n"
        "bra MODK_WAIT;\n"
        "MODK_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned); completion
// is signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, uint64_t src_gmem, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

// Per-warp ring state that lives across tiles (mbarrier phases must keep counting).
struct WarpRing {
    uint32_t data;    // shared address of stage 0
    uint32_t bars;    // shared address of mbarrier 0
    uint32_t slot;    // next stage to fill / drain (fill and drain advance in lock step per tile)
    uint32_t phases;  // bit s = parity to wait for on stage s
};
#endif

// ---- one tile = one warp ------------------------------------------------------------------------

struct TileGeom {
    uint64_t dst_al;    // 16-byte aligned address of chunk 0
    uint64_t src_al;    // 16-byte aligned address of the granule holding chunk 0's first source byte
    uint64_t dst_addr;  // address of the entry's first destination byte
    uint64_t src_addr;  // address of the entry's first source byte
    uint32_t len;
    uint32_t h0;        // dst_addr & 15
    uint32_t shift;     // byte offset of chunk data inside its first source granule (0 = co-aligned)
    uint32_t c_begin, c_end;  // chunk range of this tile
    uint32_t f_lo, f_hi;      // chunks in [f_lo, f_hi) are interior: whole-granule loads, 128-bit store
};

// Edge chunk (first / last chunk of an entry, or one whose source granules would leave the
// source buffer): byte-granular and predicated.  Out of line -- at most a couple per entry.
__device__ __noinline__ void edge_chunk(const uint8_t* src_entry, uint8_t* dst_entry, long long pos0,
                                        uint32_t len, uint32_t s, uint32_t two)
{
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const long long pos = pos0 + b;
        if (pos >= 0 && pos < (long long)len)
            w[b >> 2] |= (uint32_t)src_entry[pos] << (8 * (b & 3));
    }
    const uint4 o = cycle_chunk(make_uint4(w[0], w[1], w[2], w[3]), s, two);
    const uint32_t r[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const long long pos = pos0 + b;
        if (pos >= 0 && pos < (long long)len)
            dst_entry[pos] = (uint8_t)(r[b >> 2] >> (8 * (b & 3)));
    }
}

// Interior chunks of the tile: kUnroll independent chunks in flight per thread (all loads of a
// group are issued before any is consumed).  A register ping-pong that issued the next group's
// loads before ciphering the current one, a bulk-async shared-memory ring and bulk L2 prefetches
// all measured SLOWER on B200 (profiles/r01_tuning.md), so the simple form stays: latency is
// covered by the 32 resident warps per SM.
// kWs < 0: source and destination are co-aligned (one load per chunk).  kWs in 0..3: the chunk
// starts kWs words (+ a runtime 0..3 bytes) into its first granule and straddles two.
// kFull: every chunk of the group is interior, so there are no per-lane predicates and all
// addresses are one 64-bit pointer per lane plus immediates.
// The source granules of one load group (kUnroll rounds) held in registers, and the two halves of
// working on them: `load` (global memory, or a shared-memory stage filled by a bulk-async copy) and
// `finish` (keystream, re-alignment, XOR, store).
template <int kWs, int kUnroll>
struct GroupRegs {
    uint4 own[kUnroll];
    uint4 nxt[kUnroll];

    // every load of the group is issued before anything consumes one
    template <bool kSmem>
    __device__ __forceinline__ void load(const TileGeom& g, const uint32_t base, const uint32_t m_hi,
                                         const uint32_t lane, const uint32_t stage)
    {
        const uint64_t sp = g.src_al + 16ull * (base + lane);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t c = base + (uint32_t)u * 32u + lane;
            own[u] = make_uint4(0u, 0u, 0u, 0u);
            nxt[u] = make_uint4(0u, 0u, 0u, 0u);
            if ((c >= g.f_lo) && (c < m_hi)) {
#if MODK_BULK
                if (kSmem) {
                    own[u] = lds128(stage + 16u * (32u * (uint32_t)u + lane));
                    if (kWs >= 0)
                        nxt[u] = lds128(stage + 16u * (32u * (uint32_t)u + lane) + 16u);
                } else
#endif
                {
                    own[u] = ldg128(sp + 512ull * u);
                    if (kWs >= 0)
                        nxt[u] = ldg128(sp + 512ull * u + 16ull);
                }
            }
        }
    }

    __device__ __forceinline__ uint4 aligned(const int u, const uint32_t bs) const
    {
        uint4 data = own[u];
        if (kWs >= 0) {
            const uint32_t w[8] = {own[u].x, own[u].y, own[u].z, own[u].w, nxt[u].x, nxt[u].y, nxt[u].z, nxt[u].w};
            constexpr int k = kWs < 0 ? 0 : kWs;
            data.x = __funnelshift_r(w[k + 0], w[k + 1], bs);
            data.y = __funnelshift_r(w[k + 1], w[k + 2], bs);
            data.z = __funnelshift_r(w[k + 2], w[k + 3], bs);
            data.w = __funnelshift_r(w[k + 3], w[k + 4], bs);
        }
        return data;
    }

    // returns the state advanced by kUnroll rounds
    __device__ __forceinline__ uint32_t finish(const TileGeom& g, const uint32_t base, const uint32_t m_hi, uint32_t v,
                                               const uint32_t lane, const uint32_t bs, const uint32_t two) const
    {
        const uint64_t dp = g.dst_al + 16ull * (base + lane);
#if MODK_INTERLEAVE
        // the kUnroll keystream chunks of the group, generated as ONE basic block so that the
        // independent 16-step chains interleave (ILP = kUnroll) instead of running one after another;
        // low bytes are packed straight from the lazy states and the states are OR-ed (see cycle_chunk)
        uint32_t st[kUnroll];
        uint32_t v0[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            v0[u] = v;
            st[u] = v;
            v = mulmod(v, kRoundJump);  // state just before this lane's chunk of the next round
        }
        uint32_t ks[kUnroll][4];
        uint32_t any = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint32_t b0 = step_lazy(st[u]);
                const uint32_t b1 = step_lazy(b0);
                const uint32_t b2 = step_lazy(b1);
                const uint32_t b3 = step_lazy(b2);
                st[u] = b3;
                any |= b0 | b1;
                any |= b2 | b3;
                ks[u][j] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t c = base + (uint32_t)u * 32u + lane;
            const uint4 data = aligned(u, bs);
            uint4 out = make_uint4(data.x ^ ks[u][0], data.y ^ ks[u][1], data.z ^ ks[u][2], data.w ^ ks[u][3]);
            if (__builtin_expect((int32_t)any < 0, 0))  // some state of the group needed a canonical subtract: redo exactly
                out = cycle_chunk_rare(data, v0[u], two);
            if ((c >= g.f_lo) && (c < m_hi))
                stg128(dp + 512ull * u, out);
        }
#else
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t c = base + (uint32_t)u * 32u + lane;
            if ((c >= g.f_lo) && (c < m_hi))
                stg128(dp + 512ull * u, cycle_chunk(aligned(u, bs), v, two));
            v = mulmod(v, kRoundJump);
        }
#endif
        return v;
    }
};

// Interior chunks of the tile.  kWs < 0: source and destination are co-aligned (one load per
// chunk).  kWs in 0..3: the chunk starts kWs words (+ a runtime 0..3 bytes) into its first granule
// and straddles two.  kPipelined: two register stages -- the loads of group i+1 are issued before
// group i is ciphered and stored, so a warp has loads in flight during its integer work (the loop
// body is the two-stage ping-pong written out so both register sets have static names).
template <int kWs, int kUnroll, bool kPipelined>
__device__ __forceinline__ void process_interior(const TileGeom& g, uint32_t v, const uint32_t lane,
                                                 const uint32_t two)
{
    const uint32_t bs = (g.shift & 3u) * 8u;
    const uint32_t m_hi = min(g.c_end, g.f_hi);
    constexpr uint32_t kGroupChunks = 32u * kUnroll;

    if (kPipelined) {
        GroupRegs<kWs, kUnroll> ra, rb;
        uint32_t base = g.c_begin;
        if (base < m_hi)
            ra.template load<false>(g, base, m_hi, lane, 0u);
#pragma unroll 1
        for (; base < m_hi; base += 2u * kGroupChunks) {
            rb.template load<false>(g, base + kGroupChunks, m_hi, lane, 0u);  // all predicates false past the tile
            v = ra.finish(g, base, m_hi, v, lane, bs, two);
            ra.template load<false>(g, base + 2u * kGroupChunks, m_hi, lane, 0u);
            v = rb.finish(g, base + kGroupChunks, m_hi, v, lane, bs, two);
        }
    } else {
#pragma unroll 1
        for (uint32_t base = g.c_begin; base < m_hi; base += kGroupChunks) {
            GroupRegs<kWs, kUnroll> r;
            r.template load<false>(g, base, m_hi, lane, 0u);
            v = r.finish(g, base, m_hi, v, lane, bs, two);
        }
    }
}

#if MODK_BULK
// Interior chunks through a per-warp shared-memory ring (MODK_BULK): lane 0 keeps kStages load
// groups (kUnroll rounds = 512 * kUnroll bytes each, +16 when chunks straddle two granules) in flight
// with ONE bulk-async copy per group; per group the warp waits for its stage, pulls it into
// registers with LDS.128, hands the stage straight back to be refilled kStages groups ahead, and
// only then runs the (interleaved) integer work and the stores.  Loads in flight cost shared
// memory instead of registers, and their number is fixed by the ring, not by the warp count.
template <int kWs, int kUnroll>
__device__ __forceinline__ void process_interior_bulk(const TileGeom& g, uint32_t v, const uint32_t lane,
                                                      const uint32_t two, WarpRing& ring)
{
    const uint32_t bs = (g.shift & 3u) * 8u;
    const uint32_t m_hi = min(g.c_end, g.f_hi);
    const uint32_t lo = max(g.c_begin, g.f_lo);
    if (lo >= m_hi)
        return;
    constexpr uint32_t kGroupChunks = 32u * kUnroll;
    constexpr uint32_t kExtra = (kWs >= 0) ? 1u : 0u;
    const uint32_t n_groups = (m_hi - g.c_begin + kGroupChunks - 1u) / kGroupChunks;

    auto issue = [&](uint32_t gi, uint32_t stage) {
        const uint32_t base = g.c_begin + kGroupChunks * gi;
        const uint32_t a = max(base, lo), b = min(base + kGroupChunks, m_hi);
        if (lane == 0 && b > a) {
            const uint32_t bytes = 16u * (b - a + kExtra);
            const uint32_t bar = ring.bars + 8u * stage;
            mbar_expect_tx(bar, bytes);
            bulk_g2s(ring.data + stage * kStageBytes + 16u * (a - base), g.src_al + 16ull * a, bytes, bar);
        }
    };

    const uint32_t first_slot = ring.slot;
    const uint32_t pre = min(n_groups, (uint32_t)kStages);
    for (uint32_t gi = 0; gi < pre; ++gi)
        issue(gi, (first_slot + gi) % (uint32_t)kStages);

#pragma unroll 1
    for (uint32_t gi = 0; gi < n_groups; ++gi) {
        const uint32_t stage = (first_slot + gi) % (uint32_t)kStages;
        const uint32_t base = g.c_begin + kGroupChunks * gi;
        if (max(base, lo) < min(base + kGroupChunks, m_hi)) {  // warp-uniform: this group fetched something
            mbar_wait(ring.bars + 8u * stage, (ring.phases >> stage) & 1u);
            ring.phases ^= 1u << stage;
        }
        GroupRegs<kWs, kUnroll> r;
        r.template load<true>(g, base, m_hi, lane, ring.data + stage * kStageBytes);
        __syncwarp();  // every lane holds its granules in registers: the stage may be overwritten
        if (gi + (uint32_t)kStages < n_groups)
            issue(gi + (uint32_t)kStages, stage);
        v = r.finish(g, base, m_hi, v, lane, bs, two);
    }
    ring.slot = (first_slot + n_groups) % (uint32_t)kStages;
}
#endif

// (negated) state just before byte (16 * tin * kChunksPerTile - h0) of an entry:
//   n0 * a^(-h0) * a^(kTileBytes * tin)
__device__ __forceinline__ uint32_t tile_start_state(int32_t key, uint32_t h0, uint32_t tin)
{
    uint32_t st = modlcg::key_to_neg_state(key);
    st = mulmod(st, c_ainv[h0]);
    return mulmod(st, mulmod(c_tw0[tin & (uint32_t)(kTw0Size - 1)], c_tw1[tin / (uint32_t)kTw0Size]));
}

#if MODK_BULK
#define MODK_RING_PARAM , WarpRing& ring
#define MODK_RING_ARG , ring
#define MODK_INTERIOR(K) process_interior_bulk<K, kUnroll>(g, v, lane, a.two, ring)
#else
#define MODK_RING_PARAM
#define MODK_RING_ARG
#define MODK_INTERIOR(K) process_interior<K, kUnroll, kPipelined>(g, v, lane, a.two)
#endif

template <int kUnroll, bool kPipelined>
__device__ __forceinline__ void run_tile(const BatchArgs& a, const uint64_t src_off, const uint64_t dst_off,
                                         const uint32_t len, const uint32_t st, const uint32_t c_begin,
                                         const uint32_t tile_chunks, const uint32_t lane MODK_RING_PARAM)
{
    TileGeom g;
    g.len = len;
    g.dst_addr = (uint64_t)a.dst + dst_off;
    g.src_addr = (uint64_t)a.src + src_off;
    g.h0 = (uint32_t)g.dst_addr & 15u;
    g.dst_al = g.dst_addr - g.h0;
    const uint64_t sv = g.src_addr - g.h0;  // source address that pairs with chunk 0, byte 0
    g.shift = (uint32_t)sv & 15u;
    g.src_al = sv - g.shift;

    const uint64_t span = (uint64_t)g.h0 + len;  // bytes from the chunk grid origin to the entry end
    const uint32_t nchunks = (uint32_t)((span + 15u) >> 4);
    g.c_begin = c_begin;
    g.c_end = min(c_begin + tile_chunks, nchunks);

    // interior chunks: all 16 destination bytes belong to the entry, and the one or two source
    // granules they need lie wholly inside the source buffer
    long long f_lo = g.h0 ? 1 : 0;
    long long f_hi = (long long)(span >> 4);
    const long long g_lo = ((long long)(a.src_lo16 - g.src_al)) >> 4;
    const long long g_hi = (((long long)(a.src_hi16 - g.src_al)) >> 4) - (g.shift ? 1 : 0);
    f_lo = max(f_lo, g_lo);
    f_hi = min(f_hi, g_hi);
    f_hi = max(f_hi, 0ll);
    f_lo = min(f_lo, f_hi);
    g.f_lo = (uint32_t)f_lo;
    g.f_hi = (uint32_t)f_hi;

    // edge chunks first (rare: skipped for tiles that are interior throughout).  The chunks below
    // f_lo and those from f_hi on are enumerated as one compact list, one lane each, so the head
    // and tail edge of a small entry are handled in the same pass.
    if (g.c_begin < g.f_lo || g.c_end > g.f_hi) {
        const uint32_t n_head = g.f_lo > g.c_begin ? min(g.c_end, g.f_lo) - g.c_begin : 0u;
        const uint32_t tail0 = max(g.c_begin, g.f_hi);
        const uint32_t n_tail = g.c_end > tail0 ? g.c_end - tail0 : 0u;
        for (uint32_t i = lane; i < n_head + n_tail; i += 32u) {
            const uint32_t c = i < n_head ? g.c_begin + i : tail0 + (i - n_head);
            edge_chunk(reinterpret_cast<const uint8_t*>(g.src_addr), reinterpret_cast<uint8_t*>(g.dst_addr),
                       16ll * (long long)c - (long long)g.h0, g.len,
                       mulmod(st, g_chunk_pow[c - g.c_begin]), a.two);
        }
    }

    const uint32_t v = mulmod(st, g_chunk_pow[lane]);
    if (g.shift == 0u) {
        MODK_INTERIOR(-1);
    } else {
        switch (g.shift >> 2) {
        case 0: MODK_INTERIOR(0); break;
        case 1: MODK_INTERIOR(1); break;
        case 2: MODK_INTERIOR(2); break;
        default: MODK_INTERIOR(3); break;
        }
    }
}

__device__ __forceinline__ TileRec load_tile_rec(const TileRec* p)
{
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    TileRec r;
    r.src_off = (uint64_t)lo.x | ((uint64_t)lo.y << 32);
    r.dst_off = (uint64_t)lo.z | ((uint64_t)lo.w << 32);
    r.len = hi.x;
    r.state = hi.y;
    r.tin = hi.z;
    r.pad = hi.w;
    return r;
}

#if MODK_BULK
// One ring per warp; its lane 0 initialises the mbarriers (one arrival each: the expect_tx).
#define MODK_RING_SETUP                                                                              \
    __shared__ __align__(128) uint8_t s_ring[kWarpsPerCta][kStages][kStageBytes];                    \
    __shared__ __align__(8) uint64_t s_bars[kWarpsPerCta][kStages];                                  \
    WarpRing ring;                                                                                   \
    ring.data = smem_u32(&s_ring[threadIdx.x >> 5][0][0]);                                           \
    ring.bars = smem_u32(&s_bars[threadIdx.x >> 5][0]);                                              \
    ring.slot = 0;                                                                                   \
    ring.phases = 0;                                                                                 \
    if ((threadIdx.x & 31u) == 0) {                                                                  \
        for (int i = 0; i < kStages; ++i)                                                            \
            mbar_init(ring.bars + 8u * i, 1u);                                                       \
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                           \
    }                                                                                                \
    __syncthreads();
#else
#define MODK_RING_SETUP
#endif

// Persistent batched kernel: warp w of the grid takes tiles w, w + W, w + 2W, ... and loads the
// record of its next tile before it starts streaming the current one.
__global__ void __launch_bounds__(kThreadsPerCta, MODK_MIN_CTAS) cycle_batch_kernel(const BatchArgs a)
{
    MODK_RING_SETUP
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * (uint32_t)kWarpsPerCta;
    uint32_t tile = blockIdx.x * (uint32_t)kWarpsPerCta + (threadIdx.x >> 5);
    if (tile >= a.n_tiles)
        return;
    TileRec cur = load_tile_rec(a.tiles + tile);
    for (;;) {
        const bool more = (a.n_tiles - tile) > stride;
        TileRec nxt = cur;
        if (more)
            nxt = load_tile_rec(a.tiles + tile + stride);
        run_tile<MODK_UNROLL, (MODK_PIPELINE != 0)>(a, cur.src_off, cur.dst_off, cur.len, cur.state, cur.tin * (uint32_t)kChunksPerTile,
                 (uint32_t)kChunksPerTile, lane MODK_RING_ARG);
        if (!more)
            break;
        cur = nxt;
        tile += stride;
    }
}

// Same kernel with the (few) descriptors in the parameter block: nothing to upload, nothing to
// allocate, so a contiguous Cycle() is a single asynchronous launch.  The per-tile state comes
// from the constant-bank jump tables instead of a tile record.
__global__ void __launch_bounds__(kThreadsPerCta, MODK_MIN_CTAS_INLINE)
cycle_inline_kernel(const BatchArgs a, const __grid_constant__ InlineDescs in)
{
    MODK_RING_SETUP
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * (uint32_t)kWarpsPerCta;
    for (uint32_t tile = blockIdx.x * (uint32_t)kWarpsPerCta + (threadIdx.x >> 5); tile < a.n_tiles;) {
        const uint32_t e = tile / a.tiles_per_entry;
        const DevDesc& d = in.d[e];
        // inline tiles are a.rounds_per_tile rounds long (short buffers are spread over more warps):
        // start from the enclosing full-size tile's state and jump the remaining rounds
        const uint32_t round0 = (tile - d.first_tile) * a.rounds_per_tile;
        const uint32_t h0 = (uint32_t)((uint64_t)a.dst + d.dst_off) & 15u;
        const uint32_t st = mulmod(tile_start_state(d.key, h0, round0 / (uint32_t)kIters),
                                   c_round_pow[round0 % (uint32_t)kIters]);
        run_tile<MODK_UNROLL_INLINE, (MODK_PIPELINE_INLINE != 0)>(a, d.src_off, d.dst_off, d.len, st, round0 * 32u, a.rounds_per_tile * 32u, lane MODK_RING_ARG);
        if ((a.n_tiles - tile) <= stride)
            break;
        tile += stride;
    }
}

// Plan kernel, one thread per tile: find the tile's entry (the last entry whose first_tile <= tile;
// entries with no tiles share their successor's first_tile and are skipped by taking the last) and
// write the tile record, jump-ahead state included.
__global__ void build_tiles_kernel(const DevDesc* __restrict__ descs, uint32_t n_descs, uint32_t dst_align,
                                   TileRec* __restrict__ tiles, uint32_t n_tiles)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles)
        return;
    uint32_t lo = 0, hi = n_descs;  // invariant: first_tile[lo] <= t < first_tile[hi] (hi == n: sentinel)
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (descs[mid].first_tile <= t)
            lo = mid;
        else
            hi = mid;
    }
    const DevDesc d = descs[lo];
    TileRec r;
    r.src_off = d.src_off;
    r.dst_off = d.dst_off;
    r.len = d.len;
    r.tin = t - d.first_tile;
    r.state = tile_start_state(d.key, (uint32_t)((dst_align + d.dst_off) & 15u), r.tin);
    r.pad = lo;
    tiles[t] = r;
}

// ---- host side --------------------------------------------------------------------------------------

cudaError_t upload_tables()
{
    static uint32_t h_tw0[kTw0Size], h_tw1[kTw1Size], h_ainv[16], h_chunk[kChunksPerTile], h_round[kIters];
    static bool built = false;
    if (!built) {
        for (int j = 0; j < kTw0Size; ++j)
            h_tw0[j] = modlcg::pow_a((uint64_t)kTileBytes * (uint64_t)j);
        for (int j = 0; j < kTw1Size; ++j)
            h_tw1[j] = modlcg::pow_a((uint64_t)kTileBytes * (uint64_t)kTw0Size * (uint64_t)j);
        for (int h = 0; h < 16; ++h)
            h_ainv[h] = modlcg::pow_a_inv((uint64_t)h);
        for (int j = 0; j < kChunksPerTile; ++j)
            h_chunk[j] = modlcg::pow_a(16ull * (uint64_t)j);
        for (int r = 0; r < kIters; ++r)
            h_round[r] = modlcg::pow_a(512ull * (uint64_t)r);
        built = true;
    }
    cudaError_t err;
    if ((err = cudaMemcpyToSymbol(c_tw0, h_tw0, sizeof(h_tw0))) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_tw1, h_tw1, sizeof(h_tw1))) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_ainv, h_ainv, sizeof(h_ainv))) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(g_chunk_pow, h_chunk, sizeof(h_chunk))) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_round_pow, h_round, sizeof(h_round))) != cudaSuccess) return err;
    return cudaSuccess;
}

cudaError_t persistent_grid(int* grid_out, bool inline_kernel)
{
    static int cached[2][64] = {};
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess)
        return err;
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    int& slot = cached[inline_kernel ? 1 : 0][dev];
    if (slot == 0) {
        int sms = 0, per_sm = 0;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
        err = inline_kernel ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cycle_inline_kernel, kThreadsPerCta, 0)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cycle_batch_kernel, kThreadsPerCta, 0);
        if (err != cudaSuccess)
            return err;
        slot = sms * (per_sm > 0 ? per_sm : 1);
    }
    *grid_out = slot;
    return cudaSuccess;
}

static cudaError_t grid_for_tiles(uint32_t n_tiles, unsigned* grid, bool inline_kernel)
{
    int cap = 0;
    cudaError_t err = persistent_grid(&cap, inline_kernel);
    if (err != cudaSuccess)
        return err;
    const unsigned want = (unsigned)((n_tiles + (uint32_t)kWarpsPerCta - 1u) / (uint32_t)kWarpsPerCta);
    static const int mode = []() {
        const char* v = getenv("MOD_GRID_MODE");  // tuning aid: 1 = one tile per warp, CTAs retire (non-persistent)
        return v ? atoi(v) : MODK_GRID_MODE;
    }();
    *grid = (mode == 1 || want < (unsigned)cap) ? want : (unsigned)cap;
    return cudaSuccess;
}

cudaError_t launch_batch(const BatchArgs& args, cudaStream_t stream)
{
    if (args.n_tiles == 0)
        return cudaSuccess;
    unsigned grid = 0;
    cudaError_t err = grid_for_tiles(args.n_tiles, &grid, false);
    if (err != cudaSuccess)
        return err;
    cycle_batch_kernel<<<grid, kThreadsPerCta, 0, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_batch_inline(const BatchArgs& args, const InlineDescs& descs, cudaStream_t stream)
{
    if (args.n_tiles == 0)
        return cudaSuccess;
    unsigned grid = 0;
    cudaError_t err = grid_for_tiles(args.n_tiles, &grid, true);
    if (err != cudaSuccess)
        return err;
    cycle_inline_kernel<<<grid, kThreadsPerCta, 0, stream>>>(args, descs);
    return cudaGetLastError();
}

cudaError_t launch_build_tiles(const DevDesc* descs, uint32_t n_descs, uint32_t dst_align, TileRec* tiles,
                               uint32_t n_tiles, cudaStream_t stream)
{
    if (n_tiles == 0)
        return cudaSuccess;
    const unsigned threads = 256;
    build_tiles_kernel<<<(n_tiles + threads - 1) / threads, threads, 0, stream>>>(descs, n_descs, dst_align,
                                                                                 tiles, n_tiles);
    return cudaGetLastError();
}

}  // namespace modk
