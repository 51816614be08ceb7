// tools/copy_skeleton_bench.cu -- which data-movement skeleton reaches the HBM copy peak on B200?
// Out-of-place copy of 1 GiB (read 1 GiB + write 1 GiB), several work decompositions, CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/copy_skeleton_bench tools/copy_skeleton_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

// A: grid-stride, thread-interleaved: consecutive threads consecutive 16 B, U loads in flight,
//    the U loads of a thread are (total threads * 16 B) apart.  Persistent grid.
template <int U>
__global__ void k_gridstride(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n)
{
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += T * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * T < n) v[u] = s[i + u * T];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * T < n) d[i + u * T] = v[u];
    }
}

// B: block-contiguous, non-persistent: block b owns [b*B*U, (b+1)*B*U) granules; U loads per thread
//    blockDim*16 B apart.
template <int U>
__global__ void k_blockchunk(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n)
{
    const size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (base + (size_t)u * blockDim.x < n) v[u] = s[base + (size_t)u * blockDim.x];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (base + (size_t)u * blockDim.x < n) d[base + (size_t)u * blockDim.x] = v[u];
}

// C: the cipher kernel's skeleton: persistent warps, warp-private tiles of R rounds (512 B each),
//    U rounds in flight, warp w takes tiles w, w+W, ...
template <int U, int R>
__global__ void k_warptile(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n)
{
    const unsigned lane = threadIdx.x & 31;
    const size_t W = (size_t)gridDim.x * (blockDim.x >> 5);
    const size_t tiles = (n + 32 * R - 1) / (32 * R);
    for (size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < tiles; t += W) {
        const size_t c0 = t * 32 * R + lane;
#pragma unroll 1
        for (int r = 0; r < R; r += U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c0 + (size_t)(r + u) * 32 < n) v[u] = s[c0 + (size_t)(r + u) * 32];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c0 + (size_t)(r + u) * 32 < n) d[c0 + (size_t)(r + u) * 32] = v[u];
        }
    }
}

// D: like C, but the 8 warps of a CTA interleave at GROUP granularity inside a CTA-wide tile of
//    8*R rounds: in step j warp w moves group (j*8 + w), so a CTA touches 8*U*512 contiguous bytes per step.
template <int U, int R>
__global__ void k_ctatile(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t tile_chunks = (size_t)32 * R * nw;
    const size_t tiles = (n + tile_chunks - 1) / tile_chunks;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
#pragma unroll 1
        for (int j = 0; j < R / U; ++j) {
            const size_t c0 = t * tile_chunks + ((size_t)j * nw + warp) * 32 * U + lane;
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c0 + (size_t)u * 32 < n) v[u] = s[c0 + (size_t)u * 32];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c0 + (size_t)u * 32 < n) d[c0 + (size_t)u * 32] = v[u];
        }
    }
}

// E: C with a register software pipeline: the loads of group i+1 are issued before the stores of group i.
template <int U, int R>
__global__ void k_warptile_pp(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n)
{
    const unsigned lane = threadIdx.x & 31;
    const size_t W = (size_t)gridDim.x * (blockDim.x >> 5);
    const size_t groups = (n + 32 * U - 1) / (32 * U);   // one group = U rounds
    const size_t gpt = R / U;                              // groups per tile
    const size_t w0 = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // this warp's group sequence: tile t = w0 + k*W, groups t*gpt .. t*gpt+gpt-1
    auto group_at = [&](size_t i) { return (w0 + (i / gpt) * W) * gpt + (i % gpt); };
    const size_t tiles = (groups + gpt - 1) / gpt;
    const size_t my_tiles = w0 < tiles ? (tiles - w0 + W - 1) / W : 0;
    const size_t my_groups = my_tiles * gpt;
    uint4 a[U], b[U];
    auto load = [&](uint4* v, size_t i) {
        const size_t c0 = group_at(i) * 32 * U + lane;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (c0 + (size_t)u * 32 < n) v[u] = s[c0 + (size_t)u * 32];
    };
    auto store = [&](const uint4* v, size_t i) {
        const size_t c0 = group_at(i) * 32 * U + lane;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (c0 + (size_t)u * 32 < n) d[c0 + (size_t)u * 32] = v[u];
    };
    if (my_groups == 0) return;
    load(a, 0);
    size_t i = 0;
    for (; i + 1 < my_groups; i += 2) {
        load(b, i + 1);
        store(a, i);
        if (i + 2 < my_groups) load(a, i + 2);
        store(b, i + 1);
    }
    if (i < my_groups) store(a, i);
}

template <typename F>
float time_ms(F launch, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char** argv)
{
    const size_t bytes = 1ull << 30, n = bytes / 16;
    uint4 *s, *d;
    cudaMalloc(&s, bytes);
    cudaMalloc(&d, bytes);
    cudaMemset(s, 1, bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](const char* name, float ms) { printf("%-44s %7.1f GB/s (read+write)\n", name, 2.0 * bytes / ms / 1e6); };
    if (argc >= 3) {  // sustained mode: copy_skeleton_bench <B|C|M> <reps>  (sample power / clocks from outside)
        const int n_reps = atoi(argv[2]);
        const char which = argv[1][0];
        float ms = 0;
        if (which == 'B') ms = time_ms([&] { k_blockchunk<4><<<(unsigned)((n + 1023) / 1024), 256>>>(s, d, n); }, n_reps);
        else if (which == 'C') ms = time_ms([&] { k_warptile<4, 16><<<sms * 3, 256>>>(s, d, n); }, n_reps);
        else ms = time_ms([&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); }, n_reps);
        printf("sustained %c x %d: %.1f GB/s\n", which, n_reps, 2.0 * bytes / ms / 1e6);
        return 0;
    }
    const int reps = 30;
    report("cudaMemcpy D2D", time_ms([&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); }, reps));
    for (int bps : {2, 4, 8}) {
        char nm[96];
        snprintf(nm, sizeof nm, "A gridstride U4, 256 thr, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_gridstride<4><<<sms * bps, 256>>>(s, d, n); }, reps));
        snprintf(nm, sizeof nm, "A gridstride U8, 256 thr, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_gridstride<8><<<sms * bps, 256>>>(s, d, n); }, reps));
        snprintf(nm, sizeof nm, "A gridstride U2, 256 thr, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_gridstride<2><<<sms * bps, 256>>>(s, d, n); }, reps));
    }
    report("B blockchunk U4, 256 thr", time_ms([&] { k_blockchunk<4><<<(unsigned)((n + 1023) / 1024), 256>>>(s, d, n); }, reps));
    report("B blockchunk U8, 256 thr", time_ms([&] { k_blockchunk<8><<<(unsigned)((n + 2047) / 2048), 256>>>(s, d, n); }, reps));
    report("B blockchunk U4, 128 thr", time_ms([&] { k_blockchunk<4><<<(unsigned)((n + 511) / 512), 128>>>(s, d, n); }, reps));
    report("B blockchunk U16, 256 thr", time_ms([&] { k_blockchunk<16><<<(unsigned)((n + 4095) / 4096), 256>>>(s, d, n); }, reps));
    for (int bps : {2, 3, 4}) {
        char nm[96];
        snprintf(nm, sizeof nm, "C warptile U4 R16, 256 thr, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_warptile<4, 16><<<sms * bps, 256>>>(s, d, n); }, reps));
        snprintf(nm, sizeof nm, "C warptile U4 R4, 256 thr, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_warptile<4, 4><<<sms * bps, 256>>>(s, d, n); }, reps));
        snprintf(nm, sizeof nm, "C warptile U8 R16, 256 thr, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_warptile<8, 16><<<sms * bps, 256>>>(s, d, n); }, reps));
    }
    for (int bps : {2, 3, 4}) {
        char nm[96];
        snprintf(nm, sizeof nm, "D ctatile U4 R16, 256 thr, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_ctatile<4, 16><<<sms * bps, 256>>>(s, d, n); }, reps));
        snprintf(nm, sizeof nm, "D ctatile U2 R16, 256 thr, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_ctatile<2, 16><<<sms * bps, 256>>>(s, d, n); }, reps));
    }
    for (int bps : {2, 3, 4}) {
        char nm[96];
        snprintf(nm, sizeof nm, "E warptile ping-pong U4 R16, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_warptile_pp<4, 16><<<sms * bps, 256>>>(s, d, n); }, reps));
        snprintf(nm, sizeof nm, "E warptile ping-pong U2 R16, %d blocks/SM", bps);
        report(nm, time_ms([&] { k_warptile_pp<2, 16><<<sms * bps, 256>>>(s, d, n); }, reps));
    }
    {   // is it warps per SM or CTAs per SM?  pattern C, U4 R16, (blocks/SM, threads/block)
        const int cfg[][2] = {{1, 512}, {2, 256}, {4, 128}, {1, 768}, {2, 384}, {3, 256}, {6, 128},
                              {1, 1024}, {2, 512}, {4, 256}, {8, 128}, {1, 256}, {1, 384}, {5, 128}, {2, 320}, {2, 192}};
        for (auto& c : cfg) {
            char nm[96];
            snprintf(nm, sizeof nm, "C U4 R16  %d blocks/SM x %4d thr (%2d warps/SM)", c[0], c[1], c[0] * c[1] / 32);
            report(nm, time_ms([&] { k_warptile<4, 16><<<sms * c[0], c[1]>>>(s, d, n); }, reps));
        }
    }
    // in place (read and write the same addresses), like a contiguous Cycle
    report("A gridstride U4 in place, 4 blocks/SM", time_ms([&] { k_gridstride<4><<<sms * 4, 256>>>(s, s, n); }, reps));
    report("C warptile U4 R16 in place, 4 blocks/SM", time_ms([&] { k_warptile<4, 16><<<sms * 4, 256>>>(s, s, n); }, reps));
    report("B blockchunk U4 in place", time_ms([&] { k_blockchunk<4><<<(unsigned)((n + 1023) / 1024), 256>>>(s, s, n); }, reps));
    return 0;
}
