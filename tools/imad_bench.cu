// tools/imad_bench.cu -- integer-pipe microbenchmark for the IMAD side of the roofline:
// thread-instructions per clock per SM for IMAD.WIDE.U32, IMAD.HI.U32, LEA.HI and the exact
// step sequence of the kernel (IMAD.WIDE + LEA.HI), measured with clock64 on a full grid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imad_bench tools/imad_bench.cu && tools/imad_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;

template <int MODE>
__global__ void __launch_bounds__(1024, 2) bench(unsigned* out, unsigned long long* cycles, unsigned seed, unsigned two)
{
    unsigned s[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c)
        s[c] = seed + threadIdx.x * 977u + c * 131u + blockIdx.x;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if (MODE == 0) {  // IMAD.WIDE.U32 only (low half feeds the next one)
                unsigned long long p = (unsigned long long)s[c] * 33614u;
                s[c] = (unsigned)p ^ (unsigned)(p >> 32);  // LOP3 keeps both halves live
            } else if (MODE == 1) {  // the kernel's step: IMAD.WIDE + LEA.HI
                unsigned long long p = (unsigned long long)s[c] * 33614u;
                s[c] = (unsigned)(p >> 32) + ((unsigned)p >> 1);
            } else if (MODE == 2) {  // IMAD.HI.U32 with a 32-bit addend
                unsigned r;
                asm volatile("mad.hi.u32 %0, %1, %2, %1;" : "=r"(r) : "r"(s[c]), "r"(two));
                s[c] = r;
            } else {  // LEA.HI only
                s[c] = s[c] + (s[c] >> 31) + 1u;
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c)
        acc ^= s[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0)
        cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char* name, int per_iter_instr)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 2, threads = 1024;
    unsigned* out;
    unsigned long long* cyc;
    cudaMalloc(&out, sizeof(unsigned) * blocks * threads);
    cudaMalloc(&cyc, sizeof(unsigned long long) * blocks);
    bench<MODE><<<blocks, threads>>>(out, cyc, 1, 2);
    bench<MODE><<<blocks, threads>>>(out, cyc, 2, 2);
    cudaDeviceSynchronize();
    unsigned long long* h = new unsigned long long[blocks];
    cudaMemcpy(h, cyc, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i)
        avg += (double)h[i];
    avg /= blocks;
    // two resident blocks per SM share it for ~avg cycles
    const double thread_instr_per_sm = 2.0 * threads * (double)kIters * kChains * per_iter_instr;
    printf("%-28s %8.1f thread-instr/clk/SM   (%d instr per chain step, %.0f cycles)\n", name,
           thread_instr_per_sm / avg, per_iter_instr, avg);
    cudaFree(out);
    cudaFree(cyc);
    delete[] h;
}

int main()
{
    run<0>("IMAD.WIDE.U32 (+LOP3)", 2);
    run<1>("step: IMAD.WIDE + LEA.HI", 2);
    run<2>("IMAD.HI.U32", 1);
    run<3>("LEA.HI-class ALU (x2)", 2);
    return 0;
}
