// Micro-benchmark: how fast can ONE CTA per SM stream 32 KB tiles made of `rows` pieces of `rb` bytes (rows * rb = 32 KB),
// the pieces `stride` bytes apart, with cp.async (16 B per thread-op, 512 threads, 3 tiles in flight)?  No processing.
// Answers whether the 128-byte-row operand tiles of the tensor-core superpixel pooling are what bounds it.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) stream_kernel(const char* __restrict__ src, size_t cta_bytes, int rows, int rb, size_t stride,
                                                        int ntiles, int tile_step, unsigned* sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const char* base = src + (size_t)blockIdx.x * cta_bytes;
    const int cpr = rb / 16;                       // chunks per row
    auto fetch = [&](int tile, int slot) {
        for (int q = threadIdx.x; q < 2048; q += 512) {
            const int row = q / cpr, ch = q - row * cpr;
            const char* g = base + (size_t)row * stride + (size_t)tile * tile_step + ch * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sbase + slot * 32768 + q * 16), "l"(g) : "memory");
        }
    };
    for (int j = 0; j < 3; ++j) { if (j < ntiles) fetch(j, j); asm volatile("cp.async.commit_group;" ::: "memory"); }
    unsigned acc = 0;
    for (int it = 0; it < ntiles; ++it) {
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncthreads();
        acc += smem[(it % 4) * 32768 + threadIdx.x * 4];
        __syncthreads();
        if (it + 3 < ntiles) fetch(it + 3, (it + 3) % 4);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (acc == 0xdeadbeef) *sink = acc;
}

int main() {
    const int ctas = 128, ntiles = 32;
    const size_t cta_bytes = 8ull << 20;           // 8 MB apart (only part of it is read): 1 GB in all
    char* d; unsigned* sink;
    cudaMalloc(&d, (size_t)ctas * cta_bytes + (64 << 20)); cudaMalloc(&sink, 4);
    cudaMemset(d, 1, (size_t)ctas * cta_bytes);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
    struct { int rows, rb; size_t stride; int step; const char* what; } cfg[] = {
        {256, 128, 32768, 128, "256 rows x 128 B, rows 32 KB apart, next tile +128 B (the pooling's map tile)"},
        {64, 512, 32768, 512, "64 rows x 512 B, rows 32 KB apart"},
        {32, 1024, 32768, 1024, "32 rows x 1 KB, rows 32 KB apart"},
        {8, 4096, 32768, 0, "8 rows x 4 KB (whole channel rows), next tile 8 rows further"},
        {1, 32768, 32768, 32768, "contiguous 32 KB"},
        {256, 128, 4096, 128, "256 rows x 128 B, rows 4 KB apart (T = 1)"},
    };
    for (auto& c : cfg) {
        int step = c.step ? c.step : 8 * 32768;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            stream_kernel<<<ctas, 512, 4 * 32768>>>(d, cta_bytes, c.rows, c.rb, c.stride, ntiles, step, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-80s %7.1f us  %5.2f TB/s  (%s)\n", c.what, best * 1e3, ctas * ntiles * 32768.0 / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
