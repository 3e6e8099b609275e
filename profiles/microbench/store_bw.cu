// Pure-store microbenchmark: what write bandwidth can a kernel reach on this GPU with 128-bit vs 256-bit stores,
// into a DRAM-sized buffer (1 GiB) and into an L2-resident one (58 MiB)?  Context for the step kernel's store phase.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_bw store_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int WIDE>
__global__ void fill_kernel(uint8_t* dst, size_t bytes) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * (WIDE ? 32 : 16);
    for (size_t off = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * (WIDE ? 32 : 16); off < bytes; off += stride) {
        if (WIDE) {
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(dst + off), "r"(0x01010101u) : "memory");
        } else {
            asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" :: "l"(dst + off), "r"(0x01010101u) : "memory");
        }
    }
}

template <int WIDE>
float run(uint8_t* buf, size_t bytes, int blocks, int threads, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) fill_kernel<WIDE><<<blocks, threads>>>(buf, bytes);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) fill_kernel<WIDE><<<blocks, threads>>>(buf, bytes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return (float)((double)bytes * reps / (ms * 1e-3) / 1e9);
}

int main() {
    uint8_t* buf;
    const size_t big = 1ull << 30, small = 58ull << 20;
    cudaMalloc(&buf, big);
    printf("{");
    bool first = true;
    for (int threads : {128, 256, 512}) for (int bps : {4, 8, 16}) {
        const int blocks = 148 * bps;
        printf("%s\n \"t%d_b%d\": {\"dram_128\": %.0f, \"dram_256\": %.0f, \"l2_128\": %.0f, \"l2_256\": %.0f}", first ? "" : ",", threads, bps,
               run<0>(buf, big, blocks, threads, 10), run<1>(buf, big, blocks, threads, 10),
               run<0>(buf, small, blocks, threads, 100), run<1>(buf, small, blocks, threads, 100));
        first = false;
    }
    printf("\n}\n");
    return 0;
}
