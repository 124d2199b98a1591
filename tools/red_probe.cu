// tools/red_probe.cu -- throughput of 128-bit vector reductions (red.global.add.v4.f32)
// against plain 128-bit loads / stores at the same addresses: each warp touches 512
// contiguous bytes per access, cells drawn pseudo-randomly from a gradient-sized buffer
// (or from a small L2-resident one).  nvcc -arch=sm_100a -O3 -o red_probe red_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(float *buf, unsigned long long cells, int iters, float *sink)
{
    const int lane = threadIdx.x & 31;
    unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long x = w * 0x9E3779B97F4A7C15ull + 12345;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < iters; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        float *p = buf + ((x >> 20) % cells) * 128 + lane * 4;
        if (OP == 0) asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
        if (OP == 1) asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
        if (OP == 2) {
            float4 v;
            asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (OP == 3) {
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 1), "f"(1.0f) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 2), "f"(1.0f) : "memory");
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 3), "f"(1.0f) : "memory");
        }
    }
    if (OP == 2 && acc.x == 123.456f) *sink = acc.y + acc.z + acc.w;
}

template <int OP>
void run(const char *name, float *buf, unsigned long long cells, float *sink)
{
    const int iters = 64, threads = 256, blocks = 148 * 64;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<OP><<<blocks, threads>>>(buf, cells, iters, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) probe<OP><<<blocks, threads>>>(buf, cells, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double bytes = 5.0 * blocks * (threads / 32) * iters * 512.0;
    printf("%-22s cells=%9llu (%7.1f MB)  %8.1f GB/s  (%.3f ms per launch, %.0f MB per launch)\n", name, cells,
           cells * 512.0 / 1e6, bytes / (ms * 1e-3) / 1e9, ms / 5, bytes / 5 / 1e6);
}

int main()
{
    float *buf, *sink;
    const unsigned long long big = 182ull * 1000 * 1000 / 512, small = 32ull * 1000 * 1000 / 512;
    cudaMalloc(&buf, big * 512);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, big * 512);
    for (unsigned long long cells : {big, small}) {
        run<0>("red.add.v4.f32", buf, cells, sink);
        run<3>("4 x red.add.f32", buf, cells, sink);
        run<1>("st.v4.f32", buf, cells, sink);
        run<2>("ld.v4.f32", buf, cells, sink);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
