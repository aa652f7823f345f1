// Host->device copy rate from default pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), alone and with a
// device->host copy running beside it.  nvcc -O2 -o tools/wc_probe tools/wc_probe.cu ; prints GB/s.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
static double run(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s, void *dst2, const void *src2, cudaStream_t s2, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaMemcpyAsync(dst, src, n, k, s);
    cudaDeviceSynchronize();
    cudaEventRecord(a, s);
    for (int i = 0; i < reps; ++i) {
        cudaMemcpyAsync(dst, src, n, k, s);
        if (dst2) cudaMemcpyAsync(dst2, src2, n, cudaMemcpyDeviceToHost, s2);
    }
    cudaEventRecord(b, s);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return (double)n * reps / (ms * 1e-3) / 1e9;
}
int main() {
    const size_t n = 256u << 20;
    void *hd, *hw, *ho, *d0, *d1;
    cudaHostAlloc(&hd, n, cudaHostAllocDefault);
    cudaHostAlloc(&hw, n, cudaHostAllocWriteCombined);
    cudaHostAlloc(&ho, n, cudaHostAllocDefault);
    cudaMalloc(&d0, n); cudaMalloc(&d1, n);
    memset(hd, 1, n); memset(hw, 1, n);
    cudaStream_t s, s2;
    cudaStreamCreate(&s); cudaStreamCreate(&s2);
    printf("H2D default pinned        : %.1f GB/s\n", run(d0, hd, n, cudaMemcpyHostToDevice, s, nullptr, nullptr, s2, 8));
    printf("H2D write-combined pinned : %.1f GB/s\n", run(d0, hw, n, cudaMemcpyHostToDevice, s, nullptr, nullptr, s2, 8));
    printf("H2D default + D2H beside  : %.1f GB/s\n", run(d0, hd, n, cudaMemcpyHostToDevice, s, ho, d1, s2, 8));
    printf("H2D WC + D2H beside       : %.1f GB/s\n", run(d0, hw, n, cudaMemcpyHostToDevice, s, ho, d1, s2, 8));
    printf("D2H alone                 : %.1f GB/s\n", run(ho, d1, n, cudaMemcpyDeviceToHost, s, nullptr, nullptr, s2, 8));
    return 0;
}
