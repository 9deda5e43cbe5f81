// bw_probe.cu -- streaming ceilings on this GPU for the access mixes the hot path has:
// pure 128-bit writes, pure reads, copy (50/50) and a 1:3 read:write mix (K2's ratio).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k_write(uint4 *o, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) o[i] = make_uint4(1, 2, 3, (unsigned)i);
}
__global__ void k_read(const uint4 *a, size_t n, unsigned *sink) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (; i < n; i += st) { uint4 v = __ldcs(a + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void k_copy(const uint4 *a, uint4 *o, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) o[i] = __ldcs(a + i);
}
// read 1 x 16B, write 3 x 16B (expanding u16 -> i64 is a 1:4 expansion; K2 is ~ 1:2.5 in bytes)
__global__ void k_mix(const uint4 *a, uint4 *o, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) { uint4 v = __ldcs(a + i); o[3 * i] = v; o[3 * i + 1] = v; o[3 * i + 2] = v; }
}
// same mix but each warp writes 3 contiguous 512-byte runs (coalesced like K2's rows)
__global__ void k_mix_rows(const uint4 *a, uint4 *o, size_t n) {
    size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    for (size_t b = w; b * 32 < n; b += nw) {
        uint4 v = __ldcs(a + b * 32 + lane);
        o[b * 96 + lane] = v; o[b * 96 + 32 + lane] = v; o[b * 96 + 64 + lane] = v;
    }
}
int main() {
    const size_t N = (size_t)1 << 26;  // 1 GiB of uint4
    uint4 *a, *o; unsigned *sink;
    cudaMalloc(&a, N * 16); cudaMalloc(&o, N * 16); cudaMalloc(&sink, 4);
    cudaMemset(a, 1, N * 16); cudaMemset(o, 0, N * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, double bytes, auto launch) {
        float best = 1e9f;
        for (int r = 0; r < 6; r++) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-28s %8.1f GB/s  (%.3f ms)\n", name, bytes / best / 1e6, best);
    };
    for (int grid : {148 * 8, 148 * 16, 148 * 32}) {
        printf("grid %d x 256\n", grid);
        run("write 1 GiB", N * 16.0, [&] { k_write<<<grid, 256>>>(o, N); });
        run("read 1 GiB", N * 16.0, [&] { k_read<<<grid, 256>>>(a, N, sink); });
        run("copy 1+1 GiB", N * 32.0, [&] { k_copy<<<grid, 256>>>(a, o, N); });
        run("mix r1:w3 (interleaved)", (N / 4) * 64.0, [&] { k_mix<<<grid, 256>>>(a, o, N / 4); });
        run("mix r1:w3 (row runs)", (N / 4) * 64.0, [&] { k_mix_rows<<<grid, 256>>>(a, o, N / 4); });
    }
    run("cudaMemsetAsync 1 GiB", N * 16.0, [&] { cudaMemsetAsync(o, 0, N * 16); });
    run("cudaMemcpy D2D 1+1 GiB", N * 32.0, [&] { cudaMemcpyAsync(o, a, N * 16, cudaMemcpyDeviceToDevice); });
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
