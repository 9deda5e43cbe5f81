// store_probe.cu -- does K2's output pattern (per cell: 9616 B ids row + 9616 B values row + 1202 B
// mask row, one CTA per cell, persistent) cap the achievable write bandwidth?  No reads, no compute.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int T, int UNROLL>
__global__ void k_rows(long long *ids, long long *vals, uint8_t *mask, int n_cells, int L, int sync_each) {
    for (int j = blockIdx.x; j < n_cells; j += gridDim.x) {
        longlong2 *pi = reinterpret_cast<longlong2 *>(ids + (size_t)j * L);
        longlong2 *pv = reinterpret_cast<longlong2 *>(vals + (size_t)j * L);
        const int np = L / 2;
#pragma unroll UNROLL
        for (int p = threadIdx.x; p < np; p += T) pi[p] = make_longlong2(p, j);
#pragma unroll UNROLL
        for (int p = threadIdx.x; p < np; p += T) pv[p] = make_longlong2(j, p);
        uint4 *pm = reinterpret_cast<uint4 *>(mask + (((size_t)j * L + 15) & ~(size_t)15));
        for (int c = threadIdx.x; c < L / 16 - 1; c += T) pm[c] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
        if (sync_each) __syncthreads();
    }
}
// same bytes, but the whole grid sweeps each output array as one contiguous stream
__global__ void k_flat(uint4 *o, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) o[i] = make_uint4(1, 2, 3, (unsigned)i);
}
int main() {
    const int C = 16384, L = 1202;
    long long *ids, *vals; uint8_t *mask;
    cudaMalloc(&ids, (size_t)C * L * 8 + 64); cudaMalloc(&vals, (size_t)C * L * 8 + 64); cudaMalloc(&mask, (size_t)C * L + 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)C * L * 17;
    auto run = [&](const char *name, auto launch) {
        float best = 1e9f;
        for (int r = 0; r < 8; r++) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-44s %8.1f GB/s  (%.1f us)\n", name, bytes / best / 1e6, best * 1e3);
    };
    for (int per_sm : {4, 8}) {
        int g = 148 * per_sm;
        printf("grid %d\n", g);
        run("rows T=256 unroll1 sync", [&] { k_rows<256, 1><<<g, 256>>>(ids, vals, mask, C, L, 1); });
        run("rows T=256 unroll1 nosync", [&] { k_rows<256, 1><<<g, 256>>>(ids, vals, mask, C, L, 0); });
        run("rows T=256 unroll4 nosync", [&] { k_rows<256, 4><<<g, 256>>>(ids, vals, mask, C, L, 0); });
        run("rows T=128 unroll4 nosync", [&] { k_rows<128, 4><<<g * 2, 128>>>(ids, vals, mask, C, L, 0); });
    }
    run("rows T=256 one CTA per cell (grid=C)", [&] { k_rows<256, 4><<<C, 256>>>(ids, vals, mask, C, L, 0); });
    run("flat sweep of the same bytes (ids only x2.1)", [&] { k_flat<<<148 * 8, 256>>>((uint4 *)ids, (size_t)(bytes / 16) > (size_t)C * L / 2 ? (size_t)C * L / 2 : 0); });
    {
        // flat sweep sized to the same bytes over ids+vals contiguous? use ids buffer repeatedly: just report GB/s for its own bytes
        const size_t n = (size_t)C * L / 2;   // uint4 count of the ids array
        float best = 1e9f;
        for (int r = 0; r < 8; r++) { cudaEventRecord(e0); k_flat<<<148 * 8, 256>>>((uint4 *)ids, n); k_flat<<<148 * 8, 256>>>((uint4 *)vals, n); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        printf("%-44s %8.1f GB/s  (%.1f us)\n", "flat ids then flat vals (2 launches)", (double)n * 32 / best / 1e6, best * 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
