// microbenchmark: what does an L2 sector miss cost in DRAM traffic, and are the other sectors of the line installed?
// run under: ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./l2gran [limit]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int MODE>
__device__ __forceinline__ unsigned long long ld(const unsigned long long *p) {
    unsigned long long v;
    if (MODE == 0) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.L2::128B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
// every thread reads 8 bytes of sector `sector` of line (hash of its index): random lines, one sector each
template <int MODE>
__global__ void one_sector(const unsigned long long *a, size_t n_lines, int sector, unsigned long long *out) {
    unsigned long long acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_lines; i += (size_t)gridDim.x * blockDim.x) {
        size_t line = (i * 2654435761ull) % n_lines; // scatter the lines over the warp
        acc += ld<MODE>(a + line * 16 + sector * 4);
    }
    if (acc == 0x1234567) *out = acc;
}
// each thread reads sector 0 of a line and, AFTER that value has arrived, sector 1 of the same line (L1 bypassed):
// DRAM bytes per line = 128 if the first miss installed the whole line, 256 (or 128 with ::64B) if it did not
template <int MODE>
__global__ void two_sectors(const unsigned long long *a, size_t n_lines, unsigned long long *out) {
    unsigned long long acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_lines; i += (size_t)gridDim.x * blockDim.x) {
        size_t line = (i * 2654435761ull) % n_lines;
        unsigned long long v0, v1;
        if (MODE == 0) {
            asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v0) : "l"(a + line * 16));
            asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v1) : "l"(a + line * 16 + 4 + (v0 >> 63)));
        } else {
            asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v0) : "l"(a + line * 16));
            asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v1) : "l"(a + line * 16 + 8 + (v0 >> 63)));
        }
        acc += v0 + v1;
    }
    if (acc == 0x1234567) *out = acc;
}
int main(int argc, char **argv) {
    if (argc > 1) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1]));
    size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity limit: %zu\n", lim);
    const size_t bytes = 48ull << 20, n_lines = bytes / 128; // 48 MB: fits in L2
    unsigned long long *a, *out, *flush;
    cudaMalloc(&a, bytes); cudaMalloc(&out, 8); cudaMalloc(&flush, 512ull << 20);
    cudaMemset(a, 1, bytes);
    for (int mode = 0; mode < 4; mode++) {
        cudaMemset(flush, 0, 512ull << 20); // evict
        cudaDeviceSynchronize();
        // pass 1: sector 0 of every line; pass 2: sector 1; pass 3: sector 0 again (must hit)
        for (int pass = 0; pass < 3; pass++) {
            const int sector = pass == 2 ? 0 : pass;
            if (mode == 0) one_sector<0><<<592, 256>>>(a, n_lines, sector, out);
            if (mode == 1) one_sector<1><<<592, 256>>>(a, n_lines, sector, out);
            if (mode == 2) one_sector<2><<<592, 256>>>(a, n_lines, sector, out);
            if (mode == 3) one_sector<3><<<592, 256>>>(a, n_lines, sector, out);
            cudaDeviceSynchronize();
        }
    }
    cudaMemset(flush, 0, 512ull << 20);
    cudaDeviceSynchronize();
    two_sectors<0><<<592, 256>>>(a, n_lines, out);
    cudaDeviceSynchronize();
    two_sectors<1><<<592, 256>>>(a, n_lines, out);  // ::64B, second read is in the OTHER 64-byte half
    cudaDeviceSynchronize();
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
