// membench.cu — isolates the streaming patterns of the fast join kernels (which loads/stores reach HBM speed?).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }

template <int MODE>  // 0 plain, 1 .cs both, 2 .cs loads only, 3 .cs stores only
__global__ void k_copy(const int4 *__restrict__ a, int4 *__restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int4 v = (MODE == 1 || MODE == 2) ? __ldcs(a + i) : a[i];
        if (MODE == 1 || MODE == 3) __stcs(b + i, v); else b[i] = v;
    }
}
// AoS rows in -> 6 SoA columns out (8,4,4,8,4,4 B), RPT rows per thread in flight
template <int MODE, int RPT>
__global__ void k_soa(const int4 *__restrict__ in, long long *c0, int *c1, int *c2, long long *c3, int *c4, int *c5, size_t n, size_t shift) {
    size_t tile = (size_t)blockDim.x * RPT;
    for (size_t t0 = blockIdx.x * tile; t0 < n; t0 += (size_t)gridDim.x * tile) {
        int4 v[RPT];
#pragma unroll
        for (int k = 0; k < RPT; k++) { size_t r = t0 + k * blockDim.x + threadIdx.x; v[k] = r < n ? (MODE ? __ldcs(in + r) : in[r]) : make_int4(0, 0, 0, 0); }
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            size_t r = t0 + k * blockDim.x + threadIdx.x; if (r >= n) continue;
            size_t p = r + shift;
            long long key = ((long long)v[k].y << 32) | (unsigned)v[k].x;
            if (MODE) { __stcs(c0 + p, key); __stcs(c1 + p, v[k].z); __stcs(c2 + p, v[k].w); __stcs(c3 + p, key); __stcs(c4 + p, v[k].z ^ 1); __stcs(c5 + p, v[k].w ^ 1); }
            else { c0[p] = key; c1[p] = v[k].z; c2[p] = v[k].w; c3[p] = key; c4[p] = v[k].z ^ 1; c5[p] = v[k].w ^ 1; }
        }
    }
}
// 3 SoA columns in -> AoS 16 B rows out, in order (no scatter)
template <int MODE, int RPT>
__global__ void k_aos(const long long *c0, const int *c1, const int *c2, int4 *out, size_t n) {
    size_t tile = (size_t)blockDim.x * RPT;
    for (size_t t0 = blockIdx.x * tile; t0 < n; t0 += (size_t)gridDim.x * tile) {
        long long a[RPT]; int b[RPT], c[RPT];
#pragma unroll
        for (int k = 0; k < RPT; k++) { size_t r = t0 + k * blockDim.x + threadIdx.x; bool ok = r < n; a[k] = ok ? (MODE ? __ldcs(c0 + r) : c0[r]) : 0; b[k] = ok ? (MODE ? __ldcs(c1 + r) : c1[r]) : 0; c[k] = ok ? (MODE ? __ldcs(c2 + r) : c2[r]) : 0; }
#pragma unroll
        for (int k = 0; k < RPT; k++) { size_t r = t0 + k * blockDim.x + threadIdx.x; if (r >= n) continue; int4 v = make_int4((int)a[k], (int)(a[k] >> 32), b[k], c[k]); if (MODE) __stcs(out + r, v); else out[r] = v; }
    }
}
// scatter: rows go to P partitions in runs of RUN rows; each block owns a cursor per partition (like k_fj_scatter's output side)
__global__ void k_runs(const int4 *__restrict__ in, int4 *out, size_t n, int P, int RUN, size_t rows_per_block) {
    size_t b0 = blockIdx.x * rows_per_block, b1 = b0 + rows_per_block < n ? b0 + rows_per_block : n;
    size_t part_rows = n / P;                       // region of partition p: [p*part_rows, ...)
    size_t blk_share = part_rows / gridDim.x;       // this block's slice inside every partition
    size_t done = 0;
    for (size_t t0 = b0; t0 < b1; t0 += (size_t)blockDim.x * 8) {
        for (int k = 0; k < 8; k++) {
            size_t i = (size_t)k * blockDim.x + threadIdx.x, r = t0 + i;
            if (r >= b1) continue;
            // element i of the tile belongs to run i / RUN -> partition (run id % P); position inside this block's slice
            size_t run = i / RUN; int p = (int)(run % P); size_t within = i % RUN;
            size_t tile_idx = (t0 - b0) / ((size_t)blockDim.x * 8);
            size_t runs_per_tile_per_part = ((size_t)blockDim.x * 8 / RUN + P - 1) / P;
            size_t off = (tile_idx * runs_per_tile_per_part + run / P) * RUN + within;
            if (off >= blk_share) continue;
            __stcs(out + (size_t)p * part_rows + (size_t)blockIdx.x * blk_share + off, __ldcs(in + r));
        }
    }
    (void)done;
}
template <typename F> static float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best = 1e30f; for (int i = 0; i < 3; i++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; } return best; }
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    size_t n = (size_t)1 << 28;
    int4 *in, *out; CK(cudaMalloc(&in, n * 16)); CK(cudaMalloc(&out, n * 32 + 4096)); CK(cudaMemset(in, 1, n * 16));
    long long *c0 = (long long *)out, *c3 = c0 + n + 64; int *c1 = (int *)(c3 + n + 64), *c2 = c1 + n + 64, *c4 = c2 + n + 64, *c5 = c4 + n + 64;
    int g = sms * 8;
    printf("copy plain      : %.1f GB/s\n", n * 32.0 / timeit([&] { k_copy<0><<<g, 256>>>(in, out, n); }) / 1e6);
    printf("copy .cs ld+st  : %.1f GB/s\n", n * 32.0 / timeit([&] { k_copy<1><<<g, 256>>>(in, out, n); }) / 1e6);
    printf("copy .cs ld     : %.1f GB/s\n", n * 32.0 / timeit([&] { k_copy<2><<<g, 256>>>(in, out, n); }) / 1e6);
    printf("copy .cs st     : %.1f GB/s\n", n * 32.0 / timeit([&] { k_copy<3><<<g, 256>>>(in, out, n); }) / 1e6);
    for (int sh = 0; sh <= 3; sh += 3) {
        printf("AoS->6 SoA plain RPT4 shift %d: %.1f GB/s\n", sh, n * 48.0 / timeit([&] { k_soa<0, 4><<<sms * 6, 256>>>(in, c0, c1, c2, c3, c4, c5, n, sh); }) / 1e6);
        printf("AoS->6 SoA .cs   RPT4 shift %d: %.1f GB/s\n", sh, n * 48.0 / timeit([&] { k_soa<1, 4><<<sms * 6, 256>>>(in, c0, c1, c2, c3, c4, c5, n, sh); }) / 1e6);
        printf("AoS->6 SoA .cs   RPT8 shift %d: %.1f GB/s\n", sh, n * 48.0 / timeit([&] { k_soa<1, 8><<<sms * 4, 256>>>(in, c0, c1, c2, c3, c4, c5, n, sh); }) / 1e6);
    }
    printf("3 SoA->AoS plain RPT8: %.1f GB/s\n", n * 32.0 / timeit([&] { k_aos<0, 8><<<sms * 4, 256>>>(c0, c1, c2, in, n); }) / 1e6);
    printf("3 SoA->AoS .cs   RPT8: %.1f GB/s\n", n * 32.0 / timeit([&] { k_aos<1, 8><<<sms * 4, 256>>>(c0, c1, c2, in, n); }) / 1e6);
    CK(cudaMemset(in, 1, n * 16));
    for (int P : {1, 8, 96, 256}) for (int RUN : {8, 21, 64}) {
        int grid = sms * 3; size_t rpb = (n + grid - 1) / grid; rpb = (rpb + 2047) / 2048 * 2048;
        float ms = timeit([&] { k_runs<<<grid, 256>>>(in, out, n, P, RUN, rpb); });
        printf("scatter runs P=%3d RUN=%2d: %.3f ms -> %.1f GB/s (r+w, approx)\n", P, RUN, ms, n * 32.0 / ms / 1e6);
    }
    return 0;
}
