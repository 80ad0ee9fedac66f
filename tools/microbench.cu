// microbench.cu — B200 memory-system probes that size the join / group-by designs (not part of the library).
//   stream copy, random 16 B gathers (table size sweep), random fp64 atomics (table size sweep).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}

__global__ void k_copy(const int4 *__restrict__ a, int4 *__restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <int U>
__global__ void k_gather16(const int4 *__restrict__ table, uint64_t nslots, size_t nprobe, unsigned long long *sink) {
    unsigned long long acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nprobe; i += stride * U) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            size_t r = i + u * stride;
            uint64_t s = __umul64hi(mix(r), nslots);
            v[u] = r < nprobe ? __ldg(&table[s]) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc += (unsigned)v[u].x + (unsigned)v[u].w;
    }
    if (acc == 0x1234567) *sink = acc;
}

// dependent 2-level gather: slot -> payload row (like slot{key,rowid} + payload[rowid])
template <int U>
__global__ void k_gather_dep(const int4 *__restrict__ table, uint64_t nslots, const int2 *__restrict__ pay, size_t nprobe, unsigned long long *sink) {
    unsigned long long acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nprobe; i += stride * U) {
        int4 v[U]; int2 p[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            size_t r = i + u * stride;
            uint64_t s = __umul64hi(mix(r), nslots);
            v[u] = r < nprobe ? __ldg(&table[s]) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; u++) p[u] = __ldg(&pay[(unsigned)v[u].z]);
#pragma unroll
        for (int u = 0; u < U; u++) acc += (unsigned)p[u].x;
    }
    if (acc == 0x1234567) *sink = acc;
}

__global__ void k_atomic_f64(double *table, uint64_t n, size_t nops) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nops; i += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&table[__umul64hi(mix(i), n)], 1.0);
}

// gather + streaming in/out like a probe: read 16 B row, gather 16 B, write 32 B
__global__ void k_probe_like(const int4 *__restrict__ in, const int4 *__restrict__ table, uint64_t nslots, int4 *__restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int4 r = in[i];
        uint64_t s = __umul64hi(mix(((uint64_t)(unsigned)r.y << 32) | (unsigned)r.x), nslots);
        int4 t = __ldg(&table[s]);
        out[2 * i] = r;
        out[2 * i + 1] = t;
    }
}

__global__ void k_fill(int4 *p, size_t n, unsigned mod) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t m = mix(i * 7 + 1);
        p[i] = make_int4((int)m, (int)(m >> 32), (int)(mix(i) % mod), 1);
    }
}

__global__ void k_cursor(unsigned long long *cur, size_t nops, int per_warp) {
    // one lane per warp (or per block) bumps a single global cursor, like a tile-level output reservation
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < nops * 32; i += stride)
        if ((threadIdx.x & 31) == 0 && (per_warp || threadIdx.x == 0)) atomicAdd(cur, 7ULL);
}

static float timeit(void (*launch)(void *), void *arg, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(arg); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; i++) {
        cudaEventRecord(a); launch(arg); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

struct A { const int4 *table; uint64_t nslots; size_t nprobe; unsigned long long *sink; int2 *pay; int grid; int4 *in, *out; double *dt; };
static void l_g1(void *p) { A *a = (A *)p; k_gather16<1><<<a->grid, 256>>>(a->table, a->nslots, a->nprobe, a->sink); }
static void l_g4(void *p) { A *a = (A *)p; k_gather16<4><<<a->grid, 256>>>(a->table, a->nslots, a->nprobe, a->sink); }
static void l_g8(void *p) { A *a = (A *)p; k_gather16<8><<<a->grid, 256>>>(a->table, a->nslots, a->nprobe, a->sink); }
static void l_d4(void *p) { A *a = (A *)p; k_gather_dep<4><<<a->grid, 256>>>(a->table, a->nslots, a->pay, a->nprobe, a->sink); }
static void l_pl(void *p) { A *a = (A *)p; k_probe_like<<<a->grid, 256>>>(a->in, a->table, a->nslots, a->out, a->nprobe); }
static void l_cp(void *p) { A *a = (A *)p; k_copy<<<a->grid, 256>>>(a->in, a->out, a->nprobe); }
static void l_at(void *p) { A *a = (A *)p; k_atomic_f64<<<a->grid, 256>>>(a->dt, a->nslots, a->nprobe); }

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    A a; a.grid = sms * 8;
    size_t max_slots = (size_t)1 << 28;  // 4 GiB of 16 B slots
    int4 *table; CK(cudaMalloc(&table, max_slots * 16));
    int2 *pay; CK(cudaMalloc(&pay, (size_t)100000000 * 8));
    CK(cudaMalloc(&a.sink, 8));
    k_fill<<<sms * 8, 256>>>(table, max_slots, 100000000u); CK(cudaDeviceSynchronize());
    CK(cudaMemset(pay, 1, (size_t)100000000 * 8));
    a.table = table; a.pay = pay;
    size_t nprobe = (size_t)1 << 28;  // 268M probes
    a.nprobe = nprobe;
    size_t nrows = (size_t)1 << 28;
    CK(cudaMalloc(&a.in, nrows * 16)); CK(cudaMalloc(&a.out, nrows * 32));
    k_fill<<<sms * 8, 256>>>(a.in, nrows, 100000000u); CK(cudaDeviceSynchronize());
    { A c = a; c.nprobe = nrows; float ms = timeit(l_cp, &c, 5); printf("copy 16B x %zu: %.3f ms  %.1f GB/s (r+w)\n", nrows, ms, nrows * 32.0 / ms / 1e6); }
    for (int lg = 20; lg <= 28; lg += 2) {
        a.nslots = (uint64_t)1 << lg;
        float m1 = timeit(l_g1, &a, 3), m4 = timeit(l_g4, &a, 3), m8 = timeit(l_g8, &a, 3), md = timeit(l_d4, &a, 3);
        printf("gather16 table %6.0f MiB: U1 %.2f  U4 %.2f  U8 %.2f Gprobe/s | dep(slot->8B pay in 800MB) U4 %.2f Gprobe/s\n",
               (double)(a.nslots * 16) / 1048576.0, nprobe / m1 / 1e6, nprobe / m4 / 1e6, nprobe / m8 / 1e6, nprobe / md / 1e6);
    }
    for (int lg = 22; lg <= 28; lg += 2) {
        a.nslots = (uint64_t)1 << lg; a.nprobe = nrows;
        float ms = timeit(l_pl, &a, 3);
        printf("probe-like (16B in, 16B gather, 32B out) table %6.0f MiB: %.3f ms  %.2f Grows/s  alg(64B/row) %.1f GB/s\n",
               (double)(a.nslots * 16) / 1048576.0, ms, nrows / ms / 1e6, nrows * 64.0 / ms / 1e6);
    }
    a.dt = (double *)table;
    for (int lg = 10; lg <= 28; lg += 3) {
        a.nslots = (uint64_t)1 << lg; a.nprobe = (size_t)1 << 27;
        float ms = timeit(l_at, &a, 3);
        printf("atomicAdd(double) table %9.3f MiB: %.2f Gops/s\n", (double)(a.nslots * 8) / 1048576.0, a.nprobe / ms / 1e6);
    }
    {
        unsigned long long *cur; CK(cudaMalloc(&cur, 8)); CK(cudaMemset(cur, 0, 8));
        for (int pw = 1; pw >= 0; pw--) {
            size_t nops = (size_t)1 << 22;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_cursor<<<sms * 8, 256>>>(cur, nops, pw); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); k_cursor<<<sms * 8, 256>>>(cur, nops, pw); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double n = pw ? (double)nops : (double)nops / 8;
            printf("single-address atomicAdd(u64), one lane per %s: %.1f Mops/s (%.3f ms for %.0f ops)\n", pw ? "warp" : "block", n / ms / 1e3, ms, n);
        }
    }
    return 0;
}
