// join_fast.cuh — radix-partitioned, L2-resident hash join for a single integer equi-key (the C2 / TPC-H shape).
//
// Why: on B200 a random 16-byte table read costs ~35 G reads/s when the table lives in HBM (each one drags a ~128-byte
// fetch: 174 GB of DRAM traffic per 1 B probe rows, profiles/r01_ncu_summary.md prof_r01h) but ~230 G reads/s when the
// touched slice of the table fits the 126 MB L2 (tools/microbench.cu, profiles/r01_microbench.txt).  So for tables
// beyond L2 both sides are range-partitioned on the TABLE SLOT (slot = mulhi(key_hash, nslots), partition = the same
// hash's high bits scaled to P, i.e. partition p owns the contiguous slot range p), rows are packed into fixed-stride
// 8-byte-word rows {key, payload...}, and build and probe walk partition after partition so that the active 16 MB
// slice of the table stays L2-resident while the rows stream through with evict-first loads/stores:
//     k_fj_hist        keys only          ->  [partition][block] histogram
//     k_fj_scatter     all columns        ->  packed rows in partition order (cp.async double-buffered column loads,
//                                             shared-memory staged tile, 16-byte run writes)
//     k_fj_build_part  packed build rows  ->  table; cooperative: EMPTY-fill a 16 MB partition group, grid barrier,
//                                             CAS-insert into it while it is still dirty in L2
//     k_fj_probe       packed probe rows  ->  output columns (one table read per probe row, warp-ballot compaction,
//                                             one global cursor bump per 2048-row tile, full-line column flush)
// Tables that fit L2 skip the partitioning: k_fj_table_init + k_fj_insert build them and k_fj_probe reads the probe rows
// straight from the input columns.  k_fj_probe_pipe (persistent, cp.async row prefetch, ticketed tiles) and
// k_fj_probe_tma (persistent, TMA-staged ring) are measurement variants kept behind environment switches.
// The table holds whole build rows inline (stride = key + payload words), so a probe is ONE L2 access; duplicate
// build keys, NULLs, composite/double keys, non-equi conditions and outer-build joins take the generic path in
// join.cu (same results, two-pass sizing).
//
// Reference behaviour preserved: AbstractBufferedJoinExec.nextRows:185-264 for INNER / LEFT / RIGHT / SEMI / ANTI
// with unique build keys and no NULLs (row multiset identical; output order is unspecified in both).
#pragma once

#include <cooperative_groups.h>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace fj {

constexpr unsigned long long KEY_EMPTY = 0x8000000000000000ULL;
constexpr int THREADS = 512;           // 16 warps per CTA: two CTAs per SM give 32 resident warps at ~60 registers
constexpr int RPT = 4;                 // rows per thread per tile
constexpr int TILE = THREADS * RPT;    // 2048 rows
constexpr int MAX_WORDS = 4;           // packed row = key word + up to 3 payload words
constexpr int MAX_P = 1024;
constexpr int MAX_DISP = 4096;         // insert gives up (generic path) beyond this displacement
enum { FL_SENTINEL = 0, FL_DUP = 1, FL_DISP = 2, FL_NULLOUT = 3, FL_COUNT = 4 };

struct Layout {
    int32_t nwords, ncols, key_col, key_i32;
    int32_t word[GSQL_MAX_COLS];
    int32_t half[GSQL_MAX_COLS];  // 0 = low 32 bits, 1 = high 32 bits, 2 = whole word
};

// Returns false when the side does not fit the packed-row format.
static bool make_layout(const int32_t *types, int ncols, int key_col, Layout *L) {
    memset(L, 0, sizeof(*L));
    L->ncols = ncols;
    L->key_col = key_col;
    L->key_i32 = types[key_col] == GSQL_T_INT32;
    int next = 1;
    int open_word = -1;  // a word whose high half is still free
    for (int c = 0; c < ncols; c++) {
        if (c == key_col) {
            L->word[c] = 0;
            L->half[c] = L->key_i32 ? 0 : 2;  // an INT32 key is stored widened; its low half is the column value
            continue;
        }
        if (types[c] == GSQL_T_INT32) {
            if (open_word >= 0) {
                L->word[c] = open_word;
                L->half[c] = 1;
                open_word = -1;
            } else {
                L->word[c] = next;
                L->half[c] = 0;
                open_word = next++;
            }
        } else {
            L->word[c] = next++;
            L->half[c] = 2;
        }
    }
    L->nwords = next;
    return next <= MAX_WORDS;
}

// Fibonacci (multiplicative) hashing: the well-mixed HIGH bits of key * phi64 are exactly what the mulhi range
// reductions (slot = mulhi(h, nslots), partition = mulhi(h, P)) consume; one 64-bit multiply instead of fmix64's two.
__device__ __forceinline__ uint64_t key_hash(unsigned long long k) { return (k ^ (k >> 32)) * 0x9E3779B97F4A7C15ULL; }

template <int W>
__device__ __forceinline__ void pack_row(const DColSet &cols, const Layout &L, int64_t r, unsigned long long (&w)[W]) {
#pragma unroll
    for (int i = 0; i < W; i++) w[i] = 0;
#pragma unroll 1
    for (int c = 0; c < L.ncols; c++) {
        const DCol &col = cols.c[c];
        unsigned long long v;
        if (col.type == GSQL_T_INT32) {
            int x = ld_stream_4(reinterpret_cast<const int *>(col.data) + r);
            v = c == L.key_col ? (unsigned long long)(long long)x : (unsigned long long)(unsigned)x;
        } else {
            v = (unsigned long long)ld_stream_8(reinterpret_cast<const long long *>(col.data) + r);
        }
        int wi = L.word[c];
        if (L.half[c] == 1) v <<= 32;
#pragma unroll
        for (int i = 0; i < W; i++)
            if (i == wi) w[i] |= v;
    }
}

// Packs the RPT rows a thread owns in a tile (rows base + k*THREADS).  With a compile-time column count NC every load
// of the tile (NC x RPT coalesced loads) is issued before the first one is consumed; NC = 0 is the generic fallback
// (column loop not unrolled: RPT loads in flight per column).
template <int W, int NC>
__device__ __forceinline__ void pack_tile_nc(const DColSet &cols, const Layout &L, int64_t base, int64_t limit, unsigned long long (&w)[RPT][W]) {
    unsigned long long v[NC > 0 ? NC : 1][RPT];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const DCol &col = cols.c[c];
        const bool is32 = col.type == GSQL_T_INT32;
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            int64_t r = base + k * THREADS;
            v[c][k] = 0;
            if (r < limit) {
                if (is32) v[c][k] = (unsigned long long)(unsigned)ld_stream_4(reinterpret_cast<const int *>(col.data) + r);
                else v[c][k] = (unsigned long long)ld_stream_8(reinterpret_cast<const long long *>(col.data) + r);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < RPT; k++)
#pragma unroll
        for (int i = 0; i < W; i++) w[k][i] = 0;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const bool sext = c == L.key_col && cols.c[c].type == GSQL_T_INT32;
        const int wi = L.word[c];
        const int sh = L.half[c] == 1 ? 32 : 0;
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            unsigned long long vv = sext ? (unsigned long long)(long long)(int)(unsigned)v[c][k] : v[c][k];
            vv <<= sh;
#pragma unroll
            for (int i = 0; i < W; i++)
                if (i == wi) w[k][i] |= vv;
        }
    }
}

template <int W>
__device__ __forceinline__ void pack_tile_generic(const DColSet &cols, const Layout &L, int64_t base, int64_t limit, unsigned long long (&w)[RPT][W]) {
#pragma unroll
    for (int k = 0; k < RPT; k++)
#pragma unroll
        for (int i = 0; i < W; i++) w[k][i] = 0;
#pragma unroll 1
    for (int c = 0; c < L.ncols; c++) {
        const DCol &col = cols.c[c];
        const bool is32 = col.type == GSQL_T_INT32, iskey = c == L.key_col;
        const int wi = L.word[c];
        const int sh = L.half[c] == 1 ? 32 : 0;
        unsigned long long v[RPT];
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            int64_t r = base + k * THREADS;
            v[k] = 0;
            if (r < limit) {
                if (is32) {
                    int x = ld_stream_4(reinterpret_cast<const int *>(col.data) + r);
                    v[k] = iskey ? (unsigned long long)(long long)x : (unsigned long long)(unsigned)x;
                } else {
                    v[k] = (unsigned long long)ld_stream_8(reinterpret_cast<const long long *>(col.data) + r);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            unsigned long long vv = v[k] << sh;
#pragma unroll
            for (int i = 0; i < W; i++)
                if (i == wi) w[k][i] |= vv;
        }
    }
}

template <int W>
__device__ __forceinline__ void pack_tile(const DColSet &cols, const Layout &L, int64_t base, int64_t limit, unsigned long long (&w)[RPT][W]) {
    switch (L.ncols) {  // warp-uniform
    case 1: pack_tile_nc<W, 1>(cols, L, base, limit, w); break;
    case 2: pack_tile_nc<W, 2>(cols, L, base, limit, w); break;
    case 3: pack_tile_nc<W, 3>(cols, L, base, limit, w); break;
    case 4: pack_tile_nc<W, 4>(cols, L, base, limit, w); break;
    default: pack_tile_generic<W>(cols, L, base, limit, w); break;
    }
}

__device__ __forceinline__ unsigned long long load_key(const DCol &col, int64_t r) {
    if (col.type == GSQL_T_INT32) return (unsigned long long)(long long)ld_stream_4(reinterpret_cast<const int *>(col.data) + r);
    return (unsigned long long)ld_stream_8(reinterpret_cast<const long long *>(col.data) + r);
}

// Partition of a key hash.  Only the high 32 bits take part (one IMAD.HI): the result may differ from
// slot / slots_per_partition for a vanishing fraction of keys, which costs those rows an access outside the resident
// slice, never correctness (slots are always addressed globally).
__device__ __forceinline__ unsigned int part_of(uint64_t h, int P) { return __umulhi((unsigned int)(h >> 32), (unsigned int)P); }

struct PartGeom {
    int64_t rows, chunk;  // rows per block (multiple of TILE)
    int32_t P, nblocks;
};

// ---- pass 1: histogram of partition ids, keys only
__global__ void __launch_bounds__(THREADS) k_fj_hist(DCol keycol, PartGeom g, int64_t *__restrict__ hist, int32_t *flags) {
    extern __shared__ unsigned int sh_hist[];
    for (int i = threadIdx.x; i < g.P; i += THREADS) sh_hist[i] = 0;
    __syncthreads();
    int64_t r0 = (int64_t)blockIdx.x * g.chunk;
    int64_t r1 = r0 + g.chunk < g.rows ? r0 + g.chunk : g.rows;
    bool sentinel = false;
    for (int64_t t0 = r0; t0 < r1; t0 += TILE) {
        unsigned long long key[RPT];
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            int64_t r = t0 + k * THREADS + threadIdx.x;
            key[k] = r < r1 ? load_key(keycol, r) : 0;
        }
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            int64_t r = t0 + k * THREADS + threadIdx.x;
            if (r < r1) {
                sentinel |= key[k] == KEY_EMPTY;
                atomicAdd(&sh_hist[part_of(key_hash(key[k]), g.P)], 1u);
            }
        }
    }
    if (sentinel) flags[FL_SENTINEL] = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < g.P; i += THREADS) hist[(int64_t)i * g.nblocks + blockIdx.x] = sh_hist[i];
}

// ---- cp.async (LDGSTS) helpers: per-thread asynchronous global -> shared copies, grouped and waited per tile
__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16_hint(void *smem_dst, const void *gsrc, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Issues the asynchronous loads of one tile: column c lands at buf + (bytes of columns < c) * TILE, element
// (k*THREADS + tid); every thread later reads back exactly the elements it issued (no block barrier needed).
// FULL tiles carry no per-row bounds checks; all per-row addresses are (per-column base) + compile-time offsets.
template <bool FULL>
__device__ __forceinline__ void tile_prefetch(const DColSet &cols, const Layout &L, int64_t t0, int n_tile, unsigned char *buf) {
    const unsigned tid = threadIdx.x;
    unsigned char *sp = buf + tid * 4u;
#pragma unroll 1
    for (int c = 0; c < L.ncols; c++) {
        const DCol &col = cols.c[c];
        if (col.type == GSQL_T_INT32) {
            const int *gp = reinterpret_cast<const int *>(col.data) + t0 + tid;
#pragma unroll
            for (int k = 0; k < RPT; k++)
                if (FULL || (int)(k * THREADS + tid) < n_tile) cp_async_4(sp + k * THREADS * 4, gp + k * THREADS);
            sp += TILE * 4;
        } else {
            const long long *gp = reinterpret_cast<const long long *>(col.data) + t0 + tid;
            unsigned char *sp8 = sp + tid * 4u;
#pragma unroll
            for (int k = 0; k < RPT; k++)
                if (FULL || (int)(k * THREADS + tid) < n_tile) cp_async_8(sp8 + k * THREADS * 8, gp + k * THREADS);
            sp += TILE * 8;
        }
    }
    cp_async_commit();
}

template <int W, bool FULL>
__device__ __forceinline__ void pack_tile_smem(const DColSet &cols, const Layout &L, int n_tile, const unsigned char *buf,
                                               unsigned long long (&w)[RPT][W]) {
#pragma unroll
    for (int k = 0; k < RPT; k++)
#pragma unroll
        for (int i = 0; i < W; i++) w[k][i] = 0;
    const unsigned tid = threadIdx.x;
    const unsigned char *sp = buf + tid * 4u;
#pragma unroll 1
    for (int c = 0; c < L.ncols; c++) {
        const bool is32 = cols.c[c].type == GSQL_T_INT32, iskey = c == L.key_col;
        const int wi = L.word[c];
        const int sh = L.half[c] == 1 ? 32 : 0;
        unsigned long long v[RPT];
        if (is32) {
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                int x = (FULL || (int)(k * THREADS + tid) < n_tile) ? *reinterpret_cast<const int *>(sp + k * THREADS * 4) : 0;
                v[k] = iskey ? (unsigned long long)(long long)x : (unsigned long long)(unsigned)x;
            }
            sp += TILE * 4;
        } else {
            const unsigned char *sp8 = sp + tid * 4u;
#pragma unroll
            for (int k = 0; k < RPT; k++)
                v[k] = (FULL || (int)(k * THREADS + tid) < n_tile) ? *reinterpret_cast<const unsigned long long *>(sp8 + k * THREADS * 8) : 0ULL;
            sp += TILE * 8;
        }
#pragma unroll
        for (int i = 0; i < W; i++)
            if (i == wi) {
#pragma unroll
                for (int k = 0; k < RPT; k++) w[k][i] |= v[k] << sh;
            }
    }
}

// ---- pass 2: pack rows and scatter them into partition order through a shared-memory staged tile.  PIPE: the
// column loads of tile t+1 are issued (cp.async into a second input buffer) before tile t is ranked, staged and
// flushed, so the HBM latency of the loads overlaps the shared-memory work of the previous tile.
template <int W, bool PIPE>
__global__ void __launch_bounds__(THREADS, 2) k_fj_scatter(const __grid_constant__ DColSet cols, const __grid_constant__ Layout L, PartGeom g,
                                                        const int64_t *__restrict__ offs, unsigned long long *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *stage = reinterpret_cast<unsigned long long *>(smem_raw);           // TILE * W
    unsigned long long *cur = stage + (size_t)TILE * W;                                      // P
    unsigned long long *delta = cur + g.P;                                                   // P
    unsigned int *hist = reinterpret_cast<unsigned int *>(delta + g.P);                      // P
    unsigned int *start = hist + g.P;                                                        // P
    unsigned short *spid = reinterpret_cast<unsigned short *>(start + g.P);                  // TILE
    unsigned char *inbuf = reinterpret_cast<unsigned char *>(spid + TILE);                   // PIPE: 2 * TILE * W * 8
    typedef cub::BlockScan<unsigned int, THREADS> BlockScan;
    __shared__ typename BlockScan::TempStorage scan_tmp;

    for (int p = threadIdx.x; p < g.P; p += THREADS) {
        cur[p] = (unsigned long long)offs[(int64_t)p * g.nblocks + blockIdx.x];
        hist[p] = 0;
    }
    __syncthreads();
    int64_t r0 = (int64_t)blockIdx.x * g.chunk;
    int64_t r1 = r0 + g.chunk < g.rows ? r0 + g.chunk : g.rows;
    if (PIPE && r0 < r1) {
        if (r1 - r0 >= TILE) tile_prefetch<true>(cols, L, r0, TILE, inbuf);
        else tile_prefetch<false>(cols, L, r0, (int)(r1 - r0), inbuf);
    }
    int it = 0;
    for (int64_t t0 = r0; t0 < r1; t0 += TILE, it++) {
        unsigned long long w[RPT][W];
        unsigned int pid[RPT], rank[RPT];
        const int n_tile = (int)(r1 - t0 < TILE ? r1 - t0 : TILE);
        const bool full = n_tile == TILE;
        if (PIPE) {
            const int64_t left = r1 - (t0 + TILE);
            unsigned char *nbuf = inbuf + (size_t)((it + 1) & 1) * TILE * W * 8;
            if (left >= TILE) tile_prefetch<true>(cols, L, t0 + TILE, TILE, nbuf);
            else if (left > 0) tile_prefetch<false>(cols, L, t0 + TILE, (int)left, nbuf);
            else cp_async_commit();
            cp_async_wait<1>();
            const unsigned char *cbuf = inbuf + (size_t)(it & 1) * TILE * W * 8;
            if (full) pack_tile_smem<W, true>(cols, L, n_tile, cbuf, w);
            else pack_tile_smem<W, false>(cols, L, n_tile, cbuf, w);
        } else {
            pack_tile<W>(cols, L, t0 + threadIdx.x, r1, w);
        }
        if (full) {
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                pid[k] = part_of(key_hash(w[k][0]), g.P);
                rank[k] = atomicAdd(&hist[pid[k]], 1u);
            }
        } else {
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                pid[k] = 0xffffffffu;
                if ((int)(k * THREADS + threadIdx.x) < n_tile) {
                    pid[k] = part_of(key_hash(w[k][0]), g.P);
                    rank[k] = atomicAdd(&hist[pid[k]], 1u);
                }
            }
        }
        __syncthreads();
        {  // exclusive scan of hist[0..P) -> start[]; delta[p] = (global cursor of p) - start[p]
            constexpr int IPT = MAX_P / THREADS;
            unsigned int v[IPT];
#pragma unroll
            for (int i = 0; i < IPT; i++) {
                int p = threadIdx.x * IPT + i;
                v[i] = p < g.P ? hist[p] : 0;
            }
            BlockScan(scan_tmp).ExclusiveSum(v, v);
#pragma unroll
            for (int i = 0; i < IPT; i++) {
                int p = threadIdx.x * IPT + i;
                if (p < g.P) {
                    start[p] = v[i];
                    unsigned long long c = cur[p];
                    delta[p] = c - v[i];
                    cur[p] = c + hist[p];
                    hist[p] = 0;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            if (full || pid[k] != 0xffffffffu) {
                unsigned int pos = start[pid[k]] + rank[k];
                if (W == 2) {
                    int4 v;
                    v.x = (int)(unsigned)w[k][0]; v.y = (int)(unsigned)(w[k][0] >> 32);
                    v.z = (int)(unsigned)w[k][W - 1]; v.w = (int)(unsigned)(w[k][W - 1] >> 32);
                    *reinterpret_cast<int4 *>(stage + (size_t)pos * 2) = v;
                } else {
#pragma unroll
                    for (int i = 0; i < W; i++) stage[(size_t)pos * W + i] = w[k][i];
                }
                spid[pos] = (unsigned short)pid[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RPT; k++) {  // consecutive threads -> consecutive addresses of a run
            const int i = k * THREADS + threadIdx.x;
            if (full || i < n_tile) {
                unsigned long long dst = delta[spid[i]] + (unsigned)i;
                if (W == 2) {
                    st_stream_16(out + dst * 2, *reinterpret_cast<const int4 *>(stage + (size_t)i * 2));
                } else {
#pragma unroll
                    for (int j = 0; j < W; j++) st_stream_8(out + dst * W + j, (long long)stage[(size_t)i * W + j]);
                }
            }
        }
        __syncthreads();  // stage / spid / delta are rewritten by the next tile
    }
}

// ---- pass 2, variant without shared-memory staging (GSQL_JOIN_SCATTER_DIRECT=1): every row is stored straight to its
// partition's run, 16 bytes at a time; the position comes from a per-block shared-memory counter per partition.  No block
// barrier in the tile loop, no staging traffic; the 16-byte stores of neighbouring rows of a run meet in L2 (the write
// frontier of a block is P sectors).  Measurement variant of r02 for the scatter's shared-memory bound (profiles/r01_ncu_summary.md).
template <int W>
__global__ void __launch_bounds__(THREADS, 2) k_fj_scatter_direct(const __grid_constant__ DColSet cols, const __grid_constant__ Layout L, PartGeom g,
                                                               const int64_t *__restrict__ offs, unsigned long long *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *base = reinterpret_cast<unsigned long long *>(smem_raw);  // P: first packed row of this block's run in partition p
    unsigned int *cnt = reinterpret_cast<unsigned int *>(base + g.P);             // P: rows of the block already placed there
    for (int p = threadIdx.x; p < g.P; p += THREADS) {
        base[p] = (unsigned long long)offs[(int64_t)p * g.nblocks + blockIdx.x];
        cnt[p] = 0;
    }
    __syncthreads();
    const int64_t r0 = (int64_t)blockIdx.x * g.chunk;
    const int64_t r1 = r0 + g.chunk < g.rows ? r0 + g.chunk : g.rows;
    for (int64_t t0 = r0; t0 < r1; t0 += TILE) {
        unsigned long long w[RPT][W];
        pack_tile<W>(cols, L, t0 + threadIdx.x, r1, w);
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            if (t0 + k * THREADS + threadIdx.x >= r1) continue;
            const unsigned int pid = part_of(key_hash(w[k][0]), g.P);
            const unsigned long long dst = base[pid] + atomicAdd(&cnt[pid], 1u);
            if (W == 2) {
                int4 v;
                v.x = (int)(unsigned)w[k][0]; v.y = (int)(unsigned)(w[k][0] >> 32);
                v.z = (int)(unsigned)w[k][W - 1]; v.w = (int)(unsigned)(w[k][W - 1] >> 32);
                *reinterpret_cast<int4 *>(out + dst * 2) = v;
            } else {
#pragma unroll
                for (int i = 0; i < W; i++) out[dst * W + i] = w[k][i];
            }
        }
    }
}

static size_t scatter_smem_bytes(int W, int P, bool pipe) {
    return (size_t)TILE * W * 8 + (size_t)P * 8 * 2 + (size_t)P * 4 * 2 + (size_t)TILE * 2 + (pipe ? (size_t)2 * TILE * W * 8 : 0);
}

// ---- table
template <int W>
__global__ void __launch_bounds__(THREADS) k_fj_table_init(unsigned long long *table, uint64_t nslots) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nslots; i += (uint64_t)gridDim.x * blockDim.x) {
        table[i * W] = KEY_EMPTY;
#pragma unroll
        for (int j = 1; j < W; j++) table[i * W + j] = 0;
    }
}

// Inserts rows (packed, or packed on the fly from columns when `packed` == nullptr).  Tiles are taken in index order
// so that concurrently running blocks work on neighbouring partitions (the table slice stays in L2).
template <int W>
__global__ void __launch_bounds__(THREADS, 2) k_fj_insert(const unsigned long long *__restrict__ packed, const __grid_constant__ DColSet cols,
                                                       const __grid_constant__ Layout L, int64_t n, unsigned long long *table, uint64_t nslots,
                                                       int32_t *flags) {
    for (int64_t t0 = (int64_t)blockIdx.x * TILE; t0 < n; t0 += (int64_t)gridDim.x * TILE) {
        unsigned long long w[RPT][W];
        if (packed) {
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                int64_t r = t0 + k * THREADS + threadIdx.x;
#pragma unroll
                for (int i = 0; i < W; i++) w[k][i] = r < n ? (unsigned long long)ld_stream_8(packed + r * W + i) : 0;
            }
        } else {
            pack_tile<W>(cols, L, t0 + threadIdx.x, n, w);
        }
        // CAS attempts in rounds: every round issues one attempt for all still-unplaced rows of the thread
        uint64_t sl[RPT];
        unsigned long long prev[RPT];
        bool pending[RPT];
        bool any = false;
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            int64_t r = t0 + k * THREADS + threadIdx.x;
            sl[k] = __umul64hi(key_hash(w[k][0]), nslots);
            pending[k] = r < n;
            if (pending[k] && w[k][0] == KEY_EMPTY) { flags[FL_SENTINEL] = 1; pending[k] = false; }
            any |= pending[k];
        }
        int disp = 0;
        while (__any_sync(0xffffffffu, any)) {
#pragma unroll
            for (int k = 0; k < RPT; k++)
                if (pending[k]) prev[k] = atomicCAS(table + sl[k] * W, KEY_EMPTY, w[k][0]);
            any = false;
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                if (!pending[k]) continue;
                if (prev[k] == KEY_EMPTY) {  // claimed: the payload words follow (visible after the kernel)
#pragma unroll
                    for (int i = 1; i < W; i++) table[sl[k] * W + i] = w[k][i];
                    pending[k] = false;
                } else if (prev[k] == w[k][0]) {  // duplicate build key: the generic (chained) path takes over
                    flags[FL_DUP] = 1;
                    pending[k] = false;
                } else {
                    if (++sl[k] == nslots) sl[k] = 0;
                }
                any |= pending[k];
            }
            if (++disp > MAX_DISP) { if (any) flags[FL_DISP] = 1; break; }
        }
    }
}

// ---- partitioned build: table initialisation fused with the inserts, one partition GROUP at a time (cooperative
// launch, one grid barrier per group).  A group's slice (~32 MB) is written EMPTY and then receives its CAS inserts
// while it is still dirty in L2, so the table crosses HBM once (the final write-back) instead of three times (init
// write-back, fetch on the first atomic, write-back again) — atomics on L2-resident lines run ~9x faster than on
// lines that miss (profiles/r01_microbench.txt).  Group g+1 is initialised BEFORE the barrier that precedes the
// inserts of group g, so linear probing that runs past the end of group g (bounded by MAX_DISP <= group size) only
// ever meets initialised slots.
template <int W>
__global__ void __launch_bounds__(THREADS, 2) k_fj_build_part(const unsigned long long *__restrict__ packed, const int64_t *__restrict__ offs,
                                                           int nblocks_hist, int P, int G, unsigned long long *table, uint64_t nslots, uint64_t spp,
                                                           int32_t *flags) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int ngroups = (P + G - 1) / G;
    const uint64_t gstride = (uint64_t)gridDim.x * THREADS, gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    auto init_group = [&](int g) {
        if (g >= ngroups) return;
        const uint64_t s0 = (uint64_t)g * G * spp;
        uint64_t s1 = s0 + (uint64_t)G * spp;
        if (s1 > nslots) s1 = nslots;
        for (uint64_t i = s0 + gtid; i < s1; i += gstride) {
            if (W == 2) {
                int4 v;
                v.x = 0; v.y = (int)0x80000000u; v.z = 0; v.w = 0;  // { KEY_EMPTY, 0 }
                *reinterpret_cast<int4 *>(table + i * 2) = v;
            } else {
                table[i * W] = KEY_EMPTY;
#pragma unroll
                for (int j = 1; j < W; j++) table[i * W + j] = 0;
            }
        }
    };
    init_group(0);
    for (int g = 0; g < ngroups; g++) {
        init_group(g + 1);
        grid.sync();
        const int pe = (g + 1) * G < P ? (g + 1) * G : P;
        const int64_t r0 = offs[(int64_t)g * G * nblocks_hist], r1 = offs[(int64_t)pe * nblocks_hist];
        for (int64_t t0 = r0 + (int64_t)blockIdx.x * THREADS; t0 < r1; t0 += (int64_t)gstride) {
            const int64_t r = t0 + threadIdx.x;
            unsigned long long w[W];
            bool pending = r < r1;
            if (pending) {
                if (W == 2) {
                    int4 v = ld_stream_16(packed + r * 2);
                    w[0] = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
                    w[W - 1] = ((unsigned long long)(unsigned)v.w << 32) | (unsigned)v.z;
                } else {
#pragma unroll
                    for (int i = 0; i < W; i++) w[i] = (unsigned long long)ld_stream_8(packed + r * W + i);
                }
                if (w[0] == KEY_EMPTY) { flags[FL_SENTINEL] = 1; pending = false; }
            } else {
#pragma unroll
                for (int i = 0; i < W; i++) w[i] = 0;
            }
            uint64_t sl = __umul64hi(key_hash(w[0]), nslots);
            int disp = 0;
            while (pending) {
                unsigned long long prev = atomicCAS(table + sl * W, KEY_EMPTY, w[0]);
                if (prev == KEY_EMPTY) {
#pragma unroll
                    for (int i = 1; i < W; i++) table[sl * W + i] = w[i];
                    pending = false;
                } else if (prev == w[0]) {
                    flags[FL_DUP] = 1;
                    pending = false;
                } else {
                    if (++sl == nslots) sl = 0;
                    if (++disp > MAX_DISP) { flags[FL_DISP] = 1; pending = false; }
                }
            }
        }
    }
}

// ---- probe
struct OutMap {
    void *data[GSQL_MAX_COLS * 2];
    uint8_t *nulls[GSQL_MAX_COLS * 2];
    int8_t side[GSQL_MAX_COLS * 2];  // 0 probe, 1 build
    int8_t word[GSQL_MAX_COLS * 2];
    int8_t half[GSQL_MAX_COLS * 2];
    int8_t is32[GSQL_MAX_COLS * 2];
    int32_t nout;
    int32_t join_type;
    int32_t stage_off[GSQL_MAX_COLS * 2];   // byte offset of column q inside a tile's shared-memory output staging
    int32_t stage_null_off[GSQL_MAX_COLS * 2];
    int32_t stage_bytes;                    // staging bytes per tile (PT_TILE rows)
    int32_t lookup_mode;                    // 0: ld.global.nc  1: + L1::no_allocate  2: cp.async gather through shared memory
};

template <int W>
__device__ __forceinline__ unsigned long long pick(const unsigned long long (&w)[W], int idx) {
    unsigned long long v = w[0];
#pragma unroll
    for (int i = 1; i < W; i++)
        if (i == idx) v = w[i];
    return v;
}


// Writes the emitted rows of one thread.  The column loop is outside; each column is dispatched ONCE (warp-uniform
// switch) to a body whose source register (probe word J / build payload word J-PW) and extraction (low 32, high 32,
// whole 64 bits) are compile-time, so a row-column costs an address computation and one store.  Consecutive lanes
// hold consecutive output positions: every store instruction is coalesced.
template <int R, int PW, int BP, int J, int H>
__device__ __forceinline__ void write_col(char *data, uint8_t *nulls, bool null_if_unmatched, const unsigned long long (&pw)[R][PW],
                                          const unsigned long long (&bp)[R][BP], const bool (&found)[R], const bool (&em)[R],
                                          const unsigned long long (&pos)[R], int32_t *flags) {
#pragma unroll
    for (int k = 0; k < R; k++) {
        if (!em[k]) continue;
        unsigned long long v = J < PW ? pw[k][J < PW ? J : 0] : bp[k][(J >= PW && J - PW < BP) ? J - PW : 0];
        const bool isnull = null_if_unmatched && !found[k];
        if (isnull) v = 0;
        if (nulls) nulls[pos[k]] = isnull ? 1 : 0;
        else if (isnull) flags[FL_NULLOUT] = 1;
        if (H == 0) st_stream_4(data + pos[k] * 4, (int)(unsigned)v);
        else if (H == 1) st_stream_4(data + pos[k] * 4, (int)(unsigned)(v >> 32));
        else st_stream_8(data + pos[k] * 8, (long long)v);
    }
}

// ---- word-wise output staging ---------------------------------------------------------------------------------
// A tile's emitted rows are compacted into shared memory as 8-byte WORD arrays (probe words, then build payload
// words): sw[j][li].  The flush then produces each output column with 16-byte stores that are 16-byte aligned in the
// OUTPUT (groups of 4 INT32 / 2 INT64-or-FP64 elements): ~2 instructions per row-column and full-sector writes
// (unaligned warp stores reach only ~half the HBM write rate on B200 — tools/membench.cu).
template <int R, int PW, int BP>
__device__ __forceinline__ void stage_words(char *staging, int tile_rows, bool want_flags, const unsigned long long (&pw)[R][PW],
                                            const unsigned long long (&bp)[R][BP], const bool (&found)[R], const bool (&em)[R],
                                            const unsigned int (&li)[R]) {
    unsigned long long *sw = reinterpret_cast<unsigned long long *>(staging);
    uint8_t *sf = reinterpret_cast<uint8_t *>(staging) + (size_t)(PW + BP) * tile_rows * 8;
#pragma unroll
    for (int k = 0; k < R; k++) {
        if (!em[k]) continue;
#pragma unroll
        for (int j = 0; j < PW; j++) sw[(size_t)j * tile_rows + li[k]] = pw[k][j];
#pragma unroll
        for (int j = 0; j < BP; j++) sw[(size_t)(PW + j) * tile_rows + li[k]] = found[k] ? bp[k][j] : 0ULL;
        if (want_flags) sf[li[k]] = found[k] ? 0 : 1;  // 1 = build side is NULL for this row
    }
}

// One element per lane: a warp reads 32 consecutive staged words (conflict-free) and writes one full, 128-byte
// ALIGNED line of an INT column (or two lines of a BIGINT column).  The lane -> element mapping is shifted so that
// line boundaries of the destination ADDRESS fall between warps (the column base only needs natural alignment);
// partially covered lines occur only at the two ends of the tile's run.  Everything per element is (per-column base) +
// compile-time offset: EPT unrolled stores plus one tail store for the `shift` elements the mapping pushed out.
template <int PW, int BP, int NT, int EPT>
__device__ __forceinline__ void flush_col(const OutMap &O, int q, const char *staging, unsigned long long base, int n, int32_t *flags) {
    constexpr int tile_rows = NT * EPT;
    const bool probe_side = O.side[q] == 0;
    const int j = probe_side ? O.word[q] : (O.word[q] == 0 ? 0 : PW + O.word[q] - 1);
    const char *src = staging + (size_t)j * tile_rows * 8;
    const int tid = (int)threadIdx.x;
    if (O.is32[q]) {
        int *dst = reinterpret_cast<int *>(O.data[q]) + base;
        const int l0 = tid - (int)(((unsigned long long)(uintptr_t)dst >> 2) & 31ULL);
        const unsigned int *sp = reinterpret_cast<const unsigned int *>(src) + (O.half[q] == 1 ? 1 : 0) + 2 * l0;  // element l at [2*l]
        int *dp = dst + l0;
#pragma unroll
        for (int k = 0; k < EPT; k++)
            if ((unsigned)(l0 + k * NT) < (unsigned)n) st_stream_4(dp + k * NT, (int)sp[2 * k * NT]);
        if (l0 + EPT * NT < n) st_stream_4(dp + EPT * NT, (int)sp[2 * EPT * NT]);
    } else {
        long long *dst = reinterpret_cast<long long *>(O.data[q]) + base;
        const int l0 = tid - (int)(((unsigned long long)(uintptr_t)dst >> 3) & 15ULL);
        const unsigned long long *sp = reinterpret_cast<const unsigned long long *>(src) + l0;
        long long *dp = dst + l0;
#pragma unroll
        for (int k = 0; k < EPT; k++)
            if ((unsigned)(l0 + k * NT) < (unsigned)n) st_stream_8(dp + k * NT, (long long)sp[k * NT]);
        if (l0 + EPT * NT < n) st_stream_8(dp + EPT * NT, (long long)sp[EPT * NT]);
    }
    if (O.nulls[q]) {  // NULL flags: only build-side columns of an outer join can be NULL here
        const uint8_t *sf = reinterpret_cast<const uint8_t *>(staging) + (size_t)(PW + BP) * tile_rows * 8;
        const bool outer = O.join_type == GSQL_JOIN_LEFT || O.join_type == GSQL_JOIN_RIGHT;
        uint8_t *nd = O.nulls[q] + base;
#pragma unroll
        for (int k = 0; k < EPT; k++)
            if (tid + k * NT < n) nd[tid + k * NT] = (!probe_side && outer) ? sf[tid + k * NT] : 0;
    } else if (!probe_side && (O.join_type == GSQL_JOIN_LEFT || O.join_type == GSQL_JOIN_RIGHT)) {
        if (tid == 0) flags[FL_NULLOUT] = 1;  // rejected on the host before launch
    }
}

template <int PW, int BP, int NT, int EPT>
__device__ __forceinline__ void flush_words(const OutMap &O, const char *staging, unsigned long long base, unsigned int cnt, int32_t *flags) {
    const int n = (int)cnt;
    // the first columns are unrolled: their descriptors become direct constant-bank operands
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (q < O.nout) flush_col<PW, BP, NT, EPT>(O, q, staging, base, n, flags);
#pragma unroll 1
    for (int q = 8; q < O.nout; q++) flush_col<PW, BP, NT, EPT>(O, q, staging, base, n, flags);
}

__host__ __device__ constexpr size_t stage_words_bytes_c(int PW, int BW, int tile_rows) {
    return (size_t)(PW + (BW > 1 ? BW - 1 : 1)) * tile_rows * 8 + (size_t)tile_rows;
}
static size_t probe_pipe_smem_bytes(int PW, int BW) { return ((stage_words_bytes_c(PW, BW, 2048) + 15) & ~(size_t)15) + (size_t)2048 * PW * 8; }
static size_t stage_words_bytes(int PW, int BW, int tile_rows) {
    int BP = BW > 1 ? BW - 1 : 1;
    return (size_t)(PW + BP) * tile_rows * 8 + (size_t)tile_rows;
}

template <int R, int PW, int BP, int J>
__device__ __forceinline__ void write_col_h(int h, char *data, uint8_t *nulls, bool nu, const unsigned long long (&pw)[R][PW],
                                            const unsigned long long (&bp)[R][BP], const bool (&found)[R], const bool (&em)[R],
                                            const unsigned long long (&pos)[R], int32_t *flags) {
    if (h == 0) write_col<R, PW, BP, J, 0>(data, nulls, nu, pw, bp, found, em, pos, flags);
    else if (h == 1) write_col<R, PW, BP, J, 1>(data, nulls, nu, pw, bp, found, em, pos, flags);
    else write_col<R, PW, BP, J, 2>(data, nulls, nu, pw, bp, found, em, pos, flags);
}

template <int R, int PW, int BP>
__device__ __forceinline__ void write_rows(const OutMap &O, const unsigned long long (&pw)[R][PW], const unsigned long long (&bp)[R][BP],
                                           const bool (&found)[R], const bool (&em)[R], const unsigned long long (&pos)[R], int32_t *flags) {
#pragma unroll 1
    for (int q = 0; q < O.nout; q++) {
        const bool probe_side = O.side[q] == 0;
        // source register index: probe words 0..PW-1, then build payload words; the build key equals the probe key
        const int j = probe_side ? O.word[q] : (O.word[q] == 0 ? 0 : PW + O.word[q] - 1);
        // an INT32 column stored whole in a word (widened key) is its low half
        const int h = O.is32[q] ? (O.half[q] == 1 ? 1 : 0) : 2;
        char *data = reinterpret_cast<char *>(O.data[q]);
        uint8_t *nulls = O.nulls[q];
        const bool nu = !probe_side;
        switch (j) {
        case 0: write_col_h<R, PW, BP, 0>(h, data, nulls, nu, pw, bp, found, em, pos, flags); break;
        case 1: write_col_h<R, PW, BP, 1>(h, data, nulls, nu, pw, bp, found, em, pos, flags); break;
        case 2: write_col_h<R, PW, BP, 2>(h, data, nulls, nu, pw, bp, found, em, pos, flags); break;
        case 3: write_col_h<R, PW, BP, 3>(h, data, nulls, nu, pw, bp, found, em, pos, flags); break;
        case 4: write_col_h<R, PW, BP, 4>(h, data, nulls, nu, pw, bp, found, em, pos, flags); break;
        case 5: write_col_h<R, PW, BP, 5>(h, data, nulls, nu, pw, bp, found, em, pos, flags); break;
        default: write_col_h<R, PW, BP, 6>(h, data, nulls, nu, pw, bp, found, em, pos, flags); break;
        }
    }
}


// Table lookups of the R rows a thread owns, organised in ROUNDS: every round issues the next slot read of all still
// unresolved rows before any result is consumed, so a tile costs (longest probe sequence) dependent L2 round trips
// instead of (sum over rows of the warp-wide longest sequence).  KEY_EMPTY rows (padding / the unbuildable key) never match.
template <int R, int PW, int BW, int BP>
__device__ __forceinline__ void lookup_rounds(const unsigned long long *__restrict__ table, uint64_t nslots, uint64_t pol,
                                              const unsigned long long (&pw)[R][PW], unsigned long long (&bp)[R][BP], bool (&found)[R],
                                              int mode = 0, int4 *gbuf = nullptr) {
    uint64_t slot[R];
    unsigned long long tk[R];
    bool pending[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
        slot[k] = __umul64hi(key_hash(pw[k][0]), nslots);
#pragma unroll
        for (int i = 0; i < BP; i++) bp[k][i] = 0;
    }
    if (BW == 2 && mode == 2) {
        // 16-byte slots gathered with cp.async.cg: the reads bypass L1 (no line is reserved per outstanding miss), land in
        // this thread's own cells of the tile's shared-memory area and are picked up after one wait
#pragma unroll
        for (int k = 0; k < R; k++) cp_async_16_hint(gbuf + k * THREADS + threadIdx.x, table + slot[k] * 2, pol);
        cp_async_commit();
        cp_async_wait<0>();
#pragma unroll
        for (int k = 0; k < R; k++) {
            int4 v = gbuf[k * THREADS + threadIdx.x];
            tk[k] = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
            bp[k][0] = ((unsigned long long)(unsigned)v.w << 32) | (unsigned)v.z;
        }
    } else {
#pragma unroll
        for (int k = 0; k < R; k++) {
            if (BW == 2) {
                int4 v = mode == 1 ? ld_keep_16_na(table + slot[k] * 2, pol) : ld_keep_16(table + slot[k] * 2, pol);
                tk[k] = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
                bp[k][0] = ((unsigned long long)(unsigned)v.w << 32) | (unsigned)v.z;
            } else {
                tk[k] = ld_keep_8(table + slot[k] * BW, pol);
            }
        }
    }
    bool any = false;
#pragma unroll
    for (int k = 0; k < R; k++) {
        pending[k] = pw[k][0] != KEY_EMPTY && tk[k] != pw[k][0] && tk[k] != KEY_EMPTY;
        any |= pending[k];
    }
    while (__any_sync(0xffffffffu, any)) {  // linear probing past other keys, all unresolved rows advance together
#pragma unroll
        for (int k = 0; k < R; k++) {
            if (pending[k]) {
                if (++slot[k] == nslots) slot[k] = 0;
                if (BW == 2) {
                    int4 v = ld_keep_16(table + slot[k] * 2, pol);
                    tk[k] = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
                    bp[k][0] = ((unsigned long long)(unsigned)v.w << 32) | (unsigned)v.z;
                } else {
                    tk[k] = ld_keep_8(table + slot[k] * BW, pol);
                }
            }
        }
        any = false;
#pragma unroll
        for (int k = 0; k < R; k++) {
            pending[k] = pending[k] && tk[k] != pw[k][0] && tk[k] != KEY_EMPTY;
            any |= pending[k];
        }
    }
#pragma unroll
    for (int k = 0; k < R; k++) {
        found[k] = pw[k][0] != KEY_EMPTY && tk[k] == pw[k][0];
        if (BW > 2 && found[k]) {
#pragma unroll
            for (int i = 1; i < BW; i++) bp[k][i - 1] = ld_keep_8(table + slot[k] * BW + i, pol);
        }
    }
}

struct ProbeShared {
    unsigned int cell[THREADS / 32][RPT];
    unsigned long long tile_base;
    unsigned int tile_total;
};

// Steps 2-4 of a probe tile, given the RPT packed probe rows of this thread in pw (rows beyond the batch carry KEY_EMPTY
// and live[k] = false): table lookups, emit decision, tile-wide compaction, staged column flush.
template <int PW, int BW>
__device__ __forceinline__ void probe_tile(unsigned long long (&pw)[RPT][PW], const bool (&live)[RPT], const unsigned long long *__restrict__ table,
                                           uint64_t nslots, uint64_t pol, const OutMap &O, unsigned long long *cursor, int32_t *flags,
                                           ProbeShared &sh, unsigned char *probe_stage) {
    constexpr int BP = BW > 1 ? BW - 1 : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long bp[RPT][BP];  // build payload words (the build key equals the probe key on a match)
    unsigned int ballot[RPT];
    bool found[RPT];
    // 2. table lookups in rounds (one L2-resident read per row per round, all rows of the thread in flight)
    lookup_rounds<RPT, PW, BW, BP>(table, nslots, pol, pw, bp, found, O.lookup_mode, reinterpret_cast<int4 *>(probe_stage));
    // 3. which rows emit (AbstractBufferedJoinExec.nextRows:185-264 for unique build keys, no NULLs)
#pragma unroll
    for (int k = 0; k < RPT; k++) {
        bool emit;
        switch (O.join_type) {
        case GSQL_JOIN_INNER: emit = found[k]; break;
        case GSQL_JOIN_SEMI: emit = found[k]; break;
        case GSQL_JOIN_ANTI: emit = !found[k]; break;
        default: emit = true; break;  // LEFT / RIGHT: unmatched probe rows are NULL-padded
        }
        emit = emit && live[k];
        ballot[k] = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) sh.cell[warp][k] = __popc(ballot[k]);
    }
    __syncthreads();
    unsigned long long my_base = 0;
    if (warp == 0) {  // exclusive scan over the 64 (warp, k) cells + one cursor bump for the tile
        constexpr int CELLS = (THREADS / 32) * RPT;
        unsigned int *flat = &sh.cell[0][0];
        unsigned int a = flat[lane * 2], b = flat[lane * 2 + 1];
        unsigned int sum = a + b, incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        unsigned int excl = incl - sum;
        unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) {  // the cursor bump's round trip overlaps the staging below: its result is only published after it
            my_base = total ? atomicAdd(cursor, (unsigned long long)total) : 0ULL;
            sh.tile_total = total;
        }
        flat[lane * 2] = excl;
        flat[lane * 2 + 1] = excl + a;
        static_assert(CELLS == 64, "cell scan assumes 64 cells");
    }
    __syncthreads();
    // 4. compact the tile's rows into shared memory (word arrays), then flush the columns with aligned full-line stores
    const bool want_flags = O.join_type == GSQL_JOIN_LEFT || O.join_type == GSQL_JOIN_RIGHT;
    bool em[RPT];
    unsigned int li[RPT];
#pragma unroll
    for (int k = 0; k < RPT; k++) {
        em[k] = (ballot[k] >> lane) & 1u;
        li[k] = sh.cell[warp][k] + __popc(ballot[k] & ((1u << lane) - 1u));
    }
    stage_words<RPT, PW, BP>(reinterpret_cast<char *>(probe_stage), TILE, want_flags, pw, bp, found, em, li);
    if (threadIdx.x == 0) sh.tile_base = my_base;
    __syncthreads();
    flush_words<PW, BP, THREADS, RPT>(O, reinterpret_cast<const char *>(probe_stage), sh.tile_base, sh.tile_total, flags);
}

// One tile per block; probe rows come from the input columns (packed == nullptr) or from packed rows.
template <int PW, int BW>
__global__ void __launch_bounds__(THREADS, 2) k_fj_probe(const unsigned long long *__restrict__ packed, const __grid_constant__ DColSet cols,
                                                      const __grid_constant__ Layout L, int64_t n, const unsigned long long *__restrict__ table,
                                                      uint64_t nslots, const __grid_constant__ OutMap O, unsigned long long *cursor, int32_t *flags) {
    __shared__ ProbeShared sh;
    extern __shared__ __align__(16) unsigned char probe_stage[];
    const uint64_t pol = l2_policy_evict_last();
    const int64_t t0 = (int64_t)blockIdx.x * TILE;
    unsigned long long pw[RPT][PW];
    bool live[RPT];
    // 1. stream the probe rows in (all RPT loads in flight)
    if (packed) {
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            int64_t r = t0 + k * THREADS + threadIdx.x;
            if (r < n) {
                if (PW == 2) {
                    int4 v = ld_stream_16(packed + r * 2);
                    pw[k][0] = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
                    pw[k][PW - 1] = ((unsigned long long)(unsigned)v.w << 32) | (unsigned)v.z;
                } else {
#pragma unroll
                    for (int i = 0; i < PW; i++) pw[k][i] = (unsigned long long)ld_stream_8(packed + r * PW + i);
                }
            } else {
#pragma unroll
                for (int i = 0; i < PW; i++) pw[k][i] = 0;
            }
        }
    } else {
        pack_tile<PW>(cols, L, t0 + threadIdx.x, n, pw);
    }
#pragma unroll
    for (int k = 0; k < RPT; k++) {
        live[k] = t0 + k * THREADS + threadIdx.x < n;
        if (!live[k]) pw[k][0] = KEY_EMPTY;
    }
    probe_tile<PW, BW>(pw, live, table, nslots, pol, O, cursor, flags, sh, probe_stage);
}

// Persistent variant for packed probe rows (the partitioned mode): blocks walk the tiles in index order (all resident
// blocks stay on neighbouring tiles, i.e. on the same L2-resident table slice) and the packed rows of the block's NEXT
// tile are fetched with cp.async into a shared-memory input buffer while the current tile is looked up, compacted and
// flushed — the HBM latency of the row stream leaves the per-tile dependency chain.  Each thread copies and reads back
// only its own rows, so the single input buffer is refilled as soon as the thread has moved its rows to registers.
template <int PW, int BW>
__global__ void __launch_bounds__(THREADS, 2) k_fj_probe_pipe(const unsigned long long *__restrict__ packed, int64_t n,
                                                           const unsigned long long *__restrict__ table, uint64_t nslots,
                                                           const __grid_constant__ OutMap O, unsigned long long *cursor,
                                                           unsigned long long *ticket, int32_t *flags) {
    __shared__ ProbeShared sh;
    extern __shared__ __align__(16) unsigned char probe_stage[];
    unsigned long long *inbuf = reinterpret_cast<unsigned long long *>(probe_stage + ((stage_words_bytes_c(PW, BW, TILE) + 15) & ~(size_t)15));
    const uint64_t pol = l2_policy_evict_last();
    const int64_t ntiles = (n + TILE - 1) / TILE;
    const unsigned tid = threadIdx.x;
    auto prefetch = [&](int64_t tile) {
        const int64_t t0 = tile * TILE;
        const unsigned long long *gp = packed + (t0 + tid) * PW;
        unsigned long long *sp = inbuf + (size_t)tid * PW;
        const int n_tile = (int)(n - t0 < TILE ? n - t0 : TILE);
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            if ((int)(k * THREADS + tid) < n_tile) {
                if (PW == 2) {
                    cp_async_16(sp + (size_t)k * THREADS * 2, gp + (size_t)k * THREADS * 2);
                } else {
#pragma unroll
                    for (int i = 0; i < PW; i++) cp_async_8(sp + (size_t)k * THREADS * PW + i, gp + (size_t)k * THREADS * PW + i);
                }
            }
        }
        cp_async_commit();
    };
    // Tiles are handed out through a global ticket, one tile ahead (the prefetch needs the next tile's index): with a
    // static tile -> block map a block that falls behind keeps reading a table slice the others have left, misses L2,
    // falls further behind (measured: 2x the DRAM reads of the one-tile-per-block kernel).
    __shared__ long long nxt[2];
    if (tid == 0) {
        nxt[0] = (long long)atomicAdd(ticket, 1ULL);
        nxt[1] = (long long)atomicAdd(ticket, 1ULL);
    }
    __syncthreads();
    int64_t tile = nxt[0];
    if (tile < ntiles) prefetch(tile);
    for (int it = 0; tile < ntiles; it++) {
        const int64_t next = nxt[(it + 1) & 1];
        const int64_t t0 = tile * TILE;
        const int n_tile = (int)(n - t0 < TILE ? n - t0 : TILE);
        unsigned long long pw[RPT][PW];
        bool live[RPT];
        cp_async_wait<0>();
#pragma unroll
        for (int k = 0; k < RPT; k++) {
            live[k] = (int)(k * THREADS + tid) < n_tile;
            if (live[k]) {
                if (PW == 2) {
                    int4 v = *reinterpret_cast<const int4 *>(inbuf + ((size_t)k * THREADS + tid) * 2);
                    pw[k][0] = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
                    pw[k][PW - 1] = ((unsigned long long)(unsigned)v.w << 32) | (unsigned)v.z;
                } else {
#pragma unroll
                    for (int i = 0; i < PW; i++) pw[k][i] = inbuf[((size_t)k * THREADS + tid) * PW + i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < PW; i++) pw[k][i] = 0;
                pw[k][0] = KEY_EMPTY;
            }
        }
        if (next < ntiles) prefetch(next);
        if (tid == 0) nxt[it & 1] = (long long)atomicAdd(ticket, 1ULL);  // slot of `tile`: everyone read it before the last barrier
        probe_tile<PW, BW>(pw, live, table, nslots, pol, O, cursor, flags, sh, probe_stage);
        __syncthreads();  // the staging area and the scan cells are rewritten by the next tile
        tile = next;
    }
}

// ------------------------------------------------------------------------------------------------ TMA-staged probe
// Persistent CTAs; a 4-stage shared-memory ring of packed probe tiles is kept full by one elected thread issuing 1-D
// bulk TMA copies (cp.async.bulk, mbarrier complete_tx), so ~48 KB of HBM reads per CTA are in flight at all times
// without holding registers.  Consumers read their rows from shared memory, issue all table reads (L2) before any is
// used, compact with warp ballots and bump the global output cursor once per tile.
constexpr int PT_THREADS = 256;
constexpr int PT_RPT = 4;
constexpr int PT_TILE = PT_THREADS * PT_RPT;  // 1024 rows
constexpr int PT_STAGES = 3;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

static size_t probe_tma_smem_bytes(int PW, int BW) { return (size_t)PT_STAGES * PT_TILE * PW * 8 + 64 + stage_words_bytes(PW, BW, PT_TILE); }

template <int PW, int BW>
__global__ void __launch_bounds__(PT_THREADS, 3) k_fj_probe_tma(const unsigned long long *__restrict__ packed, int64_t n,
                                                             const unsigned long long *__restrict__ table, uint64_t nslots,
                                                             const __grid_constant__ OutMap O, unsigned long long *cursor,
                                                             unsigned long long *ticket, int32_t *flags) {
    // `ticket` = tile ticket counter: tiles are handed out in index order so that all resident
    // CTAs stay within a narrow window of tiles (one or two partitions -> the table slice stays in L2).
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *ring = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *bars = ring + (size_t)PT_STAGES * PT_TILE * PW;
    char *staging = reinterpret_cast<char *>(bars + 8);
    __shared__ unsigned int cell[2][PT_THREADS / 32][PT_RPT];
    __shared__ unsigned long long tile_base[2];
    __shared__ unsigned int tile_total[2];
    __shared__ long long stage_tile[PT_STAGES];
    constexpr int BP = BW > 1 ? BW - 1 : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (n + PT_TILE - 1) / PT_TILE;
    const uint64_t pol_keep = l2_policy_evict_last();
    uint64_t pol_stream = 0;

    auto issue = [&](int stage) {  // called by thread 0 only: take the next tile ticket and start its bulk copy
        int64_t tile = (int64_t)atomicAdd(ticket, 1ULL);
        stage_tile[stage] = tile < ntiles ? tile : -1;
        if (tile >= ntiles) return;
        int64_t r0 = tile * PT_TILE;
        int64_t rows = n - r0 < PT_TILE ? n - r0 : PT_TILE;
        uint32_t bytes = (uint32_t)(((size_t)rows * PW * 8 + 15) & ~(size_t)15);  // the buffer has 16 B of slack
        mbar_expect_tx(&bars[stage], bytes);
        tma_load_1d(ring + (size_t)stage * PT_TILE * PW, packed + (size_t)r0 * PW, bytes, &bars[stage], pol_stream);
    };
    if (threadIdx.x == 0) {
        pol_stream = l2_policy_evict_first();
        for (int s = 0; s < PT_STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int s = 0; s < PT_STAGES; s++) issue(s);
    __syncthreads();

    for (int64_t i = 0;; i++) {
        const int stage = (int)(i % PT_STAGES);
        const int64_t tile = stage_tile[stage];
        if (tile < 0) break;  // tickets are monotonic: the first exhausted stage ends this CTA
        const uint32_t parity = (uint32_t)((i / PT_STAGES) & 1);
        const int db = (int)(i & 1);
        const int64_t t0 = tile * PT_TILE;
        mbar_wait(&bars[stage], parity);

        unsigned long long pw[PT_RPT][PW];
        unsigned long long bp[PT_RPT][BP];
        bool found[PT_RPT], em[PT_RPT];
        unsigned int ballot[PT_RPT];
        const unsigned long long *src = ring + (size_t)stage * PT_TILE * PW;
#pragma unroll
        for (int k = 0; k < PT_RPT; k++) {
            int idx = k * PT_THREADS + threadIdx.x;
            if (PW == 2) {
                int4 v = *reinterpret_cast<const int4 *>(src + (size_t)idx * 2);
                pw[k][0] = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
                pw[k][PW - 1] = ((unsigned long long)(unsigned)v.w << 32) | (unsigned)v.z;
            } else {
#pragma unroll
                for (int w = 0; w < PW; w++) pw[k][w] = src[(size_t)idx * PW + w];
            }
            if (t0 + idx >= n) pw[k][0] = KEY_EMPTY;
        }
        lookup_rounds<PT_RPT, PW, BW, BP>(table, nslots, pol_keep, pw, bp, found);
#pragma unroll
        for (int k = 0; k < PT_RPT; k++) {
            bool live = t0 + k * PT_THREADS + threadIdx.x < n;
            bool e;
            switch (O.join_type) {
            case GSQL_JOIN_INNER: e = found[k]; break;
            case GSQL_JOIN_SEMI: e = found[k]; break;
            case GSQL_JOIN_ANTI: e = !found[k]; break;
            default: e = true; break;  // LEFT / RIGHT: unmatched probe rows are NULL-padded
            }
            em[k] = e && live;
            ballot[k] = __ballot_sync(0xffffffffu, em[k]);
            if (lane == 0) cell[db][warp][k] = __popc(ballot[k]);
        }
        __syncthreads();  // (A) every thread holds its rows in registers: the stage can be refilled; cells are complete
        if (threadIdx.x == 0) issue(stage);
        if (warp == 0) {  // exclusive scan over the 32 (warp, k) cells + one cursor bump for the tile
            unsigned int *flat = &cell[db][0][0];
            unsigned int a = flat[lane], incl = a;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) {
                tile_base[db] = total ? atomicAdd(cursor, (unsigned long long)total) : 0ULL;
                tile_total[db] = total;
            }
            flat[lane] = incl - a;
            static_assert((PT_THREADS / 32) * PT_RPT == 32, "cell scan assumes 32 cells");
        }
        __syncthreads();  // (B)
        unsigned int li[PT_RPT];
#pragma unroll
        for (int k = 0; k < PT_RPT; k++) li[k] = cell[db][warp][k] + __popc(ballot[k] & ((1u << lane) - 1u));
        stage_words<PT_RPT, PW, BP>(staging, PT_TILE, O.join_type == GSQL_JOIN_LEFT || O.join_type == GSQL_JOIN_RIGHT, pw, bp, found, em, li);
        __syncthreads();  // (C) the tile's output is dense in shared memory
        flush_words<PW, BP, PT_THREADS, PT_RPT>(O, staging, tile_base[db], tile_total[db], flags);
        // no barrier needed here: the next staging writes come after the next iteration's (A) and (B)
    }
}

}  // namespace fj

struct JoinFast {
    bool eligible = false;   // decided at create: shape supports the packed single-key path
    bool enabled = false;    // table built and usable
    fj::Layout bl, pl;       // build / probe packed-row layouts
    int P = 1;               // partitions (1 = table small enough to stay in L2 without partitioning)
    uint64_t nslots = 0;
    DevBuf table, flags, cursor;
    int64_t part_bytes = 16ll << 20;
    int64_t sub_batch = 256ll << 20;  // probe rows per partition+probe round (bounds scratch memory)
    int64_t part_min_rows = 1ll << 20;  // smaller probe batches skip the partitioning passes
};
