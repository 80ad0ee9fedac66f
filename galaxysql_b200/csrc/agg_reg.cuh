// agg_reg.cuh — REGISTER accumulators for the NULL-free, very-low-cardinality group-by with fp64 sums (the TPC-H Q1
// shape as the plan hands it over: 2 integer keys, SUM / AVG over DOUBLE columns and fused derived expressions,
// COUNT(*), an integer row filter).  Included by agg.cu after agg_lane.cuh.
//
// Why another kernel: k_agg_lane keeps every accumulator in shared memory and decodes a plan per step; r01 measured it
// at 12.3 warp-instructions per ROW (~390 thread instructions) and 21 % issue utilisation: 22.5 ms for the 26.4 GB of
// config 3, 18 % of the HBM roofline.  Here a thread owns RG_G x NSRC fp64 accumulators and RG_G counters in REGISTERS;
// a row costs its loads, one compare per group and NSRC predicated DADDs per group — ~100 thread instructions with
// NSRC x RG_G independent dependency chains.  The input columns of a tile reach shared memory through cp.async, one tile
// ahead of the accumulation (double buffer), so the HBM latency of the stream is never on a thread's critical path and no
// register holds a load in flight.  The (at most RG_G) group keys live in a per-block dictionary in shared memory that is
// only written when a new key shows up (the first tiles; warp-cooperative under a block lock, no block barrier in the row
// loop); a ninth key makes the row take the generic path on the spot (find_group_kv + L2 atomics), counted in
// C_FALLBACK so that the host stops choosing this kernel when that is common.  Everything that depends on the plan —
// the number of sums, the expression kind of each — is a template parameter or a warp-uniform branch on a
// __grid_constant__ descriptor; nothing is interpreted per row.
// When a block finishes, the register accumulators are reduced over the block (shuffles, then one shared-memory pass)
// and one thread per group folds them into the global table — exactly the merge of the other privatised kernels.
//
// Reference behaviour: AggOpenHashMap.putChunk (EX/operator/util/AggOpenHashMap.java:100-139) with CountRow,
// Double2DoubleSum (LittleNum2DoubleSum.java:40-64) and SpecificType2DoubleAvgV2 (:51-84); same groups, sums added in a
// different order (within the north_star's 1e-6 relative tolerance), counts bit-exact.
#pragma once

namespace {

constexpr int RG_THREADS = 256;
constexpr int RG_RPT = 4;       // rows per thread per tile
constexpr int RG_TILE = RG_THREADS * RG_RPT;
constexpr int RG_G = 8;         // groups a block can hold in registers (6 when there are 5-6 sums: 128 registers, two blocks per SM, no spills)
constexpr int RG_MAX_SRC = 6;   // distinct fp64 sums
constexpr int RG_MAX_USED = 8;  // distinct input columns a tile stages

struct RegSrc {
    int32_t kind;        // 0: column a; GSQL_EXPR_MUL_1MINUS: a*(1-b); GSQL_EXPR_MUL_1MINUS_1PLUS: a*(1-b)*(1+c)
    int32_t ua, ub, uc;  // staged-column slots
    int32_t oa, ob, oc;  // byte offsets of the operands' regions inside a tile buffer (host-flattened: with the source index a
    int32_t pad;         // compile-time constant after unrolling, these are direct constant-bank operands, no table walk)
};

struct RegPlan {
    int32_t nsrc, nkeys, nused, tile_bytes;
    RegSrc src[RG_MAX_SRC];
    int32_t used_col[RG_MAX_USED];   // input column of each staged slot
    int32_t used_off[RG_MAX_USED];   // byte offset of the slot's RG_TILE elements inside a tile buffer
    int32_t used_w[RG_MAX_USED];     // 4 or 8
    int32_t key_u[2];
    int32_t key_off[2], key_w[2];    // flattened copies for the row loop
    int32_t rf_u;                    // slot of the row-filter column, -1: no filter
    int32_t rf_off, rf_w, rf_neg;    // the filter as an interval test: pass = (lo <= x && x <= hi) != neg
    int64_t rf_lo, rf_hi;
    int32_t agg_src[GSQL_MAX_AGGS];  // per aggregate: the sum it reports, -1 for COUNT / COUNT(*)
    int32_t bulk;                    // 1: every staged column is 16-byte aligned at row0 -> full tiles arrive by bulk copy
    int32_t bulk_bytes;              // bytes one full tile brings in (sum of RG_TILE * used_w)
};

__device__ __forceinline__ void rg_cp_async_4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void rg_cp_async_8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void rg_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rg_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- bulk-copy staging (cp.async.bulk + mbarrier complete_tx; SASS: UBLKCP / SYNCS): one thread starts a whole tile —
// one instruction per staged column — instead of every thread issuing four predicated LDGSTS per column with their
// address arithmetic (~30% of the instructions the kernel executed per row, r02 SASS count)
__device__ __forceinline__ uint32_t rg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rg_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rg_mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rg_mbar_wait(unsigned long long *bar, uint32_t parity) {
    const uint32_t addr = rg_smem_u32(bar);
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void rg_bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(rg_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(rg_smem_u32(bar)), "l"(pol)
                 : "memory");
}
// One FULL tile, called by one thread: arm the stage's barrier with the tile's byte count, then one copy per column.
__device__ __forceinline__ void rg_bulk_issue(const AggParams &P, const RegPlan &L, int64_t t0, unsigned char *buf, unsigned long long *bar, uint64_t pol) {
    rg_mbar_expect_tx(bar, (uint32_t)L.bulk_bytes);
#pragma unroll 1
    for (int u = 0; u < L.nused; u++) {
        const DCol &c = P.in.c[L.used_col[u]];
        const uint32_t w = (uint32_t)L.used_w[u];
        rg_bulk_load(buf + L.used_off[u], reinterpret_cast<const unsigned char *>(c.data) + (size_t)(P.row0 + t0) * w, RG_TILE * w, bar, pol);
    }
}

// Asynchronous copy (LDGSTS) of one tile's staged columns into `buf`: element (k * RG_THREADS + tid) of every column is
// copied — and later read back — by the same thread, so no block barrier is needed, only the thread's own wait_group.
__device__ __forceinline__ void rg_prefetch(const AggParams &P, const RegPlan &L, int64_t t0, unsigned char *buf) {
    const int tid = threadIdx.x;
    const int64_t left = P.rows - t0;
#pragma unroll 1
    for (int u = 0; u < L.nused; u++) {
        const DCol &c = P.in.c[L.used_col[u]];
        unsigned char *dst = buf + L.used_off[u];
        if (L.used_w[u] == 4) {
            const int *src = reinterpret_cast<const int *>(c.data) + P.row0 + t0;
#pragma unroll
            for (int k = 0; k < RG_RPT; k++) {
                const int i = k * RG_THREADS + tid;
                if (i < left) rg_cp_async_4(dst + (size_t)i * 4, src + i);
            }
        } else {
            const long long *src = reinterpret_cast<const long long *>(c.data) + P.row0 + t0;
#pragma unroll
            for (int k = 0; k < RG_RPT; k++) {
                const int i = k * RG_THREADS + tid;
                if (i < left) rg_cp_async_8(dst + (size_t)i * 8, src + i);
            }
        }
    }
    rg_cp_async_commit();
}

// acc += v when gk == G0 — one ISETP and one predicated DADD (the `hit ? v : 0.0` form compiles to two FSELs and a DADD)
template <int G0>
__device__ __forceinline__ void rg_pred_add(double &acc, double v, int gk) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\t@p add.f64 %0, %0, %1;\n\t}" : "+d"(acc) : "d"(v), "r"(gk), "n"(G0));
}

__device__ __forceinline__ void reg_decode_key(const AggParams &P, const RegPlan &L, unsigned long long k, int64_t (&kv)[GSQL_MAX_KEYS], bool (&kn)[GSQL_MAX_KEYS]) {
    for (int c = 0; c < GSQL_MAX_KEYS; c++) { kv[c] = 0; kn[c] = false; }
    if (L.nkeys == 2) {
        kv[0] = (int64_t)(int32_t)(unsigned int)k;
        kv[1] = (int64_t)(int32_t)(unsigned int)(k >> 32);
    } else {
        kv[0] = L.key_w[0] == 4 ? (int64_t)(int32_t)(unsigned int)k : (int64_t)k;
    }
}

// Branch-free accumulate: G x NSRC select-and-add (ptxas turns the predicated fp64 add into two FSELs and a DADD).  The
// alternative — one divergent arm per group — was measured slower on the Q1 shape (10.2 vs 9.6 ms for 600 M rows, r02).
template <int NSRC, int G, int G0>
__device__ __forceinline__ void rg_accumulate(double (&acc)[G][NSRC], unsigned int (&cnt)[G], const double (&v)[NSRC], int gk) {
    if constexpr (G0 < G) {
        cnt[G0] += gk == G0 ? 1u : 0u;
#pragma unroll
        for (int j = 0; j < NSRC; j++) rg_pred_add<G0>(acc[G0][j], v[j], gk);
        rg_accumulate<NSRC, G, G0 + 1>(acc, cnt, v, gk);
    }
}

// One-hot accumulate: acc[g][j] = fma(v[j], m_g, acc[g][j]) with m_g = 1.0 for the row's group and 0.0 for the others —
// one DFMA per (group, sum) instead of a DADD and two FSELs.  Bit-identical to the select form for FINITE v: v * 1.0 is
// exact, v * 0.0 is a signed zero, and an accumulator that starts at +0.0 is never -0.0, so acc + (+-0.0) == acc.  A
// non-finite v (Inf * 0.0 = NaN would leak into the other groups) takes the select form: the caller tests the exponents.
template <int NSRC, int G>
__device__ __forceinline__ void rg_accumulate_onehot(double (&acc)[G][NSRC], unsigned int (&cnt)[G], const double (&v)[NSRC], int gk) {
#pragma unroll
    for (int g = 0; g < G; g++) {
        const bool hit = gk == g;
        cnt[g] += hit ? 1u : 0u;
        const double m = __hiloint2double(hit ? 0x3ff00000 : 0, 0);
#pragma unroll
        for (int j = 0; j < NSRC; j++) acc[g][j] = __fma_rn(v[j], m, acc[g][j]);
    }
}

template <int NSRC, int G>
__global__ void __launch_bounds__(RG_THREADS, 2) k_agg_reg(const __grid_constant__ AggParams P, const __grid_constant__ RegPlan L) {
    extern __shared__ __align__(16) unsigned char rg_smem[];  // two tile buffers
    __shared__ unsigned long long skey[G];
    __shared__ int s_ng, s_lock;
    __shared__ double red[RG_THREADS / 32][G][NSRC];
    __shared__ unsigned long long redc[RG_THREADS / 32][G];
    __shared__ __align__(8) unsigned long long tile_bar[2];  // one per tile buffer (bulk staging only)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool bulk = L.bulk != 0;
    uint64_t pol_stream = 0;
    if (tid == 0) {
        s_ng = 0;
        s_lock = 0;
        if (bulk) {
            rg_mbar_init(&tile_bar[0], 1);
            rg_mbar_init(&tile_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        }
    }
    double acc[G][NSRC];
    unsigned int cnt[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
        cnt[g] = 0;
#pragma unroll
        for (int j = 0; j < NSRC; j++) acc[g][j] = 0.0;
    }
    __syncthreads();
    unsigned long long fallback_rows = 0;
    const int64_t ntiles = (P.rows + RG_TILE - 1) / RG_TILE;
    int64_t tile = blockIdx.x;
    // A full tile of an aligned batch arrives by bulk copy (thread 0 issues, everyone waits on the stage's mbarrier); the
    // ragged last tile — and every tile of an unaligned batch — by per-thread cp.async.  Every thread commits exactly one
    // cp.async group per staged tile either way (an empty one for a bulk tile), so wait_group<1> below keeps its meaning.
    auto stage_tile = [&](int64_t tl, int st) {
        const int64_t s0 = tl * RG_TILE;
        unsigned char *dst = rg_smem + (size_t)st * L.tile_bytes;
        if (bulk && P.rows - s0 >= RG_TILE) {
            if (tid == 0) rg_bulk_issue(P, L, s0, dst, &tile_bar[st], pol_stream);
            rg_cp_async_commit();
        } else {
            rg_prefetch(P, L, s0, dst);
        }
    };
    if (tile < ntiles) stage_tile(tile, 0);
    unsigned long long kk[G];  // this thread's copy of the block's key dictionary
    int ng = 0;
    uint32_t bar_phase = 0;  // bit s: parity the next wait on stage s uses
    for (int it = 0; tile < ntiles; tile += gridDim.x, it++) {
        const int64_t t0 = tile * RG_TILE;
        const int64_t next = tile + gridDim.x;
        const int st = it & 1;
        // the next tile's columns start their way from HBM before this tile is touched: the memory latency of the stream
        // is hidden behind the accumulation of a whole tile instead of being paid once per tile
        if (next < ntiles) stage_tile(next, st ^ 1);
        else rg_cp_async_commit();
        rg_cp_async_wait<1>();
        const unsigned char *buf = rg_smem + (size_t)st * L.tile_bytes;
        const int64_t left = P.rows - t0;
        if (bulk && left >= RG_TILE) {
            rg_mbar_wait(&tile_bar[st], (bar_phase >> st) & 1u);
            bar_phase ^= 1u << st;
        }
        {  // refresh the register copy of the dictionary when another warp has added keys
            const int now = *reinterpret_cast<volatile int *>(&s_ng);
            if (now != ng) {
#pragma unroll
                for (int g = 0; g < G; g++)
                    if (g < now) kk[g] = *reinterpret_cast<volatile unsigned long long *>(&skey[g]);
                ng = now;
            }
        }
#pragma unroll 1
        for (int k = 0; k < RG_RPT; k++) {
            const int i = k * RG_THREADS + tid;
            bool pass = i < left;
            const unsigned char *row8 = buf + (size_t)i * 8, *row4 = buf + (size_t)i * 4;
            unsigned long long key = L.key_w[0] == 4 ? (unsigned long long)*reinterpret_cast<const unsigned int *>(row4 + L.key_off[0])
                                                     : *reinterpret_cast<const unsigned long long *>(row8 + L.key_off[0]);
            if (L.nkeys == 2) key |= (unsigned long long)*reinterpret_cast<const unsigned int *>(row4 + L.key_off[1]) << 32;
            if (L.rf_u >= 0) {  // fused scan-side predicate as an interval test (NULL-free column: checked by the plan)
                const long long x = L.rf_w == 4 ? (long long)*reinterpret_cast<const int *>(row4 + L.rf_off) : *reinterpret_cast<const long long *>(row8 + L.rf_off);
                pass = pass && ((x >= L.rf_lo && x <= L.rf_hi) != (L.rf_neg != 0));
            }
            int gid = -1;
#pragma unroll
            for (int g = 0; g < G; g++)
                if (g < ng && kk[g] == key) gid = g;
            // ---- a key this thread has not seen: add it to the block's dictionary (first tiles only); warp-cooperative, one
            // distinct key at a time, lane 0 takes the block's lock — no block barrier anywhere in the row loop
            unsigned need = __ballot_sync(0xffffffffu, pass && gid < 0);
            while (need) {
                const int leader = __ffs(need) - 1;
                const unsigned long long lk = __shfl_sync(0xffffffffu, key, leader);
                int got = -1;
                if (lane == 0) {
                    while (atomicCAS(&s_lock, 0, 1) != 0) {}
                    const int cur = *reinterpret_cast<volatile int *>(&s_ng);
                    for (int g = 0; g < cur; g++)
                        if (*reinterpret_cast<volatile unsigned long long *>(&skey[g]) == lk) got = g;
                    if (got < 0 && cur < G) {
                        *reinterpret_cast<volatile unsigned long long *>(&skey[cur]) = lk;
                        __threadfence_block();
                        *reinterpret_cast<volatile int *>(&s_ng) = cur + 1;
                        got = cur;
                    } else if (got < 0) {
                        got = -2;  // the dictionary is full: these rows take the generic path
                    }
                    __threadfence_block();
                    atomicExch(&s_lock, 0);
                }
                got = __shfl_sync(0xffffffffu, got, 0);
                if (pass && gid < 0 && key == lk) gid = got;
                if (got >= 0) {  // every lane learns the new entry
                    const int now = *reinterpret_cast<volatile int *>(&s_ng);
#pragma unroll
                    for (int g = 0; g < G; g++)
                        if (g < now) kk[g] = *reinterpret_cast<volatile unsigned long long *>(&skey[g]);
                    ng = now;
                }
                need = __ballot_sync(0xffffffffu, pass && gid == -1);
            }
            if (pass && gid == -2) {  // a ninth key: the generic path, right here
                fallback_rows++;
                int64_t kv[GSQL_MAX_KEYS];
                bool kn[GSQL_MAX_KEYS];
                reg_decode_key(P, L, key, kv, kn);
                const int64_t r = P.row0 + t0 + i;
                const int g2 = find_group_kv(P, kv, kn, digest_of_keys(P, kv, kn));
                if (g2 < 0) {
                    unsigned long long o = atomicAdd(&P.counters[C_OVERFLOW], 1ULL);
                    P.overflow_rows[o] = r;
                } else {
                    for (int a = 0; a < P.naggs; a++) accumulate(P, P.agg[a], g2, r);
                }
                pass = false;
            }
            // ---- accumulate: the row's values come out of shared memory (a dead row reads its own, in-buffer cell and is
            // masked by gk = -1), one compare per group, NSRC predicated adds
            double v[NSRC];
#pragma unroll
            for (int j = 0; j < NSRC; j++) {
                const RegSrc &sr = L.src[j];
                double x = *reinterpret_cast<const double *>(row8 + sr.oa);
                if (sr.kind != 0) {  // block-uniform
                    x = x * (1.0 - *reinterpret_cast<const double *>(row8 + sr.ob));
                    if (sr.kind == GSQL_EXPR_MUL_1MINUS_1PLUS) x = x * (1.0 + *reinterpret_cast<const double *>(row8 + sr.oc));
                }
                v[j] = x;
            }
            const int gk = pass ? gid : -1;
            int worst = 0;  // largest exponent field among the row's values: 0x7ff = Inf / NaN
#pragma unroll
            for (int j = 0; j < NSRC; j++) worst = max(worst, __double2hiint(v[j]) & 0x7ff00000);
            if (worst != 0x7ff00000) rg_accumulate_onehot<NSRC, G>(acc, cnt, v, gk);
            else rg_accumulate<NSRC, G, 0>(acc, cnt, v, gk);
        }
        // a bulk copy rewrites cells other threads read: everyone must be done with this buffer before thread 0 refills it
        // (at the top of the iteration after next's staging call, i.e. the very next statement executed by thread 0)
        if (bulk) __syncthreads();
    }
    rg_cp_async_wait<0>();
    if (fallback_rows) atomicAdd(&P.counters[C_FALLBACK], fallback_rows);
    // ---- merge: reduce over the warp with shuffles, over the block through shared memory, then one thread per group
    __syncthreads();
    const int ngf = s_ng;
#pragma unroll
    for (int g = 0; g < G; g++) {
        if (g >= ngf) break;
        unsigned int c = cnt[g];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) redc[warp][g] = c;
#pragma unroll
        for (int j = 0; j < NSRC; j++) {
            const double sm = warp_sum_f64(acc[g][j]);
            if (lane == 0) red[warp][g][j] = sm;
        }
    }
    __syncthreads();
    if (tid < ngf) {
        const int g = tid;
        unsigned long long c = 0;
        double sums[NSRC];
#pragma unroll
        for (int j = 0; j < NSRC; j++) sums[j] = 0.0;
        for (int w = 0; w < RG_THREADS / 32; w++) {
            c += redc[w][g];
#pragma unroll
            for (int j = 0; j < NSRC; j++) sums[j] += red[w][g][j];
        }
        if (c) {  // (a key whose rows all failed the filter never entered the dictionary)
            int64_t kv[GSQL_MAX_KEYS];
            bool kn[GSQL_MAX_KEYS];
            reg_decode_key(P, L, skey[g], kv, kn);
            // the merge may exceed gcap by at most blocks x G groups: covered by the arrays' slack (ignore_cap)
            const int gl = find_group_kv(P, kv, kn, digest_of_keys(P, kv, kn), true);
            for (int a = 0; a < P.naggs; a++) {
                const AggDev &ag = P.agg[a];
                double sv = 0.0;
#pragma unroll
                for (int j = 0; j < NSRC; j++)
                    if (L.agg_src[a] == j) sv = sums[j];
                switch (ag.kind) {
                case GSQL_AGG_COUNT_STAR: case GSQL_AGG_COUNT:
                    atomicAdd(reinterpret_cast<unsigned long long *>(&ag.l[gl]), c);
                    break;
                case GSQL_AGG_SUM:
                    atomicAdd(&ag.d[gl], sv);
                    ag.has[gl] = 1;
                    break;
                default:  // AVG
                    atomicAdd(&ag.d[gl], sv);
                    atomicAdd(reinterpret_cast<unsigned long long *>(&ag.l[gl]), c);
                    ag.has[gl] = 1;
                    break;
                }
            }
        }
    }
}

}  // namespace

struct AggReg {
    bool shape_ok = false;  // decided at create
    bool enabled = false;   // adaptive
    int64_t rows_seen = 0, rows_fallback = 0;
};

// Create-time check of everything that does not depend on the batch.
static void agg_reg_check(AggReg *F, const gsql_agg_spec &spec, int nkeys, int naggs, const gsql_agg_call *aggs, const int32_t *in_type) {
    F->shape_ok = F->enabled = false;
    if (getenv("GSQL_AGG_NO_REG") && atoi(getenv("GSQL_AGG_NO_REG"))) return;
    if (nkeys < 1 || nkeys > 2) return;
    for (int k = 0; k < nkeys; k++) {
        const int t = spec.input_types[spec.groups[k]];
        if (t == GSQL_T_FP64) return;
        if (nkeys == 2 && t != GSQL_T_INT32) return;  // two keys must share one 64-bit word
    }
    for (int a = 0; a < naggs; a++) {
        if (aggs[a].filter_arg >= 0) return;
        switch (aggs[a].kind) {
        case GSQL_AGG_COUNT_STAR: case GSQL_AGG_COUNT: break;
        case GSQL_AGG_SUM: case GSQL_AGG_AVG:
            if (in_type[a] != GSQL_T_FP64) return;
            if (aggs[a].cols[0] < spec.n_input_cols && spec.input_types[aggs[a].cols[0]] != GSQL_T_FP64) return;
            break;
        default: return;
        }
    }
    if (spec.expected_groups > 4096) return;  // the planner expects far more groups than a block's registers hold
    F->shape_ok = F->enabled = true;
}

// Per-batch plan (NULL buffers are a property of the batch).  false = this batch takes another kernel.
static bool agg_reg_plan(RegPlan *Lp, const gsql_agg_spec &spec, int nkeys, int naggs, const gsql_agg_call *aggs, const DColSet &in) {
    RegPlan &L = *Lp;
    memset(&L, 0, sizeof(L));
    L.nkeys = nkeys;
    L.rf_u = -1;
    auto has_nulls = [&](int col) { return in.c[col].nulls != nullptr; };
    bool too_many = false;
    auto slot_of = [&](int col) -> int {  // staged slot of an input column
        for (int u = 0; u < L.nused; u++)
            if (L.used_col[u] == col) return u;
        if (L.nused == RG_MAX_USED) { too_many = true; return 0; }
        L.used_col[L.nused] = col;
        L.used_w[L.nused] = in.c[col].type == GSQL_T_INT32 ? 4 : 8;
        return L.nused++;
    };
    for (int k = 0; k < nkeys; k++) {
        if (has_nulls(spec.groups[k])) return false;
        L.key_u[k] = slot_of(spec.groups[k]);
    }
    if (spec.row_filter_op != GSQL_CMP_NONE) {
        if (has_nulls(spec.row_filter_col)) return false;
        L.rf_u = slot_of(spec.row_filter_col);
    }
    for (int a = 0; a < naggs; a++) {
        L.agg_src[a] = -1;
        const gsql_agg_call &c = aggs[a];
        if (c.kind == GSQL_AGG_COUNT) {  // COUNT(x...) == COUNT(*) exactly when no argument can be NULL
            for (int q = 0; q < c.ncols; q++) {
                const int col = c.cols[q];
                if (col < spec.n_input_cols) { if (has_nulls(col)) return false; }
                else {
                    const gsql_derived_col &d = spec.derived[col - spec.n_input_cols];
                    if (has_nulls(d.a) || has_nulls(d.b) || (d.kind == GSQL_EXPR_MUL_1MINUS_1PLUS && has_nulls(d.c))) return false;
                }
            }
            continue;
        }
        if (c.kind == GSQL_AGG_COUNT_STAR) continue;
        int kind = 0, oa, ob = 0, oc = 0;
        const int col = c.cols[0];
        if (col < spec.n_input_cols) {
            oa = col;
        } else {
            const gsql_derived_col &d = spec.derived[col - spec.n_input_cols];
            kind = d.kind; oa = d.a; ob = d.b; oc = d.kind == GSQL_EXPR_MUL_1MINUS_1PLUS ? d.c : d.b;
        }
        const int ops[3] = {oa, ob, oc};
        for (int q = 0; q < (kind == 0 ? 1 : 3); q++)
            if (has_nulls(ops[q]) || in.c[ops[q]].type != GSQL_T_FP64) return false;
        RegSrc s;
        s.kind = kind;
        s.ua = slot_of(oa);
        s.ub = kind ? slot_of(ob) : 0;
        s.uc = kind == GSQL_EXPR_MUL_1MINUS_1PLUS ? slot_of(oc) : 0;
        int at = -1;
        for (int j = 0; j < L.nsrc; j++)
            if (L.src[j].kind == s.kind && L.src[j].ua == s.ua && L.src[j].ub == s.ub && L.src[j].uc == s.uc) at = j;
        if (at < 0) {
            if (L.nsrc == RG_MAX_SRC) return false;
            at = L.nsrc;
            L.src[L.nsrc++] = s;
        }
        L.agg_src[a] = at;
    }
    if (too_many || L.nsrc == 0) return false;  // (pure COUNT(*) shapes stay on the lane kernel)
    int off = 0;  // 8-byte columns first: every region keeps its natural alignment
    for (int pass = 0; pass < 2; pass++)
        for (int u = 0; u < L.nused; u++)
            if ((L.used_w[u] == 8) == (pass == 0)) {
                L.used_off[u] = off;
                off += RG_TILE * L.used_w[u];
            }
    L.tile_bytes = (off + 15) & ~15;
    L.bulk_bytes = off;
    for (int j = 0; j < L.nsrc; j++) {
        L.src[j].oa = L.used_off[L.src[j].ua];
        L.src[j].ob = L.used_off[L.src[j].ub];
        L.src[j].oc = L.used_off[L.src[j].uc];
    }
    for (int k = 0; k < nkeys; k++) {
        L.key_off[k] = L.used_off[L.key_u[k]];
        L.key_w[k] = L.used_w[L.key_u[k]];
    }
    if (L.rf_u >= 0) {  // every comparison with a constant is an interval test on integers
        L.rf_off = L.used_off[L.rf_u];
        L.rf_w = L.used_w[L.rf_u];
        const int64_t v = spec.row_filter_value, mn = INT64_MIN, mx = INT64_MAX;
        L.rf_lo = mn; L.rf_hi = mx; L.rf_neg = 0;
        switch (spec.row_filter_op) {
        case GSQL_CMP_LE: L.rf_hi = v; break;
        case GSQL_CMP_LT: if (v == mn) { L.rf_lo = 1; L.rf_hi = 0; } else L.rf_hi = v - 1; break;
        case GSQL_CMP_GE: L.rf_lo = v; break;
        case GSQL_CMP_GT: if (v == mx) { L.rf_lo = 1; L.rf_hi = 0; } else L.rf_lo = v + 1; break;
        case GSQL_CMP_EQ: L.rf_lo = L.rf_hi = v; break;
        default: L.rf_lo = L.rf_hi = v; L.rf_neg = 1; break;
        }
    }
    return true;
}
