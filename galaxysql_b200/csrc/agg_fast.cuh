// agg_fast.cuh — shared-memory privatised group-by for low-cardinality keys (the TPC-H Q1 shape), included by agg.cu.
//
// Every CTA owns a small open-addressing table {key values, NULL flags, accumulators} in shared memory.  A warp
// step takes 32 rows: each lane finds (or inserts) its key's slot in the CTA table, lanes with the same slot are
// grouped with __match_any_sync, the group's values are reduced inside the warp, and ONE lane per distinct key
// updates the shared-memory accumulators — so 600 M rows x 8 accumulators never touch an L2 atomic.  When the CTA
// finishes its rows the <= S partial groups are merged into the global table (find_group_kv + global atomics).
// The path is adaptive: a row whose key does not fit the CTA table (table full / more than 8 probes) goes through the
// generic global path on the spot, and the host stops using this kernel when too many rows do that.
//
// Reference behaviour: AggOpenHashMap.putChunk (EX/operator/util/AggOpenHashMap.java:100-139) — same groups, same
// NULL rules; floating sums are accumulated in a different order (within the north_star's 1e-6 relative tolerance).
#pragma once

namespace {

constexpr int AF_THREADS = 512;
constexpr int AF_MAX_PROBES = 8;
enum { C_FALLBACK = 2 };  // counters[2]: rows that bypassed the CTA table

struct SmemLayout {
    int32_t S;  // slots, power of two
    int32_t off_state;
    int32_t off_kv[GSQL_MAX_KEYS];
    int32_t off_kn[GSQL_MAX_KEYS];
    int32_t off_acc[GSQL_MAX_AGGS];
    int32_t off_cnt[GSQL_MAX_AGGS];  // AVG row count, -1 when unused
    int32_t off_has[GSQL_MAX_AGGS];
    int32_t total;
};

__device__ __forceinline__ int smem_find_or_insert(char *sm, const SmemLayout &L, int nkeys, const int64_t (&kv)[GSQL_MAX_KEYS],
                                                   const bool (&kn)[GSQL_MAX_KEYS], unsigned long long d) {
    if (nkeys == 0) return 0;
    int *state = reinterpret_cast<int *>(sm + L.off_state);
    int s = (int)(gsql_fmix64(d) & (unsigned long long)(L.S - 1));
    int probes = 0;
    while (probes < AF_MAX_PROBES) {
        int st = *reinterpret_cast<volatile int *>(&state[s]);
        if (st == 2) {
            bool eq = true;
            for (int c = 0; c < nkeys && eq; c++) {
                bool n = *reinterpret_cast<volatile uint8_t *>(sm + L.off_kn[c] + s) != 0;
                long long v = *reinterpret_cast<volatile long long *>(sm + L.off_kv[c] + (size_t)s * 8);
                if (n != kn[c] || (!n && v != kv[c])) eq = false;
            }
            if (eq) return s;
            s = (s + 1) & (L.S - 1);
            probes++;
        } else if (st == 0) {
            if (atomicCAS(&state[s], 0, 1) == 0) {
                for (int c = 0; c < nkeys; c++) {
                    *reinterpret_cast<long long *>(sm + L.off_kv[c] + (size_t)s * 8) = kv[c];
                    *reinterpret_cast<uint8_t *>(sm + L.off_kn[c] + s) = kn[c] ? 1 : 0;
                }
                __threadfence_block();
                *reinterpret_cast<volatile int *>(&state[s]) = 2;
                return s;
            }
        } else {
            __nanosleep(10);  // another lane is publishing this slot
        }
    }
    return -1;
}

// Sum of `v` over the lanes of `peers` (every lane of the warp calls this; each gets the sum of ITS peer group).
__device__ __forceinline__ double peer_sum_f64(double v, unsigned peers) {
    double s = 0.0;
    unsigned rem = peers;
    while (__any_sync(0xffffffffu, rem != 0)) {
        int src = rem ? __ffs(rem) - 1 : 0;
        double x = __shfl_sync(0xffffffffu, v, src);
        if (rem) {
            s += x;
            rem &= rem - 1;
        }
    }
    return s;
}
__device__ __forceinline__ long long peer_sum_i64(long long v, unsigned peers) {
    long long s = 0;
    unsigned rem = peers;
    while (__any_sync(0xffffffffu, rem != 0)) {
        int src = rem ? __ffs(rem) - 1 : 0;
        long long x = __shfl_sync(0xffffffffu, v, src);
        if (rem) {
            s = (long long)((unsigned long long)s + (unsigned long long)x);
            rem &= rem - 1;
        }
    }
    return s;
}
__device__ __forceinline__ long long peer_minmax_i64(long long v, unsigned peers, bool mx) {
    long long s = mx ? (long long)0x8000000000000000ULL : 0x7fffffffffffffffLL;
    unsigned rem = peers;
    while (__any_sync(0xffffffffu, rem != 0)) {
        int src = rem ? __ffs(rem) - 1 : 0;
        long long x = __shfl_sync(0xffffffffu, v, src);
        if (rem) {
            s = mx ? (x > s ? x : s) : (x < s ? x : s);
            rem &= rem - 1;
        }
    }
    return s;
}

__global__ void __launch_bounds__(AF_THREADS, 2) k_agg_smem(const __grid_constant__ AggParams P, const __grid_constant__ SmemLayout L) {
    extern __shared__ __align__(16) char sm[];
    // ---- init the CTA table
    for (int i = threadIdx.x; i < L.S; i += AF_THREADS) {
        reinterpret_cast<int *>(sm + L.off_state)[i] = P.nkeys == 0 ? 2 : 0;
        for (int a = 0; a < P.naggs; a++) {
            long long init = P.agg[a].kind == GSQL_AGG_MIN ? 0x7fffffffffffffffLL : P.agg[a].kind == GSQL_AGG_MAX ? (long long)0x8000000000000000ULL : 0;
            reinterpret_cast<long long *>(sm + L.off_acc[a])[i] = init;
            if (L.off_cnt[a] >= 0) reinterpret_cast<long long *>(sm + L.off_cnt[a])[i] = 0;
            reinterpret_cast<uint8_t *>(sm + L.off_has[a])[i] = 0;
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (AF_THREADS / 32);
    const int64_t warp_id = (int64_t)blockIdx.x * (AF_THREADS / 32) + (threadIdx.x >> 5);
    unsigned int fallback_rows = 0;
    for (int64_t b = warp_id * 32; b < P.rows; b += warps_total * 32) {
        const int64_t i = b + lane;
        const int64_t r = P.row0 + i;
        const bool live = i < P.rows && row_passes(P, r);
        int64_t kv[GSQL_MAX_KEYS];
        bool kn[GSQL_MAX_KEYS];
        unsigned long long d = 0;
        int slot = -2;
        if (live) {
            if (P.nkeys > 0) d = load_group_key(P, r, kv, kn);
            slot = smem_find_or_insert(sm, L, P.nkeys, kv, kn, d);
            if (slot == -1) {  // does not fit the CTA table: the generic path, right here
                fallback_rows++;
                int gid = find_group_kv(P, kv, kn, d);
                if (gid < 0) {
                    unsigned long long o = atomicAdd(&P.counters[C_OVERFLOW], 1ULL);
                    P.overflow_rows[o] = r;
                } else {
                    for (int a = 0; a < P.naggs; a++) accumulate(P, P.agg[a], gid, r);
                }
            }
        }
        const bool useful = slot >= 0;
        const unsigned peers = __match_any_sync(0xffffffffu, slot);
        const bool leader = useful && lane == __ffs(peers) - 1;
#pragma unroll 1
        for (int a = 0; a < P.naggs; a++) {
            const AggDev &ag = P.agg[a];
            bool ok = useful;
            if (ok && ag.filter_col >= 0) {
                const DCol &f = P.in.c[ag.filter_col];
                if (f.type == GSQL_T_INT64 && !in_null(f, r) && reinterpret_cast<const int64_t *>(f.data)[r] < 1) ok = false;
            }
            if (ok) {
                for (int q = 0; q < ag.ncols; q++)
                    if (val_null(P, ag.cols[q], r)) ok = false;  // COUNT: any NULL arg; others: the single argument
            }
            const unsigned okmask = __ballot_sync(0xffffffffu, ok);
            const unsigned long long cnt = __popc(peers & okmask);
            long long *acc = reinterpret_cast<long long *>(sm + L.off_acc[a]);
            switch (ag.kind) {
            case GSQL_AGG_COUNT_STAR:
            case GSQL_AGG_COUNT:
                if (leader && cnt) atomicAdd(reinterpret_cast<unsigned long long *>(&acc[slot]), cnt);
                break;
            case GSQL_AGG_SUM:
            case GSQL_AGG_AVG: {
                double v = ok ? val_f64(P, ag.cols[0], r) : 0.0;
                double s = peer_sum_f64(v, peers);
                if (leader && cnt) {
                    atomicAdd(reinterpret_cast<double *>(&acc[slot]), s);
                    if (L.off_cnt[a] >= 0) atomicAdd(reinterpret_cast<unsigned long long *>(sm + L.off_cnt[a]) + slot, cnt);
                    reinterpret_cast<uint8_t *>(sm + L.off_has[a])[slot] = 1;
                }
                break;
            }
            case GSQL_AGG_SUM0: {
                long long v = ok ? val_i64(P, ag.cols[0], r) : 0;
                long long s = peer_sum_i64(v, peers);
                if (leader && cnt) atomicAdd(reinterpret_cast<unsigned long long *>(&acc[slot]), (unsigned long long)s);
                break;
            }
            default: {  // MIN / MAX on the order-preserving int64 image
                const bool mx = ag.kind == GSQL_AGG_MAX;
                long long ident = mx ? (long long)0x8000000000000000ULL : 0x7fffffffffffffffLL;
                long long v = ident;
                if (ok) v = ag.in_type == GSQL_T_FP64 ? dbl_sortable(val_f64(P, ag.cols[0], r), mx) : val_i64(P, ag.cols[0], r);
                long long s = peer_minmax_i64(v, peers, mx);
                if (leader && cnt) {
                    if (mx) atomicMax(&acc[slot], s);
                    else atomicMin(&acc[slot], s);
                    reinterpret_cast<uint8_t *>(sm + L.off_has[a])[slot] = 1;
                }
            }
            }
        }
    }
    if (fallback_rows) atomicAdd(&P.counters[C_FALLBACK], (unsigned long long)fallback_rows);
    __syncthreads();
    // ---- merge the CTA's partial groups into the global table
    for (int s = threadIdx.x; s < L.S; s += AF_THREADS) {
        if (reinterpret_cast<int *>(sm + L.off_state)[s] != 2) continue;
        int64_t kv[GSQL_MAX_KEYS];
        bool kn[GSQL_MAX_KEYS];
        for (int c = 0; c < P.nkeys; c++) {
            kv[c] = *reinterpret_cast<long long *>(sm + L.off_kv[c] + (size_t)s * 8);
            kn[c] = *reinterpret_cast<uint8_t *>(sm + L.off_kn[c] + s) != 0;
        }
        // a partial group without any contribution (all its rows filtered per aggregate) still has to exist as a group
        // the merge may exceed gcap by at most CTAs x S groups: covered by the arrays' slack (ignore_cap)
        int gid = P.nkeys == 0 ? 0 : find_group_kv(P, kv, kn, digest_of_keys(P, kv, kn), true);
        for (int a = 0; a < P.naggs; a++) {
            const AggDev &ag = P.agg[a];
            long long v = reinterpret_cast<long long *>(sm + L.off_acc[a])[s];
            bool has = reinterpret_cast<uint8_t *>(sm + L.off_has[a])[s] != 0;
            switch (ag.kind) {
            case GSQL_AGG_COUNT_STAR: case GSQL_AGG_COUNT: case GSQL_AGG_SUM0:
                if (v) atomicAdd(reinterpret_cast<unsigned long long *>(&ag.l[gid]), (unsigned long long)v);
                break;
            case GSQL_AGG_SUM:
                if (has) { atomicAdd(&ag.d[gid], __longlong_as_double(v)); ag.has[gid] = 1; }
                break;
            case GSQL_AGG_AVG:
                if (has) {
                    atomicAdd(&ag.d[gid], __longlong_as_double(v));
                    atomicAdd(reinterpret_cast<unsigned long long *>(&ag.l[gid]), (unsigned long long)reinterpret_cast<long long *>(sm + L.off_cnt[a])[s]);
                    ag.has[gid] = 1;
                }
                break;
            default:
                if (has) {
                    if (ag.kind == GSQL_AGG_MAX) atomicMax(reinterpret_cast<long long *>(&ag.l[gid]), v);
                    else atomicMin(reinterpret_cast<long long *>(&ag.l[gid]), v);
                    ag.has[gid] = 1;
                }
            }
        }
    }
}

}  // namespace

struct AggFast {
    bool eligible = false;  // shape supported by the shared-memory kernel
    bool enabled = false;   // still profitable (few rows bypass the CTA tables)
    SmemLayout L;
    int64_t rows_seen = 0, rows_fallback = 0;
};

// Decides eligibility and the shared-memory layout (host).
static void agg_fast_plan(AggFast *F, int nkeys, int naggs, const gsql_agg_call *aggs, const int32_t *in_type) {
    F->eligible = false;
    for (int a = 0; a < naggs; a++)
        if (aggs[a].kind == GSQL_AGG_SUM && in_type[a] != GSQL_T_FP64) return;  // exact 128-bit SUM(int) stays generic
    int per_slot = 4 + nkeys * 9;
    for (int a = 0; a < naggs; a++) per_slot += 8 + 1 + (aggs[a].kind == GSQL_AGG_AVG ? 8 : 0);
    int S = 1;
    if (nkeys > 0) {
        S = 1024;
        while (S > 16 && (size_t)S * per_slot > 40 * 1024) S >>= 1;
    }
    SmemLayout &L = F->L;
    memset(&L, 0, sizeof(L));
    L.S = S;
    int off = 0;
    for (int a = 0; a < naggs; a++) { L.off_acc[a] = off; off += S * 8; }
    for (int a = 0; a < naggs; a++) {
        if (aggs[a].kind == GSQL_AGG_AVG) { L.off_cnt[a] = off; off += S * 8; }
        else L.off_cnt[a] = -1;
    }
    for (int c = 0; c < nkeys; c++) { L.off_kv[c] = off; off += S * 8; }
    L.off_state = off;
    off += S * 4;
    for (int c = 0; c < nkeys; c++) { L.off_kn[c] = off; off += S; }
    for (int a = 0; a < naggs; a++) { L.off_has[a] = off; off += S; }
    L.total = (off + 15) & ~15;
    F->eligible = true;
    F->enabled = true;
}
