// agg_reg.cuh — REGISTER accumulators for the NULL-free, very-low-cardinality group-by with fp64 sums (the TPC-H Q1
// shape as the plan hands it over: 2 integer keys, SUM / AVG over DOUBLE columns and fused derived expressions,
// COUNT(*), an integer row filter).  Included by agg.cu after agg_lane.cuh.
//
// Why another kernel: k_agg_lane keeps every accumulator in shared memory and decodes a plan per step; r01 measured it
// at 12.3 warp-instructions per ROW (~390 thread instructions) and 21 % issue utilisation: 22.5 ms for the 26.4 GB of
// config 3, 18 % of the HBM roofline.  Here a thread owns RG_G x NSRC fp64 accumulators and RG_G counters in REGISTERS;
// a row costs its loads, one compare per group and NSRC predicated DADDs per group — ~100 thread instructions with
// NSRC x RG_G independent dependency chains — and no shared-memory traffic at all in steady state.  The (at most RG_G)
// group keys live in a per-block dictionary in shared memory that is only written when a new key shows up (the first
// tiles); a ninth key makes the row take the generic path on the spot (find_group_kv + L2 atomics), counted in
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
constexpr int RG_RPT = 2;       // rows per thread per tile: all loads of both rows are in flight together
constexpr int RG_G = 8;         // groups a block can hold in registers (6 when there are 5-6 sums: 128 registers, two blocks per SM, no spills)
constexpr int RG_MAX_SRC = 6;   // distinct fp64 sums

struct RegSrc {
    int32_t kind;  // 0: column a; GSQL_EXPR_MUL_1MINUS: a*(1-b); GSQL_EXPR_MUL_1MINUS_1PLUS: a*(1-b)*(1+c)
    int32_t a, b, c;
};

struct RegPlan {
    int32_t nsrc, nkeys;
    RegSrc src[RG_MAX_SRC];
    int32_t keycol[2];
    int32_t agg_src[GSQL_MAX_AGGS];  // per aggregate: the sum it reports, -1 for COUNT / COUNT(*)
};

__device__ __forceinline__ double reg_f64(const DCol &c, int64_t r) { return __longlong_as_double(ld_stream_8(reinterpret_cast<const long long *>(c.data) + r)); }

__device__ __forceinline__ double reg_src(const AggParams &P, const RegSrc &s, int64_t r) {
    double x = reg_f64(P.in.c[s.a], r);
    if (s.kind != 0) {  // warp-uniform
        x = x * (1.0 - reg_f64(P.in.c[s.b], r));
        if (s.kind == GSQL_EXPR_MUL_1MINUS_1PLUS) x = x * (1.0 + reg_f64(P.in.c[s.c], r));
    }
    return x;
}

__device__ __forceinline__ unsigned long long reg_key(const AggParams &P, const RegPlan &L, int64_t r) {
    unsigned long long k = 0;
    const DCol &c0 = P.in.c[L.keycol[0]];
    if (c0.type == GSQL_T_INT32) k = (unsigned long long)(unsigned int)ld_stream_4(reinterpret_cast<const int *>(c0.data) + r);
    else k = (unsigned long long)ld_stream_8(reinterpret_cast<const long long *>(c0.data) + r);
    if (L.nkeys == 2)  // two INT keys share the word (checked on the host)
        k |= (unsigned long long)(unsigned int)ld_stream_4(reinterpret_cast<const int *>(P.in.c[L.keycol[1]].data) + r) << 32;
    return k;
}

__device__ __forceinline__ void reg_decode_key(const AggParams &P, const RegPlan &L, unsigned long long k, int64_t (&kv)[GSQL_MAX_KEYS], bool (&kn)[GSQL_MAX_KEYS]) {
    for (int c = 0; c < GSQL_MAX_KEYS; c++) { kv[c] = 0; kn[c] = false; }
    if (L.nkeys == 2) {
        kv[0] = (int64_t)(int32_t)(unsigned int)k;
        kv[1] = (int64_t)(int32_t)(unsigned int)(k >> 32);
    } else {
        kv[0] = P.in.c[L.keycol[0]].type == GSQL_T_INT32 ? (int64_t)(int32_t)(unsigned int)k : (int64_t)k;
    }
}

template <int NSRC, int G>
__global__ void __launch_bounds__(RG_THREADS, 2) k_agg_reg(const __grid_constant__ AggParams P, const __grid_constant__ RegPlan L) {
    __shared__ unsigned long long skey[G];
    __shared__ int s_ng, s_elect;
    __shared__ double red[RG_THREADS / 32][G][NSRC];
    __shared__ unsigned long long redc[RG_THREADS / 32][G];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_ng = 0; s_elect = 0x7fffffff; }
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
    constexpr int TILE = RG_THREADS * RG_RPT;
    const int64_t ntiles = (P.rows + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = tile * TILE;
        unsigned long long key[RG_RPT];
        double v[RG_RPT][NSRC];
        bool pass[RG_RPT];
        // ---- every load of the tile is issued before anything is consumed
#pragma unroll
        for (int k = 0; k < RG_RPT; k++) {
            const int64_t i = t0 + k * RG_THREADS + tid;
            pass[k] = i < P.rows;
            const int64_t r = P.row0 + (pass[k] ? i : 0);
            key[k] = reg_key(P, L, r);
#pragma unroll
            for (int j = 0; j < NSRC; j++) v[k][j] = reg_src(P, L.src[j], r);
            if (pass[k] && P.rf_op != GSQL_CMP_NONE) pass[k] = row_passes(P, r);
        }
        // ---- group of each row: position of its key in the block's dictionary
        int gid[RG_RPT];
        while (true) {
            const int ng = s_ng;
            bool miss = false;
#pragma unroll
            for (int k = 0; k < RG_RPT; k++) {
                gid[k] = -1;
#pragma unroll
                for (int g = 0; g < G; g++)
                    if (g < ng && skey[g] == key[k]) gid[k] = g;
                miss |= pass[k] && gid[k] < 0;
            }
            const bool room = ng < G;
            if (!__syncthreads_or(miss && room)) break;  // steady state: one barrier per tile, no dictionary write
            if (miss) atomicMin(&s_elect, tid);
            __syncthreads();
            if (tid == s_elect) {  // one new key per round (first tiles only)
                unsigned long long nk = 0;
#pragma unroll
                for (int k = RG_RPT - 1; k >= 0; k--)
                    if (pass[k] && gid[k] < 0) nk = key[k];
                skey[s_ng] = nk;
                s_ng = s_ng + 1;
                s_elect = 0x7fffffff;
            }
            __syncthreads();
        }
        // ---- accumulate: one compare per group, NSRC predicated adds
        const int ng = s_ng;
#pragma unroll
        for (int k = 0; k < RG_RPT; k++) {
            if (pass[k] && gid[k] < 0) {  // a ninth key: the generic path, right here
                fallback_rows++;
                int64_t kv[GSQL_MAX_KEYS];
                bool kn[GSQL_MAX_KEYS];
                reg_decode_key(P, L, key[k], kv, kn);
                const int64_t r = P.row0 + t0 + k * RG_THREADS + tid;
                const int g2 = find_group_kv(P, kv, kn, digest_of_keys(P, kv, kn));
                if (g2 < 0) {
                    unsigned long long o = atomicAdd(&P.counters[C_OVERFLOW], 1ULL);
                    P.overflow_rows[o] = r;
                } else {
                    for (int a = 0; a < P.naggs; a++) accumulate(P, P.agg[a], g2, r);
                }
                continue;
            }
            const int gk = pass[k] ? gid[k] : -1;
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (g >= 4 && ng <= 4) break;  // block-uniform: dbgen's Q1 has 4 groups
                const bool hit = gk == g;
                cnt[g] += hit ? 1u : 0u;
#pragma unroll
                for (int j = 0; j < NSRC; j++) acc[g][j] += hit ? v[k][j] : 0.0;
            }
        }
    }
    if (fallback_rows) atomicAdd(&P.counters[C_FALLBACK], fallback_rows);
    // ---- merge: reduce over the warp with shuffles, over the block through shared memory, then one thread per group
    const int ng = s_ng;
#pragma unroll
    for (int g = 0; g < G; g++) {
        if (g >= ng) break;
        unsigned int c = cnt[g];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) redc[warp][g] = c;
#pragma unroll
        for (int j = 0; j < NSRC; j++) {
            const double s = warp_sum_f64(acc[g][j]);
            if (lane == 0) red[warp][g][j] = s;
        }
    }
    __syncthreads();
    if (tid < ng) {
        const int g = tid;
        unsigned long long c = 0;
        double s[NSRC];
#pragma unroll
        for (int j = 0; j < NSRC; j++) s[j] = 0.0;
        for (int w = 0; w < RG_THREADS / 32; w++) {
            c += redc[w][g];
#pragma unroll
            for (int j = 0; j < NSRC; j++) s[j] += red[w][g][j];
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
                    if (L.agg_src[a] == j) sv = s[j];
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
    auto has_nulls = [&](int col) { return in.c[col].nulls != nullptr; };
    for (int k = 0; k < nkeys; k++) {
        L.keycol[k] = spec.groups[k];
        if (has_nulls(spec.groups[k])) return false;
    }
    if (spec.row_filter_op != GSQL_CMP_NONE && has_nulls(spec.row_filter_col)) return false;
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
        RegSrc s;
        const int col = c.cols[0];
        if (col < spec.n_input_cols) {
            s.kind = 0; s.a = col; s.b = s.c = 0;
            if (has_nulls(col) || in.c[col].type != GSQL_T_FP64) return false;
        } else {
            const gsql_derived_col &d = spec.derived[col - spec.n_input_cols];
            s.kind = d.kind; s.a = d.a; s.b = d.b; s.c = d.kind == GSQL_EXPR_MUL_1MINUS_1PLUS ? d.c : 0;
            const int ops[3] = {s.a, s.b, s.c};
            for (int q = 0; q < (d.kind == GSQL_EXPR_MUL_1MINUS_1PLUS ? 3 : 2); q++)
                if (has_nulls(ops[q]) || in.c[ops[q]].type != GSQL_T_FP64) return false;
        }
        int at = -1;
        for (int j = 0; j < L.nsrc; j++)
            if (L.src[j].kind == s.kind && L.src[j].a == s.a && L.src[j].b == s.b && L.src[j].c == s.c) at = j;
        if (at < 0) {
            if (L.nsrc == RG_MAX_SRC) return false;
            at = L.nsrc;
            L.src[L.nsrc++] = s;
        }
        L.agg_src[a] = at;
    }
    if (L.nsrc == 0) {  // pure COUNT(*): still needs one (unused) source slot for the template
        return false;
    }
    return true;
}
