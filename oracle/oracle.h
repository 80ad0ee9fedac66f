/*
 * oracle.h — CPU restatement of GalaxySQL's MPP vectorised operator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under galaxysql_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker / the timed CPU baseline.
 *
 * The reference is Java (polardbx-executor); no JDK exists in the build container, so the
 * reference itself cannot be executed here.  Parity is pinned by the reference's own
 * known-answer tests (HashJoinTest, HashAggExecTest, SpilledHashAggExecTest,
 * ChunkRowOpenHashMapTest) ported to tests/golden/ — see tests/test_oracle_golden.py.
 * At the hash / partition-id level the reference holds no golden values (fastutil's
 * HashCommon is an un-vendored transitive dependency, version not pinned in any pom):
 * "parity unpinned" for raw mix()/murmurHash3() outputs; they are restated from fastutil's
 * published source and only cross-checked oracle <-> GPU.
 *
 * Path aliases in citations:  EX/ = polardbx-executor/src/main/java/com/alibaba/polardbx/executor/
 */
#ifndef GSQL_ORACLE_H
#define GSQL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_T_INT32 = 0, ORC_T_INT64 = 1, ORC_T_FP64 = 2, ORC_T_DEC128 = 3 /* output only: int128 LE, scale 0 */ };
enum { ORC_JOIN_INNER = 0, ORC_JOIN_LEFT = 1, ORC_JOIN_RIGHT = 2, ORC_JOIN_SEMI = 3, ORC_JOIN_ANTI = 4 };
enum {
    ORC_AGG_COUNT_STAR = 0, /* CountRow */
    ORC_AGG_COUNT = 1,      /* Count (all listed cols non-NULL) */
    ORC_AGG_SUM = 2,        /* Double2DoubleSum | Int/Long2DecimalSum (by input type) */
    ORC_AGG_AVG = 3,        /* Double2DoubleAvg (SpecificType2DoubleAvgV2) */
    ORC_AGG_MIN = 4, ORC_AGG_MAX = 5,
    ORC_AGG_SUM0 = 6        /* Long2LongSum0 */
};
enum { ORC_OK = 0, ORC_ERR_MORE_THAN_ONE_ROW = 1, ORC_ERR_UNSUPPORTED = 2 };

#define ORC_MAX_KEYS 8

typedef struct orc_col {
    int32_t type;
    int32_t _pad;
    const void *data;
    const uint8_t *nulls; /* NULL => no nulls; else 1 byte per row like Java boolean[] */
} orc_col;

typedef struct orc_join_spec {
    int32_t join_type;
    int32_t max_one_row;  /* singleJoin */
    int32_t build_outer;  /* buildOuterInput */
    int32_t nkeys;
    int32_t outer_key[ORC_MAX_KEYS];
    int32_t inner_key[ORC_MAX_KEYS];
    int32_t key_type[ORC_MAX_KEYS]; /* unified type */
    int32_t n_anti_operands;        /* 0 => antiJoinOperands == null */
    int32_t anti_operands[ORC_MAX_KEYS]; /* outer column indices (InputRefExpression) */
    /* otherCondition, restricted form: conjunction of (joinrow col != int const).  n_cond = 0 => null */
    int32_t n_cond;
    int32_t cond_col[4];
    int64_t cond_ne_value[4];
} orc_join_spec;

typedef struct orc_agg_call {
    int32_t kind;
    int32_t ncols;
    int32_t cols[4];
    int32_t filter_arg; /* -1 none */
} orc_agg_call;

typedef struct orc_result orc_result;

/* ---- scalar restatements (exported so tests can pin them) ---- */
int32_t orc_mix(int32_t x);
int32_t orc_murmur_hash3(int32_t x);
int32_t orc_array_size(int32_t expected, float f);
int32_t orc_max_fill(int32_t n, float f);
int32_t orc_partition(int32_t hash, int32_t nparts);

/* ---- vector forms ---- */
void orc_hash_rows(const orc_col *keycols, int32_t nkeys, const int32_t *unified_types, int64_t rows, int32_t *out);
void orc_partition_ids(const int32_t *hash, int64_t rows, int32_t nparts, int32_t *out);

/* ---- operators (single-threaded, faithful order) ---- */
int orc_hash_join(const orc_join_spec *spec, const orc_col *outer, int32_t n_outer, int64_t outer_rows,
                  const orc_col *inner, int32_t n_inner, int64_t inner_rows, orc_result **out);
int orc_hash_agg(const orc_col *in, int32_t ncols, int64_t rows, const int32_t *groups, int32_t ngroups,
                 const orc_agg_call *aggs, int32_t naggs, int32_t expected_groups, int32_t chunk_size,
                 orc_result **out);
/* Local / remote hash-partition exchange: rows routed to nparts consumers, stable order inside a partition. */
int orc_partition_exchange(const orc_col *in, int32_t ncols, int64_t rows, const int32_t *channels, int32_t nchannels,
                           const int32_t *key_types, int32_t nparts, orc_result **out, int64_t *part_counts);

/* ChunkRowOpenHashMap put/get (EX/operator/util/ChunkRowOpenHashMap.java:96-166) — pins mix+probing via
 * ChunkRowOpenHashMapTest.java:33-69 */
void orc_chunk_row_open_hash_map(const orc_col *build, int32_t ncols, int32_t build_rows, const orc_col *probe,
                                 int32_t probe_rows, int32_t *put_result, int32_t *get_result);

/* ---- results ---- */
int64_t orc_result_rows(const orc_result *r);
int32_t orc_result_ncols(const orc_result *r);
int32_t orc_result_col(const orc_result *r, int32_t i, const void **data, const uint8_t **nulls);
void orc_result_free(orc_result *r);

/* ---- reference-shaped multi-threaded CPU baseline (bench.py only) ----
 * P driver threads, 1000-row chunks.  Returns seconds via out params. */
int orc_mt_join(const orc_join_spec *spec, const orc_col *outer, int32_t n_outer, int64_t outer_rows,
                const orc_col *inner, int32_t n_inner, int64_t inner_rows, int32_t nthreads, int32_t chunk_rows,
                double *build_seconds, double *probe_seconds, int64_t *out_rows, uint64_t *checksum);
int orc_mt_hash_agg(const orc_col *in, int32_t ncols, int64_t rows, const int32_t *groups, int32_t ngroups,
                    const orc_agg_call *aggs, int32_t naggs, int32_t expected_groups, int32_t nthreads,
                    int32_t chunk_rows, double *seconds, int64_t *out_groups);

#ifdef __cplusplus
}
#endif
#endif
