/*
 * oracle.c — CPU restatement of GalaxySQL's MPP vectorised operator hot path (see oracle.h header note).
 * TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Every function cites the reference lines it follows.  Aliases:
 *   EX/  = polardbx-executor/src/main/java/com/alibaba/polardbx/executor/
 *   OPT/ = polardbx-optimizer/src/main/java/com/alibaba/polardbx/optimizer/
 * fastutil HashCommon (mix, murmurHash3, arraySize, maxFill, nextPowerOfTwo) is NOT in /root/reference
 * (transitive dependency, version unpinned in the poms); restated from its published source.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------
 * fastutil HashCommon (it.unimi.dsi.fastutil.HashCommon) — call sites:
 *   EX/operator/util/ConcurrentRawHashTable.java:55,93,114 ; GroupOpenHashMap.java:90,92,143 ;
 *   EX/utils/ExecUtils.java:1026,1029
 * ---------------------------------------------------------------------------------------------- */
int32_t orc_mix(int32_t x) {
    uint32_t h = (uint32_t)x * 0x9E3779B9u; /* INT_PHI */
    return (int32_t)(h ^ (h >> 16));
}

int32_t orc_murmur_hash3(int32_t xi) {
    uint32_t x = (uint32_t)xi;
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return (int32_t)x;
}

static int64_t next_pow2_i64(int64_t x) {
    if (x == 0) return 1;
    x--;
    x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
    return (x | x >> 32) + 1;
}

/* HashCommon.arraySize(expected, f) = max(2, nextPowerOfTwo((long)ceil(expected / f))) */
int32_t orc_array_size(int32_t expected, float f) {
    int64_t s = next_pow2_i64((int64_t)ceil((double)((float)expected / f)));
    if (s < 2) s = 2;
    return (int32_t)s;
}

/* HashCommon.maxFill(n, f) = min((int)ceil(n * f), n - 1) */
int32_t orc_max_fill(int32_t n, float f) {
    int32_t m = (int32_t)ceil((double)((float)n * f));
    return m < n - 1 ? m : n - 1;
}

/* EX/utils/ExecUtils.java:1019-1031 */
int32_t orc_partition(int32_t hash, int32_t nparts) {
    int is_pow2 = (nparts & -nparts) == nparts;
    if (is_pow2) {
        return orc_murmur_hash3(hash) & (nparts - 1);
    }
    return (orc_murmur_hash3(hash) & 0x7fffffff) % nparts;
}

void orc_partition_ids(const int32_t *hash, int64_t rows, int32_t nparts, int32_t *out) {
    for (int64_t i = 0; i < rows; i++) out[i] = orc_partition(hash[i], nparts);
}

/* ------------------------------------------------------------------------------------------------
 * Block accessors.  EX/chunk/AbstractBlock.java:27-47 (isNull != null && isNull[pos])
 * ---------------------------------------------------------------------------------------------- */
static inline int col_is_null(const orc_col *c, int64_t r) { return c->nulls != NULL && c->nulls[r] != 0; }
static inline int64_t col_i64(const orc_col *c, int64_t r) {
    switch (c->type) {
    case ORC_T_INT32: return ((const int32_t *)c->data)[r];
    case ORC_T_INT64: return ((const int64_t *)c->data)[r];
    default: return (int64_t)((const double *)c->data)[r];
    }
}
static inline double col_f64(const orc_col *c, int64_t r) {
    switch (c->type) {
    case ORC_T_INT32: return (double)((const int32_t *)c->data)[r];
    case ORC_T_INT64: return (double)((const int64_t *)c->data)[r];
    default: return ((const double *)c->data)[r];
    }
}
static inline int64_t f64_bits(double d) {
    int64_t b;
    if (d != d) return 0x7ff8000000000000LL; /* Double.doubleToLongBits canonicalises NaN */
    memcpy(&b, &d, 8);
    return b;
}
static int type_width(int t) { return t == ORC_T_INT32 ? 4 : t == ORC_T_DEC128 ? 16 : 8; }

/* A key value converted to the join's / exchange's unified type (EX/chunk/Converters.java:34-46,94-131:
 * identity when types are equal, else per-row DataType.convertFrom — integer widening / (double) cast). */
typedef struct { int is_null; int64_t i; double d; } keyval;

static inline keyval key_at(const orc_col *c, int64_t r, int unified) {
    keyval k;
    k.is_null = col_is_null(c, r);
    k.i = 0; k.d = 0;
    if (k.is_null) return k;
    if (unified == ORC_T_FP64) k.d = col_f64(c, r);
    else k.i = col_i64(c, r);
    return k;
}

/* Block.hashCode(position): NULL -> 0 (EX/chunk/Block.java:113-118);
 * IntegerBlock.java:112-117 (value); LongBlock.java:110-115 (Long.hashCode = (int)(v ^ v>>>32));
 * DoubleBlock.java:111-116 (Double.hashCode = Long.hashCode(doubleToLongBits)). */
static inline int32_t key_hash(const keyval *k, int unified) {
    if (k->is_null) return 0;
    if (unified == ORC_T_INT32) return (int32_t)k->i;
    uint64_t v = unified == ORC_T_FP64 ? (uint64_t)f64_bits(k->d) : (uint64_t)k->i;
    return (int32_t)(uint32_t)(v ^ (v >> 32));
}

/* Block.equals: NULL==NULL true, NULL vs value false, else == (IntegerBlock.java:120-134,
 * LongBlock.java:69-85, DoubleBlock.java:77-91 — Java double ==, so -0.0 == 0.0 and NaN != NaN). */
static inline int key_equal(const keyval *a, const keyval *b, int unified) {
    if (a->is_null && b->is_null) return 1;
    if (a->is_null != b->is_null) return 0;
    if (unified == ORC_T_FP64) return a->d == b->d;
    return a->i == b->i;
}

/* Chunk.hashCode(pos) / hashCodeVector: h = h*31 + block.hashCode  (EX/chunk/Chunk.java:116-130) */
static inline int32_t row_hash(const orc_col *keycols, int32_t nkeys, const int32_t *types, int64_t r) {
    uint32_t h = 0;
    for (int32_t c = 0; c < nkeys; c++) {
        keyval k = key_at(&keycols[c], r, types[c]);
        h = h * 31u + (uint32_t)key_hash(&k, types[c]);
    }
    return (int32_t)h;
}

void orc_hash_rows(const orc_col *keycols, int32_t nkeys, const int32_t *types, int64_t rows, int32_t *out) {
    for (int64_t r = 0; r < rows; r++) out[r] = row_hash(keycols, nkeys, types, r);
}

static inline int rows_key_equal(const orc_col *a, int64_t ra, const orc_col *b, int64_t rb, int32_t nkeys,
                                 const int32_t *types) {
    for (int32_t c = 0; c < nkeys; c++) {
        keyval ka = key_at(&a[c], ra, types[c]);
        keyval kb = key_at(&b[c], rb, types[c]);
        if (!key_equal(&ka, &kb, types[c])) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Result container (stands in for the List<Chunk> an Executor emits; row order = emission order)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int32_t type; char *data; uint8_t *nulls; } res_col;
struct orc_result { int64_t rows; int64_t cap; int32_t ncols; res_col *cols; };

static orc_result *res_new(int32_t ncols, const int32_t *types, int64_t cap) {
    orc_result *r = (orc_result *)calloc(1, sizeof(*r));
    if (cap < 16) cap = 16;
    r->ncols = ncols; r->cap = cap;
    r->cols = (res_col *)calloc((size_t)(ncols > 0 ? ncols : 1), sizeof(res_col));
    for (int32_t i = 0; i < ncols; i++) {
        r->cols[i].type = types[i];
        r->cols[i].data = (char *)malloc((size_t)cap * type_width(types[i]));
        r->cols[i].nulls = (uint8_t *)calloc((size_t)cap, 1);
    }
    return r;
}
static void res_reserve_row(orc_result *r) {
    if (r->rows < r->cap) return;
    int64_t nc = r->cap * 2;
    for (int32_t i = 0; i < r->ncols; i++) {
        r->cols[i].data = (char *)realloc(r->cols[i].data, (size_t)nc * type_width(r->cols[i].type));
        r->cols[i].nulls = (uint8_t *)realloc(r->cols[i].nulls, (size_t)nc);
        memset(r->cols[i].nulls + r->cap, 0, (size_t)(nc - r->cap));
    }
    r->cap = nc;
}
/* Block.writePositionTo(position, builder) — copies value or appendNull */
static inline void res_put_from(orc_result *r, int32_t oc, const orc_col *src, int64_t sr) {
    res_col *c = &r->cols[oc];
    int w = type_width(c->type);
    if (col_is_null(src, sr)) {
        c->nulls[r->rows] = 1;
        memset(c->data + r->rows * w, 0, (size_t)w);
    } else {
        c->nulls[r->rows] = 0;
        memcpy(c->data + r->rows * w, (const char *)src->data + sr * w, (size_t)w);
    }
}
static inline void res_put_null(orc_result *r, int32_t oc) {
    res_col *c = &r->cols[oc];
    int w = type_width(c->type);
    c->nulls[r->rows] = 1;
    memset(c->data + r->rows * w, 0, (size_t)w);
}
static inline void res_put_raw(orc_result *r, int32_t oc, const void *v) {
    res_col *c = &r->cols[oc];
    int w = type_width(c->type);
    c->nulls[r->rows] = 0;
    memcpy(c->data + r->rows * w, v, (size_t)w);
}
int64_t orc_result_rows(const orc_result *r) { return r->rows; }
int32_t orc_result_ncols(const orc_result *r) { return r->ncols; }
int32_t orc_result_col(const orc_result *r, int32_t i, const void **data, const uint8_t **nulls) {
    *data = r->cols[i].data; *nulls = r->cols[i].nulls;
    return r->cols[i].type;
}
void orc_result_free(orc_result *r) {
    if (!r) return;
    for (int32_t i = 0; i < r->ncols; i++) { free(r->cols[i].data); free(r->cols[i].nulls); }
    free(r->cols); free(r);
}

/* ------------------------------------------------------------------------------------------------
 * Hash join.  Build: ExecUtils.buildOneChunk (EX/utils/ExecUtils.java:914-944),
 * ConcurrentRawHashTable (EX/operator/util/ConcurrentRawHashTable.java:50-116),
 * Synchronizer.initHashTable/buildHashTable (EX/operator/ParallelHashJoinExec.java:388-426).
 * Probe: AbstractBufferedJoinExec.nextRows (EX/operator/AbstractBufferedJoinExec.java:185-264),
 * AbstractHashJoinExec.matchInit/matchNext (AbstractHashJoinExec.java:80-106),
 * row builders AbstractJoinExec.java:174-227 and ParallelHashJoinExec.java:168-201,233-271.
 * ---------------------------------------------------------------------------------------------- */
static float select_load_factor(int32_t size) { /* ConcurrentRawHashTable.java:67-75 */
    if (size >= 100000000) return 0.75f;
    if (size >= 10000000) return 0.5f;
    return 0.25f;
}

typedef struct {
    int32_t n, mask;
    int32_t *keys;  /* bucket heads, NOT_EXISTS = -1 */
    int32_t *links; /* positionLinks */
} raw_table;

static void raw_table_init(raw_table *t, int32_t size) {
    t->n = orc_array_size(size, select_load_factor(size));
    t->mask = t->n - 1;
    t->keys = (int32_t *)malloc((size_t)t->n * 4);
    memset(t->keys, 0xff, (size_t)t->n * 4);
    t->links = (int32_t *)malloc((size_t)(size > 0 ? size : 1) * 4);
    memset(t->links, 0xff, (size_t)(size > 0 ? size : 1) * 4);
}
static void raw_table_free(raw_table *t) { free(t->keys); free(t->links); }

static int cols_contain_null(const orc_col *cols, int32_t n, int64_t rows) { /* AbstractJoinExec.java:136-156 */
    for (int32_t j = 0; j < n; j++)
        for (int64_t k = 0; k < rows; k++)
            if (col_is_null(&cols[j], k)) return 1;
    return 0;
}

/* restricted otherCondition: AND_i (joinrow[col_i] is NULL or != value_i); AbstractJoinExec.java:227-250 */
static int eval_cond(const orc_join_spec *s, const orc_col *left, int32_t nleft, int64_t lrow, const orc_col *right,
                     int64_t rrow) {
    for (int32_t i = 0; i < s->n_cond; i++) {
        int32_t c = s->cond_col[i];
        const orc_col *col = c < nleft ? &left[c] : &right[c - nleft];
        int64_t row = c < nleft ? lrow : rrow;
        if (col_is_null(col, row)) continue;
        if (col_i64(col, row) == s->cond_ne_value[i]) return 0;
    }
    return 1;
}

int orc_hash_join(const orc_join_spec *spec, const orc_col *outer, int32_t n_outer, int64_t outer_rows,
                  const orc_col *inner, int32_t n_inner, int64_t inner_rows, orc_result **out) {
    const int jt = spec->join_type;
    const int outer_join = jt == ORC_JOIN_LEFT || jt == ORC_JOIN_RIGHT;           /* AbstractJoinExec.java:83 */
    const int semi_join = (jt == ORC_JOIN_SEMI || jt == ORC_JOIN_ANTI) && !spec->max_one_row; /* :84 */
    const int single_join = spec->max_one_row;
    const int build_outer = spec->build_outer;
    if ((jt == ORC_JOIN_SEMI || jt == ORC_JOIN_ANTI) && single_join) return ORC_ERR_UNSUPPORTED;
    if (build_outer && (semi_join || spec->n_cond > 0)) return ORC_ERR_UNSUPPORTED;
    if (inner_rows > 0x7fffffff || outer_rows > 0x7fffffff) return ORC_ERR_UNSUPPORTED;

    const orc_col *build = build_outer ? outer : inner;
    const orc_col *probe = build_outer ? inner : outer;
    const int32_t n_build = build_outer ? n_outer : n_inner;
    const int32_t n_probe = build_outer ? n_inner : n_outer;
    const int64_t build_rows = build_outer ? outer_rows : inner_rows;
    const int64_t probe_rows = build_outer ? inner_rows : outer_rows;
    const int32_t *bk = build_outer ? spec->outer_key : spec->inner_key;
    const int32_t *pk = build_outer ? spec->inner_key : spec->outer_key;
    const int32_t nkeys = spec->nkeys;

    orc_col bkey[ORC_MAX_KEYS], pkey[ORC_MAX_KEYS];
    for (int32_t i = 0; i < nkeys; i++) { bkey[i] = build[bk[i]]; pkey[i] = probe[pk[i]]; }

    /* output schema — AbstractJoinExec.java:103-120 */
    int32_t otypes[64]; int32_t nout = 0;
    const int32_t right_cols = single_join ? 1 : n_inner;
    if (semi_join) {
        for (int32_t i = 0; i < n_outer; i++) otypes[nout++] = outer[i].type;
    } else if (single_join) {
        for (int32_t i = 0; i < n_outer; i++) otypes[nout++] = outer[i].type;
        otypes[nout++] = inner[0].type;
    } else if (jt == ORC_JOIN_RIGHT) {
        for (int32_t i = 0; i < n_inner; i++) otypes[nout++] = inner[i].type;
        for (int32_t i = 0; i < n_outer; i++) otypes[nout++] = outer[i].type;
    } else {
        for (int32_t i = 0; i < n_outer; i++) otypes[nout++] = outer[i].type;
        for (int32_t i = 0; i < n_inner; i++) otypes[nout++] = inner[i].type;
    }
    orc_result *res = res_new(nout, otypes, probe_rows + 16);
    *out = res;

    /* pass-through / pass-nothing — ParallelHashJoinExec.buildConsume:107-128,
     * AbstractBufferedJoinExec.doSpecialCheckForSemiJoin:290-310 */
    int pass_nothing = 0, pass_through = 0;
    if (build_rows == 0 && jt == ORC_JOIN_INNER) pass_nothing = 1;
    if (semi_join) {
        if (build_rows == 0) {
            if (jt == ORC_JOIN_SEMI) pass_nothing = 1; else pass_through = 1;
        } else if (jt == ORC_JOIN_ANTI && spec->n_anti_operands > 0 && n_build == 1) {
            if (cols_contain_null(build, n_build, build_rows)) pass_nothing = 1;
        }
    }
    if (pass_through) {
        for (int64_t r = 0; r < probe_rows; r++) {
            res_reserve_row(res);
            for (int32_t i = 0; i < n_probe; i++) res_put_from(res, i, &probe[i], r);
            res->rows++;
        }
        return ORC_OK;
    }
    if (pass_nothing) return ORC_OK;

    /* ---- build ---- */
    raw_table t;
    raw_table_init(&t, (int32_t)build_rows);
    for (int64_t p = 0; p < build_rows; p++) {
        int has_null = 0; /* ExecUtils.checkJoinKeysNotNull:955-962 */
        for (int32_t c = 0; c < nkeys; c++) if (col_is_null(&bkey[c], p)) { has_null = 1; break; }
        if (has_null) continue;
        int32_t h = row_hash(bkey, nkeys, spec->key_type, p);
        int32_t s = orc_mix(h) & t.mask;
        t.links[p] = t.keys[s]; /* put(): getAndSet returns previous head or NOT_EXISTS */
        t.keys[s] = (int32_t)p;
    }
    uint8_t *used = NULL; /* joinNullRowBitSet — ParallelHashJoinExec.java:437-465 */
    if (build_outer) used = (uint8_t *)calloc((size_t)(build_rows > 0 ? build_rows : 1), 1);

    int rc = ORC_OK;
    /* ---- probe ---- */
    for (int64_t r = 0; r < probe_rows && rc == ORC_OK; r++) {
        int matched = 0;
        int32_t h = row_hash(pkey, nkeys, spec->key_type, r);
        int32_t m = t.keys[orc_mix(h) & t.mask];
        while (m != -1 && !rows_key_equal(bkey, m, pkey, r, nkeys, spec->key_type)) m = t.links[m];
        for (; m != -1;) {
            /* join row sides for the condition: leftSide/rightSide(outerRow, innerRow) */
            int cond_ok = 1;
            if (spec->n_cond > 0) {
                /* not build_outer here: outer = probe, inner = build */
                if (jt == ORC_JOIN_RIGHT) cond_ok = eval_cond(spec, inner, n_inner, m, outer, r);
                else cond_ok = eval_cond(spec, outer, n_outer, r, inner, m);
            }
            if (cond_ok) {
                if (jt == ORC_JOIN_INNER || jt == ORC_JOIN_LEFT) {
                    res_reserve_row(res);
                    int32_t col = 0;
                    if (build_outer) { /* ParallelHashJoinExec.buildJoinRow:233-252 */
                        used[m] = 1;
                        for (int32_t i = 0; i < n_outer; i++) res_put_from(res, col++, &build[i], m);
                        for (int32_t i = 0; i < right_cols; i++) res_put_from(res, col++, &probe[i], r);
                    } else { /* AbstractJoinExec.buildJoinRow:174-186 */
                        for (int32_t i = 0; i < n_outer; i++) res_put_from(res, col++, &probe[i], r);
                        for (int32_t i = 0; i < right_cols; i++) res_put_from(res, col++, &build[i], m);
                    }
                    res->rows++;
                } else if (jt == ORC_JOIN_RIGHT) {
                    res_reserve_row(res);
                    int32_t col = 0;
                    if (build_outer) { /* ParallelHashJoinExec.buildRightJoinRow:254-271 */
                        used[m] = 1;
                        for (int32_t i = 0; i < n_inner; i++) res_put_from(res, col++, &probe[i], r);
                        for (int32_t i = 0; i < n_outer; i++) res_put_from(res, col++, &build[i], m);
                    } else { /* AbstractJoinExec.buildRightJoinRow:188-199 */
                        for (int32_t i = 0; i < n_inner; i++) res_put_from(res, col++, &build[i], m);
                        for (int32_t i = 0; i < n_outer; i++) res_put_from(res, col++, &probe[i], r);
                    }
                    res->rows++;
                }
                if (single_join && matched) { rc = ORC_ERR_MORE_THAN_ONE_ROW; break; } /* :217-219 */
                matched = 1;
                if (semi_join) break;
            }
            /* matchNext */
            m = t.links[m];
            while (m != -1 && !rows_key_equal(bkey, m, pkey, r, nkeys, spec->key_type)) m = t.links[m];
        }
        if (rc != ORC_OK) break;
        if (outer_join && !build_outer && !matched) { /* outputNullRowInTime() == !buildOuterInput */
            res_reserve_row(res);
            int32_t col = 0;
            if (jt != ORC_JOIN_RIGHT) { /* buildLeftNullRow:201-213 */
                for (int32_t i = 0; i < n_outer; i++) res_put_from(res, col++, &probe[i], r);
                for (int32_t i = 0; i < right_cols; i++) res_put_null(res, col++);
            } else { /* buildRightNullRow:215-225 */
                for (int32_t i = 0; i < n_inner; i++) res_put_null(res, col++);
                for (int32_t i = 0; i < n_outer; i++) res_put_from(res, col++, &probe[i], r);
            }
            res->rows++;
        }
        if (semi_join) {
            int emit = 0;
            if (jt == ORC_JOIN_SEMI && matched) emit = 1;
            else if (jt == ORC_JOIN_ANTI && !matched) {
                emit = 1; /* checkAntiJoinOperands — AbstractJoinExec.java:126-136 */
                for (int32_t i = 0; i < spec->n_anti_operands; i++)
                    if (col_is_null(&probe[spec->anti_operands[i]], r)) { emit = 0; break; }
            }
            if (emit) { /* buildSemiJoinRow:270-275 */
                res_reserve_row(res);
                for (int32_t i = 0; i < n_outer; i++) res_put_from(res, i, &probe[i], r);
                res->rows++;
            }
        }
    }
    /* unmatched build rows of an outer build — ParallelHashJoinExec.nextJoinNullRows:168-201 */
    if (rc == ORC_OK && build_outer && outer_join) {
        for (int64_t p = 0; p < build_rows; p++) {
            if (used[p]) continue;
            res_reserve_row(res);
            int32_t col = 0;
            if (jt != ORC_JOIN_RIGHT) {
                for (int32_t j = 0; j < n_outer; j++) res_put_from(res, col++, &build[j], p);
                for (int32_t j = 0; j < right_cols; j++) res_put_null(res, col++);
            } else {
                for (int32_t j = 0; j < n_inner; j++) res_put_null(res, col++);
                for (int32_t j = 0; j < n_outer; j++) res_put_from(res, col++, &build[j], p);
            }
            res->rows++;
        }
    }
    free(used);
    raw_table_free(&t);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * ChunkRowOpenHashMap — EX/operator/util/ChunkRowOpenHashMap.java:96-166 (single-threaded order).
 * put returns the previous equal row (or -1); get returns the newest equal row.
 * ---------------------------------------------------------------------------------------------- */
void orc_chunk_row_open_hash_map(const orc_col *build, int32_t ncols, int32_t build_rows, const orc_col *probe,
                                 int32_t probe_rows, int32_t *put_result, int32_t *get_result) {
    int32_t types[16];
    for (int i = 0; i < ncols; i++) types[i] = build[i].type;
    int32_t n = orc_array_size(build_rows, select_load_factor(build_rows));
    int32_t mask = n - 1;
    int32_t *keys = (int32_t *)malloc((size_t)n * 4);
    memset(keys, 0xff, (size_t)n * 4);
    for (int32_t p = 0; p < build_rows; p++) {
        int32_t h = orc_mix(row_hash(build, ncols, types, p)) & mask;
        int32_t k;
        int found = 0;
        while ((k = keys[h]) != -1) {
            if (rows_key_equal(build, k, build, p, ncols, types)) { found = 1; break; }
            h = (h + 1) & mask;
        }
        if (found) { put_result[p] = keys[h]; keys[h] = p; }
        else { keys[h] = p; put_result[p] = -1; }
    }
    for (int32_t r = 0; r < probe_rows; r++) {
        int32_t h = orc_mix(row_hash(probe, ncols, types, r)) & mask;
        int32_t k;
        int32_t got = -1;
        while ((k = keys[h]) != -1) {
            if (rows_key_equal(build, k, probe, r, ncols, types)) { got = k; break; }
            h = (h + 1) & mask;
        }
        get_result[r] = got;
    }
    free(keys);
}

/* ------------------------------------------------------------------------------------------------
 * Hash aggregation.  GroupOpenHashMap (EX/operator/util/GroupOpenHashMap.java:82-203),
 * AggOpenHashMap.putChunk/buildChunks (AggOpenHashMap.java:100-194), aggregators under EX/calc/aggfunctions/.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nkeys;
    int32_t key_types[ORC_MAX_KEYS];
    /* TypedBuffer: appended group keys, columnar (EX/operator/util/TypedBuffer.java:60-96) */
    int64_t *kval[ORC_MAX_KEYS]; /* raw 8-byte payload (int widened / double bits as stored double) */
    double *kdbl[ORC_MAX_KEYS];
    uint8_t *knull[ORC_MAX_KEYS];
    int32_t group_count, group_cap;
    int32_t *keys; /* slot -> group id, NOT_EXISTS = -1 */
    int32_t n, mask, size, max_fill;
    float f;
} group_map;

static void gm_init(group_map *g, int32_t nkeys, const int32_t *key_types, int32_t expected) {
    memset(g, 0, sizeof(*g));
    g->nkeys = nkeys;
    for (int i = 0; i < nkeys; i++) g->key_types[i] = key_types[i];
    g->f = 0.75f;                                  /* DEFAULT_LOAD_FACTOR, AggOpenHashMap.java:59 */
    g->n = orc_array_size(expected, g->f);         /* GroupOpenHashMap.java:90 */
    g->mask = g->n - 1;
    g->max_fill = orc_max_fill(g->n, g->f);        /* :92 */
    g->keys = (int32_t *)malloc((size_t)g->n * 4);
    memset(g->keys, 0xff, (size_t)g->n * 4);
    g->group_cap = 1024;
    for (int i = 0; i < nkeys; i++) {
        g->kval[i] = (int64_t *)malloc((size_t)g->group_cap * 8);
        g->kdbl[i] = (double *)malloc((size_t)g->group_cap * 8);
        g->knull[i] = (uint8_t *)malloc((size_t)g->group_cap);
    }
}
static void gm_free(group_map *g) {
    free(g->keys);
    for (int i = 0; i < g->nkeys; i++) { free(g->kval[i]); free(g->kdbl[i]); free(g->knull[i]); }
}
static inline keyval gm_key(const group_map *g, int32_t gid, int32_t c) {
    keyval k; k.is_null = g->knull[c][gid]; k.i = g->kval[c][gid]; k.d = g->kdbl[c][gid];
    return k;
}
static int gm_equals(const group_map *g, int32_t gid, const orc_col *kc, int64_t r) { /* TypedBuffer.equals */
    for (int32_t c = 0; c < g->nkeys; c++) {
        keyval a = gm_key(g, gid, c);
        keyval b = key_at(&kc[c], r, g->key_types[c]);
        if (!key_equal(&a, &b, g->key_types[c])) return 0;
    }
    return 1;
}
static int32_t gm_hash_group(const group_map *g, int32_t gid) {
    uint32_t h = 0;
    for (int32_t c = 0; c < g->nkeys; c++) {
        keyval k = gm_key(g, gid, c);
        h = h * 31u + (uint32_t)key_hash(&k, g->key_types[c]);
    }
    return (int32_t)h;
}
static int32_t gm_append_group(group_map *g, const orc_col *kc, int64_t r) { /* appendGroup:189-192 */
    if (g->group_count == g->group_cap) {
        g->group_cap *= 2;
        for (int i = 0; i < g->nkeys; i++) {
            g->kval[i] = (int64_t *)realloc(g->kval[i], (size_t)g->group_cap * 8);
            g->kdbl[i] = (double *)realloc(g->kdbl[i], (size_t)g->group_cap * 8);
            g->knull[i] = (uint8_t *)realloc(g->knull[i], (size_t)g->group_cap);
        }
    }
    int32_t gid = g->group_count++;
    for (int32_t c = 0; c < g->nkeys; c++) {
        keyval k = kc ? key_at(&kc[c], r, g->key_types[c]) : (keyval){1, 0, 0};
        g->knull[c][gid] = (uint8_t)k.is_null; g->kval[c][gid] = k.i; g->kdbl[c][gid] = k.d;
    }
    return gid;
}
static void gm_rehash(group_map *g) { /* rehash:171-187 — re-insert groups in id order */
    g->n *= 2; g->mask = g->n - 1; g->max_fill = orc_max_fill(g->n, g->f); g->size = 0;
    g->keys = (int32_t *)realloc(g->keys, (size_t)g->n * 4);
    memset(g->keys, 0xff, (size_t)g->n * 4);
    for (int32_t gid = 0; gid < g->group_count; gid++) {
        int32_t h = orc_mix(gm_hash_group(g, gid)) & g->mask;
        while (g->keys[h] != -1) h = (h + 1) & g->mask;
        g->keys[h] = gid;
        g->size++;
    }
}
/* doInnerPutArray:142-169.  `*is_new` tells the caller to append initial aggregator values. */
static int32_t gm_inner_put(group_map *g, const orc_col *kc, int64_t r, int *is_new) {
    int32_t h = orc_mix(row_hash(kc, g->nkeys, g->key_types, r)) & g->mask;
    int32_t k;
    *is_new = 0;
    while ((k = g->keys[h]) != -1) {
        if (gm_equals(g, k, kc, r)) return k;
        h = (h + 1) & g->mask;
    }
    int32_t gid = gm_append_group(g, kc, r);
    *is_new = 1;
    g->keys[h] = gid;
    if (g->size++ >= g->max_fill) gm_rehash(g);
    return gid;
}

/* aggregator state, one per call: covers NullableLong/NullableDouble/Long group states (OPT/state/) */
typedef struct {
    int32_t kind, in_type, out_type;
    int64_t *l;          /* count / long value / min-max long */
    __int128 *wide;      /* exact SUM(int) — equals LittleNum2DecimalSum's long + overflow-to-Decimal escape */
    double *d;           /* double sum / min / max */
    uint8_t *isnull;
    int32_t cap;
} agg_state;

static void as_grow(agg_state *a, int32_t need) {
    if (need <= a->cap) return;
    int32_t nc = a->cap ? a->cap : 1024;
    while (nc < need) nc *= 2;
    a->l = (int64_t *)realloc(a->l, (size_t)nc * 8);
    a->wide = (__int128 *)realloc(a->wide, (size_t)nc * 16);
    a->d = (double *)realloc(a->d, (size_t)nc * 8);
    a->isnull = (uint8_t *)realloc(a->isnull, (size_t)nc);
    a->cap = nc;
}
static void as_append_init(agg_state *a, int32_t gid) { /* Aggregator.appendInitValue */
    as_grow(a, gid + 1);
    a->l[gid] = 0; a->wide[gid] = 0; a->d[gid] = 0;
    /* Count/CountRow/Sum0 append(0L); the others appendNull() */
    a->isnull[gid] = !(a->kind == ORC_AGG_COUNT || a->kind == ORC_AGG_COUNT_STAR || a->kind == ORC_AGG_SUM0);
}
static double java_max(double a, double b) { /* Math.max: NaN wins, +0.0 > -0.0 */
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return signbit(a) ? b : a;
    return a > b ? a : b;
}
static double java_min(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return signbit(a) ? a : b;
    return a < b ? a : b;
}

static void as_accumulate(agg_state *a, const orc_agg_call *call, int32_t gid, const orc_col *in, int64_t r) {
    switch (a->kind) {
    case ORC_AGG_COUNT_STAR: a->l[gid] += 1; return; /* CountRow.java:43-45 */
    case ORC_AGG_COUNT: /* Count.java:44-56 */
        for (int32_t i = 0; i < call->ncols; i++) if (col_is_null(&in[call->cols[i]], r)) return;
        a->l[gid] += 1;
        return;
    default: break;
    }
    const orc_col *c = &in[call->cols[0]];
    if (col_is_null(c, r)) return;
    switch (a->kind) {
    case ORC_AGG_SUM:
        if (a->in_type == ORC_T_FP64) { /* LittleNum2DoubleSum.java:40-53: value + oldValue */
            double v = col_f64(c, r);
            if (a->isnull[gid]) { a->d[gid] = v; a->isnull[gid] = 0; }
            else a->d[gid] = v + a->d[gid];
        } else { /* LittleNum2DecimalSum.java:45-70 (exact) */
            a->wide[gid] += (__int128)col_i64(c, r);
            a->isnull[gid] = 0;
        }
        return;
    case ORC_AGG_AVG: { /* SpecificType2DoubleAvgV2.java:51-65 */
        double v = col_f64(c, r);
        if (a->isnull[gid]) { a->d[gid] = v; a->l[gid] = 1; a->isnull[gid] = 0; }
        else { a->d[gid] = a->d[gid] + v; a->l[gid] += 1; }
        return;
    }
    case ORC_AGG_SUM0: /* Long2LongSum0.java:44-53 (wraps) */
        a->l[gid] = (int64_t)((uint64_t)a->l[gid] + (uint64_t)col_i64(c, r));
        return;
    case ORC_AGG_MIN:
    case ORC_AGG_MAX: { /* Int2IntMax.java:50-64 and siblings */
        int is_max = a->kind == ORC_AGG_MAX;
        if (a->in_type == ORC_T_FP64) {
            double v = col_f64(c, r);
            if (a->isnull[gid]) a->d[gid] = v;
            else a->d[gid] = is_max ? java_max(v, a->d[gid]) : java_min(v, a->d[gid]);
        } else {
            int64_t v = col_i64(c, r);
            if (a->isnull[gid]) a->l[gid] = v;
            else a->l[gid] = is_max ? (v > a->l[gid] ? v : a->l[gid]) : (v < a->l[gid] ? v : a->l[gid]);
        }
        a->isnull[gid] = 0;
        return;
    }
    default: return;
    }
}

static int agg_out_type(int kind, int in_type) {
    switch (kind) {
    case ORC_AGG_COUNT_STAR: case ORC_AGG_COUNT: case ORC_AGG_SUM0: return ORC_T_INT64;
    case ORC_AGG_SUM: return in_type == ORC_T_FP64 ? ORC_T_FP64 : ORC_T_DEC128;
    case ORC_AGG_AVG: return ORC_T_FP64;
    default: return in_type;
    }
}

/* AggOpenHashMap filter handling :114-131 — Boolean or Long objects only; Integer objects never filter */
static inline int filter_pass(const orc_col *in, int32_t filter_arg, int64_t r) {
    if (filter_arg < 0) return 1;
    const orc_col *c = &in[filter_arg];
    if (col_is_null(c, r)) return 1;
    if (c->type == ORC_T_INT64) return ((const int64_t *)c->data)[r] >= 1;
    return 1;
}

typedef struct {
    group_map gm;
    agg_state *st;
    int32_t naggs;
    const orc_agg_call *aggs;
    orc_col keycols[ORC_MAX_KEYS];
    int32_t ngroups_keys;
} agg_map;

static int am_init(agg_map *m, const orc_col *in, const int32_t *groups, int32_t ngroups, const orc_agg_call *aggs,
                   int32_t naggs, int32_t expected) {
    int32_t kt[ORC_MAX_KEYS];
    for (int i = 0; i < ngroups; i++) kt[i] = in[groups[i]].type;
    gm_init(&m->gm, ngroups, kt, expected);
    m->ngroups_keys = ngroups;
    m->naggs = naggs; m->aggs = aggs;
    m->st = (agg_state *)calloc((size_t)(naggs > 0 ? naggs : 1), sizeof(agg_state));
    for (int32_t i = 0; i < naggs; i++) {
        m->st[i].kind = aggs[i].kind;
        m->st[i].in_type = aggs[i].ncols > 0 ? in[aggs[i].cols[0]].type : ORC_T_INT64;
        m->st[i].out_type = agg_out_type(aggs[i].kind, m->st[i].in_type);
        if (aggs[i].kind == ORC_AGG_AVG && m->st[i].in_type != ORC_T_FP64) return ORC_ERR_UNSUPPORTED;
        if (aggs[i].kind == ORC_AGG_SUM0 && m->st[i].in_type != ORC_T_INT64) return ORC_ERR_UNSUPPORTED;
    }
    if (ngroups == 0) { /* noGroupBy: AggOpenHashMap.java:93-96 appends one empty group */
        gm_append_group(&m->gm, NULL, 0);
        for (int32_t i = 0; i < naggs; i++) as_append_init(&m->st[i], 0);
    }
    return ORC_OK;
}
static void am_free(agg_map *m) {
    gm_free(&m->gm);
    for (int32_t i = 0; i < m->naggs; i++) { free(m->st[i].l); free(m->st[i].wide); free(m->st[i].d); free(m->st[i].isnull); }
    free(m->st);
}
/* putChunk over rows [r0, r1) of `in` (a chunk); keycols already offset-free (same row index space) */
static void am_put_chunk(agg_map *m, const orc_col *in, const orc_col *keycols, int64_t r0, int64_t r1,
                         int32_t *gid_buf) {
    if (m->ngroups_keys == 0) {
        for (int64_t r = r0; r < r1; r++) gid_buf[r - r0] = 0;
    } else {
        for (int64_t r = r0; r < r1; r++) {
            int is_new;
            int32_t gid = gm_inner_put(&m->gm, keycols, r, &is_new);
            if (is_new) for (int32_t i = 0; i < m->naggs; i++) as_append_init(&m->st[i], gid);
            gid_buf[r - r0] = gid;
        }
    }
    for (int32_t ai = 0; ai < m->naggs; ai++)
        for (int64_t r = r0; r < r1; r++)
            if (filter_pass(in, m->aggs[ai].filter_arg, r))
                as_accumulate(&m->st[ai], &m->aggs[ai], gid_buf[r - r0], in, r);
}
/* buildChunks: group keys then writeResultTo per aggregator, in group-id order (AggOpenHashMap.java:160-194) */
static orc_result *am_build_result(agg_map *m) {
    int32_t types[64]; int32_t nout = 0;
    for (int32_t i = 0; i < m->ngroups_keys; i++) types[nout++] = m->gm.key_types[i];
    for (int32_t i = 0; i < m->naggs; i++) types[nout++] = m->st[i].out_type;
    orc_result *res = res_new(nout, types, m->gm.group_count + 16);
    for (int32_t gid = 0; gid < m->gm.group_count; gid++) {
        res_reserve_row(res);
        int32_t col = 0;
        for (int32_t c = 0; c < m->ngroups_keys; c++, col++) {
            if (m->gm.knull[c][gid]) { res_put_null(res, col); continue; }
            if (types[col] == ORC_T_INT32) { int32_t v = (int32_t)m->gm.kval[c][gid]; res_put_raw(res, col, &v); }
            else if (types[col] == ORC_T_INT64) res_put_raw(res, col, &m->gm.kval[c][gid]);
            else res_put_raw(res, col, &m->gm.kdbl[c][gid]);
        }
        for (int32_t i = 0; i < m->naggs; i++, col++) {
            agg_state *a = &m->st[i];
            if (a->isnull[gid]) { res_put_null(res, col); continue; }
            switch (a->kind) {
            case ORC_AGG_COUNT_STAR: case ORC_AGG_COUNT: case ORC_AGG_SUM0:
                res_put_raw(res, col, &a->l[gid]); break;
            case ORC_AGG_SUM:
                if (a->in_type == ORC_T_FP64) res_put_raw(res, col, &a->d[gid]);
                else res_put_raw(res, col, &a->wide[gid]);
                break;
            case ORC_AGG_AVG: { /* SpecificType2DoubleAvgV2.writeResultTo:70-84; DoubleType divide by 0 -> NULL */
                if (a->l[gid] == 0) res_put_null(res, col);
                else { double avg = a->d[gid] / (double)a->l[gid]; res_put_raw(res, col, &avg); }
                break;
            }
            default: /* MIN / MAX */
                if (a->in_type == ORC_T_FP64) res_put_raw(res, col, &a->d[gid]);
                else if (a->in_type == ORC_T_INT32) { int32_t v = (int32_t)a->l[gid]; res_put_raw(res, col, &v); }
                else res_put_raw(res, col, &a->l[gid]);
            }
        }
        res->rows++;
    }
    return res;
}

int orc_hash_agg(const orc_col *in, int32_t ncols, int64_t rows, const int32_t *groups, int32_t ngroups,
                 const orc_agg_call *aggs, int32_t naggs, int32_t expected_groups, int32_t chunk_size,
                 orc_result **out) {
    (void)ncols;
    agg_map m;
    int rc = am_init(&m, in, groups, ngroups, aggs, naggs, expected_groups);
    if (rc != ORC_OK) { am_free(&m); return rc; }
    orc_col keycols[ORC_MAX_KEYS];
    for (int32_t i = 0; i < ngroups; i++) keycols[i] = in[groups[i]];
    if (chunk_size <= 0) chunk_size = 1000; /* CHUNK_SIZE default, ConnectionParams.java:1088-1089 */
    int32_t *gid_buf = (int32_t *)malloc((size_t)chunk_size * 4);
    for (int64_t r0 = 0; r0 < rows; r0 += chunk_size) { /* HashAggExec.consumeChunk:133-145 per chunk */
        int64_t r1 = r0 + chunk_size < rows ? r0 + chunk_size : rows;
        am_put_chunk(&m, in, keycols, r0, r1, gid_buf);
    }
    free(gid_buf);
    *out = am_build_result(&m);
    am_free(&m);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * Hash-partition exchange.  PartitioningExchanger.consumeChunk (EX/mpp/operator/PartitioningExchanger.java:71-135),
 * HashBucketFunction / HashPartitionFunction (PartitionedOutputCollector.java:253-302).
 * Output = input rows grouped by destination, original order kept inside each destination.
 * ---------------------------------------------------------------------------------------------- */
int orc_partition_exchange(const orc_col *in, int32_t ncols, int64_t rows, const int32_t *channels, int32_t nchannels,
                           const int32_t *key_types, int32_t nparts, orc_result **out, int64_t *part_counts) {
    orc_col keycols[ORC_MAX_KEYS];
    for (int32_t i = 0; i < nchannels; i++) keycols[i] = in[channels[i]];
    int32_t *pid = (int32_t *)malloc((size_t)(rows > 0 ? rows : 1) * 4);
    memset(part_counts, 0, (size_t)nparts * 8);
    for (int64_t r = 0; r < rows; r++) {
        pid[r] = orc_partition(row_hash(keycols, nchannels, key_types, r), nparts);
        part_counts[pid[r]]++;
    }
    int64_t *cursor = (int64_t *)malloc((size_t)nparts * 8);
    int64_t acc = 0;
    for (int32_t p = 0; p < nparts; p++) { cursor[p] = acc; acc += part_counts[p]; }
    int32_t types[64];
    for (int32_t i = 0; i < ncols; i++) types[i] = in[i].type;
    orc_result *res = res_new(ncols, types, rows + 16);
    res->rows = 0;
    for (int64_t r = 0; r < rows; r++) {
        int64_t dst = cursor[pid[r]]++;
        int64_t save = res->rows;
        res->rows = dst; /* res_put_from writes at res->rows */
        for (int32_t i = 0; i < ncols; i++) res_put_from(res, i, &in[i], r);
        res->rows = save;
    }
    res->rows = rows;
    free(cursor); free(pid);
    *out = res;
    return ORC_OK;
}

/* ================================================================================================
 * Reference-shaped multi-threaded CPU baseline (bench.py cpu_baseline / --impl reference only).
 * P = min(cores, 16) driver threads (EX/utils/ExecUtils.java:375-383, cap :168), CHUNK_SIZE-row chunks.
 * Join: one shared CAS chained table built over chunk ranges (ParallelHashJoinExec.java:406-426) and probed
 * by P drivers.  Generous to the reference: build-row access is a direct index (no ChunksIndex binary
 * search, ChunksIndex.java:53-57), hash vectors reuse buffers, no object allocation.
 * ============================================================================================== */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    const orc_join_spec *spec;
    const orc_col *build, *probe, *bkey, *pkey;
    int32_t n_build, n_probe;
    int64_t build_rows, probe_rows;
    raw_table *t;
    int32_t tid, nthreads, chunk_rows;
    pthread_barrier_t *bar;
    int64_t out_rows;
    uint64_t checksum;
    double t_build, t_probe;
} mtj_arg;

static void *mtj_worker(void *vp) {
    mtj_arg *a = (mtj_arg *)vp;
    const orc_join_spec *spec = a->spec;
    const int32_t nkeys = spec->nkeys, chunk = a->chunk_rows;
    raw_table *t = a->t;
    int32_t *hashes = (int32_t *)malloc((size_t)chunk * 4);

    pthread_barrier_wait(a->bar);
    double t0 = now_s();
    /* build: chunk-range partition per driver (buildHashTable:406-426) */
    int64_t nchunks = (a->build_rows + chunk - 1) / chunk;
    int64_t per = (nchunks + a->nthreads - 1) / a->nthreads;
    int64_t c0 = per * a->tid, c1 = c0 + per < nchunks ? c0 + per : nchunks;
    for (int64_t c = c0; c < c1; c++) {
        int64_t r0 = c * chunk, r1 = r0 + chunk < a->build_rows ? r0 + chunk : a->build_rows;
        for (int64_t r = r0; r < r1; r++) hashes[r - r0] = row_hash(a->bkey, nkeys, spec->key_type, r);
        for (int64_t r = r0; r < r1; r++) {
            int has_null = 0;
            for (int32_t k = 0; k < nkeys; k++) if (col_is_null(&a->bkey[k], r)) { has_null = 1; break; }
            if (has_null) continue;
            int32_t s = orc_mix(hashes[r - r0]) & t->mask;
            /* put(): CAS empty else getAndSet (ConcurrentRawHashTable.java:82-104) */
            int32_t prev = __atomic_exchange_n(&t->keys[s], (int32_t)r, __ATOMIC_ACQ_REL);
            t->links[r] = prev;
        }
    }
    pthread_barrier_wait(a->bar);
    a->t_build = now_s() - t0;

    /* probe: each driver takes chunks tid, tid+P, ... ; INNER only (C2) */
    t0 = now_s();
    const int32_t nout = a->n_probe + a->n_build;
    char *obuf = (char *)malloc((size_t)chunk * 8 * (size_t)nout); /* BlockBuilders of one output chunk */
    int32_t opos = 0;
    int64_t out_rows = 0;
    uint64_t cks = 0;
    int64_t pchunks = (a->probe_rows + chunk - 1) / chunk;
    for (int64_t c = a->tid; c < pchunks; c += a->nthreads) {
        int64_t r0 = c * chunk, r1 = r0 + chunk < a->probe_rows ? r0 + chunk : a->probe_rows;
        for (int64_t r = r0; r < r1; r++) hashes[r - r0] = row_hash(a->pkey, nkeys, spec->key_type, r);
        for (int64_t r = r0; r < r1; r++) {
            int32_t m = t->keys[orc_mix(hashes[r - r0]) & t->mask];
            for (; m != -1; m = t->links[m]) {
                if (!rows_key_equal(a->bkey, m, a->pkey, r, nkeys, spec->key_type)) continue;
                int32_t col = 0;
                for (int32_t i = 0; i < a->n_probe; i++, col++) {
                    int w = type_width(a->probe[i].type);
                    memcpy(obuf + ((size_t)col * chunk + opos) * 8, (const char *)a->probe[i].data + r * w, (size_t)w);
                }
                for (int32_t i = 0; i < a->n_build; i++, col++) {
                    int w = type_width(a->build[i].type);
                    memcpy(obuf + ((size_t)col * chunk + opos) * 8, (const char *)a->build[i].data + (int64_t)m * w, (size_t)w);
                }
                out_rows++;
                if (++opos == chunk) { /* chunk full -> handed downstream */
                    for (int32_t q = 0; q < nout; q++) cks += *(uint64_t *)(obuf + ((size_t)q * chunk + (chunk - 1)) * 8);
                    opos = 0;
                }
            }
        }
    }
    pthread_barrier_wait(a->bar);
    a->t_probe = now_s() - t0;
    a->out_rows = out_rows; a->checksum = cks;
    free(obuf); free(hashes);
    return NULL;
}

int orc_mt_join(const orc_join_spec *spec, const orc_col *outer, int32_t n_outer, int64_t outer_rows,
                const orc_col *inner, int32_t n_inner, int64_t inner_rows, int32_t nthreads, int32_t chunk_rows,
                double *build_seconds, double *probe_seconds, int64_t *out_rows, uint64_t *checksum) {
    if (spec->join_type != ORC_JOIN_INNER || spec->build_outer || inner_rows > 0x7fffffff) return ORC_ERR_UNSUPPORTED;
    orc_col bkey[ORC_MAX_KEYS], pkey[ORC_MAX_KEYS];
    for (int32_t i = 0; i < spec->nkeys; i++) { bkey[i] = inner[spec->inner_key[i]]; pkey[i] = outer[spec->outer_key[i]]; }
    raw_table t;
    raw_table_init(&t, (int32_t)inner_rows);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    mtj_arg *args = (mtj_arg *)calloc((size_t)nthreads, sizeof(mtj_arg));
    for (int32_t i = 0; i < nthreads; i++) {
        args[i] = (mtj_arg){spec, inner, outer, bkey, pkey, n_inner, n_outer, inner_rows, outer_rows, &t,
                            i, nthreads, chunk_rows, &bar, 0, 0, 0, 0};
        pthread_create(&th[i], NULL, mtj_worker, &args[i]);
    }
    *out_rows = 0; *checksum = 0; *build_seconds = 0; *probe_seconds = 0;
    for (int32_t i = 0; i < nthreads; i++) {
        pthread_join(th[i], NULL);
        *out_rows += args[i].out_rows; *checksum += args[i].checksum;
        if (args[i].t_build > *build_seconds) *build_seconds = args[i].t_build;
        if (args[i].t_probe > *probe_seconds) *probe_seconds = args[i].t_probe;
    }
    pthread_barrier_destroy(&bar);
    free(th); free(args);
    raw_table_free(&t);
    return ORC_OK;
}

/* Group-by in the reference's local-MPP shape: P producer drivers hash-partition their chunks
 * (PartitioningExchanger.consumeChunk:71-135 copies rows into per-destination chunks), P single-threaded
 * HashAggExec consumers aggregate (LocalExecutionPlanner.visitHashAgg:1487-1535). */
typedef struct {
    const orc_col *in; int32_t ncols; int64_t rows;
    const int32_t *groups; int32_t ngroups;
    const orc_agg_call *aggs; int32_t naggs; int32_t expected;
    int32_t tid, nthreads, chunk_rows;
    pthread_barrier_t *bar;
    /* per (producer, consumer) partition buffers: columns copied row-wise */
    char ***pbuf;      /* [producer][consumer] -> packed column-major block */
    int64_t **pcount;  /* [producer][consumer] */
    int64_t **pcap;
    int64_t out_groups;
    double seconds;
} mta_arg;

static void *mta_worker(void *vp) {
    mta_arg *a = (mta_arg *)vp;
    const int32_t P = a->nthreads, chunk = a->chunk_rows;
    orc_col keycols[ORC_MAX_KEYS]; int32_t kt[ORC_MAX_KEYS];
    for (int32_t i = 0; i < a->ngroups; i++) { keycols[i] = a->in[a->groups[i]]; kt[i] = keycols[i].type; }
    size_t roww = 0;
    for (int32_t i = 0; i < a->ncols; i++) roww += (size_t)type_width(a->in[i].type);

    pthread_barrier_wait(a->bar);
    double t0 = now_s();
    /* phase 1: this driver partitions its chunk range */
    int64_t nchunks = (a->rows + chunk - 1) / chunk;
    for (int64_t c = a->tid; c < nchunks; c += P) {
        int64_t r0 = c * chunk, r1 = r0 + chunk < a->rows ? r0 + chunk : a->rows;
        for (int64_t r = r0; r < r1; r++) {
            int32_t p = a->ngroups ? orc_partition(row_hash(keycols, a->ngroups, kt, r), P) : 0;
            int64_t n = a->pcount[a->tid][p];
            if (n == a->pcap[a->tid][p]) {
                int64_t nc = a->pcap[a->tid][p] ? a->pcap[a->tid][p] * 2 : 4096;
                a->pbuf[a->tid][p] = (char *)realloc(a->pbuf[a->tid][p], (size_t)nc * roww);
                a->pcap[a->tid][p] = nc;
            }
            char *dst = a->pbuf[a->tid][p] + (size_t)n * roww; /* row-wise append = per-column appendTo */
            for (int32_t i = 0; i < a->ncols; i++) {
                int w = type_width(a->in[i].type);
                memcpy(dst, (const char *)a->in[i].data + r * w, (size_t)w);
                dst += w;
            }
            a->pcount[a->tid][p] = n + 1;
        }
    }
    pthread_barrier_wait(a->bar);
    /* phase 2: consumer `tid` aggregates partition tid from every producer (no NULLs in baseline data) */
    agg_map m;
    orc_col *cols = (orc_col *)calloc((size_t)a->ncols, sizeof(orc_col));
    char **colbuf = (char **)calloc((size_t)a->ncols, sizeof(char *));
    for (int32_t i = 0; i < a->ncols; i++) colbuf[i] = (char *)malloc((size_t)chunk * 8);
    for (int32_t i = 0; i < a->ncols; i++) { cols[i].type = a->in[i].type; cols[i].data = colbuf[i]; cols[i].nulls = NULL; }
    am_init(&m, cols, a->groups, a->ngroups, a->aggs, a->naggs, a->expected);
    orc_col kc[ORC_MAX_KEYS];
    for (int32_t i = 0; i < a->ngroups; i++) kc[i] = cols[a->groups[i]];
    int32_t *gid_buf = (int32_t *)malloc((size_t)chunk * 4);
    for (int32_t prod = 0; prod < P; prod++) {
        int64_t n = a->pcount[prod][a->tid];
        const char *src = a->pbuf[prod][a->tid];
        for (int64_t r0 = 0; r0 < n; r0 += chunk) {
            int64_t r1 = r0 + chunk < n ? r0 + chunk : n;
            for (int64_t r = r0; r < r1; r++) { /* rebuild a columnar chunk */
                const char *row = src + (size_t)r * roww;
                for (int32_t i = 0; i < a->ncols; i++) {
                    int w = type_width(a->in[i].type);
                    memcpy(colbuf[i] + (size_t)(r - r0) * w, row, (size_t)w);
                    row += w;
                }
            }
            am_put_chunk(&m, cols, kc, 0, r1 - r0, gid_buf);
        }
    }
    orc_result *res = am_build_result(&m);
    a->out_groups = res->rows;
    orc_result_free(res);
    pthread_barrier_wait(a->bar);
    a->seconds = now_s() - t0;
    am_free(&m);
    for (int32_t i = 0; i < a->ncols; i++) free(colbuf[i]);
    free(colbuf); free(cols); free(gid_buf);
    return NULL;
}

int orc_mt_hash_agg(const orc_col *in, int32_t ncols, int64_t rows, const int32_t *groups, int32_t ngroups,
                    const orc_agg_call *aggs, int32_t naggs, int32_t expected_groups, int32_t nthreads,
                    int32_t chunk_rows, double *seconds, int64_t *out_groups) {
    const int32_t P = nthreads;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)P);
    char ***pbuf = (char ***)calloc((size_t)P, sizeof(char **));
    int64_t **pcount = (int64_t **)calloc((size_t)P, sizeof(int64_t *));
    int64_t **pcap = (int64_t **)calloc((size_t)P, sizeof(int64_t *));
    for (int32_t i = 0; i < P; i++) {
        pbuf[i] = (char **)calloc((size_t)P, sizeof(char *));
        pcount[i] = (int64_t *)calloc((size_t)P, 8);
        pcap[i] = (int64_t *)calloc((size_t)P, 8);
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)P);
    mta_arg *args = (mta_arg *)calloc((size_t)P, sizeof(mta_arg));
    for (int32_t i = 0; i < P; i++) {
        args[i] = (mta_arg){in, ncols, rows, groups, ngroups, aggs, naggs, expected_groups, i, P, chunk_rows, &bar,
                            pbuf, pcount, pcap, 0, 0};
        pthread_create(&th[i], NULL, mta_worker, &args[i]);
    }
    *seconds = 0; *out_groups = 0;
    for (int32_t i = 0; i < P; i++) {
        pthread_join(th[i], NULL);
        *out_groups += args[i].out_groups;
        if (args[i].seconds > *seconds) *seconds = args[i].seconds;
    }
    for (int32_t i = 0; i < P; i++) {
        for (int32_t j = 0; j < P; j++) free(pbuf[i][j]);
        free(pbuf[i]); free(pcount[i]); free(pcap[i]);
    }
    free(pbuf); free(pcount); free(pcap); free(th); free(args);
    pthread_barrier_destroy(&bar);
    return ORC_OK;
}
