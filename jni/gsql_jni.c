/*
 * gsql_jni.c — thin JNI shim between com.alibaba.polardbx.executor.operator.gpu.GpuNative and libgsql_gpu.so.
 *
 * NOT compiled in this repository (the build image has no JDK, hence no <jni.h>); build where the CN is built:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include gsql_jni.c -L<dir> -lgsql_gpu -o libgsql_jni.so
 *
 * A "staging" object is a set of pinned host column buffers (gsql_host_alloc) that Java fills chunk by chunk with
 * GetPrimitiveArrayCritical + memcpy ("pins Chunk blocks into device memory" = array -> pinned -> cudaMemcpyAsync inside
 * the library).  Non-zero gsql_status becomes TddlRuntimeException(ErrorCode.ERR_EXECUTOR, gsql_last_error) — or
 * GpuMoreThanOneRowException for GSQL_E_MORE_THAN_ONE_ROW, which the operator maps to
 * ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW (AbstractBufferedJoinExec.java:217-219).
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "gsql_gpu.h"

typedef struct staging {
    int32_t ncols;
    int32_t types[GSQL_MAX_COLS * 2];
    int64_t cap, rows;
    void *data[GSQL_MAX_COLS * 2];
    uint8_t *nulls[GSQL_MAX_COLS * 2];
    gsql_col cols[GSQL_MAX_COLS * 2];
    gsql_batch batch;
} staging;

static int width(int t) { return t == GSQL_T_INT32 ? 4 : t == GSQL_T_DEC128 ? 16 : 8; }

static void throw_status(JNIEnv *env, gsql_ctx *ctx, int st) {
    const char *cls = st == GSQL_E_MORE_THAN_ONE_ROW ? "com/alibaba/polardbx/executor/operator/gpu/GpuMoreThanOneRowException"
                                                     : "com/alibaba/polardbx/executor/operator/gpu/GpuExecutorException";
    (*env)->ThrowNew(env, (*env)->FindClass(env, cls), ctx ? gsql_last_error(ctx) : "gsql error");
}

static gsql_batch *as_batch(staging *s) {
    for (int i = 0; i < s->ncols; i++) {
        s->cols[i].type = s->types[i];
        s->cols[i].data = s->data[i];
        s->cols[i].nulls = s->nulls[i];
    }
    s->batch.rows = s->rows;
    s->batch.ncols = s->ncols;
    s->batch.mem = GSQL_MEM_HOST;
    s->batch.cols = s->cols;
    return &s->batch;
}

static int staging_reserve(staging *s, int64_t rows) {
    if (rows <= s->cap) return 0;
    int64_t cap = s->cap ? s->cap : 4096;
    while (cap < rows) cap *= 2;
    for (int i = 0; i < s->ncols; i++) {
        void *nd, *nn;
        if (gsql_host_alloc((size_t)cap * width(s->types[i]), &nd) != GSQL_OK) return -1;
        if (gsql_host_alloc((size_t)cap, &nn) != GSQL_OK) return -1;
        if (s->rows) {
            memcpy(nd, s->data[i], (size_t)s->rows * width(s->types[i]));
            memcpy(nn, s->nulls[i], (size_t)s->rows);
        }
        gsql_host_free(s->data[i]);
        gsql_host_free(s->nulls[i]);
        s->data[i] = nd;
        s->nulls[i] = (uint8_t *)nn;
    }
    s->cap = cap;
    return 0;
}

JNIEXPORT jlong JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_ctxCreate(JNIEnv *env, jclass c, jint device) {
    gsql_ctx *ctx = NULL;
    int st = gsql_ctx_create(device, &ctx);
    if (st != GSQL_OK) throw_status(env, NULL, st); /* no CPU fallback: the planner must not have chosen this operator */
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_ctxDestroy(JNIEnv *env, jclass c, jlong ctx) {
    gsql_ctx_destroy((gsql_ctx *)(intptr_t)ctx);
}

JNIEXPORT jlong JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_stagingCreate(JNIEnv *env, jclass c, jintArray types, jint cap) {
    staging *s = (staging *)calloc(1, sizeof(staging));
    s->ncols = (*env)->GetArrayLength(env, types);
    (*env)->GetIntArrayRegion(env, types, 0, s->ncols, (jint *)s->types);
    staging_reserve(s, cap);
    return (jlong)(intptr_t)s;
}

/* Block arrays -> pinned staging.  columns[i] is int[] / long[] / double[] (IntegerBlock.intArray():217,
 * LongBlock.longArray():191, DoubleBlock), nulls[i] is boolean[] or null (AbstractBlock.nulls():116). */
JNIEXPORT void JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_stagingAppend(JNIEnv *env, jclass c, jlong h, jobjectArray columns,
                                                                                             jobjectArray nulls, jint offset, jint rows) {
    staging *s = (staging *)(intptr_t)h;
    if (staging_reserve(s, s->rows + rows)) return;
    for (int i = 0; i < s->ncols; i++) {
        int w = width(s->types[i]);
        jarray col = (jarray)(*env)->GetObjectArrayElement(env, columns, i);
        void *p = (*env)->GetPrimitiveArrayCritical(env, col, NULL);
        memcpy((char *)s->data[i] + (size_t)s->rows * w, (char *)p + (size_t)offset * w, (size_t)rows * w);
        (*env)->ReleasePrimitiveArrayCritical(env, col, p, JNI_ABORT);
        jarray nl = nulls ? (jarray)(*env)->GetObjectArrayElement(env, nulls, i) : NULL;
        if (nl) {
            void *q = (*env)->GetPrimitiveArrayCritical(env, nl, NULL);
            memcpy(s->nulls[i] + s->rows, (char *)q + offset, (size_t)rows); /* jboolean is one byte */
            (*env)->ReleasePrimitiveArrayCritical(env, nl, q, JNI_ABORT);
        } else {
            memset(s->nulls[i] + s->rows, 0, (size_t)rows);
        }
    }
    s->rows += rows;
}

JNIEXPORT jint JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_stagingRows(JNIEnv *env, jclass c, jlong h) {
    return (jint)((staging *)(intptr_t)h)->rows;
}
JNIEXPORT void JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_stagingReset(JNIEnv *env, jclass c, jlong h) {
    ((staging *)(intptr_t)h)->rows = 0;
}
JNIEXPORT void JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_stagingDestroy(JNIEnv *env, jclass c, jlong h) {
    staging *s = (staging *)(intptr_t)h;
    if (!s) return;
    for (int i = 0; i < s->ncols; i++) {
        gsql_host_free(s->data[i]);
        gsql_host_free(s->nulls[i]);
    }
    free(s);
}

/* ---- join ------------------------------------------------------------------------------------------------- */
static void fill_ints(JNIEnv *env, jintArray a, int32_t *dst, int32_t *n, int max) {
    *n = a ? (*env)->GetArrayLength(env, a) : 0;
    if (*n > max) *n = max;
    if (*n) (*env)->GetIntArrayRegion(env, a, 0, *n, (jint *)dst);
}

JNIEXPORT jlong JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_joinCreate(
    JNIEnv *env, jclass c, jlong ctx, jint joinType, jboolean maxOneRow, jboolean buildOuter, jintArray outerKeys, jintArray innerKeys,
    jintArray keyTypes, jintArray outerTypes, jintArray innerTypes, jintArray antiOperands, jintArray condCols, jlongArray condNe,
    jlong expectedBuildRows) {
    gsql_join_spec s;
    memset(&s, 0, sizeof(s));
    int32_t n;
    s.join_type = joinType;
    s.max_one_row = maxOneRow;
    s.build_outer = buildOuter;
    fill_ints(env, outerKeys, s.outer_key, &s.nkeys, GSQL_MAX_KEYS);
    fill_ints(env, innerKeys, s.inner_key, &n, GSQL_MAX_KEYS);
    fill_ints(env, keyTypes, s.key_type, &n, GSQL_MAX_KEYS);
    fill_ints(env, outerTypes, s.outer_types, &s.n_outer_cols, GSQL_MAX_COLS);
    fill_ints(env, innerTypes, s.inner_types, &s.n_inner_cols, GSQL_MAX_COLS);
    fill_ints(env, antiOperands, s.anti_operands, &s.n_anti_operands, GSQL_MAX_KEYS);
    fill_ints(env, condCols, s.cond_col, &s.n_cond, 4);
    if (s.n_cond) (*env)->GetLongArrayRegion(env, condNe, 0, s.n_cond, (jlong *)s.cond_ne_value);
    s.expected_build_rows = expectedBuildRows;
    gsql_join *j = NULL;
    int st = gsql_join_create((gsql_ctx *)(intptr_t)ctx, &s, &j);
    if (st != GSQL_OK) throw_status(env, (gsql_ctx *)(intptr_t)ctx, st);
    return (jlong)(intptr_t)j;
}

/* ---- aggregation with the Project / Filter under it fused in (gsql_agg_spec.derived / row_filter_*) ------------------ */
JNIEXPORT jlong JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_aggCreateFused(
    JNIEnv *env, jclass c, jlong ctx, jintArray inputTypes, jintArray groups, jintArray aggKinds, jobjectArray aggCols, jintArray filterArgs,
    jlong expectedGroups, jobjectArray derived, jlongArray rowFilter) {
    gsql_agg_spec s;
    memset(&s, 0, sizeof(s));
    int32_t kinds[GSQL_MAX_AGGS], fargs[GSQL_MAX_AGGS], n;
    fill_ints(env, inputTypes, s.input_types, &s.n_input_cols, GSQL_MAX_COLS);
    fill_ints(env, groups, s.groups, &s.ngroups, GSQL_MAX_KEYS);
    fill_ints(env, aggKinds, kinds, &s.naggs, GSQL_MAX_AGGS);
    fill_ints(env, filterArgs, fargs, &n, GSQL_MAX_AGGS);
    for (int i = 0; i < s.naggs; i++) {
        jintArray cols = (jintArray)(*env)->GetObjectArrayElement(env, aggCols, i);
        s.aggs[i].kind = kinds[i];
        s.aggs[i].filter_arg = i < n ? fargs[i] : -1;
        fill_ints(env, cols, s.aggs[i].cols, &s.aggs[i].ncols, 4);
    }
    s.expected_groups = expectedGroups;
    s.n_derived = derived ? (*env)->GetArrayLength(env, derived) : 0;
    if (s.n_derived > GSQL_MAX_DERIVED) s.n_derived = GSQL_MAX_DERIVED;
    for (int i = 0; i < s.n_derived; i++) {
        int32_t d[4] = {0, 0, 0, 0}, nd;
        fill_ints(env, (jintArray)(*env)->GetObjectArrayElement(env, derived, i), d, &nd, 4);
        s.derived[i].kind = d[0];
        s.derived[i].a = d[1];
        s.derived[i].b = d[2];
        s.derived[i].c = d[3];
    }
    s.row_filter_col = -1;
    s.row_filter_op = GSQL_CMP_NONE;
    if (rowFilter) {
        jlong f[3];
        (*env)->GetLongArrayRegion(env, rowFilter, 0, 3, f);
        s.row_filter_col = (int32_t)f[0];
        s.row_filter_op = (int32_t)f[1];
        s.row_filter_value = f[2];
    }
    gsql_agg *a = NULL;
    int st = gsql_agg_create((gsql_ctx *)(intptr_t)ctx, &s, &a);
    if (st != GSQL_OK) throw_status(env, (gsql_ctx *)(intptr_t)ctx, st);
    return (jlong)(intptr_t)a;
}

/* The remaining entry points (joinBuildConsume / joinBuildFinish / joinProbe / joinUnmatchedBuild / agg* / xchg* /
 * stagingColumn / stagingNulls) follow the same pattern: as_batch(staging) in, gsql_* call, throw_status on error; for
 * outputs the shim calls the function with the staging's capacity, grows it on GSQL_E_CAPACITY and calls again. */
JNIEXPORT jint JNICALL Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_joinProbe(JNIEnv *env, jclass c, jlong jh, jlong ph, jlong oh) {
    gsql_join *j = (gsql_join *)(intptr_t)jh;
    staging *p = (staging *)(intptr_t)ph, *o = (staging *)(intptr_t)oh;
    int64_t rows = 0;
    for (;;) {
        o->rows = 0;
        gsql_batch *ob = as_batch(o);
        int st = gsql_join_probe(j, as_batch(p), ob, o->cap, &rows);
        if (st == GSQL_E_CAPACITY) {
            if (staging_reserve(o, rows)) return -1;
            continue;
        }
        if (st != GSQL_OK) {
            throw_status(env, NULL, st);
            return -1;
        }
        o->rows = rows;
        return (jint)rows;
    }
}
