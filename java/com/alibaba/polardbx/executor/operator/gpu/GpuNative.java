/*
 * JNI surface of libgsql_gpu.so (include/gsql_gpu.h).  jni/gsql_jni.c defines one Java_..._GpuNative_<name> per native
 * below (tests/test_jni_boundary.py checks the two lists against each other and compiles the shim against a jni.h stub;
 * the build image has no JDK, so this file itself is compiled where the CN is built).
 */
package com.alibaba.polardbx.executor.operator.gpu;

public final class GpuNative {
    static {
        System.loadLibrary("gsql_jni"); // links libgsql_gpu.so
    }

    private GpuNative() {
    }

    public static final int T_INT32 = 0, T_INT64 = 1, T_FP64 = 2, T_DEC128 = 3;

    /** gsql_ctx_create; one context per operator instance. Throws GpuExecutorException when no GPU is usable. */
    public static native long ctxCreate(int device);

    public static native void ctxDestroy(long ctx);

    /** Number of CUDA devices visible to the process (0 = the planner must keep the stock operators). */
    public static native int deviceCount();

    /**
     * A staging batch: pinned host memory owned by the native side (gsql_host_alloc), filled from Block arrays with
     * GetPrimitiveArrayCritical + memcpy.  columns[i] is int[] / long[] / double[] (IntegerBlock.intArray(),
     * LongBlock.longArray(), DoubleBlock.doubleArray()); nulls[i] is boolean[] or null (AbstractBlock.nulls()).
     * The staging remembers per column whether any appended row was NULL and hands the library `nulls = NULL`
     * otherwise (AbstractBlock.mayHaveNull() == false), so NULL-free inputs reach the packed-row fast paths.
     */
    public static native long stagingCreate(int[] types, int capacityRows);

    public static native void stagingAppend(long staging, Object[] columns, boolean[][] nulls, int arrayOffset, int rows);

    /** Same through a selection vector (Chunk.selection(), Chunk.java:57-79): row i is element selection[i]. */
    public static native void stagingAppendSelected(long staging, Object[] columns, boolean[][] nulls, int[] selection,
                                                    int rows);

    public static native int stagingRows(long staging);

    public static native void stagingReset(long staging);

    public static native void stagingDestroy(long staging);

    /** Copies rows [from, from+rows) of staged column `col` into a fresh Java array (int[] / long[] / double[]). */
    public static native Object stagingColumn(long staging, int col, int from, int rows);

    /** NULL flags of the same rows, or null when none of them is NULL. */
    public static native boolean[] stagingNulls(long staging, int col, int from, int rows);

    // ---- hash join (gsql_join_*)
    public static native long joinCreate(long ctx, int joinType, boolean maxOneRow, boolean buildOuter, int[] outerKeys,
                                         int[] innerKeys, int[] keyTypes, int[] outerTypes, int[] innerTypes,
                                         int[] antiOperands, int[] condCols, long[] condNeValues, long expectedBuildRows);

    public static native void joinBuildConsume(long join, long staging);

    public static native void joinBuildFinish(long join);

    /** Probes the staged rows; results land in `outStaging` (grown as needed). Returns the output row count. */
    public static native int joinProbe(long join, long probeStaging, long outStaging);

    public static native int joinUnmatchedBuild(long join, long outStaging);

    /** Bytes the handle holds in HBM (reported to the query MemoryPool). */
    public static native long joinDeviceBytes(long join);

    public static native void joinDestroy(long join);

    // ---- hash aggregation (gsql_agg_*)
    public static native long aggCreate(long ctx, int[] inputTypes, int[] groups, int[] aggKinds, int[][] aggCols,
                                        int[] filterArgs, long expectedGroups);

    /**
     * Same, with the Project / Filter operators that sit directly under the HashAgg fused into the kernel
     * (VectorizedProjectExec.java:40-143 / VectorizedFilterExec): derived columns {kind, a, b, c} are addressed by the
     * aggregates as input column n_input_cols + i (kind 1: a*(1-b), kind 2: a*(1-b)*(1+c)); rowFilter = {col, op, value}
     * keeps the rows with `col op value` (op: 1 <=, 2 <, 3 >=, 4 >, 5 =, 6 <>; NULL never passes), or null.
     */
    public static native long aggCreateFused(long ctx, int[] inputTypes, int[] groups, int[] aggKinds, int[][] aggCols,
                                             int[] filterArgs, long expectedGroups, int[][] derived, long[] rowFilter);

    public static native void aggConsume(long agg, long staging);

    public static native long aggFinish(long agg);

    public static native int aggNext(long agg, long outStaging, int maxRows);

    public static native void aggDestroy(long agg);

    // ---- vectorised filter / project (gsql_scan_*): programs are flattened {op, arg} pairs + one constant per step
    public static native long scanCreate(long ctx, int[] inputTypes, int[] filterOps, int[] filterArgs, long[] filterConsts,
                                         int[][] outOps, int[][] outArgs, long[][] outConsts);

    public static native int scanApply(long scan, long inStaging, long outStaging);

    public static native void scanDestroy(long scan);

    /**
     * PagesSerde wire format (gsql_serde_serialize): the staging batch as a stream of framed pages of at most pageRows rows
     * each — int32 positionCount | int8 marker (0 = UNCOMPRESSED) | int32 uncompressedSize | int32 sizeInBytes | raw page.
     */
    public static native byte[] serdeSerialize(long ctx, long staging, int pageRows);

    /** Decodes every framed page in pages[offset, offset + length) into outStaging (grown as needed); returns the rows. */
    public static native int serdeDeserialize(long ctx, byte[] pages, int offset, int length, long outStaging);

    // ---- local hash-partition exchange (gsql_xchg_partition)
    public static native long xchgCreate(long ctx, int[] types, int[] channels, int[] keyTypes, int nparts, int mode);

    /** Rows of `inStaging` grouped by consumer into `outStaging`; partCounts[p] = rows of consumer p. */
    public static native void xchgPartition(long xchg, long inStaging, long outStaging, long[] partCounts);

    public static native void xchgDestroy(long xchg);
}
