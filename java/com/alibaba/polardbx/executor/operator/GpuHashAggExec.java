/*
 * Drop-in for HashAggExec (operator/HashAggExec.java:37,74-91,133-162; AbstractHashAggExec.java:57-63) backed by
 * gsql_agg_*.  Lives in the operator package because AbstractExecutor's template methods doOpen / doNextChunk / doClose
 * are package-private (AbstractExecutor.java:87-91).  Compiled where the CN is built (no JDK in this repository's
 * build image) — see INTEGRATION.md.
 */
package com.alibaba.polardbx.executor.operator;

import com.alibaba.polardbx.executor.chunk.Chunk;
import com.alibaba.polardbx.executor.chunk.GpuChunks;
import com.alibaba.polardbx.executor.operator.gpu.GpuAggSpec;
import com.alibaba.polardbx.executor.operator.gpu.GpuDevices;
import com.alibaba.polardbx.executor.operator.gpu.GpuNative;
import com.alibaba.polardbx.executor.operator.gpu.GpuTypes;
import com.alibaba.polardbx.optimizer.context.ExecutionContext;
import com.alibaba.polardbx.optimizer.core.datatype.DataType;
import com.google.common.collect.ImmutableList;
import com.google.common.util.concurrent.ListenableFuture;

import java.util.List;

public class GpuHashAggExec extends AbstractExecutor implements ConsumerExecutor {
    /** rows accumulated before a batch crosses JNI: 1000-row chunks are far too small for a kernel launch */
    static final int GPU_BATCH_ROWS = 1 << 20;

    private final List<DataType> inputTypes;
    private final List<DataType> outputColumns;
    private final int[] inputCodes;
    private final int[] groups;
    private final GpuAggSpec spec; // kinds / columns / filter args derived from the plan's AggregateCalls
    private final int expectedGroups;

    private long ctx, agg, in, out;
    private boolean finished;

    public GpuHashAggExec(List<DataType> inputTypes, int[] groups, GpuAggSpec spec, List<DataType> outputColumns,
                          int expectedGroups, ExecutionContext context) {
        super(context);
        this.inputTypes = inputTypes;
        this.inputCodes = GpuTypes.codes(inputTypes);
        this.groups = groups;
        this.spec = spec;
        this.outputColumns = outputColumns;
        this.expectedGroups = expectedGroups;
    }

    @Override
    public synchronized void openConsume() {
        if (agg != 0) {
            return; // LocalExchanger.openConsume opens every consumer once, but stay idempotent
        }
        ctx = GpuNative.ctxCreate(GpuDevices.deviceForThisDriver(context));
        agg = GpuNative.aggCreate(ctx, inputCodes, groups, spec.kinds, spec.cols, spec.filterArgs, expectedGroups);
        in = GpuNative.stagingCreate(inputCodes, GPU_BATCH_ROWS + chunkLimit);
        out = GpuNative.stagingCreate(GpuTypes.codes(outputColumns), chunkLimit);
    }

    /** Several exchanger threads may feed one consumer (asyncConsume): serialised like ParallelHashJoinExec:158. */
    @Override
    public synchronized void consumeChunk(Chunk chunk) {
        GpuChunks.append(in, chunk, inputCodes); // Block arrays -> pinned staging (GetPrimitiveArrayCritical inside)
        if (GpuNative.stagingRows(in) >= GPU_BATCH_ROWS) {
            GpuNative.aggConsume(agg, in);
            GpuNative.stagingReset(in);
        }
    }

    @Override
    public synchronized void buildConsume() {
        if (GpuNative.stagingRows(in) > 0) {
            GpuNative.aggConsume(agg, in);
            GpuNative.stagingReset(in);
        }
        GpuNative.aggFinish(agg);
    }

    @Override
    Chunk doNextChunk() {
        int rows = GpuNative.aggNext(agg, out, chunkLimit);
        if (rows == 0) {
            finished = true;
            return null;
        }
        return GpuChunks.toChunk(out, outputColumns, 0, rows);
    }

    @Override
    public synchronized void closeConsume(boolean force) {
        if (agg == 0) {
            return;
        }
        GpuNative.aggDestroy(agg);
        GpuNative.stagingDestroy(in);
        GpuNative.stagingDestroy(out);
        GpuNative.ctxDestroy(ctx);
        agg = in = out = ctx = 0;
    }

    @Override
    void doOpen() {
    }

    @Override
    void doClose() {
        closeConsume(true);
    }

    @Override
    public List<DataType> getDataTypes() {
        return outputColumns;
    }

    @Override
    public List<Executor> getInputs() {
        return ImmutableList.of();
    }

    @Override
    public boolean produceIsFinished() {
        return finished;
    }

    @Override
    public ListenableFuture<?> produceIsBlocked() {
        return NOT_BLOCKED;
    }

    @Override
    public boolean needsInput() {
        return true;
    }

    @Override
    public boolean consumeIsFinished() {
        return false;
    }

    @Override
    public ListenableFuture<?> consumeIsBlocked() {
        return ConsumerExecutor.NOT_BLOCKED;
    }
}
