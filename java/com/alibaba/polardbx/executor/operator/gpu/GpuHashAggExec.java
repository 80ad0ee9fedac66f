/*
 * Drop-in for HashAggExec (polardbx-executor/.../operator/HashAggExec.java:37,74-91,133-162) backed by gsql_agg_*.
 * NOT compiled here (no JDK in the build image) — see INTEGRATION.md.
 */
package com.alibaba.polardbx.executor.operator.gpu;

import com.alibaba.polardbx.executor.chunk.Block;
import com.alibaba.polardbx.executor.chunk.Chunk;
import com.alibaba.polardbx.executor.operator.AbstractExecutor;
import com.alibaba.polardbx.executor.operator.ConsumerExecutor;
import com.alibaba.polardbx.executor.operator.Executor;
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
    private final int[] groups;
    private final GpuAggSpec spec; // kinds / columns / filter args derived from List<Aggregator> at plan time
    private final int expectedGroups;

    private long ctx, agg, in, out;
    private boolean finished;

    public GpuHashAggExec(List<DataType> inputTypes, int[] groups, GpuAggSpec spec, List<DataType> outputColumns,
                          int expectedGroups, ExecutionContext context) {
        super(context);
        this.inputTypes = inputTypes;
        this.groups = groups;
        this.spec = spec;
        this.outputColumns = outputColumns;
        this.expectedGroups = expectedGroups;
    }

    @Override
    public void openConsume() {
        ctx = GpuNative.ctxCreate(GpuDevices.deviceForThisDriver(context));
        agg = GpuNative.aggCreate(ctx, GpuTypes.codes(inputTypes), groups, spec.kinds, spec.cols, spec.filterArgs,
            expectedGroups);
        in = GpuNative.stagingCreate(GpuTypes.codes(inputTypes), GPU_BATCH_ROWS + chunkLimit);
        out = GpuNative.stagingCreate(GpuTypes.codes(outputColumns), chunkLimit);
    }

    @Override
    public void consumeChunk(Chunk chunk) {
        GpuChunks.append(in, chunk); // Block arrays -> pinned staging (GetPrimitiveArrayCritical inside)
        if (GpuNative.stagingRows(in) >= GPU_BATCH_ROWS) {
            GpuNative.aggConsume(agg, in);
            GpuNative.stagingReset(in);
        }
    }

    @Override
    public void buildConsume() {
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
    public void closeConsume(boolean force) {
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
        if (agg != 0) {
            closeConsume(true);
        }
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
