/*
 * Drop-in for ParallelHashJoinExec (polardbx-executor/.../operator/ParallelHashJoinExec.java:64-85,107-166;
 * AbstractBufferedJoinExec.java:116-264) backed by gsql_join_*.  One object is both the build-side ConsumerExecutor and
 * the probe-side Executor, exactly like AbstractHashJoinExec.  NOT compiled here (no JDK in the build image).
 */
package com.alibaba.polardbx.executor.operator.gpu;

import com.alibaba.polardbx.common.exception.TddlRuntimeException;
import com.alibaba.polardbx.common.exception.code.ErrorCode;
import com.alibaba.polardbx.executor.chunk.Chunk;
import com.alibaba.polardbx.executor.operator.AbstractExecutor;
import com.alibaba.polardbx.executor.operator.ConsumerExecutor;
import com.alibaba.polardbx.executor.operator.Executor;
import com.alibaba.polardbx.optimizer.context.ExecutionContext;
import com.alibaba.polardbx.optimizer.core.datatype.DataType;
import com.alibaba.polardbx.optimizer.core.join.EquiJoinKey;
import com.google.common.collect.ImmutableList;
import com.google.common.util.concurrent.ListenableFuture;
import org.apache.calcite.rel.core.JoinRelType;

import java.util.ArrayDeque;
import java.util.List;

public class GpuParallelHashJoinExec extends AbstractExecutor implements ConsumerExecutor {
    static final int GPU_BATCH_ROWS = 1 << 20;

    private final Executor outerInput, innerInput;
    private final JoinRelType joinType;
    private final boolean maxOneRow, buildOuterInput;
    private final List<EquiJoinKey> joinKeys;
    private final GpuJoinCondition condition;      // restricted otherCondition (col != const ...), null if none
    private final int[] antiJoinOperands;          // InputRef indices, null = NOT EXISTS
    private final List<DataType> dataTypes;

    private long ctx, join, build, probe, out;
    private final ArrayDeque<Chunk> pending = new ArrayDeque<>();
    private boolean probeDone, nullRowsDone, finished;
    private ListenableFuture<?> blocked = NOT_BLOCKED;

    public GpuParallelHashJoinExec(Executor outerInput, Executor innerInput, JoinRelType joinType, boolean maxOneRow,
                                   List<EquiJoinKey> joinKeys, GpuJoinCondition condition, int[] antiJoinOperands,
                                   boolean buildOuterInput, List<DataType> dataTypes, ExecutionContext context) {
        super(context);
        this.outerInput = outerInput;
        this.innerInput = innerInput;
        this.joinType = joinType;
        this.maxOneRow = maxOneRow;
        this.joinKeys = joinKeys;
        this.condition = condition;
        this.antiJoinOperands = antiJoinOperands;
        this.buildOuterInput = buildOuterInput;
        this.dataTypes = dataTypes;
    }

    private Executor buildInput() {
        return buildOuterInput ? outerInput : innerInput;
    }

    private Executor probeInput() {
        return buildOuterInput ? innerInput : outerInput;
    }

    @Override
    public void openConsume() {
        ctx = GpuNative.ctxCreate(GpuDevices.deviceForThisDriver(context));
        join = GpuNative.joinCreate(ctx, GpuTypes.joinType(joinType), maxOneRow, buildOuterInput,
            joinKeys.stream().mapToInt(EquiJoinKey::getOuterIndex).toArray(),
            joinKeys.stream().mapToInt(EquiJoinKey::getInnerIndex).toArray(),
            joinKeys.stream().mapToInt(k -> GpuTypes.code(k.getUnifiedType())).toArray(),
            GpuTypes.codes(outerInput.getDataTypes()), GpuTypes.codes(innerInput.getDataTypes()), antiJoinOperands,
            condition == null ? null : condition.cols, condition == null ? null : condition.neValues, 0L);
        build = GpuNative.stagingCreate(GpuTypes.codes(buildInput().getDataTypes()), GPU_BATCH_ROWS + chunkLimit);
        probe = GpuNative.stagingCreate(GpuTypes.codes(probeInput().getDataTypes()), GPU_BATCH_ROWS + chunkLimit);
        out = GpuNative.stagingCreate(GpuTypes.codes(dataTypes), GPU_BATCH_ROWS + chunkLimit);
    }

    @Override
    public void consumeChunk(Chunk chunk) { // ParallelHashJoinExec.consumeChunk:157-166 (copies; keeps no reference)
        GpuChunks.append(build, chunk);
        if (GpuNative.stagingRows(build) >= GPU_BATCH_ROWS) {
            GpuNative.joinBuildConsume(join, build);
            GpuNative.stagingReset(build);
        }
    }

    @Override
    public void buildConsume() {
        if (GpuNative.stagingRows(build) > 0) {
            GpuNative.joinBuildConsume(join, build);
            GpuNative.stagingReset(build);
        }
        GpuNative.joinBuildFinish(join);
    }

    @Override
    void doOpen() {
        probeInput().open();
    }

    private void probeBatch() {
        int rows;
        try {
            rows = GpuNative.joinProbe(join, probe, out);
        } catch (GpuMoreThanOneRowException e) {
            throw new TddlRuntimeException(ErrorCode.ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW);
        }
        GpuNative.stagingReset(probe);
        for (int from = 0; from < rows; from += chunkLimit) {
            pending.add(GpuChunks.toChunk(out, dataTypes, from, Math.min(chunkLimit, rows - from)));
        }
    }

    @Override
    Chunk doNextChunk() {
        long start = System.currentTimeMillis();
        while (pending.isEmpty() && !finished) {
            if (!probeDone) {
                Chunk c = probeInput().nextChunk();
                if (c != null) {
                    GpuChunks.append(probe, c);
                    // flush on size, or when the 1 s driver quantum is nearly spent (AbstractBufferedJoinExec.java:128)
                    if (GpuNative.stagingRows(probe) >= GPU_BATCH_ROWS || System.currentTimeMillis() - start > 500) {
                        probeBatch();
                    }
                    continue;
                }
                if (!probeInput().produceIsFinished()) {
                    blocked = probeInput().produceIsBlocked();
                    if (GpuNative.stagingRows(probe) > 0) {
                        probeBatch();
                    }
                    return pending.poll();
                }
                probeDone = true;
                if (GpuNative.stagingRows(probe) > 0) {
                    probeBatch();
                }
                continue;
            }
            if (buildOuterInput && !nullRowsDone) { // nextJoinNullRows:168-201
                nullRowsDone = true;
                int rows = GpuNative.joinUnmatchedBuild(join, out);
                for (int from = 0; from < rows; from += chunkLimit) {
                    pending.add(GpuChunks.toChunk(out, dataTypes, from, Math.min(chunkLimit, rows - from)));
                }
                continue;
            }
            finished = true;
        }
        return pending.poll();
    }

    @Override
    public void closeConsume(boolean force) {
    }

    @Override
    void doClose() {
        probeInput().close();
        if (join != 0) {
            GpuNative.joinDestroy(join);
            GpuNative.stagingDestroy(build);
            GpuNative.stagingDestroy(probe);
            GpuNative.stagingDestroy(out);
            GpuNative.ctxDestroy(ctx);
            join = 0;
        }
    }

    @Override
    public List<DataType> getDataTypes() {
        return dataTypes;
    }

    @Override
    public List<Executor> getInputs() {
        return ImmutableList.of(innerInput, outerInput);
    }

    @Override
    public boolean produceIsFinished() {
        return finished && pending.isEmpty();
    }

    @Override
    public ListenableFuture<?> produceIsBlocked() {
        return blocked;
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
