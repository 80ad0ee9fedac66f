/*
 * Drop-in for ParallelHashJoinExec (operator/ParallelHashJoinExec.java:64-85,107-166; AbstractBufferedJoinExec.java:116-264)
 * backed by gsql_join_*.  One object is both the build-side ConsumerExecutor and the probe-side Executor, exactly like
 * AbstractHashJoinExec.  The probe instances of one join share the built table through GpuJoinShared (the role of
 * ParallelHashJoinExec.Synchronizer): the last buildConsume() builds it, probes start after the build pipeline's future
 * completes (LocalExecutionPlanner.java:961).  Lives in the operator package because AbstractExecutor's template
 * methods are package-private (AbstractExecutor.java:87-91).  Compiled where the CN is built — see INTEGRATION.md.
 */
package com.alibaba.polardbx.executor.operator;

import com.alibaba.polardbx.executor.chunk.Chunk;
import com.alibaba.polardbx.executor.chunk.GpuChunks;
import com.alibaba.polardbx.executor.operator.gpu.GpuDevices;
import com.alibaba.polardbx.executor.operator.gpu.GpuJoinCondition;
import com.alibaba.polardbx.executor.operator.gpu.GpuNative;
import com.alibaba.polardbx.executor.operator.gpu.GpuTypes;
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

    /** State shared by the instances of one join (one per probe driver): the context, the handle, the build staging. */
    public static final class GpuJoinShared {
        final int consumers;
        long ctx, join, build;
        int opened, built, closed;
        boolean nullRowsTaken;

        public GpuJoinShared(int consumers) {
            this.consumers = consumers;
        }
    }

    private final GpuJoinShared shared;
    private final Executor outerInput, innerInput;
    private final JoinRelType joinType;
    private final boolean maxOneRow, buildOuterInput;
    private final List<EquiJoinKey> joinKeys;
    private final GpuJoinCondition condition;      // restricted otherCondition (col <> const ...), null if none
    private final int[] antiJoinOperands;          // InputRef indices, null = NOT EXISTS
    private final List<DataType> dataTypes;
    private final int[] buildCodes, probeCodes;

    private long probe, out;
    private final ArrayDeque<Chunk> pending = new ArrayDeque<>();
    private boolean probeDone, nullRowsDone, finished;
    private ListenableFuture<?> blocked = NOT_BLOCKED;

    public GpuParallelHashJoinExec(GpuJoinShared shared, Executor outerInput, Executor innerInput, JoinRelType joinType,
                                   boolean maxOneRow, List<EquiJoinKey> joinKeys, GpuJoinCondition condition,
                                   int[] antiJoinOperands, boolean buildOuterInput, List<DataType> dataTypes,
                                   ExecutionContext context) {
        super(context);
        this.shared = shared;
        this.outerInput = outerInput;
        this.innerInput = innerInput;
        this.joinType = joinType;
        this.maxOneRow = maxOneRow;
        this.joinKeys = joinKeys;
        this.condition = condition;
        this.antiJoinOperands = antiJoinOperands;
        this.buildOuterInput = buildOuterInput;
        this.dataTypes = dataTypes;
        this.buildCodes = GpuTypes.codes(buildInput().getDataTypes());
        this.probeCodes = GpuTypes.codes(probeInput().getDataTypes());
    }

    private Executor buildInput() {
        return buildOuterInput ? outerInput : innerInput;
    }

    private Executor probeInput() {
        return buildOuterInput ? innerInput : outerInput;
    }

    @Override
    public void openConsume() {
        synchronized (shared) {
            if (shared.opened++ > 0) {
                return;
            }
            shared.ctx = GpuNative.ctxCreate(GpuDevices.deviceForThisDriver(context));
            shared.join = GpuNative.joinCreate(shared.ctx, GpuTypes.joinType(joinType), maxOneRow, buildOuterInput,
                joinKeys.stream().mapToInt(EquiJoinKey::getOuterIndex).toArray(),
                joinKeys.stream().mapToInt(EquiJoinKey::getInnerIndex).toArray(),
                joinKeys.stream().mapToInt(k -> GpuTypes.code(k.getUnifiedType())).toArray(),
                GpuTypes.codes(outerInput.getDataTypes()), GpuTypes.codes(innerInput.getDataTypes()), antiJoinOperands,
                condition == null ? null : condition.cols, condition == null ? null : condition.neValues, 0L);
            shared.build = GpuNative.stagingCreate(buildCodes, GPU_BATCH_ROWS + chunkLimit);
        }
    }

    @Override
    public void consumeChunk(Chunk chunk) { // ParallelHashJoinExec.consumeChunk:157-166 (copies; keeps no reference)
        synchronized (shared) {
            GpuChunks.append(shared.build, chunk, buildCodes);
            if (GpuNative.stagingRows(shared.build) >= GPU_BATCH_ROWS) {
                GpuNative.joinBuildConsume(shared.join, shared.build);
                GpuNative.stagingReset(shared.build);
            }
        }
    }

    @Override
    public void buildConsume() { // every consumer instance gets this exactly once (LocalExchanger.buildConsume:97-121)
        synchronized (shared) {
            if (++shared.built < shared.consumers) {
                return;
            }
            if (GpuNative.stagingRows(shared.build) > 0) {
                GpuNative.joinBuildConsume(shared.join, shared.build);
                GpuNative.stagingReset(shared.build);
            }
            GpuNative.joinBuildFinish(shared.join);
            context.getMemoryPool().getMemoryAllocatorCtx().allocateReservedMemory(GpuNative.joinDeviceBytes(shared.join));
        }
    }

    @Override
    void doOpen() {
        probeInput().open();
        probe = GpuNative.stagingCreate(probeCodes, GPU_BATCH_ROWS + chunkLimit);
        out = GpuNative.stagingCreate(GpuTypes.codes(dataTypes), GPU_BATCH_ROWS + chunkLimit);
    }

    private void emit(int rows) {
        for (int from = 0; from < rows; from += chunkLimit) {
            pending.add(GpuChunks.toChunk(out, dataTypes, from, Math.min(chunkLimit, rows - from)));
        }
    }

    private void probeBatch() {
        int rows;
        synchronized (shared) { // a gsql handle is thread-compatible: the probe drivers take turns on it
            rows = GpuNative.joinProbe(shared.join, probe, out); // GpuMoreThanOneRowException carries
        }                                                        // ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW
        GpuNative.stagingReset(probe);
        emit(rows);
    }

    @Override
    Chunk doNextChunk() {
        long start = System.currentTimeMillis();
        while (pending.isEmpty() && !finished) {
            if (!probeDone) {
                Chunk c = probeInput().nextChunk();
                if (c != null) {
                    GpuChunks.append(probe, c, probeCodes);
                    // flush on size, or when the 1 s driver quantum is nearly spent (TaskExecutor.java:414)
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
            if (buildOuterInput && !nullRowsDone) { // nextJoinNullRows:168-201 — emitted once, by the last prober
                nullRowsDone = true;
                synchronized (shared) {
                    if (++shared.closed == shared.consumers && !shared.nullRowsTaken) {
                        shared.nullRowsTaken = true;
                        emit(GpuNative.joinUnmatchedBuild(shared.join, out));
                    }
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
        if (probe != 0) {
            GpuNative.stagingDestroy(probe);
            GpuNative.stagingDestroy(out);
            probe = out = 0;
        }
        synchronized (shared) {
            if (--shared.opened == 0 && shared.join != 0) {
                GpuNative.joinDestroy(shared.join);
                GpuNative.stagingDestroy(shared.build);
                GpuNative.ctxDestroy(shared.ctx);
                shared.join = shared.build = shared.ctx = 0;
            }
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
