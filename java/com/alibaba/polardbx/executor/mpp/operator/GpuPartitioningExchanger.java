/*
 * GPU twin of PartitioningExchanger (mpp/operator/PartitioningExchanger.java:46-135): rows of the incoming chunks are
 * routed to consumer ExecUtils.partition(Chunk.hashCode(partitionChannels)) — bit-exact with the stock exchanger, so a
 * plan may mix both — but whole batches are partitioned by one gsql_xchg_partition call (histogram, scan, scatter on
 * the device) instead of per-row appendTo calls into per-destination builders.  Chosen by LocalExchangeConsumerFactory
 * for LocalExchangeMode.PARTITION when GPU operators are enabled and every column has a GPU block type.
 */
package com.alibaba.polardbx.executor.mpp.operator;

import com.alibaba.polardbx.executor.chunk.Chunk;
import com.alibaba.polardbx.executor.chunk.GpuChunks;
import com.alibaba.polardbx.executor.mpp.execution.buffer.OutputBufferMemoryManager;
import com.alibaba.polardbx.executor.operator.ConsumerExecutor;
import com.alibaba.polardbx.executor.operator.gpu.GpuDevices;
import com.alibaba.polardbx.executor.operator.gpu.GpuNative;
import com.alibaba.polardbx.executor.operator.gpu.GpuTypes;
import com.alibaba.polardbx.optimizer.context.ExecutionContext;
import com.alibaba.polardbx.optimizer.core.datatype.DataType;

import java.util.List;
import java.util.concurrent.atomic.AtomicBoolean;

public class GpuPartitioningExchanger extends LocalExchanger {
    static final int GPU_BATCH_ROWS = 1 << 20;

    private final List<DataType> types;
    private final int[] typeCodes, channels, keyTypes;
    private final List<AtomicBoolean> consumings;
    private final ExecutionContext context;
    private final int chunkLimit;
    private long ctx, xchg, in, out;

    public GpuPartitioningExchanger(OutputBufferMemoryManager bufferMemoryManager, List<ConsumerExecutor> executors,
                                    LocalExchangersStatus status, boolean asyncConsume, List<DataType> types,
                                    List<Integer> partitionChannels, List<DataType> keyTargetTypes,
                                    ExecutionContext context, int chunkLimit) {
        super(bufferMemoryManager, executors, status, asyncConsume);
        this.types = types;
        this.typeCodes = GpuTypes.codes(types);
        this.channels = partitionChannels.stream().mapToInt(Integer::intValue).toArray();
        this.keyTypes = new int[channels.length];
        for (int i = 0; i < channels.length; i++) { // keyTargetTypes empty = no conversion (PartitioningExchanger.java:58-66)
            this.keyTypes[i] = keyTargetTypes.isEmpty() ? typeCodes[channels[i]] : GpuTypes.code(keyTargetTypes.get(i));
        }
        this.consumings = status.getConsumings();
        this.context = context;
        this.chunkLimit = chunkLimit;
    }

    private void ensureOpen() {
        if (xchg == 0) {
            ctx = GpuNative.ctxCreate(GpuDevices.deviceForThisDriver(context));
            xchg = GpuNative.xchgCreate(ctx, typeCodes, channels, keyTypes, executors.size(), 0 /* GSQL_XCHG_HASH */);
            in = GpuNative.stagingCreate(typeCodes, GPU_BATCH_ROWS + chunkLimit);
            out = GpuNative.stagingCreate(typeCodes, GPU_BATCH_ROWS + chunkLimit);
        }
    }

    @Override
    public void consumeChunk(Chunk chunk) {
        ensureOpen();
        GpuChunks.append(in, chunk, typeCodes);
        if (GpuNative.stagingRows(in) >= GPU_BATCH_ROWS) {
            flush();
        }
    }

    private void flush() {
        int rows = GpuNative.stagingRows(in);
        if (rows == 0) {
            return;
        }
        long[] counts = new long[executors.size()];
        GpuNative.xchgPartition(xchg, in, out, counts);
        GpuNative.stagingReset(in);
        int from = 0;
        for (int p = 0; p < counts.length; p++) {
            for (int done = 0; done < counts[p]; done += chunkLimit) {
                Chunk part = GpuChunks.toChunk(out, types, from + done, (int) Math.min(chunkLimit, counts[p] - done));
                deliver(p, part);
            }
            from += (int) counts[p];
        }
    }

    private void deliver(int partition, Chunk chunk) { // same consumer protection as PartitioningExchanger.java:111-131
        if (asyncConsume) {
            executors.get(partition).consumeChunk(chunk);
            return;
        }
        AtomicBoolean consuming = consumings.get(partition);
        while (true) {
            if (consuming.compareAndSet(false, true)) {
                try {
                    executors.get(partition).consumeChunk(chunk);
                } finally {
                    consuming.set(false);
                }
                return;
            }
        }
    }

    @Override
    public void buildConsume() {
        if (xchg != 0) {
            flush();
        }
        super.buildConsume();
    }

    @Override
    public void closeConsume(boolean force) {
        if (xchg != 0) {
            GpuNative.xchgDestroy(xchg);
            GpuNative.stagingDestroy(in);
            GpuNative.stagingDestroy(out);
            GpuNative.ctxDestroy(ctx);
            xchg = in = out = ctx = 0;
        }
        super.closeConsume(force);
    }

    @Override
    public boolean consumeIsFinished() {
        return false;
    }
}
