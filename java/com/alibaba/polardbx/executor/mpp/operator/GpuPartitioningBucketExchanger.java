/*
 * GPU twin of PartitioningBucketExchanger (mpp/operator/PartitioningBucketExchanger.java:52-200), the local exchange in
 * front of the bucketed (hybrid) hash join: every row goes to bucket
 *     ExecUtils.partition(Chunk.hashCode(partitionChannels), executors * bucketNum)        (HashBucketFunction,
 *     PartitionedOutputCollector.java:282-300 — the same function as the plain partitioning exchanger, over more parts),
 * chunks are built per bucket — a consumer never sees two buckets in one chunk — and bucket b is delivered to consumer
 * b % executors (sendChunk:161-178).  One gsql_xchg_partition call with nparts = executors * bucketNum groups a whole
 * batch by bucket on the device; bucket ids are bit-exact with the stock exchanger.
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

public class GpuPartitioningBucketExchanger extends LocalExchanger {
    static final int GPU_BATCH_ROWS = 1 << 20;

    private final List<DataType> types;
    private final int[] typeCodes, channels, keyTypes;
    private final List<AtomicBoolean> consumings;
    private final ExecutionContext context;
    private final int chunkLimit;
    private final int totalBucketNum;
    private long ctx, xchg, in, out;

    /** Same argument list as PartitioningBucketExchanger's constructor (:52-60). */
    public GpuPartitioningBucketExchanger(OutputBufferMemoryManager bufferMemoryManager, List<ConsumerExecutor> executors,
                                          LocalExchangersStatus status, boolean asyncConsume, List<DataType> types,
                                          List<Integer> partitionChannels, List<DataType> keyTargetTypes, int bucketNum,
                                          int chunkLimit, ExecutionContext context) {
        super(bufferMemoryManager, executors, status, asyncConsume);
        this.totalBucketNum = executors.size() * bucketNum;
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
            xchg = GpuNative.xchgCreate(ctx, typeCodes, channels, keyTypes, totalBucketNum, 0 /* GSQL_XCHG_HASH */);
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
        long[] counts = new long[totalBucketNum];
        GpuNative.xchgPartition(xchg, in, out, counts);
        GpuNative.stagingReset(in);
        int from = 0;
        for (int bucket = 0; bucket < counts.length; bucket++) { // one bucket per chunk, bucket b -> consumer b % n
            for (int done = 0; done < counts[bucket]; done += chunkLimit) {
                Chunk part = GpuChunks.toChunk(out, types, from + done, (int) Math.min(chunkLimit, counts[bucket] - done));
                deliver(bucket % executors.size(), part);
            }
            from += (int) counts[bucket];
        }
    }

    private void deliver(int partition, Chunk chunk) { // same consumer protection as PartitioningBucketExchanger.java:161-178
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
