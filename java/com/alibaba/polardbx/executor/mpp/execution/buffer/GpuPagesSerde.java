/*
 * PagesSerde (mpp/execution/buffer/PagesSerde.java:40-130) with the encode / decode of PagesSerdeUtil.writeRawPage /
 * readRawPage and the *BlockEncoding classes (chunk/LongBlockEncoding.java:47-73, EncoderUtil.java:43-150) done by
 * gsql_serde_* on the GPU: same bytes on the wire, so a GPU task and a stock Java task can sit on either side of an
 * MppExchange.  Only the UNCOMPRESSED form is produced (PagesSerde falls back to it whenever LZ4 saves < 20 %,
 * PagesSerde.java:41,86-89); a COMPRESSED page received from a Java producer is handed to the stock PagesSerde.
 * INT / BIGINT / DOUBLE columns (GpuTypes.supported).
 */
package com.alibaba.polardbx.executor.mpp.execution.buffer;

import com.alibaba.polardbx.executor.chunk.Chunk;
import com.alibaba.polardbx.executor.chunk.GpuChunks;
import com.alibaba.polardbx.executor.operator.gpu.GpuNative;
import com.alibaba.polardbx.executor.operator.gpu.GpuTypes;
import com.alibaba.polardbx.optimizer.core.datatype.DataType;
import io.airlift.slice.Slice;
import io.airlift.slice.Slices;

import java.util.ArrayList;
import java.util.List;

public class GpuPagesSerde implements AutoCloseable {
    /** bytes in front of every raw page: positionCount, marker, uncompressedSize, sizeInBytes (PagesSerdeUtil.java:36-58) */
    static final int FRAME_BYTES = 4 + 1 + 4 + 4;

    private final List<DataType> types;
    private final int[] typeCodes;
    private final PagesSerde stock; // for pages that arrive LZ4-compressed from a Java producer
    private final long ctx;
    private long in, out;

    public GpuPagesSerde(int device, List<DataType> types, PagesSerde stock) {
        this.types = types;
        this.typeCodes = GpuTypes.codes(types);
        this.stock = stock;
        this.ctx = GpuNative.ctxCreate(device);
        this.in = GpuNative.stagingCreate(typeCodes, 4096);
        this.out = GpuNative.stagingCreate(typeCodes, 4096);
    }

    /** PagesSerde.serialize(boolean, Chunk): a local chunk travels as the object, a remote one as bytes. */
    public synchronized SerializedChunk serialize(boolean localChunk, Chunk page) {
        if (localChunk) {
            return new SerializedChunk(page, (int) page.getSizeInBytes(), ChunkCompression.UNCOMPRESSED, page.getPositionCount());
        }
        GpuNative.stagingReset(in);
        GpuChunks.append(in, page, typeCodes);
        return split(GpuNative.serdeSerialize(ctx, in, Math.max(page.getPositionCount(), 1))).get(0);
    }

    /**
     * The batched form the GPU operators use: every row gathered in `staging` leaves as pages of pageRows rows, one kernel
     * launch for all of them.
     */
    public synchronized List<SerializedChunk> serializeStaging(long staging, int pageRows) {
        return split(GpuNative.serdeSerialize(ctx, staging, pageRows));
    }

    private static List<SerializedChunk> split(byte[] framed) {
        List<SerializedChunk> pages = new ArrayList<>();
        Slice all = Slices.wrappedBuffer(framed);
        int at = 0;
        while (at < framed.length) {
            int positionCount = all.getInt(at); // Slice is little-endian, like SliceOutput.writeInt on the Java side
            int uncompressedSize = all.getInt(at + 5);
            int sizeInBytes = all.getInt(at + 9);
            pages.add(new SerializedChunk(all.slice(at + FRAME_BYTES, sizeInBytes), ChunkCompression.UNCOMPRESSED, positionCount, uncompressedSize));
            at += FRAME_BYTES + sizeInBytes;
        }
        return pages;
    }

    /** PagesSerde.deserialize(SerializedChunk). */
    public synchronized Chunk deserialize(SerializedChunk serializedChunk) {
        if (serializedChunk.getPage() != null) {
            return serializedChunk.getPage();
        }
        if (serializedChunk.getCompression() != ChunkCompression.UNCOMPRESSED) {
            return stock.deserialize(serializedChunk);
        }
        Slice raw = serializedChunk.getSlice();
        byte[] framed = new byte[FRAME_BYTES + raw.length()];
        Slice f = Slices.wrappedBuffer(framed);
        f.setInt(0, serializedChunk.getPositionCount());
        f.setByte(4, 0);
        f.setInt(5, serializedChunk.getUncompressedSizeInBytes());
        f.setInt(9, raw.length());
        f.setBytes(FRAME_BYTES, raw);
        int rows = GpuNative.serdeDeserialize(ctx, framed, 0, framed.length, out);
        return GpuChunks.toChunk(out, types, 0, rows);
    }

    @Override
    public synchronized void close() {
        if (in != 0) {
            GpuNative.stagingDestroy(in);
            GpuNative.stagingDestroy(out);
            GpuNative.ctxDestroy(ctx);
            in = out = 0;
        }
    }
}
