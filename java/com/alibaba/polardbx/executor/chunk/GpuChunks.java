package com.alibaba.polardbx.executor.chunk;

import com.alibaba.polardbx.common.utils.timezone.InternalTimeZone;
import com.alibaba.polardbx.executor.operator.gpu.GpuExecutorException;
import com.alibaba.polardbx.executor.operator.gpu.GpuNative;
import com.alibaba.polardbx.executor.operator.gpu.GpuTypes;
import com.alibaba.polardbx.optimizer.core.datatype.DataType;

import java.util.List;

/**
 * Chunk <-> staging.  Lives in the chunk package because a Block's window onto its backing array is package-private
 * (AbstractBlock.arrayOffset, AbstractBlock.java:29) — the same reason the *BlockEncoding classes live here.
 */
public final class GpuChunks {
    private GpuChunks() {
    }

    /**
     * Appends a chunk to a staging batch (copies; no reference to the chunk survives the call).  Blocks expose their
     * backing arrays (IntegerBlock.intArray():217, LongBlock.longArray():191, DoubleBlock.doubleArray():190,
     * AbstractBlock.nulls():116): one GetPrimitiveArrayCritical + memcpy per column from arrayOffset.  A chunk that
     * carries a selection vector (Chunk.java:57-79) is gathered through it on the native side; a column that is not
     * one of the three primitive blocks, or an IntegerBlock with its own selection (IntegerBlock.java:38,69-74), is
     * materialised position by position through the Block interface first.  typeCodes = GpuTypes.codes(input types).
     */
    public static void append(long staging, Chunk chunk, int[] typeCodes) {
        final int n = chunk.getPositionCount();
        if (n == 0) {
            return;
        }
        final int cols = chunk.getBlockCount();
        final int[] selection = chunk.isSelectionInUse() ? chunk.selection() : null;
        Object[] arrays = new Object[cols];
        boolean[][] nulls = new boolean[cols][];
        int offset = -1;
        for (int c = 0; c < cols; c++) {
            Block b = chunk.getBlock(c);
            int blockOffset = 0;
            boolean direct = true;
            if (b instanceof IntegerBlock && ((IntegerBlock) b).getSelection() == null) {
                arrays[c] = ((IntegerBlock) b).intArray();
            } else if (b instanceof LongBlock) {
                arrays[c] = ((LongBlock) b).longArray();
            } else if (b instanceof DoubleBlock) {
                arrays[c] = ((DoubleBlock) b).doubleArray();
            } else if (b instanceof TimestampBlock) {
                arrays[c] = ((TimestampBlock) b).getPacked(); // DATETIME / TIMESTAMP: MySQL packed longs, a BIGINT column to the GPU
            } else if (b instanceof DateBlock && ((DateBlock) b).getSelection() == null) {
                arrays[c] = ((DateBlock) b).getPacked();
            } else {
                direct = false;
            }
            if (direct) {
                AbstractBlock ab = (AbstractBlock) b;
                blockOffset = ab.arrayOffset;
                nulls[c] = ab.mayHaveNull() ? ab.nulls() : null;
                if (offset < 0) {
                    offset = blockOffset;
                }
            }
            if (!direct || blockOffset != offset) {
                // materialise this column so that it starts at the common offset (rare: sliced or foreign blocks)
                materialise(b, typeCodes[c], c, n, selection, arrays, nulls, Math.max(offset, 0));
            }
        }
        if (offset < 0) {
            offset = 0;
        }
        if (selection != null) {
            GpuNative.stagingAppendSelected(staging, arrays, nulls, selection, n);
        } else {
            GpuNative.stagingAppend(staging, arrays, nulls, offset, n);
        }
    }

    private static void materialise(Block b, int code, int c, int n, int[] selection, Object[] arrays, boolean[][] nulls,
                                    int offset) {
        // when a selection is in use the native side indexes with selection[i]: lay the values out at those indexes
        int span = offset + n;
        if (selection != null) {
            for (int s : selection) {
                span = Math.max(span, s + 1);
            }
        }
        boolean[] nl = null;
        Object out;
        if (code == GpuNative.T_INT32) {
            int[] v = new int[span];
            for (int i = 0; i < n; i++) {
                int at = selection != null ? selection[i] : offset + i;
                if (b.isNull(i)) {
                    (nl == null ? nl = new boolean[span] : nl)[at] = true;
                } else {
                    v[at] = b.getInt(i);
                }
            }
            out = v;
        } else if (code == GpuNative.T_FP64) {
            double[] v = new double[span];
            for (int i = 0; i < n; i++) {
                int at = selection != null ? selection[i] : offset + i;
                if (b.isNull(i)) {
                    (nl == null ? nl = new boolean[span] : nl)[at] = true;
                } else {
                    v[at] = b.getDouble(i);
                }
            }
            out = v;
        } else {
            long[] v = new long[span];
            for (int i = 0; i < n; i++) {
                int at = selection != null ? selection[i] : offset + i;
                if (b.isNull(i)) {
                    (nl == null ? nl = new boolean[span] : nl)[at] = true;
                } else if (b instanceof DateBlock) {
                    v[at] = ((DateBlock) b).getPackedLong(i); // (honours the block's own selection)
                } else if (b instanceof TimestampBlock) {
                    v[at] = ((TimestampBlock) b).getPackedLong(i);
                } else {
                    v[at] = b.getLong(i);
                }
            }
            out = v;
        }
        arrays[c] = out;
        nulls[c] = nl;
    }

    /** Rows [from, from+rows) of a staging batch as a fresh Chunk of IntegerBlock / LongBlock / DoubleBlock / DateBlock / TimestampBlock. */
    public static Chunk toChunk(long staging, List<DataType> types, int from, int rows) {
        Block[] blocks = new Block[types.size()];
        for (int c = 0; c < blocks.length; c++) {
            Object values = GpuNative.stagingColumn(staging, c, from, rows);
            boolean[] nulls = GpuNative.stagingNulls(staging, c, from, rows); // null = no NULL among these rows
            switch (GpuTypes.code(types.get(c))) {
            case GpuNative.T_INT32:
                blocks[c] = new IntegerBlock(0, rows, nulls, (int[]) values);
                break;
            case GpuNative.T_INT64: {
                DataType type = types.get(c);
                Class<?> clazz = type.getDataClass();
                if (clazz == java.sql.Date.class) { // as DateBlockBuilder.build():146-149 / TimestampBlockBuilder.build()
                    blocks[c] = new DateBlock(0, rows, nulls, (long[]) values, type, InternalTimeZone.DEFAULT_TIME_ZONE);
                } else if (clazz == java.sql.Timestamp.class) {
                    blocks[c] = new TimestampBlock(0, rows, nulls, (long[]) values, type, InternalTimeZone.DEFAULT_TIME_ZONE);
                } else {
                    blocks[c] = new LongBlock(0, rows, nulls, (long[]) values);
                }
                break;
            }
            case GpuNative.T_FP64:
                blocks[c] = new DoubleBlock(0, rows, nulls, (double[]) values);
                break;
            default:
                throw new GpuExecutorException("no Block form for output column " + c);
            }
        }
        return new Chunk(rows, blocks);
    }
}
