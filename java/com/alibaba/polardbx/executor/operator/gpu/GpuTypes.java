package com.alibaba.polardbx.executor.operator.gpu;

import com.alibaba.polardbx.optimizer.core.datatype.DataType;
import com.alibaba.polardbx.optimizer.core.datatype.DataTypeUtil;
import com.alibaba.polardbx.optimizer.core.datatype.DataTypes;
import org.apache.calcite.rel.core.JoinRelType;

import java.util.List;

/** DataType / JoinRelType <-> gsql_type / gsql_join_type (include/gsql_gpu.h). */
public final class GpuTypes {
    private GpuTypes() {
    }

    /** -1 when the type has no GPU block form (the planner keeps the stock operator). */
    public static int code(DataType type) {
        if (DataTypeUtil.equalsSemantically(type, DataTypes.IntegerType)) {
            return GpuNative.T_INT32; // IntegerBlock
        }
        if (DataTypeUtil.equalsSemantically(type, DataTypes.LongType)) {
            return GpuNative.T_INT64; // LongBlock
        }
        if (DataTypeUtil.equalsSemantically(type, DataTypes.DoubleType)) {
            return GpuNative.T_FP64; // DoubleBlock
        }
        if (isPackedTime(type)) {
            return GpuNative.T_INT64; // DateBlock / TimestampBlock: long[] packed
        }
        return -1;
    }

    /**
     * DATE and DATETIME / TIMESTAMP columns: DateBlock and TimestampBlock hold one MySQL packed long per row
     * (chunk/DateBlock.java:43, chunk/TimestampBlock.java:47; BlockBuilders.java:73-76 picks them by data class), hash it
     * with Long.hashCode like LongBlock (DateBlock.java:138-142, TimestampBlock.java:105-109) and compare rows by that long
     * (DateBlock.java:174-191).  On the GPU such a column IS a BIGINT column: join keys, group keys, exchange partitioning
     * and pass-through are bit-identical; GpuChunks rebuilds the Block class from the DataType on the way back.
     * Arithmetic and comparisons with literals are NOT served (a literal would have to be packed the same way):
     * GpuExpression only lets such a column through as a bare input reference.
     */
    public static boolean isPackedTime(DataType type) {
        Class<?> clazz = type.getDataClass();
        return clazz == java.sql.Date.class || clazz == java.sql.Timestamp.class;
    }

    /** INT / BIGINT / DOUBLE: the types expressions may compute with. */
    public static boolean isNumeric(DataType type) {
        return code(type) >= 0 && !isPackedTime(type);
    }

    public static boolean supported(List<DataType> types) {
        for (DataType t : types) {
            if (code(t) < 0) {
                return false;
            }
        }
        return true;
    }

    public static int[] codes(List<DataType> types) {
        int[] out = new int[types.size()];
        for (int i = 0; i < out.length; i++) {
            out[i] = code(types.get(i));
            if (out[i] < 0) {
                throw new GpuExecutorException("no GPU block type for " + types.get(i));
            }
        }
        return out;
    }

    /** gsql_join_type; -1 = not served. */
    public static int joinType(JoinRelType t) {
        switch (t) {
        case INNER:
            return 0;
        case LEFT:
            return 1;
        case RIGHT:
            return 2;
        case SEMI:
            return 3;
        case ANTI:
            return 4;
        default:
            return -1;
        }
    }
}
