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
        return -1;
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
