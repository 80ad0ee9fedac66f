package com.alibaba.polardbx.executor.operator.gpu;

import com.alibaba.polardbx.common.properties.ConnectionParams;
import com.alibaba.polardbx.optimizer.context.ExecutionContext;
import com.alibaba.polardbx.optimizer.core.datatype.DataType;
import com.alibaba.polardbx.optimizer.core.join.EquiJoinKey;
import com.alibaba.polardbx.optimizer.core.rel.HashAgg;
import com.alibaba.polardbx.optimizer.utils.CalciteUtils;
import org.apache.calcite.rel.core.Join;
import org.apache.calcite.rel.core.JoinRelType;
import org.apache.calcite.rex.RexNode;

import java.util.List;

/**
 * Planner-time decisions (LocalExecutionPlanner.visitHashJoin:852 / visitHashAgg:1487): a GPU factory is chosen only
 * when the whole operator can run on the device; otherwise the stock factory is used.  There is no CPU fallback
 * inside a GPU operator.
 */
public final class GpuSupport {
    private GpuSupport() {
    }

    public static boolean enabled(ExecutionContext context) {
        return context.getParamManager().getBoolean(ConnectionParams.ENABLE_GPU_OPERATORS) && GpuDevices.count() > 0;
    }

    public static boolean joinSupported(Join join, List<EquiJoinKey> keys, RexNode otherCond, boolean maxOneRow,
                                        List<RexNode> antiOperands, ExecutionContext context) {
        if (!enabled(context)) {
            return false;
        }
        JoinRelType t = join.getJoinType();
        if (GpuTypes.joinType(t) < 0 || keys.isEmpty() || keys.size() > 8) {
            return false;
        }
        if ((t == JoinRelType.SEMI || t == JoinRelType.ANTI) && maxOneRow) {
            return false; // single semi/anti joins: GSQL_E_UNSUPPORTED at create
        }
        if (!GpuTypes.supported(CalciteUtils.getTypes(join.getOuter().getRowType()))
            || !GpuTypes.supported(CalciteUtils.getTypes(join.getInner().getRowType()))) {
            return false;
        }
        for (EquiJoinKey k : keys) {
            if (GpuTypes.code(k.getUnifiedType()) < 0 || k.isNullSafeEqual()) {
                return false;
            }
        }
        if (antiOperands != null) {
            for (RexNode o : antiOperands) {
                if (!(o instanceof org.apache.calcite.rex.RexInputRef)) {
                    return false; // NOT IN over expressions: stock operator
                }
            }
        }
        return GpuJoinCondition.convertible(otherCond);
    }

    public static boolean aggSupported(HashAgg agg, List<DataType> inputTypes, ExecutionContext context) {
        if (!enabled(context) || agg.getGroupSet().cardinality() > 8) {
            return false;
        }
        if (!GpuTypes.supported(CalciteUtils.getTypes(agg.getRowType()))) {
            return false; // e.g. SUM(BIGINT) -> DECIMAL output
        }
        for (int g : agg.getGroupSet()) {
            if (GpuTypes.code(inputTypes.get(g)) < 0) {
                return false;
            }
        }
        GpuAggSpec spec = GpuAggSpec.tryConvert(agg.getAggCallList(), inputTypes);
        return spec != null && !spec.producesDecimal(inputTypes);
    }
}
