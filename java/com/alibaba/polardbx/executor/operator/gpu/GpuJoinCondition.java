package com.alibaba.polardbx.executor.operator.gpu;

import org.apache.calcite.rex.RexCall;
import org.apache.calcite.rex.RexInputRef;
import org.apache.calcite.rex.RexLiteral;
import org.apache.calcite.rex.RexNode;
import org.apache.calcite.sql.SqlKind;

import java.util.ArrayList;
import java.util.List;

/**
 * The restricted otherCondition of gsql_join_spec: AND_i (joinRow[cols[i]] IS NULL OR joinRow[cols[i]] <> neValues[i])
 * over integer columns of the join row.  tryConvert returns null for anything else, and the planner then keeps the
 * stock ParallelHashJoinExec (AbstractJoinExec.java:227-250 evaluates arbitrary IExpressions row by row).
 */
public final class GpuJoinCondition {
    public final int[] cols;
    public final long[] neValues;

    private GpuJoinCondition(int[] cols, long[] neValues) {
        this.cols = cols;
        this.neValues = neValues;
    }

    /** otherCond == null means "no condition" and is represented by a null GpuJoinCondition: check `convertible` first. */
    public static boolean convertible(RexNode otherCond) {
        return otherCond == null || tryConvert(otherCond) != null;
    }

    public static GpuJoinCondition tryConvert(RexNode otherCond) {
        if (otherCond == null) {
            return null;
        }
        List<RexNode> conjuncts = new ArrayList<>();
        flattenAnd(otherCond, conjuncts);
        if (conjuncts.isEmpty() || conjuncts.size() > 4) {
            return null;
        }
        int[] cols = new int[conjuncts.size()];
        long[] values = new long[conjuncts.size()];
        for (int i = 0; i < cols.length; i++) {
            RexNode n = conjuncts.get(i);
            if (n.getKind() != SqlKind.NOT_EQUALS) {
                return null;
            }
            RexCall call = (RexCall) n;
            RexNode a = call.getOperands().get(0), b = call.getOperands().get(1);
            if (!(a instanceof RexInputRef) || !(b instanceof RexLiteral)) {
                return null;
            }
            Object v = ((RexLiteral) b).getValue3();
            if (!(v instanceof Number) || ((Number) v).doubleValue() != (double) ((Number) v).longValue()) {
                return null;
            }
            cols[i] = ((RexInputRef) a).getIndex();
            values[i] = ((Number) v).longValue();
        }
        return new GpuJoinCondition(cols, values);
    }

    private static void flattenAnd(RexNode n, List<RexNode> out) {
        if (n.getKind() == SqlKind.AND) {
            for (RexNode o : ((RexCall) n).getOperands()) {
                flattenAnd(o, out);
            }
        } else {
            out.add(n);
        }
    }
}
