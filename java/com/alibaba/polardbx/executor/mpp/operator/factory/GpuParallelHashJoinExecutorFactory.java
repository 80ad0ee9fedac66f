/*
 * GPU twin of ParallelHashJoinExecutorFactory (mpp/operator/factory/ParallelHashJoinExecutorFactory.java:52-118): same
 * constructor arguments, same executor-per-probe-driver shape; selected in LocalExecutionPlanner.visitHashJoin when
 * GpuSupport.joinSupported(...) holds (INTEGRATION.md shows the patch).
 */
package com.alibaba.polardbx.executor.mpp.operator.factory;

import com.alibaba.polardbx.executor.operator.Executor;
import com.alibaba.polardbx.executor.operator.GpuParallelHashJoinExec;
import com.alibaba.polardbx.executor.operator.gpu.GpuJoinCondition;
import com.alibaba.polardbx.executor.operator.util.EquiJoinUtils;
import com.alibaba.polardbx.executor.utils.RuntimeStatHelper;
import com.alibaba.polardbx.optimizer.context.ExecutionContext;
import com.alibaba.polardbx.optimizer.core.datatype.DataType;
import com.alibaba.polardbx.optimizer.core.join.EquiJoinKey;
import com.alibaba.polardbx.optimizer.utils.CalciteUtils;
import org.apache.calcite.rel.core.Join;
import org.apache.calcite.rel.core.JoinRelType;
import org.apache.calcite.rex.RexCall;
import org.apache.calcite.rex.RexInputRef;
import org.apache.calcite.rex.RexNode;

import java.util.ArrayList;
import java.util.List;

public class GpuParallelHashJoinExecutorFactory extends ExecutorFactory {
    private final Join join;
    private final RexNode otherCond, equalCond;
    private final boolean maxOneRow, driverBuilder;
    private final List<RexNode> operands;
    private final int probeParallelism;
    private final List<Executor> executors = new ArrayList<>();

    public GpuParallelHashJoinExecutorFactory(Join join, RexNode otherCond, RexNode equalCond, boolean maxOneRow,
                                              List<RexNode> operands, ExecutorFactory build, ExecutorFactory probe,
                                              int probeParallelism, int numPartitions, boolean driverBuilder) {
        this.join = join;
        this.otherCond = otherCond;
        this.equalCond = equalCond;
        this.maxOneRow = maxOneRow;
        this.operands = operands;
        this.probeParallelism = probeParallelism;
        this.driverBuilder = driverBuilder;
        addInput(build);
        addInput(probe);
    }

    @Override
    public Executor createExecutor(ExecutionContext context, int index) {
        return createAllExecutor(context).get(index);
    }

    @Override
    public List<Executor> getAllExecutors(ExecutionContext context) {
        return createAllExecutor(context);
    }

    private synchronized List<Executor> createAllExecutor(ExecutionContext context) {
        if (executors.isEmpty()) {
            GpuParallelHashJoinExec.GpuJoinShared shared = new GpuParallelHashJoinExec.GpuJoinShared(probeParallelism);
            List<EquiJoinKey> joinKeys = EquiJoinUtils
                .buildEquiJoinKeys(join, join.getOuter(), join.getInner(), (RexCall) equalCond, join.getJoinType());
            GpuJoinCondition condition = GpuJoinCondition.tryConvert(otherCond); // convertible: checked by GpuSupport
            int[] antiOperands = null;
            if (operands != null && join.getJoinType() == JoinRelType.ANTI && !operands.isEmpty()) {
                antiOperands = operands.stream().mapToInt(o -> ((RexInputRef) o).getIndex()).toArray();
            }
            List<DataType> dataTypes = CalciteUtils.getTypes(join.getRowType());
            for (int i = 0; i < probeParallelism; i++) {
                Executor inner, outerInput;
                if (driverBuilder) {
                    outerInput = getInputs().get(0).createExecutor(context, i);
                    inner = getInputs().get(1).createExecutor(context, i);
                } else {
                    inner = getInputs().get(0).createExecutor(context, i);
                    outerInput = getInputs().get(1).createExecutor(context, i);
                }
                GpuParallelHashJoinExec exec = new GpuParallelHashJoinExec(shared, outerInput, inner, join.getJoinType(),
                    maxOneRow, joinKeys, condition, antiOperands, driverBuilder, dataTypes, context);
                exec.setId(join.getRelatedId());
                if (context.getRuntimeStatistics() != null) {
                    RuntimeStatHelper.registerStatForExec(join, exec, context);
                }
                executors.add(exec);
            }
        }
        return executors;
    }
}
