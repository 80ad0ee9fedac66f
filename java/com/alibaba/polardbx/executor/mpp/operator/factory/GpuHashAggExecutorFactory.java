/*
 * GPU twin of HashAggExecutorFactory (mpp/operator/factory/HashAggExecutorFactory.java:49-106): same constructor
 * arguments (minus the spiller: the GPU path does not spill) and the same expected-group clamp; selected in
 * LocalExecutionPlanner.visitHashAgg when GpuSupport.aggSupported(...) holds.
 */
package com.alibaba.polardbx.executor.mpp.operator.factory;

import com.alibaba.polardbx.executor.operator.Executor;
import com.alibaba.polardbx.executor.operator.GpuHashAggExec;
import com.alibaba.polardbx.executor.operator.gpu.GpuAggSpec;
import com.alibaba.polardbx.executor.utils.RuntimeStatHelper;
import com.alibaba.polardbx.optimizer.context.ExecutionContext;
import com.alibaba.polardbx.optimizer.core.datatype.DataType;
import com.alibaba.polardbx.optimizer.core.rel.HashAgg;
import com.alibaba.polardbx.optimizer.utils.CalciteUtils;

import java.util.ArrayList;
import java.util.List;

public class GpuHashAggExecutorFactory extends ExecutorFactory {
    private static final int MIN_HASH_TABLE_SIZE = 1024, MAX_HASH_TABLE_SIZE = 131064; // HashAggExecutorFactory.java:76-84

    private final HashAgg hashAgg;
    private final int parallelism, taskNumber;
    private final Integer rowCount;
    private final List<DataType> inputDataTypes;
    private final List<Executor> executors = new ArrayList<>();

    public GpuHashAggExecutorFactory(HashAgg hashAgg, int parallelism, int taskNumber, Integer rowCount,
                                     List<DataType> inputDataTypes) {
        this.hashAgg = hashAgg;
        this.parallelism = parallelism;
        this.taskNumber = taskNumber;
        this.rowCount = rowCount;
        this.inputDataTypes = inputDataTypes;
    }

    @Override
    public Executor createExecutor(ExecutionContext context, int index) {
        return createAllExecutors(context).get(index);
    }

    @Override
    public List<Executor> getAllExecutors(ExecutionContext context) {
        return createAllExecutors(context);
    }

    private synchronized List<Executor> createAllExecutors(ExecutionContext context) {
        if (executors.isEmpty()) {
            int[] groups = HashAggExecutorFactory.convertFrom(hashAgg.getGroupSet());
            int expected = rowCount == null ? MIN_HASH_TABLE_SIZE : rowCount / (taskNumber * parallelism);
            expected = Math.max(MIN_HASH_TABLE_SIZE, Math.min(MAX_HASH_TABLE_SIZE, expected));
            GpuAggSpec spec = GpuAggSpec.tryConvert(hashAgg.getAggCallList(), inputDataTypes); // non-null: GpuSupport
            List<DataType> outputDataTypes = CalciteUtils.getTypes(hashAgg.getRowType());
            for (int j = 0; j < parallelism; j++) {
                GpuHashAggExec exec = new GpuHashAggExec(inputDataTypes, groups, spec, outputDataTypes, expected, context);
                exec.setId(hashAgg.getRelatedId());
                if (context.getRuntimeStatistics() != null) {
                    RuntimeStatHelper.registerStatForExec(hashAgg, exec, context);
                }
                executors.add(exec);
            }
        }
        return executors;
    }
}
