package com.alibaba.polardbx.executor.operator.gpu;

import com.alibaba.polardbx.optimizer.context.ExecutionContext;

import java.util.concurrent.atomic.AtomicInteger;

/** Which GPU an operator instance runs on: instances of one query are spread round-robin over the visible devices. */
public final class GpuDevices {
    private static final AtomicInteger NEXT = new AtomicInteger();
    private static volatile int count = -1;

    private GpuDevices() {
    }

    public static int count() {
        int c = count;
        if (c < 0) {
            try {
                c = GpuNative.deviceCount();
            } catch (UnsatisfiedLinkError e) {
                c = 0; // libgsql_jni.so not installed: the planner keeps the stock operators
            }
            count = c;
        }
        return c;
    }

    public static int deviceForThisDriver(ExecutionContext context) {
        int n = count();
        if (n <= 0) {
            throw new GpuExecutorException("no CUDA device: the planner must not select GPU operators");
        }
        return Math.floorMod(NEXT.getAndIncrement(), n);
    }
}
