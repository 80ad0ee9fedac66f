package com.alibaba.polardbx.executor.operator.gpu;

import com.alibaba.polardbx.common.exception.TddlRuntimeException;
import com.alibaba.polardbx.common.exception.code.ErrorCode;

/** Any non-zero gsql_status other than GSQL_E_MORE_THAN_ONE_ROW; thrown by the JNI shim with gsql_last_error as message. */
public class GpuExecutorException extends TddlRuntimeException {
    public GpuExecutorException(String message) {
        super(ErrorCode.ERR_EXECUTOR, message);
    }
}
