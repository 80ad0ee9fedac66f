package com.alibaba.polardbx.executor.operator.gpu;

import com.alibaba.polardbx.common.exception.TddlRuntimeException;
import com.alibaba.polardbx.common.exception.code.ErrorCode;

/** GSQL_E_MORE_THAN_ONE_ROW: a single (scalar sub-query) join met a second match (AbstractBufferedJoinExec.java:217-219). */
public class GpuMoreThanOneRowException extends TddlRuntimeException {
    public GpuMoreThanOneRowException(String message) {
        super(ErrorCode.ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW);
    }
}
