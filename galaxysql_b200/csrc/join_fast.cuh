// join_fast.cuh — single-integer-key specialisation of the join (filled in by join_fast.cu).
#pragma once
#include "common.cuh"

struct JoinFast {
    bool enabled = false;
};
