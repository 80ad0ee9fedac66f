// agg_fast.cuh — specialised group-by kernels (shared-memory privatised low-cardinality path, key-in-slot
// high-cardinality path).  Included by agg.cu after its common definitions.
#pragma once

struct AggFast {
    bool enabled = false;
};
