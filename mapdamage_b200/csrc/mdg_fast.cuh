// Fast counting path (placeholder until the slot-owner kernel lands).
#pragma once
#include "../../include/mapdamage_b200.h"
#include "mdg_device.cuh"

namespace mdg {

struct FastPlan {
    bool enabled = false;
};

inline cudaError_t plan_fast(FastPlan &, const mdg_config &, int, size_t) { return cudaSuccess; }
inline int launch_fast(const FastPlan &, const DevBatch &, const DevRef &, const CountParams &, const CountTables &,
                       cudaStream_t, int64_t *)
{
    return 0;
}
inline void free_fast(FastPlan &) {}

}  // namespace mdg
