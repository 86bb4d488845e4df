// placeholder until the register-resident N=512 kernel lands
#pragma once
#include <vector>
#include "common.cuh"
static inline bool fast512_supported(const DevPlan &) { return false; }
static inline int fast512_prepare(DevPlan &, const std::vector<float> &, std::vector<void *> &, int *) { return B200FEAT_EUNSUPPORTED; }
static inline int fast512_launch(const DevPlan &, const DevBatch &, int, int, cudaStream_t) { return 1; }
