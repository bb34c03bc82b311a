// train_simt.cuh -- loss / backward / Adam kernels (filled in after the forward path).
#pragma once
#include "common.cuh"
namespace cvb {
struct TrainWork {};
static inline void train_work_free(TrainWork* w) { delete w; }
}  // namespace cvb
