// velo_jet.cuh — device-side include of the residual functors (velo_functors.h: dual numbers, per-pose rotation pack, the five
// costfunctions.h functors as linear maps); shared by the visual residual kernel and the fixed-block evaluation of the
// device-resident solve.
#pragma once
#include "velo_common.cuh"
#include "velo_functors.h"
