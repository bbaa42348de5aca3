// model.cuh -- device-side view of the model state (structure of arrays in HBM).
//
// Layout: every field is one contiguous Float64 array of `ns` (land) or `nrs` (river) slots,
// ns/nrs = n/nriv rounded up to 32 doubles so that each layer slab of a layered field starts
// 256-byte aligned. Layered fields are LAYER-major: layer k of slot p at [k*ns + p] (the
// transpose of Julia's Vector{SVector{N}}; the ABI set/get transposes). Land slots are ordered
// by topological-depth level of the land drainage forest (then node id), river slots by the
// level of the river forest, so a wavefront stage touches a contiguous slot range.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/wflow_b200.h"

namespace wfb {

struct DevFields {
#define X(name, kind) double* name;
  WFLOWB200_FIELDS(X)
#undef X
  int32_t* number_of_layers;   // land
  int32_t* n_unsatlayers;      // land
  int32_t* riv_land_slot;      // river slot -> land slot
  // double-buffered discharges of the skewed wavefront (see routing.cu); buffer 0 aliases the
  // canonical field when `*_phase` == 0
  double* olf_q2;
  double* riv_q2;
  double* ssf_q2;
};

struct DevNet {
  int32_t n;                   // nodes
  int32_t n_levels;            // wavefront levels
  const int32_t* level_ptr;    // n_levels + 1 slot offsets
  const int32_t* level_of;     // slot -> level
  const int32_t* up_ptr;       // slot -> CSR offsets of upstream SLOTS
  const int32_t* up_idx;       // upstream slots, ordered by ascending NODE ID (the reference's
                               // left-fold order, utils.jl:472-477)
};

struct KCfg {
  int32_t n, nriv, ns, nrs;
  int32_t gash, has_lai, snow, glacier, soil_infiltration_reduction, kv_profile;
  double qroot;                // KIN_WAVE_MIN_FLOW^0.2
};

}  // namespace wfb
