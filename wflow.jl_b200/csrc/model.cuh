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

// One routing domain (land or river) as the wavefront kernels see it. Slots are ordered by
// (chunk, level, node id); a chunk is a connected piece of the drainage forest with ONE outlet
// node, walked by one CTA.
struct DevNet {
  int32_t n;                    // nodes
  int32_t n_levels;             // wavefront levels of the whole domain
  int32_t n_chunks;
  const int32_t* level_of;      // slot -> level
  const int32_t* up_ptr;        // slot -> CSR offsets of its upstream edges
  const int32_t* up_idx;        // upstream SLOTS, ordered by ascending NODE ID (the reference's
                                // left-fold order, utils.jl:472-477)
  const int32_t* up_chunk;      // per edge: producer chunk if the upstream node is another
                                // chunk's outlet, else -1
  const int32_t* chunk_ptr;     // n_chunks + 1 slot offsets
  const int32_t* chunk_l0;      // first level of a chunk
  const int32_t* chunk_l1;      // last level (= level of its outlet node)
  const int32_t* chunk_clp_off; // offsets into clp
  const int32_t* clp;           // per chunk: absolute slot offsets of its levels (l1-l0+2 entries)
  const int32_t* chunk_inl_ptr; // n_chunks + 1 offsets into the inlet lists
  const int32_t* inl_level;     // level of the receiving node of an inlet edge
  const int32_t* inl_src;       // producer chunk of an inlet edge
  const int32_t* outlet_chunk;  // slot -> chunk id if the slot is a chunk outlet that feeds
                                // another chunk, else -1
};

struct KCfg {
  int32_t n, nriv, ns, nrs;
  int32_t gash, has_lai, snow, glacier, soil_infiltration_reduction, kv_profile;
  double qroot;                // KIN_WAVE_MIN_FLOW^0.2
};

}  // namespace wfb
