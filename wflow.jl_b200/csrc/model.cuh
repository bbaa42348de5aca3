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
};

// One routing domain (land or river) as the wavefront kernels see it. Slots are ordered by
// (chunk, level, node id); a chunk is a set of connected pieces of the drainage forest (one
// outlet node each) with at most WFB_CHUNK_NODES = 32 nodes in total: one WARP walks it, one
// lane per node.
#define WFB_CHUNK_NODES 32
// piece depth of the chunks (network.hpp: build_chunks; 0 = one connected piece per chunk):
// see api.cu: build_networks for the measurements behind these values
#define WFB_PIECE_DEPTH_LAND 0
#define WFB_PIECE_DEPTH_LAND_WIDE 6   // land domains with >= WFB_PIECE_WIDE_LEVEL nodes per level
#define WFB_PIECE_WIDE_LEVEL 2048
#define WFB_PIECE_DEPTH_RIVER 0
// the subsurface sweep and the surface kernel run overlapped below this mean level width
#define WFB_OVERLAP_MAX_LEVEL_WIDTH 1750
#define WFB_NO_EDGE 0xffu
struct DevNet {
  int32_t n;                    // nodes
  int32_t n_levels;             // wavefront levels of the whole domain
  int32_t n_chunks;
  int32_t max_inlets;           // largest number of inlet edges of any chunk
  int32_t n_outlets;            // publishing piece roots of the whole domain
  const int4* chunk_meta;       // per chunk: x = first slot, y = nodes | levels << 8, z = offset
                                // of its inlet list, w = number of inlet edges
  const int32_t* node_out;      // per slot: outlet number if the node drains into another chunk
                                // (it publishes its discharge of every sub-step), else -1
  const unsigned long long* node_edges;  // per slot: up to 8 upstream sources, one byte each,
                                // ordered by ascending upstream NODE ID (the reference's
                                // left-fold order, utils.jl:472-477): lane of the source inside
                                // the chunk (< 32), 32 + k for the chunk's k-th inlet edge, or
                                // WFB_NO_EDGE
  const uint8_t* node_level;    // per slot: level inside its chunk (0 = the chunk's first level)
  const int32_t* inl_src;       // per inlet edge: outlet number of the producer
  const uint8_t* inl_level;     // per inlet edge: level (inside the chunk) of the receiving node
};

// The land domain as the single-sub-step subsurface kernel sees it (network.hpp: bands,
// fragments, bundles): one warp walks a bundle row by row, one lane per node of the row.
#define WFB_BAND_DEPTH 4
struct DevBands {
  int32_t n_bundles, n_outlets, max_inlets;
  const int32_t* slot;     // n_bundles * WFB_BAND_DEPTH * 32: land slot or -1
  const uint4* src;        // per entry: 8 x 16-bit upstream sources in ascending node id: lane in
                           // the previous row (< 32), 0x8000 | k for the bundle's k-th inlet,
                           // 0xffff none
  const int32_t* out;      // per entry: outlet number if the node feeds another bundle, else -1
  const int32_t* inl_ptr;  // n_bundles + 1
  const int32_t* inl_out;  // per inlet: outlet number of the producer
};

struct KCfg {
  int32_t n, nriv, ns, nrs;
  int32_t gash, has_lai, snow, glacier, soil_infiltration_reduction, kv_profile;
  double qroot;                // KIN_WAVE_MIN_FLOW^0.2
};

}  // namespace wfb
