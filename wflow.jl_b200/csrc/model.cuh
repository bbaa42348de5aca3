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
  const double* fp_depth;      // floodplain profile depths (cfg.fp_levels values), device
  int32_t* nlayers_kv;         // land (KvLayeredExponential only)
  int32_t* riv_reservoir;      // river slot -> reservoir (0-based) or -1; nullptr without reservoirs
  int32_t* res_land_slot;      // reservoir -> land slot of its outlet cell
  int32_t* res_river_slot;     // reservoir -> river slot of its node
  // local-inertial river flow: the staggered grid by river slot
  int32_t* li_dst_slot;        // slot of the node the leaving edge ends in, -1 none, -2 ghost
  int32_t* li_in_ptr;          // edges entering a node (= their source slots), CSR by slot,
  int32_t* li_in_idx;          // ascending source NODE id (sum_at order)
  // 2-D local-inertial overland flow: EdgeConnectivity by land slot (-1: no active neighbour) and
  // domain.land.network.river_indices (land slot -> river slot, -1: no river in the cell)
  // (all by NODE id: the 2-D state li_land_* is kept in node order, see local_inertial.cu)
  int32_t *edge_x_up, *edge_x_down, *edge_y_up, *edge_y_down;
  int32_t* lil_river_slot;     // node -> river slot or -1
  int32_t* land_slot_of_node;  // node -> land slot (olf_h / olf_storage, the vertical fields)
  int32_t* res_land_node;      // reservoir -> node of its outlet cell
  double *lil_h, *lil_storage; // work arrays: h and storage by node during a model step,
  double* lil_cell_area;       // x_length * y_length,
  int32_t *lil_xu_eff, *lil_yu_eff;  // edge_x_up / edge_y_up, -1 also where the flow width is zero
  uint8_t* land_is_res_outlet; // land slot is a reservoir outlet (nullptr without reservoirs)
  int32_t* olf_newton_trace;   // land / river, or nullptr: Newton iterations of kinematic_wave
  int32_t* riv_newton_trace;   // per node since wflowb200_newton_trace(h, 1)
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
  // cut edges of a shard that is part of a basin (wflowb200_exchange_*): an inlet edge with
  // inl_src >= n_outlets is IMPORT inl_src - n_outlets (its slots live in WaveLaunch::imports and
  // are written by another GPU); node_export: per slot, the export the node feeds or -1 (nullptr:
  // the shard exports nothing)
  const int32_t* node_export;
};

struct KCfg {
  int32_t n, nriv, ns, nrs, nres;
  int32_t gash, has_lai, snow, glacier, soil_infiltration_reduction, kv_profile;
  double qroot;                // KIN_WAVE_MIN_FLOW^0.2
  int32_t river_routing;       // 0 kinematic wave, 1 local inertial
  int32_t land_routing;        // 0 kinematic wave, 1 local inertial (2-D, with river_routing = 1)
  int32_t kw_root_each_substep; // 1: u_prev = pow(q_prev, 0.2) before every solve, like the
                               // reference (default 0: carried, see routing.cu: KwState)
  int32_t fp_levels;           // 1-D floodplain: levels of the profile (0: none) and their depths
  double fp_depth[16];
};

}  // namespace wfb
