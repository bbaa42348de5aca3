// network.hpp -- host-side construction of the drainage-network indexing artefacts.
//
// Reproduces, bit-exactly, what the reference builds at model initialisation
// (all paths under /root/reference/Wflow/src):
//   flowgraph                      routing/utils.jl:6-33
//   topological_sort_by_dfs        Graphs.jl 1.14.0 (call sites network.jl:99,244; utils.jl:66)
//   stream_order                   subdomains.jl:32-47
//   kinwave_set_subdomains         subdomains.jl:169-255 (+ subbasins :55-82, fillnodata_upstream
//                                  :8-24, graph_from_nodes :127-143, subbasins_order :93-120)
//   filter_upstream_nodes          utils.jl:61-71 (lists indexed by TOPOSORT POSITION)
// and adds the B200-specific artefact: topological-depth levels (distance to the basin outlet)
// with a level-major device ordering, so that every drainage edge spans exactly one level.
//
// Every graph on this path is a forest with out-degree <= 1, stored as `down` (1-based
// downstream id, 0 = none) plus a CSR of in-neighbours sorted ascending (Graphs.jl keeps
// adjacency lists sorted). Algorithms are O(n) and written for this representation; they are
// not a transcription of Graphs.jl.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace wfb {

struct Network {
  int64_t n = 0;
  std::vector<uint8_t> ldd;            // after the pit fix-up of flowgraph
  std::vector<int64_t> down;           // 1-based downstream node id, 0 = pit / none
  std::vector<int64_t> in_ptr, in_idx; // CSR of in-neighbours (1-based ids, ascending)
  std::vector<int64_t> order;          // topological_sort_by_dfs (1-based ids)
  std::vector<int64_t> streamorder;    // Strahler
  std::vector<int64_t> up_ptr, up_idx; // upstream_nodes by toposort position
  // sub-domain partition (order_of_subdomains / order_subdomain / subdomain_indices)
  std::vector<int64_t> lvl_ptr, lvl_idx, sub_ptr, sub_order, sub_indices;
  // B200 wavefront: topological-depth levels, and the partition of the forest into CHUNKS of
  // at most 32 nodes that one warp walks on its own. A chunk holds one or more PIECES: connected
  // parts of the forest with one outlet node (root) each; the roots that drain into another
  // chunk publish their discharge (outlet numbers, in execution order of the chunks).
  int64_t n_wave_levels = 0;
  std::vector<int64_t> wave_level_ptr; // histogram offsets of the levels (n_wave_levels + 1)
  std::vector<int64_t> node_level;     // node id - 1 -> level
  std::vector<int64_t> perm;           // device slot -> node id (1-based); chunk-major, then
                                       // level, then node id
  std::vector<int64_t> slot_of;        // node id - 1 -> device slot (0-based)
  int64_t n_chunks = 0;
  std::vector<int64_t> chunk_of_node;  // node id - 1 -> chunk (in execution order)
  std::vector<int64_t> chunk_ptr;      // n_chunks + 1 slot offsets
  std::vector<int64_t> chunk_l0, chunk_l1;  // first / last (= outlet) level of a chunk
  std::vector<int64_t> chunk_outlet;   // root node ids (1-based) of all pieces, in chunk order
  std::vector<int64_t> out_of_node;    // node id - 1 -> outlet number, -1 if it does not publish
  int64_t n_outlets = 0;
  std::vector<int64_t> chunk_clp_off;  // n_chunks + 1 offsets into clp
  std::vector<int64_t> clp;            // per chunk: (l1 - l0 + 2) absolute slot offsets of its levels
  // EdgeConnectivity (network.jl:27-33,136-153; built on request): 1-based index of the active
  // neighbour in the four grid directions, n + 1 where there is none
  std::vector<int64_t> edge_x_up, edge_x_down, edge_y_up, edge_y_down;
};

// Build `down` from a gridded LDD (flowgraph). `indices` holds 2n CartesianIndex pairs.
// Returns false (and sets err) on a cycle.
bool build_graph(Network& nw, int64_t d1, int64_t d2, const int64_t* indices, const uint8_t* ldd,
                 int64_t n, std::string& err);
// EdgeConnectivity(network::NetworkLand) (network.jl:136-153) of the cells in `indices`.
bool build_edge_connectivity(Network& nw, int64_t d1, int64_t d2, const int64_t* indices, int64_t n,
                             std::string& err);
// order, stream order (unless `streamorder_override` given), upstream CSR, partition, wavefront.
bool build_artifacts(Network& nw, int nthreads, int min_streamorder,
                     const int64_t* streamorder_override, std::string& err);
// Cut the forest into pieces of at most `cap` nodes (bottom-up: when the not-yet-cut upstream
// tree of a node would exceed the cap its largest children are cut off; every pit closes a
// piece) and, with piece_depth > 0, of at most piece_depth levels (every node whose distance to
// its outlet is a multiple of piece_depth closes a piece); pack pieces with the same root level
// into chunks of at most `cap` nodes (piece_depth == 0: one piece per chunk); derive the device
// slot order (chunk, level, node id). Shallow chunks keep the lanes of a warp busy: a chunk of
// L levels walks L - 1 + S stages for S sub-steps.
void build_chunks(Network& nw, int64_t cap, int64_t piece_depth);

}  // namespace wfb
