/*
 * wflow_b200.h -- C ABI of libwflow_b200.so: the B200 (sm_100a) implementation of Wflow.jl's
 * per-timestep hot path for the `sbm` model type (SBM vertical land update + kinematic-wave
 * routing of subsurface, overland and river flow).
 *
 * The reference has NO plugin/FFI API for this path (pure Julia, multiple dispatch). The entry
 * points below are what Julia methods of the reference's `update!`-style functions would
 * `ccall` (INTEGRATION.md shows the shim); each one cites the Julia function it replaces.
 * All paths are relative to /root/reference/Wflow/src.
 *
 * Conventions: `extern "C"`, plain pointers and sizes, every call returns int32 status
 * (0 = ok; message via wflowb200_last_error). Never throws across the ABI. Host arrays are
 * owned by the caller and never retained after the call returns. Indices crossing the ABI are
 * Julia `Int` (int64, 1-based). One handle per GPU; calls on a handle are not re-entrant.
 * There is NO CPU fallback: every entry point fails with WFLOWB200_ERR_CUDA if no device.
 */
#ifndef WFLOW_B200_H
#define WFLOW_B200_H
#include <stdint.h>
#include "wflow_b200_fields.h"

#ifdef __cplusplus
extern "C" {
#endif

#define WFLOWB200_OK 0
#define WFLOWB200_ERR_ARG 1
#define WFLOWB200_ERR_CUDA 2
#define WFLOWB200_ERR_STATE 3
#define WFLOWB200_ERR_GRAPH 4 /* cycle in the drainage graph (routing/utils.jl:26-31) */

typedef struct WflowB200 WflowB200;

/* Field ids: enum in WFLOWB200_FIELDS order. */
enum {
#define X(name, kind) WFLOWB200_F_##name,
  WFLOWB200_FIELDS(X)
#undef X
  WFLOWB200_NUM_FIELDS
};

/* int64 per-cell arrays */
#define WFLOWB200_I_number_of_layers 0   /* SbmSoilParameters.number_of_layers  soil/soil.jl:91 */
#define WFLOWB200_I_n_unsatlayers 1      /* SbmSoilVariables.n_unsatlayers      soil/soil.jl:21 */
#define WFLOWB200_I_nlayers_kv 2         /* KvLayeredExponential.nlayers_kv    soil/soil.jl:236 */

/* The flags of config_structure.jl:62-115 (ModelSection) that change the hot path. */
typedef struct {
  int64_t n;                   /* active land cells  length(domain.land.network.indices)       */
  int64_t nriv;                /* river cells        length(domain.river.network.indices)      */
  int32_t n_layers;            /* maximum_number_of_layers = len(soil_layer__thickness)+1      */
  int32_t device;              /* CUDA device ordinal                                          */
  int32_t gash;                /* 1: Gash (dt >= 23 h), 0: modified Rutter         sbm.jl:26-33 */
  int32_t has_lai;             /* !isnothing(leaf_area_index)                  canopy.jl:65,128 */
  int32_t snow;                /* snow__flag                                                   */
  int32_t glacier;             /* glacier__flag (only with snow, sbm.jl:41-54)                 */
  int32_t soil_infiltration_reduction; /* soil_infiltration_reduction__flag                    */
  int32_t kv_profile;          /* 0 exponential, 1 exponential_constant, 2 layered,
                                  3 layered_exponential        soil.jl:213-244, utils.jl:727-789 */
  int32_t adaptive;            /* kinematic_wave__adaptive_time_step_flag                      */
  int32_t nthreads;            /* Threads.nthreads() the artefacts should be built for
                                  (subdomains.jl:176: 1 -> single sub-domain)                  */
  int32_t land_streamorder_min;  /* land_streamorder__min_count  (5)                           */
  int32_t river_streamorder_min; /* river_streamorder__min_count (6)                           */
  double dt_land;              /* land_kinematic_wave__time_step       (3600 s)                */
  double dt_river;             /* river_kinematic_wave__time_step      (900 s)                 */
  double dt_ssf;               /* subsurface_kinematic_wave__time_step (86400 s)               */
  double ssf_alpha_coefficient;/* subsurface_kinematic_wave__alpha_coefficient                 */
  double kin_wave_min_flow_qroot; /* KIN_WAVE_MIN_FLOW^0.2 as the host evaluates it
                                  (routing/utils.jl:2); 0 -> library uses pow(1e-30, 0.2)      */
  /* B200 tuning, fixed at create; 0 = automatic (chosen from the shape of the domain) */
  int32_t wave_piece_depth_land; /* levels per piece of the land chunks; -1: one connected piece */
  int32_t vertical_slices;       /* unused (kept for ABI stability)                             */
  int32_t unsat_inline_iters;    /* Brooks-Corey loops up to this many trips run in line (4)    */
  int32_t snow_gravitational_transport; /* snow_gravitational_transport__flag: lateral snow
                                  transport between snow and glacier model     sbm.jl:98-100 */
  /* river_routing = "local_inertial" (config_structure.jl; surface_staggered_scheme.jl): the
   * river flow of update_river_flow_model! on the staggered grid instead of the kinematic wave */
  int32_t river_routing;       /* 0 kinematic_wave, 1 local_inertial                           */
  int32_t li_froude_limit;     /* river_water_flow__froude_limit_flag                          */
  int32_t li_ghost_nodes;      /* 1: every pit drains to a ghost node (the model), 0: it has no
                                  leaving edge (the reference's unit tests)                    */
  int32_t reserved_;
  double li_alpha;             /* river_local_inertial_flow__alpha_coefficient (0.7)           */
  double li_h_thresh;          /* river_water_flow_threshold__depth (1e-3 m)                   */
  /* floodplain_1d__flag: with the local-inertial river (floodplain.jl, surface_staggered_scheme.jl:
   * 440-533,674-712) or with the kinematic-wave river (surface_kinwave.jl:387-432,567-601;
   * with or without reservoirs): the flood depths of the FloodPlainProfile; fp_levels = 0: none */
  int32_t fp_levels;           /* length(profile.depth), <= 16                                 */
  int32_t reserved2_;
  double fp_depth[16];         /* profile.depth [m], ascending from 0                          */
  /* land_routing = "local_inertial" (needs river_routing = 1, no 1-D floodplain): 2-D overland
   * flow on the staggered grid coupled to the local-inertial river with its subgrid channel,
   * update_overland_flow_model!(overland, river, domain, clock, dt)
   * (surface_staggered_scheme.jl:1153-1194); the fields li_land_* exist only then */
  int32_t land_routing;        /* 0 kinematic_wave, 1 local_inertial                           */
  int32_t li_land_froude_limit;/* land_surface_water_flow__froude_limit_flag                   */
  double li_land_alpha;        /* land_local_inertial_flow__alpha_coefficient (0.7)            */
  double li_land_theta;        /* land_local_inertial_flow__theta_coefficient (1.0)            */
  double li_land_h_thresh;     /* land_surface_water_flow_threshold__depth (1e-3 m)            */
} WflowB200Config;

/* The drainage network as the Julia model holds it (network.jl:48-81,175-208). */
typedef struct {
  int64_t d1, d2;                    /* size(subcatch_2d): Julia dimension order               */
  const int64_t* indices;            /* 2n: CartesianIndex (i, j) pairs, 1-based, column-major
                                        ascending (utils.jl:85-99)                             */
  const uint8_t* ldd;                /* n: land local_drain_direction (PCRaster 1..9)          */
  const int64_t* river_land_indices; /* nriv: NetworkRiver.land_indices, 1-based, ascending    */
  /* reservoir__flag (domain.jl:96-109): the river node of every reservoir outlet, i.e. the
   * positions of the non-zero entries of NetworkRiver.reservoir_indices, reservoir 1 first.
   * Reservoir outlets are dropped from the upstream lists of the land and river kinematic
   * waves; their outflow becomes the inflow of the downstream river node
   * (surface_kinwave.jl:441-489). nres = 0 / NULL: no reservoirs. */
  int64_t nres;
  const int64_t* reservoir_river_indices; /* nres, 1-based river node ids                      */
  /* A shard that is PART of a drainage basin, cut at confluences (the reference cuts its basins
   * the same way into sub-domains for its threads, subdomains.jl:169-255): discharge crosses
   * the shard's border on CUT EDGES. An IMPORT is a cut edge whose upstream node lives on another
   * shard; an EXPORT is one whose downstream node does (here the node is a pit of `ldd`).
   * import_dst: the local node (1-based) the edge ends in; import_pos: the position of the edge
   * among ALL upstream sources of that node in ascending GLOBAL node id (0-based) -- the
   * reference sums upstream values in that order (utils.jl:472-477) and the shards must too.
   * export_src: the local node (1-based) the edge leaves. All zero / NULL: no cut edges.
   * Values travel through wflowb200_exchange_* below. */
  int64_t n_land_imports;
  const int64_t* land_import_dst;
  const int64_t* land_import_pos;
  int64_t n_land_exports;
  const int64_t* land_export_src;
  int64_t n_river_imports;
  const int64_t* river_import_dst;
  const int64_t* river_import_pos;
  int64_t n_river_exports;
  const int64_t* river_export_src;
} WflowB200Domain;

/* artefact ids for wflowb200_get_artifact (all returned as 1-based int64, reference layout) */
#define WFLOWB200_A_ORDER 0            /* network.order (topological_sort_by_dfs)              */
#define WFLOWB200_A_STREAMORDER 1      /* network.streamorder                                  */
#define WFLOWB200_A_UPSTREAM_PTR 2     /* upstream_nodes as CSR by TOPOSORT POSITION: ptr n+1
                                          (0-based offsets)                                    */
#define WFLOWB200_A_UPSTREAM_IDX 3     /* ... node ids, ascending within a list                */
#define WFLOWB200_A_SUBDOMAIN_LEVEL_PTR 4 /* order_of_subdomains CSR offsets                   */
#define WFLOWB200_A_SUBDOMAIN_LEVEL_IDX 5 /* order_of_subdomains sub-domain ids                */
#define WFLOWB200_A_SUBDOMAIN_PTR 6    /* CSR offsets of order_subdomain / subdomain_indices   */
#define WFLOWB200_A_SUBDOMAIN_ORDER 7  /* order_subdomain (node ids)                           */
#define WFLOWB200_A_SUBDOMAIN_INDICES 8 /* subdomain_indices (toposort positions)              */
#define WFLOWB200_A_LDD 9              /* ldd after flowgraph's pit fix-up                     */
#define WFLOWB200_A_WAVE_LEVEL_PTR 10  /* B200 wavefront: histogram offsets of the
                                          topological-depth levels                             */
#define WFLOWB200_A_WAVE_PERM 11       /* node id held by each device slot (chunk, level, id)  */
#define WFLOWB200_A_WAVE_NODE_LEVEL 12 /* level of every node (0-based), by node id            */
#define WFLOWB200_A_WAVE_CHUNK_PTR 13  /* slot offsets of the chunks (0-based, n_chunks + 1)   */
#define WFLOWB200_A_WAVE_CHUNK_OUTLET 14 /* outlet node id of every chunk                      */
/* EdgeConnectivity of the land network (network.jl:27-33,136-153; land_routing = 1): index of the
 * active neighbour of every cell, n + 1 where there is none */
#define WFLOWB200_A_EDGE_X_UP 15       /* CartesianIndex(1, 0)                                 */
#define WFLOWB200_A_EDGE_X_DOWN 16     /* CartesianIndex(-1, 0)                                */
#define WFLOWB200_A_EDGE_Y_UP 17       /* CartesianIndex(0, 1)                                 */
#define WFLOWB200_A_EDGE_Y_DOWN 18     /* CartesianIndex(0, -1)                                */
#define WFLOWB200_DOMAIN_LAND 0
#define WFLOWB200_DOMAIN_RIVER 1

/* ---- lifetime ------------------------------------------------------------------------- */

/* Replaces the construction of NetworkLand/NetworkRiver artefacts (network.jl:87-133,214-278,
 * subdomains.jl:169-255, utils.jl:61-71) and allocates every model array in HBM (NaN-filled,
 * like `fill(MISSING_VALUE, n)`). */
int32_t wflowb200_create(const WflowB200Config* cfg, const WflowB200Domain* dom, WflowB200** out);
void wflowb200_destroy(WflowB200* h);
const char* wflowb200_last_error(const WflowB200* h); /* h may be NULL: error of a failed create */

/* ---- field table ---------------------------------------------------------------------- */
int32_t wflowb200_num_fields(void);
const char* wflowb200_field_name(int32_t field_id);
int32_t wflowb200_field_kind(int32_t field_id);
int32_t wflowb200_field_id(const char* name); /* -1 if unknown */

/* ---- host <-> device state (BMI get_value_ptr/set_value bmi.jl:208-248; set_states!
 *      sbm_model.jl:104-193; output writers io.jl:815-885) ----------------------------- */

/* Copy a host array into the device field. Element (cell c, layer k) is read from
 * src[c*stride_cell + k*stride_layer]; Julia's Vector{SVector{N,Float64}} is
 * (stride_cell = N, stride_layer = 1). Scalars: stride_cell = 1, stride_layer ignored. */
int32_t wflowb200_set_field(WflowB200* h, int32_t field_id, const double* src,
                            int64_t stride_cell, int64_t stride_layer);
int32_t wflowb200_get_field(WflowB200* h, int32_t field_id, double* dst, int64_t stride_cell,
                            int64_t stride_layer);
int32_t wflowb200_set_field_i64(WflowB200* h, int32_t which, const int64_t* src);
int32_t wflowb200_get_field_i64(WflowB200* h, int32_t which, int64_t* dst);

/* update_forcing! hand-off (io.jl:108-160 -> AtmosphericForcing forcing.jl:2-10): three host
 * vectors [m s-1, m s-1, K]; copied H2D on the library's copy stream -- straight from the
 * caller's arrays when they are page-locked (cudaHostRegister / cudaMallocHost), else through
 * the library's pinned staging buffer. The arrays may be reused as soon as the call returns;
 * the next update_* call waits for the copy on the device. */
int32_t wflowb200_set_forcing(WflowB200* h, const double* precipitation,
                              const double* potential_evaporation, const double* temperature);

/* Staging ring for the forcing of the coming steps (the reader thread of io.jl:108-160 runs
 * ahead of the model): `depth` slabs of (P, PET, T) in HBM. put() starts the H2D copy of one slab
 * on the copy stream and returns at once when the arrays are page-locked (it waits, on the
 * device, until the step that last used the slot has consumed it); use() makes the next
 * update_* call take its forcing from that slab instead of wflowb200_set_forcing's. */
int32_t wflowb200_forcing_ring_create(WflowB200* h, int32_t depth);
int32_t wflowb200_forcing_ring_put(WflowB200* h, int32_t slot, const double* precipitation,
                                   const double* potential_evaporation, const double* temperature);
int32_t wflowb200_forcing_ring_use(WflowB200* h, int32_t slot);

/* Cyclic leaf_area_index (update_cyclic!, io.jl:187-227): all n_slabs slabs (12 monthly or 366
 * daily, [slab][cell] in node order) are staged in HBM once; use() copies one slab into the
 * model's leaf_area_index on the device when the month / day changes -- no host traffic. */
int32_t wflowb200_set_cyclic_lai(WflowB200* h, const double* table, int32_t n_slabs);
int32_t wflowb200_use_cyclic_lai(WflowB200* h, int32_t slab);

/* Output gather (write_output io.jl:815-899 reads a handful of model vectors per step): the
 * listed fields, concatenated in this order (node order, layered fields cell-major), with ONE
 * device-to-host copy through page-locked memory. dst holds the sum of the fields' sizes. */
int32_t wflowb200_get_fields(WflowB200* h, const int32_t* field_ids, int32_t n_ids, double* dst);
/* The same without blocking: dst must be page-locked (cudaHostRegister / cudaMallocHost); the
 * fields are packed on the compute stream, copied on the copy stream, and the next update_* call
 * may be issued at once (the output writer of step s runs while step s + 1 computes).
 * wait_outputs blocks until every pending copy has landed. */
int32_t wflowb200_get_fields_async(WflowB200* h, const int32_t* field_ids, int32_t n_ids,
                                   double* dst_pinned);
int32_t wflowb200_wait_outputs(WflowB200* h);

/* ---- the hot path --------------------------------------------------------------------- */

/* update_land_hydrology_model!(land, routing, domain, config, dt)            sbm.jl:82-132 */
int32_t wflowb200_update_land_hydrology_model(WflowB200* h, double dt);
/* recharge / water-table hand-off between soil and subsurface flow     sbm_model.jl:74-81 */
int32_t wflowb200_exchange_recharge(WflowB200* h);
/* update_subsurface_flow_model!(ssf, soil, domain, dt)  lateral_subsurface_flow.jl:279-304 */
int32_t wflowb200_update_subsurface_flow_model(WflowB200* h, double dt);
/* update_soil_water_storage!(soil, external_models, dt)               soil/soil.jl:1294-1392 */
int32_t wflowb200_update_soil_water_storage(WflowB200* h, double dt);
/* update_lateral_inflow!(overland, ...)               routing/surface/surface_kinwave.jl:740-766 */
int32_t wflowb200_update_lateral_inflow_overland(WflowB200* h);
/* update_overland_flow_model!(overland, domain.land, dt)              surface_kinwave.jl:347-385
 * land_routing = 1: update_overland_flow_model!(overland, river, domain, clock, dt) -- the 2-D
 * local-inertial overland flow AND the local-inertial river flow of the model step: per sub-step
 * dt_s = min(stable_timestep(river), stable_timestep(land)), the x / y edge flows of every cell
 * (update_directional_flow!, local_inertial_flow of de Almeida et al. 2012), the reservoirs'
 * overland inflow, river edge flow, reservoirs, and the water depth and storage of land and
 * river cells (subgrid channel, bankfull spill) -- all sub-steps inside ONE persistent kernel
 * surface_staggered_scheme.jl:1022-1043,1153-1546; surface_process.jl:123-159 */
int32_t wflowb200_update_overland_flow_model(WflowB200* h, double dt);
/* update_bc_overland_flow_model!(overland, (; soil, runoff, subsurface_flow), domain, dt): runoff
 * = (net_runoff + net_runoff_river) * area, plus the subsurface flow to the river at river cells
 * (land_routing = 1)                                       surface_staggered_scheme.jl:1080-1097 */
int32_t wflowb200_update_bc_overland_flow_model(WflowB200* h);
/* update_lateral_inflow!(river, ...)                                  surface_kinwave.jl:710-734 */
int32_t wflowb200_update_lateral_inflow_river(WflowB200* h);
/* update_inflow!(reservoir, river_flow, (; overland_flow, subsurface_flow), network): overland
 * and subsurface flow into the reservoirs                             surface_kinwave.jl:772-805
 * land_routing = 1: update_inflow!(reservoir, river_flow, subsurface_flow, network) -- the
 * subsurface flow only, the overland inflow is formed inside the overland routing
 * (update_inflow_reservoir!)                        surface_staggered_scheme.jl:1103-1114,1301-1319 */
int32_t wflowb200_update_inflow_reservoir(WflowB200* h);
/* update_river_flow_model!(river, domain, clock, dt), incl. the reservoirs on the river
 * (update_reservoir_model! reservoir.jl:585-634: simple, modified_puls, free_weir without a
 * linked lower reservoir, observed outflow; linear storage curve)    surface_kinwave.jl:613-662
 * river_routing = 1: update_river_flow_model!(::RiverFlowModel{<:LocalInertial}) -- adaptive
 * sub-steps dt_s = alpha min(L / sqrt(g h)), edge flow, reservoirs, node depth and storage, all
 * sub-steps of the model step inside ONE persistent kernel   surface_staggered_scheme.jl:326-383,
 * 627-661, 723-759, 800-838, 1004-1020; surface_process.jl:88-115 */
/* land_routing = 1: WFLOWB200_ERR_STATE -- the river is routed together with the land by
 * wflowb200_update_overland_flow_model (one scheme, one sub-step length for both). */
int32_t wflowb200_update_river_flow_model(WflowB200* h, double dt);
/* update_total_water_storage!(land, domain, routing)                          sbm.jl:143-182 */
int32_t wflowb200_update_total_water_storage(WflowB200* h);
/* update_model!(model::AbstractModel{<:SbmModel}): all of the above, in order, state resident
 * on the device                                                          sbm_model.jl:60-92 */
int32_t wflowb200_update_model(WflowB200* h, double dt);
/* Self-test of the device arithmetic the kernels are built on (csrc/device_math.cuh), over n
 * pseudo-random arguments: out6 = { max relative difference of the remaining store and of the
 * summed flux between the loop engine's tracked-power trips and the reference loop
 * (soil_process.jl:74-90), max relative difference of pow(x, c) = exp(c log x)
 * against libdevice's pow for x in (0, 1], c in [1, 40], max ulp distance of the guard-free
 * division vs IEEE `/`, of the branch-free Julia min/max vs their definition, max |difference|
 * of cld(x, 2e-4) vs Julia's formula }. No handle needed. */
int32_t wflowb200_selftest_math(int32_t device, int64_t n, double* out6);

/* block until all device work of this handle is done; WFLOWB200_ERR_STATE if a wavefront kernel
 * gave up a bounded wait (its grid was not co-resident: MPS share, second context, debugger) */
int32_t wflowb200_synchronize(WflowB200* h);

/* ---- sharded domains (one handle per GPU / per shard) ---------------------------------------
 * Basin-aligned shards need no exchange of fluxes (subdomains.jl:177-199: the reference threads
 * over the same units). With kinematic_wave__adaptive_time_step_flag every routing sub-step
 * additionally needs a statistic of the WHOLE domain -- the type-7 quantile of the Courant
 * steps (surface_kinwave.jl:674-704) and their minimum (lateral_subsurface_flow.jl:314-344) --
 * which the shards reduce: histograms of the radix select, counts and minima are all-reduced,
 * so every shard takes exactly the sub-steps the unsharded domain takes.
 *   NCCL, one process per GPU: rank 0 calls comm_unique_id, the host distributes the 128 bytes
 *   (MPI / torch.distributed), every rank calls comm_init_nccl(handle, rank, world, id).
 *   One process, several handles (several shards on one GPU, tests): group_create(n), every
 *   handle joins; the handles are then stepped from n host threads at the same time. */
typedef struct WflowB200Group WflowB200Group;
int32_t wflowb200_comm_unique_id(char* out128);
int32_t wflowb200_comm_init_nccl(WflowB200* h, int32_t rank, int32_t world, const char* id128);
int32_t wflowb200_group_create(int32_t n_handles, WflowB200Group** out);
int32_t wflowb200_group_join(WflowB200Group* g, WflowB200* h);
void wflowb200_group_destroy(WflowB200Group* g);

/* ---- cut edges: discharge across shards of ONE basin over NVLink peer memory ---------------
 * (kinematic-wave routing with fixed internal time steps; update_model only.)
 * Every sub-step value that crosses a cut edge -- subsurface flow (and its share for the river),
 * overland flow of each of the S_land sub-steps, river discharge of each of the S_river sub-steps --
 * is ONE 8-byte store of the producing lane straight into the consumer GPU's memory
 * (st.relaxed.sys through the NVLink mapping), into the slot the consuming lane polls: the same
 * data-is-flag slots the chunks of one GPU use among themselves, so the wavefronts of the shards
 * run as ONE skewed wavefront without any collective on the data path. The slots are double
 * buffered by the parity of the model step; a shard resets the other buffer at the start of a
 * step, behind a barrier of the communicator (wflowb200_comm_init_nccl / group_join: required).
 *   1. every shard: exchange_prepare(h, dt) -> its import buffer (device pointer for handles of
 *      the same process, CUDA IPC handle for other processes) and size;
 *   2. every producer: exchange_open_peer(h, peer, ...) for each shard it exports to, then
 *      exchange_bind(h, domain, export, peer, import of that peer) for each export;
 *   3. update_model as usual, all shards in the same step. */
int32_t wflowb200_exchange_prepare(WflowB200* h, double dt, uint64_t* device_ptr,
                                   void* ipc_handle64, int64_t* bytes);
/* peer: any id >= 0 chosen by the caller (< 64). device_ptr != 0: the peer's buffer is
 * addressable as is (same process; peer_device = its CUDA device, peer access is enabled);
 * otherwise ipc_handle64 is opened. n_*_imports: the PEER's import counts (layout of its buffer). */
int32_t wflowb200_exchange_open_peer(WflowB200* h, int32_t peer, uint64_t device_ptr,
                                     int32_t peer_device, const void* ipc_handle64,
                                     int64_t peer_n_land_imports, int64_t peer_n_river_imports);
/* export `export_index` (0-based, order of *_export_src) of `domain` feeds import
 * `peer_import_index` (0-based, order of the peer's *_import_dst) of the same domain on `peer` */
int32_t wflowb200_exchange_bind(WflowB200* h, int32_t domain, int64_t export_index, int32_t peer,
                                int64_t peer_import_index);

/* Select between kernel organisations that give identical results (the parity tests run every
 * one): "fuse_soil_storage", "overlap_subsurface" (-1 automatic, 0, 1), "overlap_subsurface_sms",
 * "fuse_surface" (0, 1), "surface_river_share", "surface_river_period", "vertical_graph" (0, 1),
 * "kinwave_root_each_substep" (0, 1: see below). */
int32_t wflowb200_set_option(WflowB200* h, const char* name, int32_t value);
/* Developer aid: with the options "vertical_timeline" = 1 and "vertical_graph" = 0, the completion
 * times [ms since the start of the last vertical update] of its kernels: [0] = 0,
 * [1] land_hydrology_kernel, [2] unsat_engine_kernel, [3] soil_column_kernel. capacity >= 4. */
int32_t wflowb200_get_vertical_timeline(WflowB200* h, double* out_ms, int32_t capacity);
/* Diagnostic: cells the last vertical update handed to the loop engine, by the trip count of the
 * loop they were suspended at: out[b], b < 8, for (2 * 2^b, 4 * 2^b] trips (the first bucket
 * from 2, the last open-ended); out[8 + b]: the LATER loops of more than 4 trips the engine ran
 * for those cells, same classes. capacity >= 16. */
int32_t wflowb200_get_unsat_buckets(WflowB200* h, int64_t* out, int32_t capacity);
/* "kinwave_root_each_substep" = 1 additionally evaluates u_prev = pow(q_prev, 0.2) before EVERY
 * kinematic-wave solve like surface_process.jl:33 (default: the fifth root is carried from the
 * previous solve of the node, within half an ulp of it; one pow per node and model step). */

/* Newton iteration counts of `kinematic_wave` (surface_process.jl:24-70) per node, summed over
 * the solves since enabling (iteration-count parity against the reference). enable = 1
 * allocates and zeroes the counters, 0 frees them. dst: n (land) / nriv (river) values in node
 * order. */
int32_t wflowb200_newton_trace(WflowB200* h, int32_t enable);
int32_t wflowb200_get_newton_trace(WflowB200* h, int32_t domain, int64_t* dst);

/* ---- artefacts and statistics --------------------------------------------------------- */

/* Copy an indexing artefact (1-based int64, reference layout). If dst is NULL only *len_out is
 * set. */
int32_t wflowb200_get_artifact(WflowB200* h, int32_t domain, int32_t artifact_id, int64_t* dst,
                               int64_t capacity, int64_t* len_out);

/* Host-only variant (no device needed): build just the indexing artefacts of a domain, e.g. to
 * check them against the Julia model's `network` fields before moving a model to the GPU. */
typedef struct WflowB200Network WflowB200Network;
int32_t wflowb200_network_build(const WflowB200Config* cfg, const WflowB200Domain* dom,
                                WflowB200Network** out);
int32_t wflowb200_network_get(const WflowB200Network* net, int32_t domain, int32_t artifact_id,
                              int64_t* dst, int64_t capacity, int64_t* len_out);
void wflowb200_network_destroy(WflowB200Network* net);

typedef struct {
  int64_t newton_calls_land, newton_iters_land, newton_maxit_land;
  int64_t newton_calls_river, newton_iters_river, newton_maxit_river;
  int64_t substeps_land, substeps_river, substeps_ssf;
  int64_t wave_levels_land, wave_levels_river;
  int64_t kernel_launches;      /* kernels of this library launched since create            */
  /* CUDA-event time [ms] per stage, accumulated over the update_model calls made while timing
   * was enabled (wflowb200_set_timing resets them): */
  double ms_land_hydrology;     /* land_hydrology_kernel alone (V1)                          */
  double ms_subsurface;         /* subsurface_wave_kernel                                    */
  double ms_soil_storage;       /* soil_water_storage_kernel (V2)                            */
  double ms_overland;           /* overland_wave_kernel                                      */
  double ms_river;              /* river_wave_kernel                                         */
  double ms_total_storage;      /* total_water_storage_kernel (V3)                           */
  double ms_glue;               /* scatter / exchange / lateral-inflow / forcing-gather      */
  int64_t timed_steps;
} WflowB200Stats;
int32_t wflowb200_get_stats(WflowB200* h, WflowB200Stats* out);
/* enable per-stage CUDA-event timing inside update_model (off by default); resets the sums */
int32_t wflowb200_set_timing(WflowB200* h, int32_t enabled);
/* Device-side stopwatch on the handle's compute stream (CUDA events): start records an event,
 * stop records another, waits for it and returns the elapsed milliseconds between the two. */
int32_t wflowb200_timer_start(WflowB200* h);
int32_t wflowb200_timer_stop(WflowB200* h, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif
