// kernels.cuh -- host-callable launchers of the CUDA kernels (vertical.cu, routing.cu).
// Each returns the number of kernel launches it issued (negative on bad arguments).
#pragma once
#include <cuda_runtime.h>
#include <functional>
#include "model.cuh"

namespace wfb {

int launch_scatter_river_depth(const DevFields& f, const KCfg& c, cudaStream_t s);
// Device scratch of the unsaturated-zone engine (vertical.cu): the operands of the suspended
// Brooks-Corey loops (one record per cell) and the lists of suspended cells, bucketed by
// log2(trip count).
#define WFB_UNSAT_BUCKETS 8
#define WFB_V_TILE 128                        // cells per tile (= CTA of land_hydrology_kernel)
struct UnsatWork {
  double *usd, *sum_ast, *kv_it, *l_sat, *c;  // ns doubles each
  int32_t* its_layer;                         // ns: trip count | layer << 24
  int32_t* list;                              // [WFB_UNSAT_BUCKETS][cap] cell slots
  unsigned* count;                            // [WFB_UNSAT_BUCKETS]
  int32_t cap;                                // capacity of one list (cells of the slice)
  int32_t inline_iters;                       // loops up to this many trips run in line
};
// Lane-private input staging of the two dense kernels. A warp of a register-limited kernel keeps
// only a few loads in flight, so a cell's ~50 input loads become ~12 serialised round trips to
// DRAM. Instead every lane asks for ALL its inputs at kernel entry with cp.async (no register is
// tied up): row r of the CTA's shared memory receives the tile's 128 values of array p[r], lane t
// reads back only its own word (no barrier, cp.async.wait_all only).
constexpr int kMaxStage = 80;
struct StageList {
  const double* p[kMaxStage];  // array (layer slab) staged in row r, nullptr: row not in use
  int32_t rows;                // rows to stage; dynamic shared memory = rows * 128 * 8 bytes
};
struct VerticalStage {
  StageList first[3];          // land_hydrology_kernel by phase
  StageList second;            // soil_column_kernel
};
void build_vertical_stage(const DevFields& f, const KCfg& c, int n_layers, VerticalStage& vs);
// engine_grid: SMs of the device (the engine launches as many CTAs as are resident at once). phase: 0 the whole update; 1 interception
// + snow only, 2 the rest (lateral snow transport runs between the two: launch_snow_transport).
// run_engine = false leaves the suspended cells unfinished (timing experiments only). tl: optional
// timing events of the kernels' completion (wflowb200_get_vertical_timeline).
int launch_land_hydrology(const DevFields& f, const KCfg& c, int n_layers, double dt,
                          const UnsatWork& w, const VerticalStage& vs, int engine_grid, int phase,
                          bool run_engine, cudaStream_t s, cudaEvent_t const* tl = nullptr);
// self-test of device_math.cuh: out[6] (device, zeroed) receives bit patterns of the maxima
int launch_selftest_math(long long n, unsigned long long* out, cudaStream_t s);
int launch_exchange_recharge(const DevFields& f, const KCfg& c, cudaStream_t s);
int launch_soil_water_storage(const DevFields& f, const KCfg& c, int n_layers, cudaStream_t s);
int launch_total_water_storage(const DevFields& f, const KCfg& c, const int32_t* riv_of_land,
                               cudaStream_t s);

// Device-side counters of the routing kernels.
struct RoutingStats {
  unsigned long long newton_calls_land, newton_iters_land, newton_maxit_land;
  unsigned long long newton_calls_river, newton_iters_river, newton_maxit_river;
};

struct WaveLaunch {
  unsigned* queue;            // device: next chunk to hand out (zeroed by the launcher)
  unsigned long long* q_out;  // device: per chunk, the value(s) its outlet publishes in every
                              // sub-step (bit patterns, S x NV each; all-ones = not yet written)
  RoutingStats* stats;        // device
  int S;                      // number of sub-steps pipelined through the wavefront
  double dt_fixed, dt_last;   // sub-step lengths: S - 1 times dt_fixed, then dt_last
  double dt;                  // model time step (for the averages)
  unsigned* done_flags;       // subsurface flow overlapped with the surface kernel: per chunk,
  unsigned done_epoch;        // the epoch stored when the chunk's results are final; the kernel
  int trigger_dependents;     // then lets the dependent (surface) kernel launch at once
  int fuse_soil_storage;      // subsurface flow: run update_soil_water_storage! for every cell
                              // right after its subsurface flow is final (saves a pass over HBM)
  int accumulate;             // adaptive sub-stepping, one launch per sub-step: this launch
                              // continues the cumulative fluxes of the previous ones
  int grid;
  size_t smem;                // dynamic shared memory of the kernel (wave_smem)
  unsigned smem_per_warp;     // bytes of a warp's region (0: the component's own layout);
                              // set when two components share one kernel
  unsigned* err;              // device: the handle's error word (bounded waits, routing.cu)
  // cut edges (wflowb200_exchange_*): this step's import slots of the component (S x NV per
  // import, written by the producers' GPUs) and, per export, the consumer's slots in ITS memory
  const unsigned long long* imports;
  unsigned long long* const* exports;
};

// Overland and river flow in ONE kernel (launch_surface_wave): the warps of the grid are split
// between the two components; a river chunk starts when the overland flow of its cells' land
// chunks is final (per-land-chunk flags), so the river wavefront trails the overland wavefront
// by a few levels instead of starting after it.
struct SurfaceSync {
  unsigned* land_done;                 // device: per land chunk, the epoch of its last finalize
  unsigned* ssf_done;                  // device: per land chunk, subsurface flow + soil water
                                       // storage of its cells are final (nullptr: the
                                       // subsurface flow ran in a kernel of its own)
  unsigned epoch;                      // this launch
  const int32_t* land_chunk_of_slot;   // land slot -> land chunk
  int period, river_share;             // warp g serves the river if g % period < river_share
  unsigned* err;                       // device: the handle's error word (bounded waits)
};
// kind: 0 overland, 1 river, 2 subsurface
int wave_block();
size_t surface_smem(int max_inlets_land, int max_inlets_river, unsigned* per_warp);
int surface_max_grid(size_t smem, int device);
// overlap_previous: programmatic dependent launch -- the kernel may start as soon as every CTA
// of the previous kernel in the stream (the subsurface sweep) has started; it then orders itself
// after that kernel's results through sync.ssf_done only
int launch_surface_wave(const DevFields& f, const KCfg& c, const DevNet& land, const DevNet& river,
                        const WaveLaunch& wl, const WaveLaunch& wr, const SurfaceSync& sync,
                        bool overlap_previous, cudaStream_t s);
void reset_surface_wave(const DevNet& land, const DevNet& river, const WaveLaunch& wl,
                        const WaveLaunch& wr, cudaStream_t s);
size_t wave_smem(int kind, int max_inlets);
int wave_max_grid(int kind, int n_layers, size_t smem, int device);  // co-resident CTAs
int launch_overland_wave(const DevFields& f, const KCfg& c, const DevNet& net, const WaveLaunch& w,
                         cudaStream_t s);
int launch_river_floodplain_wave(const DevFields& f, const KCfg& c, const DevNet& net,
                                 const WaveLaunch& w, cudaStream_t s);
int launch_river_wave(const DevFields& f, const KCfg& c, const DevNet& net, const WaveLaunch& w,
                      cudaStream_t s);
int launch_subsurface_wave(const DevFields& f, const KCfg& c, const DevNet& net, int n_layers,
                           const WaveLaunch& w, cudaStream_t s);
// lateral_snow_transport! (surface_process.jl:9-19): accucapacityflux of snow storage and snow
// water over the land network + flux_in!; kind 3 of wave_smem / wave_max_grid
int launch_snow_transport(const DevFields& f, const KCfg& c, const DevNet& net, const WaveLaunch& w,
                          cudaStream_t s);
// Local-inertial river flow (local_inertial.cu): all sub-steps of a model step in one persistent
// kernel with grid barriers.
struct LiLaunch {
  double dt;                    // model time step
  double alpha, h_thresh;       // stability coefficient, depth threshold for flow at an edge
  int froude_limit;
  int fp_levels;                // 1-D floodplain: levels of the profile (0: none) and their depths
  double fp_depth[16];
  unsigned* barrier;            // device: {arrivals, generation}
  unsigned long long* dt_bits;  // device: 2 slots for the minimum Courant step (bit patterns)
  unsigned* err;                // device: the handle's error word (bounded barrier waits)
  int* substeps;                // device: number of sub-steps of the model step (out)
  int grid;                     // co-resident CTAs
  // 2-D local-inertial overland flow coupled to the river (land_routing = 1)
  double land_alpha, land_theta, land_h_thresh;
  int land_froude_limit;
};
int li_max_grid(int device);
// update_overland_flow_model!(overland, river, ...) (surface_staggered_scheme.jl:1153-1194): the
// 2-D local-inertial overland flow and the local-inertial river flow, all sub-steps of a model
// step in one persistent kernel
int lil_max_grid(int device);
int launch_local_inertial_land_river(const DevFields& f, const KCfg& c, const LiLaunch& w, cudaStream_t s);
// update_bc_overland_flow_model! (surface_staggered_scheme.jl:1080-1097)
int launch_bc_overland_flow(const DevFields& f, const KCfg& c, cudaStream_t s);
int launch_local_inertial_river(const DevFields& f, const KCfg& c, const LiLaunch& w, cudaStream_t s);
int launch_lateral_inflow_overland(const DevFields& f, const KCfg& c, cudaStream_t s);
int launch_lateral_inflow_river(const DevFields& f, const KCfg& c, cudaStream_t s);
int launch_inflow_reservoir(const DevFields& f, const KCfg& c, cudaStream_t s);

// adaptive time step statistics (surface_kinwave.jl:674-704, lateral_subsurface_flow.jl:314-344)
// `work` holds n doubles; results are written to out[0] (count) / keys in work.
int launch_stable_timesteps_surface(const double* q, const double* alpha, const double* len, int n,
                                    double* work, unsigned long long* count, cudaStream_t s);
int launch_stable_timestep_ssf(const DevFields& f, const KCfg& c, double* out_min,
                               unsigned long long* count, cudaStream_t s);
// Statistics.quantile! (type 7) of the `*count` positive values in `work` (Statistics 1.11.1,
// call site surface_kinwave.jl:698): radix select of the two order statistics it interpolates
// between. `state` holds QUANTILE_STATE_WORDS 64-bit words (device); after the launches
// state[0] = k, state[1] = bits of v[j], state[2] = bits of v[j+1] (sorted, 1-based j),
// state[3] = bits of gamma.
#define WFB_QUANTILE_STATE_WORDS (16 + 256)
// `reduce` (sharded domains, else nullptr): in-place all-reduce over the shards of n 64-bit words
// in device memory, ordered on the stream; op 0 = sum, 1 = minimum. After the launches
// state[10] = k of all shards.
typedef std::function<int(unsigned long long*, int, int)> ShardReduce;
int launch_quantile7(const double* work, const unsigned long long* count, int n_max, double p,
                     unsigned long long* state, cudaStream_t s, const ShardReduce* reduce);

// host<->device layout conversion through the slot permutation
int launch_gather_field(double* dst, const double* staged, const int32_t* node_of_slot, int n,
                        int ns, int layers, long long stride_cell, long long stride_layer,
                        cudaStream_t s);
int launch_scatter_field(double* staged, const double* src, const int32_t* node_of_slot, int n,
                         int ns, int layers, long long stride_cell, long long stride_layer,
                         cudaStream_t s);
int launch_gather_forcing(const DevFields& f, const double* staged, const int32_t* node_of_slot,
                          int n, cudaStream_t s);
int launch_fill(double* p, long long count, double v, cudaStream_t s);

}  // namespace wfb
