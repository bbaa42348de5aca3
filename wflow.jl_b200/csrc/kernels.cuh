// kernels.cuh -- host-callable launchers of the CUDA kernels (vertical.cu, routing.cu).
// Each returns the number of kernel launches it issued (negative on bad arguments).
#pragma once
#include <cuda_runtime.h>
#include "model.cuh"

namespace wfb {

int launch_scatter_river_depth(const DevFields& f, const KCfg& c, cudaStream_t s);
int launch_land_hydrology(const DevFields& f, const KCfg& c, int n_layers, double dt,
                          cudaStream_t s);
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
  unsigned* queue;          // device: next chunk to hand out (zeroed by the launcher)
  int* progress;            // device: per chunk, sub-steps completed at its outlet (zeroed)
  double* q_out;            // device: per chunk, the outlet discharge of every sub-step (S each)
  RoutingStats* stats;      // device
  const double* dts;        // device: sub-step lengths (S doubles)
  int S;                    // number of sub-steps pipelined through the wavefront
  double dt;                // model time step (for the averages)
  int grid;
  int block;
  long long* prof;          // optional: 6 x n_chunks int64 (start ns, end ns, wait cycles,
                            // process cycles, stages, nodes) written by thread 0 of each chunk
};

int wave_max_grid(int kind, int n_layers, int block, int device);  // co-resident blocks
int launch_overland_wave(const DevFields& f, const KCfg& c, const DevNet& net, const WaveLaunch& w,
                         cudaStream_t s);
int launch_river_wave(const DevFields& f, const KCfg& c, const DevNet& net, const WaveLaunch& w,
                      cudaStream_t s);
int launch_subsurface_wave(const DevFields& f, const KCfg& c, const DevNet& net, int n_layers,
                           const WaveLaunch& w, cudaStream_t s);
int launch_lateral_inflow_overland(const DevFields& f, const KCfg& c, cudaStream_t s);
int launch_lateral_inflow_river(const DevFields& f, const KCfg& c, cudaStream_t s);

// adaptive time step statistics (surface_kinwave.jl:674-704, lateral_subsurface_flow.jl:314-344)
// `work` holds n doubles; results are written to out[0] (count) / keys in work.
int launch_stable_timesteps_surface(const double* q, const double* alpha, const double* len, int n,
                                    double* work, unsigned long long* count, cudaStream_t s);
int launch_stable_timestep_ssf(const DevFields& f, const KCfg& c, double* out_min,
                               unsigned long long* count, cudaStream_t s);

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
