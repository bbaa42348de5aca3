// floodplain.cuh -- FloodPlainProfile lookups (routing/surface/floodplain.jl:287-354) shared by the
// local-inertial river (local_inertial.cu) and the kinematic-wave river (routing.cu).
#pragma once
#include "device_math.cuh"

namespace wfb {
namespace {

// tables are [level][river slot]
struct FpTables {
  const double *storage, *width, *flow_area, *perimeter;
  int nrs, levels;
  const double* depth;  // device
};
// interpolation_indices: the last level with v[l] <= x (and the next one)
__device__ __forceinline__ void fp_indices_depth(const FpTables& t, double x, int& i1, int& i2) {
  int a = 0;
  for (int l = 0; l < t.levels; ++l)
    if (__ldg(t.depth + l) <= x) a = l;
  i1 = a;
  i2 = a == t.levels - 1 ? a : a + 1;
}
// compute_floodplain_flow_area (flood flow area minus the channel's share)
__device__ __forceinline__ double fp_flow_area(const FpTables& t, double h, int p, int i1, int i2) {
  const double channel_area = __ldg(t.width + p) * h;
  const double delta_h = h - __ldg(t.depth + i1);
  const double flow_area = __ldg(t.flow_area + i1 * t.nrs + p) + (__ldg(t.width + i2 * t.nrs + p) * delta_h);
  return jmax(flow_area - channel_area, 0.0);
}
__device__ __forceinline__ double fp_wetted_perimeter(const FpTables& t, double h, int p, int i1) {
  const double delta_h = h - __ldg(t.depth + i1);
  return __ldg(t.perimeter + i1 * t.nrs + p) + 2.0 * delta_h;
}
// compute_flood_depth
__device__ __forceinline__ double fp_flood_depth(const FpTables& t, double flood_storage,
                                                 double flow_length, int p) {
  int a = 0;
  for (int l = 0; l < t.levels; ++l)
    if (__ldg(t.storage + l * t.nrs + p) <= flood_storage) a = l;
  const int i2 = a == t.levels - 1 ? a : a + 1;
  const double delta_A = (flood_storage - __ldg(t.storage + a * t.nrs + p)) / flow_length;
  const double delta_h = delta_A / __ldg(t.width + i2 * t.nrs + p);
  return __ldg(t.depth + a) + delta_h;
}


// manning_flow                                                    surface_process.jl:166-169
__device__ __forceinline__ double manning_flow(double mannings_n, double hydraulic_radius, double slope,
                                               double area) {
  return cbrt(hydraulic_radius * hydraulic_radius) * sqrt(slope) * area / mannings_n;
}

}  // namespace
}  // namespace wfb
