// soil_storage.cuh -- update_soil_water_storage! for one cell (soil/soil.jl:1294-1392), shared by
// soil_water_storage_kernel (vertical.cu) and the fused land-routing kernel (routing.cu), which
// runs it for the cells of a chunk right after their subsurface flow is final.
#pragma once
#include "device_math.cuh"
#include "model.cuh"

namespace wfb {

template <int N>
__device__ __forceinline__ void soil_water_storage_cell(const DevFields& f, const int ns, const int i) {
  const int nu = f.n_unsatlayers[i];
  const int nl = f.number_of_layers[i];
  const double theta_s = __ldg(f.theta_s + i), theta_r = __ldg(f.theta_r + i);
  const double te = theta_s - theta_r;
  const double rd = __ldg(f.rooting_depth + i);
  const double zi = f.water_table_depth[i];
  double usd = 0.0, rootstore_unsat = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double uld = f.unsaturated_layer_depth[k * ns + i];
    const double ult = f.unsaturated_layer_thickness[k * ns + i];
    const double alt = __ldg(f.actual_layer_thickness + k * ns + i);
    const double cld = __ldg(f.cumulative_layer_depth + k * ns + i);
    if (k < nu) {
      usd += uld;
      rootstore_unsat += jmin(1.0, (jmax(0.0, rd - cld) / ult)) * uld;
    }
    if (k < nl) {
      const double vwc = k < nu ? (uld + (alt - ult) * te) / alt + theta_r : theta_s;
      f.volumetric_water_content[k * ns + i] = vwc;
      f.relative_volumetric_water_content[k * ns + i] = (vwc / theta_s) / 1e-2;
    }
  }
  const double exf = f.ssf_exfiltwater_average[i];
  const double sbm_runoff = jmax(0.0, exf + f.saturation_excess_water[i] + f.runoff_land[i] +
                                          f.infiltration_excess[i]);
  const double rootstore_sat = jmax(0.0, rd - zi) * te;
  const double rzs = rootstore_sat + rootstore_unsat;
  const double vwc_rz = rzs / rd + theta_r;
  const double d_soil = __ldg(f.soil_thickness + i);
  const double satwd = (d_soil - zi) * te;
  const double drainable = (d_soil - zi) * jmax(theta_s - __ldg(f.theta_fc + i), 0.02);
  f.unsaturated_store_capacity[i] = __ldg(f.soil_water_capacity + i) - satwd - usd;
  f.unsaturated_store_depth[i] = usd;
  f.saturated_water_depth[i] = satwd;
  f.drainable_water_depth[i] = drainable;
  f.exfiltration_saturated_water[i] = exf;
  f.runoff[i] = sbm_runoff;
  f.root_zone_storage[i] = rzs;
  f.volumetric_water_content_root_zone[i] = vwc_rz;
  f.relative_volumetric_water_content_root_zone[i] = (vwc_rz / theta_s) / 1e-2;
  f.total_soil_water_storage[i] = satwd + usd;
  const double net_runoff = sbm_runoff - f.actual_open_water_evaporation_land[i];
  f.net_runoff[i] = net_runoff;
  // update_lateral_inflow!(overland)  surface_kinwave.jl:740-766 (no drains / demand), fused:
  f.olf_inwater[i] = (net_runoff + 0.0) * __ldg(f.area + i) + 0.0;
}

}  // namespace wfb
