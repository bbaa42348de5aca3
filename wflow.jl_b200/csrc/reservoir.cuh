// reservoir.cuh -- one reservoir update (routing/surface/reservoir.jl), shared by the river
// kinematic wave (routing.cu) and the local-inertial river flow (local_inertial.cu).
#pragma once
#include "device_math.cuh"
#include "model.cuh"

namespace wfb {

// ---- reservoirs on the river (routing/surface/reservoir.jl) ------------------------------------
// One reservoir sits on a river node; its state lives in HBM (a handful of reservoirs per
// domain), the lane of its node updates it once per sub-step in this out-of-line function.
// update_reservoir_model!(reservoir, river variables, network, v, dt)  surface_kinwave.jl:441-489
// + update_reservoir_model!(reservoir_model, i, inflow, dt)            reservoir.jl:585-634
// Returns the outflow [m3 s-1], which becomes qin of the downstream river node.
static __device__ __noinline__ double reservoir_step(const DevFields& f, const int i, const double q_river,
                                              const double dt) {
  const double storage0 = f.res_storage[i], area = f.res_area[i];
  const double inflow_ext = f.res_external_inflow[i];
  double inflow;
  if (inflow_ext < 0.0) {  // abstraction limited to 98 % of the storage
    const double abstraction = jmin(-inflow_ext, (storage0 / dt) * 0.98);
    f.res_actual_external_abstraction_cumulative[i] += abstraction * dt;
    inflow = -abstraction;
  } else {
    inflow = inflow_ext;
  }
  inflow = q_river + f.res_inflow_overland[i] + f.res_inflow_subsurface[i] + inflow;
  // limit reservoir evaporation based on total available volume
  const double precipitation = f.res_precipitation[i] * area;
  const double available_storage = storage0 + (inflow + precipitation) * dt;
  const double potential_evaporation = f.res_evaporation[i] * area;
  const double evaporation = jmin(available_storage / dt, potential_evaporation);
  const double outflow_obs = f.res_outflow_obs[i];
  const int type = (int)f.res_outflow_curve_type[i];
  const double max_storage = f.res_maximum_storage[i];
  double outflow = 0.0, storage = storage0;
  if (outflow_obs == outflow_obs) {            // update_reservoir_outflow_obs  reservoir.jl:556-577
    const double storage_input = jmax(storage0 / dt + precipitation - evaporation + inflow, 0.0);
    outflow = jmin(outflow_obs, storage_input);
    storage = (storage_input - outflow) * dt;
    if (max_storage == max_storage) {
      const double overflow = jmax(0.0, (storage - max_storage) / dt);
      storage -= overflow * dt;
      outflow += overflow;
    }
  } else if (type == 2) {                      // update_reservoir_free_weir, no linked lower
    const double storage_input =               // reservoir (diff_wl = 0)      reservoir.jl:487-553
        jmax(storage0 / dt + precipitation - evaporation + inflow, 0.0);
    const double wl = f.res_waterlevel[i], thr = f.res_threshold[i];
    if (wl > thr) {
      const double dh = wl - thr;
      outflow = f.res_rating_curve_coefficient[i] * jpow(dh, f.res_rating_curve_exponent[i]);
      outflow = jmin(outflow, dh * area / dt);
    }
    storage = (storage_input - outflow) * dt;
  } else if (type == 3) {                      // update_reservoir_modified_puls  reservoir.jl:427-456
    const double res_factor = area / (dt * sqrt(f.res_rating_curve_coefficient[i]));
    const double si_factor = storage0 / dt + precipitation - evaporation + inflow;
    const double si_factor_adj = si_factor - area * f.res_threshold[i] / dt;
    if (si_factor_adj > 0.0) {
      const double qs = -res_factor + sqrt((res_factor * res_factor + 4 * si_factor_adj));
      outflow = qs > 0.0 ? 0.25 * (qs * qs) : 0.0;
    }
    outflow = jmin(outflow, si_factor);
    storage = (si_factor - outflow) * dt;
  } else if (type == 4) {                      // update_reservoir_simple       reservoir.jl:389-421
    storage = storage0 + (inflow + precipitation - evaporation) * dt;
    storage = jmax(storage, 0.0);
    const double fill_fraction = storage / max_storage;
    const double fac = scurve(fill_fraction, f.res_target_minimum_fraction[i], 1.0, 30.0);
    const double demand_release = jmin(fac * f.res_demand[i], storage / dt);
    storage -= demand_release * dt;
    const double release_wanted =
        jmax(0.0, (storage - max_storage * f.res_target_full_fraction[i]) / dt);
    const double overflow_q = jmax(0.0, (storage - max_storage) / dt);
    const double release_realized =
        jmin(release_wanted, overflow_q + f.res_maximum_release[i] - demand_release);
    storage -= release_realized * dt;
    outflow = release_realized + demand_release;
  }
  // linear storage curve (ReservoirProfileType.linear)
  f.res_waterlevel[i] = f.res_waterlevel[i] + (storage - storage0) / area;
  f.res_storage[i] = storage;
  f.res_outflow[i] = outflow;
  f.res_inflow_cumulative[i] += inflow * dt;
  f.res_outflow_cumulative[i] += outflow * dt;
  f.res_actevap_cumulative[i] += evaporation / area * dt;
  return outflow;
}


}  // namespace wfb
