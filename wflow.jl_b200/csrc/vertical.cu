// vertical.cu -- the SBM vertical land-surface update as fused sm_100a elementwise kernels.
//
// One thread per land slot; every input array is read once and every reference-visible output
// array is written once (the ~30 sweeps of the reference collapse into three kernels):
//   land_hydrology_kernel    update_land_hydrology_model!     sbm.jl:82-132
//   soil_water_storage_kernel update_soil_water_storage!      soil/soil.jl:1294-1392
//   total_water_storage_kernel update_total_water_storage!    sbm.jl:143-182
// HBM-bound: consecutive threads touch consecutive doubles of every SoA array, so each warp
// load/store is a fully used 256-byte transaction; the layered state lives in registers
// (template N) between sub-processes. Arithmetic order follows the reference expression by
// expression (no FMA contraction: -fmad=false) so results match the Julia code to the last
// bits that libm differences allow. All reference paths are under /root/reference/Wflow/src.
#include <cstdio>
#include "device_math.cuh"
#include "kernels.cuh"
#include "model.cuh"

namespace wfb {

namespace {

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// hydraulic_conductivity_at_depth, KvExponential / KvExponentialConstant   utils.jl:727-760
__device__ __forceinline__ double kv_at_depth(int profile, double kvfac, double kv_0, double f,
                                              double z_exp, double z) {
  if (profile == 1 && !(z < z_exp)) return kvfac * kv_0 * exp(-f * z_exp);
  return kvfac * kv_0 * exp(-f * z);
}

// unsatzone_flow_layer                                           soil/soil_process.jl:51-92
// split in two: the part every cell executes once per layer (`setup`: the transfer of water
// above saturation and the number `its` of explicit sub-iterations), and the sub-iteration
// loop itself, whose trip count is data dependent (0 for a dry layer, > 100 for a wet one).
#ifdef WFB_UNSAT_HIST
__device__ unsigned long long g_unsat_hist[40];
#endif
struct UnsatTask {
  double usd, sum_ast, kv_it, l_sat, c;
  int its;
};
__device__ __forceinline__ UnsatTask unsatzone_flow_setup(double usd, double kv_z, double l_sat,
                                                          double c, double dt) {
  UnsatTask t;
  t.l_sat = l_sat; t.c = c; t.kv_it = 0.0; t.its = 0;
  if (usd <= 0.0) { t.usd = 0.0; t.sum_ast = 0.0; return t; }
  const double st_sat = jmax(0.0, usd - l_sat);
  const double st = kv_z * bounded_power(usd / l_sat, c);
  const double sum_ast = jmin(st, st_sat / dt);
  usd -= sum_ast * dt;
  const double remainder = jmin((st - sum_ast) * dt, usd);
  const int its = (int)jcld(remainder, 2e-4);
  t.usd = usd; t.sum_ast = sum_ast; t.its = its;
  t.kv_it = kv_z / (double)its;
#ifdef WFB_UNSAT_HIST
  { int b = 0; while ((1 << b) <= its && b < 30) ++b; atomicAdd(&g_unsat_hist[b], 1ull); atomicAdd(&g_unsat_hist[32], (unsigned long long)its); }
#endif
  return t;
}
__device__ __forceinline__ void unsatzone_flow_iterate(UnsatTask& t, double dt) {
  double usd = t.usd, sum_ast = t.sum_ast;
  for (int k = 0; k < t.its; ++k) {
    const double st = t.kv_it * bounded_power(usd / t.l_sat, t.c);
    const double st_max = usd / dt;
    if (st < st_max) { usd -= st * dt; sum_ast += st; }
    else { usd = 0.0; sum_ast += st_max; break; }
  }
  t.usd = usd; t.sum_ast = sum_ast;
}

// The sub-iteration loops of the 256 cells of a CTA, re-balanced: with one cell per lane a warp
// runs as long as its wettest cell while most lanes idle (measured: 2.7 of 32 lanes active in
// this loop, 80 % of the kernel's time). The CTA therefore counting-sorts its 256 tasks by
// trip count in shared memory and lane r executes the task of rank r, so the lanes of a warp
// run loops of nearly equal length; results travel back through shared memory. Every thread
// of the CTA must call this (it synchronises), also those without a task (its = 0).
constexpr int kVertBlock = 256;
struct UnsatShared {
  double usd[kVertBlock], sum_ast[kVertBlock], kv_it[kVertBlock], l_sat[kVertBlock], c[kVertBlock];
  int its[kVertBlock];
  int bin[kVertBlock];          // histogram / running offsets of the trip counts (clamped)
  unsigned short owner[kVertBlock];
  int warp_sum[kVertBlock / 32];
  int cta_max;
};
__device__ __forceinline__ void unsatzone_flow_balanced(UnsatTask& t, double dt, UnsatShared& sh) {
  const int tid = (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // cheap uniform exit: nothing to balance when every loop is short
  const int wmax = __reduce_max_sync(0xffffffffu, t.its);
  if (tid == 0) sh.cta_max = 0;
  __syncthreads();
  if (lane == 0 && wmax > 0) atomicMax(&sh.cta_max, wmax);
  sh.bin[tid] = 0;
  __syncthreads();
  if (sh.cta_max <= 2) {  // same decision in every thread
    unsatzone_flow_iterate(t, dt);
    return;
  }
  const int key = t.its < kVertBlock - 1 ? t.its : kVertBlock - 1;
  atomicAdd(&sh.bin[key], 1);
  sh.usd[tid] = t.usd; sh.sum_ast[tid] = t.sum_ast; sh.kv_it[tid] = t.kv_it;
  sh.l_sat[tid] = t.l_sat; sh.c[tid] = t.c; sh.its[tid] = t.its;
  __syncthreads();
  // exclusive scan of the 256 bins (descending key order: longest loops first)
  const int cnt = sh.bin[kVertBlock - 1 - tid];
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) sh.warp_sum[wid] = incl;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int w2 = 0; w2 < kVertBlock / 32; ++w2)
    if (w2 < wid) base += sh.warp_sum[w2];
  __syncthreads();
  sh.bin[kVertBlock - 1 - tid] = base + incl - cnt;
  __syncthreads();
  const int rank = atomicAdd(&sh.bin[key], 1);
  sh.owner[rank] = (unsigned short)tid;
  __syncthreads();
  const int j = sh.owner[tid];
  UnsatTask u;
  u.usd = sh.usd[j]; u.sum_ast = sh.sum_ast[j]; u.kv_it = sh.kv_it[j]; u.l_sat = sh.l_sat[j];
  u.c = sh.c[j]; u.its = sh.its[j];
  unsatzone_flow_iterate(u, dt);
  sh.usd[j] = u.usd; sh.sum_ast[j] = u.sum_ast;
  __syncthreads();
  t.usd = sh.usd[tid]; t.sum_ast = sh.sum_ast[tid];
  __syncthreads();  // the arrays are reused by the next layer
}

// rwu_reduction_feddes                                         soil/soil_process.jl:183-200
__device__ __forceinline__ double rwu_reduction_feddes(double h, double h1, double h2, double h3,
                                                       double h4, double alpha_h1) {
  if (h < h4) return 0.0;
  if (h < h3) return (h - h4) / (h3 - h4);
  if (alpha_h1 == 0.0) {
    if (h < h2) return 1.0;
    if (h < h1) return (h1 - h) / (h1 - h2);
    return 0.0;
  }
  return 1.0;
}

}  // namespace

// Two passes of the same code. LIGHT (HEAVY = false): one thread per land slot; a cell whose
// Brooks-Corey loop needs more than kLightIters sub-iterations in some layer is appended to
// `heavy_list` and left untouched (its read-modify-write states are stored only after the
// unsaturated-zone section; the pure outputs it has written by then are rewritten with the
// same values later). HEAVY: one thread per entry of the list, loops re-balanced over the CTA
// (unsatzone_flow_balanced). With per-cell independent forcing ~1 task in 40 runs 16-250
// sub-iterations while 9 in 10 run one; a single pass leaves 2.7 of 32 lanes busy in the loop.
constexpr int kLightIters = 3;
template <int N, bool HEAVY>
__global__ void __launch_bounds__(kVertBlock)
land_hydrology_kernel(const DevFields f, const KCfg c, const double dt, int32_t* heavy_list,
                      unsigned* heavy_count) {
  __shared__ UnsatShared sh;
  const int i_raw = blockIdx.x * blockDim.x + threadIdx.x;
  int i;
  bool tail = false;
  if (HEAVY) {
    // the threads past the end of the list compute a copy of its last cell (they take part in
    // the CTA-wide re-balancing) and store nothing
    const int count = (int)*heavy_count;
    if ((int)(blockIdx.x * blockDim.x) >= count) return;
    tail = i_raw >= count;
    i = heavy_list[tail ? count - 1 : i_raw];
  } else {
    if (i_raw >= c.n) return;
    i = i_raw;
  }
  const int ns = c.ns;
  double st_canopy = 0.0, st_snoww = 0.0, st_gstore = 0.0, st_snow = 0.0, st_tsoil = 0.0;

  // ---- forcing ---------------------------------------------------------------------------
  const double P = __ldg(f.precipitation + i);
  const double PET = __ldg(f.potential_evaporation + i);
  const double T = __ldg(f.temperature + i);

  // ---- interception (canopy.jl:54-163, rainfall_interception.jl:9-130) ----------------------
  double cmax, gap;
  if (c.has_lai) {
    const double lai = __ldg(f.leaf_area_index + i);
    cmax = __ldg(f.storage_specific_leaf + i) * lai + __ldg(f.storage_wood + i);
    gap = exp(-__ldg(f.light_extinction_coefficient + i) * lai);
    if (!tail) f.maximum_canopy_storage[i] = cmax;
    if (!tail) f.canopy_gap_fraction[i] = gap;
  } else {
    cmax = __ldg(f.maximum_canopy_storage + i);
    gap = __ldg(f.canopy_gap_fraction + i);
  }
  const double kc = __ldg(f.crop_coefficient + i);
  const double canopy_potevap = kc * PET * (1.0 - gap);
  double throughfall, interception, stemflow;
  if (c.gash) {
    double e_r;
    if (c.has_lai) {
      const double canopyfraction = 1.0 - gap;
      const double ewet = canopyfraction * PET * kc;
      const double thr = 1e-4 * (1e-3 * (1.0 / dt));  // to_SI(1e-4, MM_PER_DT; dt)
      e_r = P > 0.0 ? jmin(0.25, ewet / jmax(thr, canopyfraction * P)) : 0.0;
      if (!tail) f.evaporation_to_precipitation_ratio[i] = e_r;
    } else {
      e_r = __ldg(f.evaporation_to_precipitation_ratio + i);
    }
    if (cmax > 0.0) {
      double frac_stem, frac_int, p_sat;
      if (gap < 1.0 / 1.1) {
        frac_stem = 0.1 * gap;
        frac_int = 1.0 - 1.1 * gap;
        // e_r == 0 gives -Inf * 0 = NaN here, and `P > NaN` is false: kept on purpose
        p_sat = e_r > frac_int ? 0.0 : -cmax / (e_r * dt) * log(1.0 - e_r / frac_int);
      } else {
        frac_stem = 1.0 - gap;
        frac_int = 0.0;
        p_sat = 0.0;
      }
      if (P > p_sat) {
        const double iwet = frac_int * p_sat - cmax / dt;
        const double isat = e_r * (P - p_sat);
        const double idry = cmax / dt;
        interception = iwet + isat + idry;
      } else {
        interception = frac_int * P;
      }
      stemflow = frac_stem * P;
      throughfall = P - interception - stemflow;
      if (interception > canopy_potevap) {
        const double drainage = interception - canopy_potevap;
        interception = canopy_potevap;
        throughfall += drainage;
      }
    } else {
      throughfall = P; interception = 0.0; stemflow = 0.0;
    }
  } else {
    double cs = f.canopy_storage[i];
    double frac_stem, p_canopy;
    if (gap < 1.0 / 1.1) {
      frac_stem = 0.1 * gap;
      p_canopy = (1.0 - gap - frac_stem) * P;
    } else {
      frac_stem = 1.0 - gap;
      p_canopy = 0.0;
    }
    stemflow = frac_stem * P;
    throughfall = gap * P;
    if (cs > cmax) { const double d = cs - cmax; cs = cmax; throughfall += d / dt; }
    cs += p_canopy * dt;
    const double max_evap = cs / dt;
    if (canopy_potevap > max_evap) { interception = max_evap; cs = 0.0; }
    else { interception = canopy_potevap; cs -= interception * dt; }
    if (cs > cmax) { const double d = cs - cmax; cs = cmax; throughfall += d / dt; }
    st_canopy = cs;
  }
  if (!tail) f.canopy_potevap[i] = canopy_potevap;
  if (!tail) f.throughfall[i] = throughfall;
  if (!tail) f.interception_rate[i] = interception;
  if (!tail) f.stemflow[i] = stemflow;

  // ---- snow (snow.jl:123-177, snow_process.jl:26-116) and glacier (glacier_process.jl:27-62)
  double water_flux_surface;
  double gfrac = 0.0;
  const bool glac = c.snow && c.glacier;
  if (c.snow) {
    const double eff = throughfall + stemflow;
    const double tti = __ldg(f.temperature_interval_snowfall + i);
    const double tt = __ldg(f.temperature_threshold_snowfall + i);
    double rainfrac;
    if (tti == 0.0) rainfrac = T > tt ? 1.0 : 0.0;
    else rainfrac = jclamp((T - (tt - tti / 2.0)) / tti, 0.0, 1.0);
    const double snowfrac = 1.0 - rainfrac;
    const double snow_precip = snowfrac * 1.0 * eff;
    const double liquid_precip = rainfrac * 1.0 * eff;
    double snow = f.snow_storage[i], snoww = f.snow_water[i];
    const double ttm = __ldg(f.temperature_threshold_melt + i);
    const double cfmax = __ldg(f.degree_day_factor + i);
    const double whc = __ldg(f.water_holding_capacity + i);
    double snow_melt;
    if (T > ttm) {
      const double pot = cfmax * (T - ttm);
      snow_melt = jmin(pot, snow / dt);
      snow -= snow_melt * dt;
      snoww += snow_melt * dt;
    } else {
      snow_melt = 0.0;
      const double potrefr = cfmax * 0.05 * (ttm - T);
      const double refr = jmin(potrefr * dt, snoww);
      snow += refr;
      snoww -= refr;
    }
    snow = jmax(snow, 0.0);
    snoww = jmax(snoww, 0.0);
    snow += snow_precip * dt;
    snoww += liquid_precip * dt;
    const double maxw = snow * whc;
    double snow_runoff;
    if (snoww > maxw) { snow_runoff = (snoww - maxw) / dt; snoww = maxw; }
    else snow_runoff = 0.0;
    if (!tail) f.effective_precip[i] = eff;
    if (!tail) f.snow_precip[i] = snow_precip;
    if (!tail) f.liquid_precip[i] = liquid_precip;
    st_snoww = snoww;
    if (!tail) f.snow_water_equivalent[i] = snoww + snow;
    if (!tail) f.snow_melt[i] = snow_melt;
    if (!tail) f.snow_runoff[i] = snow_runoff;
    double gmelt = 0.0;
    if (glac) {
      gfrac = __ldg(f.glacier_fraction + i);
      double gstore = f.glacier_store[i];
      const double maxrate = 8.0 * WFB_MM_PER_DAY;  // glacier.jl:97
      double s2g = gfrac > 0.0 ? __ldg(f.glacier_snow_to_ice_fraction + i) * snow : 0.0;
      s2g = jmin(s2g, maxrate);
      snow -= s2g * gfrac * dt;
      gstore += s2g * dt;
      const double gttm = __ldg(f.glacier_temperature_threshold_melt + i);
      const double pot = T > gttm ? __ldg(f.glacier_degree_day_factor + i) * (T - gttm) : 0.0;
      gmelt = snow < 1e-2 ? jmin(pot, gstore / dt) : 0.0;
      gstore -= gmelt * dt;
      st_gstore = gstore;
      if (!tail) f.glacier_melt[i] = gmelt;
    }
    st_snow = snow;
    water_flux_surface = snow_runoff + gmelt * gfrac;  // runoff.jl:48-58
  } else {
    water_flux_surface = throughfall + stemflow;       // runoff.jl:37-46
  }
  if (!tail) f.runoff_water_flux_surface[i] = water_flux_surface;

  // ---- open-water runoff (runoff.jl:61-111) ------------------------------------------------
  const double rf = __ldg(f.river_fraction + i), wf = __ldg(f.water_fraction + i);
  const double h_land = __ldg(f.olf_h + i);
  const double h_river = __ldg(f.waterdepth_river + i);  // refreshed by scatter_river_depth_kernel
  if (!tail) f.waterdepth_land[i] = h_land;
  const double runoff_river = jmin(1.0, rf) * water_flux_surface;
  const double runoff_land = jmin(1.0, wf) * water_flux_surface;
  const double aeow_river = rf * jmin(h_river / dt, PET);
  const double aeow_land = wf * jmin(h_land / dt, PET);
  if (!tail) f.runoff_river[i] = runoff_river;
  if (!tail) f.runoff_land[i] = runoff_land;
  if (!tail) f.actual_open_water_evaporation_river[i] = aeow_river;
  if (!tail) f.actual_open_water_evaporation_land[i] = aeow_land;
  if (!tail) f.net_runoff_river[i] = runoff_river - aeow_river;

  // ---- soil boundary conditions (soil.jl:643-682) ------------------------------------------
  const double soil_fraction = jmax(gap - wf - rf - gfrac, 0.0);
  const double pot_transp = jmax(0.0, canopy_potevap - interception);
  const double pot_soilevap0 = soil_fraction * PET;
  const double wfs = jmax(water_flux_surface - runoff_river - runoff_land, 0.0);
  if (!tail) f.soil_fraction[i] = soil_fraction;
  if (!tail) f.potential_transpiration[i] = pot_transp;
  if (!tail) f.potential_soilevaporation[i] = pot_soilevap0;
  if (!tail) f.soil_water_flux_surface[i] = wfs;

  // ---- state -> diagnostics (soil.jl:1400-1436) --------------------------------------------
  const double theta_s = __ldg(f.theta_s + i), theta_r = __ldg(f.theta_r + i);
  const double theta_fc = __ldg(f.theta_fc + i);
  const double theta_e = theta_s - theta_r;
  const double d_soil = __ldg(f.soil_thickness + i);
  const double swc = __ldg(f.soil_water_capacity + i);
  const double satwd = f.saturated_water_depth[i];
  const int nlayers = f.number_of_layers[i];
  double uld[N], ult[N], alt[N], cld[N + 1];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    uld[k] = f.unsaturated_layer_depth[k * ns + i];
    alt[k] = __ldg(f.actual_layer_thickness + k * ns + i);
    cld[k] = __ldg(f.cumulative_layer_depth + k * ns + i);
  }
  cld[N] = __ldg(f.cumulative_layer_depth + N * ns + i);
  double ustore_depth = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k < nlayers) ustore_depth += uld[k];
  const double zi = jmax(0.0, d_soil - satwd / theta_e);
  const double theta_d = jmax(theta_s - theta_fc, 0.02);  // lower_bound_drainable_porosity
  double drainable = (d_soil - zi) * theta_d;
  double ustore_cap = swc - satwd - ustore_depth;
  int n_unsat = N;
#pragma unroll
  for (int k = 0; k < N; ++k) {  // set_layerthickness utils.jl:390-404
    double t = qnan();
    if (zi > cld[k + 1]) t = alt[k];
    else if (zi - cld[k] > 0.0) t = zi - cld[k];
    ult[k] = t;
    n_unsat -= (t != t) ? 1 : 0;
    if (!tail) f.unsaturated_layer_thickness[k * ns + i] = t;
  }
  if (!tail) f.water_table_depth[i] = zi;
  if (!tail) f.n_unsatlayers[i] = n_unsat;
  if (!tail) f.total_soil_water_storage[i] = satwd + ustore_depth;

  // ---- soil temperature, infiltration (soil.jl:685-755, soil_process.jl:16-41,229-244) -------
  double f_red = 1.0;
  if (c.snow) {
    double tsoil = f.soil_surface_temperature[i];
    tsoil = tsoil + __ldg(f.w_soil + i) * (T - tsoil);
    st_tsoil = tsoil;
    if (c.soil_infiltration_reduction) {
      const double cf = __ldg(f.cf_soil + i);
      const double bb = 1.0 / (1.0 - cf);
      f_red = scurve(tsoil, 0.0 + 273.15, bb, 8.0) + cf;
    }
  }
  if (!tail) f.f_infiltration_reduction[i] = f_red;
  const double pathfrac = __ldg(f.compacted_soil_area_fraction + i);
  const double cap_soil = __ldg(f.infiltration_capacity_soil + i);
  const double cap_path = __ldg(f.infiltration_capacity_compacted_soil + i);
  const double soilinf = wfs * (1.0 - pathfrac);
  const double pathinf = wfs * pathfrac;
  const double max_infiltsoil = jmin(cap_soil * f_red, soilinf);
  const double max_infiltpath = jmin(cap_path * f_red, pathinf);
  const double infiltration = jmin(max_infiltpath + max_infiltsoil, jmax(0.0, ustore_cap / dt));
  const double infiltration_excess = (soilinf - max_infiltsoil) + (pathinf - max_infiltpath);
  if (!tail) f.infiltration[i] = infiltration;
  if (!tail) f.infiltration_excess[i] = infiltration_excess;

  // ---- unsaturated zone flow, Brooks-Corey (soil.jl:764-804) -------------------------------
  const double kv_0 = __ldg(f.kv_0 + i);
  const double fpar = __ldg(f.hydraulic_conductivity_scale_parameter + i);
  const double z_exp = c.kv_profile == 1 ? __ldg(f.z_exp + i) : 0.0;
  double bc[N], kvfac[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    bc[k] = __ldg(f.brooks_corey_exponent + k * ns + i);
    kvfac[k] = __ldg(f.vertical_hydraulic_conductivity_factor + k * ns + i);
  }
  double transfer = 0.0;
  {
    double z = 0.0, flow = 0.0;
    bool heavy = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      UnsatTask t;
      t.usd = 0.0; t.sum_ast = 0.0; t.kv_it = 0.0; t.l_sat = 1.0; t.c = 1.0; t.its = 0;
      const bool in_layer = k < n_unsat && !heavy;
      if (in_layer) {
        z = (k == 0) ? ult[0] : z + ult[k];
        const double l_sat = ult[k] * theta_e;
        const double kv_z = kv_at_depth(c.kv_profile, kvfac[k], kv_0, fpar, z_exp, z);
        const double usd = (k == 0) ? uld[k] + infiltration * dt : uld[k] + flow * dt;
        t = unsatzone_flow_setup(usd, kv_z, l_sat, bc[k], dt);
        if (tail) t.its = 0;
      }
      if (HEAVY) {
        unsatzone_flow_balanced(t, dt, sh);
      } else if (t.its > kLightIters) {
        heavy = true;
      } else {
        unsatzone_flow_iterate(t, dt);
      }
      if (in_layer) {
        uld[k] = t.usd;
        flow = t.sum_ast;
      }
    }
    if (!HEAVY && heavy) {  // left to the second pass
      heavy_list[atomicAdd(heavy_count, 1u)] = i;
      return;
    }
    if (n_unsat > 0) transfer = flow;
  }
  // the read-modify-write states of the sections above
  if (!tail) {
    if (!c.gash) f.canopy_storage[i] = st_canopy;
    if (c.snow) {
      f.snow_water[i] = st_snoww;
      f.snow_storage[i] = st_snow;
      f.soil_surface_temperature[i] = st_tsoil;
      if (c.glacier) f.glacier_store[i] = st_gstore;
    }
  }
  if (!tail) f.transfer[i] = transfer;

  // ---- soil evaporation (soil.jl:814-856, soil_process.jl:247-294) --------------------------
  double soilevap_sat, soil_evaporation;
  {
    double pot = pot_soilevap0;
    double evu;
    if (n_unsat == 0) evu = 0.0;
    else if (n_unsat == 1) evu = pot * jmin(1.0, uld[0] / (zi * theta_e));
    else evu = pot * jmin(1.0, uld[0] / (ult[0] * theta_e));
    evu = jmin(evu, uld[0] / dt);
    pot -= evu;
    uld[0] = uld[0] - evu * dt;
    if (n_unsat == 0 || n_unsat == 1) {
      const double e = pot * jmin(1.0, (alt[0] - zi) / alt[0]);
      soilevap_sat = jmin(e, (alt[0] - zi) * theta_d / dt);  // deliberately not clamped at 0
    } else {
      soilevap_sat = 0.0;
    }
    soil_evaporation = evu + soilevap_sat;
    drainable -= soilevap_sat * dt;
  }
  if (!tail) f.soil_evaporation_saturated_zone[i] = soilevap_sat;
  if (!tail) f.soil_evaporation[i] = soil_evaporation;

  // ---- transpiration (soil.jl:865-975) -----------------------------------------------------
  const double rd = __ldg(f.rooting_depth + i);
  const double h1 = __ldg(f.h1 + i), h2 = __ldg(f.h2 + i), h4 = __ldg(f.h4 + i);
  const double alpha_h1 = __ldg(f.alpha_h1 + i);
  const double hb = __ldg(f.air_entry_pressure + i);
  double h3;
  {
    const double tpot_daily = pot_transp / WFB_MM_PER_DAY;  // feddes_h3 soil_process.jl:166-176
    const double h3_high = __ldg(f.h3_high + i), h3_low = __ldg(f.h3_low + i);
    if (tpot_daily <= 1.0) h3 = h3_low;
    else if (tpot_daily < 5.0) h3 = h3_low + (h3_high - h3_low) * (tpot_daily - 1.0) / (5.0 - 1.0);
    else h3 = h3_high;
  }
  if (!tail) f.h3[i] = h3;
  double rootf[N];
#pragma unroll
  for (int k = 0; k < N; ++k) rootf[k] = __ldg(f.rootfraction + k * ns + i);
  double sum_rf = 0.0, rf_lowest = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k < n_unsat) {
      double rfu;
      if (k == n_unsat - 1 && zi < rd) {
        const double rootlength = jmin(alt[k], rd - cld[k]);
        rfu = rootf[k] * (ult[k] / rootlength);
      } else {
        rfu = rootf[k];
      }
      sum_rf += rfu;
      rf_lowest = rfu;
    }
  }
  double actevapustore = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k < n_unsat) {
      const double rfu = (k < n_unsat - 1) ? rootf[k] : rf_lowest;
      const double rfs = rd > 0.0 ? jmax(1.0 / sum_rf, 1.0) * rfu : 0.0;
      const double vwc = jmax(uld[k] / ult[k], 1e-7);
      // head_brooks_corey soil_process.jl:113-130
      const double par_lambda = 2.0 / (bc[k] - 3.0);
      const double head = par_lambda > 0.0 ? hb / jpow(vwc / theta_e, 1.0 / par_lambda) : hb;
      const double alpha = rwu_reduction_feddes(head, h1, h2, h3, h4, alpha_h1);
      const double availcap = jmin(1.0, jmax(0.0, (rd - cld[k]) / ult[k]));
      const double maxextr = uld[k] * availcap / dt;
      const double layer = jmin(alpha * rfs * pot_transp, maxextr);
      uld[k] = uld[k] - layer * dt;
      actevapustore += layer;
    }
  }
  const double wetroots = scurve(zi, rd, 1.0, __ldg(f.wet_root_distribution_parameter + i));
  const double alpha_sat = rwu_reduction_feddes(0.0, h1, h2, h3, h4, alpha_h1);
  const double restpottrans = pot_transp - actevapustore;
  const double ae_sat = jmin(restpottrans * wetroots * alpha_sat, drainable / dt);
  drainable -= ae_sat * dt;
  const double transpiration = actevapustore + ae_sat;
  if (!tail) f.actual_evaporation_unsaturated_store[i] = actevapustore;
  if (!tail) f.actual_evaporation_saturated_zone[i] = ae_sat;
  if (!tail) f.transpiration[i] = transpiration;

  // ---- actual infiltration and excess water (soil.jl:987-1043, 1178-1192) -------------------
  double excess = 0.0;
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    if (k < n_unsat) {
      excess = jmax(0.0, uld[k] - ult[k] * theta_e);
      uld[k] = uld[k] - excess;
      if (k > 0) uld[k - 1] = uld[k - 1] + excess;
    }
  }
  const double actual_infiltration = infiltration - excess / dt;
  if (!tail) f.actual_infiltration[i] = actual_infiltration;
  if (!tail) f.saturation_excess_water[i] = (wfs - actual_infiltration) - infiltration_excess;
  double actinf_soil, actinf_path;
  if (actual_infiltration > 0.0) {  // soil_process.jl:297-323
    actinf_soil = actual_infiltration * max_infiltsoil / (max_infiltpath + max_infiltsoil);
    actinf_path = actual_infiltration * max_infiltpath / (max_infiltpath + max_infiltsoil);
  } else {
    actinf_soil = 0.0; actinf_path = 0.0;
  }
  if (!tail) f.actual_infiltration_soil[i] = actinf_soil;
  if (!tail) f.actual_infiltration_compacted_soil[i] = actinf_path;
  if (!tail) f.excess_water_soil[i] = jmax(wfs * (1.0 - pathfrac) - actinf_soil, 0.0);
  if (!tail) f.excess_water_compacted_soil[i] = jmax(wfs * pathfrac - actinf_path, 0.0);

  // ---- recompute stores, capillary flux, leakage, recharge (soil.jl:1194-1209) --------------
  ustore_depth = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k < nlayers) ustore_depth += uld[k];
  ustore_cap = swc - satwd - ustore_depth;
  if (!tail) f.unsaturated_store_depth[i] = ustore_depth;
  if (!tail) f.unsaturated_store_capacity[i] = ustore_cap;
  double act_capflux = 0.0;
  if (n_unsat > 0) {  // capillary_flux! soil.jl:1050-1111
    double kvfac_nu = kvfac[0];
#pragma unroll
    for (int k = 1; k < N; ++k)
      if (k == n_unsat - 1) kvfac_nu = kvfac[k];
    const double ksat = kv_at_depth(c.kv_profile, kvfac_nu, kv_0, fpar, z_exp, zi);
    double mc = jmin(ksat, actevapustore);
    mc = jmin(mc, ustore_cap / dt);
    mc = jmin(mc, drainable / dt);
    const double maxcapflux = jmax(0.0, mc);
    double capflux = 0.0;
    if (zi > rd) {
      const double hmax = __ldg(f.cap_hmax + i);
      capflux = maxcapflux * jpow(1.0 - jmin(zi, hmax) / hmax, __ldg(f.cap_n + i));
    }
    double net = capflux;
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
      if (k < n_unsat) {
        const double toadd = jmin(net, jmax((ult[k] * theta_e - uld[k]) / dt, 0.0));
        uld[k] = uld[k] + toadd * dt;
        net -= toadd;
        act_capflux += toadd;
      }
    }
  }
  if (!tail) f.actual_capillary_flux[i] = act_capflux;
  double kvfac_nl = kvfac[0];
#pragma unroll
  for (int k = 1; k < N; ++k)
    if (k == nlayers - 1) kvfac_nl = kvfac[k];
  const double deepksat = kv_at_depth(c.kv_profile, kvfac_nl, kv_0, fpar, z_exp, d_soil);
  const double deeptransfer = jmin(drainable / dt, deepksat);
  const double leakage = jmax(0.0, jmin(__ldg(f.maximum_leakage + i), deeptransfer));
  if (!tail) f.actual_leakage[i] = leakage;
  if (!tail) f.recharge[i] = (transfer - act_capflux - leakage - ae_sat - soilevap_sat);
  // total AET (soil.jl:1206-1209) + interception (sbm.jl:130)
  double aet = soil_evaporation + transpiration + aeow_river + aeow_land + 0.0;
  aet += interception;
  if (!tail) f.actual_evapotranspiration[i] = aet;
  if (!tail) f.drainable_water_depth[i] = drainable;
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (!tail) f.unsaturated_layer_depth[k * ns + i] = uld[k];
}

// update_bc_open_water_runoff_model!: river h -> land grid                 runoff.jl:77-79
__global__ void scatter_river_depth_kernel(const DevFields f, const KCfg c) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.nriv) return;
  f.waterdepth_river[f.riv_land_slot[r]] = f.riv_h[r];
}

// recharge / water-table hand-off                                      sbm_model.jl:74-81
__global__ void exchange_recharge_kernel(const DevFields f, const KCfg c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  f.recharge_rate[i] = f.recharge[i];
  f.ssf_water_table_depth[i] = f.water_table_depth[i];
}

// update_soil_water_storage!                                        soil/soil.jl:1294-1392
template <int N>
__global__ void __launch_bounds__(256)
soil_water_storage_kernel(const DevFields f, const KCfg c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  const int ns = c.ns;
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

// update_total_water_storage!                                                sbm.jl:143-182
// river_slot_of_land[i] >= 0 marks a river cell (the reference's serial scatter loop is folded
// into the per-cell kernel through the inverse map).
__global__ void __launch_bounds__(256)
total_water_storage_kernel(const DevFields f, const KCfg c, const int32_t* __restrict__ riv_of_land) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  double total = 0.0;
  const int r = riv_of_land[i];
  if (r >= 0) total = (f.riv_h[r] * f.riv_flow_width[r] * f.riv_flow_length[r]) / f.area[i];
  const bool glac = c.snow && c.glacier;
  const double snow = c.snow ? f.snow_storage[i] : 0.0;
  const double snoww = c.snow ? f.snow_water[i] : 0.0;
  const double gl = glac ? f.glacier_store[i] * f.glacier_fraction[i] : 0.0;
  total += (((snow + snoww) + gl) + f.canopy_storage[i]) + 0.0;
  const double sub_surface = f.unsaturated_store_depth[i] + f.saturated_water_depth[i];
  const double lateral = f.olf_h[i] * (1.0 - f.river_fraction[i]);
  total += sub_surface + lateral;
  f.total_storage[i] = total;
}

// ---- launchers ----------------------------------------------------------------------------
#define WFB_DISPATCH_N(NN, ...)                       \
  switch (NN) {                                       \
    case 1: { constexpr int N = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int N = 2; __VA_ARGS__; break; } \
    case 3: { constexpr int N = 3; __VA_ARGS__; break; } \
    case 4: { constexpr int N = 4; __VA_ARGS__; break; } \
    case 5: { constexpr int N = 5; __VA_ARGS__; break; } \
    case 6: { constexpr int N = 6; __VA_ARGS__; break; } \
    case 7: { constexpr int N = 7; __VA_ARGS__; break; } \
    case 8: { constexpr int N = 8; __VA_ARGS__; break; } \
    default: return -1;                               \
  }

int launch_scatter_river_depth(const DevFields& f, const KCfg& c, cudaStream_t s) {
  if (c.nriv == 0) return 0;
  scatter_river_depth_kernel<<<(c.nriv + 255) / 256, 256, 0, s>>>(f, c);
  return 1;
}

#ifdef WFB_UNSAT_HIST
void dump_unsat_hist() {
  unsigned long long h[40];
  cudaMemcpyFromSymbol(h, g_unsat_hist, sizeof(h));
  unsigned long long z[40] = {0};
  cudaMemcpyToSymbol(g_unsat_hist, z, sizeof(z));
  fprintf(stderr, "unsat sub-iteration histogram (sum its %llu):", h[32]);
  for (int b = 0; b < 31; ++b) if (h[b]) fprintf(stderr, " [<%d]=%llu", 1 << b, h[b]);
  fprintf(stderr, "\n");
}
#endif
int launch_land_hydrology(const DevFields& f, const KCfg& c, int n_layers, double dt,
                          int32_t* heavy_list, unsigned* heavy_count, cudaStream_t s) {
  const int grid = (c.n + kVertBlock - 1) / kVertBlock;
  cudaMemsetAsync(heavy_count, 0, sizeof(unsigned), s);
  WFB_DISPATCH_N(n_layers, (land_hydrology_kernel<N, false><<<grid, kVertBlock, 0, s>>>(
                               f, c, dt, heavy_list, heavy_count)));
  // second pass over the listed cells; CTAs beyond the end of the list exit at once
  WFB_DISPATCH_N(n_layers, (land_hydrology_kernel<N, true><<<grid, kVertBlock, 0, s>>>(
                               f, c, dt, heavy_list, heavy_count)));
  return 2;
}

int launch_exchange_recharge(const DevFields& f, const KCfg& c, cudaStream_t s) {
  exchange_recharge_kernel<<<(c.n + 255) / 256, 256, 0, s>>>(f, c);
  return 1;
}

int launch_soil_water_storage(const DevFields& f, const KCfg& c, int n_layers, cudaStream_t s) {
  const int grid = (c.n + 255) / 256;
  WFB_DISPATCH_N(n_layers, (soil_water_storage_kernel<N><<<grid, 256, 0, s>>>(f, c)));
  return 1;
}

int launch_total_water_storage(const DevFields& f, const KCfg& c, const int32_t* riv_of_land,
                               cudaStream_t s) {
  total_water_storage_kernel<<<(c.n + 255) / 256, 256, 0, s>>>(f, c, riv_of_land);
  return 1;
}

}  // namespace wfb
