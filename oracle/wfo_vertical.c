/*
 * wfo_vertical.c -- CPU ORACLE (test infrastructure, see wfo.h): SBM vertical land update.
 * Sweep-by-sweep restatement of update_land_hydrology_model! (Wflow/src/sbm.jl:82-132),
 * update_soil_water_storage! (soil/soil.jl:1294-1392) and update_total_water_storage!
 * (sbm.jl:143-182). One OpenMP parallel-for per reference sweep.
 */
#include "wfo.h"
#include "wfo_math.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * field table
 * ---------------------------------------------------------------------------------------- */
static const char* k_names[] = {
#define X(name, kind) #name,
    WFO_FIELDS(X)
#undef X
};
static const int k_kinds[] = {
#define X(name, kind) kind,
    WFO_FIELDS(X)
#undef X
};
int wfo_num_fields(void) { return (int)(sizeof(k_kinds) / sizeof(int)); }
const char* wfo_field_name(int id) { return k_names[id]; }
int wfo_field_kind(int id) { return k_kinds[id]; }
wfo_model* wfo_new(void) { return (wfo_model*)calloc(1, sizeof(wfo_model)); }
void wfo_free(wfo_model* m) {
  if (m) { free(m->scratch); free(m->riv_reservoir); free(m); }
}
wfo_config* wfo_cfg(wfo_model* m) { return &m->cfg; }
int wfo_set_ptr(wfo_model* m, const char* name, double* p) {
#define X(nm, kind) if (strcmp(name, #nm) == 0) { m->nm = p; return 0; }
  WFO_FIELDS(X)
#undef X
  return -1;
}
int wfo_set_iptr(wfo_model* m, const char* name, int64_t* p) {
  if (!strcmp(name, "number_of_layers")) { m->number_of_layers = p; return 0; }
  if (!strcmp(name, "n_unsatlayers")) { m->n_unsatlayers = p; return 0; }
  if (!strcmp(name, "nlayers_kv")) { m->nlayers_kv = p; return 0; }
  if (!strcmp(name, "newton_trace_land")) { m->newton_trace_land = p; return 0; }
  if (!strcmp(name, "reservoir_river_indices")) {
    m->reservoir_river_indices = p;
    free(m->riv_reservoir);
    m->riv_reservoir = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m->cfg.nriv > 0 ? m->cfg.nriv : 1));
    for (int64_t r = 0; r < m->cfg.nriv; ++r) m->riv_reservoir[r] = -1;
    for (int64_t i = 0; i < m->cfg.nres; ++i) m->riv_reservoir[p[i]] = i;
    return 0;
  }
  if (!strcmp(name, "newton_trace_river")) { m->newton_trace_river = p; return 0; }
  if (!strcmp(name, "river_land_indices")) { m->river_land_indices = p; return 0; }
  if (!strcmp(name, "edge_x_up")) { m->edge_x_up = p; return 0; }
  if (!strcmp(name, "edge_x_down")) { m->edge_x_down = p; return 0; }
  if (!strcmp(name, "edge_y_up")) { m->edge_y_up = p; return 0; }
  if (!strcmp(name, "edge_y_down")) { m->edge_y_down = p; return 0; }
  if (!strcmp(name, "land_river_indices")) { m->land_river_indices = p; return 0; }
  return -1;
}
void wfo_set_network(wfo_model* m, int which, const wfo_network* net) {
  if (which == 0) m->land = *net; else m->river = *net;
  int64_t need = m->cfg.n > m->cfg.nriv ? m->cfg.n : m->cfg.nriv;
  free(m->scratch);
  m->scratch = (double*)malloc(sizeof(double) * (size_t)(need > 0 ? need : 1));
}

/* ------------------------------------------------------------------------------------------
 * scalar kernels
 * ---------------------------------------------------------------------------------------- */

/* vegetation/rainfall_interception.jl:9-70 */
void wfo_rainfall_interception_gash(double cmax, double e_r, double gap, double p, double cs,
                                    double maxevap, double dt, double out[4]) {
  double throughfall, interception, stem_flow;
  if (cmax > 0.0) {
    double fraction_stemflow, fraction_interception, p_sat;
    if (gap < 1.0 / 1.1) {
      fraction_stemflow = 0.1 * gap;
      fraction_interception = 1.0 - 1.1 * gap;
      if (e_r > fraction_interception) p_sat = 0.0;
      else p_sat = -cmax / (e_r * dt) * log(1.0 - e_r / fraction_interception);
    } else {
      fraction_stemflow = 1.0 - gap;
      fraction_interception = 0.0;
      p_sat = 0.0;
    }
    int large_storms = p > p_sat; /* false when p_sat is NaN (IEEE), as in the reference */
    if (large_storms) {
      double iwet = fraction_interception * p_sat - cmax / dt;
      double isat = e_r * (p - p_sat);
      double idry = cmax / dt;
      interception = iwet + isat + idry;
    } else {
      interception = fraction_interception * p;
    }
    stem_flow = fraction_stemflow * p;
    throughfall = p - interception - stem_flow;
    if (interception > maxevap) {
      double canopy_drainage = interception - maxevap;
      interception = maxevap;
      throughfall += canopy_drainage;
    }
  } else {
    throughfall = p; interception = 0.0; stem_flow = 0.0;
  }
  out[0] = throughfall; out[1] = interception; out[2] = stem_flow; out[3] = cs;
}

/* vegetation/rainfall_interception.jl:78-130 */
void wfo_rainfall_interception_modrut(double p, double pe, double cs, double gap, double cmax,
                                      double dt, double out[4]) {
  double fraction_stemflow, p_canopy;
  if (gap < 1.0 / 1.1) {
    fraction_stemflow = 0.1 * gap;
    p_canopy = (1.0 - gap - fraction_stemflow) * p;
  } else {
    fraction_stemflow = 1.0 - gap;
    p_canopy = 0.0;
  }
  double stemflow = fraction_stemflow * p;
  double throughfall = gap * p;
  if (cs > cmax) {
    double drain = cs - cmax;
    cs = cmax;
    throughfall += drain / dt;
  }
  cs += p_canopy * dt;
  double max_evap = cs / dt, evap;
  if (pe > max_evap) { evap = max_evap; cs = 0.0; }
  else { evap = pe; cs -= evap * dt; }
  if (cs > cmax) {
    double drain = cs - cmax;
    cs = cmax;
    throughfall += drain / dt;
  }
  out[0] = throughfall; out[1] = evap; out[2] = stemflow; out[3] = cs;
}

/* snow/snow_process.jl:90-116 (rfcf = sfcf = 1) */
void wfo_precipitation_hbv(double p, double t, double tti, double tt, double out[2]) {
  double rainfrac;
  if (tti == 0.0) rainfrac = (t > tt) ? 1.0 : 0.0;
  else {
    double frac = (t - (tt - tti / 2.0)) / tti;
    rainfrac = jl_clamp(frac, 0.0, 1.0);
  }
  double snowfrac = 1.0 - rainfrac;
  out[0] = snowfrac * 1.0 * p;
  out[1] = rainfrac * 1.0 * p;
}

/* snow/snow_process.jl:26-70 (cfr = 0.05) */
void wfo_snowpack_hbv(double snow, double snowwater, double snow_precip, double liquid_precip,
                      double t, double ttm, double cfmax, double whc, double dt, double out[5]) {
  const double cfr = 0.05;
  double snow_melt;
  if (t > ttm) {
    double pot = cfmax * (t - ttm);
    snow_melt = jl_min(pot, snow / dt);
    snow -= snow_melt * dt;
    snowwater += snow_melt * dt;
  } else {
    snow_melt = 0.0;
    double potrefr = cfmax * cfr * (ttm - t);
    double refr = jl_min(potrefr * dt, snowwater);
    snow += refr;
    snowwater -= refr;
  }
  snow = jl_max(snow, 0.0);
  snowwater = jl_max(snowwater, 0.0);
  snow += snow_precip * dt;
  snowwater += liquid_precip * dt;
  double maxw = snow * whc, runoff;
  if (snowwater > maxw) { runoff = (snowwater - maxw) / dt; snowwater = maxw; }
  else runoff = 0.0;
  out[0] = snow; out[1] = snowwater; out[2] = snowwater + snow; out[3] = snow_melt;
  out[4] = runoff;
}

/* glacier/glacier_process.jl:27-62 */
void wfo_glacier_hbv(double gfrac, double gstore, double snow, double t, double ttm,
                     double cfmax, double sifrac, double maxrate, double dt, double out[4]) {
  double s2g = gfrac > 0.0 ? sifrac * snow : 0.0;
  s2g = jl_min(s2g, maxrate);
  snow -= s2g * gfrac * dt;
  gstore += s2g * dt;
  double pot = (t > ttm) ? cfmax * (t - ttm) : 0.0;
  double melt = (snow < 1e-2) ? jl_min(pot, gstore / dt) : 0.0;
  gstore -= melt * dt;
  out[0] = snow; out[1] = s2g; out[2] = gstore; out[3] = melt;
}

/* soil/soil_process.jl:16-41 */
void wfo_infiltration(double pot, double pathfrac, double cap_soil, double cap_path,
                      double ustorecap, double f_red, double dt, double out[2]) {
  double soilinf = pot * (1.0 - pathfrac);
  double pathinf = pot * pathfrac;
  double max_infiltsoil = jl_min(cap_soil * f_red, soilinf);
  double max_infiltpath = jl_min(cap_path * f_red, pathinf);
  out[0] = jl_min(max_infiltpath + max_infiltsoil, jl_max(0.0, ustorecap / dt));
  out[1] = (soilinf - max_infiltsoil) + (pathinf - max_infiltpath);
}

/* soil/soil_process.jl:51-92 */
void wfo_unsatzone_flow_layer(double usd, double kv_z, double l_sat, double c, double dt,
                              double out[2]) {
  if (usd <= 0.0) { out[0] = 0.0; out[1] = 0.0; return; }
  double st_sat = jl_max(0.0, usd - l_sat);
  double st = kv_z * jl_bounded_power(usd / l_sat, c);
  double sum_ast = jl_min(st, st_sat / dt);
  usd -= sum_ast * dt;
  double remainder = jl_min((st - sum_ast) * dt, usd);
  int64_t its = (int64_t)wfo_cld(remainder, 2e-4);
  for (int64_t k = 0; k < its; ++k) {
    st = (kv_z / (double)its) * jl_bounded_power(usd / l_sat, c);
    double st_max = usd / dt;
    if (st < st_max) { usd -= st * dt; sum_ast += st; }
    else { usd = 0.0; sum_ast += st_max; break; }
  }
  out[0] = usd; out[1] = sum_ast;
}

/* soil/soil_process.jl:99-106 */
double wfo_vwc_brooks_corey(double h, double hb, double ts, double tr, double c) {
  if (h < hb) {
    double par_lambda = 2.0 / (c - 3.0);
    return (ts - tr) * jl_pow(hb / h, par_lambda) + tr;
  }
  return ts;
}

/* soil/soil_process.jl:113-130 */
double wfo_head_brooks_corey(double vwc, double ts, double tr, double c, double hb) {
  double par_lambda = 2.0 / (c - 3.0);
  if (par_lambda > 0.0) return hb / jl_pow(vwc / (ts - tr), 1.0 / par_lambda);
  return hb;
}

/* soil/soil_process.jl:166-176 ; from_SI(x, MM_PER_DAY) = x / ((1/86400)*1e-3) (units.jl:55-68) */
double wfo_feddes_h3(double h3_high, double h3_low, double tpot) {
  double tpot_daily = tpot / WFO_MM_PER_DAY;
  if (tpot_daily <= 1.0) return h3_low;
  if (tpot_daily < 5.0) return h3_low + (h3_high - h3_low) * (tpot_daily - 1.0) / (5.0 - 1.0);
  return h3_high;
}

/* soil/soil_process.jl:183-200 */
double wfo_rwu_reduction_feddes(double h, double h1, double h2, double h3, double h4,
                                double alpha_h1) {
  if (h < h4) return 0.0;
  if (h < h3) return (h - h4) / (h3 - h4);
  if (alpha_h1 == 0.0) {
    if (h < h2) return 1.0;
    if (h < h1) return (h1 - h) / (h1 - h2);
    return 0.0;
  }
  return 1.0;
}

/* soil/soil_process.jl:210-213 */
double wfo_soil_temperature(double tsoil, double w, double t) { return tsoil + w * (t - tsoil); }

/* utils.jl:27-30 */
double wfo_scurve(double x, double a, double b, double c) { return 1.0 / (b + exp(-c * (x - a))); }

/* soil/soil_process.jl:229-244 */
double wfo_infiltration_reduction_factor(double tsoil, double cf, int modelsnow, int flag) {
  if (modelsnow && flag) {
    double bb = 1.0 / (1.0 - cf);
    return wfo_scurve(tsoil, 0.0 + 273.15, bb, 8.0) + cf;
  }
  return 1.0;
}

/* soil/soil_process.jl:247-271 */
double wfo_soil_evaporation_unsaturated_store(double pot, double usd, double ust, int64_t nu,
                                              double zi, double theta_e) {
  if (nu == 0) return 0.0;
  if (nu == 1) return pot * jl_min(1.0, usd / (zi * theta_e));
  return pot * jl_min(1.0, usd / (ust * theta_e));
}

/* soil/soil_process.jl:274-294 (no clamp at zero, test/land_process.jl:444-452) */
double wfo_soil_evaporation_saturated_store(double pot, int64_t nu, double lt, double zi,
                                            double theta_d, double dt) {
  if (nu == 0 || nu == 1) {
    double e = pot * jl_min(1.0, (lt - zi) / lt);
    return jl_min(e, (lt - zi) * theta_d / dt);
  }
  return 0.0;
}

/* soil/soil_process.jl:297-323 */
void wfo_actual_infiltration_soil_path(double pot, double act, double pathfrac, double cap_soil,
                                       double cap_path, double f_red, double out[2]) {
  double soilinf = pot * (1.0 - pathfrac);
  double pathinf = pot * pathfrac;
  if (act > 0.0) {
    double max_infiltsoil = jl_min(cap_soil * f_red, soilinf);
    double max_infiltpath = jl_min(cap_path * f_red, pathinf);
    out[0] = act * max_infiltsoil / (max_infiltpath + max_infiltsoil);
    out[1] = act * max_infiltpath / (max_infiltpath + max_infiltsoil);
  } else { out[0] = 0.0; out[1] = 0.0; }
}

/* utils.jl:390-404 : layer thickness above a reference depth; NaN for inactive layers */
static void set_layerthickness(double ref_depth, const double* cum_depth, const double* thickness,
                               int64_t N, double* out) {
  for (int64_t k = 0; k < N; ++k) {
    out[k] = thickness[k] * NAN;
    if (ref_depth > cum_depth[k + 1]) out[k] = thickness[k];
    else if (ref_depth - cum_depth[k] > 0.0) out[k] = ref_depth - cum_depth[k];
  }
}
static int64_t number_of_active_layers(const double* t, int64_t N) {
  int64_t c = 0;
  for (int64_t k = 0; k < N; ++k) c += isnan(t[k]) ? 1 : 0;
  return N - c;
}

/* utils.jl:727-789 ; layer index n is 1-based like the reference */
static double kv_at_depth(const wfo_model* m, double z, int64_t i, int64_t n1) {
  const int64_t N = m->cfg.N;
  const double* kvfac = m->vertical_hydraulic_conductivity_factor + i * N;
  switch (m->cfg.kv_profile) {
    case 0:
      return kvfac[n1 - 1] * m->kv_0[i] * exp(-m->hydraulic_conductivity_scale_parameter[i] * z);
    case 1:
      if (z < m->z_exp[i])
        return kvfac[n1 - 1] * m->kv_0[i] * exp(-m->hydraulic_conductivity_scale_parameter[i] * z);
      return kvfac[n1 - 1] * m->kv_0[i] *
             exp(-m->hydraulic_conductivity_scale_parameter[i] * m->z_exp[i]);
    case 2:
      return kvfac[n1 - 1] * m->kv[i * N + n1 - 1];
    default:
      if (z < m->z_layered[i]) return kvfac[n1 - 1] * m->kv[i * N + n1 - 1];
      {
        int64_t nn = m->nlayers_kv[i];
        return kvfac[nn - 1] * m->kv[i * N + nn - 1] *
               exp(-m->hydraulic_conductivity_scale_parameter[i] * (z - m->z_layered[i]));
      }
  }
}

#define PFOR _Pragma("omp parallel for schedule(static)")

/* ------------------------------------------------------------------------------------------
 * sweeps
 * ---------------------------------------------------------------------------------------- */

/* vegetation/canopy.jl:146-163 */
static void update_canopy_parameters(wfo_model* m) {
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    m->maximum_canopy_storage[i] =
        m->storage_specific_leaf[i] * m->leaf_area_index[i] + m->storage_wood[i];
    m->canopy_gap_fraction[i] = exp(-m->light_extinction_coefficient[i] * m->leaf_area_index[i]);
  }
}

/* vegetation/canopy.jl:54-98 (Gash) and :115-143 (Rutter) */
static void update_interception_model(wfo_model* m, double dt) {
  const int64_t n = m->cfg.n;
  if (m->cfg.has_lai) update_canopy_parameters(m);
  if (m->cfg.gash) {
    if (m->cfg.has_lai) {
      /* to_SI(1e-4, MM_PER_DT; dt_val = dt) = 1e-4 * (1e-3 * dt^-1)   units.jl:256-277 */
      const double thr = 1e-4 * (1e-3 * (1.0 / dt));
      PFOR for (int64_t i = 0; i < n; ++i) {
        double canopyfraction = 1.0 - m->canopy_gap_fraction[i];
        double ewet = canopyfraction * m->potential_evaporation[i] * m->crop_coefficient[i];
        m->evaporation_to_precipitation_ratio[i] =
            m->precipitation[i] > 0.0
                ? jl_min(0.25, ewet / jl_max(thr, canopyfraction * m->precipitation[i]))
                : 0.0;
      }
    }
    PFOR for (int64_t i = 0; i < n; ++i) {
      double o[4];
      m->canopy_potevap[i] =
          m->crop_coefficient[i] * m->potential_evaporation[i] * (1.0 - m->canopy_gap_fraction[i]);
      wfo_rainfall_interception_gash(m->maximum_canopy_storage[i],
                                     m->evaporation_to_precipitation_ratio[i],
                                     m->canopy_gap_fraction[i], m->precipitation[i],
                                     m->canopy_storage[i], m->canopy_potevap[i], dt, o);
      m->throughfall[i] = o[0]; m->interception_rate[i] = o[1]; m->stemflow[i] = o[2];
      m->canopy_storage[i] = o[3];
    }
  } else {
    PFOR for (int64_t i = 0; i < n; ++i) {
      double o[4];
      m->canopy_potevap[i] =
          m->crop_coefficient[i] * m->potential_evaporation[i] * (1.0 - m->canopy_gap_fraction[i]);
      wfo_rainfall_interception_modrut(m->precipitation[i], m->canopy_potevap[i],
                                       m->canopy_storage[i], m->canopy_gap_fraction[i],
                                       m->maximum_canopy_storage[i], dt, o);
      m->throughfall[i] = o[0]; m->interception_rate[i] = o[1]; m->stemflow[i] = o[2];
      m->canopy_storage[i] = o[3];
    }
  }
}

/* snow/snow.jl:123-177 */
static void update_snow_model(wfo_model* m, double dt) {
  const int64_t n = m->cfg.n;
  if (!m->cfg.snow) return;
  PFOR for (int64_t i = 0; i < n; ++i) m->effective_precip[i] = m->throughfall[i] + m->stemflow[i];
  PFOR for (int64_t i = 0; i < n; ++i) {
    double o[2];
    wfo_precipitation_hbv(m->effective_precip[i], m->temperature[i],
                          m->temperature_interval_snowfall[i],
                          m->temperature_threshold_snowfall[i], o);
    m->snow_precip[i] = o[0]; m->liquid_precip[i] = o[1];
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    double o[5];
    wfo_snowpack_hbv(m->snow_storage[i], m->snow_water[i], m->snow_precip[i],
                     m->liquid_precip[i], m->temperature[i], m->temperature_threshold_melt[i],
                     m->degree_day_factor[i], m->water_holding_capacity[i], dt, o);
    m->snow_storage[i] = o[0]; m->snow_water[i] = o[1]; m->snow_water_equivalent[i] = o[2];
    m->snow_melt[i] = o[3]; m->snow_runoff[i] = o[4];
  }
}

/* accucapacityflux!                                              routing/utils.jl:82-109 */
void wfo_accucapacityflux(double* flux, double* material, const int64_t* order,
                          const int64_t* down, int64_t n, const double* capacity, double dt) {
  for (int64_t k = 0; k < n; ++k) {
    const int64_t v = order[k];
    const double flux_val = jl_min(material[v] / dt, capacity[v]);
    const double material_update = flux_val * dt;
    material[v] -= material_update;
    flux[v] = flux_val;
    if (down[v] >= 0) material[down[v]] += material_update;
  }
}

/* lateral_snow_transport!(snow, domain, dt)        routing/surface/surface_process.jl:9-19
 * (serial over the whole land network, like the reference) */
static void lateral_snow_transport(wfo_model* m, double dt) {
  if (!m->cfg.snow || !m->cfg.snow_transport) return;
  const int64_t n = m->cfg.n;
  const wfo_network* nw = &m->land;
  double* cap1 = (double*)malloc(sizeof(double) * (size_t)n * 3);
  double* cap2 = cap1 + n;
  double* flux2 = cap2 + n;
  const double snow_storage_max = 10.0, tan80 = 5.67;
  for (int64_t i = 0; i < n; ++i) {
    const double snowflux_frac = jl_min(0.5, m->slope[i] / tan80) *
                                 jl_min(1.0, m->snow_storage[i] / snow_storage_max);
    cap1[i] = snowflux_frac * m->snow_storage[i] / dt;   /* maxflux */
    cap2[i] = m->snow_water[i] * snowflux_frac / dt;
  }
  wfo_accucapacityflux(m->snow_out, m->snow_storage, nw->order, nw->down, n, cap1, dt);
  wfo_accucapacityflux(flux2, m->snow_water, nw->order, nw->down, n, cap2, dt);
  for (int64_t i = 0; i < n; ++i) m->snow_out[i] += flux2[i];
  /* flux_in!(snow_in, snow_out, network)                         routing/utils.jl:161-167 */
  for (int64_t k = 0; k < n; ++k) {
    const int64_t v = nw->order[k];
    double ssum = 0.0;
    for (int64_t u = nw->up_ptr[k]; u < nw->up_ptr[k + 1]; ++u) ssum += m->snow_out[nw->up_idx[u]];
    m->snow_in[v] = ssum;
  }
  free(cap1);
}

/* glacier/glacier.jl:122-154 ; only active when snow && glacier (sbm.jl:41-54) */
static void update_glacier_model(wfo_model* m, double dt) {
  if (!(m->cfg.snow && m->cfg.glacier)) return;
  const double maxrate = 8.0 * WFO_MM_PER_DAY; /* glacier.jl:97 */
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    double o[4];
    wfo_glacier_hbv(m->glacier_fraction[i], m->glacier_store[i], m->snow_storage[i],
                    m->temperature[i], m->glacier_temperature_threshold_melt[i],
                    m->glacier_degree_day_factor[i], m->glacier_snow_to_ice_fraction[i], maxrate,
                    dt, o);
    m->snow_storage[i] = o[0]; m->glacier_store[i] = o[2]; m->glacier_melt[i] = o[3];
  }
}

/* surfacewater/runoff.jl:37-111 */
static void update_open_water_runoff(wfo_model* m, double dt) {
  const int64_t n = m->cfg.n;
  const int glac = m->cfg.snow && m->cfg.glacier;
  if (m->cfg.snow) {
    PFOR for (int64_t i = 0; i < n; ++i)
      m->runoff_water_flux_surface[i] =
          m->snow_runoff[i] + (glac ? m->glacier_melt[i] * m->glacier_fraction[i] : 0.0 * 0.0);
  } else {
    PFOR for (int64_t i = 0; i < n; ++i)
      m->runoff_water_flux_surface[i] = m->throughfall[i] + m->stemflow[i];
  }
  PFOR for (int64_t i = 0; i < n; ++i) m->waterdepth_land[i] = m->olf_h[i];
  for (int64_t r = 0; r < m->cfg.nriv; ++r) m->waterdepth_river[m->river_land_indices[r]] = m->riv_h[r];
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->runoff_river[i] = jl_min(1.0, m->river_fraction[i]) * m->runoff_water_flux_surface[i];
    m->runoff_land[i] = jl_min(1.0, m->water_fraction[i]) * m->runoff_water_flux_surface[i];
    m->actual_open_water_evaporation_river[i] =
        m->river_fraction[i] * jl_min(m->waterdepth_river[i] / dt, m->potential_evaporation[i]);
    m->actual_open_water_evaporation_land[i] =
        m->water_fraction[i] * jl_min(m->waterdepth_land[i] / dt, m->potential_evaporation[i]);
    m->net_runoff_river[i] = m->runoff_river[i] - m->actual_open_water_evaporation_river[i];
  }
}

/* soil/soil.jl:643-682 (paddy / irrigation terms are Zeros without water demand) */
static void update_bc_soil_model(wfo_model* m) {
  const int glac = m->cfg.snow && m->cfg.glacier;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    double gf = glac ? m->glacier_fraction[i] : 0.0;
    m->soil_fraction[i] = jl_max(
        m->canopy_gap_fraction[i] - m->water_fraction[i] - m->river_fraction[i] - gf, 0.0);
    m->potential_transpiration[i] = jl_max(0.0, m->canopy_potevap[i] - m->interception_rate[i]);
    m->potential_soilevaporation[i] = m->soil_fraction[i] * m->potential_evaporation[i];
    m->soil_water_flux_surface[i] = jl_max(
        m->runoff_water_flux_surface[i] - m->runoff_river[i] - m->runoff_land[i], 0.0);
  }
}

/* soil/soil.jl:700-708 */
static void unsaturated_store_depth(wfo_model* m) {
  const int64_t N = m->cfg.N;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    double s = 0.0;
    for (int64_t k = 0; k < m->number_of_layers[i]; ++k) s += m->unsaturated_layer_depth[i * N + k];
    m->unsaturated_store_depth[i] = s;
  }
}

/* soil/soil.jl:1400-1436 */
void wfo_update_diagnostic_vars(wfo_model* m) {
  const int64_t N = m->cfg.N;
  unsaturated_store_depth(m);
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    m->water_table_depth[i] = jl_max(
        0.0, m->soil_thickness[i] - m->saturated_water_depth[i] / (m->theta_s[i] - m->theta_r[i]));
    m->drainable_water_depth[i] = (m->soil_thickness[i] - m->water_table_depth[i]) *
                                  jl_max(m->theta_s[i] - m->theta_fc[i], 0.02);
    m->unsaturated_store_capacity[i] =
        m->soil_water_capacity[i] - m->saturated_water_depth[i] - m->unsaturated_store_depth[i];
    set_layerthickness(m->water_table_depth[i], m->cumulative_layer_depth + i * (N + 1),
                       m->actual_layer_thickness + i * N, N,
                       m->unsaturated_layer_thickness + i * N);
    m->n_unsatlayers[i] = number_of_active_layers(m->unsaturated_layer_thickness + i * N, N);
    m->total_soil_water_storage[i] = m->saturated_water_depth[i] + m->unsaturated_store_depth[i];
  }
}

/* soil/soil.jl:764-804 */
static void unsaturated_zone_flow(wfo_model* m, double dt) {
  const int64_t N = m->cfg.N;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    double* uld = m->unsaturated_layer_depth + i * N;
    const double* ult = m->unsaturated_layer_thickness + i * N;
    const int64_t nu = m->n_unsatlayers[i];
    if (nu > 0) {
      double z = 0.0, flow_rate = 0.0;
      for (int64_t k = 0; k < nu; ++k) {
        z = (k == 0) ? ult[0] : z + ult[k]; /* cumsum(ult) */
        double l_sat = ult[k] * (m->theta_s[i] - m->theta_r[i]);
        double kv_z = kv_at_depth(m, z, i, k + 1);
        double usd = (k == 0) ? uld[k] + m->infiltration[i] * dt : uld[k] + flow_rate * dt;
        double o[2];
        wfo_unsatzone_flow_layer(usd, kv_z, l_sat, m->brooks_corey_exponent[i * N + k], dt, o);
        uld[k] = o[0]; flow_rate = o[1];
      }
      m->transfer[i] = flow_rate;
    } else {
      m->transfer[i] = 0.0;
    }
  }
}

/* soil/soil.jl:814-856 */
static void soil_evaporation(wfo_model* m, double dt) {
  const int64_t N = m->cfg.N;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    double* uld = m->unsaturated_layer_depth + i * N;
    const double* ult = m->unsaturated_layer_thickness + i * N;
    double potsoilevap = m->potential_soilevaporation[i];
    double evu = wfo_soil_evaporation_unsaturated_store(potsoilevap, uld[0], ult[0],
                                                        m->n_unsatlayers[i], m->water_table_depth[i],
                                                        m->theta_s[i] - m->theta_r[i]);
    evu = jl_min(evu, uld[0] / dt);
    potsoilevap -= evu;
    uld[0] = uld[0] - evu * dt;
    double theta_d = jl_max(m->theta_s[i] - m->theta_fc[i], 0.02);
    double evs = wfo_soil_evaporation_saturated_store(potsoilevap, m->n_unsatlayers[i],
                                                      m->actual_layer_thickness[i * N],
                                                      m->water_table_depth[i], theta_d, dt);
    m->soil_evaporation_saturated_zone[i] = evs;
    m->soil_evaporation[i] = evu + evs;
    m->drainable_water_depth[i] -= evs * dt;
  }
}

/* soil/soil.jl:865-975 */
static void transpiration(wfo_model* m, double dt) {
  const int64_t N = m->cfg.N;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    double* uld = m->unsaturated_layer_depth + i * N;
    const double* ult = m->unsaturated_layer_thickness + i * N;
    const double* alt = m->actual_layer_thickness + i * N;
    const double* cld = m->cumulative_layer_depth + i * (N + 1);
    const double* rf = m->rootfraction + i * N;
    const int64_t nu = m->n_unsatlayers[i];
    const double rd = m->rooting_depth[i];
    const double pt = m->potential_transpiration[i];
    m->h3[i] = wfo_feddes_h3(m->h3_high[i], m->h3_low[i], pt);
    double sum_rf = 0.0, rf_lowest = 0.0;
    for (int64_t k = 0; k < nu; ++k) {
      double rfu;
      if (k == nu - 1 && m->water_table_depth[i] < rd) {
        double rootlength = jl_min(alt[k], rd - cld[k]);
        rfu = rf[k] * (ult[k] / rootlength);
      } else rfu = rf[k];
      sum_rf += rfu;
      rf_lowest = rfu;
    }
    double actevapustore = 0.0;
    for (int64_t k = 0; k < nu; ++k) {
      double rfu = (k < nu - 1) ? rf[k] : rf_lowest;
      double rfs = rd > 0.0 ? jl_max(1.0 / sum_rf, 1.0) * rfu : 0.0;
      double vwc = jl_max(uld[k] / ult[k], 1e-7);
      double head = wfo_head_brooks_corey(vwc, m->theta_s[i], m->theta_r[i],
                                          m->brooks_corey_exponent[i * N + k],
                                          m->air_entry_pressure[i]);
      double alpha = wfo_rwu_reduction_feddes(head, m->h1[i], m->h2[i], m->h3[i], m->h4[i],
                                              m->alpha_h1[i]);
      double availcap = jl_min(1.0, jl_max(0.0, (rd - cld[k]) / ult[k]));
      double maxextr = uld[k] * availcap / dt;
      double layer = jl_min(alpha * rfs * pt, maxextr);
      double nuld = uld[k] - layer * dt;
      actevapustore += layer;
      uld[k] = nuld;
    }
    double wetroots = wfo_scurve(m->water_table_depth[i], rd, 1.0,
                                 m->wet_root_distribution_parameter[i]);
    double alpha = wfo_rwu_reduction_feddes(0.0, m->h1[i], m->h2[i], m->h3[i], m->h4[i],
                                            m->alpha_h1[i]);
    double rest = pt - actevapustore;
    double aesat = jl_min(rest * wetroots * alpha, m->drainable_water_depth[i] / dt);
    m->actual_evaporation_unsaturated_store[i] = actevapustore;
    m->actual_evaporation_saturated_zone[i] = aesat;
    m->drainable_water_depth[i] -= aesat * dt;
    m->transpiration[i] = actevapustore + aesat;
  }
}

/* soil/soil.jl:987-1017 */
static void actual_infiltration(wfo_model* m, double dt) {
  const int64_t N = m->cfg.N;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    double* uld = m->unsaturated_layer_depth + i * N;
    const double* ult = m->unsaturated_layer_thickness + i * N;
    double excess = 0.0;
    for (int64_t k = m->n_unsatlayers[i] - 1; k >= 0; --k) {
      excess = jl_max(0.0, uld[k] - ult[k] * (m->theta_s[i] - m->theta_r[i]));
      uld[k] = uld[k] - excess;
      if (k > 0) uld[k - 1] = uld[k - 1] + excess;
    }
    m->actual_infiltration[i] = m->infiltration[i] - excess / dt;
  }
}

/* soil/soil.jl:1050-1111 */
static void capillary_flux(wfo_model* m, double dt) {
  const int64_t N = m->cfg.N;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    double* uld = m->unsaturated_layer_depth + i * N;
    const double* ult = m->unsaturated_layer_thickness + i * N;
    const int64_t nu = m->n_unsatlayers[i];
    if (nu > 0) {
      double ksat = kv_at_depth(m, m->water_table_depth[i], i, nu);
      double mc = jl_min(ksat, m->actual_evaporation_unsaturated_store[i]);
      mc = jl_min(mc, m->unsaturated_store_capacity[i] / dt);
      mc = jl_min(mc, m->drainable_water_depth[i] / dt);
      double maxcapflux = jl_max(0.0, mc);
      double capflux = 0.0;
      if (m->water_table_depth[i] > m->rooting_depth[i])
        capflux = maxcapflux *
                  jl_pow(1.0 - jl_min(m->water_table_depth[i], m->cap_hmax[i]) / m->cap_hmax[i],
                         m->cap_n[i]);
      double net = capflux, act = 0.0;
      for (int64_t k = nu - 1; k >= 0; --k) {
        double toadd = jl_min(
            net, jl_max((ult[k] * (m->theta_s[i] - m->theta_r[i]) - uld[k]) / dt, 0.0));
        uld[k] = uld[k] + toadd * dt;
        net -= toadd;
        act += toadd;
      }
      m->actual_capillary_flux[i] = act;
    } else {
      m->actual_capillary_flux[i] = 0.0;
    }
  }
}

/* soil/soil.jl:1150-1211 */
static void update_soil_water_flow(wfo_model* m, double dt) {
  const int64_t n = m->cfg.n;
  wfo_update_diagnostic_vars(m);
  if (m->cfg.snow) { /* soil.jl:685-697 */
    PFOR for (int64_t i = 0; i < n; ++i)
      m->soil_surface_temperature[i] = wfo_soil_temperature(
          m->soil_surface_temperature[i], m->w_soil[i], m->temperature[i]);
  }
  PFOR for (int64_t i = 0; i < n; ++i)
    m->f_infiltration_reduction[i] = wfo_infiltration_reduction_factor(
        m->soil_surface_temperature[i], m->cf_soil[i], m->cfg.snow,
        m->cfg.soil_infiltration_reduction);
  PFOR for (int64_t i = 0; i < n; ++i) {
    double o[2];
    wfo_infiltration(m->soil_water_flux_surface[i], m->compacted_soil_area_fraction[i],
                     m->infiltration_capacity_soil[i], m->infiltration_capacity_compacted_soil[i],
                     m->unsaturated_store_capacity[i], m->f_infiltration_reduction[i], dt, o);
    m->infiltration[i] = o[0]; m->infiltration_excess[i] = o[1];
  }
  unsaturated_zone_flow(m, dt);
  soil_evaporation(m, dt);
  transpiration(m, dt);
  actual_infiltration(m, dt);
  PFOR for (int64_t i = 0; i < n; ++i)
    m->saturation_excess_water[i] =
        (m->soil_water_flux_surface[i] - m->actual_infiltration[i]) - m->infiltration_excess[i];
  PFOR for (int64_t i = 0; i < n; ++i) {
    double o[2];
    wfo_actual_infiltration_soil_path(
        m->soil_water_flux_surface[i], m->actual_infiltration[i],
        m->compacted_soil_area_fraction[i], m->infiltration_capacity_soil[i],
        m->infiltration_capacity_compacted_soil[i], m->f_infiltration_reduction[i], o);
    m->actual_infiltration_soil[i] = o[0]; m->actual_infiltration_compacted_soil[i] = o[1];
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->excess_water_soil[i] = jl_max(
        m->soil_water_flux_surface[i] * (1.0 - m->compacted_soil_area_fraction[i]) -
            m->actual_infiltration_soil[i], 0.0);
    m->excess_water_compacted_soil[i] = jl_max(
        m->soil_water_flux_surface[i] * m->compacted_soil_area_fraction[i] -
            m->actual_infiltration_compacted_soil[i], 0.0);
  }
  unsaturated_store_depth(m);
  PFOR for (int64_t i = 0; i < n; ++i)
    m->unsaturated_store_capacity[i] =
        m->soil_water_capacity[i] - m->saturated_water_depth[i] - m->unsaturated_store_depth[i];
  capillary_flux(m, dt);
  /* leakage! soil.jl:1118-1136 */
  PFOR for (int64_t i = 0; i < n; ++i) {
    double deepksat = kv_at_depth(m, m->soil_thickness[i], i, m->number_of_layers[i]);
    double deeptransfer = jl_min(m->drainable_water_depth[i] / dt, deepksat);
    m->actual_leakage[i] = jl_max(0.0, jl_min(m->maximum_leakage[i], deeptransfer));
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->recharge[i] = (m->transfer[i] - m->actual_capillary_flux[i] - m->actual_leakage[i] -
                      m->actual_evaporation_saturated_zone[i] -
                      m->soil_evaporation_saturated_zone[i]);
    m->actual_evapotranspiration[i] =
        m->soil_evaporation[i] + m->transpiration[i] + m->actual_open_water_evaporation_river[i] +
        m->actual_open_water_evaporation_land[i] + 0.0;
  }
}

/* sbm.jl:82-132 */
void wfo_update_land_hydrology_model(wfo_model* m, double dt) {
  update_interception_model(m, dt);
  update_snow_model(m, dt);
  lateral_snow_transport(m, dt);   /* snow_gravitational_transport__flag, sbm.jl:98-100 */
  update_glacier_model(m, dt);
  update_open_water_runoff(m, dt);
  update_bc_soil_model(m);
  update_soil_water_flow(m, dt);
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i)
    m->actual_evapotranspiration[i] += m->interception_rate[i];
}

/* soil/soil.jl:1294-1392 */
void wfo_update_soil_water_storage(wfo_model* m, double dt) {
  const int64_t N = m->cfg.N;
  (void)dt;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    const double* uld = m->unsaturated_layer_depth + i * N;
    const double* ult = m->unsaturated_layer_thickness + i * N;
    const double* alt = m->actual_layer_thickness + i * N;
    const double* cld = m->cumulative_layer_depth + i * (N + 1);
    const int64_t nu = m->n_unsatlayers[i];
    const double rd = m->rooting_depth[i];
    const double te = m->theta_s[i] - m->theta_r[i];
    double usd = 0.0;
    for (int64_t k = 0; k < nu; ++k) usd += uld[k];
    double exf = m->ssf_exfiltwater_average[i];
    double sbm_runoff = jl_max(0.0, exf + m->saturation_excess_water[i] + m->runoff_land[i] +
                                        m->infiltration_excess[i]);
    for (int64_t k = 0; k < m->number_of_layers[i]; ++k) {
      double vwc;
      if (k < nu) vwc = (uld[k] + (alt[k] - ult[k]) * te) / alt[k] + m->theta_r[i];
      else vwc = m->theta_s[i];
      m->volumetric_water_content[i * N + k] = vwc;
      m->relative_volumetric_water_content[i * N + k] = (vwc / m->theta_s[i]) / 1e-2;
    }
    double rootstore_unsat = 0.0;
    for (int64_t k = 0; k < nu; ++k)
      rootstore_unsat += jl_min(1.0, (jl_max(0.0, rd - cld[k]) / ult[k])) * uld[k];
    double rootstore_sat = jl_max(0.0, rd - m->water_table_depth[i]) * te;
    double rzs = rootstore_sat + rootstore_unsat;
    double vwc_rz = rzs / rd + m->theta_r[i];
    double satwaterdepth = (m->soil_thickness[i] - m->water_table_depth[i]) * te;
    double drainable = (m->soil_thickness[i] - m->water_table_depth[i]) *
                       jl_max(m->theta_s[i] - m->theta_fc[i], 0.02);
    m->unsaturated_store_capacity[i] = m->soil_water_capacity[i] - satwaterdepth - usd;
    m->unsaturated_store_depth[i] = usd;
    m->saturated_water_depth[i] = satwaterdepth;
    m->drainable_water_depth[i] = drainable;
    m->exfiltration_saturated_water[i] = exf;
    m->runoff[i] = sbm_runoff;
    m->root_zone_storage[i] = rzs;
    m->volumetric_water_content_root_zone[i] = vwc_rz;
    m->relative_volumetric_water_content_root_zone[i] = (vwc_rz / m->theta_s[i]) / 1e-2;
    m->total_soil_water_storage[i] = satwaterdepth + usd;
  }
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i)
    m->net_runoff[i] = m->runoff[i] - m->actual_open_water_evaporation_land[i];
}

/* sbm.jl:143-182 */
void wfo_update_total_water_storage(wfo_model* m) {
  const int64_t n = m->cfg.n;
  const int glac = m->cfg.snow && m->cfg.glacier;
  PFOR for (int64_t i = 0; i < n; ++i) m->total_storage[i] = 0.0;
  for (int64_t r = 0; r < m->cfg.nriv; ++r) {
    int64_t li = m->river_land_indices[r];
    m->total_storage[li] = (m->riv_h[r] * m->riv_flow_width[r] * m->riv_flow_length[r]) / m->area[li];
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    double snow = m->cfg.snow ? m->snow_storage[i] : 0.0;
    double snoww = m->cfg.snow ? m->snow_water[i] : 0.0;
    double gl = glac ? m->glacier_store[i] * m->glacier_fraction[i] : 0.0 * 0.0;
    m->total_storage[i] += (((snow + snoww) + gl) + m->canopy_storage[i]) + 0.0;
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    double sub_surface = m->unsaturated_store_depth[i] + m->saturated_water_depth[i];
    double lateral = m->olf_h[i] * (1.0 - m->river_fraction[i]);
    m->total_storage[i] += sub_surface + lateral;
  }
}

/* test hook: run one reference sweep by name (struct-level known-answer tests, test/soil.jl) */
int wfo_sweep(wfo_model* m, const char* name, double dt) {
  if (!strcmp(name, "update_bc_soil_model")) {
    /* soil_fraction is an input of update_bc_soil_model! (soil.jl:658): keep the caller's */
    double keep = m->soil_fraction[0];
    update_bc_soil_model(m);
    if (m->cfg.n == 1 && !isnan(keep)) {
      m->soil_fraction[0] = keep;
      m->potential_soilevaporation[0] = keep * m->potential_evaporation[0];
    }
    return 0;
  }
  if (!strcmp(name, "unsaturated_zone_flow")) { unsaturated_zone_flow(m, dt); return 0; }
  if (!strcmp(name, "soil_evaporation")) { soil_evaporation(m, dt); return 0; }
  if (!strcmp(name, "transpiration")) { transpiration(m, dt); return 0; }
  if (!strcmp(name, "capillary_flux")) { capillary_flux(m, dt); return 0; }
  if (!strcmp(name, "actual_infiltration")) { actual_infiltration(m, dt); return 0; }
  if (!strcmp(name, "update_interception_model")) { update_interception_model(m, dt); return 0; }
  if (!strcmp(name, "update_snow_model")) { update_snow_model(m, dt); return 0; }
  return -1;
}

void wfo_get_stats(wfo_model* m, int64_t out[9]) {
  out[0] = m->newton_iters_land; out[1] = m->newton_iters_river;
  out[2] = m->newton_calls_land; out[3] = m->newton_calls_river;
  out[4] = m->newton_maxit_land; out[5] = m->newton_maxit_river;
  out[6] = m->substeps_land; out[7] = m->substeps_river; out[8] = m->substeps_ssf;
}
