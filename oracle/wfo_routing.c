/*
 * wfo_routing.c -- CPU ORACLE (test infrastructure, see wfo.h): kinematic-wave routing
 * (lateral subsurface, overland, river), walked exactly like the reference: serial levels of
 * `order_of_subdomains`, threads over the sub-domains of one level, serial toposort walk
 * inside a sub-domain (Wflow/src/routing/surface/surface_kinwave.jl:293-341,492-566;
 * routing/subsurface/lateral_subsurface_flow.jl:198-273).
 */
#include "wfo.h"
#include "wfo_math.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int wfo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

#define PFOR _Pragma("omp parallel for schedule(static)")

static double kin_wave_min_flow_qroot(void) {
  /* routing/utils.jl:2 : KIN_WAVE_MIN_FLOW^0.2 (Julia Float64^Float64 = correctly rounded pow;
   * the exact value 1e-6 is representable to <1ulp, pow() returns it) */
  static double v = 0.0;
  if (v == 0.0) v = pow(WFO_KIN_WAVE_MIN_FLOW, 0.2);
  return v;
}

/* Julia cld(x::Float64, y::Float64) = round((x - mod(x, -y)) / y)  (Base div.jl) */
double wfo_cld(double x, double y) {
  double ny = -y;
  double r = fmod(x, ny), md;
  if (r == 0.0) md = copysign(r, ny);
  else if ((r > 0.0) != (ny > 0.0)) md = r + ny;
  else md = r;
  return rint((x - md) / y);
}

/* Julia round(v; sigdigits = 12) for v >= 0 (Base floatfuncs.jl: hidigit = 1+floor(log10|v|),
 * _round_invstep with 10.0^digits) */
double wfo_round_sigdigits12(double v) {
  if (v == 0.0 || !isfinite(v)) return v;
  static const double p10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,
                               1e8,  1e9,  1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                               1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  int h = 1 + (int)floor(log10(fabs(v)));
  int digits = 12 - h;
  if (digits >= 0) {
    double sm = digits <= 22 ? p10[digits] : pow(10.0, (double)digits);
    double y = rint(v * sm);
    double r = y / sm;
    return isfinite(r) ? r : v;
  } else {
    double s = -digits <= 22 ? p10[-digits] : pow(10.0, (double)(-digits));
    double y = rint(v / s);
    double r = y * s;
    return isfinite(r) ? r : v;
  }
}

/* routing/surface/surface_process.jl:24-70 */
void wfo_kinematic_wave(double q_in, double q_prev, double q_lat, double alpha, double dt,
                        double dx, double out[2], int64_t* iters) {
  int64_t it = 0;
  if (q_in + q_prev + q_lat == 0.0) { /* `≈ 0.0` with atol = 0 is exact equality */
    out[0] = 0.0; out[1] = 0.0;
    if (iters) *iters = 0;
    return;
  }
  const double qroot = kin_wave_min_flow_qroot();
  double dt_dx = dt / dx;
  double u_prev = q_prev >= 0.0 ? jl_pow(q_prev, 0.2) : 0.0;
  double constant_term = dt_dx * q_in + alpha * u_prev * u_prev * u_prev + dt * q_lat;
  double u = u_prev > 0.0 ? u_prev : cbrt(constant_term / alpha);
  const int max_iters = 3000;
  const double epsilon = 1.0e-12;
  double const_1 = 5.0 * dt_dx, const_2 = 3.0 * alpha;
  for (int k = 0; k < max_iters; ++k) {
    double u2 = u * u;
    double u3 = u2 * u;
    double f_u = u3 * (dt_dx * u2 + alpha) - constant_term;
    if (fabs(f_u) <= epsilon) break;
    double df_u = u2 * (const_1 * u2 + const_2);
    u -= f_u / df_u;
    if (isnan(u) || u <= 0.0) u = qroot;
    ++it;
  }
  u = jl_max(u, qroot);
  double u3 = u * u * u;
  out[1] = alpha * u3;
  out[0] = u3 * u * u;
  if (iters) *iters = it;
}

/* routing/subsurface/subsurface_process.jl:6-51 ; profile 0 exponential, 1 exponential_constant */
double wfo_ssf_celerity(double zi, double slope, double sy, double kh_0, double f, double z_exp,
                        int profile) {
  double z = zi;
  if (profile == 1) z = zi < z_exp ? zi : z_exp;
  return (kh_0 * exp(-f * z) * slope) / sy;
}

/* routing/subsurface/subsurface_process.jl:57-78 */
double wfo_kw_ssf_newton_raphson(double q, double constant_term, double celerity, double dt,
                                 double dx) {
  const double epsilon = 1.0e-12;
  const int max_iters = 3000;
  int count = 0;
  double dt_dx = dt / dx;
  double celerity_inv = 1.0 / celerity;
  double df = dt_dx + celerity_inv;
  for (;;) {
    double f = dt_dx * q + celerity_inv * q - constant_term;
    q -= (f / df);
    if (isnan(q)) q = 0.0;
    q = jl_max(q, WFO_KIN_WAVE_MIN_FLOW);
    if (fabs(f) <= epsilon || count >= max_iters) break;
    ++count;
  }
  return q;
}

/* utils.jl:1090-1131 */
void wfo_water_table_change(wfo_model* m, double net_flux, double sy, int64_t i, double dt,
                            double out[2]) {
  const int64_t N = m->cfg.N;
  const double* ult = m->unsaturated_layer_thickness + i * N;
  const double* uld = m->unsaturated_layer_depth + i * N;
  double theta_e = m->theta_s[i] - m->theta_r[i];
  double dh;
  if (net_flux <= 0.0) {
    dh = net_flux * dt / sy;
  } else {
    dh = 0.0;
    for (int64_t k = m->n_unsatlayers[i] - 1; k >= 0; --k) {
      double capacity = jl_max(ult[k] * theta_e - uld[k], 0.0) / dt;
      double flux_layer = jl_min(net_flux, capacity);
      if (capacity <= net_flux) dh += ult[k];
      else {
        double syd = theta_e - (uld[k] / ult[k]);
        dh += flux_layer * dt / syd;
      }
      net_flux -= flux_layer;
      if (net_flux == 0.0) break;
    }
  }
  out[0] = dh;
  out[1] = jl_max(net_flux, 0.0);
}

/* soil/soil.jl:1213-1259 */
static void update_ustorelayerdepth(wfo_model* m, double zi_prev, double zi, int64_t i) {
  const int64_t N = m->cfg.N;
  double* uld = m->unsaturated_layer_depth + i * N;
  double* ult = m->unsaturated_layer_thickness + i * N;
  int64_t nu_prev = m->n_unsatlayers[i];
  double ult_prev[16], ult_new[16];
  for (int64_t k = 0; k < N; ++k) ult_prev[k] = ult[k];
  /* set_layerthickness (utils.jl:390-404) */
  const double* cum = m->cumulative_layer_depth + i * (N + 1);
  const double* alt = m->actual_layer_thickness + i * N;
  int64_t nu = N;
  for (int64_t k = 0; k < N; ++k) {
    ult_new[k] = alt[k] * NAN;
    if (zi > cum[k + 1]) ult_new[k] = alt[k];
    else if (zi - cum[k] > 0.0) ult_new[k] = zi - cum[k];
    if (isnan(ult_new[k])) --nu;
  }
  if (zi < zi_prev) {
    for (int64_t k1 = nu; k1 <= nu_prev; ++k1) { /* 1-based layer index */
      if (k1 == 0) continue;
      if (isnan(ult_new[k1 - 1])) uld[k1 - 1] = 0.0;
      else {
        double scale = ult_new[k1 - 1] / ult_prev[k1 - 1];
        uld[k1 - 1] = scale * uld[k1 - 1];
      }
    }
  } else {
    for (int64_t k1 = nu_prev; k1 <= nu; ++k1) {
      if (k1 == 0) continue;
      double tp = isnan(ult_prev[k1 - 1]) ? 0.0 : ult_prev[k1 - 1];
      double delta = ult_new[k1 - 1] - tp;
      uld[k1 - 1] = uld[k1 - 1] + delta * (m->theta_fc[i] - m->theta_r[i]);
    }
  }
  m->n_unsatlayers[i] = nu;
  for (int64_t k = 0; k < N; ++k) ult[k] = ult_new[k];
  m->water_table_depth[i] = zi;
}

/* kinematic_wave_ssf(..., kh_profile::KhLayered, ...)      subsurface_process.jl:183-228
 * celerity from the equivalent conductivity kh of the step (ssf_celerity :47-51); no inner
 * sub-iterations; the excess above the soil column uses the effective specific yield sy_d */
static void kinematic_wave_ssf_layered(wfo_model* m, double q_in, double q_prev, double zi_prev,
                                       double q_net_bnds, double slope, double sy, double d,
                                       double dt, double dx, double dw, double q_max, int64_t i,
                                       double out[4]) {
  double q = (q_prev + q_in) / 2.0;
  const double celerity = (slope * m->ssf_kh[i]) / sy;
  const double constant_term = (dt / dx) * (q_in + q_net_bnds) + q_prev / celerity;
  q = wfo_kw_ssf_newton_raphson(q, constant_term, celerity, dt, dx);
  q = jl_min(q, (q_max * dw));
  const double net_flux = (q_in + q_net_bnds - q) / (dw * dx);
  double o[2];
  wfo_water_table_change(m, net_flux, sy, i, dt, o);
  const double dh = o[0], exfilt = o[1];
  double zi = zi_prev - dh;
  const double sy_d = dh > 0.0 ? (net_flux - exfilt) * dt / dh : sy;
  if (zi > d) {
    const double q_excess = (dw * dx) * sy_d * (zi - d) / dt;
    q = jl_max(q - q_excess, WFO_KIN_WAVE_MIN_FLOW);
  }
  zi = jl_clamp(zi, 0.0, d);
  update_ustorelayerdepth(m, zi_prev, zi, i);
  out[0] = q; out[1] = zi; out[2] = exfilt; out[3] = net_flux;
}

/* kh_layered_profile!(soil, subsurface_flow, kv_profile::KvLayered / KvLayeredExponential)
 * utils.jl:792-895: equivalent horizontal conductivity of the saturated part of the column */
void wfo_kh_layered_profile(wfo_model* m) {
  const int prof = m->cfg.kv_profile;
  if (prof < 2) return;
  const int64_t N = m->cfg.N;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    const double* kv = m->kv + i * N;
    const double* cld = m->cumulative_layer_depth + i * (N + 1);   /* _sumlayers[n] = cld[n] */
    const double* alt = m->actual_layer_thickness + i * N;
    const int64_t mm = m->number_of_layers[i];
    const double d = m->soil_thickness[i], zi = m->water_table_depth[i];
    const double ratio = m->ssf_khfrac[i];
    if (d > zi) {
      double transmissivity = 0.0;
      int64_t n = m->n_unsatlayers[i] > 1 ? m->n_unsatlayers[i] : 1;
      if (prof == 2) {
        transmissivity += (cld[n] - zi) * kv[n - 1];
        n += 1;
        while (n <= mm) { transmissivity += alt[n - 1] * kv[n - 1]; n += 1; }
      } else {
        const double f = m->hydraulic_conductivity_scale_parameter[i], zl = m->z_layered[i];
        const int64_t j = m->nlayers_kv[i];
        if (zi >= zl) {
          const double zt = d - zl;
          transmissivity += kv[j - 1] / f * (exp(-f * (zi - zl)) - exp(-f * zt));
          n = mm;
        } else {
          transmissivity += (cld[n] - zi) * kv[n - 1];
        }
        n += 1;
        while (n <= mm) {
          if (n > j) {
            const double zt = d - zl;
            transmissivity += kv[j - 1] / f * (1.0 - exp(-f * zt));
            n = mm;
          } else {
            transmissivity += alt[n - 1] * kv[n - 1];
          }
          n += 1;
        }
      }
      m->ssf_kh[i] = (transmissivity / (d - zi)) * ratio;
    } else {
      m->ssf_kh[i] = kv[mm - 1] * ratio;
    }
  }
}

/* routing/subsurface/subsurface_process.jl:89-172 (KhExponential / KhExponentialConstant) */
void wfo_kinematic_wave_ssf(wfo_model* m, double q_in, double q_prev, double zi_prev,
                            double q_net_bnds, double slope, double sy, double d, double dt,
                            double dx, double dw, double q_max, int64_t i, double out[4]) {
  if (q_in + q_prev == 0.0 && q_net_bnds <= 0.0) {
    out[0] = 0.0; out[1] = d; out[2] = 0.0; out[3] = 0.0;
    return;
  }
  const int prof = m->cfg.kv_profile;
  if (prof >= 2) {
    kinematic_wave_ssf_layered(m, q_in, q_prev, zi_prev, q_net_bnds, slope, sy, d, dt, dx, dw,
                               q_max, i, out);
    return;
  }
  const double kh_0 = m->kh_0[i], f = m->hydraulic_conductivity_scale_parameter[i];
  const double z_exp = prof == 1 ? m->z_exp[i] : 0.0;
  double q = (q_prev + q_in) / 2.0;
  double celerity = wfo_ssf_celerity(zi_prev, slope, sy, kh_0, f, z_exp, prof);
  double constant_term = (dt / dx) * (q_in + q_net_bnds) + q_prev / celerity;
  q = wfo_kw_ssf_newton_raphson(q, constant_term, celerity, dt, dx);
  q = jl_min(q, (q_max * dw));
  double net_flux = (q_in + q_net_bnds - q) / (dw * dx);
  double o[2];
  wfo_water_table_change(m, net_flux, sy, i, dt, o);
  double dh = o[0], exfilt = o[1];
  double zi = zi_prev - dh;
  if (zi > d) {
    double q_excess = (dw * dx) * sy * (zi - d) / dt;
    q = jl_max(q - q_excess, WFO_KIN_WAVE_MIN_FLOW);
  }
  zi = jl_clamp(zi, 0.0, d);
  const double max_delta_zi = 0.1;
  int64_t its = (int64_t)ceil(wfo_round_sigdigits12(fabs(zi - zi_prev) / max_delta_zi));
  if (its > 1) {
    double dt_s = dt / (double)its;
    double q_sum = 0.0, exfilt_sum = 0.0, net_flux_sum = 0.0;
    for (int64_t k = 0; k < its; ++k) {
      celerity = wfo_ssf_celerity(zi_prev, slope, sy, kh_0, f, z_exp, prof);
      constant_term = (dt_s / dx) * q_in + q_prev / celerity + q_net_bnds * (dt_s / dx);
      q = wfo_kw_ssf_newton_raphson(q_prev, constant_term, celerity, dt_s, dx);
      q = jl_min(q, (q_max * dw));
      net_flux = (q_in + q_net_bnds - q) / (dw * dx);
      wfo_water_table_change(m, net_flux, sy, i, dt_s, o);
      dh = o[0]; exfilt = o[1];
      zi = zi_prev - dh;
      if (zi > d) {
        double q_excess = (dw * dx) * sy * (zi - d) / dt_s;
        q = jl_max(q - q_excess, WFO_KIN_WAVE_MIN_FLOW);
      }
      zi = jl_clamp(zi, 0.0, d);
      update_ustorelayerdepth(m, zi_prev, zi, i);
      exfilt_sum += exfilt;
      net_flux_sum += net_flux;
      q_sum += q;
      q_prev = q;
      zi_prev = zi;
    }
    q = q_sum / (double)its;
    exfilt = exfilt_sum / (double)its;
    net_flux = net_flux_sum / (double)its;
  } else {
    update_ustorelayerdepth(m, zi_prev, zi, i);
  }
  out[0] = q; out[1] = zi; out[2] = exfilt; out[3] = net_flux;
}

/* Statistics.quantile! (type 7, alpha = beta = 1) on v[0..k) ; sorts v */
static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}
static double quantile7(double* v, int64_t n, double p) {
  qsort(v, (size_t)n, sizeof(double), cmp_double);
  double mm = 1.0 + p * (1.0 - 1.0 - 1.0);
  double aleph = (double)n * p + mm;
  int64_t j = (int64_t)trunc(aleph);
  if (j < 1) j = 1;
  if (j > n - 1) j = n - 1;
  double g = jl_clamp(aleph - (double)j, 0.0, 1.0);
  double a, b;
  if (n == 1) { a = v[0]; b = v[0]; }
  else { a = v[j - 1]; b = v[j]; }
  if (isfinite(a) && isfinite(b)) return a + g * (b - a);
  return (1.0 - g) * a + g * b;
}

/* routing/surface/surface_kinwave.jl:674-704 */
double wfo_stable_timestep_surface(const double* q, const double* alpha, const double* len,
                                   int64_t n, double p, double* work) {
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (q[i] > WFO_KIN_WAVE_MIN_FLOW) {
      double c = 1.0 / (alpha[i] * 0.6 * jl_pow(q[i], (0.6 - 1.0)));
      work[k++] = len[i] / c;
    }
  }
  if (k == 1) return work[0];
  if (k > 0) return quantile7(work, k, p);
  return 600.0;
}

/* routing/subsurface/lateral_subsurface_flow.jl:314-344 */
static double stable_timestep_ssf(wfo_model* m) {
  const int prof = m->cfg.kv_profile;
  int64_t k = 0;
  double dt_min = INFINITY;
  for (int64_t i = 0; i < m->cfg.n; ++i) {
    if (m->ssf_water_table_depth[i] > 0.0) {
      ++k;
      double c = prof >= 2 ? (m->slope[i] * m->ssf_kh[i]) / m->specific_yield[i]
                           : wfo_ssf_celerity(m->ssf_water_table_depth[i], m->slope[i],
                                              m->specific_yield[i], m->kh_0[i],
                                              m->hydraulic_conductivity_scale_parameter[i],
                                              prof == 1 ? m->z_exp[i] : 0.0, prof);
      dt_min = jl_min(dt_min, m->flow_length[i] / c);
    }
  }
  if (k == 0) dt_min = 0.5;
  return dt_min * m->cfg.ssf_alpha_coefficient;
}

/* routing/timestepping.jl:11-16 */
static double check_timestepsize(double dt_s, double t, double dt) {
  if (t + dt_s > dt) dt_s = dt - t;
  return dt_s;
}

/* ---------------------------------------------------------------------------------------- */
/* lateral subsurface flow                                                                  */
/* ---------------------------------------------------------------------------------------- */

/* lateral_subsurface_flow.jl:198-273 */
static void kinwave_subsurface_update(wfo_model* m, double dt) {
  const wfo_network* nw = &m->land;
  for (int64_t lv = 0; lv < nw->n_levels; ++lv) {
    const int64_t s0 = nw->level_ptr[lv], s1 = nw->level_ptr[lv + 1];
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t s = s0; s < s1; ++s) {
      const int64_t sub = nw->level_sub[s];
      for (int64_t e = nw->sub_ptr[sub]; e < nw->sub_ptr[sub + 1]; ++e) {
        const int64_t v = nw->sub_nodes[e], pos = nw->sub_pos[e];
        double qin = 0.0, tor = 0.0;
        for (int64_t u = nw->up_ptr[pos]; u < nw->up_ptr[pos + 1]; ++u) {
          const int64_t j = nw->up_idx[u];
          qin += m->ssf_q[j] * (1.0 - m->flow_fraction_to_river[j]);
          tor += m->ssf_q[j] * m->flow_fraction_to_river[j];
        }
        m->ssf_q_in[v] = qin;
        m->ssf_to_river_cumulative[v] += tor * dt;
        double o[4];
        wfo_kinematic_wave_ssf(m, m->ssf_q_in[v], m->ssf_q[v], m->ssf_water_table_depth[v],
                               m->ssf_q_net_bnds[v], m->slope[v], m->specific_yield[v],
                               m->ssf_soil_thickness[v], dt, m->flow_length[v], m->flow_width[v],
                               m->ssf_q_max[v], v, o);
        m->ssf_q[v] = o[0];
        m->ssf_water_table_depth[v] = o[1];
        m->ssf_q_in_cumulative[v] += m->ssf_q_in[v] * dt;
        m->ssf_q_cumulative[v] += m->ssf_q[v] * dt;
        m->ssf_exfiltwater_cumulative[v] += o[2] * dt;
        m->ssf_q_net_cumulative[v] += o[3] * m->area[v] * dt;
        m->ssf_head[v] = m->ssf_top[v] - m->ssf_water_table_depth[v];
        m->ssf_storage[v] = m->specific_yield[v] *
                            (m->ssf_soil_thickness[v] - m->ssf_water_table_depth[v]) * m->area[v];
      }
    }
  }
}

/* sbm_model.jl:74-81 */
void wfo_exchange_recharge(wfo_model* m) {
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    m->recharge_rate[i] = m->recharge[i];
    m->ssf_water_table_depth[i] = m->water_table_depth[i];
  }
  wfo_kh_layered_profile(m);   /* sbm_model.jl:84, layered conductivity profiles only */
}

/* lateral_subsurface_flow.jl:279-304 ; groundwater.jl:606-638 ; boundary_conditions.jl:12-21,219-236 */
void wfo_update_subsurface_flow_model(wfo_model* m, double dt) {
  const int64_t n = m->cfg.n;
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->ssf_to_river_cumulative[i] = 0.0;
    m->recharge_flux_cumulative[i] = 0.0;
    m->ssf_exfiltwater_cumulative[i] = 0.0;
    m->ssf_q_in_cumulative[i] = 0.0;
    m->ssf_q_cumulative[i] = 0.0;
    m->ssf_q_net_cumulative[i] = 0.0;
  }
  double t = 0.0;
  m->substeps_ssf = 0;
  while (t < dt) {
    double dt_s = m->cfg.adaptive ? stable_timestep_ssf(m) : m->cfg.dt_ssf;
    dt_s = check_timestepsize(dt_s, t, dt);
    PFOR for (int64_t i = 0; i < n; ++i) {
      m->ssf_q_net_bnds[i] = 0.0;
      double flux = m->recharge_rate[i] * m->area[i];
      if (m->ssf_water_table_depth[i] >= m->ssf_soil_thickness[i]) flux = jl_max(0.0, flux);
      m->recharge_flux[i] = flux;
      m->recharge_flux_cumulative[i] += flux * dt_s;
      m->ssf_q_net_bnds[i] += flux;
    }
    kinwave_subsurface_update(m, dt_s);
    t += dt_s;
    m->substeps_ssf++;
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->recharge_flux_average[i] = m->recharge_flux_cumulative[i] / dt;
    m->ssf_q_in_average[i] = m->ssf_q_in_cumulative[i] / dt;
    m->ssf_q_average[i] = m->ssf_q_cumulative[i] / dt;
    m->ssf_q_net_average[i] = m->ssf_q_net_cumulative[i] / dt;
    m->ssf_exfiltwater_average[i] = m->ssf_exfiltwater_cumulative[i] / dt;
    m->ssf_to_river_average[i] = m->ssf_to_river_cumulative[i] / dt;
  }
}

/* ---------------------------------------------------------------------------------------- */
/* overland flow                                                                            */
/* ---------------------------------------------------------------------------------------- */

/* surface_kinwave.jl:293-341 */
static void kinwave_land_update(wfo_model* m, double dt) {
  const wfo_network* nw = &m->land;
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) m->olf_qin[i] = 0.0;
  for (int64_t lv = 0; lv < nw->n_levels; ++lv) {
    const int64_t s0 = nw->level_ptr[lv], s1 = nw->level_ptr[lv + 1];
    int64_t it_sum = 0, calls = 0, maxit = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : it_sum, calls) reduction(max : maxit)
    for (int64_t s = s0; s < s1; ++s) {
      const int64_t sub = nw->level_sub[s];
      for (int64_t e = nw->sub_ptr[sub]; e < nw->sub_ptr[sub + 1]; ++e) {
        const int64_t v = nw->sub_nodes[e], pos = nw->sub_pos[e];
        double tor = 0.0, qin = 0.0;
        for (int64_t u = nw->up_ptr[pos]; u < nw->up_ptr[pos + 1]; ++u) {
          const int64_t j = nw->up_idx[u];
          tor += m->olf_q[j] * m->flow_fraction_to_river[j];
          qin += m->olf_q[j] * (1.0 - m->flow_fraction_to_river[j]);
        }
        m->olf_to_river_cumulative[v] += tor * dt;
        if (m->surface_flow_width[v] > 0.0) m->olf_qin[v] = qin;
        double o[2];
        int64_t it;
        wfo_kinematic_wave(m->olf_qin[v], m->olf_q[v], m->olf_qlat[v], m->olf_alpha[v], dt,
                           m->flow_length[v], o, &it);
        it_sum += it; calls += 1; if (it > maxit) maxit = it;
        if (m->newton_trace_land) m->newton_trace_land[v] += it;
        m->olf_q[v] = o[0];
        if (m->surface_flow_width[v] > 0.0) m->olf_h[v] = o[1] / m->surface_flow_width[v];
        m->olf_storage[v] = m->flow_length[v] * m->surface_flow_width[v] * m->olf_h[v];
        m->olf_q_cumulative[v] += m->olf_q[v] * dt;
        m->olf_qin_cumulative[v] += m->olf_qin[v] * dt;
      }
    }
    m->newton_iters_land += it_sum; m->newton_calls_land += calls;
    if (maxit > m->newton_maxit_land) m->newton_maxit_land = maxit;
  }
}

/* surface_kinwave.jl:740-766 (no drains, no water demand) */
void wfo_update_lateral_inflow_overland(wfo_model* m) {
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i)
    m->olf_inwater[i] = (m->net_runoff[i] + 0.0) * m->area[i] + 0.0;
}

/* surface_kinwave.jl:347-385 */
void wfo_update_overland_flow_model(wfo_model* m, double dt) {
  const int64_t n = m->cfg.n;
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->olf_qlat[i] = m->olf_inwater[i] / m->flow_length[i];
    m->olf_q_cumulative[i] = 0.0;
    m->olf_qin_cumulative[i] = 0.0;
    m->olf_to_river_cumulative[i] = 0.0;
  }
  double t = 0.0;
  m->substeps_land = 0;
  while (t < dt) {
    double dt_s = m->cfg.adaptive
                      ? wfo_stable_timestep_surface(m->olf_q, m->olf_alpha, m->flow_length, n,
                                                    0.02, m->scratch)
                      : m->cfg.dt_land;
    dt_s = check_timestepsize(dt_s, t, dt);
    kinwave_land_update(m, dt_s);
    t += dt_s;
    m->substeps_land++;
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->olf_q_average[i] = m->olf_q_cumulative[i] / dt;
    m->olf_to_river_average[i] = m->olf_to_river_cumulative[i] / dt;
    m->olf_qin_average[i] = m->olf_qin_cumulative[i] / dt;
  }
}

/* ---------------------------------------------------------------------------------------- */
/* river flow                                                                               */
/* ---------------------------------------------------------------------------------------- */

/* ---------------------------------------------------------------------------------------- */
/* reservoirs                                                    routing/surface/reservoir.jl */
/* ---------------------------------------------------------------------------------------- */

/* update_reservoir_simple                                                reservoir.jl:389-421 */
static void update_reservoir_simple(wfo_model* m, int64_t i, double precipitation,
                                    double evaporation, double inflow, double dt, double* outflow,
                                    double* storage_out) {
  double storage = m->res_storage[i] + (inflow + precipitation - evaporation) * dt;
  storage = jl_max(storage, 0.0);
  const double fill_fraction = storage / m->res_maximum_storage[i];
  const double fac = wfo_scurve(fill_fraction, m->res_target_minimum_fraction[i], 1.0, 30.0);
  const double demand_release = jl_min(fac * m->res_demand[i], storage / dt);
  storage -= demand_release * dt;
  const double release_wanted = jl_max(
      0.0, (storage - m->res_maximum_storage[i] * m->res_target_full_fraction[i]) / dt);
  const double overflow_q = jl_max(0.0, (storage - m->res_maximum_storage[i]) / dt);
  const double release_realized =
      jl_min(release_wanted, overflow_q + m->res_maximum_release[i] - demand_release);
  storage -= release_realized * dt;
  *outflow = release_realized + demand_release;
  *storage_out = storage;
}

/* update_reservoir_modified_puls                                         reservoir.jl:427-456 */
static void update_reservoir_modified_puls(wfo_model* m, int64_t i, double precipitation,
                                           double evaporation, double inflow, double dt,
                                           double* outflow_out, double* storage_out) {
  const double res_factor = m->res_area[i] / (dt * sqrt(m->res_rating_curve_coefficient[i]));
  const double si_factor = m->res_storage[i] / dt + precipitation - evaporation + inflow;
  const double si_factor_adj = si_factor - m->res_area[i] * m->res_threshold[i] / dt;
  double outflow;
  if (si_factor_adj > 0.0) {
    const double qs = -res_factor + sqrt((res_factor * res_factor + 4 * si_factor_adj));
    outflow = qs > 0.0 ? 0.25 * (qs * qs) : 0.0;
  } else {
    outflow = 0.0;
  }
  outflow = jl_min(outflow, si_factor);
  *outflow_out = outflow;
  *storage_out = (si_factor - outflow) * dt;
}

/* update_reservoir_free_weir without a linked lower reservoir (lower_reservoir_ind = 0:
 * diff_wl = 0)                                                           reservoir.jl:487-553 */
static void update_reservoir_free_weir(wfo_model* m, int64_t i, double precipitation,
                                       double evaporation, double inflow, double dt,
                                       double* outflow_out, double* storage_out) {
  const double storage_input =
      jl_max(m->res_storage[i] / dt + precipitation - evaporation + inflow, 0.0);
  double outflow;
  if (m->res_waterlevel[i] > m->res_threshold[i]) {
    const double dh = m->res_waterlevel[i] - m->res_threshold[i];
    outflow = m->res_rating_curve_coefficient[i] * jl_pow(dh, m->res_rating_curve_exponent[i]);
    const double maxflow = dh * m->res_area[i] / dt;
    outflow = jl_min(outflow, maxflow);
  } else {
    outflow = 0.0;
  }
  *outflow_out = outflow;
  *storage_out = (storage_input - outflow) * dt;
}

/* update_reservoir_outflow_obs                                           reservoir.jl:556-577 */
static void update_reservoir_outflow_obs(wfo_model* m, int64_t i, double precipitation,
                                         double evaporation, double inflow, double dt,
                                         double* outflow_out, double* storage_out) {
  const double storage_input =
      jl_max(m->res_storage[i] / dt + precipitation - evaporation + inflow, 0.0);
  double outflow = jl_min(m->res_outflow_obs[i], storage_input);
  double storage = (storage_input - outflow) * dt;
  if (!isnan(m->res_maximum_storage[i])) {
    const double overflow = jl_max(0.0, (storage - m->res_maximum_storage[i]) / dt);
    storage -= overflow * dt;
    outflow += overflow;
  }
  *outflow_out = outflow;
  *storage_out = storage;
}

/* update_reservoir_model!(reservoir_model, i, inflow, dt)                reservoir.jl:585-634
 * (linear storage curve; ReservoirOutflowType rating_curve (H-Q tables) is not restated) */
static void update_reservoir_model_i(wfo_model* m, int64_t i, double inflow, double dt) {
  const double precipitation = m->res_precipitation[i] * m->res_area[i];
  const double available_storage = m->res_storage[i] + (inflow + precipitation) * dt;
  const double potential_evaporation = m->res_evaporation[i] * m->res_area[i];
  const double evaporation = jl_min(available_storage / dt, potential_evaporation);
  double outflow = 0.0, storage = m->res_storage[i];
  const int type = (int)m->res_outflow_curve_type[i];
  if (!isnan(m->res_outflow_obs[i]))
    update_reservoir_outflow_obs(m, i, precipitation, evaporation, inflow, dt, &outflow, &storage);
  else if (type == 2)
    update_reservoir_free_weir(m, i, precipitation, evaporation, inflow, dt, &outflow, &storage);
  else if (type == 3)
    update_reservoir_modified_puls(m, i, precipitation, evaporation, inflow, dt, &outflow, &storage);
  else if (type == 4)
    update_reservoir_simple(m, i, precipitation, evaporation, inflow, dt, &outflow, &storage);
  const double waterlevel = m->res_waterlevel[i] + (storage - m->res_storage[i]) / m->res_area[i];
  m->res_storage[i] = storage;
  m->res_waterlevel[i] = waterlevel;
  m->res_outflow[i] = outflow;
  m->res_inflow_cumulative[i] += inflow * dt;
  m->res_outflow_cumulative[i] += outflow * dt;
  m->res_actevap_cumulative[i] += evaporation / m->res_area[i] * dt;
}

/* update_reservoir_model!(reservoir, river variables, network, v, dt)  surface_kinwave.jl:441-489 */
static void update_reservoir_at_node(wfo_model* m, int64_t v, double dt) {
  const int64_t i = m->riv_reservoir[v];
  if (i < 0) return;
  const double inflow_ext = m->res_external_inflow[i];
  double inflow;
  if (inflow_ext < 0.0) {
    const double abstraction = jl_min(-inflow_ext, (m->res_storage[i] / dt) * 0.98);
    m->res_actual_external_abstraction_cumulative[i] += abstraction * dt;
    inflow = -abstraction;
  } else {
    inflow = inflow_ext;
  }
  const double net_inflow =
      m->riv_q[v] + m->res_inflow_overland[i] + m->res_inflow_subsurface[i] + inflow;
  update_reservoir_model_i(m, i, net_inflow, dt);
  const int64_t j = m->river.down[v];   /* a reservoir without a downstream node is an error in
                                           the reference; the wrapper rejects it */
  if (j >= 0) m->riv_qin[j] = m->res_outflow[i];
}

/* test hooks of the reference's reservoir unit tests (test/reservoir.jl) */
void wfo_update_reservoir_model(wfo_model* m, int64_t i, double inflow, double dt) {
  update_reservoir_model_i(m, i, inflow, dt);
}
void wfo_update_reservoir_at_node(wfo_model* m, int64_t v, double dt) { update_reservoir_at_node(m, v, dt); }
/* local_inertial_flow(q0, zs0, zs1, hf, A, R, length, mannings_n_sq, froude_limit, dt) */
static double local_inertial_flow(double q0, double zs0, double zs1, double hf, double A, double R,
                                  double length, double mannings_n_sq, int froude_limit, double dt);
double wfo_local_inertial_flow(double q0, double zs0, double zs1, double hf, double A, double R,
                               double length, double mannings_n_sq, int froude_limit, double dt) {
  return local_inertial_flow(q0, zs0, zs1, hf, A, R, length, mannings_n_sq, froude_limit, dt);
}

/* update_inflow!(reservoir, river_flow, external_models, network)      surface_kinwave.jl:772-805 */
void wfo_update_inflow_reservoir(wfo_model* m) {
  for (int64_t i = 0; i < m->cfg.nres; ++i) {
    const int64_t li = m->river_land_indices[m->reservoir_river_indices[i]];
    m->res_inflow_overland[i] = m->olf_q_average[li];
    m->res_inflow_subsurface[i] = m->ssf_q_average[li];
    if (m->cfg.river_routing == 1) {  /* staggered schemes: to_river is included  :303-321 */
      m->res_inflow_overland[i] = m->olf_q_average[li] + m->olf_to_river_average[li];
      m->res_inflow_subsurface[i] = m->ssf_q_average[li] + m->ssf_to_river_average[li];
    }
  }
}

/* surface_kinwave.jl:492-566 (no floodplain) */
static void kinwave_river_update(wfo_model* m, double dt) {
  const wfo_network* nw = &m->river;
  PFOR for (int64_t i = 0; i < m->cfg.nriv; ++i) m->riv_qin[i] = 0.0;
  for (int64_t lv = 0; lv < nw->n_levels; ++lv) {
    const int64_t s0 = nw->level_ptr[lv], s1 = nw->level_ptr[lv + 1];
    int64_t it_sum = 0, calls = 0, maxit = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : it_sum, calls) reduction(max : maxit)
    for (int64_t s = s0; s < s1; ++s) {
      const int64_t sub = nw->level_sub[s];
      for (int64_t e = nw->sub_ptr[sub]; e < nw->sub_ptr[sub + 1]; ++e) {
        const int64_t v = nw->sub_nodes[e], pos = nw->sub_pos[e];
        double qs = 0.0;
        for (int64_t u = nw->up_ptr[pos]; u < nw->up_ptr[pos + 1]; ++u) qs += m->riv_q[nw->up_idx[u]];
        m->riv_qin[v] += qs;
        double inflow;
        if (m->riv_external_inflow[v] < 0.0) {
          double abstraction = jl_min(-m->riv_external_inflow[v], (m->riv_storage[v] / dt) * 0.80);
          m->riv_actual_external_abstraction_cumulative[v] += abstraction * dt;
          inflow = -abstraction / m->riv_flow_length[v];
        } else {
          inflow = m->riv_external_inflow[v] / m->riv_flow_length[v];
        }
        inflow -= m->riv_abstraction[v] / m->riv_flow_length[v];
        if (m->cfg.fp_levels > 0)                          /* surface_kinwave.jl:530-532 */
          inflow += m->riv_floodplain_water_exchange[v] / m->riv_flow_length[v];
        double o[2];
        int64_t it;
        wfo_kinematic_wave(m->riv_qin[v], m->riv_q[v], m->riv_qlat[v] + inflow, m->riv_alpha[v], dt,
                           m->riv_flow_length[v], o, &it);
        it_sum += it; calls += 1; if (it > maxit) maxit = it;
        if (m->newton_trace_river) m->newton_trace_river[v] += it;
        m->riv_q[v] = o[0];
        if (m->cfg.nres > 0) update_reservoir_at_node(m, v, dt);
        m->riv_h[v] = o[1] / m->riv_flow_width[v];
        m->riv_storage[v] = m->riv_flow_length[v] * o[1];
        m->riv_q_cumulative[v] += m->riv_q[v] * dt;
        m->riv_qin_cumulative[v] += m->riv_qin[v] * dt;
      }
    }
    m->newton_iters_river += it_sum; m->newton_calls_river += calls;
    if (maxit > m->newton_maxit_river) m->newton_maxit_river = maxit;
  }
}

/* test hook: one kinwave_river_update! with sub-step dt (routing_process.jl:290-383) */
void wfo_kinwave_river_update(wfo_model* m, double dt) { kinwave_river_update(m, dt); }

/* surface_kinwave.jl:710-734 */
void wfo_update_lateral_inflow_river(wfo_model* m) {
  PFOR for (int64_t r = 0; r < m->cfg.nriv; ++r) {
    const int64_t li = m->river_land_indices[r];
    m->riv_inwater[r] = ((m->ssf_to_river_average[li] + m->olf_to_river_average[li]) +
                         m->net_runoff_river[li] * m->area[li]) + 0.0 * m->area[li];
  }
}

static void li_update_river_flow_model(wfo_model* m, double dt);
void wfo_accucapacityflux(double* flux, double* material, const int64_t* order,
                          const int64_t* down, int64_t n, const double* capacity, double dt);

/* surface_kinwave.jl:613-662 */
void wfo_update_river_flow_model(wfo_model* m, double dt) {
  if (m->cfg.river_routing == 1) { li_update_river_flow_model(m, dt); return; }
  const int64_t n = m->cfg.nriv;
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->riv_qlat[i] = m->riv_inwater[i] / m->riv_flow_length[i];
    m->riv_q_cumulative[i] = 0.0;
    m->riv_actual_external_abstraction_cumulative[i] = 0.0;
    m->riv_qin_cumulative[i] = 0.0;
    if (m->cfg.fp_levels > 0) { m->fp_q_cumulative[i] = 0.0; m->fp_qin_cumulative[i] = 0.0; }
  }
  for (int64_t i = 0; i < m->cfg.nres; ++i) {  /* set_reservoir_vars!  surface_kinwave.jl:227-237 */
    m->res_inflow_cumulative[i] = 0.0;
    m->res_actual_external_abstraction_cumulative[i] = 0.0;
    m->res_outflow_cumulative[i] = 0.0;
    m->res_actevap_cumulative[i] = 0.0;
  }
  double t = 0.0;
  m->substeps_river = 0;
  while (t < dt) {
    double dt_s = m->cfg.adaptive
                      ? wfo_stable_timestep_surface(m->riv_q, m->riv_alpha, m->riv_flow_length, n,
                                                    0.05, m->scratch)
                      : m->cfg.dt_river;
    dt_s = check_timestepsize(dt_s, t, dt);
    if (m->cfg.fp_levels > 0) wfo_river_channel_floodplain_exchange(m, dt_s);
    kinwave_river_update(m, dt_s);
    if (m->cfg.fp_levels > 0) wfo_update_floodplain_model(m, dt_s);
    t += dt_s;
    m->substeps_river++;
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->riv_q_average[i] = m->riv_q_cumulative[i] / dt;
    m->riv_actual_external_abstraction_average[i] =
        m->riv_actual_external_abstraction_cumulative[i] / dt;
    m->riv_qin_average[i] = m->riv_qin_cumulative[i] / dt;
  }
  if (m->cfg.fp_levels > 0) {                              /* surface_kinwave.jl:650-659 */
    PFOR for (int64_t i = 0; i < n; ++i) {
      m->fp_q_average[i] = m->fp_q_cumulative[i] / dt;
      m->riv_q_channel_average[i] = m->riv_q_average[i];
      m->riv_q_average[i] = m->riv_q_channel_average[i] + m->fp_q_average[i];
      m->fp_qin_average[i] = m->fp_qin_cumulative[i] / dt;
      m->riv_qin_average[i] = m->riv_qin_average[i] + m->fp_qin_average[i];
    }
  }
  for (int64_t i = 0; i < m->cfg.nres; ++i) {  /* average_reservoir_vars!  surface_kinwave.jl:244-258 */
    m->res_outflow_average[i] = m->res_outflow_cumulative[i] / dt;
    m->res_inflow_average[i] = m->res_inflow_cumulative[i] / dt;
    m->res_actual_external_abstraction_average[i] =
        m->res_actual_external_abstraction_cumulative[i] / dt;
  }
}

/* ---------------------------------------------------------------------------------------- */
/* local-inertial river flow (no floodplain)        routing/surface/surface_staggered_scheme.jl */
/* ---------------------------------------------------------------------------------------- */
#define WFO_G 9.80665 /* GRAVITATIONAL_ACCELERATION  Wflow.jl:77 */

/* local_inertial_flow                                           surface_process.jl:88-115 */
static double local_inertial_flow(double q0, double zs0, double zs1, double hf, double A, double R,
                                  double length, double mannings_n_sq, int froude_limit, double dt) {
  const double slope = (zs1 - zs0) / length;
  const double pow_R = cbrt(R * R * R * R);
  double q = ((q0 - WFO_G * A * dt * slope) /
              (1.0 + WFO_G * dt * mannings_n_sq * fabs(q0) / (pow_R * A)));
  const double fr = ((q / A) / sqrt(WFO_G * hf)) * (double)froude_limit;
  if ((fabs(fr) > 1.0) && (q > 0.0)) q = sqrt(WFO_G * hf) * A;
  if ((fabs(fr) > 1.0) && (q < 0.0)) q = -sqrt(WFO_G * hf) * A;
  return q;
}

/* node the edge leaving node i ends in: a river node, -1 (no edge), or -2 (ghost node of a pit) */
static int64_t li_dst(const wfo_model* m, int64_t i) {
  const int64_t d = m->river.down[i];
  if (d >= 0) return d;
  return m->cfg.li_ghost_nodes ? -2 : -1;
}

/* stable_timestep(::RiverFlowModel{<:LocalInertial})   surface_staggered_scheme.jl:1004-1020 */
double wfo_li_stable_timestep(wfo_model* m) {
  double dt_min = INFINITY;
  for (int64_t i = 0; i < m->cfg.nriv; ++i) {
    const double dt = m->cfg.li_alpha * m->riv_flow_length[i] / sqrt(WFO_G * m->riv_h[i]);
    dt_min = dt < dt_min ? dt : dt_min;
  }
  return isinf(dt_min) ? 60.0 : dt_min;
}

/* update_river_channel_flow!(::RiverFlowModel{<:LocalInertial})            :326-383 */
void wfo_li_update_river_channel_flow(wfo_model* m, double dt) {
  PFOR for (int64_t i = 0; i < m->cfg.nriv; ++i) {
    const int64_t d = li_dst(m, i);
    if (d == -1) continue;                                   /* no edge leaves this node */
    if (m->cfg.nres > 0 && m->riv_reservoir[i] >= 0) continue; /* not in active_e          */
    const double q_previous = m->riv_q[i];
    const double zs_src = m->li_zb[i] + m->riv_h[i];
    const double h_dst = d == -2 ? m->li_ghost_h[i] : m->riv_h[d];
    const double zs_dst = (d == -2 ? m->li_zb[i] : m->li_zb[d]) + h_dst;
    const double zs_at_edge = jl_max(zs_src, zs_dst);
    const double hf = zs_at_edge - m->li_zb_at_edge[i];
    m->li_zs_at_edge[i] = zs_at_edge;
    m->li_water_depth_at_edge[i] = hf;
    const double w = m->li_flow_width_at_edge[i];
    const double A = w * hf;
    const double R = A / (2.0 * hf + w);
    double q = hf > m->cfg.li_h_thresh
                   ? local_inertial_flow(q_previous, zs_src, zs_dst, hf, A, R,
                                         m->li_flow_length_at_edge[i],
                                         m->li_mannings_n_sq_at_edge[i], m->cfg.li_froude_limit, dt)
                   : 0.0;
    if (m->riv_h[i] <= 0.0) q = jl_min(q, 0.0);
    if (h_dst <= 0.0) q = jl_max(q, 0.0);
    m->riv_q[i] = q;
    m->riv_q_cumulative[i] += q * dt;
  }
}

/* ---- 1-D floodplain of the local-inertial river --------------------------------------------
 * profile tables are [node][level] (Julia: profile.x[level, node])                           */
#define FP(m, tab, i, l) ((m)->tab[(i) * (m)->cfg.fp_levels + (l)])
/* interpolation_indices                                                floodplain.jl:287-300 */
static void fp_interpolation_indices(const double* v, int64_t n, double x, int64_t* i1, int64_t* i2) {
  int64_t a = 0;
  for (int64_t i = 0; i < n; ++i)
    if (v[i] <= x) a = i;
  *i1 = a;
  *i2 = a == n - 1 ? a : a + 1;
}
/* compute_floodplain_flow_area (compute_flood_flow_area minus the channel)   :306-318,343-354 */
static double fp_floodplain_flow_area(const wfo_model* m, double h, int64_t idx, int64_t i1, int64_t i2) {
  const double channel_area = FP(m, fp_profile_width, idx, 0) * h;
  const double delta_h = h - m->cfg.fp_depth[i1];
  const double flow_area = FP(m, fp_profile_flow_area, idx, i1) + (FP(m, fp_profile_width, idx, i2) * delta_h);
  return jl_max(flow_area - channel_area, 0.0);
}
/* compute_wetted_perimeter                                                          :324-328 */
static double fp_wetted_perimeter(const wfo_model* m, double h, int64_t idx, int64_t i1) {
  const double delta_h = h - m->cfg.fp_depth[i1];
  return FP(m, fp_profile_wetted_perimeter, idx, i1) + 2.0 * delta_h;
}
/* compute_flood_depth                                                               :331-341 */
static double fp_flood_depth(const wfo_model* m, double flood_storage, double flow_length, int64_t i) {
  int64_t i1, i2;
  fp_interpolation_indices(&FP(m, fp_profile_storage, i, 0), m->cfg.fp_levels, flood_storage, &i1, &i2);
  const double delta_A = (flood_storage - FP(m, fp_profile_storage, i, i1)) / flow_length;
  const double delta_h = delta_A / FP(m, fp_profile_width, i, i2);
  return m->cfg.fp_depth[i1] + delta_h;
}

/* update_floodplain_flow!(::LocalInertial)              surface_staggered_scheme.jl:440-533.
 * floodplain q_previous .= q happens at the top of update_river_channel_flow! (:336-339), so
 * fp_q still holds the previous sub-step's value here; zs_src / zs_dst / zs_at_edge are the
 * values that sweep stored for the edge (the depths have not changed since). A ghost node
 * copies the profile of its pit (floodplain.jl:176 index_pit) and holds no floodplain water. */
void wfo_li_update_floodplain_flow(wfo_model* m, double dt) {
  PFOR for (int64_t i = 0; i < m->cfg.nriv; ++i) {
    const int64_t d = li_dst(m, i);
    if (d == -1) continue;
    if (m->cfg.nres > 0 && m->riv_reservoir[i] >= 0) continue;
    const double q_previous = m->fp_q[i];
    const double zs_src = m->li_zb[i] + m->riv_h[i];
    const double h_dst = d == -2 ? m->li_ghost_h[i] : m->riv_h[d];
    const double zs_dst = (d == -2 ? m->li_zb[i] : m->li_zb[d]) + h_dst;
    const int64_t i_dst = d == -2 ? i : d;   /* profile of the ghost node = the pit's */
    const double hf = jl_max(m->li_zs_at_edge[i] - m->fp_zb_at_edge[i], 0.0);
    m->fp_water_depth_at_edge[i] = hf;
    int64_t i1, i2;
    fp_interpolation_indices(m->cfg.fp_depth, m->cfg.fp_levels, hf, &i1, &i2);
    const double a_src = fp_floodplain_flow_area(m, hf, i, i1, i2);
    const double a_dst = fp_floodplain_flow_area(m, hf, i_dst, i1, i2);
    const double A = jl_min(a_src, a_dst);
    const double R = a_src < a_dst ? a_src / fp_wetted_perimeter(m, hf, i, i1)
                                   : a_dst / fp_wetted_perimeter(m, hf, i_dst, i1);
    double q = A > 1.0e-05
                   ? local_inertial_flow(q_previous, zs_src, zs_dst, hf, A, R,
                                         m->li_flow_length_at_edge[i],
                                         m->fp_mannings_n_sq_at_edge[i], m->cfg.li_froude_limit, dt)
                   : 0.0;
    if (m->fp_h[i] <= 0.0) q = jl_min(q, 0.0);
    if ((d == -2 ? 0.0 : m->fp_h[d]) <= 0.0) q = jl_max(q, 0.0);
    if (q * m->riv_q[i] < 0.0) q = 0.0;       /* opposite to the channel flow */
    m->fp_q[i] = q;
    m->fp_q_cumulative[i] += q * dt;
  }
}

/* update_water_depth_and_storage!(floodplain_model, river_flow_model, ...)         :674-712 */
void wfo_li_update_floodplain_water_depth_and_storage(wfo_model* m, double dt) {
  const wfo_network* nw = &m->river;
  PFOR for (int64_t i = 0; i < m->cfg.nriv; ++i) {
    if (m->cfg.nres > 0 && m->riv_reservoir[i] >= 0) continue;
    double q_src = 0.0;
    for (int64_t u = nw->in_ptr[i]; u < nw->in_ptr[i + 1]; ++u) q_src += m->fp_q[nw->in_idx[u]];
    const double q_dst = li_dst(m, i) == -1 ? 0.0 : 0.0 + m->fp_q[i];
    m->fp_storage[i] += (q_src - q_dst) * dt;
    if (m->fp_storage[i] < 0.0) {
      m->fp_error[i] += fabs(m->fp_storage[i]);
      m->fp_storage[i] = 0.0;
    }
    const double storage_total = m->riv_storage[i] + m->fp_storage[i];
    if (storage_total > m->li_bankfull_storage[i]) {
      const double flood_storage = storage_total - m->li_bankfull_storage[i];
      const double h = fp_flood_depth(m, flood_storage, m->riv_flow_length[i], i);
      m->riv_h[i] = m->li_bankfull_depth[i] + h;
      m->riv_storage[i] = m->riv_h[i] * m->riv_flow_width[i] * m->riv_flow_length[i];
      m->fp_storage[i] = jl_max(storage_total - m->riv_storage[i], 0.0);
      m->fp_h[i] = m->fp_storage[i] > 0.0 ? h : 0.0;
    } else {
      m->riv_h[i] = storage_total / (m->riv_flow_length[i] * m->riv_flow_width[i]);
      m->riv_storage[i] = storage_total;
      m->fp_h[i] = 0.0;
      m->fp_storage[i] = 0.0;
    }
  }
}

/* ---- 1-D floodplain of the kinematic-wave river ------------------------------------------- */
/* manning_flow                                                    surface_process.jl:166-169 */
static double manning_flow(double mannings_n, double hydraulic_radius, double slope, double area) {
  return cbrt(hydraulic_radius * hydraulic_radius) * sqrt(slope) * area / mannings_n;
}

/* river_channel_floodplain_exchange!                               surface_kinwave.jl:567-601 */
void wfo_river_channel_floodplain_exchange(wfo_model* m, double dt) {
  PFOR for (int64_t i = 0; i < m->cfg.nriv; ++i) {
    const double storage_total = m->riv_storage[i] + m->fp_storage[i];
    double delta_river_storage;
    if (storage_total > m->li_bankfull_storage[i]) {
      const double flood_storage = storage_total - m->li_bankfull_storage[i];
      const double h = fp_flood_depth(m, flood_storage, m->riv_flow_length[i], i);
      const double river_storage = (m->li_bankfull_depth[i] + h) * m->riv_flow_width[i] * m->riv_flow_length[i];
      delta_river_storage = river_storage - m->riv_storage[i];
      m->fp_storage[i] = jl_max(storage_total - river_storage, 0.0);
      m->fp_h[i] = m->fp_storage[i] > 0.0 ? h : 0.0;
    } else {
      delta_river_storage = jl_max(storage_total - m->riv_storage[i], 0.0);
      m->fp_h[i] = 0.0;
      m->fp_storage[i] = 0.0;
    }
    m->riv_floodplain_water_exchange[i] = delta_river_storage / dt;
  }
}

/* update_floodplain_model!(::KinematicWave, ::FloodPlainModel{<:Manning})          :387-432 */
void wfo_update_floodplain_model(wfo_model* m, double dt) {
  const wfo_network* nw = &m->river;
  const int64_t n = m->cfg.nriv;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t v = nw->order[k];
    const int64_t ds = nw->down[v];
    double cap = 0.0;
    if (m->fp_h[v] > 0.0) {
      int64_t i1, i2;
      fp_interpolation_indices(m->cfg.fp_depth, m->cfg.fp_levels, m->fp_h[v], &i1, &i2);
      const double flow_area = fp_floodplain_flow_area(m, m->fp_h[v], v, i1, i2);
      const double flow_area_ds = ds >= 0 ? fp_floodplain_flow_area(m, m->fp_h[v], ds, i1, i2) : flow_area;
      if (flow_area > 1.0e-05 && flow_area_ds > 1.0e-05) {
        const double wetted_perimeter = fp_wetted_perimeter(m, m->fp_h[v], v, i1);
        const double hydraulic_radius = flow_area / wetted_perimeter;
        cap = manning_flow(m->fp_mannings_n[v], hydraulic_radius, m->fp_slope[v], flow_area);
      }
    }
    m->fp_flow_capacity[v] = cap;
  }
  /* q .= accucapacityflux(storage, network, flow_capacity, dt): `accucapacityflux` only
   * allocates the flux and calls accucapacityflux! on `storage` itself
   * (routing/utils.jl:131-135), so the floodplain storage moves downstream here */
  wfo_accucapacityflux(m->fp_q, m->fp_storage, nw->order, nw->down, n, m->fp_flow_capacity, dt);
  for (int64_t i = 0; i < n; ++i) m->fp_q_cumulative[i] += m->fp_q[i] * dt;
  for (int64_t k = 0; k < n; ++k) {   /* flux_in!  routing/utils.jl:161-167 */
    const int64_t v = nw->order[k];
    double ssum = 0.0;
    for (int64_t u = nw->up_ptr[k]; u < nw->up_ptr[k + 1]; ++u) ssum += m->fp_q[nw->up_idx[u]];
    m->fp_qin[v] = ssum;
  }
  for (int64_t i = 0; i < n; ++i) m->fp_qin_cumulative[i] += m->fp_qin[i] * dt;
}

/* update_bc_reservoir_model!                                                :627-661 */
void wfo_li_update_bc_reservoir_model(wfo_model* m, double dt) {
  const wfo_network* nw = &m->river;
  for (int64_t v = 0; v < m->cfg.nres; ++v) {
    const int64_t i = m->reservoir_river_indices[v];
    double q_in = 0.0;  /* sum_at(q, edges_at_node.src[i]): edges entering the reservoir node */
    for (int64_t u = nw->in_ptr[i]; u < nw->in_ptr[i + 1]; ++u) q_in += m->riv_q[nw->in_idx[u]];
    if (m->cfg.fp_levels > 0) {              /* get_inflow_reservoir :291-299 */
      double q_fp = 0.0;
      for (int64_t u = nw->in_ptr[i]; u < nw->in_ptr[i + 1]; ++u) q_fp += m->fp_q[nw->in_idx[u]];
      q_in += q_fp;
    }
    double inflow;
    if (m->res_external_inflow[v] < 0.0) {
      const double abstraction = jl_min(-m->res_external_inflow[v], (m->res_storage[v] / dt) * 0.98);
      m->res_actual_external_abstraction_cumulative[v] += abstraction * dt;
      inflow = -abstraction;
    } else {
      inflow = m->res_external_inflow[v];
    }
    const double net_inflow = q_in + m->res_inflow_overland[v] + m->res_inflow_subsurface[v] + inflow;
    update_reservoir_model_i(m, v, net_inflow, dt);
    m->riv_q[i] = m->res_outflow[v];
    m->riv_q_cumulative[i] += m->riv_q[i] * dt;
  }
}

/* update_water_depth_and_storage!(river_flow_model, domain, dt)            :723-759 */
void wfo_li_update_water_depth_and_storage(wfo_model* m, double dt) {
  const wfo_network* nw = &m->river;
  PFOR for (int64_t i = 0; i < m->cfg.nriv; ++i) {
    if (m->cfg.nres > 0 && m->riv_reservoir[i] >= 0) continue;  /* not in active_n */
    double q_src = 0.0;
    for (int64_t u = nw->in_ptr[i]; u < nw->in_ptr[i + 1]; ++u) q_src += m->riv_q[nw->in_idx[u]];
    const double q_dst = li_dst(m, i) == -1 ? 0.0 : 0.0 + m->riv_q[i];
    m->riv_storage[i] += (q_src - q_dst + m->riv_inwater[i] - m->riv_abstraction[i]) * dt;
    if (m->riv_storage[i] < 0.0) {
      m->li_error[i] = m->li_error[i] + fabs(m->riv_storage[i]);
      m->riv_storage[i] = 0.0;
    }
    double inflow;
    if (m->riv_external_inflow[i] < 0.0) {
      const double abstraction = jl_min(-m->riv_external_inflow[i], m->riv_storage[i] / dt * 0.80);
      m->riv_actual_external_abstraction_cumulative[i] += abstraction * dt;
      inflow = -abstraction;
    } else {
      inflow = m->riv_external_inflow[i];
    }
    m->riv_storage[i] += inflow * dt;
    m->riv_h[i] = m->riv_storage[i] / (m->riv_flow_length[i] * m->riv_flow_width[i]);
  }
}

/* update_river_flow_model!(::RiverFlowModel{<:AbstractStaggeredRoutingMethod})   :800-838 */
static void li_update_river_flow_model(wfo_model* m, double dt) {
  const int64_t n = m->cfg.nriv;
  for (int64_t i = 0; i < m->cfg.nres; ++i) {
    m->res_inflow_cumulative[i] = 0.0;
    m->res_actual_external_abstraction_cumulative[i] = 0.0;
    m->res_outflow_cumulative[i] = 0.0;
    m->res_actevap_cumulative[i] = 0.0;
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->riv_q_cumulative[i] = 0.0;
    m->riv_actual_external_abstraction_cumulative[i] = 0.0;
    if (m->cfg.fp_levels > 0) m->fp_q_cumulative[i] = 0.0;
  }
  double t = 0.0;
  m->substeps_river = 0;
  while (t < dt) {
    double dt_s = wfo_li_stable_timestep(m);
    dt_s = check_timestepsize(dt_s, t, dt);
    wfo_li_update_river_channel_flow(m, dt_s);
    if (m->cfg.fp_levels > 0) wfo_li_update_floodplain_flow(m, dt_s);
    wfo_li_update_bc_reservoir_model(m, dt_s);
    wfo_li_update_water_depth_and_storage(m, dt_s);
    if (m->cfg.fp_levels > 0) wfo_li_update_floodplain_water_depth_and_storage(m, dt_s);
    t += dt_s;
    m->substeps_river++;
  }
  PFOR for (int64_t i = 0; i < n; ++i) {
    m->riv_q_average[i] = m->riv_q_cumulative[i] / dt;
    m->riv_actual_external_abstraction_average[i] =
        m->riv_actual_external_abstraction_cumulative[i] / dt;
  }
  for (int64_t i = 0; i < m->cfg.nres; ++i) {
    m->res_outflow_average[i] = m->res_outflow_cumulative[i] / dt;
    m->res_inflow_average[i] = m->res_inflow_cumulative[i] / dt;
    m->res_actual_external_abstraction_average[i] =
        m->res_actual_external_abstraction_cumulative[i] / dt;
  }
  if (m->cfg.fp_levels > 0) {  /* :826-835 */
    PFOR for (int64_t i = 0; i < n; ++i) {
      m->fp_q_average[i] = m->fp_q_cumulative[i] / dt;
      m->riv_q_channel_average[i] = m->riv_q_average[i];
      m->riv_q_average[i] = m->riv_q_channel_average[i] + m->fp_q_average[i];
    }
  }
}

/* ---------------------------------------------------------------------------------------- */
/* 2-D local-inertial overland flow coupled to the local-inertial river                        */
/*                                          routing/surface/surface_staggered_scheme.jl:840-1546 */
/* ---------------------------------------------------------------------------------------- */
/* The reference's flow vectors hold n + 1 entries, the last one (the edge to "outside") stays 0 */
static inline double lil_at(const double* q, int64_t idx, int64_t n) { return idx < n ? q[idx] : 0.0; }

/* local_inertial_flow(theta, q0, qd, qu, zs0, zs1, hf, width, length, mannings_n_sq,
 * froude_limit, dt): rectangular area, de Almeida et al. (2012)      surface_process.jl:123-159 */
double wfo_local_inertial_flow_rect(double theta, double q0, double qd, double qu, double zs0,
                                    double zs1, double hf, double width, double length,
                                    double mannings_n_sq, int froude_limit, double dt) {
  const double slope = (zs1 - zs0) / length;
  const double pow_hf = cbrt(hf * hf * hf * hf * hf * hf * hf);
  double q = (((theta * q0 + 0.5 * (1.0 - theta) * (qu + qd)) - WFO_G * hf * width * dt * slope) /
              (1.0 + WFO_G * dt * mannings_n_sq * fabs(q0) / (pow_hf * width)));
  if (froude_limit) {
    const double fr = (q / width / hf) / sqrt(WFO_G * hf);
    if (fabs(fr) > 1.0 && q > 0.0) q = hf * sqrt(WFO_G * hf) * width;
    else if (fabs(fr) > 1.0 && q < 0.0) q = -hf * sqrt(WFO_G * hf) * width;
  }
  return q;
}

/* stable_timestep(::OverlandFlowModel{<:LocalInertial}, ::LandParameters)          :1022-1043 */
double wfo_lil_stable_timestep(wfo_model* m) {
  double dt_min = INFINITY;
  for (int64_t i = 0; i < m->cfg.n; ++i) {
    if (m->land_river_indices[i] >= 0) continue;                 /* river_location[i] != 0 */
    const double dt = m->cfg.li_land_alpha * jl_min(m->li_land_x_length[i], m->li_land_y_length[i]) /
                      sqrt(WFO_G * m->olf_h[i]);
    dt_min = dt < dt_min ? dt : dt_min;
  }
  return isinf(dt_min) ? 60.0 : dt_min;
}

/* update_directional_flow!(model, domain, i, dt, is_x_direction)                   :1201-1271 */
void wfo_lil_update_directional_flow(wfo_model* m, int64_t i, double dt, int is_x) {
  const int64_t n = m->cfg.n;
  const int64_t up = is_x ? m->edge_x_up[i] : m->edge_y_up[i];
  const int64_t down = is_x ? m->edge_x_down[i] : m->edge_y_down[i];
  const double width_at_edge = is_x ? m->li_land_ywidth_at_edge[i] : m->li_land_xwidth_at_edge[i];
  const double z_max_at_edge = is_x ? m->li_land_zx_max_at_edge[i] : m->li_land_zy_max_at_edge[i];
  const double* length_vec = is_x ? m->li_land_x_length : m->li_land_y_length;
  double* q_current = is_x ? m->li_land_qx : m->li_land_qy;
  const double* q_prev = is_x ? m->li_land_qx0 : m->li_land_qy0;
  double* q_cumulative = is_x ? m->li_land_qx_cumulative : m->li_land_qy_cumulative;
  if (up < n && width_at_edge != 0.0) {
    const double zs_current = m->li_land_z[i] + m->olf_h[i];
    const double zs_upstream = m->li_land_z[up] + m->olf_h[up];
    const double zs_max_at_edge = jl_max(zs_current, zs_upstream);
    const double water_depth_at_edge = (zs_max_at_edge - z_max_at_edge);
    if (water_depth_at_edge > m->cfg.li_land_h_thresh) {
      const double length_at_edge = 0.5 * (length_vec[i] + length_vec[up]);
      double q = wfo_local_inertial_flow_rect(m->cfg.li_land_theta, q_prev[i], lil_at(q_prev, down, n),
                                              lil_at(q_prev, up, n), zs_current, zs_upstream,
                                              water_depth_at_edge, width_at_edge, length_at_edge,
                                              m->li_land_mannings_n_sq_at_edge[i],
                                              m->cfg.li_land_froude_limit, dt);
      if (m->olf_h[i] <= 0.0) q = jl_min(q, 0.0);
      if (m->olf_h[up] <= 0.0) q = jl_max(q, 0.0);
      q_current[i] = q;
    } else {
      q_current[i] = 0.0;
    }
    q_cumulative[i] += q_current[i] * dt;
  }
}

/* local_inertial_update_fluxes!                                                    :1276-1295 */
void wfo_lil_update_fluxes(wfo_model* m, double dt) {
  const int64_t n = m->cfg.n;
  PFOR for (int64_t i = 0; i < n; ++i) { m->li_land_qx0[i] = m->li_land_qx[i]; m->li_land_qy0[i] = m->li_land_qy[i]; }
  PFOR for (int64_t i = 0; i < n; ++i) {
    wfo_lil_update_directional_flow(m, i, dt, 1);
    wfo_lil_update_directional_flow(m, i, dt, 0);
  }
}

/* qx[ind_x_down] - qx[i] + qy[ind_y_down] - qy[i] as the reference writes it (:1343-1345,1433-1434) */
static inline double lil_net_land_flow(const wfo_model* m, int64_t i) {
  const int64_t n = m->cfg.n;
  return lil_at(m->li_land_qx, m->edge_x_down[i], n) - m->li_land_qx[i] +
         lil_at(m->li_land_qy, m->edge_y_down[i], n) - m->li_land_qy[i];
}

/* update_inflow_reservoir!                                                         :1301-1319 */
void wfo_lil_update_inflow_reservoir(wfo_model* m) {
  for (int64_t i = 0; i < m->cfg.nres; ++i) {
    const int64_t j = m->river_land_indices[m->reservoir_river_indices[i]];
    m->res_inflow_overland[i] = m->li_land_runoff[j] + (lil_net_land_flow(m, j));
  }
}

/* compute_river_storage_change                                                     :1325-1352 */
double wfo_lil_compute_river_storage_change(wfo_model* m, int64_t i, double dt) {
  const wfo_network* nw = &m->river;
  const int64_t r = m->land_river_indices[i];
  double q_src = 0.0;     /* sum_at(q, edges_at_node.src[r]): the edges entering the node */
  for (int64_t u = nw->in_ptr[r]; u < nw->in_ptr[r + 1]; ++u) q_src += m->riv_q[nw->in_idx[u]];
  const double q_dst = li_dst(m, r) == -1 ? 0.0 : 0.0 + m->riv_q[r];
  const double net_river_flow = q_src - q_dst;
  const double net_land_flow = lil_net_land_flow(m, i);
  const double net_flow = net_river_flow + net_land_flow + m->li_land_runoff[i] - m->riv_abstraction[r];
  return net_flow * dt;
}

/* compute_external_inflow -> (inflow, abstraction_to_add)                          :1359-1382 */
static void lil_compute_external_inflow(const wfo_model* m, int64_t i, int64_t r, double dt, double out[2]) {
  if (m->riv_external_inflow[r] < 0.0) {
    const double available_volume = m->olf_storage[i] >= m->li_bankfull_storage[r]
                                        ? m->li_bankfull_storage[r] : m->riv_storage[r];
    const double abstraction = jl_min(-m->riv_external_inflow[r], available_volume / dt * 0.80);
    out[0] = -abstraction; out[1] = abstraction;
  } else {
    out[0] = m->riv_external_inflow[r]; out[1] = 0.0;
  }
}

/* compute_water_depths -> (river_h, land_h, river_storage)                         :1388-1416 */
void wfo_lil_compute_water_depths(wfo_model* m, double total_storage, int64_t r, int64_t i, double out[3]) {
  if (total_storage >= m->li_bankfull_storage[r]) {
    const double river_h = m->li_bankfull_depth[r] + (total_storage - m->li_bankfull_storage[r]) /
                                                         (m->li_land_x_length[i] * m->li_land_y_length[i]);
    out[0] = river_h;
    out[1] = river_h - m->li_bankfull_depth[r];
    out[2] = river_h * m->riv_flow_length[r] * m->riv_flow_width[r];
  } else {
    out[0] = total_storage / (m->riv_flow_length[r] * m->riv_flow_width[r]);
    out[1] = 0.0;
    out[2] = total_storage;
  }
}

/* compute_land_storage_change                                                      :1422-1437 */
double wfo_lil_compute_land_storage_change(wfo_model* m, int64_t i, double dt) {
  return (lil_net_land_flow(m, i) + m->li_land_runoff[i]) * dt;
}

/* update_river_and_land_storage_and_depth!                                         :1443-1485 */
void wfo_lil_update_river_and_land_storage_and_depth(wfo_model* m, int64_t i, double dt) {
  const int64_t r = m->land_river_indices[i];
  m->olf_storage[i] += wfo_lil_compute_river_storage_change(m, i, dt);
  if (m->olf_storage[i] < 0.0) {
    m->li_land_error[i] += fabs(m->olf_storage[i]);
    m->olf_storage[i] = 0.0;
  }
  double io[2], d[3];
  lil_compute_external_inflow(m, i, r, dt, io);
  m->olf_storage[i] += io[0] * dt;
  m->riv_actual_external_abstraction_cumulative[r] += io[1] * dt;
  wfo_lil_compute_water_depths(m, m->olf_storage[i], r, i, d);
  m->riv_h[r] = d[0];
  m->olf_h[i] = d[1];
  m->riv_storage[r] = d[2];
}

/* update_land_storage_and_depth!                                                   :1491-1514 */
void wfo_lil_update_land_storage_and_depth(wfo_model* m, int64_t i, double dt) {
  m->olf_storage[i] += wfo_lil_compute_land_storage_change(m, i, dt);
  if (m->olf_storage[i] < 0.0) {
    m->li_land_error[i] += fabs(m->olf_storage[i]);
    m->olf_storage[i] = 0.0;
  }
  m->olf_h[i] = m->olf_storage[i] / (m->li_land_x_length[i] * m->li_land_y_length[i]);
}

/* local_inertial_update_water_depth!                                               :1520-1546 */
void wfo_lil_update_water_depth(wfo_model* m, double dt) {
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i) {
    const int64_t r = m->land_river_indices[i];
    if (r >= 0) {
      if (!(m->cfg.nres > 0 && m->riv_reservoir[r] >= 0))      /* !reservoir_outlet[i] */
        wfo_lil_update_river_and_land_storage_and_depth(m, i, dt);
    } else {
      wfo_lil_update_land_storage_and_depth(m, i, dt);
    }
  }
}

/* update_bc_overland_flow_model!                                                   :1080-1097 */
void wfo_update_bc_overland_flow_model(wfo_model* m) {
  PFOR for (int64_t i = 0; i < m->cfg.n; ++i)
    m->li_land_runoff[i] = (m->net_runoff[i] + m->net_runoff_river[i]) * m->area[i];
  for (int64_t r = 0; r < m->cfg.nriv; ++r) {
    const int64_t li = m->river_land_indices[r];
    m->li_land_runoff[li] += m->ssf_to_river_average[li];      /* get_flux_to_river  lsf.jl:346 */
  }
}

/* update_overland_flow_model!(overland, river, domain, clock, dt; update_h = false) :1153-1194 */
void wfo_lil_update_overland_flow_model(wfo_model* m, double dt) {
  const int64_t n = m->cfg.n, nriv = m->cfg.nriv;
  for (int64_t i = 0; i < m->cfg.nres; ++i) {                  /* set_reservoir_vars! */
    m->res_inflow_cumulative[i] = 0.0;
    m->res_actual_external_abstraction_cumulative[i] = 0.0;
    m->res_outflow_cumulative[i] = 0.0;
    m->res_actevap_cumulative[i] = 0.0;
  }
  PFOR for (int64_t i = 0; i < nriv; ++i) {                    /* set_flow_vars!(river) */
    m->riv_q_cumulative[i] = 0.0;
    m->riv_actual_external_abstraction_cumulative[i] = 0.0;
  }
  PFOR for (int64_t i = 0; i < n; ++i) {                       /* set_flow_vars!(overland) :1127-1132 */
    m->li_land_qx_cumulative[i] = 0.0;
    m->li_land_qy_cumulative[i] = 0.0;
  }
  double t = 0.0;
  m->substeps_river = 0;
  while (t < dt) {
    const double dt_river = wfo_li_stable_timestep(m);
    const double dt_land = wfo_lil_stable_timestep(m);
    double dt_s = jl_min(dt_river, dt_land);
    dt_s = check_timestepsize(dt_s, t, dt);
    wfo_lil_update_fluxes(m, dt_s);
    wfo_lil_update_inflow_reservoir(m);
    /* staggered_scheme_river_update!(river, domain, dt_s, update_h = false)      :762-794 */
    wfo_li_update_river_channel_flow(m, dt_s);
    wfo_li_update_bc_reservoir_model(m, dt_s);
    wfo_lil_update_water_depth(m, dt_s);
    t += dt_s;
    m->substeps_river++;
  }
  m->substeps_land = m->substeps_river;
  PFOR for (int64_t i = 0; i < nriv; ++i) {                    /* average_flow_vars!(river) */
    m->riv_q_average[i] = m->riv_q_cumulative[i] / dt;
    m->riv_actual_external_abstraction_average[i] =
        m->riv_actual_external_abstraction_cumulative[i] / dt;
  }
  PFOR for (int64_t i = 0; i < n; ++i) {                       /* average_flow_vars!(overland) :1138-1147 */
    m->li_land_qx_average[i] = m->li_land_qx_cumulative[i] / dt;
    m->li_land_qy_average[i] = m->li_land_qy_cumulative[i] / dt;
  }
  for (int64_t i = 0; i < m->cfg.nres; ++i) {                  /* average_reservoir_vars! */
    m->res_outflow_average[i] = m->res_outflow_cumulative[i] / dt;
    m->res_inflow_average[i] = m->res_inflow_cumulative[i] / dt;
    m->res_actual_external_abstraction_average[i] =
        m->res_actual_external_abstraction_cumulative[i] / dt;
  }
}

/* surface_routing.jl:7-46; :62-86 with local-inertial land AND river routing */
void wfo_surface_routing(wfo_model* m, double dt) {
  if (m->cfg.land_routing == 1) {
    wfo_update_bc_overland_flow_model(m);
    /* update_inflow!(reservoir, river_flow, subsurface_flow, network)  surface_staggered_scheme.jl:1103-1114 */
    for (int64_t i = 0; i < m->cfg.nres; ++i) {
      const int64_t li = m->river_land_indices[m->reservoir_river_indices[i]];
      m->res_inflow_subsurface[i] = m->ssf_q_average[li] + m->ssf_to_river_average[li];
    }
    wfo_lil_update_overland_flow_model(m, dt);
    return;
  }
  wfo_update_lateral_inflow_overland(m);
  wfo_update_overland_flow_model(m, dt);
  wfo_update_lateral_inflow_river(m);
  wfo_update_inflow_reservoir(m);
  wfo_update_river_flow_model(m, dt);
}

/* sbm_model.jl:60-92 */
void wfo_update_model(wfo_model* m, double dt) {
  wfo_update_land_hydrology_model(m, dt);
  wfo_exchange_recharge(m);
  wfo_update_subsurface_flow_model(m, dt);
  wfo_update_soil_water_storage(m, dt);
  wfo_surface_routing(m, dt);
  wfo_update_total_water_storage(m);
}
