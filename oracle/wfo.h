/*
 * wfo.h -- CPU ORACLE for the wflow_sbm + kinematic-wave hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (wflow.jl_b200/, include/) may include,
 * link or call this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, as the checker / reported CPU baseline.
 *
 * It is a plain-C restatement of the reference Julia algorithm (Deltares/Wflow.jl,
 * /root/reference/Wflow/src), sweep by sweep in the reference's own order, Float64,
 * compiled with -ffp-contract=off. Every function cites the reference file:line it follows.
 *
 * Parity status: pinned at unit level by the reference's own known-answer tests
 * (tests/test_oracle_golden.py: Wflow/test/land_process.jl, soil.jl, routing_process.jl,
 * utils.jl, subdomains.jl). End-to-end parity on real data is UNPINNED (the reference is pure
 * Julia, `julia` is not installed here and the Moselle test data is a DVC pointer).
 *
 * Layout: structure of arrays; layered fields are CELL-major like Julia's
 * Vector{SVector{N,Float64}} (value of layer k of cell i at [i*N + k]); integers are int64;
 * all node indices inside this oracle are 0-based (the Python wrapper converts 1-based
 * artefacts).
 */
#ifndef WFO_H
#define WFO_H
#include <stdint.h>

/* kind: 0 = land scalar (n), 1 = land layered (n*N), 2 = land layered+1 (n*(N+1)),
 *       3 = river scalar (nriv), 4 = reservoir scalar (nres), 5 = river x floodplain level,
 *       6 = land scalar of the local-inertial overland flow (n with land_routing = 1, else 0) */
#define WFO_FIELDS(X) \
  /* forcing (forcing.jl:2-10) */ \
  X(precipitation, 0) X(potential_evaporation, 0) X(temperature, 0) \
  /* vegetation parameters (vegetation/parameters.jl) */ \
  X(leaf_area_index, 0) X(storage_specific_leaf, 0) X(storage_wood, 0) \
  X(light_extinction_coefficient, 0) X(canopy_gap_fraction, 0) X(maximum_canopy_storage, 0) \
  X(crop_coefficient, 0) X(rooting_depth, 0) \
  /* interception (canopy.jl:4-29) */ \
  X(evaporation_to_precipitation_ratio, 0) X(canopy_potevap, 0) X(interception_rate, 0) \
  X(canopy_storage, 0) X(stemflow, 0) X(throughfall, 0) \
  /* snow (snow.jl:4-45) */ \
  X(temperature_threshold_snowfall, 0) X(temperature_interval_snowfall, 0) \
  X(temperature_threshold_melt, 0) X(degree_day_factor, 0) X(water_holding_capacity, 0) \
  X(snow_storage, 0) X(snow_water, 0) X(snow_water_equivalent, 0) X(snow_melt, 0) \
  X(snow_runoff, 0) X(effective_precip, 0) X(snow_precip, 0) X(liquid_precip, 0) \
  X(snow_in, 0) X(snow_out, 0) \
  /* glacier (glacier.jl) */ \
  X(glacier_temperature_threshold_melt, 0) X(glacier_degree_day_factor, 0) \
  X(glacier_snow_to_ice_fraction, 0) X(glacier_fraction, 0) X(glacier_store, 0) \
  X(glacier_melt, 0) \
  /* open water runoff (runoff.jl:4-27) */ \
  X(runoff_water_flux_surface, 0) X(waterdepth_land, 0) X(waterdepth_river, 0) \
  X(runoff_river, 0) X(net_runoff_river, 0) X(runoff_land, 0) \
  X(actual_open_water_evaporation_land, 0) X(actual_open_water_evaporation_river, 0) \
  /* shared land parameters (domain.jl:218-261) */ \
  X(river_fraction, 0) X(water_fraction, 0) X(area, 0) X(slope, 0) X(flow_length, 0) \
  X(flow_width, 0) X(surface_flow_width, 0) X(flow_fraction_to_river, 0) \
  /* soil parameters (soil.jl:87-150) */ \
  X(theta_s, 0) X(theta_r, 0) X(theta_fc, 0) X(soil_water_capacity, 0) \
  X(vertical_hydraulic_conductivity_factor, 1) X(air_entry_pressure, 0) X(soil_thickness, 0) \
  X(actual_layer_thickness, 1) X(cumulative_layer_depth, 2) \
  X(infiltration_capacity_compacted_soil, 0) X(infiltration_capacity_soil, 0) \
  X(maximum_leakage, 0) X(cap_hmax, 0) X(cap_n, 0) X(brooks_corey_exponent, 1) X(w_soil, 0) \
  X(cf_soil, 0) X(compacted_soil_area_fraction, 0) X(wet_root_distribution_parameter, 0) \
  X(rootfraction, 1) X(h1, 0) X(h2, 0) X(h3_high, 0) X(h3_low, 0) X(h4, 0) X(alpha_h1, 0) \
  X(soil_fraction, 0) X(kv_0, 0) X(hydraulic_conductivity_scale_parameter, 0) X(z_exp, 0) \
  X(kv, 1) X(z_layered, 0) \
  /* soil boundary conditions (soil.jl:203-211) */ \
  X(soil_water_flux_surface, 0) X(potential_transpiration, 0) X(potential_soilevaporation, 0) \
  /* soil variables (soil.jl:4-84) */ \
  X(h3, 0) X(unsaturated_store_capacity, 0) X(unsaturated_layer_depth, 1) \
  X(unsaturated_layer_thickness, 1) X(saturated_water_depth, 0) X(drainable_water_depth, 0) \
  X(water_table_depth, 0) X(transpiration, 0) X(actual_evaporation_unsaturated_store, 0) \
  X(soil_evaporation, 0) X(soil_evaporation_saturated_zone, 0) X(actual_capillary_flux, 0) \
  X(actual_evaporation_saturated_zone, 0) X(actual_evapotranspiration, 0) \
  X(actual_infiltration, 0) X(actual_infiltration_soil, 0) \
  X(actual_infiltration_compacted_soil, 0) X(infiltration, 0) X(infiltration_excess, 0) \
  X(saturation_excess_water, 0) X(exfiltration_saturated_water, 0) X(excess_water_soil, 0) \
  X(excess_water_compacted_soil, 0) X(runoff, 0) X(net_runoff, 0) \
  X(volumetric_water_content, 1) X(relative_volumetric_water_content, 1) \
  X(root_zone_storage, 0) X(volumetric_water_content_root_zone, 0) \
  X(relative_volumetric_water_content_root_zone, 0) X(unsaturated_store_depth, 0) \
  X(transfer, 0) X(recharge, 0) X(actual_leakage, 0) X(total_storage, 0) \
  X(total_soil_water_storage, 0) X(soil_surface_temperature, 0) X(f_infiltration_reduction, 0) \
  /* lateral subsurface flow (lateral_subsurface_flow.jl:2-54) + recharge BC */ \
  X(kh_0, 0) X(ssf_khfrac, 0) X(ssf_kh, 0) X(ssf_soil_thickness, 0) X(specific_yield, 0) X(ssf_top, 0) \
  X(ssf_water_table_depth, 0) X(ssf_head, 0) X(ssf_exfiltwater_cumulative, 0) \
  X(ssf_exfiltwater_average, 0) X(ssf_q, 0) X(ssf_q_cumulative, 0) X(ssf_q_average, 0) \
  X(ssf_q_in, 0) X(ssf_q_in_cumulative, 0) X(ssf_q_in_average, 0) X(ssf_q_max, 0) \
  X(ssf_to_river_cumulative, 0) X(ssf_to_river_average, 0) X(ssf_q_net_bnds, 0) \
  X(ssf_q_net_cumulative, 0) X(ssf_q_net_average, 0) X(ssf_storage, 0) \
  X(recharge_rate, 0) X(recharge_flux, 0) X(recharge_flux_cumulative, 0) \
  X(recharge_flux_average, 0) \
  /* overland flow (surface_kinwave.jl:154-185) */ \
  X(olf_alpha, 0) X(olf_inwater, 0) X(olf_q, 0) X(olf_qlat, 0) X(olf_qin, 0) \
  X(olf_qin_cumulative, 0) X(olf_qin_average, 0) X(olf_q_cumulative, 0) X(olf_q_average, 0) \
  X(olf_storage, 0) X(olf_h, 0) X(olf_to_river_cumulative, 0) X(olf_to_river_average, 0) \
  /* river flow (surface_kinwave.jl:5-29) */ \
  X(riv_flow_length, 3) X(riv_flow_width, 3) X(riv_alpha, 3) X(riv_external_inflow, 3) \
  X(riv_abstraction, 3) X(riv_actual_external_abstraction_cumulative, 3) \
  X(riv_actual_external_abstraction_average, 3) X(riv_inwater, 3) X(riv_q, 3) \
  X(riv_qlat, 3) X(riv_qin, 3) X(riv_qin_cumulative, 3) X(riv_qin_average, 3) \
  X(riv_q_cumulative, 3) X(riv_q_average, 3) X(riv_storage, 3) X(riv_h, 3) \
  /* local-inertial river flow on the staggered grid (surface_staggered_scheme.jl:1-40,157-186): \
   * edge i is the edge leaving node i (init_staggered_river_flow :225-233), so the edge \
   * arrays are river-sized; riv_q / riv_q_cumulative / riv_q_average hold the edge discharge; \
   * li_ghost_h: water depth of the ghost node downstream of a pit (riverdepth_bc) */ \
  X(li_zb, 3) X(li_zb_at_edge, 3) X(li_mannings_n_sq_at_edge, 3) X(li_flow_length_at_edge, 3) \
  X(li_flow_width_at_edge, 3) X(li_ghost_h, 3) X(li_error, 3) X(li_zs_at_edge, 3) \
  X(li_water_depth_at_edge, 3) \
  /* 1-D floodplain of the local-inertial river (floodplain_1d__flag; floodplain.jl:150-215, \
   * surface_staggered_scheme.jl:440-533,674-712): variables by node / by the edge leaving the \
   * node; fp_profile_*: FloodPlainProfile tables (floodplain.jl:5-22), one row of fp_levels \
   * values per node (Julia: profile.x[level, node]) */ \
  X(fp_h, 3) X(fp_storage, 3) X(fp_q, 3) X(fp_q_cumulative, 3) X(fp_q_average, 3) X(fp_error, 3) \
  X(fp_water_depth_at_edge, 3) X(fp_mannings_n_sq_at_edge, 3) X(fp_zb_at_edge, 3) \
  X(li_bankfull_storage, 3) X(li_bankfull_depth, 3) X(riv_q_channel_average, 3) \
  /* 1-D floodplain of the KINEMATIC-WAVE river (floodplain.jl:238-262 FloodPlainParameters / \
   * Variables; surface_kinwave.jl:387-432,567-601): Manning flow capacity + accucapacityflux */ \
  X(fp_mannings_n, 3) X(fp_slope, 3) X(fp_flow_capacity, 3) X(fp_qin, 3) \
  X(fp_qin_cumulative, 3) X(fp_qin_average, 3) X(riv_floodplain_water_exchange, 3) \
  X(fp_profile_storage, 5) X(fp_profile_width, 5) X(fp_profile_flow_area, 5) \
  X(fp_profile_wetted_perimeter, 5) \
  /* reservoirs (routing/surface/reservoir.jl:5-44 parameters, 200-217 variables, 251-272 BC); \
   * res_outflow_curve_type holds ReservoirOutflowType as a number (2 free_weir, 3 \
   * modified_puls, 4 simple) */ \
  X(res_area, 4) X(res_outflow_curve_type, 4) X(res_maximum_storage, 4) X(res_threshold, 4) \
  X(res_rating_curve_coefficient, 4) X(res_rating_curve_exponent, 4) X(res_maximum_release, 4) \
  X(res_demand, 4) X(res_target_minimum_fraction, 4) X(res_target_full_fraction, 4) \
  X(res_inflow_subsurface, 4) X(res_inflow_overland, 4) X(res_inflow_cumulative, 4) \
  X(res_inflow_average, 4) X(res_external_inflow, 4) \
  X(res_actual_external_abstraction_cumulative, 4) X(res_actual_external_abstraction_average, 4) \
  X(res_precipitation, 4) X(res_evaporation, 4) X(res_waterlevel, 4) X(res_storage, 4) \
  X(res_outflow, 4) X(res_outflow_cumulative, 4) X(res_outflow_average, 4) X(res_outflow_obs, 4) \
  X(res_actevap_cumulative, 4) \
  /* 2-D local-inertial overland flow (land_routing = "local_inertial"): \
   * LocalInertialOverlandFlowParameters / Variables / BC (surface_staggered_scheme.jl:840-891, \
   * 963-968) and x_length / y_length of LandParameters (domain.jl). The reference's flow vectors \
   * hold n + 1 entries whose last one stays 0 (the edge "outside"); here they hold n and index \
   * n reads as 0. The model's h and storage live in olf_h / olf_storage \
   * (overland_flow.variables.h / .storage whatever the routing method). */ \
  X(li_land_xwidth_at_edge, 6) X(li_land_ywidth_at_edge, 6) X(li_land_zx_max_at_edge, 6) \
  X(li_land_zy_max_at_edge, 6) X(li_land_mannings_n_sq_at_edge, 6) X(li_land_z, 6) \
  X(li_land_x_length, 6) X(li_land_y_length, 6) X(li_land_runoff, 6) \
  X(li_land_qx0, 6) X(li_land_qy0, 6) X(li_land_qx, 6) X(li_land_qy, 6) \
  X(li_land_qx_cumulative, 6) X(li_land_qy_cumulative, 6) X(li_land_qx_average, 6) \
  X(li_land_qy_average, 6) X(li_land_error, 6)

/* Per-domain network artefacts needed to walk the routing in the reference's order
 * (network.jl:48-81): all 0-based here. */
typedef struct {
  int64_t n;
  const int64_t* up_ptr;      /* CSR by TOPOSORT POSITION (utils.jl:61-71), size n+1 */
  const int64_t* up_idx;      /* upstream node ids, ascending per position            */
  int64_t n_levels;           /* length(order_of_subdomains)                          */
  const int64_t* level_ptr;   /* CSR: subdomain ids per level                         */
  const int64_t* level_sub;
  int64_t n_sub;
  const int64_t* sub_ptr;     /* CSR: per subdomain, its nodes in walk order          */
  const int64_t* sub_nodes;   /* order_subdomain[m]   (node id v)                     */
  const int64_t* sub_pos;     /* subdomain_indices[m] (toposort position n)           */
  const int64_t* down;        /* outneighbors(graph, v): downstream node or -1 (pit)  */
  const int64_t* order;       /* network.order (topological_sort_by_dfs), node ids    */
  const int64_t* in_ptr;      /* inneighbors(graph, v) by NODE id, CSR (n+1), ascending */
  const int64_t* in_idx;
} wfo_network;

typedef struct {
  int64_t n, nriv, N;         /* land cells, river cells, maximum_number_of_layers    */
  int64_t nres;               /* reservoirs (reservoir__flag)                         */
  int32_t gash;               /* 1: Gash (dt >= 23 h), 0: modified Rutter  sbm.jl:26  */
  int32_t has_lai;            /* cyclic LAI present (canopy.jl:65,128)                */
  int32_t snow, glacier;      /* snow__flag, glacier__flag                            */
  int32_t soil_infiltration_reduction;
  int32_t kv_profile;         /* 0 exponential, 1 exponential_constant, 2 layered, 3 layered_exponential */
  int32_t adaptive;           /* kinematic_wave__adaptive_time_step_flag              */
  int32_t snow_transport;     /* snow_gravitational_transport__flag                   */
  int32_t river_routing;      /* 0 kinematic_wave, 1 local_inertial                   */
  int32_t li_froude_limit;    /* river_water_flow__froude_limit_flag                  */
  int32_t li_ghost_nodes;     /* pits drain to a ghost node (the model; 0: unit tests) */
  double li_alpha, li_h_thresh; /* river_local_inertial_flow__alpha_coefficient, river_water_flow_threshold__depth */
  int32_t nthreads;           /* OpenMP threads (0 = default)                         */
  double dt_land, dt_river, dt_ssf, ssf_alpha_coefficient;
  int32_t fp_levels;          /* floodplain_1d__flag: flood depths of the profile, 0 = none */
  double fp_depth[16];        /* profile.depth                                          */
  /* land_routing = "local_inertial" (with river_routing = "local_inertial"): 2-D overland flow */
  int32_t land_routing;       /* 0 kinematic_wave, 1 local_inertial                     */
  int32_t li_land_froude_limit; /* land_surface_water_flow__froude_limit_flag           */
  double li_land_alpha;       /* land_local_inertial_flow__alpha_coefficient (0.7)      */
  double li_land_theta;       /* land_local_inertial_flow__theta_coefficient (1.0)      */
  double li_land_h_thresh;    /* land_surface_water_flow_threshold__depth (1e-3)        */
} wfo_config;

typedef struct wfo_model {
  wfo_config cfg;
#define X(name, kind) double* name;
  WFO_FIELDS(X)
#undef X
  int64_t* number_of_layers;  /* n */
  int64_t* n_unsatlayers;     /* n */
  int64_t* nlayers_kv;        /* n (layered_exponential only) */
  int64_t* river_land_indices;/* nriv, 0-based land index of each river cell */
  int64_t* reservoir_river_indices; /* nres, 0-based river node of each reservoir     */
  int64_t* riv_reservoir;     /* nriv: NetworkRiver.reservoir_indices - 1 (-1 = none), derived */
  /* EdgeConnectivity of the land network (network.jl:27-33,136-153), 0-based, n = no neighbour */
  int64_t *edge_x_up, *edge_x_down, *edge_y_up, *edge_y_down;
  int64_t* land_river_indices;/* n: NetworkLand.river_indices - 1 (-1: the cell holds no river) */
  int64_t* newton_trace_land; /* n / nriv or NULL: Newton iterations of kinematic_wave per node, */
  int64_t* newton_trace_river;/* summed over the sub-steps (iteration-count parity tests)       */
  wfo_network land, river;
  double* scratch;            /* max(n, nriv) doubles: stable_timesteps */
  /* statistics */
  int64_t newton_iters_land, newton_iters_river, newton_calls_land, newton_calls_river;
  int64_t newton_maxit_land, newton_maxit_river;
  int64_t substeps_land, substeps_river, substeps_ssf;
  double dt_hist[8];
} wfo_model;

#ifdef __cplusplus
extern "C" {
#endif
/* field table */
int  wfo_num_fields(void);
const char* wfo_field_name(int id);
int  wfo_field_kind(int id);
wfo_model* wfo_new(void);
void wfo_free(wfo_model*);
int  wfo_set_ptr(wfo_model*, const char* name, double* p);
int  wfo_set_iptr(wfo_model*, const char* name, int64_t* p);
void wfo_set_network(wfo_model*, int which, const wfo_network* net);
wfo_config* wfo_cfg(wfo_model*);

/* model-level sweeps */
void wfo_update_land_hydrology_model(wfo_model*, double dt);   /* sbm.jl:82-132  */
void wfo_exchange_recharge(wfo_model*);                        /* sbm_model.jl:74-84 */
void wfo_kh_layered_profile(wfo_model*);                       /* utils.jl:792-895   */
void wfo_update_subsurface_flow_model(wfo_model*, double dt);  /* lateral_subsurface_flow.jl:279-304 */
void wfo_update_soil_water_storage(wfo_model*, double dt);     /* soil.jl:1294-1392 */
void wfo_surface_routing(wfo_model*, double dt);               /* surface_routing.jl:7-46 */
void wfo_update_overland_flow_model(wfo_model*, double dt);
void wfo_update_river_flow_model(wfo_model*, double dt);
void wfo_update_lateral_inflow_overland(wfo_model*);
void wfo_update_lateral_inflow_river(wfo_model*);
void wfo_update_inflow_reservoir(wfo_model*);                  /* surface_kinwave.jl:772-805 */
void wfo_kinwave_river_update(wfo_model*, double dt_s);        /* test hook: one sub-step    */
/* local-inertial river flow: stable_timestep (:1004-1020), update_river_channel_flow! (:326-383),
 * update_bc_reservoir_model! (:627-661), update_water_depth_and_storage! (:723-759) */
double wfo_li_stable_timestep(wfo_model*);
void wfo_li_update_river_channel_flow(wfo_model*, double dt_s);
void wfo_li_update_bc_reservoir_model(wfo_model*, double dt_s);
void wfo_li_update_water_depth_and_storage(wfo_model*, double dt_s);
void wfo_li_update_floodplain_flow(wfo_model*, double dt_s);
void wfo_li_update_floodplain_water_depth_and_storage(wfo_model*, double dt_s);
/* 2-D local-inertial overland flow coupled to the local-inertial river (surface_staggered_scheme.jl):
 * stable_timestep (:1022-1043), update_directional_flow! (:1201-1271; i 0-based),
 * local_inertial_update_fluxes! (:1276-1295), update_inflow_reservoir! (:1301-1319),
 * local_inertial_update_water_depth! (:1520-1546) and its per-cell parts (:1325-1514),
 * update_bc_overland_flow_model! (:1080-1097), update_overland_flow_model! (:1153-1194) */
double wfo_lil_stable_timestep(wfo_model*);
void wfo_lil_update_directional_flow(wfo_model*, int64_t i, double dt_s, int is_x_direction);
void wfo_lil_update_fluxes(wfo_model*, double dt_s);
void wfo_lil_update_inflow_reservoir(wfo_model*);
double wfo_lil_compute_river_storage_change(wfo_model*, int64_t i, double dt_s);
double wfo_lil_compute_land_storage_change(wfo_model*, int64_t i, double dt_s);
void wfo_lil_compute_water_depths(wfo_model*, double total_storage, int64_t river_idx, int64_t i,
                                  double out[3]);
void wfo_lil_update_river_and_land_storage_and_depth(wfo_model*, int64_t i, double dt_s);
void wfo_lil_update_land_storage_and_depth(wfo_model*, int64_t i, double dt_s);
void wfo_lil_update_water_depth(wfo_model*, double dt_s);
void wfo_update_bc_overland_flow_model(wfo_model*);
void wfo_lil_update_overland_flow_model(wfo_model*, double dt);
double wfo_local_inertial_flow_rect(double theta, double q0, double qd, double qu, double zs0,
                                    double zs1, double hf, double width, double length,
                                    double mannings_n_sq, int froude_limit, double dt);
/* test hooks: update_reservoir_model!(reservoir, i, inflow, dt) reservoir.jl:585-634 and
 * update_reservoir_model!(reservoir, river variables, network, v, dt) surface_kinwave.jl:441-489 */
void wfo_update_reservoir_model(wfo_model*, int64_t i, double inflow, double dt);
void wfo_update_reservoir_at_node(wfo_model*, int64_t v, double dt);
double wfo_local_inertial_flow(double q0, double zs0, double zs1, double hf, double A, double R,
                               double length, double mannings_n_sq, int froude_limit, double dt);
void wfo_river_channel_floodplain_exchange(wfo_model*, double dt_s);
void wfo_update_floodplain_model(wfo_model*, double dt_s);
void wfo_update_total_water_storage(wfo_model*);               /* sbm.jl:143-182 */
void wfo_update_model(wfo_model*, double dt);                  /* sbm_model.jl:60-92 */
void wfo_update_diagnostic_vars(wfo_model*);                   /* soil.jl:1400-1436 */
int  wfo_sweep(wfo_model*, const char* name, double dt);                   /* test hook */
void wfo_get_stats(wfo_model*, int64_t out[9]);
/* OpenMP team size of the sweeps (a launcher such as torchrun exports OMP_NUM_THREADS=1);
 * returns the value in effect. n <= 0 only queries. */
int wfo_set_num_threads(int n);

/* scalar kernels exported for the known-answer tests */
void wfo_rainfall_interception_gash(double cmax, double e_r, double gap, double p, double cs,
                                    double maxevap, double dt, double out[4]);
void wfo_rainfall_interception_modrut(double p, double pe, double cs, double gap, double cmax,
                                      double dt, double out[4]);
void wfo_precipitation_hbv(double p, double t, double tti, double tt, double out[2]);
void wfo_snowpack_hbv(double snow, double snowwater, double snow_precip, double liquid_precip,
                      double t, double ttm, double cfmax, double whc, double dt, double out[5]);
void wfo_glacier_hbv(double gfrac, double gstore, double snow, double t, double ttm,
                     double cfmax, double sifrac, double maxrate, double dt, double out[4]);
void wfo_infiltration(double pot, double pathfrac, double cap_soil, double cap_path,
                      double ustorecap, double f_red, double dt, double out[2]);
void wfo_unsatzone_flow_layer(double usd, double kv_z, double l_sat, double c, double dt,
                              double out[2]);
double wfo_vwc_brooks_corey(double h, double hb, double ts, double tr, double c);
double wfo_head_brooks_corey(double vwc, double ts, double tr, double c, double hb);
double wfo_feddes_h3(double h3_high, double h3_low, double tpot);
double wfo_rwu_reduction_feddes(double h, double h1, double h2, double h3, double h4,
                                double alpha_h1);
double wfo_soil_temperature(double tsoil, double w, double t);
double wfo_infiltration_reduction_factor(double tsoil, double cf, int modelsnow, int flag);
double wfo_soil_evaporation_unsaturated_store(double pot, double usd, double ust, int64_t nu,
                                              double zi, double theta_e);
double wfo_soil_evaporation_saturated_store(double pot, int64_t nu, double lt, double zi,
                                            double theta_d, double dt);
void wfo_actual_infiltration_soil_path(double pot, double act, double pathfrac, double cap_soil,
                                       double cap_path, double f_red, double out[2]);
double wfo_scurve(double x, double a, double b, double c);
void wfo_kinematic_wave(double q_in, double q_prev, double q_lat, double alpha, double dt,
                        double dx, double out[2], int64_t* iters);
double wfo_kw_ssf_newton_raphson(double q, double constant_term, double celerity, double dt,
                                 double dx);
double wfo_ssf_celerity(double zi, double slope, double sy, double kh_0, double f, double z_exp,
                        int profile);
void wfo_kinematic_wave_ssf(wfo_model* m, double q_in, double q_prev, double zi_prev,
                            double q_net_bnds, double slope, double sy, double d, double dt,
                            double dx, double dw, double q_max, int64_t i, double out[4]);
void wfo_water_table_change(wfo_model* m, double net_flux, double sy, int64_t i, double dt,
                            double out[2]);
double wfo_stable_timestep_surface(const double* q, const double* alpha, const double* len,
                                   int64_t n, double p, double* work);
double wfo_round_sigdigits12(double v);
/* accucapacityflux!(flux, material, network, capacity, dt)  routing/utils.jl:82-109; `order`
 * and `down` 0-based (down < 0: pit); material is updated in place */
void wfo_accucapacityflux(double* flux, double* material, const int64_t* order,
                          const int64_t* down, int64_t n, const double* capacity, double dt);
double wfo_cld(double x, double y);
#ifdef __cplusplus
}
#endif
#endif
