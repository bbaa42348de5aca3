/*
 * wflow_b200_fields.h -- the Float64 arrays of the Julia model structs that live on the device.
 *
 * X(name, kind): kind 0 = land scalar (n), 1 = land layered (n x N), 2 = land layered+1
 * (n x (N+1)), 3 = river scalar (nriv), 4 = reservoir scalar (nres), 5 = river x floodplain profile
 * level (nriv x fp_levels, a node's levels contiguous like Julia's profile.x[level, node]), 6 = land
 * scalar of the 2-D local-inertial overland flow (n values with land_routing = 1, else none: no HBM
 * is spent on them). Names are the reference's struct field names; a
 * component prefix is added where two structs share a name (snow_/glacier_/ssf_/olf_/riv_/
 * recharge_/runoff_/soil_).
 *
 * Reference structs (all under /root/reference/Wflow/src):
 *   AtmosphericForcing forcing.jl:2-10; VegetationParameters vegetation/parameters.jl:2-19;
 *   InterceptionVariables canopy.jl:4-16, GashParameters :19-23; SnowHbv* snow/snow.jl:4-45;
 *   Glacier* glacier/glacier.jl:4-60; OpenWaterRunoff* surfacewater/runoff.jl:4-27;
 *   LandParameters domain.jl:2-29; SbmSoilParameters soil/soil.jl:87-150, SbmSoilBC :203-211,
 *   SbmSoilVariables :4-84, Kv* :213-244 (kv, z_layered: the layered profiles); LateralSsf* routing/subsurface/
 *   lateral_subsurface_flow.jl:2-54, RechargeVariables boundary_conditions.jl:204-213;
 *   OverLandFlowVariables routing/surface/surface_kinwave.jl:154-178, LandFlowBC :181-185;
 *   RiverFlowVariables :5-29, RiverFlowBC routing/surface/surface_flow.jl:9-34;
 *   RiverFlowStaggeredParameters / Variables routing/surface/surface_staggered_scheme.jl:1-40,
 *   157-186 (li_*: local-inertial river flow; edge i is the edge leaving node i, so the edge
 *   arrays are river-sized and riv_q holds the edge discharge; li_ghost_h: water depth of the
 *   ghost node downstream of a pit, riverdepth_bc);
 *   FloodPlainProfile / FloodPlainStaggeredParameters / Variables routing/surface/floodplain.jl:
 *   5-22,150-163,216-262 (fp_*: 1-D floodplain of the local-inertial river -- edge arrays by the
 *   edge leaving the node -- or of the kinematic-wave river -- fp_mannings_n, fp_slope,
 *   fp_flow_capacity, fp_qin*, riv_floodplain_water_exchange (RiverFlowBC); li_bankfull_*:
 *   bankfull_storage / bankfull_depth of the river parameters);
 *   ReservoirParameters routing/surface/reservoir.jl:5-44, ReservoirVariables :200-217,
 *   ReservoirBC :251-272 (res_outflow_curve_type holds ReservoirOutflowType as a number:
 *   2 free_weir, 3 modified_puls, 4 simple);
 *   LocalInertialOverlandFlowParameters / Variables / BC routing/surface/surface_staggered_scheme.jl:
 *   840-891,963-968 and x_length / y_length of LandParameters (li_land_*: the 2-D local-inertial
 *   overland flow. The reference's flow vectors qx, qy, ... hold n + 1 entries whose last one -- the
 *   edge to "outside" -- stays 0; here they hold n. The model's h and storage live in olf_h /
 *   olf_storage: overland_flow.variables.h / .storage whatever the routing method).
 */
#ifndef WFLOW_B200_FIELDS_H
#define WFLOW_B200_FIELDS_H

#define WFLOWB200_FIELDS(X) \
  X(precipitation, 0) X(potential_evaporation, 0) X(temperature, 0) \
  X(leaf_area_index, 0) X(storage_specific_leaf, 0) X(storage_wood, 0) \
  X(light_extinction_coefficient, 0) X(canopy_gap_fraction, 0) X(maximum_canopy_storage, 0) \
  X(crop_coefficient, 0) X(rooting_depth, 0) \
  X(evaporation_to_precipitation_ratio, 0) X(canopy_potevap, 0) X(interception_rate, 0) \
  X(canopy_storage, 0) X(stemflow, 0) X(throughfall, 0) \
  X(temperature_threshold_snowfall, 0) X(temperature_interval_snowfall, 0) \
  X(temperature_threshold_melt, 0) X(degree_day_factor, 0) X(water_holding_capacity, 0) \
  X(snow_storage, 0) X(snow_water, 0) X(snow_water_equivalent, 0) X(snow_melt, 0) \
  X(snow_runoff, 0) X(effective_precip, 0) X(snow_precip, 0) X(liquid_precip, 0) \
  X(snow_in, 0) X(snow_out, 0) \
  X(glacier_temperature_threshold_melt, 0) X(glacier_degree_day_factor, 0) \
  X(glacier_snow_to_ice_fraction, 0) X(glacier_fraction, 0) X(glacier_store, 0) \
  X(glacier_melt, 0) \
  X(runoff_water_flux_surface, 0) X(waterdepth_land, 0) X(waterdepth_river, 0) \
  X(runoff_river, 0) X(net_runoff_river, 0) X(runoff_land, 0) \
  X(actual_open_water_evaporation_land, 0) X(actual_open_water_evaporation_river, 0) \
  X(river_fraction, 0) X(water_fraction, 0) X(area, 0) X(slope, 0) X(flow_length, 0) \
  X(flow_width, 0) X(surface_flow_width, 0) X(flow_fraction_to_river, 0) \
  X(theta_s, 0) X(theta_r, 0) X(theta_fc, 0) X(soil_water_capacity, 0) \
  X(vertical_hydraulic_conductivity_factor, 1) X(air_entry_pressure, 0) X(soil_thickness, 0) \
  X(actual_layer_thickness, 1) X(cumulative_layer_depth, 2) \
  X(infiltration_capacity_compacted_soil, 0) X(infiltration_capacity_soil, 0) \
  X(maximum_leakage, 0) X(cap_hmax, 0) X(cap_n, 0) X(brooks_corey_exponent, 1) X(w_soil, 0) \
  X(cf_soil, 0) X(compacted_soil_area_fraction, 0) X(wet_root_distribution_parameter, 0) \
  X(rootfraction, 1) X(h1, 0) X(h2, 0) X(h3_high, 0) X(h3_low, 0) X(h4, 0) X(alpha_h1, 0) \
  X(soil_fraction, 0) X(kv_0, 0) X(hydraulic_conductivity_scale_parameter, 0) X(z_exp, 0) \
  X(kv, 1) X(z_layered, 0) \
  X(soil_water_flux_surface, 0) X(potential_transpiration, 0) X(potential_soilevaporation, 0) \
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
  X(kh_0, 0) X(ssf_khfrac, 0) X(ssf_kh, 0) X(ssf_soil_thickness, 0) X(specific_yield, 0) X(ssf_top, 0) \
  X(ssf_water_table_depth, 0) X(ssf_head, 0) X(ssf_exfiltwater_cumulative, 0) \
  X(ssf_exfiltwater_average, 0) X(ssf_q, 0) X(ssf_q_cumulative, 0) X(ssf_q_average, 0) \
  X(ssf_q_in, 0) X(ssf_q_in_cumulative, 0) X(ssf_q_in_average, 0) X(ssf_q_max, 0) \
  X(ssf_to_river_cumulative, 0) X(ssf_to_river_average, 0) X(ssf_q_net_bnds, 0) \
  X(ssf_q_net_cumulative, 0) X(ssf_q_net_average, 0) X(ssf_storage, 0) \
  X(recharge_rate, 0) X(recharge_flux, 0) X(recharge_flux_cumulative, 0) \
  X(recharge_flux_average, 0) \
  X(olf_alpha, 0) X(olf_inwater, 0) X(olf_q, 0) X(olf_qlat, 0) X(olf_qin, 0) \
  X(olf_qin_cumulative, 0) X(olf_qin_average, 0) X(olf_q_cumulative, 0) X(olf_q_average, 0) \
  X(olf_storage, 0) X(olf_h, 0) X(olf_to_river_cumulative, 0) X(olf_to_river_average, 0) \
  X(riv_flow_length, 3) X(riv_flow_width, 3) X(riv_alpha, 3) X(riv_external_inflow, 3) \
  X(riv_abstraction, 3) X(riv_actual_external_abstraction_cumulative, 3) \
  X(riv_actual_external_abstraction_average, 3) X(riv_inwater, 3) X(riv_q, 3) \
  X(riv_qlat, 3) X(riv_qin, 3) X(riv_qin_cumulative, 3) X(riv_qin_average, 3) \
  X(riv_q_cumulative, 3) X(riv_q_average, 3) X(riv_storage, 3) X(riv_h, 3) \
  X(li_zb, 3) X(li_zb_at_edge, 3) X(li_mannings_n_sq_at_edge, 3) X(li_flow_length_at_edge, 3) \
  X(li_flow_width_at_edge, 3) X(li_ghost_h, 3) X(li_error, 3) X(li_zs_at_edge, 3) \
  X(li_water_depth_at_edge, 3) \
  X(fp_h, 3) X(fp_storage, 3) X(fp_q, 3) X(fp_q_cumulative, 3) X(fp_q_average, 3) X(fp_error, 3) \
  X(fp_water_depth_at_edge, 3) X(fp_mannings_n_sq_at_edge, 3) X(fp_zb_at_edge, 3) \
  X(li_bankfull_storage, 3) X(li_bankfull_depth, 3) X(riv_q_channel_average, 3) \
  X(fp_mannings_n, 3) X(fp_slope, 3) X(fp_flow_capacity, 3) X(fp_qin, 3) \
  X(fp_qin_cumulative, 3) X(fp_qin_average, 3) X(riv_floodplain_water_exchange, 3) \
  X(fp_profile_storage, 5) X(fp_profile_width, 5) X(fp_profile_flow_area, 5) \
  X(fp_profile_wetted_perimeter, 5) \
  X(res_area, 4) X(res_outflow_curve_type, 4) X(res_maximum_storage, 4) X(res_threshold, 4) \
  X(res_rating_curve_coefficient, 4) X(res_rating_curve_exponent, 4) X(res_maximum_release, 4) \
  X(res_demand, 4) X(res_target_minimum_fraction, 4) X(res_target_full_fraction, 4) \
  X(res_inflow_subsurface, 4) X(res_inflow_overland, 4) X(res_inflow_cumulative, 4) \
  X(res_inflow_average, 4) X(res_external_inflow, 4) \
  X(res_actual_external_abstraction_cumulative, 4) X(res_actual_external_abstraction_average, 4) \
  X(res_precipitation, 4) X(res_evaporation, 4) X(res_waterlevel, 4) X(res_storage, 4) \
  X(res_outflow, 4) X(res_outflow_cumulative, 4) X(res_outflow_average, 4) X(res_outflow_obs, 4) \
  X(res_actevap_cumulative, 4) \
  X(li_land_xwidth_at_edge, 6) X(li_land_ywidth_at_edge, 6) X(li_land_zx_max_at_edge, 6) \
  X(li_land_zy_max_at_edge, 6) X(li_land_mannings_n_sq_at_edge, 6) X(li_land_z, 6) \
  X(li_land_x_length, 6) X(li_land_y_length, 6) X(li_land_runoff, 6) \
  X(li_land_qx0, 6) X(li_land_qy0, 6) X(li_land_qx, 6) X(li_land_qy, 6) \
  X(li_land_qx_cumulative, 6) X(li_land_qy_cumulative, 6) X(li_land_qx_average, 6) \
  X(li_land_qy_average, 6) X(li_land_error, 6)

#endif
