"""
Pins the CPU oracle (oracle/) against the reference's own known-answer tests
(/root/reference/Wflow/test/*.jl, transcribed; file:line cited per test). Tolerance is the
reference's own `≈` (isapprox default rtol = sqrt(eps) ≈ 1.5e-8); we use 1e-9.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import network as nw
from oracle import oracle as orc

RT = 1e-9


def approx(x):
    return pytest.approx(x, rel=RT, abs=1e-300)


def test_gash_land_process_1_60():
    dt = 86400.0
    p = 2.0833333333333333e-7
    out = orc.call_out("wfo_rainfall_interception_gash", 4, 0.0, 0.11, 0.24, p, 0.0015,
                       4.6296296296296295e-8, dt)
    assert out == (p, 0.0, 0.0, 0.0015)
    out = orc.call_out("wfo_rainfall_interception_gash", 4, 0.003, 0.11, 0.24, p, 0.0015,
                       4.6296296296296295e-8, dt)
    assert out[0] == approx(1.5703703703703703e-7)
    assert out[1] == approx(4.6296296296296295e-8)
    assert out[2] == approx(5.0e-9)
    assert out[3] == 0.0015
    out = orc.call_out("wfo_rainfall_interception_gash", 4, 0.003, 0.11, 0.24,
                       1.1574074074074074e-8, 0.0015, 4.6296296296296295e-8, dt)
    assert out[0] == approx(2.7777777777777776e-9)
    assert out[1] == approx(8.518518518518518e-9)
    assert out[2] == approx(2.7777777777777777e-10)


def test_gash_zero_ratio_nan_path():
    # App. A-3 of SURVEY: E/R == 0 -> -Smax/(0*dt)*log(1) = -Inf*0 = NaN -> small-storm branch
    out = orc.call_out("wfo_rainfall_interception_gash", 4, 0.003, 0.0, 0.24, 1e-7, 0.0, 1e-6,
                       86400.0)
    fi = 1.0 - 1.1 * 0.24
    assert out[1] == fi * 1e-7


def test_modrut_land_process_62_99():
    dt = 86400.0
    out = orc.call_out("wfo_rainfall_interception_modrut", 4, 9.953703703703703e-8,
                       4.398148148148148e-8, 0.0015, 0.45, 0.0028, dt)
    assert out[0] == approx(4.4791666666666666e-8)
    assert out[1] == approx(4.398148148148148e-8)
    assert out[2] == approx(4.479166666666667e-9)
    assert out[3] == approx(0.002043)
    out = orc.call_out("wfo_rainfall_interception_modrut", 4, 1.1574074074074074e-8,
                       4.398148148148148e-8, out[3], 0.95, 0.0028, dt)
    assert out[0] == approx(1.099537037037037e-8)
    assert out[1] == approx(2.3645833333333334e-8)
    assert out[2] == approx(5.787037037037037e-10)
    assert out[3] == pytest.approx(0.0, abs=1e-18)


def test_precipitation_hbv_land_process_101_140():
    p = 3.4837962962962964e-7
    out = orc.call_out("wfo_precipitation_hbv", 2, p, 273.69, 2.0, 273.15)
    assert out[0] == approx(8.012731481481482e-8)
    assert out[1] == approx(2.682523148148148e-7)
    assert orc.call_out("wfo_precipitation_hbv", 2, p, 273.69, 0.0, 273.15) == (0.0, p)
    assert orc.call_out("wfo_precipitation_hbv", 2, p, 272.15, 0.0, 273.15) == (p, 0.0)


def test_snowpack_hbv_land_process_142_191():
    args = (0.2015, 0.015, 8.012731481481482e-8, 2.682523148148148e-7)
    out = orc.call_out("wfo_snowpack_hbv", 5, *args, 273.69, 273.15, 2.8935185185185185e-8, 0.10,
                       86400.0)
    for got, want in zip(out, (0.207073, 0.0207073, 0.22778030000000002, 1.5625e-8,
                               2.1782060185185186e-7)):
        assert got == approx(want)
    # the reference test rebinds `snow_water` to the first call's output before the second call
    args = (0.2015, out[1], 8.012731481481482e-8, 2.682523148148148e-7)
    out = orc.call_out("wfo_snowpack_hbv", 5, *args, 272.65, 273.15, 2.8935185185185185e-8, 0.10,
                       86400.0)
    for got, want in zip(out, (0.20848550000000002, 0.02084855, 0.22933404999999998, 0.0,
                               2.658940972222222e-7)):
        assert got == pytest.approx(want, rel=RT, abs=1e-18)


def test_glacier_hbv_land_process_193_219():
    out = orc.call_out("wfo_glacier_hbv", 4, 0.35, 0.5, 0.0095, 278.15, 273.15,
                       3.935185185185185e-8, 2.3148148148148148e-6, 9.259259259259259e-8, 86400.0)
    for got, want in zip(out, (0.008835, 2.199074074074074e-8, 0.4849, 1.9675925925925924e-7)):
        assert got == approx(want)


def test_infiltration_land_process_221_241():
    out = orc.call_out("wfo_infiltration", 2, 3.18287037037037e-7, 0.2, 5.787037037037037e-7,
                       5.787037037037037e-8, 0.0235, 1.0, 86400.0)
    assert out[1] == approx(5.787037037037037e-9)


def test_unsatzone_flow_layer_land_process_243_261():
    # The reference ASSIGNS (does not assert) usd_new/sum_ast here; usd_new matches, and
    # the assigned sum_ast (1.6255e-7) equals sum_ast*dt of the restated value.
    out = orc.call_out("wfo_unsatzone_flow_layer", 2, 0.043500000000000004, 2.962962962962963e-6,
                       0.135, 12.6, 86400.0)
    assert out[0] == approx(0.04349983744545384)
    assert out[1] * 86400.0 == pytest.approx(1.6255454615829024e-7, rel=1e-6)
    assert orc.call_out("wfo_unsatzone_flow_layer", 2, 0.0, 2.962962962962963e-6, 0.135, 12.6,
                        86400.0) == (0.0, 0.0)


def test_brooks_corey_land_process_263_293():
    L = orc.lib()
    h = L.wfo_head_brooks_corey(0.25, 0.6, 0.15, 10.5, -0.1)
    assert h == approx(-0.9062998208338441)
    assert L.wfo_vwc_brooks_corey(h, -0.1, 0.6, 0.15, 10.5) == approx(0.25 + 0.15)
    h = L.wfo_head_brooks_corey(0.25, 0.6, 0.15, 2.0, -0.1)
    assert h == -0.1
    assert L.wfo_vwc_brooks_corey(h, -0.1, 0.6, 0.15, 2.0) == approx(0.6)


def test_feddes_land_process_295_361():
    L = orc.lib()
    assert L.wfo_feddes_h3(-3.0, -6.0, 5.787037037037037e-9) == -6.0
    assert L.wfo_feddes_h3(-3.0, -6.0, 3.472222222222222e-8) == approx(-4.5)
    assert L.wfo_feddes_h3(-3.0, -6.0, 8.680555555555556e-8) == approx(-3.0)
    f = lambda h, a: L.wfo_rwu_reduction_feddes(h, -0.1, -1.0, -3.0, -150.0, a)
    for a, want in ((0.0, (0.0, 1.4 / 1.47, 1.0, 4 / 9, 0.0)), (0.5, (0.0, 1.4 / 1.47, 1.0, 1.0, 1.0))):
        for h, w in zip((-160.0, -10.0, -1.5, -0.5, -0.05), want):
            assert f(h, a) == pytest.approx(w, rel=RT, abs=1e-15)


def test_soil_temperature_and_infiltration_reduction_land_process_363_392():
    L = orc.lib()
    assert L.wfo_soil_temperature(274.15, 2.0, 274.65) == approx(275.15)
    assert L.wfo_infiltration_reduction_factor(273.25, 0.3, 1, 1) == approx(0.8325096069489968)
    assert L.wfo_infiltration_reduction_factor(273.25, 0.3, 1, 0) == 1.0


def test_soil_evaporation_land_process_394_464():
    L = orc.lib()
    a = (3.49537037037037e-9, 0.00123, 0.1)
    assert L.wfo_soil_evaporation_unsaturated_store(*a, 0, 0.3, 0.241) == 0.0
    assert L.wfo_soil_evaporation_unsaturated_store(*a, 1, 0.3, 0.241) == approx(5.946480713078224e-11)
    assert L.wfo_soil_evaporation_unsaturated_store(*a, 2, 0.3, 0.241) == approx(1.783944213923467e-10)
    # pinned NEGATIVE value: no clamp at zero
    assert L.wfo_soil_evaporation_saturated_store(1.4467592592592592e-9, 0, 0.1, 0.3,
                                                  0.32205961644649506, 86400.0) == approx(-7.455083714039237e-7)
    assert L.wfo_soil_evaporation_saturated_store(1.4467592592592592e-9, 2, 0.1, 0.3,
                                                  0.32205961644649506, 86400.0) == 0.0


def test_actual_infiltration_soil_path_land_process_466_501():
    out = orc.call_out("wfo_actual_infiltration_soil_path", 2, 1.883101851851852e-8,
                       1.883101851851852e-8, 0.1, 2.645787037037037e-6, 5.787037037037037e-8, 0.9)
    assert out[0] == approx(1.6947916666666665e-8)
    assert out[1] == approx(1.883101851851852e-9)
    assert orc.call_out("wfo_actual_infiltration_soil_path", 2, 1.883101851851852e-8, 0.0, 0.1,
                        2.645787037037037e-6, 5.787037037037037e-8, 0.9) == (0.0, 0.0)


def test_scurve_utils_66_77():
    L = orc.lib()
    out = L.wfo_scurve(2.0, 0.0, 3.0, 2.5)
    assert out == approx(0.3325863502664285)
    f = math.pi
    assert f * L.wfo_scurve(2.0, 0.0 + math.log(f) / 2.5, f * 3.0, 2.5) == approx(out)


def test_julia_numeric_helpers():
    L = orc.lib()
    # cld(x, y) = round((x - mod(x, -y))/y): exact ceil-division on floats
    assert L.wfo_cld(0.0, 2e-4) == 0.0
    assert L.wfo_cld(1e-9, 2e-4) == 1.0
    assert L.wfo_cld(2e-4, 2e-4) == 1.0
    assert L.wfo_cld(4.1e-4, 2e-4) == 3.0
    # round(v; sigdigits = 12)
    assert L.wfo_round_sigdigits12(2.0000000000003) == 2.0
    assert L.wfo_round_sigdigits12(2.00000000001) == 2.00000000001
    assert L.wfo_round_sigdigits12(0.0) == 0.0
    assert L.wfo_round_sigdigits12(0.123456789012345) == 0.123456789012


# ---------------------------------------------------------------------------------------------
# struct-level tests (Wflow/test/soil.jl), driven through a 1-cell OracleModel
# ---------------------------------------------------------------------------------------------
def one_cell(N, fields, cfg=None, ints=None):
    c = dict(n=1, nriv=0, N=N, gash=1, has_lai=0, snow=0, glacier=0, kv_profile=0)
    c.update(cfg or {})
    f = {}
    for k, v in fields.items():
        a = np.array(v, dtype=np.float64)
        f[k] = a.reshape(1, -1) if a.size > 1 else a.reshape(1)
    for k, v in (ints or {}).items():
        f[k] = np.array(v, dtype=np.int64).reshape(1)
    empty = dict(order=np.zeros(0, np.int64), up_ptr=np.zeros(1, np.int64), up_idx=np.zeros(0, np.int64),
                 order_of_subdomains=[], order_subdomain=[], subdomain_indices=[])
    land = dict(order=np.array([1]), up_ptr=np.array([0, 0]), up_idx=np.zeros(0, np.int64),
                order_of_subdomains=[np.array([1])], order_subdomain=[np.array([1])],
                subdomain_indices=[np.array([1])])
    return orc.OracleModel(c, f, land, empty)


def test_update_bc_soil_model_soil_1_64():
    m = one_cell(1, dict(soil_fraction=[0.2397957498236932],
                         potential_evaporation=[6.712962769799762e-9],
                         canopy_potevap=[5.103222828877092e-9],
                         interception_rate=[3.408447494157563e-10],
                         canopy_gap_fraction=[0.3487189230509198], water_fraction=[0.0],
                         river_fraction=[0.0], runoff_land=[0.0], runoff_river=[0.0],
                         runoff_water_flux_surface=[2.8572021597728237e-9]))
    m.sweep("update_bc_soil_model")
    assert m.f["potential_transpiration"][0] == approx(4.762378079461335e-9)
    assert m.f["potential_soilevaporation"][0] == approx(1.6097399409226706e-9)
    assert m.f["soil_water_flux_surface"][0] == approx(2.8572021597728237e-9)


def test_unsaturated_zone_flow_soil_66_106():
    nan = np.nan
    m = one_cell(6, dict(
        unsaturated_layer_thickness=[0.05, 0.005081648613929929, nan, nan, nan, nan],
        unsaturated_layer_depth=[0.0012855527211118947, 0.00020814868098590806, 0.0, 0.0, 0.0, 0.0],
        theta_s=[0.4414711594581604], theta_r=[0.08942600339651108], kv_0=[4.21465379220468e-6],
        hydraulic_conductivity_scale_parameter=[3.3079576678574085],
        vertical_hydraulic_conductivity_factor=[1.0] * 6,
        brooks_corey_exponent=[9.121646881103516, 9.247220993041992, 9.514554023742676,
                               9.675407409667969, 9.831438064575195, 9.716856956481934],
        infiltration=[2.635886044866978e-10]), ints=dict(n_unsatlayers=[2]))
    m.sweep("unsaturated_zone_flow", 86400.0)
    np.testing.assert_allclose(m.f["unsaturated_layer_depth"][0],
                               [0.0013083267609636298, 0.00020814799974448514, 0, 0, 0, 0], rtol=RT)
    assert m.f["transfer"][0] == approx(8.065015493937412e-15)


def test_soil_evaporation_soil_108_137():
    nan = np.nan
    m = one_cell(6, dict(
        potential_soilevaporation=[3.2407407407407403e-8],
        unsaturated_layer_thickness=[0.05, 0.021472680450878443, nan, nan, nan, nan],
        unsaturated_layer_depth=[0.001537249298366254, 4.8213268138994254e-11, 0.0, 0.0, 0.0, 0.0],
        water_table_depth=[0.07147268045087844], theta_s=[0.44], theta_r=[0.09], theta_fc=[0.275],
        actual_layer_thickness=[0.05, 0.1, 0.05, 0.2, 0.8, 0.8],
        drainable_water_depth=[0.32113323174500624]), ints=dict(n_unsatlayers=[2]))
    m.sweep("soil_evaporation", 86400.0)
    assert m.f["soil_evaporation_saturated_zone"][0] == 0
    assert m.f["soil_evaporation"][0] == approx(2.846757959937507e-9)
    assert m.f["drainable_water_depth"][0] == approx(0.32113323174500624)


def test_transpiration_soil_139_190():
    nan = np.nan
    m = one_cell(4, dict(
        h3_high=[-4.0], h3_low=[-10.0], potential_transpiration=[5.965093586654767e-10],
        water_table_depth=[0.10689587841733061],
        unsaturated_layer_thickness=[0.1, 0.006895878417330607, nan, nan],
        unsaturated_layer_depth=[0.010932797715287601, 0.000862043215499364, 0.0, 0.0],
        rooting_depth=[0.453],
        rootfraction=[0.22075055187637968, 0.6622516556291391, 0.11699779249448124, 0.0],
        actual_layer_thickness=[0.1, 0.3, 0.2, nan],
        cumulative_layer_depth=[0.0, 0.1, 0.4, 0.6, nan],
        brooks_corey_exponent=[9.53970437651816, 10.007558316712927, 10.603868189606647,
                               10.662998826419395],
        theta_s=[0.4790319800376892], theta_r=[0.17089612782001495], air_entry_pressure=[-0.1],
        h1=[0.0], h2=[-1.0], h3=[-10.0], h4=[-160.0], alpha_h1=[1.0],
        wet_root_distribution_parameter=[-500000.0], drainable_water_depth=[0.07240310797113221]),
        ints=dict(n_unsatlayers=[2]))
    m.sweep("transpiration", 86400.0)
    assert m.f["actual_evaporation_unsaturated_store"][0] == approx(5.965093586654767e-10)
    assert m.f["actual_evaporation_saturated_zone"][0] == pytest.approx(0.0, abs=1e-18)
    assert m.f["drainable_water_depth"][0] == approx(0.07240310797113221)
    assert m.f["transpiration"][0] == approx(5.965093586654767e-10)


def test_capillary_flux_soil_192_228():
    m = one_cell(6, dict(
        rooting_depth=[0.38410000000000005], kv_0=[2.335691087962963e-5],
        hydraulic_conductivity_scale_parameter=[1.29274],
        vertical_hydraulic_conductivity_factor=[1.0] * 6, water_table_depth=[1.2663358900000001],
        unsaturated_layer_thickness=[0.05, 0.1, 0.05, 0.2, 0.8, 0.06633589243],
        unsaturated_layer_depth=[0.008874129508377954, 0.018187210520563293, 0.00854476050597162,
                                 0.017279406870498257, 0.14778493107547983, 0.01251769175241036],
        actual_evaporation_unsaturated_store=[6.120061149641203e-9],
        unsaturated_store_capacity=[0.3131741519821792],
        drainable_water_depth=[0.16902603585525694], cap_hmax=[2.0], cap_n=[2.0],
        theta_s=[0.4868114888668], theta_r=[0.0711537748575]), ints=dict(n_unsatlayers=[6]))
    m.sweep("capillary_flux", 86400.0)
    assert m.f["actual_capillary_flux"][0] == approx(8.235506588899334e-10)


def test_update_soil_water_storage_soil_230_309():
    m = one_cell(4, dict(
        runoff=[0.0], water_table_depth=[1.2445135404970034],
        unsaturated_layer_thickness=[0.1, 0.3, 0.8, 0.044513540497003304],
        unsaturated_layer_depth=[0.014408928105784874, 0.01946108750081377, 0.12591094235311145,
                                 0.007223489814154271],
        actual_layer_thickness=[0.1, 0.3, 0.8, 0.8], theta_s=[0.4831694066524056],
        theta_r=[0.12372369319200516], theta_fc=[0.28599992944511926], rooting_depth=[0.38],
        cumulative_layer_depth=[0.0, 0.1, 0.4, 1.2, 2.0], soil_thickness=[2.0],
        soil_water_capacity=[0.7188914269208908], saturation_excess_water=[0.0],
        infiltration_excess=[0.0], runoff_land=[0.0], actual_open_water_evaporation_land=[0.0],
        ssf_exfiltwater_average=[0.0]),
        ints=dict(n_unsatlayers=[4], number_of_layers=[4]))
    m.update_soil_water_storage(86400.0)
    f = m.f
    assert f["runoff"][0] == pytest.approx(0.0, abs=1e-18)
    assert f["unsaturated_store_capacity"][0] == approx(0.28033060970129986)
    assert f["saturated_water_depth"][0] == approx(0.27155636944572653)
    assert f["drainable_water_depth"][0] == approx(0.14895887025738955)
    assert f["exfiltration_saturated_water"][0] == 0.0
    np.testing.assert_allclose(f["volumetric_water_content"][0],
                               [0.2678129742498539, 0.1885939848613844, 0.2811123711333945,
                                0.4721985172668562], rtol=RT)
    np.testing.assert_allclose(f["relative_volumetric_water_content"][0],
                               [55.42837989378741, 39.03268341595556, 58.18091279434587,
                                97.72939072000436], rtol=RT)
    assert f["volumetric_water_content_root_zone"][0] == approx(0.20944108733203426)
    assert f["relative_volumetric_water_content_root_zone"][0] == approx(43.34734038380604)
    assert f["total_soil_water_storage"][0] == approx(0.43856081721959095)


def test_water_table_change_utils_241_277():
    m = one_cell(5, dict(
        unsaturated_layer_depth=[0.1, 0.125, 0.15, 0.17500000000000002, 0.2],
        unsaturated_layer_thickness=[0.11, 0.145, 0.17, 0.20500000000000002, 0.24],
        theta_s=[0.98], theta_r=[0.02]), ints=dict(n_unsatlayers=[5]))
    out = (C.c_double * 2)()
    orc.lib().wfo_water_table_change(m.h, -1.1574074074074073e-5, 0.43, 0, 86400.0, out)
    assert out[0] == approx(-2.3255813953488373) and out[1] == 0.0
    orc.lib().wfo_water_table_change(m.h, 5.787037037037037e-7, 0.43, 0, 86400.0, out)
    assert out[0] == approx(0.4243119266055048) and out[1] == 0.0


# ---------------------------------------------------------------------------------------------
# routing (Wflow/test/routing_process.jl)
# ---------------------------------------------------------------------------------------------
def test_kinematic_wave_routing_process_1_20():
    it = C.c_int64()
    out = (C.c_double * 2)()
    orc.lib().wfo_kinematic_wave(1.104e-6, 0.0, 1.142e-6, 2.586, 600.0, 1061.375, out, C.byref(it))
    assert out[0] == approx(1.09308660753423e-6)
    assert out[1] == approx(0.0006852061693892164)
    assert it.value == 2  # 2 Newton iterations for this case
    orc.lib().wfo_kinematic_wave(0.0, 0.0, 0.0, 2.586, 600.0, 1061.375, out, C.byref(it))
    assert (out[0], out[1]) == (0.0, 0.0)


def test_ssf_celerity_routing_process_22_37():
    L = orc.lib()
    assert L.wfo_ssf_celerity(0.3, 0.00586, 0.274, 0.0002795374658372667, 1.8001038115471601,
                              0.0, 0) == approx(3.4838105601686665e-6)
    assert L.wfo_ssf_celerity(0.3, 0.00586, 0.274, 0.0002795374658372667, 1.8001038115471601,
                              0.2, 1) == approx(4.170921791220723e-6)


def test_kw_ssf_newton_raphson_routing_process_39_48():
    assert orc.lib().wfo_kw_ssf_newton_raphson(0.008738344907407408, 77.774, 0.0001418287037037037,
                                               86400.0, 1103.816) == approx(0.01090947420564454)


SSF_SHARED = dict(actual_layer_thickness=[0.1, 0.3, 0.8, 0.8],
                  cumulative_layer_depth=[0.0, 0.1, 0.4, 1.2, 2.0],
                  theta_s=[0.48642662167549133], theta_r=[0.11939866840839386],
                  theta_fc=[0.28219206182657536], kh_0=[0.002379589787235966],
                  hydraulic_conductivity_scale_parameter=[1.0141291422769427])


def _ssf(m, *a):
    out = (C.c_double * 4)()
    orc.lib().wfo_kinematic_wave_ssf(m.h, *a, 0, out)
    return tuple(out)


def test_kinematic_wave_ssf_routing_process_100_253():
    slope, sy, d, dt = 0.4522336721420288, 0.20423455984891598, 2.0, 86400.0
    dx, dw = 1117.0150713112287, 517.495693771673
    m = one_cell(4, dict(SSF_SHARED,
                         unsaturated_layer_thickness=[0.1, 0.3, 0.11983408703759733, np.nan],
                         unsaturated_layer_depth=[0.0001909439890049523, 0.01627933934181815,
                                                  0.019508197676020186, 0.0],
                         water_table_depth=[0.5198340870375974]),
                 ints=dict(n_unsatlayers=[3], number_of_layers=[4]))
    # first case exercises the inner sub-iteration loop (|dzi| > 0.1 m)
    q, zi, exf, nf = _ssf(m, 0.0, 0.30038365579798126, 0.0005198340870375973,
                          0.005618827458801466, slope, sy, d, dt, dx, dw, 0.0009215296489248933)
    assert q == approx(0.23130576097772237)
    assert zi == approx(0.1656875455413981)
    assert exf == pytest.approx(0.0, abs=1e-18)
    assert nf == approx(-3.904277181728481e-7)
    # q_in + q_prev == 0 and q_net <= 0
    assert _ssf(m, 0.0, 0.0, 0.0, 0.0, slope, sy, d, dt, dx, dw, 0.0009215296489248933) == (
        0.0, d, 0.0, 0.0)
    # exponential-constant profile
    m2 = one_cell(4, dict(SSF_SHARED, z_exp=[0.2],
                          unsaturated_layer_thickness=[0.1, 0.3, 0.348312461531486, np.nan],
                          unsaturated_layer_depth=[0.0001909439890049523, 0.01627933934181815,
                                                   0.058425012193036086, 0.0],
                          water_table_depth=[0.748312461531486]),
                  cfg=dict(kv_profile=1), ints=dict(n_unsatlayers=[3], number_of_layers=[4]))
    q, zi, exf, nf = _ssf(m2, 0.0, 0.627032986563781, 0.748312461531486, 0.008957349820205272,
                          slope, sy, d, dt, dx, dw, 0.0017762382461437296)
    assert q == approx(0.5171363105669935)
    assert zi == approx(1.1202203724020348)
    assert exf == pytest.approx(0.0, abs=1e-18)
    assert nf == approx(-8.791255611224121e-7)


def test_kinwave_river_update_routing_process_290_383():
    """2-node river graph (1 -> 2) with a SIMPLE RESERVOIR on node 1 that has a negative external
    inflow (abstraction): stable time step (type-7 quantile), one kinwave_river_update!, river
    q / h / storage / q_cumulative and the reservoir's water level, storage, outflow and actual
    evaporation. As in the reference's unit test node 1 stays in node 2's upstream list."""
    L = orc.lib()
    q = np.array([0.5499295110293246, 3.0005238507869465])
    alpha = np.array([2.544585458995107, 2.5507721996678145])
    length = np.array([1059.8125, 951.96875])
    width = np.array([94.73094177246094, 94.73094177246094])
    work = np.zeros(2)
    dt = L.wfo_stable_timestep_surface(q.ctypes.data, alpha.ctypes.data, length.ctypes.data, 2,
                                       0.05, work.ctypes.data)
    assert dt == approx(994.6119029285007)

    class G:  # outneighbors: 1 -> 2
        down = np.array([2, 0])
    river = dict(graph=G, order=np.array([1, 2]), up_ptr=np.array([0, 0, 1]), up_idx=np.array([1]),
                 order_of_subdomains=[np.array([1])], order_subdomain=[np.array([1, 2])],
                 subdomain_indices=[np.array([1, 2])])

    class G1:
        down = np.array([0, 0])
    land = dict(graph=G1, order=np.array([1, 2]), up_ptr=np.zeros(3, np.int64), up_idx=np.zeros(0, np.int64),
                order_of_subdomains=[np.array([1, 2])], order_subdomain=[np.array([1]), np.array([2])],
                subdomain_indices=[np.array([1]), np.array([2])])
    f = dict(riv_q=q, riv_alpha=alpha, riv_flow_length=length, riv_flow_width=width,
             riv_qlat=np.zeros(2), riv_qin=np.zeros(2), riv_storage=np.zeros(2), riv_h=np.zeros(2),
             riv_external_inflow=np.zeros(2), riv_abstraction=np.zeros(2),
             river_land_indices=np.array([0, 1]), reservoir_river_indices=np.array([0]),
             res_external_inflow=[-0.1], res_precipitation=[2.0833332436504178e-10],
             res_evaporation=[5.324074170655674e-9], res_inflow_overland=[0.0],
             res_inflow_subsurface=[0.02735223635554672], res_area=[1.498462875e6],
             res_outflow_curve_type=[4.0], res_maximum_release=[24.007999420166016],
             res_demand=[3.000999927520752], res_target_minimum_fraction=[0.07482631504535675],
             res_target_full_fraction=[0.7536525130271912], res_maximum_storage=[6.2e7],
             res_waterlevel=[29.656373296565086], res_storage=[4.443897439204416e7],
             res_outflow_obs=[np.nan])
    f = {k: np.asarray(v, dtype=np.int64 if k.endswith("indices") else np.float64)
         for k, v in f.items()}
    m = orc.OracleModel(dict(n=2, nriv=2, nres=1, N=1), f, land, river)
    L.wfo_kinwave_river_update(m.h, dt)
    assert m.f["riv_q"] == approx(np.array([0.37903337592185243, 3.1969698861305855]))
    assert m.f["riv_h"] == approx(np.array([0.01500830539624011, 0.05407828963124342]))
    assert m.f["riv_storage"] == approx(np.array([1506.7893805755937, 4876.828625285123]))
    assert m.f["riv_q_cumulative"] == approx(np.array([376.9911072990474, 3179.744302049454]))
    assert m.f["res_waterlevel"][0] == approx(29.654579645252387)
    assert m.f["res_storage"][0] == approx(4.443628667214138e7)
    assert m.f["res_outflow"][0] == approx(3.0009999145314317)
    assert m.f["res_actevap_cumulative"][0] == approx(5.295387542208319e-6)


def test_local_inertial_river_with_reservoir_routing_process_385_528():
    """3-node graph 1 -> 2 -> 3, a simple reservoir on node 2: stable time step, channel flow at
    the edges, reservoir boundary condition, water depth and storage -- each sub-step of the
    local-inertial update as the reference's unit test checks it."""
    L = orc.lib()

    class G:
        down = np.array([2, 3, 0])
    river = dict(graph=G, order=np.array([1, 2, 3]), up_ptr=np.array([0, 0, 1, 2]),
                 up_idx=np.array([1, 2]), order_of_subdomains=[np.array([1])],
                 order_subdomain=[np.array([1, 2, 3])], subdomain_indices=[np.array([1, 2, 3])])

    class G1:
        down = np.array([0, 0, 0])
    land = dict(graph=G1, order=np.array([1, 2, 3]), up_ptr=np.zeros(4, np.int64),
                up_idx=np.zeros(0, np.int64), order_of_subdomains=[np.array([1])],
                order_subdomain=[np.array([1, 2, 3])], subdomain_indices=[np.array([1, 2, 3])])
    w = 94.73094177246094
    f = dict(riv_inwater=[0.012816561479797707, 0.01544689027914484, -0.0004760637654763434],
             riv_external_inflow=np.zeros(3), riv_abstraction=np.zeros(3),
             li_zb=[315.1000061035156, 314.3999938964844, 278.8000183105469],
             li_zb_at_edge=[315.1000061035156, 314.3999938964844, 0.0],
             li_mannings_n_sq_at_edge=[0.0008999999597668652, 0.0008999999597668652, 0.0],
             li_flow_length_at_edge=[831.578125, 1005.890625, 1.0],
             li_flow_width_at_edge=[w, w, w], li_ghost_h=np.zeros(3),
             riv_flow_width=[w, w, w], riv_flow_length=[603.34375, 1059.8125, 951.96875],
             riv_h=[0.04484241735240722, 0.0, 0.07939389691400389],
             riv_storage=[2562.9827873416416, 0.0, 7159.812778536053],
             riv_q=[0.534560186001325, 0.0, 0.0],
             river_land_indices=np.array([0, 1, 2]), reservoir_river_indices=np.array([1]),
             res_external_inflow=[0.0], res_inflow_overland=[0.0],
             res_inflow_subsurface=[0.04279912663469156],
             res_precipitation=[2.0833332436504186e-10], res_evaporation=[5.324074170655674e-9],
             res_area=[1.498462875e6], res_outflow_curve_type=[4.0],
             res_maximum_release=[24.007999420166016], res_demand=[3.000999927520752],
             res_target_minimum_fraction=[0.07482631504535675],
             res_target_full_fraction=[0.7536525130271912], res_maximum_storage=[6.2e7],
             res_waterlevel=[29.6558203236325], res_storage=[4.443814578263413e7],
             res_outflow_obs=[np.nan])
    f = {k: np.asarray(v, dtype=np.int64 if k.endswith("indices") else np.float64)
         for k, v in f.items()}
    m = orc.OracleModel(dict(n=3, nriv=3, nres=1, N=1, river_routing=1, li_froude_limit=1,
                             li_ghost_nodes=0, li_alpha=1.0, li_h_thresh=0.001), f, land, river)
    dt = L.wfo_li_stable_timestep(m.h)
    assert dt == approx(909.829412320351)
    L.wfo_li_update_river_channel_flow(m.h, dt)
    assert m.f["li_zs_at_edge"][0] == approx(315.14484852086804)
    assert m.f["li_water_depth_at_edge"][0] == approx(0.04484241735241312)
    assert m.f["riv_q"][:2] == approx(np.array([0.534558444239785, 0.0]))
    assert m.f["riv_q_cumulative"][:2] == approx(np.array([486.35699517356477, 0.0]))
    L.wfo_li_update_bc_reservoir_model(m.h, dt)
    assert m.f["riv_q"][1] == approx(3.0009999145276134)
    assert m.f["riv_q"][1] == m.f["res_outflow"][0]
    assert m.f["riv_q_cumulative"][1] == approx(2730.397988608082)
    assert m.f["riv_q_cumulative"][1] == m.f["res_outflow_cumulative"][0]
    assert m.f["res_storage"][0] == approx(4.443593370702217e7)
    assert m.f["res_waterlevel"][0] == approx(29.654344093791327)
    assert m.f["res_inflow_cumulative"][0] == approx(525.2968994074305)
    assert m.f["res_actevap_cumulative"][0] == approx(4.843999273837613e-6)
    L.wfo_li_update_water_depth_and_storage(m.h, dt)
    assert m.f["riv_storage"] == approx(np.array([2088.286676767209, 0.0, 9889.777630328164]))
    assert m.f["riv_h"] == approx(np.array([0.03653704705843743, 0.0, 0.10966599406601261]))


def test_local_inertial_river_with_floodplain_routing_process_530_690():
    """3-node graph 1 -> 2 -> 3 with a 1-D floodplain (6-level profile): stable time step, channel
    flow, floodplain flow (profile interpolation, opposite-direction rule), river and floodplain
    water depth and storage (bankfull redistribution) -- each sub-step as the reference checks it."""
    L = orc.lib()

    class G:
        down = np.array([2, 3, 0])
    river = dict(graph=G, order=np.array([1, 2, 3]), up_ptr=np.array([0, 0, 1, 2]),
                 up_idx=np.array([1, 2]), order_of_subdomains=[np.array([1])],
                 order_subdomain=[np.array([1, 2, 3])], subdomain_indices=[np.array([1, 2, 3])])

    class G1:
        down = np.array([0, 0, 0])
    land = dict(graph=G1, order=np.array([1, 2, 3]), up_ptr=np.zeros(4, np.int64),
                up_idx=np.zeros(0, np.int64), order_of_subdomains=[np.array([1])],
                order_subdomain=[np.array([1, 2, 3])], subdomain_indices=[np.array([1, 2, 3])])
    w = 149.17837524414062
    zb = 165.10784077644348
    # FloodPlainProfile tables as in the reference test: rows = levels, columns = nodes
    storage = np.array([[0.0, 0.0, 0.0], [77602.0, 281141.0, 111313.0],
                        [189506.0, 609512.0, 256357.0], [346960.0, 981178.0, 420515.0],
                        [526346.0, 1.39783e6, 602100.0], [747343.0, 1.87239e6, 814606.0]])
    width = np.array([[w, w, w],
                      [341.5768913342503, 666.778728923476, 758.7636596016615],
                      [492.562310866575, 778.7935519733185, 988.6905953775695],
                      [693.0574965612104, 881.4757828423199, 1118.9809351368624],
                      [789.5944979367263, 988.1614230127849, 1237.7718606880392],
                      [972.7515818431912, 1125.5153603853994, 1448.5444669293854]])
    flow_area = np.array([[0.0, 0.0, 0.0],
                          [170.78844566712516, 333.389364461738, 379.38182980083076],
                          [417.0696011004127, 722.7861404483972, 873.7271274896154],
                          [763.5983493810179, 1163.5240318695571, 1433.2175950580468],
                          [1158.395598349381, 1657.6047433759495, 2052.103525402066],
                          [1644.7713892709767, 2220.3624235686493, 2776.375758866759]])
    perimeter = np.array([[192.3985160901097, 517.6003536793354, 609.5852843575209],
                          [193.3985160901097, 518.6003536793354, 610.5852843575209],
                          [345.38393562243436, 631.6151767291778, 841.5122201334289],
                          [546.8791213170698, 735.2974075981792, 972.8025598927218],
                          [644.4161226925856, 842.9830477686443, 1092.5934854438985],
                          [828.5732065990505, 981.3369851412588, 1304.3660916852448]])
    f = dict(riv_inwater=[0.001214603164946035, 0.0069799723162954465, 0.00022016439899281862],
             riv_external_inflow=np.zeros(3), riv_abstraction=np.zeros(3),
             li_zb=[zb, zb, zb], li_zb_at_edge=[zb, zb, 0.0],
             li_mannings_n_sq_at_edge=[0.0008999999597668652, 0.0008999999597668652, 0.0],
             li_flow_length_at_edge=[648.828125, 568.34375, 1.0],
             li_flow_width_at_edge=[w, w, w], li_ghost_h=np.zeros(3),
             li_bankfull_storage=[107921.00118967971, 200292.17449130033, 69688.46493603183],
             li_bankfull_depth=[1.592156171798706] * 3,
             riv_flow_width=[w, w, w], riv_flow_length=[454.375, 843.28125, 293.40625],
             riv_h=[1.8817912224982847, 1.8197068314233162, 1.7619620034455687],
             riv_storage=[127553.31189184493, 228917.89427333255, 77120.8437153619],
             riv_q=[137.1750567107639, 133.74797838974442, 0.0],
             fp_q=[3.306660222819796, 6.14041420989245, 0.0],
             fp_h=[0.2896350506995787, 0.22755065962461016, 0.16980583164686275],
             fp_storage=[25320.207706612186, 99321.92021301284, 30370.81429688439],
             fp_mannings_n_sq_at_edge=[0.005184, 0.005184, 0.0],
             fp_zb_at_edge=[166.6999969482422, 166.6999969482422, 0.0],
             fp_profile_storage=storage.T, fp_profile_width=width.T,
             fp_profile_flow_area=flow_area.T, fp_profile_wetted_perimeter=perimeter.T,
             river_land_indices=np.array([0, 1, 2]))
    f = {k: np.ascontiguousarray(v, dtype=np.int64 if k.endswith("indices") else np.float64)
         for k, v in f.items()}
    m = orc.OracleModel(dict(n=3, nriv=3, nres=0, N=1, river_routing=1, li_froude_limit=1,
                             li_ghost_nodes=0, li_alpha=0.7, li_h_thresh=0.001,
                             fp_depth=[0.0, 0.5, 1.0, 1.5, 2.0, 2.5]), f, land, river)
    dt = L.wfo_li_stable_timestep(m.h)
    L.wfo_li_update_river_channel_flow(m.h, dt)
    assert dt == approx(49.40931052556788)
    assert m.f["li_zs_at_edge"][:2] == approx(np.array([166.98963199894177, 166.92754760786679]))
    assert m.f["li_water_depth_at_edge"][:2] == approx(np.array([1.8817912224982933, 1.8197068314233036]))
    assert m.f["riv_q"][:2] == approx(np.array([137.1827776559179, 133.7538757670657]))
    assert m.f["riv_q_cumulative"][:2] == approx(np.array([6778.106459961183, 6608.686781773178]))
    L.wfo_li_update_floodplain_flow(m.h, dt)
    assert m.f["fp_water_depth_at_edge"][:2] == approx(np.array([0.2896350506995873, 0.22755065962459753]))
    assert m.f["fp_q"][:2] == approx(np.array([3.3074672215578524, 6.1421232455587536]))
    assert m.f["fp_q_cumulative"][:2] == approx(np.array([163.41967500308917, 303.4780747261213]))
    L.wfo_li_update_bc_reservoir_model(m.h, dt)
    L.wfo_li_update_water_depth_and_storage(m.h, dt)
    assert m.f["riv_storage"] == approx(np.array([120775.26544458869, 229087.6588271402, 83729.54137530623]))
    assert m.f["riv_h"] == approx(np.array([1.7817948514048583, 1.8210563184054411, 1.9129493838748897]))
    L.wfo_li_update_floodplain_water_depth_and_storage(m.h, dt)
    assert m.f["riv_storage"] == approx(np.array([124521.73491986065, 228924.54042985922, 78479.82701991481]))
    assert m.f["riv_h"] == approx(np.array([1.8370664336896243, 1.8197596628390198, 1.793010379352563]))
    assert m.f["fp_storage"] == approx(np.array([21410.318556337137, 99344.98021057082, 35924.006727001935]))
    assert m.f["fp_h"] == approx(np.array([0.2449102618909183, 0.22760349104031377, 0.20085420755385677]))


def test_kinwave_river_with_floodplain_routing_process_858_987():
    """2-node river (1 -> 2) with the kinematic wave's 1-D floodplain: channel-floodplain exchange,
    kinwave_river_update! with the exchange as lateral inflow, Manning flow capacity of the
    floodplain and its accucapacityflux -- step by step as the reference's unit test."""
    L = orc.lib()
    q = np.array([296.52948301601174, 192.8313119108856])
    alpha = np.array([34.0789466790827, 34.0789466790827])
    length = np.array([750.953125, 851.8125])
    width = np.array([229.91920471191406, 229.91920471191406])
    work = np.zeros(2)
    dt = L.wfo_stable_timestep_surface(q.ctypes.data, alpha.ctypes.data, length.ctypes.data, 2,
                                       0.05, work.ctypes.data)
    assert dt == approx(1602.881460805217)

    class G:
        down = np.array([2, 0])
    river = dict(graph=G, order=np.array([1, 2]), up_ptr=np.array([0, 0, 1]), up_idx=np.array([1]),
                 order_of_subdomains=[np.array([1])], order_subdomain=[np.array([1, 2])],
                 subdomain_indices=[np.array([1, 2])])

    class G1:
        down = np.array([0, 0])
    land = dict(graph=G1, order=np.array([1, 2]), up_ptr=np.zeros(3, np.int64), up_idx=np.zeros(0, np.int64),
                order_of_subdomains=[np.array([1, 2])], order_subdomain=[np.array([1]), np.array([2])],
                subdomain_indices=[np.array([1]), np.array([2])])
    storage = np.array([[0.0, 0.0], [86329.2726379633, 97924.02628183365],
                        [172659.2726379633, 207762.02628183365], [258989.2726379633, 386716.02628183365],
                        [369518.2726379633, 605009.0262818336], [724648.2726379633, 869843.0262818336]])
    w0 = 229.91920471191406
    pwidth = np.array([[w0, w0], [w0, w0], [229.9211418821914, 257.89243524836746],
                       [229.9211418821914, 420.1722796977034], [294.369904912507, 512.5376770122533],
                       [945.8113647239966, 621.8128989654414]])
    flow_area = np.array([[0.0, 0.0], [114.95960235595703, 114.95960235595703],
                          [229.9201732970527, 243.90581998014076], [344.8807442381484, 453.99195982899244],
                          [492.06569669440194, 710.260798335119], [964.9713790564002, 1021.1672478178398]])
    perimeter = np.array([[0.0, 0.0], [1.0, 1.0], [2.00193717027733, 29.9732305364534],
                          [3.00193717027733, 193.25307498578934], [68.45070020059296, 286.61847230033925],
                          [720.8921600120825, 396.8936942535273]])
    f = dict(riv_q=q, riv_alpha=alpha, riv_flow_length=length, riv_flow_width=width,
             riv_qlat=[3.9326064945614956e-5, 2.1713344971219756e-7], riv_qin=np.zeros(2),
             riv_h=[4.509741437854894, 3.4835130995322534],
             riv_storage=[778645.3962305915, 682239.2566234164],
             riv_external_inflow=np.zeros(2), riv_abstraction=np.zeros(2),
             li_bankfull_depth=[2.1051321029663086, 2.1051321029663086],
             li_bankfull_storage=[363469.04651181493, 412286.02275520907],
             fp_q=[1.2861909826521447, 1.9846650910027395], fp_h=np.zeros(2), fp_storage=np.zeros(2),
             fp_mannings_n=[0.072, 0.072], fp_slope=[1.0e-5, 1.0e-5],
             fp_profile_storage=storage.T, fp_profile_width=pwidth.T,
             fp_profile_flow_area=flow_area.T, fp_profile_wetted_perimeter=perimeter.T,
             river_land_indices=np.array([0, 1]))
    f = {k: np.ascontiguousarray(v, dtype=np.int64 if k.endswith("indices") else np.float64)
         for k, v in f.items()}
    m = orc.OracleModel(dict(n=2, nriv=2, nres=0, N=1, fp_depth=[0.0, 0.5, 1.0, 1.5, 2.0, 2.5]),
                        f, land, river)
    L.wfo_river_channel_floodplain_exchange(m.h, dt)
    assert m.f["riv_h"] == approx(np.array([4.509741437854894, 3.4835130995322534]))
    assert m.f["riv_storage"] == approx(np.array([778645.3962305915, 682239.2566234164]))
    assert m.f["fp_h"] == approx(np.array([2.0642836103410205, 1.1737631111525133]))
    assert m.f["fp_storage"] == approx(np.array([58760.1445203583, 40074.01437791623]))
    L.wfo_kinwave_river_update(m.h, dt)
    assert m.f["riv_h"] == approx(np.array([2.872002930358695, 3.1138609375883397]))
    assert m.f["riv_storage"] == approx(np.array([495875.84798393055, 609843.6005807515]))
    assert m.f["riv_q"] == approx(np.array([139.7837242183345, 159.9486203252721]))
    assert m.f["riv_q_cumulative"] == approx(np.array([224056.7400718776, 256378.67820075116]))
    L.wfo_update_floodplain_model(m.h, dt)
    assert m.f["fp_storage"] == approx(np.array([52745.30941770246, 41649.967280840756]))
    assert m.f["fp_q"] == approx(np.array([3.752513987924126, 2.7693140810933157]))
    assert m.f["fp_q_cumulative"] == approx(np.array([6014.835102655834, 4438.882199731312]))


def test_local_inertial_long_channel_macdonald_routing_process_1189_1318():
    """The reference's analytical check of the local-inertial river scheme: MacDonald (1997), a
    1000 m channel (dx = 5 m, width 10 m, Manning n 0.03, 20 m3/s in at the upper end, the
    analytical depth as downstream boundary); bed levels by integrating the analytical slope (QuadGK
    in the reference, scipy.integrate.quad here: hence the reference's own `isapprox` tolerance,
    1.5e-8). Run to a steady state (|change of the mean depth| <= 1e-12) with
    stable_timestep / update_river_channel_flow! / update_water_depth_and_storage!; the mean
    absolute error against the analytical profile is the reference's number."""
    from scipy.integrate import quad
    G = 9.80665
    L_, dx = 1000.0, 5.0
    n = int(L_ / dx)
    h = lambda x: np.cbrt(4 / G) * (1.0 + 0.5 * math.exp(-16.0 * (x / L_ - 0.5) ** 2))
    h_acc = lambda x: -np.cbrt(4 / G) * 16.0 / L_ * (x / L_ - 0.5) * math.exp(-16 * (x / L_ - 0.5) ** 2)
    slope = lambda x: ((1.0 - 4.0 / (G * h(x) ** 3)) * h_acc(x)
                       + 0.36 * (2 * h(x) + 10.0) ** (4.0 / 3.0) / ((10.0 * h(x)) ** (10.0 / 3.0)))
    x = np.arange(dx, L_ + dx / 2, dx)
    h_a = np.array([h(xi) for xi in x])
    zb = np.array([quad(slope, xi, L_, epsabs=0, epsrel=1e-12)[0] for xi in x])
    down = np.concatenate([np.arange(2, n + 1), [0]])      # node i -> i + 1; node n has no edge

    class G_:
        pass
    G_.down = down
    river = dict(graph=G_, order=np.arange(1, n + 1), up_ptr=np.concatenate([[0, 0], np.arange(1, n)]),
                 up_idx=np.arange(1, n), order_of_subdomains=[np.array([1])],
                 order_subdomain=[np.arange(1, n + 1)], subdomain_indices=[np.arange(1, n + 1)])

    class G1:
        down = np.zeros(1, dtype=np.int64)
    land = dict(graph=G1, order=np.array([1]), up_ptr=np.zeros(2, np.int64), up_idx=np.zeros(0, np.int64),
                order_of_subdomains=[np.array([1])], order_subdomain=[np.array([1])],
                subdomain_indices=[np.array([1])])
    d = np.where(down > 0, down - 1, np.arange(n))
    f = dict(riv_inwater=np.zeros(n), riv_external_inflow=np.zeros(n), riv_abstraction=np.zeros(n),
             li_zb=zb, li_zb_at_edge=np.maximum(zb, zb[d]),
             li_mannings_n_sq_at_edge=np.full(n, 0.03 * 0.03), li_flow_length_at_edge=np.full(n, dx),
             li_flow_width_at_edge=np.full(n, 10.0), li_ghost_h=np.zeros(n),
             riv_flow_width=np.full(n, 10.0), riv_flow_length=np.full(n, dx),
             riv_h=np.concatenate([np.zeros(n - 1), [h_a[-1]]]), riv_storage=np.zeros(n),
             riv_q=np.zeros(n), river_land_indices=np.zeros(n, dtype=np.int64))
    m = orc.OracleModel(dict(n=1, nriv=n, N=1, river_routing=1, li_froude_limit=1, li_ghost_nodes=0,
                             li_alpha=0.7, li_h_thresh=1e-3), f, land, river)
    L = orc.lib()
    for it in range(200000):
        m.f["riv_inwater"][0] = 20.0
        h0 = m.f["riv_h"].mean()
        dt = L.wfo_li_stable_timestep(m.h)
        L.wfo_li_update_river_channel_flow(m.h, dt)
        L.wfo_li_update_water_depth_and_storage(m.h, dt)
        m.f["riv_h"][-1] = h_a[-1]   # node n is not in active_n: its depth is the boundary condition
        if abs(h0 - m.f["riv_h"].mean()) <= 1e-12:
            break
    else:
        raise AssertionError("no steady state")
    assert np.mean(np.abs(m.f["riv_h"] - h_a)) == pytest.approx(0.01873574206931199, rel=1.5e-8)


def _lil_nets(nriv_down):
    """a river chain given by `down` (1-based, 0 = pit) and a land domain without drainage"""
    nr = len(nriv_down)

    class G:
        down = np.array(nriv_down)
    up_ptr, up_idx = [0], []
    for v in range(1, nr + 1):
        up_idx += [u + 1 for u in range(nr) if nriv_down[u] == v]
        up_ptr.append(len(up_idx))
    river = dict(graph=G, order=np.arange(1, nr + 1), up_ptr=np.array(up_ptr),
                 up_idx=np.array(up_idx, dtype=np.int64), order_of_subdomains=[np.array([1])],
                 order_subdomain=[np.arange(1, nr + 1)], subdomain_indices=[np.arange(1, nr + 1)])

    class G1:
        down = np.array([0, 0, 0])
    land = dict(graph=G1, order=np.array([1, 2, 3]), up_ptr=np.zeros(4, np.int64),
                up_idx=np.zeros(0, np.int64), order_of_subdomains=[np.array([1])],
                order_subdomain=[np.array([1, 2, 3])], subdomain_indices=[np.array([1, 2, 3])])
    return land, river


def _as_fields(f):
    ints = ("indices", "edge_x_up", "edge_x_down", "edge_y_up", "edge_y_down")
    return {k: np.asarray(v, dtype=np.int64 if k.endswith(ints) else np.float64) for k, v in f.items()}


def _reservoir_model(nriv_down, **res):
    """one reservoir on river node 1 of a chain given by `down`"""
    land, river = _lil_nets(nriv_down)
    nr = len(nriv_down)
    f = dict(river_land_indices=np.arange(nr), reservoir_river_indices=[0],
             riv_q=np.zeros(nr), riv_qin=np.zeros(nr), res_outflow_obs=[np.nan],
             res_external_inflow=[0.0], res_inflow_overland=[0.0], res_inflow_subsurface=[0.0],
             res_threshold=[0.0], res_rating_curve_coefficient=[0.0], res_rating_curve_exponent=[0.0],
             res_maximum_storage=[np.nan], res_maximum_release=[np.nan], res_demand=[np.nan],
             res_target_minimum_fraction=[np.nan], res_target_full_fraction=[np.nan])
    f.update(res)
    return orc.OracleModel(dict(n=3, nriv=nr, nres=1, N=1), _as_fields(f), land, river)


def test_update_reservoir_simple_reservoir_1_58():
    """update_reservoir_model!(res, 1, 100.0, dt), simple reservoir, without and with an observed
    outflow (test/reservoir.jl:1-58)."""
    L = orc.lib()
    dt = 86400.0
    kw = dict(res_precipitation=[4.861111111111111e-8], res_evaporation=[1.736111111111111e-8],
              res_demand=[52.523], res_maximum_release=[420.184], res_maximum_storage=[25_000_000.0],
              res_area=[1885665.353626924], res_target_full_fraction=[0.8],
              res_target_minimum_fraction=[0.2425554726620697], res_outflow_curve_type=[4.0],
              res_storage=[1.925e7], res_waterlevel=[10.208598234556407])
    m = _reservoir_model([2, 0], **kw)
    L.wfo_update_reservoir_model(m.h, 0, 100.0, dt)
    assert m.f["res_outflow"][0] == approx(91.3783714867453)
    assert m.f["res_storage"][0] == approx(2.0e7)
    assert m.f["res_actevap_cumulative"][0] == approx(0.0014999999999999998)
    assert m.f["res_outflow_cumulative"][0] == m.f["res_outflow"][0] * dt
    m = _reservoir_model([2, 0], **dict(kw, res_outflow_obs=[80.0]))
    L.wfo_update_reservoir_model(m.h, 0, 100.0, dt)
    assert m.f["res_outflow"][0] == approx(80.0)
    assert m.f["res_storage"][0] == approx(2.0983091296454795e7)


def test_update_reservoir_modified_puls_reservoir_60_115():
    """Modified Puls approach (outflow_curve_type = 3), linear storage curve (test/reservoir.jl:60-115)."""
    L = orc.lib()
    area = 180510409.0
    m = _reservoir_model([2, 0], res_precipitation=[2.3148148148148148e-7],
                         res_evaporation=[3.7037037037037036e-8], res_area=[area], res_threshold=[0.0],
                         res_outflow_curve_type=[3.0], res_rating_curve_coefficient=[0.22],
                         res_rating_curve_exponent=[2.0], res_waterlevel=[18.5],
                         res_storage=[area * 18.5])          # initialize_storage, linear
    L.wfo_update_reservoir_model(m.h, 0, 2500.0, 86400.0)
    assert m.f["res_outflow"][0] == approx(85.14292808113598)
    assert m.f["res_storage"][0] == approx(3.55111879238499e9)
    assert m.f["res_waterlevel"][0] == approx(19.672653848925634)
    assert m.f["res_storage"][0] / area == approx(19.672653848925634)   # waterlevel(linear, ...)
    assert m.f["res_actevap_cumulative"][0] == approx(0.0032)


def test_update_reservoir_model_at_node_reservoir_117_221():
    """update_reservoir_model!(reservoir, river variables, network, v, dt): limited abstraction,
    overland / subsurface / river inflow, observed outflow and the simple rule; the outflow becomes
    qin of the downstream node (test/reservoir.jl:117-221)."""
    L = orc.lib()
    m = _reservoir_model([2, 0], res_external_inflow=[-1.0], res_inflow_overland=[0.02],
                         res_inflow_subsurface=[0.04], res_precipitation=[5.787037037037037e-9],
                         res_evaporation=[1.1574074074074074e-9], res_outflow_curve_type=[4.0],
                         res_area=[6.0e4], res_waterlevel=[1.0], res_storage=[4.5e7],
                         res_outflow=[3.0], res_outflow_obs=[1.0], riv_q=[0.04, 0.04])
    L.wfo_update_reservoir_at_node(m.h, 0, 1000.0)
    assert m.f["riv_qin"][1] == approx(1.0)
    assert m.f["res_actual_external_abstraction_cumulative"][0] == approx(1e3)
    assert m.f["res_storage"][0] == approx(4.4998100277777776e7)
    assert m.f["res_waterlevel"][0] == approx(0.9683379629629354)
    assert m.f["res_outflow"][0] == approx(1.0)
    m = _reservoir_model([2, 0], res_external_inflow=[-1.0], res_inflow_overland=[0.0],
                         res_inflow_subsurface=[0.00041241066945499203],
                         res_precipitation=[8.101853611016715e-10], res_evaporation=[6.134257548385196e-9],
                         res_outflow_curve_type=[4.0], res_area=[9.069779e4], res_maximum_release=[1.74],
                         res_demand=[0.2175], res_target_minimum_fraction=[0.358469158],
                         res_target_full_fraction=[0.83492106199], res_maximum_storage=[3.3e7],
                         res_waterlevel=[3.0266425035195113], res_storage=[2.7450978618928656e7],
                         riv_q=[0.00012002923701686638, 0.21747539140212965])
    L.wfo_update_reservoir_at_node(m.h, 0, 1000.0)
    assert m.f["riv_qin"][1] == approx(0.21749985206208133)
    assert m.f["res_actual_external_abstraction_cumulative"][0] == approx(1000.0)
    assert m.f["res_storage"][0] == approx(2.744976116863499e7)
    assert m.f["res_waterlevel"][0] == approx(3.013219350720886)
    assert m.f["res_outflow"][0] == approx(0.21749985206208133)


def test_local_inertial_flow_general_area_routing_process_1320_1344():
    """local_inertial_flow(q0, zs0, zs1, hf, A, R, ...), the general-area method of the river
    (surface_process.jl:88-115)."""
    q = orc.lib().wfo_local_inertial_flow(0.0004713562869434079, 206.10117949049967, 201.9003737619653,
                                          0.0011733869840497846, 0.04970535373017763,
                                          0.0011733219820725962, 533.453125, 0.0008999999597668652, 1,
                                          89.29563868855615)
    assert q == approx(0.005331926324969742)


def test_local_inertial_flow_rectangular_routing_process_1346_1373():
    """local_inertial_flow(theta, q0, qd, qu, ...), the rectangular-area method of the overland
    flow (surface_process.jl:123-159; de Almeida et al. 2012)."""
    q = orc.lib().wfo_local_inertial_flow_rect(
        1.0, 0.0001769756305800402, 0.0, 0.0, 601.4761297394623, 601.4730243288751,
        0.00310727852479431, 620.6649135473787, 926.602742473319, 0.1773345894316103, 1,
        49.774905820268735)
    assert q == approx(0.00017992597962222483)


def test_update_directional_flow_routing_process_989_1048():
    """2-D local-inertial overland flow in the x direction at edge 2 of three cells
    | land | land | river |: stable time step of the land cells and the edge flow. The reference's
    flow vectors hold n + 1 entries (the last one, the edge to outside, is 0): n here."""
    L = orc.lib()
    land, river = _lil_nets([0])
    f = dict(li_land_runoff=[0.0, 0.0, 0.003001456821567986],
             li_land_ywidth_at_edge=[926.6857061478484, 869.7426481339323, 812.7995901200163],
             li_land_zx_max_at_edge=[257.3280029296875, 232.67100524902344, 232.67100524902344],
             li_land_mannings_n_sq_at_edge=[0.24167056670421605, 0.2883451232664811, 0.3928782408368683],
             li_land_z=[257.3280029296875, 227.5050048828125, 232.67100524902344],
             li_land_qx0=[0.0, -3.6332616217117395, -0.7525806207906618],
             li_land_qx=[0.0, -3.63341089804407, -0.7526187151790501],
             olf_h=[0.0, 1.3754708010382453, 0.11735139800699446],
             olf_storage=[0.0, 783157.9568615163, 237954.47204911432],
             li_land_x_length=[614.4202561305977] * 3, li_land_y_length=[926.6857061478484] * 3,
             edge_x_up=[1, 2, 3], edge_x_down=[3, 0, 1], edge_y_up=[3, 3, 3], edge_y_down=[3, 3, 3],
             land_river_indices=[-1, -1, 0], river_land_indices=[2],
             riv_flow_length=[1000.0], riv_h=[0.0])
    m = orc.OracleModel(dict(n=3, nriv=1, N=1, river_routing=1, land_routing=1, li_land_alpha=0.7,
                             li_land_theta=1.0, li_land_h_thresh=1e-3, li_land_froude_limit=1),
                        _as_fields(f), land, river)
    dt = L.wfo_lil_stable_timestep(m.h)
    assert dt == approx(117.10556654947368)
    m.f["li_land_qx0"][:] = m.f["li_land_qx"]
    L.wfo_lil_update_directional_flow(m.h, 1, dt, 1)
    assert m.f["li_land_qx"][1] == approx(-3.633493490896127)
    assert m.f["li_land_qx_cumulative"][1] == approx(-3.633493490896127 * dt)


def test_local_inertial_update_water_depth_routing_process_1050_1187():
    """River and land water depth and storage of the coupled local-inertial overland / river flow
    (subgrid channel), cells | land | land | river |. The reference's single river node has one
    entering edge (q = 56.69) and one leaving edge (q = 53.71); edge i is the edge leaving node i
    here, so the entering edge belongs to an upstream node and the leaving one ends in the ghost
    node of the pit."""
    L = orc.lib()
    land, river = _lil_nets([2, 0])
    f = dict(li_land_qx=[0.0, -3.63341089804407, -0.7526187151790501],
             li_land_qy=[0.0, -0.7369647824685662, 0.0],
             olf_storage=[0.0, 783157.9568615163, 237954.47204911432],
             olf_h=[0.0, 1.3754708010382453, 0.11735139800699446],
             li_land_runoff=[0.0, 0.0, 0.003001456821567986],
             li_land_x_length=[614.4202561305977] * 3, li_land_y_length=[926.6857061478484] * 3,
             edge_x_up=[1, 2, 3], edge_x_down=[3, 0, 1], edge_y_up=[3, 3, 3], edge_y_down=[3, 3, 3],
             land_river_indices=[-1, -1, 1], river_land_indices=[0, 2],
             li_bankfull_storage=[1.0, 171137.5821314017], li_bankfull_depth=[1.0, 1.3683528900146484],
             riv_q=[56.685647296907476, 53.70963118023338],
             riv_h=[0.0, 1.485704288021643], riv_storage=[0.0, 185814.5230442402],
             riv_flow_width=[113.88611602783203] * 2, riv_flow_length=[1098.1875] * 2,
             riv_external_inflow=[0.0, 0.0], riv_abstraction=[0.0, 0.0], li_ghost_h=[0.0, 0.0])
    m = orc.OracleModel(dict(n=3, nriv=2, N=1, river_routing=1, land_routing=1, li_ghost_nodes=1,
                             li_alpha=0.7, li_land_alpha=0.7, li_land_theta=1.0,
                             li_land_h_thresh=1e-3), _as_fields(f), land, river)
    dt = L.wfo_li_stable_timestep(m.h)
    assert dt == approx(201.394687315008)
    sc = L.wfo_lil_compute_river_storage_change(m.h, 2, dt)
    assert sc == approx(19.782071832453088)
    river_h, land_h, river_storage = 1.4857390315391559, 0.11738614152450744, 185818.86835722585
    out = (orc.C.c_double * 3)()
    L.wfo_lil_compute_water_depths(m.h, m.f["olf_storage"][2] + sc, 1, 2, out)
    assert list(out) == approx([river_h, land_h, river_storage])
    L.wfo_lil_update_river_and_land_storage_and_depth(m.h, 2, dt)
    assert m.f["riv_h"][1] == approx(river_h)
    assert m.f["olf_h"][2] == approx(land_h)
    assert m.f["riv_storage"][1] == approx(river_storage)
    assert m.f["riv_actual_external_abstraction_cumulative"][1] == 0.0
    assert L.wfo_lil_compute_land_storage_change(m.h, 1, dt) == approx(880.1704436259577)
    L.wfo_lil_update_land_storage_and_depth(m.h, 1, dt)
    assert m.f["olf_storage"][1] == approx(784038.1273051423)
    assert m.f["olf_h"][1] == approx(1.3770166561681556)


def test_edge_connectivity_network_136_153():
    """EdgeConnectivity of a masked 3 x 3 raster: DIRS / NEIGHBORS order (network.jl:3,
    connectivity.jl:61-66), n + 1 where there is no active neighbour."""
    mask = np.array([[1, 1, 0], [1, 1, 1], [0, 1, 1]], dtype=bool)
    idx, rev = nw.active_indices(mask)
    e = nw.edge_connectivity(idx, 3, 3)
    n = len(idx)
    # column-major numbering: (1,1)=1 (2,1)=2 (1,2)=3 (2,2)=4 (3,2)=5 (2,3)=6 (3,3)=7
    assert n == 7
    assert list(e["ind_x_up"]) == [2, 8, 4, 5, 8, 7, 8]      # CartesianIndex(1, 0)
    assert list(e["ind_x_down"]) == [8, 1, 8, 3, 4, 8, 6]    # CartesianIndex(-1, 0)
    assert list(e["ind_y_up"]) == [3, 4, 8, 6, 7, 8, 8]      # CartesianIndex(0, 1)
    assert list(e["ind_y_down"]) == [8, 8, 1, 2, 8, 4, 5]    # CartesianIndex(0, -1)


def test_accucapacityflux_routing_process_255_288():
    """PCRaster accucapacity examples on a 6-node graph (lateral snow transport's engine)."""
    L = orc.lib()
    down = np.array([4, 5, 5, 6, 6, 0], dtype=np.int64) - 1
    order = np.arange(6, dtype=np.int64)
    dt = 86400.0

    def run(material, capacity):
        mat = np.array(material, dtype=np.float64)
        cap = np.array(capacity, dtype=np.float64)
        flux = np.zeros(6)
        L.wfo_accucapacityflux(flux.ctypes.data, mat.ctypes.data, order.ctypes.data,
                               down.ctypes.data, 6, cap.ctypes.data, dt)
        return flux, mat
    flux, mat = run(dt * np.array([0.5, 2.0, 2.0, 0.5, 2.0, 0.5]), np.full(6, 1.5))
    assert np.array_equal(mat, dt * np.array([0.0, 0.5, 0.5, 0.0, 3.5, 1.5]))
    assert np.array_equal(flux, np.array([0.5, 1.5, 1.5, 1, 1.5, 1.5]))
    flux, mat = run(dt * np.full(6, 10.0), [2, 30, 30, 2, 30, 2])
    assert np.array_equal(mat, dt * np.array([8.0, 0.0, 0.0, 10.0, 0.0, 40.0]))
    assert np.array_equal(flux, np.array([2, 10, 10, 2, 30, 2], dtype=np.float64))


# ---------------------------------------------------------------------------------------------
# indexing artefacts (Wflow/test/subdomains.jl:48-87)
# ---------------------------------------------------------------------------------------------
def test_streamorder_subbasins_subdomains_48_87():
    down = np.zeros(16, dtype=np.int64)
    for a, b in ((1, 3), (2, 3), (3, 4), (4, 5), (5, 6), (7, 9), (8, 9), (9, 4), (10, 12),
                 (11, 12), (12, 16), (13, 15), (14, 15), (15, 16), (16, 5)):
        down[a - 1] = b
    g = nw.DiGraph1(down)
    toposort = nw.topological_sort_by_dfs(g)
    strord = nw.stream_order(g, toposort)
    subbas = nw.subbasins(g, strord, toposort, 2)
    subbas_fill = nw.fillnodata_upstream(g, toposort, subbas, 0)
    graph_subbas = nw.graph_from_nodes(g, subbas, subbas_fill)
    toposort_subbas = nw.topological_sort_by_dfs(graph_subbas)
    dist = nw.distances_undirected(graph_subbas, int(toposort_subbas[-1]))
    max_dist = max(int(dist.max()), 1)
    order = nw.subbasins_order(graph_subbas, int(toposort_subbas[-1]), max_dist)
    assert strord.tolist() == [1, 1, 2, 3, 4, 4, 1, 1, 2, 1, 1, 2, 1, 1, 2, 3]
    assert strord[toposort[-1] - 1] == 4
    assert subbas.tolist() == [0, 0, 5, 6, 0, 7, 0, 0, 4, 0, 0, 2, 0, 0, 1, 3]
    assert subbas_fill.tolist() == [5, 5, 5, 6, 7, 7, 4, 4, 4, 2, 2, 2, 1, 1, 1, 3]
    assert toposort_subbas.tolist() == [5, 4, 6, 2, 1, 3, 7]
    assert max_dist == 2
    assert [list(o) for o in order] == [[1, 2, 4, 5], [3, 6], [7]]


def test_subdomains_single_thread_subdomains_25_31():
    down = np.array([2, 3, 0, 3], dtype=np.int64)
    g = nw.DiGraph1(down)
    order = nw.topological_sort_by_dfs(g)
    so = nw.stream_order(g, order)
    a, b, c = nw.kinwave_set_subdomains(g, order, np.array([3]), so, 5, 1)
    assert [x.tolist() for x in a] == [[1]]
    assert b[0].tolist() == [1, 2, 3, 4]
    assert c[0].tolist() == order.tolist()
