// vertical.cu -- the SBM vertical land-surface update as sm_100a elementwise kernels.
//
// One thread per land slot; the ~30 sweeps of the reference collapse into
//   land_hydrology_kernel     update_land_hydrology_model!, first half (sbm.jl:82-132):
//                             interception, snow, glacier, open water, soil boundary conditions,
//                             diagnostics, infiltration, unsaturated-zone flow
//   unsat_engine_kernel       the long Brooks-Corey sub-iteration loops of unsatzone_flow_layer
//                             (soil_process.jl:51-92) of the cells that have them, see "the
//                             unsaturated-zone engine" below
//   soil_column_kernel        second half: soil evaporation, transpiration, actual infiltration,
//                             capillary flux, leakage, recharge, AET   soil/soil.jl:814-1209
//   soil_water_storage_kernel update_soil_water_storage!               soil/soil.jl:1294-1392
//   total_water_storage_kernel update_total_water_storage!             sbm.jl:143-182
// Consecutive threads touch consecutive doubles of every SoA array, so each warp load/store is a
// fully used 256-byte transaction; the layered state lives in registers (template N) between
// sub-processes. The two dense kernels ask for ALL their inputs at kernel entry (cp.async into
// lane-private shared-memory rows, "input staging" below): at 4-5 register-limited CTAs per SM a
// warp keeps too few ordinary loads in flight, and a cell's ~50 loads cost ~12 DRAM round trips. Arithmetic order follows the reference expression by expression (no FMA
// contraction: -fmad=false) so results match the Julia code to the last bits that libm
// differences allow. All reference paths are under /root/reference/Wflow/src.
//
// The unsaturated-zone engine. unsatzone_flow_layer runs `its = cld(remainder, 2e-4 m)` explicit
// sub-iterations, each a dependent div -> log -> exp chain. After a few wet days the trip count
// is 1 for 80 % of the (cell, layer) calls and 64-400 for 1.5 % of them -- and those 1.5 % hold
// 40 % of all iterations. With one cell per lane a warp runs as long as its wettest cell, and a
// CTA as long as its wettest warp. The engine therefore takes the long loops OUT of the
// per-cell kernel: land_hydrology_kernel runs loops of up to `inline_iters` trips in line and
// SUSPENDS a cell at the first longer one (the loop's operands go to a per-cell scratch record,
// the cell id to a list bucketed by log2(its)); unsat_engine_kernel takes every suspended cell
// through its remaining layers, one lane per cell, 32 cells of the same bucket per warp, longest
// buckets first, with trips that track the power instead of re-evaluating log and exp
// (unsatzone_flow_iterate_fast); soil_column_kernel then finishes every cell from arrays that
// are reference-visible outputs anyway.
#include <cstdio>
#include "device_math.cuh"
#include "kernels.cuh"
#include "model.cuh"
#include "soil_storage.cuh"

namespace wfb {

namespace {

constexpr int kTile = WFB_V_TILE;  // cells per tile = threads per CTA of land_hydrology_kernel

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// v[n] for a run-time n without local memory (the arrays live in registers)
template <int N>
__device__ __forceinline__ double pick(const double (&v)[N], int n) {
  double r = v[0];
#pragma unroll
  for (int k = 1; k < N; ++k)
    if (k == n) r = v[k];
  return r;
}

// The vertical conductivity profile of one cell (soil.jl:213-244). k[n] is the layer's
// vertical_hydraulic_conductivity_factor, for the layered profiles already multiplied with the
// layer's kv (the reference only ever uses the product, in this order).
template <int N>
struct KvCol {
  double k[N];
  double kv_0, f, zx;  // zx: z_exp (exponential_constant) or z_layered (layered_exponential)
  int nk;              // nlayers_kv - 1 (layered_exponential)
};
template <int N>
__device__ __forceinline__ KvCol<N> load_kvcol(const DevFields& f, const KCfg& c, int i) {
  KvCol<N> kv;
  const int ns = c.ns, prof = c.kv_profile;
  kv.kv_0 = prof < 2 ? __ldg(f.kv_0 + i) : 0.0;
  kv.f = prof != 2 ? __ldg(f.hydraulic_conductivity_scale_parameter + i) : 0.0;
  kv.zx = prof == 1 ? __ldg(f.z_exp + i) : (prof == 3 ? __ldg(f.z_layered + i) : 0.0);
  kv.nk = prof == 3 ? f.nlayers_kv[i] - 1 : 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double fac = __ldg(f.vertical_hydraulic_conductivity_factor + k * ns + i);
    kv.k[k] = prof >= 2 ? fac * __ldg(f.kv + k * ns + i) : fac;
  }
  return kv;
}
// hydraulic_conductivity_at_depth for layer n (0-based), all four profiles   utils.jl:727-789
template <int N>
__device__ __forceinline__ double kv_at_depth(int profile, const KvCol<N>& kv, double kn, double z) {
  if (profile == 0) return kn * kv.kv_0 * exp(-kv.f * z);
  if (profile == 1) return kn * kv.kv_0 * exp(-kv.f * (z < kv.zx ? z : kv.zx));
  if (profile == 2) return kn;
  if (z < kv.zx) return kn;
  return pick<N>(kv.k, kv.nk) * exp(-kv.f * (z - kv.zx));
}

// unsatzone_flow_layer                                           soil/soil_process.jl:51-92
// split in two: the part every cell executes once per layer (`setup`: the transfer of water
// above saturation and the number `its` of explicit sub-iterations), and the sub-iteration
// loop itself, whose trip count is data dependent (0 for a dry layer, > 100 for a wet one).
struct UnsatTask {
  double usd, sum_ast, kv_it, l_sat, c;
  int its;
};
__device__ __forceinline__ UnsatTask unsatzone_flow_setup(double usd, double kv_z, double l_sat,
                                                          double c, double dt, const Divisor& ddt) {
  UnsatTask t;
  t.l_sat = l_sat; t.c = c; t.kv_it = 0.0; t.its = 0;
  if (usd <= 0.0) { t.usd = 0.0; t.sum_ast = 0.0; return t; }
  const double st_sat = jmax(0.0, usd - l_sat);
  const double st = kv_z * bounded_power(fdiv(usd, l_sat), c);
  const double sum_ast = jmin(st, st_sat / ddt);
  usd -= sum_ast * dt;
  const double remainder = jmin((st - sum_ast) * dt, usd);
  const int its = (int)jcld_pos(remainder, 2e-4);
  t.usd = usd; t.sum_ast = sum_ast; t.its = its;
  t.kv_it = its > 0 ? fdiv(kv_z, (double)its) : 0.0;
  return t;
}
__device__ __forceinline__ void unsatzone_flow_iterate(UnsatTask& t, double dt, const Divisor& ddt) {
  double usd = t.usd, sum_ast = t.sum_ast;
  const Divisor dl(t.l_sat);
  for (int k = 0; k < t.its; ++k) {
    const double st = t.kv_it * bounded_power(usd / dl, t.c);
    const double st_max = usd / ddt;
    if (st < st_max) { usd -= st * dt; sum_ast += st; }
    else { usd = 0.0; sum_ast += st_max; break; }
  }
  t.usd = usd; t.sum_ast = sum_ast;
}

// The same loop with a short dependency chain, for the LONG loops of the engine. The engine lasts
// as long as the trips of its wettest cell, one after the other: ~1600 trips x (div -> log -> exp
// with libdevice: ~600 cycles; with a tracked power but the reference's comparisons as branches
// on the chain: ~150 cycles). A trip drains the fraction e = st dt / usd of the layer. With
// x = usd / l_sat and r = x^(c-1):   e = (kv_it dt / l_sat) r,   x' = x (1 - e),
// r' = r (1 - e)^(c-1), hence   e' = e (1 - e)^(c-1).
// For e <= 1/64 the power is its binomial series in e with coefficients that depend on c only
// (twelve terms, Estrin form, explicit FMAs), so e itself is tracked from trip to trip: five
// dependent operations and no branch. usd' = usd - usd e and st = usd e / dt follow beside the
// chain. c > 1 makes e fall from trip to trip, so e <= 1/64 (which also implies the reference's
// st < usd / dt) needs testing only when tracking starts. Tracking starts from the reference
// expression (bounded_power) and restarts from it every kResync trips, so e stays within ~1e-14 of
// the reference's value; an oversaturated layer (x > 1), an underflowing power or a large e take
// the reference path trip by trip. Largest difference to the reference loop over random tasks:
// wflowb200_selftest_math measures it on the device (tests/test_gpu_parity.py).
constexpr int kResync = 64;
__device__ __forceinline__ void unsatzone_flow_iterate_fast(UnsatTask& t, double dt,
                                                            const Divisor& ddt) {
  double usd = t.usd, sum_ast = t.sum_ast;
  const Divisor dl(t.l_sat);
  const double a = t.kv_it * dt / dl;
  const double inv_dt = 1.0 / ddt;
  const double m = t.c - 1.0;
  // (1 - e)^m = sum_j b_j e^j,  b_0 = 1,  b_j = -b_{j-1} (m - j + 1) / j
  double b[12];
  b[0] = 1.0;
#pragma unroll
  for (int j = 1; j < 12; ++j) b[j] = b[j - 1] * (-(m - (double)(j - 1)) / (double)j);
  const double e_max = m > 0.0 ? jmin(1.0 / 64.0, 0.175 / m) : -1.0;  // twelve terms suffice
  double e = 0.0;
  int fresh = 0;  // trips for which the tracked e may still be used
  for (int k = 0; k < t.its; ++k) {   // ONE flat loop: the lanes of a warp are at different trips
    if (fresh > 0) {
      const double sd = usd * e;
      sum_ast += usd * (e * inv_dt);
      usd -= sd;
      --fresh;
    } else {
      const double x = usd / dl;
      const double p = bounded_power(x, t.c);  // the reference expression
      const double st = t.kv_it * p;
      const double st_max = usd / ddt;
      if (st < st_max) { usd -= st * dt; sum_ast += st; }
      else { usd = 0.0; sum_ast += st_max; break; }
      if (x <= 1.0 && p > 1.0e-280) {
        e = a * (p / x);
        if (e <= e_max) fresh = kResync;
      }
    }
    if (fresh > 0) {  // e of the next trip
      const double e2 = e * e, e4 = e2 * e2, e8 = e4 * e4;
      const double p0 = fma(e, b[1], b[0]), p1 = fma(e, b[3], b[2]), p2 = fma(e, b[5], b[4]);
      const double p3 = fma(e, b[7], b[6]), p4 = fma(e, b[9], b[8]), p5 = fma(e, b[11], b[10]);
      const double q0 = fma(e2, p1, p0), q1 = fma(e2, p3, p2), q2 = fma(e2, p5, p4);
      e *= fma(e8, q2, fma(e4, q1, q0));
    }
  }
  t.usd = usd; t.sum_ast = sum_ast;
}

// ---- the unsaturated-zone engine: suspended loops ------------------------------------------
constexpr int kBuckets = WFB_UNSAT_BUCKETS;
__device__ __forceinline__ int bucket_of(int its) {  // <= 4 -> 0, (4,8] -> 1 ... by log2, clamped
  const int b = 30 - __clz(its - 1);                 // its in (2^(b+1), 2^(b+2)]
  return b < 0 ? 0 : (b >= kBuckets ? kBuckets - 1 : b);
}
// record the loop of cell i (layer k) and append the cell to the list of its bucket
// (warp-aggregated: one atomic per warp and bucket)
__device__ __forceinline__ void suspend_cell(const UnsatWork& w, int i, int k,
                                             const UnsatTask& t, bool suspend) {
  const unsigned active = __ballot_sync(0xffffffffu, suspend);
  if (!suspend) return;
  w.usd[i] = t.usd; w.sum_ast[i] = t.sum_ast; w.kv_it[i] = t.kv_it; w.l_sat[i] = t.l_sat;
  w.c[i] = t.c;
  w.its_layer[i] = t.its | (k << 24);
  const int b = bucket_of(t.its);
  const unsigned peers = __match_any_sync(active, b);
  const int lane = (int)threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(w.count + b, (unsigned)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  const unsigned pos = base + (unsigned)__popc(peers & ((1u << lane) - 1u));
  w.list[((size_t)b) * (size_t)w.cap + pos] = i;
}

// warp tile j of the concatenated bucket lists, longest bucket first -> (bucket, first entry)
__device__ __forceinline__ bool tile_of(const unsigned* cnt, int j, int& b, int& first, int& n) {
  for (b = kBuckets - 1; b >= 0; --b) {
    const int tiles = ((int)cnt[b] + 31) >> 5;
    if (j < tiles) { first = j << 5; n = (int)cnt[b]; return true; }
    j -= tiles;
  }
  return false;
}

// rwu_reduction_feddes                                         soil/soil_process.jl:183-200
__device__ __forceinline__ double rwu_reduction_feddes(double h, double h1, double h2, double h3,
                                                       double h4, double alpha_h1) {
  if (h < h4) return 0.0;
  if (h < h3) return fdiv(h - h4, h3 - h4);
  if (alpha_h1 == 0.0) {
    if (h < h2) return 1.0;
    if (h < h1) return fdiv(h1 - h, h1 - h2);
    return 0.0;
  }
  return 1.0;
}

// The unsaturated-zone layers k0.. of one cell (soil.jl:764-804): `flow` enters layer k0, z is
// the depth of the bottom of layer k0 - 1. Returns true when the cell finished all its layers
// (uld[] and transfer are then final); false when it was suspended at a long loop.
template <int N>
__device__ __forceinline__ bool unsat_layers(const KCfg& c, const UnsatWork& w, int i,
                                             int k0, int n_unsat, double z, double flow,
                                             double first_inflow, double (&uld)[N],
                                             const double (&ult)[N], const double (&bc)[N],
                                             const KvCol<N>& kv, double theta_e, double dt,
                                             const Divisor& ddt, bool live, double& transfer) {
  bool suspended = false;
  int k_susp = 0;
  UnsatTask t_susp;
  t_susp.usd = t_susp.sum_ast = t_susp.kv_it = 0.0; t_susp.l_sat = 1.0; t_susp.c = 1.0; t_susp.its = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k >= k0 && k < n_unsat && !suspended) {
      z = (k == 0) ? ult[0] : z + ult[k];
      const double l_sat = ult[k] * theta_e;
      const double kv_z = kv_at_depth<N>(c.kv_profile, kv, kv.k[k], z);
      const double usd = (k == 0) ? uld[k] + first_inflow * dt : uld[k] + flow * dt;
      UnsatTask t = unsatzone_flow_setup(usd, kv_z, l_sat, bc[k], dt, ddt);
      if (!live) t.its = 0;
      if (t.its > w.inline_iters) {
        suspended = true; k_susp = k; t_susp = t;
      } else {
        unsatzone_flow_iterate(t, dt, ddt);
        uld[k] = t.usd;
        flow = t.sum_ast;
      }
    }
  }
  suspend_cell(w, i, k_susp, t_susp, suspended);
  transfer = n_unsat > 0 ? flow : 0.0;
  return !suspended;
}

// ---- lane-private input staging (kernels.cuh: StageList) -------------------------------------
// Rows of the two dense kernels. The rows every configuration reads come first, the rows only
// some configurations read last, so that the default configurations allocate no row they do
// not use (N = 4: 43 rows = 5 CTAs per SM for the first half, 48 rows = 4 for the second).
extern __shared__ double stage_rows[];
namespace st1 {  // land_hydrology_kernel
enum : int { snow_storage, snow_water, ttm, cfmax, whc, tsoil, w_soil, cf_soil, rf, wf, olf_h,
             h_river, theta_s, theta_r, d_soil, swc, satwd, pathfrac, cap_soil, cap_path, kv_0,
             kv_f, layered };
template <int N> struct Rows {
  static constexpr int kvfac = layered, uld = kvfac + N, alt = uld + N, bc = alt + N, cld = bc + N,
                       core = cld + N + 1,
                       canopy_storage = core, cmax = core + 1, gap = core + 2, e_r = core + 3,
                       kv_zx = core + 4, kvlay = core + 5, all = kvlay + N;
};
}  // namespace st1
namespace st2 {  // soil_column_kernel
enum : int { d_soil, swc, satwd, zi, pot_soilevap, pot_transp, infiltration, infiltration_excess,
             wfs, f_red, pathfrac, cap_soil, cap_path, aeow_river, aeow_land, interception,
             theta_fc, rooting_depth, h1, h2, h4, alpha_h1, air_entry, h3_high, h3_low, wet_root,
             cap_hmax, cap_n, max_leakage, kv_0, kv_f, layered };
template <int N> struct Rows {
  static constexpr int kvfac = layered, alt = kvfac + N, cld = alt + N, rootf = cld + N + 1,
                       core = rootf + N,
                       kv_zx = core, khfrac = core + 1, kvlay = core + 2, all = kvlay + N;
};
}  // namespace st2
static_assert(st1::Rows<8>::all <= kMaxStage && st2::Rows<8>::all <= kMaxStage, "kMaxStage");

// A lane copies 16 bytes: lanes 0-15 bring the warp's 32 values of row r, lanes 16-31 those of
// row r + 1 (half the requests of an 8-byte copy per lane, and they bypass the L1).
__device__ __forceinline__ void stage_issue(const StageList& sl, const int tile) {
  const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
  const int half = lane >> 4, first = warp * 32 + (lane & 15) * 2;  // first of this lane's 2 cells
  const size_t src = (size_t)tile * kTile + first;
  const unsigned dst = (unsigned)__cvta_generic_to_shared(stage_rows) + (unsigned)first * 8u;
#pragma unroll 4
  for (int r = half; r < sl.rows; r += 2) {
    const double* const p = sl.p[r];
    if (p)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)r * (kTile * 8u)),
                   "l"(p + src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
// (whole warps must get here: a lane reads what other lanes of its warp asked for)
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
}
// this lane's value of row r
__device__ __forceinline__ double staged(const int r) { return stage_rows[r * kTile + (int)threadIdx.x]; }

template <int N, class Rows>
__device__ __forceinline__ KvCol<N> staged_kvcol(const DevFields& f, const KCfg& c, const int i,
                                                 const int kv_0, const int kv_f) {
  KvCol<N> kv;
  const int prof = c.kv_profile;
  kv.kv_0 = prof < 2 ? staged(kv_0) : 0.0;
  kv.f = prof != 2 ? staged(kv_f) : 0.0;
  kv.zx = (prof == 1 || prof == 3) ? staged(Rows::kv_zx) : 0.0;
  kv.nk = prof == 3 ? f.nlayers_kv[i] - 1 : 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double fac = staged(Rows::kvfac + k);
    kv.k[k] = prof >= 2 ? fac * staged(Rows::kvlay + k) : fac;
  }
  return kv;
}

template <int N>
struct SoilColumn {
  double theta_s, theta_e, d_soil, swc, satwd, zi;
  int nlayers, n_unsat;
  double uld[N], ult[N], alt[N], cld[N + 1], bc[N];
  KvCol<N> kv;
  double pot_soilevap, pot_transp, infiltration, infiltration_excess, wfs, max_infiltsoil,
      max_infiltpath, pathfrac, transfer, aeow_river, aeow_land, interception;
};

// kh_layered_profile!                                                      utils.jl:792-895
// equivalent horizontal conductivity of the saturated part of the column (layered profiles)
template <int N>
__device__ __forceinline__ double kh_layered(int profile, const SoilColumn<N>& s, const double (&kvl)[N],
                                             double fpar, double z_layered, int nlayers_kv,
                                             double khfrac) {
  const int m = s.nlayers;  // 1-based count
  if (!(s.d_soil > s.zi)) return pick<N>(kvl, m - 1) * khfrac;  // (both profiles: kv[i][m] * ratio)
  double transmissivity = 0.0;
  int n = s.n_unsat > 1 ? s.n_unsat : 1;  // 1-based layer
  if (profile == 2) {
    transmissivity += (pick<N + 1>(s.cld, n) - s.zi) * pick<N>(kvl, n - 1);
    n += 1;
    for (; n <= m; ++n) transmissivity += pick<N>(s.alt, n - 1) * pick<N>(kvl, n - 1);
  } else {
    const double zt = s.d_soil - z_layered;
    const double kvj = pick<N>(kvl, nlayers_kv - 1);
    if (s.zi >= z_layered) {
      transmissivity += kvj / fpar * (exp(-fpar * (s.zi - z_layered)) - exp(-fpar * zt));
      n = m;
    } else {
      transmissivity += (pick<N + 1>(s.cld, n) - s.zi) * pick<N>(kvl, n - 1);
    }
    n += 1;
    while (n <= m) {
      if (n > nlayers_kv) {
        transmissivity += kvj / fpar * (1.0 - exp(-fpar * zt));
        n = m;
      } else {
        transmissivity += pick<N>(s.alt, n - 1) * pick<N>(kvl, n - 1);
      }
      n += 1;
    }
  }
  return (transmissivity / (s.d_soil - s.zi)) * khfrac;
}

// update_land_hydrology_model!, second half, for one cell: soil evaporation, transpiration,
// actual infiltration, capillary flux, leakage, recharge, AET      soil/soil.jl:814-1209
template <int N>
__device__ __forceinline__ void soil_column_cell(const DevFields& f, const KCfg& c, const int i,
                                                 const double dt, const Divisor& ddt,
                                                 SoilColumn<N>& s) {
  const int ns = c.ns;
  const double theta_e = s.theta_e;
  const double theta_d = jmax(s.theta_s - staged(st2::theta_fc), 0.02);
  const double zi = s.zi;
  const int n_unsat = s.n_unsat, nlayers = s.nlayers;
  double (&uld)[N] = s.uld;
  const double (&ult)[N] = s.ult;
  const double (&alt)[N] = s.alt;
  const double (&cld)[N + 1] = s.cld;
  double drainable = (s.d_soil - zi) * theta_d;

  // ---- soil evaporation (soil.jl:814-856, soil_process.jl:247-294) --------------------------
  double soilevap_sat, soil_evaporation;
  {
    double pot = s.pot_soilevap;
    double evu;
    if (n_unsat == 0) evu = 0.0;
    else if (n_unsat == 1) evu = pot * jmin(1.0, fdiv(uld[0], zi * theta_e));
    else evu = pot * jmin(1.0, fdiv(uld[0], ult[0] * theta_e));
    evu = jmin(evu, uld[0] / ddt);
    pot -= evu;
    uld[0] = uld[0] - evu * dt;
    if (n_unsat == 0 || n_unsat == 1) {
      const double e = pot * jmin(1.0, fdiv(alt[0] - zi, alt[0]));
      soilevap_sat = jmin(e, (alt[0] - zi) * theta_d / ddt);  // deliberately not clamped at 0
    } else {
      soilevap_sat = 0.0;
    }
    soil_evaporation = evu + soilevap_sat;
    drainable -= soilevap_sat * dt;
  }
  f.soil_evaporation_saturated_zone[i] = soilevap_sat;
  f.soil_evaporation[i] = soil_evaporation;

  // ---- transpiration (soil.jl:865-975) -----------------------------------------------------
  const double pot_transp = s.pot_transp;
  const double rd = staged(st2::rooting_depth);
  const double h1 = staged(st2::h1), h2 = staged(st2::h2), h4 = staged(st2::h4);
  const double alpha_h1 = staged(st2::alpha_h1);
  const double hb = staged(st2::air_entry);
  double h3;
  {
    const double tpot_daily = fdiv(pot_transp, WFB_MM_PER_DAY);  // feddes_h3 soil_process.jl:166-176
    const double h3_high = staged(st2::h3_high), h3_low = staged(st2::h3_low);
    if (tpot_daily <= 1.0) h3 = h3_low;
    else if (tpot_daily < 5.0) h3 = h3_low + (h3_high - h3_low) * (tpot_daily - 1.0) / (5.0 - 1.0);
    else h3 = h3_high;
  }
  f.h3[i] = h3;
  double rootf[N];
#pragma unroll
  for (int k = 0; k < N; ++k) rootf[k] = staged(st2::Rows<N>::rootf + k);
  // root fraction of the lowest unsaturated layer, rescaled to its unsaturated part; the layer
  // is selected with pick() -- indexing the register arrays with n_unsat - 1 would move them
  // to local memory
  double sum_rf = 0.0, rf_lowest = 0.0;
  if (n_unsat > 0) {
    const int kl = n_unsat - 1;
    rf_lowest = pick<N>(rootf, kl);
    if (zi < rd) {
      const double rootlength = jmin(pick<N>(alt, kl), rd - pick<N + 1>(cld, kl));
      rf_lowest = rf_lowest * fdiv(pick<N>(ult, kl), rootlength);
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k < n_unsat - 1) sum_rf += rootf[k];  // the same left fold: upper layers, then the lowest
  if (n_unsat > 0) sum_rf += rf_lowest;
  double actevapustore = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k < n_unsat) {
      const double rfu = (k < n_unsat - 1) ? rootf[k] : rf_lowest;
      const double rfs = rd > 0.0 ? jmax(1.0 / sum_rf, 1.0) * rfu : 0.0;
      const double vwc = jmax(fdiv(uld[k], ult[k]), 1e-7);
      // head_brooks_corey soil_process.jl:113-130
      const double par_lambda = 2.0 / (s.bc[k] - 3.0);
      const double head = par_lambda > 0.0 ? hb / jpow(fdiv(vwc, theta_e), 1.0 / par_lambda) : hb;
      const double alpha = rwu_reduction_feddes(head, h1, h2, h3, h4, alpha_h1);
      const double availcap = jmin(1.0, jmax(0.0, fdiv(rd - cld[k], ult[k])));
      const double maxextr = uld[k] * availcap / ddt;
      const double layer = jmin(alpha * rfs * pot_transp, maxextr);
      uld[k] = uld[k] - layer * dt;
      actevapustore += layer;
    }
  }
  const double wetroots = scurve(zi, rd, 1.0, staged(st2::wet_root));
  const double alpha_sat = rwu_reduction_feddes(0.0, h1, h2, h3, h4, alpha_h1);
  const double restpottrans = pot_transp - actevapustore;
  const double ae_sat = jmin(restpottrans * wetroots * alpha_sat, drainable / ddt);
  drainable -= ae_sat * dt;
  const double transpiration = actevapustore + ae_sat;
  f.actual_evaporation_unsaturated_store[i] = actevapustore;
  f.actual_evaporation_saturated_zone[i] = ae_sat;
  f.transpiration[i] = transpiration;

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
  const double wfs = s.wfs, pathfrac = s.pathfrac;
  const double actual_infiltration = s.infiltration - excess / ddt;
  f.actual_infiltration[i] = actual_infiltration;
  f.saturation_excess_water[i] = (wfs - actual_infiltration) - s.infiltration_excess;
  double actinf_soil, actinf_path;
  if (actual_infiltration > 0.0) {  // soil_process.jl:297-323
    const Divisor dsum(s.max_infiltpath + s.max_infiltsoil);
    actinf_soil = actual_infiltration * s.max_infiltsoil / dsum;
    actinf_path = actual_infiltration * s.max_infiltpath / dsum;
  } else {
    actinf_soil = 0.0; actinf_path = 0.0;
  }
  f.actual_infiltration_soil[i] = actinf_soil;
  f.actual_infiltration_compacted_soil[i] = actinf_path;
  f.excess_water_soil[i] = jmax(wfs * (1.0 - pathfrac) - actinf_soil, 0.0);
  f.excess_water_compacted_soil[i] = jmax(wfs * pathfrac - actinf_path, 0.0);

  // ---- recompute stores, capillary flux, leakage, recharge (soil.jl:1194-1209) --------------
  double ustore_depth = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k < nlayers) ustore_depth += uld[k];
  const double ustore_cap = s.swc - s.satwd - ustore_depth;
  f.unsaturated_store_depth[i] = ustore_depth;
  f.unsaturated_store_capacity[i] = ustore_cap;
  double act_capflux = 0.0;
  if (n_unsat > 0) {  // capillary_flux! soil.jl:1050-1111
    const double ksat = kv_at_depth<N>(c.kv_profile, s.kv, pick<N>(s.kv.k, n_unsat - 1), zi);
    double mc = jmin(ksat, actevapustore);
    mc = jmin(mc, ustore_cap / ddt);
    mc = jmin(mc, drainable / ddt);
    const double maxcapflux = jmax(0.0, mc);
    double capflux = 0.0;
    if (zi > rd) {
      const double hmax = staged(st2::cap_hmax);
      capflux = maxcapflux * jpow(1.0 - fdiv(jmin(zi, hmax), hmax), staged(st2::cap_n));
    }
    double net = capflux;
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
      if (k < n_unsat) {
        const double toadd = jmin(net, jmax((ult[k] * theta_e - uld[k]) / ddt, 0.0));
        uld[k] = uld[k] + toadd * dt;
        net -= toadd;
        act_capflux += toadd;
      }
    }
  }
  f.actual_capillary_flux[i] = act_capflux;
  const double deepksat = kv_at_depth<N>(c.kv_profile, s.kv, pick<N>(s.kv.k, nlayers - 1), s.d_soil);
  const double deeptransfer = jmin(drainable / ddt, deepksat);
  const double leakage = jmax(0.0, jmin(staged(st2::max_leakage), deeptransfer));
  f.actual_leakage[i] = leakage;
  const double recharge = (s.transfer - act_capflux - leakage - ae_sat - soilevap_sat);
  f.recharge[i] = recharge;
  // the recharge / water-table hand-off to the subsurface flow (sbm_model.jl:74-81), fused:
  // exchange_recharge_kernel would re-read both arrays
  f.recharge_rate[i] = recharge;
  f.ssf_water_table_depth[i] = zi;
  // total AET (soil.jl:1206-1209) + interception (sbm.jl:130)
  double aet = soil_evaporation + transpiration + s.aeow_river + s.aeow_land + 0.0;
  aet += s.interception;
  f.actual_evapotranspiration[i] = aet;
  f.drainable_water_depth[i] = drainable;
#pragma unroll
  for (int k = 0; k < N; ++k) f.unsaturated_layer_depth[k * ns + i] = uld[k];
  // kh_layered_profile! (sbm_model.jl:84): the layered profiles' equivalent horizontal
  // conductivity of this step, from the water table and layers of the diagnostics above
#ifndef WFB_NO_KH
  if (c.kv_profile >= 2) {
    double kvl[N];
#pragma unroll
    for (int k = 0; k < N; ++k) kvl[k] = staged(st2::Rows<N>::kvlay + k);
    f.ssf_kh[i] = kh_layered<N>(c.kv_profile, s, kvl, s.kv.f, s.kv.zx, s.kv.nk + 1,
                                staged(st2::Rows<N>::khfrac));
  }
#endif
}

}  // namespace

// (Finishing every never-suspended cell inside land_hydrology_kernel, straight from registers,
// was built and measured: 128 registers with spills, 4 CTAs per SM, 436 us against 203 + 190 us
// for the two halves as kernels of their own -- see DESIGN.md section 5.1.)
// the loop engine's trips with the tracked power (unsatzone_flow_iterate_fast); 0: every trip
// with the reference expression
#ifndef WFB_ENGINE_FAST_TRIPS
#define WFB_ENGINE_FAST_TRIPS 1
#endif
#ifndef WFB_V_MINBLOCKS
#define WFB_V_MINBLOCKS 5
#endif
// update_land_hydrology_model!                                                sbm.jl:82-132
template <int N>
__global__ void __launch_bounds__(WFB_V_TILE, WFB_V_MINBLOCKS)
land_hydrology_kernel(const DevFields f, const KCfg c, const UnsatWork w, const StageList sl,
                      const double dt, const int tile_begin, const int phase) {
  // phase 0: the whole update. With lateral snow transport (snow_gravitational_transport__flag,
  // sbm.jl:98-100) a pass over the drainage network sits between the snow and the glacier
  // model: phase 1 = interception + snow, phase 2 = everything after the transport.
  // one tile of 128 consecutive slots per CTA
  const int i = (tile_begin + (int)blockIdx.x) * kTile + (int)threadIdx.x;
  if (i >= c.ns) return;    // whole warps (ns is a multiple of 32)
  // lanes of the padding slots [n, ns) run along on the padding values (they must take part in
  // the warp-aggregated suspension) and never suspend; their stores land in the padding
  const bool live = i < c.n;
  const int ns = c.ns;
  using R = st1::Rows<N>;
  // the few inputs that are not staged are asked for first, then everything else at once
  const double P = __ldg(f.precipitation + i);
  const double PET = __ldg(f.potential_evaporation + i);
  const double T = __ldg(f.temperature + i);
  const bool first_part = phase != 2;
  const bool with_lai = first_part && c.has_lai, with_snow = first_part && c.snow;
  const double lai = with_lai ? __ldg(f.leaf_area_index + i) : 0.0;
  const double sl_leaf = with_lai ? __ldg(f.storage_specific_leaf + i) : 0.0;
  const double s_wood = with_lai ? __ldg(f.storage_wood + i) : 0.0;
  const double k_ext = with_lai ? __ldg(f.light_extinction_coefficient + i) : 0.0;
  const double kc = first_part ? __ldg(f.crop_coefficient + i) : 0.0;
  const double tti = with_snow ? __ldg(f.temperature_interval_snowfall + i) : 0.0;
  const double tt = with_snow ? __ldg(f.temperature_threshold_snowfall + i) : 0.0;
  const int nlayers = phase != 1 ? f.number_of_layers[i] : 0;
  stage_issue(sl, tile_begin + (int)blockIdx.x);
  const Divisor ddt(dt);
  double st_canopy = 0.0, st_snoww = 0.0, st_gstore = 0.0, st_snow = 0.0, st_tsoil = 0.0;
  SoilColumn<N> s;

  // ---- forcing ---------------------------------------------------------------------------
  // ---- interception (canopy.jl:54-163, rainfall_interception.jl:9-130) ----------------------
  double cmax, gap;
  double canopy_potevap, throughfall = 0.0, interception, stemflow = 0.0;
  double snow = 0.0, snow_runoff = 0.0;
  if (phase == 2) {  // results of phase 1, from their (reference-visible) arrays
    gap = f.canopy_gap_fraction[i];
    canopy_potevap = f.canopy_potevap[i];
    interception = f.interception_rate[i];
    snow_runoff = f.snow_runoff[i];
    snow = f.snow_storage[i];       // after lateral_snow_transport!
  } else {
  if (c.has_lai) {
    cmax = sl_leaf * lai + s_wood;
    gap = exp(-k_ext * lai);
    f.maximum_canopy_storage[i] = cmax;
    f.canopy_gap_fraction[i] = gap;
  } else {
    stage_wait();
    cmax = staged(R::cmax);
    gap = staged(R::gap);
  }
  canopy_potevap = kc * PET * (1.0 - gap);
  if (c.gash) {
    double e_r;
    if (c.has_lai) {
      const double canopyfraction = 1.0 - gap;
      const double ewet = canopyfraction * PET * kc;
      const double thr = 1e-4 * (1e-3 * (1.0 / ddt));  // to_SI(1e-4, MM_PER_DT; dt)
      e_r = P > 0.0 ? jmin(0.25, fdiv(ewet, jmax(thr, canopyfraction * P))) : 0.0;
      f.evaporation_to_precipitation_ratio[i] = e_r;
    } else {
      e_r = staged(R::e_r);
    }
    if (cmax > 0.0) {
      double frac_stem, frac_int, p_sat;
      if (gap < 1.0 / 1.1) {
        frac_stem = 0.1 * gap;
        frac_int = 1.0 - 1.1 * gap;
        // e_r == 0 gives -Inf * 0 = NaN here, and `P > NaN` is false: kept on purpose
        p_sat = e_r > frac_int ? 0.0 : -cmax / (e_r * dt) * log(1.0 - fdiv(e_r, frac_int));
      } else {
        frac_stem = 1.0 - gap;
        frac_int = 0.0;
        p_sat = 0.0;
      }
      if (P > p_sat) {
        const double iwet = frac_int * p_sat - cmax / ddt;
        const double isat = e_r * (P - p_sat);
        const double idry = cmax / ddt;
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
    stage_wait();
    double cs = staged(R::canopy_storage);
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
    if (cs > cmax) { const double d = cs - cmax; cs = cmax; throughfall += d / ddt; }
    cs += p_canopy * dt;
    const double max_evap = cs / ddt;
    if (canopy_potevap > max_evap) { interception = max_evap; cs = 0.0; }
    else { interception = canopy_potevap; cs -= interception * dt; }
    if (cs > cmax) { const double d = cs - cmax; cs = cmax; throughfall += d / ddt; }
    st_canopy = cs;
  }
  f.canopy_potevap[i] = canopy_potevap;
  f.throughfall[i] = throughfall;
  f.interception_rate[i] = interception;
  f.stemflow[i] = stemflow;

  // ---- snow (snow.jl:123-177, snow_process.jl:26-116) ---------------------------------------
  if (c.snow) {
    const double eff = throughfall + stemflow;
    double rainfrac;
    if (tti == 0.0) rainfrac = T > tt ? 1.0 : 0.0;
    else rainfrac = jclamp(fdiv(T - (tt - tti / 2.0), tti), 0.0, 1.0);
    const double snowfrac = 1.0 - rainfrac;
    const double snow_precip = snowfrac * 1.0 * eff;
    const double liquid_precip = rainfrac * 1.0 * eff;
    stage_wait();
    snow = staged(st1::snow_storage);
    double snoww = staged(st1::snow_water);
    const double ttm = staged(st1::ttm);
    const double cfmax = staged(st1::cfmax);
    const double whc = staged(st1::whc);
    double snow_melt;
    if (T > ttm) {
      const double pot = cfmax * (T - ttm);
      snow_melt = jmin(pot, snow / ddt);
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
    if (snoww > maxw) { snow_runoff = (snoww - maxw) / ddt; snoww = maxw; }
    else snow_runoff = 0.0;
    f.effective_precip[i] = eff;
    f.snow_precip[i] = snow_precip;
    f.liquid_precip[i] = liquid_precip;
    st_snoww = snoww;
    f.snow_water_equivalent[i] = snoww + snow;
    f.snow_melt[i] = snow_melt;
    f.snow_runoff[i] = snow_runoff;
  }
  if (phase == 1) {  // the states lateral_snow_transport! works on; the rest follows in phase 2
    stage_wait();
    if (!c.gash) f.canopy_storage[i] = st_canopy;
    if (c.snow) { f.snow_water[i] = st_snoww; f.snow_storage[i] = snow; }
    return;
  }
  }  // phase != 2

  // ---- glacier (glacier.jl:122-154, glacier_process.jl:27-62) and the surface water flux ------
  double water_flux_surface;
  double gfrac = 0.0;
  const bool glac = c.snow && c.glacier;
  if (c.snow) {
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
      gmelt = snow < 1e-2 ? jmin(pot, gstore / ddt) : 0.0;
      gstore -= gmelt * dt;
      st_gstore = gstore;
      f.glacier_melt[i] = gmelt;
    }
    st_snow = snow;
    water_flux_surface = snow_runoff + gmelt * gfrac;  // runoff.jl:48-58
  } else {
    water_flux_surface = throughfall + stemflow;       // runoff.jl:37-46
  }
  f.runoff_water_flux_surface[i] = water_flux_surface;

  // ---- open-water runoff (runoff.jl:61-111) ------------------------------------------------
  stage_wait();
  const double rf = staged(st1::rf), wf = staged(st1::wf);
  const double h_land = staged(st1::olf_h);
  const double h_river = staged(st1::h_river);  // refreshed by scatter_river_depth_kernel
  f.waterdepth_land[i] = h_land;
  const double runoff_river = jmin(1.0, rf) * water_flux_surface;
  const double runoff_land = jmin(1.0, wf) * water_flux_surface;
  const double aeow_river = rf * jmin(h_river / ddt, PET);
  const double aeow_land = wf * jmin(h_land / ddt, PET);
  f.runoff_river[i] = runoff_river;
  f.runoff_land[i] = runoff_land;
  f.actual_open_water_evaporation_river[i] = aeow_river;
  f.actual_open_water_evaporation_land[i] = aeow_land;
  f.net_runoff_river[i] = runoff_river - aeow_river;

  // ---- soil boundary conditions (soil.jl:643-682) ------------------------------------------
  const double soil_fraction = jmax(gap - wf - rf - gfrac, 0.0);
  const double pot_transp = jmax(0.0, canopy_potevap - interception);
  const double pot_soilevap0 = soil_fraction * PET;
  const double wfs = jmax(water_flux_surface - runoff_river - runoff_land, 0.0);
  f.soil_fraction[i] = soil_fraction;
  f.potential_transpiration[i] = pot_transp;
  f.potential_soilevaporation[i] = pot_soilevap0;
  f.soil_water_flux_surface[i] = wfs;

  // ---- state -> diagnostics (soil.jl:1400-1436) --------------------------------------------
  s.theta_s = staged(st1::theta_s);
  s.theta_e = s.theta_s - staged(st1::theta_r);
  const double theta_e = s.theta_e;
  s.d_soil = staged(st1::d_soil);
  s.swc = staged(st1::swc);
  s.satwd = staged(st1::satwd);
  s.nlayers = nlayers;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    s.uld[k] = staged(R::uld + k);
    s.alt[k] = staged(R::alt + k);
    s.cld[k] = staged(R::cld + k);
  }
  s.cld[N] = staged(R::cld + N);
  double ustore_depth = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k < s.nlayers) ustore_depth += s.uld[k];
  const double zi = jmax(0.0, s.d_soil - fdiv(s.satwd, theta_e));
  s.zi = zi;
  const double ustore_cap = s.swc - s.satwd - ustore_depth;
  int n_unsat = N;
#pragma unroll
  for (int k = 0; k < N; ++k) {  // set_layerthickness utils.jl:390-404
    double t = qnan();
    if (zi > s.cld[k + 1]) t = s.alt[k];
    else if (zi - s.cld[k] > 0.0) t = zi - s.cld[k];
    s.ult[k] = t;
    n_unsat -= (t != t) ? 1 : 0;
    f.unsaturated_layer_thickness[k * ns + i] = t;
  }
  s.n_unsat = n_unsat;
  f.water_table_depth[i] = zi;
  f.n_unsatlayers[i] = n_unsat;
  f.total_soil_water_storage[i] = s.satwd + ustore_depth;

  // ---- soil temperature, infiltration (soil.jl:685-755, soil_process.jl:16-41,229-244) -------
  double f_red = 1.0;
  if (c.snow) {
    double tsoil = staged(st1::tsoil);
    tsoil = tsoil + staged(st1::w_soil) * (T - tsoil);
    st_tsoil = tsoil;
    if (c.soil_infiltration_reduction) {
      const double cf = staged(st1::cf_soil);
      const double bb = fdiv(1.0, 1.0 - cf);
      f_red = scurve(tsoil, 0.0 + 273.15, bb, 8.0) + cf;
    }
  }
  f.f_infiltration_reduction[i] = f_red;
  const double pathfrac = staged(st1::pathfrac);
  const double cap_soil = staged(st1::cap_soil);
  const double cap_path = staged(st1::cap_path);
  const double soilinf = wfs * (1.0 - pathfrac);
  const double pathinf = wfs * pathfrac;
  const double max_infiltsoil = jmin(cap_soil * f_red, soilinf);
  const double max_infiltpath = jmin(cap_path * f_red, pathinf);
  const double infiltration = jmin(max_infiltpath + max_infiltsoil, jmax(0.0, ustore_cap / ddt));
  const double infiltration_excess = (soilinf - max_infiltsoil) + (pathinf - max_infiltpath);
  f.infiltration[i] = infiltration;
  f.infiltration_excess[i] = infiltration_excess;

  // ---- unsaturated zone flow, Brooks-Corey (soil.jl:764-804) -------------------------------
  s.kv = staged_kvcol<N, R>(f, c, i, st1::kv_0, st1::kv_f);
#pragma unroll
  for (int k = 0; k < N; ++k) s.bc[k] = staged(R::bc + k);
  double transfer;
  const bool done = unsat_layers<N>(c, w, i, 0, n_unsat, 0.0, 0.0, infiltration, s.uld, s.ult,
                                    s.bc, s.kv, theta_e, dt, ddt, live, transfer);
  // the read-modify-write states of the sections above (phase 2: canopy and snow water were
  // written by phase 1)
  if (!c.gash && phase == 0) f.canopy_storage[i] = st_canopy;
  if (c.snow) {
    if (phase == 0) f.snow_water[i] = st_snoww;
    if (phase == 0 || c.glacier) f.snow_storage[i] = st_snow;
    f.soil_surface_temperature[i] = st_tsoil;
    if (c.glacier) f.glacier_store[i] = st_gstore;
  }
  // the second half runs in soil_column_kernel, after the loop engine has finished the
  // suspended cells' layers (nothing below is needed from registers: it is all in arrays)
  if (done) f.transfer[i] = transfer;
#pragma unroll
  for (int k = 0; k < N; ++k) f.unsaturated_layer_depth[k * ns + i] = s.uld[k];
  (void)pot_soilevap0; (void)infiltration_excess; (void)max_infiltsoil; (void)max_infiltpath;
  (void)aeow_river; (void)aeow_land;
}

// What soil_column_cell needs, re-read from the reference-visible arrays the first half wrote
// (s.uld, s.ult, s.bc, s.kv, s.theta_s, s.theta_e, s.n_unsat may already be loaded).
template <int N>
__device__ __forceinline__ void load_column_rest(const DevFields& f, const KCfg& c, const int i,
                                                 SoilColumn<N>& s) {
  const int ns = c.ns;
  using R = st2::Rows<N>;
  s.d_soil = staged(st2::d_soil);
  s.swc = staged(st2::swc);
  s.satwd = staged(st2::satwd);
  s.zi = staged(st2::zi);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    s.alt[k] = staged(R::alt + k);
    s.cld[k] = staged(R::cld + k);
  }
  s.cld[N] = staged(R::cld + N);
  s.pot_soilevap = staged(st2::pot_soilevap);
  s.pot_transp = staged(st2::pot_transp);
  s.infiltration = staged(st2::infiltration);
  s.infiltration_excess = staged(st2::infiltration_excess);
  s.wfs = staged(st2::wfs);
  const double f_red = staged(st2::f_red);
  s.pathfrac = staged(st2::pathfrac);
  // infiltration! soil_process.jl:16-41, the same expressions as in land_hydrology_kernel
  s.max_infiltsoil = jmin(staged(st2::cap_soil) * f_red, s.wfs * (1.0 - s.pathfrac));
  s.max_infiltpath = jmin(staged(st2::cap_path) * f_red, s.wfs * s.pathfrac);
  s.aeow_river = staged(st2::aeow_river);
  s.aeow_land = staged(st2::aeow_land);
  s.interception = staged(st2::interception);
}

// update_land_hydrology_model!, second half (soil evaporation ... recharge, AET), every cell:
// everything it needs from the first half is a reference-visible output array (or an input).
#ifndef WFB_VC_MINBLOCKS
#define WFB_VC_MINBLOCKS 4
#endif
template <int N>
__global__ void __launch_bounds__(WFB_V_TILE, WFB_VC_MINBLOCKS)
soil_column_kernel(const DevFields f, const KCfg c, const UnsatWork w, const StageList sl,
                   const double dt, const int tile_begin) {
  const int i = (tile_begin + (int)blockIdx.x) * kTile + (int)threadIdx.x;
  if (i >= c.ns) return;   // whole warps take part in the staging
  const int ns = c.ns;
  stage_issue(sl, tile_begin + (int)blockIdx.x);
  const Divisor ddt(dt);
  SoilColumn<N> s;
  s.theta_s = __ldg(f.theta_s + i);
  s.theta_e = s.theta_s - __ldg(f.theta_r + i);
  s.n_unsat = f.n_unsatlayers[i];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    s.uld[k] = f.unsaturated_layer_depth[k * ns + i];
    s.ult[k] = f.unsaturated_layer_thickness[k * ns + i];
    s.bc[k] = __ldg(f.brooks_corey_exponent + k * ns + i);
  }
  s.transfer = f.transfer[i];
  s.nlayers = f.number_of_layers[i];
  stage_wait();
  if (i >= c.n) return;
  s.kv = staged_kvcol<N, st2::Rows<N>>(f, c, i, st2::kv_0, st2::kv_f);
  load_column_rest<N>(f, c, i, s);
  soil_column_cell<N>(f, c, i, dt, ddt, s);
}

// The loop engine: one lane per suspended cell, 32 cells of one bucket (trip counts within a factor
// of two) per warp, longest buckets first. A lane runs the long loop its cell was suspended at
// (tracked trips), then the layers below with loops of any length in line. Per-layer operands are
// fetched when their layer is reached. What bounds the kernel is the latency of every warp's
// dependent instruction stream at the four warps per scheduler that 102 registers allow (ncu,
// profiles/r2final_engine_ncu.md: 44e6 warp instructions over 2368 resident warps, one issued
// every ~10 cycles; a loop itself has at most ~850 trips). 42 % of the instructions are issued for
// the 4.5-8 lanes still in the layers below the suspended one; sending those loops to a further
// launch, regrouped by trip count, was built and measured -- scratch traffic and launches ate
// the gain.
#ifndef WFB_ENGINE_MINBLOCKS
#define WFB_ENGINE_MINBLOCKS 4
#endif
template <int N>
__global__ void __launch_bounds__(128, WFB_ENGINE_MINBLOCKS)
unsat_engine_kernel(const DevFields f, const KCfg c, const UnsatWork w, const double dt) {
  const unsigned* cnt = w.count;
  const int32_t* list = w.list;
  unsigned* const queue = w.count + 2 * kBuckets;  // next warp tile to hand out
  const int lane = (int)threadIdx.x & 31;
  const int ns = c.ns;
  const Divisor ddt(dt);
  // warp tiles are handed out from a queue, longest buckets first: a warp that finishes a short
  // tile takes the next one, so every resident warp stays busy until the lists are empty (with a
  // fixed assignment a CTA lived as long as its longest tile while its other warps idled)
  for (;;) {
    int j = 0;
    if (lane == 0) j = (int)atomicAdd(queue, 1u);
    j = __shfl_sync(0xffffffffu, j, 0);
    int b, first, n;
    if (!tile_of(cnt, j, b, first, n)) break;
    const int e = first + lane;
    if (e < n) {   // (no `continue`: the warp meets again at the queue)
      const int i = list[(size_t)b * (size_t)w.cap + e];
      UnsatTask t;
      t.usd = w.usd[i]; t.sum_ast = w.sum_ast[i]; t.kv_it = w.kv_it[i]; t.l_sat = w.l_sat[i];
      t.c = w.c[i];
      const int code = w.its_layer[i];
      t.its = code & 0xffffff;
      const int kl = code >> 24;  // the layer the cell was suspended at
#if WFB_ENGINE_FAST_TRIPS
      unsatzone_flow_iterate_fast(t, dt, ddt);
#else
      unsatzone_flow_iterate(t, dt, ddt);
#endif
      f.unsaturated_layer_depth[kl * ns + i] = t.usd;
      double flow = t.sum_ast;
      const int n_unsat = f.n_unsatlayers[i];
      if (kl + 1 < n_unsat) {  // the layers below (soil.jl:770-801)
        const KvCol<N> kv = load_kvcol<N>(f, c, i);
        const double theta_e = __ldg(f.theta_s + i) - __ldg(f.theta_r + i);
        double z = 0.0;
        for (int k = 0; k <= kl; ++k) {  // same left-to-right sum as the first pass
          const double ultk = f.unsaturated_layer_thickness[k * ns + i];
          z = (k == 0) ? ultk : z + ultk;
        }
        for (int k = kl + 1; k < n_unsat; ++k) {
          const double ultk = f.unsaturated_layer_thickness[k * ns + i];
          z = z + ultk;
          const double l_sat = ultk * theta_e;
          const double kv_z = kv_at_depth<N>(c.kv_profile, kv, pick<N>(kv.k, k), z);
          const double usd = f.unsaturated_layer_depth[k * ns + i] + flow * dt;
          UnsatTask tk = unsatzone_flow_setup(usd, kv_z, l_sat,
                                              __ldg(f.brooks_corey_exponent + k * ns + i), dt, ddt);
          // diagnostic (wflowb200_get_unsat_buckets): later loops of the cell by trip count
          if (tk.its > 4) atomicAdd(w.count + kBuckets + bucket_of(tk.its), 1u);
#if WFB_ENGINE_FAST_TRIPS
          unsatzone_flow_iterate_fast(tk, dt, ddt);
#else
          unsatzone_flow_iterate(tk, dt, ddt);
#endif
          f.unsaturated_layer_depth[k * ns + i] = tk.usd;
          flow = tk.sum_ast;
        }
      }
      f.transfer[i] = flow;  // n_unsat > 0 for a suspended cell
    }
    __syncwarp();
  }
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
  soil_water_storage_cell<N>(f, c.ns, i);
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

// ---- self-test of device_math.cuh (wflowb200_selftest_math) ----------------------------------
// The largest relative difference of jpow(x, c) = exp(c log x) for x in (0, 1], c in [1, 40]
// against libdevice's pow, and checks of fdiv, jmin, jmax and jcld_pos against the plain
// formulations; out[0], out[1]: largest relative difference of the remaining store / the summed
// flux between the loop engine's fast trips and the reference loop.
__device__ __forceinline__ double ulp_dist(double a, double b) {
  if (a == b || (a != a && b != b)) return 0.0;
  if (a != a || b != b) return 1e300;
  const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
  if ((ia < 0) != (ib < 0)) return 1e300;
  const long long d = ia - ib;
  return (double)(d < 0 ? -d : d);
}
__device__ __forceinline__ double u01_hash(unsigned long long i, unsigned long long salt) {
  unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + salt * 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
__global__ void selftest_math_kernel(long long n, unsigned long long* out) {
  double w_exp = 0.0, w_log = 0.0, w_pow = 0.0, w_div = 0.0, w_mm = 0.0, w_cld = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double u = u01_hash(i, 1), v = u01_hash(i, 2), t = u01_hash(i, 3);
    // the loop engine's fast trips against the reference loop, on tasks shaped like the model's
    {
      UnsatTask ta;
      ta.l_sat = 0.02 + 0.28 * u;
      ta.c = 8.0 + 4.0 * v;
      const double dtl = 86400.0;
      const Divisor ddl(dtl);
      const double usd0 = ta.l_sat * (0.3 + 0.7 * t);
      const double kv_z = (0.001 + 0.3 * u01_hash(i, 4)) / dtl;
      ta = unsatzone_flow_setup(usd0, kv_z, ta.l_sat, ta.c, dtl, ddl);
      if (ta.its > 0 && ta.its < 400) {
        UnsatTask tb = ta;
        unsatzone_flow_iterate(ta, dtl, ddl);
        unsatzone_flow_iterate_fast(tb, dtl, ddl);
        if (ta.usd > 1e-12) w_exp = fmax(w_exp, fabs(tb.usd - ta.usd) / ta.usd);
        w_log = fmax(w_log, fabs(tb.sum_ast - ta.sum_ast) / ta.sum_ast);
      }
    }
    // jpow(x, c) = exp(c log x) against libdevice's pow (different algorithm, both < 1 ulp-ish)
    const double xp = (i & 1) ? v : 1.0 - v * v * v, cp = 1.0 + 39.0 * t;
    const double want = pow(xp, cp), got = jpow(xp, cp);
    if (want > 1e-290) w_pow = fmax(w_pow, fabs(got - want) / want);
    // fdiv == IEEE division for normal operands; zero numerator gives zero
    const double a = (i % 7 == 0) ? 0.0 : ldexp(0.5 + u, (int)(v * 600.0) - 300);
    const double bb = ldexp(0.5 + t, (int)(u * 400.0) - 200);
    w_div = fmax(w_div, ulp_dist(fdiv(a, bb), a / bb));
    w_div = fmax(w_div, ulp_dist(a / Divisor(bb), a / bb));
    // jmin / jmax against the branchy Julia formulation (NaN propagation)
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double ma = (i % 11 == 0) ? nan : u - 0.5, mb = (i % 13 == 0) ? nan : v - 0.5;
    const double rmin = (ma != ma || mb != mb) ? nan : (ma < mb ? ma : mb);
    const double rmax = (ma != ma || mb != mb) ? nan : (ma > mb ? ma : mb);
    w_mm = fmax(w_mm, ulp_dist(jmin(ma, mb), rmin));
    w_mm = fmax(w_mm, ulp_dist(jmax(ma, mb), rmax));
    // cld(x, 2e-4) against Julia's round((x - mod(x, -y)) / y)
    double xc = u * 0.3;
    if ((i & 7) == 3) xc = 2e-4 * (double)(1 + (int)(v * 1000.0));
    w_cld = fmax(w_cld, fabs(jcld_pos(xc, 2e-4) - jcld(xc, 2e-4)));
  }
  double w[6] = {w_exp, w_log, w_pow, w_div, w_mm, w_cld};
  for (int q = 0; q < 6; ++q) {
    for (int o = 16; o > 0; o >>= 1) w[q] = fmax(w[q], __shfl_xor_sync(0xffffffffu, w[q], o));
    if ((threadIdx.x & 31) == 0)
      atomicMax(out + q, (unsigned long long)__double_as_longlong(w[q]));  // w >= 0
  }
}
int launch_selftest_math(long long n, unsigned long long* out, cudaStream_t s) {
  selftest_math_kernel<<<148 * 4, 256, 0, s>>>(n, out);
  return 1;
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

// The staging lists of the two dense kernels for this configuration (see the row enums above).
template <int N>
static void build_vertical_stage_n(const DevFields& f, const KCfg& c, VerticalStage& vs) {
  const size_t ns = (size_t)c.ns;
  auto put = [](StageList& l, int row, const double* p) {
    l.p[row] = p;
    if (row + 1 > l.rows) l.rows = row + 1;
  };
  auto layered = [&](StageList& l, int row0, const double* p, int layers) {
    for (int k = 0; k < layers; ++k) put(l, row0 + k, p + (size_t)k * ns);
  };
  vs = VerticalStage{};
  using R1 = st1::Rows<N>;
  for (int phase = 0; phase < 3; ++phase) {
    StageList& l = vs.first[phase];
    if (phase != 2) {  // interception and snow
      if (!c.has_lai) {
        put(l, R1::cmax, f.maximum_canopy_storage);
        put(l, R1::gap, f.canopy_gap_fraction);
        if (c.gash) put(l, R1::e_r, f.evaporation_to_precipitation_ratio);
      }
      if (!c.gash) put(l, R1::canopy_storage, f.canopy_storage);
      if (c.snow) {
        put(l, st1::snow_storage, f.snow_storage);
        put(l, st1::snow_water, f.snow_water);
        put(l, st1::ttm, f.temperature_threshold_melt);
        put(l, st1::cfmax, f.degree_day_factor);
        put(l, st1::whc, f.water_holding_capacity);
      }
    }
    if (phase == 1) continue;
    if (c.snow) {
      put(l, st1::tsoil, f.soil_surface_temperature);
      put(l, st1::w_soil, f.w_soil);
      if (c.soil_infiltration_reduction) put(l, st1::cf_soil, f.cf_soil);
    }
    put(l, st1::rf, f.river_fraction);
    put(l, st1::wf, f.water_fraction);
    put(l, st1::olf_h, f.olf_h);
    put(l, st1::h_river, f.waterdepth_river);
    put(l, st1::theta_s, f.theta_s);
    put(l, st1::theta_r, f.theta_r);
    put(l, st1::d_soil, f.soil_thickness);
    put(l, st1::swc, f.soil_water_capacity);
    put(l, st1::satwd, f.saturated_water_depth);
    put(l, st1::pathfrac, f.compacted_soil_area_fraction);
    put(l, st1::cap_soil, f.infiltration_capacity_soil);
    put(l, st1::cap_path, f.infiltration_capacity_compacted_soil);
    if (c.kv_profile < 2) put(l, st1::kv_0, f.kv_0);
    if (c.kv_profile != 2) put(l, st1::kv_f, f.hydraulic_conductivity_scale_parameter);
    if (c.kv_profile == 1) put(l, R1::kv_zx, f.z_exp);
    if (c.kv_profile == 3) put(l, R1::kv_zx, f.z_layered);
    layered(l, R1::kvfac, f.vertical_hydraulic_conductivity_factor, N);
    if (c.kv_profile >= 2) layered(l, R1::kvlay, f.kv, N);
    layered(l, R1::uld, f.unsaturated_layer_depth, N);
    layered(l, R1::alt, f.actual_layer_thickness, N);
    layered(l, R1::bc, f.brooks_corey_exponent, N);
    layered(l, R1::cld, f.cumulative_layer_depth, N + 1);
  }
  using R2 = st2::Rows<N>;
  StageList& l = vs.second;
  put(l, st2::d_soil, f.soil_thickness);
  put(l, st2::swc, f.soil_water_capacity);
  put(l, st2::satwd, f.saturated_water_depth);
  put(l, st2::zi, f.water_table_depth);
  put(l, st2::pot_soilevap, f.potential_soilevaporation);
  put(l, st2::pot_transp, f.potential_transpiration);
  put(l, st2::infiltration, f.infiltration);
  put(l, st2::infiltration_excess, f.infiltration_excess);
  put(l, st2::wfs, f.soil_water_flux_surface);
  put(l, st2::f_red, f.f_infiltration_reduction);
  put(l, st2::pathfrac, f.compacted_soil_area_fraction);
  put(l, st2::cap_soil, f.infiltration_capacity_soil);
  put(l, st2::cap_path, f.infiltration_capacity_compacted_soil);
  put(l, st2::aeow_river, f.actual_open_water_evaporation_river);
  put(l, st2::aeow_land, f.actual_open_water_evaporation_land);
  put(l, st2::interception, f.interception_rate);
  put(l, st2::theta_fc, f.theta_fc);
  put(l, st2::rooting_depth, f.rooting_depth);
  put(l, st2::h1, f.h1);
  put(l, st2::h2, f.h2);
  put(l, st2::h4, f.h4);
  put(l, st2::alpha_h1, f.alpha_h1);
  put(l, st2::air_entry, f.air_entry_pressure);
  put(l, st2::h3_high, f.h3_high);
  put(l, st2::h3_low, f.h3_low);
  put(l, st2::wet_root, f.wet_root_distribution_parameter);
  put(l, st2::cap_hmax, f.cap_hmax);
  put(l, st2::cap_n, f.cap_n);
  put(l, st2::max_leakage, f.maximum_leakage);
  if (c.kv_profile < 2) put(l, st2::kv_0, f.kv_0);
  if (c.kv_profile != 2) put(l, st2::kv_f, f.hydraulic_conductivity_scale_parameter);
  if (c.kv_profile == 1) put(l, R2::kv_zx, f.z_exp);
  if (c.kv_profile == 3) put(l, R2::kv_zx, f.z_layered);
  layered(l, R2::kvfac, f.vertical_hydraulic_conductivity_factor, N);
  layered(l, R2::alt, f.actual_layer_thickness, N);
  layered(l, R2::cld, f.cumulative_layer_depth, N + 1);
  layered(l, R2::rootf, f.rootfraction, N);
  if (c.kv_profile >= 2) {
    put(l, R2::khfrac, f.ssf_khfrac);
    layered(l, R2::kvlay, f.kv, N);
  }
  // more than 48 KB of dynamic shared memory has to be asked for once per kernel
  const int most = (int)sizeof(double) * kTile * kMaxStage;
  cudaFuncSetAttribute(land_hydrology_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
  cudaFuncSetAttribute(soil_column_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
}

void build_vertical_stage(const DevFields& f, const KCfg& c, int n_layers, VerticalStage& vs) {
  switch (n_layers) {
    case 1: build_vertical_stage_n<1>(f, c, vs); break;
    case 2: build_vertical_stage_n<2>(f, c, vs); break;
    case 3: build_vertical_stage_n<3>(f, c, vs); break;
    case 4: build_vertical_stage_n<4>(f, c, vs); break;
    case 5: build_vertical_stage_n<5>(f, c, vs); break;
    case 6: build_vertical_stage_n<6>(f, c, vs); break;
    case 7: build_vertical_stage_n<7>(f, c, vs); break;
    case 8: build_vertical_stage_n<8>(f, c, vs); break;
    default: vs = VerticalStage{}; break;
  }
}

int launch_scatter_river_depth(const DevFields& f, const KCfg& c, cudaStream_t s) {
  if (c.nriv == 0) return 0;
  scatter_river_depth_kernel<<<(c.nriv + 255) / 256, 256, 0, s>>>(f, c);
  return 1;
}

// land_hydrology_kernel (dense) -> unsat_engine_kernel (the suspended cells' loops) ->
// soil_column_kernel (dense), one stream. Measured on B200 (1000^2, scripts/vertical_timeline.py):
// running the engine on a side stream UNDER the dense kernels, in 1 to 8 slices, is slower than
// this plain sequence -- soil_column_kernel takes 400 instead of 190 us and the engine 200
// instead of 70 us when they share the SMs (all three are bound by FP64 dependency chains, not
// by DRAM), so nothing is hidden. tl (optional timing events): [0] start, [1] land_hydrology,
// [2] unsat_engine, [3] soil_column done.
int launch_land_hydrology(const DevFields& f, const KCfg& c, int n_layers, double dt,
                          const UnsatWork& w, const VerticalStage& vs, int engine_grid, int phase,
                          bool run_engine, cudaStream_t s, cudaEvent_t const* tl) {
  int launches = 0;
  const int n_tiles = (c.ns + kTile - 1) / kTile;
  const StageList& s1 = vs.first[phase];
  const size_t row = sizeof(double) * kTile, smem1 = s1.rows * row, smem2 = vs.second.rows * row;
  if (tl) cudaEventRecord(tl[0], s);
  if (phase == 1) {  // interception + snow of every cell: no loops, no engine
    WFB_DISPATCH_N(n_layers, (land_hydrology_kernel<N><<<n_tiles, kTile, smem1, s>>>(f, c, w, s1, dt, 0, 1)));
    return launches + 1;
  }
  cudaMemsetAsync(w.count, 0, (2 * kBuckets + 1) * sizeof(unsigned), s);
  WFB_DISPATCH_N(n_layers, (land_hydrology_kernel<N><<<n_tiles, kTile, smem1, s>>>(f, c, w, s1, dt, 0, phase)));
  if (tl) cudaEventRecord(tl[1], s);
  if (run_engine) {  // as many CTAs as are resident at once: the engine hands out its work itself
    static int per_sm[9] = {0};
    if (per_sm[n_layers] == 0) {
      int nb = 0;
      WFB_DISPATCH_N(n_layers, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, unsat_engine_kernel<N>, 128, 0));
      per_sm[n_layers] = nb > 0 ? nb : 1;
    }
    const int grid = engine_grid * per_sm[n_layers];   // engine_grid = SMs of the device
    WFB_DISPATCH_N(n_layers, (unsat_engine_kernel<N><<<grid, 128, 0, s>>>(f, c, w, dt)));
  }
  if (tl) cudaEventRecord(tl[2], s);
  WFB_DISPATCH_N(n_layers, (soil_column_kernel<N><<<n_tiles, kTile, smem2, s>>>(f, c, w, vs.second, dt, 0)));
  if (tl) cudaEventRecord(tl[3], s);
  return launches + 2 + (run_engine ? 1 : 0);
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
