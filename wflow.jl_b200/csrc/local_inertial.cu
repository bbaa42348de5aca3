// local_inertial.cu -- local-inertial river flow on the staggered grid (sm_100a), the river part
// of BASELINE config #4: update_river_flow_model!(::RiverFlowModel{<:LocalInertial})
// (routing/surface/surface_staggered_scheme.jl:800-838) with
//   stable_timestep                     :1004-1020   dt_s = alpha min_i L_i / sqrt(g h_i)
//   update_river_channel_flow!          :326-383     edge flow, local_inertial_flow
//                                                     (surface_process.jl:88-115)
//   update_bc_reservoir_model!          :627-661     reservoirs (their node leaves the active
//                                                     set, the edge carries the outflow)
//   update_water_depth_and_storage!     :723-759     node storage and depth
// and, with floodplain_1d__flag, the 1-D floodplain
//   update_floodplain_flow!             :440-533     edge flow over the FloodPlainProfile tables
//                                                     (floodplain.jl:287-354)
//   update_water_depth_and_storage!(floodplain, ...) :674-712  bankfull redistribution
// in the same two phases (an edge's floodplain flow needs only that edge's channel flow and
// the depths of the previous sub-step), so the floodplain adds no grid barrier.
// All reference paths are under /root/reference/Wflow/src.
//
// The scheme is explicit: every sub-step is edge-parallel, then node-parallel, and needs ONE
// global number first (the minimum Courant step of all nodes) -- a few hundred sub-steps per
// model day with a few microseconds of work each. Launching three kernels and reading the time
// step back per sub-step would cost more than the arithmetic, so the WHOLE model step runs in
// one persistent kernel: the grid is co-resident, phases are separated by a grid barrier, the
// minimum is an atomicMin on the bit pattern (positive doubles order like integers), and every
// thread evaluates the `while t < dt` loop (surface_staggered_scheme.jl:816-821) identically.
// River state is a few MB and stays in L2; values written by other CTAs are read past the L1.
//
// Edge i is the edge leaving node i (init_staggered_river_flow, :225-233), so edge arrays are
// river-sized; a pit drains into a ghost node whose depth is the boundary condition
// li_ghost_h and whose bed level equals the pit's (get_river_parameters :62-72).
#include <algorithm>
#include "device_math.cuh"
#include "kernels.cuh"
#include "model.cuh"
#include "reservoir.cuh"
#include "floodplain.cuh"

namespace wfb {

namespace {

constexpr double kG = 9.80665;  // GRAVITATIONAL_ACCELERATION  Wflow.jl:77
constexpr int kLiBlock = 256;

__device__ __forceinline__ unsigned li_ld_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier (all CTAs are resident). bar[0]: arrivals, bar[1]: generation. Bounded: a grid that
// is not co-resident raises the handle's error word instead of hanging (api.cu: check_device_error).
__device__ __forceinline__ bool li_grid_barrier(unsigned* bar, unsigned n_blocks, unsigned& gen,
                                                unsigned* err) {
  __shared__ int ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    ok = 1;
    __threadfence();
    if (atomicAdd(&bar[0], 1u) == n_blocks - 1u) {
      atomicExch(&bar[0], 0u);
      __threadfence();
      atomicAdd(&bar[1], 1u);
    } else {
      unsigned polls = 0;
      while (li_ld_u32(&bar[1]) == gen) {
        if ((++polls & 1023u) == 0u) {
          if (polls >= (1u << 24)) { atomicOr(err, 2u); ok = 0; break; }
          if (li_ld_u32(err) != 0u) { ok = 0; break; }
        }
      }
    }
    __threadfence();
  }
  ++gen;
  __syncthreads();
  return ok != 0;
}

// local_inertial_flow                                             surface_process.jl:88-115
__device__ __forceinline__ double local_inertial_flow(double q0, double zs0, double zs1, double hf,
                                                      double A, double R, double length,
                                                      double mannings_n_sq, int froude_limit,
                                                      double dt) {
  const double slope = (zs1 - zs0) / length;
  const double pow_R = cbrt(R * R * R * R);
  double q = ((q0 - kG * A * dt * slope) / (1.0 + kG * dt * mannings_n_sq * fabs(q0) / (pow_R * A)));
  const double fr = ((q / A) / sqrt(kG * hf)) * (double)froude_limit;
  if ((fabs(fr) > 1.0) && (q > 0.0)) q = sqrt(kG * hf) * A;
  if ((fabs(fr) > 1.0) && (q < 0.0)) q = -sqrt(kG * hf) * A;
  return q;
}


}  // namespace

__global__ void __launch_bounds__(kLiBlock, 4)
local_inertial_river_kernel(const __grid_constant__ DevFields f, const KCfg c, const LiLaunch w) {
  const int n = c.nriv;
  const int stride = (int)(gridDim.x * blockDim.x);
  const int tid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const unsigned n_blocks = gridDim.x;
  unsigned gen = li_ld_u32(&w.barrier[1]);
  __shared__ unsigned long long s_min;
  const double dt = w.dt;
  const unsigned long long inf_bits = 0x7ff0000000000000ull;
  const bool floodplain = w.fp_levels > 0;
  const FpTables fp{f.fp_profile_storage, f.fp_profile_width, f.fp_profile_flow_area,
                    f.fp_profile_wetted_perimeter, c.nrs, w.fp_levels, f.fp_depth};

  // set_reservoir_vars! / set_flow_vars!                          surface_kinwave.jl:227-237,269-275
  for (int p = tid; p < n; p += stride) {
    f.riv_q_cumulative[p] = 0.0;
    f.riv_actual_external_abstraction_cumulative[p] = 0.0;
    if (floodplain) f.fp_q_cumulative[p] = 0.0;
  }
  for (int i = tid; i < c.nres; i += stride) {
    f.res_inflow_cumulative[i] = 0.0;
    f.res_actual_external_abstraction_cumulative[i] = 0.0;
    f.res_outflow_cumulative[i] = 0.0;
    f.res_actevap_cumulative[i] = 0.0;
  }

  double t = 0.0;
  int count = 0;
  bool alive = true;
  while (t < dt && alive) {
    // ---- stable_timestep: alpha L / sqrt(g h), minimum over the nodes ------------------------
    if (threadIdx.x == 0) s_min = inf_bits;
    __syncthreads();
    double mine = __longlong_as_double((long long)inf_bits);
    for (int p = tid; p < n; p += stride) {
      const double h = __ldcg(f.riv_h + p);
      const double d = w.alpha * __ldg(f.riv_flow_length + p) / sqrt(kG * h);
      mine = d < mine ? d : mine;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, mine, o);
      mine = other < mine ? other : mine;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(&s_min, (unsigned long long)__double_as_longlong(mine));
    __syncthreads();
    unsigned long long* slot = w.dt_bits + (count & 1);
    if (threadIdx.x == 0 && s_min != inf_bits) atomicMin(slot, s_min);
    alive = li_grid_barrier(w.barrier, n_blocks, gen, w.err);
    if (!alive) break;
    const unsigned long long mb = __ldcg(slot);
    if (tid == 0) w.dt_bits[(count + 1) & 1] = inf_bits;  // the slot of the next sub-step
    double dt_s = mb == inf_bits ? 60.0 : __longlong_as_double((long long)mb);
    if (t + dt_s > dt) dt_s = dt - t;  // check_timestepsize  routing/timestepping.jl:11-16

    // ---- update_river_channel_flow!: the edge leaving every active node ------------------------
    // The scheme is bound by the bytes it moves: a discharge that did not change is not written
    // back, a discharge of zero is not added to its cumulative value (x + 0.0 == x), the edge
    // diagnostics (zs_at_edge, water_depth_at_edge) are written by the last sub-step only (every
    // sub-step overwrites them), and the floodplain profile is only looked up where water stands
    // above the floodplain's bed.
    const bool last = !(t + dt_s < dt);
    for (int p = tid; p < n; p += stride) {
      const int d = f.li_dst_slot[p];
      if (d == -1) continue;
      if (f.riv_reservoir && f.riv_reservoir[p] >= 0) continue;  // not in active_e
      const double q_previous = f.riv_q[p];
      const double h_src = __ldcg(f.riv_h + p);
      const double zb = __ldg(f.li_zb + p);
      const double zs_src = zb + h_src;
      const double h_dst = d == -2 ? __ldg(f.li_ghost_h + p) : __ldcg(f.riv_h + d);
      const double zs_dst = (d == -2 ? zb : __ldg(f.li_zb + d)) + h_dst;
      const double zs_at_edge = jmax(zs_src, zs_dst);
      const double hf = zs_at_edge - __ldg(f.li_zb_at_edge + p);
      if (last) {
        f.li_zs_at_edge[p] = zs_at_edge;
        f.li_water_depth_at_edge[p] = hf;
      }
      double q = 0.0;
      if (hf > w.h_thresh) {
        const double width = __ldg(f.li_flow_width_at_edge + p);
        const double A = width * hf;
        const double R = A / (2.0 * hf + width);
        q = local_inertial_flow(q_previous, zs_src, zs_dst, hf, A, R, __ldg(f.li_flow_length_at_edge + p),
                                __ldg(f.li_mannings_n_sq_at_edge + p), w.froude_limit, dt_s);
      }
      if (h_src <= 0.0) q = jmin(q, 0.0);
      if (h_dst <= 0.0) q = jmax(q, 0.0);
      if (__double_as_longlong(q) != __double_as_longlong(q_previous)) f.riv_q[p] = q;
      if (q != 0.0) f.riv_q_cumulative[p] += q * dt_s;
      if (floodplain) {  // update_floodplain_flow! of the same edge                 :440-533
        const double q_fp_previous = f.fp_q[p];
        const double hfp = jmax(zs_at_edge - __ldg(f.fp_zb_at_edge + p), 0.0);
        if (last) f.fp_water_depth_at_edge[p] = hfp;
        double qf = 0.0;
        // (no water above the floodplain's bed and an empty first level of the profile: the flow
        // area min(a_src, a_dst) is 0, the flow 0.0 -- without the lookups)
        if (hfp > 0.0 || __ldg(f.fp_profile_flow_area + p) > 0.0) {
          const int pd = d == -2 ? p : d;  // a ghost node copies the profile of its pit
          int i1, i2;
          fp_indices_depth(fp, hfp, i1, i2);
          const double a_src = fp_flow_area(fp, hfp, p, i1, i2);
          const double a_dst = fp_flow_area(fp, hfp, pd, i1, i2);
          const double A_fp = jmin(a_src, a_dst);
          const double R_fp = a_src < a_dst ? a_src / fp_wetted_perimeter(fp, hfp, p, i1)
                                            : a_dst / fp_wetted_perimeter(fp, hfp, pd, i1);
          qf = A_fp > 1.0e-05
                   ? local_inertial_flow(q_fp_previous, zs_src, zs_dst, hfp, A_fp, R_fp,
                                         __ldg(f.li_flow_length_at_edge + p),
                                         __ldg(f.fp_mannings_n_sq_at_edge + p), w.froude_limit, dt_s)
                   : 0.0;
          if (__ldcg(f.fp_h + p) <= 0.0) qf = jmin(qf, 0.0);
          if ((d == -2 ? 0.0 : __ldcg(f.fp_h + d)) <= 0.0) qf = jmax(qf, 0.0);
          if (qf * q < 0.0) qf = 0.0;  // opposite to the channel flow
        }
        if (__double_as_longlong(qf) != __double_as_longlong(q_fp_previous)) f.fp_q[p] = qf;
        if (qf != 0.0) f.fp_q_cumulative[p] += qf * dt_s;
      }
    }
    alive = li_grid_barrier(w.barrier, n_blocks, gen, w.err);
    if (!alive) break;

    // ---- update_bc_reservoir_model!: one thread per reservoir ----------------------------------
    if (c.nres > 0) {
      for (int v = tid; v < c.nres; v += stride) {
        const int p = f.res_river_slot[v];
        double q_in = 0.0;  // sum_at(q, edges_at_node.src[i])
        for (int e = f.li_in_ptr[p]; e < f.li_in_ptr[p + 1]; ++e) q_in += __ldcg(f.riv_q + f.li_in_idx[e]);
        if (floodplain) {  // get_inflow_reservoir                                      :291-299
          double q_fp = 0.0;
          for (int e = f.li_in_ptr[p]; e < f.li_in_ptr[p + 1]; ++e) q_fp += __ldcg(f.fp_q + f.li_in_idx[e]);
          q_in += q_fp;
        }
        const double outflow = reservoir_step(f, v, q_in, dt_s);
        f.riv_q[p] = outflow;
        f.riv_q_cumulative[p] += outflow * dt_s;
      }
      alive = li_grid_barrier(w.barrier, n_blocks, gen, w.err);
      if (!alive) break;
    }

    // ---- update_water_depth_and_storage! --------------------------------------------------------
    for (int p = tid; p < n; p += stride) {
      if (f.riv_reservoir && f.riv_reservoir[p] >= 0) continue;  // not in active_n
      // both sums over the entering edges in one walk (ascending source id each, like sum_at):
      // the index loads are shared and the two gathers of an edge are in flight together
      double q_src = 0.0, qf_src = 0.0;
      for (int e = f.li_in_ptr[p], e1 = f.li_in_ptr[p + 1]; e < e1; ++e) {
        const int u = f.li_in_idx[e];
        q_src += __ldcg(f.riv_q + u);
        if (floodplain) qf_src += __ldcg(f.fp_q + u);
      }
      const int dst = f.li_dst_slot[p];
      const double q_dst = dst == -1 ? 0.0 : 0.0 + f.riv_q[p];
      double storage = f.riv_storage[p];
      storage += (q_src - q_dst + f.riv_inwater[p] - __ldg(f.riv_abstraction + p)) * dt_s;
      if (storage < 0.0) {
        f.li_error[p] = f.li_error[p] + fabs(storage);
        storage = 0.0;
      }
      const double ext = __ldg(f.riv_external_inflow + p);
      double inflow;
      if (ext < 0.0) {
        const double abstraction = jmin(-ext, storage / dt_s * 0.80);
        f.riv_actual_external_abstraction_cumulative[p] += abstraction * dt_s;
        inflow = -abstraction;
      } else {
        inflow = ext;
      }
      storage += inflow * dt_s;
      const double length = __ldg(f.riv_flow_length + p), width = __ldg(f.riv_flow_width + p);
      double h_new = storage / (length * width);
      if (floodplain) {  // update_water_depth_and_storage!(floodplain, ...)           :674-712
        const double qf_dst = dst == -1 ? 0.0 : 0.0 + f.fp_q[p];
        double fs = f.fp_storage[p];
        fs += (qf_src - qf_dst) * dt_s;
        if (fs < 0.0) {
          f.fp_error[p] += fabs(fs);
          fs = 0.0;
        }
        const double storage_total = storage + fs;
        const double bankfull = __ldg(f.li_bankfull_storage + p);
        double fh;
        if (storage_total > bankfull) {
          const double hh = fp_flood_depth(fp, storage_total - bankfull, length, p);
          h_new = __ldg(f.li_bankfull_depth + p) + hh;
          storage = h_new * width * length;
          fs = jmax(storage_total - storage, 0.0);
          fh = fs > 0.0 ? hh : 0.0;
        } else {
          h_new = storage_total / (length * width);
          storage = storage_total;
          fh = 0.0;
          fs = 0.0;
        }
        f.fp_storage[p] = fs;
        f.fp_h[p] = fh;
      }
      f.riv_storage[p] = storage;
      f.riv_h[p] = h_new;
    }
    t += dt_s;
    ++count;
    // (no barrier here: the next sub-step's minimum reads only this thread's own depths, and its
    // edge phase comes after the barrier that follows the minimum)
  }
  // average_flow_vars! / average_reservoir_vars!                   surface_kinwave.jl:244-258,283-290
  for (int p = tid; p < n; p += stride) {
    const double q_av = f.riv_q_cumulative[p] / dt;
    f.riv_q_average[p] = q_av;
    f.riv_actual_external_abstraction_average[p] = f.riv_actual_external_abstraction_cumulative[p] / dt;
    if (floodplain) {  // :826-835
      const double qf_av = f.fp_q_cumulative[p] / dt;
      f.fp_q_average[p] = qf_av;
      f.riv_q_channel_average[p] = q_av;
      f.riv_q_average[p] = q_av + qf_av;
    }
  }
  for (int i = tid; i < c.nres; i += stride) {
    f.res_outflow_average[i] = f.res_outflow_cumulative[i] / dt;
    f.res_inflow_average[i] = f.res_inflow_cumulative[i] / dt;
    f.res_actual_external_abstraction_average[i] = f.res_actual_external_abstraction_cumulative[i] / dt;
  }
  if (tid == 0) *w.substeps = count;
}

// ---------------------------------------------------------------------------------------------
// 2-D local-inertial overland flow coupled to the local-inertial river (land_routing = 1):
// update_overland_flow_model!(overland, river, domain, clock, dt; update_h = false)
// (surface_staggered_scheme.jl:1153-1194) with
//   stable_timestep (land)              :1022-1043   alpha min(dx, dy) / sqrt(g h), non-river cells
//   local_inertial_update_fluxes!       :1276-1295   x / y edge flow of every cell
//     update_directional_flow!          :1201-1271   local_inertial_flow(theta, ...), de Almeida
//                                                     et al. 2012 (surface_process.jl:123-159)
//   update_inflow_reservoir!            :1301-1319   overland flow into the reservoirs
//   staggered_scheme_river_update!      :762-794     river edge flow + reservoirs (update_h = false)
//   local_inertial_update_water_depth!  :1520-1546   land cells; river cells with their subgrid
//                                                     channel (bankfull spill), :1325-1514
// The same persistent-kernel organisation as the river alone: per sub-step one global minimum,
// an edge phase (the two edges of every land cell AND the river edges: both read only depths),
// the reservoirs, a node phase -- separated by grid barriers. The thread that owns a LAND cell
// also owns the water depth and storage of the river node in it, so the Courant steps of the next
// sub-step are formed right where the new depths are computed and need no pass of their own.
// Values another thread may have written since the last barrier are read past the L1 (__ldcg).
namespace {

// local_inertial_flow(theta, q0, qd, qu, zs0, zs1, hf, width, length, mannings_n_sq, froude_limit,
// dt): flow through a rectangular area                              surface_process.jl:123-159
__device__ __forceinline__ double local_inertial_flow_rect(double theta, double q0, double qd, double qu,
                                                           double zs0, double zs1, double hf,
                                                           double width, double length,
                                                           double mannings_n_sq, int froude_limit,
                                                           double dt) {
  // fdiv: the correctly rounded quotient without nvcc's guard, which sends every ZERO numerator
  // (still water: the common case) through a ~90-instruction slow path; every divisor here is a
  // normal number (cell lengths, hf > h_thresh, width != 0, 1 + a non-negative term)
  const double slope = fdiv(zs1 - zs0, length);
  const double pow_hf = cbrt(hf * hf * hf * hf * hf * hf * hf);
  double q = fdiv((theta * q0 + 0.5 * (1.0 - theta) * (qu + qd)) - kG * hf * width * dt * slope,
                  1.0 + fdiv(kG * dt * mannings_n_sq * fabs(q0), pow_hf * width));
  if (froude_limit) {
    const double celerity = sqrt(kG * hf);
    const double fr = fdiv(fdiv(fdiv(q, width), hf), celerity);
    if (fabs(fr) > 1.0 && q > 0.0) q = hf * celerity * width;
    else if (fabs(fr) > 1.0 && q < 0.0) q = -hf * celerity * width;
  }
  return q;
}

}  // namespace

__global__ void __launch_bounds__(kLiBlock, 4)
local_inertial_land_river_kernel(const __grid_constant__ DevFields f, const KCfg c, const LiLaunch w) {
  // The 2-D state lives in NODE order (the reference's own order: column-major over the raster),
  // not in the slot order of the wavefront kernels: a cell's x neighbours are its neighbours in
  // memory and its y neighbours sit one raster column away, so the neighbour loads of a warp are
  // coalesced. h and storage (olf_h / olf_storage, slot order, shared with the vertical update)
  // are gathered into node-ordered work arrays at the start of the model step and scattered back
  // at its end.
  const int n = c.n, nriv = c.nriv;
  const int stride = (int)(gridDim.x * blockDim.x);
  const int tid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const unsigned n_blocks = gridDim.x;
  unsigned gen = li_ld_u32(&w.barrier[1]);
  __shared__ unsigned long long s_min[2];   // river, land
  const double dt = w.dt;
  const unsigned long long inf_bits = 0x7ff0000000000000ull;
  const double inf = __longlong_as_double((long long)inf_bits);
  double* const land_h = f.lil_h;
  double* const land_storage = f.lil_storage;

  // The Courant steps of the cells this thread owns: river nodes alpha L / sqrt(g h) (:1004-1020),
  // non-river land cells alpha min(dx, dy) / sqrt(g h) (:1022-1043)
  double mine_river = inf, mine_land = inf;
  auto courant = [&](const int v, const int r, const double h_land, const double h_river) {
    // (a dry cell gives alpha L / 0 = Inf, which never is the minimum: skipped)
    if (r >= 0) {
      if (h_river > 0.0) {
        const double d = fdiv(w.alpha * __ldg(f.riv_flow_length + r), sqrt(kG * h_river));
        mine_river = d < mine_river ? d : mine_river;
      }
    } else if (h_land > 0.0) {
      const double d = fdiv(w.land_alpha * jmin(__ldg(f.li_land_x_length + v), __ldg(f.li_land_y_length + v)),
                            sqrt(kG * h_land));
      mine_land = d < mine_land ? d : mine_land;
    }
  };

  // set_reservoir_vars! / set_flow_vars! (river :269-275, overland :1127-1132); qx0 .= qx, qy0 .= qy
  // of the first sub-step (:1284-1285)
  for (int p = tid; p < nriv; p += stride) f.riv_q_cumulative[p] = 0.0;
  for (int v = tid; v < n; v += stride) {
    f.li_land_qx_cumulative[v] = 0.0;
    f.li_land_qy_cumulative[v] = 0.0;
    f.li_land_qx0[v] = f.li_land_qx[v];
    f.li_land_qy0[v] = f.li_land_qy[v];
    const int slot = f.land_slot_of_node[v];
    const double h0 = f.olf_h[slot];
    land_h[v] = h0;
    land_storage[v] = f.olf_storage[slot];
    const int r = f.lil_river_slot[v];
    if (r >= 0) f.riv_actual_external_abstraction_cumulative[r] = 0.0;
    courant(v, r, h0, r >= 0 ? f.riv_h[r] : 0.0);
    // the edges that carry flow (:1236: upstream_idx <= n && width_at_edge != 0) and the cell's
    // area, once per model step (the widths are fields the host may have set since the last one)
    f.lil_xu_eff[v] = f.li_land_ywidth_at_edge[v] != 0.0 ? f.edge_x_up[v] : -1;
    f.lil_yu_eff[v] = f.li_land_xwidth_at_edge[v] != 0.0 ? f.edge_y_up[v] : -1;
    f.lil_cell_area[v] = f.li_land_x_length[v] * f.li_land_y_length[v];
  }
  for (int i = tid; i < c.nres; i += stride) {
    f.res_inflow_cumulative[i] = 0.0;
    f.res_actual_external_abstraction_cumulative[i] = 0.0;
    f.res_outflow_cumulative[i] = 0.0;
    f.res_actevap_cumulative[i] = 0.0;
  }

  double t = 0.0;
  int count = 0;
  bool alive = true;
  while (t < dt && alive) {
    // ---- dt_s = min(stable_timestep(river), stable_timestep(land))                  :1172-1175 ----
    if (threadIdx.x == 0) { s_min[0] = inf_bits; s_min[1] = inf_bits; }
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) {
      const double a = __shfl_xor_sync(0xffffffffu, mine_river, o);
      const double b = __shfl_xor_sync(0xffffffffu, mine_land, o);
      mine_river = a < mine_river ? a : mine_river;
      mine_land = b < mine_land ? b : mine_land;
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&s_min[0], (unsigned long long)__double_as_longlong(mine_river));
      atomicMin(&s_min[1], (unsigned long long)__double_as_longlong(mine_land));
    }
    __syncthreads();
    unsigned long long* slot = w.dt_bits + 2 * (count & 1);
    if (threadIdx.x == 0) {
      if (s_min[0] != inf_bits) atomicMin(slot, s_min[0]);
      if (s_min[1] != inf_bits) atomicMin(slot + 1, s_min[1]);
    }
    alive = li_grid_barrier(w.barrier, n_blocks, gen, w.err);
    if (!alive) break;
    const unsigned long long mb_river = __ldcg(slot), mb_land = __ldcg(slot + 1);
    if (tid == 0) {  // the slots of the next sub-step
      w.dt_bits[2 * ((count + 1) & 1)] = inf_bits;
      w.dt_bits[2 * ((count + 1) & 1) + 1] = inf_bits;
    }
    const double dt_river = mb_river == inf_bits ? 60.0 : __longlong_as_double((long long)mb_river);
    const double dt_land = mb_land == inf_bits ? 60.0 : __longlong_as_double((long long)mb_land);
    double dt_s = jmin(dt_river, dt_land);
    if (t + dt_s > dt) dt_s = dt - t;  // check_timestepsize  routing/timestepping.jl:11-16

    const bool last = !(t + dt_s < dt);   // the last sub-step of the model step
    // ---- local_inertial_update_fluxes!: the x and the y edge of every cell          :1276-1295 ----
    for (int v = tid; v < n; v += stride) {
      // batch 1: the cell's own values and its edges (xu / yu: -1 also where the effective flow
      // width is zero, see the prologue); batch 2: depth and elevation of the two upstream cells (a
      // missing neighbour reads the cell itself). Everything else -- widths, roughness, lengths,
      // the neighbours' previous flows -- is only loaded for an edge that is wet: most edges of a
      // basin are dry, and the scheme is bound by the bytes it moves.
      const int xu = f.lil_xu_eff[v], yu = f.lil_yu_eff[v];
      const double h_v = __ldcg(land_h + v);
      const double z_v = __ldg(f.li_land_z + v);
      const double zx_max = __ldg(f.li_land_zx_max_at_edge + v), zy_max = __ldg(f.li_land_zy_max_at_edge + v);
      const double q0x = __ldcg(f.li_land_qx0 + v), q0y = __ldcg(f.li_land_qy0 + v);
      const int xu_c = xu >= 0 ? xu : v, yu_c = yu >= 0 ? yu : v;
      const double h_xu = __ldcg(land_h + xu_c), h_yu = __ldcg(land_h + yu_c);
      const double z_xu = __ldg(f.li_land_z + xu_c), z_yu = __ldg(f.li_land_z + yu_c);
      const double zs_v = z_v + h_v;
      if (xu >= 0) {  // update_directional_flow!, x direction                        :1201-1271
        const double zs_up = z_xu + h_xu;
        const double hf = (jmax(zs_v, zs_up) - zx_max);
        double q = 0.0;
        if (hf > w.land_h_thresh) {
          const int xd = f.edge_x_down[v];
          const double length_at_edge = 0.5 * (__ldg(f.li_land_x_length + v) + __ldg(f.li_land_x_length + xu));
          q = local_inertial_flow_rect(w.land_theta, q0x, xd >= 0 ? __ldcg(f.li_land_qx0 + xd) : 0.0,
                                       __ldcg(f.li_land_qx0 + xu), zs_v, zs_up, hf,
                                       __ldg(f.li_land_ywidth_at_edge + v), length_at_edge,
                                       __ldg(f.li_land_mannings_n_sq_at_edge + v), w.land_froude_limit, dt_s);
          if (h_v <= 0.0) q = jmin(q, 0.0);
          if (h_xu <= 0.0) q = jmax(q, 0.0);
        }
        // qx[v] still holds q0x (qx0 .= qx): a flow that did not change is not written, and a flow
        // of zero adds nothing to the cumulative flow (x + 0.0 == x)
        if (__double_as_longlong(q) != __double_as_longlong(q0x)) __stcg(f.li_land_qx + v, q);
        if (q != 0.0) f.li_land_qx_cumulative[v] += q * dt_s;
      }
      if (yu >= 0) {  // y direction
        const double zs_up = z_yu + h_yu;
        const double hf = (jmax(zs_v, zs_up) - zy_max);
        double q = 0.0;
        if (hf > w.land_h_thresh) {
          const int yd = f.edge_y_down[v];
          const double length_at_edge = 0.5 * (__ldg(f.li_land_y_length + v) + __ldg(f.li_land_y_length + yu));
          q = local_inertial_flow_rect(w.land_theta, q0y, yd >= 0 ? __ldcg(f.li_land_qy0 + yd) : 0.0,
                                       __ldcg(f.li_land_qy0 + yu), zs_v, zs_up, hf,
                                       __ldg(f.li_land_xwidth_at_edge + v), length_at_edge,
                                       __ldg(f.li_land_mannings_n_sq_at_edge + v), w.land_froude_limit, dt_s);
          if (h_v <= 0.0) q = jmin(q, 0.0);
          if (h_yu <= 0.0) q = jmax(q, 0.0);
        }
        if (__double_as_longlong(q) != __double_as_longlong(q0y)) __stcg(f.li_land_qy + v, q);
        if (q != 0.0) f.li_land_qy_cumulative[v] += q * dt_s;
      }
    }
    // ---- update_river_channel_flow!: the edge leaving every active river node        :326-383 ----
    for (int p = tid; p < nriv; p += stride) {
      const int d = f.li_dst_slot[p];
      if (d == -1) continue;
      if (f.riv_reservoir && f.riv_reservoir[p] >= 0) continue;  // not in active_e
      const double q_previous = __ldcg(f.riv_q + p);
      const double h_src = __ldcg(f.riv_h + p);
      const double zb = __ldg(f.li_zb + p);
      const double zs_src = zb + h_src;
      const double h_dst = d == -2 ? __ldg(f.li_ghost_h + p) : __ldcg(f.riv_h + d);
      const double zs_dst = (d == -2 ? zb : __ldg(f.li_zb + d)) + h_dst;
      const double zs_at_edge = jmax(zs_src, zs_dst);
      const double hf = zs_at_edge - __ldg(f.li_zb_at_edge + p);
      if (last) {  // (every sub-step overwrites the edge diagnostics: the last one's stay)
        f.li_zs_at_edge[p] = zs_at_edge;
        f.li_water_depth_at_edge[p] = hf;
      }
      double q = 0.0;
      if (hf > w.h_thresh) {
        const double width = __ldg(f.li_flow_width_at_edge + p);
        const double A = width * hf;
        const double R = A / (2.0 * hf + width);
        q = local_inertial_flow(q_previous, zs_src, zs_dst, hf, A, R, __ldg(f.li_flow_length_at_edge + p),
                                __ldg(f.li_mannings_n_sq_at_edge + p), w.froude_limit, dt_s);
      }
      if (h_src <= 0.0) q = jmin(q, 0.0);
      if (h_dst <= 0.0) q = jmax(q, 0.0);
      if (__double_as_longlong(q) != __double_as_longlong(q_previous)) __stcg(f.riv_q + p, q);
      if (q != 0.0) __stcg(f.riv_q_cumulative + p, __ldcg(f.riv_q_cumulative + p) + q * dt_s);
    }
    alive = li_grid_barrier(w.barrier, n_blocks, gen, w.err);
    if (!alive) break;

    // ---- update_inflow_reservoir! (:1301-1319) + update_bc_reservoir_model! (:627-661) -----------
    if (c.nres > 0) {
      for (int i = tid; i < c.nres; i += stride) {
        const int p = f.res_river_slot[i];
        const int j = f.res_land_node[i];
        const int xd = f.edge_x_down[j], yd = f.edge_y_down[j];
        const double net_land_flow = (xd >= 0 ? __ldcg(f.li_land_qx + xd) : 0.0) - __ldcg(f.li_land_qx + j) +
                                     (yd >= 0 ? __ldcg(f.li_land_qy + yd) : 0.0) - __ldcg(f.li_land_qy + j);
        f.res_inflow_overland[i] = f.li_land_runoff[j] + (net_land_flow);
        double q_in = 0.0;  // sum_at(q, edges_at_node.src[i])
        for (int e = f.li_in_ptr[p]; e < f.li_in_ptr[p + 1]; ++e) q_in += __ldcg(f.riv_q + f.li_in_idx[e]);
        const double outflow = reservoir_step(f, i, q_in, dt_s);
        __stcg(f.riv_q + p, outflow);
        __stcg(f.riv_q_cumulative + p, __ldcg(f.riv_q_cumulative + p) + outflow * dt_s);
      }
      alive = li_grid_barrier(w.barrier, n_blocks, gen, w.err);
      if (!alive) break;
    }

    // ---- local_inertial_update_water_depth!                                         :1520-1546 ----
    mine_river = inf;
    mine_land = inf;
    for (int v = tid; v < n; v += stride) {
      const int r = f.lil_river_slot[v];
      const int xd = f.edge_x_down[v], yd = f.edge_y_down[v];
      const double qx_v = __ldcg(f.li_land_qx + v), qy_v = __ldcg(f.li_land_qy + v);
      const double qx_xd = __ldcg(f.li_land_qx + (xd >= 0 ? xd : v));   // (a missing neighbour reads
      const double qy_yd = __ldcg(f.li_land_qy + (yd >= 0 ? yd : v));   // the cell itself: unused)
      const double net_land_flow = (xd >= 0 ? qx_xd : 0.0) - qx_v + (yd >= 0 ? qy_yd : 0.0) - qy_v;
      if (!last) {  // qx0 .= qx, qy0 .= qy of the next sub-step (:1284-1285): this thread's own edges
        __stcg(f.li_land_qx0 + v, qx_v);
        __stcg(f.li_land_qy0 + v, qy_v);
      }
      const double runoff = f.li_land_runoff[v];
      double storage = land_storage[v];
      if (r < 0) {  // update_land_storage_and_depth!                                  :1491-1514
        storage += (net_land_flow + runoff) * dt_s;
        if (storage < 0.0) {
          f.li_land_error[v] += fabs(storage);
          storage = 0.0;
        }
        const double h_new = fdiv(storage, f.lil_cell_area[v]);
        land_storage[v] = storage;
        __stcg(land_h + v, h_new);
        courant(v, r, h_new, 0.0);
        continue;
      }
      if (f.riv_reservoir && f.riv_reservoir[r] >= 0) {  // reservoir outlet: h stays as it is
        courant(v, r, 0.0, f.riv_h[r]);
        continue;
      }
      // update_river_and_land_storage_and_depth!                                      :1443-1485
      double q_src = 0.0;  // compute_river_storage_change                             :1325-1352
      for (int e = f.li_in_ptr[r]; e < f.li_in_ptr[r + 1]; ++e) q_src += __ldcg(f.riv_q + f.li_in_idx[e]);
      const double q_dst = f.li_dst_slot[r] == -1 ? 0.0 : 0.0 + __ldcg(f.riv_q + r);
      const double net_river_flow = q_src - q_dst;
      const double net_flow = net_river_flow + net_land_flow + runoff - __ldg(f.riv_abstraction + r);
      storage += net_flow * dt_s;
      if (storage < 0.0) {
        f.li_land_error[v] += fabs(storage);
        storage = 0.0;
      }
      const double bankfull_storage = __ldg(f.li_bankfull_storage + r);
      const double ext = __ldg(f.riv_external_inflow + r);
      double inflow;  // compute_external_inflow                                       :1359-1382
      if (ext < 0.0) {
        const double available_volume = storage >= bankfull_storage ? bankfull_storage : f.riv_storage[r];
        const double abstraction = jmin(-ext, available_volume / dt_s * 0.80);
        f.riv_actual_external_abstraction_cumulative[r] += abstraction * dt_s;
        inflow = -abstraction;
      } else {
        inflow = ext;   // (the cumulative abstraction grows by 0.0 * dt: unchanged)
      }
      storage += inflow * dt_s;
      const double length = __ldg(f.riv_flow_length + r), width = __ldg(f.riv_flow_width + r);
      double river_h, h_new, river_storage;  // compute_water_depths                   :1388-1416
      if (storage >= bankfull_storage) {
        const double bankfull_depth = __ldg(f.li_bankfull_depth + r);
        river_h = bankfull_depth + fdiv(storage - bankfull_storage, f.lil_cell_area[v]);
        h_new = river_h - bankfull_depth;
        river_storage = river_h * length * width;
      } else {
        river_h = fdiv(storage, length * width);
        h_new = 0.0;
        river_storage = storage;
      }
      land_storage[v] = storage;
      __stcg(f.riv_h + r, river_h);
      __stcg(land_h + v, h_new);
      f.riv_storage[r] = river_storage;
      courant(v, r, h_new, river_h);
    }
    t += dt_s;
    ++count;
    // (no barrier here: the next sub-step's minimum is formed from this thread's own new depths,
    // and its edge phase comes after the barrier that follows the minimum)
  }
  // average_flow_vars! (river :283-290, overland :1138-1147) / average_reservoir_vars!; h and
  // storage back into the arrays the vertical update and the output read (this thread's own cells)
  for (int p = tid; p < nriv; p += stride) f.riv_q_average[p] = __ldcg(f.riv_q_cumulative + p) / dt;
  for (int v = tid; v < n; v += stride) {
    f.li_land_qx_average[v] = f.li_land_qx_cumulative[v] / dt;
    f.li_land_qy_average[v] = f.li_land_qy_cumulative[v] / dt;
    const int slot = f.land_slot_of_node[v];
    f.olf_h[slot] = __ldcg(land_h + v);
    f.olf_storage[slot] = land_storage[v];
    const int r = f.lil_river_slot[v];
    if (r >= 0)
      f.riv_actual_external_abstraction_average[r] = f.riv_actual_external_abstraction_cumulative[r] / dt;
  }
  for (int i = tid; i < c.nres; i += stride) {
    f.res_outflow_average[i] = f.res_outflow_cumulative[i] / dt;
    f.res_inflow_average[i] = f.res_inflow_cumulative[i] / dt;
    f.res_actual_external_abstraction_average[i] = f.res_actual_external_abstraction_cumulative[i] / dt;
  }
  if (tid == 0) *w.substeps = count;
}

int lil_max_grid(int device) {
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, local_inertial_land_river_kernel, kLiBlock, 0) !=
      cudaSuccess)
    return -1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}

int launch_local_inertial_land_river(const DevFields& f, const KCfg& c, const LiLaunch& w, cudaStream_t s) {
  static const unsigned long long inf4[4] = {0x7ff0000000000000ull, 0x7ff0000000000000ull,
                                             0x7ff0000000000000ull, 0x7ff0000000000000ull};
  cudaMemcpyAsync(w.dt_bits, inf4, sizeof(inf4), cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(w.barrier, 0, sizeof(unsigned), s);  // arrivals; the generation keeps counting
  local_inertial_land_river_kernel<<<w.grid, kLiBlock, 0, s>>>(f, c, w);
  return 1;
}

// update_bc_overland_flow_model!                              surface_staggered_scheme.jl:1080-1097
__global__ void bc_overland_flow_kernel(const DevFields f, const KCfg c) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;   // node order (the 2-D state's order)
  if (v >= c.n) return;
  const int p = f.land_slot_of_node[v];
  double runoff = (f.net_runoff[p] + f.net_runoff_river[p]) * f.area[p];
  if (f.lil_river_slot[v] >= 0) runoff += f.ssf_to_river_average[p];  // get_flux_to_river  lsf.jl:346
  f.li_land_runoff[v] = runoff;
}

int launch_bc_overland_flow(const DevFields& f, const KCfg& c, cudaStream_t s) {
  bc_overland_flow_kernel<<<(c.n + 255) / 256, 256, 0, s>>>(f, c);
  return 1;
}

int li_max_grid(int device) {
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, local_inertial_river_kernel, kLiBlock, 0) !=
      cudaSuccess)
    return -1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}

int launch_local_inertial_river(const DevFields& f, const KCfg& c, const LiLaunch& w, cudaStream_t s) {
  static const unsigned long long inf2[2] = {0x7ff0000000000000ull, 0x7ff0000000000000ull};
  cudaMemcpyAsync(w.dt_bits, inf2, sizeof(inf2), cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(w.barrier, 0, sizeof(unsigned), s);  // arrivals; the generation keeps counting
  local_inertial_river_kernel<<<w.grid, kLiBlock, 0, s>>>(f, c, w);
  return 1;
}

}  // namespace wfb
