// api.cu -- the C ABI of libwflow_b200.so (include/wflow_b200.h): handle lifetime, HBM layout,
// host<->device field transfer through the slot permutation, and the orchestration of the hot
// path (update_model!, sbm_model.jl:60-92). No CPU fallback: every entry point needs the device.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wflow_b200.h"
#include "kernels.cuh"
#include "model.cuh"
#include "network.hpp"

using namespace wfb;

namespace {

const char* kFieldNames[] = {
#define X(name, kind) #name,
    WFLOWB200_FIELDS(X)
#undef X
};
const int kFieldKinds[] = {
#define X(name, kind) kind,
    WFLOWB200_FIELDS(X)
#undef X
};
std::string g_create_error;

// Cut edges of a shard that is part of a basin (wflowb200_exchange_*). Kinds of values:
// 0 overland flow (land, 2 values per sub-step), 1 river flow (river, 1), 2 subsurface flow (land, 2).
struct ExchangeLayout {
  int64_t n_imp[2] = {0, 0};     // imports of the land / river domain
  size_t kind_off[3] = {0, 0, 0};
  size_t words = 0;              // one parity
  void set(int64_t n_land, int64_t n_river, const int (&S)[3]) {
    n_imp[0] = n_land; n_imp[1] = n_river;
    kind_off[0] = 0;
    kind_off[1] = kind_off[0] + (size_t)n_land * S[0] * 2;
    kind_off[2] = kind_off[1] + (size_t)n_river * S[1] * 1;
    words = kind_off[2] + (size_t)n_land * S[2] * 2;
  }
};
struct ExchangePeer {
  unsigned long long* base = nullptr;
  bool ipc = false;
  ExchangeLayout lay;
};
struct Exchange {
  bool prepared = false;         // exchange_prepare was called: the shard takes part in the barrier
  int64_t n_exp[2] = {0, 0};
  int S[3] = {0, 0, 0};
  ExchangeLayout lay;            // of this shard's import buffer
  unsigned long long* imp = nullptr;            // device: [2 parities][lay.words]
  std::map<int, ExchangePeer> peers;
  std::vector<std::pair<int, int64_t>> bound[2];  // per export: (peer, import of the peer), -1 unbound
  unsigned long long** d_exp[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  bool tables_current = false;
  uint64_t step = 0;
  bool in_update_model = false;
  unsigned long long* d_token = nullptr;  // the barrier's all-reduced word
  bool active() const { return lay.n_imp[0] + lay.n_imp[1] + n_exp[0] + n_exp[1] > 0; }
};

struct DomainDev {
  Network nw;
  int32_t* node_of_slot = nullptr;  // device, 0-based node id per slot
  std::vector<void*> dev_arrays;    // everything uploaded for DevNet (freed together)
  unsigned long long* q_out = nullptr;  // per chunk x S x NV published outlet values
  size_t q_out_words = 0;
  DevNet dev{};
  int32_t* chunk_of_slot = nullptr;     // device, land only: slot -> chunk (fused surface kernel)
  unsigned* chunk_done = nullptr;       // device, land only: per chunk, epoch of its last finalize
  unsigned* chunk_ssf_done = nullptr;   // the same for subsurface flow + soil water storage
};

// NetworkLand + NetworkRiver artefacts (network.jl:87-133,214-278; domain.jl:80-125).
int32_t build_networks(const WflowB200Config* cfg, const WflowB200Domain* dom, Network& land,
                       Network& river, std::string& errmsg) {
  std::string err;
  const int64_t n = cfg->n, nriv = cfg->nriv;
  if (!build_graph(land, dom->d1, dom->d2, dom->indices, dom->ldd, n, err) ||
      !build_artifacts(land, cfg->nthreads, cfg->land_streamorder_min, nullptr, err)) {
    errmsg = "land network: " + err;
    return WFLOWB200_ERR_GRAPH;
  }
  if (cfg->land_routing == 1 && !build_edge_connectivity(land, dom->d1, dom->d2, dom->indices, n, err)) {
    errmsg = "land network: " + err;
    return WFLOWB200_ERR_GRAPH;
  }
  std::vector<int64_t> ridx(2 * (size_t)nriv), rso(nriv);
  std::vector<uint8_t> rldd(nriv);
  for (int64_t r = 0; r < nriv; ++r) {
    const int64_t li = dom->river_land_indices[r];
    if (li < 1 || li > n || (r && li <= dom->river_land_indices[r - 1])) {
      errmsg = "river_land_indices must be ascending 1-based land indices";
      return WFLOWB200_ERR_ARG;
    }
    ridx[2 * r] = dom->indices[2 * (li - 1)];
    ridx[2 * r + 1] = dom->indices[2 * (li - 1) + 1];
    rldd[r] = land.ldd[li - 1];
    rso[r] = land.streamorder[li - 1];  // network.jl:245
  }
  if (!build_graph(river, dom->d1, dom->d2, ridx.data(), rldd.data(), nriv, err) ||
      !build_artifacts(river, cfg->nthreads, cfg->river_streamorder_min, rso.data(), err)) {
    errmsg = "river network: " + err;
    return WFLOWB200_ERR_GRAPH;
  }
  // Shallow multi-piece chunks cut the warp-stages of a sweep (24 sub-steps: -31 % at depth 4,
  // one sub-step: -72 %) but add chunk-to-chunk hand-offs to the dependent chain of the sweep.
  // Measured on B200: a loss on a latency-bound domain (1000^2: 1000 nodes per level, +16 % step
  // time at depth 6), a gain on a throughput-bound one, the more the wider its levels are (a
  // chunk of depth d keeps its lanes busy S / (S + d - 1) of its stages: 1/d for the one sub-step
  // of the subsurface flow). Step times by depth: 2500 nodes per level 6: 11.6 ms, 4: 12.2;
  // 3536: 6: 21.3, 5: 20.4, 4: 20.4, 3: 21.3; 5000: 4: 14.9, 3: 15.1; 10000 (1250 levels):
  // 6: 20.3, 5: 19.2, 4: 18.1, 3: 17.3, 2: 16.8, 1: 23.6. So the depth follows the mean level width.
  int64_t depth_land = WFB_PIECE_DEPTH_LAND;
  if (land.n_wave_levels > 0) {
    const int64_t width = land.n / land.n_wave_levels;
    if (width >= 8500) depth_land = 2;
    else if (width >= 6000) depth_land = 3;
    else if (width >= 4500) depth_land = 4;
    else if (width >= 3000) depth_land = 5;
    else if (width >= WFB_PIECE_WIDE_LEVEL) depth_land = WFB_PIECE_DEPTH_LAND_WIDE;
  }
  if (cfg->wave_piece_depth_land > 0) depth_land = cfg->wave_piece_depth_land;
  else if (cfg->wave_piece_depth_land < 0) depth_land = 0;
  build_chunks(land, WFB_CHUNK_NODES, depth_land);
  build_chunks(river, WFB_CHUNK_NODES, WFB_PIECE_DEPTH_RIVER);
  return WFLOWB200_OK;
}

int32_t copy_artifact(const Network& nw, int32_t id, int64_t* dst, int64_t capacity,
                      int64_t* len_out, std::string& errmsg) {
  std::vector<int64_t> tmp;
  const std::vector<int64_t>* src = nullptr;
  switch (id) {
    case WFLOWB200_A_ORDER: src = &nw.order; break;
    case WFLOWB200_A_STREAMORDER: src = &nw.streamorder; break;
    case WFLOWB200_A_UPSTREAM_PTR: src = &nw.up_ptr; break;
    case WFLOWB200_A_UPSTREAM_IDX: src = &nw.up_idx; break;
    case WFLOWB200_A_SUBDOMAIN_LEVEL_PTR: src = &nw.lvl_ptr; break;
    case WFLOWB200_A_SUBDOMAIN_LEVEL_IDX: src = &nw.lvl_idx; break;
    case WFLOWB200_A_SUBDOMAIN_PTR: src = &nw.sub_ptr; break;
    case WFLOWB200_A_SUBDOMAIN_ORDER: src = &nw.sub_order; break;
    case WFLOWB200_A_SUBDOMAIN_INDICES: src = &nw.sub_indices; break;
    case WFLOWB200_A_LDD: tmp.assign(nw.ldd.begin(), nw.ldd.end()); src = &tmp; break;
    case WFLOWB200_A_WAVE_LEVEL_PTR: src = &nw.wave_level_ptr; break;
    case WFLOWB200_A_WAVE_PERM: src = &nw.perm; break;
    case WFLOWB200_A_WAVE_NODE_LEVEL: src = &nw.node_level; break;
    case WFLOWB200_A_WAVE_CHUNK_PTR: src = &nw.chunk_ptr; break;
    case WFLOWB200_A_WAVE_CHUNK_OUTLET: src = &nw.chunk_outlet; break;
    case WFLOWB200_A_EDGE_X_UP: src = &nw.edge_x_up; break;
    case WFLOWB200_A_EDGE_X_DOWN: src = &nw.edge_x_down; break;
    case WFLOWB200_A_EDGE_Y_UP: src = &nw.edge_y_up; break;
    case WFLOWB200_A_EDGE_Y_DOWN: src = &nw.edge_y_down; break;
    default: errmsg = "bad artefact id"; return WFLOWB200_ERR_ARG;
  }
  *len_out = (int64_t)src->size();
  if (dst) {
    if (capacity < *len_out) { errmsg = "artefact buffer too small"; return WFLOWB200_ERR_ARG; }
    memcpy(dst, src->data(), src->size() * sizeof(int64_t));
  }
  return WFLOWB200_OK;
}

}  // namespace

// ---- reductions across the shards of one model domain ------------------------------------------
// With adaptive internal time steps every routing sub-step needs a statistic of the WHOLE domain
// (surface_kinwave.jl:674-704: a type-7 quantile; lateral_subsurface_flow.jl:314-344: a minimum).
// A sharded domain reduces the partial results of its shards: over NCCL when every shard lives in
// its own process / GPU (wflowb200_comm_init_nccl), through a host rendezvous when the shards
// are handles of one process (wflowb200_group_*: tests, several shards on one GPU).
struct ShardComm {
  virtual ~ShardComm() {}
  // in-place all-reduce of n unsigned 64-bit words in device memory, ordered on s; op 0 sum, 1 min
  virtual int allreduce(unsigned long long* dev, int n, int op, cudaStream_t s) = 0;
};

struct NcclApi {  // resolved at run time: the library has no link-time dependency on NCCL
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
  }
};
static NcclApi g_nccl;

struct NcclShardComm : ShardComm {
  ncclComm_t comm = nullptr;
  ~NcclShardComm() override { if (comm) g_nccl.CommDestroy(comm); }
  int allreduce(unsigned long long* dev, int n, int op, cudaStream_t s) override {
    return g_nccl.AllReduce(dev, dev, (size_t)n, ncclUint64, op == 0 ? ncclSum : ncclMin, comm, s) ==
                   ncclSuccess ? 0 : -1;
  }
};

struct WflowB200Group {  // host rendezvous of the handles of one process
  std::mutex m;
  std::condition_variable cv;
  int n = 0, arrived = 0;
  unsigned long long generation = 0;
  std::vector<unsigned long long> acc, result[2];
};

struct GroupShardComm : ShardComm {
  WflowB200Group* g = nullptr;
  std::vector<unsigned long long> host;
  int allreduce(unsigned long long* dev, int n, int op, cudaStream_t s) override {
    host.resize(n);
    if (cudaMemcpyAsync(host.data(), dev, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s) !=
            cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
      return -1;
    {
      std::unique_lock<std::mutex> lk(g->m);
      if (g->arrived == 0) g->acc = host;
      else
        for (int k = 0; k < n; ++k)
          g->acc[k] = op == 0 ? g->acc[k] + host[k] : std::min(g->acc[k], host[k]);
      const unsigned long long gen = g->generation;
      if (++g->arrived == g->n) {
        g->result[gen & 1] = g->acc;
        g->arrived = 0;
        ++g->generation;
        g->cv.notify_all();
      } else {
        g->cv.wait(lk, [&] { return g->generation != gen; });
      }
      host = g->result[gen & 1];
    }
    return cudaMemcpyAsync(dev, host.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice, s) ==
                   cudaSuccess ? 0 : -1;
  }
};

struct WflowB200Network {
  Network land, river;
};

// Kernel organisations that are chosen from the shape of the domain (mean level width) and can be
// overridden per handle through wflowb200_set_option (the parity tests exercise every one).
struct Tuning {
  int fuse_soil_storage = -1; // -1 automatic: update_soil_water_storage! inside the subsurface sweep
  int overlap_ssf = -1;       // -1 automatic: subsurface sweep and surface kernel overlapped
  int ssf_overlap_sms = 0;    // 0 = 13/32 of the SMs
  int fuse_surface = 1;       // overland + river in one kernel
  int grid_percent = 100;     // share of the co-resident capacity the wavefront kernels may take
                              // (several handles with cut edges on ONE GPU must all be resident)
  int river_share = 1, river_period = 3;  // warps of the surface kernel serving the river
  int use_graph = 1;          // vertical update as one CUDA graph launch
  int run_engine = 1;         // 0: skip the loop engine (timing experiments; results invalid)
  int timeline = 0;           // record completion events of the vertical kernels (graph off)
};

struct WflowB200 {
  WflowB200Config cfg{};
  Tuning tune{};
  int n = 0, nriv = 0, N = 0, ns = 0, nrs = 0, nres = 0, nress = 0;
  int32_t* res_ident = nullptr;      // identity slot map of the reservoir fields
  int32_t* land_ident = nullptr;     // ... and of the node-ordered li_land_* fields (kind 6)
  DomainDev land, river;
  DomainDev land_full;               // device arrays only (nw unused): land chunks without the
  DomainDev* snow_net = nullptr;     // reservoir cut, for lateral snow transport
  int grid_snow = 0;
  size_t smem_snow = 0;
  int grid_li = 0;                   // local-inertial river flow: co-resident CTAs
  unsigned* d_li_barrier = nullptr;  // {arrivals, generation}
  unsigned long long* d_li_dt = nullptr;
  int* d_li_substeps = nullptr;
  int grid_lil = 0;                  // 2-D local-inertial overland + river flow: co-resident CTAs
  DevFields f{};
  KCfg kc{};
  double* pool = nullptr;  // one HBM allocation holding every Float64 field
  size_t pool_doubles = 0;
  std::vector<double*> field_ptr;
  int32_t* riv_of_land = nullptr;
  double* d_stage = nullptr;  // device staging for set/get (n*(N+1) doubles, >= 3n)
  size_t stage_doubles = 0;
  double* h_pinned = nullptr;  // pinned host staging for forcing (3n doubles)
  double* d_forcing = nullptr; // device staging for forcing (3n doubles)
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  ShardComm* comm = nullptr;   // reductions across the shards of the domain (adaptive mode)
  Exchange xchg;               // cut edges
  cudaEvent_t forcing_ready = nullptr, forcing_consumed = nullptr;
  bool forcing_pending = false;
  // staging ring of forcing slabs in HBM (wflowb200_forcing_ring_*): the host uploads the slabs of
  // the coming steps on the copy stream while the current step computes
  double* d_ring = nullptr;
  int ring_depth = 0, ring_use = -1;
  std::vector<cudaEvent_t> ring_ready;     // per slot: its H2D copy has landed
  std::vector<cudaEvent_t> ring_consumed;  // per slot: the gather that read it has run
  // cyclic parameters staged once (io.jl:187-227): leaf_area_index [slabs x n], node order
  double* d_lai_table = nullptr;
  int lai_slabs = 0;
  double* h_out_pinned = nullptr;          // pinned staging of wflowb200_get_fields
  size_t out_pinned_doubles = 0;
  double* d_out[2] = {nullptr, nullptr};   // device staging of the asynchronous output gather
  size_t out_doubles[2] = {0, 0};
  int out_parity = 0;
  cudaEvent_t out_gathered = nullptr, out_copied[2] = {nullptr, nullptr};
  unsigned* d_queue = nullptr;
  UnsatWork unsat{};                 // scratch of the unsaturated-zone engine
  double* d_unsat_pool = nullptr;
  int32_t* d_unsat_its = nullptr;
  int32_t* d_unsat_list = nullptr;
  unsigned* d_err = nullptr;         // device error word (bounded waits of the wavefront kernels)
  unsigned* d_unsat_count = nullptr;
  VerticalStage v_stage{};           // input staging lists of the dense vertical kernels
  int engine_grid = 0;
  cudaEvent_t tl_ev[4] = {};         // timeline of the vertical kernels (developer aid)
  cudaGraphExec_t v_graph = nullptr;  // the vertical update of one step, captured once per dt
  double v_graph_dt = 0.0;
  int v_graph_launches = 0;
  RoutingStats* d_stats = nullptr;
  unsigned long long* d_count = nullptr;
  double* d_min = nullptr;
  double* d_work = nullptr;                 // adaptive sub-stepping: per-node stable time steps
  unsigned long long* d_qstate = nullptr;   // state of the quantile select
  int grid_olf = 0, grid_riv = 0, grid_ssf = 0;
  int grid_surface = 0;                // overland + river in one kernel
  size_t smem_surface = 0;
  unsigned smem_surface_per_warp = 0;
  bool fuse_surface = true;
  unsigned surface_epoch = 0;
  unsigned long long* ssf_q_out = nullptr;  // outlet values of the subsurface warps (fused routing)
  size_t ssf_q_out_words = 0;
  int64_t launches = 0;
  int64_t sub_land = 0, sub_river = 0, sub_ssf = 0;
  bool timing = false;
  size_t smem_olf = 0, smem_riv = 0, smem_ssf = 0;
  cudaEvent_t ev[12] = {};
  cudaEvent_t tm[2] = {};
  double ms[7] = {};
  int64_t timed_steps = 0;
  std::string err;
};

namespace {

constexpr int kMaxSub = 8192;

int32_t fail(WflowB200* h, int32_t code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
#define CUDA_TRY(h, expr)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (expr);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(h, WFLOWB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// Every entry point runs on the handle's device whatever the caller's current device is (a Julia
// task may migrate between threads, each with its own current device), and restores it.
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define WFB_ENTER(h)                     \
  if (!(h)) return WFLOWB200_ERR_ARG;    \
  DeviceGuard device_guard_((h)->cfg.device)

// The wavefront kernels raise the error word when one of their bounded waits expired
// (routing.cu: spin_expired): the grid was not co-resident and the step's results are invalid.
int32_t check_device_error(WflowB200* h) {
  unsigned e = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&e, h->d_err, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (e != 0)
    return fail(h, WFLOWB200_ERR_STATE,
                "a wavefront wait timed out: the routing grid was not co-resident on the device "
                "(MPS share, second context or debugger?); the results of this step are invalid");
  return WFLOWB200_OK;
}

// device slot -> host element of a field kind: the drainage-order permutation of the land / river
// fields, the identity for reservoirs and for the node-ordered 2-D overland-flow fields (kind 6)
const int32_t* slot_map_of(const WflowB200* h, int kind) {
  if (kind == 4) return h->res_ident;
  if (kind == 6) return h->land_ident;
  return ((kind == 3 || kind == 5) ? h->river : h->land).node_of_slot;
}
int layers_of(const WflowB200* h, int kind) {
  return kind == 1 ? h->N : kind == 2 ? h->N + 1 : kind == 5 ? std::max(h->cfg.fp_levels, 1) : 1;
}
// slots / elements / slot map of a field kind (0-2 land, 3 river, 4 reservoir, 5 river x level)
// kind 6: land scalars of the 2-D local-inertial overland flow, present with land_routing = 1 only
int slots_of(const WflowB200* h, int kind) {
  if (kind == 6) return h->cfg.land_routing == 1 ? h->ns : 0;
  return (kind == 3 || kind == 5) ? h->nrs : kind == 4 ? h->nress : h->ns;
}
int count_of(const WflowB200* h, int kind) {
  if (kind == 6) return h->cfg.land_routing == 1 ? h->n : 0;
  return (kind == 3 || kind == 5) ? h->nriv : kind == 4 ? h->nres : h->n;
}

template <class T>
cudaError_t upload_i32(const std::vector<T>& src, int32_t** dst, int64_t offset) {
  std::vector<int32_t> tmp(src.size());
  for (size_t i = 0; i < src.size(); ++i) tmp[i] = (int32_t)(src[i] + offset);
  cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(tmp.size(), 1) * sizeof(int32_t));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, tmp.data(), tmp.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
}

template <class T>
cudaError_t upload_raw(const std::vector<T>& src, const T** dst, std::vector<void*>& owned) {
  T* ptr = nullptr;
  cudaError_t e = cudaMalloc((void**)&ptr, std::max<size_t>(src.size(), 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  owned.push_back(ptr);
  *dst = ptr;
  return cudaMemcpy(ptr, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// Device copies of the wavefront artefacts of one domain.
// outlet_res[v] (by node id, may be empty): node v is a reservoir outlet (domain.jl:96-109).
// cut_outlets: edges leaving an outlet are dropped (the land kinematic waves: filter_upstream_nodes,
// utils.jl:61-71); otherwise (river) they stay -- they carry the reservoir's outflow -- but
// move to the END of the receiving node's fold: the reference forms qin[v] = outflow first and
// then adds sum_at(q, upstream_nodes) (surface_kinwave.jl:481-482,517), and a + b == b + a.
// imp_dst / imp_pos / exp_src: the shard's cut edges of this domain (WflowB200Domain).
int32_t upload_domain(WflowB200* h, DomainDev& d, const Network& nw,
                      const std::vector<uint8_t>& outlet_res, bool cut_outlets,
                      const std::vector<int64_t>& imp_dst = {}, const std::vector<int64_t>& imp_pos = {},
                      const std::vector<int64_t>& exp_src = {}) {
  const int64_t n = nw.n;
  // imports by receiving node, ascending position
  std::vector<std::vector<std::pair<int64_t, int64_t>>> imports_of;  // node -> (pos, import)
  if (!imp_dst.empty()) {
    imports_of.resize(n);
    for (size_t k = 0; k < imp_dst.size(); ++k) {
      if (imp_dst[k] < 1 || imp_dst[k] > n || imp_pos[k] < 0 || imp_pos[k] > 7)
        return fail(h, WFLOWB200_ERR_ARG, "import: bad node or position");
      imports_of[imp_dst[k] - 1].emplace_back(imp_pos[k], (int64_t)k);
    }
    for (auto& l : imports_of) std::sort(l.begin(), l.end());
  }
  CUDA_TRY(h, upload_i32(nw.perm, &d.node_of_slot, -1));
  std::vector<int4> meta(nw.n_chunks);
  std::vector<unsigned long long> node_edges(n, ~0ull);
  std::vector<uint8_t> node_level(n), inl_level;
  std::vector<int32_t> inl_src, node_out(n, -1);
  int64_t max_inlets = 0;
  for (int64_t c = 0; c < nw.n_chunks; ++c) {
    const int64_t p0 = nw.chunk_ptr[c], nn = nw.chunk_ptr[c + 1] - p0;
    const int64_t nlev = nw.chunk_l1[c] - nw.chunk_l0[c] + 1;
    if (nn > WFB_CHUNK_NODES || nlev > WFB_CHUNK_NODES)
      return fail(h, WFLOWB200_ERR_STATE, "chunk exceeds WFB_CHUNK_NODES");
    const int64_t i0 = (int64_t)inl_src.size();
    int64_t k = 0;  // inlet number inside the chunk
    for (int64_t p = p0; p < p0 + nn; ++p) {
      const int64_t v = nw.perm[p] - 1;  // in-neighbours are already ascending by node id
      node_level[p] = (uint8_t)(nw.node_level[v] - nw.chunk_l0[c]);
      node_out[p] = (int32_t)nw.out_of_node[v];
      const int64_t deg = nw.in_ptr[v + 1] - nw.in_ptr[v];
      const int64_t n_imp_v = imports_of.empty() ? 0 : (int64_t)imports_of[v].size();
      if (deg + n_imp_v > 8) return fail(h, WFLOWB200_ERR_GRAPH, "node with more than 8 upstream nodes");
      unsigned long long code = ~0ull;
      int64_t srcs[8];   // node id, or -(import + 1)
      int64_t ns_ = 0;
      for (int pass = 0; pass < 2; ++pass)   // ordinary sources, then reservoir outlets
        for (int64_t e = 0; e < deg; ++e) {
          const int64_t u = nw.in_idx[nw.in_ptr[v] + e] - 1;
          const bool is_res = !outlet_res.empty() && outlet_res[u];
          if (is_res != (pass == 1) || (is_res && cut_outlets)) continue;
          srcs[ns_++] = u;
        }
      for (int64_t q = 0; q < n_imp_v; ++q) {  // a cut edge takes its place in the global order
        const int64_t pos = imports_of[v][q].first;
        if (pos > ns_) return fail(h, WFLOWB200_ERR_ARG, "import: position beyond the node's sources");
        for (int64_t e = ns_; e > pos; --e) srcs[e] = srcs[e - 1];
        srcs[pos] = -(imports_of[v][q].second + 1);
        ++ns_;
      }
      for (int64_t e = 0; e < ns_; ++e) {
        const int64_t u = srcs[e];
        if (u < 0) {  // import: an inlet edge fed by another GPU
          if (WFB_CHUNK_NODES + k >= (int64_t)WFB_NO_EDGE)
            return fail(h, WFLOWB200_ERR_STATE, "too many inlet edges in one chunk");
          const unsigned long long b = (unsigned long long)(WFB_CHUNK_NODES + k++);
          inl_level.push_back(node_level[p]);
          inl_src.push_back((int32_t)(nw.n_outlets + (-u - 1)));
          code = (code & ~(0xffull << (8 * e))) | (b << (8 * e));
          continue;
        }
        const int64_t cu = nw.chunk_of_node[u];
        unsigned long long b;
        if (cu != c) {
          if (WFB_CHUNK_NODES + k >= (int64_t)WFB_NO_EDGE)
            return fail(h, WFLOWB200_ERR_STATE, "too many inlet edges in one chunk");
          b = (unsigned long long)(WFB_CHUNK_NODES + k++);
          inl_level.push_back(node_level[p]);
          if (nw.out_of_node[u] < 0)
            return fail(h, WFLOWB200_ERR_STATE, "inlet edge from a node that does not publish");
          inl_src.push_back((int32_t)nw.out_of_node[u]);
        } else {
          b = (unsigned long long)(nw.slot_of[u] - p0);
        }
        code = (code & ~(0xffull << (8 * e))) | (b << (8 * e));
      }
      node_edges[p] = code;
    }
    max_inlets = std::max(max_inlets, k);
    meta[c] = make_int4((int)p0, (int)(nn | (nlev << 8)), (int)i0, (int)k);
  }
  CUDA_TRY(h, upload_raw(meta, &d.dev.chunk_meta, d.dev_arrays));
  CUDA_TRY(h, upload_raw(node_edges, &d.dev.node_edges, d.dev_arrays));
  CUDA_TRY(h, upload_raw(node_level, &d.dev.node_level, d.dev_arrays));
  CUDA_TRY(h, upload_raw(node_out, &d.dev.node_out, d.dev_arrays));
  d.dev.n_outlets = (int32_t)nw.n_outlets;
  CUDA_TRY(h, upload_raw(inl_src, &d.dev.inl_src, d.dev_arrays));
  CUDA_TRY(h, upload_raw(inl_level, &d.dev.inl_level, d.dev_arrays));
  d.dev.node_export = nullptr;
  if (!exp_src.empty()) {
    std::vector<int32_t> node_export(n, -1);
    for (size_t k = 0; k < exp_src.size(); ++k) {
      const int64_t v = exp_src[k];
      if (v < 1 || v > n || nw.down[v - 1] != 0 || node_export[nw.slot_of[v - 1]] >= 0)
        return fail(h, WFLOWB200_ERR_ARG, "export: the node must be a distinct pit of the shard's ldd");
      node_export[nw.slot_of[v - 1]] = (int32_t)k;
    }
    CUDA_TRY(h, upload_raw(node_export, &d.dev.node_export, d.dev_arrays));
  }
  {
    std::vector<int64_t> cos(n);
    for (int64_t p = 0; p < n; ++p) cos[p] = nw.chunk_of_node[nw.perm[p] - 1];
    CUDA_TRY(h, upload_i32(cos, &d.chunk_of_slot, 0));
    CUDA_TRY(h, cudaMalloc((void**)&d.chunk_done, std::max<size_t>(nw.n_chunks, 1) * sizeof(unsigned)));
    CUDA_TRY(h, cudaMemset(d.chunk_done, 0, std::max<size_t>(nw.n_chunks, 1) * sizeof(unsigned)));
    CUDA_TRY(h, cudaMalloc((void**)&d.chunk_ssf_done, std::max<size_t>(nw.n_chunks, 1) * sizeof(unsigned)));
    CUDA_TRY(h, cudaMemset(d.chunk_ssf_done, 0, std::max<size_t>(nw.n_chunks, 1) * sizeof(unsigned)));
  }
  d.dev.n = (int32_t)n;
  d.dev.n_levels = (int32_t)nw.n_wave_levels;
  d.dev.n_chunks = (int32_t)nw.n_chunks;
  d.dev.max_inlets = (int32_t)max_inlets;
  return WFLOWB200_OK;
}

void free_domain(DomainDev& d) {
  cudaFree(d.chunk_of_slot);
  cudaFree(d.chunk_done);
  cudaFree(d.chunk_ssf_done);
  cudaFree(d.node_of_slot);
  for (void* p : d.dev_arrays) cudaFree(p);
  cudaFree(d.q_out);
}

// The reference's `while t < dt` sub-stepping with a fixed internal step
// (surface_kinwave.jl:371-379; routing/timestepping.jl:11-16), evaluated on the host.
int fixed_substeps(double dt, double dt_fixed, std::vector<double>& out) {
  out.clear();
  if (!(dt_fixed > 0.0)) return -1;
  double t = 0.0;
  while (t < dt) {
    double dt_s = dt_fixed;
    if (t + dt_s > dt) dt_s = dt - t;
    out.push_back(dt_s);
    t += dt_s;
    if ((int)out.size() > kMaxSub) return -1;
  }
  return (int)out.size();
}

int32_t wait_forcing(WflowB200* h) {
  if (h->ring_use >= 0) {  // a slab of the staging ring
    const int k = h->ring_use;
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ring_ready[k], 0));
    h->launches += launch_gather_forcing(h->f, h->d_ring + (size_t)k * 3 * h->n,
                                         h->land.node_of_slot, h->n, h->stream);
    CUDA_TRY(h, cudaEventRecord(h->ring_consumed[k], h->stream));
    h->ring_use = -1;
    h->forcing_pending = false;
  }
  if (h->forcing_pending) {
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->forcing_ready, 0));
    h->launches += launch_gather_forcing(h->f, h->d_forcing, h->land.node_of_slot, h->n, h->stream);
    CUDA_TRY(h, cudaEventRecord(h->forcing_consumed, h->stream));
    h->forcing_pending = false;
  }
  return WFLOWB200_OK;
}

int32_t check_launch(WflowB200* h, int rc, const char* what) {
  if (rc < 0) return fail(h, WFLOWB200_ERR_CUDA, std::string(what) + ": launch failed (" +
                                                     std::to_string(rc) + ")");
  h->launches += rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(h, WFLOWB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return WFLOWB200_OK;
}

// Cut edges: this step's import slots and export targets of one component.
int32_t bind_exchange(WflowB200* h, int kind, WaveLaunch& w) {
  w.imports = nullptr;
  w.exports = nullptr;
  Exchange& x = h->xchg;
  if (!x.active()) return WFLOWB200_OK;
  if (!x.prepared || !x.tables_current || !x.in_update_model)
    return fail(h, WFLOWB200_ERR_STATE,
                "a shard with cut edges runs through wflowb200_update_model after "
                "wflowb200_exchange_prepare / open_peer / bind (every export bound)");
  if (w.S != x.S[kind])
    return fail(h, WFLOWB200_ERR_ARG, "cut edges: the time step differs from wflowb200_exchange_prepare");
  const int par = (int)(x.step & 1);
  if (x.imp) w.imports = x.imp + (size_t)par * x.lay.words + x.lay.kind_off[kind];
  w.exports = x.d_exp[kind][par];
  return WFLOWB200_OK;
}

// Run one routing component with the fixed-step skewed wavefront. kind: 0 overland, 1 river,
// 2 subsurface; nv = values published per node and sub-step.
template <class Launch>
int32_t run_wave(WflowB200* h, DomainDev& d, double dt, double dt_fixed, int kind, int nv,
                 int max_grid, size_t smem, int64_t& substeps, Launch launch, const char* what,
                 double dt_single = 0.0, bool accumulate = false, bool fuse_soil_storage = false) {
  std::vector<double> dts;
  int S;
  if (dt_single > 0.0) { dts.assign(1, dt_single); S = 1; }  // one adaptive sub-step
  else S = fixed_substeps(dt, dt_fixed, dts);
  if (S <= 0) return fail(h, WFLOWB200_ERR_ARG, std::string(what) + ": bad internal time step");
  const size_t need = (size_t)std::max<int64_t>(d.nw.n_outlets, 1) * (size_t)S * (size_t)nv;
  if (need > d.q_out_words) {  // published values of every sub-step, per outlet
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(d.q_out);
    d.q_out = nullptr;
    d.q_out_words = 0;
    CUDA_TRY(h, cudaMalloc((void**)&d.q_out, need * sizeof(unsigned long long)));
    d.q_out_words = need;
  }
  WaveLaunch w{};
  w.queue = h->d_queue + kind * 32;
  w.q_out = d.q_out;
  w.stats = h->d_stats;
  w.S = S;
  w.dt_fixed = dts[0];
  w.dt_last = dts[S - 1];
  w.dt = dt;
  w.accumulate = accumulate ? 1 : 0;
  w.fuse_soil_storage = fuse_soil_storage ? 1 : 0;
  w.grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)max_grid * h->tune.grid_percent / 100, d.nw.n_chunks));
  w.smem = smem;
  w.err = h->d_err;
  int32_t rc = bind_exchange(h, kind, w);
  if (rc) return rc;
  rc = check_launch(h, launch(w), what);
  if (rc) return rc;
  substeps = S;
  return WFLOWB200_OK;
}

// The sub-step plan and buffers of one component, without launching it (fused kernels).
int32_t prepare_wave(WflowB200* h, DomainDev& d, double dt, double dt_fixed, int kind, int nv,
                     WaveLaunch& w, int64_t& substeps, const char* what) {
  std::vector<double> dts;
  const int S = fixed_substeps(dt, dt_fixed, dts);
  if (S <= 0) return fail(h, WFLOWB200_ERR_ARG, std::string(what) + ": bad internal time step");
  const size_t need = (size_t)std::max<int64_t>(d.nw.n_outlets, 1) * (size_t)S * (size_t)nv;
  if (need > d.q_out_words) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(d.q_out);
    d.q_out = nullptr;
    d.q_out_words = 0;
    CUDA_TRY(h, cudaMalloc((void**)&d.q_out, need * sizeof(unsigned long long)));
    d.q_out_words = need;
  }
  w = WaveLaunch{};
  w.queue = h->d_queue + kind * 32;
  w.q_out = d.q_out;
  w.stats = h->d_stats;
  w.S = S;
  w.dt_fixed = dts[0];
  w.dt_last = dts[S - 1];
  w.dt = dt;
  w.err = h->d_err;
  substeps = S;
  return bind_exchange(h, kind, w);
}

// Adaptive internal time stepping (kinematic_wave__adaptive_time_step_flag): the reference's
// `while t < dt` loop (surface_kinwave.jl:371-379,640-655; lateral_subsurface_flow.jl:289-299).
// Every sub-step needs a statistic of the whole domain's state after the previous one (a
// type-7 quantile of the Courant steps for the surface components, their minimum for the
// subsurface), so sub-steps cannot be pipelined through the levels: one wavefront launch per
// sub-step, the statistic reduced on the device and read back (8-32 bytes) to drive the loop.
template <class Launch>
int32_t run_wave_adaptive(WflowB200* h, DomainDev& d, double dt, int kind, int nv, int max_grid,
                          size_t smem, int64_t& substeps, Launch launch, const char* what) {
  double t = 0.0;
  int64_t count = 0;
  while (t < dt) {
    double dt_s;
    if (kind == 2) {  // stable_timestep(::LateralSSF)  lateral_subsurface_flow.jl:314-344
      int32_t rc = check_launch(h, launch_stable_timestep_ssf(h->f, h->kc, h->d_min, h->d_count, h->stream), what);
      if (rc) return rc;
      if (h->comm && (h->comm->allreduce((unsigned long long*)h->d_min, 1, 1, h->stream) ||
                      h->comm->allreduce(h->d_count, 1, 0, h->stream)))
        return fail(h, WFLOWB200_ERR_CUDA, std::string(what) + ": shard reduction failed");
      double mn = 0.0;
      unsigned long long k = 0;
      CUDA_TRY(h, cudaMemcpyAsync(&mn, h->d_min, 8, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaMemcpyAsync(&k, h->d_count, 8, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      dt_s = (k == 0 ? 0.5 : mn) * h->cfg.ssf_alpha_coefficient;
    } else {          // stable_timestep (surface)       surface_kinwave.jl:674-704
      const bool riv = kind == 1;
      const int n = riv ? h->nriv : h->n;
      int32_t rc = check_launch(h, launch_stable_timesteps_surface(
                                       riv ? h->f.riv_q : h->f.olf_q, riv ? h->f.riv_alpha : h->f.olf_alpha,
                                       riv ? h->f.riv_flow_length : h->f.flow_length, n, h->d_work,
                                       h->d_count, h->stream), what);
      if (rc) return rc;
      bool reduce_failed = false;
      const ShardReduce reduce = [&](unsigned long long* p, int cnt, int op) {
        const int e = h->comm->allreduce(p, cnt, op, h->stream);
        reduce_failed |= e != 0;
        return e;
      };
      rc = check_launch(h, launch_quantile7(h->d_work, h->d_count, n, riv ? 0.05 : 0.02, h->d_qstate,
                                            h->stream, h->comm ? &reduce : nullptr), what);
      if (rc) return rc;
      if (reduce_failed) return fail(h, WFLOWB200_ERR_CUDA, std::string(what) + ": shard reduction failed");
      unsigned long long st[11];
      CUDA_TRY(h, cudaMemcpyAsync(st, h->d_qstate, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      double a, b, g;
      memcpy(&a, &st[1], 8); memcpy(&b, &st[2], 8); memcpy(&g, &st[3], 8);
      if (st[10] == 0) dt_s = 600.0;       // k of ALL shards
      else if (st[10] == 1) dt_s = a;
      else dt_s = (std::isfinite(a) && std::isfinite(b)) ? a + g * (b - a) : (1.0 - g) * a + g * b;
    }
    if (!(dt_s > 0.0)) return fail(h, WFLOWB200_ERR_STATE, std::string(what) + ": stable time step is not positive");
    if (t + dt_s > dt) dt_s = dt - t;  // check_timestepsize  routing/timestepping.jl:11-16
    int64_t one = 0;
    int32_t rc = run_wave(h, d, dt, 0.0, kind, nv, max_grid, smem, one, launch, what, dt_s, count > 0);
    if (rc) return rc;
    t += dt_s;
    if (++count > kMaxSub) return fail(h, WFLOWB200_ERR_STATE, std::string(what) + ": too many adaptive sub-steps");
  }
  substeps = count;
  return WFLOWB200_OK;
}

}  // namespace

// update_land_hydrology_model! as ONE graph launch (a memset and three kernels), captured once
// per time step length.
static int32_t launch_vertical(WflowB200* h, double dt) {
  const bool transport = h->cfg.snow_gravitational_transport != 0;
  auto issue_phase = [&](int phase) {
    return launch_land_hydrology(h->f, h->kc, h->N, dt, h->unsat, h->v_stage, h->engine_grid, phase,
                                 h->tune.run_engine != 0, h->stream,
                                 (h->tune.timeline && !h->tune.use_graph) ? h->tl_ev : nullptr);
  };
  if (transport) {
    // interception + snow, lateral_snow_transport! over the land network (sbm.jl:98-100), then
    // the rest of the update (below: as a graph like the one-phase update)
    int32_t rc = check_launch(h, issue_phase(1), "update_land_hydrology_model (snow)");
    if (rc) return rc;
    int64_t one = 0;
    DomainDev& sd = *h->snow_net;
    const size_t need = (size_t)std::max<int64_t>(h->land.nw.n_outlets, 1) * 3;
    if (need > h->land.q_out_words) {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      cudaFree(h->land.q_out);
      h->land.q_out = nullptr; h->land.q_out_words = 0;
      CUDA_TRY(h, cudaMalloc((void**)&h->land.q_out, need * sizeof(unsigned long long)));
      h->land.q_out_words = need;
    }
    WaveLaunch w{};
    w.queue = h->d_queue + 2 * 32;
    w.q_out = h->land.q_out;
    w.stats = h->d_stats;
    w.S = 1;
    w.dt_fixed = w.dt_last = w.dt = dt;
    w.grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->grid_snow, h->land.nw.n_chunks));
    w.smem = h->smem_snow;
    w.err = h->d_err;
    (void)one;
    rc = check_launch(h, launch_snow_transport(h->f, h->kc, sd.dev, w, h->stream), "lateral_snow_transport");
    if (rc) return rc;
  }
  auto issue = [&]() { return issue_phase(transport ? 2 : 0); };
  if (!h->tune.use_graph) return check_launch(h, issue(), "update_land_hydrology_model");
  if (!h->v_graph || h->v_graph_dt != dt) {
    if (h->v_graph) { cudaGraphExecDestroy(h->v_graph); h->v_graph = nullptr; }
    cudaGraph_t g = nullptr;
    CUDA_TRY(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = issue();
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    if (rc < 0 || e != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      return fail(h, WFLOWB200_ERR_CUDA, std::string("update_land_hydrology_model: graph capture failed: ") +
                                             cudaGetErrorString(e));
    }
    e = cudaGraphInstantiate(&h->v_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
      h->v_graph = nullptr;
      return fail(h, WFLOWB200_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    }
    h->v_graph_dt = dt;
    h->v_graph_launches = rc;
  }
  CUDA_TRY(h, cudaGraphLaunch(h->v_graph, h->stream));
  h->launches += h->v_graph_launches;
  return WFLOWB200_OK;
}

extern "C" {

int32_t wflowb200_num_fields(void) { return WFLOWB200_NUM_FIELDS; }
const char* wflowb200_field_name(int32_t id) {
  return (id >= 0 && id < WFLOWB200_NUM_FIELDS) ? kFieldNames[id] : nullptr;
}
int32_t wflowb200_field_kind(int32_t id) {
  return (id >= 0 && id < WFLOWB200_NUM_FIELDS) ? kFieldKinds[id] : -1;
}
int32_t wflowb200_field_id(const char* name) {
  for (int i = 0; i < WFLOWB200_NUM_FIELDS; ++i)
    if (!strcmp(name, kFieldNames[i])) return i;
  return -1;
}
const char* wflowb200_last_error(const WflowB200* h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

int32_t wflowb200_create(const WflowB200Config* cfg, const WflowB200Domain* dom, WflowB200** out) {
  if (!cfg || !dom || !out) return fail(nullptr, WFLOWB200_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->n <= 0 || cfg->nriv < 0 || cfg->n_layers < 1 || cfg->n_layers > 8)
    return fail(nullptr, WFLOWB200_ERR_ARG, "n > 0, nriv >= 0, 1 <= n_layers <= 8 required");
  if (cfg->n > 0x7fffff00LL) return fail(nullptr, WFLOWB200_ERR_ARG, "n exceeds int32 slots");
  if (cfg->kv_profile < 0 || cfg->kv_profile > 3)
    return fail(nullptr, WFLOWB200_ERR_ARG, "kv_profile must be 0 .. 3");
  if (cfg->snow_gravitational_transport && !cfg->snow)
    return fail(nullptr, WFLOWB200_ERR_ARG, "snow_gravitational_transport needs the snow model");
  if (cfg->land_routing != 0 &&
      (cfg->land_routing != 1 || cfg->river_routing != 1 || cfg->fp_levels > 0 ||
       cfg->snow_gravitational_transport))
    return fail(nullptr, WFLOWB200_ERR_ARG,
                "land_routing: 0 or 1; the 2-D local-inertial overland flow needs river_routing = 1 "
                "and runs without 1-D floodplain and lateral snow transport");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, WFLOWB200_ERR_CUDA, "no CUDA device: libwflow_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, WFLOWB200_ERR_CUDA, "no such CUDA device");
  DeviceGuard device_guard_(cfg->device);

  WflowB200* h = new WflowB200();
  h->cfg = *cfg;
  if (cfg->fp_levels < 0 || cfg->fp_levels > 16) {
    g_create_error = "fp_levels: 0 .. 16";
    delete h;
    return WFLOWB200_ERR_ARG;
  }
  h->n = (int)cfg->n; h->nriv = (int)cfg->nriv; h->N = cfg->n_layers;
  h->ns = (h->n + 31) / 32 * 32;
  h->nrs = (h->nriv + 31) / 32 * 32;
  h->nres = (int)(dom->reservoir_river_indices ? dom->nres : 0);
  h->nress = (h->nres + 31) / 32 * 32;
  if (h->nres < 0 || h->nres > h->nriv)
    return (delete h, fail(nullptr, WFLOWB200_ERR_ARG, "0 <= nres <= nriv required"));
  if (h->cfg.kin_wave_min_flow_qroot == 0.0) h->cfg.kin_wave_min_flow_qroot = std::pow(1e-30, 0.2);
  auto bail = [&](int32_t code) { g_create_error = h->err; wflowb200_destroy(h); return code; };

  // ---- indexing artefacts (host) ---------------------------------------------------------
  {
    int32_t rc = build_networks(cfg, dom, h->land.nw, h->river.nw, h->err);
    if (rc) return bail(rc);
  }

  // ---- HBM layout ------------------------------------------------------------------------
  size_t total = 0;
  std::vector<size_t> off(WFLOWB200_NUM_FIELDS);
  for (int i = 0; i < WFLOWB200_NUM_FIELDS; ++i) {
    off[i] = total;
    const int k = kFieldKinds[i];
    total += (size_t)slots_of(h, k) * layers_of(h, k);
  }
  h->pool_doubles = total;
#define TRY_CREATE(expr)                                                       \
  do {                                                                         \
    cudaError_t e_ = (expr);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      h->err = std::string(#expr) + ": " + cudaGetErrorString(e_);             \
      return bail(WFLOWB200_ERR_CUDA);                                         \
    }                                                                          \
  } while (0)
  TRY_CREATE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  TRY_CREATE(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  TRY_CREATE(cudaEventCreateWithFlags(&h->forcing_ready, cudaEventDisableTiming));
  TRY_CREATE(cudaEventCreateWithFlags(&h->forcing_consumed, cudaEventDisableTiming));
  for (auto& e : h->ev) TRY_CREATE(cudaEventCreate(&e));
  for (auto& e : h->tm) TRY_CREATE(cudaEventCreate(&e));
  TRY_CREATE(cudaMalloc((void**)&h->pool, total * sizeof(double)));
  h->field_ptr.resize(WFLOWB200_NUM_FIELDS);
  for (int i = 0; i < WFLOWB200_NUM_FIELDS; ++i) h->field_ptr[i] = h->pool + off[i];
  {
    int i = 0;
#define X(name, kind) h->f.name = h->field_ptr[i++];
    WFLOWB200_FIELDS(X)
#undef X
  }
  // MISSING_VALUE everywhere, then the reference's non-NaN defaults
  launch_fill(h->pool, (long long)total, NAN, h->stream);
  auto fill = [&](double* p, int kind, double v) {
    launch_fill(p, (long long)slots_of(h, kind) * layers_of(h, kind), v, h->stream);
  };
  fill(h->f.canopy_storage, 0, 0.0);            // canopy.jl:11
  fill(h->f.snow_in, 0, 0.0);                   // snow.jl:17-19
  fill(h->f.snow_out, 0, 0.0);
  fill(h->f.li_error, 3, 0.0);                  // surface_staggered_scheme.jl:181-185
  fill(h->f.li_zs_at_edge, 3, 0.0);
  fill(h->f.li_water_depth_at_edge, 3, 0.0);
  for (double* p : {h->f.fp_h, h->f.fp_storage, h->f.fp_q, h->f.fp_q_cumulative, h->f.fp_q_average,
                    h->f.fp_error, h->f.fp_water_depth_at_edge, h->f.riv_q_channel_average,
                    h->f.fp_flow_capacity, h->f.fp_qin, h->f.fp_qin_cumulative, h->f.fp_qin_average,
                    h->f.riv_floodplain_water_exchange})
    fill(p, 3, 0.0);                            // floodplain.jl:216-236
  if (cfg->land_routing == 1)                   // surface_staggered_scheme.jl:840-865,963-968
    for (double* p : {h->f.li_land_runoff, h->f.li_land_qx0, h->f.li_land_qy0, h->f.li_land_qx,
                      h->f.li_land_qy, h->f.li_land_qx_cumulative, h->f.li_land_qy_cumulative,
                      h->f.li_land_qx_average, h->f.li_land_qy_average, h->f.li_land_error})
      fill(p, 6, 0.0);
  fill(h->f.waterdepth_river, 0, 0.0);          // runoff.jl:26
  fill(h->f.unsaturated_store_depth, 0, 0.0);   // soil.jl:71
  fill(h->f.total_storage, 0, 0.0);             // soil.jl:77
  fill(h->f.f_infiltration_reduction, 0, 1.0);  // soil.jl:83
  fill(h->f.soil_surface_temperature, 0, 10.0 + 273.15);  // soil.jl:81
  {
    double* zeros_land[] = {h->f.ssf_exfiltwater_cumulative, h->f.ssf_exfiltwater_average,
                            h->f.ssf_q_cumulative, h->f.ssf_q_average, h->f.ssf_q_in_cumulative,
                            h->f.ssf_q_in_average, h->f.ssf_to_river_cumulative,
                            h->f.ssf_to_river_average, h->f.ssf_q_net_cumulative,
                            h->f.ssf_q_net_average, h->f.recharge_flux,
                            h->f.recharge_flux_cumulative, h->f.olf_inwater, h->f.olf_q,
                            h->f.olf_qlat, h->f.olf_qin, h->f.olf_qin_cumulative,
                            h->f.olf_qin_average, h->f.olf_q_cumulative, h->f.olf_q_average,
                            h->f.olf_storage, h->f.olf_h, h->f.olf_to_river_cumulative,
                            h->f.olf_to_river_average};
    for (double* p : zeros_land) fill(p, 0, 0.0);
    double* zeros_riv[] = {h->f.riv_external_inflow, h->f.riv_abstraction,
                           h->f.riv_actual_external_abstraction_cumulative,
                           h->f.riv_actual_external_abstraction_average, h->f.riv_inwater,
                           h->f.riv_q, h->f.riv_qlat, h->f.riv_qin, h->f.riv_qin_cumulative,
                           h->f.riv_qin_average, h->f.riv_q_cumulative, h->f.riv_q_average,
                           h->f.riv_storage, h->f.riv_h};
    for (double* p : zeros_riv) fill(p, 3, 0.0);
    double* zeros_res[] = {h->f.res_inflow_cumulative, h->f.res_inflow_average,   // reservoir.jl:200-272
                           h->f.res_external_inflow, h->f.res_actual_external_abstraction_cumulative,
                           h->f.res_actual_external_abstraction_average, h->f.res_outflow_cumulative,
                           h->f.res_outflow_average, h->f.res_actevap_cumulative};
    if (h->nres > 0)
      for (double* p : zeros_res) fill(p, 4, 0.0);
  }
  TRY_CREATE(cudaMalloc((void**)&h->f.number_of_layers, (size_t)h->ns * sizeof(int32_t)));
  TRY_CREATE(cudaMalloc((void**)&h->f.n_unsatlayers, (size_t)h->ns * sizeof(int32_t)));
  TRY_CREATE(cudaMalloc((void**)&h->f.nlayers_kv, (size_t)h->ns * sizeof(int32_t)));
  TRY_CREATE(cudaMemset(h->f.nlayers_kv, 0, (size_t)h->ns * sizeof(int32_t)));
  TRY_CREATE(cudaMemset(h->f.number_of_layers, 0, (size_t)h->ns * sizeof(int32_t)));
  TRY_CREATE(cudaMemset(h->f.n_unsatlayers, 0, (size_t)h->ns * sizeof(int32_t)));

  {
    std::vector<uint8_t> res_land, res_riv;
    if (h->nres > 0) {
      res_land.assign(h->n, 0);
      res_riv.assign(h->nriv, 0);
      for (int i = 0; i < h->nres; ++i) {
        const int64_t r = dom->reservoir_river_indices[i];
        if (r < 1 || r > h->nriv || res_riv[r - 1]) {
          h->err = "reservoir_river_indices must be distinct 1-based river node ids";
          return bail(WFLOWB200_ERR_ARG);
        }
        if (h->river.nw.down[r - 1] == 0) {   // surface_kinwave.jl:483-488
          h->err = "a reservoir without a downstream river node is not supported";
          return bail(WFLOWB200_ERR_GRAPH);
        }
        res_riv[r - 1] = 1;
        res_land[dom->river_land_indices[r - 1] - 1] = 1;
      }
    }
    auto vec = [](const int64_t* p, int64_t k) {
      return (p && k > 0) ? std::vector<int64_t>(p, p + k) : std::vector<int64_t>();
    };
    const int64_t nli = dom->n_land_imports, nle = dom->n_land_exports;
    const int64_t nri = dom->n_river_imports, nre = dom->n_river_exports;
    if (nli < 0 || nle < 0 || nri < 0 || nre < 0 || (nli && (!dom->land_import_dst || !dom->land_import_pos)) ||
        (nle && !dom->land_export_src) || (nri && (!dom->river_import_dst || !dom->river_import_pos)) ||
        (nre && !dom->river_export_src)) {
      h->err = "cut edges: counts and arrays do not agree";
      return bail(WFLOWB200_ERR_ARG);
    }
    if (nli + nle + nri + nre > 0 &&
        (cfg->adaptive || h->nres > 0 || cfg->snow_gravitational_transport || cfg->river_routing != 0 ||
         cfg->fp_levels > 0 || cfg->land_routing != 0)) {
      h->err = "cut edges are supported for kinematic-wave routing with fixed internal time steps, "
               "without reservoirs and lateral snow transport";
      return bail(WFLOWB200_ERR_ARG);
    }
    h->xchg.lay.n_imp[0] = nli; h->xchg.lay.n_imp[1] = nri;
    h->xchg.n_exp[0] = nle; h->xchg.n_exp[1] = nre;
    h->xchg.bound[0].assign(nle, {-1, -1});
    h->xchg.bound[1].assign(nre, {-1, -1});
    if (upload_domain(h, h->land, h->land.nw, res_land, true, vec(dom->land_import_dst, nli),
                      vec(dom->land_import_pos, nli), vec(dom->land_export_src, nle)) ||
        upload_domain(h, h->river, h->river.nw, res_riv, false, vec(dom->river_import_dst, nri),
                      vec(dom->river_import_pos, nri), vec(dom->river_export_src, nre)))
      return bail(h->err.find("import") != std::string::npos || h->err.find("export") != std::string::npos
                      ? WFLOWB200_ERR_ARG : WFLOWB200_ERR_CUDA);
    h->snow_net = &h->land;
    if (cfg->snow_gravitational_transport && h->nres > 0) {
      // accucapacityflux! walks the FULL land graph (routing/utils.jl:82-109): a second set of
      // chunk edges without the reservoir cut
      if (upload_domain(h, h->land_full, h->land.nw, std::vector<uint8_t>(), false))
        return bail(WFLOWB200_ERR_CUDA);
      h->snow_net = &h->land_full;
      std::vector<uint8_t> by_slot(h->ns, 0);
      for (int v = 0; v < h->n; ++v) by_slot[h->land.nw.slot_of[v]] = res_land[v];
      const uint8_t* dptr = nullptr;
      if (upload_raw(by_slot, &dptr, h->land_full.dev_arrays) != cudaSuccess) return bail(WFLOWB200_ERR_CUDA);
      h->f.land_is_res_outlet = const_cast<uint8_t*>(dptr);
    }
  }
  {
    std::vector<int64_t> riv_land_slot(h->nriv), riv_of_land(h->n, -1);
    for (int p = 0; p < h->nriv; ++p) {
      const int64_t rnode = h->river.nw.perm[p] - 1;
      const int64_t lnode = dom->river_land_indices[rnode] - 1;
      riv_land_slot[p] = h->land.nw.slot_of[lnode];
      riv_of_land[riv_land_slot[p]] = p;
    }
    TRY_CREATE(upload_i32(riv_land_slot, &h->f.riv_land_slot, 0));
    TRY_CREATE(upload_i32(riv_of_land, &h->riv_of_land, 0));
    if (h->nres > 0) {  // NetworkRiver.reservoir_indices by river slot; the outlet's land slot
      std::vector<int64_t> riv_res(std::max(h->nrs, 1), -1), res_land(h->nres), ident(h->nres);
      for (int i = 0; i < h->nres; ++i) {
        const int64_t rnode = dom->reservoir_river_indices[i] - 1;
        const int64_t p = h->river.nw.slot_of[rnode];
        riv_res[p] = i;
        res_land[i] = riv_land_slot[p];
        ident[i] = i;
      }
      TRY_CREATE(upload_i32(riv_res, &h->f.riv_reservoir, 0));
      TRY_CREATE(upload_i32(res_land, &h->f.res_land_slot, 0));
      {
        std::vector<int64_t> res_riv_slot(h->nres);
        for (int i = 0; i < h->nres; ++i)
          res_riv_slot[i] = h->river.nw.slot_of[dom->reservoir_river_indices[i] - 1];
        TRY_CREATE(upload_i32(res_riv_slot, &h->f.res_river_slot, 0));
      }
      TRY_CREATE(upload_i32(ident, &h->res_ident, 0));
    }
  }
  h->stage_doubles = std::max<size_t>(std::max<size_t>((size_t)h->n * (h->N + 1), (size_t)3 * h->n),
                                      (size_t)std::max(h->nriv * std::max(cfg->fp_levels, 1), h->nres));
  TRY_CREATE(cudaMalloc((void**)&h->d_stage, h->stage_doubles * sizeof(double)));
  TRY_CREATE(cudaMalloc((void**)&h->d_forcing, (size_t)3 * h->n * sizeof(double)));
  TRY_CREATE(cudaMallocHost((void**)&h->h_pinned, (size_t)3 * h->n * sizeof(double)));
  TRY_CREATE(cudaMalloc((void**)&h->d_queue, 3 * 32 * sizeof(unsigned)));
  {
    const size_t ns = (size_t)h->ns;
    TRY_CREATE(cudaMalloc((void**)&h->d_unsat_pool, 5 * ns * sizeof(double)));
    TRY_CREATE(cudaMalloc((void**)&h->d_unsat_its, ns * sizeof(int32_t)));
    TRY_CREATE(cudaMalloc((void**)&h->d_unsat_list, (size_t)WFB_UNSAT_BUCKETS * ns * sizeof(int32_t)));
    TRY_CREATE(cudaMalloc((void**)&h->d_unsat_count, (2 * WFB_UNSAT_BUCKETS + 1) * sizeof(unsigned)));
    UnsatWork& u = h->unsat;
    u.usd = h->d_unsat_pool; u.sum_ast = u.usd + ns; u.kv_it = u.sum_ast + ns;
    u.l_sat = u.kv_it + ns; u.c = u.l_sat + ns;
    u.its_layer = h->d_unsat_its;
    u.list = h->d_unsat_list;
    u.count = h->d_unsat_count;
    u.cap = (int32_t)ns;
    // in-line trip limit (the first half's in-line loops beyond the first trip run with ~4 of 32
    // lanes; the engine regroups them, but pays scratch traffic per suspended cell). Measured on
    // B200, 1000^2: V1 of steps 12-14 [us] 2: 365-376, 3: 374, 4: 374-381, 6: 385, 8: 386-390;
    // mean over steps 10-59 (the basin keeps getting wetter) 2: 436, 4: 417, 8: 427
    u.inline_iters = cfg->unsat_inline_iters > 0 ? cfg->unsat_inline_iters : 4;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    h->engine_grid = std::max(1, sms);
  }
  TRY_CREATE(cudaMalloc((void**)&h->d_err, sizeof(unsigned)));
  TRY_CREATE(cudaMemset(h->d_err, 0, sizeof(unsigned)));
  TRY_CREATE(cudaMalloc((void**)&h->d_stats, sizeof(RoutingStats)));
  TRY_CREATE(cudaMemset(h->d_stats, 0, sizeof(RoutingStats)));
  TRY_CREATE(cudaMalloc((void**)&h->d_count, sizeof(unsigned long long)));
  TRY_CREATE(cudaMalloc((void**)&h->d_min, sizeof(double)));
  if (cfg->adaptive) {
    TRY_CREATE(cudaMalloc((void**)&h->d_work, (size_t)h->ns * sizeof(double)));
    TRY_CREATE(cudaMalloc((void**)&h->d_qstate, WFB_QUANTILE_STATE_WORDS * sizeof(unsigned long long)));
  }

  h->kc.n = h->n; h->kc.nriv = h->nriv; h->kc.ns = h->ns; h->kc.nrs = h->nrs; h->kc.nres = h->nres;
  h->kc.gash = cfg->gash; h->kc.has_lai = cfg->has_lai; h->kc.snow = cfg->snow;
  h->kc.glacier = cfg->glacier;
  h->kc.soil_infiltration_reduction = cfg->soil_infiltration_reduction;
  h->kc.kv_profile = cfg->kv_profile;
  h->kc.qroot = h->cfg.kin_wave_min_flow_qroot;
  build_vertical_stage(h->f, h->kc, h->N, h->v_stage);
  h->kc.river_routing = cfg->river_routing;
  h->kc.fp_levels = cfg->fp_levels;
  for (int l = 0; l < 16; ++l) h->kc.fp_depth[l] = cfg->fp_depth[l];
  {
    double* d = nullptr;
    TRY_CREATE(cudaMalloc((void**)&d, 16 * sizeof(double)));
    TRY_CREATE(cudaMemcpy(d, cfg->fp_depth, 16 * sizeof(double), cudaMemcpyHostToDevice));
    h->f.fp_depth = d;
  }
  if (cfg->river_routing == 1 || cfg->fp_levels > 0) {   // the staggered grid by river slot (network.jl:281-293);
                                                           // the kinematic wave's floodplain needs the downstream slot
    const Network& rn = h->river.nw;
    std::vector<int64_t> dst(std::max(h->nrs, 1), -1), in_ptr(h->nriv + 1, 0), in_idx;
    for (int p = 0; p < h->nriv; ++p) {
      const int64_t v = rn.perm[p] - 1;
      dst[p] = rn.down[v] > 0 ? rn.slot_of[rn.down[v] - 1] : (cfg->li_ghost_nodes ? -2 : -1);
      for (int64_t e = rn.in_ptr[v]; e < rn.in_ptr[v + 1]; ++e)   // ascending source node id
        in_idx.push_back(rn.slot_of[rn.in_idx[e] - 1]);
      in_ptr[p + 1] = (int64_t)in_idx.size();
    }
    TRY_CREATE(upload_i32(dst, &h->f.li_dst_slot, 0));
    TRY_CREATE(upload_i32(in_ptr, &h->f.li_in_ptr, 0));
    TRY_CREATE(upload_i32(in_idx, &h->f.li_in_idx, 0));
    TRY_CREATE(cudaMalloc((void**)&h->d_li_barrier, 2 * sizeof(unsigned)));
    TRY_CREATE(cudaMemset(h->d_li_barrier, 0, 2 * sizeof(unsigned)));
    TRY_CREATE(cudaMalloc((void**)&h->d_li_dt, 4 * sizeof(unsigned long long)));
    TRY_CREATE(cudaMalloc((void**)&h->d_li_substeps, sizeof(int)));
    TRY_CREATE(cudaMemset(h->d_li_substeps, 0, sizeof(int)));
    h->grid_li = li_max_grid(cfg->device);
    if (h->grid_li <= 0) { h->err = "occupancy query failed"; return bail(WFLOWB200_ERR_CUDA); }
  }
  h->kc.land_routing = cfg->land_routing;
  if (cfg->land_routing == 1) {   // the 2-D state is kept in NODE order (local_inertial.cu)
    const Network& ln = h->land.nw;   // EdgeConnectivity (network.jl:136-153): -1 = no neighbour
    auto by_node = [&](const std::vector<int64_t>& src) {
      std::vector<int64_t> e(std::max(h->n, 1), -1);
      for (int v = 0; v < h->n; ++v) e[v] = src[v] <= h->n ? src[v] - 1 : -1;
      return e;
    };
    TRY_CREATE(upload_i32(by_node(ln.edge_x_up), &h->f.edge_x_up, 0));
    TRY_CREATE(upload_i32(by_node(ln.edge_x_down), &h->f.edge_x_down, 0));
    TRY_CREATE(upload_i32(by_node(ln.edge_y_up), &h->f.edge_y_up, 0));
    TRY_CREATE(upload_i32(by_node(ln.edge_y_down), &h->f.edge_y_down, 0));
    std::vector<int64_t> river_of_node(h->n, -1), ident(h->n);   // domain.land.network.river_indices
    for (int r = 0; r < h->nriv; ++r)
      river_of_node[dom->river_land_indices[r] - 1] = h->river.nw.slot_of[r];
    for (int v = 0; v < h->n; ++v) ident[v] = v;
    TRY_CREATE(upload_i32(river_of_node, &h->f.lil_river_slot, 0));
    TRY_CREATE(upload_i32(ln.slot_of, &h->f.land_slot_of_node, 0));
    TRY_CREATE(upload_i32(ident, &h->land_ident, 0));
    if (h->nres > 0) {
      std::vector<int64_t> res_node(h->nres);
      for (int i = 0; i < h->nres; ++i)
        res_node[i] = dom->river_land_indices[dom->reservoir_river_indices[i] - 1] - 1;
      TRY_CREATE(upload_i32(res_node, &h->f.res_land_node, 0));
    }
    TRY_CREATE(cudaMalloc((void**)&h->f.lil_h, (size_t)h->n * sizeof(double)));
    TRY_CREATE(cudaMalloc((void**)&h->f.lil_storage, (size_t)h->n * sizeof(double)));
    TRY_CREATE(cudaMalloc((void**)&h->f.lil_cell_area, (size_t)h->n * sizeof(double)));
    TRY_CREATE(cudaMalloc((void**)&h->f.lil_xu_eff, (size_t)h->n * sizeof(int32_t)));
    TRY_CREATE(cudaMalloc((void**)&h->f.lil_yu_eff, (size_t)h->n * sizeof(int32_t)));
    h->grid_lil = lil_max_grid(cfg->device);
    if (h->grid_lil <= 0) { h->err = "occupancy query failed"; return bail(WFLOWB200_ERR_CUDA); }
  }

  // persistent cooperative grids: as many co-resident CTAs as the device holds
  h->smem_olf = wave_smem(0, h->land.dev.max_inlets);
  const bool kw_floodplain = cfg->fp_levels > 0 && cfg->river_routing == 0;
  h->smem_riv = wave_smem(kw_floodplain ? 4 : 1, h->river.dev.max_inlets);
  h->smem_ssf = wave_smem(2, h->land.dev.max_inlets);
  h->grid_olf = wave_max_grid(0, h->N, h->smem_olf, cfg->device);
  h->grid_riv = wave_max_grid(kw_floodplain ? 4 : 1, h->N, h->smem_riv, cfg->device);
  h->grid_ssf = wave_max_grid(2, h->N, h->smem_ssf, cfg->device);
  if (cfg->snow_gravitational_transport) {
    h->smem_snow = wave_smem(3, std::max(h->land.dev.max_inlets, h->land_full.dev.max_inlets));
    h->grid_snow = wave_max_grid(3, h->N, h->smem_snow, cfg->device);
    if (h->grid_snow <= 0) { h->err = "occupancy query failed"; return bail(WFLOWB200_ERR_CUDA); }
  }
  h->smem_surface = surface_smem(h->land.dev.max_inlets, h->river.dev.max_inlets,
                                 &h->smem_surface_per_warp);
  h->grid_surface = surface_max_grid(h->smem_surface, cfg->device);
  h->fuse_surface = h->grid_surface > 0 && h->nriv > 0 && !cfg->adaptive && cfg->river_routing == 0 &&
                    !kw_floodplain;   // (the river with floodplain publishes three values: own kernel)
  if (h->grid_olf <= 0 || h->grid_riv <= 0 || h->grid_ssf <= 0) {
    h->err = "occupancy query failed";
    return bail(WFLOWB200_ERR_CUDA);
  }
  TRY_CREATE(cudaStreamSynchronize(h->stream));
#undef TRY_CREATE
  *out = h;
  return WFLOWB200_OK;
}

static void free_exchange(WflowB200* h);
static int32_t exchange_begin_step(WflowB200* h);
void wflowb200_destroy(WflowB200* h) {
  if (!h) return;
  DeviceGuard device_guard_(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  free_exchange(h);
  delete h->comm;
  h->comm = nullptr;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  cudaFree(h->pool); cudaFree(h->f.number_of_layers); cudaFree(h->f.n_unsatlayers);
  cudaFree(h->f.nlayers_kv); cudaFree(h->f.olf_newton_trace); cudaFree(h->f.riv_newton_trace);
  cudaFree((void*)h->f.fp_depth);
  cudaFree(h->f.riv_reservoir); cudaFree(h->f.res_land_slot); cudaFree(h->res_ident);
  cudaFree(h->f.res_river_slot); cudaFree(h->f.li_dst_slot); cudaFree(h->f.li_in_ptr);
  cudaFree(h->f.edge_x_up); cudaFree(h->f.edge_x_down); cudaFree(h->f.edge_y_up); cudaFree(h->f.edge_y_down);
  cudaFree(h->f.lil_river_slot); cudaFree(h->f.land_slot_of_node); cudaFree(h->f.res_land_node);
  cudaFree(h->f.lil_h); cudaFree(h->f.lil_storage); cudaFree(h->land_ident);
  cudaFree(h->f.lil_cell_area); cudaFree(h->f.lil_xu_eff); cudaFree(h->f.lil_yu_eff);
  cudaFree(h->f.li_in_idx); cudaFree(h->d_li_barrier); cudaFree(h->d_li_dt); cudaFree(h->d_li_substeps);
  cudaFree(h->f.riv_land_slot); cudaFree(h->riv_of_land); cudaFree(h->d_stage);
  cudaFree(h->d_forcing); cudaFreeHost(h->h_pinned); cudaFree(h->d_queue);
  cudaFree(h->d_ring); cudaFree(h->d_lai_table); cudaFreeHost(h->h_out_pinned);
  cudaFree(h->d_out[0]); cudaFree(h->d_out[1]);
  if (h->out_gathered) cudaEventDestroy(h->out_gathered);
  for (auto e : h->out_copied) if (e) cudaEventDestroy(e);
  for (auto e : h->ring_ready) if (e) cudaEventDestroy(e);
  for (auto e : h->ring_consumed) if (e) cudaEventDestroy(e);
  cudaFree(h->d_unsat_pool); cudaFree(h->d_unsat_its); cudaFree(h->d_unsat_list);
  cudaFree(h->d_unsat_count);
  cudaFree(h->d_err);
  if (h->v_graph) cudaGraphExecDestroy(h->v_graph);
  for (auto e : h->tl_ev) if (e) cudaEventDestroy(e);
  cudaFree(h->d_stats); cudaFree(h->d_count); cudaFree(h->d_min);
  cudaFree(h->d_work); cudaFree(h->d_qstate); cudaFree(h->ssf_q_out);
  free_domain(h->land); free_domain(h->river); free_domain(h->land_full);
  if (h->forcing_ready) cudaEventDestroy(h->forcing_ready);
  if (h->forcing_consumed) cudaEventDestroy(h->forcing_consumed);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->tm) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
}

static int32_t field_common(WflowB200* h, int32_t id, int64_t sc, int64_t sl, int& kind,
                            int& layers, int& count, size_t& extent) {
  if (id < 0 || id >= WFLOWB200_NUM_FIELDS) return fail(h, WFLOWB200_ERR_ARG, "bad field id");
  kind = kFieldKinds[id];
  layers = layers_of(h, kind);
  count = count_of(h, kind);
  if (layers == 1) {
    if (sc != 1) return fail(h, WFLOWB200_ERR_ARG, "scalar fields need stride_cell == 1");
  } else if (!((sc == layers && sl == 1) || (sc == 1 && sl == count))) {
    return fail(h, WFLOWB200_ERR_ARG,
                "layered fields need a dense layout: (stride_cell = layers, stride_layer = 1) "
                "or (stride_cell = 1, stride_layer = n)");
  }
  extent = (size_t)count * layers;
  return WFLOWB200_OK;
}

int32_t wflowb200_set_field(WflowB200* h, int32_t id, const double* src, int64_t sc, int64_t sl) {
  WFB_ENTER(h);
  int kind, layers, count; size_t extent;
  int32_t rc = field_common(h, id, sc, sl, kind, layers, count, extent);
  if (rc) return rc;
  if (!src) return fail(h, WFLOWB200_ERR_ARG, "null src");
  if (count == 0) return WFLOWB200_OK;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, src, extent * sizeof(double), cudaMemcpyHostToDevice,
                              h->stream));
  const int32_t* slot_map = slot_map_of(h, kind);
  h->launches += launch_gather_field(h->field_ptr[id], h->d_stage, slot_map, count,
                                     slots_of(h, kind), layers, sc, sl, h->stream);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return WFLOWB200_OK;
}

int32_t wflowb200_get_field(WflowB200* h, int32_t id, double* dst, int64_t sc, int64_t sl) {
  WFB_ENTER(h);
  int kind, layers, count; size_t extent;
  int32_t rc = field_common(h, id, sc, sl, kind, layers, count, extent);
  if (rc) return rc;
  if (!dst) return fail(h, WFLOWB200_ERR_ARG, "null dst");
  if (count == 0) return WFLOWB200_OK;
  rc = wait_forcing(h);
  if (rc) return rc;
  const int32_t* slot_map = slot_map_of(h, kind);
  h->launches += launch_scatter_field(h->d_stage, h->field_ptr[id], slot_map, count,
                                      slots_of(h, kind), layers, sc, sl, h->stream);
  CUDA_TRY(h, cudaMemcpyAsync(dst, h->d_stage, extent * sizeof(double), cudaMemcpyDeviceToHost,
                              h->stream));
  return check_device_error(h);
}

int32_t wflowb200_set_field_i64(WflowB200* h, int32_t which, const int64_t* src) {
  WFB_ENTER(h);
  if (!src) return WFLOWB200_ERR_ARG;
  if (which < 0 || which > 2) return fail(h, WFLOWB200_ERR_ARG, "bad int field");
  std::vector<int32_t> tmp(h->n);
  for (int p = 0; p < h->n; ++p) tmp[p] = (int32_t)src[h->land.nw.perm[p] - 1];
  int32_t* dst = which == 0 ? h->f.number_of_layers : which == 1 ? h->f.n_unsatlayers : h->f.nlayers_kv;
  CUDA_TRY(h, cudaMemcpyAsync(dst, tmp.data(), tmp.size() * sizeof(int32_t),
                              cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return WFLOWB200_OK;
}

int32_t wflowb200_get_field_i64(WflowB200* h, int32_t which, int64_t* dst) {
  WFB_ENTER(h);
  if (!dst) return WFLOWB200_ERR_ARG;
  if (which < 0 || which > 2) return fail(h, WFLOWB200_ERR_ARG, "bad int field");
  std::vector<int32_t> tmp(h->n);
  const int32_t* src = which == 0 ? h->f.number_of_layers : which == 1 ? h->f.n_unsatlayers : h->f.nlayers_kv;
  CUDA_TRY(h, cudaMemcpyAsync(tmp.data(), src, tmp.size() * sizeof(int32_t),
                              cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int p = 0; p < h->n; ++p) dst[h->land.nw.perm[p] - 1] = tmp[p];
  return WFLOWB200_OK;
}

int32_t wflowb200_set_forcing(WflowB200* h, const double* P, const double* PET, const double* T) {
  WFB_ENTER(h);
  if (!P || !PET || !T) return WFLOWB200_ERR_ARG;
  // the previous forcing must have been consumed before the staging buffers are reused
  CUDA_TRY(h, cudaEventSynchronize(h->forcing_consumed));
  const size_t nb = (size_t)h->n * sizeof(double);
  // Page-locked caller arrays (cudaHostRegister / cudaMallocHost) are copied straight to the
  // device; the call returns when the copy has left them, so the caller may reuse them at once
  // (same contract as the staged path). Pageable arrays go through the pinned staging buffer.
  auto pinned = [](const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  if (pinned(P) && pinned(PET) && pinned(T)) {
    CUDA_TRY(h, cudaMemcpyAsync(h->d_forcing, P, nb, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_forcing + h->n, PET, nb, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_forcing + 2 * (size_t)h->n, T, nb, cudaMemcpyHostToDevice,
                                h->copy_stream));
    CUDA_TRY(h, cudaEventRecord(h->forcing_ready, h->copy_stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    h->forcing_pending = true;
    return WFLOWB200_OK;
  }
  memcpy(h->h_pinned, P, nb);
  memcpy(h->h_pinned + h->n, PET, nb);
  memcpy(h->h_pinned + 2 * (size_t)h->n, T, nb);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_forcing, h->h_pinned, 3 * nb, cudaMemcpyHostToDevice,
                              h->copy_stream));
  CUDA_TRY(h, cudaEventRecord(h->forcing_ready, h->copy_stream));
  h->forcing_pending = true;
  return WFLOWB200_OK;
}

// ---- forcing / cyclic-parameter staging and output gather (the steps either side of the path:
//      io.jl:108-227 update_forcing! / update_cyclic!, io.jl:815-899 write_output) ---------------
int32_t wflowb200_forcing_ring_create(WflowB200* h, int32_t depth) {
  WFB_ENTER(h);
  if (depth < 1 || depth > 1024) return fail(h, WFLOWB200_ERR_ARG, "1 <= depth <= 1024 required");
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
  cudaFree(h->d_ring);
  h->d_ring = nullptr;
  for (auto e : h->ring_ready) cudaEventDestroy(e);
  for (auto e : h->ring_consumed) cudaEventDestroy(e);
  h->ring_ready.assign(depth, nullptr);
  h->ring_consumed.assign(depth, nullptr);
  CUDA_TRY(h, cudaMalloc((void**)&h->d_ring, (size_t)depth * 3 * h->n * sizeof(double)));
  for (int k = 0; k < depth; ++k) {
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ring_ready[k], cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ring_consumed[k], cudaEventDisableTiming));
  }
  h->ring_depth = depth;
  h->ring_use = -1;
  return WFLOWB200_OK;
}

int32_t wflowb200_forcing_ring_put(WflowB200* h, int32_t slot, const double* P, const double* PET,
                                   const double* T) {
  WFB_ENTER(h);
  if (!P || !PET || !T) return WFLOWB200_ERR_ARG;
  if (slot < 0 || slot >= h->ring_depth) return fail(h, WFLOWB200_ERR_ARG, "bad ring slot");
  // the slab may still be read by the gather of an earlier step
  CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ring_consumed[slot], 0));
  const size_t nb = (size_t)h->n * sizeof(double);
  double* dst = h->d_ring + (size_t)slot * 3 * h->n;
  // page-locked caller arrays stream straight over; pageable ones are staged by the driver
  CUDA_TRY(h, cudaMemcpyAsync(dst, P, nb, cudaMemcpyHostToDevice, h->copy_stream));
  CUDA_TRY(h, cudaMemcpyAsync(dst + h->n, PET, nb, cudaMemcpyHostToDevice, h->copy_stream));
  CUDA_TRY(h, cudaMemcpyAsync(dst + 2 * (size_t)h->n, T, nb, cudaMemcpyHostToDevice, h->copy_stream));
  CUDA_TRY(h, cudaEventRecord(h->ring_ready[slot], h->copy_stream));
  return WFLOWB200_OK;
}

int32_t wflowb200_forcing_ring_use(WflowB200* h, int32_t slot) {
  WFB_ENTER(h);
  if (slot < 0 || slot >= h->ring_depth) return fail(h, WFLOWB200_ERR_ARG, "bad ring slot");
  h->ring_use = slot;
  return WFLOWB200_OK;
}

int32_t wflowb200_set_cyclic_lai(WflowB200* h, const double* table, int32_t n_slabs) {
  WFB_ENTER(h);
  if (!table || n_slabs < 1) return WFLOWB200_ERR_ARG;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(h->d_lai_table);
  h->d_lai_table = nullptr;
  CUDA_TRY(h, cudaMalloc((void**)&h->d_lai_table, (size_t)n_slabs * h->n * sizeof(double)));
  CUDA_TRY(h, cudaMemcpy(h->d_lai_table, table, (size_t)n_slabs * h->n * sizeof(double),
                         cudaMemcpyHostToDevice));
  h->lai_slabs = n_slabs;
  return WFLOWB200_OK;
}

int32_t wflowb200_use_cyclic_lai(WflowB200* h, int32_t slab) {
  WFB_ENTER(h);
  if (slab < 0 || slab >= h->lai_slabs) return fail(h, WFLOWB200_ERR_ARG, "bad LAI slab");
  h->launches += launch_gather_field(h->f.leaf_area_index, h->d_lai_table + (size_t)slab * h->n,
                                     h->land.node_of_slot, h->n, h->ns, 1, 1, 1, h->stream);
  return WFLOWB200_OK;
}

int32_t wflowb200_get_fields(WflowB200* h, const int32_t* ids, int32_t n_ids, double* dst) {
  WFB_ENTER(h);
  if (!ids || !dst || n_ids < 1) return WFLOWB200_ERR_ARG;
  size_t total = 0;
  for (int k = 0; k < n_ids; ++k) {
    if (ids[k] < 0 || ids[k] >= WFLOWB200_NUM_FIELDS) return fail(h, WFLOWB200_ERR_ARG, "bad field id");
    const int kind = kFieldKinds[ids[k]];
    total += (size_t)count_of(h, kind) * layers_of(h, kind);
  }
  int32_t rc = wait_forcing(h);
  if (rc) return rc;
  if (total > h->stage_doubles) {  // the device staging grows to the largest request
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_stage);
    h->d_stage = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&h->d_stage, total * sizeof(double)));
    h->stage_doubles = total;
  }
  if (total > h->out_pinned_doubles) {
    cudaFreeHost(h->h_out_pinned);
    h->h_out_pinned = nullptr;
    CUDA_TRY(h, cudaMallocHost((void**)&h->h_out_pinned, total * sizeof(double)));
    h->out_pinned_doubles = total;
  }
  size_t off = 0;
  for (int k = 0; k < n_ids; ++k) {  // node order, layered fields cell-major (Vector{SVector{N}})
    const int kind = kFieldKinds[ids[k]];
    const int layers = layers_of(h, kind), count = count_of(h, kind);
    if (count == 0) continue;
    const int32_t* slot_map = slot_map_of(h, kind);
    h->launches += launch_scatter_field(h->d_stage + off, h->field_ptr[ids[k]], slot_map, count,
                                        slots_of(h, kind), layers, layers, 1, h->stream);
    off += (size_t)count * layers;
  }
  // ONE device-to-host copy for the whole set, through page-locked memory
  CUDA_TRY(h, cudaMemcpyAsync(h->h_out_pinned, h->d_stage, total * sizeof(double),
                              cudaMemcpyDeviceToHost, h->stream));
  rc = check_device_error(h);
  if (rc) return rc;
  memcpy(dst, h->h_out_pinned, total * sizeof(double));
  return WFLOWB200_OK;
}

// The same gather without blocking the compute stream: the fields are packed on the compute
// stream into one of two device staging buffers, the device-to-host copy runs on the copy stream
// (dst must be page-locked), and the next update_* calls may be enqueued at once.
// wflowb200_wait_outputs blocks until every pending copy has landed in its dst.
int32_t wflowb200_get_fields_async(WflowB200* h, const int32_t* ids, int32_t n_ids, double* dst) {
  WFB_ENTER(h);
  if (!ids || !dst || n_ids < 1) return WFLOWB200_ERR_ARG;
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, dst) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
    cudaGetLastError();
    return fail(h, WFLOWB200_ERR_ARG, "get_fields_async needs a page-locked destination");
  }
  size_t total = 0;
  for (int k = 0; k < n_ids; ++k) {
    if (ids[k] < 0 || ids[k] >= WFLOWB200_NUM_FIELDS) return fail(h, WFLOWB200_ERR_ARG, "bad field id");
    const int kind = kFieldKinds[ids[k]];
    total += (size_t)count_of(h, kind) * layers_of(h, kind);
  }
  int32_t rc = wait_forcing(h);
  if (rc) return rc;
  if (!h->out_gathered) {
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->out_gathered, cudaEventDisableTiming));
    for (auto& e : h->out_copied) CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const int b = h->out_parity;
  h->out_parity ^= 1;
  if (total > h->out_doubles[b]) {
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    cudaFree(h->d_out[b]);
    h->d_out[b] = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&h->d_out[b], total * sizeof(double)));
    h->out_doubles[b] = total;
  }
  // the copy that last used this staging buffer must have left it
  CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->out_copied[b], 0));
  size_t off = 0;
  for (int k = 0; k < n_ids; ++k) {
    const int kind = kFieldKinds[ids[k]];
    const int layers = layers_of(h, kind), count = count_of(h, kind);
    if (count == 0) continue;
    const int32_t* slot_map = slot_map_of(h, kind);
    h->launches += launch_scatter_field(h->d_out[b] + off, h->field_ptr[ids[k]], slot_map, count,
                                        slots_of(h, kind), layers, layers, 1, h->stream);
    off += (size_t)count * layers;
  }
  CUDA_TRY(h, cudaEventRecord(h->out_gathered, h->stream));
  CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->out_gathered, 0));
  CUDA_TRY(h, cudaMemcpyAsync(dst, h->d_out[b], total * sizeof(double), cudaMemcpyDeviceToHost,
                              h->copy_stream));
  CUDA_TRY(h, cudaEventRecord(h->out_copied[b], h->copy_stream));
  return WFLOWB200_OK;
}

int32_t wflowb200_wait_outputs(WflowB200* h) {
  WFB_ENTER(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
  return WFLOWB200_OK;
}

int32_t wflowb200_update_land_hydrology_model(WflowB200* h, double dt) {
  WFB_ENTER(h);
  int32_t rc = wait_forcing(h);
  if (rc) return rc;
  if (h->nriv > 0) {  // river h -> land grid (runoff.jl:77-79)
    if ((rc = check_launch(h, launch_scatter_river_depth(h->f, h->kc, h->stream), "scatter"))) return rc;
  }
  return launch_vertical(h, dt);
}

int32_t wflowb200_exchange_recharge(WflowB200* h) {
  WFB_ENTER(h);
  return check_launch(h, launch_exchange_recharge(h->f, h->kc, h->stream), "exchange_recharge");
}

int32_t wflowb200_update_subsurface_flow_model(WflowB200* h, double dt) {
  WFB_ENTER(h);
  if (h->cfg.adaptive)
    return run_wave_adaptive(h, h->land, dt, 2, 2, h->grid_ssf, h->smem_ssf, h->sub_ssf,
                             [&](const WaveLaunch& w) {
                               return launch_subsurface_wave(h->f, h->kc, h->land.dev, h->N, w, h->stream);
                             }, "update_subsurface_flow_model");
  return run_wave(h, h->land, dt, h->cfg.dt_ssf, 2, 2, h->grid_ssf, h->smem_ssf, h->sub_ssf,
                  [&](const WaveLaunch& w) {
                    return launch_subsurface_wave(h->f, h->kc, h->land.dev, h->N, w, h->stream);
                  }, "update_subsurface_flow_model");
}

// update_subsurface_flow_model! with update_soil_water_storage! fused into the end of every
// cell's subsurface flow (update_model path, fixed internal time step)
static int32_t update_subsurface_and_soil_storage(WflowB200* h, double dt, bool* fused) {
  *fused = false;
  // Fusing pays where the sweep is bound by its dependent chain and leaves the memory system
  // idle (1000^2: -0.05 ms per step); on wide, throughput-bound domains the separate
  // bandwidth-bound kernel is faster (3536^2: +0.8 ms when fused). Same criterion as the piece
  // depth of the chunks: the mean level width.
  const bool wide = h->land.nw.n_wave_levels > 0 &&
                    h->land.nw.n / h->land.nw.n_wave_levels >= WFB_PIECE_WIDE_LEVEL;
  const bool want = h->tune.fuse_soil_storage >= 0 ? h->tune.fuse_soil_storage != 0 : !wide;
  if (h->cfg.adaptive || !want)
    return wflowb200_update_subsurface_flow_model(h, dt);
  *fused = true;
  return run_wave(h, h->land, dt, h->cfg.dt_ssf, 2, 2, h->grid_ssf, h->smem_ssf, h->sub_ssf,
                  [&](const WaveLaunch& w) {
                    return launch_subsurface_wave(h->f, h->kc, h->land.dev, h->N, w, h->stream);
                  }, "update_subsurface_flow_model", 0.0, false, true);
}

int32_t wflowb200_update_soil_water_storage(WflowB200* h, double dt) {
  WFB_ENTER(h);
  (void)dt;
  return check_launch(h, launch_soil_water_storage(h->f, h->kc, h->N, h->stream),
                      "update_soil_water_storage");
}

int32_t wflowb200_update_lateral_inflow_overland(WflowB200* h) {
  WFB_ENTER(h);
  return check_launch(h, launch_lateral_inflow_overland(h->f, h->kc, h->stream),
                      "update_lateral_inflow(overland)");
}

static LiLaunch li_launch(WflowB200* h, double dt) {
  LiLaunch w{};
  w.dt = dt;
  w.alpha = h->cfg.li_alpha > 0.0 ? h->cfg.li_alpha : 0.7;
  w.h_thresh = h->cfg.li_h_thresh;
  w.froude_limit = h->cfg.li_froude_limit;
  w.fp_levels = h->cfg.fp_levels;
  for (int l = 0; l < 16; ++l) w.fp_depth[l] = h->cfg.fp_depth[l];
  w.barrier = h->d_li_barrier;
  w.dt_bits = h->d_li_dt;
  w.err = h->d_err;
  w.substeps = h->d_li_substeps;
  w.land_alpha = h->cfg.li_land_alpha > 0.0 ? h->cfg.li_land_alpha : 0.7;
  w.land_theta = h->cfg.li_land_theta;
  w.land_h_thresh = h->cfg.li_land_h_thresh;
  w.land_froude_limit = h->cfg.li_land_froude_limit;
  return w;
}

int32_t wflowb200_update_bc_overland_flow_model(WflowB200* h) {
  WFB_ENTER(h);
  if (h->cfg.land_routing != 1)
    return fail(h, WFLOWB200_ERR_STATE, "update_bc_overland_flow_model: land_routing is not local_inertial");
  return check_launch(h, launch_bc_overland_flow(h->f, h->kc, h->stream), "update_bc_overland_flow_model");
}

int32_t wflowb200_update_overland_flow_model(WflowB200* h, double dt) {
  WFB_ENTER(h);
  if (h->cfg.land_routing == 1) {  // 2-D local-inertial overland flow + local-inertial river flow
    LiLaunch w = li_launch(h, dt);
    w.grid = std::max(1, std::min(h->grid_lil, (h->n + 255) / 256));
    return check_launch(h, launch_local_inertial_land_river(h->f, h->kc, w, h->stream),
                        "update_overland_flow_model (local inertial)");
  }
  if (h->cfg.adaptive)
    return run_wave_adaptive(h, h->land, dt, 0, 2, h->grid_olf, h->smem_olf, h->sub_land,
                             [&](const WaveLaunch& w) {
                               return launch_overland_wave(h->f, h->kc, h->land.dev, w, h->stream);
                             }, "update_overland_flow_model");
  return run_wave(h, h->land, dt, h->cfg.dt_land, 0, 2, h->grid_olf, h->smem_olf, h->sub_land,
                  [&](const WaveLaunch& w) {
                    return launch_overland_wave(h->f, h->kc, h->land.dev, w, h->stream);
                  }, "update_overland_flow_model");
}

int32_t wflowb200_update_lateral_inflow_river(WflowB200* h) {
  WFB_ENTER(h);
  return check_launch(h, launch_lateral_inflow_river(h->f, h->kc, h->stream),
                      "update_lateral_inflow(river)");
}

int32_t wflowb200_update_inflow_reservoir(WflowB200* h) {
  WFB_ENTER(h);
  if (h->nres == 0) return WFLOWB200_OK;
  return check_launch(h, launch_inflow_reservoir(h->f, h->kc, h->stream), "update_inflow(reservoir)");
}

int32_t wflowb200_update_river_flow_model(WflowB200* h, double dt) {
  WFB_ENTER(h);
  if (h->nriv == 0) return WFLOWB200_OK;
  if (h->cfg.land_routing == 1)   // the river is routed together with the land (one scheme, one dt_s)
    return fail(h, WFLOWB200_ERR_STATE, "land_routing = local_inertial: the river flow is part of "
                                        "wflowb200_update_overland_flow_model");
  if (h->cfg.river_routing == 1) {  // local-inertial river flow: the whole model step in one kernel
    LiLaunch w = li_launch(h, dt);
    w.grid = std::max(1, std::min(h->grid_li, (h->nriv + 255) / 256));
    return check_launch(h, launch_local_inertial_river(h->f, h->kc, w, h->stream),
                        "update_river_flow_model (local inertial)");
  }
  const bool fp = h->cfg.fp_levels > 0;   // the kinematic wave's 1-D floodplain: 3 values per node
  const int nv = fp ? 3 : 1;
  auto launch = [&](const WaveLaunch& w) {
    return fp ? launch_river_floodplain_wave(h->f, h->kc, h->river.dev, w, h->stream)
              : launch_river_wave(h->f, h->kc, h->river.dev, w, h->stream);
  };
  if (h->cfg.adaptive)
    return run_wave_adaptive(h, h->river, dt, 1, nv, h->grid_riv, h->smem_riv, h->sub_river, launch,
                             "update_river_flow_model");
  return run_wave(h, h->river, dt, h->cfg.dt_river, 1, nv, h->grid_riv, h->smem_riv, h->sub_river,
                  launch, "update_river_flow_model");
}

// update_overland_flow_model! + update_lateral_inflow!(river) + update_river_flow_model! in one
// launch (routing.cu: surface_wave_kernel); same results as the three separate entry points.
static int32_t update_surface_fused(WflowB200* h, double dt) {
  WaveLaunch wl{}, wr{};
  int32_t rc;
  if ((rc = prepare_wave(h, h->land, dt, h->cfg.dt_land, 0, 2, wl, h->sub_land, "update_overland_flow_model"))) return rc;
  if ((rc = prepare_wave(h, h->river, dt, h->cfg.dt_river, 1, 1, wr, h->sub_river, "update_river_flow_model"))) return rc;
  wl.smem = wr.smem = h->smem_surface;
  wl.smem_per_warp = wr.smem_per_warp = h->smem_surface_per_warp;
  const int64_t warps_needed = h->land.nw.n_chunks + h->river.nw.n_chunks;
  wl.grid = wr.grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->grid_surface * h->tune.grid_percent / 100,
                                                                 (warps_needed + 7) / 8));
  SurfaceSync sync{};
  sync.land_done = h->land.chunk_done;
  sync.epoch = ++h->surface_epoch;
  sync.land_chunk_of_slot = h->land.chunk_of_slot;
  sync.period = h->tune.river_period;
  sync.river_share = h->tune.river_share;
  sync.err = h->d_err;
  return check_launch(h, launch_surface_wave(h->f, h->kc, h->land.dev, h->river.dev, wl, wr, sync, false, h->stream),
                      "update_overland_flow_model + update_river_flow_model");
}

// The subsurface sweep (with update_soil_water_storage! fused into it) and the surface kernel
// OVERLAPPED: the subsurface kernel runs on part of the SMs and flags every land chunk whose
// results are final; the surface kernel is launched programmatically dependent (it may start as
// soon as every CTA of the subsurface grid is resident, so it can only take the SMs that grid
// leaves free and the subsurface sweep always makes progress: no deadlock), its overland warps
// wait for their chunk's flag. Three wavefronts then follow each other through the levels.
static int32_t update_routing_overlapped(WflowB200* h, double dt, bool* done) {
  *done = false;
  // pays while the sweeps are bound by their dependent chains; measured step times with / without
  // the overlap: 1000^2 2.5 / 3.1 ms, 1500^2 5.03 / 5.21, 2000^2 8.54 / 8.24, 2500^2 12.3 / 12.1,
  // 3536^2 24.8 / 22.4 -> on below 1750 nodes per level
  const bool wide = h->land.nw.n_wave_levels > 0 &&
                    h->land.nw.n / h->land.nw.n_wave_levels >= WFB_OVERLAP_MAX_LEVEL_WIDTH;
  const bool want = h->tune.overlap_ssf >= 0 ? h->tune.overlap_ssf != 0 : !wide;
  if (!want || !h->fuse_surface || !h->tune.fuse_surface || h->cfg.adaptive) return WFLOWB200_OK;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device);
  // SMs (= CTAs) of the subsurface sweep; measured at 1000^2 (routing ms): 44 SMs 2.38, 52: 1.91, 60: 1.83, 64: 1.89, 74: 2.11, 84: 2.38
  int ssf_ctas = h->tune.ssf_overlap_sms > 0 ? h->tune.ssf_overlap_sms : (sms * 13) / 32;
  if (ssf_ctas < 1 || ssf_ctas >= sms) ssf_ctas = (sms * 13) / 32;
  if (h->tune.grid_percent < 100) {  // the handle shares the GPU with other handles
    sms = std::max(2, sms * h->tune.grid_percent / 100);
    ssf_ctas = std::max(1, std::min(ssf_ctas * h->tune.grid_percent / 100, sms - 1));
  }
  WaveLaunch ws{}, wl{}, wr{};
  int32_t rc;
  if ((rc = prepare_wave(h, h->land, dt, h->cfg.dt_land, 0, 2, wl, h->sub_land, "update_overland_flow_model"))) return rc;
  if ((rc = prepare_wave(h, h->river, dt, h->cfg.dt_river, 1, 1, wr, h->sub_river, "update_river_flow_model"))) return rc;
  // the subsurface flow walks the same land chunks as the overland flow: its own outlet buffer
  std::vector<double> dts;
  const int S = fixed_substeps(dt, h->cfg.dt_ssf, dts);
  if (S <= 0) return fail(h, WFLOWB200_ERR_ARG, "update_subsurface_flow_model: bad internal time step");
  const size_t need_ssf = (size_t)std::max<int64_t>(h->land.nw.n_outlets, 1) * 2 * (size_t)S;
  if (need_ssf > h->ssf_q_out_words) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->ssf_q_out);
    h->ssf_q_out = nullptr;
    h->ssf_q_out_words = 0;
    CUDA_TRY(h, cudaMalloc((void**)&h->ssf_q_out, need_ssf * sizeof(unsigned long long)));
    h->ssf_q_out_words = need_ssf;
  }
  ws.queue = h->d_queue + 2 * 32;
  ws.q_out = h->ssf_q_out;
  ws.stats = h->d_stats;
  ws.S = S;
  ws.dt_fixed = dts[0];
  ws.dt_last = dts[S - 1];
  ws.dt = dt;
  ws.smem = h->smem_ssf;
  ws.err = h->d_err;
  if ((rc = bind_exchange(h, 2, ws))) return rc;
  ws.grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(ssf_ctas, h->grid_ssf), h->land.nw.n_chunks));
  ws.fuse_soil_storage = 1;
  ws.done_flags = h->land.chunk_ssf_done;
  ws.done_epoch = ++h->surface_epoch;
  ws.trigger_dependents = 1;
  h->sub_ssf = S;
  wl.smem = wr.smem = h->smem_surface;
  wl.smem_per_warp = wr.smem_per_warp = h->smem_surface_per_warp;
  const int64_t warps_needed = h->land.nw.n_chunks + h->river.nw.n_chunks;
  const int64_t room = (int64_t)(sms - ws.grid) * 2;  // two surface CTAs per free SM
  wl.grid = wr.grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(h->grid_surface, room),
                                                                 (warps_needed + 7) / 8));
  SurfaceSync sync{};
  sync.land_done = h->land.chunk_done;
  sync.ssf_done = h->land.chunk_ssf_done;
  sync.epoch = h->surface_epoch;
  sync.land_chunk_of_slot = h->land.chunk_of_slot;
  sync.period = h->tune.river_period;
  sync.river_share = h->tune.river_share;
  sync.err = h->d_err;
  // every reset before the first kernel: nothing may sit between the two launches
  reset_surface_wave(h->land.dev, h->river.dev, wl, wr, h->stream);
  if ((rc = check_launch(h, launch_subsurface_wave(h->f, h->kc, h->land.dev, h->N, ws, h->stream),
                         "update_subsurface_flow_model"))) return rc;
  rc = check_launch(h, launch_surface_wave(h->f, h->kc, h->land.dev, h->river.dev, wl, wr, sync, true, h->stream),
                    "update_overland_flow_model + update_river_flow_model");
  *done = rc == WFLOWB200_OK;
  return rc;
}

int32_t wflowb200_update_total_water_storage(WflowB200* h) {
  WFB_ENTER(h);
  return check_launch(h, launch_total_water_storage(h->f, h->kc, h->riv_of_land, h->stream),
                      "update_total_water_storage");
}

// stage timing of update_model (set_timing): the ten events bracket the stages
static int32_t book_stage_times(WflowB200* h) {
  if (!h->timing) return WFLOWB200_OK;
  CUDA_TRY(h, cudaEventSynchronize(h->ev[9]));
  float d[9];
  for (int i = 0; i < 9; ++i) cudaEventElapsedTime(&d[i], h->ev[i], h->ev[i + 1]);
  h->ms[0] += d[1]; h->ms[1] += d[3]; h->ms[2] += d[4]; h->ms[3] += d[5]; h->ms[4] += d[7];
  h->ms[5] += d[8]; h->ms[6] += d[0] + d[2] + d[6];
  h->timed_steps++;
  return WFLOWB200_OK;
}

static int32_t update_model_step(WflowB200* h, double dt);
int32_t wflowb200_update_model(WflowB200* h, double dt) {
  WFB_ENTER(h);
  int32_t rc = exchange_begin_step(h);
  if (rc) return rc;
  h->xchg.in_update_model = true;
  rc = update_model_step(h, dt);
  h->xchg.in_update_model = false;
  if (rc == WFLOWB200_OK) h->xchg.step++;
  return rc;
}
static int32_t update_model_step(WflowB200* h, double dt) {
  int32_t rc;
  auto mark = [&](int i) { if (h->timing) cudaEventRecord(h->ev[i], h->stream); };
  mark(0);
  if ((rc = wait_forcing(h))) return rc;
  if (h->nriv > 0) {  // river h -> land grid (runoff.jl:77-79)
    if ((rc = check_launch(h, launch_scatter_river_depth(h->f, h->kc, h->stream), "scatter"))) return rc;
  }
  mark(1);
  if ((rc = launch_vertical(h, dt))) return rc;
  mark(2);
  // wflowb200_exchange_recharge: soil_column_kernel has already written recharge_rate and the
  // subsurface water table depth (the separate entry point stays for the fine-grained sequence)
  mark(3);
  bool fused = false;
  if ((rc = update_routing_overlapped(h, dt, &fused))) return rc;
  if (fused) {  // the three wavefronts overlapped (timed as "subsurface")
    mark(4); mark(5); mark(6); mark(7); mark(8);
    if ((rc = wflowb200_update_total_water_storage(h))) return rc;
    mark(9);
    return book_stage_times(h);
  }
  bool soil_fused = false;
  if ((rc = update_subsurface_and_soil_storage(h, dt, &soil_fused))) return rc;
  mark(4);
  if (!soil_fused && (rc = wflowb200_update_soil_water_storage(h, dt))) return rc;  // also fills olf_inwater
  mark(5);
  if (h->cfg.land_routing == 1) {  // surface_routing! with local-inertial land AND river routing
    if ((rc = wflowb200_update_bc_overland_flow_model(h))) return rc;   // surface_routing.jl:62-86
    if ((rc = wflowb200_update_inflow_reservoir(h))) return rc;
    if ((rc = wflowb200_update_overland_flow_model(h, dt))) return rc;  // (timed as "overland")
    mark(6); mark(7); mark(8);
    if ((rc = wflowb200_update_total_water_storage(h))) return rc;
    mark(9);
    return book_stage_times(h);
  }
  if (h->fuse_surface && h->tune.fuse_surface) {  // overland and river wavefronts overlapped (timed as "overland")
    if ((rc = update_surface_fused(h, dt))) return rc;
    mark(6);
    mark(7);
  } else {
    if ((rc = wflowb200_update_overland_flow_model(h, dt))) return rc;
    mark(6);
    if ((rc = wflowb200_update_lateral_inflow_river(h))) return rc;
    if ((rc = wflowb200_update_inflow_reservoir(h))) return rc;
    mark(7);
    if ((rc = wflowb200_update_river_flow_model(h, dt))) return rc;
  }
  mark(8);
  if ((rc = wflowb200_update_total_water_storage(h))) return rc;
  mark(9);
  return book_stage_times(h);
}

int32_t wflowb200_selftest_math(int32_t device, int64_t n, double* out6) {
  if (!out6 || n <= 0) return WFLOWB200_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess)
    return fail(nullptr, WFLOWB200_ERR_CUDA, "no CUDA device: libwflow_b200 has no CPU fallback");
  unsigned long long* d = nullptr;
  if (cudaMalloc((void**)&d, 6 * sizeof(unsigned long long)) != cudaSuccess)
    return fail(nullptr, WFLOWB200_ERR_CUDA, "cudaMalloc failed");
  cudaMemset(d, 0, 6 * sizeof(unsigned long long));
  launch_selftest_math(n, d, nullptr);
  const cudaError_t e = cudaMemcpy(out6, d, 6 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(nullptr, WFLOWB200_ERR_CUDA, cudaGetErrorString(e));
  return WFLOWB200_OK;
}

int32_t wflowb200_synchronize(WflowB200* h) {
  WFB_ENTER(h);
  return check_device_error(h);
}

int32_t wflowb200_newton_trace(WflowB200* h, int32_t enable) {
  WFB_ENTER(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(h->f.olf_newton_trace); cudaFree(h->f.riv_newton_trace);
  h->f.olf_newton_trace = h->f.riv_newton_trace = nullptr;
  if (enable) {
    CUDA_TRY(h, cudaMalloc((void**)&h->f.olf_newton_trace, (size_t)h->ns * sizeof(int32_t)));
    CUDA_TRY(h, cudaMalloc((void**)&h->f.riv_newton_trace, (size_t)std::max(h->nrs, 32) * sizeof(int32_t)));
    CUDA_TRY(h, cudaMemset(h->f.olf_newton_trace, 0, (size_t)h->ns * sizeof(int32_t)));
    CUDA_TRY(h, cudaMemset(h->f.riv_newton_trace, 0, (size_t)std::max(h->nrs, 32) * sizeof(int32_t)));
  }
  return WFLOWB200_OK;
}

int32_t wflowb200_get_newton_trace(WflowB200* h, int32_t domain, int64_t* dst) {
  WFB_ENTER(h);
  if (!dst) return WFLOWB200_ERR_ARG;
  const bool riv = domain == WFLOWB200_DOMAIN_RIVER;
  const int32_t* src = riv ? h->f.riv_newton_trace : h->f.olf_newton_trace;
  if (!src) return fail(h, WFLOWB200_ERR_STATE, "newton trace is not enabled");
  const DomainDev& d = riv ? h->river : h->land;
  const int n = riv ? h->nriv : h->n;
  std::vector<int32_t> tmp(n);
  CUDA_TRY(h, cudaMemcpyAsync(tmp.data(), src, tmp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost,
                              h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int p = 0; p < n; ++p) dst[d.nw.perm[p] - 1] = tmp[p];
  return WFLOWB200_OK;
}

int32_t wflowb200_comm_unique_id(char* out128) {
  if (!out128) return WFLOWB200_ERR_ARG;
  if (!g_nccl.load()) return fail(nullptr, WFLOWB200_ERR_STATE, "libnccl.so.2 not found");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return fail(nullptr, WFLOWB200_ERR_CUDA, "ncclGetUniqueId failed");
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return WFLOWB200_OK;
}

int32_t wflowb200_comm_init_nccl(WflowB200* h, int32_t rank, int32_t world, const char* id128) {
  WFB_ENTER(h);
  if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(h, WFLOWB200_ERR_ARG, "bad rank / world");
  if (!g_nccl.load()) return fail(h, WFLOWB200_ERR_STATE, "libnccl.so.2 not found");
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NcclShardComm* c = new NcclShardComm();
  if (g_nccl.CommInitRank(&c->comm, world, id, rank) != ncclSuccess) {
    c->comm = nullptr;
    delete c;
    return fail(h, WFLOWB200_ERR_CUDA, "ncclCommInitRank failed");
  }
  delete h->comm;
  h->comm = c;
  return WFLOWB200_OK;
}

int32_t wflowb200_group_create(int32_t n_handles, WflowB200Group** out) {
  if (!out || n_handles < 1) return WFLOWB200_ERR_ARG;
  *out = new WflowB200Group();
  (*out)->n = n_handles;
  return WFLOWB200_OK;
}
void wflowb200_group_destroy(WflowB200Group* g) { delete g; }
int32_t wflowb200_group_join(WflowB200Group* g, WflowB200* h) {
  WFB_ENTER(h);
  if (!g) return WFLOWB200_ERR_ARG;
  GroupShardComm* c = new GroupShardComm();
  c->g = g;
  delete h->comm;
  h->comm = c;
  return WFLOWB200_OK;
}

// ---- cut edges ------------------------------------------------------------------------------
static void free_exchange(WflowB200* h) {
  Exchange& x = h->xchg;
  for (auto& kv : x.peers)
    if (kv.second.ipc && kv.second.base) cudaIpcCloseMemHandle(kv.second.base);
  x.peers.clear();
  for (auto& k : x.d_exp)
    for (auto& t : k) { cudaFree(t); t = nullptr; }
  cudaFree(x.imp); x.imp = nullptr;
  cudaFree(x.d_token); x.d_token = nullptr;
}

int32_t wflowb200_exchange_prepare(WflowB200* h, double dt, uint64_t* device_ptr,
                                   void* ipc_handle64, int64_t* bytes) {
  WFB_ENTER(h);
  Exchange& x = h->xchg;
  if (x.prepared) return fail(h, WFLOWB200_ERR_STATE, "exchange_prepare: called twice");
  if (h->cfg.adaptive) return fail(h, WFLOWB200_ERR_ARG, "exchange: fixed internal time steps only");
  std::vector<double> dts;
  const double fixed[3] = {h->cfg.dt_land, h->cfg.dt_river, h->cfg.dt_ssf};
  for (int k = 0; k < 3; ++k) {
    x.S[k] = fixed_substeps(dt, fixed[k], dts);
    if (x.S[k] <= 0) return fail(h, WFLOWB200_ERR_ARG, "exchange_prepare: bad internal time step");
  }
  x.lay.set(x.lay.n_imp[0], x.lay.n_imp[1], x.S);
  CUDA_TRY(h, cudaMalloc((void**)&x.d_token, sizeof(unsigned long long)));
  CUDA_TRY(h, cudaMemset(x.d_token, 0, sizeof(unsigned long long)));
  if (x.lay.words > 0) {
    CUDA_TRY(h, cudaMalloc((void**)&x.imp, 2 * x.lay.words * sizeof(unsigned long long)));
    CUDA_TRY(h, cudaMemset(x.imp, 0xff, 2 * x.lay.words * sizeof(unsigned long long)));
    CUDA_TRY(h, cudaDeviceSynchronize());
    if (ipc_handle64) {
      static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
      cudaIpcMemHandle_t ih;
      CUDA_TRY(h, cudaIpcGetMemHandle(&ih, x.imp));
      memcpy(ipc_handle64, &ih, 64);
    }
  } else if (ipc_handle64) {
    memset(ipc_handle64, 0, 64);
  }
  if (device_ptr) *device_ptr = (uint64_t)(uintptr_t)x.imp;
  if (bytes) *bytes = (int64_t)(2 * x.lay.words * sizeof(unsigned long long));
  x.prepared = true;
  x.tables_current = false;
  return WFLOWB200_OK;
}

int32_t wflowb200_exchange_open_peer(WflowB200* h, int32_t peer, uint64_t device_ptr,
                                     int32_t peer_device, const void* ipc_handle64,
                                     int64_t peer_n_land_imports, int64_t peer_n_river_imports) {
  WFB_ENTER(h);
  Exchange& x = h->xchg;
  if (!x.prepared) return fail(h, WFLOWB200_ERR_STATE, "exchange_open_peer: exchange_prepare first");
  if (peer < 0 || peer >= 64 || x.peers.count(peer) || peer_n_land_imports < 0 || peer_n_river_imports < 0)
    return fail(h, WFLOWB200_ERR_ARG, "exchange_open_peer: bad or repeated peer");
  ExchangePeer pr;
  pr.lay.set(peer_n_land_imports, peer_n_river_imports, x.S);
  if (device_ptr) {
    if (peer_device != h->cfg.device) {
      int can = 0;
      CUDA_TRY(h, cudaDeviceCanAccessPeer(&can, h->cfg.device, peer_device));
      if (!can) return fail(h, WFLOWB200_ERR_CUDA, "exchange_open_peer: no peer access between the devices");
      const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(h, WFLOWB200_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
    pr.base = (unsigned long long*)(uintptr_t)device_ptr;
  } else {
    if (!ipc_handle64) return fail(h, WFLOWB200_ERR_ARG, "exchange_open_peer: no pointer and no IPC handle");
    cudaIpcMemHandle_t ih;
    memcpy(&ih, ipc_handle64, 64);
    void* ptr = nullptr;
    CUDA_TRY(h, cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
    pr.base = (unsigned long long*)ptr;
    pr.ipc = true;
  }
  x.peers[peer] = pr;
  x.tables_current = false;
  return WFLOWB200_OK;
}

int32_t wflowb200_exchange_bind(WflowB200* h, int32_t domain, int64_t export_index, int32_t peer,
                                int64_t peer_import_index) {
  WFB_ENTER(h);
  Exchange& x = h->xchg;
  if (domain < 0 || domain > 1 || export_index < 0 || export_index >= x.n_exp[domain])
    return fail(h, WFLOWB200_ERR_ARG, "exchange_bind: no such export");
  auto it = x.peers.find(peer);
  if (it == x.peers.end() || peer_import_index < 0 || peer_import_index >= it->second.lay.n_imp[domain])
    return fail(h, WFLOWB200_ERR_ARG, "exchange_bind: no such peer or import");
  x.bound[domain][export_index] = {peer, peer_import_index};
  x.tables_current = false;
  return WFLOWB200_OK;
}

// per (kind, parity): where every export of the kind's domain stores its S x NV values
static int32_t build_export_tables(WflowB200* h) {
  Exchange& x = h->xchg;
  static const int kDomainOfKind[3] = {0, 1, 0}, kNv[3] = {2, 1, 2};
  for (int kind = 0; kind < 3; ++kind) {
    const int dom = kDomainOfKind[kind];
    const int64_t ne = x.n_exp[dom];
    for (int par = 0; par < 2; ++par) {
      cudaFree(x.d_exp[kind][par]);
      x.d_exp[kind][par] = nullptr;
      if (ne == 0) continue;
      std::vector<unsigned long long*> tab(ne);
      for (int64_t e = 0; e < ne; ++e) {
        const auto& b = x.bound[dom][e];
        if (b.first < 0) return fail(h, WFLOWB200_ERR_STATE, "exchange: an export is not bound to a peer");
        const ExchangePeer& pr = x.peers[b.first];
        tab[e] = pr.base + (size_t)par * pr.lay.words + pr.lay.kind_off[kind] +
                 (size_t)b.second * x.S[kind] * kNv[kind];
      }
      CUDA_TRY(h, cudaMalloc((void**)&x.d_exp[kind][par], ne * sizeof(unsigned long long*)));
      CUDA_TRY(h, cudaMemcpy(x.d_exp[kind][par], tab.data(), ne * sizeof(unsigned long long*),
                             cudaMemcpyHostToDevice));
    }
  }
  x.tables_current = true;
  return WFLOWB200_OK;
}

// Start of a model step of a shard that takes part in an exchange: every shard has finished the
// previous step (barrier), then the import slots of the NEXT step are emptied -- the producers
// write them after the next barrier at the earliest, this step's were emptied a step ago.
static int32_t exchange_begin_step(WflowB200* h) {
  Exchange& x = h->xchg;
  if (!x.prepared) {
    if (x.active()) return fail(h, WFLOWB200_ERR_STATE, "a shard with cut edges needs wflowb200_exchange_prepare");
    return WFLOWB200_OK;
  }
  if (!h->comm) return fail(h, WFLOWB200_ERR_STATE, "exchange: the shards need a communicator (comm_init_nccl / group_join)");
  if (!x.tables_current) {
    const int32_t rc = build_export_tables(h);
    if (rc) return rc;
  }
  if (h->comm->allreduce(x.d_token, 1, 0, h->stream))
    return fail(h, WFLOWB200_ERR_CUDA, "exchange: barrier failed");
  if (x.imp) {
    const int next = (int)((x.step + 1) & 1);
    CUDA_TRY(h, cudaMemsetAsync(x.imp + (size_t)next * x.lay.words, 0xff,
                                x.lay.words * sizeof(unsigned long long), h->stream));
  }
  return WFLOWB200_OK;
}

int32_t wflowb200_get_unsat_buckets(WflowB200* h, int64_t* out, int32_t capacity) {
  WFB_ENTER(h);
  if (!out || capacity < 2 * WFB_UNSAT_BUCKETS) return WFLOWB200_ERR_ARG;
  unsigned cnt[2 * WFB_UNSAT_BUCKETS];
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemcpy(cnt, h->d_unsat_count, sizeof(cnt), cudaMemcpyDeviceToHost));
  for (int b = 0; b < 2 * WFB_UNSAT_BUCKETS; ++b) out[b] = cnt[b];
  return WFLOWB200_OK;
}

int32_t wflowb200_get_vertical_timeline(WflowB200* h, double* out_ms, int32_t capacity) {
  WFB_ENTER(h);
  const int n = 4;
  if (!out_ms || capacity < n) return fail(h, WFLOWB200_ERR_ARG, "timeline buffer too small");
  if (!h->tune.timeline || h->tune.use_graph || !h->tl_ev[0])
    return fail(h, WFLOWB200_ERR_STATE, "set vertical_timeline = 1 and vertical_graph = 0 first");
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int k = 0; k < n; ++k) {
    float ms = 0.f;
    out_ms[k] = cudaEventElapsedTime(&ms, h->tl_ev[0], h->tl_ev[k]) == cudaSuccess ? ms : -1.0;
  }
  cudaGetLastError();
  return WFLOWB200_OK;
}

int32_t wflowb200_set_option(WflowB200* h, const char* name, int32_t value) {
  WFB_ENTER(h);
  if (!name) return WFLOWB200_ERR_ARG;
  Tuning& t = h->tune;
  const std::string k(name);
  if (k == "fuse_soil_storage") t.fuse_soil_storage = value;
  else if (k == "overlap_subsurface") t.overlap_ssf = value;
  else if (k == "overlap_subsurface_sms") t.ssf_overlap_sms = value;
  else if (k == "fuse_surface") t.fuse_surface = value != 0;
  else if (k == "surface_river_share") { if (value >= 1 && value < t.river_period) t.river_share = value; }
  else if (k == "surface_river_period") { if (value >= 2 && value > t.river_share) t.river_period = value; }
  else if (k == "vertical_graph") t.use_graph = value != 0;
  else if (k == "vertical_timeline") {
    t.timeline = value != 0;
    for (auto& e : h->tl_ev)
      if (!e && t.timeline) CUDA_TRY(h, cudaEventCreate(&e));
  }
  else if (k == "vertical_engine") {   // timing experiments only: 0 leaves suspended cells unfinished
    t.run_engine = value != 0;
    if (h->v_graph) { cudaGraphExecDestroy(h->v_graph); h->v_graph = nullptr; }
  }
  else if (k == "wave_grid_percent") { if (value >= 1 && value <= 100) t.grid_percent = value; }
  else if (k == "kinwave_root_each_substep") h->kc.kw_root_each_substep = value != 0;
  else return fail(h, WFLOWB200_ERR_ARG, "unknown option: " + k);
  return WFLOWB200_OK;
}

int32_t wflowb200_get_artifact(WflowB200* h, int32_t domain, int32_t id, int64_t* dst,
                               int64_t capacity, int64_t* len_out) {
  if (!h || !len_out) return WFLOWB200_ERR_ARG;  // host data only: no device guard
  const Network& nw = domain == WFLOWB200_DOMAIN_RIVER ? h->river.nw : h->land.nw;
  return copy_artifact(nw, id, dst, capacity, len_out, h->err);
}

int32_t wflowb200_network_build(const WflowB200Config* cfg, const WflowB200Domain* dom,
                                WflowB200Network** out) {
  if (!cfg || !dom || !out) return fail(nullptr, WFLOWB200_ERR_ARG, "null argument");
  WflowB200Network* net = new WflowB200Network();
  int32_t rc = build_networks(cfg, dom, net->land, net->river, g_create_error);
  if (rc) { delete net; *out = nullptr; return rc; }
  *out = net;
  return WFLOWB200_OK;
}
int32_t wflowb200_network_get(const WflowB200Network* net, int32_t domain, int32_t id,
                              int64_t* dst, int64_t capacity, int64_t* len_out) {
  if (!net || !len_out) return WFLOWB200_ERR_ARG;
  const Network& nw = domain == WFLOWB200_DOMAIN_RIVER ? net->river : net->land;
  return copy_artifact(nw, id, dst, capacity, len_out, g_create_error);
}
void wflowb200_network_destroy(WflowB200Network* net) { delete net; }

int32_t wflowb200_get_stats(WflowB200* h, WflowB200Stats* out) {
  WFB_ENTER(h);
  if (!out) return WFLOWB200_ERR_ARG;
  RoutingStats rs{};
  { int32_t rc = check_device_error(h); if (rc) return rc; }
  CUDA_TRY(h, cudaMemcpy(&rs, h->d_stats, sizeof(rs), cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof(*out));
  out->newton_calls_land = (int64_t)rs.newton_calls_land;
  out->newton_iters_land = (int64_t)rs.newton_iters_land;
  out->newton_maxit_land = (int64_t)rs.newton_maxit_land;
  out->newton_calls_river = (int64_t)rs.newton_calls_river;
  out->newton_iters_river = (int64_t)rs.newton_iters_river;
  out->newton_maxit_river = (int64_t)rs.newton_maxit_river;
  if (h->d_li_substeps && h->cfg.river_routing == 1) {
    int cnt = 0;
    cudaMemcpy(&cnt, h->d_li_substeps, sizeof(int), cudaMemcpyDeviceToHost);
    h->sub_river = cnt;
    if (h->cfg.land_routing == 1) h->sub_land = cnt;
  }
  out->substeps_land = h->sub_land; out->substeps_river = h->sub_river;
  out->substeps_ssf = h->sub_ssf;
  out->wave_levels_land = h->land.nw.n_wave_levels;
  out->wave_levels_river = h->river.nw.n_wave_levels;
  out->kernel_launches = h->launches;
  out->ms_land_hydrology = h->ms[0]; out->ms_subsurface = h->ms[1];
  out->ms_soil_storage = h->ms[2]; out->ms_overland = h->ms[3]; out->ms_river = h->ms[4];
  out->ms_total_storage = h->ms[5]; out->ms_glue = h->ms[6];
  out->timed_steps = h->timed_steps;
  return WFLOWB200_OK;
}

int32_t wflowb200_set_timing(WflowB200* h, int32_t enabled) {
  WFB_ENTER(h);
  h->timing = enabled != 0;
  for (auto& m : h->ms) m = 0.0;
  h->timed_steps = 0;
  return WFLOWB200_OK;
}

int32_t wflowb200_timer_start(WflowB200* h) {
  WFB_ENTER(h);
  CUDA_TRY(h, cudaEventRecord(h->tm[0], h->stream));
  return WFLOWB200_OK;
}
int32_t wflowb200_timer_stop(WflowB200* h, double* elapsed_ms) {
  WFB_ENTER(h);
  if (!elapsed_ms) return WFLOWB200_ERR_ARG;
  CUDA_TRY(h, cudaEventRecord(h->tm[1], h->stream));
  CUDA_TRY(h, cudaEventSynchronize(h->tm[1]));
  float ms = 0.f;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->tm[0], h->tm[1]));
  *elapsed_ms = ms;
  return WFLOWB200_OK;
}

}  // extern "C"
