// network.cpp -- see network.hpp. Host C++ (init-time only, O(n)).
#include "network.hpp"

#include <algorithm>
#include <cmath>
#include <utility>

namespace wfb {
namespace {

// PCRaster LDD (1..9) -> CartesianIndex offset (utils.jl:2-12)
const int kDi[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};
const int kDj[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1};
constexpr uint8_t kPit = 5;

// CSR of in-neighbours, ascending source id inside each list.
void build_in_csr(const std::vector<int64_t>& down, std::vector<int64_t>& ptr,
                  std::vector<int64_t>& idx) {
  const int64_t n = (int64_t)down.size();
  ptr.assign(n + 1, 0);
  for (int64_t v = 0; v < n; ++v)
    if (down[v]) ptr[down[v]]++;
  for (int64_t v = 0; v < n; ++v) ptr[v + 1] += ptr[v];
  idx.assign(ptr[n], 0);
  std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
  for (int64_t v = 0; v < n; ++v)  // ascending v => ascending lists
    if (down[v]) idx[fill[down[v] - 1]++] = v + 1;
}

// Graphs.jl topological_sort_by_dfs on an out-degree<=1 graph: from every still-white vertex
// (ascending id) walk downstream over white vertices; the walked chain finishes in reverse;
// the final order is the reverse of the finishing sequence. Detects cycles (grey hit).
bool toposort_dfs(const std::vector<int64_t>& down, std::vector<int64_t>& order) {
  const int64_t n = (int64_t)down.size();
  std::vector<uint8_t> color(n, 0);
  std::vector<int64_t> verts;
  verts.reserve(n);
  std::vector<int64_t> chain;
  for (int64_t v = 0; v < n; ++v) {
    if (color[v]) continue;
    chain.clear();
    int64_t u = v;
    for (;;) {
      color[u] = 1;
      chain.push_back(u);
      const int64_t d = down[u];
      if (d == 0) break;
      if (color[d - 1] == 1) return false;  // "The input graph contains at least one loop."
      if (color[d - 1] == 2) break;
      u = d - 1;
    }
    for (auto it = chain.rbegin(); it != chain.rend(); ++it) {
      color[*it] = 2;
      verts.push_back(*it + 1);
    }
  }
  order.assign(verts.rbegin(), verts.rend());
  return true;
}

void stream_order(const Network& nw, std::vector<int64_t>& so) {
  so.assign(nw.n, 1);
  for (int64_t v1 : nw.order) {
    const int64_t a = nw.in_ptr[v1 - 1], b = nw.in_ptr[v1];
    if (b > a) {
      int64_t mx = 0, cnt = 0;
      for (int64_t e = a; e < b; ++e) {
        const int64_t s = so[nw.in_idx[e] - 1];
        if (s > mx) { mx = s; cnt = 1; }
        else if (s == mx) ++cnt;
      }
      so[v1 - 1] = cnt > 1 ? mx + 1 : mx;
    }
  }
}

// One basin of kinwave_set_subdomains, in local ids 1..k (ascending global id).
struct BasinWork {
  std::vector<int64_t> down, order, so, subbas, fill, in_ptr, in_idx;
};

}  // namespace

bool build_graph(Network& nw, int64_t d1, int64_t d2, const int64_t* indices, const uint8_t* ldd,
                 int64_t n, std::string& err) {
  nw.n = n;
  nw.ldd.assign(ldd, ldd + n);
  nw.down.assign(n, 0);
  std::vector<int64_t> lin(n);
  for (int64_t v = 0; v < n; ++v) {
    const int64_t i = indices[2 * v], j = indices[2 * v + 1];
    if (i < 1 || i > d1 || j < 1 || j > d2) { err = "index outside raster"; return false; }
    lin[v] = (j - 1) * d1 + (i - 1);
    if (v && lin[v] <= lin[v - 1]) { err = "indices must be column-major ascending"; return false; }
  }
  for (int64_t v = 0; v < n; ++v) {
    const uint8_t l = nw.ldd[v];
    if (l == kPit) continue;
    bool ok = l >= 1 && l <= 9;
    int64_t to = 0;
    if (ok) {
      const int64_t ti = indices[2 * v] + kDi[l - 1], tj = indices[2 * v + 1] + kDj[l - 1];
      ok = ti >= 1 && ti <= d1 && tj >= 1 && tj <= d2;
      if (ok) {
        const int64_t tl = (tj - 1) * d1 + (ti - 1);
        auto it = std::lower_bound(lin.begin(), lin.end(), tl);
        ok = it != lin.end() && *it == tl;
        to = (int64_t)(it - lin.begin()) + 1;
      }
    }
    if (!ok) { nw.ldd[v] = kPit; continue; }  // invalid direction -> pit (routing/utils.jl:20-24)
    nw.down[v] = to;
  }
  build_in_csr(nw.down, nw.in_ptr, nw.in_idx);
  if (!toposort_dfs(nw.down, nw.order)) {
    err = "One or more cycles detected in flow graph.";
    return false;
  }
  return true;
}

// NEIGHBORS (routing/subsurface/connectivity.jl:61-66) paired with DIRS (network.jl:3):
// (0,-1) ind_y_down, (-1,0) ind_x_down, (1,0) ind_x_up, (0,1) ind_y_up
bool build_edge_connectivity(Network& nw, int64_t d1, int64_t d2, const int64_t* indices, int64_t n,
                             std::string& err) {
  std::vector<int64_t> lin(n);
  for (int64_t v = 0; v < n; ++v) {
    const int64_t i = indices[2 * v], j = indices[2 * v + 1];
    if (i < 1 || i > d1 || j < 1 || j > d2) { err = "index outside raster"; return false; }
    lin[v] = (j - 1) * d1 + (i - 1);
  }
  auto neighbour = [&](int64_t v, int di, int dj) -> int64_t {
    const int64_t ti = indices[2 * v] + di, tj = indices[2 * v + 1] + dj;
    if (ti < 1 || ti > d1 || tj < 1 || tj > d2) return n + 1;
    const int64_t tl = (tj - 1) * d1 + (ti - 1);
    auto it = std::lower_bound(lin.begin(), lin.end(), tl);   // indices are column-major ascending
    return (it != lin.end() && *it == tl) ? (int64_t)(it - lin.begin()) + 1 : n + 1;
  };
  nw.edge_x_up.resize(n); nw.edge_x_down.resize(n); nw.edge_y_up.resize(n); nw.edge_y_down.resize(n);
  for (int64_t v = 0; v < n; ++v) {
    nw.edge_y_down[v] = neighbour(v, 0, -1);
    nw.edge_x_down[v] = neighbour(v, -1, 0);
    nw.edge_x_up[v] = neighbour(v, 1, 0);
    nw.edge_y_up[v] = neighbour(v, 0, 1);
  }
  return true;
}

bool build_artifacts(Network& nw, int nthreads, int min_sto, const int64_t* so_override,
                     std::string& err) {
  const int64_t n = nw.n;
  if (so_override) nw.streamorder.assign(so_override, so_override + n);
  else stream_order(nw, nw.streamorder);

  // upstream_nodes, indexed by toposort position (utils.jl:61-71)
  nw.up_ptr.assign(n + 1, 0);
  nw.up_idx.clear();
  nw.up_idx.reserve(nw.in_idx.size());
  for (int64_t k = 0; k < n; ++k) {
    const int64_t v1 = nw.order[k];
    for (int64_t e = nw.in_ptr[v1 - 1]; e < nw.in_ptr[v1]; ++e) nw.up_idx.push_back(nw.in_idx[e]);
    nw.up_ptr[k + 1] = (int64_t)nw.up_idx.size();
  }

  // ---- sub-domain partition (subdomains.jl:169-255) ------------------------------------
  nw.lvl_ptr.clear(); nw.lvl_idx.clear(); nw.sub_ptr.clear(); nw.sub_order.clear();
  nw.sub_indices.clear();
  if (nthreads <= 1 || n == 0) {
    nw.lvl_ptr = {0, 1};
    nw.lvl_idx = {1};
    nw.sub_ptr = {0, n};
    nw.sub_order = nw.order;
    nw.sub_indices.resize(n);
    for (int64_t k = 0; k < n; ++k) nw.sub_indices[k] = k + 1;
  } else {
    std::vector<int64_t> index_toposort(n);
    for (int64_t k = 0; k < n; ++k) index_toposort[nw.order[k] - 1] = k + 1;
    // basins: pits numbered in ascending node id; labels pushed upstream
    std::vector<int64_t> basin(n, 0);
    int64_t n_pits = 0;
    for (int64_t v = 0; v < n; ++v)
      if (nw.ldd[v] == kPit) basin[v] = ++n_pits;
    for (int64_t k = n - 1; k >= 0; --k) {
      const int64_t v = nw.order[k] - 1, d = nw.down[v];
      if (d && basin[v] == 0 && basin[d - 1] != 0) basin[v] = basin[d - 1];
    }
    // group nodes per basin, ascending id (counting sort)
    std::vector<int64_t> bptr(n_pits + 2, 0);
    for (int64_t v = 0; v < n; ++v) bptr[basin[v] + 1]++;  // basin 0 (unreachable) -> slot 1
    for (int64_t b = 0; b <= n_pits; ++b) bptr[b + 1] += bptr[b];
    std::vector<int64_t> bnodes(n), bfill(bptr.begin(), bptr.end() - 1);
    for (int64_t v = 0; v < n; ++v) bnodes[bfill[basin[v]]++] = v;  // 0-based global ids
    std::vector<int64_t> local_of(n, 0);

    // per-level lists of (global) sub-domain ids, merged over basins by level index
    std::vector<std::vector<int64_t>> levels;
    nw.sub_ptr.push_back(0);
    int64_t total_subbas = 0;
    BasinWork w;
    std::vector<int64_t> sg_down, sg_order, node_of, sdown, sorder, sin_ptr, sin_idx, depth;
    for (int64_t b = 1; b <= n_pits; ++b) {
      const int64_t* bas = bnodes.data() + bptr[b];
      const int64_t k = bptr[b + 1] - bptr[b];
      for (int64_t l = 0; l < k; ++l) local_of[bas[l]] = l + 1;
      w.down.assign(k, 0);
      w.so.resize(k);
      for (int64_t l = 0; l < k; ++l) {
        const int64_t d = nw.down[bas[l]];
        w.down[l] = d ? local_of[d - 1] : 0;
        w.so[l] = nw.streamorder[bas[l]];
      }
      toposort_dfs(w.down, w.order);
      // subbasins (subdomains.jl:55-82)
      w.subbas.assign(k, 0);
      int64_t n_lab = 0;
      for (int64_t v1 : w.order) {
        if (w.so[v1 - 1] < min_sto) continue;
        const int64_t d = w.down[v1 - 1];
        if (d) { if (w.so[v1 - 1] != w.so[d - 1]) w.subbas[v1 - 1] = ++n_lab; }
        else w.subbas[v1 - 1] = ++n_lab;
      }
      const int64_t n_subbas = std::max<int64_t>(n_lab, 1);
      std::vector<std::vector<int64_t>> v_subbas;
      if (n_subbas > 1) {
        // fillnodata_upstream (subdomains.jl:8-24)
        w.fill = w.subbas;
        for (int64_t q = k - 1; q >= 0; --q) {
          const int64_t v = w.order[q] - 1, d = w.down[v];
          if (d && w.fill[v] == 0 && w.fill[d - 1] != 0) w.fill[v] = w.fill[d - 1];
        }
        // graph_from_nodes (subdomains.jl:127-143)
        node_of.assign(n_subbas + 1, 0);
        for (int64_t l = 0; l < k; ++l)
          if (w.subbas[l]) node_of[w.subbas[l]] = l + 1;
        sdown.assign(n_subbas, 0);
        for (int64_t s = 1; s <= n_subbas; ++s) {
          const int64_t d = w.down[node_of[s] - 1];
          if (d) sdown[s - 1] = w.fill[d - 1];
        }
        toposort_dfs(sdown, sorder);
        const int64_t outlet = sorder.back();
        build_in_csr(sdown, sin_ptr, sin_idx);
        // distances(Graph(graph_subbas), outlet): hop count in the tree rooted at the outlet
        depth.assign(n_subbas, -1);
        depth[outlet - 1] = 0;
        int64_t max_dist = 0;
        {
          std::vector<int64_t> frontier{outlet}, nxt;
          while (!frontier.empty()) {
            nxt.clear();
            for (int64_t u : frontier) {
              for (int64_t e = sin_ptr[u - 1]; e < sin_ptr[u]; ++e) {
                const int64_t x = sin_idx[e];
                if (depth[x - 1] < 0) { depth[x - 1] = depth[u - 1] + 1; nxt.push_back(x); }
              }
              const int64_t d = sdown[u - 1];
              if (d && depth[d - 1] < 0) { depth[d - 1] = depth[u - 1] + 1; nxt.push_back(d); }
            }
            frontier.swap(nxt);
          }
          for (int64_t x : depth) max_dist = std::max(max_dist, x);
          max_dist = std::max<int64_t>(max_dist, 1);
        }
        // subbasins_order (subdomains.jl:93-120)
        std::vector<std::vector<int64_t>> ord(max_dist + 1);
        ord[0] = {outlet};
        for (int64_t i = 0; i < max_dist; ++i)
          for (int64_t s : ord[i])
            for (int64_t e = sin_ptr[s - 1]; e < sin_ptr[s]; ++e) ord[i + 1].push_back(sin_idx[e]);
        // headwater sub-basins move to the last group. Julia iterates `order[i]` while
        // `filter!` shrinks it: after a removal the element that slid into the current slot
        // is skipped.
        for (int64_t i = 0; i < max_dist; ++i) {
          auto& lst = ord[i];
          size_t pos = 0;
          while (pos < lst.size()) {
            const int64_t s = lst[pos++];
            if (sin_ptr[s] == sin_ptr[s - 1]) {
              ord[max_dist].push_back(s);
              lst.erase(std::remove(lst.begin(), lst.end(), s), lst.end());
            }
          }
        }
        v_subbas.assign(ord.rbegin(), ord.rend());
      } else {
        v_subbas = {{1}};
      }
      for (size_t m = 0; m < v_subbas.size(); ++m) {
        if (levels.size() <= m) levels.emplace_back();
        for (int64_t s : v_subbas[m]) levels[m].push_back(s + total_subbas);
      }
      total_subbas += n_subbas;
      // per sub-basin traversal order
      if (n_subbas > 1) {
        // group local nodes per sub-basin id, ascending
        std::vector<int64_t> sptr(n_subbas + 2, 0), snodes(k);
        for (int64_t l = 0; l < k; ++l) sptr[w.fill[l] + 1]++;
        for (int64_t s = 0; s <= n_subbas; ++s) sptr[s + 1] += sptr[s];
        std::vector<int64_t> sfill(sptr.begin(), sptr.end() - 1);
        for (int64_t l = 0; l < k; ++l) snodes[sfill[w.fill[l]]++] = l;
        std::vector<int64_t> sub_local(k, 0);
        for (int64_t s = 1; s <= n_subbas; ++s) {
          const int64_t* sn = snodes.data() + sptr[s];
          const int64_t ks = sptr[s + 1] - sptr[s];
          for (int64_t l = 0; l < ks; ++l) sub_local[sn[l]] = l + 1;
          sg_down.assign(ks, 0);
          for (int64_t l = 0; l < ks; ++l) {
            const int64_t d = w.down[sn[l]];
            sg_down[l] = (d && w.fill[d - 1] == s) ? sub_local[d - 1] : 0;
          }
          toposort_dfs(sg_down, sg_order);
          for (int64_t q = 0; q < ks; ++q) {
            const int64_t g = bas[sn[sg_order[q] - 1]];
            nw.sub_order.push_back(g + 1);
            nw.sub_indices.push_back(index_toposort[g]);
          }
          nw.sub_ptr.push_back((int64_t)nw.sub_order.size());
        }
      } else {
        for (int64_t q = 0; q < k; ++q) {
          const int64_t g = bas[w.order[q] - 1];
          nw.sub_order.push_back(g + 1);
          nw.sub_indices.push_back(index_toposort[g]);
        }
        nw.sub_ptr.push_back((int64_t)nw.sub_order.size());
      }
    }
    nw.lvl_ptr.push_back(0);
    for (auto& lv : levels) {
      nw.lvl_idx.insert(nw.lvl_idx.end(), lv.begin(), lv.end());
      nw.lvl_ptr.push_back((int64_t)nw.lvl_idx.size());
    }
    if ((int64_t)nw.sub_order.size() != n) {
      err = "sub-domain partition does not cover the domain (node not draining to a pit)";
      return false;
    }
  }

  // ---- B200 wavefront: level = (max distance to outlet) - (distance to outlet) -----------
  std::vector<int64_t> dist(n, 0);
  int64_t dmax = 0;
  for (int64_t k = n - 1; k >= 0; --k) {
    const int64_t v = nw.order[k] - 1, d = nw.down[v];
    dist[v] = d ? dist[d - 1] + 1 : 0;
    dmax = std::max(dmax, dist[v]);
  }
  nw.n_wave_levels = n ? dmax + 1 : 0;
  nw.wave_level_ptr.assign(nw.n_wave_levels + 1, 0);
  nw.node_level.assign(n, 0);
  for (int64_t v = 0; v < n; ++v) {
    nw.node_level[v] = dmax - dist[v];
    nw.wave_level_ptr[dmax - dist[v] + 1]++;
  }
  for (int64_t l = 0; l < nw.n_wave_levels; ++l) nw.wave_level_ptr[l + 1] += nw.wave_level_ptr[l];
  return true;
}

void build_chunks(Network& nw, int64_t cap, int64_t piece_depth) {
  const int64_t n = nw.n;
  if (cap < 1) cap = 1;
  std::vector<int64_t> dist(n, 0);
  for (int64_t k = n - 1; k >= 0; --k) {
    const int64_t v = nw.order[k] - 1, d = nw.down[v];
    dist[v] = d ? dist[d - 1] + 1 : 0;
  }
  // Bottom-up (upstream first): acc[v] = size of the not-yet-cut upstream tree of v. A piece
  // may hold at most `cap` nodes (one node per lane of the warp that walks it), so when the
  // tree rooted at v would exceed the cap its largest uncut children are cut off (each becomes
  // the root of a piece of its own) until it fits. Pits and, with piece_depth > 0, the nodes at
  // a multiple of piece_depth from their outlet are roots as well.
  std::vector<int64_t> acc(n, 1);
  std::vector<uint8_t> cut(n, 0);
  std::vector<std::pair<int64_t, int64_t>> kids;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t v = nw.order[k] - 1;
    int64_t total = 1;
    for (int64_t e = nw.in_ptr[v]; e < nw.in_ptr[v + 1]; ++e)
      if (!cut[nw.in_idx[e] - 1]) total += acc[nw.in_idx[e] - 1];
    if (total > cap) {
      kids.clear();
      for (int64_t e = nw.in_ptr[v]; e < nw.in_ptr[v + 1]; ++e)
        if (!cut[nw.in_idx[e] - 1]) kids.emplace_back(acc[nw.in_idx[e] - 1], nw.in_idx[e] - 1);
      std::sort(kids.begin(), kids.end(), [](const auto& a, const auto& b) {
        return a.first != b.first ? a.first > b.first : a.second < b.second;
      });
      for (const auto& kd : kids) {
        if (total <= cap) break;
        cut[kd.second] = 1;
        total -= kd.first;
      }
    }
    acc[v] = total;
    if (nw.down[v] == 0 || (piece_depth > 0 && dist[v] % piece_depth == 0)) cut[v] = 1;
  }
  // piece roots in execution order: ascending root level, then node id. A producer's root is
  // exactly one level above the node it feeds, which lies at or above its consumer's root
  // level, so this is a topological order of the piece DAG and of the chunks packed from it
  // (dynamic scheduling in this order cannot deadlock).
  std::vector<int64_t> roots;
  for (int64_t v = 0; v < n; ++v)
    if (cut[v]) roots.push_back(v);
  std::sort(roots.begin(), roots.end(), [&](int64_t a, int64_t b) {
    return nw.node_level[a] != nw.node_level[b] ? nw.node_level[a] < nw.node_level[b] : a < b;
  });
  // pack: consecutive pieces with the same root level share a chunk while they fit
  nw.chunk_of_node.assign(n, -1);
  nw.chunk_outlet.assign(roots.size(), 0);
  nw.out_of_node.assign(n, -1);
  nw.n_outlets = 0;
  int64_t nc = 0, fill = 0, cur_level = -1;
  for (size_t i = 0; i < roots.size(); ++i) {
    const int64_t r = roots[i];
    const bool join = piece_depth > 0 && nc > 0 && nw.node_level[r] == cur_level &&
                      fill + acc[r] <= cap;
    if (!join) { ++nc; fill = 0; cur_level = nw.node_level[r]; }
    fill += acc[r];
    nw.chunk_of_node[r] = nc - 1;
    nw.chunk_outlet[i] = r + 1;
    if (nw.down[r] != 0) nw.out_of_node[r] = nw.n_outlets++;
  }
  nw.n_chunks = nc;
  for (int64_t k = n - 1; k >= 0; --k) {  // downstream -> upstream
    const int64_t v = nw.order[k] - 1;
    if (!cut[v]) nw.chunk_of_node[v] = nw.chunk_of_node[nw.down[v] - 1];
  }
  // slot order: (chunk, level, node id)
  std::vector<int64_t> nodes(n);
  for (int64_t v = 0; v < n; ++v) nodes[v] = v;
  std::sort(nodes.begin(), nodes.end(), [&](int64_t a, int64_t b) {
    if (nw.chunk_of_node[a] != nw.chunk_of_node[b]) return nw.chunk_of_node[a] < nw.chunk_of_node[b];
    if (nw.node_level[a] != nw.node_level[b]) return nw.node_level[a] < nw.node_level[b];
    return a < b;
  });
  nw.perm.assign(n, 0);
  nw.slot_of.assign(n, 0);
  for (int64_t p = 0; p < n; ++p) { nw.perm[p] = nodes[p] + 1; nw.slot_of[nodes[p]] = p; }
  nw.chunk_ptr.assign(nw.n_chunks + 1, 0);
  nw.chunk_l0.assign(nw.n_chunks, 0);
  nw.chunk_l1.assign(nw.n_chunks, 0);
  nw.chunk_clp_off.assign(nw.n_chunks + 1, 0);
  nw.clp.clear();
  int64_t p = 0;
  for (int64_t c = 0; c < nw.n_chunks; ++c) {
    nw.chunk_ptr[c] = p;
    int64_t q = p;
    while (q < n && nw.chunk_of_node[nodes[q]] == c) ++q;
    const int64_t l0 = nw.node_level[nodes[p]];
    const int64_t l1 = nw.node_level[nodes[q - 1]];  // slots are level-ascending inside a chunk
    nw.chunk_l0[c] = l0;
    nw.chunk_l1[c] = l1;
    nw.chunk_clp_off[c] = (int64_t)nw.clp.size();
    for (int64_t l = l0; l <= l1; ++l) {
      nw.clp.push_back(p);
      while (p < q && nw.node_level[nodes[p]] == l) ++p;
    }
    nw.clp.push_back(p);
  }
  nw.chunk_ptr[nw.n_chunks] = p;
  nw.chunk_clp_off[nw.n_chunks] = (int64_t)nw.clp.size();
}

}  // namespace wfb
