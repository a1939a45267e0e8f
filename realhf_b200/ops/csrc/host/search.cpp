// Allocation search: MCMC over (device mesh, dp/tp/pp) choices per model function call, scored by a discrete-event
// simulation of the multi-iteration dataflow graph with device-mesh exclusivity, parameter-reallocation cost and a
// per-GPU memory model.
//
// Same job as the reference's `mdm_search` (csrc/search/{search,simulate,rpc,device_mesh}.cpp: MCMC :122-345,
// simulator simulate.cpp:22-205, param-sync pseudo tasks rpc.cpp:129-214, memory rpc.cpp:21-80), re-derived:
//   * GPUs are bits of a 128-bit set, mesh overlap is an AND, a mesh is busy until its last GPU frees up;
//   * parameter reallocation cost is analytic for NVSwitch (every destination GPU pulls its shard at link bandwidth,
//     all in parallel), replacing the reference's hand-measured lookup table keyed by strings (rpc.cpp:113-127);
//   * the memory cap is a parameter (180 GB B200) instead of the hard-coded 80 GiB (simulate.cpp:13);
//   * exposed through pybind11 with plain structs, so the Python driver (search/engine.py) stays small.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <queue>
#include <random>
#include <string>
#include <vector>

namespace py = pybind11;

namespace {

struct GpuSet {
  uint64_t w[2] = {0, 0};
  bool overlaps(const GpuSet& o) const { return (w[0] & o.w[0]) || (w[1] & o.w[1]); }
  bool equals(const GpuSet& o) const { return w[0] == o.w[0] && w[1] == o.w[1]; }
  bool subset_of(const GpuSet& o) const { return !(w[0] & ~o.w[0]) && !(w[1] & ~o.w[1]); }
  void set(int i) { w[i >> 6] |= (1ull << (i & 63)); }
  bool test(int i) const { return (w[i >> 6] >> (i & 63)) & 1; }
  int count() const { return __builtin_popcountll(w[0]) + __builtin_popcountll(w[1]); }
};

struct Candidate {
  int mesh = 0;              // index into meshes
  int dp = 1, tp = 1, pp = 1;
  double time_us = 0;        // execution time of the MFC under this layout
  double mem_static = 0;     // bytes per GPU that persist while the role lives on this layout (weights [+ optimizer])
  double mem_active = 0;     // bytes per GPU only during the call (activations, KV cache, gradients)
};

struct Rpc {
  std::string name;
  int role = 0;              // index of the model role
  int kind = 0;              // 0 generate, 1 inference, 2 train_step
  std::vector<Candidate> cands;
};

struct Problem {
  std::vector<Rpc> rpcs;
  std::vector<GpuSet> meshes;
  std::vector<std::pair<int, int>> edges;  // producer -> consumer (rpc indices)
  std::vector<double> role_bytes;          // bf16 weight bytes per role
  int n_gpus = 8;
  double mem_cap = 180e9;
  double link_bw = 770e9;    // bytes/s per GPU per direction (measured NVLink 5 peer copy)
  double realloc_latency_us = 30.0;
  int n_iters = 2;
  // optional: planned reallocation times, (role, src mesh, dp, tp, pp, dst mesh, dp, tp, pp) -> us; overrides the closed form
  std::map<std::array<int, 9>, double> realloc_table;
};

struct SimResult {
  double time_us = 0;        // per DFG iteration (steady state)
  double max_mem = 0;
  double cost = 0;           // time with the memory penalty applied
  std::vector<double> start, end;
};

// Time to give `dst` layout a copy of the role's weights from `src` layout.
double realloc_cost_us(const Problem& P, int role, const Candidate& src, const Candidate& dst) {
  if (src.mesh == dst.mesh && src.tp == dst.tp && src.pp == dst.pp && src.dp == dst.dp) return 0.0;
  if (!P.realloc_table.empty()) {
    auto it = P.realloc_table.find({role, src.mesh, src.dp, src.tp, src.pp, dst.mesh, dst.dp, dst.tp, dst.pp});
    if (it != P.realloc_table.end()) return it->second;
  }
  const double shard = P.role_bytes[role] / (double)(dst.tp * dst.pp);
  // a destination GPU that already holds the same TP/PP shard (same mesh, same tp/pp, only dp differs) copies locally
  const bool same_shards = src.mesh == dst.mesh && src.tp == dst.tp && src.pp == dst.pp;
  if (same_shards) return P.realloc_latency_us;
  return P.realloc_latency_us + shard / P.link_bw * 1e6;
}

SimResult simulate(const Problem& P, const std::vector<int>& choice) {
  const int R = (int)P.rpcs.size();
  const int n_roles = (int)P.role_bytes.size();
  std::vector<const Candidate*> c(R);
  for (int i = 0; i < R; ++i) c[i] = &P.rpcs[i].cands[choice[i]];
  // the train MFC of each role (if any) defines where its weights live permanently
  std::vector<int> train_of(n_roles, -1);
  for (int i = 0; i < R; ++i)
    if (P.rpcs[i].kind == 2) train_of[P.rpcs[i].role] = i;
  // per-MFC extra time: parameter reallocation in (pre-hook); the way back only drops the copy
  std::vector<double> extra(R, 0.0);
  std::vector<double> extra_mem(R, 0.0);
  for (int i = 0; i < R; ++i) {
    const int t = train_of[P.rpcs[i].role];
    if (t >= 0 && t != i) {
      extra[i] = realloc_cost_us(P, P.rpcs[i].role, *c[t], *c[i]);
      // a replica with unsharded layers (tp = pp = 1 on both sides) on GPUs that all belong to the training mesh holds exactly the
      // training shard: the runtime aliases the training buffer, the replica costs no memory (model_worker.py::_param_realloc)
      const bool aliased = c[t]->tp == 1 && c[t]->pp == 1 && c[i]->tp == 1 && c[i]->pp == 1 &&
                           P.meshes[c[i]->mesh].subset_of(P.meshes[c[t]->mesh]);
      if (extra[i] > 0 && !aliased) extra_mem[i] = P.role_bytes[P.rpcs[i].role] / (double)(c[i]->tp * c[i]->pp);
    }
  }
  // ---- list scheduling over n_iters iterations
  const int T = R * P.n_iters;
  std::vector<std::vector<int>> succ(T);
  std::vector<int> indeg(T, 0);
  auto id = [&](int it, int r) { return it * R + r; };
  for (int it = 0; it < P.n_iters; ++it) {
    for (auto& e : P.edges) {
      succ[id(it, e.first)].push_back(id(it, e.second));
      ++indeg[id(it, e.second)];
    }
    if (it + 1 < P.n_iters) {
      for (int r = 0; r < R; ++r) {
        // next iteration of the same MFC follows this one; users of a role wait for that role's train step
        succ[id(it, r)].push_back(id(it + 1, r));
        ++indeg[id(it + 1, r)];
        const int t = train_of[P.rpcs[r].role];
        if (t >= 0 && t != r) {
          succ[id(it, t)].push_back(id(it + 1, r));
          ++indeg[id(it + 1, r)];
        }
      }
    }
  }
  std::vector<double> ready(T, 0.0), start(T, 0.0), end(T, 0.0);
  std::vector<double> gpu_free(P.n_gpus, 0.0);
  using Item = std::pair<double, int>;
  std::priority_queue<Item, std::vector<Item>, std::greater<Item>> q;
  for (int t = 0; t < T; ++t)
    if (indeg[t] == 0) q.push({0.0, t});
  int done = 0;
  while (!q.empty()) {
    auto [rt, t] = q.top();
    q.pop();
    const int r = t % R;
    const GpuSet& m = P.meshes[c[r]->mesh];
    double st = rt;
    for (int g = 0; g < P.n_gpus; ++g)
      if (m.test(g)) st = std::max(st, gpu_free[g]);
    const double en = st + c[r]->time_us + extra[r];
    for (int g = 0; g < P.n_gpus; ++g)
      if (m.test(g)) gpu_free[g] = en;
    start[t] = st;
    end[t] = en;
    ++done;
    for (int s : succ[t]) {
      ready[s] = std::max(ready[s], en);
      if (--indeg[s] == 0) q.push({ready[s], s});
    }
  }
  SimResult res;
  res.start = start;
  res.end = end;
  double total_end = 0, first_end = 0;
  for (int r = 0; r < R; ++r) {
    total_end = std::max(total_end, end[id(P.n_iters - 1, r)]);
    first_end = std::max(first_end, end[id(0, r)]);
  }
  res.time_us = P.n_iters > 1 ? (total_end - first_end) / (double)(P.n_iters - 1) : total_end;
  if (done != T) res.time_us = std::numeric_limits<double>::infinity();  // cyclic input
  // ---- memory per GPU: persistent weights (+ optimizer) of every role at its home layout + the largest transient
  std::vector<double> stat(P.n_gpus, 0.0), act(P.n_gpus, 0.0);
  std::vector<char> role_counted(n_roles, 0);
  for (int r = 0; r < R; ++r) {
    const int role = P.rpcs[r].role;
    const int home = train_of[role] >= 0 ? train_of[role] : r;
    if (home == r && !role_counted[role]) {
      role_counted[role] = 1;
      const GpuSet& m = P.meshes[c[r]->mesh];
      for (int g = 0; g < P.n_gpus; ++g)
        if (m.test(g)) stat[g] += c[r]->mem_static;
    }
    const GpuSet& m = P.meshes[c[r]->mesh];
    for (int g = 0; g < P.n_gpus; ++g)
      if (m.test(g)) act[g] = std::max(act[g], c[r]->mem_active + extra_mem[r]);
  }
  for (int g = 0; g < P.n_gpus; ++g) res.max_mem = std::max(res.max_mem, stat[g] + act[g]);
  const double over = std::max(0.0, res.max_mem - P.mem_cap) / P.mem_cap;
  res.cost = res.time_us * (1.0 + 10.0 * over) + (over > 0 ? 1e7 : 0.0);
  return res;
}

struct SearchOut {
  std::vector<int> choice;
  double cost, time_us, max_mem;
};

std::vector<SearchOut> mcmc(const Problem& P, double beta, double time_limit_s, uint64_t seed, int64_t max_moves, int top_k) {
  const int R = (int)P.rpcs.size();
  std::mt19937_64 rng(seed);
  std::vector<int> cur(R, 0);
  SimResult cr = simulate(P, cur);
  std::vector<SearchOut> best;
  auto record = [&](const std::vector<int>& ch, const SimResult& r) {
    for (auto& b : best)
      if (b.choice == ch) return;
    best.push_back({ch, r.cost, r.time_us, r.max_mem});
    std::sort(best.begin(), best.end(), [](const SearchOut& a, const SearchOut& b) { return a.cost < b.cost; });
    if ((int)best.size() > top_k) best.pop_back();
  };
  record(cur, cr);
  const auto t0 = std::chrono::steady_clock::now();
  std::uniform_real_distribution<double> U(0.0, 1.0);
  const double scale = std::max(1.0, cr.cost);
  for (int64_t it = 0; it < max_moves; ++it) {
    if ((it & 255) == 0) {
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el > time_limit_s) break;
    }
    const int r = (int)(rng() % R);
    const int nc = (int)P.rpcs[r].cands.size();
    if (nc <= 1) continue;
    std::vector<int> nxt = cur;
    nxt[r] = (int)(rng() % nc);
    if (nxt[r] == cur[r]) continue;
    SimResult nr = simulate(P, nxt);
    const double d = (nr.cost - cr.cost) / scale;
    if (d <= 0 || U(rng) < std::exp(-beta * d)) {
      cur.swap(nxt);
      cr = nr;
      record(cur, cr);
    }
  }
  return best;
}

Problem problem_from_py(const py::dict& d) {
  Problem P;
  P.n_gpus = d["n_gpus"].cast<int>();
  if (P.n_gpus > 128) throw std::invalid_argument("at most 128 GPUs");
  if (d.contains("mem_cap")) P.mem_cap = d["mem_cap"].cast<double>();
  if (d.contains("link_bw")) P.link_bw = d["link_bw"].cast<double>();
  if (d.contains("n_iters")) P.n_iters = d["n_iters"].cast<int>();
  if (d.contains("realloc_latency_us")) P.realloc_latency_us = d["realloc_latency_us"].cast<double>();
  P.role_bytes = d["role_bytes"].cast<std::vector<double>>();
  if (d.contains("realloc_table")) {
    for (auto item : d["realloc_table"].cast<py::list>()) {
      py::tuple kv = item.cast<py::tuple>();
      auto k = kv[0].cast<std::vector<int>>();
      if (k.size() != 9) throw std::invalid_argument("realloc_table keys have 9 integers");
      std::array<int, 9> a;
      std::copy(k.begin(), k.end(), a.begin());
      P.realloc_table[a] = kv[1].cast<double>();
    }
  }
  for (auto m : d["meshes"].cast<std::vector<std::vector<int>>>()) {
    GpuSet s;
    for (int g : m) s.set(g);
    P.meshes.push_back(s);
  }
  P.edges = d["edges"].cast<std::vector<std::pair<int, int>>>();
  for (auto item : d["rpcs"].cast<py::list>()) {
    py::dict rd = item.cast<py::dict>();
    Rpc r;
    r.name = rd["name"].cast<std::string>();
    r.role = rd["role"].cast<int>();
    r.kind = rd["kind"].cast<int>();
    for (auto ci : rd["cands"].cast<py::list>()) {
      py::tuple t = ci.cast<py::tuple>();
      Candidate c;
      c.mesh = t[0].cast<int>(); c.dp = t[1].cast<int>(); c.tp = t[2].cast<int>(); c.pp = t[3].cast<int>();
      c.time_us = t[4].cast<double>(); c.mem_static = t[5].cast<double>(); c.mem_active = t[6].cast<double>();
      r.cands.push_back(c);
    }
    if (r.cands.empty()) throw std::invalid_argument("rpc without candidates: " + r.name);
    P.rpcs.push_back(r);
  }
  return P;
}

py::dict out_to_py(const SearchOut& o) {
  py::dict d;
  d["choice"] = o.choice;
  d["cost"] = o.cost;
  d["time_us"] = o.time_us;
  d["max_mem"] = o.max_mem;
  return d;
}

}  // namespace

void bind_search(py::module_& m) {
  m.def("simulate_allocation", [](py::dict prob, std::vector<int> choice) {
    Problem P = problem_from_py(prob);
    SimResult r = simulate(P, choice);
    py::dict d;
    d["time_us"] = r.time_us; d["max_mem"] = r.max_mem; d["cost"] = r.cost; d["start"] = r.start; d["end"] = r.end;
    return d;
  });
  m.def("mcmc_search", [](py::dict prob, double beta, double time_limit_s, uint64_t seed, int64_t max_moves, int top_k) {
    Problem P = problem_from_py(prob);
    std::vector<SearchOut> res;
    {
      py::gil_scoped_release rel;
      res = mcmc(P, beta, time_limit_s, seed, max_moves, top_k);
    }
    py::list out;
    for (auto& o : res) out.append(out_to_py(o));
    return out;
  }, py::arg("problem"), py::arg("beta") = 1.0, py::arg("time_limit_s") = 5.0, py::arg("seed") = 1, py::arg("max_moves") = 25000000,
     py::arg("top_k") = 10);
  m.def("multi_mcmc_search", [](py::dict prob, std::vector<double> betas, double time_limit_s, uint64_t seed, int top_k) {
    Problem P = problem_from_py(prob);
    std::vector<SearchOut> all;
    {
      py::gil_scoped_release rel;
      for (size_t i = 0; i < betas.size(); ++i) {
        auto res = mcmc(P, betas[i], time_limit_s / (double)betas.size(), seed + i, 25000000, top_k);
        all.insert(all.end(), res.begin(), res.end());
      }
    }
    std::sort(all.begin(), all.end(), [](const SearchOut& a, const SearchOut& b) { return a.cost < b.cost; });
    py::list out;
    for (size_t i = 0; i < all.size() && (int)i < top_k; ++i) out.append(out_to_py(all[i]));
    return out;
  }, py::arg("problem"), py::arg("betas"), py::arg("time_limit_s") = 10.0, py::arg("seed") = 1, py::arg("top_k") = 10);
  m.def("parameter_sync_cost", [](double role_bytes, std::vector<int> src, std::vector<int> dst, bool same_mesh, double link_bw) {
    Problem P;
    P.role_bytes = {role_bytes};
    P.link_bw = link_bw;
    Candidate s, t;
    s.dp = src[0]; s.tp = src[1]; s.pp = src[2]; s.mesh = 0;
    t.dp = dst[0]; t.tp = dst[1]; t.pp = dst[2]; t.mesh = same_mesh ? 0 : 1;
    return realloc_cost_us(P, 0, s, t);
  });
}
