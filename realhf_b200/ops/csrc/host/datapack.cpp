// Host-side packing helpers (C++17, no CUDA).
//
//  * partition_balanced: split a sequence of token counts into k contiguous parts so that the
//    largest part is as small as possible (binary search on the cap + greedy feasibility), then
//    place every boundary as close as possible to the ideal prefix i*total/k inside the window
//    that keeps the optimal cap.  O(n log(sum)) instead of the reference's O(n^2 k) numba DP
//    (reference: realhf/base/datapack.py:12-112).
//  * reorder_to_balanced_batches: longest-first bin packing of sequences into batches of a fixed
//    number of sequences (reference: realhf/base/datapack.py:116-143).
//  * merge_intervals: coalesce touching [a,b) intervals (reference: csrc/interval_op/interval_op.cpp).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <stdexcept>
#include <vector>

namespace py = pybind11;
using i64 = int64_t;

namespace {

// Latest boundary positions when every part greedily takes as much as fits under `cap`.
// Returns false if infeasible.  hi[i] = end index (exclusive) of part i.
bool greedy_left(const std::vector<i64>& pre, int k, i64 min_size, i64 cap, std::vector<i64>* hi) {
  const i64 n = (i64)pre.size() - 1;
  i64 start = 0;
  for (int p = 0; p < k; ++p) {
    const i64 must_leave = (i64)(k - 1 - p) * min_size;
    i64 lo_end = start + min_size;
    if (lo_end > n - must_leave) return false;
    if (pre[lo_end] - pre[start] > cap) return false;
    // largest end with sum <= cap
    i64 end = std::upper_bound(pre.begin() + lo_end, pre.begin() + (n - must_leave) + 1, pre[start] + cap) -
              pre.begin() - 1;
    if (p == k - 1) {
      if (end < n) return false;
      end = n;
    }
    if (hi) (*hi)[p] = end;
    start = end;
  }
  return true;
}

}  // namespace

std::vector<i64> partition_balanced(const std::vector<i64>& nums, int k, i64 min_size) {
  const i64 n = (i64)nums.size();
  if (k <= 0) throw std::invalid_argument("k must be positive");
  if (n < (i64)k * min_size) throw std::invalid_argument("array shorter than k * min_size");
  std::vector<i64> pre(n + 1, 0);
  for (i64 i = 0; i < n; ++i) {
    if (nums[i] < 0) throw std::invalid_argument("negative length");
    pre[i + 1] = pre[i] + nums[i];
  }
  i64 lo = 0, hi_cap = pre[n];
  while (lo < hi_cap) {
    i64 mid = lo + (hi_cap - lo) / 2;
    if (greedy_left(pre, k, min_size, mid, nullptr)) hi_cap = mid; else lo = mid + 1;
  }
  const i64 cap = lo;
  std::vector<i64> latest(k);
  greedy_left(pre, k, min_size, cap, &latest);
  // earliest feasible ends: greedy from the right
  std::vector<i64> earliest(k);
  {
    i64 end = n;
    for (int p = k - 1; p >= 0; --p) {
      earliest[p] = end;
      const i64 must_leave = (i64)p * min_size;
      i64 hi_start = end - min_size;
      // smallest start with pre[end]-pre[start] <= cap
      i64 start = std::lower_bound(pre.begin() + must_leave, pre.begin() + hi_start + 1, pre[end] - cap) - pre.begin();
      if (start > hi_start) start = hi_start;
      if (p == 0) start = 0;
      end = start;
    }
  }
  std::vector<i64> bounds(k + 1, 0);
  bounds[k] = n;
  i64 prev = 0;
  for (int p = 0; p < k - 1; ++p) {
    i64 wlo = std::max(earliest[p], prev + min_size);
    i64 whi = std::min(latest[p], n - (i64)(k - 1 - p) * min_size);
    // the chosen end must also keep this part under the cap
    whi = std::min<i64>(whi, std::upper_bound(pre.begin(), pre.end(), pre[prev] + cap) - pre.begin() - 1);
    if (whi < wlo) whi = wlo;
    const double ideal = (double)pre[n] * (p + 1) / k;
    i64 pos = std::lower_bound(pre.begin() + wlo, pre.begin() + whi + 1, (i64)ideal) - pre.begin();
    if (pos > whi) pos = whi;
    if (pos > wlo && std::abs((double)pre[pos - 1] - ideal) <= std::abs((double)pre[pos] - ideal)) --pos;
    bounds[p + 1] = pos;
    prev = pos;
  }
  return bounds;
}

std::pair<std::vector<i64>, i64> reorder_to_balanced_batches(const std::vector<i64>& seqlens, i64 n_seqs_per_batch) {
  const i64 n = (i64)seqlens.size();
  const i64 n_bins = (n + n_seqs_per_batch - 1) / n_seqs_per_batch;
  std::vector<i64> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](i64 a, i64 b) { return seqlens[a] > seqlens[b]; });
  // min-heap on (tokens, bin id) over bins that still have room
  using E = std::pair<i64, i64>;
  std::priority_queue<E, std::vector<E>, std::greater<E>> heap;
  for (i64 b = 0; b < n_bins; ++b) heap.push({0, b});
  std::vector<std::vector<i64>> bins(n_bins);
  std::vector<i64> tokens(n_bins, 0);
  for (i64 idx : order) {
    auto [t, b] = heap.top();
    heap.pop();
    bins[b].push_back(idx);
    tokens[b] = t + seqlens[idx];
    if ((i64)bins[b].size() < n_seqs_per_batch) heap.push({tokens[b], b});
  }
  std::vector<i64> bin_order(n_bins);
  std::iota(bin_order.begin(), bin_order.end(), 0);
  std::stable_sort(bin_order.begin(), bin_order.end(), [&](i64 a, i64 b) { return tokens[a] > tokens[b]; });
  std::vector<i64> out;
  out.reserve(n);
  for (i64 b : bin_order) out.insert(out.end(), bins[b].begin(), bins[b].end());
  i64 max_diff = n_bins ? (*std::max_element(tokens.begin(), tokens.end()) - *std::min_element(tokens.begin(), tokens.end())) : 0;
  return {out, max_diff};
}

std::vector<std::pair<i64, i64>> merge_intervals(std::vector<std::pair<i64, i64>> iv) {
  std::vector<std::pair<i64, i64>> out;
  for (auto& p : iv) {
    if (!out.empty() && out.back().second == p.first) out.back().second = p.second;
    else out.push_back(p);
  }
  return out;
}

void bind_search(py::module_& m);  // search.cpp

PYBIND11_MODULE(host_ext, m) {
  m.doc() = "realhf_b200 host-side native helpers";
  m.def("partition_balanced", &partition_balanced, py::arg("nums"), py::arg("k"), py::arg("min_size") = 1);
  m.def("reorder_to_balanced_batches", &reorder_to_balanced_batches);
  m.def("merge_intervals", &merge_intervals);
  bind_search(m);
}
