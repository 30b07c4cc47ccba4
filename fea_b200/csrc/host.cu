// Host-side partition helper of the multi-GPU path (SURVEY.md §8(b) group 4, §8(e)): one parallel
// pass over the GLOBAL connectivity (host memory) that tells a rank which elements touch its node
// range and which nodes those elements reference.  fea_b200/dist.py:plan_slab calls it once per
// solve; with numpy the same scan costs ~0.2 s at 2.56 M hex8 elements (three boolean passes over
// 164 MB per range), a sixth of an 8-GPU solve -- here it is one multi-threaded read of the array.
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "fea_b200.h"

namespace {

struct RangeStat {
  int64_t count, first_id, last_id, min_node, max_node;
};

template <typename Index>
void scan_chunk(const Index* el, int64_t e0, int64_t e1, int npe, const int64_t* ranges, int n_ranges, RangeStat* out) {
  for (int r = 0; r < n_ranges; ++r) out[r] = RangeStat{0, -1, -1, INT64_MAX, INT64_MIN};
  for (int64_t e = e0; e < e1; ++e) {
    const Index* row = el + e * npe;
    int64_t mn = row[0], mx = row[0];
    for (int a = 1; a < npe; ++a) {
      mn = std::min<int64_t>(mn, row[a]);
      mx = std::max<int64_t>(mx, row[a]);
    }
    for (int r = 0; r < n_ranges; ++r) {
      const int64_t lo = ranges[2 * r], hi = ranges[2 * r + 1];
      if (mx < lo || mn >= hi) continue;  // cannot have a node inside [lo, hi)
      bool touch = false;
      for (int a = 0; a < npe && !touch; ++a) touch = row[a] >= lo && row[a] < hi;
      if (!touch) continue;
      RangeStat& s = out[r];
      if (s.count == 0) s.first_id = e;
      s.last_id = e;
      ++s.count;
      s.min_node = std::min(s.min_node, mn);
      s.max_node = std::max(s.max_node, mx);
    }
  }
}

template <typename Index>
int scan(const Index* el, int64_t n_elem, int npe, const int64_t* ranges, int n_ranges, int64_t* out) {
  unsigned hw = std::thread::hardware_concurrency();
  int n_threads = (int)std::min<int64_t>(hw == 0 ? 4 : std::min(hw, 32u), std::max<int64_t>(1, n_elem / 65536));
  std::vector<RangeStat> stats((size_t)n_threads * n_ranges);
  std::vector<std::thread> pool;
  const int64_t per = (n_elem + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const int64_t e0 = std::min<int64_t>(n_elem, t * per), e1 = std::min<int64_t>(n_elem, e0 + per);
    RangeStat* dst = stats.data() + (size_t)t * n_ranges;
    if (t == n_threads - 1) {
      scan_chunk(el, e0, e1, npe, ranges, n_ranges, dst);
    } else {
      pool.emplace_back([=] { scan_chunk(el, e0, e1, npe, ranges, n_ranges, dst); });
    }
  }
  for (auto& th : pool) th.join();
  for (int r = 0; r < n_ranges; ++r) {
    RangeStat tot{0, -1, -1, INT64_MAX, INT64_MIN};
    for (int t = 0; t < n_threads; ++t) {  // chunks are in ascending element order
      const RangeStat& s = stats[(size_t)t * n_ranges + r];
      if (s.count == 0) continue;
      if (tot.count == 0) tot.first_id = s.first_id;
      tot.last_id = s.last_id;
      tot.count += s.count;
      tot.min_node = std::min(tot.min_node, s.min_node);
      tot.max_node = std::max(tot.max_node, s.max_node);
    }
    int64_t* o = out + 5 * r;
    o[0] = tot.count, o[1] = tot.first_id, o[2] = tot.last_id, o[3] = tot.min_node, o[4] = tot.max_node;
  }
  return FEA_OK;
}

}  // namespace

extern "C" int fea_slab_scan(const void* elements_host, int32_t index_bytes, int64_t n_elem, int32_t nodes_per_elem,
                             const int64_t* ranges_host, int32_t n_ranges, int64_t* out_host) {
  if (!elements_host || !ranges_host || !out_host || n_elem < 0 || nodes_per_elem < 1 || n_ranges < 1)
    return FEA_ERR_INVALID;
  if (index_bytes == 8)
    return scan(static_cast<const int64_t*>(elements_host), n_elem, nodes_per_elem, ranges_host, n_ranges, out_host);
  if (index_bytes == 4)
    return scan(static_cast<const int32_t*>(elements_host), n_elem, nodes_per_elem, ranges_host, n_ranges, out_host);
  return FEA_ERR_INVALID;
}
