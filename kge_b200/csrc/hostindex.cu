// hostindex.cu — host-side (CPU) label plumbing next to the scoring path: the key -> all-values index that
// KvsAll training and filtered entity ranking build their label / filter coordinates from.
//
// Replaces the reference's KvsAllIndex (indexing.py:10-194: numpy argsort + np.unique + a numba dict) and the
// Python collate loops that walk it (train_KvsAll.py:116-203, util.py:6-30) with plain C++: a sort, a unique
// pass and binary searches; results are returned in CSR form (row offsets + column ids), which is what the
// device epilogues consume instead of `coord_to_sparse_tensor(...).to_dense()` (util.py:32-60).
// No device code here; the functions run without a GPU.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>
#include "common.cuh"

namespace b200kge {
namespace {

inline bool col_ok(int c) { return c >= 0 && c <= 2; }

// index of (k0, k1) in the sorted unique key list, or -1
inline int64_t find_key(const int64_t* keys, int64_t num_keys, int64_t k0, int64_t k1) {
  int64_t lo = 0, hi = num_keys;
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    const int64_t a = keys[2 * mid], b = keys[2 * mid + 1];
    if (a < k0 || (a == k0 && b < k1)) lo = mid + 1;
    else hi = mid;
  }
  if (lo < num_keys && keys[2 * lo] == k0 && keys[2 * lo + 1] == k1) return lo;
  return -1;
}

}  // namespace
}  // namespace b200kge

using namespace b200kge;

extern "C" {

int b200kge_kvsall_index_build(const int64_t* triples, int64_t n, int key_col0, int key_col1, int value_col,
                               int64_t* keys_out, int64_t* offsets_out, int64_t* values_out, int64_t* num_keys) {
  if ((!triples && n > 0) || !offsets_out || !num_keys || (n > 0 && (!keys_out || !values_out))) {
    set_error("null operand");
    return B200KGE_ERR_INVALID;
  }
  if (n < 0 || !col_ok(key_col0) || !col_ok(key_col1) || !col_ok(value_col) || key_col0 == key_col1 ||
      value_col == key_col0 || value_col == key_col1) {
    set_error("key/value columns must be a permutation of (0,1,2)");
    return B200KGE_ERR_INVALID;
  }
  // sort by (key0, key1, value): indexing.py:178-194 (sort by value, then stable by key1, then stable by key0)
  std::vector<int64_t> order((size_t)n);
  std::iota(order.begin(), order.end(), (int64_t)0);
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    const int64_t* x = triples + 3 * a;
    const int64_t* y = triples + 3 * b;
    if (x[key_col0] != y[key_col0]) return x[key_col0] < y[key_col0];
    if (x[key_col1] != y[key_col1]) return x[key_col1] < y[key_col1];
    return x[value_col] < y[value_col];
  });
  // unique keys + start offset of each (np.unique(..., axis=0, return_index=True), indexing.py:39-42)
  int64_t nk = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t* t = triples + 3 * order[(size_t)i];
    if (i == 0 || t[key_col0] != keys_out[2 * (nk - 1)] || t[key_col1] != keys_out[2 * (nk - 1) + 1]) {
      keys_out[2 * nk] = t[key_col0];
      keys_out[2 * nk + 1] = t[key_col1];
      offsets_out[nk] = i;
      ++nk;
    }
    values_out[i] = t[value_col];          // duplicates are kept, as in the reference
  }
  offsets_out[nk] = n;
  *num_keys = nk;
  return 0;
}

int b200kge_kvsall_lookup(const int64_t* keys, const int64_t* offsets, const int64_t* values, int64_t num_keys,
                          const int64_t* query_keys, int64_t nq, int64_t col_shift, int64_t* offsets_out,
                          int64_t* cols_out) {
  if (!offsets_out || (nq > 0 && !query_keys) || (num_keys > 0 && (!keys || !offsets || !values)) || nq < 0 || num_keys < 0) {
    set_error("null operand");
    return B200KGE_ERR_INVALID;
  }
  // KvsAllIndex.get_all (indexing.py:113-166): absent keys contribute nothing
  int64_t total = 0;
  for (int64_t i = 0; i < nq; ++i) {
    offsets_out[i] = total;
    const int64_t k = find_key(keys, num_keys, query_keys[2 * i], query_keys[2 * i + 1]);
    if (k < 0) continue;
    const int64_t b = offsets[k], e = offsets[k + 1];
    if (cols_out)
      for (int64_t j = b; j < e; ++j) cols_out[total + (j - b)] = values[j] + col_shift;
    total += e - b;
  }
  offsets_out[nq] = total;
  return 0;
}

int b200kge_kvsall_gather(const int64_t* keys, const int64_t* offsets, const int64_t* values, int64_t num_keys,
                          const int64_t* examples, int64_t nb, int64_t* queries_out, int64_t* offsets_out,
                          int64_t* cols_out) {
  if (!offsets_out || (nb > 0 && (!examples || !keys || !offsets || !values || !queries_out)) || nb < 0) {
    set_error("null operand");
    return B200KGE_ERR_INVALID;
  }
  // the collate function of KvsAll training for one query type (train_KvsAll.py:116-203): example = key index
  int64_t total = 0;
  for (int64_t i = 0; i < nb; ++i) {
    const int64_t k = examples[i];
    if (k < 0 || k >= num_keys) { set_error("example index %lld out of range [0, %lld)", (long long)k, (long long)num_keys); return B200KGE_ERR_INVALID; }
    offsets_out[i] = total;
    queries_out[2 * i] = keys[2 * k];
    queries_out[2 * i + 1] = keys[2 * k + 1];
    const int64_t b = offsets[k], e = offsets[k + 1];
    if (cols_out)
      for (int64_t j = b; j < e; ++j) cols_out[total + (j - b)] = values[j];
    total += e - b;
  }
  offsets_out[nb] = total;
  return 0;
}

}  // extern "C"
