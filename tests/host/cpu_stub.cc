// cpu_stub.cc -- TEST INFRASTRUCTURE: stands in for libxfeat_b200.so when the host classes (XFBmatcher, XFBvocabulary) are
// checked on a machine without a GPU (`pytest -m "not gpu"`).  Every distance / tree walk the host code asks for is answered
// by the C oracle (oracle/matcher_oracle.c), so what these tests exercise is the HOST logic: visiting order, pair-list
// construction, accept / reject replay, map bookkeeping.  The product never links this file.
#include <cstring>
#include <vector>

#include "xfeat_b200.h"

extern "C" {
int mo_descriptor_distance(const float* a, const float* b);
void mo_distance_matrix(const float* A, int n1, const float* B, int n2, int32_t* out);
void mo_bruteforce(const float* A, int n1, const float* B, int n2, const int32_t* groupA, const int32_t* groupB, int init, int32_t* best_idx,
                   int32_t* best_dist, int32_t* second_dist, int32_t* best_idx_rev, int32_t* best_dist_rev);
void mo_bow_transform(const float* desc, int n, const uint8_t* node_desc, const int32_t* child_start, const int32_t* child_index, int L, int levelsup,
                      int32_t* leaf, int32_t* nid);
}

static std::vector<uint8_t> g_desc;
static std::vector<int32_t> g_start, g_child;
static int g_L = 0;

extern "C" {
const char* xfb_last_error(const xfb_ctx*) { return "cpu stub"; }
int xfb_distance_matrix(xfb_ctx*, const float* A, int n1, const float* B, int n2, int32_t* out) { mo_distance_matrix(A, n1, B, n2, out); return XFB_OK; }
int xfb_distance_pairs(xfb_ctx*, const float* A, int n1, const float* B, int n2, const int32_t* ia, const int32_t* ib, int n, int32_t* out) {
  for (int p = 0; p < n; ++p) {
    if (ia[p] < 0 || ia[p] >= n1 || ib[p] < 0 || ib[p] >= n2) return XFB_ERR_ARG;
    out[p] = mo_descriptor_distance(A + (size_t)ia[p] * 64, B + (size_t)ib[p] * 64);
  }
  return XFB_OK;
}
int xfb_match(xfb_ctx*, const float* A, int n1, const float* B, int n2, const int32_t* ga, const int32_t* gb, int init, int32_t* bi, int32_t* bd,
              int32_t* sd, int32_t* ri, int32_t* rd) {
  std::vector<int32_t> t0(n1), t1(n1), t2(n1), t3(n2), t4(n2);
  mo_bruteforce(A, n1, B, n2, ga, gb, init, t0.data(), t1.data(), t2.data(), t3.data(), t4.data());
  if (bi) std::memcpy(bi, t0.data(), sizeof(int32_t) * n1);
  if (bd) std::memcpy(bd, t1.data(), sizeof(int32_t) * n1);
  if (sd) std::memcpy(sd, t2.data(), sizeof(int32_t) * n1);
  if (ri) std::memcpy(ri, t3.data(), sizeof(int32_t) * n2);
  if (rd) std::memcpy(rd, t4.data(), sizeof(int32_t) * n2);
  return XFB_OK;
}
int xfb_vocab_load(xfb_ctx*, const uint8_t* nd, const int32_t* cs, const int32_t* ci, int n, int m, int L) {
  g_desc.assign(nd, nd + (size_t)n * 32); g_start.assign(cs, cs + n + 1); g_child.assign(ci, ci + m); g_L = L;
  return XFB_OK;
}
int xfb_bow_transform(xfb_ctx*, const float* d, int n, int levelsup, int32_t* leaf, int32_t* nid) {
  mo_bow_transform(d, n, g_desc.data(), g_start.data(), g_child.data(), g_L, levelsup, leaf, nid);
  return XFB_OK;
}
}
