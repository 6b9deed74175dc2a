// stitch_driver.cpp -- a C++ caller of the multi-rank entry points of include/zq_b200.h as the archiver's host code
// would use them for ONE file cut across the ranks: the ranks are threads, the all-gather a
// barrier (zq_dist_create_cb), every piece is fragmented by the checker's chunker (oracle/zq_oracle.c, linked here:
// TEST INFRASTRUCTURE) and zq_dist_stitch_fragments decides which fragments are the stream's.  Exit code 0 iff the
// kept fragments of all ranks together equal the fragments of the whole stream.
// usage: stitch_driver <world> <kind: 0 text-like, 1 zero run across every cut>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "zq_b200.h"

extern "C" long long zqo_fragment(const uint8_t* in, uint64_t n, int fragment, uint32_t blocksize, uint32_t* frag_len,
                                  uint32_t* frag_hits, uint64_t cap);

namespace {

struct Gather {   // all-gather among the threads of this process: two generations of a counting barrier
  int world; std::mutex m; std::condition_variable cv; int arrived = 0; long gen = 0;
  std::vector<std::vector<uint8_t>> slot;
  explicit Gather(int w) : world(w), slot(w) {}
  void wait() {
    std::unique_lock<std::mutex> l(m);
    const long g = gen;
    if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(l, [&] { return gen != g; });
  }
};
struct Rank { Gather* g; int rank; };

int allgather(void* user, const void* in, void* out, size_t bytes) {
  Rank* r = static_cast<Rank*>(user);
  r->g->slot[r->rank].assign(static_cast<const uint8_t*>(in), static_cast<const uint8_t*>(in) + bytes);
  r->g->wait();
  for (int k = 0; k < r->g->world; ++k) memcpy(static_cast<uint8_t*>(out) + (size_t)k * bytes, r->g->slot[k].data(), bytes);
  r->g->wait();
  return 0;
}

std::vector<uint32_t> fragment(const uint8_t* p, uint64_t n, int frag) {
  std::vector<uint32_t> len(n / 64 + 16), hits(n / 64 + 16);
  const long long k = zqo_fragment(p, n, frag, (1u << 26) - 4096, len.data(), hits.data(), len.size());
  len.resize(k < 0 ? 0 : (size_t)k);
  return len;
}

}  // namespace

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 3, kind = argc > 2 ? atoi(argv[2]) : 0, frag = 1;
  const uint64_t total = 700000, overlap = 3 * (8128u << frag);
  std::vector<uint8_t> data(total);
  uint32_t x = 12345;
  for (uint64_t i = 0; i < total; ++i) {          // word-like bytes: a small alphabet with repeats
    x = x * 1664525u + 1013904223u;
    data[i] = (uint8_t)("etaoin shrdlu\n"[(x >> 24) % 14]);
  }
  if (kind == 1)                                   // zero pages across every cut: the chains never meet by themselves
    for (int r = 1; r < world; ++r) {
      uint64_t lo, hi; zq_dist_shard_range(total, r, world, &lo, &hi);
      memset(&data[lo - 60000], 0, 120000);
    }
  const std::vector<uint32_t> want = fragment(data.data(), total, frag);
  Gather g(world);
  std::vector<std::vector<uint32_t>> kept(world);
  std::vector<zq_stitch> rec(world);
  std::vector<int> rounds(world, 0), rc(world, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back([&, r] {
    Rank me{&g, r};
    zq_dist* zd = zq_dist_create_cb(r, world, allgather, &me);
    uint64_t lo, hi; zq_dist_shard_range(total, r, world, &lo, &hi);
    const uint64_t avail = r + 1 < world ? (hi + overlap < total ? hi + overlap : total) : total;
    uint64_t from = lo;
    zq_stitch st;
    std::vector<uint32_t> len;
    for (;;) {
      len = fragment(&data[from], avail - from, frag);
      do { rc[r] = zq_dist_stitch_fragments(zd, total, lo, hi, from, avail, len.data(), len.size(), &st); ++rounds[r]; }
      while (rc[r] == ZQ_OK && st.again && !st.restart);
      if (rc[r] != ZQ_OK || !st.again) break;
      from = st.restart_at;
    }
    if (rc[r] == ZQ_OK) kept[r].assign(len.begin() + st.first_keep, len.begin() + st.first_keep + st.n_keep);
    else fprintf(stderr, "rank %d: %s\n", r, zq_dist_last_error(zd));
    rec[r] = st;
    zq_dist_destroy(zd);
  });
  for (auto& t : th) t.join();
  std::vector<uint32_t> got;
  uint64_t at = 0;
  bool ok = true;
  for (int r = 0; r < world; ++r) {
    ok = ok && rc[r] == ZQ_OK && rec[r].global_first == got.size() && rec[r].begin == at && rec[r].global_total == want.size();
    for (uint32_t l : kept[r]) { got.push_back(l); at += l; }
    ok = ok && rec[r].end == at;
  }
  ok = ok && got == want;
  int maxr = 0;
  for (int r = 0; r < world; ++r) maxr = rounds[r] > maxr ? rounds[r] : maxr;
  printf("%s world %d kind %d: %zu fragments, %zu expected, %d call(s) on the busiest rank\n", ok ? "OK" : "FAILED", world, kind,
         got.size(), want.size(), maxr);
  return ok ? 0 : 1;
}
