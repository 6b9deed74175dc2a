// frag_emu.cpp -- TEST INFRASTRUCTURE ONLY: the dedup fragmenter's device code (zq_fragment.cuh) on the host through
// tests/emu/simt_emu.h.  The round loop below mirrors zq_fragment_ex (zq_api.cu) step by step: segments, round 0,
// constant-run table, stitching rounds until no exit changes, gather.
#include <cuda_runtime.h>   // the shim

#include <algorithm>

#include "zq_fragment.cuh"

using namespace zqdev;

extern "C" long emu_fragment(const uint8_t* data, uint64_t n, int fragment, uint32_t blocksize, uint64_t seg,
                             uint32_t* frag_len, uint32_t* frag_hits, uint64_t frag_cap, uint32_t* rounds_out) {
  uint64_t maxf64 = fragment <= 19 ? ((uint64_t)8128 << fragment) : (uint64_t)blocksize - 12;
  if (maxf64 > (uint64_t)blocksize - 12) maxf64 = blocksize - 12;
  uint64_t minf64 = fragment <= 25 ? ((uint64_t)64 << fragment) : maxf64;
  if (minf64 > maxf64) minf64 = maxf64;
  const u32 maxf = (u32)maxf64, minf = (u32)minf64;
  const u32 thresh = fragment <= 22 ? (1u << (22 - fragment)) : 0u;
  if (seg < minf) seg = minf;
  std::vector<ZqSeg> segs;
  for (uint64_t p = 0; p < n; p += seg) {
    ZqSeg sg; sg.begin = p; sg.end = std::min<uint64_t>(n, p + seg); sg.file_end = n; sg.first = p == 0; sg.pad = 0;
    segs.push_back(sg);
  }
  const int nseg = (int)segs.size();
  if (nseg == 0) return 0;
  const u32 cap = (u32)((seg + maxf) / std::max<u32>(minf, 1) + 3);
  const size_t per = (size_t)nseg * cap;
  std::vector<u64> ex[2] = {std::vector<u64>(nseg), std::vector<u64>(nseg)}, ent(nseg), bnd[2] = {std::vector<u64>(per), std::vector<u64>(per)};
  std::vector<u32> hit[2] = {std::vector<u32>(per), std::vector<u32>(per)}, cnt[2] = {std::vector<u32>(nseg), std::vector<u32>(nseg)};
  std::vector<u32> cv(nseg, FRAG_NOT_CONST), Lv(256), Hv(256);
  std::vector<u64> ce(nseg, 0);
  std::vector<u32> rf(nseg, 0);
  u32 flags[2] = {0, 0};
  int cur = 0, round = 0;
  for (; round <= nseg; ++round) {
    flags[0] = 0;
    const int P = cur, N = cur ^ 1;
    emu::launch((nseg + FRAG_THREADS - 1) / FRAG_THREADS, FRAG_THREADS, 0, [&] {
      k_fragment_round(data, segs.data(), nseg, round, minf, maxf, thresh, cap, ex[P].data(), ex[N].data(), ent.data(), bnd[P].data(),
                       hit[P].data(), cnt[P].data(), bnd[N].data(), hit[N].data(), cnt[N].data(), &flags[0], &flags[1], cv.data(),
                       ce.data(), rf.data(), Lv.data(), Hv.data());
    });
    cur ^= 1;
    if (round == 0) {
      bool any = false;
      for (int k = nseg - 1; k >= 0; --k)
        if (cv[k] < 256) { ce[k] = (k + 1 < nseg && !segs[k + 1].first && cv[k + 1] == cv[k]) ? ce[k + 1] : segs[k].end; any = true; }
      for (int k = 0; k < nseg; ++k) rf[k] = (cv[k] < 256 && k > 0 && !segs[k].first && cv[k - 1] == cv[k]) ? rf[k - 1] : (u32)k;
      if (any) emu::launch(1, 256, 0, [&] { k_fragment_const_table(minf, maxf, thresh, Lv.data(), Hv.data()); });
    }
    if (flags[1]) return -1;
    if (!flags[0]) break;
  }
  if (rounds_out) *rounds_out = (u32)round + 1;
  uint64_t total = 0;
  for (int k = 0; k < nseg; ++k) {
    u64 prev = ent[k];
    for (u32 q = 0; q < cnt[cur][k]; ++q) {
      if (total >= frag_cap) return -2;
      const u64 b = bnd[cur][(size_t)k * cap + q];
      frag_len[total] = (u32)(b - prev); frag_hits[total] = hit[cur][(size_t)k * cap + q]; prev = b; ++total;
    }
  }
  return (long)total;
}
