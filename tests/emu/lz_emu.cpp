// lz_emu.cpp -- TEST INFRASTRUCTURE ONLY: the -m2 device pipeline (k_suffix_sort -> k_lz77_sa) compiled for the
// host through tests/emu/simt_emu.h; tests/test_lz_emu.py compares the suffix array and the LZ77 stream with
// the oracle.  Build: g++ -O1 -std=c++17 -Itests/emu/shim -Izpaqfranz_b200/csrc -shared -fPIC tests/emu/lz_emu.cpp
#include <cuda_runtime.h>   // the shim

#include "zq_lz77.cuh"

using namespace zqdev;

// args = makeConfig's args[0..8] of the method.  sa_out: n u32 (may be null).  Returns stream length or <0.
extern "C" long emu_lz_sa(const uint8_t* data, uint32_t n, const int* args, uint32_t* sa_out, uint8_t* out, uint32_t cap) {
  const bool idx16 = n <= 65536;
  const u32 w = idx16 ? 2 : 4;
  std::vector<u8> in(data, data + n); in.resize(n + 64);
  std::vector<u8> work(zq_work_bytes(n, w) + 256);
  ZqUnit u; memset(&u, 0, sizeof u);
  u.n = n; u.idx16 = idx16; u.lz_cap = cap;
  ZqPlan pl; memset(&pl, 0, sizeof pl);
  for (int k = 0; k < 9; ++k) pl.args[k] = args[k];
  pl.lz_level = args[1] & 3; pl.use_sa = 1;
  int todo = 0;
  const size_t scr = (((size_t)n + 1) + 63) & ~(size_t)63;
  std::vector<u64> kbuf(2 * scr); std::vector<u32> vbuf(6 * scr);
  emu::launch(1, 256, sizeof(SortSmem<256>), [&] {
    k_suffix_sort<256, 4>(in.data(), &u, &todo, 1, work.data(), kbuf.data(), vbuf.data(), scr);
  });
  if (sa_out)
    for (u32 i = 0; i < n; ++i) sa_out[i] = idx16 ? ((const u16*)work.data())[i] : ((const u32*)work.data())[i];
  u32 lzlen = 0, err = 0, next = 0;
  emu::launch(1, 32, 0, [&] {
    if (idx16) k_lz77_sa<u16, true, 6>(in.data(), &u, &pl, &todo, 1, work.data(), out, &lzlen, &err, &next);
    else k_lz77_sa<u32, true, 6>(in.data(), &u, &pl, &todo, 1, work.data(), out, &lzlen, &err, &next);
  });
  if (err) return -(long)err;
  return (long)lzlen;
}
