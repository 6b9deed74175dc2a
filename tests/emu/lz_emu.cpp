// lz_emu.cpp -- TEST INFRASTRUCTURE ONLY: the -m2 device pipeline (k_suffix_sort -> k_lz77_sa) compiled for the
// host through tests/emu/simt_emu.h; tests/test_lz_emu.py compares the suffix array and the LZ77 stream with
// the oracle.  Build: g++ -O1 -std=c++17 -Itests/emu/shim -Izpaqfranz_b200/csrc -shared -fPIC tests/emu/lz_emu.cpp
#include <cuda_runtime.h>   // the shim

#include "zq_lz77.cuh"
#include "zq_lz77_scan.cuh"
#include "zq_sufsort16.cuh"

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
    if (idx16) k_lz77_sa<u16>(in.data(), &u, &pl, &todo, 1, work.data(), out, &lzlen, &err, &next);
    else k_lz77_sa<u32>(in.data(), &u, &pl, &todo, 1, work.data(), out, &lzlen, &err, &next);
  });
  if (err) return -(long)err;
  return (long)lzlen;
}


// The position-parallel form of the same parse (zq_lz77_scan.cuh): k_suffix_sort -> k_lz_scan<0> -> k_lz_scan<1> ->
// k_lz_walk -> k_lz_emit, with the grid / block shapes the host uses.  Returns stream length or <0.
template <typename IdxT>
static long run_scan_pipeline(std::vector<u8>& in, ZqUnit& u, ZqPlan& pl, std::vector<u8>& work, uint8_t* out, uint32_t* ntok_out) {
  int todo = 0;
  u32 tile_first[2] = {0, lzs_tiles(u.n)};
  u32 ctr = 0;
  emu::launch(2, LZS_NT, sizeof(LzsSmem<IdxT>) + (LZS_NT / 32) * sizeof(LzsQueue<IdxT>), [&] { k_lz_scan<IdxT, 0>(&u, &pl, &todo, tile_first, 1, 0u, work.data(), &ctr); });
  ctr = 0;
  emu::launch(2, LZS_NT, sizeof(LzsSmem<IdxT>) + (LZS_NT / 32) * sizeof(LzsQueue<IdxT>), [&] { k_lz_scan<IdxT, 1>(&u, &pl, &todo, tile_first, 1, lzs_tiles(u.n), work.data(), &ctr); });
  const size_t tokcap = u.n / std::max(pl.args[2], 1) + u.n / 4096 + 4 + 32;
  std::vector<LzToken> tok(tokcap);
  std::vector<u64> bitpos(tokcap);
  u64 tok_off = 0; u32 ntok = 0, next = 0, lzlen = 0, err = 0;
  emu::launch(1, 128, 0, [&] { k_lz_walk<IdxT, false>(in.data(), &u, &pl, &todo, 1, work.data(), &tok_off, tok.data(), &ntok, &next); });
  next = 0;
  emu::launch(1, 128, 0, [&] { k_lz_walk<IdxT, true>(in.data(), &u, &pl, &todo, 1, work.data(), &tok_off, tok.data(), &ntok, &next); });
  if (ntok_out) *ntok_out = ntok;
  std::vector<u8> lz(((size_t)u.lz_cap + 15) / 16 * 16 + 16, 0);   // malloc'ed: 16-byte aligned like the device arena
  emu::launch(1, LZE_NT, sizeof(LzeSmem), [&] { k_lz_emit(in.data(), &u, &pl, &todo, 1, &tok_off, tok.data(), &ntok, bitpos.data(), lz.data(), &lzlen, &err); });
  if (err) return -(long)err;
  memcpy(out, lz.data(), lzlen);
  return (long)lzlen;
}

extern "C" long emu_lz_scan(const uint8_t* data, uint32_t n, const int* args, uint8_t* out, uint32_t cap, uint32_t* ntok_out) {
  const bool idx16 = n <= 65536;
  const u32 w = idx16 ? 2 : 4;
  std::vector<u8> in(data, data + n); in.resize(n + 64);
  std::vector<u8> work(zq_work_bytes_scan(n, w) + 256);
  ZqUnit u; memset(&u, 0, sizeof u);
  u.n = n; u.idx16 = idx16; u.lz_cap = cap; u.want_pk = 1;
  ZqPlan pl; memset(&pl, 0, sizeof pl);
  for (int k = 0; k < 9; ++k) pl.args[k] = args[k];
  pl.lz_level = args[1] & 3; pl.use_sa = 1;
  int todo = 0;
  const size_t scr = (((size_t)n + 1) + 63) & ~(size_t)63;
  std::vector<u64> kbuf(2 * scr); std::vector<u32> vbuf(6 * scr);
  emu::launch(1, 256, sizeof(SortSmem<256>), [&] {
    k_suffix_sort<256, 4>(in.data(), &u, &todo, 1, work.data(), kbuf.data(), vbuf.data(), scr);
  });
  return idx16 ? run_scan_pipeline<u16>(in, u, pl, work, out, ntok_out) : run_scan_pipeline<u32>(in, u, pl, work, out, ntok_out);
}

// Shared-memory suffix sort of a block of up to 64 KiB (zq_sufsort16.cuh), followed by k_suffix_sort for what it hands
// back.  Outputs the four arrays + packed rows the parse kernels read; returns 1 if the general sorter was needed.
extern "C" int emu_sort16(const uint8_t* data, uint32_t n, uint16_t* sa, uint16_t* isa, uint16_t* lcp, uint8_t* bwt, uint32_t* pk) {
  std::vector<u8> in(data, data + n); in.resize(n + 64);
  std::vector<u8> work(zq_work_bytes_scan(n, 2) + 256);
  ZqUnit u; memset(&u, 0, sizeof u);
  u.n = n; u.idx16 = 1; u.want_pk = 1;
  int todo = 0;
  u32 flag = 7, next = 0;
  emu::launch(1, S16_NT, sizeof(Sort16Smem), [&] { k_suffix_sort16(in.data(), &u, &todo, 1, work.data(), &flag, &next); });
  if (flag) {
    const size_t scr = (((size_t)n + 1) + 63) & ~(size_t)63;
    std::vector<u64> kbuf(2 * scr); std::vector<u32> vbuf(6 * scr);
    emu::launch(1, 256, sizeof(SortSmem<256>), [&] { k_suffix_sort<256, 4>(in.data(), &u, &todo, 1, work.data(), kbuf.data(), vbuf.data(), scr, &flag); });
  }
  const u64 stride = zq_work_stride(n, 2);
  memcpy(sa, work.data(), 2 * (size_t)n); memcpy(isa, work.data() + stride, 2 * (size_t)n);
  memcpy(lcp, work.data() + 2 * stride, 2 * (size_t)n); memcpy(bwt, work.data() + 2 * stride + zq_work_stride(n, 2), n);
  memcpy(pk, work.data() + zq_work_bytes(n, 2), 4 * (size_t)n);
  return (int)flag;
}

#include "zq_frame.cuh"

// hash-table LZ77 parse (-m1 and the low-redundancy forms of -m2..-m4): k_lz77_hash over a zeroed 2^args[5] table
extern "C" long emu_lz_hash(const uint8_t* data, uint32_t n, const int* args, uint8_t* out, uint32_t cap) {
  std::vector<u8> in(data, data + n); in.resize(n + 64);
  std::vector<u32> ht((size_t)1 << args[5], 0u);
  ZqUnit u; memset(&u, 0, sizeof u);
  u.n = n; u.lz_cap = cap;
  ZqPlan pl; memset(&pl, 0, sizeof pl);
  for (int k = 0; k < 9; ++k) pl.args[k] = args[k];
  pl.lz_level = args[1] & 3;
  int todo = 0;
  u32 lzlen = 0, err = 0, next = 0;
  emu::launch(1, 32, 0, [&] { k_lz77_hash<8>(in.data(), &u, &pl, &todo, 1, (u8*)ht.data(), out, &lzlen, &err, &next); });
  if (err) return -(long)err;
  return (long)lzlen;
}

// BWT pre-pass (level 3): k_suffix_sort then k_bwt_stream; n + 5 bytes
extern "C" long emu_bwt(const uint8_t* data, uint32_t n, uint8_t* out) {
  const bool idx16 = n <= 65536;
  const u32 w = idx16 ? 2 : 4;
  std::vector<u8> in(data, data + n); in.resize(n + 64);
  std::vector<u8> work(zq_work_bytes(n, w) + 256);
  ZqUnit u; memset(&u, 0, sizeof u);
  u.n = n; u.idx16 = idx16;
  int todo = 0;
  const size_t scr = (((size_t)n + 1) + 63) & ~(size_t)63;
  std::vector<u64> kbuf(2 * scr); std::vector<u32> vbuf(6 * scr);
  emu::launch(1, 256, sizeof(SortSmem<256>), [&] { k_suffix_sort<256, 4>(in.data(), &u, &todo, 1, work.data(), kbuf.data(), vbuf.data(), scr); });
  u32 lzlen = 0;
  emu::launch(1, 256, 0, [&] { k_bwt_stream(in.data(), &u, &todo, 1, work.data(), out, &lzlen); });
  return (long)lzlen;
}

// E8E9 filter in place (one thread per block); level >= 5 byte-gap period analysis (one CTA per block)
extern "C" void emu_e8e9(uint8_t* data, uint32_t n) {
  std::vector<u8> in(data, data + n); in.resize(n + 64);
  ZqUnit u; memset(&u, 0, sizeof u);
  u.n = n;
  int todo = 0;
  emu::launch(1, 64, 0, [&] { k_e8e9(in.data(), &u, &todo, 1); });
  memcpy(data, in.data(), n);
}
extern "C" void emu_gap_periods(const uint8_t* data, uint32_t n, int* periods2) {
  std::vector<u8> in(data, data + n); in.resize(n + 64);
  const u64 off = 0;
  emu::launch(1, 256, 0, [&] { k_gap_periods(in.data(), &off, &n, 1, periods2); });
}
