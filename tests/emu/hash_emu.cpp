// hash_emu.cpp -- TEST INFRASTRUCTURE ONLY: the digest kernels (zq_sha1.cuh, zq_hashes.cuh) on the host through
// tests/emu/simt_emu.h; the launch sequences mirror zq_sha1 / zq_sha256 / zq_xxh3_128 / zq_blake3 in zq_api.cu.
#include <cuda_runtime.h>   // the shim

#include <algorithm>

#include "zq_sha1.cuh"
#include "zq_hashes.cuh"
#include "zq_hashes2.cuh"
#include "zq_hashes3.cuh"
#include <cmath>

using namespace zqdev;

// kind: 0 SHA-1 (20 B), 1 SHA-256 (32 B), 2 XXH3-128 (16 B), 3 BLAKE3 (32 B), 4 CRC-32 (4 B), 5 XXH64 (8 B), 6 MD5 (16 B),
// 7 SHA3-256 (32 B)
extern "C" int emu_hash(int kind, const uint8_t* base, const uint64_t* off, const uint64_t* len, int n, uint8_t* digests) {
  if (n <= 0) return 0;
  if (kind == 0) {
    emu::launch((n + 127) / 128, 128, 0, [&] { k_sha1_many(base, off, nullptr, len, n, digests); });
  } else if (kind == 1) {
    emu::launch((n + 127) / 128, 128, 0, [&] { k_sha256_many(base, off, len, n, digests); });
  } else if (kind == 2) {
    emu::launch((n + 3) / 4, 128, 0, [&] { k_xxh3_128_many(base, off, len, n, digests); });
  } else if (kind == 4) {
    // tables exactly as crc_tables() in zq_api.cu builds them
    static CrcTables t;
    for (u32 v = 0; v < 256; ++v) { u32 c = v; for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u))); t.T[0][v] = c; }
    for (u32 v = 0; v < 256; ++v) for (int k = 1; k < 4; ++k) t.T[k][v] = (t.T[k - 1][v] >> 8) ^ t.T[0][t.T[k - 1][v] & 255];
    for (int k = 0; k < 4; ++k) for (u32 v = 0; v < 256; ++v) { u32 s = v << (8 * k); for (u32 z = 0; z < CRC_CHUNK; ++z) s = t.T[0][s & 255] ^ (s >> 8); t.Z[k][v] = s; }
    std::vector<u64> first(n + 1);
    u64 tot = 0;
    for (int i = 0; i < n; ++i) { first[i] = tot; tot += (len[i] + CRC_CHUNK - 1) / CRC_CHUNK; }
    first[n] = tot;
    std::vector<u32> part(tot + 1);
    if (tot) emu::launch((unsigned)((tot + 127) / 128), 128, 0, [&] { k_crc32_chunks(base, off, len, first.data(), n, tot, &t, part.data()); });
    emu::launch((n + 127) / 128, 128, 0, [&] { k_crc32_fold(len, first.data(), n, &t, part.data(), (u32*)digests); });
  } else if (kind == 5) {
    emu::launch((n + 127) / 128, 128, 0, [&] { k_xxh64_many(base, off, len, n, (u64*)digests); });
  } else if (kind == 6) {
    static Md5Consts K;     // as zq_md5 in zq_api.cu builds it
    for (int i = 0; i < 64; ++i) K.K[i] = (uint32_t)(long long)floor(fabs(sin((double)(i + 1))) * 4294967296.0);
    emu::launch((n + 127) / 128, 128, 0, [&] { k_md5_many(base, off, len, n, &K, digests); });
  } else if (kind == 7) {
    emu::launch((n + 127) / 128, 128, 0, [&] { k_sha3_256_many(base, off, len, n, digests); });
  } else {
    std::vector<u64> first(n + 1);
    std::vector<int> multi;
    u64 tot = 0;
    for (int i = 0; i < n; ++i) {
      first[i] = tot;
      const u64 k = len[i] ? (len[i] + 1023) / 1024 : 1;
      if (k > 1) multi.push_back(i);
      tot += k;
    }
    first[n] = tot;
    std::vector<u32> cvA(tot * 8 + 8), cvB(tot * 8 + 8);
    emu::launch((unsigned)((tot + 127) / 128), 128, 0, [&] { k_blake3_chunks(base, off, len, first.data(), n, tot, cvA.data(), digests); });
    if (!multi.empty())
      emu::launch((unsigned)std::min<size_t>(multi.size(), 8), 256, 0,
                  [&] { k_blake3_tree(first.data(), multi.data(), (int)multi.size(), cvA.data(), cvB.data(), digests); });
  }
  return 0;
}

// the fragment index kernels (k_dedup_insert / k_dedup_lookup) as zq_dedup_first launches them
extern "C" int emu_dedup_first(const uint8_t* sha1, uint32_t n, uint32_t* first) {
  u32 slots = 1024;
  while (slots < 2 * (u64)n) slots <<= 1;
  std::vector<u32> tab(slots, 0xffffffffu);
  const u32* dg = (const u32*)sha1;
  emu::launch((n + 255) / 256, 256, 0, [&] { k_dedup_insert(dg, n, tab.data(), slots - 1); });
  emu::launch((n + 255) / 256, 256, 0, [&] { k_dedup_lookup(dg, n, tab.data(), slots - 1, first); });
  return 0;
}

// k_sha1_continue / k_sha256_continue: `words` = 5 or 8
extern "C" int emu_sha_continue(int words, uint32_t* state, const uint8_t* data, uint64_t nblocks) {
  std::vector<uint4> al(nblocks * 4 + 1);            // the kernels read 16-byte words
  memcpy(al.data(), data, nblocks * 64);
  if (words == 5) emu::launch(1, 32, 0, [&] { k_sha1_continue((const u8*)al.data(), nblocks, state); });
  else emu::launch(1, 32, 0, [&] { k_sha256_continue((const u8*)al.data(), nblocks, state); });
  return 0;
}
