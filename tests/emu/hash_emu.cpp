// hash_emu.cpp -- TEST INFRASTRUCTURE ONLY: the digest kernels (zq_sha1.cuh, zq_hashes.cuh) on the host through
// tests/emu/simt_emu.h; the launch sequences mirror zq_sha1 / zq_sha256 / zq_xxh3_128 / zq_blake3 in zq_api.cu.
#include <cuda_runtime.h>   // the shim

#include <algorithm>

#include "zq_sha1.cuh"
#include "zq_hashes.cuh"

using namespace zqdev;

// kind: 0 SHA-1 (20 B), 1 SHA-256 (32 B), 2 XXH3-128 (16 B), 3 BLAKE3 (32 B); buffers base+off[i], len[i]
extern "C" int emu_hash(int kind, const uint8_t* base, const uint64_t* off, const uint64_t* len, int n, uint8_t* digests) {
  if (n <= 0) return 0;
  if (kind == 0) {
    emu::launch((n + 127) / 128, 128, 0, [&] { k_sha1_many(base, off, nullptr, len, n, digests); });
  } else if (kind == 1) {
    emu::launch((n + 127) / 128, 128, 0, [&] { k_sha256_many(base, off, len, n, digests); });
  } else if (kind == 2) {
    emu::launch((n + 3) / 4, 128, 0, [&] { k_xxh3_128_many(base, off, len, n, digests); });
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
