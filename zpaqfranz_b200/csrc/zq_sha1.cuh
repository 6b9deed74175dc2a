// zq_sha1.cuh -- batched SHA-1 (FIPS 180-4), one THREAD per independent buffer.
//
// Replaces libzpaq::SHA1::write/result (Z:12637-12819; SHA-NI variant Z:11313-12251) as used for the
// block trailer (compressBlock, Z:20278-20287) and the fragment dedup key (Z:122569-122573).
// A single SHA-1 stream is a strict 80-round dependency chain per 64-byte block, so the parallelism
// is across buffers: lanes of a warp hash 32 different buffers in lock step (no divergence: every
// full block executes the same 80 rounds).  Integer-ALU bound; ~1 B read per input byte.
#pragma once
#include "zq_common.cuh"

namespace zqdev {

__device__ __forceinline__ u32 rol32(u32 x, int k) { return __funnelshift_l(x, x, k); }
__device__ __forceinline__ u32 bswap32(u32 x) { return __byte_perm(x, 0, 0x0123); }

struct Sha1State { u32 h0, h1, h2, h3, h4; };

__device__ __forceinline__ void sha1_init(Sha1State& s) {
  s.h0 = 0x67452301u; s.h1 = 0xEFCDAB89u; s.h2 = 0x98BADCFEu; s.h3 = 0x10325476u; s.h4 = 0xC3D2E1F0u;
}

// w[16]: big-endian message words; fully unrolled so w[] stays in registers
__device__ __forceinline__ void sha1_rounds(Sha1State& s, u32 (&w)[16]) {
  u32 a = s.h0, b = s.h1, c = s.h2, d = s.h3, e = s.h4;
#pragma unroll
  for (int t = 0; t < 80; ++t) {
    u32 wt;
    if (t < 16) wt = w[t];
    else {
      wt = rol32(w[(t - 3) & 15] ^ w[(t - 8) & 15] ^ w[(t - 14) & 15] ^ w[t & 15], 1);
      w[t & 15] = wt;
    }
    u32 f, k;
    if (t < 20) { f = d ^ (b & (c ^ d)); k = 0x5A827999u; }
    else if (t < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
    else if (t < 60) { f = (b & c) | (d & (b | c)); k = 0x8F1BBCDCu; }
    else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
    u32 tmp = rol32(a, 5) + f + e + k + wt;
    e = d; d = c; c = rol32(b, 30); b = a; a = tmp;
  }
  s.h0 += a; s.h1 += b; s.h2 += c; s.h3 += d; s.h4 += e;
}

// Hash p[0..n) (any alignment) and store the 20-byte digest.
__device__ void sha1_buffer(const u8* __restrict__ p, u64 n, u8* __restrict__ digest) {
  Sha1State st; sha1_init(st);
  u32 w[16];
  const u64 nfull = n >> 6;
  const u32 sh = (u32)((uintptr_t)p & 3) * 8;
  const u32* __restrict__ pa = (const u32*)((uintptr_t)p & ~(uintptr_t)3);
  if (sh == 0) {
    if (((uintptr_t)p & 15) == 0) {
      const uint4* __restrict__ p4 = (const uint4*)p;
      for (u64 b = 0; b < nfull; ++b) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v = __ldg(p4 + b * 4 + q);
          w[4 * q] = bswap32(v.x); w[4 * q + 1] = bswap32(v.y); w[4 * q + 2] = bswap32(v.z); w[4 * q + 3] = bswap32(v.w);
        }
        sha1_rounds(st, w);
      }
    } else {
      for (u64 b = 0; b < nfull; ++b) {
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q] = bswap32(__ldg(pa + b * 16 + q));
        sha1_rounds(st, w);
      }
    }
  } else {
    for (u64 b = 0; b < nfull; ++b) {
      u32 lo = __ldg(pa + b * 16);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        u32 hi = __ldg(pa + b * 16 + q + 1);  // last one still overlaps the buffer (sh != 0)
        w[q] = bswap32(__funnelshift_r(lo, hi, sh));
        lo = hi;
      }
      sha1_rounds(st, w);
    }
  }
  // tail: r bytes, 0x80, zeros, 64-bit big-endian bit count
  const u32 r = (u32)(n & 63);
  const u8* t = p + (nfull << 6);
#pragma unroll
  for (int q = 0; q < 16; ++q) w[q] = 0;
  for (u32 k = 0; k < r; ++k) {
    const u32 byte = t[k];
#pragma unroll
    for (int q = 0; q < 16; ++q) if ((int)(k >> 2) == q) w[q] |= byte << (24 - 8 * (k & 3));
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) if ((int)(r >> 2) == q) w[q] |= 0x80u << (24 - 8 * (r & 3));
  const u64 bits = n << 3;
  if (r >= 56) {
    sha1_rounds(st, w);
#pragma unroll
    for (int q = 0; q < 16; ++q) w[q] = 0;
  }
  w[14] = (u32)(bits >> 32); w[15] = (u32)bits;
  sha1_rounds(st, w);
  const u32 hs[5] = {st.h0, st.h1, st.h2, st.h3, st.h4};
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    digest[4 * q] = hs[q] >> 24; digest[4 * q + 1] = hs[q] >> 16; digest[4 * q + 2] = hs[q] >> 8; digest[4 * q + 3] = hs[q];
  }
}

// n independent buffers base[off[i] .. +len[i]) -> digests[20*i..]; len32 or len64 (one non-null)
__global__ void __launch_bounds__(128) k_sha1_many(const u8* __restrict__ base, const u64* __restrict__ off,
                                                   const u32* __restrict__ len32, const u64* __restrict__ len64,
                                                   int n, u8* __restrict__ digests) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 len = len64 ? len64[i] : (u64)len32[i];
  sha1_buffer(base + off[i], len, digests + (size_t)i * 20);
}

// unit-descriptor flavour used by the block compressor
__global__ void __launch_bounds__(128) k_sha1_units(const u8* __restrict__ base, const ZqUnit* __restrict__ units,
                                                    int n, u8* __restrict__ digests) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sha1_buffer(base + units[i].in_off, units[i].n, digests + (size_t)i * 20);
}

}  // namespace zqdev
