// zq_hashes.cuh -- the per-file / per-fragment digests of the archiver besides SHA-1:
//   SHA-256     libzpaq::SHA256 (Z:12828-13015)              one THREAD per buffer (serial compression chain)
//   XXH3-128    XXH3_128bits, seed 0, default secret (Z:24635-27034; printed high64||low64, Z:67189)
//               one WARP per buffer: the 128 products of a 1 KiB block are summed by the lanes, then scrambled
//   BLAKE3      blake3_hasher_* (Z:21747-22470, portable path)  chunk-parallel: one thread per 1 KiB chunk, then a
//               per-buffer tree reduction (one CTA per buffer, pairwise with the odd node carried up)
// All integer; algorithms from the published specifications (FIPS 180-4, xxHash v0.8 XXH3, BLAKE3 paper).
#pragma once
#include "zq_common.cuh"

namespace zqdev {

// ------------------------------------------------------------------------------------------ SHA-256
__constant__ u32 kSha256K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ u32 rotr32(u32 x, int k) { return __funnelshift_r(x, x, k); }

__device__ void sha256_rounds(u32 (&st)[8], u32 (&w)[16]) {
  u32 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
  for (int t = 0; t < 64; ++t) {
    u32 wt;
    if (t < 16) wt = w[t];
    else {
      const u32 w15 = w[(t - 15) & 15], w2 = w[(t - 2) & 15];
      const u32 s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3), s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      wt = w[t & 15] + s0 + w[(t - 7) & 15] + s1;
      w[t & 15] = wt;
    }
    const u32 S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25), ch = g ^ (e & (f ^ g));
    const u32 t1 = h + S1 + ch + kSha256K[t] + wt;
    const u32 S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22), mj = (a & b) | (c & (a | b));
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + S0 + mj;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

__global__ void __launch_bounds__(128) k_sha256_many(const u8* __restrict__ base, const u64* __restrict__ off,
                                                     const u64* __restrict__ len, int n, u8* __restrict__ digests) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8* __restrict__ p = base + off[i];
  const u64 L = len[i];
  u32 st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  u32 w[16];
  const u64 nfull = L >> 6;
  for (u64 b = 0; b < nfull; ++b) {
    const u8* q = p + (b << 6);
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = (u32)q[4 * k] << 24 | (u32)q[4 * k + 1] << 16 | (u32)q[4 * k + 2] << 8 | q[4 * k + 3];
    sha256_rounds(st, w);
  }
  const u32 r = (u32)(L & 63);
  const u8* t = p + (nfull << 6);
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = 0;
  for (u32 k = 0; k <= r; ++k) {
    const u32 byte = k < r ? (u32)t[k] : 0x80u;
#pragma unroll
    for (int q = 0; q < 16; ++q) if ((int)(k >> 2) == q) w[q] |= byte << (24 - 8 * (k & 3));
  }
  if (r >= 56) {
    sha256_rounds(st, w);
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = 0;
  }
  w[14] = (u32)((L << 3) >> 32); w[15] = (u32)(L << 3);
  sha256_rounds(st, w);
  u8* d = digests + (size_t)i * 32;
#pragma unroll
  for (int k = 0; k < 8; ++k) { d[4 * k] = st[k] >> 24; d[4 * k + 1] = st[k] >> 16; d[4 * k + 2] = st[k] >> 8; d[4 * k + 3] = st[k]; }
}

// ------------------------------------------------------------------------------------------ XXH3-128
__constant__ u8 kXxhSecret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c, 0xde, 0xd4, 0x6d, 0xe9,
    0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f, 0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78,
    0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21, 0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6,
    0x81, 0x3a, 0x26, 0x4c, 0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8, 0xa8, 0xfa, 0x76, 0x3f,
    0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d, 0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31,
    0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64, 0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff,
    0xfa, 0x13, 0x63, 0xeb, 0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce, 0x45, 0xcb, 0x3a, 0x8f,
    0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e};

constexpr u64 XP1 = 0x9E3779B185EBCA87ull, XP2 = 0xC2B2AE3D27D4EB4Full, XP3 = 0x165667B19E3779F9ull,
              XP4 = 0x85EBCA77C2B2AE63ull, XP5 = 0x27D4EB2F165667C5ull;
constexpr u32 XQ1 = 0x9E3779B1u, XQ2 = 0x85EBCA77u, XQ3 = 0xC2B2AE3Du;

__device__ __forceinline__ u64 rd64(const u8* p) { u64 v = 0; for (int k = 7; k >= 0; --k) v = v << 8 | p[k]; return v; }
__device__ __forceinline__ u32 rd32(const u8* p) { return (u32)p[0] | (u32)p[1] << 8 | (u32)p[2] << 16 | (u32)p[3] << 24; }
__device__ __forceinline__ u64 sec64(int o) { return rd64(kXxhSecret + o); }
__device__ __forceinline__ u64 mul128_fold64(u64 a, u64 b) { return (a * b) ^ __umul64hi(a, b); }
__device__ __forceinline__ u64 xxh64_avalanche(u64 h) { h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32; return h; }
__device__ __forceinline__ u64 xxh3_avalanche(u64 h) { h ^= h >> 37; h *= 0x165667919E3779F9ull; h ^= h >> 32; return h; }
__device__ __forceinline__ u64 bswap64(u64 x) { return ((u64)__byte_perm((u32)x, 0, 0x0123) << 32) | __byte_perm((u32)(x >> 32), 0, 0x0123); }
__device__ __forceinline__ u64 mix16(const u8* in, int so) { return mul128_fold64(rd64(in) ^ sec64(so), rd64(in + 8) ^ sec64(so + 8)); }
struct X128 { u64 lo, hi; };
__device__ __forceinline__ X128 mix32(X128 a, const u8* i1, const u8* i2, int so) {
  a.lo += mix16(i1, so); a.lo ^= rd64(i2) + rd64(i2 + 8);
  a.hi += mix16(i2, so + 16); a.hi ^= rd64(i1) + rd64(i1 + 8);
  return a;
}

// inputs of at most 240 bytes (seed 0): executed by one lane
__device__ X128 xxh3_128_short(const u8* __restrict__ in, u32 len) {
  X128 h;
  if (len == 0) { h.lo = xxh64_avalanche(sec64(64) ^ sec64(72)); h.hi = xxh64_avalanche(sec64(80) ^ sec64(88)); return h; }
  if (len <= 3) {
    const u32 c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
    const u32 cl = (c1 << 16) | (c2 << 24) | c3 | (len << 8);
    const u32 sw = __byte_perm(cl, 0, 0x0123);
    const u32 chh = __funnelshift_l(sw, sw, 13);
    const u64 fl = (u64)(rd32(kXxhSecret) ^ rd32(kXxhSecret + 4)), fh = (u64)(rd32(kXxhSecret + 8) ^ rd32(kXxhSecret + 12));
    h.lo = xxh64_avalanche((u64)cl ^ fl); h.hi = xxh64_avalanche((u64)chh ^ fh);
    return h;
  }
  if (len <= 8) {
    const u64 in64 = (u64)rd32(in) + ((u64)rd32(in + len - 4) << 32);
    const u64 keyed = in64 ^ (sec64(16) ^ sec64(24));
    const u64 m = XP1 + ((u64)len << 2);
    u64 lo = keyed * m, hi = __umul64hi(keyed, m);
    hi += lo << 1; lo ^= hi >> 3;
    lo ^= lo >> 35; lo *= 0x9FB21C651E98DF25ull; lo ^= lo >> 28;
    h.lo = lo; h.hi = xxh3_avalanche(hi);
    return h;
  }
  if (len <= 16) {
    const u64 fl = sec64(32) ^ sec64(40), fh = sec64(48) ^ sec64(56);
    const u64 ilo = rd64(in); u64 ihi = rd64(in + len - 8);
    const u64 x = ilo ^ ihi ^ fl;
    u64 mlo = x * XP1, mhi = __umul64hi(x, XP1);
    mlo += (u64)(len - 1) << 54;
    ihi ^= fh;
    mhi += ihi + (u64)(u32)ihi * (u64)(XQ2 - 1);
    mlo ^= bswap64(mhi);
    u64 hlo = mlo * XP2, hhi = __umul64hi(mlo, XP2) + mhi * XP2;
    h.lo = xxh3_avalanche(hlo); h.hi = xxh3_avalanche(hhi);
    return h;
  }
  X128 acc; acc.lo = (u64)len * XP1; acc.hi = 0;
  if (len <= 128) {
    if (len > 32) {
      if (len > 64) {
        if (len > 96) acc = mix32(acc, in + 48, in + len - 64, 96);
        acc = mix32(acc, in + 32, in + len - 48, 64);
      }
      acc = mix32(acc, in + 16, in + len - 32, 32);
    }
    acc = mix32(acc, in, in + len - 16, 0);
  } else {
    const int rounds = (int)len / 32;
    for (int i = 0; i < 4; ++i) acc = mix32(acc, in + 32 * i, in + 32 * i + 16, 32 * i);
    acc.lo = xxh3_avalanche(acc.lo); acc.hi = xxh3_avalanche(acc.hi);
    for (int i = 4; i < rounds; ++i) acc = mix32(acc, in + 32 * i, in + 32 * i + 16, 3 + 32 * (i - 4));
    acc = mix32(acc, in + len - 16, in + len - 32, 136 - 17 - 16);
  }
  h.lo = xxh3_avalanche(acc.lo + acc.hi);
  h.hi = 0 - xxh3_avalanche(acc.lo * XP1 + acc.hi * XP4 + (u64)len * XP2);
  return h;
}

// one warp per buffer; lanes l = 8*s + a handle accumulator a of stripes s, s+4, s+8, s+12 of each block
__global__ void __launch_bounds__(128) k_xxh3_128_many(const u8* __restrict__ base, const u64* __restrict__ off,
                                                       const u64* __restrict__ len, int n, u8* __restrict__ digests) {
  const int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wi >= n) return;
  const u32 lane = lane_id();
  const u8* __restrict__ in = base + off[wi];
  const u64 L = len[wi];
  X128 h;
  if (L <= 240) {
    if (lane == 0) h = xxh3_128_short(in, (u32)L);
  } else {
    const u32 a = lane & 7, sgrp = lane >> 3;
    const u64 init[8] = {XQ3, XP1, XP2, XP3, XP4, XQ2, XP5, XQ1};
    u64 acc = 0;                      // this lane's partial sum for accumulator a; lanes 0-7 also hold the running value
    u64 run = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) if ((int)a == k) run = init[k];
    const u64 nblocks = (L - 1) / 1024;
    auto stripe = [&](const u8* sp, int so) {   // adds stripe contribution for accumulator a into acc
      const u64 dv = rd64(sp + 8 * a), dk = dv ^ sec64(so + 8 * a);
      const u64 dn = rd64(sp + 8 * (a ^ 1));       // acc[a] += data_val of the neighbour lane a^1
      acc += dn + (u64)(u32)dk * (u64)(u32)(dk >> 32);
    };
    auto fold = [&]() {   // sum the four stripe groups into lanes 0-7, add to the running accumulators
      acc += __shfl_xor_sync(ZQ_FULL, acc, 8);
      acc += __shfl_xor_sync(ZQ_FULL, acc, 16);
      run += acc; acc = 0;
    };
    for (u64 b = 0; b < nblocks; ++b) {
      const u8* bp = in + b * 1024;
#pragma unroll
      for (int q = 0; q < 4; ++q) stripe(bp + 64 * (sgrp + 4 * q), 8 * (sgrp + 4 * q));
      fold();
      run = (run ^ (run >> 47) ^ sec64(128 + 8 * a)) * (u64)XQ1;   // scramble with secret[192-64 ..]
    }
    const u64 nstripes = ((L - 1) - 1024 * nblocks) / 64;
    const u8* bp = in + nblocks * 1024;
    for (u64 s = sgrp; s < nstripes; s += 4) stripe(bp + 64 * s, (int)(8 * s));
    if (sgrp == 0) stripe(in + L - 64, 192 - 64 - 7);
    fold();
    // merge (lanes 0-7 hold the accumulators)
    const u64 a0 = __shfl_sync(ZQ_FULL, run, 0), a1 = __shfl_sync(ZQ_FULL, run, 1), a2 = __shfl_sync(ZQ_FULL, run, 2),
              a3 = __shfl_sync(ZQ_FULL, run, 3), a4 = __shfl_sync(ZQ_FULL, run, 4), a5 = __shfl_sync(ZQ_FULL, run, 5),
              a6 = __shfl_sync(ZQ_FULL, run, 6), a7 = __shfl_sync(ZQ_FULL, run, 7);
    auto merge = [&](int so, u64 start) {
      u64 r = start;
      r += mul128_fold64(a0 ^ sec64(so), a1 ^ sec64(so + 8));
      r += mul128_fold64(a2 ^ sec64(so + 16), a3 ^ sec64(so + 24));
      r += mul128_fold64(a4 ^ sec64(so + 32), a5 ^ sec64(so + 40));
      r += mul128_fold64(a6 ^ sec64(so + 48), a7 ^ sec64(so + 56));
      return xxh3_avalanche(r);
    };
    h.lo = merge(11, L * XP1);
    h.hi = merge(192 - 64 - 11, ~(L * XP2));
  }
  if (lane == 0) {
    u8* d = digests + (size_t)wi * 16;
    for (int k = 0; k < 8; ++k) { d[k] = (u8)(h.hi >> (56 - 8 * k)); d[8 + k] = (u8)(h.lo >> (56 - 8 * k)); }
  }
}

// ------------------------------------------------------------------------------------------ BLAKE3
__device__ __forceinline__ void b3_g(u32& a, u32& b, u32& c, u32& d, u32 x, u32 y) {
  a = a + b + x; d = rotr32(d ^ a, 16); c = c + d; b = rotr32(b ^ c, 12);
  a = a + b + y; d = rotr32(d ^ a, 8); c = c + d; b = rotr32(b ^ c, 7);
}
// out[0..7] = first half of the compression output (chaining value / first 32 digest bytes)
__device__ void b3_compress(const u32 (&cv)[8], const u32 (&m)[16], u64 counter, u32 blen, u32 flags, u32 (&out)[8]) {
  u32 v0 = cv[0], v1 = cv[1], v2 = cv[2], v3 = cv[3], v4 = cv[4], v5 = cv[5], v6 = cv[6], v7 = cv[7];
  u32 v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
  u32 v12 = (u32)counter, v13 = (u32)(counter >> 32), v14 = blen, v15 = flags;
  // message schedule: the fixed permutation applied r times, fully unrolled as index tables
  const int S[7][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8},
                        {3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1}, {10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6},
                        {12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4}, {9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7},
                        {11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13}};
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    b3_g(v0, v4, v8, v12, m[S[r][0]], m[S[r][1]]);
    b3_g(v1, v5, v9, v13, m[S[r][2]], m[S[r][3]]);
    b3_g(v2, v6, v10, v14, m[S[r][4]], m[S[r][5]]);
    b3_g(v3, v7, v11, v15, m[S[r][6]], m[S[r][7]]);
    b3_g(v0, v5, v10, v15, m[S[r][8]], m[S[r][9]]);
    b3_g(v1, v6, v11, v12, m[S[r][10]], m[S[r][11]]);
    b3_g(v2, v7, v8, v13, m[S[r][12]], m[S[r][13]]);
    b3_g(v3, v4, v9, v14, m[S[r][14]], m[S[r][15]]);
  }
  out[0] = v0 ^ v8; out[1] = v1 ^ v9; out[2] = v2 ^ v10; out[3] = v3 ^ v11;
  out[4] = v4 ^ v12; out[5] = v5 ^ v13; out[6] = v6 ^ v14; out[7] = v7 ^ v15;
}

// thread t -> chunk t of the flattened chunk list; chunk_first[b] = first flattened chunk of buffer b
__global__ void __launch_bounds__(128)
k_blake3_chunks(const u8* __restrict__ base, const u64* __restrict__ off, const u64* __restrict__ len,
                const u64* __restrict__ chunk_first, int nbuf, u64 nchunks, u32* __restrict__ cvs, u8* __restrict__ digests) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  int lo = 0, hi = nbuf - 1;   // last b with chunk_first[b] <= t
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (chunk_first[mid] <= t) lo = mid; else hi = mid - 1; }
  const int b = lo;
  const u64 ci = t - chunk_first[b];
  const u64 L = len[b];
  const u64 cbeg = ci * 1024;
  const u32 clen = (u32)min((u64)1024, L - cbeg);
  const bool single = chunk_first[b + 1] - chunk_first[b] == 1;
  const u8* __restrict__ p = base + off[b] + cbeg;
  u32 cv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  const u32 nblk = clen ? (clen + 63) / 64 : 1;
  for (u32 k = 0; k < nblk; ++k) {
    u32 m[16];
    const u32 bl = min(64u, clen - k * 64);
    const u8* q = p + k * 64;
    if (bl == 64) {
#pragma unroll
      for (int w = 0; w < 16; ++w) m[w] = rd32(q + 4 * w);
    } else {
#pragma unroll
      for (int w = 0; w < 16; ++w) {
        u32 x = 0;
        for (int z = 3; z >= 0; --z) x = x << 8 | ((u32)(4 * w + z) < bl ? (u32)q[4 * w + z] : 0u);
        m[w] = x;
      }
    }
    u32 flags = (k == 0 ? 1u : 0u) | (k == nblk - 1 ? 2u : 0u);
    if (single && k == nblk - 1) flags |= 8u;
    u32 o[8];
    b3_compress(cv, m, ci, bl, flags, o);
#pragma unroll
    for (int w = 0; w < 8; ++w) cv[w] = o[w];
  }
  if (single) {
    u8* d = digests + (size_t)b * 32;
#pragma unroll
    for (int w = 0; w < 8; ++w) { d[4 * w] = cv[w]; d[4 * w + 1] = cv[w] >> 8; d[4 * w + 2] = cv[w] >> 16; d[4 * w + 3] = cv[w] >> 24; }
  } else {
#pragma unroll
    for (int w = 0; w < 8; ++w) cvs[t * 8 + w] = cv[w];
  }
}

// one CTA per multi-chunk buffer: pairwise parent compressions level by level, odd node carried up
__global__ void __launch_bounds__(256)
k_blake3_tree(const u64* __restrict__ chunk_first, const int* __restrict__ multi, int nmulti, u32* __restrict__ cvA,
              u32* __restrict__ cvB, u8* __restrict__ digests) {
  for (int t = blockIdx.x; t < nmulti; t += gridDim.x) {
    const int b = multi[t];
    const u64 o = chunk_first[b];
    u64 cnt = chunk_first[b + 1] - o;
    u32* src = cvA + o * 8; u32* dst = cvB + o * 8;
    const u32 iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    while (cnt > 1) {
      const u64 np = cnt >> 1;
      const bool root = cnt == 2;
      for (u64 k = threadIdx.x; k < np; k += blockDim.x) {
        u32 m[16], out[8];
#pragma unroll
        for (int w = 0; w < 16; ++w) m[w] = src[k * 16 + w];
        b3_compress(iv, m, 0, 64, 4u | (root ? 8u : 0u), out);
        if (root) {
          u8* d = digests + (size_t)b * 32;
#pragma unroll
          for (int w = 0; w < 8; ++w) { d[4 * w] = out[w]; d[4 * w + 1] = out[w] >> 8; d[4 * w + 2] = out[w] >> 16; d[4 * w + 3] = out[w] >> 24; }
        } else {
#pragma unroll
          for (int w = 0; w < 8; ++w) dst[k * 8 + w] = out[w];
        }
      }
      if ((cnt & 1) && threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) dst[np * 8 + w] = src[(cnt - 1) * 8 + w];
      cnt = np + (cnt & 1);
      __syncthreads();
      u32* tmp = src; src = dst; dst = tmp;
    }
    __syncthreads();
  }
}

}  // namespace zqdev
