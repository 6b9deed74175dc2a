// zq_hashes3.cuh -- MD5 and SHA3-256 of many buffers: the last two per-file digests Jidac::updatehash can be asked for
// (-md5 / -sha3: MD5::add Z:21616, SHA3::add Z:21337).  One thread per buffer -- both are
// strictly sequential per 64 / 136-byte block; 32 buffers run in lock step per warp.  Written from the specifications
// (RFC 1321; FIPS 202 with the 0x06 domain byte): the sine table and the round constants are computed, not copied.
#pragma once
#include "zq_common.cuh"
#include "zq_sha1.cuh"
#include "zq_hashes.cuh"

namespace zqdev {

__device__ __forceinline__ u32 md5_rotl(u32 x, int k) { return __funnelshift_l(x, x, k); }

struct Md5Consts { u32 K[64]; };   // floor(2^32 * |sin(i + 1)|), filled by the host

__device__ __forceinline__ void md5_block(u32 st[4], const u32 w[16], const u32* __restrict__ K) {
  u32 a = st[0], b = st[1], c = st[2], d = st[3];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    u32 f; int g, r;
    if (i < 16) { f = (b & c) | (~b & d); g = i; r = (i & 3) == 0 ? 7 : (i & 3) == 1 ? 12 : (i & 3) == 2 ? 17 : 22; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; r = (i & 3) == 0 ? 5 : (i & 3) == 1 ? 9 : (i & 3) == 2 ? 14 : 20; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; r = (i & 3) == 0 ? 4 : (i & 3) == 1 ? 11 : (i & 3) == 2 ? 16 : 23; }
    else { f = c ^ (b | ~d); g = (7 * i) & 15; r = (i & 3) == 0 ? 6 : (i & 3) == 1 ? 10 : (i & 3) == 2 ? 15 : 21; }
    const u32 t = d; d = c; c = b;
    b = b + md5_rotl(a + f + K[i] + w[g], r);
    a = t;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

// out[16 b ..] = MD5(buffer b)
__global__ void __launch_bounds__(128)
k_md5_many(const u8* __restrict__ base, const u64* __restrict__ off, const u64* __restrict__ len, int n, const Md5Consts* __restrict__ C,
           u8* __restrict__ out) {
  __shared__ u32 K[64];
  if (threadIdx.x < 64) K[threadIdx.x] = C->K[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8* __restrict__ p = base + off[i];
  const u64 L = len[i];
  u32 st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  u32 w[16];
  u64 pos = 0;
  for (; pos + 64 <= L; pos += 64) {
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = (u32)p[pos + 4 * k] | (u32)p[pos + 4 * k + 1] << 8 | (u32)p[pos + 4 * k + 2] << 16 | (u32)p[pos + 4 * k + 3] << 24;
    md5_block(st, w, K);
  }
  // tail: 0x80, zeros, 64-bit little-endian bit count
  const u32 rem = (u32)(L - pos);
  for (int pass = 0; pass < 2; ++pass) {
    bool last = true;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      u32 v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const u32 q = (u32)pass * 64 + 4 * k + b;
        const u32 c = q < rem ? (u32)p[pos + q] : q == rem ? 0x80u : 0u;
        v |= c << (8 * b);
      }
      w[k] = v;
    }
    const bool fits = rem + 9 <= 64;            // length field fits the first tail block
    if (pass == 0 && !fits) last = false;
    if (last) { w[14] = (u32)(L << 3); w[15] = (u32)((L << 3) >> 32); }
    md5_block(st, w, K);
    if (last) break;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int b = 0; b < 4; ++b) out[(size_t)i * 16 + 4 * k + b] = (u8)(st[k] >> (8 * b));
}

// ---- SHA3-256 (Keccak-f[1600], rate 136 bytes) ------------------------------------------------------------------
__device__ __forceinline__ u64 k_rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }

__device__ __forceinline__ void keccak_f(u64 s[25]) {
  // round constants from the degree-8 LFSR of FIPS 202 section 3.2.5; rotation offsets from the (x, y) walk of 3.2.2
  u32 lfsr = 1;
  for (int round = 0; round < 24; ++round) {
    u64 c[5];
#pragma unroll
    for (int x = 0; x < 5; ++x) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
#pragma unroll
    for (int x = 0; x < 5; ++x) {
      const u64 d = c[(x + 4) % 5] ^ k_rotl(c[(x + 1) % 5], 1);
#pragma unroll
      for (int y = 0; y < 25; y += 5) s[y + x] ^= d;
    }
    // rho + pi along the orbit of (1, 0) under (x, y) -> (y, 2x + 3y)
    u64 cur = s[1];
    int x = 1, y = 0;
#pragma unroll
    for (int t = 0; t < 24; ++t) {
      const int nx = y, ny = (2 * x + 3 * y) % 5;
      x = nx; y = ny;
      const u64 nxt = s[5 * y + x];
      s[5 * y + x] = k_rotl(cur, ((t + 1) * (t + 2) / 2) & 63);
      cur = nxt;
    }
#pragma unroll
    for (int y5 = 0; y5 < 25; y5 += 5) {
      u64 r[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) r[q] = s[y5 + q];
#pragma unroll
      for (int q = 0; q < 5; ++q) s[y5 + q] = r[q] ^ (~r[(q + 1) % 5] & r[(q + 2) % 5]);
    }
    u64 rc = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (lfsr & 1) rc ^= 1ull << ((1 << j) - 1);
      lfsr = (lfsr << 1) ^ ((lfsr & 0x80u) ? 0x171u : 0u);
      lfsr &= 0xffu | 0x100u;
      lfsr &= 0xffu;
    }
    s[0] ^= rc;
  }
}

// out[32 b ..] = SHA3-256(buffer b)
__global__ void __launch_bounds__(128)
k_sha3_256_many(const u8* __restrict__ base, const u64* __restrict__ off, const u64* __restrict__ len, int n, u8* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8* __restrict__ p = base + off[i];
  const u64 L = len[i];
  u64 s[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) s[k] = 0;
  u64 pos = 0;
  const u32 RATE = 136;
  for (; pos + RATE <= L; pos += RATE) {
#pragma unroll
    for (int k = 0; k < 17; ++k) {
      u64 v = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) v |= (u64)p[pos + 8 * k + b] << (8 * b);
      s[k] ^= v;
    }
    keccak_f(s);
  }
  const u32 rem = (u32)(L - pos);
#pragma unroll
  for (int k = 0; k < 17; ++k) {
    u64 v = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const u32 q = 8 * k + b;
      u32 c = q < rem ? (u32)p[pos + q] : 0u;
      if (q == rem) c ^= 0x06u;
      if (q == RATE - 1) c ^= 0x80u;
      v |= (u64)c << (8 * b);
    }
    s[k] ^= v;
  }
  keccak_f(s);
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int b = 0; b < 8; ++b) out[(size_t)i * 32 + 8 * k + b] = (u8)(s[k] >> (8 * b));
}

// ---- the fragment index: which earlier fragment has the same SHA-1 (HTIndex::find, Z:71567-71604) --------------------
// Open addressing over a table of fragment numbers; a slot belongs to one digest for good and holds the SMALLEST
// fragment number seen with it (atomicMin), so the answer does not depend on the order the threads arrive in.
static const u32 DEDUP_EMPTY = 0xffffffffu;

__device__ __forceinline__ bool dedup_same(const u32* __restrict__ dg, u32 a, const u32 (&w)[5]) {
  const u32* q = dg + 5 * (size_t)a;
  return q[0] == w[0] && q[1] == w[1] && q[2] == w[2] && q[3] == w[3] && q[4] == w[4];
}

__device__ __forceinline__ u32 dedup_slot(const u32 (&w)[5], u32 mask) { return (w[0] ^ (w[1] * 0x9E3779B1u)) & mask; }

__global__ void __launch_bounds__(256) k_dedup_insert(const u32* __restrict__ dg, u32 n, u32* tab, u32 mask) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 w[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) w[k] = dg[5 * (size_t)i + k];
  for (u32 slot = dedup_slot(w, mask);; slot = (slot + 1) & mask) {
    const u32 cur = atomicCAS(&tab[slot], DEDUP_EMPTY, i);
    if (cur == DEDUP_EMPTY) return;
    if (dedup_same(dg, cur, w)) { atomicMin(&tab[slot], i); return; }
  }
}

__global__ void __launch_bounds__(256) k_dedup_lookup(const u32* __restrict__ dg, u32 n, const u32* __restrict__ tab, u32 mask,
                                                      u32* __restrict__ first) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 w[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) w[k] = dg[5 * (size_t)i + k];
  for (u32 slot = dedup_slot(w, mask);; slot = (slot + 1) & mask) {
    const u32 cur = tab[slot];
    if (cur == DEDUP_EMPTY) { first[i] = i; return; }      // cannot happen after k_dedup_insert; keeps the loop finite
    if (dedup_same(dg, cur, w)) { first[i] = cur; return; }
  }
}

// ---- one stream, continued: the compression function over whole 64-byte blocks from a given chaining value ---------
// (libzpaq::SHA1 / SHA256 as STREAMING classes, Z:12637 / Z:12828: write() in pieces, result() at the end.)  A single
// stream is a dependency chain -- one thread; the point is bounded memory on the host, not speed: batches of buffers go
// through k_sha1_many / k_sha256_many.
__global__ void k_sha1_continue(const u8* __restrict__ p, u64 nblocks, u32* __restrict__ state) {
  if (blockIdx.x | threadIdx.x) return;
  Sha1State st; st.h0 = state[0]; st.h1 = state[1]; st.h2 = state[2]; st.h3 = state[3]; st.h4 = state[4];
  u32 w[16];
  const uint4* __restrict__ p4 = (const uint4*)p;    // device staging buffer: 16-byte aligned
  for (u64 b = 0; b < nblocks; ++b) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = p4[b * 4 + q];
      w[4 * q] = bswap32(v.x); w[4 * q + 1] = bswap32(v.y); w[4 * q + 2] = bswap32(v.z); w[4 * q + 3] = bswap32(v.w);
    }
    sha1_rounds(st, w);
  }
  state[0] = st.h0; state[1] = st.h1; state[2] = st.h2; state[3] = st.h3; state[4] = st.h4;
}
__global__ void k_sha256_continue(const u8* __restrict__ p, u64 nblocks, u32* __restrict__ state) {
  if (blockIdx.x | threadIdx.x) return;
  u32 st[8], w[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) st[k] = state[k];
  const uint4* __restrict__ p4 = (const uint4*)p;
  for (u64 b = 0; b < nblocks; ++b) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = p4[b * 4 + q];
      w[4 * q] = bswap32(v.x); w[4 * q + 1] = bswap32(v.y); w[4 * q + 2] = bswap32(v.z); w[4 * q + 3] = bswap32(v.w);
    }
    sha256_rounds(st, w);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) state[k] = st[k];
}

}  // namespace zqdev
