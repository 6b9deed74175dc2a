// zq_sufsort16.cuh -- suffix array + inverse + capped LCP + BWT + packed rows of a block of up to 64 KiB, built
// ENTIRELY in shared memory by one CTA (the 64 KiB unit of the headline benchmark).  Same result as divsufsort
// (Z:19121): the suffix array of a string is unique.
//
// Why a second sorter: k_suffix_sort (zq_sufsort.cuh) is a general prefix-doubling sort whose 40 bytes of scratch per
// input byte live in global memory -- 22 MB of DRAM traffic per 64 KiB block.  Text-like and random blocks do not need
// doubling at all: their suffixes differ within a few characters.  This kernel sorts by the first 10-12 characters
// with a network over 64-bit words (bin number inside the batch | following stream bits | 16-bit suffix index) held in
// shared memory, then orders the few suffixes that still tie (groups of 2..128) by comparing the block directly.  A
// block where that stops paying -- a group of more than 128 suffixes sharing the key, or two suffixes sharing more than
// 512 characters -- is handed to k_suffix_sort through a flag; nothing is approximated.
//
//   shared memory (~225 KB, one CTA per SM):  block 64 KiB | 8192 bin starts | 16384 x u64 sort buffer
//   1. the bytes arrive in the (still unused) sort buffer by bulk async copy; every byte is replaced by its rank among
//      the byte values that occur (b = 1..8 bits, order preserving) and the block is kept as that packed big-endian bit
//      stream only (the "dense form", S16Dense below): 64 bits at bit offset b*i = the first 64/b characters of suffix i
//   2. histogram of the first 13 stream bits of every suffix (shared-memory atomics), exclusive scan: the start row of
//      every bin (2.6 characters of the word corpus: ~3 500 bins of 19 rows on average).  Every position is dealt to
//      its bin's rows of the (global, L2-resident) sa array -- a counting sort on 13 bits.
//   3. per batch (whole bins, as many as fit the buffer): rows read back (coalesced), keyed, sorted by the all-ascending
//      network cut short at the size of the largest bin (s16_bitonic_bins); tie runs dealt to the lanes of a warp and
//      ordered by window compares; then the rows are final: sa / lcp / bwt / pk written in row order, isa scattered.
// The forms this one replaced (raw bytes: 7 per sort word, 165 bins for text; the complete network; tie runs ordered by
// the thread whose row starts one) and their times are in profiles/README.md (r04a); the code is gone.
#pragma once
#include "zq_sufsort.cuh"

namespace zqdev {

constexpr int S16_NT = 1024;
constexpr u32 S16_BINS = 8192;
#ifndef S16_BUF_ELEMS
#define S16_BUF_ELEMS 16384
#endif
constexpr u32 S16_BUF = S16_BUF_ELEMS;   // sort buffer capacity (elements) = rows per batch at most
constexpr u32 S16_MAXGROUP = 128;    // largest tie group ordered by direct comparison
constexpr u32 S16_MAXDEPTH = 512;    // longest common prefix followed by direct comparison
#ifndef S16_TIE_ROWS
#define S16_TIE_ROWS 512
#endif
constexpr u32 S16_TIE_CHUNK = S16_TIE_ROWS;   // rows a warp searches for tie-run heads at a time (16 per lane)

struct Sort16Smem {
  u32 bins[S16_BINS + 8];
  u32 cnt, prev_idx, fallback, pad0;     // pad0: end row of the current batch
  u32 wsum[32], wlo[32], whi[32], wmx[32];
  u32 blo, bhi, pad2[2];                 // first / last non-empty bin of the current batch
  ZqMbar bar;
  u64 pad1;
  u32 cbits, crcp, kused, bmx;           // bits per character, 65536 / cbits + 1, characters in use; largest bin of the batch
  u8 rank[256], inv[256], present[256];  // byte -> dense code (its rank among the bytes that occur), code -> byte
  alignas(16) u8 text[65536 + 64];       // the block as a packed character stream (S16Dense below)
  alignas(16) u64 buf[S16_BUF];
};

__device__ __forceinline__ void s16_cx(u64& a, u64& b, bool up) {
  const bool sw = (a > b) == up;
  const u64 lo = sw ? b : a, hi = sw ? a : b;
  a = lo; b = hi;
}

// ---- the sort network: bitonic merges in their all-ascending ("flip") form, cut short ------------------------------
// A merge phase of size k first compares every element with its mirror image inside the k-block (i with i ^ (k-1)),
// then runs the strides k/4 ... 1; every comparator puts the smaller word at the lower index, so padding and whole
// sorted runs never move against the final order.  That allows stopping early: the batch arrives ordered by bin (the
// bin number leads the sort word) and no bin is larger than K, so after sorting the aligned chunks of K words every
// bin lies sorted inside a chunk or split over two neighbours -- one more phase of size 2K merges the pairs (0,1),
// (2,3) ..., the same phase shifted by K the pairs (1,2), (3,4) ..., and the batch is sorted: log^2(K)/2 + 2 log(2K)
// comparator stages instead of log^2(P)/2 (65 instead of 105 for K = 512 in a batch of 16 384).
__device__ __forceinline__ void s16f_reg8(u64* __restrict__ S, u32 base, bool head) {
  // 8 consecutive words per thread = 64 bytes: with every lane reading its 16-byte pieces in the same order, lanes
  // L and L+2 hit the same banks (4-way conflict).  Every other lane pair therefore starts with its upper half: slot e
  // then holds word e ^ 4, which only turns around the pairs that are 4 apart (direction flag `lo4`).
#ifdef S16_REG8_PLAIN
  const u32 rot = 0;
#else
  const u32 rot = (threadIdx.x >> 1) & 1u;
#endif
  const bool lo4 = rot == 0;
  u64 v[8];
  const ulonglong2* src = reinterpret_cast<const ulonglong2*>(S + base);
#pragma unroll
  for (int q = 0; q < 4; ++q) { const ulonglong2 t = src[q ^ (2 * rot)]; v[2 * q] = t.x; v[2 * q + 1] = t.y; }
  if (head) {
    s16_cx(v[0], v[1], true); s16_cx(v[2], v[3], true); s16_cx(v[4], v[5], true); s16_cx(v[6], v[7], true);
    s16_cx(v[0], v[3], true); s16_cx(v[1], v[2], true); s16_cx(v[4], v[7], true); s16_cx(v[5], v[6], true);
    s16_cx(v[0], v[1], true); s16_cx(v[2], v[3], true); s16_cx(v[4], v[5], true); s16_cx(v[6], v[7], true);
    s16_cx(v[0], v[7], lo4); s16_cx(v[1], v[6], lo4); s16_cx(v[2], v[5], lo4); s16_cx(v[3], v[4], lo4);
  } else {
    s16_cx(v[0], v[4], lo4); s16_cx(v[1], v[5], lo4); s16_cx(v[2], v[6], lo4); s16_cx(v[3], v[7], lo4);
  }
  s16_cx(v[0], v[2], true); s16_cx(v[1], v[3], true); s16_cx(v[4], v[6], true); s16_cx(v[5], v[7], true);
  s16_cx(v[0], v[1], true); s16_cx(v[2], v[3], true); s16_cx(v[4], v[5], true); s16_cx(v[6], v[7], true);
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(S + base);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q ^ (2 * rot)] = make_ulonglong2(v[2 * q], v[2 * q + 1]);
}

// one merge phase of size k >= 16 over S[0..len), len a multiple of k (every thread of the CTA calls it)
// Passes whose smallest stride is 8 words make the two halves of a half-warp read the same banks (lanes 0-7 words
// 0..7 of one 32-word block, lanes 8-15 words 0..7 of the next): every second group of 8 lanes therefore takes its
// words in the opposite order (`sw`), which turns around the comparators inside a pair of slots (direction `!sw`).
#ifdef S16_NO_STAGGER
#define S16_SW(t) false
#else
#define S16_SW(t) ((((t) >> 3) & 1u) != 0)
#endif
__device__ void s16f_phase(u64* __restrict__ S, u32 len, u32 k) {
  const u32 tid = threadIdx.x;
  if (k == 16) {
    for (u32 t = tid; t < len / 2; t += S16_NT) {
      const u32 r = t & 7u, blk = (t - r) * 2;
      const bool sw = S16_SW(t);
      const u32 i0 = sw ? blk + 15 - r : blk + r, i1 = sw ? blk + r : blk + 15 - r;
      u64 x = S[i0], y = S[i1];
      s16_cx(x, y, !sw);
      S[i0] = x; S[i1] = y;
    }
    __syncthreads();
  } else {
    const u32 h = k >> 2;                                 // mirror stage and stride k/4 in one pass
    if (h == 8) {
      for (u32 t = tid; t < len / 4; t += S16_NT) {
        const u32 r = t & 7u, blk = (t - r) * 4;
        const bool sw = S16_SW(t);
        const u32 ia = blk + r, ib = ia + 8, id = blk + 31 - r, ic = id - 8;
        const u32 i0 = sw ? ib : ia, i1 = sw ? ia : ib, i2 = sw ? id : ic, i3 = sw ? ic : id;
        u64 v0 = S[i0], v1 = S[i1], v2 = S[i2], v3 = S[i3];
        s16_cx(v0, v3, true); s16_cx(v1, v2, true);       // (a,d),(b,c) -- or (b,c),(a,d)
        s16_cx(v0, v1, !sw); s16_cx(v2, v3, !sw);         // (a,b),(c,d) -- or (b,a),(d,c)
        S[i0] = v0; S[i1] = v1; S[i2] = v2; S[i3] = v3;
      }
    } else {
      for (u32 t = tid; t < len / 4; t += S16_NT) {
        const u32 r = t & (h - 1), blk = (t - r) * 4;
        const u32 ia = blk + r, ib = ia + h, id = blk + k - 1 - r, ic = id - h;
        u64 a = S[ia], b = S[ib], c = S[ic], d = S[id];
        s16_cx(a, d, true); s16_cx(b, c, true);
        s16_cx(a, b, true); s16_cx(c, d, true);
        S[ia] = a; S[ib] = b; S[ic] = c; S[id] = d;
      }
    }
    __syncthreads();
    u32 j = h >> 1;
    while (j >= 8) {
      if (j >= 32) {
        const u32 g = j >> 1;
        for (u32 t = tid; t < len / 4; t += S16_NT) {
          const u32 i = 4 * t - 3 * (t & (g - 1));
          u64 a = S[i], b = S[i + g], c = S[i + j], d = S[i + j + g];
          s16_cx(a, c, true); s16_cx(b, d, true);
          s16_cx(a, b, true); s16_cx(c, d, true);
          S[i] = a; S[i + g] = b; S[i + j] = c; S[i + j + g] = d;
        }
        j >>= 2;
      } else if (j == 16) {                               // strides 16 and 8
        for (u32 t = tid; t < len / 4; t += S16_NT) {
          const u32 i = 4 * t - 3 * (t & 7u);
          const bool sw = S16_SW(t);
          const u32 i0 = sw ? i + 8 : i, i1 = sw ? i : i + 8;
          u64 v0 = S[i0], v1 = S[i1], v2 = S[i0 + 16], v3 = S[i1 + 16];
          s16_cx(v0, v2, true); s16_cx(v1, v3, true);     // (a,c),(b,d) either way
          s16_cx(v0, v1, !sw); s16_cx(v2, v3, !sw);       // (a,b),(c,d) -- or (b,a),(d,c)
          S[i0] = v0; S[i1] = v1; S[i0 + 16] = v2; S[i1 + 16] = v3;
        }
        j >>= 2;
      } else {                                            // stride 8 alone
        for (u32 t = tid; t < len / 2; t += S16_NT) {
          const u32 i = 2 * t - (t & 7u);
          const bool sw = S16_SW(t);
          const u32 i0 = sw ? i + 8 : i, i1 = sw ? i : i + 8;
          u64 x = S[i0], y = S[i1];
          s16_cx(x, y, !sw);
          S[i0] = x; S[i1] = y;
        }
        j >>= 1;
      }
      __syncthreads();
    }
  }
  for (u32 t = tid; t < len / 8; t += S16_NT) s16f_reg8(S, 8 * t, false);
  __syncthreads();
}

// S[0..P) ascending, P a power of two >= 16, given that the words arrive grouped by their leading (bin) bits in
// ascending order and no group is larger than K (a power of two >= 16)
__device__ void s16_bitonic_bins(u64* __restrict__ S, u32 P, u32 K) {
  const u32 tid = threadIdx.x;
  for (u32 t = tid; t < P / 8; t += S16_NT) s16f_reg8(S, 8 * t, true);
  __syncthreads();
  const u32 Kc = min(K, P);
  for (u32 k = 16; k <= Kc; k <<= 1) s16f_phase(S, P, k);
  if (Kc < P) {
    s16f_phase(S, P, 2 * Kc);
    s16f_phase(S + Kc, P - 2 * Kc, 2 * Kc);
  }
}

// ---- the block as a stream of dense characters -------------------------------------------------------------------------
// Text uses few of the 256 byte values (28 in the word corpus), so 13 bits of raw prefix tell only 165 bins apart and
// a 64-bit sort word holds 7 characters.  Here every byte is first replaced by its rank among the bytes that occur in
// the block (order preserving, b = 1..8 bits) and the block is kept in shared memory ONLY as that bit stream, most
// significant bit first: the 64 bits at bit offset b*i are the first 64/b characters of suffix i as one big-endian
// number.  The same 13 bin bits now separate ~3 500 bins (2.6 characters at b = 5), batches fill the sort buffer
// almost exactly, a sort word holds 10-12 characters (a third of the ties), and comparisons / LCPs run a window of
// 64/b characters at a time.  Positions past the end read as character 0; every comparison checks the suffix ends.
// the 32 bits at bit offset sh (0..31) of the 64-bit number hi:lo.  Written out rather than __funnelshift_l: that
// intrinsic is a volatile asm statement, which pins every shift behind its own loads and keeps the compiler from
// overlapping the loads of neighbouring positions
__device__ __forceinline__ u32 s16_bits32(u32 hi, u32 lo, u32 sh) { return (u32)(((((u64)hi << 32) | lo) << sh) >> 32); }

struct S16Dense {
  const u32* D; u32 b, rcp, n;
  __device__ __forceinline__ u32 hi32(u32 i) const {           // first 32 bits of suffix i
    const u32 bit = b * i, j = bit >> 5, sh = bit & 31u;
    return s16_bits32(D[j], D[j + 1], sh);
  }
  __device__ __forceinline__ u64 win(u32 i) const {            // first 64 bits of suffix i
    const u32 bit = b * i, j = bit >> 5, sh = bit & 31u;
    const u32 w0 = D[j], w1 = D[j + 1], w2 = D[j + 2];
    return ((u64)s16_bits32(w0, w1, sh) << 32) | s16_bits32(w1, w2, sh);
  }
  __device__ __forceinline__ u32 code(u32 i) const { return hi32(i) >> (32u - b); }
  __device__ __forceinline__ u32 chars(u64 x) const {          // whole characters two windows share, x = their XOR
    const u32 w = (64u * rcp) >> 16;                           // 64 / b
    return x ? min(((u32)__clzll((long long)x) * rcp) >> 16, w) : w;
  }
};

// suffix a < suffix b, both known to share their first `from` characters
__device__ __forceinline__ bool s16d_less(const S16Dense& d, u32 a, u32 b, u32 from, u32* deep) {
  const u32 w = (64u * d.rcp) >> 16;
  u32 k = from;
  for (;;) {
    const u32 pa = a + k, pb = b + k;
    if (pa >= d.n || pb >= d.n) return pa >= d.n;               // the suffix that ends first is the smaller one
    const u64 wa = d.win(pa), wb = d.win(pb);
    const u32 L = d.chars(wa ^ wb);
    if (L >= d.n - max(pa, pb)) return a > b;                   // equal up to the end of the shorter one
    if (L < w) return wa < wb;                                  // the first difference lies in a character both have
    k += w;
    if (k > S16_MAXDEPTH) { *deep = 1; return a < b; }
  }
}

// common prefix length of suffixes a and b, capped
__device__ __forceinline__ u32 s16d_lcp(const S16Dense& d, u32 a, u32 b, u32 cap) {
  const u32 w = (64u * d.rcp) >> 16;
  const u32 lim = min(cap, d.n - max(a, b));
  u32 l = 0;
  while (l < lim) {
    const u32 L = d.chars(d.win(a + l) ^ d.win(b + l));
    l += L;
    if (L < w) break;
  }
  return min(l, lim);
}

// One block.  The raw bytes are staged in the (still unused) sort buffer.
__device__ void suffix_sort16d_load(const u8* __restrict__ gT, u32 n, Sort16Smem& sm) {
  const u32 tid = threadIdx.x;
  u8* __restrict__ R = reinterpret_cast<u8*>(sm.buf);
  const bool aligned = ((uintptr_t)gT & 15) == 0;
  const u32 nb = n & ~15u;
  if (aligned && nb) {
    if (tid == 0) { zq_mbar_expect_tx(&sm.bar, nb); zq_bulk_g2s(R, gT, nb, &sm.bar); }
  }
  for (u32 i = (aligned ? nb : 0) + tid; i < n; i += S16_NT) R[i] = gT[i];
  for (u32 b = tid; b < S16_BINS + 8; b += S16_NT) sm.bins[b] = 0;
  if (tid < 256) sm.present[tid] = 0;
  if (tid == 0) { sm.fallback = 0; sm.prev_idx = 0; }
  __syncthreads();
}

__device__ bool suffix_sort16d_rest(u32 n, u8* __restrict__ w, bool want_pk, Sort16Smem& sm) {
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* __restrict__ S = sm.buf;
  u32* __restrict__ Dw = reinterpret_cast<u32*>(sm.text);
  // 1. alphabet: which bytes occur, their ranks
  {
    const u8* __restrict__ R = reinterpret_cast<const u8*>(sm.buf);
    for (u32 i = tid; i < n; i += S16_NT) sm.present[R[i]] = 1;
    __syncthreads();
    if (warp == 0) {
      u32 c = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) c += sm.present[lane * 8 + q];
      u32 inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(ZQ_FULL, inc, o); if (lane >= (u32)o) inc += t; }
      u32 run = inc - c;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const u32 v = lane * 8 + q;
        sm.rank[v] = (u8)run;
        if (sm.present[v]) { sm.inv[run] = (u8)v; ++run; }
      }
      if (lane == 31) {
        const u32 k = inc, b = max(1u, (u32)zq_bitlen(k - 1));
        sm.kused = k; sm.cbits = b; sm.crcp = 65536u / b + 1u;
      }
    }
    __syncthreads();
    // 2. the packed stream: word j holds stream bits 32j .. 32j+31
    const u32 b = sm.cbits;
    const u32 nwords = (n * b + 31) / 32 + 4;
    for (u32 j = tid; j < nwords; j += S16_NT) {
      const u32 bit0 = 32 * j;
      u32 c = bit0 / b;                                    // first character with bits in this word
      u32 word = 0;
      for (u32 at = c * b; at < bit0 + 32; at += b, ++c) {
        const u32 code = c < n ? (u32)sm.rank[R[c]] : 0u;
        const int sh = (int)(bit0 + 32) - (int)(at + b);   // distance of the character's last bit from the word's
        word |= sh >= 0 ? code << sh : code >> (-sh);
      }
      Dw[j] = word;
    }
    __syncthreads();
  }
  S16Dense d; d.D = Dw; d.b = sm.cbits; d.rcp = sm.crcp; d.n = n;
  // 3. bins: the first 13 bits of every suffix
  for (u32 i = tid; i < n; i += S16_NT) atomicAdd(&sm.bins[d.hi32(i) >> 19], 1u);
  __syncthreads();
  {
    u32 loc[8], s = 0, mx = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { loc[q] = sm.bins[tid * 8 + q]; s += loc[q]; mx = max(mx, loc[q]); }
    u32 inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(ZQ_FULL, inc, o); if (lane >= (u32)o) inc += t; }
    mx = __reduce_max_sync(ZQ_FULL, mx);
    if (lane == 31) sm.wsum[warp] = inc;
    if (lane == 0 && mx > S16_BUF) sm.fallback = 1;
    __syncthreads();
    u32 pre = 0;
    for (u32 q = 0; q < warp; ++q) pre += sm.wsum[q];
    u32 run = pre + inc - s;
#pragma unroll
    for (int q = 0; q < 8; ++q) { sm.bins[tid * 8 + q] = run; run += loc[q]; }
    if (tid == S16_NT - 1) sm.bins[S16_BINS] = run;     // = n
    __syncthreads();
  }
  if (sm.fallback) return false;
  const u64 stride = zq_work_stride(n, 2);
  u16* __restrict__ o_sa = (u16*)w; u16* __restrict__ o_isa = (u16*)(w + stride);
  u32* cursor = reinterpret_cast<u32*>(sm.buf);
  for (u32 b = tid; b < S16_BINS; b += S16_NT) cursor[b] = sm.bins[b];
  __syncthreads();
  {
    u32 i = tid;
    for (; i + 3 * S16_NT < n; i += 4 * S16_NT) {         // four positions in flight: the atomics' latency overlaps
      const u32 b0 = d.hi32(i) >> 19, b1 = d.hi32(i + S16_NT) >> 19, b2 = d.hi32(i + 2 * S16_NT) >> 19, b3 = d.hi32(i + 3 * S16_NT) >> 19;
      const u32 r0 = atomicAdd(&cursor[b0], 1u), r1 = atomicAdd(&cursor[b1], 1u), r2 = atomicAdd(&cursor[b2], 1u), r3 = atomicAdd(&cursor[b3], 1u);
      o_sa[r0] = (u16)i; o_sa[r1] = (u16)(i + S16_NT); o_sa[r2] = (u16)(i + 2 * S16_NT); o_sa[r3] = (u16)(i + 3 * S16_NT);
    }
    for (; i < n; i += S16_NT) o_sa[atomicAdd(&cursor[d.hi32(i) >> 19], 1u)] = (u16)i;
  }
  __syncthreads();
  u16* __restrict__ o_lcp = (u16*)(w + 2 * stride);
  u8* __restrict__ o_bwt = w + 2 * stride + zq_work_stride(n, 2);
  u32* __restrict__ o_pk = (u32*)(w + zq_work_bytes(n, 2));
  u32 rowbase = 0;
  while (rowbase < n) {
    const u32 limit = rowbase + S16_BUF;
    {
      u32 best = 0, blo = S16_BINS, bhi = 0, bmx = 0;     // ... and the largest bin of the batch
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const u32 b = tid * 8 + q, st = sm.bins[b], e = sm.bins[b + 1];
        if (e <= limit) best = max(best, e);
        if (e > st && st >= rowbase && e <= limit) { blo = min(blo, b); bhi = max(bhi, b); bmx = max(bmx, e - st); }
      }
      best = __reduce_max_sync(ZQ_FULL, best); blo = __reduce_min_sync(ZQ_FULL, blo); bhi = __reduce_max_sync(ZQ_FULL, bhi);
      bmx = __reduce_max_sync(ZQ_FULL, bmx);
      if (lane == 0) { sm.wsum[warp] = best; sm.wlo[warp] = blo; sm.whi[warp] = bhi; sm.wmx[warp] = bmx; }
      __syncthreads();
      best = 0; blo = S16_BINS; bhi = 0; bmx = 0;
      for (u32 q = 0; q < S16_NT / 32; ++q) { best = max(best, sm.wsum[q]); blo = min(blo, sm.wlo[q]); bhi = max(bhi, sm.whi[q]); bmx = max(bmx, sm.wmx[q]); }
      __syncthreads();
      sm.pad0 = best; sm.blo = blo; sm.bhi = bhi; sm.bmx = bmx;
    }
    const u32 rowend = sm.pad0;
    // sort words: bin number inside the batch (nb bits) | the 48 - nb stream bits after the bin bits | index.  Equal
    // keys share 61 - nb bits = at least `from` whole characters.
    const u32 blo = sm.blo, nb = (u32)zq_bitlen(sm.bhi - sm.blo), from = ((61u - nb) * d.rcp) >> 16;
#pragma unroll 4
    for (u32 j = tid; j < rowend - rowbase; j += S16_NT) {
      const u32 i = o_sa[rowbase + j];
      const u64 p64 = d.win(i);
      const u64 rel = (u64)((u32)(p64 >> 51) - blo);
      S[j] = (((rel << (48 - nb)) | ((p64 << 13) >> (16 + nb))) << 16) | i;
    }
    const u32 m = rowend - rowbase;
    u32 P = 16;
    while (P < m) P <<= 1;
    for (u32 t = m + tid; t < P; t += S16_NT) S[t] = ~0ull;
    __syncthreads();
    {
      u32 K = 16;
      while (K < sm.bmx) K <<= 1;
      s16_bitonic_bins(S, P, K);
    }
    for (u32 c0 = warp * S16_TIE_CHUNK; c0 < m; c0 += (S16_NT / 32) * S16_TIE_CHUNK) {
      u32 hm[S16_TIE_CHUNK / 32], total = 0;
#pragma unroll
      for (u32 r = 0; r < S16_TIE_CHUNK / 32; ++r) {
        const u32 j = c0 + r * 32 + lane;
        bool head = false;
        if (j + 1 < m) {
          const u64 kj = S[j] >> 16;
          head = (j == 0 || (S[j - 1] >> 16) != kj) && (S[j + 1] >> 16) == kj;
        }
        hm[r] = __ballot_sync(ZQ_FULL, head);
        total += (u32)__popc(hm[r]);
      }
      for (u32 h = lane; h < total; h += 32) {
        u32 j = 0, hh = h;
#pragma unroll
        for (u32 r = 0; r < S16_TIE_CHUNK / 32; ++r) {
          const u32 pc = (u32)__popc(hm[r]);
          if (hh < pc) {
            u32 mk = hm[r];
            for (u32 q = 0; q < hh; ++q) mk &= mk - 1;
            j = c0 + r * 32 + (u32)(__ffs(mk) - 1);
            hh = 0xffffffffu;
          } else if (hh != 0xffffffffu) hh -= pc;
        }
        const u64 kj = S[j] >> 16;
        u32 e = j + 1;
        while (e < m && e - j <= S16_MAXGROUP && (S[e] >> 16) == kj) ++e;
        if (e - j > S16_MAXGROUP) { sm.fallback = 1; continue; }
        u32 deep = 0;
        for (u32 x = j + 1; x < e; ++x) {
          const u32 v = (u32)S[x] & 0xffffu;
          u32 y = x;
          while (y > j && s16d_less(d, v, (u32)S[y - 1] & 0xffffu, from, &deep)) { S[y] = S[y - 1]; --y; }
          S[y] = (kj << 16) | v;
        }
        if (deep) sm.fallback = 1;
      }
    }
    __syncthreads();
    if (sm.fallback) return false;
    const u32 prev_last = sm.prev_idx;
    for (u32 j = tid; j < m; j += S16_NT) {
      const u32 b = (u32)S[j] & 0xffffu, row = rowbase + j;
      u32 l = 0;
      if (row > 0) l = s16d_lcp(d, j ? ((u32)S[j - 1] & 0xffffu) : prev_last, b, ZQ_LCP_CAP);
      const u32 bw = b > 0 ? (u32)sm.inv[d.code(b - 1)] : 0u;
      o_sa[row] = (u16)b; o_lcp[row] = (u16)l; o_bwt[row] = (u8)bw; o_isa[b] = (u16)row;
      if (want_pk) o_pk[row] = LzsPack<u16>::make(b, l, bw);
    }
    __syncthreads();
    if (tid == 0) sm.prev_idx = (u32)S[m - 1] & 0xffffu;
    rowbase = rowend;
    __syncthreads();
  }
  return true;
}

// Persistent CTAs pull blocks from a counter.  need_old[t] = 1 for the blocks left to k_suffix_sort (too large for
// 16-bit indices, or not text-like enough for this method).
__global__ void __launch_bounds__(S16_NT, 1)
k_suffix_sort16(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const int* __restrict__ todo, int ntodo,
                u8* __restrict__ work_base, u32* __restrict__ need_old, u32* next_unit) {
  ZQ_DYN_SMEM(smem_raw);
  Sort16Smem& sm = *reinterpret_cast<Sort16Smem*>(smem_raw);
  __shared__ int s_t;
  if (threadIdx.x == 0) zq_mbar_init(&sm.bar, 1);
  __syncthreads();
  u32 uses = 0;
  for (;;) {
    if (threadIdx.x == 0) s_t = (int)atomicAdd(next_unit, 1u);
    zq_fence_async_smem();     // the sort buffer was written with ordinary stores; the next block's bulk copy lands in it
    __syncthreads();
    const int t = s_t;
    __syncthreads();
    if (t >= ntodo) break;
    const ZqUnit u = units[todo[t]];
    bool ok = u.idx16 != 0 && u.n > 0;
    if (ok) {
      const u8* gT = in_base + u.in_off;
      suffix_sort16d_load(gT, u.n, sm);
      if ((((uintptr_t)gT & 15) == 0) && (u.n & ~15u)) { zq_mbar_wait(&sm.bar, uses & 1u); ++uses; }
      __syncthreads();
      ok = suffix_sort16d_rest(u.n, work_base + u.work_off, u.want_pk != 0, sm);
    }
    __syncthreads();
    if (threadIdx.x == 0) need_old[t] = ok || u.n == 0 ? 0u : 1u;
  }
}

}  // namespace zqdev
