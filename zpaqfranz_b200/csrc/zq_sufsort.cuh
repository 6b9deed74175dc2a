// zq_sufsort.cuh -- suffix array + inverse + capped LCP of one block per CTA.
//
// Replaces libzpaq::divsufsort (Z:17721-19147, entry Z:19121) as called from LZBuffer's constructor
// (Z:19372).  The suffix array of a string is unique (unsigned bytes, a shorter suffix sorts before
// any longer suffix it prefixes), so any correct construction is bit-exact with divsufsort's output.
//
// Algorithm (GPU-first, not divsufsort's induced sorting): prefix doubling restricted to the still
// ambiguous suffixes.
//   1. LSD radix sort of all suffixes by their first 4 bytes.
//   2. rank[i] = start index of i's group; suffixes in groups of size 1 are final.
//   3. Round h = 4, 8, 16, ...: gather the m not-yet-unique suffixes (in SA order), key them by
//      (rank[i], rank[i+h]+1 or 0 past the end), radix sort the compacted list, write it back into
//      the same SA slots, split groups where keys differ, drop the now-unique ones.  Stops when m = 0.
//   The final rank[] is the inverse suffix array the LZ77 parser needs (LZBuffer's isa, Z:19405).
//   4. bwt[x] = T[sa[x]-1]; lcp[x] = min(LCP(sa[x-1], sa[x]), ZQ_LCP_CAP): lets the parser get match lengths of SA
//      neighbours by a running minimum instead of byte compares (exact for everything <= 255, which
//      is all the parser's scan needs, Z:19424).
// One CTA of 1024 threads per block; all arrays live in global memory (L1/L2 resident for 64 KiB
// blocks: 1 CTA per SM leaves ~150 KB of L1), radix ranking state in shared memory.
#pragma once
#include "zq_common.cuh"

namespace zqdev {

#ifdef ZQ_SORT_PROF
// phase cycle counters (thread 0 of every CTA): 0 initial keys+sort, 1 first rank, 2 round keys, 3 round sort,
// 4 round rerank, 5 outputs (lcp/bwt/isa), 6 number of rounds, 7 sum of m over rounds
__device__ unsigned long long g_sort_prof[8];
#define ZQ_PROF_T0 long long prof_t = clock64();
#define ZQ_PROF(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_sort_prof[k], (unsigned long long)(now_ - prof_t)); prof_t = now_; } } while (0)
#define ZQ_PROF_ADD(k, v) do { if (threadIdx.x == 0) atomicAdd(&g_sort_prof[k], (unsigned long long)(v)); } while (0)
#else
#define ZQ_PROF_T0
#define ZQ_PROF(k)
#define ZQ_PROF_ADD(k, v)
#endif

#ifndef ZQ_SORT_ITEMS
#define ZQ_SORT_ITEMS 8
#endif
constexpr int SORT_ITEMS = ZQ_SORT_ITEMS;
constexpr int RANK_ITEMS = 4;            // consecutive elements per thread in the ranking loops
constexpr int SORT_MAXD = 256;           // digits per radix pass (8 bits)
constexpr u32 ZQ_LCP_CAP = 256;

// packed SA row for the LZ77 scan kernels: one shared-memory load per scan step
template <typename IdxT> struct LzsPack;
template <> struct LzsPack<u16> {
  typedef u32 T;
  static __device__ __forceinline__ u32 make(u32 sa, u32 lcp, u32 bw) { return sa | (min(lcp, 255u) << 16) | (bw << 24); }
  static __device__ __forceinline__ u32 sa(u32 w) { return w & 0xffffu; }
#ifdef LZS_SHIFT_FIELDS
  static __device__ __forceinline__ u32 lcp(u32 w) { return (w >> 16) & 255u; }
#else   // one byte-permute instead of shift + mask: the scan kernels are bound by the integer pipe
  static __device__ __forceinline__ u32 lcp(u32 w) { return __byte_perm(w, 0u, 0x4442); }
#endif
  static __device__ __forceinline__ u32 bwt(u32 w) { return w >> 24; }
};
template <> struct LzsPack<u32> {
  typedef u64 T;
  static __device__ __forceinline__ u64 make(u32 sa, u32 lcp, u32 bw) { return sa | ((u64)min(lcp, 255u) << 32) | ((u64)bw << 40); }
  static __device__ __forceinline__ u32 sa(u64 w) { return (u32)w; }
  static __device__ __forceinline__ u32 lcp(u64 w) { return (u32)(w >> 32) & 255u; }
  static __device__ __forceinline__ u32 bwt(u64 w) { return (u32)(w >> 40) & 255u; }
};

template <int NT>
struct SortSmem {
  u32 wcount[(NT / 32) * SORT_MAXD];  // per-warp digit counters, then per-warp scatter bases
  u32 hist[SORT_MAXD];
  u32 base[SORT_MAXD];
  u32 wsum[32];
  u32 wsum2[32];
  u32 misc[4];
  u32 hist_next[SORT_MAXD]; // digit histogram of the NEXT pass, gathered while this pass writes out
  u32 tot[SORT_MAXD];      // digit totals of the current tile
  u32 tstart[SORT_MAXD];   // exclusive scan of tot: where a digit's run starts inside the staged tile
  u64 stage_k[NT * SORT_ITEMS];   // the tile in digit order, so that the global write-out is coalesced runs
  u32 stage_v[NT * SORT_ITEMS];
};

// Per-CTA scratch (global memory), sized for the largest block of the wave.
struct SortScratch {
  u64* kA; u64* kB;   // keys
  u32* vA; u32* vB;   // suffix indices
  u32* pA; u32* pB;   // SA slots of the compacted (unsorted) suffixes
  u32* sa; u32* rank; // working suffix array and rank (= inverse SA at the end)
};

template <int NT>
__device__ __forceinline__ u32 block_scan_incl_add(u32 v, SortSmem<NT>& sm, u32& total) {
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(ZQ_FULL, v, o); if (lane >= (u32)o) v += t; }
  if (lane == 31) sm.wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    u32 w = lane < NT / 32 ? sm.wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(ZQ_FULL, w, o); if (lane >= (u32)o) w += t; }
    sm.wsum[lane] = w;
  }
  __syncthreads();
  const u32 pre = warp ? sm.wsum[warp - 1] : 0;
  total = sm.wsum[31];
  __syncthreads();
  return v + pre;
}

template <int NT>
__device__ __forceinline__ u32 block_scan_incl_max(u32 v, SortSmem<NT>& sm, u32& total) {
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(ZQ_FULL, v, o); if (lane >= (u32)o) v = max(v, t); }
  if (lane == 31) sm.wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    u32 w = lane < NT / 32 ? sm.wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(ZQ_FULL, w, o); if (lane >= (u32)o) w = max(w, t); }
    sm.wsum[lane] = w;
  }
  __syncthreads();
  const u32 pre = warp ? sm.wsum[warp - 1] : 0;
  total = sm.wsum[31];
  __syncthreads();
  return max(v, pre);
}

// Both scans of the ranking loops in one pass: inclusive max of `vmax` and inclusive sum of `vadd` over the
// CTA's threads in thread order; totals of the whole CTA in tmax / tadd.
template <int NT>
__device__ __forceinline__ void block_scan_max_add(u32& vmax, u32& vadd, SortSmem<NT>& sm, u32& tmax, u32& tadd) {
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 a = __shfl_up_sync(ZQ_FULL, vmax, o), b = __shfl_up_sync(ZQ_FULL, vadd, o);
    if (lane >= (u32)o) { vmax = max(vmax, a); vadd += b; }
  }
  if (lane == 31) { sm.wsum[warp] = vmax; sm.wsum2[warp] = vadd; }
  __syncthreads();
  if (warp == 0) {
    u32 a = lane < NT / 32 ? sm.wsum[lane] : 0, b = lane < NT / 32 ? sm.wsum2[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 ta = __shfl_up_sync(ZQ_FULL, a, o), tb = __shfl_up_sync(ZQ_FULL, b, o);
      if (lane >= (u32)o) { a = max(a, ta); b += tb; }
    }
    sm.wsum[lane] = a; sm.wsum2[lane] = b;
  }
  __syncthreads();
  if (warp) { vmax = max(vmax, sm.wsum[warp - 1]); vadd += sm.wsum2[warp - 1]; }
  tmax = sm.wsum[31]; tadd = sm.wsum2[31];
  __syncthreads();
}

// lanes of the warp holding the same 8-bit digit (among active lanes): 8 ballots.  (match.any is
// emulated by a loop over the distinct values on this architecture -- up to 32 rounds per call.)
__device__ __forceinline__ u32 warp_peers8(u32 d, bool act) {
  u32 peers = __ballot_sync(ZQ_FULL, act);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const u32 bal = __ballot_sync(ZQ_FULL, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

// One stable LSD pass on digit (key >> shift) & 255 over m (key,value) pairs. Returns false (and
// moves nothing) when every key has the same digit.
template <int NT>
// have_hist: sm.hist_next already holds this pass's histogram (counted by the previous pass);
// next_shift >= 0: count the next pass's digits while writing out.
__device__ bool radix_pass(const u64* __restrict__ kin, const u32* __restrict__ vin,
                           u64* __restrict__ kout, u32* __restrict__ vout, u32 m, int shift, SortSmem<NT>& sm,
                           bool have_hist = false, int next_shift = -1) {
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (u32 d = tid; d < SORT_MAXD; d += NT) sm.hist[d] = have_hist ? sm.hist_next[d] : 0u;
  if (tid == 0) sm.misc[0] = 0;
  __syncthreads();
  if (!have_hist) {
    for (u32 b = 0; b < m; b += NT) {
      const u32 idx = b + tid;
      const bool act = idx < m;
      const u32 d = act ? (u32)(kin[idx] >> shift) & 255u : 0u;
      const u32 peers = warp_peers8(d, act);
      if (act && lane == (u32)(__ffs(peers) - 1)) atomicAdd(&sm.hist[d], __popc(peers));
    }
  }
  for (u32 d = tid; d < SORT_MAXD; d += NT) sm.hist_next[d] = 0;
  __syncthreads();
  for (u32 d = tid; d < SORT_MAXD; d += NT) if (sm.hist[d] == m) sm.misc[0] = 1;
  __syncthreads();
  if (sm.misc[0]) return false;
  if (warp == 0) {  // exclusive scan of the 256-bin histogram: 8 bins per lane
    u32 loc[8], s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { loc[q] = sm.hist[lane * 8 + q]; s += loc[q]; }
    u32 inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(ZQ_FULL, inc, o); if (lane >= (u32)o) inc += t; }
    u32 run = inc - s;
#pragma unroll
    for (int q = 0; q < 8; ++q) { sm.base[lane * 8 + q] = run; run += loc[q]; }
  }
  __syncthreads();
  for (u32 tile = 0; tile < m; tile += (NT * SORT_ITEMS)) {
    for (u32 d = lane; d < SORT_MAXD; d += 32) sm.wcount[warp * SORT_MAXD + d] = 0;
    __syncwarp();
    u64 k[SORT_ITEMS]; u32 v[SORT_ITEMS]; u32 r[SORT_ITEMS];
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
      const u32 idx = tile + warp * (SORT_ITEMS * 32) + it * 32 + lane;
      const bool act = idx < m;
      u32 d = 0u;
      if (act) { k[it] = kin[idx]; v[it] = vin[idx]; d = (u32)(k[it] >> shift) & 255u; }
      const u32 peers = warp_peers8(d, act);
      u32 cnt = 0;
      if (act) { cnt = sm.wcount[warp * SORT_MAXD + d]; r[it] = (d << 16) | (cnt + __popc(peers & lanemask_lt())); }
      else r[it] = 0xffffffffu;
      __syncwarp();
      if (act && lane == (u32)(__ffs(peers) - 1)) sm.wcount[warp * SORT_MAXD + d] = cnt + __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    for (u32 d = tid; d < SORT_MAXD; d += NT) {   // per digit: offsets of each warp's share inside the tile
      u32 run = 0;
#pragma unroll 8
      for (int w = 0; w < NT / 32; ++w) { const u32 c = sm.wcount[w * SORT_MAXD + d]; sm.wcount[w * SORT_MAXD + d] = run; run += c; }
      sm.tot[d] = run;
    }
    __syncthreads();
    if (warp == 0) {
      u32 loc[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { loc[q] = sm.tot[lane * 8 + q]; sum += loc[q]; }
      u32 inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(ZQ_FULL, inc, o); if (lane >= (u32)o) inc += t; }
      u32 run = inc - sum;
#pragma unroll
      for (int q = 0; q < 8; ++q) { sm.tstart[lane * 8 + q] = run; run += loc[q]; }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
      if (r[it] != 0xffffffffu) {
        const u32 d = r[it] >> 16;
        const u32 lp = sm.tstart[d] + sm.wcount[warp * SORT_MAXD + d] + (r[it] & 0xffffu);
        sm.stage_k[lp] = k[it]; sm.stage_v[lp] = v[it];
      }
    }
    __syncthreads();
    const u32 tcount = min((u32)(NT * SORT_ITEMS), m - tile);
    for (u32 sidx = tid; sidx < tcount; sidx += NT) {
      const u64 kk = sm.stage_k[sidx];
      const u32 d = (u32)(kk >> shift) & 255u;
      const u32 gp = sm.base[d] + (sidx - sm.tstart[d]);
      kout[gp] = kk; vout[gp] = sm.stage_v[sidx];
      if (next_shift >= 0) atomicAdd(&sm.hist_next[(u32)(kk >> next_shift) & 255u], 1u);
    }
    __syncthreads();
    for (u32 d = tid; d < SORT_MAXD; d += NT) sm.base[d] += sm.tot[d];
    __syncthreads();
  }
  return true;
}

// LSD sort on key bits [lo, hi); ping-pongs between (kA,vA) and (kB,vB); returns which holds the result
template <int NT>
__device__ int radix_sort_bits(u64*& kA, u32*& vA, u64*& kB, u32*& vB, u32 m, int lo, int hi, SortSmem<NT>& sm) {
  bool have = false;
  for (int s = lo; s < hi; s += 8) {
    const int nxt = s + 8 < hi ? s + 8 : -1;
    const bool moved = radix_pass<NT>(kA, vA, kB, vB, m, s, sm, have, nxt);
    if (moved) { u64* tk = kA; kA = kB; kB = tk; u32* tv = vA; vA = vB; vB = tv; }
    have = moved && nxt >= 0;     // a skipped pass (all digits equal) counted nothing for its successor
    __syncthreads();
  }
  return 0;
}

// ---- doubling rounds: tile sort ------------------------------------------------------------------------
// The compacted list of a round is already grouped by rank (it is in SA order), so only the order INSIDE
// each group is open.  Instead of five global radix passes over (rank, rank2) the list is cut, at group
// boundaries, into tiles of <= NT*SORT_ITEMS elements that are sorted in shared memory by one 64-bit word
// (rank | rank2 | suffix) with a bitonic network: comparators whose partners lie inside one 256-element
// region are run by the warp owning the region (warp-level sync only), the few wider ones block-wide.
constexpr u32 TILE_REGION = 256;

__device__ __forceinline__ void bitonic_cmpx(u64* __restrict__ S, u32 i, u32 j, u32 k) {
  const u64 a = S[i], b = S[i + j];
  const bool up = (i & k) == 0;
  if ((a > b) == up) { S[i] = b; S[i + j] = a; }
}

template <int NT>
__device__ void tile_sort(u64* __restrict__ kA, u32* __restrict__ vA, u32 s, u32 cnt, int bits2, int bitsv, SortSmem<NT>& sm) {
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* __restrict__ S = sm.stage_k;
  u32 P = 64;
  while (P < cnt) P <<= 1;
  const u64 vmask = (1ull << bitsv) - 1, k2mask = (1ull << bits2) - 1;
  for (u32 t = tid; t < P; t += NT) {
    u64 key = ~0ull;
    if (t < cnt) { const u64 k = kA[s + t]; key = ((((k >> 32) << bits2) | (k & 0xffffffffull)) << bitsv) | vA[s + t]; }
    S[t] = key;
  }
  __syncthreads();
  for (u32 k = 2; k <= P; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      if (j >= TILE_REGION) {
        for (u32 t = tid; t < P / 2; t += NT) bitonic_cmpx(S, 2 * t - (t & (j - 1)), j, k);
        __syncthreads();
      } else {
        for (u32 base = warp * TILE_REGION; base < P; base += (NT / 32) * TILE_REGION) {
          const u32 ncmp = min(TILE_REGION, P - base) / 2;
          for (u32 t = lane; t < ncmp; t += 32) bitonic_cmpx(S, base + 2 * t - (t & (j - 1)), j, k);
        }
        __syncwarp();
        if (j == 1 && k < P && k >= TILE_REGION) __syncthreads();   // the next merge starts with block-wide comparators
      }
    }
  }
  __syncthreads();
  for (u32 t = tid; t < cnt; t += NT) {
    const u64 key = S[t];
    vA[s + t] = (u32)(key & vmask);
    kA[s + t] = ((key >> (bitsv + bits2)) << 32) | ((key >> bitsv) & k2mask);
  }
  __syncthreads();
}

// Sorts the round's list kA/vA[0..m) (grouped by the high word) inside its groups by the low word.
template <int NT>
__device__ void sort_round_tiles(u64* kA, u32* vA, u64* kB, u32* vB, u32 m, int bits2, int bitsv, SortSmem<NT>& sm) {
  const u32 tid = threadIdx.x, lane = tid & 31;
  constexpr u32 TILE = NT * SORT_ITEMS;
  u32 s = 0;
  while (s < m) {
    if (tid < 32) {   // warp 0 places the tile's end on a group boundary
      u32 e = min(s + TILE, m), big_end = 0;
      if (e < m) {
        const u32 g = (u32)(kA[e] >> 32);
        if ((u32)(kA[e - 1] >> 32) == g) {
          const u32 back = e - 1 - s;   // how far a lane may look back
          const bool same = lane <= back && (u32)(kA[e - 1 - min(lane, back)] >> 32) == g;
          const u32 differ = ~__ballot_sync(ZQ_FULL, same);
          u32 start;
          if (differ) start = e - (u32)(__ffs(differ) - 1);
          else {   // the group reaches further back than a warp sees: binary search its first element
            u32 lo = s, hi = e - 32;
            while (lo < hi) { const u32 mid = (lo + hi) >> 1; if ((u32)(kA[mid] >> 32) >= g) hi = mid; else lo = mid + 1; }
            start = lo;
          }
          if (start > s) e = start;
          else {   // one group larger than a tile: find where it ends
            u32 a = e, b = m;
            while (a < b) { const u32 mid = (a + b) >> 1; if ((u32)(kA[mid] >> 32) > g) b = mid; else a = mid + 1; }
            big_end = a;
          }
        }
      }
      if (lane == 0) { sm.misc[1] = e; sm.misc[2] = big_end; }
    }
    __syncthreads();
    const u32 e = sm.misc[1], big_end = sm.misc[2];
    __syncthreads();
    if (big_end) {   // all elements share the rank: LSD passes on the second key only, then back into place
      const u32 len = big_end - s;
      u64 *a = kA + s, *b = kB + s; u32 *va = vA + s, *vb = vB + s;
      radix_sort_bits<NT>(a, va, b, vb, len, 0, bits2, sm);
      if (a != kA + s) {
        for (u32 t = tid; t < len; t += NT) { kA[s + t] = a[t]; vA[s + t] = va[t]; }
        __syncthreads();
      }
      s = big_end;
    } else {
      tile_sort<NT>(kA, vA, s, e - s, bits2, bitsv, sm);
      s = e;
    }
  }
}

// 4 bytes at any address (little endian) from the two aligned words around it; reads up to 7 bytes past p.
__device__ __forceinline__ u32 ld32_unaligned(const u8* p) {
  const uintptr_t q = (uintptr_t)p;
  const u32* w = (const u32*)(q & ~(uintptr_t)3);
  return __funnelshift_r(w[0], w[1], (u32)(q & 3) * 8);
}

// Builds the suffix array of T[0..n) in scratch (sa, rank = inverse SA) and writes sa | isa | lcp to
// the unit's work region `w` with index width 2 (idx16) or 4 bytes.
template <int NT>
__device__ void suffix_sort_block(const u8* __restrict__ T, u32 n, u8* __restrict__ w, bool idx16,
                                  SortScratch sc, SortSmem<NT>& sm, bool want_pk) {
  u32* __restrict__ sa = sc.sa; u32* __restrict__ rank = sc.rank;
  const u32 tid = threadIdx.x;
  if (n == 0) return;
  u64 *kA = sc.kA, *kB = sc.kB; u32 *vA = sc.vA, *vB = sc.vB; u32 *pA = sc.pA, *pB = sc.pB;
  ZQ_PROF_T0
  const u32 nshort = n < 3 ? n : 3;  // suffixes shorter than the 4-byte key: n-1, n-2, n-3
  // 1. keys = first 4 bytes (zero padded). Input order puts the short suffixes first, shortest
  //    first, so the stable sort leaves them ahead of equal-keyed longer suffixes (implicit
  //    terminator is smaller than any byte).
  for (u32 j = tid; j < n; j += NT) {
    const u32 i = j < nshort ? n - 1 - j : j - nshort;
    u32 key = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) key = (key << 8) | (i + b < n ? (u32)T[i + b] : 0u);
    kA[j] = key; vA[j] = i;
  }
  __syncthreads();
  radix_sort_bits<NT>(kA, vA, kB, vB, n, 0, 32, sm);
  ZQ_PROF(0);
  // 2. group heads -> rank, sa; compact the ambiguous positions
  u32 m = 0;
  {
    u32 carry_max = 0, carry_cnt = 0;
    for (u32 b = 0; b < n; b += NT * RANK_ITEMS) {
      const u32 x0 = b + tid * RANK_ITEMS;
      // keys/suffixes of x0-1 .. x0+RANK_ITEMS (the neighbours decide head / nexthead)
      u64 kk[RANK_ITEMS + 2]; u32 vv[RANK_ITEMS + 2];
#pragma unroll
      for (int q = 0; q < RANK_ITEMS + 2; ++q) {
        const u32 x = x0 + q - 1;
        const bool ok = x0 + q >= 1 && x < n;
        kk[q] = ok ? kA[x] : 0; vv[q] = ok ? vA[x] : 0;
      }
      bool headq[RANK_ITEMS + 1];
#pragma unroll
      for (int q = 0; q <= RANK_ITEMS; ++q) {   // headq[q]: element x0+q starts a group (elements past the end do)
        const u32 x = x0 + q;
        headq[q] = x >= n || x == 0 || kk[q] != kk[q + 1] || vv[q + 1] + 4 > n || vv[q] + 4 > n;
      }
      u32 lmax = 0, ladd = 0; u32 gq[RANK_ITEMS]; u32 aq[RANK_ITEMS];
#pragma unroll
      for (int q = 0; q < RANK_ITEMS; ++q) {
        const u32 x = x0 + q;
        const bool act = x < n;
        if (act && headq[q]) lmax = max(lmax, x);
        gq[q] = lmax;
        const bool amb = act && !(headq[q] && headq[q + 1]);
        ladd += amb ? 1u : 0u;
        aq[q] = amb ? ladd : 0u;
      }
      u32 smax = lmax, sadd = ladd, tmax, tadd;
      block_scan_max_add<NT>(smax, sadd, sm, tmax, tadd);
      // exclusive prefixes of this thread
      const u32 up = __shfl_up_sync(ZQ_FULL, smax, 1);
      const u32 exmax = max(carry_max, (tid & 31) ? up : ((tid >> 5) ? sm.wsum[(tid >> 5) - 1] : 0u));
      const u32 exadd = carry_cnt + sadd - ladd;
#pragma unroll
      for (int q = 0; q < RANK_ITEMS; ++q) {
        const u32 x = x0 + q;
        if (x < n) {
          const u32 i = vv[q + 1];
          sa[x] = i;
          const u32 g = max(gq[q], exmax);
          rank[i] = g;
          if (aq[q]) { const u32 c = exadd + aq[q] - 1; pA[c] = x; kB[c] = (u64)g << 32; vB[c] = i; }   // next round's list
        }
      }
      carry_max = max(carry_max, tmax);
      carry_cnt += tadd;
      __syncthreads();
    }
    m = carry_cnt;
    { u64* tk = kA; kA = kB; kB = tk; u32* tv = vA; vA = vB; vB = tv; }
  }
  __syncthreads();
  ZQ_PROF(1);
  // 3. doubling rounds over the ambiguous suffixes only
  const int bits_rank = zq_bitlen(n - 1), bits_key2 = zq_bitlen(n);
  for (u32 h = 4; m > 0; h <<= 1) {
    // the list (kA = rank << 32, vA = suffix, pA = SA slot) was written by the previous ranking pass: only the
    // second key is gathered here, four independent elements per thread
    for (u32 j0 = tid; j0 < m; j0 += NT * 4) {
      u32 ii[4], k2[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { const u32 j = j0 + q * NT; ii[q] = j < m ? vA[j] : 0xffffffffu; }
#pragma unroll
      for (int q = 0; q < 4; ++q) k2[q] = (ii[q] != 0xffffffffu && ii[q] + h < n) ? rank[ii[q] + h] + 1u : 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) { const u32 j = j0 + q * NT; if (j < m) kA[j] |= k2[q]; }
    }
    __syncthreads();
    ZQ_PROF(2); ZQ_PROF_ADD(6, 1); ZQ_PROF_ADD(7, m);
    if (2 * bits_rank + bits_key2 <= 64) sort_round_tiles<NT>(kA, vA, kB, vB, m, bits_key2, bits_rank, sm);
    else {
      radix_sort_bits<NT>(kA, vA, kB, vB, m, 0, bits_key2, sm);
      radix_sort_bits<NT>(kA, vA, kB, vB, m, 32, 32 + bits_rank, sm);
    }
    ZQ_PROF(3);
    u32 carry_max = 0, carry_cnt = 0;
    for (u32 b = 0; b < m; b += NT * RANK_ITEMS) {
      const u32 j0 = b + tid * RANK_ITEMS;
      u64 kk[RANK_ITEMS + 2]; u32 ii[RANK_ITEMS], xx[RANK_ITEMS];
#pragma unroll
      for (int q = 0; q < RANK_ITEMS + 2; ++q) {
        const u32 j = j0 + q - 1;
        kk[q] = (j0 + q >= 1 && j < m) ? kA[j] : 0;
      }
#pragma unroll
      for (int q = 0; q < RANK_ITEMS; ++q) {
        const bool act = j0 + q < m;
        ii[q] = act ? vA[j0 + q] : 0; xx[q] = act ? pA[j0 + q] : 0;
      }
      bool headq[RANK_ITEMS + 1];
#pragma unroll
      for (int q = 0; q <= RANK_ITEMS; ++q) { const u32 j = j0 + q; headq[q] = j >= m || j == 0 || kk[q] != kk[q + 1]; }
      u32 lmax = 0, ladd = 0; u32 gq[RANK_ITEMS]; u32 aq[RANK_ITEMS];
#pragma unroll
      for (int q = 0; q < RANK_ITEMS; ++q) {
        const bool act = j0 + q < m;
        if (act && headq[q]) lmax = max(lmax, xx[q]);
        gq[q] = lmax;
        const bool amb = act && !(headq[q] && headq[q + 1]);
        ladd += amb ? 1u : 0u;
        aq[q] = amb ? ladd : 0u;
      }
      u32 smax = lmax, sadd = ladd, tmax, tadd;
      block_scan_max_add<NT>(smax, sadd, sm, tmax, tadd);
      const u32 up = __shfl_up_sync(ZQ_FULL, smax, 1);
      const u32 exmax = max(carry_max, (tid & 31) ? up : ((tid >> 5) ? sm.wsum[(tid >> 5) - 1] : 0u));
      const u32 exadd = carry_cnt + sadd - ladd;
#pragma unroll
      for (int q = 0; q < RANK_ITEMS; ++q) {
        if (j0 + q < m) {   // keys were materialised before: safe to update rank in place
          sa[xx[q]] = ii[q];
          const u32 g = max(gq[q], exmax);
          rank[ii[q]] = g;
          if (aq[q]) { const u32 c = exadd + aq[q] - 1; pB[c] = xx[q]; kB[c] = (u64)g << 32; vB[c] = ii[q]; }
        }
      }
      carry_max = max(carry_max, tmax);
      carry_cnt += tadd;
      __syncthreads();
    }
    m = carry_cnt;
    { u32* t = pA; pA = pB; pB = t; u64* tk = kA; kA = kB; kB = tk; u32* tv = vA; vA = vB; vB = tv; }
    __syncthreads();
    ZQ_PROF(4);
  }
  // 4. outputs: sa, isa (= rank) and the capped LCP of SA neighbours
  const u64 stride = zq_work_stride(n, idx16 ? 2 : 4);
  u16* __restrict__ lcp = (u16*)(w + 2 * stride);
  u8* __restrict__ bwt = w + 2 * stride + zq_work_stride(n, 2);
  for (u32 x = tid; x < n; x += NT) {
    u32 l = 0;
    const u32 b = sa[x];
    if (x > 0) {
      const u32 a = sa[x - 1];
      const u32 hi = max(a, b);
      const u32 lim = min(ZQ_LCP_CAP, n - hi);
      // four bytes per step while both (re-aligned) word pairs stay inside the block, then bytewise
      bool done = false;
      while (l + 4 <= lim && hi + l + 8 <= n) {
        const u32 d = ld32_unaligned(T + a + l) ^ ld32_unaligned(T + b + l);
        if (d) { l += (u32)(__ffs(d) - 1) >> 3; done = true; break; }
        l += 4;
      }
      if (!done) while (l < lim && T[a + l] == T[b + l]) ++l;
    }
    lcp[x] = (u16)l;
    bwt[x] = b > 0 ? T[b - 1] : (u8)0;
    const u32 bw = b > 0 ? (u32)T[b - 1] : 0u;
    if (idx16) {
      ((u16*)w)[x] = (u16)b; ((u16*)(w + stride))[x] = (u16)rank[x];
      if (want_pk) ((u32*)(w + zq_work_bytes(n, 2)))[x] = LzsPack<u16>::make(b, l, bw);
    } else {
      ((u32*)w)[x] = b; ((u32*)(w + stride))[x] = rank[x];
      if (want_pk) ((u64*)(w + zq_work_bytes(n, 4)))[x] = LzsPack<u32>::make(b, l, bw);
    }
  }
  ZQ_PROF(5);
}

// Grid-stride over the units of a wave that need a suffix array; scratch is per CTA.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_suffix_sort(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const int* __restrict__ todo, int ntodo,
              u8* __restrict__ work_base, u64* kbuf, u32* vbuf, u64 scratch_elems, const u32* __restrict__ only = nullptr) {
  ZQ_DYN_SMEM(smem_raw);
  SortSmem<NT>& sm = *reinterpret_cast<SortSmem<NT>*>(smem_raw);
  SortScratch sc;
  sc.kA = kbuf + (u64)blockIdx.x * 2 * scratch_elems; sc.kB = sc.kA + scratch_elems;
  u32* vb = vbuf + (u64)blockIdx.x * 6 * scratch_elems;
  sc.vA = vb; sc.vB = vb + scratch_elems; sc.pA = vb + 2 * scratch_elems; sc.pB = vb + 3 * scratch_elems;
  sc.sa = vb + 4 * scratch_elems; sc.rank = vb + 5 * scratch_elems;
  for (int t = blockIdx.x; t < ntodo; t += gridDim.x) {
    if (only && !only[t]) continue;     // already sorted by k_suffix_sort16 (zq_sufsort16.cuh)
    const ZqUnit u = units[todo[t]];
    suffix_sort_block<NT>(in_base + u.in_off, u.n, work_base + u.work_off, u.idx16 != 0, sc, sm, u.want_pk != 0);
    __syncthreads();
  }
}

}  // namespace zqdev
