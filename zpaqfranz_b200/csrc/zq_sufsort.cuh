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

constexpr int SORT_ITEMS = 8;
constexpr int SORT_MAXD = 256;           // digits per radix pass (8 bits)
constexpr u32 ZQ_LCP_CAP = 256;

template <int NT>
struct SortSmem {
  u32 wcount[(NT / 32) * SORT_MAXD];  // per-warp digit counters, then per-warp scatter bases
  u32 hist[SORT_MAXD];
  u32 base[SORT_MAXD];
  u32 wsum[32];
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

// Builds the suffix array of T[0..n) in scratch (sa, rank = inverse SA) and writes sa | isa | lcp to
// the unit's work region `w` with index width 2 (idx16) or 4 bytes.
template <int NT>
__device__ void suffix_sort_block(const u8* __restrict__ T, u32 n, u8* __restrict__ w, bool idx16,
                                  SortScratch sc, SortSmem<NT>& sm) {
  u32* __restrict__ sa = sc.sa; u32* __restrict__ rank = sc.rank;
  const u32 tid = threadIdx.x;
  if (n == 0) return;
  u64 *kA = sc.kA, *kB = sc.kB; u32 *vA = sc.vA, *vB = sc.vB; u32 *pA = sc.pA, *pB = sc.pB;
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
  // 2. group heads -> rank, sa; compact the ambiguous positions
  u32 m = 0;
  {
    u32 carry_max = 0, carry_cnt = 0;
    for (u32 b = 0; b < n; b += NT) {
      const u32 x = b + tid;
      const bool act = x < n;
      u32 i = 0; bool head = false, nexthead = true;
      if (act) {
        i = vA[x];
        const u64 kx = kA[x];
        head = x == 0 || kA[x - 1] != kx || i + 4 > n || vA[x - 1] + 4 > n;
        if (x + 1 < n) nexthead = kA[x + 1] != kx || i + 4 > n || vA[x + 1] + 4 > n;
      }
      u32 tot;
      u32 g = block_scan_incl_max<NT>(head ? x : 0u, sm, tot);
      g = max(g, carry_max);
      carry_max = max(carry_max, tot);
      const bool amb = act && !(head && nexthead);
      u32 tot2;
      const u32 inc = block_scan_incl_add<NT>(amb ? 1u : 0u, sm, tot2);
      if (act) { sa[x] = i; rank[i] = g; }
      if (amb) pA[carry_cnt + inc - 1] = x;
      carry_cnt += tot2;
    }
    m = carry_cnt;
  }
  __syncthreads();
  // 3. doubling rounds over the ambiguous suffixes only
  const int bits_rank = zq_bitlen(n - 1), bits_key2 = zq_bitlen(n);
  for (u32 h = 4; m > 0; h <<= 1) {
    for (u32 j = tid; j < m; j += NT) {
      const u32 i = sa[pA[j]];
      const u32 k2 = i + h < n ? rank[i + h] + 1u : 0u;
      kA[j] = ((u64)rank[i] << 32) | k2;
      vA[j] = i;
    }
    __syncthreads();
    radix_sort_bits<NT>(kA, vA, kB, vB, m, 0, bits_key2, sm);
    radix_sort_bits<NT>(kA, vA, kB, vB, m, 32, 32 + bits_rank, sm);
    u32 carry_max = 0, carry_cnt = 0;
    for (u32 b = 0; b < m; b += NT) {
      const u32 j = b + tid;
      const bool act = j < m;
      u32 i = 0, x = 0; bool head = false, nexthead = true;
      if (act) {
        i = vA[j]; x = pA[j];
        const u64 kj = kA[j];
        head = j == 0 || kA[j - 1] != kj;
        if (j + 1 < m) nexthead = kA[j + 1] != kj;
      }
      u32 tot;
      u32 g = block_scan_incl_max<NT>(head ? x : 0u, sm, tot);
      g = max(g, carry_max);
      carry_max = max(carry_max, tot);
      const bool amb = act && !(head && nexthead);
      u32 tot2;
      const u32 inc = block_scan_incl_add<NT>(amb ? 1u : 0u, sm, tot2);
      if (act) { sa[x] = i; rank[i] = g; }   // keys were materialised before: safe to update in place
      if (amb) pB[carry_cnt + inc - 1] = x;
      carry_cnt += tot2;
    }
    m = carry_cnt;
    { u32* t = pA; pA = pB; pB = t; }
    __syncthreads();
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
      const u32 lim = min(ZQ_LCP_CAP, n - max(a, b));
      while (l < lim && T[a + l] == T[b + l]) ++l;
    }
    lcp[x] = (u16)l;
    bwt[x] = b > 0 ? T[b - 1] : (u8)0;
    if (idx16) { ((u16*)w)[x] = (u16)b; ((u16*)(w + stride))[x] = (u16)rank[x]; }
    else { ((u32*)w)[x] = b; ((u32*)(w + stride))[x] = rank[x]; }
  }
}

// Grid-stride over the units of a wave that need a suffix array; scratch is per CTA.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_suffix_sort(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const int* __restrict__ todo, int ntodo,
              u8* __restrict__ work_base, u64* kbuf, u32* vbuf, u64 scratch_elems) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortSmem<NT>& sm = *reinterpret_cast<SortSmem<NT>*>(smem_raw);
  SortScratch sc;
  sc.kA = kbuf + (u64)blockIdx.x * 2 * scratch_elems; sc.kB = sc.kA + scratch_elems;
  u32* vb = vbuf + (u64)blockIdx.x * 6 * scratch_elems;
  sc.vA = vb; sc.vB = vb + scratch_elems; sc.pA = vb + 2 * scratch_elems; sc.pB = vb + 3 * scratch_elems;
  sc.sa = vb + 4 * scratch_elems; sc.rank = vb + 5 * scratch_elems;
  for (int t = blockIdx.x; t < ntodo; t += gridDim.x) {
    const ZqUnit u = units[todo[t]];
    suffix_sort_block<NT>(in_base + u.in_off, u.n, work_base + u.work_off, u.idx16 != 0, sc, sm);
    __syncthreads();
  }
}

}  // namespace zqdev
