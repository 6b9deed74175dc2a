// zq_fragment.cuh -- content-defined fragmenter of the dedup stage.
//
// Replaces the chunker loop inlined in Jidac::add (Z:122457-122561; canonical form Z:95604-95633,
// constants Z:121626-121631): per byte  pred = o1[c1]; h = (h+c+1) * (c==pred ? 314159265 : 271828182);
// o1[c1] = c; c1 = c;  a fragment ends when h < 2^(22-fragment) and size >= MIN, or size >= MAX, or EOF;
// all state resets at every fragment start.
//
// The machine is sequential per file, so parallelism is (a) across files and (b) inside a file by
// speculate-and-stitch, exact by construction: files are cut into fixed segments; in round 0 every
// segment runs a fresh machine from its first byte up to its first boundary at or past the segment
// end (its "exit").  In later rounds a segment whose predecessor's exit changed restarts from that
// exit and stops as soon as it lands on a boundary of its own previous chain (both machines start
// fresh there, so the old tail is the true continuation).  Rounds repeat until no exit changes;
// segment 0 of each file is true from the start, so the fix point is the reference's chain.
// One thread per segment; the 256-byte o1 table of each thread lives in shared memory.
//
// Constant runs (zero pages, sparse files) never re-synchronise: every machine started inside the run cuts
// fragments of the same length L(v) relative to ITS start, so a speculative chain is only right if its start is
// congruent to the true one.  A fresh machine fed nothing but the byte v is a fixed sequence, so L(v) and its hit
// count are tabulated once (k_fragment_const_table); round 0 marks the segments that hold a single byte value,
// and from round 1 on a fragment that starts inside such a segment and fits before the end of the constant run is
// emitted in closed form instead of being scanned; a segment deep inside a run takes its entry straight from the
// run's entry (entry_run + j*L(v)), so a whole run settles in one round once its entry is final.
#pragma once
#include "zq_common.cuh"

namespace zqdev {

struct ZqSeg {
  u64 begin, end;     // this segment's byte range in the arena (end == next segment's begin or file end)
  u64 file_end;       // end of the file the segment belongs to
  u32 first;          // 1: first segment of its file (entry is always `begin`)
  u32 pad;
};

constexpr int FRAG_THREADS = 128;
constexpr u32 FRAG_NOT_CONST = 256;

// L(v), hits(v): length and predictor hits of the fragment a fresh machine cuts from an endless run of byte v.
// Only o1[0] (the virtual byte in front of the fragment is 0) and o1[v] are ever touched.
__global__ void __launch_bounds__(256)
k_fragment_const_table(u32 minf, u32 maxf, u32 thresh, u32* __restrict__ Lv, u32* __restrict__ Hv) {
  const u32 v = threadIdx.x;
  u32 o1_0 = 0, o1_v = 0, c1 = 0, h = 0, hits = 0, sz = 0;
  for (;;) {
    const u32 pred = (c1 == 0 || v == 0) ? o1_0 : o1_v;
    const bool hit = v == pred;
    h = (h + v + 1u) * (hit ? 314159265u : 271828182u);
    hits += hit;
    if (c1 == 0 || v == 0) o1_0 = v; else o1_v = v;
    c1 = v; ++sz;
    if (sz >= maxf || (thresh && h < thresh && sz >= minf)) break;
  }
  Lv[v] = sz; Hv[v] = hits;
}

__global__ void __launch_bounds__(FRAG_THREADS)
k_fragment_round(const u8* __restrict__ base, const ZqSeg* __restrict__ segs, int nseg, int round,
                 u32 minf, u32 maxf, u32 thresh /* 0 = never */, u32 cap,
                 const u64* __restrict__ exit_prev, u64* __restrict__ exit_next,
                 u64* __restrict__ entry, const u64* __restrict__ bnd_prev, const u32* __restrict__ hits_prev,
                 const u32* __restrict__ cnt_prev, u64* __restrict__ bnd_next, u32* __restrict__ hits_next,
                 u32* __restrict__ cnt_next, u32* __restrict__ changed, u32* __restrict__ overflow,
                 u32* __restrict__ constv /* round 0 writes, later rounds read */, const u64* __restrict__ const_end,
                 const u32* __restrict__ run_first, const u32* __restrict__ Lv, const u32* __restrict__ Hv) {
  __shared__ u8 o1s[FRAG_THREADS][256 + 4];   // +4: stagger banks between threads
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nseg) return;
  u8* o1 = o1s[threadIdx.x];
  const ZqSeg sg = segs[k];
  u64 e = (round == 0 || sg.first) ? sg.begin : exit_prev[k - 1];
  const u32 cv = round ? constv[k] : FRAG_NOT_CONST;
  const u64 cend = round ? const_end[k] : 0;
  if (cv < FRAG_NOT_CONST && run_first[k] != (u32)k) {   // inside a constant run: boundaries are entry_run + j*L(v)
    const u32 r = run_first[k];
    const u64 er = segs[r].first ? segs[r].begin : exit_prev[r - 1];
    const u64 L = Lv[cv];
    if (er <= sg.begin) {
      const u64 cand = er + (sg.begin - er + L - 1) / L * L;
      if (cand <= cend) e = cand;
    }
  }
  const u64* oldb = bnd_prev + (u64)k * cap; const u32* oldh = hits_prev + (u64)k * cap;
  u64* nb = bnd_next + (u64)k * cap; u32* nh = hits_next + (u64)k * cap;
  const u32 oldn = round ? cnt_prev[k] : 0;
  if (round && e == entry[k]) {   // same entry as last time: chain unchanged
    for (u32 q = 0; q < oldn; ++q) { nb[q] = oldb[q]; nh[q] = oldh[q]; }
    cnt_next[k] = oldn; exit_next[k] = exit_prev[k];
    return;
  }
  entry[k] = e;
  u32 cnt = 0, op = 0;
  u64 pos = e, ex = e;
  bool merged = false;
  // round 0 (entry == begin): does the segment hold a single byte value?
  const u64 seg_stop = min(sg.end, sg.file_end);
  const u32 v0 = sg.begin < seg_stop ? (u32)base[sg.begin] : 0u;
  bool same = round == 0;
  while (pos < sg.end && pos < sg.file_end) {   // a fragment starts inside this segment
    u32 hits = 0;
    if (cv < FRAG_NOT_CONST && pos + Lv[cv] <= cend) {   // wholly inside a constant run: closed form
      pos += Lv[cv]; hits = Hv[cv];
    } else {
      for (int q = 0; q < 256; q += 4) *(u32*)(o1 + q) = 0;
      u32 h = 0, sz = 0, c1 = 0;
      for (;;) {
        const u32 c = base[pos];
        if (same && pos < seg_stop && c != v0) same = false;
        ++pos;
        const bool hit = c == o1[c1];
        h = (h + c + 1u) * (hit ? 314159265u : 271828182u);
        hits += hit;
        o1[c1] = (u8)c; c1 = c; ++sz;
        if (pos >= sg.file_end || sz >= maxf || (thresh && h < thresh && sz >= minf)) break;
      }
    }
    if (cnt < cap) { nb[cnt] = pos; nh[cnt] = hits; } else atomicOr(overflow, 1u);
    ++cnt;
    ex = pos;
    // landed on a boundary of the previous chain? then its tail is the continuation
    while (op < oldn && oldb[op] < pos) ++op;
    if (op < oldn && oldb[op] == pos) {
      for (u32 q = op + 1; q < oldn; ++q) { if (cnt < cap) { nb[cnt] = oldb[q]; nh[cnt] = oldh[q]; } ++cnt; }
      if (oldn) ex = max(ex, oldb[oldn - 1]);
      merged = true;
      break;
    }
  }
  (void)merged;
  if (round == 0) constv[k] = (same && sg.begin < seg_stop) ? v0 : FRAG_NOT_CONST;
  cnt_next[k] = min(cnt, cap);
  exit_next[k] = ex;
  if (!round || ex != exit_prev[k]) atomicOr(changed, 1u);
}

// fragment table in file order: frag k of segment s starts at the previous boundary
__global__ void k_fragment_gather(const ZqSeg* __restrict__ segs, int nseg, u32 cap, const u64* __restrict__ bnd,
                                  const u32* __restrict__ hits, const u32* __restrict__ cnt, const u64* __restrict__ first_out,
                                  const u64* __restrict__ entry, u32* __restrict__ frag_len, u32* __restrict__ frag_hits,
                                  u64* __restrict__ frag_off) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nseg) return;
  u64 prev = entry[k];
  const u64 o = first_out[k];
  for (u32 q = 0; q < cnt[k]; ++q) {
    const u64 b = bnd[(u64)k * cap + q];
    frag_len[o + q] = (u32)(b - prev); frag_hits[o + q] = hits[(u64)k * cap + q]; frag_off[o + q] = prev;
    prev = b;
  }
}

// The fragmenter's order-1 table as it stands when a fragment ends (Z:122186: o1 starts at zero with every
// fragment; o1[previous byte] = byte): entry v holds the byte that followed the LAST occurrence of v as
// "previous byte" (the virtual byte in front of the fragment is 0), 0 if v never preceded anything.  The
// archiver's type heuristics read it (Z:122608-122634).  One CTA per fragment, 256 threads.
__global__ void __launch_bounds__(256)
k_fragment_o1(const u8* __restrict__ base, const u64* __restrict__ frag_off, const u32* __restrict__ frag_len, int nfrag,
              u8* __restrict__ o1_out) {
  __shared__ int last[256];
  for (int f = blockIdx.x; f < nfrag; f += gridDim.x) {
    const u8* d = base + frag_off[f];
    const int n = (int)frag_len[f];
    last[threadIdx.x] = -1;
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += 256) {
      const u32 c1 = j ? d[j - 1] : 0u;
      atomicMax(&last[c1], j);
    }
    __syncthreads();
    const int l = last[threadIdx.x];
    o1_out[(u64)f * 256 + threadIdx.x] = l >= 0 ? d[l] : (u8)0;
    __syncthreads();
  }
}

}  // namespace zqdev
