// zq_frame.cuh -- assembles the finished ZPAQ block of an UNMODELED method (n components == 0:
// store, -m1, -m2): what Compressor::{writeTag,startBlock,startSegment,postProcess,compress,
// endSegment,endBlock} (Z:15970-16187) and the unmodeled branch of Encoder::compress
// (Z:15590-15600) write.  One CTA per block; pure byte movement (HBM bound, ~2 B moved per
// output byte).
//
//   prefix  = tag[13] 'z' 'P' 'Q' lvl 1 header 01 filename 00 comment 00 00      (built on the host)
//   payload = selector/PCOMP bytes (host) followed by the pre-pass stream (device)
//   body    = payload cut in <= 65536-byte chunks, each preceded by its big-endian u32 length
//   trailer = 00 00 00 00 (FD sha1[20] | FE) FF
#pragma once
#include "zq_common.cuh"

namespace zqdev {

__host__ __device__ inline u64 unmodeled_block_size(u32 prefix_len, u64 payload_total, bool sha) {
  const u64 nchunks = (payload_total + 65535) >> 16;
  return prefix_len + payload_total + 4 * nchunks + 4 + (sha ? 21 : 1) + 1;
}

__device__ __forceinline__ void frame_trailer(u8* tr, const u8* sha1, int ui) {
  if (threadIdx.x < 4) tr[threadIdx.x] = 0;
  if (sha1) {
    if (threadIdx.x == 4) tr[4] = 253;
    if (threadIdx.x >= 32 && threadIdx.x < 52) tr[5 + threadIdx.x - 32] = sha1[(size_t)ui * 20 + threadIdx.x - 32];
    if (threadIdx.x == 5) tr[25] = 255;
  } else {
    if (threadIdx.x == 4) tr[4] = 254;
    if (threadIdx.x == 5) tr[5] = 255;
  }
}

// Modeled blocks (n components > 0): prefix | arithmetic-coded bytes | trailer.
__global__ void __launch_bounds__(256)
k_frame(const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans, const int* __restrict__ todo, int ntodo,
        const u8* __restrict__ blob, const u8* __restrict__ in_base, const u8* __restrict__ lz_base,
        const u32* __restrict__ lz_len, const u8* __restrict__ coded_base, const u32* __restrict__ coded_len,
        const u8* __restrict__ sha1 /* 20 B per unit or null */, const u64* __restrict__ out_off, u8* __restrict__ out_base) {
  for (int t = blockIdx.x; t < ntodo; t += gridDim.x) {
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const ZqPlan pl = plans[u.plan];
    u8* __restrict__ out = out_base + out_off[ui];
    if (pl.modeled) {
      for (u32 k = threadIdx.x; k < u.prefix_len; k += blockDim.x) out[k] = blob[u.prefix_off + k];
      const u32 cl = coded_len[ui];
      const u8* __restrict__ src = coded_base + u.coded_off;
      u8* __restrict__ body = out + u.prefix_len;
      for (u32 k = threadIdx.x; k < cl; k += blockDim.x) body[k] = src[k];
      frame_trailer(body + cl, sha1, ui);
      continue;
    }
    const u8* __restrict__ stream; u32 slen;
    if (pl.lz_level) { stream = lz_base + u.lz_off; slen = lz_len[ui]; }
    else { stream = in_base + u.in_off; slen = u.n; }
    for (u32 k = threadIdx.x; k < u.prefix_len; k += blockDim.x) out[k] = blob[u.prefix_off + k];
    const u32 plen = pl.payload_len;
    const u64 total = (u64)plen + slen;
    const u64 nchunks = (total + 65535) >> 16;
    u8* __restrict__ body = out + u.prefix_len;
    for (u64 c = threadIdx.x; c < nchunks; c += blockDim.x) {
      const u64 rem = total - (c << 16);
      const u32 cl = rem < 65536 ? (u32)rem : 65536u;
      u8* h = body + c * 65540;
      h[0] = cl >> 24; h[1] = cl >> 16; h[2] = cl >> 8; h[3] = cl;
    }
    for (u64 s = threadIdx.x; s < total; s += blockDim.x) {
      const u8 b = s < plen ? blob[pl.payload_off + s] : stream[s - plen];
      body[s + 4 * ((s >> 16) + 1)] = b;
    }
    frame_trailer(body + total + 4 * nchunks, sha1, ui);
  }
}

// E8E9 pre-filter (e8e9(), Z:19162): x86 CALL/JMP rel32 operands become absolute, scanning
// backwards; a transform changes bytes a later (lower) position tests, so it is sequential per
// block: one thread per block (rare: only "exe" typed blocks ask for it).
__global__ void __launch_bounds__(64)
k_e8e9(u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const int* __restrict__ todo, int ntodo) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntodo) return;
  const ZqUnit u = units[todo[t]];
  u8* buf = in_base + u.in_off;
  for (int i = (int)u.n - 5; i >= 0; --i) {
    if ((buf[i] & 254) == 0xe8 && ((buf[i + 4] + 1) & 254) == 0) {
      const u32 a = ((u32)buf[i + 1] | (u32)buf[i + 2] << 8 | (u32)buf[i + 3] << 16) + (u32)i;
      buf[i + 1] = (u8)a; buf[i + 2] = (u8)(a >> 8); buf[i + 3] = (u8)(a >> 16);
    }
  }
}

// Byte-gap period analysis of compressBlock's levels >= 5 (Z:20355-20388): histogram of the distance
// to the previous occurrence of the same byte value (first occurrences count from position 0, as the
// reference's zero-initialised table does), then up to two periods picked by the same double
// arithmetic.  One 256-thread CTA per block: thread b follows byte value b through the block (bytes are
// staged in shared memory, every thread reads the same one = broadcast), histogram in shared memory.
__global__ void __launch_bounds__(256)
k_gap_periods(const u8* __restrict__ in_base, const u64* __restrict__ off, const u32* __restrict__ len, int n, int* __restrict__ periods) {
  __shared__ u32 gap[4096];
  __shared__ u8 tile[4096];
  for (int t = blockIdx.x; t < n; t += gridDim.x) {
    const u8* __restrict__ p = in_base + off[t];
    const u32 L = len[t];
    for (u32 k = threadIdx.x; k < 4096; k += blockDim.x) gap[k] = 0;
    __syncthreads();
    const u32 mine = threadIdx.x;
    u32 last = 0;
    for (u32 base = 0; base < L; base += 4096) {
      const u32 cnt = min(4096u, L - base);
      for (u32 k = threadIdx.x; k < cnt; k += blockDim.x) tile[k] = p[base + k];
      __syncthreads();
      for (u32 k = 0; k < cnt; ++k) {
        if (tile[k] == mine) {
          const u32 i = base + k, d = i - last;
          if (d > 0 && d < 4096) atomicAdd(&gap[d], 1u);
          last = i;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      int per[2] = {0, 0};
      int n1 = (int)L - (int)gap[1] - (int)gap[2] - (int)gap[3];
      for (int rep = 0; rep < 2; ++rep) {
        int period = 0, tt = 0;
        double best = 0;
        for (int j = 5; j < 4096 && tt < n1; ++j) {
          const double sc = (double)(int)gap[j] / (256.0 + n1 - tt);
          if (sc > best) best = sc, period = j;
          tt += (int)gap[j];
        }
        if (!(period > 4 && best > 0.1)) break;
        per[rep] = period;
        n1 -= (int)gap[period];
        gap[period] = 0;
      }
      periods[2 * t] = per[0]; periods[2 * t + 1] = per[1];
    }
    __syncthreads();
  }
}

}  // namespace zqdev
