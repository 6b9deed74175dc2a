// zq_common.cuh -- shared device helpers and descriptors for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define ZQ_FULL 0xffffffffu

// One compression unit (== one libzpaq::compressBlock call, Z:20255) as the kernels see it.
struct ZqUnit {
  u64 in_off;    // byte offset of the input in the device input arena
  u64 work_off;  // byte offset of this unit's sa | isa | lcp region in the per-wave work arena
  u64 lz_off;    // byte offset into the pre-pass stream buffer
  u64 model_off; // byte offset of this unit's component tables + VM memory in the model arena
  u64 coded_off; // byte offset of the arithmetic coder's output
  u32 n;         // input length
  u32 plan;      // index into the plan table
  u32 lz_cap;    // capacity reserved at lz_off
  u32 prefix_off, prefix_len;  // block prefix (tag .. segment header) in the blob
  u32 idx16;     // 1: sa/isa stored as u16 (n <= 65536), 0: u32
  u32 coded_cap; // capacity reserved at coded_off
  u32 pad;
};

// Per (method, block size class) constants (== makeConfig's args, Z:19620-19628).
struct ZqPlan {
  int args[9];
  u32 payload_off, payload_len;  // selector + PCOMP bytes that precede the stream, in the blob
  u32 lz_level;                  // 0 none, 1 var-length codes, 2 byte codes, 3 BWT
  u32 use_sa;
  u32 e8e9;
  u32 modeled;                   // ncomp > 0
  u32 cm_plan;                   // index into the ZqCmPlan table when modeled
  u32 pad;
};

// per-unit work region: sa | isa (index width w = 2 or 4) | lcp (u16) | bwt (u8), each padded to 128 B
__host__ __device__ inline u64 zq_work_stride(u32 n, u32 w) { return (((u64)n + 1) * w + 127) & ~(u64)127; }
__host__ __device__ inline u64 zq_work_bytes(u32 n, u32 w) { return 2 * zq_work_stride(n, w) + zq_work_stride(n, 2) + zq_work_stride(n, 1); }

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
#ifdef ZQ_EMU   // host SIMT emulator build (tests/emu): no PTX, dynamic shared memory is a plain pointer
__device__ __forceinline__ u32 lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
#define ZQ_DYN_SMEM(name) unsigned char* name = emu::dyn_smem
#else
__device__ __forceinline__ u32 lanemask_lt() {
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
#define ZQ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
// floor(log2(x))+1, 0 for 0 (== reference lg(), Z:19269)
__device__ __forceinline__ int zq_bitlen(u32 x) { return 32 - __clz(x); }
