// jit_kernel_check.cpp -- TEST INFRASTRUCTURE ONLY: the generated KERNELS (zq_ctx_kernel, zq_code_kernel of zq_jit.cpp)
// run under the SIMT emulator with the argument layout zq_api.cu builds for them (per-group arrays of stream offset /
// length, model offset, context offset, coded offset / capacity, unit ids): several blocks of different lengths in
// one launch, one block per thread (stride 1) or per warp (stride 32).  tests/test_jit.py compares every block's
// coded bytes with the oracle.  Compiled per model with -DZQ_JIT_GENERATED="\"<generated source>\"".
#include <cuda_runtime.h>   // the shim

#include "zq_cm_host.h"

#define __CUDACC__ 1        // the kernels of the generated translation unit
#include ZQ_JIT_GENERATED
#undef __CUDACC__

typedef uint8_t u8; typedef uint32_t u32; typedef uint64_t u64;

namespace {
void host_fill(const ZqCmFill& f, const zq::CmTables& t, u8* region) {
  u8* dst = region + f.off;
  switch (f.kind) {
    case ZQ_FILL_ZERO: memset(dst, 0, f.bytes); break;
    case ZQ_FILL_MATCHBUF: memset(dst, 0, f.bytes); dst[0] = 1; break;
    case ZQ_FILL_U32: for (u64 k = 0; k + 4 <= f.bytes; k += 4) memcpy(dst + k, &f.value, 4); break;
    case ZQ_FILL_U16: { const uint16_t v = (uint16_t)f.value; for (u64 k = 0; k + 2 <= f.bytes; k += 2) memcpy(dst + k, &v, 2); } break;
    case ZQ_FILL_SSE:
      for (u64 k = 0; k + 4 <= f.bytes; k += 4) {
        const u32 j = (u32)(k / 4), v = (u32)t.squash[(j & 31) * 64 - 992 + 2048] << 17 | f.value;
        memcpy(dst + k, &v, 4);
      }
      break;
    case ZQ_FILL_ICM: memcpy(dst, t.icm_init, f.bytes); break;
    case ZQ_FILL_ISSE: memcpy(dst, t.isse_init, f.bytes); break;
  }
}
}  // namespace

// nunits blocks: streams back to back in `streams` (soff/slen), shared payload.  out: coded bytes of block t at
// out + t * cap, lengths in out_len.  Returns 0 or a negative code.
extern "C" int jit_kernels(const u8* header, u32 hlen, const u8* payload, u32 plen, const u8* streams, const u64* soff, const u32* slen,
                           int nunits, int stride, u8* out, u32 cap, u32* out_len) {
  try {
    size_t used = 0;
    zq::Assembled code = zq::parse_block_header(header, hlen, &used);
    std::vector<ZqCmFill> fills;
    ZqCmPlan cp = zq::make_cm_plan(code, fills);
    const zq::CmTables& tab = zq::cm_tables();
    const u64 mstride = (cp.model_bytes + 255) & ~(u64)255;
    std::vector<u8> modelbuf((size_t)mstride * nunits + 512);
    u8* mbase = (u8*)(((uintptr_t)modelbuf.data() + 255) & ~(uintptr_t)255);
    std::vector<u64> moff(nunits), coff(nunits), cdoff(nunits);
    std::vector<u32> cdcap(nunits, cap), lens((size_t)nunits + 8, 0);
    std::vector<int> ids(nunits);
    u64 ctx_elems = 0;
    for (int t = 0; t < nunits; ++t) {
      moff[t] = (u64)t * mstride;
      for (u32 j = 0; j < cp.fill_count; ++j) host_fill(fills[cp.fill_first + j], tab, mbase + moff[t]);
      coff[t] = ctx_elems; ctx_elems += ((u64)plen + slen[t]) * cp.n;
      cdoff[t] = (u64)t * cap;
      ids[t] = nunits - 1 - t;            // unit ids need not be the launch order
    }
    std::vector<u32> ctx(ctx_elems + 16, 0xdeadbeefu);
    u32 err = 0;
    unsigned long long mo = cp.m_off, ho = cp.h_off, ro = cp.r_off;
    emu::launch((nunits + 63) / 64, 64, 0, [&] {
      zq_ctx_kernel(payload, plen, streams, (const unsigned long long*)soff, slen, nunits, mbase, (const unsigned long long*)moff.data(), mo, ho, ro,
                    ctx.data(), (const unsigned long long*)coff.data(), &err);
    });
    if (err) return -3;
    emu::launch((nunits * stride + 127) / 128, 128, 0, [&] {
      zq_code_kernel(payload, plen, streams, (const unsigned long long*)soff, slen, nunits, stride, mbase, (const unsigned long long*)moff.data(),
                     ctx.data(), (const unsigned long long*)coff.data(), (const u8*)&tab, out, (const unsigned long long*)cdoff.data(), cdcap.data(),
                     ids.data(), lens.data(), &err);
    });
    if (err) return -4;
    for (int t = 0; t < nunits; ++t) out_len[t] = lens[ids[t]];
    return 0;
  } catch (const zq::Error& e) {
    fprintf(stderr, "jit_kernels: %s\n", e.msg.c_str());
    return -100;
  }
}
