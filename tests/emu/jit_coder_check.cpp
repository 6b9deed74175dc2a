// jit_coder_check.cpp -- TEST INFRASTRUCTURE ONLY: runs a generated translation unit (zq_jit.cpp: zq_hcomp +
// zq_encode_block) on the host: contexts from the translated HCOMP, block coded by the generated straight-line model,
// model region initialised as k_cm_init would (the fills of make_cm_plan).  tests/test_jit.py compares the coded bytes
// with the oracle.  Compiled per model with -DZQ_JIT_GENERATED="\"<generated source>\"".
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "zq_cm_host.h"

#define ZQ_JIT_FN static inline
#include ZQ_JIT_GENERATED

namespace {
void host_fill(const ZqCmFill& f, const zq::CmTables& t, uint8_t* region) {
  uint8_t* dst = region + f.off;
  switch (f.kind) {
    case ZQ_FILL_ZERO: memset(dst, 0, f.bytes); break;
    case ZQ_FILL_MATCHBUF: memset(dst, 0, f.bytes); dst[0] = 1; break;
    case ZQ_FILL_U32: for (uint64_t k = 0; k + 4 <= f.bytes; k += 4) memcpy(dst + k, &f.value, 4); break;
    case ZQ_FILL_U16: { const uint16_t v = (uint16_t)f.value; for (uint64_t k = 0; k + 2 <= f.bytes; k += 2) memcpy(dst + k, &v, 2); } break;
    case ZQ_FILL_SSE:
      for (uint64_t k = 0; k + 4 <= f.bytes; k += 4) {
        const uint32_t j = (uint32_t)(k / 4), v = (uint32_t)t.squash[(j & 31) * 64 - 992 + 2048] << 17 | f.value;
        memcpy(dst + k, &v, 4);
      }
      break;
    case ZQ_FILL_ICM: memcpy(dst, t.icm_init, f.bytes); break;
    case ZQ_FILL_ISSE: memcpy(dst, t.isse_init, f.bytes); break;
  }
}
}  // namespace

extern "C" long jit_encode(const uint8_t* header, uint32_t hlen, const uint8_t* payload, uint32_t plen, const uint8_t* stream,
                           uint32_t slen, uint8_t* out, uint32_t cap) {
  try {
    size_t used = 0;
    zq::Assembled code = zq::parse_block_header(header, hlen, &used);
    std::vector<ZqCmFill> fills;
    ZqCmPlan cp = zq::make_cm_plan(code, fills);
    const zq::CmTables& tab = zq::cm_tables();
    std::vector<uint8_t> modelbuf((size_t)cp.model_bytes + 512);
    uint8_t* model = (uint8_t*)(((uintptr_t)modelbuf.data() + 255) & ~(uintptr_t)255);
    for (uint32_t j = 0; j < cp.fill_count; ++j) host_fill(fills[cp.fill_first + j], tab, model);
    // contexts: H[0..n) after every byte but the last (what zq_ctx_kernel stores)
    const uint32_t K = plen + slen;
    std::vector<unsigned> ctx((size_t)(K ? K : 1) * cp.n + 1, 0);
    unsigned char* M = model + cp.m_off; unsigned* H = (unsigned*)(model + cp.h_off); unsigned* R = (unsigned*)(model + cp.r_off);
    ZqJitVm v; memset(&v, 0, sizeof v);
    int err = 0;
    const unsigned hmask = (1u << cp.hh) - 1;
    for (uint32_t k = 0; k + 1 < K; ++k) {
      zq_hcomp(v, M, H, R, k < plen ? payload[k] : stream[k - plen], err);
      for (int i = 0; i < cp.n; ++i) ctx[(size_t)k * cp.n + i] = H[i & hmask];
    }
    if (err) return -3;
    int overflow = 0;
    const unsigned n = zq_encode_block(payload, plen, stream, slen, ctx.data(), model, tab.stretch, tab.squash, tab.dt, tab.dt2k, tab.ns, out, cap, &overflow);
    return overflow ? -2 : (long)n;
  } catch (const zq::Error& e) {
    fprintf(stderr, "jit_encode: %s\n", e.msg.c_str());
    return -100;
  }
}
