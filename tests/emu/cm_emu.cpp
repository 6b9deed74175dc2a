// cm_emu.cpp -- TEST INFRASTRUCTURE ONLY: runs the device code of the context-mixing coder
// (zpaqfranz_b200/csrc/zq_cm.cuh, compiled for the host through tests/emu/simt_emu.h) on one block and
// returns the arithmetic coder's bytes, so tests/test_cm_emu.py can compare them with the oracle on CPU.
// Build: g++ -O1 -std=c++17 -Itests/emu/shim -Izpaqfranz_b200/csrc -shared -fPIC tests/emu/cm_emu.cpp
//        zpaqfranz_b200/csrc/zq_cm_host.cpp zpaqfranz_b200/csrc/zq_config.cpp
#include <cuda_runtime.h>   // the shim

#include "zq_decode.cuh"
#include "zq_cm_host.h"

using namespace zqdev;

namespace {
void host_fill(const ZqCmFill& f, const zq::CmTables& t, u8* region) {
  u8* dst = region + f.off;
  switch (f.kind) {
    case ZQ_FILL_ZERO: memset(dst, 0, f.bytes); break;
    case ZQ_FILL_MATCHBUF: memset(dst, 0, f.bytes); dst[0] = 1; break;
    case ZQ_FILL_U32: for (u64 k = 0; k + 4 <= f.bytes; k += 4) memcpy(dst + k, &f.value, 4); break;
    case ZQ_FILL_U16: { const u16 v = (u16)f.value; for (u64 k = 0; k + 2 <= f.bytes; k += 2) memcpy(dst + k, &v, 2); } break;
    case ZQ_FILL_SSE:
      for (u64 k = 0; k + 4 <= f.bytes; k += 4) {
        const u32 j = (u32)(k / 4);
        const u32 v = (u32)t.squash[(j & 31) * 64 - 992 + 2048] << 17 | f.value;
        memcpy(dst + k, &v, 4);
      }
      break;
    case ZQ_FILL_ICM: memcpy(dst, t.icm_init, f.bytes); break;
    case ZQ_FILL_ISSE: memcpy(dst, t.isse_init, f.bytes); break;
  }
}
}  // namespace

// header: block header bytes (hsize .. hcomp 0); payload: selector (+PCOMP) bytes coded before the stream.
// Returns the number of coded bytes written to out, or a negative error.
extern "C" long emu_cm_encode(const uint8_t* header, uint32_t hlen, const uint8_t* payload, uint32_t plen,
                              const uint8_t* stream, uint32_t slen, uint8_t* out, uint32_t cap, int threads, int prefetch) {
  (void)prefetch;
  try {
    size_t used = 0;
    zq::Assembled code = zq::parse_block_header(header, hlen, &used);
    std::vector<ZqCmFill> fills;
    ZqCmPlan cp = zq::make_cm_plan(code, fills);
    const zq::CmTables& tab = zq::cm_tables();
    std::vector<u8> blob(payload, payload + plen);
    cp.hcomp_off = (u32)blob.size(); cp.hcomp_len = (u32)code.hcomp.size();
    blob.insert(blob.end(), code.hcomp.begin(), code.hcomp.end());
    blob.resize(blob.size() + 16);   // like the device blob: slack after the last byte code
    u8* model = (u8*)aligned_alloc(256, (size_t)cp.model_bytes + 256);
    for (u32 j = 0; j < cp.fill_count; ++j) host_fill(fills[cp.fill_first + j], tab, model);
    ZqUnit u; memset(&u, 0, sizeof u);
    u.n = slen; u.plan = 0; u.coded_cap = cap;
    ZqPlan pl; memset(&pl, 0, sizeof pl);
    pl.payload_off = 0; pl.payload_len = plen; pl.lz_level = 0; pl.modeled = 1; pl.cm_plan = 0;
    int todo = 0;
    u32 coded_len = 0, err = 0, next = 0, lzlen = slen;
    static_assert(sizeof(CmTablesDev) == sizeof(zq::CmTables), "table layout");
    const CmTablesDev* dtab = (const CmTablesDev*)&tab;
    // bit 4 of the flags: contexts precomputed (the role the translated program plays on the device); here by the
    // interpreter, byte-major H[0..n) after every byte but the last
    std::vector<u32> ctxbuf;
    const u32* ctxp = nullptr; const u64* ctxo = nullptr; u64 ctx_zero = 0;
    if (prefetch & 16) {
      const u32 K = plen + slen;
      std::vector<u8> M((size_t)1 << cp.hm, 0), prog(code.hcomp); prog.resize(prog.size() + 16);
      std::vector<u32> H((size_t)1 << cp.hh, 0), R(256, 0);
      CmVm vm; memset(&vm, 0, sizeof vm);
      vm.m = M.data(); vm.h = H.data(); vm.r = R.data(); vm.mmask = (u32)M.size() - 1; vm.hmask = (u32)H.size() - 1;
      vm.code = prog.data(); vm.len = (int)code.hcomp.size();
      ctxbuf.assign((size_t)(K ? K : 1) * (cp.n ? cp.n : 1), 0);
      for (u32 k = 0; k + 1 < K; ++k) {
        cm_vm_run_switch<false>(vm, k < plen ? payload[k] : stream[k - plen], nullptr);
        for (int i = 0; i < cp.n; ++i) ctxbuf[(size_t)k * cp.n + i] = H[i & vm.hmask];
      }
      ctxp = ctxbuf.data(); ctxo = &ctx_zero;
    }
    const unsigned block = (unsigned)threads < 64 ? 64u : (unsigned)threads & ~63u;   // whole (coder, context) pairs
    emu::launch(1, block, sizeof(CmSmem) + (block / 64) * sizeof(CmUnitSmem), [&] {
      if (prefetch & 16) k_cm_encode<0, true>(stream, &u, &pl, &cp, &todo, 1, dtab, blob.data(), stream, &lzlen, model, out, &coded_len, &err, &next, prefetch & 1, (prefetch >> 1) & 1, ctxp, ctxo);
      else if (prefetch & 8) k_cm_encode<2, false>(stream, &u, &pl, &cp, &todo, 1, dtab, blob.data(), stream, &lzlen, model, out, &coded_len, &err, &next, prefetch & 1, (prefetch >> 1) & 1, ctxp, ctxo);
      else if (prefetch & 4) k_cm_encode<1, false>(stream, &u, &pl, &cp, &todo, 1, dtab, blob.data(), stream, &lzlen, model, out, &coded_len, &err, &next, prefetch & 1, (prefetch >> 1) & 1, ctxp, ctxo);
      else k_cm_encode<0, false>(stream, &u, &pl, &cp, &todo, 1, dtab, blob.data(), stream, &lzlen, model, out, &coded_len, &err, &next, prefetch & 1, (prefetch >> 1) & 1, ctxp, ctxo);
    });
    free(model);
    if (err) return -(long)err;
    return (long)coded_len;
  } catch (const zq::Error& e) {
    fprintf(stderr, "emu_cm_encode: %s\n", e.msg.c_str());
    return -100;
  }
}
// Decodes the coded data of one block (everything after the segment header) with the device decoder.
// Returns the number of restored bytes or -(1000 + device error code).
static std::vector<ZqDecSeg> g_segs;     // segment table and input position of the last emu_cm_decode
static u32 g_consumed = 0;
extern "C" unsigned emu_cm_decode_segments(uint32_t* out3, unsigned cap, uint32_t* consumed) {
  for (size_t k = 0; k < g_segs.size() && k < cap; ++k) { out3[3 * k] = g_segs[k].unit; out3[3 * k + 1] = g_segs[k].out_end; out3[3 * k + 2] = g_segs[k].trailer; }
  if (consumed) *consumed = g_consumed;
  return (unsigned)g_segs.size();
}
extern "C" long emu_cm_decode(const uint8_t* header, uint32_t hlen, const uint8_t* coded, uint32_t clen, uint8_t* out, uint32_t cap, int fast) {
  (void)fast;
  try {
    size_t used = 0;
    zq::Assembled code = zq::parse_block_header(header, hlen, &used);
    std::vector<ZqCmFill> fills;
    ZqCmPlan cp;
    if (code.ncomp > 0) cp = zq::make_cm_plan(code, fills);
    else { memset(&cp, 0, sizeof cp); cp.hh = code.hh; cp.hm = code.hm; cp.fill_first = 0; }
    zq::add_pcomp_region(cp, code.ph, code.pm, fills);
    const zq::CmTables& tab = zq::cm_tables();
    std::vector<u8> blob(code.hcomp.begin(), code.hcomp.end());
    cp.hcomp_off = 0; cp.hcomp_len = (u32)code.hcomp.size();
    blob.resize(blob.size() + 16);
    u8* model = (u8*)aligned_alloc(256, (size_t)cp.model_bytes + 256);
    for (u32 j = 0; j < cp.fill_count; ++j) host_fill(fills[cp.fill_first + j], tab, model);
    ZqDecUnit u; memset(&u, 0, sizeof u);
    u.data_off = 0; u.data_len = clen; u.out_off = 0; u.model_off = 0; u.out_cap = cap; u.plan = 0;
    ZqDecResult res; memset(&res, 0, sizeof res);
    u32 next = 0, nseg = 0;
    std::vector<ZqDecSeg> segs(64);
    const CmTablesDev* dtab = (const CmTablesDev*)&tab;
    const size_t smem = sizeof(CmSmem) + sizeof(CmUnitSmem);
    emu::launch(1, 32, smem, [&] {
      if (fast & 8) k_cm_decode<2>(coded, &u, &cp, 1, dtab, blob.data(), model, out, &res, &next, fast & 1, segs.data(), &nseg, 64u, 0u);
      else if (fast & 4) k_cm_decode<1>(coded, &u, &cp, 1, dtab, blob.data(), model, out, &res, &next, fast & 1, segs.data(), &nseg, 64u, 0u);
      else k_cm_decode<0>(coded, &u, &cp, 1, dtab, blob.data(), model, out, &res, &next, fast & 1, segs.data(), &nseg, 64u, 0u);
    });
    free(model);
    segs.resize(nseg < 64 ? nseg : 64);
    g_segs = segs; g_consumed = res.consumed;
    if (res.error) return -(1000 + (long)res.error);
    return (long)res.out_len;
  } catch (const zq::Error& e) {
    fprintf(stderr, "emu_cm_decode: %s\n", e.msg.c_str());
    return -100;
  }
}
extern "C" unsigned long long emu_collectives() { return emu::collectives; }
