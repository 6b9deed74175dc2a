// zq_cm_host.cpp -- see zq_cm_host.h.
#include "zq_cm_host.h"

#include <cmath>
#include <cstring>
#include <mutex>

namespace zq {
namespace {

// ---- bit-history state table: the ZPAQ specification's generator ---------------------------------
// A state stands for a pair (n0, n1) of bounded zero/one counts, plus which bit came last where both
// are kept.  States are numbered by increasing n0+n1.
int state_count(int n0, int n1) {
  static const int bound[6] = {20, 48, 15, 8, 6, 5};
  if (n0 < n1) std::swap(n0, n1);
  if (n1 < 0 || n1 >= 6 || n0 > bound[n1]) return 0;
  return 1 + ((n1 > 0 && n0 + n1 <= 17) ? 1 : 0);
}
int discounted(int v) { return (v >= 1) + (v >= 2) + (v >= 3) + (v >= 4) + (v >= 5) + (v >= 7) + (v >= 8); }
void advance(int& n0, int& n1, int y) {
  if (n0 < n1) { advance(n1, n0, 1 - y); return; }
  if (y) { ++n1; n0 = discounted(n0); } else { ++n0; n1 = discounted(n1); }
  while (!state_count(n0, n1)) {
    if (n1 < 2) --n0;
    else { n0 = (n0 * (n1 - 1) + n1 / 2) / n1; --n1; }
  }
}
void build_state_table(uint8_t ns[1024]) {
  const int N = 50;
  static uint8_t id[50][50][2];
  memset(id, 0, sizeof id);
  int next = 0;
  for (int tot = 0; tot < N; ++tot)
    for (int n1 = 0; n1 <= tot; ++n1) {
      const int n0 = tot - n1, k = state_count(n0, n1);
      if (k) { id[n0][n1][0] = (uint8_t)next; id[n0][n1][1] = (uint8_t)(next + k - 1); next += k; }
    }
  memset(ns, 0, 1024);
  for (int n0 = 0; n0 < N; ++n0)
    for (int n1 = 0; n1 < N; ++n1)
      for (int y = 0; y < state_count(n0, n1); ++y) {
        const int s = id[n0][n1][y];
        for (int bit = 0; bit < 2; ++bit) {
          int a = n0, b = n1;
          advance(a, b, bit);
          ns[s * 4 + bit] = id[a][b][bit];
        }
        ns[s * 4 + 2] = (uint8_t)n0;
        ns[s * 4 + 3] = (uint8_t)n1;
      }
}

int clamp512k(int x) { return x < -(1 << 19) ? -(1 << 19) : x >= (1 << 19) ? (1 << 19) - 1 : x; }

CmTables g_tab;
std::once_flag g_once;
std::string g_tab_err;

void build_tables() {
  CmTables& t = g_tab;
  t.dt2k[0] = 0;
  for (int i = 1; i < 256; ++i) t.dt2k[i] = 2048 / i;
  for (int i = 0; i < 1024; ++i) t.dt[i] = (1 << 17) / (i * 2 + 3) * 2;
  for (int i = 0; i < 32768; ++i)
    t.stretch[i] = (int16_t)((int)(std::log((i + 0.5) / (32767.5 - i)) * 64 + 0.5 + 100000) - 100000);
  for (int i = 0; i < 4096; ++i) t.squash[i] = (uint16_t)(int)(32768.0 / (1 + std::exp((i - 2048) * (-1.0 / 64))));
  // the reference's own known-answer test for these tables (Z:14942-14950)
  uint32_t stsum = 0, sqsum = 0;
  for (int i = 32767; i >= 0; --i) stsum = stsum * 3 + (uint32_t)(int)t.stretch[i];
  for (int i = 4095; i >= 0; --i) sqsum = sqsum * 3 + (uint32_t)t.squash[i];
  if (stsum != 3887533746u || sqsum != 2278286169u) { g_tab_err = "stretch/squash table checksum mismatch"; return; }
  build_state_table(t.ns);
  for (int j = 0; j < 256; ++j) {
    const int init = ((t.ns[j * 4 + 3] * 2 + 1) << 22) / (t.ns[j * 4 + 2] + t.ns[j * 4 + 3] + 1);
    t.icm_init[j] = (uint32_t)init;
    t.isse_init[2 * j] = 1u << 15;
    t.isse_init[2 * j + 1] = (uint32_t)clamp512k(t.stretch[init >> 8] * 1024);
  }
}

uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }
const int kCompBytes[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};

}  // namespace

const CmTables& cm_tables() {
  std::call_once(g_once, build_tables);
  if (!g_tab_err.empty()) throw Error(g_tab_err);
  return g_tab;
}

ZqCmPlan make_cm_plan(const Assembled& code, std::vector<ZqCmFill>& fills) {
  ZqCmPlan p;
  memset(&p, 0, sizeof p);
  if (code.ncomp > ZQ_CM_MAXCOMP) throw Error("models with more than 64 components have no device path yet");
  p.n = code.ncomp; p.hh = code.hh; p.hm = code.hm;
  if (p.hh > 24 || p.hm > 28) throw Error("HCOMP memory too large for the device path");
  p.fill_first = (uint32_t)fills.size();
  uint64_t off = 0;
  auto add_fill = [&](uint64_t o, uint64_t bytes, uint32_t kind, uint32_t value) {
    ZqCmFill f; f.off = o; f.bytes = bytes; f.kind = kind; f.value = value; fills.push_back(f);
  };
  p.m_off = off; off = align256(off + ((uint64_t)1 << p.hm));
  p.h_off = off; off = align256(off + ((uint64_t)4 << p.hh));
  p.r_off = off; off = align256(off + 1024);
  add_fill(0, off, ZQ_FILL_ZERO, 0);
  const uint8_t* cp = code.comp.data();
  int level[ZQ_CM_MAXCOMP];
  for (int i = 0; i < p.n; ++i) {
    ZqCmComp& c = p.comp[i];
    c.type = cp[0];
    const int len = kCompBytes[c.type];
    c.a1 = len > 1 ? cp[1] : 0; c.a2 = len > 2 ? cp[2] : 0; c.a3 = len > 3 ? cp[3] : 0;
    c.a4 = len > 4 ? cp[4] : 0; c.a5 = len > 5 ? cp[5] : 0;
    auto need = [&](bool ok, const char* msg) { if (!ok) throw Error(msg); };
    int lv = 0;
    switch (c.type) {
      case ZQ_CONS: break;
      case ZQ_CM: {
        need(c.a1 <= 32, "max size for CM is 32");
        need(c.a1 <= 28, "CM too large for the device path");
        const uint64_t bytes = (uint64_t)4 << c.a1;
        c.cm_off = off; c.cm_mask = (uint32_t)(((uint64_t)1 << c.a1) - 1);
        add_fill(off, bytes, ZQ_FILL_U32, 0x80000000u);
        off = align256(off + bytes);
        break;
      }
      case ZQ_ICM: {
        need(c.a1 <= 26, "max size for ICM is 26");
        const uint64_t hb = (uint64_t)64 << c.a1;
        c.ht_off = off; c.ht_mask = (uint32_t)(hb - 1);
        add_fill(off, hb, ZQ_FILL_ZERO, 0); off = align256(off + hb);
        c.cm_off = off; c.cm_mask = 255;
        add_fill(off, 1024, ZQ_FILL_ICM, 0); off = align256(off + 1024);
        break;
      }
      case ZQ_MATCH: {
        need(c.a1 <= 32 && c.a2 <= 32, "max size for MATCH is 32 32");
        need(c.a1 <= 28 && c.a2 <= 30, "MATCH too large for the device path");
        const uint64_t ib = (uint64_t)4 << c.a1, bb = (uint64_t)1 << c.a2;
        c.cm_off = off; c.cm_mask = (uint32_t)(((uint64_t)1 << c.a1) - 1);
        add_fill(off, ib, ZQ_FILL_ZERO, 0); off = align256(off + ib);
        c.ht_off = off; c.ht_mask = (uint32_t)(bb - 1);
        add_fill(off, bb, ZQ_FILL_MATCHBUF, 0); off = align256(off + bb);
        break;
      }
      case ZQ_AVG:
        need(c.a1 < i, "AVG j >= i"); need(c.a2 < i, "AVG k >= i");
        lv = 1 + std::max(level[c.a1], level[c.a2]);
        break;
      case ZQ_MIX2: {
        need(c.a1 <= 32, "max size for MIX2 is 32");
        need(c.a3 < i, "MIX2 k >= i"); need(c.a2 < i, "MIX2 j >= i");
        need(c.a1 <= 28, "MIX2 too large for the device path");
        const uint64_t bytes = (uint64_t)2 << c.a1;
        c.cm_off = off; c.cm_mask = (uint32_t)(((uint64_t)1 << c.a1) - 1);
        add_fill(off, bytes, ZQ_FILL_U16, 32768); off = align256(off + bytes);
        lv = 1 + std::max(level[c.a2], level[c.a3]);
        break;
      }
      case ZQ_MIX: {
        need(c.a1 <= 32, "max size for MIX is 32");
        need(c.a2 < i, "MIX j >= i");
        need(c.a3 >= 1 && c.a3 <= i - c.a2, "MIX m not in 1..i-j");
        need(c.a1 <= 26, "MIX too large for the device path");
        const uint64_t bytes = ((uint64_t)4 * c.a3) << c.a1;
        c.cm_off = off; c.cm_mask = (uint32_t)(((uint64_t)1 << c.a1) - 1);   // rows-1
        add_fill(off, bytes, ZQ_FILL_U32, (uint32_t)(65536 / c.a3)); off = align256(off + bytes);
        for (int j = 0; j < c.a3; ++j) lv = std::max(lv, 1 + level[c.a2 + j]);
        if (i < ZQ_CM_LANES) p.mix_mask |= 1u << i;
        break;
      }
      case ZQ_ISSE: {
        need(c.a1 <= 32, "max size for ISSE is 32");
        need(c.a2 < i, "ISSE j >= i");
        need(c.a1 <= 26, "ISSE too large for the device path");
        const uint64_t hb = (uint64_t)64 << c.a1;
        c.ht_off = off; c.ht_mask = (uint32_t)(hb - 1);
        add_fill(off, hb, ZQ_FILL_ZERO, 0); off = align256(off + hb);
        c.cm_off = off; c.cm_mask = 511;
        add_fill(off, 2048, ZQ_FILL_ISSE, 0); off = align256(off + 2048);
        lv = 1 + level[c.a2];
        break;
      }
      case ZQ_SSE: {
        need(c.a1 <= 32, "max size for SSE is 32");
        need(c.a2 < i, "SSE j >= i");
        need(c.a3 <= c.a4 * 4, "SSE start > limit*4");
        need(c.a1 <= 24, "SSE too large for the device path");
        const uint64_t bytes = (uint64_t)128 << c.a1;
        c.cm_off = off; c.cm_mask = (uint32_t)(((uint64_t)32 << c.a1) - 1);
        add_fill(off, bytes, ZQ_FILL_SSE, c.a3); off = align256(off + bytes);
        lv = 1 + level[c.a2];
        break;
      }
      default: throw Error("unknown component type");
    }
    level[i] = lv; c.level = (uint8_t)lv;
    p.nlevels = std::max(p.nlevels, lv + 1);
    cp += len;
  }
  if (p.n > ZQ_CM_LANES) {   // one lane evaluates such a model component by component: its per-component registers
    p.wide_off = off;
    add_fill(off, (uint64_t)p.n * 64, ZQ_FILL_ZERO, 0); off = align256(off + (uint64_t)p.n * 64);
  }
  p.model_bytes = align256(off);
  p.fill_count = (uint32_t)fills.size() - p.fill_first;
  p.chain = (p.n == 2 && p.comp[0].type == ZQ_ICM && p.comp[1].type == ZQ_ISSE && p.comp[1].a2 == 0) ? 1u : 0u;
  return p;
}

void add_pcomp_region(ZqCmPlan& p, int ph, int pm, std::vector<ZqCmFill>& fills) {
  if (ph > 26 || pm > 30) throw Error("PCOMP memory too large for the device path");
  p.ph = ph; p.pm = pm;
  uint64_t off = p.model_bytes;
  const uint64_t begin = off;
  p.pm_off = off; off = align256(off + ((uint64_t)1 << pm));
  p.ph_off = off; off = align256(off + ((uint64_t)4 << ph));
  p.pr_off = off; off = align256(off + 1024);
  p.pcode_off = off; off = align256(off + 65536 + 512);
  ZqCmFill f; f.off = begin; f.bytes = off - begin; f.kind = ZQ_FILL_ZERO; f.value = 0;
  fills.push_back(f);
  p.model_bytes = off;
  p.fill_count = (uint32_t)fills.size() - p.fill_first;
}

Assembled parse_block_header(const uint8_t* hdr, size_t avail, size_t* consumed) {
  Assembled a;
  if (avail < 8) throw Error("unexpected end of file");
  const size_t hsize = hdr[0] + 256u * hdr[1];
  if (avail < hsize + 2) throw Error("unexpected end of file");
  a.hh = hdr[2]; a.hm = hdr[3]; a.ph = hdr[4]; a.pm = hdr[5]; a.ncomp = hdr[6];
  size_t p = 7;
  for (int i = 0; i < a.ncomp; ++i) {
    if (p >= hsize + 2) throw Error("COMP overflows header");
    const int type = hdr[p];
    if (type < 1 || type > 9) throw Error("Invalid component type");
    const int sz = kCompBytes[type];
    if (p + sz > hsize + 2) throw Error("COMP overflows header");
    a.comp.insert(a.comp.end(), hdr + p, hdr + p + sz);
    p += sz;
  }
  if (p >= hsize + 2 || hdr[p] != 0) throw Error("missing COMP END");
  ++p;
  if (hsize + 2 <= p) throw Error("missing HCOMP");
  a.hcomp.assign(hdr + p, hdr + hsize + 2);
  if (a.hcomp.empty() || a.hcomp.back() != 0) throw Error("missing HCOMP END");
  a.header.assign(hdr, hdr + hsize + 2);
  *consumed = hsize + 2;
  return a;
}

}  // namespace zq
