// zq_cm_wide.cuh -- models of 33..64 components (ZQ_CM_MAXCOMP).
//
// The coder warp of zq_cm.cuh maps component i to lane i, which ends at 32.  The built-in models stay below that
// except one: method 5 on text with two periodic models detected (33 components; log files, CSV, fixed-width
// records).  Such blocks are coded by ONE lane that evaluates the components in index order -- literally
// Predictor::predict0 / update0 (Z:15041-15137, Z:15139-15251): index order is a valid evaluation order because
// every input has a lower index.  It is the same arithmetic as the lane code (cm_predict / cm_update), statement by
// statement, over the same tables in the block's model region, so k_cm_init applies unchanged; the per-component
// registers of the lane code live in 64 B of the model region per component (ZqCmPlan::wide_off).  Slow by
// design (no parallelism inside the block; blocks still run side by side): a correct device path instead of
// ZQ_E_UNSUPPORTED for the rare wide model, not a fast one.
#pragma once

namespace zqdev {

struct CmWideComp {   // what CmLane keeps in registers, per component
  int p; u32 cxt, pn; int w0, w1;
  u32 ca, cb, cc, pos;          // MATCH: length, offset, predicted bit, write position
  u32 rowpos, rowok, pad;
  u8 row[16];                   // ICM/ISSE: the current hash row
};

struct CmWide {
  const ZqCmPlan* cp; u8* model; CmWideComp* st;
  int n, c8, hmap4;
};

__device__ __noinline__ void cmw_setup(CmWide& W, const ZqCmPlan& cp, u8* model) {
  W.cp = &cp; W.model = model; W.st = (CmWideComp*)(model + cp.wide_off);
  W.n = cp.n; W.c8 = 1; W.hmap4 = 1;
  for (int i = 0; i < W.n; ++i)   // (the region itself was zeroed by k_cm_init)
    if (cp.comp[i].type == ZQ_CONS) W.st[i].p = ((int)cp.comp[i].a1 - 128) * 4;
}

// Predictor::find (Z:15254) + switch of the cached row; see cm_row_switch
__device__ __forceinline__ void cmw_row_switch(CmWideComp& s, u8* ht, u32 ht_mask, u32 chkshift, u32 cxt) {
  uint4* rc = (uint4*)s.row;
  if (s.rowok) *(uint4*)(ht + s.rowpos) = *rc;
  const u32 chk = (cxt >> chkshift) & 255u;
  const u32 h0 = (cxt * 16u) & (ht_mask - 15u), h1 = h0 ^ 16u, h2 = h0 ^ 32u;
  const uint4 r0 = *(const uint4*)(ht + h0), r1 = *(const uint4*)(ht + h1), r2 = *(const uint4*)(ht + h2);
  u32 r; uint4 row;
  if ((r0.x & 255u) == chk) { r = h0; row = r0; }
  else if ((r1.x & 255u) == chk) { r = h1; row = r1; }
  else if ((r2.x & 255u) == chk) { r = h2; row = r2; }
  else {
    const u32 p0 = (r0.x >> 8) & 255u, p1 = (r1.x >> 8) & 255u, p2 = (r2.x >> 8) & 255u;
    if (p0 <= p1 && p0 <= p2) r = h0; else if (p1 < p2) r = h1; else r = h2;
    row = make_uint4(chk, 0, 0, 0);
  }
  *rc = row;
  s.rowpos = r; s.rowok = 1;
}

// p for the next bit.  h: the context machine's H[] (component i reads h[i & hmask]).
__device__ __noinline__ int cmw_predict(CmWide& W, const u32* h, u32 hmask, const CmSmem& T) {
  const int c8 = W.c8, hmap4 = W.hmap4;
  const bool nib = c8 == 1 || (c8 & 0xf0) == 16;
  for (int i = 0; i < W.n; ++i) {
    const ZqCmComp& c = W.cp->comp[i];
    CmWideComp& s = W.st[i];
    u32* cm = (u32*)(W.model + c.cm_off);
    u8* ht = W.model + c.ht_off;
    const u32 hi = h[(u32)i & hmask];
    switch (c.type) {
      case ZQ_CM:
        s.cxt = (hi ^ (u32)hmap4) & c.cm_mask;
        s.pn = cm[s.cxt];
        s.p = T.stretch[s.pn >> 17];
        break;
      case ZQ_ICM:
        if (nib) cmw_row_switch(s, ht, c.ht_mask, (u32)c.a1 + 2, hi + 16u * (u32)c8);
        s.cxt = s.row[hmap4 & 15];
        s.pn = cm[s.cxt];
        s.p = T.stretch[s.pn >> 8];
        break;
      case ZQ_ISSE: {
        if (nib) cmw_row_switch(s, ht, c.ht_mask, (u32)c.a1 + 2, hi + 16u * (u32)c8);
        s.cxt = s.row[hmap4 & 15];
        const int2 w = *(const int2*)(cm + s.cxt * 2);
        s.w0 = w.x; s.w1 = w.y;
        s.p = cm_clamp2k((s.w0 * W.st[c.a2].p + s.w1 * 64) >> 16);
        break;
      }
      case ZQ_MATCH:
        if (s.ca == 0) s.p = 0;
        else {
          s.cc = (ht[(s.pos - s.cb) & c.ht_mask] >> (7 - s.cxt)) & 1u;
          s.p = T.stretch[(T.dt2k[s.ca] * (s.cc ? -1 : 1)) & 32767];
        }
        break;
      case ZQ_AVG:
        s.p = (W.st[c.a1].p * (int)c.a3 + W.st[c.a2].p * (256 - (int)c.a3)) >> 8;
        break;
      case ZQ_MIX2:
        s.cxt = (hi + ((u32)c8 & c.a5)) & c.cm_mask;
        s.w0 = ((const u16*)cm)[s.cxt];
        s.p = (s.w0 * W.st[c.a2].p + (65536 - s.w0) * W.st[c.a3].p) >> 16;
        break;
      case ZQ_MIX: {
        s.cxt = ((hi + ((u32)c8 & c.a5)) & c.cm_mask) * c.a3;
        const int* wt = (const int*)cm + s.cxt;
        int sum = 0;
        for (u32 j = 0; j < c.a3; ++j) sum += (wt[j] >> 8) * W.st[c.a2 + j].p;
        s.p = cm_clamp2k(sum >> 8);
        break;
      }
      case ZQ_SSE: {
        u32 cx = (hi + (u32)c8) * 32u;
        int pq = min(max(W.st[c.a2].p + 992, 0), 1983);
        const int wt = pq & 63;
        pq >>= 6;
        cx += (u32)pq;
        const u32 e0 = cm[cx & c.cm_mask], e1 = cm[(cx + 1) & c.cm_mask];
        s.p = T.stretch[((e0 >> 10) * (u32)(64 - wt) + (e1 >> 10) * (u32)wt) >> 13];
        cx += (u32)(wt >> 5);
        s.cxt = cx & c.cm_mask;
        s.pn = (wt >> 5) ? e1 : e0;
        break;
      }
      default: break;   // CONS: set once
    }
  }
  return T.squash[W.st[W.n - 1].p + 2048];
}

// train on bit y; true when a byte was completed
__device__ __noinline__ bool cmw_update(CmWide& W, const u32* h, u32 hmask, const CmSmem& T, int y) {
  for (int i = 0; i < W.n; ++i) {
    const ZqCmComp& c = W.cp->comp[i];
    CmWideComp& s = W.st[i];
    u32* cm = (u32*)(W.model + c.cm_off);
    u8* ht = W.model + c.ht_off;
    switch (c.type) {
      case ZQ_CM: case ZQ_SSE: {
        const u32 limit = c.type == ZQ_CM ? c.a2 * 4u : c.a4 * 4u;
        const u32 count = s.pn & 0x3ffu;
        const int error = y * 32767 - (int)(s.pn >> 17);
        s.pn += (u32)((error * T.dt[count]) & -1024) + (count < limit ? 1u : 0u);
        cm[s.cxt] = s.pn;
        break;
      }
      case ZQ_ICM:
        s.row[W.hmap4 & 15] = T.ns[s.cxt * 4 + y];
        s.pn += (u32)(((int)(y * 32767 - (int)(s.pn >> 8))) >> 2);
        cm[s.cxt] = s.pn;
        break;
      case ZQ_ISSE: {
        s.row[W.hmap4 & 15] = T.ns[s.cxt * 4 + y];
        const int err = y * 32767 - (int)T.squash[s.p + 2048];
        int2 w;
        w.x = cm_clamp512k(s.w0 + ((err * W.st[c.a2].p + (1 << 12)) >> 13));
        w.y = cm_clamp512k(s.w1 + ((err + 16) >> 5));
        *(int2*)(cm + s.cxt * 2) = w;
        break;
      }
      case ZQ_MIX2: {
        const int err = ((y * 32767 - (int)T.squash[s.p + 2048]) * (int)c.a4) >> 5;
        int w = s.w0 + ((err * (W.st[c.a2].p - W.st[c.a3].p) + (1 << 12)) >> 13);
        w = min(max(w, 0), 65535);
        ((u16*)cm)[s.cxt] = (u16)w;
        break;
      }
      case ZQ_MIX: {
        const int err = ((y * 32767 - (int)T.squash[s.p + 2048]) * (int)c.a4) >> 4;
        int* wt = (int*)cm + s.cxt;
        for (u32 j = 0; j < c.a3; ++j) wt[j] = cm_clamp512k(wt[j] + ((err * W.st[c.a2 + j].p + (1 << 12)) >> 13));
        break;
      }
      case ZQ_MATCH: {
        const u32 hm = c.ht_mask;
        if ((int)s.cc != y) s.ca = 0;
        if (++s.cxt == 8) {
          ht[s.pos & hm] = (u8)(W.c8 * 2 + y);
          s.cxt = 0;
          s.pos = (s.pos + 1) & hm;
          const u32 hi = h[(u32)i & hmask];
          if (s.ca == 0) {
            s.cb = s.pos - cm[hi & c.cm_mask];
            if (s.cb & hm)
              while (s.ca < 255 && ht[(s.pos - s.ca - 1) & hm] == ht[(s.pos - s.ca - s.cb - 1) & hm]) ++s.ca;
          } else s.ca += s.ca < 255;
          cm[hi & c.cm_mask] = s.pos;
        }
        break;
      }
      default: break;
    }
  }
  W.c8 += W.c8 + y;
  if (W.c8 >= 256) { W.hmap4 = 1; W.c8 = 1; return true; }
  if (W.c8 >= 16 && W.c8 < 32) W.hmap4 = (W.hmap4 & 0xf) << 5 | y << 4 | 1;
  else W.hmap4 = (W.hmap4 & 0x1f0) | (((W.hmap4 & 0xf) * 2 + y) & 0xf);
  return false;
}

}  // namespace zqdev
