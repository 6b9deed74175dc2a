/* oracle/zq_oracle_cm.c -- TEST INFRASTRUCTURE ONLY (see zq_oracle.c).
 *
 * Plain-C restatement of the context-mixing half of libzpaq's block compressor:
 *
 *   ZPAQL virtual machine (HCOMP)      ZPAQL::run0/execute    Z:14232-14467
 *   model-independent tables           Predictor::init        Z:14926-14941 (formulas Z:61905-61909)
 *   bit-history state table            StateTable             Z:61750-61850 (the ZPAQ spec's generator)
 *   component init / predict / update  Predictor::*0, find    Z:14956-15270, train Z:13161
 *   arithmetic coder                   Encoder::encode/compress Z:15557-15589
 *   modeled block framing              Compressor::*          Z:15970-16187
 *
 * Pinned against oracle/_ref (the reference compiled here) by tests/test_oracle_pinned.py, and its
 * tables against the reference's own checksums 3887533746 / 2278286169 (Z:14949-14950).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

enum { T_CONS = 1, T_CM, T_ICM, T_MATCH, T_AVG, T_MIX2, T_MIX, T_ISSE, T_SSE };
static const int compsize[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};

/* ------------------------------------------------------------------ tables */
static int g_tables;
static int dt2k[256], dtab[1024];
static u16 squasht[4096];
static short stretcht[32768];
static u8 ns[1024];

static int nstates(int n0, int n1) {
  static const int bound[6] = {20, 48, 15, 8, 6, 5};
  if (n0 < n1) { int t = n0; n0 = n1; n1 = t; }
  if (n0 < 0 || n1 < 0 || n1 >= 6 || n0 > bound[n1]) return 0;
  return 1 + (n1 > 0 && n0 + n1 <= 17);
}
static void nxt(int* n0, int* n1, int y) {
  if (*n0 < *n1) { nxt(n1, n0, 1 - y); return; }
  int* inc = y ? n1 : n0; int* dis = y ? n0 : n1;
  ++*inc;
  { int v = *dis; *dis = (v >= 1) + (v >= 2) + (v >= 3) + (v >= 4) + (v >= 5) + (v >= 7) + (v >= 8); }
  while (!nstates(*n0, *n1)) {
    if (*n1 < 2) --*n0;
    else { *n0 = (*n0 * (*n1 - 1) + (*n1 / 2)) / *n1; --*n1; }
  }
}
static void make_tables(void) {
  if (g_tables) return;
  g_tables = 1;
  for (int i = 1; i < 256; ++i) dt2k[i] = 2048 / i;
  for (int i = 0; i < 1024; ++i) dtab[i] = (1 << 17) / (i * 2 + 3) * 2;
  for (int i = 0; i < 32768; ++i) stretcht[i] = (short)((int)(log((i + 0.5) / (32767.5 - i)) * 64 + 0.5 + 100000) - 100000);
  for (int i = 0; i < 4096; ++i) squasht[i] = (u16)(int)(32768.0 / (1 + exp((i - 2048) * (-1.0 / 64))));
  enum { N = 50 };
  static u8 t[N][N][2];
  int state = 0;
  for (int i = 0; i < N; ++i)
    for (int n1 = 0; n1 <= i; ++n1) {
      int n0 = i - n1, k = nstates(n0, n1);
      if (k) { t[n0][n1][0] = state; t[n0][n1][1] = state + k - 1; state += k; }
    }
  memset(ns, 0, sizeof ns);
  for (int n0 = 0; n0 < N; ++n0)
    for (int n1 = 0; n1 < N; ++n1)
      for (int y = 0; y < nstates(n0, n1); ++y) {
        int s = t[n0][n1][y], a = n0, b = n1;
        nxt(&a, &b, 0); ns[s * 4] = t[a][b][0];
        a = n0; b = n1;
        nxt(&a, &b, 1); ns[s * 4 + 1] = t[a][b][1];
        ns[s * 4 + 2] = n0; ns[s * 4 + 3] = n1;
      }
}
static int squash(int x) { return squasht[x + 2048]; }
static int stretch(int x) { return stretcht[x]; }
static int clamp2k(int x) { return x < -2048 ? -2048 : x > 2047 ? 2047 : x; }
static int clamp512k(int x) { return x < -(1 << 19) ? -(1 << 19) : x >= (1 << 19) ? (1 << 19) - 1 : x; }
static int cminit(int s) { return ((ns[s * 4 + 3] * 2 + 1) << 22) / (ns[s * 4 + 2] + ns[s * 4 + 3] + 1); }

/* exported for the tests: table checksums as the reference computes them, and raw tables */
void zqo_table_sums(u32* stsum, u32* sqsum) {
  make_tables();
  u32 a = 0, b = 0;
  for (int i = 32767; i >= 0; --i) a = a * 3 + (u32)stretch(i);
  for (int i = 4095; i >= 0; --i) b = b * 3 + (u32)squash(i - 2048);
  *stsum = a; *sqsum = b;
}
void zqo_state_table(u8* out1024) { make_tables(); memcpy(out1024, ns, 1024); }

/* ------------------------------------------------------------------ ZPAQL VM */
typedef struct {
  const u8* code; int len;   /* program incl. trailing 0 */
  u8* m; u32 msize;          /* M, size power of 2 */
  u32* h; u32 hsize;         /* H */
  u32 r[256];
  u32 a, b, c, d; int f;
  int error;
  u8* out; u64 outlen, outcap;  /* OUT sink (PCOMP only) */
} VM;

static int vm_init(VM* v, const u8* code, int len, int hbits, int mbits) {
  memset(v, 0, sizeof *v);
  v->code = code; v->len = len;
  v->hsize = 1u << hbits; v->msize = 1u << mbits;
  v->h = (u32*)calloc(v->hsize, 4); v->m = (u8*)calloc(v->msize, 1);
  return v->h && v->m ? 0 : -1;
}
static void vm_free(VM* v) { free(v->h); free(v->m); }

static void vm_run(VM* v, u32 input) {
  const u8* P = v->code;
  int pc = 0;
  u32 a = input, b = v->b, c = v->c, d = v->d; int f = v->f;
#define MB v->m[b & (v->msize - 1)]
#define MC v->m[c & (v->msize - 1)]
#define HD v->h[d & (v->hsize - 1)]
  for (;;) {
    if (pc < 0 || pc >= v->len) { v->error = 1; break; }
    const int op = P[pc++];
    if (op == 56) break; /* HALT */
    if (op >= 64 && op < 240 && op != 0) {
      const int src = op & 7, grp = op >> 3;
      u32 x;
      switch (src) { case 0: x = a; break; case 1: x = b; break; case 2: x = c; break; case 3: x = d; break;
                     case 4: x = MB; break; case 5: x = MC; break; case 6: x = HD; break; default: x = P[pc++]; }
      if (grp >= 8 && grp <= 14) {  /* dst = src */
        switch (grp - 8) { case 0: a = x; break; case 1: b = x; break; case 2: c = x; break; case 3: d = x; break;
                           case 4: MB = (u8)x; break; case 5: MC = (u8)x; break; default: HD = x; }
      } else if (grp >= 16 && grp <= 29) {
        switch (grp - 16) {
          case 0: a += x; break; case 1: a -= x; break; case 2: a *= x; break;
          case 3: a = x ? a / x : 0; break; case 4: a = x ? a % x : 0; break;
          case 5: a &= x; break; case 6: a &= ~x; break; case 7: a |= x; break; case 8: a ^= x; break;
          case 9: a <<= (x & 31); break; case 10: a >>= (x & 31); break;
          case 11: f = a == x; break; case 12: f = a < x; break; default: f = a > x;
        }
      } else { v->error = 1; break; }
      continue;
    }
    switch (op) {
      case 1: ++a; break; case 2: --a; break; case 3: a = ~a; break; case 4: a = 0; break;
      case 7: a = v->r[P[pc++]]; break;
      case 8: { u32 t = a; a = b; b = t; } break;
      case 9: ++b; break; case 10: --b; break; case 11: b = ~b; break; case 12: b = 0; break;
      case 15: b = v->r[P[pc++]]; break;
      case 16: { u32 t = a; a = c; c = t; } break;
      case 17: ++c; break; case 18: --c; break; case 19: c = ~c; break; case 20: c = 0; break;
      case 23: c = v->r[P[pc++]]; break;
      case 24: { u32 t = a; a = d; d = t; } break;
      case 25: ++d; break; case 26: --d; break; case 27: d = ~d; break; case 28: d = 0; break;
      case 31: d = v->r[P[pc++]]; break;
      case 32: { u8 t = MB; MB = (u8)a; a = (a & ~255u) | t; } break;
      case 33: ++MB; break; case 34: --MB; break; case 35: MB = ~MB; break; case 36: MB = 0; break;
      case 39: if (f) pc += ((P[pc] + 128) & 255) - 127; else ++pc; break;
      case 40: { u8 t = MC; MC = (u8)a; a = (a & ~255u) | t; } break;
      case 41: ++MC; break; case 42: --MC; break; case 43: MC = ~MC; break; case 44: MC = 0; break;
      case 47: if (!f) pc += ((P[pc] + 128) & 255) - 127; else ++pc; break;
      case 48: { u32 t = HD; HD = a; a = t; } break;
      case 49: ++HD; break; case 50: --HD; break; case 51: HD = ~HD; break; case 52: HD = 0; break;
      case 55: v->r[P[pc++]] = a; break;
      case 57: if (v->out && v->outlen < v->outcap) v->out[v->outlen] = (u8)a; ++v->outlen; break;
      case 59: a = (a + MB + 512) * 773; break;
      case 60: HD = (HD + a + 512) * 773; break;
      case 63: pc += ((P[pc] + 128) & 255) - 127; break;
      case 255: pc = P[pc] + 256 * P[pc + 1]; break;
      default: v->error = 1;
    }
    if (v->error) break;
  }
  v->a = a; v->b = b; v->c = c; v->d = d; v->f = f;
#undef MB
#undef MC
#undef HD
}

/* ------------------------------------------------------------------ predictor */
typedef struct {
  u32 limit, cxt, a, b, c;
  u32* cm; u32 cmsize;   /* power of 2 (or 256 / 512) */
  u8* ht; u32 htsize;
  u16* a16; u32 a16size;
} Comp;

typedef struct {
  int n; const u8* desc;   /* component descriptors */
  int c8, hmap4;
  int p[256]; u32 h[256];
  Comp comp[256];
  VM vm;
} Pred;

static u32 ht_find(Comp* cr, int sizebits, u32 cxt) {
  u8* ht = cr->ht;
  const int chk = (cxt >> sizebits) & 255;
  const u32 h0 = (cxt * 16) & (cr->htsize - 16), h1 = h0 ^ 16, h2 = h0 ^ 32;
  if (ht[h0] == chk) return h0;
  if (ht[h1] == chk) return h1;
  if (ht[h2] == chk) return h2;
  u32 r;
  if (ht[h0 + 1] <= ht[h1 + 1] && ht[h0 + 1] <= ht[h2 + 1]) r = h0;
  else if (ht[h1 + 1] < ht[h2 + 1]) r = h1;
  else r = h2;
  memset(ht + r, 0, 16); ht[r] = (u8)chk;
  return r;
}

static int pred_init(Pred* pr, const u8* header) {
  make_tables();
  memset(pr, 0, sizeof *pr);
  pr->n = header[6]; pr->desc = header + 7; pr->c8 = 1; pr->hmap4 = 1;
  const u8* cp = pr->desc;
  for (int i = 0; i < pr->n; ++i) {
    Comp* cr = &pr->comp[i];
    switch (cp[0]) {
      case T_CONS: pr->p[i] = (cp[1] - 128) * 4; break;
      case T_CM:
        cr->cmsize = 1u << cp[1]; cr->cm = (u32*)malloc(4 * (size_t)cr->cmsize); cr->limit = cp[2] * 4;
        if (!cr->cm) return -1;
        for (u32 j = 0; j < cr->cmsize; ++j) cr->cm[j] = 0x80000000u;
        break;
      case T_ICM:
        cr->limit = 1023; cr->cmsize = 256; cr->cm = (u32*)malloc(1024);
        cr->htsize = 64u << cp[1]; cr->ht = (u8*)calloc(cr->htsize, 1);
        if (!cr->cm || !cr->ht) return -1;
        for (int j = 0; j < 256; ++j) cr->cm[j] = cminit(j);
        break;
      case T_MATCH:
        cr->cmsize = 1u << cp[1]; cr->cm = (u32*)calloc(cr->cmsize, 4);
        cr->htsize = 1u << cp[2]; cr->ht = (u8*)calloc(cr->htsize, 1);
        if (!cr->cm || !cr->ht) return -1;
        cr->ht[0] = 1;
        break;
      case T_AVG: break;
      case T_MIX2:
        cr->c = 1u << cp[1]; cr->a16size = cr->c; cr->a16 = (u16*)malloc(2 * (size_t)cr->c);
        if (!cr->a16) return -1;
        for (u32 j = 0; j < cr->c; ++j) cr->a16[j] = 32768;
        break;
      case T_MIX: {
        const int m = cp[3];
        cr->c = 1u << cp[1]; cr->cmsize = (u32)m << cp[1]; cr->cm = (u32*)malloc(4 * (size_t)cr->cmsize);
        if (!cr->cm) return -1;
        for (u32 j = 0; j < cr->cmsize; ++j) cr->cm[j] = 65536 / m;
        break;
      }
      case T_ISSE:
        cr->htsize = 64u << cp[1]; cr->ht = (u8*)calloc(cr->htsize, 1);
        cr->cmsize = 512; cr->cm = (u32*)malloc(2048);
        if (!cr->cm || !cr->ht) return -1;
        for (int j = 0; j < 256; ++j) { cr->cm[j * 2] = 1 << 15; cr->cm[j * 2 + 1] = (u32)clamp512k(stretch(cminit(j) >> 8) * 1024); }
        break;
      case T_SSE:
        cr->cmsize = 32u << cp[1]; cr->cm = (u32*)malloc(4 * (size_t)cr->cmsize); cr->limit = cp[4] * 4;
        if (!cr->cm) return -1;
        for (u32 j = 0; j < cr->cmsize; ++j) cr->cm[j] = (u32)squash((j & 31) * 64 - 992) << 17 | cp[3];
        break;
      default: return -2;
    }
    cp += compsize[cp[0]];
  }
  return 0;
}
static void pred_free(Pred* pr) {
  for (int i = 0; i < 256; ++i) { free(pr->comp[i].cm); free(pr->comp[i].ht); free(pr->comp[i].a16); }
}

static int pred_predict(Pred* pr) {
  const u8* cp = pr->desc;
  int* p = pr->p; const int c8 = pr->c8, hmap4 = pr->hmap4;
  for (int i = 0; i < pr->n; ++i) {
    Comp* cr = &pr->comp[i];
    switch (cp[0]) {
      case T_CONS: break;
      case T_CM:
        cr->cxt = pr->h[i] ^ hmap4;
        p[i] = stretch(cr->cm[cr->cxt & (cr->cmsize - 1)] >> 17);
        break;
      case T_ICM:
        if (c8 == 1 || (c8 & 0xf0) == 16) cr->c = ht_find(cr, cp[1] + 2, pr->h[i] + 16 * c8);
        cr->cxt = cr->ht[cr->c + (hmap4 & 15)];
        p[i] = stretch(cr->cm[cr->cxt & 255] >> 8);
        break;
      case T_MATCH:
        if (cr->a == 0) p[i] = 0;
        else {
          cr->c = (cr->ht[(cr->limit - cr->b) & (cr->htsize - 1)] >> (7 - cr->cxt)) & 1;
          p[i] = stretch(dt2k[cr->a] * ((int)cr->c * -2 + 1) & 32767);
        }
        break;
      case T_AVG: p[i] = (p[cp[1]] * cp[3] + p[cp[2]] * (256 - cp[3])) >> 8; break;
      case T_MIX2: {
        cr->cxt = (pr->h[i] + (c8 & cp[5])) & (cr->c - 1);
        const int w = cr->a16[cr->cxt];
        p[i] = (w * p[cp[2]] + (65536 - w) * p[cp[3]]) >> 16;
        break;
      }
      case T_MIX: {
        const int m = cp[3];
        cr->cxt = ((pr->h[i] + (c8 & cp[5])) & (cr->c - 1)) * m;
        const int* wt = (const int*)&cr->cm[cr->cxt];
        int s = 0;
        for (int j = 0; j < m; ++j) s += (wt[j] >> 8) * p[cp[2] + j];
        p[i] = clamp2k(s >> 8);
        break;
      }
      case T_ISSE: {
        if (c8 == 1 || (c8 & 0xf0) == 16) cr->c = ht_find(cr, cp[1] + 2, pr->h[i] + 16 * c8);
        cr->cxt = cr->ht[cr->c + (hmap4 & 15)];
        const int* wt = (const int*)&cr->cm[cr->cxt * 2];
        p[i] = clamp2k((wt[0] * p[cp[2]] + wt[1] * 64) >> 16);
        break;
      }
      case T_SSE: {
        cr->cxt = (pr->h[i] + c8) * 32;
        int pq = p[cp[2]] + 992;
        if (pq < 0) pq = 0;
        if (pq > 1983) pq = 1983;
        const int wt = pq & 63;
        pq >>= 6;
        cr->cxt += pq;
        const u32 m = cr->cmsize - 1;
        p[i] = stretch((int)(((cr->cm[cr->cxt & m] >> 10) * (64 - wt) + (cr->cm[(cr->cxt + 1) & m] >> 10) * wt) >> 13));
        cr->cxt += wt >> 5;
        break;
      }
    }
    cp += compsize[cp[0]];
  }
  return squash(p[pr->n - 1]);
}

static void train(Comp* cr, int y) {
  u32* pn = &cr->cm[cr->cxt & (cr->cmsize - 1)];
  const u32 count = *pn & 0x3ff;
  const int error = y * 32767 - (int)(*pn >> 17);
  *pn += (u32)((error * dtab[count]) & -1024) + (count < cr->limit);
}

static void pred_update(Pred* pr, int y) {
  const u8* cp = pr->desc;
  int* p = pr->p;
  for (int i = 0; i < pr->n; ++i) {
    Comp* cr = &pr->comp[i];
    switch (cp[0]) {
      case T_CM: case T_SSE: train(cr, y); break;
      case T_ICM: {
        u8* slot = &cr->ht[cr->c + (pr->hmap4 & 15)];
        *slot = ns[*slot * 4 + y];
        u32* pn = &cr->cm[cr->cxt & 255];
        *pn += (u32)((int)(y * 32767 - (*pn >> 8)) >> 2);
        break;
      }
      case T_MATCH: {
        const u32 hm = cr->htsize - 1;
        if ((int)cr->c != y) cr->a = 0;
        cr->ht[cr->limit & hm] += cr->ht[cr->limit & hm] + y;
        if (++cr->cxt == 8) {
          cr->cxt = 0;
          ++cr->limit;
          cr->limit &= (1u << cp[2]) - 1;
          if (cr->a == 0) {
            cr->b = cr->limit - cr->cm[pr->h[i] & (cr->cmsize - 1)];
            if (cr->b & hm)
              while (cr->a < 255 && cr->ht[(cr->limit - cr->a - 1) & hm] == cr->ht[(cr->limit - cr->a - cr->b - 1) & hm]) ++cr->a;
          } else cr->a += cr->a < 255;
          cr->cm[pr->h[i] & (cr->cmsize - 1)] = cr->limit;
        }
        break;
      }
      case T_MIX2: {
        const int err = (y * 32767 - squash(p[i])) * cp[4] >> 5;
        int w = cr->a16[cr->cxt];
        w += (err * (p[cp[2]] - p[cp[3]]) + (1 << 12)) >> 13;
        if (w < 0) w = 0;
        if (w > 65535) w = 65535;
        cr->a16[cr->cxt] = (u16)w;
        break;
      }
      case T_MIX: {
        const int m = cp[3];
        const int err = (y * 32767 - squash(p[i])) * cp[4] >> 4;
        int* wt = (int*)&cr->cm[cr->cxt];
        for (int j = 0; j < m; ++j) wt[j] = clamp512k(wt[j] + ((err * p[cp[2] + j] + (1 << 12)) >> 13));
        break;
      }
      case T_ISSE: {
        const int err = y * 32767 - squash(p[i]);
        int* wt = (int*)&cr->cm[cr->cxt * 2];
        wt[0] = clamp512k(wt[0] + ((err * p[cp[2]] + (1 << 12)) >> 13));
        wt[1] = clamp512k(wt[1] + ((err + 16) >> 5));
        cr->ht[cr->c + (pr->hmap4 & 15)] = ns[cr->cxt * 4 + y];
        break;
      }
      default: break;
    }
    cp += compsize[cp[0]];
  }
  pr->c8 += pr->c8 + y;
  if (pr->c8 >= 256) {
    vm_run(&pr->vm, pr->c8 - 256);
    pr->hmap4 = 1; pr->c8 = 1;
    for (int i = 0; i < pr->n; ++i) pr->h[i] = pr->vm.h[i & (pr->vm.hsize - 1)];
  } else if (pr->c8 >= 16 && pr->c8 < 32)
    pr->hmap4 = (pr->hmap4 & 0xf) << 5 | y << 4 | 1;
  else
    pr->hmap4 = (pr->hmap4 & 0x1f0) | (((pr->hmap4 & 0xf) * 2 + y) & 0xf);
}

/* ------------------------------------------------------------------ arithmetic coder + framing */
typedef struct { u8* out; u64 cap, len; int overflow; u32 low, high; } Enc;
static void eput(Enc* e, int c) { if (e->len < e->cap) e->out[e->len++] = (u8)c; else e->overflow = 1; }
static void encode_bit(Enc* e, int y, int p) {
  const u32 mid = e->low + (u32)(((u64)(e->high - e->low) * (u32)p) >> 16);
  if (y) e->high = mid; else e->low = mid + 1;
  while ((e->high ^ e->low) < 0x1000000u) {
    eput(e, e->high >> 24);
    e->high = e->high << 8 | 255;
    e->low = e->low << 8;
    e->low += (e->low == 0);
  }
}
static void encode_byte(Enc* e, Pred* pr, int c) {
  encode_bit(e, 0, 0);
  for (int i = 7; i >= 0; --i) {
    const int p = pred_predict(pr) * 2 + 1;
    const int y = (c >> i) & 1;
    encode_bit(e, y, p);
    pred_update(pr, y);
  }
}

/* One complete modeled block (n components > 0):
 * tag | "zPQ" 1 1 | header | 01 filename 00 comment 00 00 | coded(selector/pcomp + stream + EOS) |
 * 00 00 00 00 | FD sha1 / FE | FF.   header = bytes as in the block (hsize .. hcomp 0). */
long long zqo_block_modeled(const u8* header, u32 hlen, const u8* pcomp, u32 plen, const char* filename,
                            const char* comment_full, const u8* stream, u64 slen, const u8* sha1, u8* out, u64 cap) {
  static const u8 tag[13] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3};
  Enc e; memset(&e, 0, sizeof e); e.out = out; e.cap = cap; e.low = 1; e.high = 0xFFFFFFFFu;
  for (int i = 0; i < 13; ++i) eput(&e, tag[i]);
  eput(&e, 'z'); eput(&e, 'P'); eput(&e, 'Q'); eput(&e, 1 + (header[6] == 0)); eput(&e, 1);
  for (u32 i = 0; i < hlen; ++i) eput(&e, header[i]);
  eput(&e, 1);
  for (const char* p = filename; p && *p; ++p) eput(&e, (u8)*p);
  eput(&e, 0);
  for (const char* p = comment_full; p && *p; ++p) eput(&e, (u8)*p);
  eput(&e, 0); eput(&e, 0);
  Pred* pr = (Pred*)malloc(sizeof(Pred));
  if (!pr) return -1;
  int rc = pred_init(pr, header);
  /* HCOMP code follows the component list and its 0 terminator */
  const u8* cp = header + 7;
  for (int i = 0; i < header[6]; ++i) cp += compsize[cp[0]];
  ++cp;
  const int codelen = (int)(header + hlen - cp);
  if (!rc) rc = vm_init(&pr->vm, cp, codelen, header[2], header[3]);
  if (rc) { pred_free(pr); free(pr); return -1; }
  if (plen) {
    encode_byte(&e, pr, 1); encode_byte(&e, pr, plen & 255); encode_byte(&e, pr, plen >> 8);
    for (u32 i = 0; i < plen; ++i) encode_byte(&e, pr, pcomp[i]);
  } else encode_byte(&e, pr, 0);
  for (u64 i = 0; i < slen; ++i) encode_byte(&e, pr, stream[i]);
  encode_bit(&e, 1, 0);  /* EOS */
  const int vmerr = pr->vm.error;
  vm_free(&pr->vm); pred_free(pr); free(pr);
  eput(&e, 0); eput(&e, 0); eput(&e, 0); eput(&e, 0);
  if (sha1) { eput(&e, 253); for (int i = 0; i < 20; ++i) eput(&e, sha1[i]); } else eput(&e, 254);
  eput(&e, 255);
  if (vmerr) return -3;
  return e.overflow ? -2 : (long long)e.len;
}
