// zq_archive.cpp -- host side of the archiver's add loop: dedup index, data-type heuristics, the "new block"
// rule and the data / fragment-table blocks, so that a front end feeding whole files gets the same "d" and "h"
// blocks the reference writes.  The per-byte work runs on the device:
// zq_fragment_ex (boundaries, hits, SHA-1, order-1 tables) and zq_compress_blocks (every block).  What is left
// here is a few hundred integer operations per fragment, restated from Jidac::add:
//   file order key        Z:121735-121754, compareFilename Z:63575
//   per-fragment analysis Z:122592-122638 (text / exe / redundancy estimate from the order-1 table)
//   new-block rule        Z:122640-122664
//   block payload, method Z:122687-122704  ("<method>,<redundancy>,<type>", name "jDC<date>d<first id>")
//   bookkeeping           Z:122774-122795
//   fragment tables       Z:122880-122911  (name "jDC<date>h<first id>", method "0")
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/zq_b200.h"

namespace {

const int kTables = 4;   // order-1 tables of the last fragments kept for comparison (ON, Z:121952)

std::string digits(uint64_t x, int n) {   // itos(x, n): zero padded decimal (Z:29411)
  std::string r;
  for (; x || n > 0; x /= 10, --n) r.insert(r.begin(), (char)('0' + x % 10));
  return r;
}
void put_le32(std::vector<uint8_t>& v, uint64_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }

struct Sha1Key {
  uint8_t b[20];
  bool operator==(const Sha1Key& o) const { return memcmp(b, o.b, 20) == 0; }
};

// What one fragment tells about its data (Z:122592-122638): updates `hits` in place.
void analyse_fragment(const uint8_t* o1, int64_t sz, const uint8_t* o1prev, unsigned& hits, int& text1, int& exe1) {
  int text = 0, exe = 0;
  int64_t h1 = sz;
  uint8_t seen[256] = {0};   // how often each predicted byte occurs in the table
  for (int i = 0; i < 256; ++i) {
    uint8_t& k = seen[o1[i]];
    if (k < 255) { const int w = k < 160 ? 160 / (k + 1) : 0; h1 -= (sz * w) >> 15; ++k; }
    if (o1[i] == ' ' && (isalnum(i) || i == '.' || i == ',')) ++text;
    if (o1[i] && (i < 9 || i == 11 || i == 12 || (i >= 14 && i <= 31) || i >= 240)) --text;
    if (i >= 192 && i < 240 && o1[i] && (o1[i] < 128 || o1[i] >= 192)) --text;
    if (o1[i] == 139) ++exe;
  }
  text1 = text >= 3;
  exe1 = exe >= 5;
  if (sz > 0) h1 = h1 * h1 / sz;
  unsigned h2 = (unsigned)h1;
  if (h2 > hits) hits = h2;
  h2 = (unsigned)((int64_t)seen[0] * sz / 256);
  if (h2 > hits) hits = h2;
  unsigned same = 0;
  for (int i = 0; i < 256 * kTables; ++i) same += o1prev[i] == o1[i & 255];
  h2 = (unsigned)((int64_t)same * sz / (256 * kTables));
  if (h2 > hits) hits = h2;
  if ((int64_t)hits > sz) hits = (unsigned)sz;
}

}  // namespace

extern "C" {

uint64_t zq_file_sort_key(const char* path, int64_t size) {
  // first bytes of the extension, case folded, then descending size in 16 KiB steps (Z:121735-121754)
  uint64_t key = 0;
  int sp = 0;
  for (const char* q = path; q && *q; ++q) {
    uint64_t ch = (unsigned char)*q;
    if (ch >= 'A' && ch <= 'Z') ch += 'a' - 'A';
    if (ch == '/') sp = 0, key = 0;
    else if (ch == '.') sp = 8, key = 0;
    else if (sp > 3) key += ch << (--sp * 8);
  }
  int64_t s = size >> 14;
  if (s >= (1 << 24)) s = (1 << 24) - 1;
  return key + (uint64_t)((1 << 24) - s - 1);
}

int zq_add_files(zq_ctx* ctx, int nfiles, const uint8_t* base, const uint64_t* off, const uint64_t* len,
                 const char* method_in, int fragment, const char* date14, uint32_t first_id,
                 uint8_t* d_out, uint64_t d_cap, uint64_t* d_len, uint8_t* h_out, uint64_t h_cap, uint64_t* h_len,
                 uint32_t* file_frags, uint64_t file_frags_cap, uint64_t* file_first, uint32_t* nblocks_out) {
  if (!ctx) return ZQ_E_NODEVICE;
  if (nfiles < 0 || (nfiles > 0 && (!base || !off || !len)) || !date14 || strlen(date14) != 14 || !d_len || !h_len || !file_first ||
      first_id < 1)
    return ZQ_E_ARG;
  // -method defaults (Z:121603-121611)
  std::string method = method_in && *method_in ? method_in : "1";
  if (method.size() == 1) method += (method[0] >= '2' && method[0] <= '9') ? "6" : "4";
  if (!strchr("0123456789xs", method[0])) return ZQ_E_METHOD;
  if (fragment < 0) fragment = 0;
  const int log_blocksize = 20 + atoi(method.c_str() + 1);
  if (log_blocksize < 20 || log_blocksize > 31) return ZQ_E_METHOD;
  const uint32_t blocksize = (1u << log_blocksize) - 4096;
  const uint32_t max_fragment = (fragment > 19 || ((uint64_t)8128 << fragment) > blocksize - 12) ? blocksize - 12 : 8128u << fragment;
  const uint32_t min_fragment = (fragment > 25 || ((uint64_t)64 << fragment) > max_fragment) ? max_fragment : 64u << fragment;

  // ---- device: fragment boundaries, hits, SHA-1 and order-1 tables of every file ----------------------
  uint64_t total = 0;
  for (int f = 0; f < nfiles; ++f) total += len[f];
  const uint64_t cap = total / std::max<uint32_t>(min_fragment, 1) + (uint64_t)nfiles + 16;
  std::vector<uint32_t> flen(cap), fhits(cap);
  std::vector<uint8_t> fsha(cap * 20), fo1(cap * 256);
  std::vector<uint64_t> ffirst(nfiles + 1, 0);
  int rc = zq_fragment_ex(ctx, nfiles, base, off, len, fragment, blocksize, flen.data(), fhits.data(), fsha.data(), fo1.data(), cap, ffirst.data());
  if (rc != ZQ_OK) return rc;
  uint8_t sha_empty[20];
  { const uint64_t z = 0; static const uint8_t none = 0; rc = zq_sha1(ctx, 1, &none, &z, &z, sha_empty); if (rc != ZQ_OK) return rc; }

  // ---- device: the fragment index (first fragment with the same digest) ------------------------------
  const uint64_t nfrag = ffirst[nfiles];
  std::vector<uint32_t> first_same(nfrag + 1), frag_id(nfrag + 1, 0);
  rc = zq_dedup_first(ctx, nfrag, fsha.data(), first_same.data());
  if (rc != ZQ_OK) return rc;
  uint32_t empty_id = 0;                // the one fragment of size 0 (only empty files have it)

  // ---- host: block assembly ---------------------------------------------------------------------------
  struct Frag { Sha1Key sha; uint32_t usize; };
  std::vector<Frag> ht;                 // new fragments; id = first_id + position
  std::vector<uint8_t> arena;           // block payloads back to back
  std::vector<uint64_t> boff; std::vector<uint32_t> blen, bfirst;   // per block: payload range, first fragment id
  std::vector<std::string> bmethod, bname;
  std::vector<uint8_t> sb;              // the block being filled
  unsigned frags = 0, redundancy = 0, text = 0, exe = 0;
  uint8_t o1prev[256 * kTables] = {0};
  static const uint8_t zero_table[256] = {0};
  uint64_t nptr = 0;

  for (int fi = 0; fi <= nfiles; ++fi) {
    if (fi < nfiles) file_first[fi] = nptr;
    const uint64_t nf = fi < nfiles ? ffirst[fi + 1] - ffirst[fi] : 0;
    uint64_t pos = fi < nfiles ? off[fi] : 0;   // where the next fragment of this file starts
    for (uint64_t fj = 0; true; ++fj) {
      int64_t sz = 0;
      unsigned hits = 0;
      uint32_t id = 0;
      const uint8_t* o1 = zero_table;
      const uint8_t* data = nullptr;
      Sha1Key key;
      if (fi < nfiles) {
        if (fj < nf) {
          const uint64_t g = ffirst[fi] + fj;
          sz = flen[g]; hits = fhits[g]; o1 = &fo1[g * 256]; memcpy(key.b, &fsha[g * 20], 20);
          data = base + pos;
          pos += flen[g];
        } else if (fj > 0) break;                      // end of file
        else memcpy(key.b, sha_empty, 20);             // an empty file is one empty fragment (Z:122562)
        if (fj < nf) { const uint64_t g = ffirst[fi] + fj; if (first_same[g] != g) id = frag_id[first_same[g]]; }
        else id = empty_id;
      }
      if (id == 0) {
        int text1 = 0, exe1 = 0;
        analyse_fragment(o1, sz, o1prev, hits, text1, exe1);
        bool newblock = false;
        if (frags > 0 && fj == 0 && fi < nfiles) {
          const int64_t esize = (int64_t)len[fi];
          const int64_t newsize = (int64_t)sb.size() + esize + (esize >> 14) + 4096 + (int64_t)frags * 4;
          if (newsize > (int64_t)(blocksize / 4) && redundancy < sb.size() / 128) newblock = true;
          if (newblock) {   // ... unless the file looks like what the block already holds
            unsigned ct = 0;
            for (int i = 0; i < 256 * kTables; ++i) if (o1prev[i] && o1prev[i] == o1[i & 255]) ++ct;
            if (ct > (unsigned)kTables * 2) newblock = false;
          }
          if (newsize >= (int64_t)blocksize) newblock = true;
        }
        if (sb.size() + (uint64_t)sz + 80 + (uint64_t)frags * 4 >= blocksize) newblock = true;
        if (fi == nfiles) newblock = true;
        if (frags < 1) newblock = false;
        if (newblock) {
          const uint32_t first = first_id + (uint32_t)ht.size() - frags;
          for (size_t i = ht.size() - frags; i < ht.size(); ++i) put_le32(sb, ht[i].usize);
          put_le32(sb, 0);
          put_le32(sb, frags);
          std::string m = method;
          if (isdigit((unsigned char)method[0])) {
            const unsigned redz = (unsigned)(redundancy / (sb.size() / 256 + 1));
            m += "," + digits(redz, 1) + "," + digits((exe > frags) * 2 + (text > frags), 1);
          }
          boff.push_back(arena.size()); blen.push_back((uint32_t)sb.size()); bfirst.push_back(first);
          bmethod.push_back(m); bname.push_back("jDC" + std::string(date14) + "d" + digits(first, 10));
          arena.insert(arena.end(), sb.begin(), sb.end());
          arena.resize((arena.size() + 15) & ~(size_t)15);
          sb.clear();
          frags = redundancy = text = exe = 0;
          memset(o1prev, 0, sizeof o1prev);
        }
        if (sz) sb.insert(sb.end(), data, data + sz);
        ++frags;
        redundancy += hits;
        exe += exe1 * 4;
        text += text1 * 2;
        if (sz >= (int64_t)min_fragment) {
          memmove(o1prev, o1prev + 256, 256 * (kTables - 1));
          memcpy(o1prev + 256 * (kTables - 1), o1, 256);
        }
      }
      if (fi < nfiles) {
        if (id == 0) {
          id = first_id + (uint32_t)ht.size();
          Frag fr; fr.sha = key; fr.usize = (uint32_t)sz;
          ht.push_back(fr);
          if (fj >= nf) empty_id = id;
        }
        if (fj < nf) frag_id[ffirst[fi] + fj] = id;
        if (file_frags) { if (nptr >= file_frags_cap) return ZQ_E_OUTPUT; file_frags[nptr] = id; }
        ++nptr;
      }
      if (sz == 0) break;
    }
  }
  file_first[nfiles] = nptr;

  // ---- device: every data block in one batch ("jDC\x01" comment, checksum on; compressThread Z:71422) ---
  const int nb = (int)boff.size();
  if (nblocks_out) *nblocks_out = (uint32_t)nb;
  *d_len = 0; *h_len = 0;
  if (nb == 0) return ZQ_OK;
  std::vector<const char*> mp(nb), np(nb), cp(nb, "jDC\x01");
  for (int i = 0; i < nb; ++i) { mp[i] = bmethod[i].c_str(); np[i] = bname[i].c_str(); }
  std::vector<uint64_t> ooff(nb); std::vector<uint32_t> olen(nb);
  arena.resize(arena.size() + 16);
  rc = zq_compress_blocks(ctx, nb, arena.data(), boff.data(), blen.data(), mp.data(), np.data(), cp.data(), 0, 1, d_out, d_cap, ooff.data(), olen.data());
  if (rc != ZQ_OK) return rc;
  *d_len = ooff[nb - 1] + olen[nb - 1];

  // ---- fragment tables: compressed size, then (SHA-1, size) per fragment, stored with method "0" ----------
  std::vector<uint8_t> tarena; std::vector<uint64_t> toff(nb); std::vector<uint32_t> tlen(nb);
  std::vector<std::string> tname(nb);
  for (int i = 0; i < nb; ++i) {
    toff[i] = tarena.size();
    put_le32(tarena, olen[i]);
    const uint32_t a = bfirst[i] - first_id, b = (i + 1 < nb ? bfirst[i + 1] : first_id + (uint32_t)ht.size()) - first_id;
    for (uint32_t j = a; j < b; ++j) { tarena.insert(tarena.end(), ht[j].sha.b, ht[j].sha.b + 20); put_le32(tarena, ht[j].usize); }
    tlen[i] = (uint32_t)(tarena.size() - toff[i]);
    tname[i] = "jDC" + std::string(date14) + "h" + digits(bfirst[i], 10);
    tarena.resize((tarena.size() + 15) & ~(size_t)15);
  }
  tarena.resize(tarena.size() + 16);
  for (int i = 0; i < nb; ++i) { np[i] = tname[i].c_str(); mp[i] = "0"; }
  rc = zq_compress_blocks(ctx, nb, tarena.data(), toff.data(), tlen.data(), mp.data(), np.data(), cp.data(), 0, 1, h_out, h_cap, ooff.data(), olen.data());
  if (rc != ZQ_OK) return rc;
  *h_len = ooff[nb - 1] + olen[nb - 1];
  return ZQ_OK;
}

int zq_journal_header(zq_ctx* ctx, const char* date14, int64_t cdata, uint32_t htsize, uint8_t* out, uint64_t cap, uint64_t* len) {
  // the transaction header: 8 bytes (size of the data blocks that follow, or -1 while the update is open), stored,
  // named jDC<date>c<first fragment id> (writeJidacHeader, Z:71521-71540)
  if (!ctx) return ZQ_E_NODEVICE;
  if (!date14 || strlen(date14) != 14 || !out || !len) return ZQ_E_ARG;
  uint8_t payload[32] = {0};
  for (int i = 0; i < 8; ++i) payload[i] = (uint8_t)((uint64_t)cdata >> (8 * i));
  const uint64_t off = 0; const uint32_t n = 8; uint64_t ooff = 0; uint32_t olen = 0;
  const std::string name = "jDC" + std::string(date14) + "c" + digits(htsize, 10);
  const char* m = "0"; const char* nm = name.c_str(); const char* cm = "jDC\x01";
  const int rc = zq_compress_blocks(ctx, 1, payload, &off, &n, &m, &nm, &cm, 0, 1, out, cap, &ooff, &olen);
  if (rc != ZQ_OK) return rc;
  *len = ooff + olen;
  return ZQ_OK;
}

int zq_journal_index(zq_ctx* ctx, const char* date14, int nrec, const int64_t* date, const char* const* name,
                     const uint8_t* attr_base, const uint64_t* attr_off, const uint32_t* attr_len,
                     const uint64_t* frag_first, const uint32_t* frags,
                     uint8_t* out, uint64_t cap, uint64_t* len, uint32_t* nblocks) {
  // the index of one transaction: per record the date, the name, and for a live file its attribute bytes and fragment
  // ids (deletions carry date 0 and nothing else); a block is closed once it passes 16 000 bytes and all of them are
  // compressed in one batch with method "1" as jDC<date>i<count> (Z:122915-123100)
  if (!ctx) return ZQ_E_NODEVICE;
  if (!date14 || strlen(date14) != 14 || nrec < 0 || !len || (nrec > 0 && (!date || !name || !out))) return ZQ_E_ARG;
  std::vector<uint8_t> arena; std::vector<uint64_t> boff; std::vector<uint32_t> blen;
  std::vector<uint8_t> is;
  auto close_block = [&]() {
    boff.push_back(arena.size()); blen.push_back((uint32_t)is.size());
    arena.insert(arena.end(), is.begin(), is.end());
    arena.resize((arena.size() + 15) & ~(size_t)15);
    is.clear();
  };
  for (int r = 0; r < nrec; ++r) {
    if (!name[r]) return ZQ_E_ARG;
    const size_t nl = strlen(name[r]);
    if (nl > 65535) return ZQ_E_ARG;
    for (int i = 0; i < 8; ++i) is.push_back((uint8_t)((uint64_t)date[r] >> (8 * i)));
    is.insert(is.end(), name[r], name[r] + nl + 1);
    if (date[r]) {
      const uint32_t na = attr_len ? attr_len[r] : 0;
      if (na > 65535 || (na && (!attr_base || !attr_off))) return ZQ_E_ARG;
      put_le32(is, na);
      if (na) is.insert(is.end(), attr_base + attr_off[r], attr_base + attr_off[r] + na);
      const uint64_t a = frag_first ? frag_first[r] : 0, b = frag_first ? frag_first[r + 1] : 0;
      if (b < a || (b > a && !frags)) return ZQ_E_ARG;
      put_le32(is, b - a);
      for (uint64_t j = a; j < b; ++j) put_le32(is, frags[j]);
    }
    if (is.size() > 16000) close_block();
  }
  if (!is.empty()) close_block();
  const int nb = (int)boff.size();
  if (nblocks) *nblocks = (uint32_t)nb;
  *len = 0;
  if (nb == 0) return ZQ_OK;
  std::vector<std::string> names(nb);
  std::vector<const char*> mp(nb, "1"), np(nb), cp(nb, "jDC\x01");
  for (int i = 0; i < nb; ++i) { names[i] = "jDC" + std::string(date14) + "i" + digits((uint64_t)i + 1, 10); np[i] = names[i].c_str(); }
  std::vector<uint64_t> ooff(nb); std::vector<uint32_t> olen(nb);
  arena.resize(arena.size() + 16);
  const int rc = zq_compress_blocks(ctx, nb, arena.data(), boff.data(), blen.data(), mp.data(), np.data(), cp.data(), 0, 1, out, cap, ooff.data(), olen.data());
  if (rc != ZQ_OK) return rc;
  *len = ooff[nb - 1] + olen[nb - 1];
  return ZQ_OK;
}

}  // extern "C"
