// libzpaq_b200.cpp -- the C++ mirror of libzpaq's compress-side surface over the C ABI (see
// include/libzpaq_b200.h for the reference lines). No compute here: it packs the inputs into one
// arena, calls zq_compress_blocks and hands the blocks to the Writers.
#include "../../include/libzpaq_b200.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/zq_b200.h"

namespace libzpaq_b200 {

void error(const char* msg) {
  if (strstr(msg, "ut of memory")) throw std::bad_alloc();
  throw std::runtime_error(msg);
}

int Reader::read(char* buf, int n) {
  int i = 0, c;
  while (i < n && (c = get()) >= 0) buf[i++] = (char)c;
  return i;
}
void Writer::write(const char* buf, int n) { for (int i = 0; i < n; ++i) put((unsigned char)buf[i]); }

void StringBuffer::write(const char* b, int n) {
  if (n <= 0) return;
  if (buf_.size() + (size_t)n > limit_) error("StringBuffer overflow");
  buf_.insert(buf_.end(), (const unsigned char*)b, (const unsigned char*)b + n);
}
int StringBuffer::read(char* b, int n) {
  size_t r = buf_.size() - rpos_;
  if (r > (size_t)n) r = n;
  if (r) memcpy(b, buf_.data() + rpos_, r);
  rpos_ += r;
  return (int)r;
}

namespace {
struct ThreadCtx {
  zq_ctx* c = nullptr;
  ~ThreadCtx() { if (c) zq_destroy(c); }
  zq_ctx* get() {
    if (!c) {
      const char* d = getenv("ZQ_DEVICE");
      c = zq_create(d ? atoi(d) : 0);
      if (!c) error(zq_last_error(nullptr));
    }
    return c;
  }
};
thread_local ThreadCtx t_ctx;
}  // namespace

// ---- digests ---------------------------------------------------------------------------------------
// The classes stream: at most kFlush bytes wait on the host.  A stream that never reaches kFlush (every segment the
// archiver hashes here is a block, <= 64 MiB, and most are far smaller) is hashed in one go by the batch kernel at
// result(); a longer one is fed to the device in whole 64-byte blocks (zq_sha*_continue) and padded at the end.
template <int WORDS, class Cont>
static void hash_flush(std::vector<unsigned char>& buf, uint32_t* st, bool& streamed, Cont cont) {
  const size_t whole = buf.size() & ~(size_t)63;
  if (!whole) return;
  zq_ctx* c = t_ctx.get();
  if (cont(c, st, buf.data(), whole / 64) != ZQ_OK) error(zq_last_error(c));
  buf.erase(buf.begin(), buf.begin() + whole);
  streamed = true;
}
template <int WORDS, class Cont>
static void hash_finish(std::vector<unsigned char>& buf, uint64_t total, uint32_t* st, Cont cont, char* out) {
  // FIPS 180-4 padding: 0x80, zeros, the bit length as a big-endian 64-bit number
  buf.push_back(0x80);
  while (buf.size() % 64 != 56) buf.push_back(0);
  for (int i = 7; i >= 0; --i) buf.push_back((unsigned char)((total * 8) >> (8 * i)));
  zq_ctx* c = t_ctx.get();
  if (cont(c, st, buf.data(), buf.size() / 64) != ZQ_OK) error(zq_last_error(c));
  for (int k = 0; k < WORDS; ++k) for (int b = 0; b < 4; ++b) out[4 * k + b] = (char)(st[k] >> (24 - 8 * b));
}

void SHA1::flush() { hash_flush<5>(buf_, st_, streamed_, zq_sha1_continue); }
void SHA256::flush() { hash_flush<8>(buf_, st_, streamed_, zq_sha256_continue); }
void SHA1::write(const char* b, int64_t n) {
  while (n > 0) {
    const size_t k = (size_t)std::min<int64_t>(n, (int64_t)(kFlush - buf_.size()));
    buf_.insert(buf_.end(), (const unsigned char*)b, (const unsigned char*)b + k);
    b += k; n -= (int64_t)k; len_ += k;
    if (buf_.size() >= kFlush) flush();
  }
}
void SHA256::write(const char* b, int64_t n) {
  while (n > 0) {
    const size_t k = (size_t)std::min<int64_t>(n, (int64_t)(kFlush - buf_.size()));
    buf_.insert(buf_.end(), (const unsigned char*)b, (const unsigned char*)b + k);
    b += k; n -= (int64_t)k; len_ += k;
    if (buf_.size() >= kFlush) flush();
  }
}
const char* SHA1::result() {
  zq_ctx* c = t_ctx.get();
  if (streamed_) hash_finish<5>(buf_, len_, st_, zq_sha1_continue, h_);
  else {
    const uint64_t off = 0, len = buf_.size();
    static const unsigned char none = 0;
    if (zq_sha1(c, 1, buf_.empty() ? &none : buf_.data(), &off, &len, (uint8_t*)h_) != ZQ_OK) error(zq_last_error(c));
  }
  static const uint32_t init[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  memcpy(st_, init, sizeof init);
  buf_.clear(); len_ = 0; streamed_ = false;
  return h_;
}
const char* SHA256::result() {
  zq_ctx* c = t_ctx.get();
  if (streamed_) hash_finish<8>(buf_, len_, st_, zq_sha256_continue, h_);
  else {
    const uint64_t off = 0, len = buf_.size();
    static const unsigned char none = 0;
    if (zq_sha256(c, 1, buf_.empty() ? &none : buf_.data(), &off, &len, (uint8_t*)h_) != ZQ_OK) error(zq_last_error(c));
  }
  static const uint32_t init[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
  memcpy(st_, init, sizeof init);
  buf_.clear(); len_ = 0; streamed_ = false;
  return h_;
}

// ---- Compressor ------------------------------------------------------------------------------------
static const unsigned char kTag[13] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3};

void Compressor::writeTag() {
  if (state_ != INIT) error("writeTag: not at the start of a block");
  out_->write((const char*)kTag, 13);
}

void Compressor::begin_block() {
  // "zPQ", level (2 if the model has no components), type 1, header (Z:16040-16046)
  if (header_.size() < 8) error("bad block header");
  out_->put('z'); out_->put('P'); out_->put('Q');
  out_->put(1 + (header_[6] == 0));
  out_->put(1);
  out_->write((const char*)header_.data(), (int)header_.size());
  have_pcomp_ = false; pcomp_.clear();
  state_ = BLOCK1;
}

void Compressor::startBlock(int level) {
  if (level < 1) error("compression level must be at least 1");
  const char* cfg = zq_model_config(level);   // min.cfg / mid.cfg as ZPAQL source
  if (!cfg) error("compression level too high");   // max.cfg: pass its source to startBlock(config, args)
  int args[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  startBlock(cfg, args, 0);
}

void Compressor::startBlock(const char* hcomp) {
  if (state_ != INIT) error("startBlock: block already open");
  const int hsize = toU16(hcomp);
  header_.assign((const unsigned char*)hcomp, (const unsigned char*)hcomp + hsize + 2);
  pcomp_default_.clear();
  begin_block();
}

void Compressor::startBlock(const char* config, int* args, Writer* pcomp_cmd) {
  if (state_ != INIT) error("startBlock: block already open");
  std::vector<unsigned char> h(1 << 16), pc(1 << 16);
  std::vector<char> cmd(1 << 12);
  uint32_t hl = (uint32_t)h.size(), pl = (uint32_t)pc.size();
  char err[256] = "";
  if (zq_assemble_config(config, args, h.data(), &hl, pc.data(), &pl, cmd.data(), cmd.size(), err, sizeof err) != ZQ_OK) error(err);
  header_.assign(h.begin(), h.begin() + hl);
  pcomp_default_.assign(pc.begin(), pc.begin() + pl);
  if (pcomp_cmd) for (const char* q = cmd.data(); *q; ++q) pcomp_cmd->put(*q);
  begin_block();
}

void Compressor::startSegment(const char* filename, const char* comment) {
  if (state_ == BLOCK2) error("more than one segment per block has no device path");
  if (state_ != BLOCK1) error("startSegment: no block open");
  filename_ = filename ? filename : ""; comment_ = comment ? comment : "";
  out_->put(1);
  out_->write(filename_.data(), (int)filename_.size()); out_->put(0);
  out_->write(comment_.data(), (int)comment_.size()); out_->put(0);
  out_->put(0);
  data_.clear();
  state_ = SEG1;
}

void Compressor::postProcess(const char* pcomp, int len) {
  if (state_ == SEG2) return;
  if (state_ != SEG1) error("postProcess: no segment open");
  if (!pcomp) pcomp_ = pcomp_default_;
  else {
    if (len == 0) { len = toU16(pcomp); pcomp += 2; }
    pcomp_.assign((const unsigned char*)pcomp, (const unsigned char*)pcomp + len);
  }
  have_pcomp_ = true;
  state_ = SEG2;
}

bool Compressor::compress(int n) {
  if (state_ == SEG1) postProcess();
  if (state_ != SEG2) error("compress: no segment open");
  char buf[1 << 14];
  while (n) {
    int nbuf = (int)sizeof buf;
    if (n >= 0 && n < nbuf) nbuf = n;
    const int nr = in_->read(buf, nbuf);
    if (nr < 0 || nr > nbuf) error("invalid read size");
    if (nr <= 0) return false;
    if (n >= 0) n -= nr;
    data_.insert(data_.end(), (unsigned char*)buf, (unsigned char*)buf + nr);
  }
  return true;
}

void Compressor::endSegment(const char* sha1string) {
  if (state_ == SEG1) postProcess();
  if (state_ != SEG2) error("endSegment: no segment open");
  if (data_.size() > 0xfffffff0ull) error("segment too large");
  zq_ctx* c = t_ctx.get();
  const uint64_t off = 0; uint64_t ooff = 0;
  const uint32_t len = (uint32_t)data_.size(); uint32_t olen = 0;
  std::vector<uint8_t> outbuf(zq_compress_bound(len) + pcomp_.size() * 2 + header_.size());
  const char* fn = filename_.c_str(); const char* cm = comment_.c_str();
  static const unsigned char none = 0;
  int rc = zq_compress_segments(c, 1, data_.empty() ? &none : data_.data(), &off, &len, header_.data(), (uint32_t)header_.size(),
                                pcomp_.empty() ? nullptr : pcomp_.data(), (uint32_t)pcomp_.size(), &fn, &cm, 1,
                                (const uint8_t*)sha1string, 0, outbuf.data(), outbuf.size(), &ooff, &olen);
  if (rc != ZQ_OK) error(zq_last_error(c));
  // the device wrote the whole block; everything up to the segment header is already on the Writer
  const size_t prefix = 5 + header_.size() + 1 + filename_.size() + 1 + comment_.size() + 2;
  out_->write((const char*)outbuf.data() + prefix, (int)(olen - prefix - 1));
  data_.clear();
  state_ = BLOCK2;
}

void Compressor::endBlock() {
  if (state_ != BLOCK2) error("endBlock: no finished segment");
  out_->put(255);
  state_ = INIT;
}

// ---- Decompresser ----------------------------------------------------------------------------------
bool Decompresser::fill(size_t need) {   // make buf_[.. need) available; false at EOF
  while (buf_.size() < need && !eof_) {
    char tmp[1 << 16];
    const int r = in_ ? in_->read(tmp, (int)sizeof tmp) : 0;
    if (r <= 0) { eof_ = true; break; }
    buf_.insert(buf_.end(), (unsigned char*)tmp, (unsigned char*)tmp + r);
  }
  return buf_.size() >= need;
}
int Decompresser::byte() { return fill(pos_ + 1) ? buf_[pos_++] : -1; }

static size_t find_locator(const std::vector<unsigned char>& b, size_t from) {
  for (size_t i = from; i + 16 <= b.size(); ++i)
    if (b[i] == kTag[0] && memcmp(&b[i], kTag, 13) == 0 && b[i + 13] == 'z' && b[i + 14] == 'P' && b[i + 15] == 'Q') return i;
  return (size_t)-1;
}

bool Decompresser::findBlock(double* memptr) {
  if (state_ != BLOCK) error("findBlock: inside a block");
  // drop what is behind us, then look for the 16-byte locator (tag + "zPQ"; findBlock's rolling hashes, Z:15421-15433)
  if (pos_) { buf_.erase(buf_.begin(), buf_.begin() + pos_); pos_ = 0; }
  size_t from = 0, at;
  while ((at = find_locator(buf_, from)) == (size_t)-1) {
    from = buf_.size() > 15 ? buf_.size() - 15 : 0;
    if (!fill(buf_.size() + 1)) return false;
  }
  blk_ = at; pos_ = at + 16;
  const int level = byte();
  if (level != 1 && level != 2) error("unsupported ZPAQ level");
  if (byte() != 1) error("unsupported ZPAQL type");
  if (!fill(pos_ + 2)) error("unexpected end of file");
  const size_t hsize = buf_[pos_] + 256u * buf_[pos_ + 1];
  if (!fill(pos_ + 2 + hsize)) error("unexpected end of file");
  hdr_ = pos_; hdr_len_ = hsize + 2;
  const unsigned char* h = &buf_[hdr_];
  if (hsize < 6) error("bad block header");
  if (level == 1 && h[6] == 0) error("ZPAQ level 1 requires at least 1 component");
  if (memptr) {   // ZPAQL::memory (Z:14192): table sizes implied by the header
    auto p2 = [](int x) { double r = 1; while (x-- > 0) r += r; return r; };
    static const int compsize[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};
    double mem = p2(h[2] + 2) + p2(h[3]) + p2(h[4] + 2) + p2(h[5]) + (double)(hsize + 300);
    size_t cp = 7;
    for (unsigned i = 0; i < h[6] && cp + 1 < hdr_len_; ++i) {
      const int t = h[cp];
      if (t < 1 || t > 9) error("unknown component type");
      const double size = p2(h[cp + 1]);
      if (t == 2) mem += 4 * size;
      else if (t == 3) mem += 64 * size + 1024;
      else if (t == 4) mem += 4 * size + p2(h[cp + 2]);
      else if (t == 6) mem += 2 * size;
      else if (t == 7) mem += 4 * size * h[cp + 3];
      else if (t == 8) mem += 64 * size + 2048;
      else if (t == 9) mem += 128 * size;
      cp += compsize[t];
    }
    *memptr = mem;
  }
  pos_ = hdr_ + hdr_len_;
  state_ = FILENAME; first_seg_ = true;
  return true;
}

void Decompresser::hcomp(Writer* out2) { if (out2 && hdr_len_) out2->write((const char*)&buf_[hdr_], (int)hdr_len_); }

bool Decompresser::findFilename(Writer* filename) {
  if (state_ != FILENAME) error("findFilename: not at a segment boundary");
  int c = byte();
  if (c == 1) {
    while (true) {
      c = byte();
      if (c == -1) error("unexpected EOF");
      if (c == 0) { state_ = COMMENT; return true; }
      if (filename) filename->put(c);
    }
  } else if (c == 255) { state_ = BLOCK; return false; }
  error("missing segment or end of block");
}

void Decompresser::readComment(Writer* comment) {
  if (state_ != COMMENT) error("readComment: not at a comment");
  state_ = DATA; decoded_ = false; expect_ = 0; have_expect_ = false;
  bool digits = true;
  while (true) {
    const int c = byte();
    if (c == -1) error("unexpected EOF");
    if (c == 0) break;
    if (digits && c >= '0' && c <= '9') { expect_ = expect_ * 10 + (c - '0'); have_expect_ = true; } else digits = false;
    if (comment) comment->put(c);
  }
  if (byte() != 0) error("missing reserved byte");
}

void Decompresser::decode_segment() {
  // The device restores a block whole -- the model, the coder's range and the post-processor carry on from one
  // segment into the next -- so the block's first segment decodes all of them and the later ones are served from
  // that result (zq_decompress_last_segments says where each one ends).
  if (first_seg_) {
    // the block runs to the next locator or to EOF; decoding stops at the last segment's end by itself
    size_t from = pos_, nxt;
    while ((nxt = find_locator(buf_, from)) == (size_t)-1) {
      from = buf_.size() > 15 ? buf_.size() - 15 : 0;
      if (from < pos_) from = pos_;
      if (!fill(buf_.size() + 1)) break;
    }
    const size_t end = nxt == (size_t)-1 ? buf_.size() : nxt;
    zq_ctx* c = t_ctx.get();
    const uint64_t off = blk_;
    const uint32_t len = (uint32_t)(end - blk_);
    uint64_t ooff = 0; uint32_t olen = 0, used = 0;
    unsigned char last_trailer[21];
    // expected size: the number that opens the comment, else (or with more segments) grow until the decoder stops complaining
    uint64_t cap = have_expect_ && expect_ <= 0xfffffff0ull ? expect_ : (uint64_t)len * 8 + 65536;
    for (;;) {
      const uint32_t e32 = (uint32_t)cap;
      block_out_.resize(cap + 16);
      const int rc = zq_decompress_blocks_ex(c, 1, buf_.data(), &off, &len, &e32, block_out_.data(), block_out_.size(), &ooff, &olen, &used, last_trailer);
      if (rc == ZQ_OK) break;
      if (rc == ZQ_E_OUTPUT && cap < 0xf0000000ull) { cap = cap * 4 < 0xfffffff0ull ? std::max<uint64_t>(cap * 4, 65536) : 0xfffffff0ull; continue; }
      error(zq_last_error(c));
    }
    block_out_.resize(olen);
    const zq_segment* zs = nullptr; uint64_t nz = 0;
    if (zq_decompress_last_segments(c, &zs, &nz) != ZQ_OK || nz == 0) error(zq_last_error(c));
    segs_.clear();
    for (uint64_t k = 0; k < nz; ++k) { Seg g; g.out_begin = zs[k].out_begin; g.out_end = zs[k].out_end; g.trailer = zs[k].trailer; segs_.push_back(g); }
    seg_idx_ = 0;
  }
  if (seg_idx_ >= segs_.size()) error("segment not found in the decoded block");
  const Seg& g = segs_[seg_idx_];
  obuf_.assign(block_out_.begin() + g.out_begin, block_out_.begin() + g.out_end);
  opos_ = 0;
  const size_t t = blk_ + g.trailer;
  if (t >= buf_.size()) error("unexpected end of file");
  memset(trailer_, 0, sizeof trailer_);
  if (buf_[t] == 253) { if (t + 21 > buf_.size()) error("unexpected end of file"); trailer_[0] = 1; memcpy(trailer_ + 1, &buf_[t + 1], 20); }
  seg_end_ = t + (buf_[t] == 253 ? 21 : 1);
  decoded_ = true;
}

bool Decompresser::decompress(int n) {
  if (state_ != DATA) error("decompress: not inside a segment");
  if (!decoded_) decode_segment();
  size_t k = obuf_.size() - opos_;
  const bool all = n < 0 || (size_t)n > k;   // the end-of-segment symbol would be decoded in this call
  if (!all) k = (size_t)n;
  if (k) {
    if (out_) out_->write((const char*)obuf_.data() + opos_, (int)k);
    if (sha1_) sha1_->write((const char*)obuf_.data() + opos_, (int64_t)k);
    opos_ += k;
  }
  if (all) { state_ = SEGEND; return false; }
  return true;
}

void Decompresser::readSegmentEnd(char* sha1string) {
  if (state_ != DATA && state_ != SEGEND) error("readSegmentEnd: not inside a segment");
  if (!decoded_) decode_segment();      // skipping still has to find the segment's end
  if (sha1string) memcpy(sha1string, trailer_, trailer_[0] ? 21 : 1);
  pos_ = seg_end_;
  obuf_.clear(); opos_ = 0; decoded_ = false; first_seg_ = false;
  ++seg_idx_;
  state_ = FILENAME;
}

void compressBlocks(int n, StringBuffer* const* ins, Writer* const* outs, const char* const* methods,
                    const char* const* filenames, const char* const* comments, bool dosha1) {
  if (n <= 0) return;
  zq_ctx* c = t_ctx.get();
  std::vector<uint64_t> off(n), ooff(n);
  std::vector<uint32_t> len(n), olen(n);
  uint64_t total = 0, bound = 0;
  for (int i = 0; i < n; ++i) {
    off[i] = total; len[i] = (uint32_t)ins[i]->size();
    total += (len[i] + 15) & ~15ull;
    bound += zq_compress_bound(len[i]);
  }
  std::vector<uint8_t> arena(total + 16), outbuf(bound);
  for (int i = 0; i < n; ++i) if (len[i]) memcpy(arena.data() + off[i], ins[i]->data(), len[i]);
  int rc = zq_compress_blocks(c, n, arena.data(), off.data(), len.data(), methods, filenames, comments, 0, dosha1 ? 1 : 0,
                              outbuf.data(), outbuf.size(), ooff.data(), olen.data());
  if (rc != ZQ_OK) error(zq_last_error(c));
  for (int i = 0; i < n; ++i) outs[i]->write((const char*)outbuf.data() + ooff[i], (int)olen[i]);
}

void compressBlock(StringBuffer* in, Writer* out, const char* method, const char* filename, const char* comment, bool dosha1) {
  compressBlocks(1, &in, &out, &method, filename ? &filename : nullptr, comment ? &comment : nullptr, dosha1);
}

void decompress(Reader* in, Writer* out) {
  // slurp the stream, split it at block starts (13-byte locator tag, Z:15973 / findBlock Z:15418)
  std::vector<uint8_t> buf;
  { char tmp[1 << 16]; int r; while ((r = in->read(tmp, sizeof tmp)) > 0) buf.insert(buf.end(), tmp, tmp + r); }
  static const unsigned char tag[13] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3};
  std::vector<uint64_t> off;
  for (size_t i = 0; i + 16 <= buf.size(); ++i)
    if (memcmp(&buf[i], tag, 13) == 0 && buf[i + 13] == 'z' && buf[i + 14] == 'P' && buf[i + 15] == 'Q') off.push_back(i);
  if (off.empty()) return;
  const int n = (int)off.size();
  std::vector<uint32_t> len(n), olen(n), expect(n);
  std::vector<uint64_t> ooff(n);
  uint64_t cap = 0;
  for (int i = 0; i < n; ++i) {
    len[i] = (uint32_t)((i + 1 < n ? off[i + 1] : buf.size()) - off[i]);
    // expected size = decimal number opening the segment comment (after the filename)
    // (untrusted input: every index is checked against the buffer; the size field is capped at what one block holds)
    size_t p = off[i] + 18;
    if (p + 2 > buf.size()) error("unexpected end of file");
    p += 2 + buf[p] + 256u * buf[p + 1];      // header
    ++p;                                      // segment marker
    if (p >= buf.size()) error("unexpected end of file");
    while (p < buf.size() && buf[p]) ++p;     // filename
    ++p;
    uint64_t e = 0;
    while (p < buf.size() && buf[p] >= '0' && buf[p] <= '9' && e <= 0xfffffff0ull) e = e * 10 + (buf[p++] - '0');
    if (e > 0xfffffff0ull) error("archive corrupted");
    expect[i] = (uint32_t)e;
    cap += e;
  }
  std::vector<uint8_t> outbuf(cap + 16);
  zq_ctx* c = t_ctx.get();
  int rc;
  for (int attempt = 0;; ++attempt) {
    // the comment's number is the FIRST segment's size: blocks of several segments (or comments without a number) need
    // more room -- grow and decode again
    rc = zq_decompress_blocks(c, n, buf.data(), off.data(), len.data(), expect.data(), outbuf.data(), outbuf.size(), ooff.data(), olen.data());
    if (rc != ZQ_E_OUTPUT || attempt >= 8) break;
    cap = 0;
    for (int i = 0; i < n; ++i) {
      const uint64_t g = std::max<uint64_t>((uint64_t)expect[i] * 4, (uint64_t)len[i] * 8 + 65536);
      expect[i] = (uint32_t)std::min<uint64_t>(g, 0xfffffff0ull);
      cap += expect[i];
    }
    outbuf.resize(cap + 16);
  }
  if (rc != ZQ_OK) error(zq_last_error(c));
  for (int i = 0; i < n; ++i) out->write((const char*)outbuf.data() + ooff[i], (int)olen[i]);
}

}  // namespace libzpaq_b200
