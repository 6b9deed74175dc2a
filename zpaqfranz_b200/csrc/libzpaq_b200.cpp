// libzpaq_b200.cpp -- the C++ mirror of libzpaq's compress-side surface over the C ABI (see
// include/libzpaq_b200.h for the reference lines). No compute here: it packs the inputs into one
// arena, calls zq_compress_blocks and hands the blocks to the Writers.
#include "../../include/libzpaq_b200.h"

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
  std::vector<uint32_t> len(n), olen(n);
  std::vector<uint64_t> ooff(n);
  uint64_t cap = 0;
  for (int i = 0; i < n; ++i) {
    len[i] = (uint32_t)((i + 1 < n ? off[i + 1] : buf.size()) - off[i]);
    // expected size = decimal number opening the segment comment (after the filename)
    size_t p = off[i] + 18;
    p += 2 + buf[p] + 256u * buf[p + 1];      // header
    ++p;                                      // segment marker
    while (p < buf.size() && buf[p]) ++p;     // filename
    ++p;
    uint64_t e = 0;
    while (p < buf.size() && buf[p] >= '0' && buf[p] <= '9') e = e * 10 + (buf[p++] - '0');
    cap += e;
  }
  std::vector<uint8_t> outbuf(cap + 16);
  zq_ctx* c = t_ctx.get();
  int rc = zq_decompress_blocks(c, n, buf.data(), off.data(), len.data(), nullptr, outbuf.data(), outbuf.size(), ooff.data(), olen.data());
  if (rc != ZQ_OK) error(zq_last_error(c));
  for (int i = 0; i < n; ++i) out->write((const char*)outbuf.data() + ooff[i], (int)olen[i]);
}

}  // namespace libzpaq_b200
