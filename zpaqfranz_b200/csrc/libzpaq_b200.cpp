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

}  // namespace libzpaq_b200
