// libzpaq_b200.h -- host-side C++ mirror of the slice of the reference's `libzpaq` namespace that
// sits on the block-compression hot path (zpaqfranz.cpp Z:12554-13477), implemented on top of the
// C ABI in zq_b200.h.  Same names, argument meaning and error behaviour as the reference so that
// callers (and tests) read the same:
//
//   libzpaq::Reader / Writer          Z:12565 / Z:12571   byte-stream interfaces
//   libzpaq::error(const char*)       Z:12560             must not return; this build throws
//   libzpaq::StringBuffer             Z:13362-13466       growable in-memory Reader+Writer
//   libzpaq::compressBlock(...)       Z:13476 / Z:20255   one input buffer -> one ZPAQ block
//
// plus the batch form the GPU wants (compressBlocks), which is what a modified
// CompressJob::appendz (Z:71364) would call, and decompress() for whole streams of such blocks
// (the streaming Decompresser class with its per-segment callbacks stays the reference's own).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace libzpaq_b200 {

// Callback for error handling: throws std::runtime_error (std::bad_alloc for "Out of memory"),
// exactly like zpaqfranz's own handler (Z:27148-27154).
[[noreturn]] void error(const char* msg);

class Reader {
 public:
  virtual int get() = 0;                   // next byte or -1 at EOF
  virtual int read(char* buf, int n);      // up to n bytes, returns count
  virtual ~Reader() {}
};

class Writer {
 public:
  virtual void put(int c) = 0;
  virtual void write(const char* buf, int n);
  virtual ~Writer() {}
};

class StringBuffer : public Reader, public Writer {
  std::vector<unsigned char> buf_;
  size_t rpos_ = 0, limit_ = (size_t)-1;
 public:
  explicit StringBuffer(size_t reserve = 0) { buf_.reserve(reserve); }
  const char* c_str() const { return (const char*)buf_.data(); }
  unsigned char* data() { return buf_.data(); }
  size_t size() const { return buf_.size(); }
  size_t remaining() const { return buf_.size() - rpos_; }
  void reset() { buf_.clear(); rpos_ = 0; }
  void resize(size_t n) { buf_.resize(n); if (rpos_ > n) rpos_ = n; }
  void setLimit(size_t n) { limit_ = n; }
  void swap(StringBuffer& o) { buf_.swap(o.buf_); std::swap(rpos_, o.rpos_); std::swap(limit_, o.limit_); }
  void put(int c) override { if (buf_.size() >= limit_) error("StringBuffer overflow"); buf_.push_back((unsigned char)c); }
  void write(const char* b, int n) override;
  int get() override { return rpos_ < buf_.size() ? buf_[rpos_++] : -1; }
  int read(char* b, int n) override;
};

// == libzpaq::compressBlock (Z:20255). Like the reference it may transform `in` in place (E8E9).
// Runs as a batch of one on the calling thread's device context (created on first use, device
// taken from ZQ_DEVICE or 0).  Throws via error() where the reference would.
void compressBlock(StringBuffer* in, Writer* out, const char* method, const char* filename = 0,
                   const char* comment = 0, bool dosha1 = true);

// Batch form: n inputs -> n blocks appended to outs[i] (same per-element semantics).
void compressBlocks(int n, StringBuffer* const* ins, Writer* const* outs, const char* const* methods,
                    const char* const* filenames, const char* const* comments, bool dosha1 = true);

// == libzpaq::decompress(Reader*, Writer*) (Z:15536) for a stream of blocks as compressBlock writes
// them (one segment per block): every block found in `in` is restored on the device and appended to out.
void decompress(Reader* in, Writer* out);

}  // namespace libzpaq_b200
