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
//   libzpaq::SHA1 / SHA256            Z:12637 / Z:12828   streaming digests (hashed on the device at result())
//   libzpaq::Compressor               Z:13325-13358       caller-driven block/segment writer
//   libzpaq::Decompresser             Z:13241-13262       block/segment reader
//   libzpaq::decompress(Reader*,Writer*) Z:13264 / Z:15536
//
// plus the batch form the GPU wants (compressBlocks), which is what a modified
// CompressJob::appendz (Z:71364) would call.  The byte-at-a-time classes buffer one segment and run it
// on the device when the segment ends; a block holds ONE segment (what every zpaqfranz call site writes).
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

// == libzpaq::Array<T> (Z:12585): zero-initialised array whose element [0] sits on a 64-byte boundary; resize()
// erases the content; no copy, no assignment.  The front end uses it for its own tables; nothing on the device path
// does, it is here so that code written against the reference's header compiles against this one.
template <typename T>
class Array {
  std::vector<unsigned char> raw_;
  T* data_ = nullptr;
  size_t n_ = 0;
  Array(const Array&) = delete;
  void operator=(const Array&) = delete;

 public:
  explicit Array(size_t sz = 0, int ex = 0) { resize(sz, ex); }
  void resize(size_t sz, int ex = 0) {      // sz << ex elements, all zero
    for (; ex > 0; --ex) { if (sz > sz * 2) error("Array too big"); sz *= 2; }
    raw_.clear(); raw_.shrink_to_fit(); data_ = nullptr; n_ = 0;
    if (sz == 0) return;
    const size_t nb = 128 + sz * sizeof(T);
    if (nb <= 128 || (nb - 128) / sizeof(T) != sz) error("Array too big");
    try { raw_.assign(nb, 0); } catch (...) { error("Out of memory"); }
    const size_t mis = reinterpret_cast<uintptr_t>(raw_.data()) & 63;
    data_ = reinterpret_cast<T*>(raw_.data() + (64 - mis));
    n_ = sz;
  }
  size_t size() const { return n_; }
  int isize() const { return int(n_); }
  T& operator[](size_t i) { if (!(n_ > 0 && i < n_)) error("09386: operator[] kaputt"); return data_[i]; }
  T& operator()(size_t i) { return data_[i & (n_ - 1)]; }     // n a power of two
};

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

inline int toU16(const char* p) { return (p[0] & 255) + 256 * (p[1] & 255); }   // Z:13479

// == libzpaq::SHA1 (Z:12637): put/write accumulate, result() returns the 20-byte digest and resets.
// The bytes are kept until result() and hashed there by the device kernel (zq_sha1).
class SHA1 {
  std::vector<unsigned char> buf_;   // bytes not hashed yet; flushed to the device in whole 64-byte blocks at kFlush
  uint64_t len_ = 0;
  uint32_t st_[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  bool streamed_ = false;            // some blocks already went through zq_sha1_continue
  char h_[20];
  void flush();
 public:
  static const size_t kFlush = 4u << 20;
  void put(int c) { buf_.push_back((unsigned char)c); ++len_; if (buf_.size() >= kFlush) flush(); }
  void write(const char* b, int64_t n);
  double size() const { return (double)len_; }
  uint64_t usize() const { return len_; }
  const char* result();
};

// == libzpaq::SHA256 (Z:12828)
class SHA256 {
  std::vector<unsigned char> buf_;
  uint64_t len_ = 0;
  uint32_t st_[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
  bool streamed_ = false;
  char h_[32];
  void flush();
 public:
  static const size_t kFlush = 4u << 20;
  void put(int c) { buf_.push_back((unsigned char)c); ++len_; if (buf_.size() >= kFlush) flush(); }
  void write(const char* b, int64_t n);
  double size() const { return (double)len_; }
  uint64_t usize() const { return len_; }
  const char* result();
};

// == libzpaq::Compressor (Z:13325-13358, Z:15970-16187).  Same call sequence and the same bytes on the
// Writer at the same calls: writeTag / startBlock / startSegment write at once; the segment's data is
// collected by compress() and coded on the device in endSegment().
class Compressor {
 public:
  Compressor() {}
  void setOutput(Writer* out) { out_ = out; }
  void writeTag();
  void startBlock(int level);                     // built-in models 1 (min.cfg) and 2 (mid.cfg)
  void startBlock(const char* hcomp);             // stored header: hsize[2] hh hm ph pm n comp.. 0 hcomp.. 0
  void startBlock(const char* config, int* args, Writer* pcomp_cmd = 0);   // ZPAQL source (Compiler, Z:15904)
  void startSegment(const char* filename = 0, const char* comment = 0);
  void setInput(Reader* i) { in_ = i; }
  void postProcess(const char* pcomp = 0, int len = 0);
  bool compress(int n = -1);                      // n bytes, or to EOF if n < 0; false at EOF
  void endSegment(const char* sha1string = 0);
  void endBlock();
  int stat(int) { return 0; }
  void hcomp(Writer* out2) { if (out2) out2->write((const char*)header_.data(), (int)header_.size()); }
  void setVerify(bool) {}                         // the device decoder is the verifier (zq_decompress_blocks)
 private:
  enum { INIT, BLOCK1, SEG1, BLOCK2, SEG2 } state_ = INIT;
  Writer* out_ = 0;
  Reader* in_ = 0;
  std::vector<unsigned char> header_, pcomp_default_, pcomp_, data_;
  std::string filename_, comment_;
  bool have_pcomp_ = false;
  void begin_block();
};

// == libzpaq::Decompresser (Z:13241-13262, Z:15418-15534).  Reads ahead to the next block locator tag
// (or EOF) and restores the segment on the device at the first decompress()/readSegmentEnd().
class Decompresser {
 public:
  Decompresser() {}
  void setInput(Reader* in) { in_ = in; }
  bool findBlock(double* memptr = 0);
  void hcomp(Writer* out2);
  bool findFilename(Writer* = 0);
  void readComment(Writer* = 0);
  void setOutput(Writer* out) { out_ = out; }
  void setSHA1(SHA1* sha1ptr) { sha1_ = sha1ptr; }
  bool decompress(int n = -1);                    // n bytes, -1 = all; true until the segment is done
  void readSegmentEnd(char* sha1string = 0);
  int stat(int) { return 0; }
  int buffered() { return (int)(buf_.size() - pos_); }
 private:
  enum { BLOCK, FILENAME, COMMENT, DATA, SEGEND } state_ = BLOCK;
  Reader* in_ = 0;
  Writer* out_ = 0;
  SHA1* sha1_ = 0;
  std::vector<unsigned char> buf_, obuf_, block_out_;   // block_out_: every segment of the current block, restored at once
  struct Seg { uint32_t out_begin, out_end, trailer; };
  std::vector<Seg> segs_;
  size_t seg_idx_ = 0;
  size_t pos_ = 0, blk_ = 0, hdr_ = 0, hdr_len_ = 0, opos_ = 0, seg_end_ = 0;
  bool eof_ = false, decoded_ = false, first_seg_ = true;
  unsigned char trailer_[21];
  uint64_t expect_ = 0; bool have_expect_ = false;
  bool fill(size_t need);
  int byte();
  void decode_segment();
};

// == libzpaq::compressBlock (Z:20255). Like the reference it may transform `in` in place (E8E9).
// Runs as a batch of one on the calling thread's device context (created on first use, device
// taken from ZQ_DEVICE or 0).  Throws via error() where the reference would.
void compressBlock(StringBuffer* in, Writer* out, const char* method, const char* filename = 0,
                   const char* comment = 0, bool dosha1 = true);

// Batch form: n inputs -> n blocks appended to outs[i] (same per-element semantics).
void compressBlocks(int n, StringBuffer* const* ins, Writer* const* outs, const char* const* methods,
                    const char* const* filenames, const char* const* comments, bool dosha1 = true);

// == libzpaq::decompress(Reader*, Writer*) (Z:15536): every block found in `in` is restored on the device
// and appended to out (all blocks of the stream in one batch).
void decompress(Reader* in, Writer* out);

}  // namespace libzpaq_b200

// The archiver's sources say `libzpaq::`: with this alias they compile against the mirror unchanged.  Define
// ZQ_NO_LIBZPAQ_ALIAS when the reference's own namespace is in the same translation unit (the parity tests).
#ifndef ZQ_NO_LIBZPAQ_ALIAS
namespace libzpaq = libzpaq_b200;
#endif
