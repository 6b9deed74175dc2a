// oracle/frontend_gpu_tail.cpp -- TEST INFRASTRUCTURE ONLY: the tail of the front-end link build.
//
// oracle/Makefile pipes the UNMODIFIED reference translation unit through sed, which renames the one line that opens
// the DEFINITION of libzpaq::compressBlock (Z:20255) to compressBlock_cpu, and appends this file: the archiver's own
// call sites (compressThread Z:71422, the metadata writers Z:71538 / 122910 / 122937 ...) then bind to the definition
// below, which forwards every block to libzqb200.so exactly like the stub in INTEGRATION.md section 1.  The result,
// oracle/_ref/libzpaqref_gpu.so, is the real zpaqfranz front end running on the device compressor; tests compare the
// archives it writes with the stock build's (tests/test_gpu_frontend.py).  No reference source is stored anywhere.
#include "zq_b200.h"

#include <atomic>
#include <vector>

static std::atomic<unsigned long long> g_gpu_blocks{0}, g_gpu_bytes{0};

namespace libzpaq {
void compressBlock(StringBuffer* in, Writer* out, const char* method, const char* filename, const char* comment, bool dosha1) {
  static thread_local zq_ctx* ctx = zq_create(0);            // one context per compressor thread
  if (!ctx) error(zq_last_error(0));
  uint64_t in_off = 0, out_off = 0;
  uint32_t in_len = (uint32_t)in->size(), out_len = 0;
  std::vector<uint8_t> buf(zq_compress_bound(in_len));
  const int rc = zq_compress_blocks(ctx, 1, (const uint8_t*)in->data(), &in_off, &in_len, &method, filename ? &filename : 0,
                                    comment ? &comment : 0, 1, dosha1, buf.data(), buf.size(), &out_off, &out_len);
  if (rc) error(zq_last_error(ctx));                          // throws like every libzpaq failure (Z:27148)
  out->write((const char*)buf.data(), (int)out_len);
  ++g_gpu_blocks; g_gpu_bytes += in_len;
}
}  // namespace libzpaq

extern "C" int zref_main(int argc, const char** argv) { return zpaqfranz_reference_main(argc, argv); }
extern "C" unsigned long long zref_gpu_blocks() { return g_gpu_blocks.load(); }
