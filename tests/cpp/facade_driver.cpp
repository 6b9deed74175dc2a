// Exercises the C++ mirror of the libzpaq classes (include/libzpaq_b200.h) the way zpaqfranz's call sites use
// the originals: compressBlock (Z:71422), Compressor driven by hand (Z:20396-20428), the Decompresser walk of
// decompressThread (Z:72697-72770) and the SHA1/SHA256 classes.  Usage: facade_driver <input> <outdir>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "libzpaq_b200.h"

using namespace libzpaq_b200;

static std::vector<char> slurp(const std::string& p) {
  std::vector<char> v; FILE* f = fopen(p.c_str(), "rb"); if (!f) { perror(p.c_str()); exit(2); }
  char buf[65536]; size_t r; while ((r = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + r);
  fclose(f); return v;
}
static void dump(const std::string& p, const char* d, size_t n) { FILE* f = fopen(p.c_str(), "wb"); fwrite(d, 1, n, f); fclose(f); }
static std::string hex(const char* d, int n) { std::string s; char b[3]; for (int i = 0; i < n; ++i) { snprintf(b, 3, "%02x", d[i] & 255); s += b; } return s; }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::vector<char> in = slurp(argv[1]);
  const std::string dir = argv[2];
  try {
    // (a) compressBlock
    StringBuffer sb; sb.write(in.data(), (int)in.size());
    StringBuffer a;
    compressBlock(&sb, &a, "2", "file_a", "jDC\x01", true);
    dump(dir + "/a.zpaq", a.c_str(), a.size());

    // digests of the input through the streaming classes
    SHA1 s1; s1.write(in.data(), (int64_t)in.size());
    char sha1[20]; memcpy(sha1, s1.result(), 20);
    SHA256 s2; for (char c : in) s2.put(c & 255);
    printf("sha1 %s\nsha256 %s\n", hex(sha1, 20).c_str(), hex(s2.result(), 32).c_str());

    // a stream longer than the classes keep on the host (kFlush): written in uneven pieces, hashed block by block
    {
      SHA1 big1; SHA256 big2;
      std::vector<char> piece;
      uint64_t total = 0, x = 12345;
      while (total < 9500000) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const size_t k = 1 + (size_t)((x >> 33) % 300000);
        piece.resize(k);
        for (size_t i = 0; i < k; ++i) piece[i] = (char)((total + i) * 2654435761u >> 13);
        big1.write(piece.data(), (int64_t)k); big2.write(piece.data(), (int64_t)k);
        total += k;
      }
      big1.put(7); big2.put(7); ++total;
      char d1[20]; memcpy(d1, big1.result(), 20);
      printf("bigsize %llu\nbigsha1 %s\nbigsha256 %s\n", (unsigned long long)total, hex(d1, 20).c_str(), hex(big2.result(), 32).c_str());
    }

    // (b) Compressor by hand, built-in model 2, checksum stored
    StringBuffer src; src.write(in.data(), (int)in.size());
    StringBuffer b;
    Compressor co;
    co.setOutput(&b); co.setInput(&src);
    co.writeTag(); co.startBlock(2); co.startSegment("file_b", "a comment");
    co.postProcess(); co.compress(1000); while (co.compress(4096)) {}
    co.endSegment(sha1); co.endBlock();
    dump(dir + "/b.zpaq", b.c_str(), b.size());

    // (c) walk both blocks like decompressThread does
    StringBuffer both; both.write(a.c_str(), (int)a.size()); both.write("garbage between blocks", 22); both.write(b.c_str(), (int)b.size());
    Decompresser d; d.setInput(&both);
    StringBuffer out;
    int nblock = 0;
    double mem = 0;
    while (d.findBlock(&mem)) {
      StringBuffer fn, cm;
      while (d.findFilename(&fn)) {
        d.readComment(&cm);
        SHA1 check; d.setOutput(&out); d.setSHA1(&check);
        if (nblock == 0) d.decompress(); else { while (d.decompress(10000)) {} }
        char tr[21]; d.readSegmentEnd(tr);
        const uint64_t sz = check.usize();
        const bool ok = tr[0] == 1 && memcmp(tr + 1, check.result(), 20) == 0;
        printf("block %d name %s comment_len %d size %llu stored_sha1 %d match %d mem %.0f\n", nblock, std::string(fn.c_str(), fn.size()).c_str(),
               (int)cm.size(), (unsigned long long)sz, tr[0], ok ? 1 : 0, mem);
      }
      ++nblock;
    }
    dump(dir + "/out.bin", out.c_str(), out.size());
    printf("blocks %d\n", nblock);

    // (e) optional: an archive given by the caller (blocks of several segments), walked segment by segment and through
    //     the free function decompress()
    if (argc > 3) {
      const std::vector<char> ar = slurp(argv[3]);
      StringBuffer src2; src2.write(ar.data(), (int)ar.size());
      Decompresser d2; d2.setInput(&src2);
      StringBuffer out2;
      int nb = 0;
      while (d2.findBlock()) {
        for (int k = 0;; ++k) {
          StringBuffer fn, cm;
          if (!d2.findFilename(&fn)) break;
          d2.readComment(&cm);
          SHA1 check; d2.setOutput(&out2); d2.setSHA1(&check);
          if (k & 1) d2.decompress(); else { while (d2.decompress(777)) {} }
          char tr[21]; d2.readSegmentEnd(tr);
          const uint64_t sz = check.usize();
          const bool ok = tr[0] == 1 && memcmp(tr + 1, check.result(), 20) == 0;
          printf("seg %d.%d name %s size %llu stored_sha1 %d match %d\n", nb, k, std::string(fn.c_str(), fn.size()).c_str(), (unsigned long long)sz, tr[0], ok ? 1 : 0);
        }
        ++nb;
      }
      dump(dir + "/out2.bin", out2.c_str(), out2.size());
      StringBuffer src3; src3.write(ar.data(), (int)ar.size());
      StringBuffer out3;
      decompress(&src3, &out3);
      dump(dir + "/out3.bin", out3.c_str(), out3.size());
    }

    // (d) errors surface as exceptions carrying the library's message
    try { Compressor bad; StringBuffer o; bad.setOutput(&o); int args[9] = {0}; bad.startBlock("comp 0 0 0 0 1 0 nosuch 1 hcomp halt end", args); printf("error none\n"); }
    catch (std::exception& e) { printf("error %s\n", e.what()); }
  } catch (std::exception& e) { printf("FAILED %s\n", e.what()); return 1; }
  return 0;
}
