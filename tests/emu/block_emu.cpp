// block_emu.cpp -- TEST INFRASTRUCTURE ONLY: one whole compressBlock on the host through tests/emu/simt_emu.h, every
// device stage in the order zq_compress_blocks launches them: k_sha1_units -> (k_e8e9) -> k_suffix_sort -> k_lz77_sa |
// k_lz77_hash | k_bwt_stream -> (k_cm_encode) -> k_frame.  The block prefix is laid out as zq_api.cu does
// (tag, "zPQ", level, 1, header, 1, filename, 0, "<n>[ comment]", 0, 0).  tests/test_block_emu.py compares the bytes
// with the reference's libzpaq::compressBlock.
#include <cuda_runtime.h>   // the shim

#include <string>

#include "zq_lz77.cuh"
#include "zq_sha1.cuh"
#include "zq_frame.cuh"
#include "zq_decode.cuh"
#include "zq_cm_host.h"

using namespace zqdev;

extern "C" long emu_block(const uint8_t* data, uint32_t n, const int* args, const uint8_t* header, uint32_t hlen,
                          const uint8_t* pcomp, uint32_t pclen, const char* filename, const char* comment, int dosha1,
                          uint8_t* out, uint32_t cap) {
  try {
    static const u8 kTag[13] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3};
    const int ncomp = header[6];
    const int lz_level = args[1] & 3;
    const bool e8 = (args[1] & 4) != 0, use_sa = lz_level == 3 || (lz_level && args[5] - args[0] >= 21);
    std::vector<u8> in(data, data + n); in.resize(n + 64);
    // blob: payload (selector + PCOMP), prefix, then HCOMP byte code for the coder
    std::vector<u8> blob;
    ZqPlan pl; memset(&pl, 0, sizeof pl);
    for (int k = 0; k < 9; ++k) pl.args[k] = args[k];
    pl.lz_level = lz_level; pl.use_sa = use_sa; pl.e8e9 = e8; pl.modeled = ncomp > 0; pl.cm_plan = 0;
    pl.payload_off = 0;
    if (pclen) { blob.push_back(1); blob.push_back(pclen & 255); blob.push_back(pclen >> 8); blob.insert(blob.end(), pcomp, pcomp + pclen); }
    else blob.push_back(0);
    pl.payload_len = (u32)blob.size();
    ZqUnit u; memset(&u, 0, sizeof u);
    u.n = n; u.idx16 = n <= 65536; u.lz_cap = n + n / 32 + 64 + 5; u.coded_cap = 2 * u.lz_cap + 2 * pl.payload_len + 1024;
    u.prefix_off = (u32)blob.size();
    blob.insert(blob.end(), kTag, kTag + 13);
    blob.push_back('z'); blob.push_back('P'); blob.push_back('Q'); blob.push_back(1 + (ncomp == 0)); blob.push_back(1);
    blob.insert(blob.end(), header, header + hlen);
    blob.push_back(1);
    if (filename) blob.insert(blob.end(), filename, filename + strlen(filename));
    blob.push_back(0);
    std::string cs = std::to_string(n);
    if (comment) { cs += " "; cs += comment; }
    blob.insert(blob.end(), cs.begin(), cs.end());
    blob.push_back(0); blob.push_back(0);
    u.prefix_len = (u32)blob.size() - u.prefix_off;
    int todo = 0;
    // 1. SHA-1 of the original bytes, then the E8E9 filter in place
    std::vector<u8> sha(20);
    if (dosha1) emu::launch(1, 128, 0, [&] { k_sha1_units(in.data(), &u, 1, sha.data()); });
    if (e8) emu::launch(1, 64, 0, [&] { k_e8e9(in.data(), &u, &todo, 1); });
    // 2. pre-pass
    std::vector<u8> lz(u.lz_cap + 64), work;
    u32 lzlen = 0, err = 0, next = 0;
    if (use_sa) {
      const u32 w = u.idx16 ? 2 : 4;
      work.resize(zq_work_bytes(n, w) + 256);
      const size_t scr = (((size_t)n + 1) + 63) & ~(size_t)63;
      std::vector<u64> kbuf(2 * scr); std::vector<u32> vbuf(6 * scr);
      emu::launch(1, 256, sizeof(SortSmem<256>), [&] { k_suffix_sort<256, 4>(in.data(), &u, &todo, 1, work.data(), kbuf.data(), vbuf.data(), scr); });
      if (lz_level == 3) emu::launch(1, 256, 0, [&] { k_bwt_stream(in.data(), &u, &todo, 1, work.data(), lz.data(), &lzlen); });
      else emu::launch(1, 32, 0, [&] {
        if (u.idx16) k_lz77_sa<u16>(in.data(), &u, &pl, &todo, 1, work.data(), lz.data(), &lzlen, &err, &next);
        else k_lz77_sa<u32>(in.data(), &u, &pl, &todo, 1, work.data(), lz.data(), &lzlen, &err, &next);
      });
    } else if (lz_level) {
      std::vector<u32> ht((size_t)1 << args[5], 0u);
      emu::launch(1, 32, 0, [&] { k_lz77_hash<8>(in.data(), &u, &pl, &todo, 1, (u8*)ht.data(), lz.data(), &lzlen, &err, &next); });
    }
    if (err) return -10 - (long)err;
    // 3. context-mixing coder
    std::vector<u8> coded(u.coded_cap + 64);
    u32 coded_len = 0;
    ZqCmPlan cp; memset(&cp, 0, sizeof cp);
    if (ncomp > 0) {
      size_t used = 0;
      zq::Assembled code = zq::parse_block_header(header, hlen, &used);
      std::vector<ZqCmFill> fills;
      cp = zq::make_cm_plan(code, fills);
      cp.hcomp_off = (u32)blob.size(); cp.hcomp_len = (u32)code.hcomp.size();
      blob.insert(blob.end(), code.hcomp.begin(), code.hcomp.end());
      blob.resize(blob.size() + 16);
      const zq::CmTables& tab = zq::cm_tables();
      std::vector<u8> model((size_t)cp.model_bytes + 512);
      u8* mp = (u8*)(((uintptr_t)model.data() + 255) & ~(uintptr_t)255);
      const CmTablesDev* dtab = (const CmTablesDev*)&tab;
      next = 0;
      emu::launch(cp.fill_count, 256, 0, [&] { k_cm_init(&u, &pl, &cp, fills.data(), &todo, 1, (int)cp.fill_count, dtab, mp); });
      emu::launch(1, 64, sizeof(CmSmem) + sizeof(CmUnitSmem), [&] {
        k_cm_encode<0, false>(in.data(), &u, &pl, &cp, &todo, 1, dtab, blob.data(), lz.data(), &lzlen, mp, coded.data(), &coded_len, &err, &next, 1, 1,
                              nullptr, nullptr);
      });
      if (err) return -20 - (long)err;
    }
    // 4. framing
    const u64 out_off = 0;
    std::vector<u8> o((size_t)cap + 64);
    emu::launch(1, 256, 0, [&] {
      k_frame(&u, &pl, &todo, 1, blob.data(), in.data(), lz.data(), &lzlen, coded.data(), &coded_len, dosha1 ? sha.data() : nullptr, &out_off, o.data());
    });
    const u64 size = ncomp > 0 ? (u64)u.prefix_len + coded_len + 4 + (dosha1 ? 21 : 1) + 1
                               : unmodeled_block_size(u.prefix_len, (u64)pl.payload_len + (lz_level ? lzlen : n), dosha1 != 0);
    if (size > cap) return -2;
    memcpy(out, o.data(), size);
    return (long)size;
  } catch (const zq::Error& e) {
    fprintf(stderr, "emu_block: %s\n", e.msg.c_str());
    return -100;
  }
}
