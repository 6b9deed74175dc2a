// zq_jit.cpp -- see zq_jit.h.
#include "zq_jit.h"

#include <dlfcn.h>

#include <cstdio>
#include <set>

namespace zq {
namespace {

int op_len(unsigned op) { return 1 + ((op & 7) == 7) + (op == 255); }

std::string num(unsigned v) { return std::to_string(v) + "u"; }

// the value of operand `src` of a two-operand instruction
std::string operand(unsigned src, unsigned imm) {
  switch (src) {
    case 0: return "a"; case 1: return "b"; case 2: return "c"; case 3: return "d";
    case 4: return "(unsigned)M[b & MM]"; case 5: return "(unsigned)M[c & MM]"; case 6: return "H[d & HM]";
    default: return num(imm);
  }
}

}  // namespace

bool jit_translate(const uint8_t* code, size_t len, bool with_out, std::string& body, std::string& why) {
  // pass 1: instruction starts and jump targets
  std::set<long> starts, targets;
  for (size_t pc = 0; pc < len; pc += op_len(code[pc])) starts.insert((long)pc);
  for (size_t pc = 0; pc < len; pc += op_len(code[pc])) {
    const unsigned op = code[pc];
    if (pc + op_len(op) > len + 1) break;
    long t = -1;
    if (op == 39 || op == 47 || op == 63) t = (long)pc + 2 + (long)(int8_t)code[pc + 1];   // ((N+128)&255)-127 from the operand byte
    else if (op == 255 && pc + 2 < len) t = (long)code[pc + 1] + 256L * code[pc + 2];
    else continue;
    if (t < 0 || t > (long)len) { targets.insert(-1); continue; }   // leaves the program: a ZPAQL error at run time
    if (t < (long)len && !starts.count(t)) { why = "jump into the middle of an instruction"; return false; }
    targets.insert(t);
  }
  auto label = [&](long t) { return t < 0 || t >= (long)len ? std::string("Lerr") : "L" + std::to_string(t); };
  // pass 2: statements
  std::string s;
  char buf[64];
  for (size_t pc = 0; pc < len; pc += op_len(code[pc])) {
    const unsigned op = code[pc], n = pc + 1 < len ? code[pc + 1] : 0;
    if (targets.count((long)pc)) { snprintf(buf, sizeof buf, "L%ld:;\n", (long)pc); s += buf; }
    s += "  ";
    if (op >= 64) {
      const unsigned src = op & 7, grp = op >> 3;
      const std::string x = operand(src, n);
      switch (grp) {
        case 8: s += "a = " + x + ";"; break; case 9: s += "b = " + x + ";"; break;
        case 10: s += "c = " + x + ";"; break; case 11: s += "d = " + x + ";"; break;
        case 12: s += "M[b & MM] = (unsigned char)(" + x + ");"; break;
        case 13: s += "M[c & MM] = (unsigned char)(" + x + ");"; break;
        case 14: s += "H[d & HM] = " + x + ";"; break;
        case 16: s += "a += " + x + ";"; break; case 17: s += "a -= " + x + ";"; break; case 18: s += "a *= " + x + ";"; break;
        case 19: s += "{ const unsigned x = " + x + "; a = x ? a / x : 0u; }"; break;
        case 20: s += "{ const unsigned x = " + x + "; a = x ? a % x : 0u; }"; break;
        case 21: s += "a &= " + x + ";"; break; case 22: s += "a &= ~(" + x + ");"; break;
        case 23: s += "a |= " + x + ";"; break; case 24: s += "a ^= " + x + ";"; break;
        case 25: s += "a <<= ((" + x + ") & 31u);"; break; case 26: s += "a >>= ((" + x + ") & 31u);"; break;
        case 27: s += "f = a == " + x + ";"; break; case 28: s += "f = a < " + x + ";"; break; case 29: s += "f = a > " + x + ";"; break;
        default:
          if (op == 255) s += "goto " + label(pc + 2 < len ? (long)code[pc + 1] + 256L * code[pc + 2] : -1) + ";";
          else s += "goto Lerr;";
      }
    } else {
      const long rel = (long)pc + 2 + (long)(int8_t)n;
      switch (op) {
        case 1: s += "++a;"; break; case 2: s += "--a;"; break; case 3: s += "a = ~a;"; break; case 4: s += "a = 0u;"; break;
        case 7: s += "a = R[" + num(n) + "];"; break;
        case 8: s += "{ const unsigned t = a; a = b; b = t; }"; break;
        case 9: s += "++b;"; break; case 10: s += "--b;"; break; case 11: s += "b = ~b;"; break; case 12: s += "b = 0u;"; break;
        case 15: s += "b = R[" + num(n) + "];"; break;
        case 16: s += "{ const unsigned t = a; a = c; c = t; }"; break;
        case 17: s += "++c;"; break; case 18: s += "--c;"; break; case 19: s += "c = ~c;"; break; case 20: s += "c = 0u;"; break;
        case 23: s += "c = R[" + num(n) + "];"; break;
        case 24: s += "{ const unsigned t = a; a = d; d = t; }"; break;
        case 25: s += "++d;"; break; case 26: s += "--d;"; break; case 27: s += "d = ~d;"; break; case 28: s += "d = 0u;"; break;
        case 31: s += "d = R[" + num(n) + "];"; break;
        case 32: s += "{ const unsigned t = M[b & MM]; M[b & MM] = (unsigned char)a; a = (a & ~255u) | t; }"; break;
        case 33: s += "M[b & MM] = (unsigned char)(M[b & MM] + 1);"; break; case 34: s += "M[b & MM] = (unsigned char)(M[b & MM] - 1);"; break;
        case 35: s += "M[b & MM] = (unsigned char)~M[b & MM];"; break; case 36: s += "M[b & MM] = 0;"; break;
        case 39: s += "if (f) goto " + label(rel) + ";"; break;
        case 40: s += "{ const unsigned t = M[c & MM]; M[c & MM] = (unsigned char)a; a = (a & ~255u) | t; }"; break;
        case 41: s += "M[c & MM] = (unsigned char)(M[c & MM] + 1);"; break; case 42: s += "M[c & MM] = (unsigned char)(M[c & MM] - 1);"; break;
        case 43: s += "M[c & MM] = (unsigned char)~M[c & MM];"; break; case 44: s += "M[c & MM] = 0;"; break;
        case 47: s += "if (!f) goto " + label(rel) + ";"; break;
        case 48: s += "{ const unsigned t = H[d & HM]; H[d & HM] = a; a = t; }"; break;
        case 49: s += "H[d & HM] = H[d & HM] + 1u;"; break; case 50: s += "H[d & HM] = H[d & HM] - 1u;"; break;
        case 51: s += "H[d & HM] = ~H[d & HM];"; break; case 52: s += "H[d & HM] = 0u;"; break;
        case 55: s += "R[" + num(n) + "] = a;"; break;
        case 56: s += "goto Lend;"; break;
        case 57: s += with_out ? "ZQ_JIT_OUT(a);" : ";"; break;
        case 59: s += "a = (a + M[b & MM] + 512u) * 773u;"; break;
        case 60: s += "H[d & HM] = (H[d & HM] + a + 512u) * 773u;"; break;
        case 63: s += "goto " + label(rel) + ";"; break;
        default: s += "goto Lerr;";
      }
    }
    snprintf(buf, sizeof buf, "   /* %zu: %u */\n", pc, op);
    s += buf;
  }
  s += "Lerr: err = 1;\nLend:;\n";   // running past the last instruction is a ZPAQL error too
  body = s;
  return true;
}

bool jit_context_source(const uint8_t* hcomp, size_t len, int hh, int hm, int ncomp, std::string& src, std::string& why) {
  std::string body;
  if (!jit_translate(hcomp, len, false, body, why)) return false;
  std::string s;
  s += "// generated by zq_jit (ZPAQL -> CUDA C); do not edit\n";
  s += "#ifndef ZQ_JIT_FN\n#define ZQ_JIT_FN static __device__ __forceinline__\n#endif\n";
  s += "#define MM " + num(hm >= 32 ? 0xffffffffu : ((1u << hm) - 1u)) + "\n";
  s += "#define HM " + num(hh >= 32 ? 0xffffffffu : ((1u << hh) - 1u)) + "\n";
  s += "#define ZQ_JIT_NCOMP " + std::to_string(ncomp) + "\n";
  s += "struct ZqJitVm { unsigned a, b, c, d, f; };\n";
  s += "ZQ_JIT_FN void zq_hcomp(ZqJitVm& v, unsigned char* M, unsigned* H, unsigned* R, unsigned input, int& err) {\n";
  s += "  unsigned a = input, b = v.b, c = v.c, d = v.d, f = v.f;\n";
  s += body;
  s += "  v.a = a; v.b = b; v.c = c; v.d = d; v.f = f;\n}\n";
  s += "#undef MM\n#undef HM\n";
  s +=
      "#ifdef __CUDACC__\n"
      "// one thread per block: the block's coded bytes are head[0..hlen) then stream; M, H, R live in the block's model\n"
      "// region (zeroed by k_cm_init); ctx receives H[0..ncomp) after every byte but the last, byte-major\n"
      "extern \"C\" __global__ void zq_ctx_kernel(const unsigned char* head, unsigned hlen, const unsigned char* sbase,\n"
      "                                         const unsigned long long* soff, const unsigned* slen, int nunits,\n"
      "                                         unsigned char* model_base, const unsigned long long* model_off,\n"
      "                                         unsigned long long m_off, unsigned long long h_off, unsigned long long r_off,\n"
      "                                         unsigned* ctx_base, const unsigned long long* ctx_off, unsigned* err_flag) {\n"
      "  const int t = blockIdx.x * blockDim.x + threadIdx.x;\n"
      "  if (t >= nunits) return;\n"
      "  unsigned char* model = model_base + model_off[t];\n"
      "  unsigned char* M = model + m_off;\n"
      "  unsigned* H = (unsigned*)(model + h_off);\n"
      "  unsigned* R = (unsigned*)(model + r_off);\n"
      "  const unsigned char* s = sbase + soff[t];\n"
      "  unsigned* out = ctx_base + ctx_off[t];\n"
      "  ZqJitVm v; v.a = v.b = v.c = v.d = v.f = 0;\n"
      "  int err = 0;\n"
      "  const unsigned K = hlen + slen[t];\n"
      "  for (unsigned k = 0; k + 1 < K; ++k) {\n"
      "    zq_hcomp(v, M, H, R, k < hlen ? head[k] : s[k - hlen], err);\n"
      "    for (int i = 0; i < ZQ_JIT_NCOMP; ++i) out[(unsigned long long)k * ZQ_JIT_NCOMP + i] = H[i & " + num(hh >= 32 ? 0xffffffffu : ((1u << hh) - 1u)) + "];\n"
      "  }\n"
      "  if (err) atomicOr(err_flag, 2u);\n"
      "}\n"
      "#endif\n";
  src = s;
  return true;
}

int jit_compile(const std::string& src, std::vector<char>& cubin, std::string& log) {
  typedef int (*create_t)(void**, const char*, const char*, int, const char* const*, const char* const*);
  typedef int (*compile_t)(void*, int, const char* const*);
  typedef int (*size_t_fn)(void*, size_t*);
  typedef int (*get_t)(void*, char*);
  typedef int (*destroy_t)(void**);
  static void* lib = nullptr;
  if (!lib) {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
  }
  if (!lib) { log = "libnvrtc not found"; return -1; }
  create_t create = (create_t)dlsym(lib, "nvrtcCreateProgram");
  compile_t compile = (compile_t)dlsym(lib, "nvrtcCompileProgram");
  size_t_fn logsize = (size_t_fn)dlsym(lib, "nvrtcGetProgramLogSize");
  get_t getlog = (get_t)dlsym(lib, "nvrtcGetProgramLog");
  size_t_fn binsize = (size_t_fn)dlsym(lib, "nvrtcGetCUBINSize");
  get_t getbin = (get_t)dlsym(lib, "nvrtcGetCUBIN");
  destroy_t destroy = (destroy_t)dlsym(lib, "nvrtcDestroyProgram");
  if (!create || !compile || !logsize || !getlog || !binsize || !getbin || !destroy) { log = "libnvrtc lacks the expected entry points"; return -1; }
  void* prog = nullptr;
  if (create(&prog, src.c_str(), "zq_jit.cu", 0, nullptr, nullptr) != 0) { log = "nvrtcCreateProgram failed"; return -2; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17"};
  const int rc = compile(prog, 3, opts);
  size_t ls = 0;
  if (logsize(prog, &ls) == 0 && ls > 1) { log.resize(ls); getlog(prog, &log[0]); }
  if (rc != 0) { destroy(&prog); return -3; }
  size_t bs = 0;
  if (binsize(prog, &bs) != 0 || bs == 0) { destroy(&prog); log += " (no cubin)"; return -4; }
  cubin.resize(bs);
  getbin(prog, cubin.data());
  destroy(&prog);
  return 0;
}

}  // namespace zq

// ---- C ABI (include/zq_b200.h) ------------------------------------------------------------------------------
#include <cstring>

#include "../../include/zq_b200.h"
#include "zq_cm_host.h"

extern "C" int zq_jit_context_source(const uint8_t* header, uint32_t header_len, char* src, uint32_t src_cap, uint32_t* src_len,
                                     char* errbuf, size_t errcap) {
  auto fail = [&](int rc, const std::string& m) {
    if (errbuf && errcap) snprintf(errbuf, errcap, "%s", m.c_str());
    return rc;
  };
  if (!header || !src_len) return fail(ZQ_E_ARG, "bad argument");
  try {
    size_t used = 0;
    zq::Assembled code = zq::parse_block_header(header, header_len, &used);
    std::string s, why;
    if (!zq::jit_context_source(code.hcomp.data(), code.hcomp.size(), code.hh, code.hm, code.ncomp, s, why))
      return fail(ZQ_E_UNSUPPORTED, why);
    *src_len = (uint32_t)s.size();
    if (src) {
      if (src_cap < s.size() + 1) return fail(ZQ_E_OUTPUT, "source buffer too small");
      memcpy(src, s.c_str(), s.size() + 1);
    }
    return ZQ_OK;
  } catch (const zq::Error& e) {
    return fail(ZQ_E_METHOD, e.msg);
  }
}

extern "C" int zq_jit_compile(const char* src, uint32_t* cubin_size, char* log, size_t logcap) {
  if (!src) return ZQ_E_ARG;
  std::vector<char> cubin;
  std::string l;
  const int rc = zq::jit_compile(src, cubin, l);
  if (log && logcap) snprintf(log, logcap, "%s", l.c_str());
  if (cubin_size) *cubin_size = (uint32_t)cubin.size();
  return rc == 0 ? ZQ_OK : rc == -1 ? ZQ_E_UNSUPPORTED : ZQ_E_METHOD;
}
