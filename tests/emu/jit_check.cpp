// jit_check.cpp -- TEST INFRASTRUCTURE ONLY: steps a generated ZPAQL->C translation (zq_jit.cpp) against the device
// interpreter (zq_cm.cuh compiled for the host) over the same input bytes and compares registers, H and M.
// Compiled per program by tests/test_jit.py with -DZQ_JIT_GENERATED="\"<generated source>\"".
#include <cuda_runtime.h>   // the shim

#include "zq_cm.cuh"

#define ZQ_JIT_FN static inline
#include ZQ_JIT_GENERATED

using namespace zqdev;

// returns 0 if every step agrees, else 1 + index of the first differing byte; *errs = (interpreter error, translated error)
extern "C" int jit_check(const uint8_t* code, int len, int hh, int hm, const uint8_t* bytes, int n, int* errs) {
  const size_t msz = (size_t)1 << hm, hsz = (size_t)1 << hh;
  std::vector<u8> M1(msz, 0), M2(msz, 0), prog(code, code + len);
  prog.resize(len + 16);
  std::vector<u32> H1(hsz, 0), H2(hsz, 0), R1(256, 0), R2(256, 0);
  CmVm vm; memset(&vm, 0, sizeof vm);
  vm.m = M1.data(); vm.h = H1.data(); vm.r = R1.data(); vm.mmask = (u32)msz - 1; vm.hmask = (u32)hsz - 1;
  vm.code = prog.data(); vm.len = len;
  ZqJitVm jv; memset(&jv, 0, sizeof jv);
  int jerr = 0;
  for (int k = 0; k < n; ++k) {
    cm_vm_run_switch<false>(vm, bytes[k], nullptr);
    zq_hcomp(jv, M2.data(), H2.data(), R2.data(), bytes[k], jerr);
    errs[0] = vm.error; errs[1] = jerr;
    if ((vm.error != 0) != (jerr != 0)) return 1 + k;
    if (vm.error) return 0;     // both stopped with a ZPAQL error at the same byte
    if (vm.a != jv.a || vm.b != jv.b || vm.c != jv.c || vm.d != jv.d || (u32)vm.f != jv.f) return 1 + k;
    if (memcmp(H1.data(), H2.data(), hsz * 4) || memcmp(R1.data(), R2.data(), 1024)) return 1 + k;
  }
  return memcmp(M1.data(), M2.data(), msz) ? 1 + n : 0;
}
