// zq_cm_host.h -- host side of the context-mixing engine: the model-independent lookup tables
// (Predictor::init, Z:14926-14941; formulas Z:61905-61909; state table generator Z:61750-61850 = the
// ZPAQ specification's) and the per-method device layout of the component tables.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "zq_cm_types.h"
#include "zq_config.h"

namespace zq {

// 78 KiB of tables shared by every block (staged to shared memory by the kernels).
struct CmTables {
  int16_t stretch[32768];
  uint16_t squash[4096];
  int32_t dt[1024];
  int32_t dt2k[256];
  uint8_t ns[1024];
  uint32_t icm_init[256];   // StateTable::cminit(j)
  uint32_t isse_init[512];  // {1<<15, clamp512k(stretch(cminit(j)>>8)*1024)}
};
// Builds the tables and checks the reference's own checksums (Z:14949-14950); throws zq::Error.
const CmTables& cm_tables();

// Device layout of one modeled method. Throws zq::Error for configurations without a device path.
ZqCmPlan make_cm_plan(const Assembled& code, std::vector<ZqCmFill>& fills);
// Decoder: append the post-processor's memory (M 2^pm, H 2^ph, R, 64 KiB code) to the model region.
void add_pcomp_region(ZqCmPlan& p, int ph, int pm, std::vector<ZqCmFill>& fills);
// Parse the COMP/HCOMP header of a block as stored in the archive (ZPAQL::read, Z:14104) into `code`.
Assembled parse_block_header(const uint8_t* hdr, size_t avail, size_t* consumed);

}  // namespace zq
