// zq_cm_types.h -- plain structs shared by host and device for the CM engine.
#pragma once
#include <stdint.h>

enum { ZQ_CONS = 1, ZQ_CM, ZQ_ICM, ZQ_MATCH, ZQ_AVG, ZQ_MIX2, ZQ_MIX, ZQ_ISSE, ZQ_SSE };
#define ZQ_CM_MAXCOMP 64   // components per model on the device; above ZQ_CM_LANES one lane evaluates them in turn (zq_cm_wide.cuh)
#define ZQ_CM_LANES 32

// One component as a lane sees it (descriptor bytes = ZPAQ COMP section, Z:13938 compsize).
struct ZqCmComp {
  uint8_t type, a1, a2, a3, a4, a5;  // cp[0..5]
  uint8_t level;                     // dependency depth: inputs are all at lower levels
  uint8_t pad;
  uint32_t cm_mask;                  // entries-1 of the u32 table (CM/SSE/MATCH index/MIX rows*m uses c)
  uint32_t ht_mask;                  // bytes-1 of the byte table (ICM/ISSE hash rows, MATCH buffer)
  uint64_t cm_off, ht_off;           // byte offsets inside the unit's model region
};

struct ZqCmPlan {
  int32_t n, nlevels;
  int32_t hh, hm;                    // HCOMP H (2^hh u32) and M (2^hm bytes)
  uint32_t hcomp_off, hcomp_len;     // HCOMP bytecode (incl. trailing 0) in the blob
  uint32_t fill_first, fill_count;   // init jobs in the fill table
  uint64_t m_off, h_off, r_off;      // VM memory inside the model region
  uint64_t model_bytes;              // size of one unit's model region (256 B aligned)
  uint32_t mix_mask;                 // lanes that are MIX components
  uint32_t chain;                    // 1: the model is ICM -> ISSE(0) (every built-in level-3 model): encoder fast path
  // decoder only: post-processor (PCOMP) machine of the block, Z:15358-15414
  int32_t ph, pm;                    // PCOMP H (2^ph u32) and M (2^pm bytes)
  uint64_t pm_off, ph_off, pr_off, pcode_off;
  uint64_t wide_off;                 // n > ZQ_CM_LANES: 64 B of per-component state each (zq_cm_wide.cuh)
  ZqCmComp comp[ZQ_CM_MAXCOMP];
};

// One initialisation job of a unit's model region.
enum { ZQ_FILL_ZERO = 0, ZQ_FILL_U32, ZQ_FILL_U16, ZQ_FILL_SSE, ZQ_FILL_ICM, ZQ_FILL_ISSE, ZQ_FILL_MATCHBUF };
struct ZqCmFill {
  uint64_t off, bytes;
  uint32_t kind, value;
};
