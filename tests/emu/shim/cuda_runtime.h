// Host shim for the SIMT emulator build (tests only): stands in for <cuda_runtime.h>.
#pragma once
#include "../simt_emu.h"
