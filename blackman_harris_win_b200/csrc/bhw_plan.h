// bhw_plan.h - CUDA-free planning helpers (see bhw_plan.cpp).
#pragma once
#include <vector>

#include "bhw_device.cuh"

namespace bhw {

SrcParams canonical_source(const SrcParams& sp, uint32_t* drop);
bool fast32_ok(const SrcParams& sp);
void build_taylor_rom(int dw, int lut, std::vector<I2>& rom);
void fill_fast_rec(const WinParams& wp, WinRec& r);
bool fast_tail_exact(const WinParams& wp);

}  // namespace bhw
