// Host-side helpers shared by schema.cpp and api.cu.
#pragma once
#include <stdint.h>

#include <string>

namespace nfx {
std::string rust_f32_display(float x);
int set_cols(int set_index);                         // columns of set 0..4 (flat() order)
int column_offset(uint32_t mask, uint32_t bit);      // first column of `bit` within `mask`, or -1
extern thread_local std::string g_thread_error;      // last error without a context
}  // namespace nfx
