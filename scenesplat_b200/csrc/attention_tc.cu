// tcgen05 / TMEM patch attention -- under construction in this commit; the entry point validates its
// arguments and reports "bad arguments" until the kernel lands (ops.patch_attention routes to the SIMT
// kernel meanwhile).
#include "tc_common.cuh"
#include "../../include/scenesplat_b200.h"

extern "C" int ss_patch_attention(const void* qkv_bf16, const int64_t* order_row, const int32_t* table, int max_patches,
                                  int patch_size, int heads, int head_dim, float scale, void* out_bf16, void* stream_) {
  (void)qkv_bf16; (void)order_row; (void)table; (void)max_patches; (void)patch_size; (void)heads; (void)head_dim;
  (void)scale; (void)out_bf16; (void)stream_;
  return SS_BAD_ARGS;
}
