// Internal state of the bf16 / tcgen05 side (shared by gemm_tc.cu and attention_tc.cu).
#pragma once
#include "gemm_tc.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <string>

namespace cgg {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


struct TcState {
  cgg_config cfg;
  std::string err;
  EncodeTiledFn encode = nullptr;
  int H4 = 0, W4 = 0, lh[3] = {0, 0, 0}, lw[3] = {0, 0, 0}, nl[3] = {0, 0, 0};
  __nv_bfloat16* wkv[3] = {nullptr, nullptr, nullptr};   // (nl*2C, C) bf16
  __nv_bfloat16* rk[3] = {nullptr, nullptr, nullptr};    // (K_l, nl*C) bf16 key-bias table
  const float* bkv[3] = {nullptr, nullptr, nullptr};     // (nl*2C) fp32, owned by the handle
  // N tiling of the mask einsum / bits GEMMs
  int q_pad = 0, ein_ntile = 0, ein_calls_per_tile = 0, bits_ntile = 0, bits_nt = 0;
  int rows_per_batch = 0;
  bool smem_attr_set = false, attn_attr_set = false;
  uint8_t* live_buf = nullptr;          // per (image, query tile, key tile) "any key unmasked" flags
  size_t live_bytes = 0;
  void free_all() {
    for (int l = 0; l < 3; ++l) { cudaFree(wkv[l]); cudaFree(rk[l]); wkv[l] = rk[l] = nullptr; }
  }
};


inline int tc_fail(TcState* t, int code, const std::string& m) { t->err = m; return code; }

#define TCU(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return tc_fail(t, CGG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

}  // namespace cgg
