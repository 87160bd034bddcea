// Throughput-mode (CGG_BF16) side: bf16 operands on tcgen05 tensor cores, TMA-fed, fp32
// accumulation in TMEM.  Internal interface used by api.cu.
#pragma once
#include "../../include/cgg_b200.h"
#include <cuda_runtime.h>

namespace cgg {

struct TcState;

TcState* tc_create(const cgg_config& cfg);
void tc_destroy(TcState* t);
const char* tc_last_error(const TcState* t);
size_t tc_workspace_bytes(const TcState* t, int batch);  // t may be null (fp32 mode) -> 0

// bf16 copies of the K/V projection weights and bias tables + TMA descriptors for the given sizes.
int tc_prepare(TcState* t, const cgg_weights* w, int H4, int W4, const int* lh, const int* lw, const int* nl,
               float* const* wkv_f32, float* const* rk_f32, float* const* bkv_f32, cudaStream_t s);

// K4: kv[b, key, :] = mem[b, :, key]^T Wkv^T + bias tables, bf16 out (B, K_l, nl*2C)
int tc_kv_project(TcState* t, int level, int batch, const void* mem_bf16, void* kv_bf16, cudaStream_t s);

// K2 (+K3): mask[b,q,p] = sum_c me[b,q,c] F[b,c,p] (bf16 out); when target_level >= 0 also the
// attention-mask bitmap / all_masked flags of that level.
int tc_mask_einsum(TcState* t, int batch, const float* me_f32, const void* mask_features_bf16, void* mask_bf16,
                   int target_level, uint32_t* bitmap, uint8_t* all_masked, void* ws, cudaStream_t s);

// K5 with bf16 K/V
int tc_attention(TcState* t, int batch, int num_keys, const float* q, const void* k, const void* v, long kv_stride,
                 long kv_bstride, const uint32_t* bitmap, const uint8_t* all_masked, float* out, cudaStream_t s);

}  // namespace cgg
