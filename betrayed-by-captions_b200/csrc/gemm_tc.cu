#include "gemm_tc.h"
#include <string>
namespace cgg {
struct TcState { cgg_config cfg; std::string err; };
TcState* tc_create(const cgg_config& cfg) { TcState* t = new TcState(); t->cfg = cfg; return t; }
void tc_destroy(TcState* t) { delete t; }
const char* tc_last_error(const TcState* t) { return t ? t->err.c_str() : ""; }
size_t tc_workspace_bytes(const TcState* t, int) { (void)t; return 0; }
int tc_prepare(TcState* t, const cgg_weights*, int, int, const int*, const int*, const int*, float* const*, float* const*, float* const*, cudaStream_t) { t->err = "bf16 path not built yet"; return CGG_ERR_UNSUPPORTED; }
int tc_kv_project(TcState* t, int, int, const void*, void*, cudaStream_t) { t->err = "bf16 path not built yet"; return CGG_ERR_UNSUPPORTED; }
int tc_mask_einsum(TcState* t, int, const float*, const void*, void*, int, uint32_t*, uint8_t*, void*, cudaStream_t) { t->err = "bf16 path not built yet"; return CGG_ERR_UNSUPPORTED; }
int tc_attention(TcState* t, int, int, const float*, const void*, const void*, long, long, const uint32_t*, const uint8_t*, float*, cudaStream_t) { t->err = "bf16 path not built yet"; return CGG_ERR_UNSUPPORTED; }
}
