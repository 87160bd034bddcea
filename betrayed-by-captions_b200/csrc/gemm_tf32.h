// Training-step contractions on tcgen05 (kind::tf32) -- internal interface used by api.cu; see gemm_tf32.cu.
#pragma once
#include "kernels.h"

namespace cgg {

struct Tf32Ctx;
Tf32Ctx* tf32_create();
void tf32_destroy(Tf32Ctx* t);
const char* tf32_last_error(const Tf32Ctx* t);
// The GemmF32 contract on tensor cores.  0 = launched; 1 = the operand strides cannot be described to TMA (or A2 is
// set): the caller runs launch_gemm_f32 instead; < 0 = error.
int launch_gemm_tf32(Tf32Ctx* t, const GemmF32& g, cudaStream_t s);

}  // namespace cgg
