// The step after the path at test time: final mask upsample + fusion-head scoring.
//   reference: simple_test upsamples the last head call's mask logits to the padded input size with
//   F.interpolate(bilinear, align_corners=False) (open_set/models/mask2former_head.py:957-964) -- 6.7 GB of fp32 per
//   16-image batch at 1024^2 -- and MaskFormerFusionHeadOpen then crops to img_shape, optionally resamples to ori_shape
//   (maskformer_fusion_head.py:412-425) and, per kept (query, class) pair, thresholds `> 0`, averages sigmoid over the
//   positive pixels and takes the bounding box (instance_postprocess_emb, :318-366; mmdet mask2bbox).
// instance_mask_stats_kernel does all of it in ONE pass that reads only the (B, Q, H/4, W/4) logits: the full-resolution
// logits are never written.  Both resampling stages follow ATen's upsample_bilinear2d index math and operation order
// (src = scale (dst + 0.5) - 0.5 clamped at 0; v = h0 (w0 a + w1 b) + h1 (w0 c + w1 d)), evaluated in fp32.
#include "kernels.h"
#include <cuda_bf16.h>
#include <limits.h>
#include <math.h>

namespace cgg {

namespace {

struct Axis { int i0, i1; float l0, l1; };

__device__ __forceinline__ Axis src_axis(float scale, int dst, int in_size) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Axis a;
  a.i0 = (int)src;
  if (a.i0 > in_size - 1) a.i0 = in_size - 1;
  a.i1 = a.i0 + ((a.i0 < in_size - 1) ? 1 : 0);
  a.l1 = src - (float)a.i0;
  a.l0 = 1.f - a.l1;
  return a;
}

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// value of the first-stage upsample (logits (h4, w4) -> (up_h, up_w)) at (y, x) given precomputed axes
template <typename T>
__device__ __forceinline__ float up1(const T* __restrict__ src, int w4, const Axis& ay, const Axis& ax) {
  const float a = ldf(src + (long)ay.i0 * w4 + ax.i0), b = ldf(src + (long)ay.i0 * w4 + ax.i1);
  const float c = ldf(src + (long)ay.i1 * w4 + ax.i0), d = ldf(src + (long)ay.i1 * w4 + ax.i1);
  return ay.l0 * (ax.l0 * a + ax.l1 * b) + ay.l1 * (ax.l0 * c + ax.l1 * d);
}

// materialised first stage: out (B*Q, up_h, up_w) fp32 = F.interpolate(logits, (up_h, up_w))   (head.py:957-964)
template <typename T>
__global__ void __launch_bounds__(256) upsample_kernel(const T* __restrict__ logits, float* __restrict__ out, int h4, int w4,
                                                       int up_h, int up_w) {
  const int x = blockIdx.x * 64 + (threadIdx.x & 63), y0 = (blockIdx.y * 4 + (threadIdx.x >> 6)) * 8;
  const long plane = blockIdx.z;
  if (x >= up_w) return;
  const float sy = (float)h4 / (float)up_h, sx = (float)w4 / (float)up_w;
  const Axis ax = src_axis(sx, x, w4);
  const T* src = logits + plane * (long)h4 * w4;
  for (int y = y0; y < y0 + 8 && y < up_h; ++y) {
    const Axis ay = src_axis(sy, y, h4);
    out[(plane * up_h + y) * (long)up_w + x] = up1(src, w4, ay, ax);
  }
}

// geom[b] = {crop_h, crop_w, out_h, out_w}: crop of the upsampled map (img_shape), final size (ori_shape when rescaling,
// else the crop).  One warp = one 32-pixel column strip over ROWS_PER_WARP output rows of one (image, query) plane.
constexpr int ROWS_PER_BLOCK = 64;
template <typename T>
__global__ void __launch_bounds__(256) instance_mask_stats_kernel(const T* __restrict__ logits, const int* __restrict__ geom,
                                                                  int Q, int h4, int w4, int up_h, int up_w,
                                                                  uint32_t* __restrict__ bits, long bits_plane, int bits_w32,
                                                                  int* __restrict__ count, float* __restrict__ sig_sum,
                                                                  int* __restrict__ bbox) {
  const int plane = blockIdx.z, b = plane / Q;
  const int crop_h = geom[4 * b], crop_w = geom[4 * b + 1], out_h = geom[4 * b + 2], out_w = geom[4 * b + 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip = blockIdx.x * 8 + warp, x = strip * 32 + lane;
  const int y_begin = blockIdx.y * ROWS_PER_BLOCK;
  if (strip * 32 >= out_w || y_begin >= out_h) return;
  const bool rescale = out_h != crop_h || out_w != crop_w;
  const float s1y = (float)h4 / (float)up_h, s1x = (float)w4 / (float)up_w;
  const float s2y = (float)crop_h / (float)out_h, s2x = (float)crop_w / (float)out_w;
  const bool x_ok = x < out_w;
  const int xc = x_ok ? x : out_w - 1;
  // x axes: second stage (out -> crop) and, for its two source columns, first stage (up -> logits)
  Axis bx = {xc, xc, 1.f, 0.f};
  if (rescale) bx = src_axis(s2x, xc, crop_w);
  const Axis ax0 = src_axis(s1x, bx.i0, w4), ax1 = src_axis(s1x, bx.i1, w4);
  const T* src = logits + (long)plane * h4 * w4;
  int cnt = 0, xmin = INT_MAX, xmax = -1, ymin = INT_MAX, ymax = -1;
  float ssum = 0.f;
  for (int y = y_begin; y < y_begin + ROWS_PER_BLOCK && y < out_h; ++y) {
    float v;
    if (rescale) {
      const Axis by = src_axis(s2y, y, crop_h);
      const Axis ay0 = src_axis(s1y, by.i0, h4), ay1 = src_axis(s1y, by.i1, h4);
      const float a = up1(src, w4, ay0, ax0), bb = up1(src, w4, ay0, ax1);
      const float c = up1(src, w4, ay1, ax0), d = up1(src, w4, ay1, ax1);
      v = by.l0 * (bx.l0 * a + bx.l1 * bb) + by.l1 * (bx.l0 * c + bx.l1 * d);
    } else {
      v = up1(src, w4, src_axis(s1y, y, h4), ax0);
    }
    const bool pos = x_ok && v > 0.f;
    const uint32_t word = __ballot_sync(0xffffffffu, pos);
    if (bits && lane == 0) bits[(long)plane * bits_plane + (long)y * bits_w32 + strip] = word;
    if (pos) {
      ++cnt;
      ssum += 1.0f / (1.0f + expf(-v));
      xmin = min(xmin, x); xmax = max(xmax, x);
      ymin = min(ymin, y); ymax = max(ymax, y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if (lane == 0 && cnt > 0) {
    atomicAdd(count + plane, cnt);
    atomicAdd(sig_sum + plane, ssum);
    atomicMin(bbox + 4 * plane, xmin);
    atomicMin(bbox + 4 * plane + 1, ymin);
    atomicMax(bbox + 4 * plane + 2, xmax + 1);
    atomicMax(bbox + 4 * plane + 3, ymax + 1);
  }
}

__global__ void stats_init_kernel(int* count, float* sig_sum, int* bbox, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  count[i] = 0;
  sig_sum[i] = 0.f;
  bbox[4 * i] = INT_MAX; bbox[4 * i + 1] = INT_MAX; bbox[4 * i + 2] = 0; bbox[4 * i + 3] = 0;
}
// empty masks: mmdet mask2bbox leaves the box at zeros
__global__ void stats_finish_kernel(const int* count, int* bbox, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (count[i] == 0) { bbox[4 * i] = 0; bbox[4 * i + 1] = 0; }
}

// in-place softmax over rows of length n (get_cls_emb_scores, maskformer_fusion_head.py:297-315), one warp per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, int rows, int n) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = x + (long)row * n;
  float m = -INFINITY;
  for (int i = lane; i < n; i += 32) m = fmaxf(m, r[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += expf(r[i] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int i = lane; i < n; i += 32) r[i] = expf(r[i] - m) / s;
}

}  // namespace

cudaError_t launch_upsample_masks(const void* logits, bool bf16, float* out, int planes, int h4, int w4, int up_h, int up_w,
                                  cudaStream_t s) {
  if (planes <= 0) return cudaSuccess;
  dim3 grid((up_w + 63) / 64, (up_h + 31) / 32, planes);
  if (bf16) upsample_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), out, h4, w4, up_h, up_w);
  else upsample_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(logits), out, h4, w4, up_h, up_w);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_instance_mask_stats(const void* logits, bool bf16, const int* geom, int B, int Q, int h4, int w4, int up_h,
                                       int up_w, int max_out_h, int max_out_w, uint32_t* bits, int* count, float* sig_sum,
                                       int* bbox, cudaStream_t s) {
  const int n = B * Q;
  if (n <= 0) return cudaSuccess;
  stats_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(count, sig_sum, bbox, n);
  const int w32 = (max_out_w + 31) / 32;
  dim3 grid((w32 + 7) / 8, (max_out_h + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK, n);
  const long plane = (long)max_out_h * w32;
  if (bf16)
    instance_mask_stats_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), geom, Q, h4, w4,
                                                                   up_h, up_w, bits, plane, w32, count, sig_sum, bbox);
  else
    instance_mask_stats_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(logits), geom, Q, h4, w4, up_h, up_w, bits,
                                                           plane, w32, count, sig_sum, bbox);
  stats_finish_kernel<<<(n + 255) / 256, 256, 0, s>>>(count, bbox, n);
  count_launch(3);
  return cudaGetLastError();
}

cudaError_t launch_softmax_rows(float* x, int rows, int n, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, rows, n);
  count_launch();
  return cudaGetLastError();
}

}  // namespace cgg
