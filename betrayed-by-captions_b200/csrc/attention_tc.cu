// K5 on tensor cores (throughput mode): masked multi-head cross-attention, flash-style, sm_100a.
//
// One CTA per (image, head, 128-query tile).  Keys stream in 128-key tiles:
//   S = Q K^T      tcgen05.mma  M=128 (queries) x N=128 (keys) x K=32 (head dim), S in TMEM
//   softmax        4 warps, thread = query row (TMEM lane): the row's 128 scores are read with
//                  tcgen05.ld, the attention-mask BITMAP words of that row are applied
//                  (bit = 1 -> -inf), row max / sum are thread-local (no shuffles), P is written
//                  to shared memory as bf16 in the UMMA core-matrix layout
//   O_j = P V      tcgen05.mma  M=128 x N=32 x K=128, into a fresh TMEM tile; the softmax thread
//                  folds it into its register accumulator with the online-softmax rescale
// K and V tiles arrive by TMA (64-byte swizzle, one head's 32 dims = 64-byte rows) straight from
// the projected K/V buffer.  Key tiles that are masked for EVERY query of the CTA are skipped
// (flags from live_tiles_kernel); rows with all_masked set ignore the bitmap (the reference's
// all-masked-row fallback, head.py:825-826).
//   warp 0: TMA producer   warp 1: MMA issuer (1 thread)   warps 2-9: softmax / epilogue -- two
//   groups of 4 warps, each owning 64 of the tile's 128 key columns with its OWN running max / sum /
//   output accumulator (merged once at the end), so the groups never synchronise per tile
#include "kernels.h"
#include "tc_ptx.cuh"
#include "tc_state.h"

namespace cgg {

namespace {

constexpr int AT_THREADS = 64 + 256;       // TMA warp + MMA warp + 8 softmax warps (two column halves)
constexpr int AT_KT = 128;                   // keys per tile
constexpr int AT_STAGES = 4;
constexpr int AT_KV_TILE_BYTES = AT_KT * 64; // 128 keys x 32 dims x bf16 = 8 KB
constexpr int AT_STAGE_BYTES = 4 * AT_KV_TILE_BYTES;   // K tile, V tile, key-bias (R) tile as hi + lo
constexpr int AT_Q_BYTES = 128 * 64;         // 128 queries x 32 dims bf16, core-matrix layout
constexpr int AT_P_BYTES = 128 * AT_KT * 2;  // 32 KB per P buffer
constexpr int AT_ONES_BYTES = 64 * 64;      // 64 key rows x 64 B: column 0 = 1.0 (row sums come out of the PV MMA)
constexpr float LOG2E = 1.4426950408889634f;

struct AttnP {
  const float* q; float* out; __nv_bfloat16* out_bf16;
  const uint32_t* bitmap; const uint8_t* all_masked; const uint8_t* live;
  int Q, K, heads, W32, ntiles, nqt;
  int has_r, r_col0, r_lo_off;   // key-bias table term: S += Q (R_hi + R_lo)^T, R columns r_col0 + head*32
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Shared-memory descriptor with an explicit layout type (0 = no swizzle / core-matrix interleave,
// 4 = SWIZZLE_64B); same bit layout as ptx::umma_desc_sw128.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int next_live(const uint8_t* live, int t, int n) {
  while (t < n && !live[t]) ++t;
  return t;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmR, const __grid_constant__ AttnP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                   // 8 KB
  uint8_t* sP = sQ + AT_Q_BYTES;                        // 2 x 32 KB
  uint8_t* sKV = sP + 2 * AT_P_BYTES;                   // stages x (K tile, V tile)
  uint8_t* sOnes = sKV + AT_STAGES * AT_STAGE_BYTES;   // above every V tile (descriptor offsets are unsigned)
  uint8_t* sLive = sOnes + AT_ONES_BYTES;               // this CTA's live-tile flags (<= 512 key tiles)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLive + 512);
  uint64_t* kv_full = bars;
  uint64_t* kv_empty = kv_full + AT_STAGES;
  uint64_t* s_full = kv_empty + AT_STAGES;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.heads, qt = blockIdx.x / p.heads, b = blockIdx.y;
  const uint8_t* live_g = p.live + ((long)b * p.nqt + qt) * p.ntiles;
  const uint8_t* live = sLive;
  const int C = p.heads * 32;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    if (p.has_r) ptx::prefetch_tmap(&tmR);
    for (int i = 0; i < AT_STAGES; ++i) { ptx::mbar_init(&kv_full[i], 1); ptx::mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&p_full[i], 256); ptx::mbar_init(&o_full[i], 1); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // barrier init / TMEM allocation above overlapped the previous kernel's tail
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  if (warp >= 2) {
    // "ones" tile, an extra N-group of the PV B operand in the V tiles' layout (64-byte rows, 64-byte
    // swizzle: 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)): column 0 of every key row is
    // 1.0, so column 32 of O = P [V | 1] is the row sum of the bf16 P the tensor core actually used.
    const int tid = threadIdx.x - 64;                       // 0..255
    uint4 z = make_uint4(0, 0, 0, 0);
    const int r = tid >> 2, cpos = tid & 3;                 // one 16-byte chunk per thread: 64 rows x 4 chunks
    if (cpos == ((r >> 1) & 3)) z.x = 0x00003F80u;          // bf16 1.0 in element 0 of logical chunk 0
    *reinterpret_cast<uint4*>(sOnes + r * 64 + cpos * 16) = z;
    fence_async_smem();
  }
  // the live-tile flags are read by three roles every tile: one copy into shared memory
  for (int i = threadIdx.x; i < p.ntiles; i += AT_THREADS) sLive[i] = live_g[i];
  if (warp >= 2 && warp < 6) {
    // Q tile -> bf16, core-matrix (no-swizzle) K-major layout: element (row, d) at
    // (row/8)*512 + (d/8)*128 + (row%8)*16 + (d%8)*2
    const int row = (warp & 3) * 32 + lane, qi = qt * 128 + row;
    float qv[32];
    if (qi < p.Q) {
      const float4* src = reinterpret_cast<const float4*>(p.q + ((long)b * p.Q + qi) * C + h * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 x = __ldg(src + i);
        qv[4 * i] = x.x; qv[4 * i + 1] = x.y; qv[4 * i + 2] = x.z; qv[4 * i + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) qv[i] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 pk;
      uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 v2 = __floats2bfloat162_rn(qv[8 * c + 2 * j], qv[8 * c + 2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&v2);
      }
      *reinterpret_cast<uint4*>(sQ + (row >> 3) * 512 + c * 128 + (row & 7) * 16) = pk;
    }
    fence_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;    // 2 buffers x 2 column halves x 64-column slots (48 used: 32 dims, sum, pad)

  if (warp == 0) {
    if (lane == 0) {
      // ------------- TMA producer: K and V tiles of every live key tile
      int it = 0;
      for (int t = next_live(live, 0, p.ntiles); t < p.ntiles; t = next_live(live, t + 1, p.ntiles), ++it) {
        const int s = it % AT_STAGES;
        const uint32_t ph = (uint32_t)(it / AT_STAGES) & 1u;
        ptx::mbar_wait(&kv_empty[s], ph ^ 1u);
        ptx::mbar_expect_tx(&kv_full[s], (p.has_r ? 4 : 2) * AT_KV_TILE_BYTES);
        uint8_t* st = sKV + s * AT_STAGE_BYTES;
        if (p.has_r) {
          ptx::tma_load_2d(st + 2 * AT_KV_TILE_BYTES, &tmR, &kv_full[s], p.r_col0 + h * 32, t * AT_KT);
          ptx::tma_load_2d(st + 3 * AT_KV_TILE_BYTES, &tmR, &kv_full[s], p.r_lo_off + p.r_col0 + h * 32, t * AT_KT);
        }
        ptx::tma_load_3d(st, &tmK, &kv_full[s], h * 32, t * AT_KT, b);
        ptx::tma_load_3d(st + AT_KV_TILE_BYTES, &tmV, &kv_full[s], h * 32, t * AT_KT, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------- MMA issuer
      const uint32_t idesc_s = ptx::umma_idesc_bf16(128, AT_KT, false, false);   // S = Q K^T
      const uint32_t idesc_o = ptx::umma_idesc_bf16(128, 48, false, true);       // O = P [V | 1] (N-major B, 2 N-groups)
      const uint32_t ones_addr = ptx::smem_u32(sOnes);
      const uint32_t q_addr = ptx::smem_u32(sQ);
      int n_live = 0;
      for (int t = 0; t < p.ntiles; ++t) n_live += live[t] ? 1 : 0;
      auto issue_qk = [&](int j) {
        const int s = j % AT_STAGES;
        ptx::mbar_wait(&kv_full[s], (uint32_t)(j / AT_STAGES) & 1u);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(sKV + s * AT_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t adesc = umma_desc(q_addr + k * 256, 128, 512, 0);          // Q: core matrices
          const uint64_t bdesc = umma_desc(k_addr + k * 32, 16, 512, 4);            // K tile: SW64, K-major
          ptx::mma_bf16_ss(tmem_S + (uint32_t)((j & 1) * 128), adesc, bdesc, idesc_s, k);
        }
        if (p.has_r) {
          // + Q R^T: positional / level / bias part of the keys, a batch-independent table that the K/V
          // projection therefore never has to add (K = Wk x + R  =>  q.K = q.(Wk x) + q.R)
          // (R is kept as a bf16 hi/lo pair so that the table itself carries no bf16 rounding)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = umma_desc(q_addr + (k & 1) * 256, 128, 512, 0);
            const uint64_t bdesc = umma_desc(k_addr + (2 + (k >> 1)) * AT_KV_TILE_BYTES + (k & 1) * 32, 16, 512, 4);
            ptx::mma_bf16_ss(tmem_S + (uint32_t)((j & 1) * 128), adesc, bdesc, idesc_s, 1u);
          }
        }
        ptx::mma_commit(&s_full[j & 1]);
      };
      if (n_live > 0) issue_qk(0);
      if (n_live > 1) issue_qk(1);
      for (int j = 0; j < n_live; ++j) {
        const int s = j % AT_STAGES;
        ptx::mbar_wait(&p_full[j & 1], (uint32_t)(j >> 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t p_addr = ptx::smem_u32(sP + (j & 1) * AT_P_BYTES);
        const uint32_t v_addr = ptx::smem_u32(sKV + s * AT_STAGE_BYTES + AT_KV_TILE_BYTES);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          // O_half = P[:, 64 keys of this half] . V[those keys, :]  (each half has its own softmax reference)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = umma_desc(p_addr + hf * 1024 + k * 256, 128, 2048, 0);         // P: core matrices
            // V: SW64, N-major; the leading-dimension offset reaches from this V slice to the ones tile
            const uint32_t vk = v_addr + hf * 4096 + k * 1024;
            const uint64_t bdesc = umma_desc(vk, (ones_addr + k * 1024) - vk, 512, 4);
            ptx::mma_bf16_ss(tmem_O + (uint32_t)(((j & 1) * 2 + hf) * 64), adesc, bdesc, idesc_o, k);
          }
        }
        ptx::mma_commit(&o_full[j & 1]);
        ptx::mma_commit(&kv_empty[s]);
        if (j + 2 < n_live) issue_qk(j + 2);
      }
    }
  } else {
    // ------------- softmax / epilogue: thread = query row
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;             // which 64 key columns of every tile
    const int row = quarter * 32 + lane, qi = qt * 128 + row;
    const bool row_ok = qi < p.Q;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool ignore_mask = !row_ok || p.bitmap == nullptr || (p.all_masked && p.all_masked[(long)b * p.Q + qi]);
    const uint32_t* brow = p.bitmap ? p.bitmap + ((long)b * p.Q + (row_ok ? qi : 0)) * p.W32 : nullptr;
    float m_run = -INFINITY;
    float o[33];                                  // 32 output dims + the running row sum (column 32 of O)
#pragma unroll
    for (int i = 0; i < 33; ++i) o[i] = 0.f;
    auto fold_o = [&](int pb, float scale) {      // o = (o + O_tile) * scale
      float ov[32];
      const uint32_t oa = tmem_O + (uint32_t)((pb * 2 + hf) * 64) + lane_off;
      tmem_ld32(oa, ov);
      uint32_t sum_bits;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(sum_bits) : "r"(oa + 32));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = (o[i] + ov[i]) * scale;
      o[32] = (o[32] + __uint_as_float(sum_bits)) * scale;
    };
    // mask words of this row for this half's 64 keys of tile t (bit = 1 -> masked); keys >= K are masked.
    // They are fetched one live tile AHEAD so the global-load latency hides behind the current tile.
    auto load_mask = [&](int t, uint32_t* mw) {
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int widx = t * 4 + hf * 2 + w;
        uint32_t word = 0u;
        if (!ignore_mask && widx < p.W32) word = __ldg(brow + widx);
        const int k0 = t * AT_KT + (hf * 2 + w) * 32;
        if (k0 + 32 > p.K) word |= (k0 >= p.K) ? 0xffffffffu : ~((1u << (p.K - k0)) - 1u);
        mw[w] = word;
      }
    };
    int it = 0;
    uint32_t mw_next[2] = {0u, 0u};
    {
      const int t0 = next_live(live, 0, p.ntiles);
      if (t0 < p.ntiles) load_mask(t0, mw_next);
    }
    for (int t = next_live(live, 0, p.ntiles); t < p.ntiles; ++it) {
      const int buf = it & 1;
      const int t_next = next_live(live, t + 1, p.ntiles);
      uint32_t mw[2] = {mw_next[0], mw_next[1]};
      if (t_next < p.ntiles) load_mask(t_next, mw_next);
      ptx::mbar_wait(&s_full[buf], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t s_addr = tmem_S + (uint32_t)(buf * 128 + hf * 64) + lane_off;
      // the 64 scores are read from TMEM ONCE, masked to -inf in registers, and reused for max and exp
      float v[64];
      tmem_ld32(s_addr, v);
      tmem_ld32(s_addr + 32, v + 32);
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        if ((mw[i >> 5] >> (i & 31)) & 1u) v[i] = -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
      const float m_new = fmaxf(m_run, mx);
      const float scale = (m_new == -INFINITY) ? 1.f : fast_exp2((m_run - m_new) * LOG2E);
      const float mneg = (m_new == -INFINITY) ? 0.f : -m_new * LOG2E;
      m_run = m_new;
      // p = exp(s - m) as bf16 into the P buffer (core-matrix layout, 16 B = 8 keys per store); exp2(-inf) = 0
      uint8_t* prow = sP + buf * AT_P_BYTES + (row >> 3) * 2048 + (row & 7) * 16 + hf * 1024;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 pk;
        uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i0 = g * 8 + 2 * j;
          __nv_bfloat162 v2 = __floats2bfloat162_rn(fast_exp2(fmaf(v[i0], LOG2E, mneg)), fast_exp2(fmaf(v[i0 + 1], LOG2E, mneg)));
          w[j] = *reinterpret_cast<uint32_t*>(&v2);
        }
        *reinterpret_cast<uint4*>(prow + g * 128) = pk;
      }
      fence_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[buf]);
      // fold the previous tile's P [V | 1] into the register accumulator, then apply this tile's rescale
      if (it > 0) {
        ptx::mbar_wait(&o_full[(it - 1) & 1], (uint32_t)((it - 1) >> 1) & 1u);
        ptx::tc_fence_after();
        fold_o((it - 1) & 1, scale);
      }
      t = t_next;
    }
    if (it > 0) {
      ptx::mbar_wait(&o_full[(it - 1) & 1], (uint32_t)((it - 1) >> 1) & 1u);
      ptx::tc_fence_after();
      fold_o((it - 1) & 1, 1.f);
    }
    float l_run = o[32];
    // merge the two column halves of the row: half 1 publishes (m, l, o) through shared memory
    // (the P buffers are idle once the last PV has been consumed), half 0 combines and stores
    float* xch = reinterpret_cast<float*>(sP) + row * 35;
    if (hf == 1) {
      xch[0] = m_run;
      xch[1] = l_run;
#pragma unroll
      for (int i = 0; i < 32; ++i) xch[2 + i] = o[i];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hf == 0) {
      const float m1 = xch[0], l1 = xch[1];
      const float m = fmaxf(m_run, m1);
      const float a0 = (m_run == -INFINITY) ? 0.f : fast_exp2((m_run - m) * LOG2E);
      const float a1 = (m1 == -INFINITY) ? 0.f : fast_exp2((m1 - m) * LOG2E);
      l_run = l_run * a0 + l1 * a1;
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = o[i] * a0 + xch[2 + i] * a1;
    }
    if (row_ok && hf == 0) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      if (p.out) {
        float4* dst = reinterpret_cast<float4*>(p.out + ((long)b * p.Q + qi) * C + h * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(o[4 * i] * inv, o[4 * i + 1] * inv, o[4 * i + 2] * inv, o[4 * i + 3] * inv);
      }
      if (p.out_bf16) {
        uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + ((long)b * p.Q + qi) * C + h * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 pk;
          uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 v2 = __floats2bfloat162_rn(o[8 * i + 2 * j] * inv, o[8 * i + 2 * j + 1] * inv);
            w[j] = *reinterpret_cast<uint32_t*>(&v2);
          }
          dst[i] = pk;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// live[b][qt][tile] = 1 iff some query row of the tile attends to some key of the key tile
// (rows under the all-masked fallback attend everywhere).
__global__ void __launch_bounds__(128) live_tiles_kernel(const uint32_t* __restrict__ bitmap,
                                                         const uint8_t* __restrict__ all_masked, int Q, int K, int W32,
                                                         int ntiles, int nqt, uint8_t* __restrict__ live) {
  ptx::grid_dep_launch();
  ptx::grid_dep_wait();
  const int t = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;
  const int qi = qt * 128 + threadIdx.x;
  int any = 0;
  if (qi < Q) {
    if (bitmap == nullptr || (all_masked && all_masked[(long)b * Q + qi])) {
      any = 1;
    } else {
      const uint32_t* brow = bitmap + ((long)b * Q + qi) * W32;
      for (int w = 0; w < 4; ++w) {
        const int widx = t * 4 + w, k0 = t * AT_KT + w * 32;
        if (widx >= W32 || k0 >= K) break;
        uint32_t valid = (k0 + 32 > K) ? ((1u << (K - k0)) - 1u) : 0xffffffffu;
        if ((~brow[widx]) & valid) any = 1;
      }
    }
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) live[((long)b * nqt + qt) * ntiles + t] = any ? 1 : 0;
}

int make_map_kv(TcState* t, CUtensorMap* m, const void* base, int K, long kv_stride, long kv_bstride, int B, int C) {
  if ((kv_stride * 2) % 16 != 0 || (kv_bstride * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15))
    return tc_fail(t, CGG_ERR_UNSUPPORTED, "K/V rows must be 16-byte aligned");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)K, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)kv_stride * 2, (cuuint64_t)kv_bstride * 2};
  cuuint32_t box[3] = {32, AT_KT, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = t->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tc_fail(t, CGG_ERR_CUDA, "cuTensorMapEncodeTiled(KV) failed: " + std::to_string((int)r));
  return CGG_OK;
}

}  // namespace

int tc_attention(TcState* t, int batch, int num_keys, const float* q, const void* k, const void* v, long kv_stride,
                 long kv_bstride, const uint32_t* bitmap, const uint8_t* all_masked, float* out, __nv_bfloat16* out_bf16,
                 cudaStream_t s, const void* r_table, long r_cols, int r_col0) {
  // r_table: (num_keys, r_cols) bf16 with columns [hi (r_cols/2) | lo (r_cols/2)]
  const int Q = t->cfg.num_queries, heads = t->cfg.num_heads, C = t->cfg.embed_dim;
  const int ntiles = (num_keys + AT_KT - 1) / AT_KT, nqt = (Q + 127) / 128;
  if ((size_t)batch * nqt * ntiles > t->live_bytes) return tc_fail(t, CGG_ERR_BAD_SHAPE, "too many key tiles");
  CUtensorMap mK, mV;
  int st = make_map_kv(t, &mK, k, num_keys, kv_stride, kv_bstride, batch, C);
  if (st != CGG_OK) return st;
  st = make_map_kv(t, &mV, v, num_keys, kv_stride, kv_bstride, batch, C);
  if (st != CGG_OK) return st;
  CUtensorMap mR = mK;
  if (r_table) {
    cuuint64_t dims[2] = {(cuuint64_t)r_cols, (cuuint64_t)num_keys};
    cuuint64_t strides[1] = {(cuuint64_t)r_cols * 2};
    cuuint32_t box[2] = {32, AT_KT};
    cuuint32_t es[2] = {1, 1};
    CUresult r = t->encode(&mR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(r_table), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return tc_fail(t, CGG_ERR_CUDA, "cuTensorMapEncodeTiled(R) failed: " + std::to_string((int)r));
  }
  const int W32 = (num_keys + 31) / 32;
  TCU(launch_pdl(live_tiles_kernel, dim3(ntiles, nqt, batch), dim3(128), 0, s, bitmap, all_masked, Q, num_keys, W32, ntiles, nqt,
                 t->live_buf));
  count_launch();
  TCU(cudaGetLastError());
  AttnP p;
  p.q = q; p.out = out; p.out_bf16 = out_bf16; p.bitmap = bitmap; p.all_masked = all_masked; p.live = t->live_buf;
  p.Q = Q; p.K = num_keys; p.heads = heads; p.W32 = W32; p.ntiles = ntiles; p.nqt = nqt;
  p.has_r = r_table ? 1 : 0; p.r_col0 = r_col0; p.r_lo_off = (int)(r_cols / 2);
  if (ntiles > 512) return tc_fail(t, CGG_ERR_BAD_SHAPE, "more than 512 key tiles");
  const size_t smem = 1024 + AT_Q_BYTES + 2 * AT_P_BYTES + AT_STAGES * AT_STAGE_BYTES + AT_ONES_BYTES + 512 + (2 * AT_STAGES + 6) * 8 + 16;
  if (!t->attn_attr_set) {
    TCU(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    t->attn_attr_set = true;
  }
  TCU(launch_pdl(attention_tc_kernel, dim3(heads * nqt, batch), dim3(AT_THREADS), smem, s, mK, mV, mR, p));
  count_launch();
  TCU(cudaGetLastError());
  return CGG_OK;
}

}  // namespace cgg
