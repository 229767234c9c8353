// CTA-pair (tcgen05 cta_group::2) GEMMs of Block.forward with the neighbouring elementwise work in the epilogue
// (vit_models/dynamic_vit.py:263-283, Mlp.forward :159-175):
//
//   MODE_LN   x' = x + (A @ W^T + bias);  h = LayerNorm(x') * gamma + beta          (attn.proj / mlp.fc2 + residual + next norm)
//   MODE_ACT  out = act(A @ W^T + bias)                                             (mlp.fc1 + GELU)
//
// bf16 operands, fp32 accumulation in TMEM, every intermediate rounded to bf16 exactly where the reference's separate
// Linear / add / LayerNorm kernels round (Linear output, residual sum), LayerNorm statistics in fp32.
//
// Two CTAs of a cluster (one TPC) work on a 256-row tile: each CTA loads its own 128 rows of A and HALF of the W tile;
// one thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads A and W from both CTAs' shared
// memory and leaves each CTA's 128 x N accumulator in its own TMEM.  Against one-CTA tiles this halves the W traffic
// from L2 per FLOP (the L2 -> SM path, ~42 B/clk/SM, is what bounds a 128-row tile at K = 384) and the shared-memory
// operand reads per MMA.
//
//   warp 0     TMA producer (both CTAs): A 128x64 + W (N/2)x64 bf16 boxes, SWIZZLE_128B, ring of stages; completion is
//              signalled on the LEADER's `full` barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1     TMEM allocation (both CTAs); MMA issue (leader only); tcgen05.commit multicast frees the ring slot in
//              both CTAs and publishes the accumulator to both epilogues
//   warps 2-9  epilogue (both CTAs), two warps per TMEM lane quadrant splitting the columns
//
// MODE_LN keeps the whole 128 x D row tile (D = 192 or 384 <= 512 TMEM columns) in the CTA, so the LayerNorm of the
// freshly produced residual stream is computed in the same kernel: pass 1 reads the accumulator and the residual tile
// (TMA-loaded into shared memory), writes x' in place, accumulates sum / sum of squares, releases TMEM (the next tile's
// MMAs start here), TMA-stores x'; pass 2 normalises in place and TMA-stores h.  The separate add+LayerNorm kernel
// (8 bytes of traffic per element) and the GEMM's own output round trip disappear.
#include <stdlib.h>
#include "d2s_tc.cuh"

namespace d2s {

constexpr int kGpBM = 128, kGpBK = 64, kGpThreads = 320;
constexpr uint32_t kGpABytes = kGpBM * 128;      // 128 rows x 64 bf16
constexpr uint32_t kGpBlkBytes = 128 * 128;      // one staged 128 x 64 bf16 block (SWIZZLE_128B)

enum { kGpModeAct = 0, kGpModeLn = 1 };

template <int MODE, int NSUB> struct GpCfg;
template <> struct GpCfg<kGpModeAct, 1> {        // fc1: 256-column tiles, accumulator double-buffered
  static constexpr int UN = 256, NSUB = 1, ACC = 2, STAGES = 4, BLOCKS = 4;      // BLOCKS: staging blocks of 16 KB
};
template <> struct GpCfg<kGpModeLn, 2> {         // D = 384: two N = 192 MMAs per k-step, one accumulator
  static constexpr int UN = 192, NSUB = 2, ACC = 1, STAGES = 3, BLOCKS = 6;
};
template <> struct GpCfg<kGpModeLn, 1> {         // D = 192
  static constexpr int UN = 192, NSUB = 1, ACC = 1, STAGES = 4, BLOCKS = 3;
};

struct GpBars {
  uint64_t full[4], empty[4], tmem_full[2], tmem_empty[2], xfull;
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are counted on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
template <bool kAccum>
__device__ __forceinline__ void mma2_ss_imm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if (kAccum)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc));
}
// arrive (count 1) on the barrier at the same shared-memory offset in both CTAs of the pair once all prior MMAs retire
__device__ __forceinline__ void mma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

struct GpParams {
  const __nv_bfloat16* bias;    // (N) or NULL
  const __nv_bfloat16* gamma;   // MODE_LN: (N) or NULL when no LayerNorm output is wanted
  const __nv_bfloat16* beta;
  float eps;
  int M, N, K, act, want_ln;
  long long* trace;             // D2S_GEMM_TRACE: device buffer for per-warp clock64 phase totals (profiling only)
  int dbg;                      // profiling switches (D2S_GEMM_DEBUG): 1 no output stores, 2 no epilogue body, 8 no operand loads
};

template <int MODE, int NSUB_, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGpThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_o,    // MODE_ACT: out; MODE_LN: x' (sum) output
                 const __grid_constant__ CUtensorMap map_x,    // MODE_LN: residual input
                 const __grid_constant__ CUtensorMap map_h,    // MODE_LN: LayerNorm output
                 const GpParams p) {
  using Cfg = GpCfg<MODE, NSUB_>;
  constexpr int UN = Cfg::UN, NSUB = Cfg::NSUB, ACC = Cfg::ACC, STAGES = Cfg::STAGES;
  constexpr int TN = UN * NSUB;                                 // output columns per tile
  constexpr uint32_t kBSub = (UN / 2) * 128;                    // this CTA's half of one W sub-tile
  constexpr uint32_t kStage = kGpABytes + NSUB * kBSub;
  constexpr uint32_t kAccCols = (TN <= 256) ? 256 : 512;        // TMEM columns per accumulator stage

  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t padb = (1024u - (raw & 1023u)) & 1023u;
  unsigned char* ring = smem_dyn + padb;
  unsigned char* blocks = ring + STAGES * kStage;               // MODE_ACT: store staging; MODE_LN: x / x' / h tile
  GpBars* bars = reinterpret_cast<GpBars*>(blocks + Cfg::BLOCKS * kGpBlkBytes);
  float* bias_s = reinterpret_cast<float*>(bars + 1);           // N floats (MODE_ACT) / 3 x TN floats (MODE_LN)
  float2* red_s = reinterpret_cast<float2*>(bias_s + (MODE == kGpModeLn ? 3 * TN : p.N));   // MODE_LN: [2][128] partial stats

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int pair_tiles = (p.M + 2 * kGpBM - 1) / (2 * kGpBM);
  const int n_tiles = p.N / TN, k_blocks = p.K / kGpBK;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tmem_full[i]), 1); mbar_init(smem_u32(&bars->tmem_empty[i]), 16); }
    mbar_init(smem_u32(&bars->xfull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (MODE == kGpModeLn) {
    for (int i = tid; i < TN; i += kGpThreads) {
      bias_s[i] = p.bias ? __bfloat162float(p.bias[i]) : 0.f;
      bias_s[TN + i] = p.gamma ? __bfloat162float(p.gamma[i]) : 1.f;
      bias_s[2 * TN + i] = p.beta ? __bfloat162float(p.beta[i]) : 0.f;
    }
  } else {
    for (int i = tid; i < p.N; i += kGpThreads) bias_s[i] = p.bias ? __bfloat162float(p.bias[i]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      // ======================================= TMA producer (both CTAs) =======================================
      uint32_t it = 0;
      for (int pt = pair; pt < pair_tiles; pt += num_pairs) {
        const int row0 = pt * 2 * kGpBM + (int)rank * kGpBM;
        for (int nt = 0; nt < n_tiles; ++nt)
          for (int kb = 0; kb < k_blocks; ++kb, ++it) {
            const uint32_t s = it % STAGES, n = it / STAGES;
            mbar_wait(smem_u32(&bars->empty[s]), (n & 1) ^ 1);
            const uint32_t full_local = smem_u32(&bars->full[s]);
            if (p.dbg & 8) { if (rank == 0) mbar_expect_tx(full_local, 0); continue; }
            if (rank == 0) mbar_expect_tx(full_local, 2 * kStage);       // both CTAs' boxes land on the leader's barrier
            const uint32_t full_leader = mapa(full_local, 0);
            const uint32_t dst = smem_u32(ring + s * kStage);
            tma_load_2d_pair(dst, &map_a, kb * kGpBK, row0, full_leader);
#pragma unroll
            for (int j = 0; j < NSUB; ++j)
              tma_load_2d_pair(dst + kGpABytes + j * kBSub, &map_w, kb * kGpBK, nt * TN + j * UN + (int)rank * (UN / 2), full_leader);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ======================================== MMA issuer (leader) ========================================
      const uint32_t idesc = make_idesc(2 * kGpBM, UN, 0);
      uint32_t it = 0, tile = 0;
      for (int pt = pair; pt < pair_tiles; pt += num_pairs)
        for (int nt = 0; nt < n_tiles; ++nt, ++tile) {
          const uint32_t as = tile % ACC, an = tile / ACC;
          mbar_wait(smem_u32(&bars->tmem_empty[as]), (an & 1) ^ 1);      // both epilogues have drained this accumulator
          tc_fence_after();
          const uint32_t d = tmem + as * kAccCols;
          for (int kb = 0; kb < k_blocks; ++kb, ++it) {
            const uint32_t s = it % STAGES, n = it / STAGES;
            mbar_wait(smem_u32(&bars->full[s]), n & 1);
            tc_fence_after();
            const uint64_t ad = make_desc_sw128(smem_u32(ring + s * kStage), 16, 1024);
#pragma unroll
            for (int j = 0; j < NSUB; ++j) {
              const uint64_t bd = make_desc_sw128(smem_u32(ring + s * kStage + kGpABytes + j * kBSub), 16, 1024);
              if (kb == 0) mma2_ss_imm<false>(d + j * UN, ad, bd, idesc); else mma2_ss_imm<true>(d + j * UN, ad, bd, idesc);
              mma2_ss_imm<true>(d + j * UN, ad + 2, bd + 2, idesc);
              mma2_ss_imm<true>(d + j * UN, ad + 4, bd + 4, idesc);
              mma2_ss_imm<true>(d + j * UN, ad + 6, bd + 6, idesc);
            }
            mma2_commit_both(smem_u32(&bars->empty[s]));                 // ring slot free in both CTAs once these retire
          }
          mma2_commit_both(smem_u32(&bars->tmem_full[as]));
        }
    }
  } else {
    // ========================================= epilogue (both CTAs) =========================================
    const int ew = warp - 2;                    // 0..7
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may touch (warp id % 4)
    const int hc = ew >> 2;                     // which half of the tile's columns
    const int r = quad * 32 + lane;             // row inside this CTA's 128-row tile
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t tile = 0;

    if (MODE == kGpModeAct) {
      long long tr[5] = {0, 0, 0, 0, 0};
      const int gtid = tid - 64 - hc * 128;     // 0..127 inside this column-half group (4 warps)
      for (int pt = pair; pt < pair_tiles; pt += num_pairs) {
        const int row0 = pt * 2 * kGpBM + (int)rank * kGpBM;
        for (int nt = 0; nt < n_tiles; ++nt, ++tile) {
          const uint32_t as = tile % ACC, an = tile / ACC;
          long long t0 = clock64();
          mbar_wait(smem_u32(&bars->tmem_full[as]), an & 1);
          tc_fence_after();
          long long t1 = clock64();
          tr[0] += t1 - t0;
          const int col0 = nt * TN + hc * (TN / 2);
          const uint32_t taddr = lane_addr + as * kAccCols + hc * (TN / 2);
#pragma unroll
          for (int sb = 0; sb < TN / 128; ++sb) {     // 64-column store blocks of this half
            unsigned char* blk = blocks + (size_t)(hc * 2 + (sb & 1)) * kGpBlkBytes;
            // the TMA store that last read this block must be done before it is overwritten
            t0 = clock64();
            if (gtid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + hc) : "memory");
            t1 = clock64();
            tr[1] += t1 - t0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (p.dbg & 2) break;
              uint32_t v[32];
              t0 = clock64();
              tmem_ld32_nowait(taddr + sb * 64 + c * 32, v);
              tmem_ld_wait();
              t1 = clock64();
              tr[2] += t1 - t0;
              uint32_t o[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const float2 bq = *reinterpret_cast<const float2*>(&bias_s[col0 + sb * 64 + c * 32 + 2 * q]);
                uint64_t xp = f2_add(f2_pack(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), f2_pack(bq.x, bq.y));
                float x0, x1;
                if (ACT == D2S_ACT_GELU) xp = gelu_erf_pair(xp);
                f2_unpack(xp, x0, x1);
                if (ACT == D2S_ACT_RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                o[q] = pack_bf16x2(x0, x1);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(blk + sw128_off(r, c * 4 + q)) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
              t0 = clock64();
              tr[3] += t0 - t1;
            }
            t0 = clock64();
            if (sb == TN / 128 - 1) {   // all TMEM reads of this accumulator are done: hand it back before the stores drain
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&bars->tmem_empty[as]), 0));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + hc) : "memory");
            if (gtid == 0 && !(p.dbg & 1)) {
              tma_store_2d(&map_o, smem_u32(blk), col0 + sb * 64, row0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            t1 = clock64();
            tr[4] += t1 - t0;
          }
        }
      }
      if (p.trace && lane == 0) {
        long long* dst = p.trace + ((size_t)blockIdx.x * 8 + ew) * 8;
        for (int i = 0; i < 5; ++i) dst[i] = tr[i];
        dst[5] = tile;
      }
    } else {
      // ---- MODE_LN: n_tiles == 1, TN == N ----
      constexpr int NCH = TN / 32;              // 32-column chunks per row; this thread owns chunks [hc*NCH/2, (hc+1)*NCH/2)
      const float* gamma_s = bias_s + TN;
      const float* beta_s = bias_s + 2 * TN;
      const uint32_t xfull = smem_u32(&bars->xfull);
      if (tid == 64 && pair < pair_tiles) {     // residual tile of the first row tile
        mbar_expect_tx(xfull, Cfg::BLOCKS * kGpBlkBytes);
        for (int b = 0; b < Cfg::BLOCKS; ++b)
          tma_load_2d(smem_u32(blocks + b * kGpBlkBytes), &map_x, b * 64, pair * 2 * kGpBM + (int)rank * kGpBM, xfull);
      }
      for (int pt = pair; pt < pair_tiles; pt += num_pairs, ++tile) {
        const int row0 = pt * 2 * kGpBM + (int)rank * kGpBM;
        mbar_wait(smem_u32(&bars->tmem_full[0]), tile & 1);
        tc_fence_after();
        mbar_wait(xfull, tile & 1);
        // ---- pass 1: x' = bf16(x + bf16(acc + bias)) in place, fp32 sum / sum of squares of the rounded values ----
        uint64_t acc_s = f2_bcast(0.f), acc_q = f2_bcast(0.f);
#pragma unroll 2
        for (int ci = 0; ci < NCH / 2; ++ci) {
          const int ch = hc * (NCH / 2) + ci;
          uint32_t v[32];
          tmem_ld32_nowait(lane_addr + ch * 32, v);
          unsigned char* blk = blocks + (size_t)(ch >> 1) * kGpBlkBytes;
          uint4 xr[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) xr[q] = *reinterpret_cast<const uint4*>(blk + sw128_off(r, (ch & 1) * 4 + q));
          tmem_ld_wait();
          const uint32_t* xw = reinterpret_cast<const uint32_t*>(xr);
          uint32_t o[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float2 bq = *reinterpret_cast<const float2*>(&bias_s[ch * 32 + 2 * q]);
            float y0, y1;
            f2_unpack(f2_add(f2_pack(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), f2_pack(bq.x, bq.y)), y0, y1);
            const uint32_t yb = pack_bf16x2(y0, y1);                    // the Linear's bf16 output
            float s0, s1;
            f2_unpack(f2_add(f2_pack(bf16_lo(xw[q]), bf16_hi(xw[q])), f2_pack(bf16_lo(yb), bf16_hi(yb))), s0, s1);
            o[q] = pack_bf16x2(s0, s1);                                 // the residual add's bf16 output
            const uint64_t sv = f2_pack(bf16_lo(o[q]), bf16_hi(o[q]));
            acc_s = f2_add(acc_s, sv);
            acc_q = f2_fma(sv, sv, acc_q);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(blk + sw128_off(r, (ch & 1) * 4 + q)) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
        // accumulator drained: the next tile's MMAs may start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&bars->tmem_empty[0]), 0));
        {
          float s0, s1, q0, q1;
          f2_unpack(acc_s, s0, s1);
          f2_unpack(acc_q, q0, q1);
          red_s[hc * 128 + r] = make_float2(s0 + s1, q0 + q1);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 64) {
          for (int b = 0; b < Cfg::BLOCKS; ++b) tma_store_2d(&map_o, smem_u32(blocks + b * kGpBlkBytes), b * 64, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // x' has left shared memory
        }
        if (p.want_ln) {
          const float2 ra = red_s[r], rb = red_s[128 + r];
          const float mean = (ra.x + rb.x) * (1.0f / TN);
          const float var = fmaxf((ra.y + rb.y) * (1.0f / TN) - mean * mean, 0.f);
          const float rstd = rsqrtf(var + p.eps);
          const uint64_t sc = f2_bcast(rstd), sh = f2_bcast(-mean * rstd);
          asm volatile("bar.sync 1, 256;" ::: "memory");                   // x' stores have read the tile; red_s consumed
          // ---- pass 2: h = (x' - mean) * rstd * gamma + beta in place ----
#pragma unroll 2
          for (int ci = 0; ci < NCH / 2; ++ci) {
            const int ch = hc * (NCH / 2) + ci;
            unsigned char* blk = blocks + (size_t)(ch >> 1) * kGpBlkBytes;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4* ptr = reinterpret_cast<uint4*>(blk + sw128_off(r, (ch & 1) * 4 + q));
              uint4 xv = *ptr;
              uint32_t* w = reinterpret_cast<uint32_t*>(&xv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = ch * 32 + q * 8 + e * 2;
                const float2 g = *reinterpret_cast<const float2*>(&gamma_s[col]);
                const float2 bt = *reinterpret_cast<const float2*>(&beta_s[col]);
                float h0, h1;
                f2_unpack(f2_fma(f2_fma(f2_pack(bf16_lo(w[e]), bf16_hi(w[e])), sc, sh), f2_pack(g.x, g.y), f2_pack(bt.x, bt.y)), h0, h1);
                w[e] = pack_bf16x2(h0, h1);
              }
              *ptr = xv;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (tid == 64) {
            for (int b = 0; b < Cfg::BLOCKS; ++b) tma_store_2d(&map_h, smem_u32(blocks + b * kGpBlkBytes), b * 64, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
        } else {
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // the tile buffer is free: fetch the next row tile's residual (same thread that waited for the stores)
        if (tid == 64 && pt + num_pairs < pair_tiles) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(xfull, Cfg::BLOCKS * kGpBlkBytes);
          for (int b = 0; b < Cfg::BLOCKS; ++b)
            tma_load_2d(smem_u32(blocks + b * kGpBlkBytes), &map_x, b * 64, (pt + num_pairs) * 2 * kGpBM + (int)rank * kGpBM, xfull);
        }
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // outstanding TMA stores read shared memory
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the peer may still signal barriers / read operands in this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

static int gp_debug() {
  const char* e = getenv("D2S_GEMM_DEBUG");
  return e ? atoi(e) : 0;
}

static long long* gp_trace() {
  const char* e = getenv("D2S_GEMM_TRACE");
  return e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 10)) : nullptr;
}

static int gp_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                     CUtensorMapL2promotion promo, const char* what) {
  GgEncodeFn enc = gg_encode_fn();
  D2S_REQUIRE(enc != nullptr, D2S_ERR_CUDA, "%s: cuTensorMapEncodeTiled is unavailable from the driver", what);
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstr[1] = {inner * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  D2S_REQUIRE(cr == CUDA_SUCCESS, D2S_ERR_CUDA, "%s: tensor map encode failed (%d)", what, (int)cr);
  return D2S_OK;
}

template <int MODE, int NSUB, int ACT>
static int gp_launch(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo, const CUtensorMap& mx, const CUtensorMap& mh,
                     const GpParams& p, cudaStream_t stream, const char* what) {
  using Cfg = GpCfg<MODE, NSUB>;
  constexpr uint32_t kStage = kGpABytes + NSUB * (Cfg::UN / 2) * 128;
  const size_t smem = 1024 + (size_t)Cfg::STAGES * kStage + (size_t)Cfg::BLOCKS * kGpBlkBytes + sizeof(GpBars) +
                      (MODE == kGpModeLn ? (size_t)3 * Cfg::UN * NSUB * 4 + 2 * 128 * sizeof(float2) : (size_t)p.N * 4);
  D2S_REQUIRE(smem <= 227 * 1024, D2S_ERR_ARG, "%s: needs %zu B of shared memory", what, smem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel<MODE, NSUB, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    attr_set = true;
  }
  const int pair_tiles = (p.M + 2 * kGpBM - 1) / (2 * kGpBM);
  const int pairs = pair_tiles < kNumSMs / 2 ? pair_tiles : kNumSMs / 2;
  gemm_pair_kernel<MODE, NSUB, ACT><<<2 * pairs, kGpThreads, smem, stream>>>(ma, mw, mo, mx, mh, p);
  count_launch();
  return check_launch(what);
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_linear_act_pair_bf16(const void* a, const void* w, const void* bias, int M, int N, int K, int act, void* out,
                                        d2s_stream_t stream) {
  const char* what = "d2s_linear_act_pair_bf16";
  D2S_REQUIRE(a && w && out, D2S_ERR_ARG, "linear_act_pair: null pointer");
  D2S_REQUIRE(M >= 0 && N >= 256 && N % 256 == 0 && N <= 4096 && K >= kGpBK && K % kGpBK == 0, D2S_ERR_ARG,
              "linear_act_pair: need N %% 256 == 0 (N <= 4096) and K %% %d == 0 (got M=%d N=%d K=%d)", kGpBK, M, N, K);
  D2S_REQUIRE(act >= D2S_ACT_NONE && act <= D2S_ACT_RELU, D2S_ERR_ARG, "linear_act_pair: bad activation %d", act);
  D2S_REQUIRE(aligned16(a) && aligned16(w) && aligned16(out), D2S_ERR_ALIGN, "linear_act_pair: pointers must be 16-byte aligned");
  if (M == 0) return D2S_OK;
  CUtensorMap ma, mw, mo;
  int rc;
  if ((rc = gp_map_2d(&ma, a, K, M, kGpBK, kGpBM, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what))) return rc;
  if ((rc = gp_map_2d(&mw, w, K, N, kGpBK, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, what))) return rc;
  if ((rc = gp_map_2d(&mo, out, N, M, 64, kGpBM, CU_TENSOR_MAP_L2_PROMOTION_NONE, what))) return rc;
  GpParams p{(const __nv_bfloat16*)bias, nullptr, nullptr, 0.f, M, N, K, act, 0, gp_trace(), gp_debug()};
  if (act == D2S_ACT_GELU) return gp_launch<kGpModeAct, 1, D2S_ACT_GELU>(ma, mw, mo, mo, mo, p, (cudaStream_t)stream, what);
  if (act == D2S_ACT_RELU) return gp_launch<kGpModeAct, 1, D2S_ACT_RELU>(ma, mw, mo, mo, mo, p, (cudaStream_t)stream, what);
  return gp_launch<kGpModeAct, 1, D2S_ACT_NONE>(ma, mw, mo, mo, mo, p, (cudaStream_t)stream, what);
}

extern "C" int d2s_linear_residual_ln_bf16(const void* a, const void* w, const void* bias, const void* x, const void* gamma,
                                           const void* beta, float eps, int M, int N, int K, void* out_sum, void* out_norm,
                                           d2s_stream_t stream) {
  const char* what = "d2s_linear_residual_ln_bf16";
  D2S_REQUIRE(a && w && x && out_sum, D2S_ERR_ARG, "linear_residual_ln: null pointer");
  D2S_REQUIRE(M >= 0 && (N == 192 || N == 384) && K >= kGpBK && K % kGpBK == 0, D2S_ERR_ARG,
              "linear_residual_ln: need N in {192, 384} (one CTA holds whole rows in TMEM) and K %% %d == 0 (got M=%d N=%d K=%d)",
              kGpBK, M, N, K);
  D2S_REQUIRE(!out_norm || (gamma && beta), D2S_ERR_ARG, "linear_residual_ln: out_norm needs gamma and beta");
  D2S_REQUIRE(aligned16(a) && aligned16(w) && aligned16(x) && aligned16(out_sum) && aligned16(out_norm), D2S_ERR_ALIGN,
              "linear_residual_ln: pointers must be 16-byte aligned");
  if (M == 0) return D2S_OK;
  CUtensorMap ma, mw, mo, mx, mh;
  int rc;
  if ((rc = gp_map_2d(&ma, a, K, M, kGpBK, kGpBM, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what))) return rc;
  if ((rc = gp_map_2d(&mw, w, K, N, kGpBK, 96, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, what))) return rc;
  if ((rc = gp_map_2d(&mx, x, N, M, 64, kGpBM, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, what))) return rc;
  if ((rc = gp_map_2d(&mo, out_sum, N, M, 64, kGpBM, CU_TENSOR_MAP_L2_PROMOTION_NONE, what))) return rc;
  if ((rc = gp_map_2d(&mh, out_norm ? out_norm : out_sum, N, M, 64, kGpBM, CU_TENSOR_MAP_L2_PROMOTION_NONE, what))) return rc;
  GpParams p{(const __nv_bfloat16*)bias, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, eps, M, N, K, 0, out_norm ? 1 : 0, gp_trace(), gp_debug()};
  if (N == 384) return gp_launch<kGpModeLn, 2, 0>(ma, mw, mo, mx, mh, p, (cudaStream_t)stream, what);
  return gp_launch<kGpModeLn, 1, 0>(ma, mw, mo, mx, mh, p, (cudaStream_t)stream, what);
}
