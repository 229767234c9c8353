// fc1 of the MLP with the activation fused into the GEMM epilogue (Mlp.forward, vit_models/dynamic_vit.py:159-175):
//     out (M,N) = GELU(A (M,K) @ W (N,K)^T + bias (N)),  bf16 in / fp32 accumulate / bf16 out.
// In the inference step the separate GELU pass was the largest non-GEMM item left (16 % of the step, 1.2 GB of traffic
// per T=197 layer); fusing it removes one write and one read of the (B,T,4D) hidden tensor.
//
// Persistent warp-specialised tcgen05 GEMM, one CTA per SM:
//   warp 0      TMA producer: A tile 128x64 and W tile 256x64 (bf16, SWIZZLE_128B) into a 3-stage ring
//   warp 1      MMA issuer: tcgen05.mma.kind::f16 M=128 N=256 K=16, accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2-9   epilogue: two warps per TMEM lane quadrant (128 output columns each): tcgen05.ld -> +bias -> GELU (erf
//               form: erfcx polynomial x one MUFU.EX2, packed fp32x2 arithmetic) -> bf16 -> swizzled shared-memory block -> TMA store (one
//               instruction per 128x64 block; per-thread global stores of 16 bytes cost 32 LSU wavefronts per warp
//               instruction and made the first version of this kernel store-bound)
// A CTA walks its 128-row tiles with the N tiles innermost, so the A tile is re-read from L2, never from HBM.
// The epilogue is the binding stage (GELU against 384 MACs per element on the tensor pipe), which is why it gets eight
// warps, packed fp32x2 arithmetic and runs concurrently with the next tile's MMAs.
#include "d2s_tc.cuh"

namespace d2s {

constexpr int kGgBM = 128, kGgBN = 256, kGgBK = 64, kGgStages = 3;
constexpr uint32_t kGgStoreBytes = 128 * 128;   // one staged store block: 128 rows x 64 bf16 columns (SWIZZLE_128B)
constexpr int kGgThreads = 320;
constexpr uint32_t kGgABytes = kGgBM * 128, kGgBBytes = kGgBN * 128, kGgStageBytes = kGgABytes + kGgBBytes;

struct GgBars {
  uint64_t full[kGgStages], empty[kGgStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __launch_bounds__(kGgThreads, 1)
gemm_bias_gelu_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_o, const __nv_bfloat16* __restrict__ bias, int M, int N, int K,
                      int act) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  unsigned char* ring = smem_dyn + pad;
  unsigned char* stage_o = ring + kGgStages * kGgStageBytes;   // 2 column halves x 2 buffers x 16 KB
  GgBars* bars = reinterpret_cast<GgBars*>(stage_o + 4 * kGgStoreBytes);
  float* bias_s = reinterpret_cast<float*>(bars + 1);          // N floats

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_tiles = (M + kGgBM - 1) / kGgBM, n_tiles = N / kGgBN, k_blocks = K / kGgBK;

  if (tid == 0) {
    for (int i = 0; i < kGgStages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tmem_full[i]), 1); mbar_init(smem_u32(&bars->tmem_empty[i]), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < N; i += kGgThreads) bias_s[i] = bias ? __bfloat162float(bias[i]) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      // ======================================= TMA producer =======================================
      uint32_t it = 0;
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x)
        for (int nt = 0; nt < n_tiles; ++nt)
          for (int kb = 0; kb < k_blocks; ++kb, ++it) {
            const uint32_t s = it % kGgStages, n = it / kGgStages;
            mbar_wait(smem_u32(&bars->empty[s]), (n & 1) ^ 1);
            const uint32_t bar = smem_u32(&bars->full[s]);
            mbar_expect_tx(bar, kGgStageBytes);
            tma_load_2d(smem_u32(ring + s * kGgStageBytes), &map_a, kb * kGgBK, mt * kGgBM, bar);
            tma_load_2d(smem_u32(ring + s * kGgStageBytes + kGgABytes), &map_w, kb * kGgBK, nt * kGgBN, bar);
          }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ======================================== MMA issuer ========================================
      const uint32_t idesc = make_idesc(kGgBM, kGgBN, 0);
      uint32_t it = 0, tile = 0;
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x)
        for (int nt = 0; nt < n_tiles; ++nt, ++tile) {
          const uint32_t as = tile & 1, an = tile >> 1;
          mbar_wait(smem_u32(&bars->tmem_empty[as]), (an & 1) ^ 1);   // the epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d = tmem + as * kGgBN;
          for (int kb = 0; kb < k_blocks; ++kb, ++it) {
            const uint32_t s = it % kGgStages, n = it / kGgStages;
            mbar_wait(smem_u32(&bars->full[s]), n & 1);
            tc_fence_after();
            const uint64_t ad = make_desc_sw128(smem_u32(ring + s * kGgStageBytes), 16, 1024);
            const uint64_t bd = make_desc_sw128(smem_u32(ring + s * kGgStageBytes + kGgABytes), 16, 1024);
            if (kb == 0) mma_ss_imm<false>(d, ad, bd, idesc); else mma_ss_imm<true>(d, ad, bd, idesc);
            mma_ss_imm<true>(d, ad + 2, bd + 2, idesc);
            mma_ss_imm<true>(d, ad + 4, bd + 4, idesc);
            mma_ss_imm<true>(d, ad + 6, bd + 6, idesc);
            mma_commit(smem_u32(&bars->empty[s]));                     // smem slot free once these MMAs retire
          }
          mma_commit(smem_u32(&bars->tmem_full[as]));
        }
    }
  } else {
    // ========================================= epilogue =========================================
    const int ew = warp - 2;                    // 0..7
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may touch (warp id % 4)
    const int hcol = ew >> 2;                   // which 128-column half of the tile
    const int r = quad * 32 + lane;             // row inside the tile
    uint32_t tile = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x)
      for (int nt = 0; nt < n_tiles; ++nt, ++tile) {
        const uint32_t as = tile & 1, an = tile >> 1;
        mbar_wait(smem_u32(&bars->tmem_full[as]), an & 1);
        tc_fence_after();
        const int col0 = nt * kGgBN + hcol * 128;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + as * kGgBN + hcol * 128;
        const int gtid = tid - 64 - hcol * 128;       // 0..127 inside this column-half group (4 warps)
#pragma unroll
        for (int sb = 0; sb < 2; ++sb) {              // two 64-column store blocks per half
          unsigned char* blk = stage_o + (size_t)(hcol * 2 + sb) * kGgStoreBytes;
          // the TMA store that last read this block (previous tile) must be done before it is overwritten
          if (gtid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + hcol) : "memory");
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32_nowait(taddr + sb * 64 + c * 32, v);
            tmem_ld_wait();
            uint32_t o[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float2 bq = *reinterpret_cast<const float2*>(&bias_s[col0 + sb * 64 + c * 32 + 2 * q]);
              uint64_t xp = f2_add(f2_pack(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), f2_pack(bq.x, bq.y));
              float x0, x1;
              if (act == D2S_ACT_GELU) xp = gelu_erf_pair(xp);
              f2_unpack(xp, x0, x1);
              if (act == D2S_ACT_RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
              o[q] = pack_bf16x2(x0, x1);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)   // 16-byte chunk (c*4 + q) of this row's 128 bytes, SWIZZLE_128B position
              *reinterpret_cast<uint4*>(blk + sw128_off(r, c * 4 + q)) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
          if (sb == 1) {   // all TMEM reads of this accumulator are done: hand it back before the stores drain
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->tmem_empty[as]));
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA unit
          asm volatile("bar.sync %0, 128;" ::"r"(1 + hcol) : "memory");
          if (gtid == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&map_o)), "r"(smem_u32(blk)), "r"(col0 + sb * 64), "r"(mt * kGgBM) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }

      }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // outstanding TMA stores read shared memory
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_linear_act_bf16(const void* a, const void* w, const void* bias, int M, int N, int K, int act, void* out,
                                   d2s_stream_t stream) {
  D2S_REQUIRE(a && w && out, D2S_ERR_ARG, "linear_act: null pointer");
  D2S_REQUIRE(M >= 0 && N >= kGgBN && N % kGgBN == 0 && N <= 4096 && K >= kGgBK && K % kGgBK == 0, D2S_ERR_ARG,
              "linear_act: need N %% %d == 0 (N <= 4096) and K %% %d == 0 (got M=%d N=%d K=%d)", kGgBN, kGgBK, M, N, K);
  D2S_REQUIRE(act >= D2S_ACT_NONE && act <= D2S_ACT_RELU, D2S_ERR_ARG, "linear_act: bad activation %d", act);
  D2S_REQUIRE(aligned16(a) && aligned16(w) && aligned16(out), D2S_ERR_ALIGN, "linear_act: pointers must be 16-byte aligned");
  if (M == 0) return D2S_OK;
  GgEncodeFn enc = gg_encode_fn();
  D2S_REQUIRE(enc != nullptr, D2S_ERR_CUDA, "linear_act: cuTensorMapEncodeTiled is unavailable from the driver");
  CUtensorMap map_a, map_w;
  const cuuint32_t estr[2] = {1, 1};
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {kGgBK, kGgBM};
    CUresult cr = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    D2S_REQUIRE(cr == CUDA_SUCCESS, D2S_ERR_CUDA, "linear_act: tensor map (A) failed (%d)", (int)cr);
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    const cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {kGgBK, kGgBN};
    CUresult cr = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    D2S_REQUIRE(cr == CUDA_SUCCESS, D2S_ERR_CUDA, "linear_act: tensor map (W) failed (%d)", (int)cr);
  }
  CUtensorMap map_o;
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)N * 2};
    const cuuint32_t box[2] = {64, kGgBM};
    CUresult cr = enc(&map_o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    D2S_REQUIRE(cr == CUDA_SUCCESS, D2S_ERR_CUDA, "linear_act: tensor map (out) failed (%d)", (int)cr);
  }
  const size_t smem = 1024 + (size_t)kGgStages * kGgStageBytes + 4 * (size_t)kGgStoreBytes + sizeof(GgBars) + (size_t)N * sizeof(float);
  D2S_REQUIRE(smem <= 227 * 1024, D2S_ERR_ARG, "linear_act: N=%d needs %zu B of shared memory", N, smem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bias_gelu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "linear_act: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int m_tiles = (M + kGgBM - 1) / kGgBM;
  const int grid = m_tiles < kNumSMs ? m_tiles : kNumSMs;
  gemm_bias_gelu_kernel<<<grid, kGgThreads, smem, (cudaStream_t)stream>>>(map_a, map_w, map_o, (const __nv_bfloat16*)bias, M, N, K, act);
  count_launch();
  return check_launch("d2s_linear_act_bf16");
}
