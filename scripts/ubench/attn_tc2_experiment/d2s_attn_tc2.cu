// Kernel family (4), tensor-core part, 128 < T <= 256 (two 128-row query tiles per (image, head) unit).
//
// Same math as d2s_attn_tc.cu (S = Q K^T -> policy softmax in TMEM -> O = P V -> normalise), different schedule:
// ONE persistent CTA per SM owns the whole TMEM (512 columns) and most of the shared memory, and runs
//   warps 0-3  softmax warpgroup 0  (TMEM columns [0,256):   S/P at +0, O at +128)
//   warps 4-7  softmax warpgroup 1  (TMEM columns [256,512): S/P at +0, O at +128)
//   warp  8    TMA producer (lane 0)
//   warps 9,10 MMA issuers  (lane 0), one per warpgroup
// Both query tiles of a unit are in flight at once, one per warpgroup (the roles swap every unit so the short tail
// tile alternates), and they share the unit's K and V.  Q/K die right after the two S-MMAs (start of the unit) and V
// after the two PV-MMAs (end of the unit), so the shared memory is a 2-deep ring of {Q0,Q1,K} stages plus a 3-deep
// ring of V buffers: the TMA producer runs one to two units ahead of the tensor cores and the loads never sit on the
// critical path (the first version of this kernel, 2 CTAs/SM with single buffers, was bound by exactly that: its
// load/MMA skeleton and its softmax time added up instead of overlapping; profiles/README.md, r01b).
#include "d2s_tc.cuh"

namespace d2s {

constexpr int kT2Threads = 352;
constexpr int kT2OCol = 128;

struct Tc2Bars {
  uint64_t qk_full[2], qk_empty[2], v_full[3], v_empty[3];
  uint64_t s_full[2], p_full[2], o_full[2], tmem_free[2];
  uint32_t tmem_base;
  uint32_t pad;
};

template <bool kPol>
__global__ void __launch_bounds__(kT2Threads, 1)
attn_tc2_fwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const float* __restrict__ policy, int num_units, int T, int H, int Tkp, int vstages, float scale,
                    float eps, __nv_bfloat16* __restrict__ out, float* __restrict__ cls_row, long long* trace) {
#define TR(role, ev) do { if (trace && blockIdx.x == 0 && it < 16) trace[((role) * 16 + it) * 8 + (ev)] = clock64(); } while (0)
  extern __shared__ unsigned char smem_dyn[];
  const int rows_b = Tkp - kTileRows;                          // rows of the second TMA box (16..128)
  const uint32_t bytes_a = kTileBytes, bytes_b = (uint32_t)rows_b * 128u;
  const uint32_t kv_bytes = (uint32_t)Tkp * 128u;              // one K or V buffer
  const uint32_t qk_stage = kTileBytes + bytes_b + kv_bytes;   // Q0 (128 rows) | Q1 (rows_b rows) | K (Tkp rows)
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  unsigned char* tiles = smem_dyn + pad;                       // SWIZZLE_128B atoms are 1024 B and address based
  unsigned char* qk_s = tiles;                                 // 2 stages
  unsigned char* v_s = qk_s + 2 * (size_t)qk_stage;            // vstages buffers
  Tc2Bars* bars = reinterpret_cast<Tc2Bars*>(v_s + (size_t)vstages * kv_bytes);
  float* pol_s = reinterpret_cast<float*>(bars + 1);           // 2 x 256
  float* cls_s = pol_s + 512;                                  // 2 x 256
  float* vsum_s = cls_s + 512;                                 // 2 x 64

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->qk_full[i]), 1);
      mbar_init(smem_u32(&bars->qk_empty[i]), 2);
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->p_full[i]), 128);
      mbar_init(smem_u32(&bars->o_full[i]), 1);
      mbar_init(smem_u32(&bars->tmem_free[i]), 128);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(smem_u32(&bars->v_full[i]), 1);
      mbar_init(smem_u32(&bars->v_empty[i]), 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 8) {
    if (lane == 0) {
      // ======================================= TMA producer =======================================
      uint32_t it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int b = unit / H, h = unit % H;
        const uint32_t sq = it & 1, nq = it >> 1;
        unsigned char* st = qk_s + (size_t)sq * qk_stage;
        TR(0, 0);
        mbar_wait(smem_u32(&bars->qk_empty[sq]), (nq & 1) ^ 1);
        TR(0, 1);
        const uint32_t bar = smem_u32(&bars->qk_full[sq]);
        mbar_expect_tx(bar, 2 * (bytes_a + bytes_b));
        // K first (the S-MMA of either tile needs all of it), then the two Q tiles
        tma_load_3d(smem_u32(st) + kTileBytes + bytes_b, &map_a, (H + h) * kTcHD, 0, b, bar);
        tma_load_3d(smem_u32(st) + kTileBytes + bytes_b + bytes_a, &map_b, (H + h) * kTcHD, kTileRows, b, bar);
        tma_load_3d(smem_u32(st), &map_a, h * kTcHD, 0, b, bar);
        tma_load_3d(smem_u32(st) + kTileBytes, &map_b, h * kTcHD, kTileRows, b, bar);
        const uint32_t sv = it % (uint32_t)vstages, nv = it / (uint32_t)vstages;
        TR(0, 2);
        mbar_wait(smem_u32(&bars->v_empty[sv]), (nv & 1) ^ 1);
        TR(0, 3);
        const uint32_t vbar = smem_u32(&bars->v_full[sv]);
        mbar_expect_tx(vbar, bytes_a + bytes_b);
        tma_load_3d(smem_u32(v_s) + sv * kv_bytes, &map_a, (2 * H + h) * kTcHD, 0, b, vbar);
        tma_load_3d(smem_u32(v_s) + sv * kv_bytes + bytes_a, &map_b, (2 * H + h) * kTcHD, kTileRows, b, vbar);
      }
    }
  } else if (warp >= 9) {
    if (lane == 0) {
      // ============================ MMA issuer of warpgroup w = warp - 9 ============================
      // Each warpgroup is its own S -> softmax -> PV pipeline; the two only share the unit's smem stages.  Warpgroup 1
      // starts half a period late (after warpgroup 0's first softmax) so that one group's MUFU-bound exp pass runs
      // against the other group's MMA / load / store phases instead of against its exp pass.
      const uint32_t w = (uint32_t)(warp - 9);
      const uint32_t idesc_s = make_idesc(128, Tkp, 0);
      const uint32_t idesc_o = make_idesc(128, kTcHD, 1);
      const int ksteps = Tkp / 16;
      const uint32_t d = tmem + 256u * w;
      uint32_t it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const uint32_t sq = it & 1, nq = it >> 1;
        const uint32_t sv = it % (uint32_t)vstages, nv = it / (uint32_t)vstages;
        const uint32_t t = (w + it) & 1;                 // query tile of this warpgroup for this unit
        unsigned char* st = qk_s + (size_t)sq * qk_stage;
        const uint64_t kd = make_desc_sw128(smem_u32(st) + kTileBytes + bytes_b, 16, 1024);
        const uint64_t qd = make_desc_sw128(smem_u32(st) + t * kTileBytes, 16, 1024);
        if (w == 0) TR(1, 0);
        mbar_wait(smem_u32(&bars->qk_full[sq]), nq & 1);
        if (w == 0) TR(1, 1);
        if (it > 0) mbar_wait(smem_u32(&bars->tmem_free[w]), (it - 1) & 1);
        else if (w == 1) mbar_wait(smem_u32(&bars->p_full[0]), 0);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < kTcHD / 16; ++ks) mma_ss(d, qd + (uint64_t)(ks * 2), kd + (uint64_t)(ks * 2), idesc_s, ks > 0);
        mma_commit(smem_u32(&bars->s_full[w]));
        mma_commit(smem_u32(&bars->qk_empty[sq]));       // count 2: the stage dies when both groups' S-MMAs are done
        if (w == 0) TR(1, 2);
        mbar_wait(smem_u32(&bars->v_full[sv]), nv & 1);
        if (w == 0) TR(1, 3);
        const uint64_t vd = make_desc_sw128(smem_u32(v_s) + sv * kv_bytes, 16, 1024);
        mbar_wait(smem_u32(&bars->p_full[w]), it & 1);
        if (w == 0) TR(1, 4);
        tc_fence_after();
        for (int ks = 0; ks < ksteps; ++ks)
          mma_ts(d + kT2OCol, d + (uint32_t)(ks * 8), vd + (uint64_t)(ks * 128), idesc_o, ks > 0);
        mma_commit(smem_u32(&bars->o_full[w]));
        mma_commit(smem_u32(&bars->v_empty[sv]));        // count 2
        if (w == 0) TR(1, 6);
      }
    }
  } else {
    // ===================================== softmax / epilogue warpgroups =====================================
    const uint32_t w = warp >> 2;                       // warpgroup
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may touch
    const int r = quad * 32 + lane;                     // row inside the tile == TMEM lane
    const int wtid = tid & 127;
    const uint32_t lane_addr = tmem + 256u * w + ((uint32_t)(quad * 32) << 16);
    float* pol = pol_s + 256 * w;
    float* cls = cls_s + 256 * w;
    float* vsum = vsum_s + 64 * w;
    const float k2 = scale * 1.4426950408889634f;
    const float c_eps = kPol ? eps / (float)T : 0.0f;
    const float eps_den = kPol ? eps : 0.0f;
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const int b = unit / H, h = unit % H;
      const int t = (int)((w + it) & 1);                // query tile of this warpgroup for this unit
      if (kPol) {
        const uint32_t sv = it % (uint32_t)vstages, nv = it / (uint32_t)vstages;
        const unsigned char* vb = v_s + (size_t)sv * kv_bytes;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");  // previous unit's readers of pol / vsum are done
        for (int j = wtid; j < 256; j += 128) pol[j] = (j < T) ? policy[(size_t)b * T + j] : 0.0f;
        mbar_wait(smem_u32(&bars->v_full[sv]), nv & 1);
        if (wtid < kTcHD) {
          // column sums of V for the eps/T term: sum_j V[j][d]
          float acc = 0.f;
          const int cchunk = wtid >> 3, within = wtid & 7;
          for (int j = 0; j < T; ++j)
            acc += __bfloat162float(*(reinterpret_cast<const __nv_bfloat16*>(vb + sw128_off(j, cchunk)) + within));
          vsum[wtid] = acc;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
      }
      const int i = t * kTileRows + r;                              // query token of this thread
      const bool warp_active = t * kTileRows + quad * 32 < T;       // whole warp beyond T: nothing to compute
      if (wtid == 0) TR(2 + w, 0);
      mbar_wait(smem_u32(&bars->s_full[w]), it & 1);
      if (wtid == 0) TR(2 + w, 1);
      tc_fence_after();
      float sum = 0.f, eps_scale = 1.0f;
      const bool want_cls = (cls_row != nullptr) && (i == 0);
      if (warp_active) {
        // ONE pass over S (TMEM reads run at ~64 B/clk/SM and are the binding resource of this kernel: reading S
        // twice, once for the row max and once for the exponentials, costs more than the exponentials themselves).
        // The exponentials are taken against a per-row reference m' = max of the row's first 32 columns instead of
        // the row max: softmax is shift invariant and bf16 keeps fp32's exponent range, so P = 2^(s - m') is as
        // accurate as 2^(s - max); the true max is tracked on the side (ALU pipe) because the reference's eps terms
        // are not shift invariant -- they are rescaled by C = 2^(max - m') below, which restores the exact formula.
        // Exponents are clamped at 2^120 (a row whose max exceeds its first 32 columns by 83 nats).
        // The row is walked in 32-column chunks with the NEXT chunk's tcgen05.ld already in flight.
        const int n32 = (Tkp + 31) >> 5;
        uint32_t va[32], vb[32];
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        auto max_chunk = [&](const uint32_t (&v)[32], int c) {
          if (c * 32 + 32 <= T) {
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[q]));
              m1 = fmaxf(m1, __uint_as_float(v[q + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[q + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[q + 3]));
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (c * 32 + q < T) m0 = fmaxf(m0, __uint_as_float(v[q]));
          }
        };
        float mxk = 0.f;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        auto exp_chunk = [&](const uint32_t (&v)[32], int c) {
          const bool full = c * 32 + 32 <= T;
          uint32_t packed[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int j = c * 32 + 2 * q;
            float e0 = ex2_approx(fminf(fmaf(__uint_as_float(v[2 * q]), k2, -mxk), 120.0f));
            float e1 = ex2_approx(fminf(fmaf(__uint_as_float(v[2 * q + 1]), k2, -mxk), 120.0f));
            if (kPol) {
              e0 *= (j == i) ? 1.0f : pol[j];
              e1 *= (j + 1 == i) ? 1.0f : pol[j + 1];
            }
            if (!full) {
              if (j >= T) e0 = 0.f;
              if (j + 1 >= T) e1 = 0.f;
            }
            if (q & 1) { s2 += e0; s3 += e1; } else { s0 += e0; s1 += e1; }
            if (want_cls) { cls[j] = e0; cls[j + 1] = e1; }
            packed[q] = pack_bf16x2_alu(e0, e1);
          }
          tmem_st16(lane_addr + (uint32_t)(c * 16), packed);  // P overlays the S columns already consumed
        };
        tmem_ld32_nowait(lane_addr, va);
        for (int c = 0; c < n32; c += 2) {
          tmem_ld_wait();
          if (c + 1 < n32) tmem_ld32_nowait(lane_addr + (uint32_t)((c + 1) * 32), vb);
          if (c == 0) {
            max_chunk(va, 0);
            mxk = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * k2;
          } else if (kPol) {
            max_chunk(va, c);
          }
          exp_chunk(va, c);
          if (c + 1 < n32) {
            tmem_ld_wait();
            if (c + 2 < n32) tmem_ld32_nowait(lane_addr + (uint32_t)((c + 2) * 32), va);
            if (kPol) max_chunk(vb, c + 1);
            exp_chunk(vb, c + 1);
          }
        }
        if (kPol) eps_scale = ex2_approx(fminf(fmaf(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), k2, -mxk), 120.0f));
        if (wtid == 0) TR(2 + w, 2);
        sum = (s0 + s1) + (s2 + s3);
      }  // rows of an idle warp are never written out; whatever their P rows hold stays in those rows
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      if (wtid == 0) TR(2 + w, 3);
      mbar_arrive(smem_u32(&bars->p_full[w]));
      const float den = sum + eps_den * eps_scale;
      const float c_eps_row = c_eps * eps_scale;
      if (cls_row != nullptr && t == 0 && quad == 0) {
        // CLS row (query 0 = lane 0 of quadrant 0): probabilities of row 0, Attention.forward's second output
        __syncwarp();
        const float den0 = __shfl_sync(0xffffffffu, den, 0);
        const float ce0 = __shfl_sync(0xffffffffu, c_eps_row, 0);
        for (int j = lane; j < T; j += 32) cls_row[(size_t)unit * T + j] = (cls[j] + ce0) / den0;
        __syncwarp();
      }
      // ---- epilogue ---------------------------------------------------------------------------------------
      mbar_wait(smem_u32(&bars->o_full[w]), it & 1);
      if (wtid == 0) TR(2 + w, 4);
      tc_fence_after();
      uint32_t ov[2][32];
      if (warp_active) {
        tmem_ld32_nowait(lane_addr + (uint32_t)kT2OCol, ov[0]);
        tmem_ld32_nowait(lane_addr + (uint32_t)(kT2OCol + 32), ov[1]);
        tmem_ld_wait();
      }
      tc_fence_before();
      if (wtid == 0) TR(2 + w, 5);
      mbar_arrive(smem_u32(&bars->tmem_free[w]));     // O is in registers: the next S-MMA may overwrite this TMEM region
      if (warp_active && i < T) {
        const float inv = 1.0f / den;
        __nv_bfloat16* orow = out + ((size_t)b * T + i) * (size_t)(H * kTcHD) + (size_t)h * kTcHD;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float o0 = __uint_as_float(ov[half][2 * q]), o1 = __uint_as_float(ov[half][2 * q + 1]);
            if (kPol) {
              o0 += c_eps_row * vsum[half * 32 + 2 * q];
              o1 += c_eps_row * vsum[half * 32 + 2 * q + 1];
            }
            o[q] = pack_bf16x2(o0 * inv, o1 * inv);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<uint4*>(orow)[half * 4 + q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

static long long* g_trace = nullptr;
static size_t tc2_smem_bytes(int Tkp, int vstages) {
  const size_t rows_b = Tkp - kTileRows;
  const size_t qk_stage = kTileBytes + rows_b * 128 + (size_t)Tkp * 128;
  return 1024 + 2 * qk_stage + (size_t)vstages * Tkp * 128 + sizeof(Tc2Bars) + (512 + 512 + 128) * sizeof(float);
}

// Called by d2s_attn_policy_fwd (d2s_attn_tc.cu) for bf16, hd == 64, 128 < T <= 256.
int attn_tc2_launch(const CUtensorMap& map_a, const CUtensorMap& map_b, const float* policy, int units, int T, int H,
                    int Tkp, float scale, float eps, void* out, float* cls_row, cudaStream_t stream) {
  const size_t limit = 227 * 1024;
  const int vstages = tc2_smem_bytes(Tkp, 3) <= limit ? 3 : 2;
  const size_t smem = tc2_smem_bytes(Tkp, vstages);
  D2S_REQUIRE(smem <= limit, D2S_ERR_ARG, "attn_policy_fwd(tc2): T=%d needs %zu B of shared memory", T, smem);
  static bool attr_set[2] = {false, false};
  const int pi = policy ? 1 : 0;
  if (!attr_set[pi]) {
    cudaError_t e = policy ? cudaFuncSetAttribute(attn_tc2_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit)
                           : cudaFuncSetAttribute(attn_tc2_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "attn_policy_fwd(tc2): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[pi] = true;
  }
  const int grid = units < kNumSMs ? units : kNumSMs;
  static long long* trace = nullptr;
  if (getenv("D2S_ATTN_TRACE")) {
    if (!trace) { cudaMalloc(&trace, 4 * 16 * 8 * 8); }
    cudaMemsetAsync(trace, 0, 4 * 16 * 8 * 8, stream);
    g_trace = trace;
  }
  if (policy)
    attn_tc2_fwd_kernel<true><<<grid, kT2Threads, smem, stream>>>(map_a, map_b, policy, units, T, H, Tkp, vstages, scale, eps,
                                                                  (__nv_bfloat16*)out, cls_row, trace);
  else
    attn_tc2_fwd_kernel<false><<<grid, kT2Threads, smem, stream>>>(map_a, map_b, policy, units, T, H, Tkp, vstages, scale, eps,
                                                                   (__nv_bfloat16*)out, cls_row, trace);
  count_launch();
  return check_launch("d2s_attn_policy_fwd(tcgen05, 2 tiles)");
}

}  // namespace d2s

extern "C" int d2s_debug_read_trace(long long* host) {
  if (!d2s::g_trace) return 1;
  cudaDeviceSynchronize();
  cudaMemcpy(host, d2s::g_trace, 4 * 16 * 8 * 8, cudaMemcpyDeviceToHost);
  return 0;
}
