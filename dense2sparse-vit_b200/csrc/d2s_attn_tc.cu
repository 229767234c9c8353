// Kernel family (4), tensor-core part: policy-masked attention on tcgen05 / TMEM (bf16 in, fp32 accumulate).
//
//   S = Q K^T        tcgen05.mma kind::f16, A = Q tile (128 x 64, K-major, SWIZZLE_128B smem),
//                    B = K (Tkp x 64, K-major, SWIZZLE_128B smem), D = S in TMEM (128 lanes x Tkp cols fp32)
//   P = policy softmax, one query row per thread straight out of TMEM (tcgen05.ld 32x32b), unnormalised
//                    a_ij = exp((s_ij - max_j s_ij) * scale) * m_ij written back to TMEM as packed bf16
//                    over the dead S columns (tcgen05.st), row sums kept in registers
//   O = P V          tcgen05.mma with A = P from TMEM, B = V (MN-major SWIZZLE_128B smem), D = O in TMEM
//   out = (O + (eps/T) * colsum(V)) / (rowsum + eps)     -- the reference's "+eps/T on every entry" term
//                    (vit_models/dynamic_vit.py:213) folded into one per-head vector; CLS row side output.
//
// Execution model: persistent CTAs (2 per SM for T > 128, 4 per SM for T <= 128), each looping over (image, head)
// units.  Warp 4 lane 0 is the control thread: it issues the TMA loads (one 3-D tensor map over the packed
// (B, T, 3*H*64) qkv buffer, box = 64 x 128 rows, SWIZZLE_128B, rows past T zero-filled by the TMA unit) and the
// tcgen05.mma instructions; warps 0-3 are the softmax/epilogue warps (TMEM lane quadrant = warp id).  Inside one
// CTA a unit is a serial chain S-MMA -> softmax -> PV-MMA -> epilogue; the loads of the NEXT unit are issued as soon
// as their buffers die (Q/K after the last S-MMA, V after the last PV-MMA) so they land during the softmax, and the
// other CTAs resident on the SM fill the tensor / MUFU / FMA pipes while this one waits.
//
// Per image-layer (T=197, H=6): 59.6 MFLOP against 611 KB of algorithmic HBM traffic: HBM-bound
// unless the QKV projection is fused; the scores never touch HBM.  The MUFU (exp2) floor is ~64 us for B=1024.
#include <cuda.h>
#include "d2s_common.cuh"

namespace d2s {

constexpr int kTcThreads = 160;  // warps 0-3: softmax/epilogue (TMEM lane quadrant = warp id); warp 4: TMA + MMA issue
constexpr int kTcHD = 64;
constexpr int kTileRows = 128;
constexpr uint32_t kTileBytes = kTileRows * 128;  // one TMA box: 128 rows x 64 bf16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA: 3-D tiled load (coordinates innermost first), completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

struct TcBars {
  uint64_t q_full[2], k_full, v_full, s_full, p_full, o_full, tmem_free;
  uint32_t tmem_base;
  uint32_t pad;
};

// kNT  : 128-row tiles per unit (1: T <= 128, 2: T <= 256)
// kPol : policy given (eps terms, masked exponentials, colsum(V))
template <int kNT, bool kPol>
__global__ void __launch_bounds__(kTcThreads, kNT == 1 ? 4 : 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap qkv_map, const float* __restrict__ policy, int num_units, int T,
                   int H, int Tkp, int box_rows, float scale, float eps, __nv_bfloat16* __restrict__ out,
                   float* __restrict__ cls_row) {
  constexpr int kTmemCols = kNT == 1 ? 128 : 256;
  constexpr int kOCol = kNT == 1 ? 64 : 128;  // O accumulator columns: beyond the packed-P columns [0, Tkp/2)
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B atoms are 1024 B and address based: align the tile region
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  unsigned char* tiles = smem_dyn + pad;
  unsigned char* q_s = tiles;                                 // kNT x (128 x 128 B)
  unsigned char* k_s = q_s + kNT * kTileBytes;                // kNT x (128 x 128 B)
  unsigned char* v_s = k_s + kNT * kTileBytes;                // kNT x (128 x 128 B)
  TcBars* bars = reinterpret_cast<TcBars*>(v_s + kNT * kTileBytes);
  float* pol_s = reinterpret_cast<float*>(bars + 1);          // 256
  float* cls_s = pol_s + 256;                                 // 256
  float* vsum_s = cls_s + 256;                                // 64

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t box_bytes = (uint32_t)box_rows * 128u;

  if (tid == 0) {
    mbar_init(smem_u32(&bars->q_full[0]), 1);
    mbar_init(smem_u32(&bars->q_full[1]), 1);
    mbar_init(smem_u32(&bars->k_full), 1);
    mbar_init(smem_u32(&bars->v_full), 1);
    mbar_init(smem_u32(&bars->s_full), 1);
    mbar_init(smem_u32(&bars->p_full), 128);
    mbar_init(smem_u32(&bars->o_full), 1);
    mbar_init(smem_u32(&bars->tmem_free), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      // =============================== control thread: TMA + MMA issue ===============================
      const uint32_t idesc_s = make_idesc(128, Tkp, 0);
      const uint32_t idesc_o = make_idesc(128, kTcHD, 1);
      const uint64_t kd = make_desc_sw128(smem_u32(k_s), 16, 1024);
      const uint64_t vd = make_desc_sw128(smem_u32(v_s), 16, 1024);
      const int ksteps = Tkp / 16;
      auto issue_qk = [&](int unit) {
        const int b = unit / H, h = unit % H;
        mbar_expect_tx(smem_u32(&bars->k_full), box_bytes * kNT);
#pragma unroll
        for (int t = 0; t < kNT; ++t)
          tma_load_3d(smem_u32(k_s) + t * kTileBytes, &qkv_map, (H + h) * kTcHD, t * kTileRows, b, smem_u32(&bars->k_full));
#pragma unroll
        for (int t = 0; t < kNT; ++t) {
          mbar_expect_tx(smem_u32(&bars->q_full[t]), box_bytes);
          tma_load_3d(smem_u32(q_s) + t * kTileBytes, &qkv_map, h * kTcHD, t * kTileRows, b, smem_u32(&bars->q_full[t]));
        }
      };
      auto issue_v = [&](int unit) {
        const int b = unit / H, h = unit % H;
        mbar_expect_tx(smem_u32(&bars->v_full), box_bytes * kNT);
#pragma unroll
        for (int t = 0; t < kNT; ++t)
          tma_load_3d(smem_u32(v_s) + t * kTileBytes, &qkv_map, (2 * H + h) * kTcHD, t * kTileRows, b, smem_u32(&bars->v_full));
      };
      int unit = blockIdx.x;
      if (unit < num_units) { issue_qk(unit); issue_v(unit); }
      uint32_t g = 0;  // tiles processed by this CTA: parity source for the per-tile barriers
      for (uint32_t it = 0; unit < num_units; unit += gridDim.x, ++it) {
        const int next = unit + gridDim.x;
#pragma unroll
        for (int t = 0; t < kNT; ++t, ++g) {
          if (t * kTileRows >= T) { --g; continue; }  // (never for kNT == 1; T <= 128 with kNT == 2 is not launched)
          if (t == 0) mbar_wait(smem_u32(&bars->k_full), it & 1);
          mbar_wait(smem_u32(&bars->q_full[t]), it & 1);
          if (g > 0) mbar_wait(smem_u32(&bars->tmem_free), (g - 1) & 1);
          tc_fence_after();
          const uint64_t qd = make_desc_sw128(smem_u32(q_s) + t * kTileBytes, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < kTcHD / 16; ++ks) mma_ss(tmem, qd + (uint64_t)(ks * 2), kd + (uint64_t)(ks * 2), idesc_s, ks > 0);
          mma_commit(smem_u32(&bars->s_full));
          const bool last = (t + 1) * kTileRows >= T;
          mbar_wait(smem_u32(&bars->p_full), g & 1);   // softmax done => the S-MMA has completed as well
          if (last && next < num_units) issue_qk(next);  // Q and K buffers are dead: prefetch during PV / epilogue / next softmax
          if (t == 0) mbar_wait(smem_u32(&bars->v_full), it & 1);
          tc_fence_after();
          for (int ks = 0; ks < ksteps; ++ks)
            mma_ts(tmem + kOCol, tmem + (uint32_t)(ks * 8), vd + (uint64_t)(ks * 128), idesc_o, ks > 0);
          mma_commit(smem_u32(&bars->o_full));
          if (last && next < num_units) {
            mbar_wait(smem_u32(&bars->o_full), g & 1);  // V is dead once the PV-MMA has completed
            issue_v(next);
          }
        }
      }
    }
  } else {
    // ===================================== softmax / epilogue warps =====================================
    const int r = tid;  // row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const float k2 = scale * 1.4426950408889634f;
    const float c_eps = kPol ? eps / (float)T : 0.0f;
    const float eps_den = kPol ? eps : 0.0f;
    const int nchunks = Tkp / 16;
    uint32_t g = 0;
    for (uint32_t it = 0, unit = blockIdx.x; (int)unit < num_units; unit += gridDim.x, ++it) {
      const int b = unit / H, h = unit % H;
      if (kPol) {
        // per-unit policy row and column sums of V (for the eps/T term): sum_j V[j][d]
        asm volatile("bar.sync 1, 128;" ::: "memory");  // previous unit's readers of pol_s / vsum_s are done
        for (int j = tid; j < 256; j += 128) pol_s[j] = (j < T) ? policy[(size_t)b * T + j] : 0.0f;
        mbar_wait(smem_u32(&bars->v_full), it & 1);
        if (tid < kTcHD) {
          float acc = 0.f;
          const int cchunk = tid >> 3, within = tid & 7;
          for (int j = 0; j < T; ++j)
            acc += __bfloat162float(*(reinterpret_cast<const __nv_bfloat16*>(v_s + sw128_off(j, cchunk)) + within));
          vsum_s[tid] = acc;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
#pragma unroll
      for (int t = 0; t < kNT; ++t, ++g) {
        if (t * kTileRows >= T) { --g; continue; }
        const int i = t * kTileRows + r;                  // query token of this thread
        const bool warp_active = t * kTileRows + warp * 32 < T;   // whole warp beyond T: nothing to compute
        mbar_wait(smem_u32(&bars->s_full), g & 1);
        tc_fence_after();
        float sum = 0.f;
        const bool want_cls = (cls_row != nullptr) && (i == 0);
        if (warp_active) {
          // pass 1: row max over the T valid columns
          float mx = -INFINITY;
          for (int ch = 0; ch < nchunks; ++ch) {
            uint32_t v[16];
            tmem_ld16_nowait(lane_addr + (uint32_t)(ch * 16), v);
            tmem_ld_wait();
            if (ch * 16 + 16 <= T) {
#pragma unroll
              for (int q = 0; q < 16; ++q) mx = fmaxf(mx, __uint_as_float(v[q]));
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q)
                if (ch * 16 + q < T) mx = fmaxf(mx, __uint_as_float(v[q]));
            }
          }
          const float mxk = mx * k2;
          // pass 2: a_ij = exp2(s*k2 - max*k2) [* mask], packed to bf16 over the consumed S columns
          for (int ch = 0; ch < nchunks; ++ch) {
            uint32_t v[16];
            tmem_ld16_nowait(lane_addr + (uint32_t)(ch * 16), v);
            tmem_ld_wait();
            float a[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int j = ch * 16 + q;
              float e = ex2_approx(fmaf(__uint_as_float(v[q]), k2, -mxk));
              if (kPol) e *= (j == i) ? 1.0f : pol_s[j];
              if (ch * 16 + 16 > T && j >= T) e = 0.f;
              a[q] = e;
              sum += e;
            }
            if (want_cls) {
#pragma unroll
              for (int q = 0; q < 16; ++q) cls_s[ch * 16 + q] = a[q];
            }
            uint32_t packed[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) packed[q] = pack_bf16x2(a[2 * q], a[2 * q + 1]);
            tmem_st8(lane_addr + (uint32_t)(ch * 8), packed);  // P overlays the S columns already consumed
          }
        }  // rows of an idle warp are never written out; whatever their P rows hold stays in those rows
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->p_full));
        const float den = sum + eps_den;
        if (cls_row != nullptr && t == 0 && warp == 0) {
          // CLS row (query 0 = lane 0 of warp 0): probabilities of row 0, Attention.forward's second output
          __syncwarp();
          const float den0 = __shfl_sync(0xffffffffu, den, 0);
          for (int j = lane; j < T; j += 32) cls_row[(size_t)unit * T + j] = (cls_s[j] + c_eps) / den0;
          __syncwarp();
        }
        // ---- epilogue -----------------------------------------------------------------------------------
        mbar_wait(smem_u32(&bars->o_full), g & 1);
        tc_fence_after();
        if (warp_active) {
          const float inv = 1.0f / den;
          __nv_bfloat16* orow = out + ((size_t)b * T + min(i, T - 1)) * (size_t)(H * kTcHD) + (size_t)h * kTcHD;
#pragma unroll
          for (int ch = 0; ch < kTcHD / 16; ++ch) {
            uint32_t v[16];
            tmem_ld16_nowait(lane_addr + (uint32_t)(kOCol + ch * 16), v);
            tmem_ld_wait();
            uint32_t w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float o0 = __uint_as_float(v[2 * q]), o1 = __uint_as_float(v[2 * q + 1]);
              if (kPol) {
                o0 += c_eps * vsum_s[ch * 16 + 2 * q];
                o1 += c_eps * vsum_s[ch * 16 + 2 * q + 1];
              }
              w[q] = pack_bf16x2(o0 * inv, o1 * inv);
            }
            if (i < T) {
              reinterpret_cast<uint4*>(orow)[ch * 2] = make_uint4(w[0], w[1], w[2], w[3]);
              reinterpret_cast<uint4*>(orow)[ch * 2 + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->tmem_free));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

int attn_simt_dispatch(const void* qkv, const float* policy, int dtype, int B, int T, int H, int hd, float scale,
                       float eps, void* out, float* cls_row, cudaStream_t stream);

static bool force_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("D2S_ATTN_FORCE_SIMT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

template <int kNT, bool kPol>
static int launch_tc(const CUtensorMap& map, const float* policy, int units, int T, int H, int Tkp, int box_rows,
                     float scale, float eps, void* out, float* cls_row, cudaStream_t stream) {
  auto kern = attn_tc_fwd_kernel<kNT, kPol>;
  const size_t smem = 1024 + 3 * (size_t)kNT * kTileBytes + sizeof(TcBars) + (256 + 256 + 64) * sizeof(float);
  static bool smem_set = false;  // one flag per instantiation; the opt-in is sticky
  if (!smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "attn_policy_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = true;
  }
  const int per_sm = kNT == 1 ? 4 : 2;
  const int grid = units < per_sm * kNumSMs ? units : per_sm * kNumSMs;
  kern<<<grid, kTcThreads, smem, stream>>>(map, policy, units, T, H, Tkp, box_rows, scale, eps, (__nv_bfloat16*)out, cls_row);
  count_launch();
  return check_launch("d2s_attn_policy_fwd(tcgen05)");
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_attn_policy_fwd(const void* qkv, const float* policy, int dtype, int B, int T, int H, int hd,
                                   float scale, float eps, void* out, float* cls_row, d2s_stream_t stream_) {
  D2S_REQUIRE(qkv && out, D2S_ERR_ARG, "attn_policy_fwd: null pointer");
  D2S_REQUIRE(dtype == D2S_F32 || dtype == D2S_BF16, D2S_ERR_ARG, "attn_policy_fwd: dtype %d unsupported", dtype);
  D2S_REQUIRE(B >= 0 && T >= 1 && H >= 1 && hd >= 1, D2S_ERR_ARG, "attn_policy_fwd: bad shape B=%d T=%d H=%d hd=%d", B, T, H, hd);
  D2S_REQUIRE((long long)B * H <= (1LL << 30), D2S_ERR_ARG, "attn_policy_fwd: B*H=%lld too large", (long long)B * H);
  D2S_REQUIRE(aligned16(qkv) && aligned16(out), D2S_ERR_ALIGN, "attn_policy_fwd: qkv/out must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (dtype == D2S_F32 || force_simt()) {
    D2S_REQUIRE((long long)B * H <= 65535, D2S_ERR_ARG, "attn_policy_fwd(simt): B*H=%lld exceeds 65535", (long long)B * H);
    if (B == 0) return D2S_OK;
    return attn_simt_dispatch(qkv, policy, dtype, B, T, H, hd, scale, eps, out, cls_row, stream);
  }
  D2S_REQUIRE(hd == kTcHD, D2S_ERR_ARG, "attn_policy_fwd(bf16): head dim %d unsupported by the tcgen05 kernel (64 only)", hd);
  D2S_REQUIRE(T <= 256, D2S_ERR_ARG, "attn_policy_fwd(bf16): T=%d exceeds 256", T);
  if (B == 0) return D2S_OK;
  const int Tkp = ceil_div(T, 16) * 16;
  const int box_rows = Tkp < kTileRows ? Tkp : kTileRows;
  EncodeFn enc = encode_fn();
  D2S_REQUIRE(enc != nullptr, D2S_ERR_CUDA, "attn_policy_fwd: cuTensorMapEncodeTiled is unavailable from the driver");
  CUtensorMap map;
  const cuuint64_t gdim[3] = {(cuuint64_t)3 * H * kTcHD, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t gstride[2] = {(cuuint64_t)3 * H * kTcHD * 2, (cuuint64_t)T * 3 * H * kTcHD * 2};
  const cuuint32_t box[3] = {(cuuint32_t)kTcHD, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  D2S_REQUIRE(cr == CUDA_SUCCESS, D2S_ERR_CUDA, "attn_policy_fwd: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  const int units = B * H;
  if (T <= kTileRows) {
    return policy ? launch_tc<1, true>(map, policy, units, T, H, Tkp, box_rows, scale, eps, out, cls_row, stream)
                  : launch_tc<1, false>(map, policy, units, T, H, Tkp, box_rows, scale, eps, out, cls_row, stream);
  }
  return policy ? launch_tc<2, true>(map, policy, units, T, H, Tkp, box_rows, scale, eps, out, cls_row, stream)
                : launch_tc<2, false>(map, policy, units, T, H, Tkp, box_rows, scale, eps, out, cls_row, stream);
}
