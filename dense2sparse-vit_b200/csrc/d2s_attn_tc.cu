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
// One CTA per (image, head, 128-query-row tile); 2 CTAs per SM (256 TMEM columns and ~70 KB smem each).
// Per image-layer (T=197, H=6): 59.6 MFLOP against 611 KB of algorithmic HBM traffic:
// HBM-bound unless the QKV projection is fused; the scores never touch HBM.
#include "d2s_common.cuh"

namespace d2s {

constexpr int kTcThreads = 160;  // warps 0-3: softmax/epilogue (TMEM lane quadrant = warp id); warp 4: alloc + MMA issue
constexpr int kTcHD = 64;
constexpr int kTmemCols = 256;
constexpr int kOCol = 128;       // O accumulator columns [128,192): beyond the packed-P columns [0, Tkp/2)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
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

__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);  // .x = lo -> low 16 bits
  return *reinterpret_cast<const uint32_t*>(&t);
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

struct TcSmemTail {
  uint64_t bar_s, bar_p, bar_o;
  uint32_t tmem_base;
  float den0;
  float pol[256];
  float cls[256];
  float vsum[kTcHD];
};

__global__ void __launch_bounds__(kTcThreads, 2)
attn_tc_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ policy, int T, int H, int Tkp,
                   float scale, float eps, __nv_bfloat16* __restrict__ out, float* __restrict__ cls_row) {
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B atoms are 1024 B and address based: align the tile region
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  unsigned char* tiles = smem_dyn + pad;
  unsigned char* q_s = tiles;                         // 128 x 128 B
  unsigned char* k_s = q_s + 128 * 128;               // Tkp x 128 B
  unsigned char* v_s = k_s + (size_t)Tkp * 128;       // Tkp x 128 B
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(v_s + (size_t)Tkp * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtile = blockIdx.x, bh = blockIdx.y, b = bh / H, h = bh % H;
  const int m0 = mtile * 128;
  const size_t tok_stride = (size_t)3 * H * kTcHD;  // elements between consecutive tokens
  const __nv_bfloat16* q_g = qkv + (size_t)b * T * tok_stride + (size_t)h * kTcHD;
  const __nv_bfloat16* k_g = q_g + (size_t)H * kTcHD;
  const __nv_bfloat16* v_g = q_g + (size_t)2 * H * kTcHD;

  // ---- stage Q tile, K, V (rows beyond T are zero-filled) -------------------------------------
  for (int idx = tid; idx < 128 * 8; idx += kTcThreads) {
    const int r = idx >> 3, c = idx & 7, row = m0 + r;
    cp_async16(smem_u32(q_s) + sw128_off(r, c), q_g + (size_t)min(row, T - 1) * tok_stride + c * 8, row < T);
  }
  for (int idx = tid; idx < Tkp * 8; idx += kTcThreads) {
    const int r = idx >> 3, c = idx & 7;
    const size_t goff = (size_t)min(r, T - 1) * tok_stride + c * 8;
    cp_async16(smem_u32(k_s) + sw128_off(r, c), k_g + goff, r < T);
    cp_async16(smem_u32(v_s) + sw128_off(r, c), v_g + goff, r < T);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int j = tid; j < 256; j += kTcThreads) tail->pol[j] = (policy && j < T) ? policy[(size_t)b * T + j] : 1.0f;

  if (tid == 0) {
    mbar_init(smem_u32(&tail->bar_s), 1);
    mbar_init(smem_u32(&tail->bar_p), 128);
    mbar_init(smem_u32(&tail->bar_o), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  fence_async_smem();  // make the staged tiles visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      // ---- S = Q K^T ---------------------------------------------------------------------------
      const uint32_t idesc_s = make_idesc(128, Tkp, 0);
      const uint64_t qd = make_desc_sw128(smem_u32(q_s), 16, 1024);
      const uint64_t kd = make_desc_sw128(smem_u32(k_s), 16, 1024);
#pragma unroll
      for (int ks = 0; ks < kTcHD / 16; ++ks) mma_ss(tmem, qd + (uint64_t)(ks * 2), kd + (uint64_t)(ks * 2), idesc_s, ks > 0);
      mma_commit(smem_u32(&tail->bar_s));
      // ---- O = P V once the softmax warps have written P --------------------------------------------
      mbar_wait(smem_u32(&tail->bar_p), 0);
      tc_fence_after();
      const uint32_t idesc_o = make_idesc(128, kTcHD, 1);
      const uint64_t vd = make_desc_sw128(smem_u32(v_s), 16, 1024);
      const int ksteps = Tkp / 16;
      for (int ks = 0; ks < ksteps; ++ks)
        mma_ts(tmem + kOCol, tmem + (uint32_t)(ks * 8), vd + (uint64_t)(ks * 128), idesc_o, ks > 0);
      mma_commit(smem_u32(&tail->bar_o));
    }
  } else {
    // ---- softmax: thread == query row ---------------------------------------------------------------
    const int r = tid;                // row inside the tile == TMEM lane
    const int i = m0 + r;             // query token
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const float k2 = scale * 1.4426950408889634f;
    const float c_eps = policy ? eps / (float)T : 0.0f;
    const float eps_den = policy ? eps : 0.0f;
    if (policy && tid < kTcHD) {
      // column sums of V for the eps/T term: sum_j V[j][d]
      float acc = 0.f;
      const int d = tid, cchunk = d >> 3, within = d & 7;
      for (int j = 0; j < T; ++j) {
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(v_s + sw128_off(j, cchunk)) + within;
        acc += __bfloat162float(*p);
      }
      tail->vsum[d] = acc;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // vsum visible to all softmax threads
    mbar_wait(smem_u32(&tail->bar_s), 0);
    tc_fence_after();
    const int nchunks = Tkp / 16;
    float mx = -INFINITY;
    for (int ch = 0; ch < nchunks; ++ch) {
      float v[16];
      tmem_ld16(lane_addr + (uint32_t)(ch * 16), v);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (ch * 16 + q < T) mx = fmaxf(mx, v[q]);
    }
    float sum = 0.f;
    const bool want_cls = (cls_row != nullptr) && (i == 0);
    for (int ch = 0; ch < nchunks; ++ch) {
      float v[16];
      tmem_ld16(lane_addr + (uint32_t)(ch * 16), v);
      uint32_t packed[8];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int j = ch * 16 + q;
        float a = 0.f;
        if (j < T) a = exp2f((v[q] - mx) * k2) * ((j == i) ? 1.0f : tail->pol[j]);
        sum += a;
        if (want_cls) tail->cls[j] = a;
        v[q] = a;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) packed[q] = pack_bf16x2(v[2 * q], v[2 * q + 1]);
      tmem_st8(lane_addr + (uint32_t)(ch * 8), packed);  // P overlays the S columns already consumed
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    mbar_arrive(smem_u32(&tail->bar_p));
    const float den = sum + eps_den;
    if (want_cls) tail->den0 = den;
    // ---- epilogue ---------------------------------------------------------------------------------------
    mbar_wait(smem_u32(&tail->bar_o), 0);
    tc_fence_after();
    const float inv = 1.0f / den;
    __nv_bfloat16* orow = out + ((size_t)b * T + min(i, T - 1)) * (size_t)(H * kTcHD) + (size_t)h * kTcHD;
#pragma unroll
    for (int ch = 0; ch < kTcHD / 16; ++ch) {
      float v[16];
      tmem_ld16(lane_addr + (uint32_t)(kOCol + ch * 16), v);
      uint32_t w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float o0 = v[2 * q], o1 = v[2 * q + 1];
        if (policy) {
          o0 += c_eps * tail->vsum[ch * 16 + 2 * q];
          o1 += c_eps * tail->vsum[ch * 16 + 2 * q + 1];
        }
        w[q] = pack_bf16x2(o0 * inv, o1 * inv);
      }
      if (i < T) {
        reinterpret_cast<uint4*>(orow)[ch * 2] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4*>(orow)[ch * 2 + 1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (cls_row != nullptr && mtile == 0) {
    const float den0 = tail->den0;
    const float c_eps = policy ? eps / (float)T : 0.0f;
    for (int j = tid; j < T; j += kTcThreads) cls_row[(size_t)bh * T + j] = (tail->cls[j] + c_eps) / den0;
  }
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

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_attn_policy_fwd(const void* qkv, const float* policy, int dtype, int B, int T, int H, int hd,
                                   float scale, float eps, void* out, float* cls_row, d2s_stream_t stream_) {
  D2S_REQUIRE(qkv && out, D2S_ERR_ARG, "attn_policy_fwd: null pointer");
  D2S_REQUIRE(dtype == D2S_F32 || dtype == D2S_BF16, D2S_ERR_ARG, "attn_policy_fwd: dtype %d unsupported", dtype);
  D2S_REQUIRE(B >= 0 && T >= 1 && H >= 1 && hd >= 1, D2S_ERR_ARG, "attn_policy_fwd: bad shape B=%d T=%d H=%d hd=%d", B, T, H, hd);
  D2S_REQUIRE((long long)B * H <= 65535, D2S_ERR_ARG, "attn_policy_fwd: B*H=%lld exceeds 65535", (long long)B * H);
  D2S_REQUIRE(aligned16(qkv) && aligned16(out), D2S_ERR_ALIGN, "attn_policy_fwd: qkv/out must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (B == 0) return D2S_OK;
  if (dtype == D2S_F32 || force_simt())
    return attn_simt_dispatch(qkv, policy, dtype, B, T, H, hd, scale, eps, out, cls_row, stream);
  D2S_REQUIRE(hd == kTcHD, D2S_ERR_ARG, "attn_policy_fwd(bf16): head dim %d unsupported by the tcgen05 kernel (64 only)", hd);
  D2S_REQUIRE(T <= 256, D2S_ERR_ARG, "attn_policy_fwd(bf16): T=%d exceeds 256", T);
  const int Tkp = ceil_div(T, 16) * 16;
  const size_t smem = 1024 + 128 * 128 + 2 * (size_t)Tkp * 128 + sizeof(TcSmemTail);
  static size_t smem_set = 0;  // opt-in is sticky per function; skipped once raised (and during graph capture)
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "attn_policy_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = 227 * 1024;
  }
  dim3 grid(ceil_div(T, 128), B * H);
  attn_tc_fwd_kernel<<<grid, kTcThreads, smem, stream>>>((const __nv_bfloat16*)qkv, policy, T, H, Tkp, scale, eps,
                                                         (__nv_bfloat16*)out, cls_row);
  count_launch();
  return check_launch("d2s_attn_policy_fwd(tcgen05)");
}
