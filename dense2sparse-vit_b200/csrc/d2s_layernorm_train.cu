// LayerNorm forward/backward for the training path (Block.forward, vit_models/dynamic_vit.py:263-283, under bf16
// autocast: fp32 residual stream in, bf16 normalised activations out for the following Linear).
//
// Why it exists: in one training step torch spends 12.0 ms in the LayerNorm weight/bias-gradient kernel, 2.7 ms in the
// forward, 1.4 ms in the input-gradient kernel and ~2 ms in the fp32->bf16 casts that follow every LayerNorm
// (profiles/r01f_launches_train_step.csv).  All three passes are plain HBM streaming:
//   fwd : read x (ex*D), write h (eh*D) + (mean, rstd)
//   bwd : read dh, x, write dx; dgamma/dbeta are reduced per CTA in shared memory and flushed with one atomicAdd per
//         column per CTA (rows per CTA = 128 -> 2*D atomics per 128 rows).
// Half-warp per row, 16-byte vectors, row kept in registers (same scheme as d2s_layernorm.cu).
#include "d2s_common.cuh"

namespace d2s {

constexpr int kLtThreads = 256;
constexpr int kLtRowsPerIter = kLtThreads / 16;   // 16 rows per CTA pass
constexpr int kLtRowsPerCta = 128;
constexpr int kLtBwdThreads = 128;                // backward: ~150 registers per thread -> smaller CTAs, 3 per SM
constexpr int kLtBwdRowsPerIter = kLtBwdThreads / 16;

template <typename T_> struct TV;
template <> struct TV<__nv_bfloat16> {
  static constexpr int kElems = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const int4 r = *reinterpret_cast<const int4*>(p);
    const uint32_t w[4] = {(uint32_t)r.x, (uint32_t)r.y, (uint32_t)r.z, (uint32_t)r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&t); }
    *reinterpret_cast<int4*>(p) = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
  }
  __device__ static float round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }
};
template <> struct TV<float> {   // 8 floats = two 16-byte vectors, so both dtypes walk the row in 8-element steps
  static constexpr int kElems = 8;
  __device__ static void load(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ static void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __device__ static float round(float f) { return f; }
};

// LayerNorm over rows `skip`.. of every (seg + skip)-row segment of x (the predictors normalise x[:, 1:], default_dynamic_vit.py
// :461 / dynamic_vit.py:846): dense row r of the output <-> row r + (r / seg + 1) * skip of x.  seg == 0: identity.
__device__ __forceinline__ long long seg_row(long long row, int seg, int skip) {
  return seg > 0 ? row + (row / seg + 1) * skip : row;
}

__device__ __forceinline__ float hw_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// h = (x - mean) * rstd * gamma + beta ; stats (rows, 2) = (mean, rstd)
template <typename TX, typename TH, int kVPL>
__global__ void __launch_bounds__(kLtThreads)
ln_fwd_kernel(const TX* __restrict__ x, const TX* __restrict__ res, const float* __restrict__ gamma,
              const float* __restrict__ beta, long long rows, int D, float eps, TX* __restrict__ out_sum, TH* __restrict__ h,
              float* __restrict__ stats, int seg, int skip) {
  const int sub = threadIdx.x & 15;
  const long long row = (long long)blockIdx.x * kLtRowsPerIter + (threadIdx.x >> 4);
  const bool ok = row < rows;
  const long long xrow = seg_row(row, seg, skip);      // x / res / out_sum live in the segmented layout, h and stats are dense
  const int nvec = D / 8;
  float v[kVPL][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < kVPL; ++k) {
    const int vi = sub + 16 * k;
    if (ok && vi < nvec) {
      TV<TX>::load(x + xrow * D + vi * 8, v[k]);
      if (res) {   // residual add folded in: s = x + res rounded to the stream's dtype (what torch's add would store), kept
        float r[8];
        TV<TX>::load(res + xrow * D + vi * 8, r);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[k][q] = TV<TX>::round(v[k][q] + r[q]);
        TV<TX>::store(out_sum + xrow * D + vi * 8, v[k]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) sum += v[k][q];
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[k][q] = 0.f;
    }
  }
  const float mean = hw_sum(sum) / (float)D;
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < kVPL; ++k)
    if (sub + 16 * k < nvec) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float d = v[k][q] - mean; var = fmaf(d, d, var); }
    }
  const float rstd = rsqrtf(hw_sum(var) / (float)D + eps);
  if (!ok) return;
  if (sub == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
#pragma unroll
  for (int k = 0; k < kVPL; ++k) {
    const int vi = sub + 16 * k;
    if (vi < nvec) {
      float g[8], b[8], o[8];
      TV<float>::load(gamma + vi * 8, g);
      TV<float>::load(beta + vi * 8, b);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = fmaf((v[k][q] - mean) * rstd, g[q], b[q]);
      TV<TH>::store(h + row * D + vi * 8, o);
    }
  }
}

// 16-byte asynchronous global -> shared copies (LDGSTS): the bytes in flight no longer sit in registers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kN> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory"); }

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dh * gamma;  dgamma += dh * xhat;  dbeta += dh
// The inputs of a pass (8 rows: x, dh and the residual-stream gradient) travel through a kStages-deep ring of THREAD-PRIVATE
// shared-memory slots filled by cp.async: every thread copies exactly the 16-byte pieces it will read itself, so the ring needs
// no barrier (cp.async.wait_group orders a thread's own copies), and two passes of loads are in flight per thread without
// costing registers.  With the loads held in registers (167 per thread, 3 CTAs of 128 threads per SM) the kernel had 18 KB in
// flight per SM and ran at 2-3.4 TB/s.  The ring also lets the second half of a pass re-read the row instead of holding it in
// registers across the reductions: 96 registers, four CTAs per SM.
template <typename TX, typename TH, int kVPL, int kStages>
__global__ void __launch_bounds__(kLtBwdThreads)
ln_bwd_kernel(const TH* __restrict__ dh, const TX* __restrict__ x, const float* __restrict__ stats,
              const float* __restrict__ gamma, const TX* __restrict__ gadd, long long rows, int D, TX* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int seg, int skip, int rows_per_cta) {
  extern __shared__ __align__(16) unsigned char ln_smem[];
  float* red = reinterpret_cast<float*>(ln_smem);   // (half-warps per CTA) x D partial column sums
  constexpr int kXB = 8 * (int)sizeof(TX), kHB = 8 * (int)sizeof(TH);          // bytes of one 8-element piece
  constexpr int kStageBytes = kLtBwdThreads * kVPL * (2 * kXB + kHB);
  unsigned char* ring = ln_smem + (size_t)kLtBwdRowsPerIter * D * sizeof(float);
  // slot of (stage, tensor, k) for this thread: consecutive threads -> consecutive 16-byte words (conflict-free)
  auto slot_x = [&](int st, int k) { return ring + (size_t)st * kStageBytes + (size_t)(k * kLtBwdThreads + threadIdx.x) * kXB; };
  auto slot_g = [&](int st, int k) { return ring + (size_t)st * kStageBytes + (size_t)kLtBwdThreads * kVPL * kXB + (size_t)(k * kLtBwdThreads + threadIdx.x) * kXB; };
  auto slot_h = [&](int st, int k) { return ring + (size_t)st * kStageBytes + (size_t)kLtBwdThreads * kVPL * 2 * kXB + (size_t)(k * kLtBwdThreads + threadIdx.x) * kHB; };
  const int sub = threadIdx.x & 15;
  const int nvec = D / 8;
  float ag[kVPL][8], ab[kVPL][8];
#pragma unroll
  for (int k = 0; k < kVPL; ++k)
#pragma unroll
    for (int q = 0; q < 8; ++q) { ag[k][q] = 0.f; ab[k][q] = 0.f; }
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  const long long row_end = min(rows, row_begin + rows_per_cta);
  const int iters = (int)((row_end - row_begin + kLtBwdRowsPerIter - 1) / kLtBwdRowsPerIter);
  auto issue = [&](int it) {
    if (it < iters) {
      const long long row = row_begin + (long long)it * kLtBwdRowsPerIter + (threadIdx.x >> 4);
      if (row < row_end) {
        const long long xrow = seg_row(row, seg, skip);    // x / gadd / dx: segmented layout; dh and stats: dense
        const int st = it % kStages;
#pragma unroll
        for (int k = 0; k < kVPL; ++k) {
          const int vi = sub + 16 * k;
          if (vi < nvec) {
#pragma unroll
            for (int c = 0; c < kXB / 16; ++c) cp_async16(slot_x(st, k) + 16 * c, reinterpret_cast<const unsigned char*>(x + xrow * D + vi * 8) + 16 * c);
#pragma unroll
            for (int c = 0; c < kHB / 16; ++c) cp_async16(slot_h(st, k) + 16 * c, reinterpret_cast<const unsigned char*>(dh + row * D + vi * 8) + 16 * c);
            if (gadd) {
#pragma unroll
              for (int c = 0; c < kXB / 16; ++c) cp_async16(slot_g(st, k) + 16 * c, reinterpret_cast<const unsigned char*>(gadd + xrow * D + vi * 8) + 16 * c);
            }
          }
        }
      }
    }
    cp_async_commit();          // one group per pass, also when empty: the wait below counts groups
  };
#pragma unroll
  for (int p = 0; p < kStages - 1; ++p) issue(p);
  for (int it = 0; it < iters; ++it) {
    issue(it + kStages - 1);
    cp_async_wait<kStages - 1>();                          // this thread's copies of pass `it` have landed
    const int st = it % kStages;
    const long long row = row_begin + (long long)it * kLtBwdRowsPerIter + (threadIdx.x >> 4);
    const bool ok = row < row_end;
    const float mean = ok ? stats[row * 2] : 0.f, rstd = ok ? stats[row * 2 + 1] : 0.f;
    const long long xrow = seg_row(row, seg, skip);
    // pass A: row sums and the column accumulators.  Nothing of the row is kept in registers across the two half-warp reductions:
    // pass B re-reads x and dh from this thread's ring slots (shared memory) and recomputes xhat and g -- 48 registers less, which
    // is what lets a fourth CTA onto the SM (the kernel is latency-bound at three CTAs of four warps).
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kVPL; ++k) {
      const int vi = sub + 16 * k;
      if (ok && vi < nvec) {
        float xv[8], dv[8], g[8];
        TV<TX>::load(reinterpret_cast<const TX*>(slot_x(st, k)), xv);
        TV<TH>::load(reinterpret_cast<const TH*>(slot_h(st, k)), dv);
        TV<float>::load(gamma + vi * 8, g);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float xh = (xv[q] - mean) * rstd, gq = dv[q] * g[q];
          s1 += gq;
          s2 = fmaf(gq, xh, s2);
          ag[k][q] = fmaf(dv[q], xh, ag[k][q]);
          ab[k][q] += dv[q];
        }
      }
    }
    s1 = hw_sum(s1) / (float)D;
    s2 = hw_sum(s2) / (float)D;
    if (ok) {
#pragma unroll
      for (int k = 0; k < kVPL; ++k) {
        const int vi = sub + 16 * k;
        if (vi < nvec) {
          float xv[8], dv[8], g[8], o[8];
          TV<TX>::load(reinterpret_cast<const TX*>(slot_x(st, k)), xv);
          TV<TH>::load(reinterpret_cast<const TH*>(slot_h(st, k)), dv);
          TV<float>::load(gamma + vi * 8, g);
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] = rstd * (dv[q] * g[q] - s1 - (xv[q] - mean) * rstd * s2);
          if (gadd) {   // gradient that reaches x directly (the residual stream's), accumulated here instead of by a torch add
            float ga[8];
            TV<TX>::load(reinterpret_cast<const TX*>(slot_g(st, k)), ga);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] += ga[q];
          }
          TV<TX>::store(dx + xrow * D + vi * 8, o);
          if (seg > 0 && row % seg == 0) {     // the rows the LayerNorm skipped receive no gradient from it: zero (or gadd's)
#pragma unroll
            for (int s = 1; s <= skip; ++s) {
              float z[8];
              if (gadd) TV<TX>::load(gadd + (xrow - s) * D + vi * 8, z);
              else {
#pragma unroll
                for (int q = 0; q < 8; ++q) z[q] = 0.f;
              }
              TV<TX>::store(dx + (xrow - s) * D + vi * 8, z);
            }
          }
        }
      }
    }
  }
  // The CTA's half-warps hold partial column sums: they meet in shared memory as plain stores (one slot per half-warp: shared
  // float atomics are CAS loops), first dgamma then dbeta through the same buffer, then one global atomic per column per CTA.
  float* slot = red + (size_t)(threadIdx.x >> 4) * D;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int k = 0; k < kVPL; ++k) {
      const int vi = sub + 16 * k;
      if (vi < nvec) {
        const float (&a)[8] = pass == 0 ? ag[k] : ab[k];
        *reinterpret_cast<float4*>(slot + vi * 8) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(slot + vi * 8 + 4) = make_float4(a[4], a[5], a[6], a[7]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += kLtBwdThreads) {
      float t = 0.f;
#pragma unroll
      for (int h = 0; h < kLtBwdRowsPerIter; ++h) t += red[(size_t)h * D + c];
      atomicAdd(pass == 0 ? &dgamma[c] : &dbeta[c], t);
    }
    __syncthreads();
  }
}

template <typename TX, typename TH>
static int ln_fwd_launch(const void* x, const void* res, const float* gamma, const float* beta, long long rows, int D, float eps,
                         void* out_sum, void* h, float* stats, int seg, int skip, cudaStream_t st) {
  const int vpl = ceil_div(D / 8, 16);
  const unsigned grid = (unsigned)((rows + kLtRowsPerIter - 1) / kLtRowsPerIter);
  if (vpl <= 3) ln_fwd_kernel<TX, TH, 3><<<grid, kLtThreads, 0, st>>>((const TX*)x, (const TX*)res, gamma, beta, rows, D, eps, (TX*)out_sum, (TH*)h, stats, seg, skip);
  else          ln_fwd_kernel<TX, TH, 6><<<grid, kLtThreads, 0, st>>>((const TX*)x, (const TX*)res, gamma, beta, rows, D, eps, (TX*)out_sum, (TH*)h, stats, seg, skip);
  count_launch();
  return check_launch("d2s_layernorm_fwd");
}
template <typename TX, typename TH, int kVPL, int kStages>
static int ln_bwd_launch_cfg(const void* dh, const void* x, const float* stats, const float* gamma, const void* gadd, long long rows,
                             int D, void* dx, float* dgamma, float* dbeta, int seg, int skip, cudaStream_t st) {
  const size_t smem = (size_t)kLtBwdRowsPerIter * D * sizeof(float) +
                      (size_t)kStages * kLtBwdThreads * kVPL * (2 * 8 * sizeof(TX) + 8 * sizeof(TH));
  // one wave of equally loaded CTAs: as many as fit the GPU at once (shared memory bound), each walking a contiguous row block
  // of at least kLtRowsPerCta rows in 8-row passes and flushing its column sums once (2 D atomics per CTA)
  const int per_sm = (int)((227 * 1024) / (smem + 1024));
  const long long slots = (long long)kNumSMs * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
  long long rpc = (rows + slots - 1) / slots;
  rpc = (rpc + kLtBwdRowsPerIter - 1) / kLtBwdRowsPerIter * kLtBwdRowsPerIter;
  if (rpc < kLtRowsPerCta / 2) rpc = kLtRowsPerCta / 2;           // at least 64 rows per 2 D column atomics
  const unsigned grid = (unsigned)((rows + rpc - 1) / rpc);
  // opt in once per device for the largest row this instantiation takes (D <= 128 kVPL): the memo does not know about D
  constexpr size_t kSmemMax = (size_t)kLtBwdRowsPerIter * (128 * kVPL) * sizeof(float) +
                              (size_t)kStages * kLtBwdThreads * kVPL * (2 * 8 * sizeof(TX) + 8 * sizeof(TH));
  static SmemOptIn opt;
  cudaError_t e = opt_in_smem(opt, ln_bwd_kernel<TX, TH, kVPL, kStages>, (int)kSmemMax);
  D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "layernorm_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  ln_bwd_kernel<TX, TH, kVPL, kStages><<<grid, kLtBwdThreads, smem, st>>>((const TH*)dh, (const TX*)x, stats, gamma, (const TX*)gadd, rows, D,
                                                                           (TX*)dx, dgamma, dbeta, seg, skip, (int)rpc);
  count_launch();
  return check_launch("d2s_layernorm_bwd");
}
template <typename TX, typename TH>
static int ln_bwd_launch(const void* dh, const void* x, const float* stats, const float* gamma, const void* gadd, long long rows,
                         int D, void* dx, float* dgamma, float* dbeta, int seg, int skip, cudaStream_t st) {
  // ring depth: two passes (one in flight while one is consumed); four CTAs of four warps then fit an SM for bf16 rows up to 384
  const int vpl = ceil_div(D / 8, 16);
  if (vpl <= 3) return ln_bwd_launch_cfg<TX, TH, 3, 2>(dh, x, stats, gamma, gadd, rows, D, dx, dgamma, dbeta, seg, skip, st);
  return ln_bwd_launch_cfg<TX, TH, 6, 2>(dh, x, stats, gamma, gadd, rows, D, dx, dgamma, dbeta, seg, skip, st);
}

static int ln_check(const char* what, long long rows, int D, int dx, int dh) {
  D2S_REQUIRE(rows >= 0 && D >= 8 && D % 8 == 0 && D <= 768, D2S_ERR_ARG, "%s: D=%d must be a multiple of 8 in [8,768]", what, D);
  D2S_REQUIRE((dx == D2S_F32 || dx == D2S_BF16) && (dh == D2S_F32 || dh == D2S_BF16), D2S_ERR_ARG, "%s: bad dtypes %d/%d", what, dx, dh);
  return D2S_OK;
}

}  // namespace d2s

using namespace d2s;

static int ln_fwd_entry(const char* what, const void* x, const void* res, int x_dtype, const float* gamma, const float* beta,
                        long long rows, int D, float eps, void* out_sum, void* h, int h_dtype, float* stats, int seg, int skip,
                        d2s_stream_t stream) {
  D2S_REQUIRE(x && gamma && beta && h && stats && (!res || out_sum), D2S_ERR_ARG, "%s: null pointer", what);
  D2S_REQUIRE(seg >= 0 && skip >= 0 && (seg == 0 || rows % seg == 0), D2S_ERR_ARG, "%s: rows=%lld must be whole segments of %d rows", what,
              rows, seg);
  int rc = ln_check(what, rows, D, x_dtype, h_dtype);
  if (rc) return rc;
  D2S_REQUIRE(aligned16(x) && aligned16(h) && aligned16(gamma) && aligned16(beta) && (!res || (aligned16(res) && aligned16(out_sum))),
              D2S_ERR_ALIGN, "%s: 16-byte alignment", what);
  if (rows == 0) return D2S_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == D2S_F32 && h_dtype == D2S_BF16) return ln_fwd_launch<float, __nv_bfloat16>(x, res, gamma, beta, rows, D, eps, out_sum, h, stats, seg, skip, st);
  if (x_dtype == D2S_F32 && h_dtype == D2S_F32) return ln_fwd_launch<float, float>(x, res, gamma, beta, rows, D, eps, out_sum, h, stats, seg, skip, st);
  if (x_dtype == D2S_BF16 && h_dtype == D2S_BF16) return ln_fwd_launch<__nv_bfloat16, __nv_bfloat16>(x, res, gamma, beta, rows, D, eps, out_sum, h, stats, seg, skip, st);
  return ln_fwd_launch<__nv_bfloat16, float>(x, res, gamma, beta, rows, D, eps, out_sum, h, stats, seg, skip, st);
}

static int ln_bwd_entry(const char* what, const void* dh, int h_dtype, const void* x, int x_dtype, const float* stats,
                        const float* gamma, const void* gadd, long long rows, int D, void* dx, float* dgamma, float* dbeta,
                        int seg, int skip, d2s_stream_t stream) {
  D2S_REQUIRE(dh && x && stats && gamma && dx && dgamma && dbeta, D2S_ERR_ARG, "%s: null pointer", what);
  D2S_REQUIRE(seg >= 0 && skip >= 0, D2S_ERR_ARG, "%s: bad segments", what);
  D2S_REQUIRE(seg == 0 || rows % seg == 0, D2S_ERR_ARG, "%s: rows=%lld must be whole segments of %d rows", what, rows, seg);
  int rc = ln_check(what, rows, D, x_dtype, h_dtype);
  if (rc) return rc;
  D2S_REQUIRE(aligned16(dh) && aligned16(x) && aligned16(dx) && aligned16(gamma) && (!gadd || aligned16(gadd)), D2S_ERR_ALIGN,
              "%s: 16-byte alignment", what);
  if (rows == 0) return D2S_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == D2S_F32 && h_dtype == D2S_BF16) return ln_bwd_launch<float, __nv_bfloat16>(dh, x, stats, gamma, gadd, rows, D, dx, dgamma, dbeta, seg, skip, st);
  if (x_dtype == D2S_F32 && h_dtype == D2S_F32) return ln_bwd_launch<float, float>(dh, x, stats, gamma, gadd, rows, D, dx, dgamma, dbeta, seg, skip, st);
  if (x_dtype == D2S_BF16 && h_dtype == D2S_BF16) return ln_bwd_launch<__nv_bfloat16, __nv_bfloat16>(dh, x, stats, gamma, gadd, rows, D, dx, dgamma, dbeta, seg, skip, st);
  return ln_bwd_launch<__nv_bfloat16, float>(dh, x, stats, gamma, gadd, rows, D, dx, dgamma, dbeta, seg, skip, st);
}

extern "C" int d2s_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, long long rows, int D,
                                 float eps, void* h, int h_dtype, float* stats, d2s_stream_t stream) {
  return ln_fwd_entry("layernorm_fwd", x, nullptr, x_dtype, gamma, beta, rows, D, eps, nullptr, h, h_dtype, stats, 0, 0, stream);
}

// The same over rows `skip`.. of every (seg + skip)-row segment of x (LayerNorm(x[:, 1:]) without the slice copy): h and stats
// are dense over the rows = n_segments * seg normalised rows.
extern "C" int d2s_layernorm_seg_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, long long rows, int D,
                                     int seg, int skip, float eps, void* h, int h_dtype, float* stats, d2s_stream_t stream) {
  D2S_REQUIRE(seg > 0, D2S_ERR_ARG, "layernorm_seg_fwd: seg=%d must be positive", seg);
  return ln_fwd_entry("layernorm_seg_fwd", x, nullptr, x_dtype, gamma, beta, rows, D, eps, nullptr, h, h_dtype, stats, seg, skip, stream);
}

extern "C" int d2s_layernorm_bwd(const void* dh, int h_dtype, const void* x, int x_dtype, const float* stats, const float* gamma,
                                 long long rows, int D, void* dx, float* dgamma, float* dbeta, d2s_stream_t stream) {
  return ln_bwd_entry("layernorm_bwd", dh, h_dtype, x, x_dtype, stats, gamma, nullptr, rows, D, dx, dgamma, dbeta, 0, 0, stream);
}

// Backward of d2s_layernorm_seg_fwd: dx has x's segmented layout; the skipped rows of every segment are written as zeros.
extern "C" int d2s_layernorm_seg_bwd(const void* dh, int h_dtype, const void* x, int x_dtype, const float* stats, const float* gamma,
                                     long long rows, int D, int seg, int skip, void* dx, float* dgamma, float* dbeta,
                                     d2s_stream_t stream) {
  D2S_REQUIRE(seg > 0, D2S_ERR_ARG, "layernorm_seg_bwd: seg=%d must be positive", seg);
  return ln_bwd_entry("layernorm_seg_bwd", dh, h_dtype, x, x_dtype, stats, gamma, nullptr, rows, D, dx, dgamma, dbeta, seg, skip, stream);
}

// Residual add folded into the LayerNorm that follows it, with autograd (Block.forward, dynamic_vit.py:276-283: x = x + branch;
// norm(x)): forward writes out_sum = x + res (dtype of x) and h = LayerNorm(out_sum); backward adds the gradient that reaches
// out_sum directly (gsum, may be NULL) to the LayerNorm's input gradient, so neither the add nor the gradient accumulation is
// a separate pass.
extern "C" int d2s_add_layernorm_fwd(const void* x, const void* res, int x_dtype, const float* gamma, const float* beta,
                                     long long rows, int D, float eps, void* out_sum, void* h, int h_dtype, float* stats,
                                     d2s_stream_t stream) {
  D2S_REQUIRE(res && out_sum, D2S_ERR_ARG, "add_layernorm_fwd: null pointer");
  return ln_fwd_entry("add_layernorm_fwd", x, res, x_dtype, gamma, beta, rows, D, eps, out_sum, h, h_dtype, stats, 0, 0, stream);
}

extern "C" int d2s_add_layernorm_bwd(const void* dh, int h_dtype, const void* xsum, int x_dtype, const float* stats,
                                     const float* gamma, const void* gsum, long long rows, int D, void* dx, float* dgamma,
                                     float* dbeta, d2s_stream_t stream) {
  return ln_bwd_entry("add_layernorm_bwd", dh, h_dtype, xsum, x_dtype, stats, gamma, gsum, rows, D, dx, dgamma, dbeta, 0, 0, stream);
}
