// Token distillation loss rows: KL(softmax(t) || softmax(s)) per token over the channel dimension, the
// F.kl_div(F.log_softmax(token_s), F.log_softmax(token_t), log_target=True) of the backbone losses (losses.py:220-225 and
// DynamicViT's DistillDiffPruningLoss) before its "batchmean" reduction.  torch runs it as two log-softmax passes, exp, sub,
// mul, sum (plus a reshape copy of the x[:, 1:] view) forward and as many backward -- ~0.5 ms of a DeiT-S training step on
// (50176, 384) fp32 tensors; here one pass reads both rows once and writes
//     kl[r]      = sum_c softmax(t)[c] * (log_softmax(t)[c] - log_softmax(s)[c])
//     diff[r, c] = softmax(s)[c] - softmax(t)[c]          (= d kl[r] / d s[r, c]: the backward is a row scaling of it)
// A warp owns a row; lane l holds the 8-element chunks l, l + 32, ... (C <= 1024) in registers.  The student / teacher tensors
// may be row-sliced views (batch stride != N * C), so no copy is made of x[:, 1:].
#include "d2s_common.cuh"

namespace d2s {

constexpr int kKlWarps = 8;
constexpr int kKlMaxK = 4;

template <typename T_> struct KlRow;
template <> struct KlRow<float> {
  __device__ static void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct KlRow<__nv_bfloat16> {
  __device__ static void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const int4 r = *reinterpret_cast<const int4*>(p);
    const uint32_t w[4] = {(uint32_t)r.x, (uint32_t)r.y, (uint32_t)r.z, (uint32_t)r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

template <typename TS, typename TT, int kK>
__global__ void __launch_bounds__(kKlWarps * 32)
token_kl_fwd_kernel(const TS* __restrict__ s, long long s_bstride, const TT* __restrict__ t, long long t_bstride, long long rows,
                    int N, int C, float* __restrict__ kl, float* __restrict__ diff) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * kKlWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const long long b = r / N;
  const int n = (int)(r - b * N);
  const TS* sp = s + b * s_bstride + (long long)n * C;
  const TT* tp = t + b * t_bstride + (long long)n * C;
  const int nchunk = C / 8;
  float sv[kK][8], tv[kK][8];
  float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
  for (int k = 0; k < kK; ++k) {
    const int j = lane + 32 * k;
    if (j < nchunk) {
      KlRow<TS>::load8(sp + j * 8, sv[k]);
      KlRow<TT>::load8(tp + j * 8, tv[k]);
#pragma unroll
      for (int q = 0; q < 8; ++q) { ms = fmaxf(ms, sv[k][q]); mt = fmaxf(mt, tv[k][q]); }
    }
  }
  ms = warp_max(ms);
  mt = warp_max(mt);
  // one exponential per element and tensor: e = exp(x - max) is kept and reused for the probabilities; the log-probabilities are
  // (x - max) - log(sum e), so the row needs two logarithms in all.  lds = (s - ms) - (t - mt) overwrites sv.
  float es = 0.f, et = 0.f;
#pragma unroll
  for (int k = 0; k < kK; ++k)
    if (lane + 32 * k < nchunk) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float ds = sv[k][q] - ms, dt = tv[k][q] - mt;
        const float xs = __expf(ds), xt = __expf(dt);
        es += xs;
        et += xt;
        sv[k][q] = xs;
        tv[k][q] = xt;
      }
    }
  es = warp_sum(es);
  et = warp_sum(et);
  // kl = sum_c pt (lpt - lps),  lpt - lps = (t - mt) - (s - ms) - (log et - log es) = log(xt / xs) - log(et / es): computed from
  // the differences of the shifted logits, which are reloaded (L1 hits) rather than kept in a third register array
  const float inv_s = 1.f / es, inv_t = 1.f / et;
  const float lratio = (mt + logf(et)) - (ms + logf(es));         // log-partition difference: lpt - lps = (t - s) - lratio
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < kK; ++k) {
    const int j = lane + 32 * k;
    if (j < nchunk) {
      float s0[8], t0[8], d[8];
      KlRow<TS>::load8(sp + j * 8, s0);
      KlRow<TT>::load8(tp + j * 8, t0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float pt = tv[k][q] * inv_t;
        acc = fmaf(pt, (t0[q] - s0[q]) - lratio, acc);
        d[q] = fmaf(sv[k][q], inv_s, -pt);
      }
      float* dp = diff + r * C + j * 8;
      *reinterpret_cast<float4*>(dp) = make_float4(d[0], d[1], d[2], d[3]);
      *reinterpret_cast<float4*>(dp + 4) = make_float4(d[4], d[5], d[6], d[7]);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) kl[r] = acc;
}

template <typename TS, typename TT>
static int token_kl_launch(const void* s, long long sb, const void* t, long long tb, long long rows, int N, int C, float* kl, float* diff,
                           cudaStream_t st) {
  const long long grid = (rows + kKlWarps - 1) / kKlWarps;
  if (C <= 512)
    token_kl_fwd_kernel<TS, TT, 2><<<(unsigned)grid, kKlWarps * 32, 0, st>>>((const TS*)s, sb, (const TT*)t, tb, rows, N, C, kl, diff);
  else
    token_kl_fwd_kernel<TS, TT, kKlMaxK><<<(unsigned)grid, kKlWarps * 32, 0, st>>>((const TS*)s, sb, (const TT*)t, tb, rows, N, C, kl, diff);
  count_launch();
  return check_launch("d2s_token_kl_fwd");
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_token_kl_fwd(const void* s, int s_dtype, long long s_batch_stride, const void* t, int t_dtype,
                                long long t_batch_stride, int B, int N, int C, float* kl_rows, float* diff, d2s_stream_t stream) {
  D2S_REQUIRE(s && t && kl_rows && diff, D2S_ERR_ARG, "token_kl_fwd: null pointer");
  D2S_REQUIRE((s_dtype == D2S_F32 || s_dtype == D2S_BF16) && (t_dtype == D2S_F32 || t_dtype == D2S_BF16), D2S_ERR_ARG,
              "token_kl_fwd: dtypes %d/%d unsupported", s_dtype, t_dtype);
  D2S_REQUIRE(B >= 0 && N >= 1 && C >= 8 && C % 8 == 0 && C <= 8 * 32 * kKlMaxK, D2S_ERR_ARG,
              "token_kl_fwd: bad shape B=%d N=%d C=%d (C %% 8 == 0, C <= %d)", B, N, C, 8 * 32 * kKlMaxK);
  D2S_REQUIRE(s_batch_stride >= (long long)N * C && t_batch_stride >= (long long)N * C && s_batch_stride % 8 == 0 && t_batch_stride % 8 == 0,
              D2S_ERR_ARG, "token_kl_fwd: batch strides %lld/%lld must cover N*C and be multiples of 8", s_batch_stride, t_batch_stride);
  D2S_REQUIRE(aligned16(s) && aligned16(t) && aligned16(diff), D2S_ERR_ALIGN, "token_kl_fwd: s/t/diff must be 16-byte aligned");
  const long long rows = (long long)B * N;
  if (rows == 0) return D2S_OK;
  D2S_REQUIRE((rows + kKlWarps - 1) / kKlWarps <= 0x7fffffffLL, D2S_ERR_ARG, "token_kl_fwd: too many rows %lld", rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (s_dtype == D2S_F32 && t_dtype == D2S_F32) return token_kl_launch<float, float>(s, s_batch_stride, t, t_batch_stride, rows, N, C, kl_rows, diff, st);
  if (s_dtype == D2S_F32) return token_kl_launch<float, __nv_bfloat16>(s, s_batch_stride, t, t_batch_stride, rows, N, C, kl_rows, diff, st);
  if (t_dtype == D2S_F32) return token_kl_launch<__nv_bfloat16, float>(s, s_batch_stride, t, t_batch_stride, rows, N, C, kl_rows, diff, st);
  return token_kl_launch<__nv_bfloat16, __nv_bfloat16>(s, s_batch_stride, t, t_batch_stride, rows, N, C, kl_rows, diff, st);
}
