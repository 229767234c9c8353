// AdamW over ONE flat fp32 parameter buffer (training step, BASELINE configs[2]; the reference builds timm's AdamW in
// mask_predictor.py / ddp_training.py through create_optimizer and calls optimizer.step() once per batch, train.py:63-66).
// torch's capturable multi-tensor AdamW is ~15 passes over the parameters plus, because the bias-correction divisors are 0-dim
// tensors, one div_ launch PER PARAMETER twice a step (364 launches for DeiT-S); here the step is one stream over
// (p, g, m, v): 16 B read + 12 B written per parameter, + 2 B for the bf16 copy of the weights the next forward's GEMMs read
// (the per-step cast pass of ops.BF16WeightCache disappears).
//     g' = g * grad_scale                               (the 1/world of the data-parallel mean folded in)
//     p  = p * (1 - lr * wd)
//     m  = m + (g' - m) * (1 - beta1)                   v = v * beta2 + (1 - beta2) * g' * g'
//     p  = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// lr and the step count t are read from device memory, so one captured CUDA graph serves every step of a schedule.
#include "d2s_common.cuh"

namespace d2s {

struct AdamHyper {
  float beta1, beta2, eps, wd, grad_scale;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamHyper& h, float decay, float step_size,
                                          float inv_bc2_sqrt) {
  g *= h.grad_scale;
  p *= decay;
  m = fmaf(g - m, 1.f - h.beta1, m);
  v = fmaf(g * g, 1.f - h.beta2, v * h.beta2);
  const float denom = fmaf(sqrtf(v), inv_bc2_sqrt, h.eps);
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  __nv_bfloat16* __restrict__ shadow, long long begin, long long end, const float* __restrict__ lr_ptr,
                  const float* __restrict__ step_ptr, AdamHyper h) {
  const float lr = __ldg(lr_ptr), t = __ldg(step_ptr);
  const float bc1 = 1.f - powf(h.beta1, t), bc2 = 1.f - powf(h.beta2, t);
  const float step_size = lr / bc1, inv_bc2_sqrt = 1.f / sqrtf(bc2), decay = 1.f - lr * h.wd;
  const long long v0 = begin >> 2, v1 = (end + 3) >> 2;          // 4-element vectors touched by [begin, end)
  for (long long i = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i << 2;
    if (e >= begin && e + 4 <= end) {
      float4 pp = *reinterpret_cast<const float4*>(p + e);
      const float4 gg = *reinterpret_cast<const float4*>(g + e);
      float4 mm = *reinterpret_cast<const float4*>(m + e);
      float4 vv = *reinterpret_cast<const float4*>(v + e);
      adamw_one(pp.x, gg.x, mm.x, vv.x, h, decay, step_size, inv_bc2_sqrt);
      adamw_one(pp.y, gg.y, mm.y, vv.y, h, decay, step_size, inv_bc2_sqrt);
      adamw_one(pp.z, gg.z, mm.z, vv.z, h, decay, step_size, inv_bc2_sqrt);
      adamw_one(pp.w, gg.w, mm.w, vv.w, h, decay, step_size, inv_bc2_sqrt);
      *reinterpret_cast<float4*>(p + e) = pp;
      *reinterpret_cast<float4*>(m + e) = mm;
      *reinterpret_cast<float4*>(v + e) = vv;
      if (shadow) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
        uint2 w;
        w.x = *reinterpret_cast<const uint32_t*>(&lo);
        w.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(shadow + e) = w;
      }
    } else {                                                       // a group boundary inside this vector
      for (long long j = max(e, begin); j < min(e + 4, end); ++j) {
        float pj = p[j], mj = m[j], vj = v[j];
        adamw_one(pj, g[j], mj, vj, h, decay, step_size, inv_bc2_sqrt);
        p[j] = pj; m[j] = mj; v[j] = vj;
        if (shadow) shadow[j] = __float2bfloat16_rn(pj);
      }
    }
  }
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_adamw_flat_f32(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long begin, long long end,
                                  const float* lr, const float* step, float beta1, float beta2, float eps, float weight_decay,
                                  float grad_scale, d2s_stream_t stream) {
  D2S_REQUIRE(p && g && m && v && lr && step, D2S_ERR_ARG, "adamw: null pointer");
  D2S_REQUIRE(begin >= 0 && end >= begin, D2S_ERR_ARG, "adamw: bad range [%lld, %lld)", begin, end);
  D2S_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, D2S_ERR_ARG,
              "adamw: betas (%g, %g) must lie in [0, 1), eps %g >= 0", beta1, beta2, eps);
  D2S_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!shadow_bf16 || aligned16(shadow_bf16)), D2S_ERR_ALIGN,
              "adamw: the flat buffers must be 16-byte aligned");
  if (end == begin) return D2S_OK;
  const long long vecs = ((end + 3) >> 2) - (begin >> 2);
  long long grid = (vecs + 255) / 256;
  const long long cap = 16LL * kNumSMs;                            // 8 resident CTAs of 256 threads per SM, two waves
  if (grid > cap) grid = cap;
  AdamHyper h{beta1, beta2, eps, weight_decay, grad_scale};
  adamw_flat_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (__nv_bfloat16*)shadow_bf16, begin, end, lr, step, h);
  count_launch();
  return check_launch("d2s_adamw_flat_f32");
}
