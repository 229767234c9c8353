// Token assembly after the patch-embedding GEMM (vit_models/dynamic_vit.py:816-824, default_dynamic_vit.py:437-442):
// x = cat(cls_token.expand(B), patches) + pos_embed in ONE pass -- the reference runs a concat (read+write of the whole
// token matrix) and then a broadcast add (another read+write).  HBM-bound: e*D*N read + e*D*(N+1) written per image.
#include "d2s_common.cuh"

namespace d2s {

template <typename T_, int VE>
__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const T_* __restrict__ patches, const T_* __restrict__ cls, const T_* __restrict__ pos, long long B,
                       int N, int D, T_* __restrict__ out) {
  const int nvec = D / VE;
  const int T = N + 1;
  const long long total = B * T * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const long long row = i / nvec;
    const int t = (int)(row % T);
    const long long b = row / T;
    const T_* src = t == 0 ? cls + (size_t)v * VE : patches + ((size_t)(b * N + (t - 1)) * D + (size_t)v * VE);
    const int4 a = t == 0 ? *reinterpret_cast<const int4*>(src) : ld_stream16(src);
    const int4 p = *reinterpret_cast<const int4*>(pos + (size_t)t * D + (size_t)v * VE);
    int4 r;
    if (VE == 8) {
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&p);
      __nv_bfloat162* r2 = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 fa = __bfloat1622float2(a2[q]), fp = __bfloat1622float2(p2[q]);
        r2[q] = __floats2bfloat162_rn(fa.x + fp.x, fa.y + fp.y);
      }
    } else {
      r.x = __float_as_int(__int_as_float(a.x) + __int_as_float(p.x));
      r.y = __float_as_int(__int_as_float(a.y) + __int_as_float(p.y));
      r.z = __float_as_int(__int_as_float(a.z) + __int_as_float(p.z));
      r.w = __float_as_int(__int_as_float(a.w) + __int_as_float(p.w));
    }
    *reinterpret_cast<int4*>(out + (size_t)row * D + (size_t)v * VE) = r;
  }
}

// im2col of non-overlapping patches: img (B,C,H,W) -> patches (B, gh*gw, C*ph*pw), k = (c, py, px) as Conv2d's weight
// is laid out (PatchEmbed.proj, dynamic_vit.py:296-302).  One 16-byte vector per thread, output order (coalesced
// writes; reads are 32-byte runs of one patch row).
template <int VE>
__global__ void __launch_bounds__(256)
patchify_kernel(const int4* __restrict__ img, long long total, int C, int Hh, int Ww, int ph, int pw, int4* __restrict__ out) {
  const int gh = Hh / ph, gw = Ww / pw;
  const int vec_per_prow = pw / VE;                 // vectors per patch row
  const int vec_per_patch = C * ph * vec_per_prow;  // vectors per output row
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kv = (int)(i % vec_per_patch);
    const long long r = i / vec_per_patch;
    const int gx = (int)(r % gw);
    const int gy = (int)((r / gw) % gh);
    const long long b = r / ((long long)gw * gh);
    const int pxv = kv % vec_per_prow;
    const int py = (kv / vec_per_prow) % ph;
    const int c = kv / (vec_per_prow * ph);
    const long long src = (((b * C + c) * Hh + (gy * ph + py)) * (long long)Ww + gx * pw) / VE + pxv;
    out[i] = ld_stream16(img + src);
  }
}

// Same im2col through shared memory: a CTA owns one band of patches (image b, patch row gy) = C * ph image rows of W pixels.
// Loads walk the image rows (W * elem contiguous bytes each), stores walk the output rows (a patch's C*ph*pw values are
// contiguous), so both sides move whole lines; the one-vector-per-thread kernel above reads 32-byte runs and spends most of
// its instructions on 64-bit index arithmetic (3.6 TB/s).
template <int VE>
__global__ void __launch_bounds__(256)
patchify_band_kernel(const int4* __restrict__ img, int C, int Hh, int Ww, int ph, int pw, int4* __restrict__ out) {
  extern __shared__ int4 band[];                      // [C * ph][W / VE]
  const int gh = Hh / ph, gw = Ww / pw;
  const int b = blockIdx.x / gh, gy = blockIdx.x - b * gh;
  const int wv = Ww / VE, pwv = pw / VE;              // vectors per image row / per patch row
  const int nvec = C * ph * wv;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int r = i / wv, xv = i - r * wv;            // r = c * ph + py
    const int c = r / ph, py = r - c * ph;
    band[i] = ld_stream16(img + ((size_t)(b * C + c) * Hh + (size_t)(gy * ph + py)) * wv + xv);
  }
  __syncthreads();
  const int vpp = C * ph * pwv;                       // vectors per patch (= per output row)
  int4* dst = out + (size_t)blockIdx.x * gw * vpp;
  for (int j = threadIdx.x; j < nvec; j += blockDim.x) {
    const int gx = j / vpp, kv = j - gx * vpp;
    const int r = kv / pwv, pxv = kv - r * pwv;       // r = c * ph + py
    dst[j] = band[r * wv + gx * pwv + pxv];
  }
}

// uint8 images: the host side of the end-to-end path ships raw pixels (half the bytes of bf16, a quarter of fp32 over PCIe) and
// torchvision's ToTensor + Normalize -- x / 255, then (x - mean[c]) / std[c], both in fp32 -- run here.  Each CTA first builds the
// 256-entry table of that expression per channel with the SAME two IEEE operations (division by 255, subtraction, division; no
// FMA contraction, no reciprocal), rounded to the output type, so every pixel is bit-identical to the torch pipeline at the cost
// of one shared-memory lookup; then the band is transposed to patch rows as in patchify_band_kernel.
template <typename OutT>
__global__ void __launch_bounds__(256)
patchify_u8_band_kernel(const uint8_t* __restrict__ img, const float* __restrict__ mean, const float* __restrict__ stdv, int C,
                        int Hh, int Ww, int ph, int pw, OutT* __restrict__ out) {
  extern __shared__ int4 band_raw[];                   // [C * ph][W] bytes, then the table
  uint8_t* band = reinterpret_cast<uint8_t*>(band_raw);
  OutT* lut = reinterpret_cast<OutT*>(band + (size_t)C * ph * Ww);      // [C][256]
  const int gh = Hh / ph, gw = Ww / pw;
  const int b = blockIdx.x / gh, gy = blockIdx.x - b * gh;
  for (int i = threadIdx.x; i < C * 256; i += blockDim.x) {
    const int c = i >> 8;
    const float x = __fdiv_rn((float)(i & 255), 255.0f);
    const float v = __fdiv_rn(__fsub_rn(x, mean[c]), stdv[c]);
    if (sizeof(OutT) == 2) reinterpret_cast<__nv_bfloat16*>(lut)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(lut)[i] = v;
  }
  const int wv = Ww / 16;                              // 16-byte vectors per image row
  const int nvec = C * ph * wv;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int r = i / wv, xv = i - r * wv;             // r = c * ph + py
    const int c = r / ph, py = r - c * ph;
    band_raw[i] = ld_stream16(img + (((size_t)(b * C + c) * Hh + (size_t)(gy * ph + py)) * Ww + (size_t)xv * 16));
  }
  __syncthreads();
  constexpr int VE = 16 / sizeof(OutT);                // output elements per 16-byte store
  const int pwv = pw / VE;                             // vectors per patch row
  const int vpp = C * ph * pwv;                        // vectors per patch (= per output row)
  const int nout = gw * vpp;
  int4* dst = reinterpret_cast<int4*>(out) + (size_t)blockIdx.x * nout;
  for (int j = threadIdx.x; j < nout; j += blockDim.x) {
    const int gx = j / vpp, kv = j - gx * vpp;
    const int r = kv / pwv, pxv = kv - r * pwv;        // r = c * ph + py
    const int c = r / ph;
    const uint8_t* px = band + (size_t)r * Ww + gx * pw + pxv * VE;
    const OutT* t = lut + c * 256;
    OutT v[VE];
#pragma unroll
    for (int q = 0; q < VE; ++q) v[q] = t[px[q]];
    dst[j] = *reinterpret_cast<const int4*>(v);
  }
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_patchify_u8(const void* img, const float* mean, const float* stdv, int out_dtype, int B, int C, int Hh, int Ww,
                               int ph, int pw, void* out, d2s_stream_t stream) {
  D2S_REQUIRE(img && mean && stdv && out, D2S_ERR_ARG, "patchify_u8: null pointer");
  D2S_REQUIRE(out_dtype == D2S_F32 || out_dtype == D2S_BF16, D2S_ERR_ARG, "patchify_u8: output dtype %d unsupported", out_dtype);
  const int ve = out_dtype == D2S_BF16 ? 8 : 4;
  D2S_REQUIRE(B >= 0 && C >= 1 && C <= 4 && ph >= 1 && pw >= ve && Hh % ph == 0 && Ww % pw == 0 && pw % ve == 0 && Ww % 16 == 0,
              D2S_ERR_ARG, "patchify_u8: bad shape B=%d C=%d H=%d W=%d patch=%dx%d (C <= 4, W %% 16 == 0, patch width %% %d == 0)", B,
              C, Hh, Ww, ph, pw, ve);
  const size_t smem = (size_t)C * ph * Ww + (size_t)C * 256 * (out_dtype == D2S_BF16 ? 2 : 4);
  D2S_REQUIRE(smem <= 48 * 1024 && (long long)B * (Hh / ph) <= 0x7fffffffLL, D2S_ERR_ARG,
              "patchify_u8: a band of C*ph image rows must fit 48 KB of shared memory (needs %zu B)", smem);
  D2S_REQUIRE(aligned16(img) && aligned16(out), D2S_ERR_ALIGN, "patchify_u8: pointers must be 16-byte aligned");
  if (B == 0) return D2S_OK;
  const unsigned grid = (unsigned)((long long)B * (Hh / ph));
  if (out_dtype == D2S_BF16)
    patchify_u8_band_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>((const uint8_t*)img, mean, stdv, C, Hh, Ww, ph,
                                                                                      pw, (__nv_bfloat16*)out);
  else
    patchify_u8_band_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>((const uint8_t*)img, mean, stdv, C, Hh, Ww, ph, pw,
                                                                              (float*)out);
  count_launch();
  return check_launch("d2s_patchify_u8");
}

extern "C" int d2s_patchify(const void* img, int dtype, int B, int C, int Hh, int Ww, int ph, int pw, void* out,
                            d2s_stream_t stream) {
  D2S_REQUIRE(img && out, D2S_ERR_ARG, "patchify: null pointer");
  D2S_REQUIRE(dtype == D2S_F32 || dtype == D2S_BF16, D2S_ERR_ARG, "patchify: dtype %d unsupported", dtype);
  const int ve = dtype == D2S_BF16 ? 8 : 4;
  D2S_REQUIRE(B >= 0 && C >= 1 && ph >= 1 && pw >= ve && Hh % ph == 0 && Ww % pw == 0 && pw % ve == 0, D2S_ERR_ARG,
              "patchify: bad shape B=%d C=%d H=%d W=%d patch=%dx%d (patch width must be a multiple of %d)", B, C, Hh, Ww, ph,
              pw, ve);
  D2S_REQUIRE(aligned16(img) && aligned16(out), D2S_ERR_ALIGN, "patchify: pointers must be 16-byte aligned");
  if (B == 0) return D2S_OK;
  const size_t band_bytes = (size_t)C * ph * Ww * (dtype == D2S_BF16 ? 2 : 4);
  if (band_bytes <= 48 * 1024 && (long long)B * (Hh / ph) <= 0x7fffffffLL) {
    const unsigned grid = (unsigned)((long long)B * (Hh / ph));
    if (dtype == D2S_BF16)
      patchify_band_kernel<8><<<grid, 256, band_bytes, (cudaStream_t)stream>>>((const int4*)img, C, Hh, Ww, ph, pw, (int4*)out);
    else
      patchify_band_kernel<4><<<grid, 256, band_bytes, (cudaStream_t)stream>>>((const int4*)img, C, Hh, Ww, ph, pw, (int4*)out);
    count_launch();
    return check_launch("d2s_patchify");
  }
  const long long total = (long long)B * C * Hh * Ww / ve;
  long long blocks = (total + 255) / 256;
  if (blocks > 32LL * kNumSMs) blocks = 32LL * kNumSMs;
  if (dtype == D2S_BF16)
    patchify_kernel<8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const int4*)img, total, C, Hh, Ww, ph, pw, (int4*)out);
  else
    patchify_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const int4*)img, total, C, Hh, Ww, ph, pw, (int4*)out);
  count_launch();
  return check_launch("d2s_patchify");
}

extern "C" int d2s_assemble_tokens(const void* patches, const void* cls, const void* pos, int dtype, int B, int N, int D,
                                   void* out, d2s_stream_t stream) {
  D2S_REQUIRE(patches && cls && pos && out, D2S_ERR_ARG, "assemble_tokens: null pointer");
  D2S_REQUIRE(dtype == D2S_F32 || dtype == D2S_BF16, D2S_ERR_ARG, "assemble_tokens: dtype %d unsupported", dtype);
  const int ve = dtype == D2S_BF16 ? 8 : 4;
  D2S_REQUIRE(B >= 0 && N >= 1 && D >= ve && D % ve == 0, D2S_ERR_ARG, "assemble_tokens: bad shape B=%d N=%d D=%d", B, N, D);
  D2S_REQUIRE(aligned16(patches) && aligned16(cls) && aligned16(pos) && aligned16(out), D2S_ERR_ALIGN,
              "assemble_tokens: pointers must be 16-byte aligned");
  if (B == 0) return D2S_OK;
  const long long total = (long long)B * (N + 1) * (D / ve);
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
  if (dtype == D2S_BF16)
    assemble_tokens_kernel<__nv_bfloat16, 8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)patches, (const __nv_bfloat16*)cls, (const __nv_bfloat16*)pos, B, N, D, (__nv_bfloat16*)out);
  else
    assemble_tokens_kernel<float, 4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float*)patches, (const float*)cls, (const float*)pos, B, N, D, (float*)out);
  count_launch();
  return check_launch("d2s_assemble_tokens");
}
