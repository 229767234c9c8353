// Head-major relayout for the training attention (Attention.forward, vit_models/dynamic_vit.py:218-221: the reference's
// reshape(B,N,3,H,hd).permute(2,0,3,1,4) leaves strided views that torch's bmm copies and, with T = 197 rows, runs on
// unaligned legacy GEMM kernels).  One pass each way:
//   split:  src (B,T,G,H,hd) token-major packed (G = 3: qkv Linear output; G = 1: attention output gradient)
//           -> dst (G,B,H,Tp,hd) head-major, rows T..Tp-1 written as zeros (Tp = round_up(T, 8): every GEMM dimension and
//           leading dimension of the batched products is then a multiple of 8 elements)
//   merge:  the inverse, dropping the padding rows (attention output, and dq/dk/dv straight into the packed qkv gradient)
// HBM-bound copies: 16-byte vectors, a head row (hd bf16 = 128 B at hd = 64) is moved by hd/8 consecutive lanes.
#include "d2s_common.cuh"

namespace d2s {

template <bool kSplit>
__global__ void __launch_bounds__(256)
heads_relayout_kernel(const int4* __restrict__ src, int4* __restrict__ dst, long long total, int B, int T, int Tp, int G, int H,
                      int vec) {
  // one thread per 16-byte vector of the head-major tensor (G,B,H,Tp,vec)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    long long r = i / vec;
    const int t = (int)(r % Tp);
    r /= Tp;
    const int h = (int)(r % H);
    r /= H;
    const int b = (int)(r % B);
    const int g = (int)(r / B);
    const size_t tok = ((((size_t)b * T + t) * G + g) * H + h) * vec + v;      // token-major (B,T,G,H,vec)
    if (kSplit) {
      dst[i] = t < T ? ld_stream16(src + tok) : make_int4(0, 0, 0, 0);
    } else if (t < T) {
      st_stream16(dst + tok, src[i]);
    }
  }
}

}  // namespace d2s

using namespace d2s;

static int heads_relayout(bool split, const void* src, void* dst, int B, int T, int Tp, int G, int H, int hd, cudaStream_t stream,
                          const char* what) {
  D2S_REQUIRE(src && dst, D2S_ERR_ARG, "%s: null pointer", what);
  D2S_REQUIRE(B >= 0 && T >= 1 && Tp >= T && G >= 1 && H >= 1 && hd >= 8 && hd % 8 == 0, D2S_ERR_ARG,
              "%s: bad shape B=%d T=%d Tp=%d G=%d H=%d hd=%d (bf16, hd %% 8 == 0)", what, B, T, Tp, G, H, hd);
  D2S_REQUIRE(aligned16(src) && aligned16(dst), D2S_ERR_ALIGN, "%s: pointers must be 16-byte aligned", what);
  if (B == 0) return D2S_OK;
  const int vec = hd / 8;
  const long long total = (long long)G * B * H * Tp * vec;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)kNumSMs * 16 ? want : (long long)kNumSMs * 16);
  if (split)
    heads_relayout_kernel<true><<<grid, 256, 0, stream>>>((const int4*)src, (int4*)dst, total, B, T, Tp, G, H, vec);
  else
    heads_relayout_kernel<false><<<grid, 256, 0, stream>>>((const int4*)src, (int4*)dst, total, B, T, Tp, G, H, vec);
  count_launch();
  return check_launch(what);
}

extern "C" int d2s_split_heads_bf16(const void* src, int B, int T, int Tp, int G, int H, int hd, void* dst, d2s_stream_t stream) {
  return heads_relayout(true, src, dst, B, T, Tp, G, H, hd, (cudaStream_t)stream, "d2s_split_heads_bf16");
}

extern "C" int d2s_merge_heads_bf16(const void* src, int B, int T, int Tp, int G, int H, int hd, void* dst, d2s_stream_t stream) {
  return heads_relayout(false, src, dst, B, T, Tp, G, H, hd, (cudaStream_t)stream, "d2s_merge_heads_bf16");
}
