// Kernel family (3): gather of kept tokens (+CLS) and its backward (scatter + zero fill).
//
// HBM-bound byte movement.  A token row is D*elem bytes (768 B for DeiT-S bf16); rows are moved as
// 16-byte vectors, one warp per row, UNROLL rows in flight per warp so that each SM keeps tens of KB
// outstanding (B200 needs ~35 KB/SM in flight to cover HBM latency at 6.5 TB/s).  The instruction
// budget is ~5 warp-instructions per 16 B at speed of light, so there is no per-vector index math:
// the row index is loaded once per row (warp-uniform) and lanes stride over the row.
//
// Algorithmic bytes per image: fwd 2*e*D*(K+1) + 8*(K+1); bwd e*D*((K+1)+T) + 8*(K+1).
#include "d2s_common.cuh"

namespace d2s {

constexpr int kGatherThreads = 256;
constexpr int kGatherWarps = kGatherThreads / 32;
constexpr int kGatherUnroll = 4;

// out[b, r, :] = x[b, src(r), :],  src(r) = prepend_cls ? (r == 0 ? 0 : idx[b, r-1] + 1) : idx[b, r]
__global__ void __launch_bounds__(kGatherThreads)
gather_rows_kernel(const int4* __restrict__ x, const int64_t* __restrict__ idx, int4* __restrict__ out,
                   int T, int K, int rows_out, int row_vecs, int prepend_cls) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp_in_img = blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
  const int warps_per_img = gridDim.x * kGatherWarps;
  const int64_t* idx_b = idx + (size_t)b * K;
  const int4* x_b = x + (size_t)b * T * row_vecs;
  int4* out_b = out + (size_t)b * rows_out * row_vecs;

  for (int r0 = warp_in_img * kGatherUnroll; r0 < rows_out; r0 += warps_per_img * kGatherUnroll) {
    int src[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const int r = r0 + u;
      int s = 0;
      if (r < rows_out) {
        if (prepend_cls) s = (r == 0) ? 0 : (int)idx_b[r - 1] + 1;
        else             s = (int)idx_b[r];
      }
      src[u] = min(max(s, 0), T - 1);  // out-of-range indices are clamped, never dereferenced wild
    }
    for (int v0 = lane; v0 < row_vecs; v0 += 32 * 2) {
      int4 buf[kGatherUnroll][2];
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int v = v0 + 32 * h;
          if (r0 + u < rows_out && v < row_vecs) buf[u][h] = ld_stream16(x_b + (size_t)src[u] * row_vecs + v);
        }
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int v = v0 + 32 * h;
          if (r0 + u < rows_out && v < row_vecs) st_stream16(out_b + (size_t)(r0 + u) * row_vecs + v, buf[u][h]);
        }
    }
  }
}

// gx[b, t, :] = gout[b, j, :] if row t was selected at output position j, else 0
__global__ void __launch_bounds__(kGatherThreads)
scatter_rows_kernel(const int4* __restrict__ gout, const int64_t* __restrict__ idx, int4* __restrict__ gx,
                    int T, int K, int rows_out, int row_vecs, int prepend_cls) {
  extern __shared__ int inv[];  // T entries: output position of input row t, or -1
  const int b = blockIdx.y;
  const int64_t* idx_b = idx + (size_t)b * K;
  for (int t = threadIdx.x; t < T; t += blockDim.x) inv[t] = -1;
  __syncthreads();
  for (int r = threadIdx.x; r < rows_out; r += blockDim.x) {
    int s;
    if (prepend_cls) s = (r == 0) ? 0 : (int)idx_b[r - 1] + 1;
    else             s = (int)idx_b[r];
    if (s >= 0 && s < T) inv[s] = r;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp_in_img = blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
  const int warps_per_img = gridDim.x * kGatherWarps;
  const int4* gout_b = gout + (size_t)b * rows_out * row_vecs;
  int4* gx_b = gx + (size_t)b * T * row_vecs;
  const int4 zero = make_int4(0, 0, 0, 0);

  for (int t0 = warp_in_img * kGatherUnroll; t0 < T; t0 += warps_per_img * kGatherUnroll) {
    int src[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) src[u] = (t0 + u < T) ? inv[t0 + u] : -1;
    for (int v0 = lane; v0 < row_vecs; v0 += 32 * 2) {
      int4 buf[kGatherUnroll][2];
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int v = v0 + 32 * h;
          buf[u][h] = zero;
          if (src[u] >= 0 && v < row_vecs) buf[u][h] = ld_stream16(gout_b + (size_t)src[u] * row_vecs + v);
        }
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int v = v0 + 32 * h;
          if (t0 + u < T && v < row_vecs) st_stream16(gx_b + (size_t)(t0 + u) * row_vecs + v, buf[u][h]);
        }
    }
  }
}

// Scalar paths for rows that are not a multiple of 16 bytes (e.g. prev_decision with D == 1).
template <typename T_>
__global__ void gather_scalar_kernel(const T_* __restrict__ x, const int64_t* __restrict__ idx, T_* __restrict__ out,
                                     int B, int T, int D, int K, int rows_out, int prepend_cls) {
  const size_t total = (size_t)B * rows_out * D;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const size_t br = i / D;
    const int r = (int)(br % rows_out);
    const int b = (int)(br / rows_out);
    int s;
    if (prepend_cls) s = (r == 0) ? 0 : (int)idx[(size_t)b * K + r - 1] + 1;
    else             s = (int)idx[(size_t)b * K + r];
    s = min(max(s, 0), T - 1);
    out[i] = x[((size_t)b * T + s) * D + d];
  }
}

template <typename T_>
__global__ void scatter_scalar_kernel(const T_* __restrict__ gout, const int64_t* __restrict__ idx, T_* __restrict__ gx,
                                      int B, int T, int D, int K, int rows_out, int prepend_cls) {
  // gx is zero-filled by the caller-side memset issued before this kernel
  const size_t total = (size_t)B * rows_out * D;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const size_t br = i / D;
    const int r = (int)(br % rows_out);
    const int b = (int)(br / rows_out);
    int s;
    if (prepend_cls) s = (r == 0) ? 0 : (int)idx[(size_t)b * K + r - 1] + 1;
    else             s = (int)idx[(size_t)b * K + r];
    if (s >= 0 && s < T) gx[((size_t)b * T + s) * D + d] = gout[i];
  }
}

// splits per image so that the grid is at least ~2 waves of 148 SMs x 2 CTAs when B is small
static int splits_for(int B, int rows) {
  const int max_useful = ceil_div(rows, kGatherWarps * kGatherUnroll);
  int s = ceil_div(4 * kNumSMs, B);
  s = s < 1 ? 1 : s;
  return s > max_useful ? max_useful : s;
}

static int check_common(const void* a, const void* b, const int64_t* idx, int dtype, int B, int T, int D, int K,
                        int prepend_cls) {
  D2S_REQUIRE(a && b && (idx || K == 0), D2S_ERR_ARG, "gather/scatter: null pointer");
  D2S_REQUIRE(dtype == D2S_F32 || dtype == D2S_BF16, D2S_ERR_ARG, "gather/scatter: dtype %d unsupported", dtype);
  D2S_REQUIRE(B >= 0 && T >= 1 && D >= 1 && K >= 0, D2S_ERR_ARG, "gather/scatter: bad shape B=%d T=%d D=%d K=%d", B, T, D, K);
  D2S_REQUIRE(B <= 65535, D2S_ERR_ARG, "gather/scatter: B=%d exceeds 65535", B);
  D2S_REQUIRE(K + (prepend_cls ? 1 : 0) <= (prepend_cls ? T : INT32_MAX), D2S_ERR_ARG,
              "gather/scatter: K=%d too large for T=%d", K, T);
  return D2S_OK;
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_gather_tokens(const void* x, int dtype, int B, int T, int D, const int64_t* idx, int K,
                                 int prepend_cls, void* out, d2s_stream_t stream_) {
  int rc = check_common(x, out, idx, dtype, B, T, D, K, prepend_cls);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rows_out = K + (prepend_cls ? 1 : 0);
  if (B == 0 || rows_out == 0) return D2S_OK;
  const size_t row_bytes = (size_t)D * elem_size(dtype);
  if (row_bytes % 16 == 0) {
    D2S_REQUIRE(aligned16(x) && aligned16(out), D2S_ERR_ALIGN, "gather: x/out must be 16-byte aligned");
    dim3 grid(splits_for(B, rows_out), B);
    gather_rows_kernel<<<grid, kGatherThreads, 0, stream>>>((const int4*)x, idx, (int4*)out, T, K, rows_out,
                                                            (int)(row_bytes / 16), prepend_cls);
  } else {
    const size_t total = (size_t)B * rows_out * D;
    const int blocks = (int)((total + 255) / 256 < (size_t)(8 * kNumSMs) ? (total + 255) / 256 : 8 * kNumSMs);
    if (dtype == D2S_F32)
      gather_scalar_kernel<float><<<blocks, 256, 0, stream>>>((const float*)x, idx, (float*)out, B, T, D, K, rows_out, prepend_cls);
    else
      gather_scalar_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)x, idx, (__nv_bfloat16*)out, B, T, D, K, rows_out, prepend_cls);
  }
  count_launch();
  return check_launch("d2s_gather_tokens");
}

extern "C" int d2s_scatter_tokens_bwd(const void* gout, int dtype, int B, int T, int D, const int64_t* idx, int K,
                                      int prepend_cls, void* gx, d2s_stream_t stream_) {
  int rc = check_common(gout, gx, idx, dtype, B, T, D, K, prepend_cls);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rows_out = K + (prepend_cls ? 1 : 0);
  if (B == 0) return D2S_OK;
  const size_t row_bytes = (size_t)D * elem_size(dtype);
  if (row_bytes % 16 == 0) {
    D2S_REQUIRE(aligned16(gout) && aligned16(gx), D2S_ERR_ALIGN, "scatter: gout/gx must be 16-byte aligned");
    D2S_REQUIRE(T <= 8192, D2S_ERR_ARG, "scatter: T=%d exceeds 8192", T);
    dim3 grid(splits_for(B, T), B);
    scatter_rows_kernel<<<grid, kGatherThreads, T * sizeof(int), stream>>>((const int4*)gout, idx, (int4*)gx, T, K,
                                                                           rows_out, (int)(row_bytes / 16), prepend_cls);
    count_launch();
  } else {
    cudaMemsetAsync(gx, 0, (size_t)B * T * row_bytes, stream);
    if (rows_out > 0) {
      const size_t total = (size_t)B * rows_out * D;
      const int blocks = (int)((total + 255) / 256 < (size_t)(8 * kNumSMs) ? (total + 255) / 256 : 8 * kNumSMs);
      if (dtype == D2S_F32)
        scatter_scalar_kernel<float><<<blocks, 256, 0, stream>>>((const float*)gout, idx, (float*)gx, B, T, D, K, rows_out, prepend_cls);
      else
        scatter_scalar_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)gout, idx, (__nv_bfloat16*)gx, B, T, D, K, rows_out, prepend_cls);
      count_launch();
    }
  }
  return check_launch("d2s_scatter_tokens_bwd");
}
