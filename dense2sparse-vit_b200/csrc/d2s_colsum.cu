// Column sums of a bf16 matrix: out[n] = sum_m dy[m, n] in fp32 -- the bias gradient of a Linear layer (training path).
// torch.autograd's reduction runs at ~2.9 TB/s on these shapes and cuBLASLt's bias-gradient epilogue kernel at ~1.9 TB/s; this
// is a plain HBM stream: a thread owns 8 consecutive columns (one 16-byte vector per row), blockDim.y row slots walk the CTA's
// row block with four rows in flight per thread, partial sums meet in shared memory and leave as one atomicAdd per column per CTA.
#include "d2s_common.cuh"

namespace d2s {

__global__ void __launch_bounds__(1024)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dy, long long M, int N, int rows_per_cta, float* __restrict__ out) {
  extern __shared__ float cs_red[];   // blockDim.y x N
  const int tx = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const long long row_end = min(M, row0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  auto add = [&](const int4& v) {
    const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      acc[2 * q] += __uint_as_float(w[q] << 16);
      acc[2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
    }
  };
  const __nv_bfloat16* base = dy + (size_t)tx * 8;
  long long r = row0 + ty;
  for (; r + 3 * ny < row_end; r += 4 * ny) {
    const int4 a = ld_stream16(base + (size_t)r * N);
    const int4 b = ld_stream16(base + (size_t)(r + ny) * N);
    const int4 c = ld_stream16(base + (size_t)(r + 2 * ny) * N);
    const int4 d = ld_stream16(base + (size_t)(r + 3 * ny) * N);
    add(a); add(b); add(c); add(d);
  }
  for (; r < row_end; r += ny) add(ld_stream16(base + (size_t)r * N));
  float* mine = cs_red + (size_t)ty * N + tx * 8;
  *reinterpret_cast<float4*>(mine) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(mine + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  __syncthreads();
  for (int c = ty * blockDim.x + tx; c < N; c += ny * blockDim.x) {
    float s = 0.f;
    for (int y = 0; y < ny; ++y) s += cs_red[(size_t)y * N + c];
    atomicAdd(out + c, s);
  }
}

}  // namespace d2s

using namespace d2s;

static int colsum_entry(const void* dy, long long M, int N, float* out, bool zero_first, d2s_stream_t stream) {
  D2S_REQUIRE(dy && out, D2S_ERR_ARG, "colsum: null pointer");
  D2S_REQUIRE(M >= 0 && N >= 8 && N % 8 == 0 && N <= 8192, D2S_ERR_ARG, "colsum: bad shape M=%lld N=%d (N %% 8 == 0, N <= 8192)", M, N);
  D2S_REQUIRE(aligned16(dy), D2S_ERR_ALIGN, "colsum: dy must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_first) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st);
    D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "colsum: memset: %s", cudaGetErrorString(e));
  }
  if (M == 0) return D2S_OK;
  const int nx = N / 8;
  int ny = 1024 / nx;                      // row slots per CTA
  if (ny > 8) ny = 8;
  if (ny < 1) ny = 1;
  const size_t smem = (size_t)ny * N * sizeof(float);
  D2S_REQUIRE(smem <= 48 * 1024, D2S_ERR_ARG, "colsum: N=%d too wide", N);
  // ~4 CTAs per SM worth of row blocks, each at least one full unrolled pass
  long long rows_per_cta = (M + 4LL * kNumSMs - 1) / (4LL * kNumSMs);
  if (rows_per_cta < 4LL * ny) rows_per_cta = 4LL * ny;
  const long long grid = (M + rows_per_cta - 1) / rows_per_cta;
  colsum_bf16_kernel<<<(unsigned)grid, dim3(nx, ny), smem, st>>>((const __nv_bfloat16*)dy, M, N, (int)rows_per_cta, out);
  count_launch();
  return check_launch("d2s_colsum_bf16");
}

extern "C" int d2s_colsum_bf16(const void* dy, long long M, int N, float* out, d2s_stream_t stream) {
  return colsum_entry(dy, M, N, out, true, stream);
}

// out += column sums: the form that lands a bias gradient straight in the parameter's (zeroed once per step) gradient slot
extern "C" int d2s_colsum_acc_bf16(const void* dy, long long M, int N, float* out, d2s_stream_t stream) {
  return colsum_entry(dy, M, N, out, false, stream);
}

// GELU backward fused with the bias gradient of the Linear in front of it (Mlp.forward, dynamic_vit.py:170-172: fc1 -> GELU):
//     du = ga * gelu'(u)   (bf16)        db[n] = sum_m du[m, n]   (fp32, of the rounded values the dW GEMM consumes)
// gelu'(x) = Phi(x) + x phi(x).  With a = |x| and E = exp(-a^2/2):  Phi(-a) = 0.5 erfcx(a / sqrt 2) E  (the degree-8 polynomial of
// gelu_erf_pair, d2s_tc.cuh) and phi(x) = E / sqrt(2 pi) share the one MUFU.EX2.  Same thread layout as colsum_bf16_kernel.
#include "d2s_tc.cuh"

namespace d2s {

// two elements at a time on the packed fp32x2 pipe (the scalar form made the kernel co-bound by FP32 issue)
__device__ __forceinline__ uint64_t gelu_grad_pair(uint64_t x) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  const uint64_t a = f2_pack(fminf(fabsf(x0), 5.6568542f), fminf(fabsf(x1), 5.6568542f));
  uint64_t r = f2_fma(f2_bcast(-3.457075050e-06f), a, f2_bcast(9.698495899e-05f));     // r = -0.5 erfcx(a / sqrt 2)
  r = f2_fma(r, a, f2_bcast(-1.174711513e-03f));
  r = f2_fma(r, a, f2_bcast(8.114228228e-03f));
  r = f2_fma(r, a, f2_bcast(-3.579151344e-02f));
  r = f2_fma(r, a, f2_bcast(1.080985674e-01f));
  r = f2_fma(r, a, f2_bcast(-2.371637582e-01f));
  r = f2_fma(r, a, f2_bcast(3.961593576e-01f));
  r = f2_fma(r, a, f2_bcast(-4.998897713e-01f));
  float e0, e1;
  f2_unpack(f2_mul(f2_mul(x, x), f2_bcast(-0.72134752044448170368f)), e0, e1);
  const uint64_t e = f2_pack(ex2_approx(e0), ex2_approx(e1));                            // exp(-x^2 / 2)
  float nh0, nh1;
  f2_unpack(f2_mul(r, e), nh0, nh1);                                                     // -Phi(-|x|)
  const uint64_t cdf = f2_pack(x0 >= 0.f ? 1.0f + nh0 : -nh0, x1 >= 0.f ? 1.0f + nh1 : -nh1);
  return f2_fma(f2_mul(x, f2_bcast(0.3989422804014327f)), e, cdf);
}

__global__ void __launch_bounds__(1024)
gelu_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ ga, long long M, int N,
                       int rows_per_cta, __nv_bfloat16* __restrict__ du, float* __restrict__ db) {
  extern __shared__ float gb_red[];   // blockDim.y x N
  const int tx = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const long long row_end = min(M, row0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  auto one = [&](long long r, const int4& uv, const int4& gv) {
    const uint32_t uw[4] = {(uint32_t)uv.x, (uint32_t)uv.y, (uint32_t)uv.z, (uint32_t)uv.w};
    const uint32_t gw[4] = {(uint32_t)gv.x, (uint32_t)gv.y, (uint32_t)gv.z, (uint32_t)gv.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float d0, d1;
      f2_unpack(f2_mul(f2_pack(__uint_as_float(gw[q] << 16), __uint_as_float(gw[q] & 0xffff0000u)),
                       gelu_grad_pair(f2_pack(__uint_as_float(uw[q] << 16), __uint_as_float(uw[q] & 0xffff0000u)))), d0, d1);
      const __nv_bfloat162 t = __floats2bfloat162_rn(d0, d1);
      o[q] = *reinterpret_cast<const uint32_t*>(&t);
      acc[2 * q] += __uint_as_float(o[q] << 16);
      acc[2 * q + 1] += __uint_as_float(o[q] & 0xffff0000u);
    }
    *reinterpret_cast<uint4*>(du + (size_t)r * N + (size_t)tx * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  const size_t col = (size_t)tx * 8;
  long long r = row0 + ty;
  for (; r + 3 * ny < row_end; r += 4 * ny) {       // four rows (eight 16-byte loads) in flight per thread
    const int4 u0 = ld_stream16(u + (size_t)r * N + col), g0 = ld_stream16(ga + (size_t)r * N + col);
    const int4 u1 = ld_stream16(u + (size_t)(r + ny) * N + col), g1 = ld_stream16(ga + (size_t)(r + ny) * N + col);
    const int4 u2 = ld_stream16(u + (size_t)(r + 2 * ny) * N + col), g2 = ld_stream16(ga + (size_t)(r + 2 * ny) * N + col);
    const int4 u3 = ld_stream16(u + (size_t)(r + 3 * ny) * N + col), g3 = ld_stream16(ga + (size_t)(r + 3 * ny) * N + col);
    one(r, u0, g0);
    one(r + ny, u1, g1);
    one(r + 2 * ny, u2, g2);
    one(r + 3 * ny, u3, g3);
  }
  for (; r < row_end; r += ny) one(r, ld_stream16(u + (size_t)r * N + col), ld_stream16(ga + (size_t)r * N + col));
  if (db == nullptr) return;
  float* mine = gb_red + (size_t)ty * N + tx * 8;
  *reinterpret_cast<float4*>(mine) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(mine + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  __syncthreads();
  for (int c = ty * blockDim.x + tx; c < N; c += ny * blockDim.x) {
    float s = 0.f;
    for (int y = 0; y < ny; ++y) s += gb_red[(size_t)y * N + c];
    atomicAdd(db + c, s);
  }
}

}  // namespace d2s

static int gelu_bwd_colsum_entry(const void* u, const void* ga, long long M, int N, void* du, float* db, bool zero_first,
                                 d2s_stream_t stream) {
  D2S_REQUIRE(u && ga && du, D2S_ERR_ARG, "gelu_bwd_colsum: null pointer");
  D2S_REQUIRE(M >= 0 && N >= 8 && N % 8 == 0 && N <= 8192, D2S_ERR_ARG, "gelu_bwd_colsum: bad shape M=%lld N=%d (N %% 8 == 0, N <= 8192)",
              M, N);
  D2S_REQUIRE(aligned16(u) && aligned16(ga) && aligned16(du), D2S_ERR_ALIGN, "gelu_bwd_colsum: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (db && zero_first) {
    cudaError_t e = cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st);
    D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "gelu_bwd_colsum: memset: %s", cudaGetErrorString(e));
  }
  if (M == 0) return D2S_OK;
  const int nx = N / 8;
  int ny = 1024 / nx;
  if (ny > 8) ny = 8;
  if (ny < 1) ny = 1;
  const size_t smem = (size_t)ny * N * sizeof(float);
  D2S_REQUIRE(smem <= 48 * 1024, D2S_ERR_ARG, "gelu_bwd_colsum: N=%d too wide", N);
  long long rows_per_cta = (M + 4LL * d2s::kNumSMs - 1) / (4LL * d2s::kNumSMs);
  if (rows_per_cta < 4LL * ny) rows_per_cta = 4LL * ny;
  const long long grid = (M + rows_per_cta - 1) / rows_per_cta;
  d2s::gelu_bwd_colsum_kernel<<<(unsigned)grid, dim3(nx, ny), smem, st>>>((const __nv_bfloat16*)u, (const __nv_bfloat16*)ga, M, N,
                                                                         (int)rows_per_cta, (__nv_bfloat16*)du, db);
  d2s::count_launch();
  return d2s::check_launch("d2s_gelu_bwd_colsum_bf16");
}

extern "C" int d2s_gelu_bwd_colsum_bf16(const void* u, const void* ga, long long M, int N, void* du, float* db, d2s_stream_t stream) {
  return gelu_bwd_colsum_entry(u, ga, M, N, du, db, true, stream);
}

extern "C" int d2s_gelu_bwd_colsum_acc_bf16(const void* u, const void* ga, long long M, int N, void* du, float* db, d2s_stream_t stream) {
  return gelu_bwd_colsum_entry(u, ga, M, N, du, db, false, stream);
}
