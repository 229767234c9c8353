// Kernel family (2): PerturbedTopK forward/backward without the (b, nS, k, d) one-hot tensor
// (reference: vit_models/peturbed_topk.py:16-80 materialises 38 MB f32 + 77 MB i64 per image).
//
// Forward.  grid = (G, B), a thread-block cluster of G CTAs per image splits the S samples.
//   phase 1 (warp per sample): perturbed = x + sigma*noise (two roundings, as torch), order-preserving
//     32-bit keys, K-th largest key by MSB-first radix descent with warp-wide counts (early exit as soon as
//     a threshold separates exactly K keys), kept bit-masks by ballot;
//   phase 2 (thread per token): token i kept at sorted position j = popc(mask below i) updates
//     cnt[j][i-j] += 1 and eg[j][i-j] += noise_i.  Cells are indexed by (j, i-j): only the band
//     0 <= i-j <= N-K can be non-zero, which halves the accumulator (58 KB for N=196, K=98) and makes
//     every cell private to thread i -- no atomics.
//   epilogue: the cluster reduces the G partial bands through distributed shared memory and writes
//     indicators = cnt / S (integer counts -> bit-exact) and egrad = eg / S / sigma (the backward's
//     expected-gradient tensor, peturbed_topk.py:77-78) densely, zeros outside the band.
// Backward is then grad_x[b,i] = sum_j gout[b,j,i] * egrad[b,j,i]  (peturbed_topk.py:79).
//
// Algorithmic bytes per image: fwd 4N + 4SN + 4KN (+4KN stash); bwd 8KN + 4N.
#include <cooperative_groups.h>
#include "d2s_common.cuh"

namespace cg = cooperative_groups;

namespace d2s {

constexpr int kPtThreads = 256;
constexpr int kPtWarps = kPtThreads / 32;
constexpr int kPtSlots = 7;            // tokens per lane: N <= 224
constexpr int kPtMaxN = 32 * kPtSlots;

struct PtBatchBuf {
  float noise[kPtWarps][kPtMaxN];
  uint32_t mask[kPtWarps][8];
  uint32_t pre[kPtWarps][8];
};

// ---- Philox4x32-10 + Box-Muller for the in-kernel RNG contract -----------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  n0 = r * cs;
  n1 = r * sn;
}

template <bool kRng>
__global__ void __launch_bounds__(kPtThreads)
ptopk_fwd_kernel(const float* __restrict__ x, const float* __restrict__ noise, uint64_t seed, int N, int K, int S,
                 float sigma, float* __restrict__ indicators, float* __restrict__ egrad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int G = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int Wd = N - K + 1;
  const int cells = K * Wd;
  float* eg = reinterpret_cast<float*>(smem_raw);                                  // cells
  uint16_t* cnt = reinterpret_cast<uint16_t*>(eg + cells);                          // cells
  PtBatchBuf* buf = reinterpret_cast<PtBatchBuf*>(smem_raw + (((size_t)cells * 6 + 15) & ~(size_t)15));  // 2 buffers

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < cells; c += kPtThreads) { eg[c] = 0.f; cnt[c] = 0; }

  // this CTA's slice of the samples
  const int per = (S + G - 1) / G;
  const int s_begin = rank * per;
  const int s_end = min(S, s_begin + per);

  float xv[kPtSlots];
#pragma unroll
  for (int e = 0; e < kPtSlots; ++e) {
    const int i = lane + 32 * e;
    xv[e] = (i < N) ? x[(size_t)b * N + i] : 0.f;
  }
  __syncthreads();

  auto load_noise = [&](int s, float (&nz)[kPtSlots]) {
    if (kRng) {
      uint32_t r[4];
      const uint32_t sg = (uint32_t)((size_t)b * S + s);
      philox4x32_10(sg, (uint32_t)lane, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
      box_muller(r[0], r[1], nz[0], nz[1]);
      box_muller(r[2], r[3], nz[2], nz[3]);
      philox4x32_10(sg, (uint32_t)lane, 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
      box_muller(r[0], r[1], nz[4], nz[5]);
      float unused;
      box_muller(r[2], r[3], nz[6], unused);
    } else {
      const float* row = noise + ((size_t)b * S + s) * N;
#pragma unroll
      for (int e = 0; e < kPtSlots; ++e) {
        const int i = lane + 32 * e;
        nz[e] = (i < N) ? __ldcs(row + i) : 0.f;
      }
    }
  };

  float nz_next[kPtSlots];
  if (s_begin + warp < s_end) load_noise(s_begin + warp, nz_next);

  int it = 0;
  for (int base = s_begin; base < s_end; base += kPtWarps, ++it) {
    PtBatchBuf& bb = buf[it & 1];
    const int s = base + warp;
    if (s < s_end) {
      float nz[kPtSlots];
#pragma unroll
      for (int e = 0; e < kPtSlots; ++e) nz[e] = nz_next[e];
      if (s + kPtWarps < s_end) load_noise(s + kPtWarps, nz_next);  // prefetch the next row of this warp

      uint32_t key[kPtSlots];
#pragma unroll
      for (int e = 0; e < kPtSlots; ++e) {
        const int i = lane + 32 * e;
        const float p = __fadd_rn(xv[e], __fmul_rn(nz[e], sigma));  // x + noise*sigma, no FMA contraction
        key[e] = (i < N) ? float_to_ordered(p) : 0u;                // 0 sorts below every real key
      }
      // K-th largest key: greedy MSB-first threshold, stop early when exactly K keys are >= threshold
      uint32_t thr = 0u;
      bool exact = false;
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = thr | (1u << bit);
        int c = 0;
#pragma unroll
        for (int e = 0; e < kPtSlots; ++e) c += (key[e] >= cand) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= K) {
          thr = cand;
          if (c == K) { exact = true; break; }
        }
      }
      uint32_t m[kPtSlots];
      if (exact) {
#pragma unroll
        for (int e = 0; e < kPtSlots; ++e) m[e] = __ballot_sync(0xffffffffu, key[e] >= thr);
      } else {
        // thr is the K-th largest key and it is tied: keep all greater keys, then ties by lowest index
        int ngt = 0;
        uint32_t gtm[kPtSlots], eqm[kPtSlots];
#pragma unroll
        for (int e = 0; e < kPtSlots; ++e) {
          gtm[e] = __ballot_sync(0xffffffffu, key[e] > thr);
          eqm[e] = __ballot_sync(0xffffffffu, key[e] == thr);
          ngt += __popc(gtm[e]);
        }
        int need = K - ngt, seen = 0;
#pragma unroll
        for (int e = 0; e < kPtSlots; ++e) {
          const int my = seen + __popc(eqm[e] & ((1u << lane) - 1u));
          const bool take = ((eqm[e] >> lane) & 1u) && (my < need);
          m[e] = gtm[e] | __ballot_sync(0xffffffffu, take);
          seen += __popc(eqm[e]);
        }
      }
      int run = 0;
#pragma unroll
      for (int e = 0; e < kPtSlots; ++e) {
        if (lane == e) { bb.mask[warp][e] = m[e]; bb.pre[warp][e] = (uint32_t)run; }
        run += __popc(m[e]);
        const int i = lane + 32 * e;
        if (i < N) bb.noise[warp][i] = nz[e];
      }
    }
    __syncthreads();
    // phase 2: thread == token.  Cells (j, i-j) with this i are touched by this thread only.
    if (tid < N) {
      const int i = tid, e = i >> 5, bit = i & 31;
      const int nvalid = min(kPtWarps, s_end - base);
      for (int w = 0; w < nvalid; ++w) {
        const uint32_t word = bb.mask[w][e];
        if ((word >> bit) & 1u) {
          const int j = (int)bb.pre[w][e] + __popc(word & ((1u << bit) - 1u));
          const int cell = j * Wd + (i - j);
          cnt[cell] = (uint16_t)(cnt[cell] + 1);
          eg[cell] += bb.noise[w][i];
        }
      }
    }
    // no second barrier: the next batch writes the other buffer (see header comment)
  }
  // ---- cluster reduction + dense write-out ------------------------------------------------------
  cluster.sync();
  const float invS_div = (float)S;
  for (int j = rank * kPtWarps + warp; j < K; j += G * kPtWarps) {
    float* ind_row = indicators + ((size_t)b * K + j) * N;
    float* eg_row = egrad ? egrad + ((size_t)b * K + j) * N : nullptr;
    for (int i = lane; i < N; i += 32) {
      const int w = i - j;
      float cv = 0.f, ev = 0.f;
      if (w >= 0 && w < Wd) {
        const int cell = j * Wd + w;
        unsigned csum = 0;
        for (int g = 0; g < G; ++g) {
          csum += cluster.map_shared_rank(cnt, g)[cell];
          ev += cluster.map_shared_rank(eg, g)[cell];
        }
        cv = (float)csum / invS_div;
        ev = ev / invS_div / sigma;
      }
      ind_row[i] = cv;
      if (eg_row) eg_row[i] = ev;
    }
  }
  cluster.sync();  // keep shared memory alive until every peer has finished reading it
}

constexpr int kBwdCols = 64, kBwdGroups = 4;
__global__ void __launch_bounds__(kBwdCols * kBwdGroups)
ptopk_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ egrad, int N, int K, float* __restrict__ gx) {
  __shared__ float part[kBwdGroups][kBwdCols];
  const int b = blockIdx.y;
  const int cx = threadIdx.x % kBwdCols, gy = threadIdx.x / kBwdCols;
  const int i = blockIdx.x * kBwdCols + cx;
  float acc = 0.f;
  if (i < N) {
    const float* g = gout + (size_t)b * K * N + i;
    const float* e = egrad + (size_t)b * K * N + i;
    for (int j = gy; j < K; j += kBwdGroups) acc = fmaf(__ldcs(g + (size_t)j * N), __ldcs(e + (size_t)j * N), acc);
  }
  part[gy][cx] = acc;
  __syncthreads();
  if (gy == 0 && i < N) gx[(size_t)b * N + i] = (part[0][cx] + part[1][cx]) + (part[2][cx] + part[3][cx]);
}

static int ptopk_launch(bool rng, const float* x, const float* noise, uint64_t seed, int B, int N, int K, int S, float sigma,
                        float* indicators, float* egrad, cudaStream_t stream) {
  D2S_REQUIRE(x && indicators && (rng || noise), D2S_ERR_ARG, "ptopk_fwd: null pointer");
  D2S_REQUIRE(N >= 1 && N <= kPtMaxN, D2S_ERR_ARG, "ptopk_fwd: N=%d outside [1,%d]", N, kPtMaxN);
  D2S_REQUIRE(K >= 1 && K <= N, D2S_ERR_ARG, "ptopk_fwd: K=%d outside [1,N=%d]", K, N);
  D2S_REQUIRE(S >= 1 && S <= 65535, D2S_ERR_ARG, "ptopk_fwd: S=%d outside [1,65535]", S);
  D2S_REQUIRE(B >= 0 && B <= 65535, D2S_ERR_ARG, "ptopk_fwd: B=%d outside [0,65535]", B);
  D2S_REQUIRE(sigma != 0.0f, D2S_ERR_ARG, "ptopk_fwd: sigma must be non-zero (the reference divides by it)");
  if (B == 0) return D2S_OK;
  const size_t cells = (size_t)K * (N - K + 1);
  const size_t smem = ((cells * 6 + 15) & ~(size_t)15) + 2 * sizeof(PtBatchBuf);
  D2S_REQUIRE(smem <= 227 * 1024, D2S_ERR_ARG, "ptopk_fwd: N=%d K=%d needs %zu B of shared memory", N, K, smem);
  // Cluster size G (CTAs per image, splitting the samples): per-CTA work ~ 1/G, the grid runs in ceil(B*G / slots)
  // waves of `slots` resident CTAs, so the time ~ waves / G; pick the power of two that minimises it (ties -> smaller G,
  // less DSMEM reduction).  B=256: G=2 would run 512 CTAs in two waves of 444 slots (time 1.0), G=8 runs five waves of
  // eighth-size CTAs (0.625).
  const int per_sm = (int)((227u * 1024u) / (smem + 1024u)) < 1 ? 1 : (int)((227u * 1024u) / (smem + 1024u));
  const long long slots = (long long)kNumSMs * per_sm;
  int G = 1;
  double best = 1e30;
  for (int g = 1; g <= 8; g *= 2) {
    if (g > 1 && g * kPtWarps > S) break;  // never more CTAs than sample batches
    const long long waves = ((long long)B * g + slots - 1) / slots;
    const double t = (double)waves / g + 0.01 * g;
    if (t < best - 1e-9) { best = t; G = g; }
  }
  auto kern = rng ? ptopk_fwd_kernel<true> : ptopk_fwd_kernel<false>;
  static SmemOptIn opt[2];
  cudaError_t e = opt_in_smem(opt[rng ? 1 : 0], kern, 227 * 1024);
  D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "ptopk_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G, B);
  cfg.blockDim = dim3(kPtThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, x, noise, seed, N, K, S, sigma, indicators, egrad);
  D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "ptopk_fwd: launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("d2s_ptopk_fwd");
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_ptopk_fwd(const float* x, const float* noise, int B, int N, int K, int S, float sigma,
                             float* indicators, float* egrad, d2s_stream_t stream) {
  return ptopk_launch(false, x, noise, 0, B, N, K, S, sigma, indicators, egrad, (cudaStream_t)stream);
}

extern "C" int d2s_ptopk_fwd_rng(const float* x, uint64_t seed, int B, int N, int K, int S, float sigma,
                                 float* indicators, float* egrad, d2s_stream_t stream) {
  return ptopk_launch(true, x, nullptr, seed, B, N, K, S, sigma, indicators, egrad, (cudaStream_t)stream);
}

extern "C" int d2s_ptopk_bwd(const float* gout, const float* egrad, int B, int N, int K, float* gx, d2s_stream_t stream) {
  D2S_REQUIRE(gout && egrad && gx, D2S_ERR_ARG, "ptopk_bwd: null pointer");
  D2S_REQUIRE(N >= 1 && K >= 1 && B >= 0 && B <= 65535, D2S_ERR_ARG, "ptopk_bwd: bad shape B=%d N=%d K=%d", B, N, K);
  if (B == 0) return D2S_OK;
  dim3 grid(ceil_div(N, kBwdCols), B);
  ptopk_bwd_kernel<<<grid, kBwdCols * kBwdGroups, 0, (cudaStream_t)stream>>>(gout, egrad, N, K, gx);
  count_launch();
  return check_launch("d2s_ptopk_bwd");
}
