// Weight + bias gradient of a Linear layer in ONE library GEMM (training path, bf16):
//
//     dW (N, K) = dY^T X        db (N) = sum_m dY[m, :]
//
// torch.autograd runs the bias gradient as a separate column reduction over dY (61 `reduce_kernel` launches, 3.7 ms of a
// 34 ms DeiT-S training step, each re-reading a tensor the dW GEMM reads anyway).  cuBLASLt can fold it into the dW GEMM's
// epilogue (CUBLASLT_EPILOGUE_BGRADB: reduce the B operand over the GEMM's k dimension); torch does not expose that, so the
// call lives here, behind the C ABI.  The dense contractions of the gradient path are library GEMMs by design.
//
// cuBLASLt is bound at first use with dlopen/dlsym (like the driver entry point for tensor maps): the shared library carries
// no link-time dependency on it, and inside a PyTorch process the already-loaded libcublasLt.so.12 is the one that is found.
#include <cublasLt.h>
#include <dlfcn.h>

#include <mutex>
#include <unordered_map>

#include "d2s_common.cuh"

namespace d2s {

struct LtApi {
  decltype(&cublasLtCreate) create;
  decltype(&cublasLtMatmulDescCreate) desc_create;
  decltype(&cublasLtMatmulDescDestroy) desc_destroy;
  decltype(&cublasLtMatmulDescSetAttribute) desc_set;
  decltype(&cublasLtMatrixLayoutCreate) layout_create;
  decltype(&cublasLtMatrixLayoutDestroy) layout_destroy;
  decltype(&cublasLtMatmulPreferenceCreate) pref_create;
  decltype(&cublasLtMatmulPreferenceDestroy) pref_destroy;
  decltype(&cublasLtMatmulPreferenceSetAttribute) pref_set;
  decltype(&cublasLtMatmulAlgoGetHeuristic) heuristic;
  decltype(&cublasLtMatmul) matmul;
  bool ok;
};

static const LtApi* lt_api() {
  static LtApi api{};
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libcublasLt.so.12", "libcublasLt.so.13", "libcublasLt.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return;
#define D2S_LT_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, #sym)); if (!api.field) return;
    D2S_LT_SYM(create, cublasLtCreate)
    D2S_LT_SYM(desc_create, cublasLtMatmulDescCreate)
    D2S_LT_SYM(desc_destroy, cublasLtMatmulDescDestroy)
    D2S_LT_SYM(desc_set, cublasLtMatmulDescSetAttribute)
    D2S_LT_SYM(layout_create, cublasLtMatrixLayoutCreate)
    D2S_LT_SYM(layout_destroy, cublasLtMatrixLayoutDestroy)
    D2S_LT_SYM(pref_create, cublasLtMatmulPreferenceCreate)
    D2S_LT_SYM(pref_destroy, cublasLtMatmulPreferenceDestroy)
    D2S_LT_SYM(pref_set, cublasLtMatmulPreferenceSetAttribute)
    D2S_LT_SYM(heuristic, cublasLtMatmulAlgoGetHeuristic)
    D2S_LT_SYM(matmul, cublasLtMatmul)
#undef D2S_LT_SYM
    api.ok = true;
  });
  return api.ok ? &api : nullptr;
}

constexpr size_t kLtWorkspace = 32u << 20;

// one plan (descriptors + chosen algorithm) per problem shape; the handle and workspace belong to the process's device
struct WgradPlan {
  cublasLtMatmulDesc_t op;
  cublasLtMatrixLayout_t a, b, c;
  cublasLtMatmulAlgo_t algo;
};
struct LtState {
  std::mutex mu;
  cublasLtHandle_t handle = nullptr;
  void* workspace = nullptr;
  int device = -1;
  std::unordered_map<unsigned long long, WgradPlan> plans;   // key: M, N, K, want_bias
};
static LtState& lt_state() {
  static LtState s;
  return s;
}

}  // namespace d2s

using namespace d2s;

#define D2S_LT_CHECK(expr, what)                                                                          \
  do {                                                                                                    \
    cublasStatus_t st_ = (expr);                                                                          \
    D2S_REQUIRE(st_ == CUBLAS_STATUS_SUCCESS, D2S_ERR_CUDA, "linear_wgrad: %s failed (cublas status %d)", what, (int)st_); \
  } while (0)

extern "C" int d2s_linear_wgrad_bf16(const void* dy, const void* x, int M, int N, int K, void* dw, void* db, d2s_stream_t stream) {
  D2S_REQUIRE(dy && x && dw, D2S_ERR_ARG, "linear_wgrad: null pointer");
  D2S_REQUIRE(M >= 1 && N >= 8 && K >= 8 && N % 8 == 0 && K % 8 == 0, D2S_ERR_ARG,
              "linear_wgrad: bad shape M=%d N=%d K=%d (bf16: N and K multiples of 8)", M, N, K);
  D2S_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dw) && (!db || aligned16(db)), D2S_ERR_ALIGN,
              "linear_wgrad: pointers must be 16-byte aligned");
  const LtApi* lt = lt_api();
  D2S_REQUIRE(lt != nullptr, D2S_ERR_CUDA, "linear_wgrad: libcublasLt could not be loaded (dlopen/dlsym)");
  LtState& s = lt_state();
  std::lock_guard<std::mutex> lock(s.mu);
  int dev = 0;
  cudaGetDevice(&dev);
  if (!s.handle) {
    D2S_LT_CHECK(lt->create(&s.handle), "cublasLtCreate");
    cudaError_t e = cudaMalloc(&s.workspace, kLtWorkspace);
    D2S_REQUIRE(e == cudaSuccess, D2S_ERR_CUDA, "linear_wgrad: workspace allocation: %s", cudaGetErrorString(e));
    s.device = dev;
  }
  D2S_REQUIRE(dev == s.device, D2S_ERR_ARG, "linear_wgrad: one device per process (first used on %d, now %d)", s.device, dev);

  const unsigned long long key = ((unsigned long long)M << 33) ^ ((unsigned long long)N << 17) ^ ((unsigned long long)K << 1) ^ (db ? 1ull : 0ull);
  auto it = s.plans.find(key);
  if (it == s.plans.end()) {
    // Row-major (r, c) == column-major (c, r).  Column-major problem: dW^T (K x N) = X^T (K x M) . dY (M x N), i.e.
    // A = x viewed (K x M, ld K, no transpose), B = dy viewed (N x M, ld N, transposed), C = dw viewed (K x N, ld K).
    WgradPlan plan{};
    D2S_LT_CHECK(lt->desc_create(&plan.op, CUBLAS_COMPUTE_32F, CUDA_R_32F), "MatmulDescCreate");
    const cublasOperation_t ta = CUBLAS_OP_N, tb = CUBLAS_OP_T;
    D2S_LT_CHECK(lt->desc_set(plan.op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)), "set TRANSA");
    D2S_LT_CHECK(lt->desc_set(plan.op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)), "set TRANSB");
    if (db) {
      const cublasLtEpilogue_t ep = CUBLASLT_EPILOGUE_BGRADB;     // db[n] = sum over the GEMM's k (= M) of B
      D2S_LT_CHECK(lt->desc_set(plan.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep)), "set EPILOGUE");
      const cudaDataType_t bt = CUDA_R_16BF;
      D2S_LT_CHECK(lt->desc_set(plan.op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &bt, sizeof(bt)), "set BIAS_DATA_TYPE");
    }
    D2S_LT_CHECK(lt->layout_create(&plan.a, CUDA_R_16BF, K, M, K), "layout A");
    D2S_LT_CHECK(lt->layout_create(&plan.b, CUDA_R_16BF, N, M, N), "layout B");
    D2S_LT_CHECK(lt->layout_create(&plan.c, CUDA_R_16BF, K, N, K), "layout C");
    if (db) {   // the heuristic must see a bias pointer of the final alignment
      D2S_LT_CHECK(lt->desc_set(plan.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &db, sizeof(db)), "set BIAS_POINTER");
    }
    cublasLtMatmulPreference_t pref;
    D2S_LT_CHECK(lt->pref_create(&pref), "PreferenceCreate");
    const size_t ws = kLtWorkspace;
    D2S_LT_CHECK(lt->pref_set(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws, sizeof(ws)), "set MAX_WORKSPACE");
    cublasLtMatmulHeuristicResult_t res{};
    int found = 0;
    cublasStatus_t hs = lt->heuristic(s.handle, plan.op, plan.a, plan.b, plan.c, plan.c, pref, 1, &res, &found);
    lt->pref_destroy(pref);
    D2S_REQUIRE(hs == CUBLAS_STATUS_SUCCESS && found >= 1, D2S_ERR_CUDA,
                "linear_wgrad: no cuBLASLt algorithm for M=%d N=%d K=%d bias=%d (status %d)", M, N, K, db ? 1 : 0, (int)hs);
    plan.algo = res.algo;
    it = s.plans.emplace(key, plan).first;
  }
  WgradPlan& plan = it->second;
  if (db) D2S_LT_CHECK(lt->desc_set(plan.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &db, sizeof(db)), "set BIAS_POINTER");
  const float alpha = 1.0f, beta = 0.0f;
  D2S_LT_CHECK(lt->matmul(s.handle, plan.op, &alpha, x, plan.a, dy, plan.b, &beta, dw, plan.c, dw, plan.c, &plan.algo, s.workspace,
                          kLtWorkspace, (cudaStream_t)stream),
               "cublasLtMatmul");
  return D2S_OK;   // (a library kernel: not counted as a d2s launch)
}
