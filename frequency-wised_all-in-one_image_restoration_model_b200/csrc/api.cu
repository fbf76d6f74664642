// Library-level entry points: version, error string, launch counter, per-class in-stream timing.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "freqair_internal.h"

static thread_local char g_err[512] = "";
// touched from the forward thread and from autograd's backward thread: atomics + one mutex around the event list
static std::atomic<int> g_prof_cls{0};
static std::mutex g_prof_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static std::atomic<int64_t> g_launches{0};

void fa_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void fa_count_launch(int) { g_launches.fetch_add(1, std::memory_order_relaxed); }

FaProfScope::FaProfScope(int cls_, cudaStream_t st_) : cls(cls_), st(st_), e0(nullptr), on(false) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const int want = g_prof_cls.load(std::memory_order_relaxed);
  if (want != 0 && want == cls) {
    on = true;
    cudaEventCreate(&e0);
    cudaEventRecord(e0, st);
  }
}
FaProfScope::~FaProfScope() {
  if (on) {
    cudaEvent_t e1;
    cudaEventCreate(&e1);
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_events.emplace_back(e0, e1);
  }
}

extern "C" {

const char* fa_version(void) { return "freqair 0.2 (sm_100a)"; }
int fa_abi_version(void) { return FREQAIR_ABI_VERSION; }
const char* fa_last_error_string(void) { return g_err; }

int fa_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  FA_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  FA_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return FA_OK;
}

int fa_prof_begin(int kernel_class) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& pr : g_prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  g_prof_events.clear();
  g_prof_cls = kernel_class;
  return FA_OK;
}

int fa_prof_end(double* total_ms, int64_t* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double tot = 0.0;
  for (auto& pr : g_prof_events) {
    cudaEventSynchronize(pr.second);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, pr.first, pr.second);
    tot += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int64_t)g_prof_events.size();
  g_prof_events.clear();
  g_prof_cls = 0;
  return FA_OK;
}

int64_t fa_launch_count(void) { return g_launches.load(); }
void fa_launch_count_reset(void) { g_launches.store(0); }

int fa_gemm(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc,
            int transA, int transB, const FaGemmEpilogue* epi, int backend, fa_stream_t stream) {
  FA_REQUIRE(A && B && C, "fa_gemm: null operand");
  FA_REQUIRE(M >= 0 && N >= 0 && K >= 0, "fa_gemm: negative dimension");
  FA_REQUIRE(backend >= 0 && backend <= 5, "fa_gemm: backend must be 0..5");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_GEMM, st);
  if (backend != 1) {
    // MMAs per product: 3 (backends 0, 2), 0 = raw truncated operands (3), 2 = A exact / B RN-rounded (4), 1 = RN-TF32 (5)
    const int passes = backend == 3 ? 0 : (backend == 4 ? 2 : (backend == 5 ? 1 : 3));
    int rc = fa_gemm_tc_launch(A, B, C, M, N, K, lda, ldb, ldc, transA, transB, epi, st, passes);
    if (rc != FA_ERR_UNSUPPORTED) return rc;
    if (backend == 2 || backend == 3) {
      fa_set_error("fa_gemm: shape M=%d N=%d K=%d tA=%d tB=%d not eligible for the tcgen05 path", M, N, K, transA, transB);
      return FA_ERR_UNSUPPORTED;
    }
  }
  if (epi && epi->a_kscale) {
    fa_set_error("fa_gemm: a_kscale needs the tcgen05 path (M=%d N=%d K=%d tA=%d tB=%d, a_k_rows_per_scale=%d)", M, N, K,
                 transA, transB, epi->a_k_rows_per_scale);
    return FA_ERR_UNSUPPORTED;
  }
  FaGemmEpilogue e2;
  if (epi && epi->a_rowsum) {
    // the SIMT kernel has no fused row-sum: take it in a separate pass over A, then contract without it
    int rc = fa_a_rowsum(A, epi->a_rowsum, M, K, lda, transA, stream);
    if (rc) return rc;
    e2 = *epi;
    e2.a_rowsum = nullptr;
    epi = &e2;
  }
  return fa_gemm_simt_launch(A, B, C, M, N, K, lda, ldb, ldc, transA, transB, epi, st);
}

static int conv_passes(int backend) { return backend == 4 ? 2 : (backend == 5 ? 1 : 3); }

int fa_conv3x3_gemm(const float* x, const float* wk, float* y, int B, int H, int W, int Cin, int Cout, int64_t ldy,
                    const FaGemmEpilogue* epi, int backend, fa_stream_t stream) {
  FA_REQUIRE(x && wk && y, "fa_conv3x3_gemm: null operand");
  FA_REQUIRE(backend == 0 || backend == 2 || backend == 4 || backend == 5, "fa_conv3x3_gemm: backend must be 0, 2, 4 or 5");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_GEMM, st);
  FaConvOperand cv{1, B, H, W, Cin};
  const int64_t T = (int64_t)B * H * W;
  FA_REQUIRE(T < (1ll << 31), "fa_conv3x3_gemm: too many tokens");
  int rc = fa_gemm_tc_launch(x, wk, y, (int)T, Cout, 9 * Cin, Cin, 9 * Cin, ldy, 0, 1, epi, st, conv_passes(backend), &cv);
  if (rc == FA_ERR_UNSUPPORTED)
    fa_set_error("fa_conv3x3_gemm: geometry B=%d H=%d W=%d Cin=%d Cout=%d not eligible (Cin %% 32, H*W %% 128, W | 128 or 128 | W)",
                 B, H, W, Cin, Cout);
  return rc;
}

int fa_conv3x3_wgrad(const float* g, int64_t ldg, const float* x, float* dwk, int B, int H, int W, int Cin, int Cout,
                     int accumulate, float* dbias, int backend, fa_stream_t stream) {
  FA_REQUIRE(g && x && dwk, "fa_conv3x3_wgrad: null operand");
  FA_REQUIRE(backend == 0 || backend == 2 || backend == 4 || backend == 5, "fa_conv3x3_wgrad: backend must be 0, 2, 4 or 5");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_GEMM, st);
  FaConvOperand cv{2, B, H, W, Cin};
  const int64_t T = (int64_t)B * H * W;
  FA_REQUIRE(T < (1ll << 31), "fa_conv3x3_wgrad: too many tokens");
  FaGemmEpilogue e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.0f;
  e.accumulate = accumulate;
  e.a_rowsum = dbias;                    // column sums of g = the bias gradient, from the same pass over g
  int rc = fa_gemm_tc_launch(g, x, dwk, Cout, 9 * Cin, (int)T, ldg, Cin, 9 * Cin, 1, 0, &e, st, conv_passes(backend), &cv);
  if (rc == FA_ERR_UNSUPPORTED)
    fa_set_error("fa_conv3x3_wgrad: geometry B=%d H=%d W=%d Cin=%d Cout=%d not eligible (Cin %% 32, W %% 32, H*W %% 128)", B, H, W,
                 Cin, Cout);
  return rc;
}

}  // extern "C"
