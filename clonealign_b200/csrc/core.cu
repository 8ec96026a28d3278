// clonealign_b200 core: device state + the C-ABI declared in include/clonealign_b200.h.
//
// One ca_handle == one TensorFlow session of the reference (R/inference-tflow.R:351-457) for one
// cell shard on one GPU.  No CPU fallback exists: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <map>
#include <mutex>
#include <tuple>
#include <stdexcept>
#include <type_traits>
#include <string>
#include <utility>
#include <vector>

#include "../../include/clonealign_b200.h"
#include "common.cuh"
#include "kernels_expgemm.cuh"
#include "kernels_fused.cuh"
#include "kernels_interp.cuh"
#include "kernels_p2p.cuh"
#include "kernels_pca.cuh"
#include "kernels_small.cuh"
#ifdef CA_EMULATE   // tests/cuda_emul: functional CPU emulation of the non-tensor kernels (test infrastructure only)
#include "kernels_tc_stub.h"
#else
#include "kernels_tc.cuh"
#endif
#include "kernels_ypass.cuh"

using namespace ca;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
namespace {

struct CaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw CaError(buf);
}

#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                                __LINE__, cudaGetErrorString(_e));                            \
  } while (0)

int report(const std::exception& e, char* err, size_t errlen) {
  if (err && errlen) {
    strncpy(err, e.what(), errlen - 1);
    err[errlen - 1] = 0;
  }
  return 1;
}

// ------------------------------------------------------------------------------------------------
// NCCL, resolved at run time (no link-time dependency; a single-GPU fit never touches it)
// ------------------------------------------------------------------------------------------------
struct Uid { char internal[128]; };   // ncclUniqueId (passed by value)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Uid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

#ifdef CA_EMULATE   // tests/cuda_emul: ranks are threads of one process, see nccl_emul.h
}  // namespace
#include "nccl_emul.h"
namespace {
NcclApi& nccl() {
  static NcclApi api;
  if (api.lib) return api;
  api.GetUniqueId = [](void* p) { return ca_emul_nccl::GetUniqueId(p); };
  api.CommInitRank = [](void** c, int w, Uid id, int r) { return ca_emul_nccl::CommInitRank(c, w, id, r); };
  api.AllReduce = [](const void* s, void* d, size_t n, int t, int o, void* c, cudaStream_t st) {
    return ca_emul_nccl::AllReduce(s, d, n, t, o, c, (void*)st);
  };
  api.CommDestroy = [](void* c) { return ca_emul_nccl::CommDestroy(c); };
  api.GetErrorString = [](int e) { return ca_emul_nccl::GetErrorString(e); };
  api.lib = (void*)&api;
  return api;
}
#else
NcclApi& nccl() {
  static NcclApi api;
  if (api.lib) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) fail("NCCL is required for world > 1 but libnccl.so.2 could not be loaded: %s", dlerror());
  auto sym = [&](const char* s) {
    void* p = dlsym(api.lib, s);
    if (!p) fail("NCCL symbol %s not found", s);
    return p;
  };
  api.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, Uid, int))sym("ncclCommInitRank");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
  api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
  api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  return api;
}
#endif
constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;
#define NCCL_OK(expr)                                                              \
  do {                                                                             \
    int _r = (expr);                                                               \
    if (_r != 0) fail("NCCL error %d at %s:%d: %s", _r, __FILE__, __LINE__, nccl().GetErrorString(_r)); \
  } while (0)

// Communicators are expensive to build (ncclCommInitRank: 0.3 - 1 s with 8 ranks) and a process usually runs several
// sessions of the same shape one after another (restarts, the set-up of a benchmark and its end-to-end run), so a
// released communicator is kept per (world, rank, device) and handed to the next session of that shape; every rank
// of a job creates and releases its sessions in the same order, so all ranks hit (or miss) the cache together.
// ca_core_shutdown() destroys what is parked.
struct CommKey {
  int world, rank, dev;
  bool operator<(const CommKey& o) const { return std::tie(world, rank, dev) < std::tie(o.world, o.rank, o.dev); }
};
std::mutex& comm_mu() { static std::mutex m; return m; }
std::map<CommKey, std::vector<void*>>& comm_pool() { static std::map<CommKey, std::vector<void*>> p; return p; }
void* comm_acquire(int world, int rank, int dev, const void* id128) {
  {
    std::lock_guard<std::mutex> lk(comm_mu());
    auto& v = comm_pool()[CommKey{world, rank, dev}];
    if (!v.empty()) { void* c = v.back(); v.pop_back(); return c; }
  }
  Uid id;
  memcpy(&id, id128, sizeof id);
  void* c = nullptr;
  NCCL_OK(nccl().CommInitRank(&c, world, id, rank));
  return c;
}
void comm_release(int world, int rank, int dev, void* c) {
  if (!c) return;
  std::lock_guard<std::mutex> lk(comm_mu());
  comm_pool()[CommKey{world, rank, dev}].push_back(c);
}

// ------------------------------------------------------------------------------------------------
// conversion kernels (ingest)
// ------------------------------------------------------------------------------------------------
template <typename Tin>
__global__ void k_ingest_colmajor(const Tin* __restrict__ in, int64_t ld_in, int64_t N, int g0, int gcount,
                                  float* __restrict__ out, int64_t ldY) {
  // in: column-major chunk, element (n, gg) at in[gg*ld_in + n]; out[n][g0+gg]
  __shared__ float tile[32][33];
  int64_t nb = (int64_t)blockIdx.x * 32;
  int gb = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int gg = gb + i;
    int64_t n = nb + threadIdx.x;
    tile[i][threadIdx.x] = (gg < gcount && n < N) ? (float)in[(int64_t)gg * ld_in + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t n = nb + i;
    int gg = gb + threadIdx.x;
    if (n < N && gg < gcount) out[n * ldY + g0 + gg] = tile[threadIdx.x][i];
  }
}
template <typename Tin>
__global__ void k_ingest_rowmajor(const Tin* __restrict__ in, int64_t ld_in, int64_t rows, int G,
                                  float* __restrict__ out, int64_t ldY) {
  int64_t r = blockIdx.y;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x)
    if (r < rows) out[r * ldY + g] = (float)in[r * ld_in + g];
}
// compressed sparse rows -> dense: one warp per cell scatters its stored values (Yf is zero-filled beforehand).
// idx / val hold the chunk's entries starting at offset `base`; *bad is set on an out-of-range gene index.
template <typename Tin>
__global__ void k_ingest_csr(const int* __restrict__ indptr, const int* __restrict__ idx, const Tin* __restrict__ val,
                             int64_t base, int64_t r0, int64_t rows, int G, float* __restrict__ Yf, int64_t ldY,
                             int* __restrict__ bad) {
  const int64_t r = r0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= r0 + rows) return;
  const int64_t a = indptr[r], b = indptr[r + 1];
  for (int64_t k = a + lane; k < b; k += 32) {
    const int g = idx[k - base];
    if (g < 0 || g >= G) { atomicOr(bad, 1); continue; }
    Yf[r * ldY + g] = (float)val[k - base];
  }
}
// flags: bit0 non-integer or negative, bit1 value > 255, bit2 value > 65535
__global__ void k_scan_y(const float* __restrict__ Y, int64_t ldY, int64_t N, int G, int* __restrict__ flags) {
  int64_t r = blockIdx.y;
  int f = 0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
    float y = Y[r * ldY + g];
    if (!(y >= 0.f) || y != floorf(y)) f |= 1;
    if (y > 255.f) f |= 2;
    if (y > 65535.f) f |= 4;
  }
  if (f) atomicOr(flags, f);
}
template <typename Tout>
__global__ void k_narrow_y(const float* __restrict__ Y, int64_t ldY, int64_t N, Tout* __restrict__ out) {
  int64_t r = blockIdx.y;
  for (int64_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ldY; g += (int64_t)gridDim.x * blockDim.x)
    out[r * ldY + g] = (Tout)Y[r * ldY + g];
}
__global__ void k_colmajor_to_rowmajor_f(const double* __restrict__ in, int64_t rows, int cols, float* __restrict__ out,
                                         int ld_out, int col_off) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int64_t r = i % rows;
  int c = (int)(i / rows);
  out[r * ld_out + col_off + c] = (float)in[i];
}

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct Prof {
  std::string name;
  cudaEvent_t a, b;
};

}  // namespace

// Device-resident inputs of a fit that do not depend on the restart: the count matrix as stored and
// everything derived from it once (library sizes, B = Y log L, multinomial constants, column sums, allele term).
// Sessions created with ca_core_create_shared read them in place (read-only), so the restarts of run_clonealign
// (R/clonealign.R:50-56) upload and preprocess Y once per device instead of once per fit.
struct ca_data {
  int dev = 0;
  int64_t N = 0, ldY = 0;
  int G = 0, C = 0, V = 0, ystore = CA_STORE_F32, poison = 0;
  double const_sum = 0.0;
  void* Y = nullptr;
  float *L = nullptr, *Bm = nullptr, *vA = nullptr, *s = nullptr, *colsum = nullptr, *snv = nullptr;
  std::vector<void*> allocs;
  std::atomic<int> refs{0};   // sessions created from it (host threads of concurrent restarts create / destroy them)
};

struct ca_handle {
  ca_config cfg{};
  ca_data* shared = nullptr;       // inputs owned by a ca_data (ca_core_create_shared), else by this handle
  bool data_only = false;          // ca_core_data_create: stop after the Y-derived part of build()
  int dev = 0, num_sms = 148;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap = true;
  int64_t N = 0, Ntot = 0, ldY = 0, Gld = 0, Nld = 0;
  int G = 0, C = 0, S = 0, K = 0, P = 0, KP = 0, SC = 0, SCp = 0, J = 0, V = 0;
  bool tc = false;
  bool interp = false;             // K = 1 univariate-interpolation path (kernels_interp.cuh)
  uint32_t variants = 0;           // enum ca_variant bits
  bool epi2 = false;               // interp path: fused Clenshaw + per-cell epilogue (kernels_fused.cuh)
  bool lean = false;               // with epi2: k_prologue / k_gene_fused / k_adam_all
  bool defer = false;              // with lean: Y-linear terms added after the per-cell kernel (late join of the Y pass)
  int y4_minb = 4;                 // k_ypass_k1_v4 register budget: sized for 4 (64 registers) or 3 (80) CTAs per SM
  bool cosched = false;            // with defer + ypass4: the Y pass starts first in the step, next to everything up to the gene kernel
  bool pending_join = false;       // a Y pass forked onto stream2 has not been joined yet
  int n_yv_blocks = 0;             // ELBO partials written by k_yv_dot (behind the per-cell kernel's in elbo_part)
  double* chi_cur = nullptr;
  float* pmm_part = nullptr;
  unsigned* ticket = nullptr;
  int gene_panels = 0;
  size_t gene_smem = 0;
  int fused_nj = 0, fused_panels = 0, fused_warps = kFusedWarps;
  size_t fused_smem = 0;
  int64_t n_cell_parts = 0;        // per-block ELBO / sum-gamma partials written by the per-cell kernel in use
  InterpPlan* iplan = nullptr;
  int n2_tj = 8, n2_ncgp = 32, n2_split_f = 1, n2_split_b = 1, n2_blocks_per_sm = kN2BlocksPerSM;   // k_interp_nodes2 launch geometry
  size_t n2_smem = 0;
  float* mm_psi = nullptr;
  double *ivals = nullptr, *icoef = nullptr;
  size_t ieval_smem = 0;
  int ieval_panels = 0;
  int ystore = CA_STORE_F32;
  int poison = 0;
  std::vector<void*> allocs;

  void* Y = nullptr;
  float *L = nullptr, *Bm = nullptr, *vA = nullptr, *s = nullptr, *colsum = nullptr, *snv = nullptr;
  double const_sum = 0.0;
  // trainable + Adam state + gradients
  float *U = nullptr, *Vm = nullptr, *chi_raw = nullptr, *u = nullptr, *loc = nullptr, *lsd = nullptr, *t = nullptr;
  float *m_U = nullptr, *v_U = nullptr, *m_V = nullptr, *v_V = nullptr, *m_chi = nullptr, *v_chi = nullptr;
  float *m_u = nullptr, *v_u = nullptr, *m_loc = nullptr, *v_loc = nullptr, *m_lsd = nullptr, *v_lsd = nullptr;
  float *m_t = nullptr, *v_t = nullptr;
  float *g_U = nullptr, *g_V = nullptr, *g_chi = nullptr, *g_u = nullptr, *g_loc = nullptr, *g_lsd = nullptr, *g_t = nullptr;
  // per-iteration scratch
  float *eps_in = nullptr, *eps = nullptr, *mu = nullptr, *logmu = nullptr, *sig = nullptr;
  float *Mx = nullptr, *shift = nullptr, *mm = nullptr, *Zx = nullptr, *Rx = nullptr, *dMx = nullptr, *dM_sum = nullptr;
  __nv_bfloat16 *MxT_hi = nullptr, *MxT_lo = nullptr;
  __half* RxT = nullptr;
  float* shift_bwd = nullptr;
  float *rowpart = nullptr, *colpart = nullptr, *YV = nullptr, *YtU = nullptr, *Fout = nullptr, *log_alpha = nullptr;
  float* ar = nullptr;
  double *gsum_part = nullptr, *elbo_part = nullptr, *gene_part = nullptr, *scal_elbo = nullptr, *cell_sum = nullptr,
         *wsq = nullptr, *elbo_dev = nullptr;
  int nCB = 1, nRB = 1, RB = 512, n_gene_blocks = 0, nsplit = 1;
  int64_t n_epi_blocks = 0;
  bool ydirty = true;
  bool inspect = false;            // test hook (ca_core_grads): also write inspection-only arrays (Z of the fused kernel)
  TcPlan tcplan;

  std::vector<float> eps_queue;   // host-fed draws, S*G floats each
  int64_t eps_q_head = 0;         // next draw to consume
  uint64_t draw = 0;
  int adam_t = 0;

  StepState* dstate = nullptr;     // device-side counters (draw, adam_t, lr_t, p2p_step): constant launch arguments
  cudaGraphExec_t g_train[2] = {nullptr, nullptr}, g_eval[2] = {nullptr, nullptr};   // replayable step / evaluation, by ydirty
  bool use_graph = false;
  void* comm = nullptr;
  // variant P2P: exchange buffer of this rank, the peers' mappings, step counter
  bool p2p = false, p2p_ready = false;
  float* p2p_buf = nullptr;
  int64_t p2p_cnt = 0, p2p_cnt_pad = 0;
  float* p2p_slots[kP2PMaxWorld] = {};
  unsigned* p2p_flags[kP2PMaxWorld] = {};
  void* p2p_mapped[kP2PMaxWorld] = {};
  unsigned* p2p_ticket = nullptr;
  int* p2p_err = nullptr;
  unsigned p2p_step = 0;
  bool prof_on = false;
  std::vector<Prof> prof;
  int launches_last_step = 0;

  template <typename T> T* alloc(size_t n, bool zero = true) {
    void* p = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    CUDA_OK(cudaMalloc(&p, bytes));
    allocs.push_back(p);
    if (zero) CUDA_OK(cudaMemsetAsync(p, 0, bytes, stream));
    return (T*)p;
  }
  void release(void* p) {
    for (auto& q : allocs)
      if (q == p) { cudaFree(p); q = nullptr; }
  }
};

namespace {

struct LaunchScope {
  ca_handle* h;
  bool on;
  size_t idx = 0;
  LaunchScope(ca_handle* h_, const char* name, int n_kernels = 1) : h(h_), on(h_->prof_on) {
    h->launches_last_step += n_kernels;
    if (on) {
      Prof p;
      p.name = name;
      CUDA_OK(cudaEventCreate(&p.a));
      CUDA_OK(cudaEventCreate(&p.b));
      CUDA_OK(cudaEventRecord(p.a, h->stream));
      h->prof.push_back(p);
      idx = h->prof.size() - 1;
    }
  }
  ~LaunchScope() {
    if (on) cudaEventRecord(h->prof[idx].b, h->stream);
  }
};
#define KCHECK() CUDA_OK(cudaGetLastError())

template <typename F>
void dispatch_y(ca_handle* h, F&& f) {
  switch (h->ystore) {
    case CA_STORE_F32: f((const float*)h->Y); break;
    case CA_STORE_U16: f((const uint16_t*)h->Y); break;
    case CA_STORE_U8: f((const uint8_t*)h->Y); break;
    default: fail("bad y_store");
  }
}

AdamHyper adam_hyper(ca_handle* h, bool apply) {
  AdamHyper a;
  int t = h->adam_t + 1;
  double b1 = 0.9, b2 = 0.999;
  a.lr_t = (float)(h->cfg.learning_rate * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t)));
  a.b1 = 0.9f;
  a.b2 = 0.999f;
  a.eps = 1e-8f;
  a.apply = apply ? 1 : 0;
  return a;
}

// ---- the Y pass (K3) ---------------------------------------------------------------------------
void run_ypass(ca_handle* h, cudaStream_t st) {
  if (h->KP == 0 || !h->ydirty) return;
  dispatch_y(h, [&](auto* Yp) {
    using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
    if (h->KP == 1) {
      LaunchScope ls(h, "ypass");
      dim3 grid(h->nCB, h->nRB);
      if (h->variants & CA_VAR_YPASS4) {
        const int64_t tiles = (int64_t)h->nCB * h->nRB;
        // persistent grid: 2 CTAs per SM (64 KB rings); fp32 storage has 128 KB rings: one per SM
        const int per_sm = std::is_same<T, float>::value ? 1 : 2;
        const unsigned g4 = (unsigned)std::min<int64_t>(tiles, per_sm * (int64_t)h->num_sms);
        if (h->y4_minb == 3) {
          auto k = k_ypass_k1_v4<T, 3>;
          CA_LAUNCH(k, g4, 256, ypass4_smem_bytes<T>(), st)(Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm, h->rowpart, h->colpart);
        } else {
          auto k = k_ypass_k1_v4<T, 4>;
          CA_LAUNCH(k, g4, 256, ypass4_smem_bytes<T>(), st)(Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm, h->rowpart, h->colpart);
        }
      } else if (h->variants & CA_VAR_YPASS3) {
        CA_LAUNCH(k_ypass_k1_v3<T>, grid, 256, 0, st)(Yp, h->ldY, h->N, h->G, h->RB, h->U, h->Vm, h->rowpart, h->colpart);
      } else if (h->variants & CA_VAR_YPASS2) {
        CA_LAUNCH(k_ypass_k1_v2<T>, grid, 256, 0, st)(Yp, h->ldY, h->N, h->G, h->RB, h->U, h->Vm, h->rowpart, h->colpart);
      } else if (st == h->stream && !getenv("CLONEALIGN_B200_YPASS_LIGHT")) {
        CA_LAUNCH(k_ypass_k1<T>, grid, 256, 0, st)(Yp, h->ldY, h->N, h->G, h->RB, h->U, h->Vm, h->rowpart, h->colpart);
      } else {
        int64_t tiles = (int64_t)h->nCB * h->nRB;
        unsigned g = (unsigned)std::min<int64_t>(tiles, 2 * (int64_t)h->num_sms);
        CA_LAUNCH(k_ypass_k1_persistent<T>, g, 256, 0, st)(Yp, h->ldY, h->N, h->G, h->RB, h->nCB, h->nRB, h->U, h->Vm, h->rowpart,
                                                    h->colpart);
      }
      KCHECK();
    } else {
      {
        LaunchScope ls(h, "ypass_rows");
        CA_LAUNCH(k_ypass_rows_generic<T>, (unsigned)ceil_div64(h->N, 8), 256, 0, st)(Yp, h->ldY, h->N, h->G, h->KP, h->Vm,
                                                                                 h->rowpart);
        KCHECK();
      }
      {
        LaunchScope ls(h, "ypass_cols");
        dim3 grid((h->G + 127) / 128, h->nRB);
        CA_LAUNCH(k_ypass_cols_generic<T>, grid, 128, 0, st)(Yp, h->ldY, h->N, h->G, h->KP, h->RB, h->U, h->colpart);
        KCHECK();
      }
    }
  });
  h->ydirty = false;
}

// ---- forward: eps -> mu, Mx -> shift -> Zx -> (Y pass) -> epilogue -------------------------------
void stage_eps(ca_handle* h, const float** eps_in) {
  *eps_in = nullptr;
  size_t per = (size_t)h->S * h->G;
  if ((size_t)h->eps_q_head * per < h->eps_queue.size()) {
    CUDA_OK(cudaMemcpyAsync(h->eps_in, h->eps_queue.data() + (size_t)h->eps_q_head * per, per * sizeof(float),
                            cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));   // source is pageable host memory owned by the queue
    h->eps_q_head++;
    if ((size_t)h->eps_q_head * per >= h->eps_queue.size()) {
      h->eps_queue.clear();
      h->eps_q_head = 0;
    }
    *eps_in = h->eps_in;
  }
}

// node sums + coefficients of the interp path (kernels_interp.cuh, k_interp_nodes2 / k_interp_coeffs2)
template <bool FWD>
void launch_interp_nodes(ca_handle* h, const float* rv, const float* shift, const float* B, int64_t R) {
  const int nsplit = FWD ? h->n2_split_f : h->n2_split_b;
  const int max_pan = FWD ? kIMaxPanF : kIMaxPanB;
  const unsigned grid = (unsigned)std::min<int64_t>((int64_t)max_pan * nsplit, (int64_t)h->n2_blocks_per_sm * h->num_sms);
  if (h->n2_tj == 8) {
    auto k = k_interp_nodes2<FWD, 8>;
    CA_LAUNCH(k, grid, kN2Threads, h->n2_smem, h->stream)(h->iplan, rv, shift, B, R, h->J, h->n2_ncgp, nsplit, max_pan, h->ivals);
  } else {
    auto k = k_interp_nodes2<FWD, 6>;
    CA_LAUNCH(k, grid, kN2Threads, h->n2_smem, h->stream)(h->iplan, rv, shift, B, R, h->J, h->n2_ncgp, nsplit, max_pan, h->ivals);
  }
  CA_LAUNCH(k_interp_coeffs2, dim3((h->J + kC2Cols - 1) / kC2Cols, kC2PanelsY), kIP * kC2Cols * kC2Lanes, 0, h->stream)(
      h->iplan, h->ivals, nsplit, max_pan, h->J, FWD ? 1 : 0, h->icoef);
}

// the partial sums of the Y pass are needed from here on: wait for the pass forked onto stream2, or run it now
void join_ypass(ca_handle* h, int mode) {
  if (h->pending_join) {
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    h->pending_join = false;
  } else if (mode != EPI_INIT) {
    run_ypass(h, h->stream);
  }
}

template <int MODE>
void launch_fused_mode(ca_handle* h, const FusedArgs& a) {
  const unsigned grid = (unsigned)h->n_cell_parts;
  switch (h->fused_nj) {
    case 1: { auto k = k_cell_fused<MODE, 1>; CA_LAUNCH(k, grid, h->fused_warps * 32, h->fused_smem, h->stream)(a); break; }
    case 2: { auto k = k_cell_fused<MODE, 2>; CA_LAUNCH(k, grid, h->fused_warps * 32, h->fused_smem, h->stream)(a); break; }
    case 3: { auto k = k_cell_fused<MODE, 3>; CA_LAUNCH(k, grid, h->fused_warps * 32, h->fused_smem, h->stream)(a); break; }
    case 4: { auto k = k_cell_fused<MODE, 4>; CA_LAUNCH(k, grid, h->fused_warps * 32, h->fused_smem, h->stream)(a); break; }
    default: fail("fused per-cell kernel: unsupported S*C");
  }
}
void launch_fused(ca_handle* h, int mode, const FusedArgs& a) {
  if (mode == EPI_TRAIN) launch_fused_mode<EPI_TRAIN>(h, a);
  else if (mode == EPI_EVAL) launch_fused_mode<EPI_EVAL>(h, a);
  else launch_fused_mode<EPI_INIT>(h, a);
}
template <int NJ>
void fused_set_smem(size_t smem) {
  CUDA_OK(cudaFuncSetAttribute(k_cell_fused<EPI_TRAIN, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(k_cell_fused<EPI_EVAL, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(k_cell_fused<EPI_INIT, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

void run_forward(ca_handle* h, int mode) {
  const float* eps_in;
  stage_eps(h, &eps_in);
  // The Y stream (HBM-bound, touches only Y, psi, W) is independent of the forward contraction (tensor / MUFU
  // bound): fork it onto a second stream so both run on the SMs at once; joined before the per-cell epilogue.
  bool joined_later = false;
  const bool want_fork = mode != EPI_INIT && (h->overlap || h->cosched) && !h->prof_on && h->ydirty && h->KP > 0;
  auto fork_ypass = [&]() {
    CUDA_OK(cudaEventRecord(h->ev_fork, h->stream));
    CUDA_OK(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    run_ypass(h, h->stream2);
    CUDA_OK(cudaEventRecord(h->ev_join, h->stream2));
    joined_later = true;
  };
  // Variant DEFER forks later, right before the per-cell kernel: two Y-pass CTAs take the whole register file of an SM, so a
  // pass started here would only push the short gene-level launches (prologue, node sums, coefficients: the head of the
  // step's critical path) behind its first wave; started together with the per-cell kernel (half a register file per
  // CTA) it shares every SM with it instead.
  if (want_fork && (!h->defer || h->cosched)) fork_ypass();   // cosched: the persistent 2-CTA-per-SM pass goes first, everything else fits next to it
  SampleMuArgs sm;
  sm.G = h->G; sm.C = h->C; sm.S = h->S; sm.K = h->K; sm.KP = h->KP; sm.SCp = h->SCp; sm.J = h->J; sm.Gld = h->Gld;
  sm.loc = h->loc; sm.lsd = h->lsd; sm.Vm = h->Vm; sm.L = h->L; sm.colsum = h->colsum; sm.chi_raw = h->chi_raw;
  sm.eps_in = eps_in; sm.seed = h->cfg.seed; sm.draw = h->draw++;
  sm.eps_out = h->eps; sm.mu = h->mu; sm.logmu = h->logmu; sm.sig = h->sig;
  sm.Mx = h->tc ? nullptr : h->Mx; sm.MxT_hi = h->tc ? h->MxT_hi : nullptr; sm.MxT_lo = h->tc ? h->MxT_lo : nullptr;
  sm.gene_part = h->gene_part;
  if (h->lean) {
    LaunchScope ls(h, "prologue");
    PrologueArgs a;
    a.N = h->N; a.G = h->G; a.C = h->C; a.K = h->K;
    a.u = h->u; a.chi_raw = h->chi_raw; a.Vm = h->Vm; a.U = h->U;
    a.log_alpha = h->log_alpha; a.mm = h->mm; a.mm_psi = h->mm_psi; a.chi_cur = h->chi_cur;
    a.scal_elbo = h->scal_elbo; a.wsq = h->wsq; a.pmm_part = h->pmm_part; a.ticket = h->ticket; a.plan = h->iplan;
    a.dirichlet_const = (double)h->C * lgamma(1.0 / h->C) - lgamma(1.0);
    a.state = h->dstate; a.lr = h->cfg.learning_rate;
    a.mu = sm;
    a.mu_vec4 = (h->C % 4 == 0) ? 1 : 0;
    CA_LAUNCH(k_prologue, 2 + kProPsiBlocks + h->n_gene_blocks, kProThreads, 0, h->stream)(a);
    KCHECK();
  } else {
    LaunchScope ls(h, "alpha");
    CA_LAUNCH(k_alpha, 1, 32, 0, h->stream)(h->u, h->C, h->chi_raw, h->K, h->log_alpha, h->scal_elbo);
    KCHECK();
  }
  if (!h->lean) {
    LaunchScope ls(h, "sample_mu");
    CA_LAUNCH(k_sample_mu, h->n_gene_blocks, 256, 0, h->stream)(sm);
    KCHECK();
  }
  if (h->KP == 0) {
    CUDA_OK(cudaMemsetAsync(h->shift, 0, sizeof(float) * h->N, h->stream));
  } else if (h->lean) {
    // W range (and sum of squares) come from k_prologue, m_n from the fused per-cell kernel
  } else if (h->K == 1 && h->P == 0) {
    LaunchScope ls(h, "shift", h->epi2 ? 1 : 2);
    CA_LAUNCH(k_minmax, 1, 1024, 0, h->stream)(h->Vm, h->G, h->mm);
    KCHECK();
    if (!h->epi2) {   // EPI2 computes m_n inside the fused per-cell kernel
      CA_LAUNCH(k_shift_k1, (unsigned)ceil_div64(h->N, 256), 256, 0, h->stream)(h->U, h->mm, h->N, h->shift);
      KCHECK();
    }
  } else {
    LaunchScope ls(h, "shift");
    CA_LAUNCH(k_shift_general, (unsigned)ceil_div64(h->N, 8), 256, 0, h->stream)(h->U, h->Vm, h->N, h->G, h->KP, h->shift);
    KCHECK();
  }
  {
    LaunchScope ls(h, "lse_fwd", h->interp ? (h->lean ? 2 : (h->epi2 ? 4 : 5)) : 1);
    if (h->interp) {
      // K = 1: Zx[n][j] = F_j(psi_n) by piecewise Chebyshev interpolation (kernels_interp.cuh)
      if (!h->lean) {
        CA_LAUNCH(k_minmax, 1, 1024, 0, h->stream)(h->U, (int)h->N, h->mm_psi);
        CA_LAUNCH(k_interp_plan, 1, 32, 0, h->stream)(h->mm, h->mm_psi, h->iplan);
      }
      launch_interp_nodes<true>(h, h->Vm, nullptr, h->Mx, h->G);
      if (!h->epi2)
        CA_LAUNCH(k_interp_eval<true>, h->num_sms, kIEvalWarps * 32, h->ieval_smem, h->stream)(h->iplan, h->icoef, h->U, h->N, h->J, h->Zx,
                                                                                    h->ieval_panels);
    } else if (h->tc) {
      tc_launch_fwd(h->tcplan, h->U, h->Vm, h->shift, h->Zx, h->stream);
    } else {
      int Jc = (mode == EPI_TRAIN) ? h->J : h->SC;   // ELBO-only passes need just Z
      dim3 grid((Jc + 63) / 64, (unsigned)ceil_div64(h->N, 64));
      CA_LAUNCH(k_expgemm<true>, grid, 256, 0, h->stream)(h->U, h->Vm, h->shift, h->Mx, h->Zx, h->N, h->G, Jc, h->J, h->KP);
    }
    KCHECK();
  }
  h->pending_join = joined_later;
  if (!h->defer) join_ypass(h, mode);
  if (h->epi2) {
    LaunchScope ls(h, mode == EPI_TRAIN ? "cell_epilogue" : (mode == EPI_EVAL ? "cell_epilogue_eval" : "gamma_init"));
    FusedArgs a;
    a.N = h->N; a.C = h->C; a.S = h->S; a.SC = h->SC; a.J = h->J; a.nCB = h->nCB; a.smem_panels = h->fused_panels;
    a.plan = h->iplan; a.coeff = h->icoef; a.mm = h->mm;
    a.U = h->U; a.Bm = h->Bm; a.vA = h->vA; a.s = h->s; a.log_alpha = h->log_alpha; a.rowpart = h->rowpart;
    a.t = h->t; a.gT = h->g_t; a.Rx = h->Rx; a.gU = h->g_U; a.YV = h->YV; a.shift = h->shift;
    a.Fout = h->inspect ? h->Fout : nullptr;     // inspection copies (ca_core_grads): not written by the timed path
    a.Zx = h->inspect ? h->Zx : nullptr;
    a.elbo_part = h->elbo_part; a.gsum_part = h->gsum_part;
    a.defer_yv = h->defer ? 1 : 0;
    // DEFER + OVERLAP: the pass may start once everything before the per-cell kernel is done (event recorded here), but
    // it is handed to the device AFTER the per-cell kernel, whose 148 persistent CTAs should be placed first
    if (want_fork && h->defer && !h->cosched) CUDA_OK(cudaEventRecord(h->ev_fork, h->stream));
    launch_fused(h, mode, a);
    KCHECK();
    if (want_fork && h->defer && !h->cosched) {
      CUDA_OK(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
      run_ypass(h, h->stream2);
      CUDA_OK(cudaEventRecord(h->ev_join, h->stream2));
      h->pending_join = true;
    }
    if (h->defer && mode == EPI_EVAL) {   // the ELBO needs sum_n psi_n (YW)_n now; a train step joins before k_gene_fused
      join_ypass(h, mode);
      LaunchScope ls2(h, "yv_dot");
      CA_LAUNCH(k_yv_dot, h->n_yv_blocks, 256, 0, h->stream)(h->N, h->nCB, h->rowpart, h->U, h->YV, h->elbo_part + h->n_cell_parts);
      KCHECK();
    }
  } else {
    LaunchScope ls(h, mode == EPI_TRAIN ? "cell_epilogue" : (mode == EPI_EVAL ? "cell_epilogue_eval" : "gamma_init"));
    EpiArgs a;
    a.N = h->N; a.Nld = h->Nld; a.C = h->C; a.S = h->S; a.SCp = h->SCp; a.J = h->J; a.K = h->K; a.KP = h->KP; a.nCB = h->nCB;
    a.fsplit = h->tc ? h->tcplan.fsplit : 1;
    a.Zx = h->Zx; a.Bm = h->Bm; a.vA = h->vA; a.s = h->s; a.shift = h->shift; a.log_alpha = h->log_alpha; a.U = h->U;
    a.rowpart = h->rowpart; a.t = h->t; a.gT = h->g_t; a.Rx = h->tc ? nullptr : h->Rx; a.gU = h->g_U; a.YV = h->YV;
    a.Fout = h->Fout; a.RxT = h->tc ? h->RxT : nullptr; a.shift_bwd = h->shift_bwd; a.elbo_part = h->elbo_part; a.gsum_part = h->gsum_part;
    size_t smem = epi_smem_bytes(h->SCp, h->C, h->J, h->tc);
    unsigned grid = (unsigned)h->n_epi_blocks;
    if (mode == EPI_TRAIN) CA_LAUNCH(k_cell_epilogue<EPI_TRAIN>, grid, kEpiWarps * 32, smem, h->stream)(a);
    else if (mode == EPI_EVAL) CA_LAUNCH(k_cell_epilogue<EPI_EVAL>, grid, kEpiWarps * 32, smem, h->stream)(a);
    else CA_LAUNCH(k_cell_epilogue<EPI_INIT>, grid, kEpiWarps * 32, smem, h->stream)(a);
    KCHECK();
  }
}

void run_train(ca_handle* h, bool apply) {
  h->launches_last_step = 0;
  run_forward(h, EPI_TRAIN);
  {
    LaunchScope ls(h, "lse_bwd", h->interp ? (h->lean ? 2 : 3) : 1);
    if (h->interp) {
      // K = 1: dMx[g][j] = H_j(w_g); the plan of this step's forward pass is still valid (psi, W unchanged)
      launch_interp_nodes<false>(h, h->U, h->shift, h->Rx, h->N);
      if (!h->lean)
        CA_LAUNCH(k_interp_eval<false>, h->num_sms, kIEvalWarps * 32, h->ieval_smem, h->stream)(h->iplan, h->icoef, h->Vm, h->G, h->J, h->dMx,
                                                                                     h->ieval_panels);
    } else if (h->tc) {
      tc_launch_bwd(h->tcplan, h->U, h->Vm, h->shift_bwd, h->dMx, h->stream);
    } else {
      dim3 grid((h->J + 63) / 64, (h->G + 63) / 64);
      CA_LAUNCH(k_expgemm<false>, grid, 256, 0, h->stream)(h->Vm, h->U, h->shift, h->Rx, h->dMx, h->G, h->N, h->J, h->J, h->KP);
    }
    KCHECK();
  }
  if (h->defer) join_ypass(h, EPI_TRAIN);   // colpart (gene gradients) and rowpart (d psi in k_adam_all) are needed from here on
  if (h->lean) {
    LaunchScope ls(h, "gene_grads", 1);
    GeneFusedArgs a;
    a.G = h->G; a.C = h->C; a.S = h->S; a.SC = h->SC; a.J = h->J; a.nRB = h->nRB; a.smem_panels = h->gene_panels;
    a.plan = h->iplan; a.coeff = h->icoef;
    a.Vm = h->Vm; a.colpart = h->colpart; a.mu = h->mu; a.sig = h->sig; a.eps = h->eps; a.lsd = h->lsd; a.L = h->L;
    a.ar = h->ar; a.YtU = h->YtU; a.dM_out = h->inspect ? h->dM_sum : nullptr;
    a.gsum_part = h->gsum_part; a.n_parts = h->n_cell_parts;
    // two 512-thread blocks per SM (64 registers, <= 78 KB of coefficients each): 32 warps keep the fp64 recurrences fed
    switch ((h->SC + 31) / 32) {
      case 1: { auto k = k_gene_fused<1>; CA_LAUNCH(k, 2 * h->num_sms + 1, kGeneWarps * 32, h->gene_smem, h->stream)(a); break; }
      case 2: { auto k = k_gene_fused<2>; CA_LAUNCH(k, 2 * h->num_sms + 1, kGeneWarps * 32, h->gene_smem, h->stream)(a); break; }
      case 3: { auto k = k_gene_fused<3>; CA_LAUNCH(k, 2 * h->num_sms + 1, kGeneWarps * 32, h->gene_smem, h->stream)(a); break; }
      default: { auto k = k_gene_fused<4>; CA_LAUNCH(k, 2 * h->num_sms + 1, kGeneWarps * 32, h->gene_smem, h->stream)(a); break; }
    }
    KCHECK();
  } else {
    LaunchScope ls(h, "gene_grads", 2);
    GeneGradArgs a;
    a.G = h->G; a.C = h->C; a.S = h->S; a.K = h->K; a.KP = h->KP; a.SCp = h->SCp; a.J = h->J; a.nsplit = h->nsplit; a.nRB = h->nRB;
    a.dMx = h->dMx; a.colpart = h->colpart; a.mu = h->mu; a.sig = h->sig; a.eps = h->eps; a.lsd = h->lsd; a.L = h->L;
    a.ar = h->ar; a.YtU = h->YtU; a.dM_out = h->dM_sum;
    CA_LAUNCH(k_gene_grads_warp, (h->G + 7) / 8, 256, 0, h->stream)(a);
    KCHECK();
    CA_LAUNCH(k_reduce_gsum, 1, 1024, 0, h->stream)(h->gsum_part, h->n_cell_parts, h->C, h->ar + (int64_t)h->G * (2 + h->KP));
    KCHECK();
  }
  if (h->cfg.world > 1 && h->p2p) {
    if (!h->p2p_ready) fail("variant p2p: ca_core_p2p_connect has not been called");
    LaunchScope ls(h, "allreduce");
    P2PArgs a;
    a.world = h->cfg.world; a.rank = h->cfg.rank; a.cnt = h->p2p_cnt; a.cnt_pad = h->p2p_cnt_pad; a.step_ctr = &h->dstate->p2p_step;
    a.src = h->ar; a.dst = h->ar; a.ticket = h->p2p_ticket; a.error = h->p2p_err;
    for (int r = 0; r < kP2PMaxWorld; ++r) { a.slots[r] = h->p2p_slots[r]; a.flags[r] = h->p2p_flags[r]; }
    CA_LAUNCH(k_p2p_allreduce, std::min(h->num_sms, 64), kP2PThreads, 0, h->stream)(a);
    KCHECK();
  } else if (h->cfg.world > 1) {
    LaunchScope ls(h, "allreduce");
    size_t cnt = (size_t)h->G * (2 + h->KP) + h->C;
    NCCL_OK(nccl().AllReduce(h->ar, h->ar, cnt, kNcclFloat32, kNcclSum, h->comm, h->stream));
  }
  {
    LaunchScope ls(h, "adam", h->lean ? 1 : (apply ? 4 : 3));
    AdamHyper hy = adam_hyper(h, apply);
    if (!h->lean) {
      CA_LAUNCH(k_wsq, 1, 1024, 0, h->stream)(h->Vm, h->G, h->K, h->KP, h->wsq);
      KCHECK();
    }
    ScalarAdamArgs sa;
    sa.G = h->G; sa.C = h->C; sa.K = h->K; sa.n_total = (double)h->Ntot; sa.wsq = h->wsq;
    sa.gsum = h->ar + (int64_t)h->G * (2 + h->KP);
    sa.chi_raw = h->chi_raw; sa.m_chi = h->m_chi; sa.v_chi = h->v_chi; sa.g_chi = h->g_chi;
    sa.u = h->u; sa.m_u = h->m_u; sa.v_u = h->v_u; sa.g_u = h->g_u; sa.h = hy;
    GeneAdamArgs ga;
    ga.G = h->G; ga.S = h->S; ga.K = h->K; ga.KP = h->KP; ga.ar = h->ar; ga.mu = h->mu; ga.logmu = h->logmu; ga.sig = h->sig;
    ga.eps = h->eps; ga.colsum = h->colsum; ga.chi_raw = h->chi_raw; ga.loc = h->loc; ga.lsd = h->lsd; ga.Vm = h->Vm;
    ga.m_loc = h->m_loc; ga.v_loc = h->v_loc; ga.m_lsd = h->m_lsd; ga.v_lsd = h->v_lsd; ga.m_V = h->m_V; ga.v_V = h->v_V;
    ga.g_loc = h->g_loc; ga.g_lsd = h->g_lsd; ga.g_V = h->g_V; ga.h = hy;
    if (h->lean) {
      AdamAllArgs aa;
      aa.ga = ga; aa.chi_cur = h->chi_cur; aa.sa = sa; aa.N = h->N; aa.C = h->C;
      aa.t = h->t; aa.m_t = h->m_t; aa.v_t = h->v_t; aa.U = h->U; aa.m_U = h->m_U; aa.v_U = h->v_U; aa.gT = h->g_t; aa.gU = h->g_U;
      aa.n_gene_blocks = (h->G + 255) / 256;
      aa.n_cell_blocks = (apply || h->defer) ? ceil_div64(ceil_div64(h->N * h->C, 4) + h->N, 256) : 0;
      aa.defer_yv = h->defer ? 1 : 0; aa.nCB = h->nCB; aa.rowpart = h->rowpart; aa.YV = h->YV;
      aa.state = h->dstate;
      CA_LAUNCH(k_adam_all, (unsigned)(aa.n_gene_blocks + aa.n_cell_blocks + 1), 256, 0, h->stream)(aa);
      KCHECK();
    } else {
    // gene kernel reads chi_raw (old) -> must precede the scalar update
    CA_LAUNCH(k_gene_adam, (h->G + 127) / 128, 128, 0, h->stream)(ga);
    KCHECK();
    CA_LAUNCH(k_scalar_adam, 1, 32, 0, h->stream)(sa);
    KCHECK();
    }
    if (apply && !h->lean) {
      int64_t tot = h->N * h->C + h->N * h->KP;
      CA_LAUNCH(k_cell_adam, (unsigned)ceil_div64(tot, 256), 256, 0, h->stream)(h->N, h->C, h->K, h->KP, h->t, h->m_t, h->v_t, h->g_t,
                                                                       h->U, h->m_U, h->v_U, h->g_U, hy);
      KCHECK();
    }
  }
  if (apply) {
    h->adam_t++;
    h->ydirty = true;
  }
}

void run_elbo_async(ca_handle* h) {
  h->launches_last_step = 0;
  run_forward(h, EPI_EVAL);
  LaunchScope ls(h, "elbo_reduce", 2);
  CA_LAUNCH(k_reduce_partials, 1, 1024, 0, h->stream)(h->elbo_part, h->n_cell_parts + h->n_yv_blocks, 1, h->cell_sum, h->const_sum);
  KCHECK();
  if (h->cfg.world > 1) NCCL_OK(nccl().AllReduce(h->cell_sum, h->cell_sum, 1, kNcclFloat64, kNcclSum, h->comm, h->stream));
  CA_LAUNCH(k_elbo_final, 1, 256, 0, h->stream)(h->cell_sum, h->gene_part, h->n_gene_blocks, h->scal_elbo, h->poison, h->elbo_dev);
  KCHECK();
}

// ---- CUDA-graph replay of the train step / the ELBO evaluation --------------------------------------
// The fused (lean) kernel set keeps everything that changes from step to step in device memory (StepState), so the
// launches of a step have constant arguments: the sequence is captured once per (kind, "Y pass needed") and replayed with
// one cudaGraphLaunch -- 7-9 launches, the fork / join of the Y-pass stream and the all-reduce of a sharded fit included.
// Not used with host-fed draws (test hook), per-kernel profiling or inspection copies; CLONEALIGN_B200_NO_GRAPH=1 disables it.
bool graph_ok(ca_handle* h) {
#ifdef CA_EMULATE
  return false;
#else
  return h->use_graph && h->lean && !h->prof_on && !h->inspect && h->eps_queue.empty();
#endif
}
#ifndef CA_EMULATE
template <typename F>
void capture_or_replay(ca_handle* h, cudaGraphExec_t& exec, F&& body, const std::function<void()>& host_effects) {
  if (!exec) {
    cudaGraph_t graph = nullptr;
    CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    try {
      body();                                  // also applies the host-side bookkeeping once
    } catch (...) {
      cudaStreamEndCapture(h->stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CUDA_OK(cudaStreamEndCapture(h->stream, &graph));
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_OK(e);
  } else {
    host_effects();
  }
  CUDA_OK(cudaGraphLaunch(exec, h->stream));
}
#endif
void train_step(ca_handle* h) {
#ifndef CA_EMULATE
  if (graph_ok(h)) {
    const int key = h->ydirty ? 1 : 0;
    const int n_launch = h->launches_last_step;
    capture_or_replay(h, h->g_train[key], [&] { run_train(h, true); },
                      [&] { h->draw++; h->adam_t++; h->ydirty = true; h->pending_join = false; (void)n_launch; });
    return;
  }
#endif
  run_train(h, true);
}
void eval_step(ca_handle* h) {
#ifndef CA_EMULATE
  if (graph_ok(h)) {
    const int key = h->ydirty ? 1 : 0;
    capture_or_replay(h, h->g_eval[key], [&] { run_elbo_async(h); },
                      [&] { h->draw++; if (h->KP > 0) h->ydirty = false; h->pending_join = false; });
    return;
  }
#endif
  run_elbo_async(h);
}

// ---- host <-> device helpers ---------------------------------------------------------------------
void upload_colmajor(ca_handle* h, const double* src, int64_t rows, int cols, float* dst, int ld_dst, int col_off) {
  if (!src || rows * cols == 0) return;
  double* tmp = nullptr;
  CUDA_OK(cudaMalloc(&tmp, sizeof(double) * rows * cols));
  CUDA_OK(cudaMemcpyAsync(tmp, src, sizeof(double) * rows * cols, cudaMemcpyHostToDevice, h->stream));
  CA_LAUNCH(k_colmajor_to_rowmajor_f, (unsigned)ceil_div64(rows * cols, 256), 256, 0, h->stream)(tmp, rows, cols, dst, ld_dst, col_off);
  KCHECK();
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaFree(tmp));
}

// device row-major float [rows][ld] (columns col_off..col_off+cols) -> host column-major double
void download_colmajor(ca_handle* h, const float* src, int64_t rows, int cols, int ld, int col_off, double* out) {
  if (!out || rows * cols == 0) return;
  std::vector<float> tmp((size_t)rows * ld);
  CUDA_OK(cudaMemcpyAsync(tmp.data(), src, sizeof(float) * rows * ld, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < cols; ++c)
    for (int64_t r = 0; r < rows; ++r) out[(int64_t)c * rows + r] = (double)tmp[(size_t)r * ld + col_off + c];
}

template <typename Tin>
void ingest_y(ca_handle* h, const Tin* Ysrc, float* Yf) {
  const ca_config& c = h->cfg;
  const int64_t N = h->N;
  const int G = h->G;
  const bool on_dev = c.y_mem == CA_Y_DEVICE;
  if (c.y_layout == CA_Y_CSR) {
    if (on_dev) fail("CSR input must be in host memory");
    if (!c.y_indptr || !c.y_indices) fail("CSR input needs y_indptr and y_indices");
    const int32_t* ip = c.y_indptr;
    if (ip[0] < 0) fail("bad CSR row offsets");
    for (int64_t r = 0; r < N; ++r)
      if (ip[r + 1] < ip[r]) fail("bad CSR row offsets");
    int *d_ip = nullptr, *d_idx = nullptr, *d_bad = nullptr;
    Tin* d_val = nullptr;
    const int64_t cap = std::max<int64_t>(1, (int64_t)(128ll << 20) / (int64_t)(sizeof(Tin) + sizeof(int)));   // entries per chunk
    CUDA_OK(cudaMalloc(&d_ip, sizeof(int) * (N + 1)));
    CUDA_OK(cudaMalloc(&d_bad, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
    CUDA_OK(cudaMemcpyAsync(d_ip, ip, sizeof(int) * (N + 1), cudaMemcpyHostToDevice, h->stream));
    int64_t r0 = 0;
    int64_t cur_cap = 0;
    while (r0 < N) {
      // rows [r0, r1) whose entries fit in one chunk (a single row longer than the chunk gets a chunk of its own)
      int64_t r1 = r0 + 1;
      while (r1 < N && (int64_t)ip[r1 + 1] - ip[r0] <= cap) ++r1;
      const int64_t base = ip[r0], cnt = (int64_t)ip[r1] - base;
      if (cnt > cur_cap) {
        if (d_idx) { CUDA_OK(cudaFree(d_idx)); CUDA_OK(cudaFree(d_val)); }
        cur_cap = std::max(cnt, cap);
        CUDA_OK(cudaMalloc(&d_idx, sizeof(int) * cur_cap));
        CUDA_OK(cudaMalloc(&d_val, sizeof(Tin) * cur_cap));
      }
      if (cnt > 0) {
        CUDA_OK(cudaMemcpyAsync(d_idx, c.y_indices + base, sizeof(int) * cnt, cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(cudaMemcpyAsync(d_val, Ysrc + base, sizeof(Tin) * cnt, cudaMemcpyHostToDevice, h->stream));
        CA_LAUNCH(k_ingest_csr<Tin>, (unsigned)ceil_div64(r1 - r0, 8), 256, 0, h->stream)(d_ip, d_idx, d_val, base, r0, r1 - r0, G, Yf,
                                                                                          h->ldY, d_bad);
        KCHECK();
        CUDA_OK(cudaStreamSynchronize(h->stream));
      }
      r0 = r1;
    }
    int hbad = 0;
    CUDA_OK(cudaMemcpyAsync(&hbad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(d_ip); cudaFree(d_bad);
    if (d_idx) { cudaFree(d_idx); cudaFree(d_val); }
    if (hbad) fail("CSR input has a gene index outside [0, G)");
  } else if (c.y_layout == CA_Y_COLMAJOR) {
    int64_t ld = c.y_ld ? c.y_ld : N;
    int gchunk = (int)std::max<int64_t>(1, std::min<int64_t>(G, (int64_t)(256ll << 20) / (int64_t)(sizeof(Tin) * N)));
    Tin* stage = nullptr;
    if (!on_dev) CUDA_OK(cudaMalloc(&stage, sizeof(Tin) * (size_t)gchunk * N));
    for (int g0 = 0; g0 < G; g0 += gchunk) {
      int gc = std::min(gchunk, G - g0);
      const Tin* src;
      int64_t ld_in;
      if (on_dev) {
        src = Ysrc + (int64_t)g0 * ld;
        ld_in = ld;
      } else {
        CUDA_OK(cudaMemcpy2DAsync(stage, sizeof(Tin) * N, Ysrc + (int64_t)g0 * ld, sizeof(Tin) * ld, sizeof(Tin) * N, gc,
                                  cudaMemcpyHostToDevice, h->stream));
        src = stage;
        ld_in = N;
      }
      dim3 grid((unsigned)ceil_div64(N, 32), (gc + 31) / 32), blk(32, 8);
      CA_LAUNCH(k_ingest_colmajor<Tin>, grid, blk, 0, h->stream)(src, ld_in, N, g0, gc, Yf, h->ldY);
      KCHECK();
      CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    if (stage) CUDA_OK(cudaFree(stage));
  } else {
    int64_t ld = c.y_ld ? c.y_ld : G;
    int64_t rchunk = std::max<int64_t>(1, std::min<int64_t>(N, (int64_t)(256ll << 20) / (int64_t)(sizeof(Tin) * ld)));
    rchunk = std::min<int64_t>(rchunk, 65535);
    Tin* stage = nullptr;
    if (!on_dev) CUDA_OK(cudaMalloc(&stage, sizeof(Tin) * (size_t)rchunk * ld));
    for (int64_t r0 = 0; r0 < N; r0 += rchunk) {
      int64_t rc = std::min(rchunk, N - r0);
      const Tin* src;
      if (on_dev) {
        src = Ysrc + r0 * ld;
      } else {
        CUDA_OK(cudaMemcpyAsync(stage, Ysrc + r0 * ld, sizeof(Tin) * (size_t)rc * ld, cudaMemcpyHostToDevice, h->stream));
        src = stage;
      }
      dim3 grid(std::min((G + 255) / 256, 64), (unsigned)rc);
      CA_LAUNCH(k_ingest_rowmajor<Tin>, grid, 256, 0, h->stream)(src, ld, rc, G, Yf + r0 * h->ldY, h->ldY);
      KCHECK();
      CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    if (stage) CUDA_OK(cudaFree(stage));
  }
}

void destroy(ca_handle* h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm) comm_release(h->cfg.world, h->cfg.rank, h->dev, h->comm);   // parked for the next session of this shape
  for (int r = 0; r < kP2PMaxWorld; ++r)
    if (h->p2p_mapped[r]) cudaIpcCloseMemHandle(h->p2p_mapped[r]);
#ifndef CA_EMULATE
  for (auto* g : {&h->g_train[0], &h->g_train[1], &h->g_eval[0], &h->g_eval[1]})
    if (*g) { cudaGraphExecDestroy(*g); *g = nullptr; }
#endif
  tc_plan_destroy(h->tcplan);
  for (auto& p : h->prof) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (void* p : h->allocs)
    if (p) cudaFree(p);
  if (h->shared) h->shared->refs--;
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

void build(ca_handle* h, const void* Y, const double* L, const double* psi_init, const double* loc_init, const double* X,
           const double* colsum_total, const double* clone_allele, const double* alt, const double* cov) {
  const ca_config& c = h->cfg;
  if (c.N <= 0 || c.G <= 0 || c.C <= 0 || c.S <= 0 || c.K < 0 || c.P < 0) fail("bad dimensions");
  if (c.K + c.P > kMaxKP) fail("K + P = %d exceeds the supported maximum of %d", c.K + c.P, kMaxKP);
  if (c.world < 1 || c.rank < 0 || c.rank >= c.world) fail("bad rank/world");
  if (c.world > 1 && !c.nccl_id) fail("world > 1 requires cfg.nccl_id");
  if (!h->shared && (!Y || !L)) fail("missing input pointer");
  if (!h->data_only && (!loc_init || (c.K > 0 && !psi_init) || (c.P > 0 && !X))) fail("missing input pointer");
  if (!h->shared && c.V > 0 && (!clone_allele || !alt || !cov)) fail("V > 0 requires clone_allele, alt and cov");
  if (h->shared) {
    const ca_data* d = h->shared;
    if (c.world != 1) fail("shared inputs are for single-shard sessions (world == 1)");
    if (c.N != d->N || c.G != d->G || c.C != d->C || c.V != d->V || c.device != d->dev)
      fail("session dimensions / device do not match the shared inputs (N %lld G %d C %d V %d device %d)", (long long)d->N, d->G, d->C,
           d->V, d->dev);
  }
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (c.device < 0 || c.device >= ndev) fail("CUDA device %d not available (%d devices)", c.device, ndev);
  h->dev = c.device;
  CUDA_OK(cudaSetDevice(h->dev));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, h->dev));
  if (prop.major != 10) fail("clonealign_b200 kernels are built for sm_100a only; device %d is sm_%d%d", h->dev, prop.major, prop.minor);
  h->num_sms = prop.multiProcessorCount;
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (c.world > 1) h->comm = comm_acquire(c.world, c.rank, h->dev, c.nccl_id);   // collective (first session of this shape)
  // Measured on B200 (profiles/r01_notes.md): co-scheduling the Y stream with the forward contraction does not pay
  // yet (the register-light Y kernel is slower than the saved time), so the fork is opt-in.
  h->overlap = getenv("CLONEALIGN_B200_OVERLAP") != nullptr || (c.variants & CA_VAR_OVERLAP);
  h->use_graph = getenv("CLONEALIGN_B200_NO_GRAPH") == nullptr;

  h->N = c.N; h->Ntot = c.N_total > 0 ? c.N_total : c.N; h->G = c.G; h->C = c.C; h->S = c.S; h->K = c.K; h->P = c.P;
  h->KP = c.K + c.P; h->SC = c.S * c.C; h->V = c.V;
  bool tc_ok = kTcAvailable && (c.K == 1 && c.P == 0 && round_up64(h->SC, 16) <= 128);
  if (c.path == CA_PATH_TENSOR && !tc_ok) fail("tensor path needs K == 1, P == 0 and S*C <= 128");
  if (c.path == CA_PATH_INTERP && !(c.K == 1 && c.P == 0)) fail("interp path needs K == 1 and P == 0");
  // path = auto: the reference's default model (K = 1, no covariates; K is forced to 1 at R/clonealign.R:226-232) runs the
  // univariate-interpolation kernel set that round 2 validated on hardware (profiles/r02_notes.md): interp + bulk-copy Y
  // pass on the stored integers, co-scheduled with the rest of the step + fused per-cell kernel + fused gene-level
  // launches + late join of the Y pass.  Explicit variant bits of the caller are kept (ypass2 / ypass3, overlap, p2p).  Other shapes: tcgen05 contractions
  // (K = 1, S*C <= 128) or the CUDA-core kernels (any K + P <= 8).
  const bool interp_ok = c.K == 1 && c.P == 0 && c.C <= kFusedMaxC && c.S * c.C <= 32 * kFusedMaxNJ;
  if (c.path == CA_PATH_AUTO && interp_ok) {
    h->cfg.path = CA_PATH_INTERP;
    if (!(h->cfg.variants & (CA_VAR_YPASS2 | CA_VAR_YPASS3 | CA_VAR_YPASS4))) h->cfg.variants |= CA_VAR_YPASS4;
    h->cfg.variants |= CA_VAR_EPI2 | CA_VAR_LEAN | CA_VAR_DEFER;
    if (h->cfg.variants & CA_VAR_YPASS4) h->cfg.variants |= CA_VAR_COSCHED;
  }
  h->interp = (c.path == CA_PATH_INTERP);
  h->variants = c.variants;
  if (c.variants & ~(uint32_t)(CA_VAR_YPASS2 | CA_VAR_EPI2 | CA_VAR_LEAN | CA_VAR_P2P | CA_VAR_OVERLAP | CA_VAR_YPASS3 | CA_VAR_DEFER | CA_VAR_YPASS4 | CA_VAR_COSCHED))
    fail("unknown kernel variant bits 0x%x", c.variants);
  if ((c.variants & CA_VAR_P2P) && c.world > kP2PMaxWorld) fail("variant p2p supports at most %d ranks", kP2PMaxWorld);
  h->p2p = (c.variants & CA_VAR_P2P) && c.world > 1;
  if ((c.variants & CA_VAR_LEAN) && !(c.variants & CA_VAR_EPI2)) fail("variant lean needs variant epi2");
  h->lean = (c.variants & CA_VAR_LEAN) != 0;
  if ((c.variants & CA_VAR_DEFER) && !(c.variants & CA_VAR_LEAN)) fail("variant defer needs variants epi2 and lean");
  h->defer = (c.variants & CA_VAR_DEFER) != 0;
  if ((c.variants & CA_VAR_YPASS2) && c.K + c.P != 1) fail("variant ypass2 needs K + P == 1");
  if ((c.variants & CA_VAR_YPASS3) && c.K + c.P != 1) fail("variant ypass3 needs K + P == 1");
  if ((c.variants & CA_VAR_YPASS3) && (c.variants & CA_VAR_YPASS2)) fail("variants ypass2 and ypass3 are alternatives");
  if ((c.variants & CA_VAR_YPASS4) && c.K + c.P != 1) fail("variant ypass4 needs K + P == 1");
  if ((c.variants & CA_VAR_YPASS4) && (c.variants & (CA_VAR_YPASS2 | CA_VAR_YPASS3))) fail("variants ypass2, ypass3 and ypass4 are alternatives");
  if ((c.variants & CA_VAR_COSCHED) && !((c.variants & CA_VAR_DEFER) && (c.variants & CA_VAR_YPASS4)))
    fail("variant cosched needs variants defer and ypass4");
  h->cosched = (c.variants & CA_VAR_COSCHED) != 0;
  if (c.variants & CA_VAR_EPI2) {
    if (!h->interp) fail("variant epi2 belongs to the interp path (path = interp)");
    if (c.C > kFusedMaxC || c.S * c.C > 32 * kFusedMaxNJ) fail("variant epi2 needs C <= %d and S*C <= %d", kFusedMaxC, 32 * kFusedMaxNJ);
    h->epi2 = true;
  }
  h->tc = !h->interp && ((c.path == CA_PATH_TENSOR) || (c.path == CA_PATH_AUTO && tc_ok));
  h->SCp = h->tc ? (int)round_up64(h->SC, 16) : h->SC;
  h->J = h->SCp * (1 + h->KP);
  h->ldY = round_up64(h->G, 16);
  h->Gld = round_up64(h->G, 64);
  h->Nld = round_up64(h->N, 64);
  const int64_t N = h->N;
  const int G = h->G, C = h->C, S = h->S, K = h->K, KP = h->KP, J = h->J;

  if (h->shared) {
    const ca_data* d = h->shared;
    h->Y = d->Y; h->ystore = d->ystore; h->L = d->L; h->Bm = d->Bm; h->vA = d->vA; h->s = d->s; h->colsum = d->colsum;
    h->snv = d->snv; h->const_sum = d->const_sum; h->poison = d->poison;
    if (h->ldY != d->ldY) fail("shared inputs: leading dimension mismatch");
  } else {
  // ---- Y -> device fp32 [N][ldY] ----
  float* Yf = h->alloc<float>((size_t)N * h->ldY);
  switch (c.y_dtype) {
    case CA_Y_F64: ingest_y<double>(h, (const double*)Y, Yf); break;
    case CA_Y_F32: ingest_y<float>(h, (const float*)Y, Yf); break;
    case CA_Y_I32: ingest_y<int>(h, (const int*)Y, Yf); break;
    case CA_Y_U8: ingest_y<uint8_t>(h, (const uint8_t*)Y, Yf); break;
    case CA_Y_U16: ingest_y<uint16_t>(h, (const uint16_t*)Y, Yf); break;
    default: fail("bad y_dtype");
  }
  // ---- narrow storage if exact ----
  int* flags = h->alloc<int>(1);
  {
    dim3 grid(std::min((G + 255) / 256, 64), 1);
    // grid.y is limited to 65535: loop over row chunks
    for (int64_t r0 = 0; r0 < N; r0 += 65535) {
      grid.y = (unsigned)std::min<int64_t>(65535, N - r0);
      CA_LAUNCH(k_scan_y, grid, 256, 0, h->stream)(Yf + r0 * h->ldY, h->ldY, grid.y, G, flags);
      KCHECK();
    }
  }
  int hflags = 0;
  CUDA_OK(cudaMemcpyAsync(&hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  int want = c.y_store;
  if (want == CA_STORE_AUTO) want = (hflags & 1) ? CA_STORE_F32 : ((hflags & 2) ? ((hflags & 4) ? CA_STORE_F32 : CA_STORE_U16) : CA_STORE_U8);
  if (want == CA_STORE_U8 && (hflags & 3)) fail("y_store = u8 requested but Y has non-integer, negative or > 255 entries");
  if (want == CA_STORE_U16 && (hflags & 5)) fail("y_store = u16 requested but Y has non-integer, negative or > 65535 entries");
  h->ystore = want;
  h->Y = Yf;

  // ---- small inputs ----
  h->L = h->alloc<float>((size_t)G * C);
  upload_colmajor(h, L, G, C, h->L, C, 0);
  std::vector<float> logL((size_t)G * C);
  for (int g = 0; g < G; ++g)
    for (int cc = 0; cc < C; ++cc) {
      double l = L[(size_t)cc * G + g];
      if (!(l > 0.0)) h->poison = 1;   // copy number 0 => 0*log(0) = NaN in the reference
      logL[(size_t)g * C + cc] = l > 0.0 ? (float)log(l) : 0.f;
    }
  float* d_logL = h->alloc<float>((size_t)G * C);
  CUDA_OK(cudaMemcpyAsync(d_logL, logL.data(), sizeof(float) * G * C, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));

  h->Bm = h->alloc<float>((size_t)N * C);
  h->vA = h->alloc<float>((size_t)N * C);
  h->s = h->alloc<float>(N);
  h->colsum = h->alloc<float>(G);
  double* cst = h->alloc<double>(N);
  CA_LAUNCH(k_setup_rows<float>, (unsigned)N, 256, 0, h->stream)(Yf, h->ldY, N, G, C, d_logL, h->s, cst, h->Bm);
  KCHECK();
  {
    double* csum = h->alloc<double>(1);
    CA_LAUNCH(k_reduce_partials, 1, 1024, 0, h->stream)(cst, N, 1, csum, 0.0);
    KCHECK();
    CUDA_OK(cudaMemcpyAsync(&h->const_sum, csum, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->release(csum);
  }
  h->release(cst);
  h->release(d_logL);
  if (colsum_total) {
    std::vector<float> cs(G);
    for (int g = 0; g < G; ++g) cs[g] = (float)colsum_total[g];
    CUDA_OK(cudaMemcpyAsync(h->colsum, cs.data(), sizeof(float) * G, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  } else {
    // colSums(Y) (R/inference-tflow.R:117) over this shard in fp64; under cell sharding the shards' sums are added
    // with one all-reduce (collective: every rank passes colsum_total == NULL or none does)
    const int RS = 64;
    double* part = h->alloc<double>((size_t)RS * G);
    double* tot = h->alloc<double>(G);
    dim3 grid((G + 127) / 128, RS);
    CA_LAUNCH(k_colsum_part<float>, grid, 128, 0, h->stream)(Yf, h->ldY, N, G, RS, part);
    KCHECK();
    CA_LAUNCH(k_colsum_final, (G + 127) / 128, 128, 0, h->stream)(part, RS, G, h->colsum, tot);
    KCHECK();
    if (c.world > 1) {
      NCCL_OK(nccl().AllReduce(tot, tot, (size_t)G, kNcclFloat64, kNcclSum, h->comm, h->stream));
      CA_LAUNCH(k_colsum_final, (G + 127) / 128, 128, 0, h->stream)(tot, 1, G, h->colsum, nullptr);
      KCHECK();
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->release(part);
    h->release(tot);
  }
  if (c.V > 0) {
    int V = c.V;
    float* d_alt = h->alloc<float>((size_t)N * V);
    float* d_cov = h->alloc<float>((size_t)N * V);
    float* d_cn = h->alloc<float>((size_t)V * C);
    upload_colmajor(h, alt, N, V, d_alt, V, 0);
    upload_colmajor(h, cov, N, V, d_cov, V, 0);
    upload_colmajor(h, clone_allele, V, C, d_cn, C, 0);
    CA_LAUNCH(k_allele, (unsigned)N, 128, 0, h->stream)(d_alt, d_cov, d_cn, N, V, C, h->vA);
    KCHECK();
    h->snv = h->alloc<float>((size_t)N * C);
    CA_LAUNCH(k_softmax_rows, (unsigned)ceil_div64(N, 128), 128, 0, h->stream)(h->vA, N, C, h->snv);
    KCHECK();
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->release(d_alt);
    h->release(d_cov);
    h->release(d_cn);
  }
  // narrow Y after the setup passes that read it as fp32
  if (h->ystore != CA_STORE_F32) {
    dim3 grid(std::min<int64_t>((h->ldY + 255) / 256, 64), 1);
    void* Yn = nullptr;
    if (h->ystore == CA_STORE_U16) Yn = h->alloc<uint16_t>((size_t)N * h->ldY, false);
    else Yn = h->alloc<uint8_t>((size_t)N * h->ldY, false);
    for (int64_t r0 = 0; r0 < N; r0 += 65535) {
      grid.y = (unsigned)std::min<int64_t>(65535, N - r0);
      if (h->ystore == CA_STORE_U16) CA_LAUNCH(k_narrow_y<uint16_t>, grid, 256, 0, h->stream)(Yf + r0 * h->ldY, h->ldY, grid.y, (uint16_t*)Yn + r0 * h->ldY);
      else CA_LAUNCH(k_narrow_y<uint8_t>, grid, 256, 0, h->stream)(Yf + r0 * h->ldY, h->ldY, grid.y, (uint8_t*)Yn + r0 * h->ldY);
      KCHECK();
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->release(Yf);
    h->Y = Yn;
  }
  }   // !shared
  if (h->data_only) {
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return;
  }

  // ---- parameters (R/inference-tflow.R:240-272) ----
  h->dstate = h->alloc<StepState>(1);
  auto z = [&](size_t n) { return h->alloc<float>(n); };
  h->U = z((size_t)N * KP + 64); h->m_U = z((size_t)N * KP); h->v_U = z((size_t)N * KP); h->g_U = z((size_t)N * KP);
  h->Vm = z((size_t)G * KP + 64); h->m_V = z((size_t)G * KP); h->v_V = z((size_t)G * KP); h->g_V = z((size_t)G * KP);
  h->chi_raw = z(K); h->m_chi = z(K); h->v_chi = z(K); h->g_chi = z(K);
  h->u = z(C); h->m_u = z(C); h->v_u = z(C); h->g_u = z(C);
  h->loc = z(G); h->m_loc = z(G); h->v_loc = z(G); h->g_loc = z(G);
  h->lsd = z(G); h->m_lsd = z(G); h->v_lsd = z(G); h->g_lsd = z(G);
  h->t = z((size_t)N * C); h->m_t = z((size_t)N * C); h->v_t = z((size_t)N * C); h->g_t = z((size_t)N * C);
  if (K > 0) upload_colmajor(h, psi_init, N, K, h->U, KP, 0);
  if (c.P > 0) upload_colmajor(h, X, N, c.P, h->U, KP, K);
  {
    std::vector<float> lf(G);
    for (int g = 0; g < G; ++g) lf[g] = (float)loc_init[g];
    CUDA_OK(cudaMemcpyAsync(h->loc, lf.data(), sizeof(float) * G, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }

  // ---- scratch ----
  h->eps_in = z((size_t)S * G); h->eps = z((size_t)S * G); h->mu = z((size_t)S * G); h->logmu = z((size_t)S * G); h->sig = z((size_t)S * G);
  h->shift = z(N + 64); h->mm = z(2); h->log_alpha = z(C);
  h->YV = z((size_t)N * std::max(KP, 1)); h->YtU = z((size_t)G * std::max(KP, 1)); h->Fout = z((size_t)N * C);
  h->dM_sum = z((size_t)G * J);
  h->n_gene_blocks = (int)ceil_div64((int64_t)G * S, h->lean ? kProThreads : 256);   // one thread per (sample, gene) pair
  // fused prologue: two 512-thread blocks fit an SM; 32 of the slots go to its scalar / range blocks, the gene blocks stride
  if (h->lean) h->n_gene_blocks = std::min(h->n_gene_blocks, std::max(1, 2 * h->num_sms - 2 - kProPsiBlocks));
  h->n_epi_blocks = ceil_div64(N, kEpiWarps);
  h->n_cell_parts = h->epi2 ? (int64_t)h->num_sms : h->n_epi_blocks;
  h->gene_part = h->alloc<double>(h->n_gene_blocks);
  h->n_yv_blocks = h->defer ? h->num_sms : 0;
  h->elbo_part = h->alloc<double>(h->n_cell_parts + h->n_yv_blocks);
  h->gsum_part = h->alloc<double>((size_t)h->n_cell_parts * C);
  h->scal_elbo = h->alloc<double>(1); h->cell_sum = h->alloc<double>(1); h->wsq = h->alloc<double>(std::max(K, 1));
  h->elbo_dev = h->alloc<double>(1);
  h->ar = z((size_t)G * (2 + KP) + C + 4);
  if (KP == 1) {
    int tile_cols = kYCB;
    if (h->variants & (CA_VAR_YPASS3 | CA_VAR_YPASS4))   // column tile of k_ypass_k1_v3 / v4: 256 threads x the columns a thread owns for this storage type
      tile_cols = h->ystore == CA_STORE_U8 ? ypass3_tile_cols<uint8_t>() : (h->ystore == CA_STORE_U16 ? ypass3_tile_cols<uint16_t>() : ypass3_tile_cols<float>());
    h->nCB = (int)ceil_div64(h->ldY, tile_cols);
    h->RB = 512;
    if (h->variants & (CA_VAR_YPASS3 | CA_VAR_YPASS4)) {
      // Size the row blocks so that the grid is (just under) a whole number of waves of the 2 CTAs an SM holds: with
      // 512-row blocks config 3 gives 5 x 196 = 980 CTAs = 3.31 waves of 296, i.e. a last wave that is one third full
      // on the kernel that bounds the step; 432-row blocks give 5 x 232 = 1160 CTAs = 3.92 waves.  Small problems get
      // enough row blocks to cover every SM (10k x 5k: 250 CTAs instead of 40).  Any multiple of 16 rows works (vector
      // loads of psi, 8 / 16 rows in flight).
      const int64_t slots = 2 * (int64_t)h->num_sms;
      const int64_t waves = std::max<int64_t>(1, ceil_div64((int64_t)h->nCB * ceil_div64(N, 512), slots));
      const int64_t nrb = std::max<int64_t>(1, waves * slots / h->nCB);
      h->RB = (int)std::min<int64_t>(1 << 20, std::max<int64_t>(16, round_up64(ceil_div64(N, nrb), 16)));
    }
  } else {
    h->nCB = 1;
    h->RB = 1024;
  }
  h->nRB = (int)ceil_div64(N, h->RB);
  h->rowpart = z((size_t)h->nCB * N * std::max(KP, 1));
  h->colpart = z((size_t)h->nRB * G * std::max(KP, 1));
  if (h->tc) {
    h->MxT_hi = h->alloc<__nv_bfloat16>((size_t)J * h->Gld);
    h->MxT_lo = h->alloc<__nv_bfloat16>((size_t)J * h->Gld);
    h->RxT = h->alloc<__half>((size_t)J * h->Nld);
    h->shift_bwd = z((size_t)h->Nld);
    tc_plan_create(h->tcplan, h->dev, N, h->Nld, G, h->Gld, h->SCp, J, h->MxT_hi, h->MxT_lo, h->RxT);
    // The overlapped Y pass must be able to share an SM with a contraction CTA (211 KB of shared memory): give it
    // the same (maximum) shared-memory carveout, otherwise the SM has to drain before it can be reconfigured.
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_persistent<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_persistent<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_persistent<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    h->nsplit = h->tcplan.nsplit;
    h->Zx = z((size_t)h->tcplan.fsplit * N * J);
    h->dMx = z((size_t)h->nsplit * G * J);
  } else {
    h->Mx = z((size_t)G * J);
    h->Zx = z((size_t)N * J);
    h->Rx = z((size_t)N * J);
    h->dMx = z((size_t)G * J);
    h->nsplit = 1;
  }
  if (h->interp) {
    h->iplan = h->alloc<InterpPlan>(1);
    h->mm_psi = z(2);
    // node-sum kernel: columns per thread, column groups per warp, slices of the reduction index (>= 4 staged chunks per
    // work item, at most three items per SM and panel: with one active panel every resident block still has work), dynamic shared memory
    h->n2_tj = n2_pick_tj(J);
    h->n2_ncgp = n2_ncg_pow2(J, h->n2_tj);
    auto n2_split = [&](int64_t R) {
      return (int)std::max<int64_t>(1, std::min<int64_t>(kN2BlocksPerSM * (int64_t)h->num_sms, ceil_div64(ceil_div64(R, kN2Chunk), 4)));
    };
    h->n2_split_f = n2_split(G);
    h->n2_split_b = n2_split(N);
    h->n2_smem = n2_smem_bytes(J, h->n2_tj);
    h->n2_blocks_per_sm = (int)std::max<size_t>(1, std::min<size_t>(kN2BlocksPerSM, (220 * 1024) / h->n2_smem));
    if (h->n2_smem > 48 * 1024) {
      if (h->n2_tj == 8) {
        CUDA_OK(cudaFuncSetAttribute(k_interp_nodes2<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->n2_smem));
        CUDA_OK(cudaFuncSetAttribute(k_interp_nodes2<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->n2_smem));
      } else {
        CUDA_OK(cudaFuncSetAttribute(k_interp_nodes2<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->n2_smem));
        CUDA_OK(cudaFuncSetAttribute(k_interp_nodes2<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->n2_smem));
      }
    }
    const size_t nodes_f = (size_t)h->n2_split_f * kIMaxPanF * kIP, nodes_b = (size_t)h->n2_split_b * kIMaxPanB * kIP;
    h->ivals = h->alloc<double>(std::max(nodes_f, nodes_b) * J, false);
    h->icoef = h->alloc<double>((size_t)std::max(kIMaxPanF, kIMaxPanB) * kIP * J);
    const size_t per_panel = (size_t)kIP * J * sizeof(double);
    h->ieval_panels = (int)std::min<size_t>(16, (200 * 1024) / per_panel);
    h->ieval_smem = (size_t)h->ieval_panels * per_panel;
    if (h->ieval_smem > 48 * 1024) {
      CUDA_OK(cudaFuncSetAttribute(k_interp_eval<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ieval_smem));
      CUDA_OK(cudaFuncSetAttribute(k_interp_eval<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ieval_smem));
    }
  }
  if (h->variants & CA_VAR_YPASS4) {
    if (const char* e = getenv("CLONEALIGN_B200_Y4_MINB")) h->y4_minb = atoi(e) == 3 ? 3 : 4;
    auto set4 = [&](auto kern, size_t bytes) { CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); };
    set4(k_ypass_k1_v4<float, 3>, ypass4_smem_bytes<float>()); set4(k_ypass_k1_v4<float, 4>, ypass4_smem_bytes<float>());
    set4(k_ypass_k1_v4<uint16_t, 3>, ypass4_smem_bytes<uint16_t>()); set4(k_ypass_k1_v4<uint16_t, 4>, ypass4_smem_bytes<uint16_t>());
    set4(k_ypass_k1_v4<uint8_t, 3>, ypass4_smem_bytes<uint8_t>()); set4(k_ypass_k1_v4<uint8_t, 4>, ypass4_smem_bytes<uint8_t>());
  }
  if (h->lean) {
    h->chi_cur = h->alloc<double>(std::max(K, 1));
    h->pmm_part = h->alloc<float>(2 * kProPsiBlocks);
    h->ticket = h->alloc<unsigned>(1);
    h->gene_panels = gene_fused_smem_panels(J);
    if (const char* e = getenv("CLONEALIGN_B200_FUSED_PANELS")) h->gene_panels = std::max(0, std::min(h->gene_panels, atoi(e)));
    h->gene_smem = gene_fused_smem_bytes(J, h->gene_panels);
    if (h->gene_smem > 48 * 1024)
    {
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene_smem));
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene_smem));
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene_smem));
      CUDA_OK(cudaFuncSetAttribute(k_gene_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gene_smem));
    }
  }
  if (h->epi2) {
    h->fused_nj = (h->SC + 31) / 32;
    // defer + overlap: 16 warps x 64 registers = half of the register file, so that one Y-pass CTA (256 threads x 128
    // registers, the other half) can be resident on the same SM while the per-cell kernel runs
    h->fused_warps = (h->defer && (c.variants & (CA_VAR_OVERLAP | CA_VAR_COSCHED))) ? kFusedWarps / 2 : kFusedWarps;
    if (h->cosched && h->y4_minb == 3) h->fused_warps = 12;     // 2 x 256 x 80 registers for the stream leave 24 K of the 64 K
    if (const char* e = getenv("CLONEALIGN_B200_FUSED_WARPS")) h->fused_warps = std::max(1, std::min(kFusedWarps, atoi(e)));
    if (h->fused_warps != kFusedWarps) {
      // the Y-pass CTA must fit next to ~200 KB of shared memory: ask for the maximum shared-memory carveout, otherwise the
      // SM would have to drain before it can be reconfigured (measured in round 1 for the contraction kernels)
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v3<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v3<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v3<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v2<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v2<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CUDA_OK(cudaFuncSetAttribute(k_ypass_k1_v2<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    // cosched: two persistent Y-pass CTAs (64 KB rings) stay resident on every SM; the per-cell CTA gets what is left
    const size_t fused_budget = h->cosched ? 92 * 1024 : 200 * 1024;
    h->fused_panels = fused_smem_panels(h->SC, C, J, fused_budget, h->fused_warps);
    if (const char* e = getenv("CLONEALIGN_B200_FUSED_PANELS"))   // test hook: force the coefficients-through-L2 branch
      h->fused_panels = std::max(0, std::min(h->fused_panels, atoi(e)));
    h->fused_smem = fused_smem_bytes(h->SC, C, J, h->fused_panels, h->fused_warps);
    if (h->fused_smem > 48 * 1024) {
      switch (h->fused_nj) {
        case 1: fused_set_smem<1>(h->fused_smem); break;
        case 2: fused_set_smem<2>(h->fused_smem); break;
        case 3: fused_set_smem<3>(h->fused_smem); break;
        default: fused_set_smem<4>(h->fused_smem); break;
      }
    }
  }
  size_t smem = epi_smem_bytes(h->SCp, C, J, h->tc);
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) fail("S*C too large for the per-cell epilogue (%zu bytes of shared memory)", smem);
    CUDA_OK(cudaFuncSetAttribute(k_cell_epilogue<EPI_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(k_cell_epilogue<EPI_EVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(k_cell_epilogue<EPI_INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (h->p2p) {
    h->p2p_cnt = (int64_t)G * (2 + KP) + C;
    h->p2p_cnt_pad = round_up64(h->p2p_cnt, 4);
    const size_t slot_bytes = sizeof(float) * 2 * (size_t)c.world * h->p2p_cnt_pad;
    // one allocation (one IPC handle): slots, then the flags on their own 256-byte line
    h->p2p_buf = (float*)h->alloc<unsigned char>(slot_bytes + 256 + sizeof(unsigned) * 2 * kP2PMaxWorld);
    h->p2p_ticket = h->alloc<unsigned>(1);
    h->p2p_err = h->alloc<int>(1);
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
}

struct ArrayRef {
  const float* p;
  int64_t rows;
  int cols, ld, off;
  bool writable;
};

bool lookup(ca_handle* h, const std::string& n, ArrayRef& r) {
  const int64_t N = h->N;
  const int G = h->G, C = h->C, K = h->K, P = h->P, KP = h->KP, SC = h->SC, J = h->J;
  auto set = [&](const float* p, int64_t rows, int cols, int ld, int off, bool w) { r = {p, rows, cols, ld, off, w}; return true; };
  if (n == "W") return set(h->Vm, G, K, KP, 0, true);
  if (n == "beta") return set(h->Vm, G, P, KP, K, true);
  if (n == "psi") return set(h->U, N, K, KP, 0, true);
  if (n == "chi_raw") return set(h->chi_raw, K, 1, 1, 0, true);
  if (n == "alpha_unconstr") return set(h->u, C, 1, 1, 0, true);
  if (n == "loc") return set(h->loc, G, 1, 1, 0, true);
  if (n == "lsd") return set(h->lsd, G, 1, 1, 0, true);
  if (n == "gamma_logits") return set(h->t, N, C, C, 0, true);
  if (n == "grad_W") return set(h->g_V, G, K, KP, 0, false);
  if (n == "grad_beta") return set(h->g_V, G, P, KP, K, false);
  if (n == "grad_psi") return set(h->g_U, N, K, KP, 0, false);
  if (n == "grad_chi_raw") return set(h->g_chi, K, 1, 1, 0, false);
  if (n == "grad_alpha_unconstr") return set(h->g_u, C, 1, 1, 0, false);
  if (n == "grad_loc") return set(h->g_loc, G, 1, 1, 0, false);
  if (n == "grad_lsd") return set(h->g_lsd, G, 1, 1, 0, false);
  if (n == "grad_gamma_logits") return set(h->g_t, N, C, C, 0, false);
  if (n == "Z") return set(h->Zx, N, SC, J, 0, false);
  if (n == "Zx") return set(h->Zx, N, J, J, 0, false);
  if (n == "R" && h->Rx) return set(h->Rx, N, SC, J, 0, false);
  if (n == "dM") return set(h->dM_sum, G, SC, J, 0, false);
  if (n == "dMx") return set(h->dM_sum, G, J, J, 0, false);
  if (n == "F") return set(h->Fout, N, C, C, 0, false);
  if (n == "YV") return set(h->YV, N, KP, KP, 0, false);
  if (n == "YtU") return set(h->YtU, G, KP, KP, 0, false);
  if (n == "B") return set(h->Bm, N, C, C, 0, false);
  if (n == "v") return set(h->vA, N, C, C, 0, false);
  if (n == "s") return set(h->s, N, 1, 1, 0, false);
  if (n == "colsum") return set(h->colsum, G, 1, 1, 0, false);
  if (n == "shift") return set(h->shift, N, 1, 1, 0, false);
  if (n == "mu_samples") return set(h->mu, h->S, G, G, 0, false);   // NOTE: returned as S x G column-major
  return false;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int ca_core_abi_version(void) { return CA_ABI_VERSION; }

int ca_core_shutdown(void) {
  std::lock_guard<std::mutex> lk(comm_mu());
  for (auto& kv : comm_pool())
    for (void* c : kv.second) {
      cudaSetDevice(kv.first.dev);
      nccl().CommDestroy(c);
    }
  comm_pool().clear();
  return 0;
}

int ca_core_device_count(int* count, char* err, size_t errlen) {
  try {
    int n = 0;
    CUDA_OK(cudaGetDeviceCount(&n));
    if (count) *count = n;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_nccl_unique_id(void* out128, char* err, size_t errlen) {
  try {
    if (!out128) fail("null output");
    NCCL_OK(nccl().GetUniqueId(out128));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_create(ca_handle** out, const ca_config* cfg, const void* Y, const double* L, const double* psi_init,
                   const double* loc_init, const double* X, const double* colsum_total, const double* clone_allele,
                   const double* alt, const double* cov, char* err, size_t errlen) {
  ca_handle* h = nullptr;
  try {
    if (!out || !cfg) fail("null argument");
    h = new ca_handle();
    h->cfg = *cfg;
    build(h, Y, L, psi_init, loc_init, X, colsum_total, clone_allele, alt, cov);
    h->cfg.nccl_id = nullptr;   // never retain caller pointers
    h->cfg.y_indptr = nullptr;
    h->cfg.y_indices = nullptr;
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    destroy(h);
    return report(e, err, errlen);
  }
}

int ca_core_data_create(ca_data** out, const ca_config* cfg, const void* Y, const double* L, const double* colsum_total,
                        const double* clone_allele, const double* alt, const double* cov, char* err, size_t errlen) {
  ca_handle* t = nullptr;
  try {
    if (!out || !cfg) fail("null argument");
    t = new ca_handle();
    t->cfg = *cfg;
    t->cfg.S = std::max(1, cfg->S);
    t->cfg.K = 0; t->cfg.P = 0; t->cfg.path = CA_PATH_CUDACORE; t->cfg.variants = 0; t->cfg.world = 1; t->cfg.rank = 0;
    t->data_only = true;
    build(t, Y, L, nullptr, nullptr, nullptr, colsum_total, clone_allele, alt, cov);
    ca_data* d = new ca_data();
    d->dev = t->dev; d->N = t->N; d->ldY = t->ldY; d->G = t->G; d->C = t->C; d->V = t->V; d->ystore = t->ystore;
    d->poison = t->poison; d->const_sum = t->const_sum;
    d->Y = t->Y; d->L = t->L; d->Bm = t->Bm; d->vA = t->vA; d->s = t->s; d->colsum = t->colsum; d->snv = t->snv;
    for (void* p : t->allocs)
      if (p) d->allocs.push_back(p);
    t->allocs.clear();     // ownership moved
    destroy(t);
    *out = d;
    return 0;
  } catch (const std::exception& e) {
    destroy(t);
    return report(e, err, errlen);
  }
}

int ca_core_data_destroy(ca_data* d, char* err, size_t errlen) {
  try {
    if (!d) return 0;
    if (d->refs.load() > 0) fail("ca_core_data_destroy: %d session(s) still use these inputs", d->refs.load());
    cudaSetDevice(d->dev);
    for (void* p : d->allocs) cudaFree(p);
    delete d;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_data_stats(ca_data* d, double* rowsum, double* colsum, double* mu_guess, char* err, size_t errlen) {
  try {
    if (!d) fail("null argument");
    CUDA_OK(cudaSetDevice(d->dev));
    const int64_t N = d->N;
    const int G = d->G;
    const int RS = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / 64));
    double *d_row = nullptr, *d_part = nullptr, *d_sums = nullptr;
    struct Guard {
      double*& a; double*& b; double*& c;
      ~Guard() { cudaFree(a); cudaFree(b); cudaFree(c); }
    } guard{d_row, d_part, d_sums};
    CUDA_OK(cudaMalloc(&d_row, sizeof(double) * N));
    CUDA_OK(cudaMalloc(&d_part, sizeof(double) * (size_t)RS * G * 2));
    CUDA_OK(cudaMalloc(&d_sums, sizeof(double) * (size_t)G * 2));
    cudaStream_t st = nullptr;   // the legacy default stream: the inputs are immutable and nothing else is in flight on them
    auto run = [&](auto* Yp) {
      using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
      CA_LAUNCH(k_stats_rows<T>, (unsigned)ceil_div64(N, 8), 256, 0, st)(Yp, d->ldY, N, G, d_row);
      KCHECK();
      CA_LAUNCH(k_stats_cols<T>, dim3((G + 127) / 128, RS), 128, 0, st)(Yp, d->ldY, N, G, RS, d_row, d_part);
      KCHECK();
    };
    switch (d->ystore) {
      case CA_STORE_F32: run((const float*)d->Y); break;
      case CA_STORE_U16: run((const uint16_t*)d->Y); break;
      case CA_STORE_U8: run((const uint8_t*)d->Y); break;
      default: fail("bad y_store");
    }
    CA_LAUNCH(k_pca_colstats_reduce, (G + 127) / 128, 128, 0, st)(d_part, RS, G, d_sums);
    KCHECK();
    std::vector<double> hs((size_t)2 * G);
    CUDA_OK(cudaMemcpyAsync(hs.data(), d_sums, sizeof(double) * 2 * G, cudaMemcpyDeviceToHost, st));
    if (rowsum) CUDA_OK(cudaMemcpyAsync(rowsum, d_row, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (int g = 0; g < G; ++g) {
      if (colsum) colsum[g] = hs[2 * (size_t)g];
      if (mu_guess) mu_guess[g] = hs[2 * (size_t)g + 1] * (double)G / (double)N;   // colMeans(Y / rowMeans(Y)), :222
    }
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

// rowSums(Y[, keep]) over the resident matrix: the cell filter of preprocess_for_clonealign (R/preprocess.R:138-139), which
// counts only the genes that survived the gene filters.  gene_keep: G bytes, non-zero = gene retained.
int ca_core_data_masked_rowsums(ca_data* d, const uint8_t* gene_keep, double* rowsum, char* err, size_t errlen) {
  try {
    if (!d || !gene_keep || !rowsum) fail("null argument");
    CUDA_OK(cudaSetDevice(d->dev));
    const int64_t N = d->N;
    const int G = d->G;
    double* d_row = nullptr;
    unsigned char* d_keep = nullptr;
    struct Guard {
      double*& a; unsigned char*& b;
      ~Guard() { cudaFree(a); cudaFree(b); }
    } guard{d_row, d_keep};
    CUDA_OK(cudaMalloc(&d_row, sizeof(double) * N));
    CUDA_OK(cudaMalloc(&d_keep, (size_t)G));
    cudaStream_t st = nullptr;
    CUDA_OK(cudaMemcpyAsync(d_keep, gene_keep, (size_t)G, cudaMemcpyHostToDevice, st));
    auto run = [&](auto* Yp) {
      using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
      CA_LAUNCH(k_stats_rows_masked<T>, (unsigned)ceil_div64(N, 8), 256, 0, st)(Yp, d->ldY, N, G, d_keep, d_row);
      KCHECK();
    };
    switch (d->ystore) {
      case CA_STORE_F32: run((const float*)d->Y); break;
      case CA_STORE_U16: run((const uint16_t*)d->Y); break;
      case CA_STORE_U8: run((const uint8_t*)d->Y); break;
      default: fail("bad y_store");
    }
    CUDA_OK(cudaMemcpyAsync(rowsum, d_row, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_create_shared(ca_handle** out, const ca_config* cfg, ca_data* data, const double* psi_init, const double* loc_init,
                          const double* X, char* err, size_t errlen) {
  ca_handle* h = nullptr;
  try {
    if (!out || !cfg || !data) fail("null argument");
    h = new ca_handle();
    h->cfg = *cfg;
    h->shared = data;
    data->refs++;
    build(h, nullptr, nullptr, psi_init, loc_init, X, nullptr, nullptr, nullptr, nullptr);
    h->cfg.nccl_id = nullptr;
    h->cfg.y_indptr = nullptr;
    h->cfg.y_indices = nullptr;
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    destroy(h);
    return report(e, err, errlen);
  }
}

int ca_core_destroy(ca_handle* h) {
  destroy(h);
  return 0;
}

int ca_core_init_gamma(ca_handle* h, char* err, size_t errlen) {
  try {
    if (!h) fail("null handle");
    CUDA_OK(cudaSetDevice(h->dev));
    run_forward(h, EPI_INIT);
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_step(ca_handle* h, char* err, size_t errlen) {
  try {
    if (!h) fail("null handle");
    CUDA_OK(cudaSetDevice(h->dev));
    train_step(h);
    return 0;   // asynchronous: the next call on this handle is stream-ordered behind it
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_grads(ca_handle* h, char* err, size_t errlen) {
  try {
    if (!h) fail("null handle");
    CUDA_OK(cudaSetDevice(h->dev));
    h->inspect = true;
    run_train(h, false);
    h->inspect = false;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_elbo(ca_handle* h, double* elbo, char* err, size_t errlen) {
  try {
    if (!h || !elbo) fail("null argument");
    CUDA_OK(cudaSetDevice(h->dev));
    eval_step(h);
    CUDA_OK(cudaMemcpyAsync(elbo, h->elbo_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    int p2p_failed = 0;
    if (h->p2p_err) CUDA_OK(cudaMemcpyAsync(&p2p_failed, h->p2p_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    if (p2p_failed) fail("variant p2p: the all-reduce kernel timed out waiting for a peer's contribution");
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

// n evaluations of the ELBO with fresh draws, queued back to back on the stream and fetched with ONE device-to-host
// copy (the 20 evaluations behind final_elbo / sd_final_elbo, R/inference-tflow.R:447-449, each a sess$run(elbo) with
// its own host round trip in the reference).  Same draw sequence, same values as n calls of ca_core_elbo; the
// parameters do not change in between, so only the first evaluation can need a Y pass.
int ca_core_elbo_many(ca_handle* h, int32_t n, double* elbo, char* err, size_t errlen) {
  double* d_out = nullptr;
  try {
    if (!h || !elbo || n < 0) fail("bad argument");
    if (n == 0) return 0;
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaMalloc(&d_out, sizeof(double) * (size_t)n));
    for (int i = 0; i < n; ++i) {
      eval_step(h);
      CUDA_OK(cudaMemcpyAsync(d_out + i, h->elbo_dev, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    CUDA_OK(cudaMemcpyAsync(elbo, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    int p2p_failed = 0;
    if (h->p2p_err) CUDA_OK(cudaMemcpyAsync(&p2p_failed, h->p2p_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaFree(d_out);
    d_out = nullptr;
    if (p2p_failed) fail("variant p2p: the all-reduce kernel timed out waiting for a peer's contribution");
    return 0;
  } catch (const std::exception& e) {
    if (d_out) { cudaStreamSynchronize(h->stream); cudaFree(d_out); }
    return report(e, err, errlen);
  }
}

int ca_core_params(ca_handle* h, double* mu, double* clone_probs, double* s, double* alpha, double* psi, double* W,
                   double* chi, double* beta, double* clone_probs_from_snv, char* err, size_t errlen) {
  try {
    if (!h) fail("null handle");
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    const int64_t N = h->N;
    const int G = h->G, C = h->C, K = h->K, KP = h->KP;
    if (mu) {   // tf$nn$softplus(qmu$distribution$loc), R/inference-tflow.R:424
      std::vector<double> tmp(G);
      download_colmajor(h, h->loc, G, 1, 1, 0, tmp.data());
      for (int g = 0; g < G; ++g) { double x = tmp[g]; mu[g] = x > 0 ? x + log1p(exp(-x)) : log1p(exp(x)); }
    }
    if (clone_probs) {   // softmax(gamma_logits), :273,424
      std::vector<double> tmp((size_t)N * C);
      download_colmajor(h, h->t, N, C, C, 0, tmp.data());
      for (int64_t n = 0; n < N; ++n) {
        double mx = -1e300, z = 0;
        for (int c = 0; c < C; ++c) mx = std::max(mx, tmp[(size_t)c * N + n]);
        for (int c = 0; c < C; ++c) z += exp(tmp[(size_t)c * N + n] - mx);
        for (int c = 0; c < C; ++c) clone_probs[(size_t)c * N + n] = exp(tmp[(size_t)c * N + n] - mx) / z;
      }
    }
    if (s) download_colmajor(h, h->s, N, 1, 1, 0, s);
    if (alpha) {   // exp(log_softmax(alpha_unconstr))
      std::vector<double> tmp(C);
      download_colmajor(h, h->u, C, 1, 1, 0, tmp.data());
      double mx = -1e300, z = 0;
      for (int c = 0; c < C; ++c) mx = std::max(mx, tmp[c]);
      for (int c = 0; c < C; ++c) z += exp(tmp[c] - mx);
      for (int c = 0; c < C; ++c) alpha[c] = exp(tmp[c] - mx) / z;
    }
    if (psi && K > 0) download_colmajor(h, h->U, N, K, KP, 0, psi);
    if (W && K > 0) download_colmajor(h, h->Vm, G, K, KP, 0, W);
    if (beta && h->P > 0) download_colmajor(h, h->Vm, G, h->P, KP, K, beta);
    if (chi && K > 0) {
      std::vector<double> tmp(K);
      download_colmajor(h, h->chi_raw, K, 1, 1, 0, tmp.data());
      for (int k = 0; k < K; ++k) chi[k] = exp(tmp[k]);
    }
    if (clone_probs_from_snv) {
      if (!h->snv) fail("clone_probs_from_snv requested but the allele-specific likelihood is not in use");
      download_colmajor(h, h->snv, N, C, C, 0, clone_probs_from_snv);
    }
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_set_eps(ca_handle* h, const float* eps, int64_t n_draws, char* err, size_t errlen) {
  try {
    if (!h || !eps || n_draws < 0) fail("bad argument");
    size_t per = (size_t)h->S * h->G;
    h->eps_queue.insert(h->eps_queue.end(), eps, eps + per * (size_t)n_draws);
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_get_eps(ca_handle* h, float* eps, char* err, size_t errlen) {
  try {
    if (!h || !eps) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaMemcpyAsync(eps, h->eps, sizeof(float) * h->S * h->G, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_get_array(ca_handle* h, const char* name, double* out, int64_t n, char* err, size_t errlen) {
  try {
    if (!h || !name || !out) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    ArrayRef r;
    if (!lookup(h, name, r)) fail("unknown array '%s'", name);
    if (n < r.rows * r.cols) fail("buffer too small for '%s': need %lld", name, (long long)(r.rows * r.cols));
    download_colmajor(h, r.p, r.rows, r.cols, r.ld, r.off, out);
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_set_array(ca_handle* h, const char* name, const double* in, int64_t n, char* err, size_t errlen) {
  try {
    if (!h || !name || !in) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    ArrayRef r;
    if (!lookup(h, name, r) || !r.writable) fail("array '%s' is not writable", name);
    if (n != r.rows * r.cols) fail("size mismatch for '%s': expected %lld", name, (long long)(r.rows * r.cols));
    upload_colmajor(h, in, r.rows, r.cols, const_cast<float*>(r.p), r.ld, r.off);
    h->ydirty = true;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_time_steps(ca_handle* h, int32_t n_steps, int32_t with_eval, double* ms, char* err, size_t errlen) {
  try {
    if (!h || !ms || n_steps < 0) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    cudaEvent_t a, b;
    CUDA_OK(cudaEventCreate(&a));
    CUDA_OK(cudaEventCreate(&b));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaEventRecord(a, h->stream));
    for (int i = 0; i < n_steps; ++i) {
      train_step(h);
      if (with_eval) eval_step(h);
    }
    CUDA_OK(cudaEventRecord(b, h->stream));
    CUDA_OK(cudaEventSynchronize(b));
    float f = 0.f;
    CUDA_OK(cudaEventElapsedTime(&f, a, b));
    *ms = (double)f;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_profile_step(ca_handle* h, char* names, size_t names_len, double* ms, int32_t cap, int32_t* n_k, char* err,
                         size_t errlen) {
  try {
    if (!h || !names || !ms || !n_k) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    for (auto& p : h->prof) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    h->prof.clear();
    h->prof_on = true;
    run_train(h, true);
    h->prof_on = false;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    std::string all;
    int k = 0;
    for (auto& p : h->prof) {
      if (k >= cap) break;
      float f = 0.f;
      CUDA_OK(cudaEventElapsedTime(&f, p.a, p.b));
      ms[k++] = (double)f;
      if (!all.empty()) all += ";";
      all += p.name;
    }
    *n_k = k;
    strncpy(names, all.c_str(), names_len - 1);
    names[names_len - 1] = 0;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_correlations(ca_handle* h, const int32_t* clone_idx, const double* L, double* out, char* err, size_t errlen) {
  try {
    if (!h || !clone_idx || !out) fail("bad argument");
    CUDA_OK(cudaSetDevice(h->dev));
    const int64_t N = h->N;
    const int G = h->G, C = h->C;
    if ((size_t)C * 128 * sizeof(float) > 48 * 1024) fail("too many clones for the correlation kernel");
    int* d_z = nullptr;
    float* d_L = nullptr;
    double *d_part = nullptr, *d_out = nullptr, *d_sums = nullptr;
    struct Guard {   // scratch is released on every exit path
      int*& z; float*& l; double*& p; double*& o; double*& s;
      ~Guard() { cudaFree(z); cudaFree(l); cudaFree(p); cudaFree(o); cudaFree(s); }
    } guard{d_z, d_L, d_part, d_out, d_sums};
    const int RS = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / 64));
    CUDA_OK(cudaMalloc(&d_z, sizeof(int) * N));
    CUDA_OK(cudaMalloc(&d_L, sizeof(float) * (size_t)G * C));
    CUDA_OK(cudaMalloc(&d_part, sizeof(double) * (size_t)RS * G * 5));
    CUDA_OK(cudaMalloc(&d_out, sizeof(double) * G));
    CUDA_OK(cudaMalloc(&d_sums, sizeof(double) * ((size_t)5 * G + 1 + C)));
    std::vector<double> tail(1 + (size_t)C, 0.0);      // number of assigned cells, then cells per clone (this shard)
    for (int64_t n = 0; n < N; ++n)
      if (clone_idx[n] >= 0 && clone_idx[n] < C) { tail[0] += 1.0; tail[1 + clone_idx[n]] += 1.0; }
    CUDA_OK(cudaMemcpyAsync(d_z, clone_idx, sizeof(int) * N, cudaMemcpyHostToDevice, h->stream));
    if (L) upload_colmajor(h, L, G, C, d_L, C, 0);
    else CUDA_OK(cudaMemcpyAsync(d_L, h->L, sizeof(float) * (size_t)G * C, cudaMemcpyDeviceToDevice, h->stream));
    dispatch_y(h, [&](auto* Yp) {
      using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
      dim3 grid((G + 127) / 128, RS);
      CA_LAUNCH(k_corr_part<T>, grid, 128, sizeof(float) * C * 128, h->stream)(Yp, h->ldY, N, G, C, d_z, d_L, RS, d_part);
      KCHECK();
    });
    CA_LAUNCH(k_corr_reduce, (G + 127) / 128, 128, 0, h->stream)(d_part, RS, G, d_sums);
    KCHECK();
    CUDA_OK(cudaMemcpyAsync(d_sums + (size_t)5 * G, tail.data(), sizeof(double) * tail.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));   // `tail` lives on this stack frame
    if (h->cfg.world > 1)   // collective: the sums over cells run over every shard
      NCCL_OK(nccl().AllReduce(d_sums, d_sums, (size_t)5 * G + 1 + C, kNcclFloat64, kNcclSum, h->comm, h->stream));
    CA_LAUNCH(k_corr_final, (G + 127) / 128, 128, 0, h->stream)(d_sums, d_L, G, C, d_out);
    KCHECK();
    CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(double) * G, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_pca_scores(ca_handle* h, int32_t max_iter, double tol, double* scores, int32_t* iters_out, char* err, size_t errlen) {
  try {
    if (!h || !scores || max_iter < 1) fail("bad argument");
    const bool sharded = h->cfg.world > 1;   // collective call: column statistics and X^T t are summed over the cell shards
    CUDA_OK(cudaSetDevice(h->dev));
    const int64_t N = h->N;
    const int G = h->G;
    if (h->Ntot < 2) fail("need at least two cells");
    const int RS = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / 64));
    std::vector<void*> tmp;
    auto dalloc = [&](size_t n) {
      void* p = nullptr;
      CUDA_OK(cudaMalloc(&p, sizeof(double) * (n ? n : 1)));
      tmp.push_back(p);
      return (double*)p;
    };
    double *part = dalloc((size_t)RS * G * 2), *mean = dalloc(G), *inv_sd = dalloc(G), *v = dalloc(G), *w = dalloc(G), *a = dalloc(G),
           *b = dalloc(1), *t = dalloc(N), *tsum = dalloc(RS), *out2 = dalloc(2), *sums = dalloc((size_t)2 * G + 2);
    int* bad = (int*)dalloc(1);
    CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), h->stream));
    int status = 0;
    std::string msg;
    try {
      dispatch_y(h, [&](auto* Yp) {
        using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
        dim3 gridc((G + 127) / 128, RS);
        CA_LAUNCH(k_pca_colstats<T>, gridc, 128, 0, h->stream)(Yp, h->ldY, N, G, RS, part);
        KCHECK();
        CA_LAUNCH(k_pca_colstats_reduce, (G + 127) / 128, 128, 0, h->stream)(part, RS, G, sums);
        KCHECK();
        if (sharded) NCCL_OK(nccl().AllReduce(sums, sums, (size_t)2 * G, kNcclFloat64, kNcclSum, h->comm, h->stream));
        CA_LAUNCH(k_pca_colstats_final, (G + 127) / 128, 128, 0, h->stream)(sums, G, (double)h->Ntot, mean, inv_sd, bad);
        KCHECK();
        int hbad = 0;
        CUDA_OK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(cudaStreamSynchronize(h->stream));
        if (hbad) fail("cannot rescale a constant/zero column to unit variance");   // prcomp(..., scale = TRUE)
        // deterministic start: v_g proportional to 1 + (g mod 7) / 7 (not orthogonal to a dominant direction in practice)
        std::vector<double> v0(G);
        double nn = 0.0;
        for (int g = 0; g < G; ++g) { v0[g] = 1.0 + (double)(g % 7) / 7.0; nn += v0[g] * v0[g]; }
        for (int g = 0; g < G; ++g) v0[g] /= sqrt(nn);
        CUDA_OK(cudaMemcpyAsync(v, v0.data(), sizeof(double) * G, cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(cudaStreamSynchronize(h->stream));
        int it = 0;
        for (; it < max_iter; ++it) {
          CA_LAUNCH(k_pca_prepare, 1, 1024, 0, h->stream)(v, mean, inv_sd, G, a, b);
          KCHECK();
          CA_LAUNCH(k_pca_rows<T>, (unsigned)ceil_div64(N, 8), 256, 0, h->stream)(Yp, h->ldY, N, G, a, b, t);
          KCHECK();
          CA_LAUNCH(k_pca_cols<T>, gridc, 128, 0, h->stream)(Yp, h->ldY, N, G, RS, t, part, tsum);
          KCHECK();
          CA_LAUNCH(k_pca_cols_reduce, (G + 1 + 127) / 128, 128, 0, h->stream)(part, tsum, RS, G, sums);
          KCHECK();
          if (sharded) NCCL_OK(nccl().AllReduce(sums, sums, (size_t)G + 1, kNcclFloat64, kNcclSum, h->comm, h->stream));
          CA_LAUNCH(k_pca_update, 1, 1024, 0, h->stream)(sums, G, mean, inv_sd, v, w, out2);
          KCHECK();
          double o2[2];
          CUDA_OK(cudaMemcpyAsync(o2, out2, sizeof o2, cudaMemcpyDeviceToHost, h->stream));
          CUDA_OK(cudaStreamSynchronize(h->stream));
          if (o2[1] < tol) { ++it; break; }
        }
        if (iters_out) *iters_out = it;
        // scores of the converged direction: t = X v
        CA_LAUNCH(k_pca_prepare, 1, 1024, 0, h->stream)(v, mean, inv_sd, G, a, b);
        KCHECK();
        CA_LAUNCH(k_pca_rows<T>, (unsigned)ceil_div64(N, 8), 256, 0, h->stream)(Yp, h->ldY, N, G, a, b, t);
        KCHECK();
      });
      // sign convention (prcomp's is arbitrary): the loading of largest magnitude is positive
      std::vector<double> hv(G);
      CUDA_OK(cudaMemcpyAsync(hv.data(), v, sizeof(double) * G, cudaMemcpyDeviceToHost, h->stream));
      CUDA_OK(cudaMemcpyAsync(scores, t, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream));
      CUDA_OK(cudaStreamSynchronize(h->stream));
      int gmax = 0;
      for (int g = 1; g < G; ++g)
        if (fabs(hv[g]) > fabs(hv[gmax])) gmax = g;
      if (hv[gmax] < 0.0)
        for (int64_t n = 0; n < N; ++n) scores[n] = -scores[n];
    } catch (const std::exception& e) {
      status = 1;
      msg = e.what();
    }
    for (void* p : tmp) cudaFree(p);
    if (status) fail("%s", msg.c_str());
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_p2p_export(ca_handle* h, void* handle64, char* err, size_t errlen) {
  try {
    if (!h || !handle64) fail("bad argument");
    if (!h->p2p) fail("ca_core_p2p_export: the session was not created with variant p2p (and world > 1)");
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaStreamSynchronize(h->stream));   // the zero-fill of the flags must have landed before a peer can signal
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t hd;
    CUDA_OK(cudaIpcGetMemHandle(&hd, h->p2p_buf));
    memcpy(handle64, &hd, 64);
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_p2p_connect(ca_handle* h, const void* handles, char* err, size_t errlen) {
  try {
    if (!h || !handles) fail("bad argument");
    if (!h->p2p) fail("ca_core_p2p_connect: the session was not created with variant p2p (and world > 1)");
    CUDA_OK(cudaSetDevice(h->dev));
    const int world = h->cfg.world;
    const size_t slot_bytes = sizeof(float) * 2 * (size_t)world * h->p2p_cnt_pad;
    for (int r = 0; r < world; ++r) {
      void* base = nullptr;
      if (r == h->cfg.rank) {
        base = h->p2p_buf;
      } else {
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)handles + 64 * (size_t)r, 64);
        CUDA_OK(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
        h->p2p_mapped[r] = base;
      }
      h->p2p_slots[r] = (float*)base;
      h->p2p_flags[r] = (unsigned*)((char*)base + slot_bytes + 256);
    }
    h->p2p_ready = true;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

// variant p2p inside ONE process (ca_core_multi_*): the peers' exchange buffers are ordinary device pointers (CUDA IPC
// handles cannot be opened by the process that exported them); peer access is enabled on demand.
int ca_core_p2p_base(ca_handle* h, void** base, char* err, size_t errlen) {
  try {
    if (!h || !base) fail("bad argument");
    if (!h->p2p) fail("ca_core_p2p_base: the session was not created with variant p2p (and world > 1)");
    CUDA_OK(cudaSetDevice(h->dev));
    CUDA_OK(cudaStreamSynchronize(h->stream));   // the zero-fill of the flags must have landed before a peer can signal
    *base = h->p2p_buf;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_p2p_connect_ptrs(ca_handle* h, void* const* bases, const int32_t* devices, char* err, size_t errlen) {
  try {
    if (!h || !bases || !devices) fail("bad argument");
    if (!h->p2p) fail("ca_core_p2p_connect_ptrs: the session was not created with variant p2p (and world > 1)");
    CUDA_OK(cudaSetDevice(h->dev));
    const int world = h->cfg.world;
    const size_t slot_bytes = sizeof(float) * 2 * (size_t)world * h->p2p_cnt_pad;
    for (int r = 0; r < world; ++r) {
      if (r != h->cfg.rank && devices[r] != h->dev) {
        int can = 0;
        CUDA_OK(cudaDeviceCanAccessPeer(&can, h->dev, devices[r]));
        if (!can) fail("variant p2p: device %d cannot access device %d", h->dev, devices[r]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_OK(e);
        cudaGetLastError();
      }
      h->p2p_slots[r] = (float*)bases[r];
      h->p2p_flags[r] = (unsigned*)((char*)bases[r] + slot_bytes + 256);
    }
    h->p2p_ready = true;
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_ypass_many(ca_handle* const* hs, int32_t n, char* err, size_t errlen) {
  try {
    if (!hs || n < 1) fail("bad argument");
    ca_handle* h0 = hs[0];
    if (!h0) fail("null handle");
    for (int i = 0; i < n; ++i) {
      ca_handle* h = hs[i];
      if (!h) fail("null handle");
      if (h->KP != 1) fail("ca_core_ypass_many needs K + P == 1");
      if ((h->variants & (CA_VAR_YPASS3 | CA_VAR_YPASS4)) && n > 1) fail("ca_core_ypass_many: the batched kernel uses the column tiling of ypass2 (sessions with variant ypass3 / ypass4 run their own pass)");
      if (h->Y != h0->Y || h->dev != h0->dev || h->N != h0->N || h->G != h0->G || h->ystore != h0->ystore)
        fail("ca_core_ypass_many: the sessions do not share one count matrix (create them with ca_core_create_shared)");
    }
    CUDA_OK(cudaSetDevice(h0->dev));
    for (int g0 = 0; g0 < n; g0 += kYMultiMax) {
      const int R = std::min(kYMultiMax, n - g0);
      ca_handle* lead = hs[g0];
      if (R == 1) {            // a lone fit: its own pass
        lead->ydirty = true;
        run_ypass(lead, lead->stream);
        continue;
      }
      // the pass runs on the first session's stream: it must see the parameter updates of the others, and their next
      // kernels must see its partial sums
      for (int r = 1; r < R; ++r) {
        CUDA_OK(cudaEventRecord(hs[g0 + r]->ev_fork, hs[g0 + r]->stream));
        CUDA_OK(cudaStreamWaitEvent(lead->stream, hs[g0 + r]->ev_fork, 0));
      }
      YMultiArgs a;
      for (int r = 0; r < kYMultiMax; ++r) {
        ca_handle* h = hs[g0 + (r < R ? r : 0)];
        a.U[r] = h->U; a.Vm[r] = h->Vm; a.rowpart[r] = h->rowpart; a.colpart[r] = h->colpart;
      }
      dispatch_y(lead, [&](auto* Yp) {
        using T = typename std::remove_const<typename std::remove_pointer<decltype(Yp)>::type>::type;
        dim3 grid(lead->nCB, lead->nRB);
        if (R == 2) { auto k = k_ypass_k1_multi<T, 2>; CA_LAUNCH(k, grid, 256, 0, lead->stream)(Yp, lead->ldY, lead->N, lead->G, lead->RB, a); }
        else if (R == 3) { auto k = k_ypass_k1_multi<T, 3>; CA_LAUNCH(k, grid, 256, 0, lead->stream)(Yp, lead->ldY, lead->N, lead->G, lead->RB, a); }
        else { auto k = k_ypass_k1_multi<T, 4>; CA_LAUNCH(k, grid, 256, 0, lead->stream)(Yp, lead->ldY, lead->N, lead->G, lead->RB, a); }
        KCHECK();
      });
      CUDA_OK(cudaEventRecord(lead->ev_join, lead->stream));
      for (int r = 0; r < R; ++r) {
        if (r > 0) CUDA_OK(cudaStreamWaitEvent(hs[g0 + r]->stream, lead->ev_join, 0));
        hs[g0 + r]->ydirty = false;
      }
    }
    return 0;
  } catch (const std::exception& e) { return report(e, err, errlen); }
}

int ca_core_describe(ca_handle* h, char* json, size_t json_len) {
  if (!h || !json || !json_len) return 1;
  const char* st = h->ystore == CA_STORE_F32 ? "f32" : (h->ystore == CA_STORE_U16 ? "u16" : "u8");
  int bpe = h->ystore == CA_STORE_F32 ? 4 : (h->ystore == CA_STORE_U16 ? 2 : 1);
  // interp path: the panel structure of the last step (device-side data: the node work is proportional to it)
  InterpPlan pl;
  memset(&pl, 0, sizeof pl);
  if (h->iplan) {
    cudaSetDevice(h->dev);
    cudaMemcpyAsync(&pl, h->iplan, sizeof pl, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
  }
  snprintf(json, json_len,
           "{\"N\": %lld, \"G\": %d, \"C\": %d, \"S\": %d, \"K\": %d, \"P\": %d, \"path\": \"%s\", \"y_store\": \"%s\", "
           "\"y_bytes_per_entry\": %d, \"ldY\": %lld, \"launches_last_step\": %d, \"nsplit\": %d, \"fsplit\": %d, "
           "\"SCp\": %d, \"J\": %d, \"world\": %d, \"rank\": %d, \"variants\": %u, \"ypass_grid\": [%d, %d], \"ypass_rows_per_block\": %d, "
           "\"num_sms\": %d, \"panels\": {\"nf_neg\": %d, \"nf_pos\": %d, \"nb\": %d, \"w_range\": [%.6g, %.6g], \"psi_range\": [%.6g, %.6g]}}",
           (long long)h->N, h->G, h->C, h->S, h->K, h->P, h->interp ? "interp" : (h->tc ? "tcgen05" : "cudacore"), st, bpe, (long long)h->ldY,
           h->launches_last_step, h->nsplit, h->tc ? h->tcplan.fsplit : 1, h->SCp, h->J, h->cfg.world, h->cfg.rank, h->variants, h->nCB, h->nRB, h->RB,
           h->num_sms, pl.nf_neg, pl.nf_pos, pl.nb, pl.wmin, pl.wmax, pl.pmin, pl.pmax);
  return 0;
}

}  // extern "C"
