// core.cu, part 1: errors, NCCL binding + communicator pool, ingest kernels.
// Part of the single translation unit core.cu (included from there; not compiled on its own).
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

// path = auto picks the tensor-copy integer Y pass from this many stored elements per rank on (see profiles/r02_notes.md section 3c)
constexpr int64_t kY7MinElements = (int64_t)1 << 25;
inline bool y7_wanted(int64_t elements) {
  if (const char* e = getenv("CLONEALIGN_B200_Y7")) return atoi(e) != 0;
  return elements >= kY7MinElements;
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

#include CA_NCCL_PROVIDER
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
