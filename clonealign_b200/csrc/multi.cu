// One fit sharded by cells over several GPUs of ONE process, driven from one host thread (include/clonealign_b200.h,
// ca_core_multi_*).  The reference runs a single TensorFlow session from the R interpreter (R/inference-tflow.R:351);
// R is single-threaded and its API must not be touched from other threads, so the library owns one worker thread per
// device: every call of the caller is handed to all workers, which issue the same per-shard C-ABI call (ca_core_*) on
// their own device -- exactly what the ranks of the one-process-per-GPU launch do, so results are identical -- and the
// caller's thread waits for all of them.  NCCL calls of different communicators are therefore never issued from one
// thread one after another (which would dead-lock without group semantics).  Written against the public C-ABI only.
#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/clonealign_b200.h"

namespace {

struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, done = false, quit = false;
  void loop() {
    for (;;) {
      std::function<void()> j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return has_job || quit; });
        if (quit && !has_job) return;
        j = std::move(job);
        has_job = false;
      }
      j();
      {
        std::lock_guard<std::mutex> lk(mu);
        done = true;
      }
      cv.notify_all();
    }
  }
};

}  // namespace

struct ca_multi {
  int n = 0;
  int64_t N = 0;
  int G = 0, C = 0, K = 0, P = 0, V = 0;
  std::vector<ca_handle*> h;
  std::vector<int64_t> a, b;
  std::vector<Worker*> w;

  // run fn(i) on worker i for every shard; returns 0 or the first failure (message in err)
  int run_all(const std::function<int(int, char*, size_t)>& fn, char* err, size_t errlen) {
    std::vector<int> st(n, 0);
    std::vector<std::string> msg(n);
    for (int i = 0; i < n; ++i) {
      Worker* wk = w[i];
      std::lock_guard<std::mutex> lk(wk->mu);
      wk->done = false;
      wk->job = [&, i] {
        char buf[1024];
        buf[0] = 0;
        st[i] = fn(i, buf, sizeof buf);
        if (st[i]) msg[i] = buf;
      };
      wk->has_job = true;
      wk->cv.notify_all();
    }
    for (int i = 0; i < n; ++i) {
      std::unique_lock<std::mutex> lk(w[i]->mu);
      w[i]->cv.wait(lk, [&] { return w[i]->done; });
    }
    for (int i = 0; i < n; ++i)
      if (st[i]) {
        if (err && errlen) snprintf(err, errlen, "shard %d (of %d): %s", i, n, msg[i].c_str());
        return st[i];
      }
    return 0;
  }
};

namespace {

int fail(char* err, size_t errlen, const char* m) {
  if (err && errlen) snprintf(err, errlen, "%s", m);
  return 1;
}
size_t y_elem(int32_t dt) {
  switch (dt) {
    case CA_Y_F64: return 8;
    case CA_Y_F32: case CA_Y_I32: return 4;
    case CA_Y_U16: return 2;
    default: return 1;
  }
}
// rows [a, b) of a column-major rows x cols double matrix as a contiguous column-major block
std::vector<double> row_block(const double* src, int64_t rows, int cols, int64_t a, int64_t b) {
  std::vector<double> out((size_t)(b - a) * (cols > 0 ? cols : 0));
  if (!src) return out;
  for (int c = 0; c < cols; ++c) memcpy(out.data() + (size_t)c * (b - a), src + (size_t)c * rows + a, sizeof(double) * (size_t)(b - a));
  return out;
}
void destroy_multi(ca_multi* m) {
  if (!m) return;
  if (!m->w.empty())
    m->run_all([&](int i, char*, size_t) { if (m->h[i]) ca_core_destroy(m->h[i]); m->h[i] = nullptr; return 0; }, nullptr, 0);
  for (Worker* wk : m->w) {
    {
      std::lock_guard<std::mutex> lk(wk->mu);
      wk->quit = true;
    }
    wk->cv.notify_all();
    if (wk->th.joinable()) wk->th.join();
    delete wk;
  }
  delete m;
}

}  // namespace

extern "C" {

int ca_core_multi_create(ca_multi** out, const ca_config* cfg, const int32_t* devices, int32_t n_devices, const void* Y,
                         const double* L, const double* psi_init, const double* loc_init, const double* X,
                         const double* clone_allele, const double* alt, const double* cov, char* err, size_t errlen) {
  if (!out || !cfg || !devices || n_devices < 1 || !Y || !L) return fail(err, errlen, "ca_core_multi_create: bad argument");
  if (cfg->y_mem != CA_Y_HOST) return fail(err, errlen, "ca_core_multi_create: Y must be in host memory");
  if (cfg->N < n_devices) return fail(err, errlen, "ca_core_multi_create: fewer cells than devices");
  ca_multi* m = new ca_multi();
  m->n = n_devices;
  m->N = cfg->N; m->G = cfg->G; m->C = cfg->C; m->K = cfg->K; m->P = cfg->P; m->V = cfg->V;
  m->h.assign(n_devices, nullptr);
  const int64_t base = cfg->N / n_devices, extra = cfg->N % n_devices;
  for (int i = 0; i < n_devices; ++i) {
    const int64_t a = i * base + (i < extra ? i : extra);
    m->a.push_back(a);
    m->b.push_back(a + base + (i < extra ? 1 : 0));
  }
  for (int i = 0; i < n_devices; ++i) {
    Worker* wk = new Worker();
    wk->th = std::thread([wk] { wk->loop(); });
    m->w.push_back(wk);
  }
  unsigned char id[128];
  memset(id, 0, sizeof id);
  if (n_devices > 1 && ca_core_nccl_unique_id(id, err, errlen)) { destroy_multi(m); return 1; }
  const int st = m->run_all([&](int i, char* e, size_t el) {
    const int64_t a = m->a[i], b = m->b[i];
    ca_config c = *cfg;
    c.N = b - a;
    c.N_total = cfg->N;
    c.rank = i;
    c.world = n_devices;
    c.device = devices[i];
    c.nccl_id = n_devices > 1 ? id : nullptr;
    const unsigned char* yp = (const unsigned char*)Y;
    const size_t es = y_elem(cfg->y_dtype);
    if (cfg->y_layout == CA_Y_COLMAJOR) {
      c.y_ld = cfg->y_ld ? cfg->y_ld : cfg->N;
      yp += (size_t)a * es;
    } else if (cfg->y_layout == CA_Y_ROWMAJOR) {
      c.y_ld = cfg->y_ld ? cfg->y_ld : cfg->G;
      yp += (size_t)a * (size_t)c.y_ld * es;
    } else {                      // CSR: the row offsets of the shard; values / indices stay absolute
      c.y_indptr = cfg->y_indptr + a;
    }
    std::vector<double> psi = row_block(psi_init, cfg->N, cfg->K, a, b), x = row_block(X, cfg->N, cfg->P, a, b);
    std::vector<double> al = row_block(alt, cfg->N, cfg->V, a, b), cv = row_block(cov, cfg->N, cfg->V, a, b);
    return ca_core_create(&m->h[i], &c, yp, L, cfg->K > 0 ? psi.data() : nullptr, loc_init, cfg->P > 0 ? x.data() : nullptr, nullptr,
                          clone_allele, cfg->V > 0 ? al.data() : nullptr, cfg->V > 0 ? cv.data() : nullptr, e, el);
  }, err, errlen);
  if (st) { destroy_multi(m); return st; }
  if (n_devices > 1 && (cfg->variants & CA_VAR_P2P)) {
    // variant p2p: every shard learns the exchange buffers of its peers (plain device pointers within one process)
    std::vector<void*> bases(n_devices, nullptr);
    int s2 = m->run_all([&](int i, char* e, size_t el) { return ca_core_p2p_base(m->h[i], &bases[i], e, el); }, err, errlen);
    if (!s2) s2 = m->run_all([&](int i, char* e, size_t el) { return ca_core_p2p_connect_ptrs(m->h[i], bases.data(), devices, e, el); }, err, errlen);
    if (s2) { destroy_multi(m); return s2; }
  }
  *out = m;
  return 0;
}

int ca_core_multi_destroy(ca_multi* m) {
  destroy_multi(m);
  return 0;
}

int ca_core_multi_size(ca_multi* m) { return m ? m->n : 0; }

int ca_core_multi_shard(ca_multi* m, int32_t i, ca_handle** out, int64_t* row_begin, int64_t* row_end) {
  if (!m || i < 0 || i >= m->n || !out) return 1;
  *out = m->h[i];
  if (row_begin) *row_begin = m->a[i];
  if (row_end) *row_end = m->b[i];
  return 0;
}

int ca_core_multi_init_gamma(ca_multi* m, char* err, size_t errlen) {
  if (!m) return fail(err, errlen, "null handle");
  return m->run_all([&](int i, char* e, size_t el) { return ca_core_init_gamma(m->h[i], e, el); }, err, errlen);
}

int ca_core_multi_step(ca_multi* m, char* err, size_t errlen) {
  if (!m) return fail(err, errlen, "null handle");
  return m->run_all([&](int i, char* e, size_t el) { return ca_core_step(m->h[i], e, el); }, err, errlen);
}

int ca_core_multi_elbo(ca_multi* m, double* elbo, char* err, size_t errlen) {
  if (!m || !elbo) return fail(err, errlen, "bad argument");
  std::vector<double> v(m->n, 0.0);
  const int st = m->run_all([&](int i, char* e, size_t el) { return ca_core_elbo(m->h[i], &v[i], e, el); }, err, errlen);
  *elbo = v[0];                  // the all-reduced value: identical on every shard
  return st;
}

int ca_core_multi_elbo_many(ca_multi* m, int32_t n, double* elbo, char* err, size_t errlen) {
  if (!m || !elbo || n < 0) return fail(err, errlen, "bad argument");
  std::vector<std::vector<double>> v(m->n, std::vector<double>((size_t)n, 0.0));
  const int st = m->run_all([&](int i, char* e, size_t el) { return ca_core_elbo_many(m->h[i], n, v[i].data(), e, el); }, err, errlen);
  if (n > 0) memcpy(elbo, v[0].data(), sizeof(double) * (size_t)n);
  return st;
}

int ca_core_multi_params(ca_multi* m, double* mu, double* clone_probs, double* s, double* alpha, double* psi, double* W,
                         double* chi, double* beta, double* clone_probs_from_snv, char* err, size_t errlen) {
  if (!m) return fail(err, errlen, "null handle");
  const int64_t N = m->N;
  return m->run_all([&](int i, char* e, size_t el) {
    const int64_t a = m->a[i], nl = m->b[i] - a;
    std::vector<double> cp(clone_probs ? (size_t)nl * m->C : 0), sv(s ? (size_t)nl : 0), ps(psi && m->K > 0 ? (size_t)nl * m->K : 0),
        snv(clone_probs_from_snv ? (size_t)nl * m->C : 0);
    const bool first = i == 0;   // gene-level / scalar outputs are replicated: shard 0 delivers them
    const int st = ca_core_params(m->h[i], first ? mu : nullptr, clone_probs ? cp.data() : nullptr, s ? sv.data() : nullptr,
                                  first ? alpha : nullptr, ps.empty() ? nullptr : ps.data(), first ? W : nullptr, first ? chi : nullptr,
                                  first ? beta : nullptr, clone_probs_from_snv ? snv.data() : nullptr, e, el);
    if (st) return st;
    auto scatter = [&](const std::vector<double>& src, double* dst, int cols) {   // shard rows -> rows a.. of the N x cols result
      if (!dst || src.empty()) return;
      for (int c = 0; c < cols; ++c) memcpy(dst + (size_t)c * N + a, src.data() + (size_t)c * nl, sizeof(double) * (size_t)nl);
    };
    scatter(cp, clone_probs, m->C);
    scatter(sv, s, 1);
    scatter(ps, psi, m->K);
    scatter(snv, clone_probs_from_snv, m->C);
    return 0;
  }, err, errlen);
}

int ca_core_multi_time_steps(ca_multi* m, int32_t n_steps, int32_t with_eval, double* ms, char* err, size_t errlen) {
  if (!m || !ms) return fail(err, errlen, "bad argument");
  std::vector<double> v(m->n, 0.0);
  const int st = m->run_all([&](int i, char* e, size_t el) { return ca_core_time_steps(m->h[i], n_steps, with_eval, &v[i], e, el); }, err, errlen);
  double mx = 0.0;
  for (double x : v) mx = x > mx ? x : mx;
  *ms = mx;                      // device time of the slowest shard
  return st;
}

}  // extern "C"
