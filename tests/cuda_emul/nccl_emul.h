// In-process stand-in for the four NCCL calls core.cu uses, for the CPU emulation (tests only): the ranks of a
// cell-sharded fit are threads of ONE process (tests/test_emul_parity.py drives them from Python threads; ctypes releases
// the GIL), a communicator is a rendezvous object looked up by the 128-byte unique id, and an all-reduce is
// "everybody deposits its pointer, barrier, everybody sums all contributions in RANK ORDER, barrier" -- i.e. every rank
// computes bit-identical sums, which is also what the real one-shot all-reduce does.
#pragma once
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

namespace ca_emul_nccl {

struct Group {
  int world = 0, joined = 0, left = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned gen = 0;
  std::vector<const void*> src;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned g = gen;
    if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
struct Comm { Group* grp; int rank; };

inline std::mutex& reg_mu() { static std::mutex m; return m; }
inline std::map<std::string, Group*>& registry() { static std::map<std::string, Group*> r; return r; }

inline int GetUniqueId(void* out128) {
  static std::mt19937_64 gen(12345);
  std::lock_guard<std::mutex> lk(reg_mu());
  uint64_t* p = (uint64_t*)out128;
  for (int i = 0; i < 16; ++i) p[i] = gen();
  return 0;
}
template <typename Uid>
inline int CommInitRank(void** comm, int world, Uid id, int rank) {
  Group* g;
  {
    std::lock_guard<std::mutex> lk(reg_mu());
    std::string key((const char*)&id, sizeof id);
    auto it = registry().find(key);
    if (it == registry().end()) {
      g = new Group();
      g->world = world;
      g->src.resize(world);
      registry()[key] = g;
    } else {
      g = it->second;
    }
    if (g->world != world) return 5;
    g->joined++;
  }
  *comm = new Comm{g, rank};
  g->barrier();   // like ncclCommInitRank: returns once every rank has joined
  return 0;
}
inline int AllReduce(const void* send, void* recv, size_t count, int dtype, int /*op: sum*/, void* comm, void* /*stream*/) {
  Comm* c = (Comm*)comm;
  Group* g = c->grp;
  const size_t esz = dtype == 8 ? 8 : 4;
  std::vector<unsigned char> mine(count * esz);
  std::memcpy(mine.data(), send, count * esz);     // in-place calls: keep this rank's contribution intact
  g->src[c->rank] = mine.data();
  g->barrier();
  if (dtype == 8) {
    double* out = (double*)recv;
    for (size_t i = 0; i < count; ++i) {
      double s = 0.0;
      for (int r = 0; r < g->world; ++r) s += ((const double*)g->src[r])[i];
      out[i] = s;
    }
  } else {
    float* out = (float*)recv;
    for (size_t i = 0; i < count; ++i) {
      float s = 0.f;
      for (int r = 0; r < g->world; ++r) s += ((const float*)g->src[r])[i];
      out[i] = s;
    }
  }
  g->barrier();
  return 0;
}
inline int CommDestroy(void* comm) {
  delete (Comm*)comm;
  return 0;
}
inline const char* GetErrorString(int) { return "emulated NCCL error"; }

}  // namespace ca_emul_nccl
