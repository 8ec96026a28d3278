// Variant P2P (world > 1): the per-step all-reduce of the gene-level gradient partials (G (2 + K + P) + C floats,
// 240 KB at 100k x 20k x 12) as ONE kernel over NVLink peer memory instead of an ncclAllReduce launch.
//
// Every rank owns a device buffer that its peers map through CUDA IPC (ca_core_p2p_export / ca_core_p2p_connect):
//     slots[2][world][cnt_pad] floats   slot (parity, r) = rank r's contribution of a step with that parity
//     flags[2][world]          uint32   flag (parity, r) = last step whose contribution from rank r is complete
// One launch per step on every rank (<= 64 co-resident blocks: nothing else of the step runs next to it):
//   push   : the rank's partial sums are written straight into slot (parity, rank) of EVERY rank's buffer
//            (16-byte stores over NVLink; NVSwitch gives every pair full bandwidth, so the 8 x 240 KB leave in ~2 us);
//   signal : when the last block has finished pushing (ticket), one system-scope fence and `flag[parity][rank] = step`
//            on every peer;
//   reduce : every block waits until all `world` flags of its OWN buffer carry this step, then sums its slice of the
//            slots in RANK ORDER into the local all-reduce buffer.
// Rank order makes the result bit-identical on every rank and independent of timing (the replicated gene-level Adam
// state cannot drift apart).  Two parities: a fast rank can start pushing step s + 1 while a slow one still reduces
// step s; it cannot reach s + 2 before the slow rank has signalled s + 1, i.e. finished reducing s.
// Replaces: nothing in the reference (single process, no collective; SURVEY.md 8e adds exactly this exchange).
#pragma once
#include "common.cuh"

#ifndef CA_SPIN_PAUSE
#define CA_SPIN_PAUSE() __nanosleep(64)
#endif

namespace ca {

constexpr int kP2PMaxWorld = 16;
constexpr int kP2PThreads = 512;

struct P2PArgs {
  int world, rank;
  int64_t cnt, cnt_pad;            // floats per contribution (cnt_pad: multiple of 4)
  unsigned* step_ctr;              // device counter of finished exchanges: this launch is exchange *step_ctr + 1 (flags start at 0)
  const float* src;                // this rank's partial sums [cnt]
  float* dst;                      // reduced result [cnt] (may alias src)
  float* slots[kP2PMaxWorld];      // base of every rank's slots[2][world][cnt_pad] (own entry = local pointer)
  unsigned* flags[kP2PMaxWorld];   // base of every rank's flags[2][world]
  unsigned* ticket;                // local, zero between launches
  int* error;                      // local: set to 1 when a peer's flag did not arrive within kP2PSpinCycles
};
constexpr long long kP2PSpinCycles = 6000000000LL;   // ~3 s of SM clock: a dead peer must not hang the GPU for ever

__global__ void __launch_bounds__(kP2PThreads) k_p2p_allreduce(P2PArgs a) {
  __shared__ int is_last;
  const unsigned step = *a.step_ctr + 1u;      // every block reads it before it arrives at the ticket
  const int par = (int)(step & 1u);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  // ---- push ----
  const int64_t nvec = a.cnt_pad / 4;
  for (int p = 0; p < a.world; ++p) {
    const int peer = (a.rank + p) % a.world;     // stagger the targets so that the ranks do not all hit rank 0 first
    float4* out = reinterpret_cast<float4*>(a.slots[peer] + ((int64_t)par * a.world + a.rank) * a.cnt_pad);
    for (int64_t i = tid; i < nvec; i += nth) {
      float4 v;
      const int64_t e = 4 * i;
      v.x = e < a.cnt ? a.src[e] : 0.f;
      v.y = e + 1 < a.cnt ? a.src[e + 1] : 0.f;
      v.z = e + 2 < a.cnt ? a.src[e + 2] : 0.f;
      v.w = e + 3 < a.cnt ? a.src[e + 3] : 0.f;
      out[i] = v;
    }
  }
  // ---- signal (last block of this rank) ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (is_last && threadIdx.x < a.world) {
    __threadfence_system();
    volatile unsigned* f = a.flags[threadIdx.x] + par * a.world + a.rank;
    *f = step;
    if (threadIdx.x == 0) { *a.ticket = 0u; *a.step_ctr = step; }
  }
  // ---- wait for every rank's contribution in the local buffer ----
  if (threadIdx.x < a.world) {
    volatile unsigned* f = a.flags[a.rank] + par * a.world + threadIdx.x;
    const long long t0 = clock64();
    while (*f != step) {
      if (clock64() - t0 > kP2PSpinCycles) {
        *a.error = 1;
        break;
      }
      CA_SPIN_PAUSE();
    }
  }
  __syncthreads();
  __threadfence_system();
  // ---- reduce in rank order ----
  // 16 bytes per thread and rank, ALL ranks' loads in flight before the first add (the first version summed rank after rank
  // through one dependent load chain per output: ~8 load latencies per element and 7 elements per thread at 8 ranks)
  const float* base = a.slots[a.rank] + (int64_t)par * a.world * a.cnt_pad;
  for (int64_t i = tid; i < nvec; i += nth) {
    float4 v[kP2PMaxWorld];
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r)
      if (r < a.world) v[r] = __ldcv(reinterpret_cast<const float4*>(base + (int64_t)r * a.cnt_pad) + i);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r)
      if (r < a.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    const int64_t e = 4 * i;
    if (e + 3 < a.cnt) {
      reinterpret_cast<float4*>(a.dst)[i] = s;        // dst (the all-reduce buffer) is 16-byte aligned
    } else {
      if (e < a.cnt) a.dst[e] = s.x;
      if (e + 1 < a.cnt) a.dst[e + 1] = s.y;
      if (e + 2 < a.cnt) a.dst[e + 2] = s.z;
    }
  }
}

}  // namespace ca
