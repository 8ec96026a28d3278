/* Plain-C client of the C-ABI (what an R .Call shim or a cgo/JNI stub does): builds an R-layout problem
 * (column-major doubles), creates a session, runs init_gamma / step / elbo / params, destroys it.
 * Without a usable CUDA device every call must fail with a message (there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "clonealign_b200.h"

int main(void) {
  enum { N = 40, G = 24, C = 3 };
  static double Y[N * G], L[G * C], psi[N], loc[G];
  char err[512] = {0};
  unsigned s = 12345u;
  for (int g = 0; g < G; ++g)
    for (int n = 0; n < N; ++n) {          /* column-major: cell index fastest */
      s = s * 1664525u + 1013904223u;
      Y[g * N + n] = (double)((s >> 24) % 7);
    }
  for (int n = 0; n < N; ++n) Y[0 * N + n] += 1.0;   /* no empty cell */
  for (int c = 0; c < C; ++c)
    for (int g = 0; g < G; ++g) L[c * G + g] = 1.0 + (double)((g + c) % 3);
  for (int n = 0; n < N; ++n) psi[n] = sin(0.7 * n);
  for (int g = 0; g < G; ++g) loc[g] = 0.5;

  int ndev = -1;
  int st = ca_core_device_count(&ndev, err, sizeof err);
  printf("abi=%d device_count_status=%d ndev=%d\n", ca_core_abi_version(), st, ndev);

  ca_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.N = N; cfg.N_total = N; cfg.G = G; cfg.C = C; cfg.S = 2; cfg.K = 1; cfg.P = 0; cfg.V = 0;
  cfg.learning_rate = 0.1; cfg.seed = 7; cfg.device = 0; cfg.rank = 0; cfg.world = 1;
  cfg.y_dtype = CA_Y_F64; cfg.y_layout = CA_Y_COLMAJOR; cfg.y_mem = CA_Y_HOST; cfg.y_store = CA_STORE_AUTO; cfg.path = CA_PATH_AUTO;
  ca_handle* h = NULL;
  st = ca_core_create(&h, &cfg, Y, L, psi, loc, NULL, NULL, NULL, NULL, NULL, err, sizeof err);
  if (st != 0) {
    printf("create_failed status=%d msg_len=%zu msg=%s\n", st, strlen(err), err);
    return (h == NULL && strlen(err) > 0) ? 0 : 2;     /* loud failure, nothing leaked */
  }
  double e0 = 0, e1 = 0, mu[G], cp[N * C];
  if (ca_core_init_gamma(h, err, sizeof err) || ca_core_elbo(h, &e0, err, sizeof err) || ca_core_step(h, err, sizeof err) ||
      ca_core_elbo(h, &e1, err, sizeof err) ||
      ca_core_params(h, mu, cp, NULL, NULL, NULL, NULL, NULL, NULL, NULL, err, sizeof err)) {
    printf("call_failed msg=%s\n", err);
    ca_core_destroy(h);
    return 3;
  }
  double rs = 0;
  for (int c = 0; c < C; ++c) rs += cp[c * N + 0];
  printf("ok elbo0=%.6f elbo1=%.6f rowsum0=%.6f\n", e0, e1, rs);
  ca_core_destroy(h);
  return (isfinite(e0) && isfinite(e1) && fabs(rs - 1.0) < 1e-5) ? 0 : 4;
}
