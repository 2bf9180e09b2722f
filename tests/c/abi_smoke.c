/* Plain-C99 client of the C ABI (what cgo compiles against): the header must be valid C, the library must link
 * without a C++ runtime on the caller's side, and without a CUDA device sphb_create must fail loudly with
 * SPHB_E_CUDA - never compute on the CPU.  Exit code 0 = behaved as specified. */
#include <stdio.h>
#include <string.h>
#include "sphb.h"

int main(void) {
  sphb_params p;
  memset(&p, 0, sizeof p);
  p.dt_half = 0.001; p.gamma = 1.66666; p.particle_mass = 1.0;
  p.hor[0] = SPHB_OPEN_LO; p.hor[1] = SPHB_OPEN_HI; p.ver[0] = SPHB_OPEN_LO; p.ver[1] = SPHB_OPEN_HI;
  p.refl_L = SPHB_OPEN_LO; p.refl_R = SPHB_OPEN_HI; p.refl_U = SPHB_OPEN_LO; p.refl_D = SPHB_OPEN_HI;
  p.kernel = SPHB_KERNEL_MONAGHAN; p.precision = 64; p.device = 0; p.flags = 0;
  double pos[64 * 2];
  for (int i = 0; i < 64; ++i) { pos[2 * i] = (i % 8 + 0.5) / 8.0; pos[2 * i + 1] = (i / 8 + 0.5) / 8.0; }
  sphb_sim* s = NULL;
  int rc = sphb_create(&p, 64, 64, pos, NULL, NULL, NULL, NULL, &s);
  if (rc == SPHB_OK) {  /* a GPU is present: one step must work and the handle must report 64 particles */
    rc = sphb_step(s, 1);
    if (rc == SPHB_OK) rc = sphb_sync(s);
    long long n = (long long)sphb_count(s);
    printf("gpu path: step rc=%d count=%lld\n", rc, n);
    sphb_destroy(s);
    return (rc == SPHB_OK && n == 64) ? 0 : 2;
  }
  printf("no device: rc=%d (%s)\n", rc, sphb_last_error(NULL));
  /* bad arguments are refused before any device work */
  p.kernel = 7;
  int rc2 = sphb_create(&p, 64, 64, pos, NULL, NULL, NULL, NULL, &s);
  return (rc == SPHB_E_CUDA && rc2 == SPHB_E_INVALID && s == NULL) ? 0 : 1;
}
