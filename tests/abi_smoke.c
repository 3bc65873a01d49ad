/* C-ABI smoke test: the boundary exercised without Python.
 *
 *   gcc -O1 -o abi_smoke tests/abi_smoke.c -ldl && ./abi_smoke <path to libtsd_b200.so> [expect-gpu]
 *
 * dlopen()s the library, resolves the entry points a Mojo / C caller of the reference's hot path binds
 * (include/tsd_b200.h), and runs:  tsd_init -> tsd_diffusion_create (8x8 latent) -> tsd_diffusion_init_random ->
 * tsd_diffusion_forward (Diffusion.forward, diffusion.mojo:309-318) twice (eager pass, CUDA-graph replay: must agree
 * bit for bit) -> tsd_diffusion_step -> tsd_conv2d against a direct loop on the host.
 * Without a GPU (`expect-gpu` absent) tsd_init must fail with TSD_ERR_NO_DEVICE and a message: there is no CPU path.
 * Exit code 0 = pass. */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tsd_b200.h"

#define SYM(name)                                                    \
  __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); \
  if (!p_##name) {                                                   \
    fprintf(stderr, "missing symbol %s\n", #name);                   \
    return 2;                                                        \
  }

static float frand(uint32_t* s) {
  *s = *s * 1664525u + 1013904223u;
  return ((float)(*s >> 8) / 8388608.0f) - 1.0f;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: abi_smoke <libtsd_b200.so> [expect-gpu]\n");
    return 2;
  }
  const int expect_gpu = argc > 2 && !strcmp(argv[2], "expect-gpu");
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) {
    fprintf(stderr, "dlopen failed: %s\n", dlerror());
    return 2;
  }
  SYM(tsd_init) SYM(tsd_shutdown) SYM(tsd_last_error) SYM(tsd_conv2d) SYM(tsd_diffusion_create) SYM(tsd_diffusion_destroy)
  SYM(tsd_diffusion_init_random) SYM(tsd_diffusion_forward) SYM(tsd_diffusion_step) SYM(tsd_launch_count)
  SYM(tsd_dist_init) SYM(tsd_dist_broadcast_context) SYM(tsd_dist_generate) SYM(tsd_dist_shutdown)

  tsd_ctx* ctx = NULL;
  int32_t rc = p_tsd_init(0, &ctx);
  if (rc != TSD_OK) {
    const char* msg = p_tsd_last_error(NULL);
    printf("tsd_init: status %d (%s)\n", rc, msg ? msg : "");
    if (expect_gpu) return 1;
    return (rc == TSD_ERR_NO_DEVICE && msg && msg[0]) ? 0 : 1;  /* no device: a loud failure is the contract */
  }

  /* op level: 3x3 conv, 32 -> 16 channels at 8x8 against the loop nest of Conv2D.forward (utils.mojo:1764-1782) */
  enum { CI = 32, CO = 16, HW = 8 };
  static float x[CI * HW * HW], w[CO * CI * 9], b[CO], y[CO * HW * HW], ref[CO * HW * HW];
  uint32_t seed = 12345;
  for (int i = 0; i < CI * HW * HW; ++i) x[i] = frand(&seed);
  for (int i = 0; i < CO * CI * 9; ++i) w[i] = frand(&seed) / 17.0f;
  for (int i = 0; i < CO; ++i) b[i] = frand(&seed);
  for (int o = 0; o < CO; ++o)
    for (int i = 0; i < HW; ++i)
      for (int j = 0; j < HW; ++j) {
        double acc = b[o];
        for (int c = 0; c < CI; ++c)
          for (int u = 0; u < 3; ++u)
            for (int v = 0; v < 3; ++v) {
              const int ii = i + u - 1, jj = j + v - 1;
              if (ii < 0 || ii >= HW || jj < 0 || jj >= HW) continue;
              acc += (double)x[(c * HW + ii) * HW + jj] * w[((o * CI + c) * 3 + u) * 3 + v];
            }
        ref[(o * HW + i) * HW + j] = (float)acc;
      }
  rc = p_tsd_conv2d(ctx, x, 1, CI, HW, HW, w, b, CO, 3, 1, 1, y);
  if (rc != TSD_OK) {
    fprintf(stderr, "tsd_conv2d: %d %s\n", rc, p_tsd_last_error(ctx));
    return 1;
  }
  double emax = 0, rmax = 0;
  for (int i = 0; i < CO * HW * HW; ++i) {
    emax = fmax(emax, fabs((double)y[i] - ref[i]));
    rmax = fmax(rmax, fabs((double)ref[i]));
  }
  printf("tsd_conv2d: rel_linf %.2e\n", emax / rmax);
  if (!(emax / rmax < 5e-3)) return 1;  /* TF32 tolerance */

  /* model level */
  tsd_diffusion_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.latent_h = 8;
  cfg.latent_w = 8;
  cfg.max_batch = 2;
  cfg.context_len = 77;
  cfg.context_dim = 768;
  tsd_diffusion* m = NULL;
  rc = p_tsd_diffusion_create(ctx, &cfg, &m);
  if (rc == TSD_OK) rc = p_tsd_diffusion_init_random(m, 1234);
  if (rc != TSD_OK) {
    fprintf(stderr, "diffusion create/init: %d %s\n", rc, p_tsd_last_error(ctx));
    return 1;
  }
  static float lat[4 * 64], cx[2 * 77 * 768], temb[320], out1[4 * 64], out2[4 * 64], nz[4 * 64], next[4 * 64];
  for (int i = 0; i < 4 * 64; ++i) { lat[i] = frand(&seed); nz[i] = frand(&seed); }
  for (int i = 0; i < 2 * 77 * 768; ++i) cx[i] = frand(&seed);
  for (int i = 0; i < 320; ++i) temb[i] = i < 160 ? 1.0f : 0.0f;
  rc = p_tsd_diffusion_forward(m, lat, cx, 1, temb, 1, 1, out1);
  if (rc == TSD_OK) rc = p_tsd_diffusion_forward(m, lat, cx, 1, temb, 1, 1, out2);
  if (rc != TSD_OK) {
    fprintf(stderr, "tsd_diffusion_forward: %d %s\n", rc, p_tsd_last_error(ctx));
    return 1;
  }
  int finite = 1;
  for (int i = 0; i < 4 * 64; ++i) finite &= isfinite(out1[i]) != 0;
  if (!finite || memcmp(out1, out2, sizeof out1) != 0) {
    fprintf(stderr, "Diffusion.forward: not finite or graph replay != eager pass\n");
    return 1;
  }
  /* one loop iteration with CFG in a single call (cond + uncond rows) */
  rc = p_tsd_diffusion_step(m, lat, cx, 2, temb, nz, 1, 7.5f, 0.9f, 0.4359f, 0.1f, 0.9f, 0.05f, 1, next);
  if (rc != TSD_OK) {
    fprintf(stderr, "tsd_diffusion_step: %d %s\n", rc, p_tsd_last_error(ctx));
    return 1;
  }
  for (int i = 0; i < 4 * 64; ++i) finite &= isfinite(next[i]) != 0;
  printf("Diffusion.forward + step ok, kernels launched so far: %lld\n", (long long)p_tsd_launch_count(ctx));
  p_tsd_diffusion_destroy(m);
  p_tsd_shutdown(ctx);
  dlclose(lib);
  return finite ? 0 : 1;
}
