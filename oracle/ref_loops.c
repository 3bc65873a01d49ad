/*
 * oracle/ref_loops.c - TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the scalar loop nests of lrmantovani10/Stable-Diffusion.mojo for the
 * denoising hot path, in the reference's own accumulation order and fp32 arithmetic, with
 * OpenMP over exactly the axes the reference hands to `parallelize`.  PARITY UNPINNED: the
 * reference ships no tests / golden vectors and cannot be built here (no Mojo toolchain), see
 * DESIGN.md.  Semantics follow SURVEY.md section 0 (memory-safety accidents of the reference are
 * not reproduced: no per-patch heap allocation, correct slicing).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libtsd_b200.so) never links or calls it.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

API int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Launchers such as torchrun export OMP_NUM_THREADS=1 to every rank; the CPU baseline is defined on ALL host cores,
 * so bench.py sets the team size explicitly.  n <= 0: the number of processors OpenMP sees.  Returns the new size. */
API int ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n <= 0) n = omp_get_num_procs();
  omp_set_dynamic(0);
  omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* Conv2D.forward, helpers/utils.mojo:1738-1811.
 * pad (utils.mojo:1383-1413) then, parallel over out channels (:1809), for every (y, x) of
 * tile_2d (:405-409, :1788-1807), for every in channel (:1771): 3x3 patch * kernel plane,
 * summed row-major (multiply().sum(), :1417, :1360-1366), accumulated over channels (:1780);
 * output = sum + bias[oc] (:1782).  x (cin,h,w), w OIHW, out (cout,ho,wo). */
API void ref_conv2d(const float* x, int cin, int h, int w, const float* wt, const float* bias, int cout,
                    int k, int pad, int stride, float* out) {
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  const int ho = (hp - k) / stride + 1, wo = (wp - k) / stride + 1;
  float* xp = (float*)calloc((size_t)cin * hp * wp, sizeof(float));
  for (int c = 0; c < cin; ++c)
    for (int i = 0; i < h; ++i)
      memcpy(xp + ((size_t)c * hp + i + pad) * wp + pad, x + ((size_t)c * h + i) * w, (size_t)w * sizeof(float));
#pragma omp parallel for schedule(dynamic, 1)
  for (int oc = 0; oc < cout; ++oc) {
    const float* kern = wt + (size_t)oc * cin * k * k;
    for (int yo = 0; yo < ho; ++yo) {
      for (int xo = 0; xo < wo; ++xo) {
        const int y = yo * stride, xx = xo * stride;
        float conv_sum = 0.0f;
        for (int ic = 0; ic < cin; ++ic) {
          const float* plane = xp + (size_t)ic * hp * wp;
          const float* kp = kern + (size_t)ic * k * k;
          float patch = 0.0f;
          for (int u = 0; u < k; ++u)
            for (int v = 0; v < k; ++v) patch += plane[(size_t)(y + u) * wp + xx + v] * kp[u * k + v];
          conv_sum += patch;
        }
        out[((size_t)oc * ho + yo) * wo + xo] = conv_sum + (bias ? bias[oc] : 0.0f);
      }
    }
  }
  free(xp);
}

/* Matrix.matmul, helpers/utils.mojo:1549-1569: per channel c, rows in parallel (:1566), k outer,
 * n inner, new[c,m,n] += a[c,m,k] * b[c,k,n] (:1561-1565).  a (c,m,k), b (c,k,n), out (c,m,n). */
API void ref_matmul(const float* a, const float* b, int c, int m, int k, int n, float* out) {
  for (int ch = 0; ch < c; ++ch) {
    const float* A = a + (size_t)ch * m * k;
    const float* B = b + (size_t)ch * k * n;
    float* O = out + (size_t)ch * m * n;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i) {
      float* orow = O + (size_t)i * n;
      for (int j = 0; j < n; ++j) orow[j] = 0.0f;
      for (int kk = 0; kk < k; ++kk) {
        const float av = A[(size_t)i * k + kk];
        const float* brow = B + (size_t)kk * n;
        for (int j = 0; j < n; ++j) orow[j] += av * brow[j];
      }
    }
  }
}

/* Linear.forward, helpers/utils.mojo:1954-1976: x.matmul(weight.transpose(1,2)) then + bias on
 * every column.  x (rows,in), w (out,in), y (rows,out). */
API void ref_linear(const float* x, int rows, int in_f, const float* w, const float* bias, int out_f, float* y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < rows; ++i) {
    float* yr = y + (size_t)i * out_f;
    for (int j = 0; j < out_f; ++j) yr[j] = 0.0f;
    /* k outer, n inner as Matrix.matmul; w^T[k][n] = w[n][k] */
    for (int kk = 0; kk < in_f; ++kk) {
      const float av = x[(size_t)i * in_f + kk];
      for (int j = 0; j < out_f; ++j) yr[j] += av * w[(size_t)j * in_f + kk];
    }
    if (bias)
      for (int j = 0; j < out_f; ++j) yr[j] += bias[j];
  }
}

/* Softmax, helpers/utils.mojo:411-448: exp with no max subtraction (:413); dim=2 divides each
 * column m[ch,:,col] by its sum (:435-445); dim=1 each row (:423-433).  In place. */
API void ref_softmax(float* m, int c, int r, int cols, int dim) {
  const size_t plane = (size_t)r * cols;
#pragma omp parallel for schedule(static)
  for (int ch = 0; ch < c; ++ch) {
    float* p = m + ch * plane;
    for (size_t i = 0; i < plane; ++i) p[i] = expf(p[i]);
    if (dim == 2) {
      for (int j = 0; j < cols; ++j) {
        float s = 0.0f;
        for (int i = 0; i < r; ++i) s += p[(size_t)i * cols + j];
        for (int i = 0; i < r; ++i) p[(size_t)i * cols + j] /= s;
      }
    } else {
      for (int i = 0; i < r; ++i) {
        float s = 0.0f;
        for (int j = 0; j < cols; ++j) s += p[(size_t)i * cols + j];
        for (int j = 0; j < cols; ++j) p[(size_t)i * cols + j] /= s;
      }
    }
  }
}

/* GroupNorm.forward, helpers/utils.mojo:1845-1885 with Matrix.sum/mean/std (:1360-1380): per
 * group one serial fp32 sum, mean = sum/N, std = sqrt(sum((x-mean)^2)/N) (biased, :1380),
 * y = (x-mean)/(std+eps)*gamma, gamma = 1 (:1833, :1868-1870).  x,y (c,h*w). */
API void ref_groupnorm(const float* x, int c, int hw, int groups, float eps, float* y) {
  const int cpg = c / groups;
  const size_t n = (size_t)cpg * hw;
#pragma omp parallel for schedule(static)
  for (int g = 0; g < groups; ++g) {
    const float* xg = x + (size_t)g * n;
    float* yg = y + (size_t)g * n;
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) sum += xg[i];
    const float mean = sum / (float)n;
    float sq = 0.0f;
    for (size_t i = 0; i < n; ++i) {
      const float d = xg[i] - mean;
      sq += d * d;
    }
    const float sd = sqrtf(sq / (float)n);
    for (size_t i = 0; i < n; ++i) yg[i] = (xg[i] - mean) / (sd + eps) * 1.0f;
  }
}

/* SiLU.forward, helpers/utils.mojo:1892-1902: x / (1 + exp(-x)). */
API void ref_silu(const float* x, size_t n, float* y) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) y[i] = x[i] / (1.0f + expf(-x[i]));
}

/* Gelu.forward, helpers/utils.mojo:1908-1919: x * 0.5 * (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))). */
API void ref_gelu(const float* x, size_t n, float* y) {
  const float k = sqrtf(2.0f / 3.14159265358979323846f);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) {
    const float v = x[i];
    const float cdf = 0.5f * (1.0f + tanhf(k * (v + 0.044715f * v * v * v)));
    y[i] = v * cdf;
  }
}

/* Upsample.forward under contract Q8 (helpers/utils.mojo:1989-2010 never launches its closure):
 * nearest x2 spatial, y[c,i,j] = x[c,i/2,j/2]. */
API void ref_upsample2x(const float* x, int c, int h, int w, float* y) {
#pragma omp parallel for schedule(static)
  for (int ch = 0; ch < c; ++ch)
    for (int i = 0; i < 2 * h; ++i)
      for (int j = 0; j < 2 * w; ++j)
        y[((size_t)ch * 2 * h + i) * 2 * w + j] = x[((size_t)ch * h + i / 2) * w + j / 2];
}

/* Attention core of Self_Attention / Cross_Attention.forward, helpers/attention.mojo:46-62 and
 * 105-115: per head (raw-reshape split, Q4) weight = q.matmul(k^T) / sqrt(d), Softmax(dim=2)
 * (or dim=1 when key_axis != 0), weight.matmul(v), merge by transpose(0,1)+reshape.
 * q (heads,tq,d), k,v (heads,tk,d), out (tq, heads*d).  The T x T score plane of one head is
 * the only temporary (the reference keeps all heads alive at once). */
API void ref_attention_core(const float* q, const float* k, const float* v, int heads, int tq, int tk, int d,
                            int key_axis, float* out) {
  const float sq = sqrtf((float)d);
  float* s = (float*)malloc((size_t)tq * tk * sizeof(float));
  float* kt = (float*)malloc((size_t)tk * d * sizeof(float));
  float* o = (float*)malloc((size_t)tq * d * sizeof(float));
  for (int hd = 0; hd < heads; ++hd) {
    const float* K = k + (size_t)hd * tk * d;
    for (int j = 0; j < tk; ++j)
      for (int e = 0; e < d; ++e) kt[(size_t)e * tk + j] = K[(size_t)j * d + e];
    ref_matmul(q + (size_t)hd * tq * d, kt, 1, tq, d, tk, s);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)tq * tk; ++i) s[i] = s[i] / sq; /* weight / sqrt(d), attention.mojo:58 */
    ref_softmax(s, 1, tq, tk, key_axis ? 1 : 2);
    ref_matmul(s, v + (size_t)hd * tk * d, 1, tq, tk, d, o);
    for (int i = 0; i < tq; ++i) memcpy(out + ((size_t)i * heads + hd) * d, o + (size_t)i * d, (size_t)d * sizeof(float));
  }
  free(s);
  free(kt);
  free(o);
}
