/* TEST INFRASTRUCTURE -- plain-C CPU restatement of the reference's two CUDA kernels.
 *
 *   oracle_bev_pool_v2_fwd  follows bev_pool_v2_kernel   (projects/mmdet3d_plugin/ops/
 *                           bev_pool_v2/src/bev_pool_cuda.cu:21-50)
 *   oracle_bev_pool_v2_bwd  follows bev_pool_grad_kernel (same file, 69-123)
 *
 * One CUDA thread of the reference == one iteration of the outer loop here, with the
 * same serial fp32 accumulation order inside (separate multiply and add are NOT used:
 * the reference compiles `psum += a * b` to an FFMA, so fmaf() is used to stay
 * bit-identical to what the GPU computes).  Optional OpenMP over the outer loop is
 * only for the timed cpu_baseline leg; results do not depend on the thread count.
 *
 * Not part of the product: only tests/, smoke() and bench.py's CPU legs may load it.
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

int oracle_set_threads(int n) {
  g_threads = n > 0 ? n : 1;
#ifdef _OPENMP
  return 1;
#else
  return 0;
#endif
}

void oracle_bev_pool_v2_fwd(int c, int n_intervals, const float *depth, const float *feat,
                            const int *ranks_depth, const int *ranks_feat, const int *ranks_bev,
                            const int *interval_starts, const int *interval_lengths, float *out) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 256)
  for (int iv = 0; iv < n_intervals; ++iv) {
    const int s = interval_starts[iv], n = interval_lengths[iv];
    float *o = out + (size_t)ranks_bev[s] * c;
    for (int ch = 0; ch < c; ++ch) {
      float acc = 0.f;
      for (int i = 0; i < n; ++i)
        acc = fmaf(feat[(size_t)ranks_feat[s + i] * c + ch], depth[ranks_depth[s + i]], acc);
      o[ch] = acc;
    }
  }
}

void oracle_bev_pool_v2_bwd(int c, int n_intervals, const float *out_grad, const float *depth,
                            const float *feat, const int *ranks_depth, const int *ranks_feat,
                            const int *ranks_bev, const int *interval_starts,
                            const int *interval_lengths, float *depth_grad, float *feat_grad) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 64)
  for (int iv = 0; iv < n_intervals; ++iv) {
    const int s = interval_starts[iv], n = interval_lengths[iv];
    for (int i = 0; i < n; ++i) {
      const float *g = out_grad + (size_t)ranks_bev[s + i] * c;
      const float *f = feat + (size_t)ranks_feat[s + i] * c;
      float acc = 0.f;
      for (int ch = 0; ch < c; ++ch) acc = fmaf(g[ch], f[ch], acc);
      depth_grad[ranks_depth[s + i]] = acc;
    }
    float *fg = feat_grad + (size_t)ranks_feat[s] * c;
    for (int ch = 0; ch < c; ++ch) {
      float acc = 0.f;
      for (int i = 0; i < n; ++i)
        acc = fmaf(out_grad[(size_t)ranks_bev[s + i] * c + ch], depth[ranks_depth[s + i]], acc);
      fg[ch] = acc;
    }
  }
}
