/*
 * CPU oracle for the sparse-LiDAR leg of the FusionDepth hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg, never by fusiondepth_b200/.
 *
 * Sequential plain-C restatement of
 *   kitti_utils.generate_depth_map      (reference kitti_utils.py:40-102, sub2ind 33-37)
 *   F.max_pool2d(.,2,ceil_mode=True) -> astype(float32) -> /100.0
 *                                        (kitti_dataset.py:105-107, mono_dataset.py:194-198)
 *   get_4beam_2channel                   (gen2channel.py:60-117)
 *
 * Parity pin: checked bit-for-bit against outputs of the reference functions run in the
 * build container (tests/make_golden.py -> tests/golden/lidar_*.npz).  The reference
 * itself holds no golden vectors for this path.
 *
 * The projection is an fp64 dot product per point: numpy's np.dot on [3,4]x[4,n]
 * resolves to a BLAS dgemm whose k-loop is an fma chain in k order, restated here with
 * fma() so the result does not depend on the host BLAS.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline double dot4(const double *P, double x, double y, double z) {
  /* k order 0..3, homogeneous coordinate forced to 1.0 (kitti_utils.py:10) */
  double acc = P[0] * x;
  acc = fma(P[1], y, acc);
  acc = fma(P[2], z, acc);
  acc = fma(P[3], 1.0, acc);
  return acc;
}

/* generate_depth_map without the final pad; depth is [H_im, W_im] fp64, zero-filled here.
 * Returns the number of points that landed inside the image. */
long fdo_depth_map(const float *pts, long n, const double *P /*3x4 row-major*/, int W_im,
                   int H_im, int vel_depth, double *depth) {
  long nvalid = 0;
  long npix = (long)H_im * W_im;
  memset(depth, 0, sizeof(double) * npix);
  int *pu = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
  int *pv = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
  double *pz = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  for (long i = 0; i < n; ++i) {
    double x = pts[4 * i], y = pts[4 * i + 1], z = pts[4 * i + 2];
    if (!(x >= 0)) continue;                                     /* :61 */
    double p0 = dot4(P, x, y, z), p1 = dot4(P + 4, x, y, z), p2 = dot4(P + 8, x, y, z);
    double u = p0 / p2, v = p1 / p2;                             /* :65 */
    double zz = vel_depth ? x : p2;                              /* :67-68 */
    u = rint(u) - 1;                                             /* np.round = half-to-even, :72-73 */
    v = rint(v) - 1;
    if (!(u >= 0 && v >= 0 && u < W_im && v < H_im)) continue;    /* :74-76 */
    pu[nvalid] = (int)u;
    pv[nvalid] = (int)v;
    pz[nvalid] = zz;
    ++nvalid;
  }
  /* fancy assignment: last point in file order wins (:79-80) */
  for (long i = 0; i < nvalid; ++i) depth[(long)pv[i] * W_im + pu[i]] = pz[i];
  /* duplicates by the quirky key v*(W-1)+u-1 (:33-37, 83-89) */
  long nkeys = (long)(H_im - 1) * (W_im - 1) + (W_im - 1) + 1 + 1; /* keys in [-1, max] */
  int *cnt = (int *)calloc(nkeys, sizeof(int));
  long *first = (long *)malloc(sizeof(long) * nkeys);
  double *mn = (double *)malloc(sizeof(double) * nkeys);
  for (long i = 0; i < nvalid; ++i) {
    long k = (long)pv[i] * (W_im - 1) + pu[i] - 1 + 1;
    if (cnt[k] == 0) {
      first[k] = i;
      mn[k] = pz[i];
    } else if (pz[i] < mn[k]) {
      mn[k] = pz[i];
    }
    cnt[k]++;
  }
  for (long k = 0; k < nkeys; ++k)
    if (cnt[k] > 1) depth[(long)pv[first[k]] * W_im + pu[first[k]]] = mn[k];
  for (long i = 0; i < npix; ++i)
    if (depth[i] < 0) depth[i] = 0;                              /* :90 */
  free(pu); free(pv); free(pz); free(cnt); free(first); free(mn);
  return nvalid;
}

/* the `shape` pad of generate_depth_map (:92-101): top |sh-H|, left xpad/2, right rest,
 * then drop two rows if sh < H.  out must hold out_h*out_w with
 * out_h = H + |sh-H| - (sh<H ? 2 : 0), out_w = sw. */
void fdo_pad(const double *depth, int H_im, int W_im, int sh, int sw, double *out) {
  int ypad = abs(sh - H_im), xpad = sw - W_im, xpad1 = xpad / 2;
  int crop = sh < H_im ? 2 : 0;
  int out_h = H_im + ypad - crop;
  memset(out, 0, sizeof(double) * (long)out_h * sw);
  for (int r = 0; r < H_im; ++r) {
    int ro = r + ypad - crop;
    if (ro < 0) continue;
    memcpy(out + (long)ro * sw + xpad1, depth + (long)r * W_im, sizeof(double) * W_im);
  }
}

/* max_pool2d(2, ceil_mode=True) on fp64 -> float32 -> /100.0f */
void fdo_pool_scale(const double *in, int H, int W, float *out) {
  int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  for (int i = 0; i < Ho; ++i)
    for (int j = 0; j < Wo; ++j) {
      double m = -INFINITY;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          int r = 2 * i + a, c = 2 * j + b;
          if (r < H && c < W) {
            double v = in[(long)r * W + c];
            if (v > m || v != v) m = v;
          }
        }
      out[(long)i * Wo + j] = (float)m / 100.0f;
    }
}

/* get_4beam_2channel: scatter form, raster order over the source window
 * rows [r0,r1) cols [c0,c1) (reference: 76..189, 2..637 at 192x640). */
void fdo_two_channel(const float *fb, int H, int W, int r0, int r1, int c0, int c1,
                     float *expanded, float *conf) {
  long n = (long)H * W;
  float *acc = (float *)calloc(n, sizeof(float));
  memset(expanded, 0, sizeof(float) * n);
  memset(conf, 0, sizeof(float) * n);
  for (int i = r0; i < r1; ++i)
    for (int j = c0; j < c1; ++j) {
      float v = fb[(long)i * W + j];
      if (v == 0) continue;
      long c = (long)i * W + j;
      expanded[c] = v; conf[c] = 1; acc[c] = 1;
      for (int dis = 1; dis <= 2; ++dis) {
        float cf = (float)(1.0 / (dis + 1));
        for (int hz = 1; hz <= dis; ++hz) {
          /* the four sign combinations in the reference's order (:72-114) */
          int xs[4] = {hz, -hz, hz, -hz};
          int ys[4] = {dis - hz, dis - hz, hz - dis, hz - dis};
          /* each later branch tests the x,y left behind by the branches before it */
          int x = hz, y = dis - hz;
          for (int b = 0; b < 4; ++b) {
            int go;
            if (b == 0) go = 1;
            else if (b == 1) go = (x != 0);
            else if (b == 2) go = (y != 0);
            else go = (x != 0 && y != 0);
            if (!go) continue;
            x = xs[b]; y = ys[b];
            long t = (long)(i + x) * W + (j + y);
            if (acc[t] == 0 || conf[t] < cf) {
              expanded[t] = v; conf[t] = cf; acc[t] = 1;
            } else if (conf[t] == cf) {
              expanded[t] += v; acc[t] += 1;
            }
          }
        }
      }
    }
  for (long t = 0; t < n; ++t) {
    float a = acc[t] == 0 ? 1.0f : acc[t];
    expanded[t] = expanded[t] / a;
  }
  free(acc);
}
