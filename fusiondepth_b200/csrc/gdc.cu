// Graph-based Depth Correction on the device (reference gdc_old.py:74-250, called per frame
// from inf_gdc.py:81 with k = 10, W_tol = 3e-5, recon_tol = 5e-4, method = 'cg' -- on the host it is a KD-tree query,
// N dense (k+2)x(k+2) solves in numpy and scipy's conjugate gradients on sparse matrices, minutes of CPU per
// sequence between stage 1 and stage 2).  Everything is fp64, as in the reference.
//
//   gdc_select_kernel    per pixel: back-projection with the predicted depth (kitti_util_from_pse.py:204-215),
//                        range / pitch filters (gdc_old.py:18-26, 54-62), the |pred - gt| < 2 anchor test (:144)
//                        -> class 0 (untouched) / 1 (pseudo-LiDAR point to correct) / 2 (LiDAR anchor) + point
//   gdc_knn_kernel       exact k nearest neighbours (brute force through shared-memory tiles, the query itself
//                        excluded), ascending distance -- what KDTree.query(k + 1)[:, 1:] returns (:171-172)
//   gdc_weights_kernel   per point the (k+2) x (k+2) constrained reconstruction system (:174-186), LU with partial
//                        pivoting in local memory
//   gdc_apply_kernel     y = A x   with A = [I - W_PLPL ; W_PLL] read straight from (neighbours, W)     (:221)
//   gdc_apply_t_kernel   z = A^T y through a column-major entry list (deterministic summation order)
//   gdc_rhs_kernel       b = [W_LPL gt ; gt - W_LL gt]                                                   (:222)
//   gdc_dot_kernel, gdc_cg_update_kernel, gdc_cg_dir_kernel   the conjugate-gradient recurrence of
//                        scipy.sparse.linalg.cg on device scalars (one block: fixed summation order)
#include "common.cuh"
#include "../../include/fusiondepth_b200.h"

namespace {

constexpr int GDC_MAXK = 16;

__global__ void gdc_select_kernel(const double* __restrict__ pred, const double* __restrict__ gt, int H, int W,
                                  double c_u, double c_v, double f_u, double f_v, double b_x, double b_y,
                                  double th_lo, double th_hi, unsigned char* __restrict__ cls,
                                  double* __restrict__ pts) {
  const long n = (long)H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double u = (double)(i % W), v = (double)(i / W);
    auto backproject = [&](double d, double& x, double& y) {
      x = __dadd_rn(__ddiv_rn(__dmul_rn(u - c_u, d), f_u), b_x);
      y = __dadd_rn(__ddiv_rn(__dmul_rn(v - c_v, d), f_v), b_y);
    };
    auto in_range = [](double x, double y, double z) {
      return z < 80.0 && z > 1.0 && x < 40.0 && x >= -40.0 && y < 2.5 && y >= -1.0;
    };
    const double d = pred[i], g = gt[i];
    double x, y, xg, yg;
    backproject(d, x, y);
    backproject(g, xg, yg);
    const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(d, d)));
    const double theta = asin(__ddiv_rn(y, r));
    const bool consider_pl = in_range(x, y, d) && theta >= th_lo && theta < th_hi;      // NaN compares false
    const bool anchor = consider_pl && in_range(xg, yg, g) && fabs(d - g) < 2.0;
    cls[i] = anchor ? 2 : (consider_pl ? 1 : 0);
    pts[3 * i] = x; pts[3 * i + 1] = y; pts[3 * i + 2] = d;
  }
}

// One thread per query point; the candidates stream through shared memory in tiles of blockDim.x points.
template <int K>
__global__ void __launch_bounds__(128) gdc_knn_kernel(const double* __restrict__ pts, int n, int* __restrict__ nbr) {
  __shared__ double sx[128], sy[128], sz[128];
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  double qx = 0, qy = 0, qz = 0;
  if (q < n) { qx = pts[3 * q]; qy = pts[3 * q + 1]; qz = pts[3 * q + 2]; }
  double bd[K];
  int bi[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { bd[j] = 1e300; bi[j] = -1; }
  for (int t0 = 0; t0 < n; t0 += blockDim.x) {
    const int c = t0 + threadIdx.x;
    if (c < n) { sx[threadIdx.x] = pts[3 * c]; sy[threadIdx.x] = pts[3 * c + 1]; sz[threadIdx.x] = pts[3 * c + 2]; }
    __syncthreads();
    const int lim = min((int)blockDim.x, n - t0);
    if (q < n) {
      for (int j = 0; j < lim; ++j) {
        const int c2 = t0 + j;
        const double dx = sx[j] - qx, dy = sy[j] - qy, dz = sz[j] - qz;
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (c2 != q && d2 < bd[K - 1]) {
          // insert (ascending distance; equal distances keep the lower index first)
          double cd = d2;
          int ci = c2;
#pragma unroll
          for (int s = 0; s < K; ++s) {
            if (cd < bd[s]) {
              const double td = bd[s]; const int ti = bi[s];
              bd[s] = cd; bi[s] = ci;
              cd = td; ci = ti;
            }
          }
        }
      }
    }
    __syncthreads();
  }
  if (q < n) {
#pragma unroll
    for (int j = 0; j < K; ++j) nbr[(long)q * K + j] = bi[j];
  }
}

// (k+2) x (k+2) system of gdc_old.py:174-184 per point, solved by LU with partial pivoting
__global__ void gdc_weights_kernel(const double* __restrict__ x_info, const int* __restrict__ nbr, int n, int k,
                                   double w_tol, double* __restrict__ Wout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int M = GDC_MAXK + 2;
  double A[M][M], b[M];
  const int m = k + 2;
  for (int r = 0; r < m; ++r) {
    b[r] = 0.0;
    for (int c = 0; c < m; ++c) A[r][c] = 0.0;
  }
  for (int j = 0; j < k; ++j) {
    const double xn = x_info[nbr[(long)i * k + j]];
    A[j][j] = 1.0 + w_tol;
    A[k + 1][j] = 1.0;
    A[j][k + 1] = 1.0;
    A[k][j] = xn;
    A[j][k] = xn;
  }
  b[k] = x_info[i];
  b[k + 1] = 1.0;
  for (int c = 0; c < m; ++c) {
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < m; ++r)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (piv != c) {
      for (int cc = 0; cc < m; ++cc) { const double t = A[c][cc]; A[c][cc] = A[piv][cc]; A[piv][cc] = t; }
      const double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1; r < m; ++r) {
      const double f = A[r][c] * inv;
      if (f != 0.0) {
        for (int cc = c + 1; cc < m; ++cc) A[r][cc] -= f * A[c][cc];
        b[r] -= f * b[c];
      }
    }
  }
  for (int r = m - 1; r >= 0; --r) {
    double s = b[r];
    for (int cc = r + 1; cc < m; ++cc) s -= A[r][cc] * b[cc];
    b[r] = s / A[r][r];
  }
  for (int j = 0; j < k; ++j) Wout[(long)i * k + j] = b[j];
}

// y [n] = A x,  x [n_pl]:  row i < n_pl: x_i - sum_{nbr < n_pl} W x_nbr;  row i >= n_pl: + sum
__global__ void gdc_apply_kernel(const double* __restrict__ Wm, const int* __restrict__ nbr, int n, int n_pl, int k,
                                 const double* __restrict__ x, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int j = 0; j < k; ++j) {
    const int c = nbr[(long)i * k + j];
    if (c < n_pl) s = __dadd_rn(s, __dmul_rn(Wm[(long)i * k + j], x[c]));
  }
  y[i] = i < n_pl ? __dsub_rn(x[i], s) : s;
}

// z [n_pl] = A^T y: column c gathers its entries e (flattened (row, slot) indices sorted by column, rows ascending)
__global__ void gdc_apply_t_kernel(const double* __restrict__ Wm, const long* __restrict__ entry,
                                   const long* __restrict__ col_ptr, int n_pl, int k, const double* __restrict__ y,
                                   double* __restrict__ z) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_pl) return;
  double s = y[c];
  for (long e = col_ptr[c]; e < col_ptr[c + 1]; ++e) {
    const long f = entry[e];
    const int row = (int)(f / k);
    const double t = __dmul_rn(Wm[f], y[row]);
    s = row < n_pl ? __dsub_rn(s, t) : __dadd_rn(s, t);
  }
  z[c] = s;
}

// b [n] = [W_LPL gt ; gt - W_LL gt],  gt_info [n - n_pl]
__global__ void gdc_rhs_kernel(const double* __restrict__ Wm, const int* __restrict__ nbr, int n, int n_pl, int k,
                               const double* __restrict__ gt_info, double* __restrict__ b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int j = 0; j < k; ++j) {
    const int c = nbr[(long)i * k + j];
    if (c >= n_pl) s = __dadd_rn(s, __dmul_rn(Wm[(long)i * k + j], gt_info[c - n_pl]));
  }
  b[i] = i < n_pl ? s : __dsub_rn(gt_info[i - n_pl], s);
}

// out[0] = sum a_i b_i, one block of 1024 threads (fixed summation order: bit-reproducible)
__device__ double block_dot(const double* __restrict__ a, const double* __restrict__ b, int n, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = __dadd_rn(s, __dmul_rn(a[i], b[i]));
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] = __dadd_rn(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  return red[0];
}
__global__ void __launch_bounds__(1024) gdc_dot_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                       int n, double* __restrict__ out) {
  __shared__ double red[1024];
  const double s = block_dot(a, b, n, red);
  if (threadIdx.x == 0) out[0] = s;
}
// alpha = rho / pq;  x += alpha p;  r -= alpha q;  scal[2] = r.r   (scal = {rho, pq, rho_new})
__global__ void __launch_bounds__(1024) gdc_cg_update_kernel(double* __restrict__ x, double* __restrict__ r,
                                                             const double* __restrict__ p,
                                                             const double* __restrict__ q, int n,
                                                             double* __restrict__ scal) {
  __shared__ double red[1024];
  const double alpha = __ddiv_rn(scal[0], scal[1]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    x[i] = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));
    r[i] = __dsub_rn(r[i], __dmul_rn(alpha, q[i]));
  }
  __syncthreads();
  const double s = block_dot(r, r, n, red);
  if (threadIdx.x == 0) scal[2] = s;
}
// p = r + (rho_new / rho) p;  rho = rho_new
__global__ void __launch_bounds__(1024) gdc_cg_dir_kernel(double* __restrict__ p, const double* __restrict__ r, int n,
                                                          double* __restrict__ scal) {
  const double beta = __ddiv_rn(scal[2], scal[0]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = __dadd_rn(__dmul_rn(beta, p[i]), r[i]);
  __syncthreads();
  if (threadIdx.x == 0) scal[0] = scal[2];
}

}  // namespace

extern "C" {

int fd_gdc_select(const double* pred, const double* gt, int H, int W, const double* calib6_host, double th_lo,
                  double th_hi, unsigned char* cls, double* points, void* stream) {
  FD_REQUIRE(H > 0 && W > 0 && calib6_host, "fd_gdc_select: bad arguments");
  const long n = (long)H * W;
  gdc_select_kernel<<<(int)fd::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
      pred, gt, H, W, calib6_host[0], calib6_host[1], calib6_host[2], calib6_host[3], calib6_host[4], calib6_host[5],
      th_lo, th_hi, cls, points);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_gdc_knn(const double* points, int n, int k, int* neighbors, void* stream) {
  FD_REQUIRE(n > k && (k == 10 || k == 8 || k == 12 || k == 16), "fd_gdc_knn: k must be 8, 10, 12 or 16 and n > k");
  const int blocks = (int)fd::cdiv(n, 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 10) gdc_knn_kernel<10><<<blocks, 128, 0, st>>>(points, n, neighbors);
  else if (k == 8) gdc_knn_kernel<8><<<blocks, 128, 0, st>>>(points, n, neighbors);
  else if (k == 12) gdc_knn_kernel<12><<<blocks, 128, 0, st>>>(points, n, neighbors);
  else gdc_knn_kernel<16><<<blocks, 128, 0, st>>>(points, n, neighbors);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_gdc_weights(const double* x_info, const int* neighbors, int n, int k, double w_tol, double* weights,
                   void* stream) {
  FD_REQUIRE(n > 0 && k > 0 && k <= GDC_MAXK, "fd_gdc_weights: k <= %d", GDC_MAXK);
  gdc_weights_kernel<<<(int)fd::cdiv(n, 64), 64, 0, (cudaStream_t)stream>>>(x_info, neighbors, n, k, w_tol, weights);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_gdc_rhs(const double* weights, const int* neighbors, int n, int n_pl, int k, const double* gt_info, double* b,
               void* stream) {
  gdc_rhs_kernel<<<(int)fd::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(weights, neighbors, n, n_pl, k, gt_info, b);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_gdc_apply(const double* weights, const int* neighbors, int n, int n_pl, int k, const double* x, double* y,
                 void* stream) {
  gdc_apply_kernel<<<(int)fd::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(weights, neighbors, n, n_pl, k, x, y);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_gdc_apply_t(const double* weights, const long* entries, const long* col_ptr, int n_pl, int k, const double* y,
                   double* z, void* stream) {
  gdc_apply_t_kernel<<<(int)fd::cdiv(n_pl, 256), 256, 0, (cudaStream_t)stream>>>(weights, entries, col_ptr, n_pl, k,
                                                                                y, z);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_gdc_dot(const double* a, const double* b, int n, double* out, void* stream) {
  gdc_dot_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a, b, n, out);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_gdc_cg_update(double* x, double* r, const double* p, const double* q, int n, double* scal, void* stream) {
  gdc_cg_update_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, r, p, q, n, scal);
  FD_CHECK_LAUNCH();
  return 0;
}

int fd_gdc_cg_dir(double* p, const double* r, int n, double* scal, void* stream) {
  gdc_cg_dir_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p, r, n, scal);
  FD_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
