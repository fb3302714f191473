// mvr_camera.cuh -- the per-view camera arithmetic shared by mvr_camera.cu (look_at kernels) and the kernels that fuse it
// (points_backward_reduce_angles_kernel).  Every user is compiled with -fmad=false: the same IEEE sequence everywhere.
#pragma once
#include "mvr_common.cuh"

namespace mvr {

__device__ __forceinline__ void normalize3(const float v[3], float eps, float o[3]) {
  const float n = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  const float d = n > eps ? n : eps;
  o[0] = v[0] / d; o[1] = v[1] / d; o[2] = v[2] / d;
}
__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ void normalize_bwd(const float v[3], float eps, const float g[3], float gv[3]) {
  const float n = sqrtf(fmaf(v[0], v[0], fmaf(v[1], v[1], v[2] * v[2])));
  if (n > eps) {
    const float inv = 1.f / n;
    const float u[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
    const float d = fmaf(u[0], g[0], fmaf(u[1], g[1], u[2] * g[2]));
#pragma unroll
    for (int i = 0; i < 3; ++i) gv[i] = (g[i] - u[i] * d) * inv;
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) gv[i] = g[i] / eps;
  }
}

// Chain rule through T = -R^T C, R = [x y z], y = n(z x x), x = n(up x z), z = n(-C), C(d, e, a): the gradients of ONE view.
// gR (9, row-major), gT (3), gC (3): any may be NULL.  Angles in degrees, as look_at_view_transform takes them.
__device__ __forceinline__ void look_at_backward_view(float azim_deg, float elev_deg, float dist, const float* gR, const float* gT,
                                                      const float* gC, float& g_azim, float& g_elev, float& g_dist) {
  const float deg = (float)(3.14159265358979323846 / 180.0);
  const float e = deg * elev_deg, a = deg * azim_deg, d = dist;
  float se, ce, sa, ca;
  sincosf(e, &se, &ce);
  sincosf(a, &sa, &ca);
  const float c[3] = {(d * ce) * sa, d * se, (d * ce) * ca};
  const float up[3] = {0.f, 1.f, 0.f};
  const float mz[3] = {-c[0], -c[1], -c[2]};
  float x[3], y[3], z[3], tx[3], ty[3], txr[3], x0[3];
  normalize3(mz, 1e-5f, z);
  cross3(up, z, tx); normalize3(tx, 1e-5f, x);
  cross3(z, x, ty); normalize3(ty, 1e-5f, y);
  x0[0] = x[0]; x0[1] = x[1]; x0[2] = x[2];
  const bool replaced = fabsf(x[0]) <= 5e-3f && fabsf(x[1]) <= 5e-3f && fabsf(x[2]) <= 5e-3f;
  if (replaced) { cross3(y, z, txr); normalize3(txr, 1e-5f, x); }
  float gx[3] = {0, 0, 0}, gy[3] = {0, 0, 0}, gz[3] = {0, 0, 0}, gc[3] = {0, 0, 0};
  float gt3[3] = {0, 0, 0};
  if (gT) { gt3[0] = gT[0]; gt3[1] = gT[1]; gt3[2] = gT[2]; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (gR) { g0 = gR[3 * r]; g1 = gR[3 * r + 1]; g2 = gR[3 * r + 2]; }
    gx[r] = g0 - gt3[0] * c[r];
    gy[r] = g1 - gt3[1] * c[r];
    gz[r] = g2 - gt3[2] * c[r];
    gc[r] = -(gt3[0] * x[r] + gt3[1] * y[r] + gt3[2] * z[r]);
    if (gC) gc[r] += gC[r];
  }
  float gt[3], tmp[3];
  if (replaced) {
    normalize_bwd(txr, 1e-5f, gx, gt);
    cross3(z, gt, tmp); gy[0] += tmp[0]; gy[1] += tmp[1]; gy[2] += tmp[2];
    cross3(gt, y, tmp); gz[0] += tmp[0]; gz[1] += tmp[1]; gz[2] += tmp[2];
    gx[0] = gx[1] = gx[2] = 0.f;
  }
  normalize_bwd(ty, 1e-5f, gy, gt);
  cross3(x0, gt, tmp); gz[0] += tmp[0]; gz[1] += tmp[1]; gz[2] += tmp[2];
  cross3(gt, z, tmp); gx[0] += tmp[0]; gx[1] += tmp[1]; gx[2] += tmp[2];
  normalize_bwd(tx, 1e-5f, gx, gt);
  cross3(gt, up, tmp); gz[0] += tmp[0]; gz[1] += tmp[1]; gz[2] += tmp[2];
  normalize_bwd(mz, 1e-5f, gz, gt);
  gc[0] -= gt[0]; gc[1] -= gt[1]; gc[2] -= gt[2];
  const float gd = gc[0] * ce * sa + gc[1] * se + gc[2] * ce * ca;
  const float ge = gc[0] * (-d * se * sa) + gc[1] * (d * ce) + gc[2] * (-d * se * ca);
  const float ga = gc[0] * (d * ce * ca) + gc[2] * (-d * ce * sa);
  g_dist = gd; g_elev = ge * deg; g_azim = ga * deg;
}

}  // namespace mvr
