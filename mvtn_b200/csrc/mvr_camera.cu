// mvr_camera.cu -- look_at_view_transform forward/backward + rotation validity check, one thread
// per view.  Replaces [upstream] renderer/cameras.py look_at_view_transform /
// camera_position_from_spherical_angles (renderer.py:79-80,122-123,168) and util.py:403-420.
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

__global__ void look_at_forward_kernel(const float* __restrict__ azim, const float* __restrict__ elev,
                                       const float* __restrict__ dist, int n, float* __restrict__ R,
                                       float* __restrict__ T, float* __restrict__ C, int* invalid_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float deg = (float)(3.14159265358979323846 / 180.0);
  const float e = deg * elev[i], a = deg * azim[i], d = dist[i];
  float se, ce, sa, ca;
  sincosf(e, &se, &ce);
  sincosf(a, &sa, &ca);
  float c[3] = {(d * ce) * sa, d * se, (d * ce) * ca};
  const float up[3] = {0.f, 1.f, 0.f};
  float mz[3] = {0.f - c[0], 0.f - c[1], 0.f - c[2]};
  float x[3], y[3], z[3], t[3];
  normalize3(mz, 1e-5f, z);
  cross3(up, z, t); normalize3(t, 1e-5f, x);
  cross3(z, x, t); normalize3(t, 1e-5f, y);
  if (fabsf(x[0]) <= 5e-3f && fabsf(x[1]) <= 5e-3f && fabsf(x[2]) <= 5e-3f) {
    cross3(y, z, t); normalize3(t, 1e-5f, x);
  }
  float r[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) { r[3 * k] = x[k]; r[3 * k + 1] = y[k]; r[3 * k + 2] = z[k]; }
#pragma unroll
  for (int k = 0; k < 9; ++k) R[9 * (size_t)i + k] = r[k];
#pragma unroll
  for (int j = 0; j < 3; ++j) T[3 * (size_t)i + j] = -((r[j] * c[0] + r[3 + j] * c[1]) + r[6 + j] * c[2]);
  if (C) { C[3 * (size_t)i] = c[0]; C[3 * (size_t)i + 1] = c[1]; C[3 * (size_t)i + 2] = c[2]; }
  if (invalid_count) {
    // util.py:403-420: allclose(R R^T, I, atol=1e-6[, rtol=1e-5]) and allclose(det R, 1)
    bool ok = true;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float s = (r[3 * p] * r[3 * q] + r[3 * p + 1] * r[3 * q + 1]) + r[3 * p + 2] * r[3 * q + 2];
        const float I = p == q ? 1.f : 0.f;
        if (!(fabsf(s - I) <= 1e-6f + 1e-5f * I)) ok = false;
      }
    const float det = r[0] * (r[4] * r[8] - r[5] * r[7]) - r[1] * (r[3] * r[8] - r[5] * r[6]) +
                      r[2] * (r[3] * r[7] - r[4] * r[6]);
    if (!(fabsf(det - 1.f) <= 1e-8f + 1e-5f)) ok = false;
    if (!ok) atomicAdd(invalid_count, 1);
  }
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

// Chain rule through T = -R^T C, R = [x y z], y = n(z x x), x = n(up x z), z = n(-C), C(d, e, a).
__global__ void look_at_backward_kernel(const float* __restrict__ azim, const float* __restrict__ elev,
                                        const float* __restrict__ dist, int n, const float* __restrict__ gR,
                                        const float* __restrict__ gT, const float* __restrict__ gC,
                                        float* __restrict__ g_azim, float* __restrict__ g_elev,
                                        float* __restrict__ g_dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float deg = (float)(3.14159265358979323846 / 180.0);
  const float e = deg * elev[i], a = deg * azim[i], d = dist[i];
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
  if (gT) { gt3[0] = gT[3 * (size_t)i]; gt3[1] = gT[3 * (size_t)i + 1]; gt3[2] = gT[3 * (size_t)i + 2]; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (gR) { g0 = gR[9 * (size_t)i + 3 * r]; g1 = gR[9 * (size_t)i + 3 * r + 1]; g2 = gR[9 * (size_t)i + 3 * r + 2]; }
    gx[r] = g0 - gt3[0] * c[r];
    gy[r] = g1 - gt3[1] * c[r];
    gz[r] = g2 - gt3[2] * c[r];
    gc[r] = -(gt3[0] * x[r] + gt3[1] * y[r] + gt3[2] * z[r]);
    if (gC) gc[r] += gC[3 * (size_t)i + r];
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
  if (g_dist) g_dist[i] = gd;
  if (g_elev) g_elev[i] = ge * deg;
  if (g_azim) g_azim[i] = ga * deg;
}

}  // namespace mvr

using namespace mvr;

extern "C" int mvr_look_at_forward(const float* azim, const float* elev, const float* dist, int n, float* R,
                                   float* T, float* C, int* invalid_count, void* stream) {
  if (n < 0 || (n > 0 && (!azim || !elev || !dist || !R || !T))) { mvr::set_error("mvr_look_at_forward: null pointer or negative n"); return -1; }
  if (invalid_count) {      // the call owns the flag: zeroed here, so that the caller does not pay a fill launch for it
    cudaError_t e = cudaMemsetAsync(invalid_count, 0, sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) { mvr::set_error("mvr_look_at_forward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (n == 0) return 0;
  MVR_LAUNCH(look_at_forward_kernel, (n + 127) / 128, 128, 0, (cudaStream_t)stream, azim, elev, dist, n, R, T, C, invalid_count);
  return mvr::check_launch("look_at_forward_kernel");
}

extern "C" int mvr_look_at_backward(const float* azim, const float* elev, const float* dist, int n,
                                    const float* gR, const float* gT, const float* gC, float* g_azim,
                                    float* g_elev, float* g_dist, void* stream) {
  if (n < 0 || (n > 0 && (!azim || !elev || !dist))) { mvr::set_error("mvr_look_at_backward: null pointer or negative n"); return -1; }
  if (n == 0) return 0;
  MVR_LAUNCH(look_at_backward_kernel, (n + 127) / 128, 128, 0, (cudaStream_t)stream, azim, elev, dist, n, gR, gT, gC, g_azim, g_elev, g_dist);
  return mvr::check_launch("look_at_backward_kernel");
}
