// mvr_camera.cu -- look_at_view_transform forward/backward + rotation validity check, one thread
// per view.  Replaces [upstream] renderer/cameras.py look_at_view_transform /
// camera_position_from_spherical_angles (renderer.py:79-80,122-123,168) and util.py:403-420.
#include "mvr_camera.cuh"

namespace mvr {

// One view: R, T, C written; false when the rotation fails the validity check (checked only when `check`).
__device__ __forceinline__ bool look_at_forward_view(const float* __restrict__ azim, const float* __restrict__ elev,
                                                     const float* __restrict__ dist, int i, float* __restrict__ R,
                                                     float* __restrict__ T, float* __restrict__ C, bool check) {
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
  bool ok = true;
  if (check) {
    // util.py:403-420: allclose(R R^T, I, atol=1e-6[, rtol=1e-5]) and allclose(det R, 1)
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
  }
  return ok;
}

__global__ void look_at_forward_kernel(const float* __restrict__ azim, const float* __restrict__ elev,
                                       const float* __restrict__ dist, int n, float* __restrict__ R,
                                       float* __restrict__ T, float* __restrict__ C, int* invalid_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!look_at_forward_view(azim, elev, dist, i, R, T, C, invalid_count != nullptr)) atomicAdd(invalid_count, 1);
}

// Up to a few thousand views (every MVTN batch): ONE CTA strides over them and owns the count, so the flag needs neither a memset in
// front of the kernel nor a copy behind it -- thread 0 stores it to invalid_count and, when given, straight into the caller's pinned
// host word (pinned memory is device-addressable under unified addressing; visible to the host once the kernel has completed).
__global__ void __launch_bounds__(256) look_at_forward_one_cta_kernel(const float* __restrict__ azim, const float* __restrict__ elev,
                                                                      const float* __restrict__ dist, int n, float* __restrict__ R,
                                                                      float* __restrict__ T, float* __restrict__ C,
                                                                      int* __restrict__ invalid_count, int* host_flag) {
  __shared__ int s_bad;
  pdl_enter();
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) bad += look_at_forward_view(azim, elev, dist, i, R, T, C, invalid_count != nullptr) ? 0 : 1;
  if (bad) atomicAdd(&s_bad, bad);
  __syncthreads();
  if (threadIdx.x == 0 && invalid_count) {
    *invalid_count = s_bad;
    if (host_flag) { *(volatile int*)host_flag = s_bad; __threadfence_system(); }
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
  float ga, ge, gd;
  look_at_backward_view(azim[i], elev[i], dist[i], gR ? gR + 9 * (size_t)i : nullptr, gT ? gT + 3 * (size_t)i : nullptr,
                        gC ? gC + 3 * (size_t)i : nullptr, ga, ge, gd);
  if (g_dist) g_dist[i] = gd;
  if (g_elev) g_elev[i] = ge;
  if (g_azim) g_azim[i] = ga;
}

}  // namespace mvr

using namespace mvr;

static const int kLookAtOneCtaMax = 4096;

// host_flag != NULL: the one-CTA kernel stores the count there as well (n <= kLookAtOneCtaMax only; the caller copies otherwise)
static int look_at_forward_impl(const float* azim, const float* elev, const float* dist, int n, float* R, float* T, float* C,
                                int* invalid_count, int* host_flag, void* stream) {
  if (n > 0 && n <= kLookAtOneCtaMax) {
    MVR_LAUNCH_PDL(look_at_forward_one_cta_kernel, 1, 256, 0, (cudaStream_t)stream, azim, elev, dist, n, R, T, C, invalid_count, host_flag);
    return mvr::check_launch("look_at_forward_one_cta_kernel");
  }
  if (invalid_count) {      // the call owns the flag: zeroed here, so that the caller does not pay a fill launch for it
    cudaError_t e = cudaMemsetAsync(invalid_count, 0, sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) { mvr::set_error("mvr_look_at_forward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (n == 0) return 0;
  MVR_LAUNCH(look_at_forward_kernel, (n + 127) / 128, 128, 0, (cudaStream_t)stream, azim, elev, dist, n, R, T, C, invalid_count);
  return mvr::check_launch("look_at_forward_kernel");
}

extern "C" int mvr_look_at_forward(const float* azim, const float* elev, const float* dist, int n, float* R,
                                   float* T, float* C, int* invalid_count, void* stream) {
  if (n < 0 || (n > 0 && (!azim || !elev || !dist || !R || !T))) { mvr::set_error("mvr_look_at_forward: null pointer or negative n"); return -1; }
  return look_at_forward_impl(azim, elev, dist, n, R, T, C, invalid_count, nullptr, stream);
}

// Is `p` pinned host memory that a kernel may store to through the same pointer?  (cudaHostAlloc / cudaHostRegister memory under
// unified addressing.)  Asked on every call (~1 us): an address can be freed and reused by pageable memory.
static bool host_flag_device_addressable(const int* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && a.devicePointer == (const void*)p;
}

// The same, and the validity flag on its way to the host behind the kernel: `host_flag` (pinned) receives invalid_count by an
// asynchronous copy and `event` (a cudaEvent_t) is recorded behind that copy -- a caller that checks the flag before it trusts the
// cameras (MVRenderer's rotation guard, util.py:403-420) then waits for this event only, and pays neither a copy call nor an event
// call of its own in front of the rasterizer launch (host time there is step time).
extern "C" int mvr_look_at_forward_flagged(const float* azim, const float* elev, const float* dist, int n, float* R, float* T,
                                           float* C, int* invalid_count, int* host_flag, void* event, void* stream) {
  if (!invalid_count || !host_flag || !event) { mvr::set_error("mvr_look_at_forward_flagged: null pointer"); return -2; }
  if (n < 0 || (n > 0 && (!azim || !elev || !dist || !R || !T))) { mvr::set_error("mvr_look_at_forward_flagged: null pointer or negative n"); return -1; }
  const bool direct = n > 0 && n <= kLookAtOneCtaMax && host_flag_device_addressable(host_flag);
  int rc = look_at_forward_impl(azim, elev, dist, n, R, T, C, invalid_count, direct ? host_flag : nullptr, stream);
  if (rc) return rc;
  cudaError_t e = cudaSuccess;
  if (!direct) e = cudaMemcpyAsync(host_flag, invalid_count, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream);
  if (e != cudaSuccess) { mvr::set_error("mvr_look_at_forward_flagged: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int mvr_look_at_backward(const float* azim, const float* elev, const float* dist, int n,
                                    const float* gR, const float* gT, const float* gC, float* g_azim,
                                    float* g_elev, float* g_dist, void* stream) {
  if (n < 0 || (n > 0 && (!azim || !elev || !dist))) { mvr::set_error("mvr_look_at_backward: null pointer or negative n"); return -1; }
  if (n == 0) return 0;
  MVR_LAUNCH(look_at_backward_kernel, (n + 127) / 128, 128, 0, (cudaStream_t)stream, azim, elev, dist, n, gR, gT, gC, g_azim, g_elev, g_dist);
  return mvr::check_launch("look_at_backward_kernel");
}
