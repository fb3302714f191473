// mvr_common.cuh -- shared device helpers for libmvr_b200 (sm_100a).
//
// Arithmetic contract (DESIGN.md "Parity"): everything that decides a fragment (projection,
// edge functions, barycentrics, depth, point distances) is IEEE fp32 in the written operation
// order.  The translation units are compiled with -fmad=false so that a*b+c is never contracted;
// where an FMA is wanted (shading, gradients: tolerance-compared) it is written as fmaf().
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvr_b200.h"

#define MVR_K_EPS 1e-8f
#define MVR_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define MVR_THREADS 256

namespace mvr {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
void prof_begin(const char* name, cudaStream_t st);
void prof_end(const char* name, cudaStream_t st);

// every kernel launch goes through this: counts it (mvr_launch_count) and, when bench.py asked for it,
// brackets it with CUDA events on the launching stream (mvr_profile_enable / mvr_profile_collect).
#define MVR_LAUNCH(kernel, grid, block, smem, st, ...)          \
  do {                                                          \
    mvr::prof_begin(#kernel, st);                               \
    kernel<<<grid, block, smem, st>>>(__VA_ARGS__);             \
    mvr::prof_end(#kernel, st);                                 \
  } while (0)

// Programmatic dependent launch (on unless MVR_PDL=0): a kernel launched through MVR_LAUNCH_PDL may be scheduled while its predecessor in
// the stream is still draining; pdl_enter() -- the FIRST statement of such a kernel -- blocks until the predecessor has completed
// and its writes are visible, then lets the kernel behind this one be scheduled in turn.  Only the launch latency is hidden; no
// kernel touches memory before its predecessor is done (0.800 -> 0.786 ms per resident mesh step at BASELINE configs[1], six edges).
// Not used on a capturing stream.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled(cudaStream_t st);
#define MVR_LAUNCH_PDL(kernel, grid, block, smem, st, ...)                                   \
  do {                                                                                       \
    mvr::prof_begin(#kernel, st);                                                            \
    if (mvr::pdl_enabled(st)) {                                                              \
      cudaLaunchConfig_t cfg_ = {};                                                          \
      cfg_.gridDim = dim3(grid); cfg_.blockDim = dim3(block);                                \
      cfg_.dynamicSmemBytes = smem; cfg_.stream = st;                                        \
      cudaLaunchAttribute at_[1];                                                            \
      at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                        \
      at_[0].val.programmaticStreamSerializationAllowed = 1;                                 \
      cfg_.attrs = at_; cfg_.numAttrs = 1;                                                   \
      cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                                        \
    } else {                                                                                 \
      kernel<<<grid, block, smem, st>>>(__VA_ARGS__);                                        \
    }                                                                                        \
    mvr::prof_end(#kernel, st);                                                              \
  } while (0)

// [upstream] rasterization_utils PixToNonSquareNdc -- NDC coordinate of the centre of pixel i.
__host__ __device__ __forceinline__ float pix_to_ndc(int i, int S1, int S2) {
  float range = 2.0f;
  if (S1 > S2) range = ((float)(S1 / S2)) * range;
  const float offset = range / 2.0f;
  return -offset + (range * (float)i + offset) / (float)S1;
}

// Inclusive range [lo, hi] of flipped pixel indices j (j = S-1-i) whose centre c(j) satisfies
// vmin <= c(j) <= vmax; empty when lo > hi.  c(j) is strictly increasing in j.  The float
// estimate is corrected with exact evaluations so the range is exact, not conservative.
__device__ __forceinline__ void ndc_range_to_pix(float vmin, float vmax, int S1, int S2, int& lo, int& hi) {
  float range = 2.0f;
  if (S1 > S2) range = ((float)(S1 / S2)) * range;
  const float offset = range / 2.0f;
  // c(j) = -offset + (range*j + offset)/S1  =>  j = ((c + offset)*S1 - offset)/range
  float flo = ceilf(((vmin + offset) * (float)S1 - offset) / range);
  float fhi = floorf(((vmax + offset) * (float)S1 - offset) / range);
  flo = fminf(fmaxf(flo, -1.0f), (float)S1);
  fhi = fminf(fmaxf(fhi, -1.0f), (float)S1);
  lo = (int)flo; hi = (int)fhi;
  if (lo < 0) lo = 0;
  if (hi > S1 - 1) hi = S1 - 1;
  // exact fix-up (at most a step or two)
  while (lo > 0 && pix_to_ndc(lo - 1, S1, S2) >= vmin) --lo;
  while (lo <= S1 - 1 && pix_to_ndc(lo, S1, S2) < vmin) ++lo;
  while (hi < S1 - 1 && pix_to_ndc(hi + 1, S1, S2) <= vmax) ++hi;
  while (hi >= 0 && pix_to_ndc(hi, S1, S2) > vmax) --hi;
}

// Inclusive range [ilo, ihi] of pixel indices in [t0, t1] whose centre lies in [vmin, vmax]; empty when
// ilo > ihi.  tab[i - t0] holds the centre of pixel i (strictly decreasing in i).  A float estimate from the inverse
// of PixToNonSquareNdc is corrected against the table, so the range is exact, not conservative: the pixel
// loops visit exactly the pixels that pass the oracle's CheckPointOutsideBoundingBox.
__device__ __forceinline__ void pixel_range(float vmin, float vmax, int S1, int S2, int t0, int t1, const float* tab,
                                            int& ilo, int& ihi) {
  float range = 2.0f;
  if (S1 > S2) range = ((float)(S1 / S2)) * range;
  const float offset = range / 2.0f;
  const float inv_range = 1.0f / range;
  // centre of pixel i is c(S1-1-i) with c(j) = -offset + (range*j + offset)/S1, so i decreases as the coordinate grows
  float jhi = floorf(((vmax + offset) * (float)S1 - offset) * inv_range);
  float jlo = ceilf(((vmin + offset) * (float)S1 - offset) * inv_range);
  jhi = fminf(fmaxf(jhi, -2.0f), (float)S1 + 1.0f);
  jlo = fminf(fmaxf(jlo, -2.0f), (float)S1 + 1.0f);
  ilo = max(S1 - 1 - (int)jhi, t0);
  ihi = min(S1 - 1 - (int)jlo, t1);
  if (ilo > t1 + 1) ilo = t1 + 1;
  if (ihi < t0 - 1) ihi = t0 - 1;
  while (ilo > t0 && tab[ilo - 1 - t0] <= vmax) --ilo;
  while (ilo <= t1 && tab[ilo - t0] > vmax) ++ilo;
  while (ihi < t1 && tab[ihi + 1 - t0] >= vmin) ++ihi;
  while (ihi >= t0 && tab[ihi - t0] < vmin) --ihi;
}

// fills tab[0..W) with the NDC x of every pixel column centre and tab[W..W+H) with the NDC y of every row centre
__device__ __forceinline__ void fill_pixel_table(float* __restrict__ tab, int H, int W, int tid, int nthreads) {
  for (int i = tid; i < W; i += nthreads) tab[i] = pix_to_ndc(W - 1 - i, W, H);
  for (int i = tid; i < H; i += nthreads) tab[W + i] = pix_to_ndc(H - 1 - i, H, W);
}

// X_view = X_world R + T, normative order ((x*R0j + y*R1j) + z*R2j) + Tj.
struct Camera {
  float r[9];
  float t[3];
};
__device__ __forceinline__ Camera load_camera(const float* __restrict__ R, const float* __restrict__ T, int n) {
  Camera c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.r[i] = __ldg(R + 9 * (size_t)n + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) c.t[i] = __ldg(T + 3 * (size_t)n + i);
  return c;
}
// Written with round-to-nearest intrinsics, which the compiler never contracts into FMAs: the result is the oracle's
// IEEE sequence whatever -fmad says for the translation unit (the backward units are compiled with contraction).
__device__ __forceinline__ void world_to_view(const Camera& c, float x, float y, float z, float& px, float& py, float& pz) {
  px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, c.r[0]), __fmul_rn(y, c.r[3])), __fmul_rn(z, c.r[6])), c.t[0]);
  py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, c.r[1]), __fmul_rn(y, c.r[4])), __fmul_rn(z, c.r[7])), c.t[1]);
  pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, c.r[2]), __fmul_rn(y, c.r[5])), __fmul_rn(z, c.r[8])), c.t[2]);
}

// Consumer-side output transform (Trainer_mvt.py:41-49 Normalize, run before the CNN): the image
// is written as (x - mean_c) * (1 / std_c), optionally rounded to bfloat16 [MVR_IMAGES_BF16]; the backward reads the
// cotangent of THAT tensor and scales it by 1 / std_c.  on == false leaves the arithmetic of the default path untouched.
struct OutNorm {
  float m0, m1, m2, s0, s1, s2;
  bool on;
};
static inline OutNorm make_out_norm(const float* host_mean_std) {
  OutNorm q = {0.f, 0.f, 0.f, 1.f, 1.f, 1.f, false};
  if (host_mean_std) {
    q.m0 = host_mean_std[0]; q.m1 = host_mean_std[1]; q.m2 = host_mean_std[2];
    q.s0 = 1.0f / host_mean_std[3]; q.s1 = 1.0f / host_mean_std[4]; q.s2 = 1.0f / host_mean_std[5];
    q.on = true;
  }
  return q;
}
static inline bool out_norm_valid(const float* host_mean_std) {
  if (!host_mean_std) return true;
  for (int i = 3; i < 6; ++i)
    if (!(host_mean_std[i] > 0.f)) return false;
  return true;
}
// planar (n,3,H,W) pixel store / cotangent load; io = offset of channel 0, plane = H*W
__device__ __forceinline__ void store_rgb(void* images, bool bf16, size_t io, size_t plane, float r, float g, float b,
                                          const OutNorm& q) {
  if (q.on) { r = (r - q.m0) * q.s0; g = (g - q.m1) * q.s1; b = (b - q.m2) * q.s2; }
  if (bf16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(images);
    o[io] = __float2bfloat16_rn(r); o[io + plane] = __float2bfloat16_rn(g); o[io + 2 * plane] = __float2bfloat16_rn(b);
  } else {
    float* o = reinterpret_cast<float*>(images);
    o[io] = r; o[io + plane] = g; o[io + 2 * plane] = b;
  }
}
__device__ __forceinline__ void load_grad_rgb(const void* grad, bool bf16, size_t io, size_t plane, const OutNorm& q,
                                              float& g0, float& g1, float& g2) {
  if (bf16) {
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(grad);
    g0 = __bfloat162float(g[io]); g1 = __bfloat162float(g[io + plane]); g2 = __bfloat162float(g[io + 2 * plane]);
  } else {
    const float* g = reinterpret_cast<const float*>(grad);
    g0 = __ldg(g + io); g1 = __ldg(g + io + plane); g2 = __ldg(g + io + 2 * plane);
  }
  if (q.on) { g0 *= q.s0; g1 *= q.s1; g2 *= q.s2; }
}

// raw bits of a pixel's three cotangent channels: loaded one iteration ahead, converted only when they are used, so that the
// loads of item i + 1 are in flight while item i is computed
struct GradRaw { unsigned int c0, c1, c2; };
__device__ __forceinline__ GradRaw load_grad_raw(const void* grad, bool bf16, size_t io, size_t plane) {
  GradRaw r;
  if (bf16) {
    const unsigned short* g = reinterpret_cast<const unsigned short*>(grad);
    r.c0 = __ldg(g + io); r.c1 = __ldg(g + io + plane); r.c2 = __ldg(g + io + 2 * plane);
  } else {
    const unsigned int* g = reinterpret_cast<const unsigned int*>(grad);
    r.c0 = __ldg(g + io); r.c1 = __ldg(g + io + plane); r.c2 = __ldg(g + io + 2 * plane);
  }
  return r;
}
__device__ __forceinline__ void grad_from_raw(const GradRaw& r, bool bf16, const OutNorm& q, float& g0, float& g1, float& g2) {
  const int sh = bf16 ? 16 : 0;
  g0 = __uint_as_float(r.c0 << sh); g1 = __uint_as_float(r.c1 << sh); g2 = __uint_as_float(r.c2 << sh);
  if (q.on) { g0 *= q.s0; g1 *= q.s1; g2 *= q.s2; }
}

__device__ __forceinline__ unsigned long long make_key(float z, int idx) {
  // z >= 0 here, so the IEEE bit pattern is monotone; +0.0f canonicalises -0.0f.
  return ((unsigned long long)__float_as_uint(z + 0.0f) << 32) | (unsigned int)idx;
}

// 64-bit min on shared memory (lexicographic (z, idx) == the oracle's priority-queue order).
__device__ __forceinline__ void smem_key_min(unsigned long long* addr, unsigned long long key) {
  unsigned long long cur = *(volatile unsigned long long*)addr;
  while (key < cur) {
    const unsigned long long old = atomicCAS(addr, cur, key);
    if (old == cur) break;
    cur = old;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Warp-wide sums of 16 values per lane in 16 shuffles instead of 80: at each butterfly step a lane keeps the half of
// its values selected by the corresponding bit of its lane id and hands the other half to its partner.  On return
// lanes 2i and 2i+1 both hold the warp total of v[i] (i = 0..15); the summation order is fixed, so the result is
// run-to-run identical.
__device__ __forceinline__ float warp_sum16_transposed(float (&v)[16]) {
  const int lane = threadIdx.x & 31;
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Block-wide sum of NV values per thread (MVR_THREADS threads); result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* s_red /* [8][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) s_red[warp * NV + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
    for (int w = 0; w < MVR_THREADS / 32; ++w) s += s_red[w * NV + threadIdx.x];
    s_red[threadIdx.x] = s;
  }
  __syncthreads();
}

}  // namespace mvr
