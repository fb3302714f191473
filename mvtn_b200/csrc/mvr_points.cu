// mvr_points.cu -- point-cloud path of MVRenderer (renderer.py:116-151) for sm_100a.
//
//   forward : points_forward_kernel -- one CTA per (view, strip of image rows).  A view has only a few
//             thousand points, so there is no global binning pass: every strip CTA streams the cloud
//             once (coalesced, L2-resident across the M views and strips), projects
//             p = (X / dist) R + T, rejects on the strip's y-range, and SCATTERS each surviving point into
//             the handful of pixel centres inside its radius with a 64-bit (z, point) min on a
//             shared-memory key per pixel.  K > 1 peels layers.  The epilogue recomputes dist2 of the
//             winner, applies the norm-weighted or alpha compositor and the background, and writes whole
//             image rows (fully coalesced planar stores).
//   backward: points_backward_kernel -- per pixel recompute from idx only; compositor backward ->
//             d dist2 -> d ndc.xy -> (dR, dT, d(1/dist)) block-reduced to one partial per (view, strip),
//             summed in fixed order; optional per-point / colour gradients via atomics.
#include "mvr_common.cuh"

namespace mvr {

constexpr int PB_ROWS = 8;     // backward strip height
constexpr int PB_VALS = 13;    // dR 9, dT 3, d inv_dist 1

struct PointsParams {
  const float* points; const float* rgb;
  const float* R; const float* T; const float* inv_dist; const float* bg_rgb;
  float radius, r2_raster, r2_weight;
  int B, Np, M, H, W, K, flags, strip_rows, n_strips;
  float* images; int* idx; float* zbuf; float* dists2;
};

__device__ __forceinline__ void project_point(const float* __restrict__ pts, int pi, float s, const Camera& cam,
                                              float& px, float& py, float& pz) {
  const float x = __ldg(pts + 3 * (size_t)pi) * s, y = __ldg(pts + 3 * (size_t)pi + 1) * s, z = __ldg(pts + 3 * (size_t)pi + 2) * s;
  world_to_view(cam, x, y, z, px, py, pz);
}

__global__ void __launch_bounds__(MVR_THREADS) points_forward_kernel(const PointsParams p) {
  extern __shared__ unsigned long long s_keys[];   // cur [rows*W] (+ prev [rows*W] + acc float4 [rows*W] when K > 1)
  const int tid = threadIdx.x;
  const int n = blockIdx.x / p.n_strips, strip = blockIdx.x % p.n_strips;
  const int b = n / p.M;
  const int y0 = strip * p.strip_rows, y1 = min(y0 + p.strip_rows, p.H) - 1;   // inclusive
  const int npix = (y1 - y0 + 1) * p.W;
  const int cap = p.strip_rows * p.W;
  unsigned long long* s_cur = s_keys;
  unsigned long long* s_prev = s_keys + cap;
  float4* s_acc = (float4*)(s_keys + 2 * (size_t)cap);
  const Camera cam = load_camera(p.R, p.T, n);
  const float s = __ldg(p.inv_dist + n);
  const float* pts = p.points + 3 * (size_t)b * p.Np;
  const bool per_point_rgb = p.flags & MVR_RGB_PER_ELEMENT;
  const bool alpha_mode = p.flags & MVR_COMPOSITE_ALPHA;
  const float* feat = per_point_rgb ? p.rgb + 3 * (size_t)b * p.Np : p.rgb;
  const float bg0 = __ldg(p.bg_rgb), bg1 = __ldg(p.bg_rgb + 1), bg2 = __ldg(p.bg_rgb + 2);
  // conservative search radius for candidate pixel centres (the exact test is dist2 < r2 below)
  const float rr = p.radius * 1.0001f + 1e-7f;
  // y-extent of the strip in NDC (pixel centres), padded by the search radius
  const float strip_ymax = pix_to_ndc(p.H - 1 - y0, p.H, p.W) + rr;
  const float strip_ymin = pix_to_ndc(p.H - 1 - y1, p.H, p.W) - rr;

  if (p.K > 1)
    for (int i = tid; i < npix; i += MVR_THREADS) { s_prev[i] = 0ull; s_acc[i] = make_float4(0.f, 0.f, 0.f, alpha_mode ? 1.f : 0.f); }

  for (int k = 0; k < p.K; ++k) {
    const bool peel = k > 0;
    for (int i = tid; i < npix; i += MVR_THREADS) s_cur[i] = MVR_EMPTY_KEY;
    __syncthreads();
    for (int pi = tid; pi < p.Np; pi += MVR_THREADS) {
      float px, py, pz;
      project_point(pts, pi, s, cam, px, py, pz);
      if (!(py <= strip_ymax && py >= strip_ymin)) continue;
      if (pz < 0.f) continue;
      int jlo, jhi;
      ndc_range_to_pix(py - rr, py + rr, p.H, p.W, jlo, jhi);
      const int yl = max(p.H - 1 - jhi, y0), yh = min(p.H - 1 - jlo, y1);
      if (yl > yh) continue;
      ndc_range_to_pix(px - rr, px + rr, p.W, p.H, jlo, jhi);
      const int xl = p.W - 1 - jhi, xh = p.W - 1 - jlo;
      const unsigned long long key = make_key(pz, pi);
      for (int yy = yl; yy <= yh; ++yy) {
        const float dy = py - pix_to_ndc(p.H - 1 - yy, p.H, p.W);
        for (int xx = xl; xx <= xh; ++xx) {
          const float dx = px - pix_to_ndc(p.W - 1 - xx, p.W, p.H);
          const float d2 = dx * dx + dy * dy;
          if (!(d2 < p.r2_raster)) continue;
          const int pix = (yy - y0) * p.W + xx;
          if (peel && key <= s_prev[pix]) continue;
          smem_key_min(&s_cur[pix], key);
        }
      }
    }
    __syncthreads();
    // ---- epilogue for layer k ----
    const bool last = k == p.K - 1;
    for (int pix = tid; pix < npix; pix += MVR_THREADS) {
      const int yi = y0 + pix / p.W, xi = pix % p.W;
      const unsigned long long key = s_cur[pix];
      int pid = -1;
      float z = -1.f, d2 = -1.f;
      float4 acc = make_float4(0.f, 0.f, 0.f, alpha_mode ? 1.f : 0.f);
      if (p.K > 1) { acc = s_acc[pix]; s_prev[pix] = key; }
      if (key != MVR_EMPTY_KEY) {
        pid = (int)(unsigned int)(key & 0xffffffffull);
        float px, py, pz;
        project_point(pts, pid, s, cam, px, py, pz);
        const float dx = px - pix_to_ndc(p.W - 1 - xi, p.W, p.H);
        const float dy = py - pix_to_ndc(p.H - 1 - yi, p.H, p.W);
        d2 = dx * dx + dy * dy;
        z = __uint_as_float((unsigned int)(key >> 32));
        const float a = 1.f - d2 / p.r2_weight;
        const float f0 = __ldg(feat + (per_point_rgb ? 3 * (size_t)pid : 0)), f1 = __ldg(feat + (per_point_rgb ? 3 * (size_t)pid : 0) + 1),
                    f2 = __ldg(feat + (per_point_rgb ? 3 * (size_t)pid : 0) + 2);
        if (alpha_mode) {   // out += cum * alpha * f ; cum *= (1 - alpha)
          const float ca = acc.w * a;
          acc.x += ca * f0; acc.y += ca * f1; acc.z += ca * f2;
          acc.w = acc.w * (1.f - a);
        } else {            // numerators and the alpha sum
          acc.x += a * f0; acc.y += a * f1; acc.z += a * f2;
          acc.w += a;
        }
      }
      if (p.K > 1) s_acc[pix] = acc;
      const size_t po = (((size_t)n * p.H + yi) * p.W + xi) * p.K + k;
      p.idx[po] = pid;
      if (p.zbuf) p.zbuf[po] = z;
      if (p.dists2) p.dists2[po] = d2;
      if (last) {
        // background where the FIRST layer is empty ([upstream] _add_background_color_to_images)
        const bool fg = (p.K > 1) ? (p.idx[po - k] >= 0) : (pid >= 0);
        float o0 = bg0, o1 = bg1, o2 = bg2;
        if (fg) {
          if (alpha_mode) { o0 = acc.x; o1 = acc.y; o2 = acc.z; }
          else { const float t = fmaxf(acc.w, 1e-4f); o0 = acc.x / t; o1 = acc.y / t; o2 = acc.z / t; }
        }
        const size_t io = ((size_t)n * 3 * p.H + yi) * p.W + xi;
        const size_t plane = (size_t)p.H * p.W;
        p.images[io] = o0; p.images[io + plane] = o1; p.images[io + 2 * plane] = o2;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
struct PointsBwdParams {
  const float* points; const float* rgb;
  const float* R; const float* T; const float* inv_dist;
  float r2_weight;
  int B, Np, M, H, W, K, flags, n_strips;
  const int* idx; const float* grad_images;
  float* partials;        // (N, n_strips, 16)
  float* grad_points; float* grad_rgb;
};

__global__ void __launch_bounds__(MVR_THREADS) points_backward_kernel(const PointsBwdParams p) {
  __shared__ float s_red[8 * PB_VALS];
  __shared__ int s_any;
  const int tid = threadIdx.x;
  const int n = blockIdx.x / p.n_strips, strip = blockIdx.x % p.n_strips;
  const int b = n / p.M;
  const int y0 = strip * PB_ROWS, y1 = min(y0 + PB_ROWS, p.H) - 1;
  const int npix = (y1 - y0 + 1) * p.W;
  const bool per_point_rgb = p.flags & MVR_RGB_PER_ELEMENT;
  const bool alpha_mode = p.flags & MVR_COMPOSITE_ALPHA;
  const float* pts = p.points + 3 * (size_t)b * p.Np;
  const float* feat = per_point_rgb ? p.rgb + 3 * (size_t)b * p.Np : p.rgb;
  if (tid == 0) s_any = 0;
  __syncthreads();
  float acc[PB_VALS];
#pragma unroll
  for (int i = 0; i < PB_VALS; ++i) acc[i] = 0.f;
  bool any = false, ctx = false;
  Camera cam; float s = 0.f;
  const size_t plane = (size_t)p.H * p.W;
  const float inv_r2 = 1.f / p.r2_weight;
  for (int pix = tid; pix < npix; pix += MVR_THREADS) {
    const int yi = y0 + pix / p.W, xi = pix % p.W;
    const int* ip = p.idx + (((size_t)n * p.H + yi) * p.W + xi) * p.K;
    if (__ldg(ip) < 0) continue;   // background pixel: masked_scatter blocks the gradient
    const size_t io = ((size_t)n * 3 * p.H + yi) * p.W + xi;
    const float g0 = __ldg(p.grad_images + io), g1 = __ldg(p.grad_images + io + plane), g2 = __ldg(p.grad_images + io + 2 * plane);
    if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;
    any = true;
    if (!ctx) { cam = load_camera(p.R, p.T, n); s = __ldg(p.inv_dist + n); ctx = true; }
    const float xf = pix_to_ndc(p.W - 1 - xi, p.W, p.H), yf = pix_to_ndc(p.H - 1 - yi, p.H, p.W);
    // pass 1: compositor totals
    float t_alpha = 0.f, tf0 = 0.f, tf1 = 0.f, tf2 = 0.f;   // norm: sum a, sum a f ; alpha: out_c
    {
      float cum = 1.f;
      for (int k = 0; k < p.K; ++k) {
        const int pid = __ldg(ip + k);
        if (pid < 0) break;
        float px, py, pz; project_point(pts, pid, s, cam, px, py, pz);
        const float dx = px - xf, dy = py - yf;
        const float a = 1.f - (dx * dx + dy * dy) / p.r2_weight;   // exactly the forward's alpha: it is clamped at 1e-4 below
        const float* f = feat + (per_point_rgb ? 3 * (size_t)pid : 0);
        const float wgt = alpha_mode ? cum * a : a;
        tf0 = fmaf(wgt, __ldg(f), tf0); tf1 = fmaf(wgt, __ldg(f + 1), tf1); tf2 = fmaf(wgt, __ldg(f + 2), tf2);
        t_alpha += a; cum *= (1.f - a);
      }
    }
    const float t = fmaxf(t_alpha, 1e-4f);
    // pass 2: per-layer gradients
    float cum = 1.f, pre0 = 0.f, pre1 = 0.f, pre2 = 0.f;
    for (int k = 0; k < p.K; ++k) {
      const int pid = __ldg(ip + k);
      if (pid < 0) break;
      const float X0 = __ldg(pts + 3 * (size_t)pid), X1 = __ldg(pts + 3 * (size_t)pid + 1), X2 = __ldg(pts + 3 * (size_t)pid + 2);
      float px, py, pz; world_to_view(cam, X0 * s, X1 * s, X2 * s, px, py, pz);
      const float dx = px - xf, dy = py - yf;
      const float a = 1.f - (dx * dx + dy * dy) / p.r2_weight;   // exactly the forward's alpha: it is clamped at 1e-4 below
      const float* f = feat + (per_point_rgb ? 3 * (size_t)pid : 0);
      const float f0 = __ldg(f), f1 = __ldg(f + 1), f2 = __ldg(f + 2);
      float ga, gf;   // d/d alpha_k ; d/d f_k (per unit grad_out, same for all channels)
      if (alpha_mode) {
        // out_c = sum_j f_jc cum_j a_j ;  d/da_k = f_kc cum_k - (sum_{j>k} f_jc cum_j a_j) / (1 - a_k + 1e-9)
        const float w = cum * a;
        pre0 = fmaf(w, f0, pre0); pre1 = fmaf(w, f1, pre1); pre2 = fmaf(w, f2, pre2);
        const float inv1m = 1.f / (1.f - a + 1e-9f);
        ga = g0 * (f0 * cum - (tf0 - pre0) * inv1m) + g1 * (f1 * cum - (tf1 - pre1) * inv1m) + g2 * (f2 * cum - (tf2 - pre2) * inv1m);
        gf = w;
        cum *= (1.f - a);
      } else {
        const float it2 = 1.f / (t * t);
        ga = (g0 * (t * f0 - tf0) + g1 * (t * f1 - tf1) + g2 * (t * f2 - tf2)) * it2;
        gf = a / t;
      }
      const float gd2 = -ga * inv_r2;
      const float gpx = 2.f * gd2 * dx, gpy = 2.f * gd2 * dy;
      const float Xs0 = X0 * s, Xs1 = X1 * s, Xs2 = X2 * s;
      acc[0] = fmaf(Xs0, gpx, acc[0]); acc[1] = fmaf(Xs0, gpy, acc[1]);
      acc[3] = fmaf(Xs1, gpx, acc[3]); acc[4] = fmaf(Xs1, gpy, acc[4]);
      acc[6] = fmaf(Xs2, gpx, acc[6]); acc[7] = fmaf(Xs2, gpy, acc[7]);
      acc[9] += gpx; acc[10] += gpy;
      const float gx0 = cam.r[0] * gpx + cam.r[1] * gpy, gx1 = cam.r[3] * gpx + cam.r[4] * gpy, gx2 = cam.r[6] * gpx + cam.r[7] * gpy;
      acc[12] += gx0 * X0 + gx1 * X1 + gx2 * X2;
      if (p.grad_points) {
        float* o = p.grad_points + 3 * ((size_t)b * p.Np + pid);
        atomicAdd(o, gx0 * s); atomicAdd(o + 1, gx1 * s); atomicAdd(o + 2, gx2 * s);
      }
      if (p.grad_rgb) {
        float* o = p.grad_rgb + (per_point_rgb ? 3 * ((size_t)b * p.Np + pid) : 0);
        atomicAdd(o, g0 * gf); atomicAdd(o + 1, g1 * gf); atomicAdd(o + 2, g2 * gf);
      }
    }
  }
  if (any) s_any = 1;
  __syncthreads();
  float* out = p.partials + ((size_t)n * p.n_strips + strip) * 16;
  if (!s_any) {
    if (tid < 16) out[tid] = 0.f;
    return;
  }
  block_sum<PB_VALS>(acc, s_red);
  if (tid < 16) out[tid] = tid < PB_VALS ? s_red[tid] : 0.f;
}

__global__ void points_backward_reduce_kernel(const float* __restrict__ partials, int N, int n_strips,
                                              float* __restrict__ gR, float* __restrict__ gT, float* __restrict__ gs) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const int v = lane & 15, par = lane >> 4;
  float s = 0.f;
  for (int t = par; t < n_strips; t += 2) s += partials[((size_t)n * n_strips + t) * 16 + v];
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  if (lane < 9) gR[9 * (size_t)n + lane] = s;
  else if (lane < 12) gT[3 * (size_t)n + lane - 9] = s;
  else if (lane == 12) gs[n] = s;
}

static int choose_strip_rows(int N, int H, int W, int K, size_t* smem_bytes) {
  const size_t bpp = K == 1 ? 8 : 32;
  int rows = (int)((64 * 1024) / ((size_t)W * bpp));
  if (rows > 32) rows = 32;
  if (rows < 1) rows = 1;
  if (rows > H) rows = H;
  while (rows > 4 && (long long)N * ((H + rows - 1) / rows) < 4 * 148) rows = (rows + 1) / 2;
  *smem_bytes = (size_t)rows * W * bpp;
  return rows;
}

}  // namespace mvr

using namespace mvr;

static int check_points_common(const char* who, int B, int Np, int M, int H, int W, int K, double radius) {
  if (B < 0 || Np < 0 || M < 0) { set_error("%s: negative size", who); return -1; }
  if (H <= 0 || W <= 0 || H > 4096 || W > 4096) { set_error("%s: image size %dx%d outside [1, 4096]", who, H, W); return -2; }
  if (K < 1 || K > 64) { set_error("%s: points_per_pixel %d outside [1, 64]", who, K); return -3; }
  if (!(radius > 0.0)) { set_error("%s: radius must be positive", who); return -4; }
  if ((int64_t)B * M > 0x7fffffffLL / (H + 1)) { set_error("%s: too many views", who); return -5; }
  return 0;
}

extern "C" size_t mvr_points_workspace_bytes(int B, int M, int H, int W, int K) {
  if (B < 0 || M < 0 || H <= 0 || W <= 0) return 0;
  const size_t n_strips = (H + PB_ROWS - 1) / PB_ROWS;
  return ((size_t)B * M * n_strips * 16 * sizeof(float) + 255) & ~(size_t)255;
}

extern "C" int mvr_points_forward(const float* points, const float* rgb, int B, int Np, int M, const float* R,
                                  const float* T, const float* inv_dist, double radius, const float* bg_rgb,
                                  int H, int W, int K, int flags, float* images, int* idx, float* zbuf,
                                  float* dists2, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_points_common("mvr_points_forward", B, Np, M, H, W, K, radius);
  if (rc) return rc;
  const int64_t N = (int64_t)B * M;
  if (N == 0) return 0;
  if ((Np > 0 && !points) || !rgb || !R || !T || !inv_dist || !bg_rgb || !images || !idx) { set_error("mvr_points_forward: null pointer"); return -6; }
  PointsParams p;
  p.points = points; p.rgb = rgb; p.R = R; p.T = T; p.inv_dist = inv_dist; p.bg_rgb = bg_rgb;
  p.radius = (float)radius;
  p.r2_raster = p.radius * p.radius;            // [upstream] rasterize_points_cpu.cpp: float radius * radius
  p.r2_weight = (float)(radius * radius);       // [upstream] points/renderer.py: python-float r * r
  p.B = B; p.Np = Np; p.M = M; p.H = H; p.W = W; p.K = K; p.flags = flags;
  size_t smem;
  p.strip_rows = choose_strip_rows((int)N, H, W, K, &smem);
  p.n_strips = (H + p.strip_rows - 1) / p.strip_rows;
  p.images = images; p.idx = idx; p.zbuf = zbuf; p.dists2 = dists2;
  if (smem > 200 * 1024) { set_error("mvr_points_forward: image too wide for one shared-memory row (%zu B)", smem); return -7; }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(points_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mvr_points_forward: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  MVR_LAUNCH(points_forward_kernel, (unsigned)(N * p.n_strips), MVR_THREADS, smem, (cudaStream_t)stream, p);
  return check_launch("points_forward_kernel");
}

extern "C" int mvr_points_backward(const float* points, const float* rgb, int B, int Np, int M, const float* R,
                                   const float* T, const float* inv_dist, double radius, int H, int W, int K,
                                   int flags, const int* idx, const float* grad_images, float* gR, float* gT,
                                   float* g_inv_dist, float* grad_points, float* grad_rgb, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  int rc = check_points_common("mvr_points_backward", B, Np, M, H, W, K, radius);
  if (rc) return rc;
  const int64_t N = (int64_t)B * M;
  if (N == 0) return 0;
  if ((Np > 0 && !points) || !rgb || !R || !T || !inv_dist || !idx || !grad_images || !gR || !gT || !g_inv_dist || !workspace) {
    set_error("mvr_points_backward: null pointer"); return -6;
  }
  const size_t need = mvr_points_workspace_bytes(B, M, H, W, K);
  if (workspace_bytes < need) { set_error("mvr_points_backward: workspace too small (%zu < %zu)", workspace_bytes, need); return -7; }
  PointsBwdParams p;
  p.points = points; p.rgb = rgb; p.R = R; p.T = T; p.inv_dist = inv_dist;
  p.r2_weight = (float)(radius * radius);
  p.B = B; p.Np = Np; p.M = M; p.H = H; p.W = W; p.K = K; p.flags = flags;
  p.n_strips = (H + PB_ROWS - 1) / PB_ROWS;
  p.idx = idx; p.grad_images = grad_images; p.partials = (float*)workspace;
  p.grad_points = grad_points; p.grad_rgb = grad_rgb;
  cudaStream_t st = (cudaStream_t)stream;
  MVR_LAUNCH(points_backward_kernel, (unsigned)(N * p.n_strips), MVR_THREADS, 0, st, p);
  rc = check_launch("points_backward_kernel");
  if (rc) return rc;
  const int wpb = 8;
  MVR_LAUNCH(points_backward_reduce_kernel, (unsigned)((N + wpb - 1) / wpb), wpb * 32, 0, st, (const float*)workspace, (int)N, p.n_strips, gR, gT, g_inv_dist);
  return check_launch("points_backward_reduce_kernel");
}
