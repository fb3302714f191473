// mvr_mesh.cu -- mesh path of MVRenderer (renderer.py:65-114) for sm_100a.
//
//   prepare : pack verts/normals/colours to float4 and faces to int4; area-weighted vertex normals
//             ([upstream] Meshes._compute_vertex_normals) once per OBJECT, not per view.
//   project : mesh_project_kernel -- every vertex of every view is projected exactly ONCE (world -> view -> NDC,
//             IEEE order) into a float4 (x_ndc, y_ndc, z_view) plane that stays L2-resident (30 MB at C2); the
//             scatter, shade and backward kernels gather it instead of re-projecting three vertices per
//             face / pixel.  The same launch writes the table of exact pixel-centre NDC coordinates.
//   scatter : mesh_scatter_kernel -- each face of each view is set up exactly ONCE (project, cull, exact pixel
//             bbox) by the CTA that owns its 1024-face chunk.  No bins: the work is flattened inside the CTA
//             through shared-memory queues so that every phase runs on full warps --
//               A  setup    thread per face   -> record + runs of <= 32 bbox pixels ("sub-items")
//               B  filter   thread per run    -> edge-function sign test  -> per-warp candidate queues
//               C  resolve  thread per candidate: exact barycentrics / depth -> 64-bit (z, face) RED.MIN on the
//                           pixel's key in a global (L2-resident) key plane
//             so a 3-pixel sliver and a large triangle cost their threads the same, and the IEEE divisions run
//             only on dense warps of pixels that are inside their face.  Triangles covering > 1024 pixels are
//             walked by the whole CTA.  K > 1 peels layers (pass k keeps keys > layer k-1).
//   shade   : mesh_shade_kernel -- one thread per pixel: winning key -> face, exact barycentrics recomputed with the
//             same operation sequence, Phong shading + hard background blend, planar (n,3,H,W) image and
//             pix_to_face (+ optional zbuf / bary / dists) written with fully coalesced stores.
//   backward: mesh_backward_kernel -- per pixel recompute (no fragment traffic), chain
//             d image -> Phong -> barycentrics -> NDC verts -> view verts -> (dR, dT, dC), block-reduced to
//             one partial per CTA and summed in fixed order (deterministic, no float atomics).
#include <cstdlib>

#include "mvr_common.cuh"

namespace mvr {

constexpr int FACES_PER_CTA = 1024;      // 4 rounds of 256 faces
constexpr int BIG_FACE_PIX = 1024;       // bbox pixels above which the whole CTA walks a face
constexpr int REC_WORDS = 12;            // x0 y0 z0 x1 y1 z1 x2 y2 z2 fid rect_xy rect_wh
constexpr int ITEM_CAP = 2048;           // sub-items per round (typically 256 faces x 1-3)
constexpr int WCAP = 320;                // candidates per warp queue
constexpr int NWARPS = MVR_THREADS / 32;
constexpr int BWD_PIX_PER_THREAD = 4;
constexpr int BWD_VALS = 15;             // dR 9, dT 3, dC 3

struct GeomLayout {
  size_t verts4, normals4, rgb4, faces4, nacc, total;
};
static GeomLayout geom_layout(int64_t tv, int64_t tf) {
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  GeomLayout g;
  size_t o = 0;
  g.verts4 = o; o = al(o + (size_t)tv * 16);
  g.normals4 = o; o = al(o + (size_t)tv * 16);
  g.rgb4 = o; o = al(o + (size_t)tv * 16);
  g.faces4 = o; o = al(o + (size_t)tf * 16);
  g.nacc = o; o = al(o + (size_t)tv * 24);
  g.total = o;
  return g;
}

struct WsLayout {
  size_t pv, tab, keys, prev, partials, total;
  int bwd_ctas_per_view;
};
// [pv | tab] are shared by the forward and the backward call (each re-projects: the workspace is scratch and may
// have been reused in between); the forward adds the key planes, the backward its per-CTA partial sums.
static WsLayout ws_layout(int B, int M, int H, int W, int K, int64_t total_verts) {
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  WsLayout w;
  const size_t N = (size_t)B * M, HW = (size_t)H * W;
  size_t o = 0;
  w.pv = o; o = al(o + (size_t)M * (size_t)total_verts * 16);
  w.tab = o; o = al(o + ((size_t)W + H) * sizeof(float));
  const size_t common = o;
  w.keys = o; o = al(o + N * HW * 8);
  w.prev = o; if (K > 1) o = al(o + N * HW * 8);
  w.bwd_ctas_per_view = ((W + 31) / 32) * ((H + 31) / 32);      // 32x32-pixel tiles
  w.partials = common;
  const size_t bwd = al(common + N * w.bwd_ctas_per_view * NWARPS * 16 * sizeof(float));
  w.total = o > bwd ? o : bwd;
  return w;
}

// ------------------------------------------------------------------------------------------------
// prepare
// ------------------------------------------------------------------------------------------------
__global__ void geom_pack_verts_kernel(const float* __restrict__ verts, const float* __restrict__ rgb, int64_t tv,
                                       float4* __restrict__ verts4, float4* __restrict__ rgb4, double* __restrict__ nacc) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= tv) return;
  verts4[v] = make_float4(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2], 0.f);
  if (rgb) rgb4[v] = make_float4(rgb[3 * v], rgb[3 * v + 1], rgb[3 * v + 2], 0.f);
  nacc[3 * v] = 0.0; nacc[3 * v + 1] = 0.0; nacc[3 * v + 2] = 0.0;
}

template <typename IdxT>
__global__ void geom_pack_faces_kernel(const IdxT* __restrict__ faces, const int* __restrict__ vert_off,
                                       const int* __restrict__ face_off, const float4* __restrict__ verts4,
                                       int4* __restrict__ faces4, double* __restrict__ nacc) {
  const int b = blockIdx.y;
  const int f0 = face_off[b], F = face_off[b + 1] - f0;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int voff = vert_off[b], V = vert_off[b + 1] - voff;
  const IdxT* fp = faces + 3 * (size_t)(f0 + f);
  int i0 = (int)fp[0], i1 = (int)fp[1], i2 = (int)fp[2];
  // out-of-range ids would fault later: clamp (garbage in, bounded garbage out)
  i0 = min(max(i0, 0), V - 1); i1 = min(max(i1, 0), V - 1); i2 = min(max(i2, 0), V - 1);
  faces4[f0 + f] = make_int4(i0, i1, i2, 0);
  const float4 v0 = verts4[voff + i0], v1 = verts4[voff + i1], v2 = verts4[voff + i2];
  // fn = cross(v2 - v1, v0 - v1), area-weighted
  const float ax = v2.x - v1.x, ay = v2.y - v1.y, az = v2.z - v1.z;
  const float bx = v0.x - v1.x, by = v0.y - v1.y, bz = v0.z - v1.z;
  const double nx = (double)(ay * bz - az * by), ny = (double)(az * bx - ax * bz), nz = (double)(ax * by - ay * bx);
  // double accumulation: order-independent to ~1e-16, i.e. run-to-run identical after rounding to fp32
  atomicAdd(nacc + 3 * (size_t)(voff + i0), nx); atomicAdd(nacc + 3 * (size_t)(voff + i0) + 1, ny); atomicAdd(nacc + 3 * (size_t)(voff + i0) + 2, nz);
  atomicAdd(nacc + 3 * (size_t)(voff + i1), nx); atomicAdd(nacc + 3 * (size_t)(voff + i1) + 1, ny); atomicAdd(nacc + 3 * (size_t)(voff + i1) + 2, nz);
  atomicAdd(nacc + 3 * (size_t)(voff + i2), nx); atomicAdd(nacc + 3 * (size_t)(voff + i2) + 1, ny); atomicAdd(nacc + 3 * (size_t)(voff + i2) + 2, nz);
}

__global__ void geom_finish_normals_kernel(const double* __restrict__ nacc, int64_t tv, float4* __restrict__ normals4) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= tv) return;
  const float x = (float)nacc[3 * v], y = (float)nacc[3 * v + 1], z = (float)nacc[3 * v + 2];
  const float n = sqrtf((x * x + y * y) + z * z);
  const float d = n > 1e-6f ? n : 1e-6f;
  normals4[v] = make_float4(x / d, y / d, z / d, 0.f);
}

__global__ void geom_get_normals_kernel(const float4* __restrict__ normals4, int64_t tv, float* __restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= tv) return;
  const float4 n = normals4[v];
  out[3 * v] = n.x; out[3 * v + 1] = n.y; out[3 * v + 2] = n.z;
}

// ------------------------------------------------------------------------------------------------
// shared device code: projection, face setup and the per-(face, pixel) test
// ------------------------------------------------------------------------------------------------
struct MeshParams {
  const float4* verts4; const float4* normals4; const float4* rgb4; const int4* faces4;
  const int* vert_off; const int* face_off;
  const float* R; const float* T; const float* Cc; const float* light; int light_stride;
  const float* obj_rgb; const float* bg_rgb;
  float k00, k11, z_clip;
  int B, M, H, W, K, flags;
  int chunks_per_view, layer, item_cap, wcap;
  float4* pv;            // (x_ndc, y_ndc, z_view, 0) of vertex v of view (b, m) at M*vert_off[b] + m*V_b + v
  float* tab;            // pixel-centre NDC coordinates: xf[W] then yf[H]
  unsigned long long* keys; unsigned long long* prev;
  float* images; int* pix_to_face; float* zbuf; float* bary; float* dists;
  long long* counters;
};

struct Face {
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
};

__device__ __forceinline__ void project_vertex(const Camera& cam, const float4 v, float k00, float k11,
                                               float& xn, float& yn, float& zv) {
  float px, py, pz;
  world_to_view(cam, v.x, v.y, v.z, px, py, pz);
  xn = (px * k00) / pz;
  yn = (py * k11) / pz;
  zv = pz;
}

// grid: x = 256-vertex chunks of the largest object, y = view m, z = object b.  Block (0,0,0) also fills the
// pixel-centre table ([upstream] PixToNonSquareNdc evaluated once per row / column instead of once per pixel).
__global__ void __launch_bounds__(MVR_THREADS) mesh_project_kernel(const float4* __restrict__ verts4,
                                                                    const int* __restrict__ vert_off,
                                                                    const float* __restrict__ R, const float* __restrict__ T,
                                                                    int M, int H, int W, float k00, float k11,
                                                                    float4* __restrict__ pv, float* __restrict__ tab) {
  const int b = blockIdx.z, m = blockIdx.y, n = b * M + m;
  if (blockIdx.x == 0 && m == 0 && b == 0) {
    fill_pixel_table(tab, H, W, threadIdx.x, MVR_THREADS);
  }
  const int voff = vert_off[b], V = vert_off[b + 1] - voff;
  const int v = blockIdx.x * MVR_THREADS + threadIdx.x;
  if (v >= V) return;
  const Camera cam = load_camera(R, T, n);
  float xn, yn, zv;
  project_vertex(cam, __ldg(verts4 + voff + v), k00, k11, xn, yn, zv);
  pv[(size_t)M * voff + (size_t)m * V + v] = make_float4(xn, yn, zv, 0.f);
}

__device__ __forceinline__ Face gather_face(const float4* __restrict__ pvn, const int4 fi) {
  const float4 a = __ldg(pvn + fi.x), b = __ldg(pvn + fi.y), c = __ldg(pvn + fi.z);
  Face f;
  f.x0 = a.x; f.y0 = a.y; f.z0 = a.z;
  f.x1 = b.x; f.y1 = b.y; f.z1 = b.z;
  f.x2 = c.x; f.y2 = c.y; f.z2 = c.z;
  return f;
}

// Face-level rejection ([upstream] clip.py near cull, CheckPointOutsideBoundingBox z_invalid,
// RasterizeMeshesNaiveCpu zero-area / back-face tests) and the exact pixel bbox (inclusive ranges).
__device__ __forceinline__ bool face_pixel_bbox(const Face& f, const MeshParams& p, const float* s_xf, const float* s_yf,
                                                int& xi_lo, int& xi_hi, int& yi_lo, int& yi_hi) {
  if (p.z_clip >= 0.f && f.z0 < p.z_clip && f.z1 < p.z_clip && f.z2 < p.z_clip) return false;
  const float zmin = fminf(fminf(f.z0, f.z1), f.z2);
  if (zmin < MVR_K_EPS) return false;
  const float face_area = (f.x0 - f.x1) * (f.y2 - f.y1) - (f.y0 - f.y1) * (f.x2 - f.x1);
  if ((p.flags & MVR_CULL_BACKFACES) && face_area < 0.f) return false;
  if (face_area <= MVR_K_EPS && face_area >= -1.0f * MVR_K_EPS) return false;
  const float xmin = fminf(fminf(f.x0, f.x1), f.x2), xmax = fmaxf(fmaxf(f.x0, f.x1), f.x2);
  const float ymin = fminf(fminf(f.y0, f.y1), f.y2), ymax = fmaxf(fmaxf(f.y0, f.y1), f.y2);
  pixel_range(xmin, xmax, p.W, p.H, 0, p.W - 1, s_xf, xi_lo, xi_hi);
  if (xi_lo > xi_hi) return false;
  pixel_range(ymin, ymax, p.H, p.W, 0, p.H - 1, s_yf, yi_lo, yi_hi);
  return yi_lo <= yi_hi;
}

struct FaceEdges {
  float A0, B0, A1, B1, A2, B2, area_p;
};
__device__ __forceinline__ FaceEdges face_edges(const Face& f) {
  FaceEdges e;
  e.A0 = f.y2 - f.y1; e.B0 = f.x2 - f.x1;   // E(p, v1, v2)
  e.A1 = f.y0 - f.y2; e.B1 = f.x0 - f.x2;   // E(p, v2, v0)
  e.A2 = f.y1 - f.y0; e.B2 = f.x1 - f.x0;   // E(p, v0, v1)
  e.area_p = ((f.x2 - f.x0) * e.A2 - (f.y2 - f.y0) * e.B2) + MVR_K_EPS;  // E(v2, v0, v1) + kEpsilon
  return e;
}

// [upstream] BarycentricCoordinatesForward (+ BarycentricPerspectiveCorrectionForward), pz, inside.
// w = plain barycentrics, b = (corrected) barycentrics.  A cheap sign filter comes first: a pixel can
// only be inside if every edge function has the sign of the area (DESIGN.md "Parity" proves the
// filter never rejects a pixel the oracle accepts).
__device__ __forceinline__ bool raster_test(const Face& f, const FaceEdges& e, bool persp, float xf, float yf,
                                            float w[3], float b[3], float& pz) {
  const float e0 = (xf - f.x1) * e.A0 - (yf - f.y1) * e.B0;
  const float e1 = (xf - f.x2) * e.A1 - (yf - f.y2) * e.B1;
  const float e2 = (xf - f.x0) * e.A2 - (yf - f.y0) * e.B2;
  if (e.area_p > 0.f) { if (!(e0 > 0.f && e1 > 0.f && e2 > 0.f)) return false; }
  else { if (!(e0 < 0.f && e1 < 0.f && e2 < 0.f)) return false; }
  w[0] = e0 / e.area_p; w[1] = e1 / e.area_p; w[2] = e2 / e.area_p;
  if (persp) {
    const float t0 = w[0] * f.z1 * f.z2, t1 = w[1] * f.z0 * f.z2, t2 = w[2] * f.z0 * f.z1;
    const float denom = fmaxf(t0 + t1 + t2, MVR_K_EPS);
    b[0] = t0 / denom; b[1] = t1 / denom; b[2] = t2 / denom;
  } else {
    b[0] = w[0]; b[1] = w[1]; b[2] = w[2];
  }
  pz = b[0] * f.z0 + b[1] * f.z1 + b[2] * f.z2;
  if (pz < 0.f) return false;
  return b[0] > 0.0f && b[1] > 0.0f && b[2] > 0.0f;
}

__device__ __forceinline__ float point_line_dist2(float px, float py, float ax, float ay, float bx, float by) {
  const float dx = bx - ax, dy = by - ay;
  const float l2 = dx * dx + dy * dy;
  if (l2 <= MVR_K_EPS) return (px - bx) * (px - bx) + (py - by) * (py - by);
  const float t = (dx * (px - ax) + dy * (py - ay)) / l2;
  const float tt = fminf(fmaxf(t, 0.00f), 1.00f);
  const float qx = ax + tt * dx, qy = ay + tt * dy;
  return (px - qx) * (px - qx) + (py - qy) * (py - qy);
}

// ------------------------------------------------------------------------------------------------
// scatter pass
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_addr_pinned(const void* ptr) {
  unsigned int a;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(ptr));
  return a;
}
__device__ __forceinline__ unsigned int lanemask_lt() {
  unsigned int m;
  asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ float lds_f32(unsigned int a) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(unsigned int a, unsigned int v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// exact test of one (face, pixel) candidate and the keyed min on the global key plane
__device__ __forceinline__ void resolve_pixel(const Face& fc, const FaceEdges& fe, int fid, unsigned int zmin_bits,
                                              bool persp, float xf, float yf, unsigned long long* key_ptr,
                                              const unsigned long long* prev_ptr) {
  const unsigned long long cur = __ldcg(key_ptr);
  // early depth reject: pz is a convex combination of the vertex depths up to a few ulp (perspective-corrected
  // barycentrics sum to 1 unless their 1e-8 denominator clamp acts, which needs z ~ 1e-4; plain barycentrics sum
  // to area/(area+1e-8), so zmin_bits is 0 for them), hence a face whose nearest vertex is clearly behind the
  // pixel's current winner cannot produce a smaller key.  A stale `cur` only makes the test less effective.
  if (zmin_bits > (unsigned int)(cur >> 32)) return;
  float w[3], b[3], pz;
  if (!raster_test(fc, fe, persp, xf, yf, w, b, pz)) return;
  const unsigned long long key = make_key(pz, fid);
  if (key >= cur) return;
  if (prev_ptr && key <= __ldcg(prev_ptr)) return;
  atomicMin(key_ptr, key);      // result unused: RED.MIN.64 resolved in L2
}

template <int MINB>
__global__ void __launch_bounds__(MVR_THREADS, MINB) mesh_scatter_kernel(const MeshParams p) {
  __shared__ float s_rec[REC_WORDS][MVR_THREADS];     // SoA face records of the current round
  __shared__ int s_items[ITEM_CAP];                    // slot | start << 8 | count << 18
  __shared__ int s_cand[NWARPS][WCAP];                 // slot | x << 8 | y << 20
  __shared__ int s_big[MVR_THREADS];
  __shared__ int s_cnt[2];                             // [0] items, [1] big faces
  __shared__ int s_wcnt[NWARPS];
  extern __shared__ float s_tab[];                     // pixel centres: xf[W], yf[H]
  const float* s_xf = s_tab;
  const float* s_yf = s_tab + p.W;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x / p.chunks_per_view, chunk = blockIdx.x % p.chunks_per_view;
  const int b = n / p.M, m = n - b * p.M;
  const int f0 = p.face_off[b], F = p.face_off[b + 1] - f0;
  const int fbeg = chunk * FACES_PER_CTA, fend = min(F, fbeg + FACES_PER_CTA);
  if (fbeg >= fend) return;
  const int voff = p.vert_off[b], V = p.vert_off[b + 1] - voff;
  const float4* pvn = p.pv + (size_t)p.M * voff + (size_t)m * V;     // this view's projected vertices
  const bool persp = p.flags & MVR_PERSPECTIVE_CORRECT;
  unsigned long long* keys = p.keys + (size_t)n * p.H * p.W;
  const unsigned long long* prev = p.layer > 0 ? p.prev + (size_t)n * p.H * p.W : nullptr;

  for (int i = tid; i < p.W + p.H; i += MVR_THREADS) s_tab[i] = __ldg(p.tab + i);
  if (tid < 2) s_cnt[tid] = 0;
  __syncthreads();
  // shared-memory byte addresses and the lane mask, pinned in registers (volatile asm is never rematerialised: the
  // compiler otherwise rebuilds them from SR_TID / SR_CgaCtaId inside the phase-B inner loop)
  const unsigned int tab_a = smem_addr_pinned(s_tab);
  const unsigned int my_cand_a = smem_addr_pinned(&s_cand[warp][0]);
  const unsigned int lt_mask = lanemask_lt();

  int n_straddle = 0, n_big = 0;
  for (int rbeg = fbeg; rbeg < fend; rbeg += MVR_THREADS) {
    // ---------------- phase A: setup, one thread per face ----------------
    const int fid = rbeg + tid;
    if (fid < fend) {
      const Face fc = gather_face(pvn, __ldg(p.faces4 + f0 + fid));
      if (p.z_clip >= 0.f && p.layer == 0) {   // every face crossing z_clip is counted, visible or not (as the oracle does)
        const int nb = (fc.z0 < p.z_clip) + (fc.z1 < p.z_clip) + (fc.z2 < p.z_clip);
        n_straddle += (nb == 1 || nb == 2);
      }
      int xl, xh, yl, yh;
      if (face_pixel_bbox(fc, p, s_xf, s_yf, xl, xh, yl, yh)) {
        const int bw = xh - xl + 1, bh = yh - yl + 1, npx = bw * bh;
        s_rec[0][tid] = fc.x0; s_rec[1][tid] = fc.y0; s_rec[2][tid] = fc.z0;
        s_rec[3][tid] = fc.x1; s_rec[4][tid] = fc.y1; s_rec[5][tid] = fc.z1;
        s_rec[6][tid] = fc.x2; s_rec[7][tid] = fc.y2; s_rec[8][tid] = fc.z2;
        s_rec[9][tid] = __int_as_float(fid);
        s_rec[10][tid] = __int_as_float(xl | (yl << 16));
        s_rec[11][tid] = __int_as_float(bw | (bh << 16));
        bool queued = false;
        if (npx <= BIG_FACE_PIX) {
          // runs of G pixels: 8 for ordinary faces, up to 32 for large ones (<= 32 runs per face)
          const int G = max(8, (npx + 31) >> 5);
          const int nsub = (npx + G - 1) / G;
          const int at = atomicAdd(&s_cnt[0], nsub);
          if (at + nsub <= p.item_cap) {
            for (int q = 0; q < nsub; ++q) s_items[at + q] = tid | ((q * G) << 8) | (min(G, npx - q * G) << 18);
            queued = true;
          } else {
            for (int q = at; q < p.item_cap; ++q) s_items[q] = 0;      // a straddling reservation leaves no garbage
          }
        }
        if (!queued) s_big[atomicAdd(&s_cnt[1], 1)] = tid;               // walked by the whole CTA below
      }
    }
    __syncthreads();
    // ---------------- phase B: sign filter over bbox pixels, per-warp candidate queues ----------------
    const int n_items = min(s_cnt[0], p.item_cap);
    const int n_bigf = s_cnt[1];
    int wcnt = 0;                                // warp-uniform: candidates queued by this warp
    for (int j0 = warp * 32; j0 < n_items; j0 += MVR_THREADS) {      // warp-uniform trip count
      const int j = j0 + lane;
      int slot = 0, count = 0;
      unsigned int xl_a = tab_a, xend_a = tab_a + 4u, xa = tab_a, ya = tab_a;     // shared-memory BYTE addresses
      float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f, cx = 0.f, cy = 0.f;
      if (j < n_items) {
        const int it = s_items[j];
        slot = it & 255; count = it >> 18;
        const int start = (it >> 8) & 1023;
        ax = s_rec[0][slot]; ay = s_rec[1][slot];
        bx = s_rec[3][slot]; by = s_rec[4][slot];
        cx = s_rec[6][slot]; cy = s_rec[7][slot];
        const int rxy = __float_as_int(s_rec[10][slot]);
        const int xl = rxy & 0xffff;
        const int bw = __float_as_int(s_rec[11][slot]) & 0xffff;
        const int row = (int)__fdividef((float)start + 0.5f, (float)bw);     // small integers: exact
        xl_a = tab_a + 4u * (unsigned)xl;
        xend_a = xl_a + 4u * (unsigned)bw;
        xa = xl_a + 4u * (unsigned)(start - row * bw);                       // address of xf of the run's first pixel
        ya = tab_a + 4u * (unsigned)(p.W + (rxy >> 16) + row);               // address of its yf
      }
      // edge coefficients with the sign of the area folded in (negation is exact and commutes with rounding), so
      // the filter below is "all three > 0" for either winding: bit-for-bit the sign test of raster_test
      float A0 = cy - by, B0 = cx - bx, A1 = ay - cy, B1 = ax - cx, A2 = by - ay, B2 = bx - ax;
      const float area_p = ((cx - ax) * A2 - (cy - ay) * B2) + MVR_K_EPS;
      if (!(area_p > 0.f)) { A0 = -A0; B0 = -B0; A1 = -A1; B1 = -B1; A2 = -A2; B2 = -B2; }
      // candidate word = slot | x << 8 | y << 20 with x = (xa - tab_a) / 4, y = (ya - tab_a) / 4 - W: constants folded
      const unsigned int cand_m = (unsigned)slot - (tab_a << 6) - ((tab_a + 4u * (unsigned)p.W) << 18);
      const int maxc = __reduce_max_sync(0xffffffffu, count);
      for (int c = 0; c < maxc; ++c) {
        const unsigned int cxa = xa, cya = ya;
        bool pass = false;
        if (c < count) {
          const float xf = lds_f32(xa), yf = lds_f32(ya);
          const float e0 = (xf - bx) * A0 - (yf - by) * B0;
          const float e1 = (xf - cx) * A1 - (yf - cy) * B1;
          const float e2 = (xf - ax) * A2 - (yf - ay) * B2;
          pass = e0 > 0.f && e1 > 0.f && e2 > 0.f;
          xa += 4u;
          if (xa == xend_a) { xa = xl_a; ya += 4u; }
        }
        const unsigned int mk = __ballot_sync(0xffffffffu, pass);
        if (mk == 0u) continue;
        if (pass) {
          const int at = wcnt + __popc(mk & lt_mask);
          if (at < p.wcap) {
            sts_u32(my_cand_a + 4u * (unsigned)at, cand_m + (cxa << 6) + (cya << 18));
          } else {                                                  // queue full: resolve in place
            const int xx = (int)((cxa - tab_a) >> 2), yy = (int)((cya - tab_a) >> 2) - p.W;
            Face fc;
            fc.x0 = ax; fc.y0 = ay; fc.z0 = s_rec[2][slot]; fc.x1 = bx; fc.y1 = by; fc.z1 = s_rec[5][slot];
            fc.x2 = cx; fc.y2 = cy; fc.z2 = s_rec[8][slot];
            resolve_pixel(fc, face_edges(fc), __float_as_int(s_rec[9][slot]), 0u, persp, s_xf[xx], s_yf[yy],
                          keys + (size_t)yy * p.W + xx, prev ? prev + (size_t)yy * p.W + xx : nullptr);
          }
        }
        wcnt += __popc(mk);
      }
    }
    if (lane == 0) s_wcnt[warp] = min(wcnt, p.wcap);
    __syncthreads();
    // ---------------- phase C: exact resolve, candidates of all warps spread over all threads ----------------
    if (tid < 2) s_cnt[tid] = 0;          // every thread read both counts before the barrier above
    {
      int pre[NWARPS + 1];
      pre[0] = 0;
#pragma unroll
      for (int wi = 0; wi < NWARPS; ++wi) pre[wi + 1] = pre[wi] + s_wcnt[wi];
      for (int j = tid; j < pre[NWARPS]; j += MVR_THREADS) {
        int wi = 0;
#pragma unroll
        for (int q = 1; q < NWARPS; ++q) wi += (j >= pre[q]);
        const int cd = s_cand[wi][j - pre[wi]];
        const int slot = cd & 255, xx = (cd >> 8) & 4095, yy = (cd >> 20) & 4095;
        Face fc;
        fc.x0 = s_rec[0][slot]; fc.y0 = s_rec[1][slot]; fc.z0 = s_rec[2][slot];
        fc.x1 = s_rec[3][slot]; fc.y1 = s_rec[4][slot]; fc.z1 = s_rec[5][slot];
        fc.x2 = s_rec[6][slot]; fc.y2 = s_rec[7][slot]; fc.z2 = s_rec[8][slot];
        const float zmin = fminf(fminf(fc.z0, fc.z1), fc.z2);
        const unsigned int zmin_bits = (persp && zmin > 1e-3f) ? __float_as_uint(zmin * 0.999999f) : 0u;
        resolve_pixel(fc, face_edges(fc), __float_as_int(s_rec[9][slot]), zmin_bits, persp, s_xf[xx], s_yf[yy],
                      keys + (size_t)yy * p.W + xx, prev ? prev + (size_t)yy * p.W + xx : nullptr);
      }
    }
    // ---------------- large faces: the whole CTA walks the bbox ----------------
    for (int q = 0; q < n_bigf; ++q) {
      const int slot = s_big[q];
      Face fc;
      fc.x0 = s_rec[0][slot]; fc.y0 = s_rec[1][slot]; fc.z0 = s_rec[2][slot];
      fc.x1 = s_rec[3][slot]; fc.y1 = s_rec[4][slot]; fc.z1 = s_rec[5][slot];
      fc.x2 = s_rec[6][slot]; fc.y2 = s_rec[7][slot]; fc.z2 = s_rec[8][slot];
      const FaceEdges fe = face_edges(fc);
      const int bfid = __float_as_int(s_rec[9][slot]);
      const int rxy = __float_as_int(s_rec[10][slot]), rwh = __float_as_int(s_rec[11][slot]);
      const int xl = rxy & 0xffff, yl = rxy >> 16, bw = rwh & 0xffff, bh = rwh >> 16;
      const float zmin = fminf(fminf(fc.z0, fc.z1), fc.z2);
      const unsigned int zmin_bits = (persp && zmin > 1e-3f) ? __float_as_uint(zmin * 0.999999f) : 0u;
      for (int y = warp; y < bh; y += NWARPS)
        for (int x = lane; x < bw; x += 32) {
          const int xx = xl + x, yy = yl + y;
          resolve_pixel(fc, fe, bfid, zmin_bits, persp, s_xf[xx], s_yf[yy], keys + (size_t)yy * p.W + xx,
                        prev ? prev + (size_t)yy * p.W + xx : nullptr);
        }
    }
    n_big += (tid == 0) ? n_bigf : 0;
    __syncthreads();
  }
  if (p.counters) {
    if (n_straddle) atomicAdd((unsigned long long*)(p.counters + MVR_CNT_STRADDLE), (unsigned long long)n_straddle);
    if (n_big) atomicAdd((unsigned long long*)(p.counters + MVR_CNT_BIG_FACES), (unsigned long long)n_big);
  }
}

// ------------------------------------------------------------------------------------------------
// shading ([upstream] shading.py phong_shading, lighting.py diffuse/specular, blending.py hard_rgb_blend)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 interp(const float b[3], const float4 a0, const float4 a1, const float4 a2) {
  return make_float3((b[0] * a0.x + b[1] * a1.x) + b[2] * a2.x, (b[0] * a0.y + b[1] * a1.y) + b[2] * a2.y,
                     (b[0] * a0.z + b[1] * a1.z) + b[2] * a2.z);
}
// 1 / max(|v|, eps) for F.normalize(v, eps).  Shading is tolerance-compared (1e-5 on images), so the reciprocal
// square root comes from the SFU (<= 2 ulp) instead of an IEEE sqrt followed by an IEEE division.
__device__ __forceinline__ float inv_norm_clamped(float x, float y, float z, float eps) {
  const float n2 = fmaf(x, x, fmaf(y, y, z * z));
  return n2 > eps * eps ? rsqrtf(n2) : __frcp_rn(eps);
}
__device__ __forceinline__ float pow64(float a) {
  a = a * a; a = a * a; a = a * a; a = a * a; a = a * a; a = a * a;
  return a;
}

struct ShadeCtx {
  float lx, ly, lz;   // normalised light direction
  float cx, cy, cz;   // camera centre
};

__device__ __forceinline__ ShadeCtx load_shade_ctx(const float* __restrict__ light, int light_stride,
                                                   const float* __restrict__ Cc, int n) {
  ShadeCtx sc;
  const float* Lp = light + (size_t)light_stride * n;
  const float lx = __ldg(Lp), ly = __ldg(Lp + 1), lz = __ldg(Lp + 2);
  const float il = inv_norm_clamped(lx, ly, lz, 1e-6f);
  sc.lx = lx * il; sc.ly = ly * il; sc.lz = lz * il;
  sc.cx = __ldg(Cc + 3 * (size_t)n); sc.cy = __ldg(Cc + 3 * (size_t)n + 1); sc.cz = __ldg(Cc + 3 * (size_t)n + 2);
  return sc;
}

__device__ __forceinline__ void phong_pixel(const float b[3], const float4 X0, const float4 X1, const float4 X2,
                                            const float4 N0, const float4 N1, const float4 N2, const float4 c0,
                                            const float4 c1, const float4 c2, const ShadeCtx& s, float out[3]) {
  const float3 P = interp(b, X0, X1, X2);
  const float3 Nn = interp(b, N0, N1, N2);
  const float3 tex = interp(b, c0, c1, c2);
  const float in = inv_norm_clamped(Nn.x, Nn.y, Nn.z, 1e-6f);
  const float nx = Nn.x * in, ny = Nn.y * in, nz = Nn.z * in;
  const float cosang = fmaf(nx, s.lx, fmaf(ny, s.ly, nz * s.lz));
  const float diff = fmaxf(cosang, 0.f);
  const float vx = s.cx - P.x, vy = s.cy - P.y, vz = s.cz - P.z;
  const float iv = inv_norm_clamped(vx, vy, vz, 1e-6f);
  const float rx = fmaf(2.f * cosang, nx, -s.lx), ry = fmaf(2.f * cosang, ny, -s.ly), rz = fmaf(2.f * cosang, nz, -s.lz);
  const float dt = fmaf(vx * iv, rx, fmaf(vy * iv, ry, (vz * iv) * rz));
  const float alpha = (dt > 0.f && cosang > 0.f) ? dt : 0.f;
  const float spec = MVR_SPECULAR * pow64(alpha);
  const float kd = fmaf(MVR_DIFFUSE, diff, MVR_AMBIENT);
  out[0] = fmaf(kd, tex.x, spec); out[1] = fmaf(kd, tex.y, spec); out[2] = fmaf(kd, tex.z, spec);
}

// ------------------------------------------------------------------------------------------------
// shade pass: one thread per pixel
// ------------------------------------------------------------------------------------------------
// tile index -> (row, column) of tiles without an integer division (small integers: the float quotient is exact)
__device__ __forceinline__ void tile_rc(int t, int tiles_x, int& ty, int& tx) {
  ty = (int)__fdividef((float)t + 0.5f, (float)tiles_x);
  tx = t - ty * tiles_x;
}

// Barycentrics of a pixel KNOWN to be inside its face, for shading only (images are compared at 1e-5): the edge
// functions come from the same projected vertices as the scatter pass, so they are bit-identical to the rasterizer's;
// only the six IEEE divisions are replaced by two SFU reciprocals (a few ulp on b, ~1e-7 on the colour).  The exact
// sequence (raster_test) is used whenever the caller asks for the barycentrics themselves.
__device__ __forceinline__ void shading_barycentrics(const Face& f, const FaceEdges& e, bool persp, float xf, float yf,
                                                     float b[3]) {
  const float e0 = (xf - f.x1) * e.A0 - (yf - f.y1) * e.B0;
  const float e1 = (xf - f.x2) * e.A1 - (yf - f.y2) * e.B1;
  const float e2 = (xf - f.x0) * e.A2 - (yf - f.y0) * e.B2;
  const float ia = __fdividef(1.0f, e.area_p);
  b[0] = e0 * ia; b[1] = e1 * ia; b[2] = e2 * ia;
  if (persp) {
    const float t0 = b[0] * f.z1 * f.z2, t1 = b[1] * f.z0 * f.z2, t2 = b[2] * f.z0 * f.z1;
    const float id = __fdividef(1.0f, fmaxf(t0 + t1 + t2, MVR_K_EPS));
    b[0] = t0 * id; b[1] = t1 * id; b[2] = t2 * id;
  }
}

// grid: x = 32x8-pixel tiles of the image, y = view m, z = object b.  EXACT: the caller wants zbuf / bary / dists.
template <bool EXACT, int MINB>
__global__ void __launch_bounds__(MVR_THREADS, MINB) mesh_shade_kernel(const MeshParams p, int tiles_x) {
  const int b = blockIdx.z, m = blockIdx.y, n = b * p.M + m;
  int ty, tx;
  tile_rc(blockIdx.x, tiles_x, ty, tx);
  const int xi = tx * 32 + (threadIdx.x & 31), yi = ty * 8 + (threadIdx.x >> 5);
  const int HW = p.H * p.W;
  if (xi >= p.W || yi >= p.H) return;
  const int pix = yi * p.W + xi;
  const int k = p.layer;
  unsigned long long* kp = p.keys + (size_t)n * HW + pix;
  const unsigned long long key = *kp;
  if (k + 1 < p.K) {            // hand the layer to the next peeling pass
    p.prev[(size_t)n * HW + pix] = key;
    *kp = MVR_EMPTY_KEY;
  }
  int fid = -1;
  float w[3] = {-1.f, -1.f, -1.f}, bb[3] = {-1.f, -1.f, -1.f}, pz = -1.f, dd = -1.f;
  float out[3];
  if (k == 0) { out[0] = __ldg(p.bg_rgb); out[1] = __ldg(p.bg_rgb + 1); out[2] = __ldg(p.bg_rgb + 2); }
  if (key != MVR_EMPTY_KEY) {
    const int f0 = p.face_off[b], voff = p.vert_off[b], V = p.vert_off[b + 1] - voff;
    const bool persp = p.flags & MVR_PERSPECTIVE_CORRECT;
    fid = (int)(unsigned int)(key & 0xffffffffull);
    const int4 fi = __ldg(p.faces4 + f0 + fid);
    // every gather of this pixel is issued before the first use (one round trip to L2 instead of three)
    const Face fc = gather_face(p.pv + (size_t)p.M * voff + (size_t)m * V, fi);
    float4 X0, X1, X2, N0, N1, N2, c0, c1, c2;
    if (k == 0) {
      X0 = __ldg(p.verts4 + voff + fi.x); X1 = __ldg(p.verts4 + voff + fi.y); X2 = __ldg(p.verts4 + voff + fi.z);
      N0 = __ldg(p.normals4 + voff + fi.x); N1 = __ldg(p.normals4 + voff + fi.y); N2 = __ldg(p.normals4 + voff + fi.z);
      if (p.flags & MVR_RGB_PER_ELEMENT) { c0 = __ldg(p.rgb4 + voff + fi.x); c1 = __ldg(p.rgb4 + voff + fi.y); c2 = __ldg(p.rgb4 + voff + fi.z); }
      else { c0 = c1 = c2 = make_float4(__ldg(p.obj_rgb), __ldg(p.obj_rgb + 1), __ldg(p.obj_rgb + 2), 0.f); }
    }
    const float xf = __ldg(p.tab + xi), yf = __ldg(p.tab + p.W + yi);
    const FaceEdges fe = face_edges(fc);
    if (EXACT) {
      // The barycentrics are recomputed with the SAME exact operation sequence as the scatter (from the same
      // projected vertices), so the fragments returned to the caller are the rasterizer's, bit for bit.
      raster_test(fc, fe, persp, xf, yf, w, bb, pz);
      if (p.dists) {
        const float e01 = point_line_dist2(xf, yf, fc.x0, fc.y0, fc.x1, fc.y1);
        const float e02 = point_line_dist2(xf, yf, fc.x0, fc.y0, fc.x2, fc.y2);
        const float e12 = point_line_dist2(xf, yf, fc.x1, fc.y1, fc.x2, fc.y2);
        dd = -fminf(fminf(e01, e02), e12);
      }
    } else {
      shading_barycentrics(fc, fe, persp, xf, yf, bb);
    }
    pz = __uint_as_float((unsigned int)(key >> 32));
    if (k == 0) {
      const ShadeCtx sc = load_shade_ctx(p.light, p.light_stride, p.Cc, n);
      phong_pixel(bb, X0, X1, X2, N0, N1, N2, c0, c1, c2, sc, out);
    }
  }
  const size_t po = ((size_t)n * HW + pix) * p.K + k;
  p.pix_to_face[po] = fid;
  if (EXACT) {
    if (p.zbuf) p.zbuf[po] = pz;
    if (p.dists) p.dists[po] = dd;
    if (p.bary) { p.bary[3 * po] = bb[0]; p.bary[3 * po + 1] = bb[1]; p.bary[3 * po + 2] = bb[2]; }
  }
  if (k == 0) {
    const size_t io = (size_t)n * 3 * HW + pix;
    p.images[io] = out[0]; p.images[io + HW] = out[1]; p.images[io + 2 * (size_t)HW] = out[2];
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct MeshBwdParams {
  const float4* verts4; const float4* normals4; const float4* rgb4; const int4* faces4;
  const int* vert_off; const int* face_off;
  const float* R; const float* T; const float* Cc; const float* light; int light_stride;
  const float* obj_rgb;
  float k00, k11;
  int B, M, H, W, K, flags, ctas_per_view, tiles_x;
  const float4* pv; const float* tab;
  const int* pix_to_face; const float* grad_images;
  float* partials;       // (N, ctas_per_view, NWARPS, 16)
  float* grad_verts; float* grad_normals;
};

__device__ __forceinline__ float rcp_fast(float x) { return __fdividef(1.0f, x); }

// d/dv of v / max(|v|, eps)
__device__ __forceinline__ void normalize_bwd3(float vx, float vy, float vz, float eps, float gx, float gy, float gz,
                                               float& ox, float& oy, float& oz) {
  const float n2 = fmaf(vx, vx, fmaf(vy, vy, vz * vz));
  if (n2 > eps * eps) {
    const float inv = rsqrtf(n2);
    const float ux = vx * inv, uy = vy * inv, uz = vz * inv;
    const float d = fmaf(ux, gx, fmaf(uy, gy, uz * gz));
    ox = fmaf(-ux, d, gx) * inv; oy = fmaf(-uy, d, gy) * inv; oz = fmaf(-uz, d, gz) * inv;
  } else {
    const float inv = 1.f / eps;
    ox = gx * inv; oy = gy * inv; oz = gz * inv;
  }
}

template <int MINB>
__global__ void __launch_bounds__(MVR_THREADS, MINB) mesh_backward_kernel(const MeshBwdParams p) {
  const int tid = threadIdx.x;
  // grid: x = 32x32-pixel tiles, y = view m, z = object b; thread (lane, warp) owns pixels (x0+lane, y0+warp+8j)
  const int b = blockIdx.z, m = blockIdx.y, n = b * p.M + m, cta = blockIdx.x;
  int tyb, txb;
  tile_rc(cta, p.tiles_x, tyb, txb);
  const int xi = txb * 32 + (tid & 31), yi0 = tyb * 32 + (tid >> 5);
  const int HW = p.H * p.W;
  const int f0 = p.face_off[b], voff = p.vert_off[b], V = p.vert_off[b + 1] - voff;
  const float4* pvn = p.pv + (size_t)p.M * voff + (size_t)m * V;
  const bool persp = p.flags & MVR_PERSPECTIVE_CORRECT;
  const bool per_vertex_rgb = p.flags & MVR_RGB_PER_ELEMENT;
  // issue every load of this thread's pixels first (memory-level parallelism), then do the math
  int fids[BWD_PIX_PER_THREAD];
  float gin[BWD_PIX_PER_THREAD][3];
#pragma unroll
  for (int j = 0; j < BWD_PIX_PER_THREAD; ++j) {
    const int yi = yi0 + 8 * j;
    fids[j] = (xi < p.W && yi < p.H) ? __ldg(p.pix_to_face + ((size_t)n * HW + (size_t)yi * p.W + xi) * p.K) : -1;
  }
#pragma unroll
  for (int j = 0; j < BWD_PIX_PER_THREAD; ++j) {
    const int pix = (yi0 + 8 * j) * p.W + xi;
    if (fids[j] >= 0) {
      const size_t io = (size_t)n * 3 * HW + pix;
      gin[j][0] = __ldg(p.grad_images + io); gin[j][1] = __ldg(p.grad_images + io + HW); gin[j][2] = __ldg(p.grad_images + io + 2 * (size_t)HW);
    } else {
      gin[j][0] = gin[j][1] = gin[j][2] = 0.f;
    }
  }
  float acc[BWD_VALS];
#pragma unroll
  for (int i = 0; i < BWD_VALS; ++i) acc[i] = 0.f;
  bool any = false;
  const ShadeCtx sc = load_shade_ctx(p.light, p.light_stride, p.Cc, n);
  const float xf = xi < p.W ? __ldg(p.tab + xi) : 0.f;
  float4 ucol = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!per_vertex_rgb) ucol = make_float4(__ldg(p.obj_rgb), __ldg(p.obj_rgb + 1), __ldg(p.obj_rgb + 2), 0.f);
#pragma unroll 1
  for (int j = 0; j < BWD_PIX_PER_THREAD; ++j) {
    const int fid = fids[j];
    const float g0 = gin[j][0], g1 = gin[j][1], g2 = gin[j][2];
    if (fid < 0 || (g0 == 0.f && g1 == 0.f && g2 == 0.f)) continue;
    any = true;
    const int yi = yi0 + 8 * j;
    const int4 fi = __ldg(p.faces4 + f0 + fid);
    // ---- forward recompute from the projected vertices (exact IEEE projection, done once per view by
    // mesh_project_kernel: for small faces the barycentrics amplify a 1-ulp change of a vertex by |xy| / area);
    // everything downstream is well conditioned and uses fast reciprocals ----
    const Face fc = gather_face(pvn, fi);
    const float4 X0 = __ldg(p.verts4 + voff + fi.x), X1 = __ldg(p.verts4 + voff + fi.y), X2 = __ldg(p.verts4 + voff + fi.z);
    const float4 N0 = __ldg(p.normals4 + voff + fi.x), N1 = __ldg(p.normals4 + voff + fi.y), N2 = __ldg(p.normals4 + voff + fi.z);
    float4 c0 = ucol, c1 = ucol, c2 = ucol;
    if (per_vertex_rgb) { c0 = __ldg(p.rgb4 + voff + fi.x); c1 = __ldg(p.rgb4 + voff + fi.y); c2 = __ldg(p.rgb4 + voff + fi.z); }
    const FaceEdges fe = face_edges(fc);
    const float yf = __ldg(p.tab + p.W + yi);
    const float e0 = (xf - fc.x1) * fe.A0 - (yf - fc.y1) * fe.B0;
    const float e1 = (xf - fc.x2) * fe.A1 - (yf - fc.y2) * fe.B1;
    const float e2 = (xf - fc.x0) * fe.A2 - (yf - fc.y0) * fe.B2;
    const float inv_area = rcp_fast(fe.area_p);
    const float w0 = e0 * inv_area, w1 = e1 * inv_area, w2 = e2 * inv_area;
    float bb[3] = {w0, w1, w2};
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, id = 1.f;
    bool clamped = false;
    if (persp) {
      t0 = w0 * fc.z1 * fc.z2; t1 = w1 * fc.z0 * fc.z2; t2 = w2 * fc.z0 * fc.z1;
      const float st = t0 + t1 + t2;
      clamped = st < MVR_K_EPS;
      id = rcp_fast(fmaxf(st, MVR_K_EPS));
      bb[0] = t0 * id; bb[1] = t1 * id; bb[2] = t2 * id;
    }
    // ---- Phong backward ----
    const float3 P = interp(bb, X0, X1, X2);
    const float3 Nn = interp(bb, N0, N1, N2);
    const float3 tex = interp(bb, c0, c1, c2);
    const float in = inv_norm_clamped(Nn.x, Nn.y, Nn.z, 1e-6f);
    const float nx = Nn.x * in, ny = Nn.y * in, nz = Nn.z * in;
    const float cosang = fmaf(nx, sc.lx, fmaf(ny, sc.ly, nz * sc.lz));
    const float diff = fmaxf(cosang, 0.f);
    const float vx = sc.cx - P.x, vy = sc.cy - P.y, vz = sc.cz - P.z;
    const float iv = inv_norm_clamped(vx, vy, vz, 1e-6f);
    const float vhx = vx * iv, vhy = vy * iv, vhz = vz * iv;
    const float rx = fmaf(2.f * cosang, nx, -sc.lx), ry = fmaf(2.f * cosang, ny, -sc.ly), rz = fmaf(2.f * cosang, nz, -sc.lz);
    const float dt = fmaf(vhx, rx, fmaf(vhy, ry, vhz * rz));
    const bool lit = cosang > 0.f;
    const float alpha = (dt > 0.f && lit) ? dt : 0.f;
    const float kd = fmaf(MVR_DIFFUSE, diff, MVR_AMBIENT);
    const float gtx = g0 * kd, gty = g1 * kd, gtz = g2 * kd;
    const float gdiff = MVR_DIFFUSE * fmaf(g0, tex.x, fmaf(g1, tex.y, g2 * tex.z));
    const float gs = MVR_SPECULAR * (g0 + g1 + g2);
    const float a2 = alpha * alpha, a4 = a2 * a2, a8 = a4 * a4, a16 = a8 * a8, a32 = a16 * a16;
    const float a63 = a32 * a16 * a8 * a4 * a2 * alpha;
    const float gdt = (dt > 0.f && lit) ? gs * 64.f * a63 : 0.f;
    const float gvhx = gdt * rx, gvhy = gdt * ry, gvhz = gdt * rz;
    const float grx = gdt * vhx, gry = gdt * vhy, grz = gdt * vhz;
    const float gcos = (lit ? gdiff : 0.f) + 2.f * fmaf(grx, nx, fmaf(gry, ny, grz * nz));
    const float gnx = fmaf(2.f * cosang, grx, gcos * sc.lx), gny = fmaf(2.f * cosang, gry, gcos * sc.ly), gnz = fmaf(2.f * cosang, grz, gcos * sc.lz);
    float gNx, gNy, gNz, gvx, gvy, gvz;
    normalize_bwd3(Nn.x, Nn.y, Nn.z, 1e-6f, gnx, gny, gnz, gNx, gNy, gNz);
    normalize_bwd3(vx, vy, vz, 1e-6f, gvhx, gvhy, gvhz, gvx, gvy, gvz);
    acc[12] += gvx; acc[13] += gvy; acc[14] += gvz;   // dC
    // d bary_i = gtex.col_i + gN.n_i + gP.X_i  with gP = -gv
    float gb0 = fmaf(gtx, c0.x, fmaf(gty, c0.y, gtz * c0.z)) + fmaf(gNx, N0.x, fmaf(gNy, N0.y, gNz * N0.z)) - fmaf(gvx, X0.x, fmaf(gvy, X0.y, gvz * X0.z));
    float gb1 = fmaf(gtx, c1.x, fmaf(gty, c1.y, gtz * c1.z)) + fmaf(gNx, N1.x, fmaf(gNy, N1.y, gNz * N1.z)) - fmaf(gvx, X1.x, fmaf(gvy, X1.y, gvz * X1.z));
    float gb2 = fmaf(gtx, c2.x, fmaf(gty, c2.y, gtz * c2.z)) + fmaf(gNx, N2.x, fmaf(gNy, N2.y, gNz * N2.z)) - fmaf(gvx, X2.x, fmaf(gvy, X2.y, gvz * X2.z));
    // ---- [upstream] BarycentricPerspectiveCorrectionBackward ----
    float dz0 = 0.f, dz1 = 0.f, dz2 = 0.f;
    if (persp) {
      // b = t / sum(t) is invariant to a common shift of d/db (its Jacobian annihilates constants), so
      // remove k = sum(b_i gb_i) first: the d/d denom term then vanishes identically instead of cancelling
      // O(1/area) terms in fp32.  Same gradient in exact arithmetic; not valid when denom was clamped.
      if (!clamped) {
        const float kk = fmaf(bb[0], gb0, fmaf(bb[1], gb1, bb[2] * gb2));
        gb0 -= kk; gb1 -= kk; gb2 -= kk;
      }
      const float gden = clamped ? -(gb0 * t0 + gb1 * t1 + gb2 * t2) * id * id : 0.f;
      const float gt0 = fmaf(gb0, id, gden), gt1 = fmaf(gb1, id, gden), gt2 = fmaf(gb2, id, gden);
      gb0 = gt0 * fc.z1 * fc.z2; gb1 = gt1 * fc.z0 * fc.z2; gb2 = gt2 * fc.z0 * fc.z1;
      dz0 = gt1 * w1 * fc.z2 + gt2 * w2 * fc.z1;
      dz1 = gt0 * w0 * fc.z2 + gt2 * w2 * fc.z0;
      dz2 = gt0 * w0 * fc.z1 + gt1 * w1 * fc.z0;
    }
    // ---- [upstream] BarycentricCoordsBackward / EdgeFunctionBackward ----
    const float ge0 = gb0 * inv_area, ge1 = gb1 * inv_area, ge2 = gb2 * inv_area;
    const float garea = -(gb0 * e0 + gb1 * e1 + gb2 * e2) * inv_area * inv_area;
    float gx0, gy0, gx1, gy1, gx2, gy2;
    // E(p,a,b): dE/da = (py-by, bx-px), dE/db = (ay-py, px-ax)
    gx1 = ge0 * (yf - fc.y2); gy1 = ge0 * (fc.x2 - xf); gx2 = ge0 * (fc.y1 - yf); gy2 = ge0 * (xf - fc.x1);          // e0 = E(p,v1,v2)
    gx2 += ge1 * (yf - fc.y0); gy2 += ge1 * (fc.x0 - xf); gx0 = ge1 * (fc.y2 - yf); gy0 = ge1 * (xf - fc.x2);        // e1 = E(p,v2,v0)
    gx0 += ge2 * (yf - fc.y1); gy0 += ge2 * (fc.x1 - xf); gx1 += ge2 * (fc.y0 - yf); gy1 += ge2 * (xf - fc.x0);      // e2 = E(p,v0,v1)
    // area = E(v2, v0, v1)
    gx0 += garea * (fc.y2 - fc.y1); gy0 += garea * (fc.x1 - fc.x2);
    gx1 += garea * (fc.y0 - fc.y2); gy1 += garea * (fc.x2 - fc.x0);
    gx2 += garea * (fc.y1 - fc.y0); gy2 += garea * (fc.x0 - fc.x1);
    // ---- projection backward + X R + T backward: x_ndc = (px k00) / pz, so px k00 = x_ndc pz ----
    const float gxn[3] = {gx0, gx1, gx2}, gyn[3] = {gy0, gy1, gy2}, gzn[3] = {dz0, dz1, dz2};
    const float xn[3] = {fc.x0, fc.x1, fc.x2}, yn[3] = {fc.y0, fc.y1, fc.y2}, zv[3] = {fc.z0, fc.z1, fc.z2};
    const float4 Xs[3] = {X0, X1, X2};
    const int vid[3] = {fi.x, fi.y, fi.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float iz = rcp_fast(zv[i]);
      const float gpx = gxn[i] * p.k00 * iz;
      const float gpy = gyn[i] * p.k11 * iz;
      const float gpz = gzn[i] - (gxn[i] * xn[i] + gyn[i] * yn[i]) * iz;
      acc[0] = fmaf(Xs[i].x, gpx, acc[0]); acc[1] = fmaf(Xs[i].x, gpy, acc[1]); acc[2] = fmaf(Xs[i].x, gpz, acc[2]);
      acc[3] = fmaf(Xs[i].y, gpx, acc[3]); acc[4] = fmaf(Xs[i].y, gpy, acc[4]); acc[5] = fmaf(Xs[i].y, gpz, acc[5]);
      acc[6] = fmaf(Xs[i].z, gpx, acc[6]); acc[7] = fmaf(Xs[i].z, gpy, acc[7]); acc[8] = fmaf(Xs[i].z, gpz, acc[8]);
      acc[9] += gpx; acc[10] += gpy; acc[11] += gpz;
      if (p.grad_verts) {
        const float* r = p.R + 9 * (size_t)n;
        float* o = p.grad_verts + 3 * (size_t)(voff + vid[i]);
        atomicAdd(o + 0, fmaf(__ldg(r + 0), gpx, fmaf(__ldg(r + 1), gpy, __ldg(r + 2) * gpz)) - bb[i] * gvx);
        atomicAdd(o + 1, fmaf(__ldg(r + 3), gpx, fmaf(__ldg(r + 4), gpy, __ldg(r + 5) * gpz)) - bb[i] * gvy);
        atomicAdd(o + 2, fmaf(__ldg(r + 6), gpx, fmaf(__ldg(r + 7), gpy, __ldg(r + 8) * gpz)) - bb[i] * gvz);
      }
      if (p.grad_normals) {
        float* o = p.grad_normals + 3 * (size_t)(voff + vid[i]);
        atomicAdd(o + 0, bb[i] * gNx); atomicAdd(o + 1, bb[i] * gNy); atomicAdd(o + 2, bb[i] * gNz);
      }
    }
  }
  // one partial per WARP, no block barrier: a warp retires as soon as its own pixels are done
  float* out = p.partials + (((size_t)n * p.ctas_per_view + cta) * NWARPS + (tid >> 5)) * 16;
  const int lane = tid & 31;
  if (!__any_sync(0xffffffffu, any)) {   // background-only warp
    if (lane < 16) out[lane] = 0.f;
    return;
  }
#pragma unroll
  for (int i = 0; i < BWD_VALS; ++i) acc[i] = warp_sum(acc[i]);
  float mine = 0.f;
#pragma unroll
  for (int i = 0; i < BWD_VALS; ++i) mine = lane == i ? acc[i] : mine;
  if (lane < 16) out[lane] = mine;
}

// fixed-order sum of the per-warp partials: one warp per view -> gR, gT, gC
__global__ void mesh_backward_reduce_kernel(const float* __restrict__ partials, int N, int n_parts,
                                            float* __restrict__ gR, float* __restrict__ gT, float* __restrict__ gC) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  // lane l < 16 owns value l of even parts, lane l >= 16 value l-16 of odd parts
  const int v = lane & 15, par = lane >> 4;
  float s = 0.f;
  for (int t = par; t < n_parts; t += 2) s += partials[((size_t)n * n_parts + t) * 16 + v];
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  if (lane < 9) gR[9 * (size_t)n + lane] = s;
  else if (lane < 12) gT[3 * (size_t)n + lane - 9] = s;
  else if (lane < 15) gC[3 * (size_t)n + lane - 12] = s;
}

}  // namespace mvr

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace mvr;

extern "C" size_t mvr_mesh_geometry_bytes(int64_t total_verts, int64_t total_faces) {
  if (total_verts < 0 || total_faces < 0) return 0;
  return geom_layout(total_verts, total_faces).total;
}

extern "C" int mvr_mesh_prepare(const float* verts, const void* faces, const int* vert_off, const int* face_off,
                                int B, int64_t total_verts, int64_t total_faces, int max_faces,
                                const float* vert_rgb, int flags, void* geometry, size_t geometry_bytes,
                                void* stream) {
  if (B < 0 || total_verts < 0 || total_faces < 0 || max_faces < 0) { set_error("mvr_mesh_prepare: negative size"); return -1; }
  if (total_verts > 0x7fffffffLL || total_faces > 0x7fffffffLL) { set_error("mvr_mesh_prepare: more than 2^31-1 packed verts/faces"); return -2; }
  const GeomLayout g = geom_layout(total_verts, total_faces);
  if (geometry_bytes < g.total) { set_error("mvr_mesh_prepare: geometry buffer too small (%zu < %zu)", geometry_bytes, g.total); return -3; }
  if (B == 0 || total_verts == 0) return 0;
  if (!verts || !vert_off || !face_off || !geometry || (total_faces > 0 && !faces)) { set_error("mvr_mesh_prepare: null pointer"); return -4; }
  if ((flags & MVR_RGB_PER_ELEMENT) && !vert_rgb) { set_error("mvr_mesh_prepare: MVR_RGB_PER_ELEMENT without vert_rgb"); return -5; }
  char* base = (char*)geometry;
  cudaStream_t st = (cudaStream_t)stream;
  float4* verts4 = (float4*)(base + g.verts4);
  float4* normals4 = (float4*)(base + g.normals4);
  float4* rgb4 = (float4*)(base + g.rgb4);
  int4* faces4 = (int4*)(base + g.faces4);
  double* nacc = (double*)(base + g.nacc);
  const int tb = 256;
  MVR_LAUNCH(geom_pack_verts_kernel, (unsigned)((total_verts + tb - 1) / tb), tb, 0, st, verts, (flags & MVR_RGB_PER_ELEMENT) ? vert_rgb : nullptr, total_verts, verts4, rgb4, nacc);
  if (total_faces > 0 && max_faces > 0) {
    dim3 grid((max_faces + tb - 1) / tb, B);
    if (flags & MVR_FACES_I64) MVR_LAUNCH(geom_pack_faces_kernel<long long>, grid, tb, 0, st, (const long long*)faces, vert_off, face_off, verts4, faces4, nacc);
    else MVR_LAUNCH(geom_pack_faces_kernel<int>, grid, tb, 0, st, (const int*)faces, vert_off, face_off, verts4, faces4, nacc);
  }
  MVR_LAUNCH(geom_finish_normals_kernel, (unsigned)((total_verts + tb - 1) / tb), tb, 0, st, nacc, total_verts, normals4);
  return check_launch("mvr_mesh_prepare");
}

extern "C" int mvr_mesh_get_normals(const void* geometry, int64_t total_verts, int64_t total_faces, float* normals, void* stream) {
  if (total_verts <= 0) return 0;
  if (!geometry || !normals) { set_error("mvr_mesh_get_normals: null pointer"); return -1; }
  const GeomLayout g = geom_layout(total_verts, total_faces);
  MVR_LAUNCH(geom_get_normals_kernel, (unsigned)((total_verts + 255) / 256), 256, 0, (cudaStream_t)stream, (const float4*)((const char*)geometry + g.normals4), total_verts, normals);
  return check_launch("mvr_mesh_get_normals");
}

extern "C" size_t mvr_mesh_workspace_bytes(int B, int M, int H, int W, int K, int64_t total_verts) {
  if (B < 0 || M < 0 || H <= 0 || W <= 0 || K < 1 || total_verts < 0) return 0;
  return ws_layout(B, M, H, W, K, total_verts).total;
}

static int check_mesh_common(const char* who, int B, int M, int H, int W, int K, int64_t tv, int64_t tf, int max_verts) {
  if (B < 0 || M < 0 || tv < 0 || tf < 0 || max_verts < 0) { set_error("%s: negative size", who); return -1; }
  if (H <= 0 || W <= 0 || H > 4096 || W > 4096) { set_error("%s: image size %dx%d outside [1, 4096]", who, H, W); return -2; }
  if (K < 1 || K > 64) { set_error("%s: faces_per_pixel %d outside [1, 64]", who, K); return -3; }
  if ((int64_t)B * M * (((int64_t)H * W + 255) / 256 + 1) > 0x7fffffffLL || B > 65535 || M > 65535) { set_error("%s: too many views", who); return -4; }
  return 0;
}

// tuning knob (profiling only): MVR_SCATTER_MINB=3 trades occupancy (3 CTAs/SM, 85 registers) for fewer
// rematerialised instructions in the scatter kernel's inner loops; default 4 CTAs/SM
static int scatter_minb() {
  static const int v = [] { const char* e = getenv("MVR_SCATTER_MINB"); return (e && atoi(e) == 3) ? 3 : 4; }();
  return v;
}

static int shade_minb() {
  static const int v = [] { const char* e = getenv("MVR_SHADE_MINB"); const int x = e ? atoi(e) : 4; return (x == 5 || x == 6) ? x : 4; }();
  return v;
}
static int backward_minb() {
  static const int v = [] { const char* e = getenv("MVR_BWD_MINB"); return (e && atoi(e) == 2) ? 2 : 3; }();
  return v;
}

// world -> NDC of every (view, vertex) + the pixel-centre table, into the front of the workspace
static int launch_project(const char* who, const GeomLayout& g, const WsLayout& w, const void* geometry,
                          const int* vert_off, const float* R, const float* T, int B, int M, int H, int W,
                          int max_verts, float k00, float k11, void* workspace, cudaStream_t st) {
  const char* gb = (const char*)geometry;
  char* wb = (char*)workspace;
  const dim3 grid((unsigned)((max_verts + MVR_THREADS - 1) / MVR_THREADS > 0 ? (max_verts + MVR_THREADS - 1) / MVR_THREADS : 1),
                  (unsigned)M, (unsigned)B);
  MVR_LAUNCH(mesh_project_kernel, grid, MVR_THREADS, 0, st, (const float4*)(gb + g.verts4), vert_off, R, T, M, H, W,
             k00, k11, (float4*)(wb + w.pv), (float*)(wb + w.tab));
  return check_launch(who);
}

extern "C" int mvr_mesh_forward(const void* geometry, const int* vert_off, const int* face_off, int B, int M,
                                int64_t total_verts, int64_t total_faces, int max_verts, int max_faces,
                                const float* R, const float* T, const float* Cc, const float* light,
                                int light_stride, const float* obj_rgb, const float* bg_rgb, float k00, float k11,
                                float z_clip, int H, int W, int K, int flags, float* images, int* pix_to_face,
                                float* zbuf, float* bary, float* dists, int64_t* counters, void* workspace,
                                size_t workspace_bytes, void* stream) {
  int rc = check_mesh_common("mvr_mesh_forward", B, M, H, W, K, total_verts, total_faces, max_verts);
  if (rc) return rc;
  const int64_t N = (int64_t)B * M;
  if (N == 0) return 0;
  if (!geometry || !vert_off || !face_off || !R || !T || !Cc || !light || !bg_rgb || !images || !pix_to_face || !workspace) {
    set_error("mvr_mesh_forward: null pointer"); return -5;
  }
  if (!(flags & MVR_RGB_PER_ELEMENT) && !obj_rgb) { set_error("mvr_mesh_forward: obj_rgb is NULL and the geometry has no per-vertex colours"); return -6; }
  const WsLayout w = ws_layout(B, M, H, W, K, total_verts);
  if (workspace_bytes < w.total) { set_error("mvr_mesh_forward: workspace too small (%zu < %zu)", workspace_bytes, w.total); return -7; }
  const int chunks_per_view = max_faces > 0 ? (max_faces + FACES_PER_CTA - 1) / FACES_PER_CTA : 0;
  if (N * (int64_t)(chunks_per_view + 1) > 0x7fffffffLL) { set_error("mvr_mesh_forward: too many face chunks"); return -8; }
  const GeomLayout g = geom_layout(total_verts, total_faces);
  const char* gb = (const char*)geometry;
  char* wb = (char*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  MeshParams p;
  p.verts4 = (const float4*)(gb + g.verts4); p.normals4 = (const float4*)(gb + g.normals4);
  p.rgb4 = (const float4*)(gb + g.rgb4); p.faces4 = (const int4*)(gb + g.faces4);
  p.vert_off = vert_off; p.face_off = face_off;
  p.R = R; p.T = T; p.Cc = Cc; p.light = light; p.light_stride = light_stride;
  p.obj_rgb = obj_rgb; p.bg_rgb = bg_rgb;
  p.k00 = k00; p.k11 = k11; p.z_clip = z_clip;
  p.B = B; p.M = M; p.H = H; p.W = W; p.K = K; p.flags = flags;
  p.chunks_per_view = chunks_per_view; p.layer = 0;
  p.item_cap = (flags & MVR_TEST_TINY_QUEUES) ? 24 : ITEM_CAP;
  p.wcap = (flags & MVR_TEST_TINY_QUEUES) ? 5 : WCAP;
  p.pv = (float4*)(wb + w.pv); p.tab = (float*)(wb + w.tab);
  p.keys = (unsigned long long*)(wb + w.keys); p.prev = (unsigned long long*)(wb + w.prev);
  p.images = images; p.pix_to_face = pix_to_face; p.zbuf = zbuf; p.bary = bary; p.dists = dists;
  p.counters = (long long*)counters;
  const size_t HW = (size_t)H * W;
  cudaError_t e = cudaMemsetAsync(p.keys, 0xFF, (size_t)N * HW * 8, st);      // every key = EMPTY
  if (e != cudaSuccess) { set_error("mvr_mesh_forward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  rc = launch_project("mesh_project_kernel", g, w, geometry, vert_off, R, T, B, M, H, W, max_verts, k00, k11, workspace, st);
  if (rc) return rc;
  const size_t tab_smem = ((size_t)W + H) * sizeof(float);
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8;
  const dim3 shade_grid((unsigned)(tiles_x * tiles_y), (unsigned)M, (unsigned)B);
  for (int k = 0; k < K; ++k) {
    p.layer = k;
    if (chunks_per_view > 0) {
      if (scatter_minb() == 3) MVR_LAUNCH(mesh_scatter_kernel<3>, (unsigned)(N * chunks_per_view), MVR_THREADS, tab_smem, st, p);
      else MVR_LAUNCH(mesh_scatter_kernel<4>, (unsigned)(N * chunks_per_view), MVR_THREADS, tab_smem, st, p);
      rc = check_launch("mesh_scatter_kernel");
      if (rc) return rc;
    }
    if (zbuf || bary || dists) MVR_LAUNCH((mesh_shade_kernel<true, 3>), shade_grid, MVR_THREADS, 0, st, p, tiles_x);
    else if (shade_minb() == 5) MVR_LAUNCH((mesh_shade_kernel<false, 5>), shade_grid, MVR_THREADS, 0, st, p, tiles_x);
    else if (shade_minb() == 6) MVR_LAUNCH((mesh_shade_kernel<false, 6>), shade_grid, MVR_THREADS, 0, st, p, tiles_x);
    else MVR_LAUNCH((mesh_shade_kernel<false, 4>), shade_grid, MVR_THREADS, 0, st, p, tiles_x);
    rc = check_launch("mesh_shade_kernel");
    if (rc) return rc;
  }
  return 0;
}

extern "C" int mvr_mesh_backward(const void* geometry, const int* vert_off, const int* face_off, int B, int M,
                                 int64_t total_verts, int64_t total_faces, int max_verts, const float* R,
                                 const float* T, const float* Cc, const float* light, int light_stride,
                                 const float* obj_rgb, float k00, float k11, int H, int W, int K, int flags,
                                 const int* pix_to_face, const float* grad_images, float* gR, float* gT, float* gC,
                                 float* grad_verts, float* grad_normals, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  int rc = check_mesh_common("mvr_mesh_backward", B, M, H, W, K, total_verts, total_faces, max_verts);
  if (rc) return rc;
  const int64_t N = (int64_t)B * M;
  if (N == 0) return 0;
  if (!geometry || !vert_off || !face_off || !R || !T || !Cc || !light || !pix_to_face || !grad_images || !gR || !gT || !gC || !workspace) {
    set_error("mvr_mesh_backward: null pointer"); return -5;
  }
  if (!(flags & MVR_RGB_PER_ELEMENT) && !obj_rgb) { set_error("mvr_mesh_backward: obj_rgb is NULL"); return -6; }
  const WsLayout w = ws_layout(B, M, H, W, K, total_verts);
  if (workspace_bytes < w.total) { set_error("mvr_mesh_backward: workspace too small (%zu < %zu)", workspace_bytes, w.total); return -7; }
  const GeomLayout g = geom_layout(total_verts, total_faces);
  const char* gb = (const char*)geometry;
  char* wb = (char*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  // the workspace is scratch (it may have served another render since the forward): project again, 2% of the step
  rc = launch_project("mesh_project_kernel", g, w, geometry, vert_off, R, T, B, M, H, W, max_verts, k00, k11, workspace, st);
  if (rc) return rc;
  MeshBwdParams p;
  p.verts4 = (const float4*)(gb + g.verts4); p.normals4 = (const float4*)(gb + g.normals4);
  p.rgb4 = (const float4*)(gb + g.rgb4); p.faces4 = (const int4*)(gb + g.faces4);
  p.vert_off = vert_off; p.face_off = face_off;
  p.R = R; p.T = T; p.Cc = Cc; p.light = light; p.light_stride = light_stride; p.obj_rgb = obj_rgb;
  p.k00 = k00; p.k11 = k11;
  p.B = B; p.M = M; p.H = H; p.W = W; p.K = K; p.flags = flags; p.ctas_per_view = w.bwd_ctas_per_view; p.tiles_x = (W + 31) / 32;
  p.pv = (const float4*)(wb + w.pv); p.tab = (const float*)(wb + w.tab);
  p.pix_to_face = pix_to_face; p.grad_images = grad_images;
  p.partials = (float*)(wb + w.partials); p.grad_verts = grad_verts; p.grad_normals = grad_normals;
  const dim3 bgrid((unsigned)w.bwd_ctas_per_view, (unsigned)M, (unsigned)B);
  if (backward_minb() == 2) MVR_LAUNCH(mesh_backward_kernel<2>, bgrid, MVR_THREADS, 0, st, p);
  else MVR_LAUNCH(mesh_backward_kernel<3>, bgrid, MVR_THREADS, 0, st, p);
  rc = check_launch("mesh_backward_kernel");
  if (rc) return rc;
  const int wpb = 8;
  MVR_LAUNCH(mesh_backward_reduce_kernel, (unsigned)((N + wpb - 1) / wpb), wpb * 32, 0, st, (const float*)(wb + w.partials), (int)N, w.bwd_ctas_per_view * NWARPS, gR, gT, gC);
  return check_launch("mesh_backward_reduce_kernel");
}
