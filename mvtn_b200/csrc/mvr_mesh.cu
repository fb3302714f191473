// mvr_mesh.cu -- mesh path of MVRenderer (renderer.py:65-114) for sm_100a.
//
//   prepare : pack verts/normals/colours to float4 and faces to int4; area-weighted vertex normals
//             ([upstream] Meshes._compute_vertex_normals) once per OBJECT, not per view.
//   coarse  : mesh_bin_kernel   -- one thread per (view, face): project, cull, exact pixel bbox ->
//             32x32-pixel tile bins.  Per-CTA shared-memory histogram + ONE global atomic per CTA
//             reserves a contiguous pool range; per-(view, chunk, tile) segment descriptors make the
//             bins exact-sized with no fixed max_faces_per_bin.  A chunk that does not fit in the pool
//             is flagged and later scanned unbinned (never drops a face).
//   fine    : mesh_fine_kernel  -- one CTA per (view, tile).  Faces of the tile are SCATTERED: each
//             thread takes a face, walks only the pixels of its (tile-clipped) bbox and does a 64-bit
//             (z, face) min into a shared-memory key per pixel; big faces are handed to whole warps.
//             K > 1 peels layers (pass k keeps keys > layer k-1).  The epilogue recomputes the
//             barycentrics of the winning face, Phong-shades, hard-blends the background and writes
//             planar (n,3,H,W) images + pix_to_face (+ optional zbuf/bary/dists) with full-line stores.
//   backward: mesh_backward_kernel -- per pixel recompute (no fragment traffic), chain
//             d image -> Phong -> barycentrics -> NDC verts -> view verts -> (dR, dT, dC), block-reduced to
//             one partial per (view, tile) and summed in fixed order (deterministic, no float atomics).
#include "mvr_common.cuh"

namespace mvr {

constexpr int TILE = 32;              // pixels per tile side (row segment = 128 B = one line)
constexpr int TILE_PIX = TILE * TILE;
constexpr int MAX_CHUNKS = 256;       // face chunks per view (bin kernel CTAs per view)
constexpr int MIN_FACES_PER_CHUNK = 2048;
constexpr int BWD_VALS = 15;          // dR 9, dT 3, dC 3

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

static int faces_per_chunk(int max_faces) {
  int fpc = (max_faces + MAX_CHUNKS - 1) / MAX_CHUNKS;
  fpc = (fpc + MVR_THREADS - 1) / MVR_THREADS * MVR_THREADS;
  return fpc < MIN_FACES_PER_CHUNK ? MIN_FACES_PER_CHUNK : fpc;
}

struct WsLayout {
  size_t counter, seg, pool, total;
  int64_t pool_cap;
  int n_tiles, tiles_x, max_chunks, fpc;
};
static WsLayout ws_layout(int B, int M, int H, int W, int64_t total_faces, int max_faces) {
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  WsLayout w;
  w.tiles_x = (W + TILE - 1) / TILE;
  w.n_tiles = w.tiles_x * ((H + TILE - 1) / TILE);
  w.fpc = faces_per_chunk(max_faces);
  w.max_chunks = max_faces > 0 ? (max_faces + w.fpc - 1) / w.fpc : 1;
  const int64_t N = (int64_t)B * M;
  w.pool_cap = 2 * (int64_t)M * total_faces + 32 * N * w.n_tiles + 4096;
  size_t o = 0;
  w.counter = o; o = al(o + 256);
  w.seg = o; o = al(o + (size_t)N * w.max_chunks * w.n_tiles * sizeof(int2));
  w.pool = o; o = al(o + (size_t)w.pool_cap * sizeof(int));
  // the backward pass reuses the front of the workspace for its (view, tile) partial sums
  size_t bwd = al((size_t)N * w.n_tiles * 16 * sizeof(float));
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
// shared device code: face setup and the per-(face, pixel) test
// ------------------------------------------------------------------------------------------------
struct MeshParams {
  const float4* verts4; const float4* normals4; const float4* rgb4; const int4* faces4;
  const int* vert_off; const int* face_off;
  const float* R; const float* T; const float* Cc; const float* light; int light_stride;
  const float* obj_rgb; const float* bg_rgb;
  float k00, k11, z_clip;
  int B, M, H, W, K, flags;
  int n_tiles, tiles_x, max_chunks, fpc;
  long long pool_cap;
  int* pool_counter; int2* seg; int* pool;
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

__device__ __forceinline__ Face load_face(const MeshParams& p, const Camera& cam, int voff, const int4 fi) {
  Face f;
  project_vertex(cam, __ldg(p.verts4 + voff + fi.x), p.k00, p.k11, f.x0, f.y0, f.z0);
  project_vertex(cam, __ldg(p.verts4 + voff + fi.y), p.k00, p.k11, f.x1, f.y1, f.z1);
  project_vertex(cam, __ldg(p.verts4 + voff + fi.z), p.k00, p.k11, f.x2, f.y2, f.z2);
  return f;
}

// Face-level rejection ([upstream] clip.py near cull, CheckPointOutsideBoundingBox z_invalid,
// RasterizeMeshesNaiveCpu zero-area / back-face tests) and the exact pixel bbox (inclusive ranges), clipped to the
// rectangle [tx0,tx1] x [ty0,ty1] and made exact with pixel-centre tables (tab[i - t0], decreasing in i).
__device__ __forceinline__ void tile_range(float vmin, float vmax, int S1, int S2, int t0, int t1, const float* tab,
                                           int& ilo, int& ihi) {
  float range = 2.0f;
  if (S1 > S2) range = ((float)(S1 / S2)) * range;
  const float offset = range / 2.0f;
  // centre of pixel i is c(S1-1-i) with c(j) = -offset + (range*j + offset)/S1, so i decreases as the coordinate grows
  float jhi = floorf(((vmax + offset) * (float)S1 - offset) / range);
  float jlo = ceilf(((vmin + offset) * (float)S1 - offset) / range);
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
__device__ __forceinline__ bool face_tile_bbox(const Face& f, const MeshParams& p, int tx0, int tx1, int ty0, int ty1,
                                               const float* s_xf, const float* s_yf, int& xi_lo, int& xi_hi,
                                               int& yi_lo, int& yi_hi) {
  if (p.z_clip >= 0.f && f.z0 < p.z_clip && f.z1 < p.z_clip && f.z2 < p.z_clip) return false;
  const float zmin = fminf(fminf(f.z0, f.z1), f.z2);
  if (zmin < MVR_K_EPS) return false;
  const float face_area = (f.x0 - f.x1) * (f.y2 - f.y1) - (f.y0 - f.y1) * (f.x2 - f.x1);
  if ((p.flags & MVR_CULL_BACKFACES) && face_area < 0.f) return false;
  if (face_area <= MVR_K_EPS && face_area >= -1.0f * MVR_K_EPS) return false;
  const float xmin = fminf(fminf(f.x0, f.x1), f.x2), xmax = fmaxf(fmaxf(f.x0, f.x1), f.x2);
  const float ymin = fminf(fminf(f.y0, f.y1), f.y2), ymax = fmaxf(fmaxf(f.y0, f.y1), f.y2);
  tile_range(xmin, xmax, p.W, p.H, tx0, tx1, s_xf, xi_lo, xi_hi);
  if (xi_lo > xi_hi) return false;
  tile_range(ymin, ymax, p.H, p.W, ty0, ty1, s_yf, yi_lo, yi_hi);
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

// exclusive scan of one int per thread over the block; returns the exclusive prefix, *total = sum
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp /* [8] */, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect s_warp reuse
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < MVR_THREADS / 32; ++w) {
    const int s = s_warp[w];
    if (w < warp) base += s;
    tot += s;
  }
  *total = tot;
  return base + inc - v;
}

// ------------------------------------------------------------------------------------------------
// coarse pass
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MVR_THREADS) mesh_bin_kernel(const MeshParams p) {
  extern __shared__ int s_dyn[];       // [n_tiles] counts, [n_tiles] exclusive offsets / cursors, [W] + [H] pixel centres
  __shared__ int s_warp[8];
  __shared__ int s_base;
  int* s_cnt = s_dyn;
  int* s_off = s_dyn + p.n_tiles;
  float* s_xf = (float*)(s_dyn + 2 * p.n_tiles);
  float* s_yf = s_xf + p.W;
  const int n = blockIdx.x / p.max_chunks, c = blockIdx.x % p.max_chunks;
  const int b = n / p.M;
  const int f0 = p.face_off[b], F = p.face_off[b + 1] - f0;
  const int fbeg = c * p.fpc, fend = min(F, fbeg + p.fpc);
  int2* seg = p.seg + ((size_t)n * p.max_chunks + c) * p.n_tiles;
  if (fbeg >= fend) {
    for (int t = threadIdx.x; t < p.n_tiles; t += MVR_THREADS) seg[t] = make_int2(0, 0);
    return;
  }
  for (int t = threadIdx.x; t < p.n_tiles; t += MVR_THREADS) s_cnt[t] = 0;
  for (int i = threadIdx.x; i < p.W; i += MVR_THREADS) s_xf[i] = pix_to_ndc(p.W - 1 - i, p.W, p.H);
  for (int i = threadIdx.x; i < p.H; i += MVR_THREADS) s_yf[i] = pix_to_ndc(p.H - 1 - i, p.H, p.W);
  __syncthreads();
  const Camera cam = load_camera(p.R, p.T, n);
  const int voff = p.vert_off[b];
  int n_straddle = 0;
  for (int f = fbeg + threadIdx.x; f < fend; f += MVR_THREADS) {
    const Face fc = load_face(p, cam, voff, __ldg(p.faces4 + f0 + f));
    int xl, xh, yl, yh;
    if (p.z_clip >= 0.f) {     // every face crossing z_clip is counted, visible or not (as the oracle does)
      const int nb = (fc.z0 < p.z_clip) + (fc.z1 < p.z_clip) + (fc.z2 < p.z_clip);
      n_straddle += (nb == 1 || nb == 2);
    }
    if (!face_tile_bbox(fc, p, 0, p.W - 1, 0, p.H - 1, s_xf, s_yf, xl, xh, yl, yh)) continue;
    const int tx0 = xl / TILE, tx1 = xh / TILE, ty0 = yl / TILE, ty1 = yh / TILE;
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&s_cnt[ty * p.tiles_x + tx], 1);
  }
  if (p.counters && n_straddle) atomicAdd((unsigned long long*)(p.counters + MVR_CNT_STRADDLE), (unsigned long long)n_straddle);
  __syncthreads();
  // exclusive scan of the tile histogram
  int carry = 0;
  for (int t0 = 0; t0 < p.n_tiles; t0 += MVR_THREADS) {
    const int t = t0 + threadIdx.x;
    const int v = t < p.n_tiles ? s_cnt[t] : 0;
    int tot;
    const int ex = block_excl_scan(v, s_warp, &tot);
    if (t < p.n_tiles) s_off[t] = carry + ex;
    carry += tot;
  }
  const int total = carry;
  if (threadIdx.x == 0) {
    int base = -1;
    if (total > 0) {
      const long long got = (long long)atomicAdd((unsigned long long*)p.pool_counter, (unsigned long long)total);
      base = (got + total <= p.pool_cap) ? (int)got : -1;
      if (p.counters) {
        atomicAdd((unsigned long long*)(p.counters + MVR_CNT_BIN_ENTRIES), (unsigned long long)total);
        if (base < 0) atomicAdd((unsigned long long*)(p.counters + MVR_CNT_BIN_OVERFLOW), 1ull);
      }
    }
    s_base = base;
  }
  __syncthreads();
  const int base = s_base;
  if (total == 0) {
    for (int t = threadIdx.x; t < p.n_tiles; t += MVR_THREADS) seg[t] = make_int2(0, 0);
    return;
  }
  if (base < 0) {
    // pool exhausted: every tile of this view scans the whole chunk unbinned (start = -1 marks it)
    for (int t = threadIdx.x; t < p.n_tiles; t += MVR_THREADS) seg[t] = make_int2(-1, fend - fbeg);
    return;
  }
  for (int t = threadIdx.x; t < p.n_tiles; t += MVR_THREADS) seg[t] = make_int2(base + s_off[t], s_cnt[t]);
  __syncthreads();
  // fill: identical arithmetic => identical tile rectangles
  for (int f = fbeg + threadIdx.x; f < fend; f += MVR_THREADS) {
    const Face fc = load_face(p, cam, voff, __ldg(p.faces4 + f0 + f));
    int xl, xh, yl, yh;
    if (!face_tile_bbox(fc, p, 0, p.W - 1, 0, p.H - 1, s_xf, s_yf, xl, xh, yl, yh)) continue;
    const int tx0 = xl / TILE, tx1 = xh / TILE, ty0 = yl / TILE, ty1 = yh / TILE;
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) {
        const int slot = atomicAdd(&s_off[ty * p.tiles_x + tx], 1);
        p.pool[base + slot] = f;
      }
  }
}

// ------------------------------------------------------------------------------------------------
// shading ([upstream] shading.py phong_shading, lighting.py diffuse/specular, blending.py hard_rgb_blend)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 interp(const float b[3], const float4 a0, const float4 a1, const float4 a2) {
  return make_float3((b[0] * a0.x + b[1] * a1.x) + b[2] * a2.x, (b[0] * a0.y + b[1] * a1.y) + b[2] * a2.y,
                     (b[0] * a0.z + b[1] * a1.z) + b[2] * a2.z);
}
__device__ __forceinline__ float inv_norm_clamped(float x, float y, float z, float eps) {
  const float n2 = fmaf(x, x, fmaf(y, y, z * z));
  const float n = sqrtf(n2);
  return 1.0f / fmaxf(n, eps);
}
__device__ __forceinline__ float pow64(float a) {
  a = a * a; a = a * a; a = a * a; a = a * a; a = a * a; a = a * a;
  return a;
}

struct ShadeCtx {
  float lx, ly, lz;   // normalised light direction
  float cx, cy, cz;   // camera centre
};

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
// fine pass
// ------------------------------------------------------------------------------------------------
// Shared-memory layout of one fine CTA (dynamic): every phase of a 256-entry chunk runs DENSE --
//   A  setup    thread per bin entry: project, cull, tile-clipped pixel bbox -> record + <= 8 sub-items
//   B  filter   thread per sub-item (a run of bbox pixels): edge-function sign test -> candidate queue
//   C  resolve  thread per candidate: exact barycentrics / depth -> 64-bit (z, face) min on the pixel key
// so a 3-pixel sliver and a tile-filling triangle cost their threads the same, and the IEEE divisions
// run only on full warps of pixels that are inside their face.
constexpr int REC_WORDS = 11;            // x0 y0 z0 x1 y1 z1 x2 y2 z2 fid rect
constexpr int ITEM_CAP = 2048;           // sub-items per chunk (typically 256 entries x 1-3)
constexpr int WCAP = 320;                // candidates per warp queue
constexpr int NWARPS = MVR_THREADS / 32;

struct FineSmem {
  unsigned long long* cur;     // [TILE_PIX]
  unsigned long long* prev;    // [TILE_PIX] (K > 1 only)
  float* rec;                  // [REC_WORDS][256]  SoA
  int* items;                  // [ITEM_CAP]  slot | start << 8 | count << 18
  int* cand;                   // [NWARPS][WCAP]  slot | pix << 8
  int* pref;                   // [MAX_CHUNKS + 1]
  int* segstart;               // [MAX_CHUNKS]
  float* xf;                   // [TILE]
  float* yf;                   // [TILE]
  int* counters;               // [0] items
  int* warp;                   // [8] scan scratch / per-warp candidate counts
};
__host__ __device__ inline size_t fine_smem_bytes(int K) {
  return (size_t)TILE_PIX * 8 * (K > 1 ? 2 : 1) + REC_WORDS * MVR_THREADS * 4 + ITEM_CAP * 4 + NWARPS * WCAP * 4 +
         (MAX_CHUNKS + 1 + MAX_CHUNKS) * 4 + 2 * TILE * 4 + 16 * 4;
}
__device__ __forceinline__ FineSmem carve_fine_smem(unsigned char* base, int K) {
  FineSmem s;
  s.cur = (unsigned long long*)base; base += TILE_PIX * 8;
  s.prev = (unsigned long long*)base; if (K > 1) base += TILE_PIX * 8;
  s.rec = (float*)base; base += REC_WORDS * MVR_THREADS * 4;
  s.items = (int*)base; base += ITEM_CAP * 4;
  s.cand = (int*)base; base += NWARPS * WCAP * 4;
  s.pref = (int*)base; base += (MAX_CHUNKS + 1) * 4;
  s.segstart = (int*)base; base += MAX_CHUNKS * 4;
  s.xf = (float*)base; base += TILE * 4;
  s.yf = (float*)base; base += TILE * 4;
  s.counters = (int*)base; base += 8 * 4;
  s.warp = (int*)base;
  return s;
}

__device__ __forceinline__ Face load_record(const float* rec, int slot) {
  Face f;
  f.x0 = rec[0 * MVR_THREADS + slot]; f.y0 = rec[1 * MVR_THREADS + slot]; f.z0 = rec[2 * MVR_THREADS + slot];
  f.x1 = rec[3 * MVR_THREADS + slot]; f.y1 = rec[4 * MVR_THREADS + slot]; f.z1 = rec[5 * MVR_THREADS + slot];
  f.x2 = rec[6 * MVR_THREADS + slot]; f.y2 = rec[7 * MVR_THREADS + slot]; f.z2 = rec[8 * MVR_THREADS + slot];
  return f;
}

// phase C body: exact test of one (face, pixel) candidate and the keyed min
__device__ __forceinline__ void resolve_candidate(const FineSmem& s, int slot, int pix, bool persp, bool peel) {
  const Face fc = load_record(s.rec, slot);
  const FaceEdges fe = face_edges(fc);
  float w[3], b[3], pz;
  if (!raster_test(fc, fe, persp, s.xf[pix & (TILE - 1)], s.yf[pix / TILE], w, b, pz)) return;
  const unsigned long long key = make_key(pz, __float_as_int(s.rec[9 * MVR_THREADS + slot]));
  if (peel && key <= s.prev[pix]) return;
  smem_key_min(&s.cur[pix], key);
}

__device__ __forceinline__ void fine_epilogue(const MeshParams& p, const FineSmem& s, const Camera& cam, const ShadeCtx& sc,
                                              int n, int k, int x0, int y0, int f0, int voff, bool persp,
                                              bool per_vertex_rgb, const float4 ucol, float bg0, float bg1, float bg2) {
  // The barycentrics are recomputed with the SAME exact operation sequence as the scatter: for sliver faces a
  // reciprocal-multiply shortcut moves them by far more than the 1e-5 image tolerance (error ~ ulp * |xy| / area).
  const int tid = threadIdx.x;
  for (int j = 0; j < TILE_PIX / MVR_THREADS; ++j) {
    const int pix = tid + j * MVR_THREADS;
    const int ly = pix / TILE, lx = pix % TILE;
    const int yi = y0 + ly, xi = x0 + lx;
    const unsigned long long key = s.cur[pix];
    if (p.K > 1) s.prev[pix] = key;   // EMPTY stays EMPTY: later layers find nothing
    if (yi >= p.H || xi >= p.W) continue;
    int fid = -1;
    float w[3] = {-1.f, -1.f, -1.f}, bb[3] = {-1.f, -1.f, -1.f}, pz = -1.f, dd = -1.f;
    float out[3] = {bg0, bg1, bg2};
    if (key != MVR_EMPTY_KEY) {
      fid = (int)(unsigned int)(key & 0xffffffffull);
      const int4 fi = __ldg(p.faces4 + f0 + fid);
      const float4 X0 = __ldg(p.verts4 + voff + fi.x), X1 = __ldg(p.verts4 + voff + fi.y), X2 = __ldg(p.verts4 + voff + fi.z);
      const float xf = s.xf[lx], yf = s.yf[ly];
      Face fc;
      project_vertex(cam, X0, p.k00, p.k11, fc.x0, fc.y0, fc.z0);
      project_vertex(cam, X1, p.k00, p.k11, fc.x1, fc.y1, fc.z1);
      project_vertex(cam, X2, p.k00, p.k11, fc.x2, fc.y2, fc.z2);
      const FaceEdges fe = face_edges(fc);
      raster_test(fc, fe, persp, xf, yf, w, bb, pz);
      pz = __uint_as_float((unsigned int)(key >> 32));
      if (p.dists) {
        const float e01 = point_line_dist2(xf, yf, fc.x0, fc.y0, fc.x1, fc.y1);
        const float e02 = point_line_dist2(xf, yf, fc.x0, fc.y0, fc.x2, fc.y2);
        const float e12 = point_line_dist2(xf, yf, fc.x1, fc.y1, fc.x2, fc.y2);
        dd = -fminf(fminf(e01, e02), e12);
      }
      if (k == 0) {
        const float4 N0 = __ldg(p.normals4 + voff + fi.x), N1 = __ldg(p.normals4 + voff + fi.y), N2 = __ldg(p.normals4 + voff + fi.z);
        float4 c0 = ucol, c1 = ucol, c2 = ucol;
        if (per_vertex_rgb) { c0 = __ldg(p.rgb4 + voff + fi.x); c1 = __ldg(p.rgb4 + voff + fi.y); c2 = __ldg(p.rgb4 + voff + fi.z); }
        phong_pixel(bb, X0, X1, X2, N0, N1, N2, c0, c1, c2, sc, out);
      }
    }
    const size_t po = (((size_t)n * p.H + yi) * p.W + xi) * p.K + k;
    p.pix_to_face[po] = fid;
    if (p.zbuf) p.zbuf[po] = pz;
    if (p.dists) p.dists[po] = dd;
    if (p.bary) { p.bary[3 * po] = bb[0]; p.bary[3 * po + 1] = bb[1]; p.bary[3 * po + 2] = bb[2]; }
    if (k == 0) {
      const size_t io = ((size_t)n * 3 * p.H + yi) * p.W + xi;
      const size_t plane = (size_t)p.H * p.W;
      p.images[io] = out[0]; p.images[io + plane] = out[1]; p.images[io + 2 * plane] = out[2];
    }
  }
}

__global__ void __launch_bounds__(MVR_THREADS, 4) mesh_fine_kernel(const MeshParams p) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const FineSmem s = carve_fine_smem(s_raw, p.K);
  const int tid = threadIdx.x;
  const int n = blockIdx.x / p.n_tiles, tile = blockIdx.x % p.n_tiles;
  const int b = n / p.M;
  const int x0 = (tile % p.tiles_x) * TILE, y0 = (tile / p.tiles_x) * TILE;
  const int x1 = min(x0 + TILE, p.W) - 1, y1 = min(y0 + TILE, p.H) - 1;   // inclusive
  const int f0 = p.face_off[b], F = p.face_off[b + 1] - f0;
  const int voff = p.vert_off[b];
  const int n_chunks = (F + p.fpc - 1) / p.fpc;
  const bool persp = p.flags & MVR_PERSPECTIVE_CORRECT;
  const Camera cam = load_camera(p.R, p.T, n);

  if (tid < TILE) {
    s.xf[tid] = pix_to_ndc(p.W - 1 - (x0 + tid), p.W, p.H);
    s.yf[tid] = pix_to_ndc(p.H - 1 - (y0 + tid), p.H, p.W);
  }
  if (tid == 0) s.counters[0] = 0;
  // bin segments of this (view, tile): one per face chunk
  int cnt = 0;
  if (tid < n_chunks) {
    const int2 sg = p.seg[((size_t)n * p.max_chunks + tid) * p.n_tiles + tile];
    cnt = sg.y; s.segstart[tid] = sg.x;
  }
  int total;
  const int ex = block_excl_scan(cnt, s.warp, &total);
  if (tid < n_chunks) s.pref[tid] = ex;
  if (tid == 0) s.pref[n_chunks] = total;
  __syncthreads();

  // light / camera for the epilogue
  ShadeCtx sc;
  {
    const float* Lp = p.light + (size_t)p.light_stride * n;
    const float lx = __ldg(Lp), ly = __ldg(Lp + 1), lz = __ldg(Lp + 2);
    const float il = inv_norm_clamped(lx, ly, lz, 1e-6f);
    sc.lx = lx * il; sc.ly = ly * il; sc.lz = lz * il;
    sc.cx = __ldg(p.Cc + 3 * (size_t)n); sc.cy = __ldg(p.Cc + 3 * (size_t)n + 1); sc.cz = __ldg(p.Cc + 3 * (size_t)n + 2);
  }
  const float bg0 = __ldg(p.bg_rgb), bg1 = __ldg(p.bg_rgb + 1), bg2 = __ldg(p.bg_rgb + 2);
  const bool per_vertex_rgb = p.flags & MVR_RGB_PER_ELEMENT;
  float4 ucol = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!per_vertex_rgb) ucol = make_float4(__ldg(p.obj_rgb), __ldg(p.obj_rgb + 1), __ldg(p.obj_rgb + 2), 0.f);

  for (int k = 0; k < p.K; ++k) {
    const bool peel = k > 0;
    for (int i = tid; i < TILE_PIX; i += MVR_THREADS) s.cur[i] = MVR_EMPTY_KEY;
    __syncthreads();
    for (int base = 0; base < total; base += MVR_THREADS) {
      // ---------------- phase A: setup ----------------
      const int i = base + tid;
      if (i < total) {
        int lo = 0, hi = n_chunks - 1;          // chunk of entry i: largest c with pref[c] <= i
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (s.pref[mid] <= i) lo = mid; else hi = mid - 1;
        }
        const int local = i - s.pref[lo];
        const int start = s.segstart[lo];
        const int fid = start < 0 ? lo * p.fpc + local : __ldg(p.pool + start + local);
        const Face fc = load_face(p, cam, voff, __ldg(p.faces4 + f0 + fid));
        int xl, xh, yl, yh;
        if (face_tile_bbox(fc, p, x0, x1, y0, y1, s.xf, s.yf, xl, xh, yl, yh)) {
          const int bw = xh - xl + 1, bh = yh - yl + 1, npx = bw * bh;
          s.rec[0 * MVR_THREADS + tid] = fc.x0; s.rec[1 * MVR_THREADS + tid] = fc.y0; s.rec[2 * MVR_THREADS + tid] = fc.z0;
          s.rec[3 * MVR_THREADS + tid] = fc.x1; s.rec[4 * MVR_THREADS + tid] = fc.y1; s.rec[5 * MVR_THREADS + tid] = fc.z1;
          s.rec[6 * MVR_THREADS + tid] = fc.x2; s.rec[7 * MVR_THREADS + tid] = fc.y2; s.rec[8 * MVR_THREADS + tid] = fc.z2;
          s.rec[9 * MVR_THREADS + tid] = __int_as_float(fid);
          s.rec[10 * MVR_THREADS + tid] = __int_as_float((xl - x0) | ((yl - y0) << 5) | ((bw - 1) << 10) | ((bh - 1) << 15));
          // runs of G pixels: 8 for ordinary faces, up to 32 for tile-sized ones (<= 32 runs per face)
          const int G = max(8, (npx + 31) >> 5);
          const int nsub = (npx + G - 1) / G;
          const int at = atomicAdd(&s.counters[0], nsub);
          if (at + nsub <= ITEM_CAP) {
            for (int q = 0; q < nsub; ++q) s.items[at + q] = tid | ((q * G) << 8) | (min(G, npx - q * G) << 18);
          } else {
            // item queue exhausted (a chunk of tile-sized faces): this thread walks its own bbox
            for (int q = at; q < ITEM_CAP; ++q) s.items[q] = 0;      // a straddling reservation leaves no garbage
            const FaceEdges fe = face_edges(fc);
            for (int yy = yl; yy <= yh; ++yy)
              for (int xx = xl; xx <= xh; ++xx) {
                float w[3], bq[3], pz;
                if (!raster_test(fc, fe, persp, s.xf[xx - x0], s.yf[yy - y0], w, bq, pz)) continue;
                const unsigned long long key = make_key(pz, fid);
                const int pix = (yy - y0) * TILE + (xx - x0);
                if (peel && key <= s.prev[pix]) continue;
                smem_key_min(&s.cur[pix], key);
              }
          }
        }
      }
      __syncthreads();
      // ---------------- phase B: sign filter over bbox pixels, per-warp candidate queues ----------------
      const int n_items = min(s.counters[0], ITEM_CAP);
      const int warp = tid >> 5, lane = tid & 31;
      int wcnt = 0;                              // warp-uniform: candidates queued by this warp
      int* wq = s.cand + warp * WCAP;
      for (int j0 = warp * 32; j0 < n_items; j0 += MVR_THREADS) {      // warp-uniform trip count
        const int j = j0 + lane;
        int slot = 0, count = 0, row = 0, col = 0, lxl = 0, lyl = 0, bw = 1;
        float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f, cx = 0.f, cy = 0.f;
        unsigned int zmin_bits = 0u;
        if (j < n_items) {
          const int it = s.items[j];
          slot = it & 255; count = it >> 18;
          const int start = (it >> 8) & 1023;
          ax = s.rec[0 * MVR_THREADS + slot]; ay = s.rec[1 * MVR_THREADS + slot];
          bx = s.rec[3 * MVR_THREADS + slot]; by = s.rec[4 * MVR_THREADS + slot];
          cx = s.rec[6 * MVR_THREADS + slot]; cy = s.rec[7 * MVR_THREADS + slot];
          const int rect = __float_as_int(s.rec[10 * MVR_THREADS + slot]);
          lxl = rect & 31; lyl = (rect >> 5) & 31; bw = ((rect >> 10) & 31) + 1;
          row = (int)__fdividef((float)start + 0.5f, (float)bw);     // small integers: exact
          col = start - row * bw;
          // early depth reject: pz is a convex combination of the vertex depths up to a few ulp (perspective-
          // corrected barycentrics sum to 1 unless their 1e-8 denominator clamp acts, which needs z ~ 1e-4;
          // plain barycentrics sum to area/(area+1e-8), so the shortcut is not used for them), hence a face
          // whose nearest vertex is clearly behind the pixel's current winner cannot produce a smaller key
          const float zmin = fminf(fminf(s.rec[2 * MVR_THREADS + slot], s.rec[5 * MVR_THREADS + slot]), s.rec[8 * MVR_THREADS + slot]);
          if (persp && zmin > 1e-3f) zmin_bits = __float_as_uint(zmin * 0.999999f);
        }
        const float A0 = cy - by, B0 = cx - bx, A1 = ay - cy, B1 = ax - cx, A2 = by - ay, B2 = bx - ax;
        const float area_p = ((cx - ax) * A2 - (cy - ay) * B2) + MVR_K_EPS;
        const int maxc = __reduce_max_sync(0xffffffffu, count);
        for (int c = 0; c < maxc; ++c) {
          const int lx = lxl + col, ly = lyl + row;
          const int pix = ly * TILE + lx;
          bool pass = c < count;
          if (pass) {
            const float xf = s.xf[lx], yf = s.yf[ly];
            const float e0 = (xf - bx) * A0 - (yf - by) * B0;
            const float e1 = (xf - cx) * A1 - (yf - cy) * B1;
            const float e2 = (xf - ax) * A2 - (yf - ay) * B2;
            pass = area_p > 0.f ? (e0 > 0.f && e1 > 0.f && e2 > 0.f) : (e0 < 0.f && e1 < 0.f && e2 < 0.f);
            if (pass) pass = zmin_bits <= ((const volatile unsigned int*)s.cur)[2 * pix + 1];
            if (++col == bw) { col = 0; ++row; }
          }
          const unsigned int m = __ballot_sync(0xffffffffu, pass);
          if (pass) {
            const int at = wcnt + __popc(m & ((1u << lane) - 1u));
            if (at < WCAP) wq[at] = slot | (pix << 8);
            else resolve_candidate(s, slot, pix, persp, peel);       // queue full: resolve in place
          }
          wcnt += __popc(m);
        }
      }
      if (lane == 0) s.warp[warp] = min(wcnt, WCAP);
      __syncthreads();
      // ---------------- phase C: exact resolve, candidates of all warps spread over all threads ----------------
      if (tid == 0) s.counters[0] = 0;          // every thread read the item count before the barrier above
      {
        int pre[NWARPS + 1];
        pre[0] = 0;
#pragma unroll
        for (int wi = 0; wi < NWARPS; ++wi) pre[wi + 1] = pre[wi] + s.warp[wi];
        for (int j = tid; j < pre[NWARPS]; j += MVR_THREADS) {
          int wi = 0;
#pragma unroll
          for (int q = 1; q < NWARPS; ++q) wi += (j >= pre[q]);
          const int cd = s.cand[wi * WCAP + (j - pre[wi])];
          resolve_candidate(s, cd & 255, cd >> 8, persp, peel);
        }
      }
      __syncthreads();
    }
    __syncthreads();
    fine_epilogue(p, s, cam, sc, n, k, x0, y0, f0, voff, persp, per_vertex_rgb, ucol, bg0, bg1, bg2);
    __syncthreads();
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
  int B, M, H, W, K, flags, n_tiles, tiles_x;
  const int* pix_to_face; const float* grad_images;
  float* partials;       // (N, n_tiles, 16)
  float* grad_verts; float* grad_normals;
};

__device__ __forceinline__ void normalize_bwd3(float vx, float vy, float vz, float eps, float gx, float gy, float gz,
                                               float& ox, float& oy, float& oz) {
  const float n = sqrtf(fmaf(vx, vx, fmaf(vy, vy, vz * vz)));
  if (n > eps) {
    const float inv = 1.f / n;
    const float ux = vx * inv, uy = vy * inv, uz = vz * inv;
    const float d = fmaf(ux, gx, fmaf(uy, gy, uz * gz));
    ox = (gx - ux * d) * inv; oy = (gy - uy * d) * inv; oz = (gz - uz * d) * inv;
  } else {
    const float inv = 1.f / eps;
    ox = gx * inv; oy = gy * inv; oz = gz * inv;
  }
}

__global__ void __launch_bounds__(MVR_THREADS) mesh_backward_kernel(const MeshBwdParams p) {
  __shared__ float s_red[8 * BWD_VALS];
  __shared__ int s_any;
  const int tid = threadIdx.x;
  const int n = blockIdx.x / p.n_tiles, tile = blockIdx.x % p.n_tiles;
  const int b = n / p.M;
  const int x0 = (tile % p.tiles_x) * TILE, y0 = (tile / p.tiles_x) * TILE;
  const int f0 = p.face_off[b], voff = p.vert_off[b];
  const bool persp = p.flags & MVR_PERSPECTIVE_CORRECT;
  const bool per_vertex_rgb = p.flags & MVR_RGB_PER_ELEMENT;
  if (tid == 0) s_any = 0;
  __syncthreads();
  float acc[BWD_VALS];
#pragma unroll
  for (int i = 0; i < BWD_VALS; ++i) acc[i] = 0.f;
  const size_t plane = (size_t)p.H * p.W;
  bool any = false;
  Camera cam; ShadeCtx sc; float4 ucol = make_float4(0.f, 0.f, 0.f, 0.f);
  bool ctx_loaded = false;
  for (int j = 0; j < TILE_PIX / MVR_THREADS; ++j) {
    const int pix = tid + j * MVR_THREADS;
    const int ly = pix / TILE, lx = pix % TILE;
    const int yi = y0 + ly, xi = x0 + lx;
    if (yi >= p.H || xi >= p.W) continue;
    const int fid = __ldg(p.pix_to_face + (((size_t)n * p.H + yi) * p.W + xi) * p.K);
    if (fid < 0) continue;
    const size_t io = ((size_t)n * 3 * p.H + yi) * p.W + xi;
    const float g0 = __ldg(p.grad_images + io), g1 = __ldg(p.grad_images + io + plane), g2 = __ldg(p.grad_images + io + 2 * plane);
    if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;
    any = true;
    if (!ctx_loaded) {
      cam = load_camera(p.R, p.T, n);
      const float* L = p.light + (size_t)p.light_stride * n;
      const float lx_ = __ldg(L), ly_ = __ldg(L + 1), lz_ = __ldg(L + 2);
      const float il = inv_norm_clamped(lx_, ly_, lz_, 1e-6f);
      sc.lx = lx_ * il; sc.ly = ly_ * il; sc.lz = lz_ * il;
      sc.cx = __ldg(p.Cc + 3 * (size_t)n); sc.cy = __ldg(p.Cc + 3 * (size_t)n + 1); sc.cz = __ldg(p.Cc + 3 * (size_t)n + 2);
      if (!per_vertex_rgb) ucol = make_float4(__ldg(p.obj_rgb), __ldg(p.obj_rgb + 1), __ldg(p.obj_rgb + 2), 0.f);
      ctx_loaded = true;
    }
    const int4 fi = __ldg(p.faces4 + f0 + fid);
    const float4 X0 = __ldg(p.verts4 + voff + fi.x), X1 = __ldg(p.verts4 + voff + fi.y), X2 = __ldg(p.verts4 + voff + fi.z);
    const float4 N0 = __ldg(p.normals4 + voff + fi.x), N1 = __ldg(p.normals4 + voff + fi.y), N2 = __ldg(p.normals4 + voff + fi.z);
    float4 c0 = ucol, c1 = ucol, c2 = ucol;
    if (per_vertex_rgb) { c0 = __ldg(p.rgb4 + voff + fi.x); c1 = __ldg(p.rgb4 + voff + fi.y); c2 = __ldg(p.rgb4 + voff + fi.z); }
    // ---- forward recompute ----
    float pv[3][3];   // view-space vertices
    world_to_view(cam, X0.x, X0.y, X0.z, pv[0][0], pv[0][1], pv[0][2]);
    world_to_view(cam, X1.x, X1.y, X1.z, pv[1][0], pv[1][1], pv[1][2]);
    world_to_view(cam, X2.x, X2.y, X2.z, pv[2][0], pv[2][1], pv[2][2]);
    Face fc;
    fc.x0 = (pv[0][0] * p.k00) / pv[0][2]; fc.y0 = (pv[0][1] * p.k11) / pv[0][2]; fc.z0 = pv[0][2];
    fc.x1 = (pv[1][0] * p.k00) / pv[1][2]; fc.y1 = (pv[1][1] * p.k11) / pv[1][2]; fc.z1 = pv[1][2];
    fc.x2 = (pv[2][0] * p.k00) / pv[2][2]; fc.y2 = (pv[2][1] * p.k11) / pv[2][2]; fc.z2 = pv[2][2];
    const FaceEdges fe = face_edges(fc);
    const float xf = pix_to_ndc(p.W - 1 - xi, p.W, p.H), yf = pix_to_ndc(p.H - 1 - yi, p.H, p.W);
    const float e0 = (xf - fc.x1) * fe.A0 - (yf - fc.y1) * fe.B0;
    const float e1 = (xf - fc.x2) * fe.A1 - (yf - fc.y2) * fe.B1;
    const float e2 = (xf - fc.x0) * fe.A2 - (yf - fc.y0) * fe.B2;
    const float inv_area = 1.f / fe.area_p;
    const float w0 = e0 * inv_area, w1 = e1 * inv_area, w2 = e2 * inv_area;
    float bb[3] = {w0, w1, w2};
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, denom = 1.f;
    if (persp) {
      t0 = w0 * fc.z1 * fc.z2; t1 = w1 * fc.z0 * fc.z2; t2 = w2 * fc.z0 * fc.z1;
      denom = fmaxf(t0 + t1 + t2, MVR_K_EPS);
      const float id = 1.f / denom;
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
    float a2 = alpha * alpha, a4 = a2 * a2, a8 = a4 * a4, a16 = a8 * a8, a32 = a16 * a16;
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
      const float id = 1.f / denom;
      // b = t / sum(t) is invariant to a common shift of d/db (its Jacobian annihilates constants), so
      // remove k = sum(b_i gb_i) first: the d/d denom term then vanishes identically instead of cancelling
      // O(1/area) terms in fp32.  Same gradient in exact arithmetic; not valid when denom was clamped.
      const bool clamped = (t0 + t1 + t2) < MVR_K_EPS;
      if (!clamped) {
        const float k = fmaf(bb[0], gb0, fmaf(bb[1], gb1, bb[2] * gb2));
        gb0 -= k; gb1 -= k; gb2 -= k;
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
    // ---- projection backward + X R + T backward ----
    const float gxn[3] = {gx0, gx1, gx2}, gyn[3] = {gy0, gy1, gy2}, gzn[3] = {dz0, dz1, dz2};
    const float4 Xs[3] = {X0, X1, X2};
    const int vid[3] = {fi.x, fi.y, fi.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float iz = 1.f / pv[i][2];
      const float gpx = gxn[i] * p.k00 * iz;
      const float gpy = gyn[i] * p.k11 * iz;
      const float gpz = gzn[i] - (gxn[i] * (pv[i][0] * p.k00) + gyn[i] * (pv[i][1] * p.k11)) * iz * iz;
      acc[0] = fmaf(Xs[i].x, gpx, acc[0]); acc[1] = fmaf(Xs[i].x, gpy, acc[1]); acc[2] = fmaf(Xs[i].x, gpz, acc[2]);
      acc[3] = fmaf(Xs[i].y, gpx, acc[3]); acc[4] = fmaf(Xs[i].y, gpy, acc[4]); acc[5] = fmaf(Xs[i].y, gpz, acc[5]);
      acc[6] = fmaf(Xs[i].z, gpx, acc[6]); acc[7] = fmaf(Xs[i].z, gpy, acc[7]); acc[8] = fmaf(Xs[i].z, gpz, acc[8]);
      acc[9] += gpx; acc[10] += gpy; acc[11] += gpz;
      if (p.grad_verts) {
        float* o = p.grad_verts + 3 * (size_t)(voff + vid[i]);
        atomicAdd(o + 0, fmaf(cam.r[0], gpx, fmaf(cam.r[1], gpy, cam.r[2] * gpz)) - bb[i] * gvx);
        atomicAdd(o + 1, fmaf(cam.r[3], gpx, fmaf(cam.r[4], gpy, cam.r[5] * gpz)) - bb[i] * gvy);
        atomicAdd(o + 2, fmaf(cam.r[6], gpx, fmaf(cam.r[7], gpy, cam.r[8] * gpz)) - bb[i] * gvz);
      }
      if (p.grad_normals) {
        float* o = p.grad_normals + 3 * (size_t)(voff + vid[i]);
        atomicAdd(o + 0, bb[i] * gNx); atomicAdd(o + 1, bb[i] * gNy); atomicAdd(o + 2, bb[i] * gNz);
      }
    }
  }
  if (any) s_any = 1;   // benign race: all writers store 1
  __syncthreads();
  float* out = p.partials + ((size_t)n * p.n_tiles + tile) * 16;
  if (!s_any) {   // uniform: background-only tile
    if (tid < 16) out[tid] = 0.f;
    return;
  }
  block_sum<BWD_VALS>(acc, s_red);
  if (tid < 16) out[tid] = tid < BWD_VALS ? s_red[tid] : 0.f;
}

// fixed-order sum of the per-tile partials: one warp per view -> gR, gT, gC
__global__ void mesh_backward_reduce_kernel(const float* __restrict__ partials, int N, int n_tiles,
                                            float* __restrict__ gR, float* __restrict__ gT, float* __restrict__ gC) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  // lane l < 16 owns value l of even tiles, lane l >= 16 value l-16 of odd tiles
  const int v = lane & 15, par = lane >> 4;
  float s = 0.f;
  for (int t = par; t < n_tiles; t += 2) s += partials[((size_t)n * n_tiles + t) * 16 + v];
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

extern "C" size_t mvr_mesh_workspace_bytes(int B, int M, int H, int W, int64_t total_faces, int max_faces) {
  if (B < 0 || M < 0 || H <= 0 || W <= 0 || total_faces < 0 || max_faces < 0) return 0;
  return ws_layout(B, M, H, W, total_faces, max_faces).total;
}

static int check_mesh_common(const char* who, int B, int M, int H, int W, int K, int64_t tv, int64_t tf) {
  if (B < 0 || M < 0 || tv < 0 || tf < 0) { set_error("%s: negative size", who); return -1; }
  if (H <= 0 || W <= 0 || H > 4096 || W > 4096) { set_error("%s: image size %dx%d outside [1, 4096]", who, H, W); return -2; }
  if (K < 1 || K > 64) { set_error("%s: faces_per_pixel %d outside [1, 64]", who, K); return -3; }
  if ((int64_t)B * M > 0x7fffffffLL / (((W + 31) / 32) * ((H + 31) / 32) + 256)) { set_error("%s: too many views", who); return -4; }
  return 0;
}

extern "C" int mvr_mesh_forward(const void* geometry, const int* vert_off, const int* face_off, int B, int M,
                                int64_t total_verts, int64_t total_faces, int max_faces, const float* R,
                                const float* T, const float* Cc, const float* light, int light_stride,
                                const float* obj_rgb, const float* bg_rgb, float k00, float k11, float z_clip,
                                int H, int W, int K, int flags, float* images, int* pix_to_face, float* zbuf,
                                float* bary, float* dists, int64_t* counters, void* workspace,
                                size_t workspace_bytes, void* stream) {
  int rc = check_mesh_common("mvr_mesh_forward", B, M, H, W, K, total_verts, total_faces);
  if (rc) return rc;
  const int64_t N = (int64_t)B * M;
  if (N == 0) return 0;
  if (!geometry || !vert_off || !face_off || !R || !T || !Cc || !light || !bg_rgb || !images || !pix_to_face || !workspace) {
    set_error("mvr_mesh_forward: null pointer"); return -5;
  }
  if (!(flags & MVR_RGB_PER_ELEMENT) && !obj_rgb) { set_error("mvr_mesh_forward: obj_rgb is NULL and the geometry has no per-vertex colours"); return -6; }
  const WsLayout w = ws_layout(B, M, H, W, total_faces, max_faces);
  if (workspace_bytes < w.total) { set_error("mvr_mesh_forward: workspace too small (%zu < %zu)", workspace_bytes, w.total); return -7; }
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
  p.n_tiles = w.n_tiles; p.tiles_x = w.tiles_x; p.max_chunks = w.max_chunks; p.fpc = w.fpc;
  p.pool_cap = (flags & MVR_TEST_TINY_POOL) ? 64 : w.pool_cap;
  p.pool_counter = (int*)(wb + w.counter); p.seg = (int2*)(wb + w.seg); p.pool = (int*)(wb + w.pool);
  p.images = images; p.pix_to_face = pix_to_face; p.zbuf = zbuf; p.bary = bary; p.dists = dists;
  p.counters = (long long*)counters;
  cudaError_t e = cudaMemsetAsync(wb + w.counter, 0, 256, st);
  if (e != cudaSuccess) { set_error("mvr_mesh_forward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  const size_t bin_smem = (2 * (size_t)w.n_tiles + W + H) * sizeof(int);
  if (bin_smem > 48 * 1024) {
    e = cudaFuncSetAttribute(mesh_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem);
    if (e != cudaSuccess) { set_error("mvr_mesh_forward: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  MVR_LAUNCH(mesh_bin_kernel, (unsigned)(N * w.max_chunks), MVR_THREADS, bin_smem, st, p);
  rc = check_launch("mesh_bin_kernel");
  if (rc) return rc;
  MVR_LAUNCH(mesh_fine_kernel, (unsigned)(N * w.n_tiles), MVR_THREADS, fine_smem_bytes(K), st, p);
  return check_launch("mesh_fine_kernel");
}

extern "C" int mvr_mesh_backward(const void* geometry, const int* vert_off, const int* face_off, int B, int M,
                                 int64_t total_verts, int64_t total_faces, const float* R, const float* T,
                                 const float* Cc, const float* light, int light_stride, const float* obj_rgb,
                                 float k00, float k11, int H, int W, int K, int flags, const int* pix_to_face,
                                 const float* grad_images, float* gR, float* gT, float* gC, float* grad_verts,
                                 float* grad_normals, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_mesh_common("mvr_mesh_backward", B, M, H, W, K, total_verts, total_faces);
  if (rc) return rc;
  const int64_t N = (int64_t)B * M;
  if (N == 0) return 0;
  if (!geometry || !vert_off || !face_off || !R || !T || !Cc || !light || !pix_to_face || !grad_images || !gR || !gT || !gC || !workspace) {
    set_error("mvr_mesh_backward: null pointer"); return -5;
  }
  if (!(flags & MVR_RGB_PER_ELEMENT) && !obj_rgb) { set_error("mvr_mesh_backward: obj_rgb is NULL"); return -6; }
  const int tiles_x = (W + TILE - 1) / TILE, n_tiles = tiles_x * ((H + TILE - 1) / TILE);
  const size_t need = (size_t)N * n_tiles * 16 * sizeof(float);
  if (workspace_bytes < need) { set_error("mvr_mesh_backward: workspace too small (%zu < %zu)", workspace_bytes, need); return -7; }
  const GeomLayout g = geom_layout(total_verts, total_faces);
  const char* gb = (const char*)geometry;
  cudaStream_t st = (cudaStream_t)stream;
  MeshBwdParams p;
  p.verts4 = (const float4*)(gb + g.verts4); p.normals4 = (const float4*)(gb + g.normals4);
  p.rgb4 = (const float4*)(gb + g.rgb4); p.faces4 = (const int4*)(gb + g.faces4);
  p.vert_off = vert_off; p.face_off = face_off;
  p.R = R; p.T = T; p.Cc = Cc; p.light = light; p.light_stride = light_stride; p.obj_rgb = obj_rgb;
  p.k00 = k00; p.k11 = k11;
  p.B = B; p.M = M; p.H = H; p.W = W; p.K = K; p.flags = flags; p.n_tiles = n_tiles; p.tiles_x = tiles_x;
  p.pix_to_face = pix_to_face; p.grad_images = grad_images;
  p.partials = (float*)workspace; p.grad_verts = grad_verts; p.grad_normals = grad_normals;
  MVR_LAUNCH(mesh_backward_kernel, (unsigned)(N * n_tiles), MVR_THREADS, 0, st, p);
  rc = check_launch("mesh_backward_kernel");
  if (rc) return rc;
  const int wpb = 8;
  MVR_LAUNCH(mesh_backward_reduce_kernel, (unsigned)((N + wpb - 1) / wpb), wpb * 32, 0, st, (const float*)workspace, (int)N, n_tiles, gR, gT, gC);
  return check_launch("mesh_backward_reduce_kernel");
}
