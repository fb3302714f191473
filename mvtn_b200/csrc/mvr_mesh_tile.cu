// mvr_mesh_tile.cu -- tile-binned forward of the mesh path for faces_per_pixel == 1 (renderer.py:89-113): a coarse binning
// pass followed by ONE fine kernel that rasterizes, resolves depth in shared memory and shades (north_star (2) + (3)).
//
//   bin   : mesh_bin_kernel<count> -- thread per (view, face): gather the three projected vertices, the rasterizer's face-level
//           rejections, a conservative pixel bbox -> range of 32x32-pixel tiles -> one 4-byte bin code per face instance
//           and per-(view, tile) counts (warp-aggregated atomics: neighbouring faces fall into the same tile).  Faces that
//           span more than 2 x 2 tiles or cross the near clip plane go to a per-view "big" list instead.
//           mesh_bin_scan_kernel -- CTA per view: exclusive scan of the (padded) counts -> tile offsets; decides which
//           facing (sign of the screen-space area) is nearer on average, so that the fine pass can visit it FIRST.
//           mesh_bin_kernel<fill> -- reads the bin codes back and writes the face ids into the tile lists, the nearer
//           facing from the front of a tile's segment and the other one from its back.
//   fine  : mesh_tile_kernel -- CTA per (view, tile).  The tile's face-id list is CONTIGUOUS in global memory and is staged
//           into shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier, double-buffered: round r + 1 travels
//           while round r is rasterized).  Per round of 256 faces the same three phases as mesh_scatter_kernel
//           (A setup clipped to the tile, B FMA-form edge filter over runs of bbox pixels, C exact IEEE resolve), but the
//           64-bit (z, face) keys of the tile's 1024 pixels live in SHARED memory: the early depth test reads the
//           current key at no cost, and because the nearer facing comes first the hidden half of a closed mesh is rejected
//           before its six IEEE divisions.  After the last round the same CTA shades its pixels straight from the shared keys:
//           no global key plane, no memset, no re-arm pass, no DRAM trip between rasterizer and shader.
//   Exactness: the fine pass evaluates exactly the oracle's sequence (raster_test) on exactly the pixels of the exact
//   bbox, and keeps the lexicographic (z, face) minimum -- order-independent, so the nondeterministic list order of the
//   binning atomics never shows.  The conservative binning only ever ADDS faces to a list.
#include "mvr_mesh_fwd.cuh"

namespace mvr {

constexpr unsigned BIN_VALID = 1u << 29, BIN_POS = 1u << 30, BIN_BIG = 1u << 31;
constexpr int T_ITEM_CAP = 1024;      // sub-items per round in the tile kernel
constexpr int T_WCAP = 256;           // candidates per warp queue
constexpr int FID_BUF = MVR_THREADS + 8;      // a round's ids start at a 16-byte aligned source address: up to 3 ids of slack

struct TileWs {
  unsigned int* codes; int* lists; int* big_list; int* tile_cnt; int* big_cnt; float* vote; int* tile_off; int* cursors; int* front_sign;
  int tiles_x, tiles;
};

// grid: x = 256-face chunks of the largest object, y = view m, z = object b
template <bool FILL>
__global__ void __launch_bounds__(MVR_THREADS) mesh_bin_kernel(const MeshParams p, const TileWs tw) {
  const int m = blockIdx.y, b = blockIdx.z, n = b * p.M + m, lane = threadIdx.x & 31;
  const int f0 = p.face_off[b], F = p.face_off[b + 1] - f0;
  const int f = blockIdx.x * MVR_THREADS + threadIdx.x;
  if (blockIdx.x * MVR_THREADS >= F) return;
  const size_t fb = (size_t)p.M * f0 + (size_t)m * F;      // first face instance of this view
  unsigned int code = 0;
  if (!FILL) {
    float zsum = 0.f;
    if (f < F) {
      const int voff = p.vert_off[b], V = p.vert_off[b + 1] - voff;
      const Face fc = gather_face(p.pv + (size_t)p.M * voff + (size_t)m * V, __ldg(p.faces4 + f0 + f));
      if (face_straddles(fc, p.z_clip)) {
        code = BIN_VALID | BIN_BIG | (1u << 28);      // bit 28: straddler (counted, clipped by every tile CTA of the view)
      } else {
        // the rasterizer's face-level rejections (face_pixel_bbox) ...
        const float zmin = fminf(fminf(fc.z0, fc.z1), fc.z2);
        const float area = (fc.x0 - fc.x1) * (fc.y2 - fc.y1) - (fc.y0 - fc.y1) * (fc.x2 - fc.x1);
        bool ok = !(p.z_clip >= 0.f && fc.z0 < p.z_clip && fc.z1 < p.z_clip && fc.z2 < p.z_clip) && !(zmin < MVR_K_EPS);
        ok = ok && !((p.flags & MVR_CULL_BACKFACES) && area < 0.f) && !(area <= MVR_K_EPS && area >= -1.0f * MVR_K_EPS);
        int xl, xh, yl, yh;
        // ... and a conservative pixel bbox (the exact one is the fine pass's job)
        ok = ok && pixel_range_conservative(fminf(fminf(fc.x0, fc.x1), fc.x2), fmaxf(fmaxf(fc.x0, fc.x1), fc.x2), p.W, p.jx_scale, p.jx_off, xl, xh);
        ok = ok && pixel_range_conservative(fminf(fminf(fc.y0, fc.y1), fc.y2), fmaxf(fmaxf(fc.y0, fc.y1), fc.y2), p.H, p.jy_scale, p.jy_off, yl, yh);
        if (ok) {
          const int tx0 = xl >> 5, tx1 = xh >> 5, ty0 = yl >> 5, ty1 = yh >> 5;
          const bool pos = area > 0.f;
          zsum = pos ? zmin : -zmin;
          code = BIN_VALID | (pos ? BIN_POS : 0u);
          if (tx1 - tx0 > 1 || ty1 - ty0 > 1) code |= BIN_BIG;
          else code |= (unsigned)tx0 | ((unsigned)ty0 << 7) | ((unsigned)(tx1 - tx0) << 14) | ((unsigned)(ty1 - ty0) << 15);
        }
      }
      tw.codes[fb + f] = code;
    }
    // which facing is nearer?  per-view sums of the nearest vertex depth of either facing, from a SAMPLE of the faces (warp 0
    // of every fourth CTA: same-address atomics serialise in L2, and an estimate is all the ordering heuristic needs)
    if (threadIdx.x < 32 && (blockIdx.x & 3) == 0) {
      const float zp = warp_sum(zsum > 0.f ? zsum : 0.f), zn = warp_sum(zsum < 0.f ? -zsum : 0.f);
      const unsigned mp = __ballot_sync(0xffffffffu, zsum > 0.f), mn = __ballot_sync(0xffffffffu, zsum < 0.f);
      if (lane == 0) {
        if (mp) { atomicAdd(tw.vote + 4 * n, zp); atomicAdd(tw.vote + 4 * n + 1, (float)__popc(mp)); }
        if (mn) { atomicAdd(tw.vote + 4 * n + 2, zn); atomicAdd(tw.vote + 4 * n + 3, (float)__popc(mn)); }
      }
    }
    if (p.counters) {
      const unsigned ms = __ballot_sync(0xffffffffu, (code >> 28) & 1u), mb = __ballot_sync(0xffffffffu, (code & BIN_BIG) && !((code >> 28) & 1u));
      if (lane == 0 && ms) atomicAdd((unsigned long long*)(p.counters + MVR_CNT_STRADDLE), (unsigned long long)__popc(ms));
      if (lane == 0 && mb) atomicAdd((unsigned long long*)(p.counters + MVR_CNT_BIG_FACES), (unsigned long long)__popc(mb));
    }
  } else if (f < F) {
    code = tw.codes[fb + f];
  }
  const bool valid = code & BIN_VALID, big = code & BIN_BIG;
  if (big) {
    if (!FILL) tw.big_list[fb + atomicAdd(tw.big_cnt + n, 1)] = f;
    code = 0;
  }
  const int T = tw.tiles;
  const int tx0 = code & 127, ty0 = (code >> 7) & 127, ntx = (code >> 14) & 1, nty = (code >> 15) & 1;
  const size_t tbase = (size_t)n * T;
  bool back = false;
  if (FILL) back = (int)((code >> 30) & 1u) != __ldg(tw.front_sign + n);
  // primary tile: warp-aggregated (faces adjacent in the index are adjacent on the surface -> mostly the same tile)
  {
    const int t = (valid && !big) ? ty0 * tw.tiles_x + tx0 : -1;
    const int akey = FILL ? (t < 0 ? -1 : 2 * t + (int)back) : t;
    const unsigned grp = __match_any_sync(0xffffffffu, akey);
    const int leader = __ffs(grp) - 1, rank = __popc(grp & ((1u << lane) - 1u));
    int base = 0;
    if (t >= 0 && lane == leader) base = atomicAdd(FILL ? tw.cursors + 2 * (tbase + t) + (int)back : tw.tile_cnt + tbase + t, __popc(grp));
    base = __shfl_sync(grp, base, leader);
    if (FILL && t >= 0) {
      const int off = tw.tile_off[tbase + t], cnt = tw.tile_cnt[tbase + t];
      tw.lists[4 * fb + 4 * tbase + (back ? off + cnt - 1 - (base + rank) : off + base + rank)] = f;
    }
  }
  // the (rare) other tiles of a face that straddles a tile border: plain atomics
  if (valid && !big && (ntx | nty)) {
    for (int dy = 0; dy <= nty; ++dy)
      for (int dx = 0; dx <= ntx; ++dx) {
        if (!(dx | dy)) continue;
        const int t = (ty0 + dy) * tw.tiles_x + tx0 + dx;
        if (!FILL) atomicAdd(tw.tile_cnt + tbase + t, 1);
        else {
          const int pos = atomicAdd(tw.cursors + 2 * (tbase + t) + (int)back, 1);
          const int off = tw.tile_off[tbase + t], cnt = tw.tile_cnt[tbase + t];
          tw.lists[4 * fb + 4 * tbase + (back ? off + cnt - 1 - pos : off + pos)] = f;
        }
      }
  }
}

// grid: one CTA per view.  tile_off = exclusive scan of the counts rounded up to 4 ids (16-byte aligned list starts for the
// bulk copies), cursors = 0, front_sign = the facing whose faces are nearer on average.
__global__ void __launch_bounds__(MVR_THREADS) mesh_bin_scan_kernel(const TileWs tw) {
  __shared__ int s_w[NWARPS];
  __shared__ int s_carry;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = tw.tiles;
  if (tid == 0) {
    s_carry = 0;
    const float zp = tw.vote[4 * n], cp = tw.vote[4 * n + 1], zn = tw.vote[4 * n + 2], cn = tw.vote[4 * n + 3];
    tw.front_sign[n] = (cn == 0.f || (cp > 0.f && zp * cn <= zn * cp)) ? 1 : 0;      // mean depth of positive-area faces <= negative
  }
  __syncthreads();
  for (int base = 0; base < T; base += MVR_THREADS) {
    const int t = base + tid;
    const int c4 = t < T ? (tw.tile_cnt[(size_t)n * T + t] + 3) & ~3 : 0;
    int incl = c4;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < NWARPS; ++q) { const int v = s_w[q]; if (q < warp) woff += v; tot += v; }
    const int carry = s_carry;
    if (t < T) {
      tw.tile_off[(size_t)n * T + t] = carry + woff + incl - c4;
      tw.cursors[2 * ((size_t)n * T + t)] = 0; tw.cursors[2 * ((size_t)n * T + t) + 1] = 0;
    }
    __syncthreads();
    if (tid == 0) s_carry = carry + tot;
    __syncthreads();
  }
}

// ---- 1-D TMA bulk copy + mbarrier (sm_90+ PTX; a non-cluster launch is a cluster of one CTA) ----
__device__ __forceinline__ void mbar_init(unsigned int mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned int mbar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int mbar, unsigned int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
}

// exact test of one (face, pixel) candidate against the tile's shared key (early depth reject: see resolve_pixel_with)
__device__ __forceinline__ void resolve_pixel_smem(const Face& fc, const FaceEdges& fe, int fid, unsigned int zmin_bits, bool persp,
                                                   float xf, float yf, unsigned long long* key_ptr) {
  const unsigned long long cur = *(volatile unsigned long long*)key_ptr;
  if (zmin_bits > (unsigned int)(cur >> 32)) return;
  float w[3], b[3], pz;
  if (!raster_test(fc, fe, persp, xf, yf, w, b, pz)) return;
  const unsigned long long key = make_key(pz, fid);
  if (key >= cur) return;
  smem_key_min(key_ptr, key);
}

// face-level rejections + exact pixel bbox restricted to the tile [x0, x1] x [y0, y1] (s_xf / s_yf: the tile's slice of the
// pixel-centre table) -- face_pixel_bbox with tile bounds
__device__ __forceinline__ bool face_tile_bbox(const Face& f, const MeshParams& p, int x0, int x1, int y0, int y1, const float* s_xf,
                                               const float* s_yf, int& xi_lo, int& xi_hi, int& yi_lo, int& yi_hi) {
  if (p.z_clip >= 0.f && f.z0 < p.z_clip && f.z1 < p.z_clip && f.z2 < p.z_clip) return false;
  const float zmin = fminf(fminf(f.z0, f.z1), f.z2);
  if (zmin < MVR_K_EPS) return false;
  const float face_area = (f.x0 - f.x1) * (f.y2 - f.y1) - (f.y0 - f.y1) * (f.x2 - f.x1);
  if ((p.flags & MVR_CULL_BACKFACES) && face_area < 0.f) return false;
  if (face_area <= MVR_K_EPS && face_area >= -1.0f * MVR_K_EPS) return false;
  const float xmin = fminf(fminf(f.x0, f.x1), f.x2), xmax = fmaxf(fmaxf(f.x0, f.x1), f.x2);
  const float ymin = fminf(fminf(f.y0, f.y1), f.y2), ymax = fmaxf(fmaxf(f.y0, f.y1), f.y2);
  pixel_range(xmin, xmax, p.W, p.H, x0, x1, s_xf, xi_lo, xi_hi);
  if (xi_lo > xi_hi) return false;
  pixel_range(ymin, ymax, p.H, p.W, y0, y1, s_yf, yi_lo, yi_hi);
  return yi_lo <= yi_hi;
}

// Hi-Z test of a face against the 8x8 blocks its (tile-local) pixel bbox touches: hidden when its nearest vertex is clearly
// behind the farthest current winner of every one of them (same conservative bound as the per-candidate early depth test of
// resolve_pixel_with: perspective-corrected depth is a convex combination of the vertex depths up to a few ulp).
__device__ __forceinline__ bool hiz_hidden(const Face& f, bool persp, const unsigned int* s_zmax, int lx0, int lx1, int ly0, int ly1) {
  const float zmin = fminf(fminf(f.z0, f.z1), f.z2);
  if (!(persp && zmin > 1e-3f)) return false;
  const unsigned int zb = __float_as_uint(zmin * 0.999999f);
  unsigned int zm = 0u;
  for (int by = ly0 >> 3; by <= (ly1 >> 3); ++by)
    for (int bx = lx0 >> 3; bx <= (lx1 >> 3); ++bx) zm = max(zm, s_zmax[by * 4 + bx]);
  return zb > zm;
}

// shading of one pixel from its final key (the body of mesh_shade_kernel for layer 0 of K = 1).  Returns false when the
// pixel belongs to mesh_shade_clipped_kernel (its winning face crosses the near plane): nothing is stored then.
template <bool EXACT, bool VRGB>
__device__ __forceinline__ bool tile_shade_pixel(const MeshParams& p, int n, int f0, int voff, const float4* __restrict__ pvn, bool persp,
                                                 unsigned long long key, float xf, float yf, int pix, const float bg[3], int HW) {
  int fid = -1;
  float w[3] = {-1.f, -1.f, -1.f}, bb[3] = {-1.f, -1.f, -1.f}, pz = -1.f, dd = -1.f;
  float out[3] = {bg[0], bg[1], bg[2]};
  if (key != MVR_EMPTY_KEY) {
    fid = (int)(unsigned int)(key & 0xffffffffull);
    const int4 fi = __ldg(p.faces4 + f0 + fid);
    const Face fc = gather_face(pvn, fi);
    if (may_clip(p.wsflags) && face_straddles(fc, p.z_clip)) return false;
    float4 X0, X1, X2, N0, N1, N2, c0, c1, c2;
    gather_xn(p.xn8, voff + fi.x, X0, N0); gather_xn(p.xn8, voff + fi.y, X1, N1); gather_xn(p.xn8, voff + fi.z, X2, N2);
    if (VRGB) { c0 = __ldg(p.rgb4 + voff + fi.x); c1 = __ldg(p.rgb4 + voff + fi.y); c2 = __ldg(p.rgb4 + voff + fi.z); }
    else { c0 = c1 = c2 = make_float4(__ldg(p.obj_rgb), __ldg(p.obj_rgb + 1), __ldg(p.obj_rgb + 2), 0.f); }
    const FaceEdges fe = face_edges(fc);
    if (EXACT) {
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
    const ShadeCtx sc = load_shade_ctx(p.light, p.light_stride, p.Cc, n);
    phong_pixel<VRGB>(bb, X0, X1, X2, N0, N1, N2, c0, c1, c2, sc, out);
  }
  const size_t po = (size_t)n * HW + pix;
  p.pix_to_face[po] = fid;
  if (EXACT) {
    if (p.zbuf) p.zbuf[po] = pz;
    if (p.dists) p.dists[po] = dd;
    if (p.bary) { p.bary[3 * po] = bb[0]; p.bary[3 * po + 1] = bb[1]; p.bary[3 * po + 2] = bb[2]; }
  }
  store_rgb(p.images, p.flags & MVR_IMAGES_BF16, (size_t)n * 3 * HW + pix, (size_t)HW, out[0], out[1], out[2], p.onorm);
  return true;
}

// grid: x = tile, y = view m, z = object b
template <bool EXACT, int MINB, bool VRGB>
__global__ void __launch_bounds__(MVR_THREADS, MINB) mesh_tile_kernel(const MeshParams p, const TileWs tw) {
  __shared__ float s_rec[REC_WORDS][MVR_THREADS];      // SoA face records of the current round
  __shared__ int s_items[T_ITEM_CAP];                  // slot | bbox row << 8
  __shared__ int s_cand[NWARPS][T_WCAP];               // slot | x << 8 | y << 20 (tile-local pixel)
  __shared__ int s_big[MVR_THREADS];
  __shared__ __align__(16) unsigned long long s_key[32 * 32];
  __shared__ __align__(16) int s_fid[2][FID_BUF];      // face ids of two rounds: TMA bulk-copy destinations
  __shared__ unsigned int s_zmax[16];                  // Hi-Z: farthest winning depth (bits) of every 8x8-pixel block of the tile
  __shared__ __align__(8) unsigned long long s_mbar[2];
  __shared__ float s_tab[64];                          // pixel centres of the tile: xf[32], yf[32]
  __shared__ int s_cnt[2];
  __shared__ int s_wcnt[NWARPS];
  const float* s_xf = s_tab;
  const float* s_yf = s_tab + 32;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = blockIdx.x, m = blockIdx.y, b = blockIdx.z, n = b * p.M + m;
  int ty, tx;
  tile_rc(t, tw.tiles_x, ty, tx);
  const int x0 = tx * 32, y0 = ty * 32, x1 = min(x0 + 31, p.W - 1), y1 = min(y0 + 31, p.H - 1);
  const int HW = p.H * p.W;
  const int f0 = p.face_off[b], F = p.face_off[b + 1] - f0;
  const int voff = p.vert_off[b], V = p.vert_off[b + 1] - voff;
  const float4* pvn = p.pv + (size_t)p.M * voff + (size_t)m * V;
  const bool persp = p.flags & MVR_PERSPECTIVE_CORRECT;
  const size_t fb = (size_t)p.M * f0 + (size_t)m * F;
  const size_t tb = (size_t)n * tw.tiles + t;
  // the tile's list: [0, n_front) the nearer facing, [n_front, n_list) the other one; each facing gets its own rounds, so
  // that the Hi-Z plane left by the first can cull the second
  const int n_list = tw.tile_cnt[tb], n_front = tw.cursors[2 * tb], n_bigv = tw.big_cnt[n];
  const int* list = tw.lists + 4 * fb + 4 * (size_t)n * tw.tiles + tw.tile_off[tb];
  const int* big_list = tw.big_list + fb;
  const int rounds_front = (n_front + MVR_THREADS - 1) / MVR_THREADS;
  const int list_rounds = rounds_front + (n_list - n_front + MVR_THREADS - 1) / MVR_THREADS;
  auto round_range = [&](int r, int& start, int& cnt) {
    if (r < rounds_front) { start = r * MVR_THREADS; cnt = min(MVR_THREADS, n_front - start); }
    else { start = n_front + (r - rounds_front) * MVR_THREADS; cnt = min(MVR_THREADS, n_list - start); }
  };

  const unsigned int mbar_a = smem_addr_pinned(&s_mbar[0]);
  const unsigned int fid_a = smem_addr_pinned(&s_fid[0][0]);
  auto issue_ids = [&](int r) {      // one thread: bulk copy of round r's ids (16-byte aligned source, size a multiple of 16)
    int start, cnt;
    round_range(r, start, cnt);
    const int a0 = start & ~3;
    const unsigned int bytes = (unsigned int)(((start - a0 + cnt + 3) & ~3) * 4);
    const unsigned int mb = mbar_a + 8u * (r & 1);
    mbar_expect_tx(mb, bytes);
    bulk_g2s(fid_a + 4u * FID_BUF * (r & 1), list + a0, bytes, mb);
  };
  if (tid == 0) {
    mbar_init(mbar_a, 1); mbar_init(mbar_a + 8u, 1);
    fence_mbar_init();
    if (list_rounds > 0) issue_ids(0);      // round 0 of the list is on its way while the CTA sets up
    s_cnt[0] = 0; s_cnt[1] = 0;
  }
  if (tid < 16) s_zmax[tid] = 0xFFFFFFFFu;
  if (tid < 32) s_tab[tid] = x0 + tid <= x1 ? __ldg(p.tab + x0 + tid) : 0.f;
  else if (tid < 64) s_tab[tid] = y0 + tid - 32 <= y1 ? __ldg(p.tab + p.W + y0 + tid - 32) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) s_key[tid + MVR_THREADS * j] = MVR_EMPTY_KEY;
  __syncthreads();
  const unsigned int my_cand_a = smem_addr_pinned(&s_cand[warp][0]);

  const int total_rounds = list_rounds + (n_bigv + MVR_THREADS - 1) / MVR_THREADS;
  for (int r = 0; r < total_rounds; ++r) {
    if (r > 0) {
      // Hi-Z: farthest winner of every 8x8 block after the previous rounds (an uncovered pixel keeps the block open; pixels
      // outside the image do not count).  16 threads per block, 4 pixels each.
      const int bq = tid >> 4, sub = tid & 15, px = ((bq & 3) << 3) + (sub & 7), py = ((bq >> 2) << 3) + (sub >> 3);
      unsigned int zm = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ly = py + 2 * j;
        const unsigned int zb = (unsigned int)(s_key[ly * 32 + px] >> 32);
        zm = max(zm, (x0 + px <= x1 && y0 + ly <= y1) ? zb : 0u);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) zm = max(zm, __shfl_xor_sync(0xffffffffu, zm, o));
      if (sub == 0) s_zmax[bq] = zm;
      __syncthreads();
    }
    // ---------------- phase A: setup, one thread per face ----------------
    int fid = -1;
    if (r < list_rounds) {
      const int buf = r & 1;
      if (tid == 0 && r + 1 < list_rounds) issue_ids(r + 1);      // the other buffer was released by the barrier that ended round r - 1
      mbar_wait(mbar_a + 8u * buf, (unsigned int)((r >> 1) & 1));
      int start, cnt;
      round_range(r, start, cnt);
      if (tid < cnt) fid = s_fid[buf][(start & 3) + tid];
    } else {
      const int i = (r - list_rounds) * MVR_THREADS + tid;
      if (i < n_bigv) fid = __ldg(big_list + i);
    }
    if (fid >= 0) {
      const Face fc = gather_face(pvn, __ldg(p.faces4 + f0 + fid));
      int xl, xh, yl, yh;
      if (face_straddles(fc, p.z_clip)) {
        s_rec[0][tid] = fc.x0; s_rec[1][tid] = fc.y0; s_rec[2][tid] = fc.z0;
        s_rec[3][tid] = fc.x1; s_rec[4][tid] = fc.y1; s_rec[5][tid] = fc.z1;
        s_rec[6][tid] = fc.x2; s_rec[7][tid] = fc.y2; s_rec[8][tid] = fc.z2;
        s_rec[9][tid] = __int_as_float(fid);
        s_big[atomicAdd(&s_cnt[1], 1)] = tid | 0x100;
      } else if (face_tile_bbox(fc, p, x0, x1, y0, y1, s_xf, s_yf, xl, xh, yl, yh) && !hiz_hidden(fc, persp, s_zmax, xl - x0, xh - x0, yl - y0, yh - y0)) {
        const int bw = xh - xl + 1, bh = yh - yl + 1, npx = bw * bh;
        s_rec[0][tid] = fc.x0; s_rec[1][tid] = fc.y0; s_rec[2][tid] = fc.z0;
        s_rec[3][tid] = fc.x1; s_rec[4][tid] = fc.y1; s_rec[5][tid] = fc.z1;
        s_rec[6][tid] = fc.x2; s_rec[7][tid] = fc.y2; s_rec[8][tid] = fc.z2;
        s_rec[9][tid] = __int_as_float(fid);
        s_rec[10][tid] = __int_as_float((xl - x0) | ((yl - y0) << 16));      // tile-local
        s_rec[11][tid] = __int_as_float(bw | (bh << 16));
        store_filter_edges(fc, p.ndc_max, &s_rec[12][tid]);
        bool queued = false;
        if (npx <= 256) {
          // one sub-item per bbox ROW (phase B turns it into the interval of candidate pixels)
          const int at = atomicAdd(&s_cnt[0], bh);
          if (at + bh <= p.item_cap) {
            for (int q = 0; q < bh; ++q) s_items[at + q] = tid | (q << 8);
            queued = true;
          } else {
            for (int q = at; q < p.item_cap; ++q) s_items[q] = 0x7fffff00;      // a straddling reservation leaves no garbage (row beyond any bbox)
          }
        }
        if (!queued) s_big[atomicAdd(&s_cnt[1], 1)] = tid;               // walked by the whole CTA below
      }
    }
    __syncthreads();
    // ---------------- phase B: scanline spans of the bbox rows (row_span) -> per-warp candidate queues ----------------
    const int n_items = min(s_cnt[0], p.item_cap);
    const int n_bigf = s_cnt[1];
    int wcnt = 0;
    for (int j0 = warp * 32; j0 < n_items; j0 += MVR_THREADS) {
      const int j = j0 + lane;
      int slot = 0, cnt = 0, xs = 0, yy = 0;      // xs, yy: tile-local
      if (j < n_items) {
        const int it = s_items[j];
        slot = it & 255;
        const int row = it >> 8;
        const int rxy = __float_as_int(s_rec[10][slot]), rwh = __float_as_int(s_rec[11][slot]);
        if (row < (rwh >> 16)) {
          yy = (rxy >> 16) + row;
          cnt = row_span(&s_rec[12][slot], MVR_THREADS, s_yf[yy], x0 + (rxy & 0xffff), rwh & 0xffff, p.W, p.jx_scale, p.jx_off, xs);
          xs -= x0;
        }
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total == 0) continue;
      const int base = wcnt + incl - cnt;
      const unsigned int word0 = (unsigned)slot | ((unsigned)xs << 8) | ((unsigned)yy << 20);
      const int maxc = __reduce_max_sync(0xffffffffu, cnt);
      for (int c = 0; c < maxc; ++c) {
        if (c < cnt) {
          const int at = base + c;
          if (at < p.wcap) {
            sts_u32(my_cand_a + 4u * (unsigned)at, word0 + ((unsigned)c << 8));
          } else {                                                  // queue full: resolve in place
            const int lx = xs + c;
            Face fc;
            fc.x0 = s_rec[0][slot]; fc.y0 = s_rec[1][slot]; fc.z0 = s_rec[2][slot];
            fc.x1 = s_rec[3][slot]; fc.y1 = s_rec[4][slot]; fc.z1 = s_rec[5][slot];
            fc.x2 = s_rec[6][slot]; fc.y2 = s_rec[7][slot]; fc.z2 = s_rec[8][slot];
            resolve_pixel_smem(fc, face_edges(fc), __float_as_int(s_rec[9][slot]), 0u, persp, s_xf[lx], s_yf[yy], &s_key[yy * 32 + lx]);
          }
        }
      }
      wcnt += total;
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
      const int total_c = pre[NWARPS];
      for (int j = tid; j < total_c; j += MVR_THREADS) {
        int wi = 0;
#pragma unroll
        for (int q = 1; q < NWARPS; ++q) wi += (j >= pre[q]);
        const int cd = s_cand[wi][j - pre[wi]];
        const int slot = cd & 255, lx = (cd >> 8) & 4095, ly = (cd >> 20) & 4095;
        // early depth reject straight from the shared key, before the face record is even read
        const float zmin = fminf(fminf(s_rec[2][slot], s_rec[5][slot]), s_rec[8][slot]);
        const unsigned int zmin_bits = (persp && zmin > 1e-3f) ? __float_as_uint(zmin * 0.999999f) : 0u;
        unsigned long long* kp = &s_key[ly * 32 + lx];
        if (zmin_bits > (unsigned int)((*(volatile unsigned long long*)kp) >> 32)) continue;
        Face fc;
        fc.x0 = s_rec[0][slot]; fc.y0 = s_rec[1][slot]; fc.z0 = s_rec[2][slot];
        fc.x1 = s_rec[3][slot]; fc.y1 = s_rec[4][slot]; fc.z1 = s_rec[5][slot];
        fc.x2 = s_rec[6][slot]; fc.y2 = s_rec[7][slot]; fc.z2 = s_rec[8][slot];
        resolve_pixel_smem(fc, face_edges(fc), __float_as_int(s_rec[9][slot]), zmin_bits, persp, s_xf[lx], s_yf[ly], kp);
      }
    }
    // ---------------- large / clipped faces: the whole CTA walks the bbox (restricted to the tile) ----------------
    for (int q = 0; q < n_bigf; ++q) {
      const int ent = s_big[q], slot = ent & 255;
      Face fc;
      fc.x0 = s_rec[0][slot]; fc.y0 = s_rec[1][slot]; fc.z0 = s_rec[2][slot];
      fc.x1 = s_rec[3][slot]; fc.y1 = s_rec[4][slot]; fc.z1 = s_rec[5][slot];
      fc.x2 = s_rec[6][slot]; fc.y2 = s_rec[7][slot]; fc.z2 = s_rec[8][slot];
      const int bfid = __float_as_int(s_rec[9][slot]);
      if (ent & 0x100) {      // near-plane clipping: the sub-triangles compete under the ORIGINAL face id
        ClipSub cs;
        clip_face(fc, p.z_clip, persp, cs);
        for (int s = 0; s < cs.ns; ++s) {
          const Face sf = cs.f[s];
          int cxl, cxh, cyl, cyh;
          if (!face_tile_bbox(sf, p, x0, x1, y0, y1, s_xf, s_yf, cxl, cxh, cyl, cyh)) continue;
          const FaceEdges sfe = face_edges(sf);
          for (int yy = cyl + warp; yy <= cyh; yy += NWARPS)
            for (int xx = cxl + lane; xx <= cxh; xx += 32)
              resolve_pixel_smem(sf, sfe, bfid, 0u, persp, s_xf[xx - x0], s_yf[yy - y0], &s_key[(yy - y0) * 32 + xx - x0]);
        }
        continue;
      }
      const FaceEdges fe = face_edges(fc);
      const int rxy = __float_as_int(s_rec[10][slot]), rwh = __float_as_int(s_rec[11][slot]);
      const int xl = rxy & 0xffff, yl = rxy >> 16, bw = rwh & 0xffff, bh = rwh >> 16;
      const float zmin = fminf(fminf(fc.z0, fc.z1), fc.z2);
      const unsigned int zmin_bits = (persp && zmin > 1e-3f) ? __float_as_uint(zmin * 0.999999f) : 0u;
      for (int y = warp; y < bh; y += NWARPS)
        for (int x = lane; x < bw; x += 32)
          resolve_pixel_smem(fc, fe, bfid, zmin_bits, persp, s_xf[xl + x], s_yf[yl + y], &s_key[(yl + y) * 32 + xl + x]);
    }
    __syncthreads();
  }
  // ---------------- shade: 4 pixels per thread straight from the shared keys ----------------
  const int lx = lane, xi = x0 + lx;
  const bool clip = may_clip(p.wsflags);      // block-uniform
  float bg[3] = {__ldg(p.bg_rgb), __ldg(p.bg_rgb + 1), __ldg(p.bg_rgb + 2)};
  if (xi <= x1) {
    const float xf = s_xf[lx];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ly = warp + 8 * j, yi = y0 + ly;
      if (yi > y1) break;
      const unsigned long long key = s_key[ly * 32 + lx];
      const int pix = yi * p.W + xi;
      // faces may cross the near plane in this call: mesh_shade_clipped_kernel needs the keys (it owns the pixels whose
      // winning face is a straddler)
      if (clip) p.keys[(size_t)n * HW + pix] = key;
      tile_shade_pixel<EXACT, VRGB>(p, n, f0, voff, pvn, persp, key, xf, s_yf[ly], pix, bg, HW);
    }
  }
}

}  // namespace mvr

using namespace mvr;

// The tile-binned forward is OPT-IN (flag MVR_FORWARD_TILED, or MVR_MESH_TILED=1 in the environment for A/B runs): measured
// on B200 it loses to the bin-free scatter + shade pair at every BASELINE configuration -- the
// forward is instruction- and latency-bound, not traffic-bound, so the ~0.1 ms binning pass and the barrier-separated phases of
// a tile CTA cost more than the key-plane traffic they remove.
bool mesh_tiled_enabled() {
  static const bool v = [] { const char* e = getenv("MVR_MESH_TILED"); return e && atoi(e) == 1; }();
  return v;
}

int launch_mesh_forward_tiled(MeshParams p, const WsLayout& w, void* workspace, int B, int M, int max_faces, bool exact,
                              cudaStream_t st) {
  char* wb = (char*)workspace;
  TileWs tw;
  tw.codes = (unsigned int*)(wb + w.codes); tw.lists = (int*)(wb + w.lists); tw.big_list = (int*)(wb + w.big_list);
  tw.tile_cnt = (int*)(wb + w.tile_cnt); tw.big_cnt = (int*)(wb + w.big_cnt); tw.vote = (float*)(wb + w.vote);
  tw.tile_off = (int*)(wb + w.tile_off); tw.cursors = (int*)(wb + w.cursors); tw.front_sign = (int*)(wb + w.front_sign);
  tw.tiles_x = w.tiles_x; tw.tiles = w.tiles;
  cudaError_t e = cudaMemsetAsync(wb + w.zero_begin, 0, w.zero_end - w.zero_begin, st);
  if (e != cudaSuccess) { set_error("mvr_mesh_forward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
  p.item_cap = (p.flags & MVR_TEST_TINY_QUEUES) ? 24 : T_ITEM_CAP;
  p.wcap = (p.flags & MVR_TEST_TINY_QUEUES) ? 5 : T_WCAP;
  const int N = B * M;
  if (max_faces > 0) {
    const dim3 bgrid((unsigned)((max_faces + MVR_THREADS - 1) / MVR_THREADS), (unsigned)M, (unsigned)B);
    MVR_LAUNCH(mesh_bin_kernel<false>, bgrid, MVR_THREADS, 0, st, p, tw);
    MVR_LAUNCH(mesh_bin_scan_kernel, (unsigned)N, MVR_THREADS, 0, st, tw);
    MVR_LAUNCH(mesh_bin_kernel<true>, bgrid, MVR_THREADS, 0, st, p, tw);
  }
  const dim3 tgrid((unsigned)w.tiles, (unsigned)M, (unsigned)B);
  const bool vrgb = p.flags & MVR_RGB_PER_ELEMENT;
  if (exact) {
    if (vrgb) MVR_LAUNCH((mesh_tile_kernel<true, 3, true>), tgrid, MVR_THREADS, 0, st, p, tw);
    else MVR_LAUNCH((mesh_tile_kernel<true, 3, false>), tgrid, MVR_THREADS, 0, st, p, tw);
  } else if (vrgb) MVR_LAUNCH((mesh_tile_kernel<false, 4, true>), tgrid, MVR_THREADS, 0, st, p, tw);
  else MVR_LAUNCH((mesh_tile_kernel<false, 4, false>), tgrid, MVR_THREADS, 0, st, p, tw);
  return check_launch("mesh_tile_kernel");
}
