// K10b: tiled-decode blend.  Replaces, per tile, `image[tile] += decode_latents(...)[centre]`, `count[tile] += 1`
// and the final `image / count` of tiled_decode (elastic_diffusion.py:303-308), with decode_latents' `/2 + 0.5`
// and clamp (ed:271) fused in.  One pass: each output pixel gathers the centre pixels of the (few) tiles that cover
// it, in ascending tile order (the reference's accumulation order, so the fp32 sum is bit-identical).
#include "common.cuh"

namespace ed {

// Where the decoded patches live.  Single GPU: one buffer.  Tiles sharded over ranks (ed_tile_blend_peer): patch p =
// j*B + b lives in the buffer of rank p / per at local index p % per; `peers` is a device array of peer-mapped base
// pointers and the centre crops are read straight from their owners over NVLink (no gathered copy).
struct PatchSrc {
  const void* local;
  const void* const* peers;
  int per;
};

// One thread produces VEC consecutive pixels of ROWS consecutive image rows of one (b, ch) plane.  ROWS rows aligned to
// ROWS lie in the same latent row when scale % ROWS == 0, so the tile-cover lookups are shared and the ROWS patch
// loads are independent (ROWS x VEC x sizeof(PT) bytes in flight per thread).
template <typename PT, int VEC, int ROWS, bool PEER>
__global__ void __launch_bounds__(256) tile_blend_kernel(const ed_tiles_t T, const PatchSrc S, float* __restrict__ image) {
  const long long patch_elems = (long long)T.CH * (T.core + 2 * T.pad) * T.scale * (T.core + 2 * T.pad) * T.scale;
  auto patch = [&](int p) -> const PT* {           // start of decoded patch p = j*B + b (all CH channels)
    if constexpr (PEER) {
      const int r = p / S.per;
      return static_cast<const PT*>(S.peers[r]) + (long long)(p - r * S.per) * patch_elems;
    } else {
      return static_cast<const PT*>(S.local) + (long long)p * patch_elems;
    }
  };
  const int Hp = T.H * T.scale, Wp = T.W * T.scale;
  const int side = (T.core + 2 * T.pad) * T.scale;   // decoded patch side in pixels
  const int padp = T.pad * T.scale;
  const int xv = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * ROWS;
  if (xv * VEC >= Wp || y0 >= Hp) return;
  const int ly = y0 / T.scale;
  const int lx = (xv * VEC) / T.scale;          // the VEC pixels share one latent column (scale % VEC == 0)
  const int r0 = __ldg(T.trow_first + ly), rn = __ldg(T.trow_cnt + ly);
  const int c0 = __ldg(T.tcol_first + lx), cn = __ldg(T.tcol_cnt + lx);
  const int icnt = rn * cn;                                          // count[...] += 1 per covering tile (ed:307)
  const float cnt = (float)icnt;
  // image / count: for power-of-two counts (1, 2, 4 - everything but the shifted last tiles of low_vram decoding) the
  // IEEE quotient equals the product with the exact reciprocal; the ~30-instruction division is kept for the rest
  const bool pow2 = (icnt & (icnt - 1)) == 0;
  const float rcp = 1.0f / cnt;
  auto load_row = [&](const PT* sp, float (&v)[VEC]) {
    if constexpr (sizeof(PT) == 4 && VEC % 4 == 0) {
#pragma unroll
      for (int e = 0; e < VEC; e += 4) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(sp + e));
        v[e] = t.x; v[e + 1] = t.y; v[e + 2] = t.z; v[e + 3] = t.w;
      }
    } else if constexpr (sizeof(PT) == 2 && VEC == 8) {
      const uint4 t = __ldcs(reinterpret_cast<const uint4*>(sp));   // 8 x 16-bit
      const PT* h = reinterpret_cast<const PT*>(&t);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = to_f32<PT>(h[e]);
    } else if constexpr (sizeof(PT) == 2 && VEC == 4) {
      const uint2 t = __ldcs(reinterpret_cast<const uint2*>(sp));   // 4 x 16-bit
      const PT* h = reinterpret_cast<const PT*>(&t);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = to_f32<PT>(h[e]);
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = to_f32<PT>(sp[e]);
    }
  };
  auto store_row = [&](float* dst, const float (&o)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
      for (int e = 0; e < VEC; e += 4) __stcs(reinterpret_cast<float4*>(dst + e), make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]));
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) dst[e] = o[e];
    }
  };
  const int rows_here = (Hp - y0) < ROWS ? (Hp - y0) : ROWS;
  for (int z = blockIdx.z; z < T.B * T.CH; z += gridDim.z) {
    const int b = z / T.CH, ch = z - b * T.CH;
    float* dst0 = image + (((long long)b * T.CH + ch) * Hp + y0) * Wp + xv * VEC;
    if (icnt == 1 && rows_here == ROWS) {
      // ---- fast path (one covering tile, the normal case): all ROWS row loads in flight, then clamp + store -------------
      const int j = r0 * T.ntc + c0;
      const int h0 = __ldg(T.tiles + j * 4 + 0), w0 = __ldg(T.tiles + j * 4 + 2);
      const PT* src = patch(j * T.B + b) + ((long long)ch * side + padp + (y0 - h0 * T.scale)) * side + padp +
                      (xv * VEC - w0 * T.scale);
      float v[ROWS][VEC];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) load_row(src + (long long)r * side, v[r]);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float p = __fadd_rn(__fmul_rn(v[r][e], 0.5f), 0.5f);         // imgs / 2 + 0.5 (ed:271); x/2 == x*0.5 exactly
          v[r][e] = fminf(fmaxf(p, 0.f), 1.f);                         // .clamp(0, 1);  0 + p and p / 1 are exact
        }
        store_row(dst0 + (long long)r * Wp, v[r]);
      }
    } else {
      // ---- general path (overlapping tiles / ragged bottom): one row at a time, tiles in ascending order ---------------
#pragma unroll 1
      for (int r = 0; r < rows_here; ++r) {
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
        for (int a = 0; a < rn; ++a)
          for (int q = 0; q < cn; ++q) {                                 // ascending tile order = the reference's += order
            const int j = (r0 + a) * T.ntc + (c0 + q);
            const int h0 = __ldg(T.tiles + j * 4 + 0), w0 = __ldg(T.tiles + j * 4 + 2);
            const int py = padp + (y0 + r - h0 * T.scale);
            const int px = padp + (xv * VEC - w0 * T.scale);
            float v[VEC];
            load_row(patch(j * T.B + b) + ((long long)ch * side + py) * side + px, v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              float p = __fadd_rn(__fmul_rn(v[e], 0.5f), 0.5f);
              p = fminf(fmaxf(p, 0.f), 1.f);
              acc[e] = __fadd_rn(acc[e], p);                             // image[...] += patch (ed:306)
            }
          }
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = pow2 ? __fmul_rn(acc[e], rcp) : __fdiv_rn(acc[e], cnt);   // image / count
        store_row(dst0 + (long long)r * Wp, acc);
      }
    }
  }
}

}  // namespace ed

using namespace ed;

static int launch_tile_blend(const ed_tiles_t* tiles, const void* patches, const void* const* peers, int world, int per,
                             int patch_dtype, float* image, void* stream_) {
  if (!tiles || (!patches && !peers) || !image) return ED_ERR_INVALID;
  if (peers && (world <= 0 || per <= 0)) return ED_ERR_INVALID;
  const ed_tiles_t& T = *tiles;
  if (!T.tiles || !T.trow_first || !T.trow_cnt || !T.tcol_first || !T.tcol_cnt || T.ntiles <= 0 || T.ntc <= 0 ||
      T.scale <= 0 || T.B <= 0 || T.CH <= 0)
    return ED_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int Wp = T.W * T.scale, Hp = T.H * T.scale;
  // peer buffers are symmetric-memory allocations (>= 256-byte aligned); a patch is a multiple of 16 bytes when scale % 4 == 0
  const bool al = ((reinterpret_cast<uintptr_t>(image) & 15) == 0) && ((reinterpret_cast<uintptr_t>(patches) & 15) == 0);
  const int vec = (al && T.scale % 8 == 0) ? 8 : (al && T.scale % 4 == 0) ? 4 : 1;
  const int rows_per = (vec == 8) ? 8 : (T.scale % 4 == 0) ? 4 : 1;
  const int cols = Wp / vec, rows = (Hp + rows_per - 1) / rows_per;
  const dim3 block(32, 8);
  const int planes = T.B * T.CH;
  const dim3 g((cols + 31) / 32, (rows + 7) / 8, planes > 65535 ? 65535 : planes);
  const PatchSrc S{patches, peers, per};
#define ED_BLEND_AS(PT, PEER)                                                                         \
  if (vec == 8 && rows_per == 8) tile_blend_kernel<PT, 8, 8, PEER><<<g, block, 0, stream>>>(T, S, image); \
  else if (vec == 4) tile_blend_kernel<PT, 4, 4, PEER><<<g, block, 0, stream>>>(T, S, image);          \
  else if (rows_per >= 4) tile_blend_kernel<PT, 1, 4, PEER><<<g, block, 0, stream>>>(T, S, image);     \
  else tile_blend_kernel<PT, 1, 1, PEER><<<g, block, 0, stream>>>(T, S, image);
#define ED_BLEND(PT)         \
  if (peers) {               \
    ED_BLEND_AS(PT, true)    \
  } else {                   \
    ED_BLEND_AS(PT, false)   \
  }
  switch (patch_dtype) {
    case ED_F32: ED_BLEND(float) break;
    case ED_F16: ED_BLEND(__half) break;
    case ED_BF16: ED_BLEND(__nv_bfloat16) break;
    default: return ED_ERR_INVALID;
  }
#undef ED_BLEND
#undef ED_BLEND_AS
  ED_LAUNCH_CHECK();
  return ED_OK;
}

extern "C" int ed_tile_blend(const ed_tiles_t* tiles, const void* patches, int patch_dtype, float* image, void* stream_) {
  return launch_tile_blend(tiles, patches, nullptr, 0, 0, patch_dtype, image, stream_);
}

extern "C" int ed_tile_blend_peer(const ed_tiles_t* tiles, const void* const* d_peer_patches, int world, int per,
                                  int patch_dtype, float* image, void* stream_) {
  return launch_tile_blend(tiles, nullptr, d_peer_patches, world, per, patch_dtype, image, stream_);
}
