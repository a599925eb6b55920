// K10b: tiled-decode blend.  Replaces, per tile, `image[tile] += decode_latents(...)[centre]`, `count[tile] += 1`
// and the final `image / count` of tiled_decode (elastic_diffusion.py:303-308), with decode_latents' `/2 + 0.5`
// and clamp (ed:271) fused in.  One pass: each output pixel gathers the centre pixels of the (few) tiles that cover
// it, in ascending tile order (the reference's accumulation order, so the fp32 sum is bit-identical).
#include "common.cuh"

namespace ed {

template <typename PT, int VEC>
__global__ void __launch_bounds__(256) tile_blend_kernel(const ed_tiles_t T, const PT* __restrict__ patches,
                                                         float* __restrict__ image) {
  const int Hp = T.H * T.scale, Wp = T.W * T.scale;
  const int side = (T.core + 2 * T.pad) * T.scale;   // decoded patch side in pixels
  const int padp = T.pad * T.scale;
  const int wv = Wp / VEC;
  const long long total = (long long)T.B * T.CH * Hp * wv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xv = (int)(i % wv);
    long long r = i / wv;
    const int y = (int)(r % Hp);
    r /= Hp;
    const int ch = (int)(r % T.CH);
    const int b = (int)(r / T.CH);
    const int ly = y / T.scale;
    const int lx = (xv * VEC) / T.scale;        // VEC consecutive pixels share one latent column (scale % VEC == 0)
    const int r0 = __ldg(T.trow_first + ly), rn = __ldg(T.trow_cnt + ly);
    const int c0 = __ldg(T.tcol_first + lx), cn = __ldg(T.tcol_cnt + lx);
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int a = 0; a < rn; ++a)
      for (int q = 0; q < cn; ++q) {
        const int j = (r0 + a) * T.ntc + (c0 + q);
        const int h0 = __ldg(T.tiles + j * 4 + 0), w0 = __ldg(T.tiles + j * 4 + 2);
        const int py = padp + (y - h0 * T.scale);
        const int px = padp + (xv * VEC - w0 * T.scale);
        const PT* src = patches + ((((long long)j * T.B + b) * T.CH + ch) * side + py) * side + px;
        float v[VEC];
        if constexpr (VEC == 4 && sizeof(PT) == 4) {
          const float4 t = __ldcs(reinterpret_cast<const float4*>(src));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) v[e] = to_f32<PT>(src[e]);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float p = __fadd_rn(__fmul_rn(v[e], 0.5f), 0.5f);          // imgs / 2 + 0.5 (ed:271); x/2 == x*0.5 exactly
          p = fminf(fmaxf(p, 0.f), 1.f);                             // .clamp(0, 1)
          acc[e] = __fadd_rn(acc[e], p);                             // image[...] += patch (ed:306)
        }
      }
    const float cnt = (float)(rn * cn);                              // count[...] += 1 per covering tile (ed:307)
    float* dst = image + (((long long)b * T.CH + ch) * Hp + y) * Wp + xv * VEC;
    if constexpr (VEC == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(__fdiv_rn(acc[0], cnt), __fdiv_rn(acc[1], cnt),
                                                    __fdiv_rn(acc[2], cnt), __fdiv_rn(acc[3], cnt));
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) dst[e] = __fdiv_rn(acc[e], cnt);
    }
  }
}

}  // namespace ed

using namespace ed;

extern "C" int ed_tile_blend(const ed_tiles_t* tiles, const void* patches, int patch_dtype, float* image,
                             void* stream_) {
  if (!tiles || !patches || !image) return ED_ERR_INVALID;
  const ed_tiles_t& T = *tiles;
  if (!T.tiles || !T.trow_first || !T.trow_cnt || !T.tcol_first || !T.tcol_cnt || T.ntiles <= 0 || T.ntc <= 0 ||
      T.scale <= 0 || T.B <= 0 || T.CH <= 0)
    return ED_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int Wp = T.W * T.scale;
  const bool vec = (T.scale % 4 == 0) && ((reinterpret_cast<uintptr_t>(image) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(patches) & 15) == 0);
  const long long total = (long long)T.B * T.CH * T.H * T.scale * (vec ? Wp / 4 : Wp);
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
#define ED_BLEND(PT)                                                                            \
  if (vec) tile_blend_kernel<PT, 4><<<(int)g, 256, 0, stream>>>(T, (const PT*)patches, image);   \
  else tile_blend_kernel<PT, 1><<<(int)g, 256, 0, stream>>>(T, (const PT*)patches, image);
  switch (patch_dtype) {
    case ED_F32: ED_BLEND(float) break;
    case ED_F16: ED_BLEND(__half) break;
    case ED_BF16: ED_BLEND(__nv_bfloat16) break;
    default: return ED_ERR_INVALID;
  }
#undef ED_BLEND
  ED_LAUNCH_CHECK();
  return ED_OK;
}
