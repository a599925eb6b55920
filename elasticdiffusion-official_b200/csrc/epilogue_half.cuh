// "Half" fused wave epilogue: the exact 1/2-ratio geometry with tiling views (plan flag ED_PLAN_HALF_FAST, every tiled
// BASELINE config), where every index of the generic kernels' per-pixel reference tables is closed-form in (y, x).
//
// Same contract and the same floating-point operation order as the direct / staged kernels (reference
// elastic_diffusion.py "ed:N": scatter ed:852-861, direction fills ed:439-440 / 634-647, CFG + DDIM ed:1031-1035, RRG
// ed:886-940 + ed:1078) - results are bit-identical - with ~4x fewer instructions per element:
//   * a thread owns 2 latent rows x 8 columns of one (batch entry, channel): exactly the 2x2 footprints of 4 low-res cells.
//     Everything per cell (direction, g * direction, the RRG low-res reference x0) is evaluated once per cell instead of
//     once per pixel, and all loads are 8/16-byte vectors: 4 x 16 B of latent, 2 x 16 B (bf16/fp16; 4 x 16 B fp32) of the
//     single covering view, 2 x 8 B of (uncond, cond) low-res scores - no per-pixel reference loads at all.
//   * view index / offset come from the per-row and per-column tables vrow_first / vrow_off / vcol_first / vcol_off
//     (4 small L1-resident loads per thread, shared by every (b, c) the thread walks).
//   * R1 == 1 (wave 2 of a repaint step, the RRG launch of the BASELINE configs; resampling_steps == 0): the owner of every
//     pixel is iteration 0 - no owner map, no shared memory, no TMA: a pure streaming kernel.
//   * R1 > 1 without a noise stream (last step; repaint_sampling=False): the (uncond, cond) scores of ALL R1 iterations of
//     the 4 cells are loaded as 2*R1 independent 8-byte vectors into shared memory slots private to the thread (no
//     barrier), then picked per pixel by owner.
// Launches with a re-noise stream (wave 1 of a repaint step) stay on the tile-staged kernel, which already runs at 0.95 of
// the HBM roofline behind its 20-tensor noise stream.
//
// Compiled twice like epilogue_staged.cuh: by nvcc into libelastic_b200.so and by g++ (tests/emu) into the host emulation.
#pragma once
#include "epilogue_staged.cuh"

#define HALF_THREADS 128   // threads per CTA of the half kernels (the per-thread slot stride of the R1 > 1 kernel depends on it)
#ifndef ED_HALF_MINB_MULTI
#define ED_HALF_MINB_MULTI 5   // same for the R1 > 1 kernel (measured: 4 -> 0.52, 5 -> 0.56, 6 -> 0.54 of the HBM roofline)
#endif
#ifndef ED_HALF_MINB
#define ED_HALF_MINB 7   // resident 128-thread CTAs per SM the R1 == 1 kernel is compiled for (register cap 65536 / (128 * MINB));
                         // measured plain / +rrg fraction of the HBM roofline: 5 -> 0.78 / 0.73, 6 -> 0.80 / 0.73, 7 -> 0.84 / 0.72, 8 -> 0.80 / 0.67
#endif

namespace ed {

// 8 consecutive elements with read-only vector loads (p aligned to 8 elements)
ED_DEVICE void ld8_ro(const float* p, float out[8]) {
  ld4_ro(p, out);
  ld4_ro(p + 4, out + 4);
}
ED_DEVICE void ld8_ro(const __half* p, float out[8]) {
  const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  __half h[8];
  memcpy(h, &t, 16);
#pragma unroll
  for (int e = 0; e < 8; ++e) out[e] = __half2float(h[e]);
}
ED_DEVICE void ld8_ro(const __nv_bfloat16* p, float out[8]) {
  const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  __nv_bfloat16 h[8];
  memcpy(h, &t, 16);
#pragma unroll
  for (int e = 0; e < 8; ++e) out[e] = __bfloat162float(h[e]);
}

// a / b through the shared reciprocal (struct DivBy) WITHOUT div_by()'s per-quotient fall-back: the range test is folded
// into `bad` (one FSETP with a predicate accumulate per quotient instead of test + branch + reconvergence), and the caller
// redoes its whole tile with IEEE division when any quotient was inf / nan / near overflow (practically never).
// EXACT = true is that slow path.
template <bool EXACT>
ED_DEVICE float half_div(const DivBy& d, float a, bool& bad) {
  if constexpr (EXACT) {
    return __fdiv_rn(a, d.b);
  } else {
    float q = __fmul_rn(a, d.y);
    float r = __fmaf_rn(d.nb, q, a);
    q = __fmaf_rn(r, d.y, q);
    r = __fmaf_rn(d.nb, q, a);
    q = __fmaf_rn(r, d.y, q);
    bad = bad || !(fabsf(q) <= 3.0e38f);
    return q;
  }
}

template <bool F16>
ED_DEVICE float round_f16(float v) {
  if constexpr (F16) return __half2float(__float2half_rn(v));
  else return v;
}

// The reference's low-res DDIM x0 of one (cell, channel) (ed:909-921): xl = low-res latent of the last iteration, ul = its
// uncond score, (lun, lco) = uncond / cond scores behind downsampled_direction (ed:688).  F16: fp16 roundings of the
// CUDA-autocast path (fp16 + fp16, 0-dim fp32 tensor * fp16 tensor -> fp16).
template <bool F16, bool EXACT>
ED_DEVICE float half_low_res_x0(float xl, float ul, float lun, float lco, float g, float sb, const DivBy& div_sa, bool& bad) {
  const float dl = round_f16<F16>(__fsub_rn(lco, lun));
  const float gl = round_f16<F16>(__fmul_rn(g, dl));
  const float el = round_f16<F16>(__fadd_rn(ul, gl));               // ed:918
  const float t1 = round_f16<F16>(__fmul_rn(sb, el));
  return half_div<EXACT>(div_sa, __fsub_rn(xl, t1), bad);          // ed:920-921
}

struct HalfScalars {
  float g, sb, sap, sd, rrg_norm, rrg_w;
  DivBy div_sa;
};

ED_DEVICE void st8(float* p, const float v[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// the 2 x 8 pixels of one (b, c): CFG + DDIM (+ RRG) given g * direction of every pixel's OWNER iteration (`gd`, per pixel:
// the two rows of a cell may have different owners) and the per-cell RRG reference; each row is stored as soon as it is
// done (a tile whose fast division went out of range is simply redone and overwritten by the same thread)
template <bool RRG, bool F16, bool EXACT>
ED_DEVICE void half_finish(const HalfScalars& K, const float xin[2][8], const float uu[2][8], const float gd[2][8],
                           const float rx0[4], float* dst, float* dst_x0, int W, bool& bad) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float res[8], x0v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float eps = __fadd_rn(uu[r][e], gd[r][e]);                                       // ed:1031
      const float x0 = half_div<EXACT>(K.div_sa, __fsub_rn(xin[r][e], __fmul_rn(K.sb, eps)), bad);   // DDIM "predicted x_0"
      x0v[e] = x0;
      float v = __fadd_rn(__fmul_rn(K.sap, x0), __fmul_rn(K.sd, eps));                       // x_{t-1}, eta = 0
      if constexpr (RRG) {
        // -d/dx0 [ w * mse(ref_up, x0) ] = -( (2/N) * (x0 - ref) * w )   (mse_loss backward, ed:932-935)
        const float grad = __fmul_rn(__fmul_rn(K.rrg_norm, __fsub_rn(x0, rx0[e >> 1])), K.rrg_w);
        v = __fadd_rn(v, -grad);                                                             // ed:1078
      }
      res[e] = v;
    }
    st8(dst + r * W, res);
    if (dst_x0) st8(dst_x0 + r * W, x0v);
  }
}

// MULTI = false: R1 == 1 (owner == 0 everywhere).  MULTI = true: R1 > 1, the scores of all iterations go through
// per-thread shared-memory slots (dynamic shared memory: blockDim.x*blockDim.y * 2*R1 * 4 * sizeof(OT) bytes).
// grid: x over W/8 column groups, y over H/2 row pairs, z over (b, c) with a grid-stride loop.
template <typename OT, bool MULTI, bool RRG, bool PEER, bool F16>
ED_DEVICE void half_body(const EpiArgs& A, int xg, int yr, uint8_t* smem_slots) {
  const ed_plan_t& P = A.P;
  const ed_step_params_t& S = *A.prm;
  const int R1 = A.R1;
  const int x0c = xg * 8, y0 = yr * 2;
  HalfScalars K;
  K.g = S.guidance; K.sb = S.sqrt_beta_t; K.sap = S.sqrt_alpha_prev; K.sd = S.sqrt_dir;
  K.rrg_norm = S.rrg_norm; K.rrg_w = S.rrg_weight;
  K.div_sa = make_div_by(S.sqrt_alpha_t);
  const int plane = P.dH * P.dW;
  const long long sample_stride = (long long)P.C * plane;
  const OT* __restrict__ out = static_cast<const OT*>(A.unet_out);
  // start of UNet-output sample `sidx`.  PEER: it lives in the buffer of rank sidx / per (read over NVLink)
  auto sample = [&](int sidx) -> const OT* {
    if constexpr (PEER) {
      const int r = sidx / A.per;
      return static_cast<const OT*>(A.peers[r]) + (long long)(sidx - r * A.per) * sample_stride;
    } else {
      return out + (long long)sidx * sample_stride;
    }
  };
  // ---- per-thread references, shared by every (b, c) this thread walks ---------------------------------------------------
  const int view = __ldg(P.vrow_first + y0) * P.nvc + __ldg(P.vcol_first + x0c);          // the single covering window
  const int voff = __ldg(P.vrow_off + y0) * P.dW + __ldg(P.vcol_off + x0c);               // pixel (y0, x0c) in its canvas plane
  const int doff = (P.g_tp + yr) * P.dW + P.g_lp + xg * 4;                                // cells (yr, 4 xg ..) in a canvas plane
  const int first_view_sample = 2 * P.B * R1;
  const int cell0 = yr * P.lw + xg * 4;
  // owner iteration of each of the 2 x 8 pixels (ed:637, 643-644 via ed_owner_map); R1 == 1: iteration 0
  int own[2][8];
  if constexpr (MULTI) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint2 o = __ldg(reinterpret_cast<const uint2*>(A.owner + (long long)(y0 + r) * P.W + x0c));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        own[r][e] = (o.x >> (8 * e)) & 0xff;
        own[r][4 + e] = (o.y >> (8 * e)) & 0xff;
      }
    }
  }
  // RRG: which pixel of each 2x2 cell the LAST iteration picked (ed:612-613, 910), and the owner at the pixel (2r, 2c)
  // nearest-DOWNsampling reads for the cell (ed:688)
  unsigned picks = 0;
  if constexpr (RRG) picks = __ldg(reinterpret_cast<const unsigned*>(A.idx + (long long)(R1 - 1) * P.lh * P.lw + cell0));
  OT* slots = reinterpret_cast<OT*>(smem_slots);   // [2*R1][HALF_THREADS][4]

  // one (b, c) plane of the thread's tile; returns true when a quotient left the fast division's range
  auto plane_tile = [&](int z, auto exact_tag) -> bool {
    constexpr bool EXACT = decltype(exact_tag)::value;
    bool bad = false;
    const int b = z >> 2, c = z & 3;                                                      // C == 4
    const float* lat = A.latent + ((long long)z * P.H + y0) * P.W + x0c;
    float xin[2][8], uu[2][8];
    ld8_ro(lat, xin[0]);
    ld8_ro(lat + P.W, xin[1]);
    const OT* vs = sample(first_view_sample + view * P.B + b) + c * plane + voff;
    ld8_ro(vs, uu[0]);
    ld8_ro(vs + P.dW, uu[1]);
    float gd[2][8], rx0[4];
    if constexpr (!MULTI) {
      float un[4], co[4];
      ld4_ro(sample(b) + c * plane + doff, un);
      ld4_ro(sample(P.B + b) + c * plane + doff, co);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float d = round_f16<F16>(__fsub_rn(co[q], un[q]));                          // ed:440 (fp16 tensor under autocast)
        const float v = round_f16<F16>(__fmul_rn(K.g, d));                                // python float * fp16 tensor -> fp16
        gd[0][2 * q] = gd[0][2 * q + 1] = gd[1][2 * q] = gd[1][2 * q + 1] = v;
        if constexpr (RRG) {
          const unsigned p = (picks >> (8 * q)) & 3u;
          const float top = (p & 1u) ? xin[0][2 * q + 1] : xin[0][2 * q];
          const float bot = (p & 1u) ? xin[1][2 * q + 1] : xin[1][2 * q];
          rx0[q] = half_low_res_x0<F16, EXACT>((p & 2u) ? bot : top, un[q], un[q], co[q], K.g, K.sb, K.div_sa, bad);
        }
      }
    } else {
      // all 2*R1 (iteration, uncond/cond) score vectors of the 4 cells: independent 8-byte (fp32: 16-byte) loads -> this
      // thread's private shared-memory slots (no barrier).  Slot stride between consecutive (iteration, s) vectors is the
      // compile-time constant HALF_THREADS * 4 elements, so once a pixel's owner is folded into its base pointer the
      // (uncond, cond) picks are two loads at immediate offsets.  Within a warp the slots are interleaved (lane l ->
      // slot 2 (l % 16) + l / 16) so that the 16-bit picks of the 32 lanes fall into 32 different banks.
      constexpr int KS = HALF_THREADS * 4;                       // elements between the vectors of ks and ks + 1
      const int tid = threadIdx.y * blockDim.x + threadIdx.x;
      const int slot = (tid & ~31) | ((tid & 15) << 1) | ((tid >> 4) & 1);
      OT* mine = slots + slot * 4;
      // cp.async (LDGSTS): every vector straight from global into its slot, all 2*R1 in flight at once, no staging
      // registers; a load-store-load-store loop would serialise the memory latency of every vector
      for (int ks = 0; ks < 2 * R1; ++ks)
        cp_async<4 * (int)sizeof(OT)>(mine + ks * KS, sample(ks * P.B + b) + c * plane + doff);
      cp_async_wait_all();
      float lun[4], lco[4];                                      // scores of the owner at pixel (2r, 2c) = row 0, even column
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const OT* pk = mine + own[r][e] * (2 * KS) + (e >> 1);  // (owner, uncond) of this pixel's cell; cond is KS further
          const float un = to_f32<OT>(pk[0]), co = to_f32<OT>(pk[KS]);
          if (r == 0 && (e & 1) == 0) {
            lun[e >> 1] = un;
            lco[e >> 1] = co;
          }
          const float d = round_f16<F16>(__fsub_rn(co, un));
          gd[r][e] = round_f16<F16>(__fmul_rn(K.g, d));
        }
      if constexpr (RRG) {
        const OT* last = mine + (R1 - 1) * (2 * KS);             // uncond scores of the last iteration (ed:910-918)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned p = (picks >> (8 * q)) & 3u;
          const float top = (p & 1u) ? xin[0][2 * q + 1] : xin[0][2 * q];
          const float bot = (p & 1u) ? xin[1][2 * q + 1] : xin[1][2 * q];
          rx0[q] = half_low_res_x0<F16, EXACT>((p & 2u) ? bot : top, to_f32<OT>(last[q]), lun[q], lco[q], K.g, K.sb, K.div_sa, bad);
        }
      }
    }
    const long long o = ((long long)z * P.H + y0) * P.W + x0c;
    half_finish<RRG, F16, EXACT>(K, xin, uu, gd, rx0, A.out_latent + o, A.out_x0 ? A.out_x0 + o : nullptr, P.W, bad);
    return bad;
  };
  const int nz = P.B * P.C;
  for (int z = blockIdx.z; z < nz; z += gridDim.z)
    if (plane_tile(z, std::false_type{})) plane_tile(z, std::true_type{});   // rare: redo the tile with IEEE division
}

// The step's flags live in DEVICE memory (a captured launch is replayed with new parameters), so RRG / fp16 semantics are
// CTA-uniform run-time branches into four specialised bodies rather than launch-time template arguments.
template <typename OT, bool MULTI, bool PEER>
__global__ void __launch_bounds__(HALF_THREADS, MULTI ? ED_HALF_MINB_MULTI : ED_HALF_MINB) wave_epilogue_half_kernel(const EpiArgs A) {
  ED_DYN_SMEM(smem_raw);
  const int xg = blockIdx.x * blockDim.x + threadIdx.x;   // group of 8 columns
  const int yr = blockIdx.y * blockDim.y + threadIdx.y;   // row pair = low-res row
  if (xg * 8 >= A.P.W || yr * 2 >= A.P.H) return;
  const int flags = A.prm->flags;
  const bool rrg = (flags & ED_FLAG_RRG) != 0;
  const bool f16 = (flags & ED_FLAG_FP16_SEM) != 0;
  if (rrg) {
    if (f16) half_body<OT, MULTI, true, PEER, true>(A, xg, yr, smem_raw);
    else half_body<OT, MULTI, true, PEER, false>(A, xg, yr, smem_raw);
  } else {
    if (f16) half_body<OT, MULTI, false, PEER, true>(A, xg, yr, smem_raw);
    else half_body<OT, MULTI, false, PEER, false>(A, xg, yr, smem_raw);
  }
}

// Launch geometry shared by the CUDA launcher and the host emulation.
struct HalfCfg {
  bool ok;
  int bx, by, grid_x, grid_y, grid_z;
  size_t smem;
};
static inline HalfCfg half_config(const ed_plan_t& P, int R1, int so) {
  HalfCfg c{};
  if (!(P.flags & ED_PLAN_HALF_FAST) || P.C != 4 || (P.W & 7) || (P.H & 1) || R1 <= 0 || P.B <= 0) return c;
  const int wg = P.W / 8, hr = P.H / 2;
  int bx = 32;
  while (bx > 1 && bx / 2 >= wg) bx /= 2;
  c.bx = bx;
  c.by = HALF_THREADS / bx;
  c.grid_x = (wg + bx - 1) / bx;
  c.grid_y = (hr + c.by - 1) / c.by;
  const long long nz = (long long)P.B * P.C;
  c.grid_z = nz > 65535 ? 65535 : (int)nz;
  c.smem = R1 > 1 ? (size_t)HALF_THREADS * 2 * R1 * 4 * so : 0;
  c.ok = c.smem <= 200 * 1024;
  return c;
}

}  // namespace ed
