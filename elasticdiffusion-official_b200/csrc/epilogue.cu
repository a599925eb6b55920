// Fused wave epilogue (K2 + K5 + K6 [+ K7] [+ K8]) and the stand-alone re-noise kernel (K7).
//
// One pass over the full-resolution latent replaces, per denoise wave (elastic_diffusion.py "ed:N"):
//   first-writer-wins scatter of the view windows          ed:852-861
//   direction = cond - uncond, nearest-up, masked fills     ed:439-440, 634-647 (loop ed:661-681)
//   eps = uncond + g * direction ; DDIM x0 / x_prev         ed:1031-1035 / 1053-1056 (diffusers 0.21.4 DDIM step)
//   undo_step (repaint re-noising)                          ed:692-704
//   reduced-resolution guidance + final add                 ed:886-940, 1078
//
// Arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) in the reference's operation order, so
// that with identical UNet outputs the result is bit-identical to eager fp32 PyTorch on CPU.
#include "common.cuh"

namespace ed {

struct EpiArgs {
  ed_plan_t P;
  const ed_step_params_t* prm;
  const float* latent;
  const void* unet_out;
  const uint8_t* idx;
  const float* noise;
  float* out_latent;
  float* out_x0;
};

// Which resampling iteration wrote target_direction[y, x] last (ed:637: every iteration overwrites where its mask is
// set; ed:643-644: what is still NaN after the last iteration takes the last iteration's value).
__device__ __forceinline__ int owner_iteration(const ed_plan_t& P, const uint8_t* __restrict__ idx, int R1, int y, int x) {
  const int rlo = __ldg(P.mrow_lo + y), rn = __ldg(P.mrow_n + y);
  const int clo = __ldg(P.mcol_lo + x), cn = __ldg(P.mcol_n + x);
  const int cells = P.lh * P.lw;
  for (int k = R1 - 1; k > 0; --k) {
    const uint8_t* t = idx + (long long)k * cells;
    for (int a = 0; a < rn; ++a) {
      const int ry = rlo + a;
      for (int e = 0; e < cn; ++e) {
        const int rx = clo + e;
        if (t[(ry >> 1) * P.lw + (rx >> 1)] == (((ry & 1) << 1) | (rx & 1))) return k;
      }
    }
  }
  // iteration 0 either owns the pixel through its own mask, or nothing does and the NaN back-fill of the last
  // iteration applies.
  if (R1 > 1) {
    const uint8_t* t = idx;
    bool hit = false;
    for (int a = 0; a < rn; ++a)
      for (int e = 0; e < cn; ++e) {
        const int ry = rlo + a, rx = clo + e;
        hit |= t[(ry >> 1) * P.lw + (rx >> 1)] == (((ry & 1) << 1) | (rx & 1));
      }
    return hit ? 0 : R1 - 1;
  }
  return 0;
}

template <typename OT>
__device__ __forceinline__ float direction_at(const ed_plan_t& P, const OT* __restrict__ out, const uint8_t* idx, int R1,
                                              int b, int c, int y, int x, bool fp16sem) {
  const int k = owner_iteration(P, idx, R1, y, x);
  const long long plane = (long long)P.dH * P.dW;
  const long long off = (long long)(P.g_tp + __ldg(P.up_row + y)) * P.dW + P.g_lp + __ldg(P.up_col + x);
  const float un = to_f32<OT>(out[(((long long)(k * 2 + 0) * P.B + b) * P.C + c) * plane + off]);
  const float co = to_f32<OT>(out[(((long long)(k * 2 + 1) * P.B + b) * P.C + c) * plane + off]);
  float d = __fsub_rn(co, un);                                       // ed:440
  if (fp16sem) d = __half2float(__float2half_rn(d));                 // fp16 tensor under CUDA autocast / ed:655
  return d;
}

template <typename OT>
__device__ __forceinline__ float local_uncond_at(const ed_plan_t& P, const OT* __restrict__ out, int first_view_sample,
                                                 int b, int c, int y, int x) {
  const int r0 = __ldg(P.vrow_first + y), rn = __ldg(P.vrow_cnt + y);
  const int c0 = __ldg(P.vcol_first + x), cn = __ldg(P.vcol_cnt + x);
  const long long plane = (long long)P.dH * P.dW;
  float u = 0.f;
  for (int a = 0; a < rn; ++a)
    for (int e = 0; e < cn; ++e) {
      const int v = (r0 + a) * P.nvc + (c0 + e);
      const int32_t* vt = P.views + v * 8;
      const int yy = P.v_tp + vt[6] + (y - vt[0]);
      const int xx = P.v_lp + vt[7] + (x - vt[2]);
      u = to_f32<OT>(out[(((long long)(first_view_sample + v * P.B + b)) * P.C + c) * plane + (long long)yy * P.dW + xx]);
      if (u != 0.f) return u;                                       // first writer wins where the value is non-zero (ed:859)
    }
  return u;
}

template <typename OT, int VEC>
__global__ void __launch_bounds__(256) wave_epilogue_kernel(const EpiArgs A) {
  const ed_plan_t& P = A.P;
  const ed_step_params_t& S = *A.prm;
  const OT* __restrict__ out = static_cast<const OT*>(A.unet_out);
  const int R1 = S.R1;
  const int flags = S.flags;
  const bool fp16sem = (flags & ED_FLAG_FP16_SEM) != 0;
  const float g = S.guidance, sb = S.sqrt_beta_t, sa = S.sqrt_alpha_t, sap = S.sqrt_alpha_prev, sd = S.sqrt_dir;
  const int first_view_sample = 2 * P.B * R1;
  const int wv = P.W / VEC;
  const long long total = (long long)P.B * P.C * P.H * wv;
  const long long numel = (long long)P.B * P.C * P.H * P.W;
  const long long plane = (long long)P.dH * P.dW;

  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xv = (int)(i % wv);
    long long r = i / wv;
    const int y = (int)(r % P.H);
    r /= P.H;
    const int c = (int)(r % P.C);
    const int b = (int)(r / P.C);
    const long long base = (((long long)b * P.C + c) * P.H + y) * P.W + xv * VEC;

    float xin[VEC], res[VEC], x0v[VEC];
    if constexpr (VEC == 4) {
      const float4 t = *reinterpret_cast<const float4*>(A.latent + base);
      xin[0] = t.x; xin[1] = t.y; xin[2] = t.z; xin[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) xin[e] = A.latent[base + e];
    }

#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int x = xv * VEC + e;
      const float u = local_uncond_at<OT>(P, out, first_view_sample, b, c, y, x);
      const float d = direction_at<OT>(P, out, A.idx, R1, b, c, y, x, fp16sem);
      float gd = __fmul_rn(g, d);
      if (fp16sem) gd = __half2float(__float2half_rn(gd));          // python float * fp16 tensor -> fp16
      const float eps = __fadd_rn(u, gd);                            // ed:1031
      const float x0 = __fdiv_rn(__fsub_rn(xin[e], __fmul_rn(sb, eps)), sa);   // DDIM "predicted x_0"
      const float xp = __fadd_rn(__fmul_rn(sap, x0), __fmul_rn(sd, eps));      // x_{t-1}, eta = 0
      x0v[e] = x0;
      res[e] = xp;

      if (flags & ED_FLAG_RRG) {
        // reference low-res x0 of the LAST resampling iteration at the cell that nearest-upsampling reads (ed:909-922)
        const int kl = R1 - 1;
        const int ur = __ldg(P.up_row + y), uc = __ldg(P.up_col + x);
        const int pick = A.idx[((long long)kl * P.lh + ur) * P.lw + uc];
        const int sr = __ldg(P.row_src + 2 * ur + (pick >> 1)), sc = __ldg(P.col_src + 2 * uc + (pick & 1));
        const float xl = A.latent[(((long long)b * P.C + c) * P.H + sr) * P.W + sc];
        const float ul = to_f32<OT>(out[(((long long)(kl * 2) * P.B + b) * P.C + c) * plane +
                                        (long long)(P.g_tp + ur) * P.dW + P.g_lp + uc]);
        // downsampled_direction = nearest-down of the filled full-res direction (ed:688)
        const float dl = direction_at<OT>(P, out, A.idx, R1, b, c, __ldg(P.down_row + ur), __ldg(P.down_col + uc), fp16sem);
        float gl = __fmul_rn(g, dl);
        float el;
        float t1;
        if (fp16sem) {
          gl = __half2float(__float2half_rn(gl));
          el = __half2float(__float2half_rn(__fadd_rn(ul, gl)));     // fp16 + fp16 (ed:918)
          t1 = __half2float(__float2half_rn(__fmul_rn(sb, el)));     // 0-dim fp32 tensor * fp16 tensor -> fp16
        } else {
          el = __fadd_rn(ul, gl);
          t1 = __fmul_rn(sb, el);
        }
        const float ref = __fdiv_rn(__fsub_rn(xl, t1), sa);          // ed:920-921
        // -d/dx0 [ w * mse(ref_up, x0) ] = -( (2/N) * (x0 - ref) * w )   (mse_loss backward, ed:932-935)
        const float grad = __fmul_rn(__fmul_rn(S.rrg_norm, __fsub_rn(x0, ref)), S.rrg_weight);
        res[e] = __fadd_rn(xp, -grad);                               // ed:1078
      }
    }

    if (flags & ED_FLAG_RENOISE) {                                   // ed:692-704, sequential like the reference
      const int n = S.n_renoise;
      for (int k = 0; k < n; ++k) {
        const float a = S.renoise_a[k], bb = S.renoise_b[k];
        float nz[VEC];
        if constexpr (VEC == 4) {
          const float4 t = __ldcs(reinterpret_cast<const float4*>(A.noise + (long long)k * numel + base));
          nz[0] = t.x; nz[1] = t.y; nz[2] = t.z; nz[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) nz[e] = A.noise[(long long)k * numel + base + e];
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) res[e] = __fadd_rn(__fmul_rn(a, res[e]), __fmul_rn(bb, nz[e]));
      }
    }

    if constexpr (VEC == 4) {
      *reinterpret_cast<float4*>(A.out_latent + base) = make_float4(res[0], res[1], res[2], res[3]);
      if (A.out_x0) *reinterpret_cast<float4*>(A.out_x0 + base) = make_float4(x0v[0], x0v[1], x0v[2], x0v[3]);
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        A.out_latent[base + e] = res[e];
        if (A.out_x0) A.out_x0[base + e] = x0v[e];
      }
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) renoise_kernel(const ed_step_params_t* __restrict__ prm, const float* __restrict__ x,
                                                      const float* __restrict__ noise, float* __restrict__ out,
                                                      long long numel) {
  const int n = prm->n_renoise;
  const long long nv = numel / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float v[VEC];
    if constexpr (VEC == 4) {
      const float4 t = reinterpret_cast<const float4*>(x)[i];
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      v[0] = x[i];
    }
    for (int k = 0; k < n; ++k) {
      const float a = prm->renoise_a[k], b = prm->renoise_b[k];
      if constexpr (VEC == 4) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(noise + (long long)k * numel) + i);
        v[0] = __fadd_rn(__fmul_rn(a, v[0]), __fmul_rn(b, t.x));
        v[1] = __fadd_rn(__fmul_rn(a, v[1]), __fmul_rn(b, t.y));
        v[2] = __fadd_rn(__fmul_rn(a, v[2]), __fmul_rn(b, t.z));
        v[3] = __fadd_rn(__fmul_rn(a, v[3]), __fmul_rn(b, t.w));
      } else {
        v[0] = __fadd_rn(__fmul_rn(a, v[0]), __fmul_rn(b, noise[(long long)k * numel + i]));
      }
    }
    if constexpr (VEC == 4) reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
    else out[i] = v[0];
  }
}

static int epi_grid(long long threads) {
  long long g = (threads + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  return g < 1 ? 1 : (int)g;
}

}  // namespace ed

using namespace ed;

extern "C" {

int ed_wave_epilogue(const ed_plan_t* plan, const ed_step_params_t* d_params, const float* latent, const void* unet_out,
                     int out_dtype, const uint8_t* idx, const float* noise, float* out_latent, float* out_x0,
                     void* stream_) {
  if (!plan || !d_params || !latent || !unet_out || !idx || !out_latent) return ED_ERR_INVALID;
  const ed_plan_t& P = *plan;
  if (!P.mrow_lo || !P.mrow_n || !P.mcol_lo || !P.mcol_n || !P.up_row || !P.up_col || !P.down_row || !P.down_col ||
      !P.views || !P.vrow_first || !P.vrow_cnt || !P.vcol_first || !P.vcol_cnt || !P.row_src || !P.col_src)
    return ED_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EpiArgs A{P, d_params, latent, unet_out, idx, noise, out_latent, out_x0};
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = (P.W % 4 == 0) && aligned(latent) && aligned(out_latent) && (!out_x0 || aligned(out_x0)) &&
                   (!noise || aligned(noise));
  const long long total = (long long)P.B * P.C * P.H * (vec ? P.W / 4 : P.W);
  const int g = epi_grid(total);
#define ED_EPI(T)                                                     \
  if (vec) wave_epilogue_kernel<T, 4><<<g, 256, 0, stream>>>(A);       \
  else wave_epilogue_kernel<T, 1><<<g, 256, 0, stream>>>(A);
  switch (out_dtype) {
    case ED_F32: ED_EPI(float) break;
    case ED_F16: ED_EPI(__half) break;
    case ED_BF16: ED_EPI(__nv_bfloat16) break;
    default: return ED_ERR_INVALID;
  }
#undef ED_EPI
  ED_LAUNCH_CHECK();
  return ED_OK;
}

int ed_renoise(const ed_step_params_t* d_params, const float* x, const float* noise, float* out, int64_t numel,
               void* stream_) {
  if (!d_params || !x || !noise || !out || numel <= 0) return ED_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = (numel % 4 == 0) && aligned(x) && aligned(noise) && aligned(out);
  if (vec) renoise_kernel<4><<<epi_grid(numel / 4), 256, 0, stream>>>(d_params, x, noise, out, numel);
  else renoise_kernel<1><<<epi_grid(numel), 256, 0, stream>>>(d_params, x, noise, out, numel);
  ED_LAUNCH_CHECK();
  return ED_OK;
}

}  // extern "C"
