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
#include <stdlib.h>

#include <atomic>

#include "epilogue_half.cuh"
#include "epilogue_staged.cuh"

#ifndef ED_EPI_MINB
#define ED_EPI_MINB 2   // CTAs of 256 threads per SM the epilogue is compiled for (register cap 65536 / (256 * MINB))
#endif

namespace ed {

// EpiArgs: epilogue_staged.cuh

static std::atomic<int> g_epilogue_mode{ED_EPILOGUE_AUTO};
static std::atomic<long long> g_launches_direct{0}, g_launches_staged{0}, g_launches_half{0};

// Which resampling iteration wrote target_direction[y, x] last (ed:637: every iteration overwrites where its mask is
// set; ed:643-644: what is still NaN after the last iteration takes the last iteration's value).  All idx bytes are
// loaded unconditionally (independent loads, one memory round trip) and the owner is the highest k whose mask hits.
__device__ __forceinline__ int owner_iteration(const ed_plan_t& P, const uint8_t* __restrict__ idx, int R1, int y, int x) {
  const int rlo = __ldg(P.mrow_lo + y), rn = __ldg(P.mrow_n + y);
  const int clo = __ldg(P.mcol_lo + x), cn = __ldg(P.mcol_n + x);
  const int cells = P.lh * P.lw;
  int owner = -1;
  if (rn == 1 && cn == 1) {   // exact 1/2 (or identity) ratio: one cell, one code
    const int cell = (rlo >> 1) * P.lw + (clo >> 1);
    const int code = ((rlo & 1) << 1) | (clo & 1);
#pragma unroll 4
    for (int k = 0; k < R1; ++k) owner = (__ldg(idx + k * cells + cell) == code) ? k : owner;
  } else {
    for (int k = 0; k < R1; ++k) {
      const uint8_t* t = idx + k * cells;
      bool hit = false;
      for (int a = 0; a < rn; ++a)
        for (int e = 0; e < cn; ++e) {
          const int ry = rlo + a, rx = clo + e;
          hit |= __ldg(t + (ry >> 1) * P.lw + (rx >> 1)) == (((ry & 1) << 1) | (rx & 1));
        }
      owner = hit ? k : owner;
    }
  }
  return owner < 0 ? R1 - 1 : owner;
}

// Per-pixel, channel-independent references.  Everything static comes from the host-built tables of the plan
// (pix_ref / cell_cand / cell_down); the only per-wave inputs are the owner map and the last iteration's pick bytes.
struct PixelRefs {
  int dir_k, dir_off;     // owner iteration of target_direction here + offset of the low-res cell inside a canvas plane
  int view, view_off;     // single covering view and the pixel's offset in its canvas plane; view = -1: several windows
                          // cover the pixel (overlapping last row / column) -> table walk per channel
  int lat_off;            // RRG: offset in a latent plane of the pixel the LAST iteration picked for that low-res cell
  int ddir_k, ddir_off;   // RRG: owner + cell offset at the full-res pixel nearest-DOWNsampling reads for that cell
};

__device__ __forceinline__ void pixel_refs(const ed_plan_t& P, const uint8_t* __restrict__ idx,
                                           const uint8_t* __restrict__ owner, int R1, int pix, bool rrg, PixelRefs& r) {
  const int4 s = __ldg(reinterpret_cast<const int4*>(P.pix_ref) + pix);
  r.dir_k = __ldg(owner + pix);
  r.dir_off = s.x;
  r.view = s.y;
  r.view_off = s.z;
  if (rrg) {
    const int cell = s.w;
    const int pick = __ldg(idx + (R1 - 1) * P.lh * P.lw + cell) & 3;
    r.lat_off = __ldg(P.cell_cand + cell * 4 + pick);
    const int2 d = __ldg(reinterpret_cast<const int2*>(P.cell_down) + cell);
    r.ddir_k = __ldg(owner + d.x);
    r.ddir_off = d.y;
  }
}

__global__ void __launch_bounds__(256) owner_map_kernel(const ed_plan_t P, int R1, const uint8_t* __restrict__ idx,
                                                        uint8_t* __restrict__ owner) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < P.W && y < P.H) owner[y * P.W + x] = (uint8_t)owner_iteration(P, idx, R1, y, x);
}

template <typename OT>
__device__ __forceinline__ float direction_val(const OT* __restrict__ out_bc, long long sample_stride, int B, int k, int off,
                                               bool fp16sem) {
  // out_bc points at channel c of batch entry b of sample 0; sample s lives s*sample_stride further
  const float un = ld_ro<OT>(out_bc + (long long)(k * 2 + 0) * B * sample_stride + off);
  const float co = ld_ro<OT>(out_bc + (long long)(k * 2 + 1) * B * sample_stride + off);
  float d = __fsub_rn(co, un);                                       // ed:440
  if (fp16sem) d = __half2float(__float2half_rn(d));                 // fp16 tensor under CUDA autocast / ed:655
  return d;
}

struct ViewWalk {   // by-value subset of the plan for the rare multi-cover path (keeps the kernel parameter out of local memory)
  const int32_t *views, *vrow_first, *vrow_cnt, *vcol_first, *vcol_cnt;
  int nvc, v_tp, v_lp, dW, B;
};

template <typename OT>
__device__ __forceinline__ float local_uncond_walk(const ViewWalk V, const OT* __restrict__ out_bc, long long sample_stride,
                                                int first_view_sample, int y, int x) {
  const int r0 = __ldg(V.vrow_first + y), rn = __ldg(V.vrow_cnt + y);
  const int c0 = __ldg(V.vcol_first + x), cn = __ldg(V.vcol_cnt + x);
  float u = 0.f;
  for (int a = 0; a < rn; ++a)
    for (int e = 0; e < cn; ++e) {
      const int v = (r0 + a) * V.nvc + (c0 + e);
      const int32_t* vt = V.views + v * 8;
      const int yy = V.v_tp + __ldg(vt + 6) + (y - __ldg(vt + 0));
      const int xx = V.v_lp + __ldg(vt + 7) + (x - __ldg(vt + 2));
      u = ld_ro<OT>(out_bc + (long long)(first_view_sample + v * V.B) * sample_stride + (long long)yy * V.dW + xx);
      if (u != 0.f) return u;                                       // first writer wins where the value is non-zero (ed:859)
    }
  return u;
}

// grid: x over W/VEC, y over H, z over (b, channel group); thread = VEC consecutive pixels of one row, CPT channels.
// CPT = 4 (all channels of an SD latent in one thread) is the throughput shape: the per-pixel refs are amortised over
// the channels, the prologue loads of all channels are independent (read-only path, results kept in registers, stores
// at the end), and the re-noise stream is fetched 4 channels x 4 steps = 16 float4 loads deep before the dependent
// FMA chain consumes them.  CPT = 1 spreads the channels over gridDim.z (small batches: more CTAs, 8 loads deep).
template <typename OT, int VEC, int CPT, bool PEER>
__global__ void __launch_bounds__(256, ED_EPI_MINB) wave_epilogue_kernel(const EpiArgs A) {
  const ed_plan_t& P = A.P;
  const int xv = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (xv * VEC >= P.W || y >= P.H) return;
  const ed_step_params_t& S = *A.prm;
  const OT* __restrict__ out = static_cast<const OT*>(A.unet_out);
  const int R1 = A.R1;
  const int flags = S.flags;
  const bool fp16sem = (flags & ED_FLAG_FP16_SEM) != 0;
  const bool rrg = (flags & ED_FLAG_RRG) != 0;
  const float g = S.guidance, sb = S.sqrt_beta_t, sa = S.sqrt_alpha_t, sap = S.sqrt_alpha_prev, sd = S.sqrt_dir;
  const float rrg_norm = S.rrg_norm, rrg_w = S.rrg_weight;
  const int n_re = (flags & ED_FLAG_RENOISE) ? S.n_renoise : 0;
  const int first_view_sample = 2 * P.B * R1;
  const long long hw = (long long)P.H * P.W;
  const long long numel = (long long)P.B * P.C * hw;
  const long long plane = (long long)P.dH * P.dW;
  const long long sample_stride = (long long)P.C * plane;
  constexpr int KB = 16 / CPT;                                         // re-noise steps fetched per batch
  // start of UNet-output sample `sidx` (all C channels).  PEER: the sample lives in the symmetric buffer of rank
  // sidx / per and is read over NVLink with plain P2P loads (no all-gather ran).
  auto sample = [&](int sidx) -> const OT* {
    if constexpr (PEER) {
      const int r = sidx / A.per;
      return static_cast<const OT*>(A.peers[r]) + (long long)(sidx - r * A.per) * sample_stride;
    } else {
      return out + (long long)sidx * sample_stride;
    }
  };

  PixelRefs ref[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) pixel_refs(P, A.idx, A.owner, R1, y * P.W + xv * VEC + e, rrg, ref[e]);

  const int groups = P.C / CPT;
  for (int z = blockIdx.z; z < P.B * groups; z += gridDim.z) {
    const int b = z / groups, c_lo = (z - b * groups) * CPT;
    const long long base0 = (((long long)b * P.C + c_lo) * P.H + y) * P.W + xv * VEC;   // channel c_lo; + cc*hw per channel
    float res[CPT][VEC];
    // ---- phase A: issue every load of the prologue (read-only path, all independent) before any arithmetic ------------
    float xin[CPT][VEC], uu[CPT][VEC], dco[CPT][VEC], dun[CPT][VEC];
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      const long long ch_off = (long long)(c_lo + cc) * plane;            // channel offset inside a sample
      if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(A.latent + base0 + cc * hw));
        xin[cc][0] = t.x; xin[cc][1] = t.y; xin[cc][2] = t.z; xin[cc][3] = t.w;
      } else {
        xin[cc][0] = __ldg(A.latent + base0 + cc * hw);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        // pixels covered by several windows (view < 0, overlapping last row / column) are patched after the load phase
        const int v_ = ref[e].view >= 0 ? ref[e].view : 0;
        uu[cc][e] = ld_ro<OT>(sample(first_view_sample + v_ * P.B + b) + ch_off + ref[e].view_off);
        dun[cc][e] = ld_ro<OT>(sample((ref[e].dir_k * 2 + 0) * P.B + b) + ch_off + ref[e].dir_off);
        dco[cc][e] = ld_ro<OT>(sample((ref[e].dir_k * 2 + 1) * P.B + b) + ch_off + ref[e].dir_off);
      }
    }
    bool multi = false;
#pragma unroll
    for (int e = 0; e < VEC; ++e) multi |= ref[e].view < 0;
    if (multi) {   // rare: first-writer-wins walk over the covering windows (ed:852-861)
      const ViewWalk V{P.views, P.vrow_first, P.vrow_cnt, P.vcol_first, P.vcol_cnt, P.nvc, P.v_tp, P.v_lp, P.dW, P.B};
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (ref[e].view < 0) {
            // first-writer-wins walk (ed:852-861), sample pointers resolved through sample() (peer-aware)
            const int x = xv * VEC + e;
            const int r0 = __ldg(V.vrow_first + y), rn = __ldg(V.vrow_cnt + y);
            const int c0 = __ldg(V.vcol_first + x), cn = __ldg(V.vcol_cnt + x);
            float u = 0.f;
            bool done = false;
            for (int a = 0; a < rn && !done; ++a)
              for (int q = 0; q < cn && !done; ++q) {
                const int v = (r0 + a) * V.nvc + (c0 + q);
                const int32_t* vt = V.views + v * 8;
                const int yy = V.v_tp + __ldg(vt + 6) + (y - __ldg(vt + 0));
                const int xx = V.v_lp + __ldg(vt + 7) + (x - __ldg(vt + 2));
                u = ld_ro<OT>(sample(first_view_sample + v * V.B + b) + (long long)(c_lo + cc) * plane + (long long)yy * V.dW + xx);
                done = (u != 0.f);                                       // first writer wins where the value is non-zero
              }
            uu[cc][e] = u;
          }
    }
    // ---- phase B: CFG + DDIM (ed:1031-1035) ------------------------------------------------------------------------------
    float x0v[CPT][VEC];
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float d = __fsub_rn(dco[cc][e], dun[cc][e]);                   // ed:440
        if (fp16sem) d = __half2float(__float2half_rn(d));             // fp16 tensor under CUDA autocast / ed:655
        float gd = __fmul_rn(g, d);
        if (fp16sem) gd = __half2float(__float2half_rn(gd));           // python float * fp16 tensor -> fp16
        const float eps = __fadd_rn(uu[cc][e], gd);                    // ed:1031
        const float x0 = __fdiv_rn(__fsub_rn(xin[cc][e], __fmul_rn(sb, eps)), sa);   // DDIM "predicted x_0"
        x0v[cc][e] = x0;
        res[cc][e] = __fadd_rn(__fmul_rn(sap, x0), __fmul_rn(sd, eps));              // x_{t-1}, eta = 0
      }
    if (A.out_x0) {
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) {
        if constexpr (VEC == 4)
          *reinterpret_cast<float4*>(A.out_x0 + base0 + cc * hw) = make_float4(x0v[cc][0], x0v[cc][1], x0v[cc][2], x0v[cc][3]);
        else
          A.out_x0[base0 + cc * hw] = x0v[cc][0];
      }
    }
    // ---- RRG (ed:886-940, 1078): second load phase (low-res reference of the last iteration), then the gradient -------
    if (rrg) {
      const int kl = R1 - 1;
      float xl[CPT][VEC], ul[CPT][VEC], lco[CPT][VEC], lun[CPT][VEC];
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) {
        const float* lat_plane = A.latent + ((long long)b * P.C + c_lo + cc) * hw;
        const long long ch_off = (long long)(c_lo + cc) * plane;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          xl[cc][e] = __ldg(lat_plane + ref[e].lat_off);               // low-res latent of the last iteration (ed:910)
          ul[cc][e] = ld_ro<OT>(sample(kl * 2 * P.B + b) + ch_off + ref[e].dir_off);   // its uncond score
          // downsampled_direction = nearest-down of the filled full-res direction (ed:688)
          lun[cc][e] = ld_ro<OT>(sample((ref[e].ddir_k * 2 + 0) * P.B + b) + ch_off + ref[e].ddir_off);
          lco[cc][e] = ld_ro<OT>(sample((ref[e].ddir_k * 2 + 1) * P.B + b) + ch_off + ref[e].ddir_off);
        }
      }
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float dl = __fsub_rn(lco[cc][e], lun[cc][e]);
          if (fp16sem) dl = __half2float(__float2half_rn(dl));
          float gl = __fmul_rn(g, dl);
          float el, t1;
          if (fp16sem) {
            gl = __half2float(__float2half_rn(gl));
            el = __half2float(__float2half_rn(__fadd_rn(ul[cc][e], gl)));   // fp16 + fp16 (ed:918)
            t1 = __half2float(__float2half_rn(__fmul_rn(sb, el)));          // 0-dim fp32 tensor * fp16 tensor -> fp16
          } else {
            el = __fadd_rn(ul[cc][e], gl);
            t1 = __fmul_rn(sb, el);
          }
          const float rx0 = __fdiv_rn(__fsub_rn(xl[cc][e], t1), sa);        // ed:920-921
          // -d/dx0 [ w * mse(ref_up, x0) ] = -( (2/N) * (x0 - ref) * w )   (mse_loss backward, ed:932-935)
          const float grad = __fmul_rn(__fmul_rn(rrg_norm, __fsub_rn(x0v[cc][e], rx0)), rrg_w);
          res[cc][e] = __fadd_rn(res[cc][e], -grad);                        // ed:1078
        }
    }
    // ed:692-704: x <- a_k x + b_k eps_k, sequential in k like the reference.  Noise is streamed once (evict-first).
    if (n_re > 0) {
      const float* nz = A.noise + base0;
      int k = 0;
      if constexpr (VEC == 4) {
        for (; k + KB <= n_re; k += KB) {
          float4 t[CPT][KB];
#pragma unroll
          for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
            for (int j = 0; j < KB; ++j)
              t[cc][j] = __ldcs(reinterpret_cast<const float4*>(nz + (long long)(k + j) * numel + cc * hw));
#pragma unroll
          for (int j = 0; j < KB; ++j) {
            const float a = S.renoise_a[k + j], bb = S.renoise_b[k + j];
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
              res[cc][0] = __fadd_rn(__fmul_rn(a, res[cc][0]), __fmul_rn(bb, t[cc][j].x));
              res[cc][1] = __fadd_rn(__fmul_rn(a, res[cc][1]), __fmul_rn(bb, t[cc][j].y));
              res[cc][2] = __fadd_rn(__fmul_rn(a, res[cc][2]), __fmul_rn(bb, t[cc][j].z));
              res[cc][3] = __fadd_rn(__fmul_rn(a, res[cc][3]), __fmul_rn(bb, t[cc][j].w));
            }
          }
        }
      }
      for (; k < n_re; ++k) {
        const float a = S.renoise_a[k], bb = S.renoise_b[k];
#pragma unroll
        for (int cc = 0; cc < CPT; ++cc) {
          if constexpr (VEC == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(nz + (long long)k * numel + cc * hw));
            res[cc][0] = __fadd_rn(__fmul_rn(a, res[cc][0]), __fmul_rn(bb, t.x));
            res[cc][1] = __fadd_rn(__fmul_rn(a, res[cc][1]), __fmul_rn(bb, t.y));
            res[cc][2] = __fadd_rn(__fmul_rn(a, res[cc][2]), __fmul_rn(bb, t.z));
            res[cc][3] = __fadd_rn(__fmul_rn(a, res[cc][3]), __fmul_rn(bb, t.w));
          } else {
            res[cc][0] = __fadd_rn(__fmul_rn(a, res[cc][0]), __fmul_rn(bb, __ldcs(nz + (long long)k * numel + cc * hw)));
          }
        }
      }
    }
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      if constexpr (VEC == 4)
        *reinterpret_cast<float4*>(A.out_latent + base0 + cc * hw) = make_float4(res[cc][0], res[cc][1], res[cc][2], res[cc][3]);
      else
        A.out_latent[base0 + cc * hw] = res[cc][0];
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

// Tile-staged kernel (epilogue_staged.cuh).  ED_ERR_UNSUPPORTED: shape / alignment outside what it handles - the caller
// then runs the direct kernel.
template <typename OT, bool RENOISE, int CPT>
static int launch_staged_as(const EpiArgs& A, int out_dtype, int sms, int origin, cudaStream_t stream) {
  const ed_plan_t& P = A.P;
  static const int max_threads = getenv("ED_STAGED_THREADS") ? atoi(getenv("ED_STAGED_THREADS")) : 128;   // measured: 128-thread CTAs are 2-6 % faster than 256
  StagedCfg cfg = staged_config(P, A.R1, (int)sizeof(OT), sms, origin, CPT, max_threads);
  if (!cfg.ok) return ED_ERR_UNSUPPORTED;
  const long long n_samples = 2LL * P.B * A.R1 + (long long)P.nv * P.B;
  CUtensorMap tm;
  // all samples of the wave as (dW, dH, samples * C) planes; box = cells x rows x CPT channels of one sample
  int rc = encode_tmap_3d(&tm, A.unet_out, out_dtype, (uint64_t)P.dW, (uint64_t)P.dH, (uint64_t)n_samples * P.C,
                          (uint32_t)cfg.g.bw, (uint32_t)cfg.g.bh, (uint32_t)CPT);
  if (rc != ED_OK) return rc;
  cfg.g.vec_views = 1;   // encode_tmap_3d checked the 16-byte alignment of unet_out; dH*dW*sizeof(OT) is a multiple of 16
  // per launch: the opt-in is per DEVICE (a process-wide "already set" flag breaks the second GPU of a process); cheap
  ED_CUDA_CHECK(cudaFuncSetAttribute(wave_epilogue_staged_kernel<OT, RENOISE, CPT>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const dim3 block(cfg.g.bx, cfg.g.by), grid(cfg.grid_x, cfg.grid_y, cfg.grid_z);
  wave_epilogue_staged_kernel<OT, RENOISE, CPT><<<grid, block, cfg.smem, stream>>>(tm, A, cfg.g);
  ED_LAUNCH_CHECK();
  return ED_OK;
}

template <typename OT>
static int launch_staged(const EpiArgs& A, int out_dtype, cudaStream_t stream) {
  int sms = 0;
  if (int rc = current_sm_count(&sms)) return rc;
  static const int origin = getenv("ED_STAGED_ORIGIN") ? atoi(getenv("ED_STAGED_ORIGIN")) : (ED_BOX_ALIGN | ED_BOX_CLAMP);
  static const int cpt_plain = getenv("ED_STAGED_CPT") ? atoi(getenv("ED_STAGED_CPT")) : 2;
  // with a noise buffer: the re-noise stream keeps the launch DRAM-bound, 4 channels per thread amortise the per-pixel
  // work best.  Without one (wave 2, no-repaint steps) the re-noise flag could not be honoured anyway: the lighter
  // instantiation, channel pairs spread over the grid for twice the resident warps
  if (A.noise) return launch_staged_as<OT, true, 4>(A, out_dtype, sms, origin, stream);
  if (cpt_plain == 4) return launch_staged_as<OT, false, 4>(A, out_dtype, sms, origin, stream);
  return launch_staged_as<OT, false, 2>(A, out_dtype, sms, origin, stream);
}

// Half kernels (epilogue_half.cuh): plan flag ED_PLAN_HALF_FAST, no noise stream.  ED_ERR_UNSUPPORTED: outside their domain.
template <typename OT, bool MULTI>
static int launch_half_as(const EpiArgs& A, const HalfCfg& cfg, cudaStream_t stream) {
  const dim3 block(cfg.bx, cfg.by), grid(cfg.grid_x, cfg.grid_y, cfg.grid_z);
  if (A.peers) {
    if (cfg.smem > 48 * 1024)
      ED_CUDA_CHECK(cudaFuncSetAttribute(wave_epilogue_half_kernel<OT, MULTI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    wave_epilogue_half_kernel<OT, MULTI, true><<<grid, block, cfg.smem, stream>>>(A);
  } else {
    if (cfg.smem > 48 * 1024)
      ED_CUDA_CHECK(cudaFuncSetAttribute(wave_epilogue_half_kernel<OT, MULTI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    wave_epilogue_half_kernel<OT, MULTI, false><<<grid, block, cfg.smem, stream>>>(A);
  }
  ED_LAUNCH_CHECK();
  return ED_OK;
}

template <typename OT>
static int launch_half(const EpiArgs& A, cudaStream_t stream) {
  const HalfCfg cfg = half_config(A.P, A.R1, (int)sizeof(OT));
  if (!cfg.ok || A.noise) return ED_ERR_UNSUPPORTED;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  // peer buffers come from symmetric-memory allocations (>= 256-byte aligned); the local pointers are checked here
  if (!aligned(A.latent) || !aligned(A.out_latent) || (A.out_x0 && !aligned(A.out_x0)) || (A.unet_out && !aligned(A.unet_out)) ||
      (A.R1 > 1 && (reinterpret_cast<uintptr_t>(A.owner) & 7)) || (reinterpret_cast<uintptr_t>(A.idx) & 3) ||
      (((long long)(A.R1 - 1) * A.P.lh * A.P.lw) & 3))
    return ED_ERR_UNSUPPORTED;
  return A.R1 > 1 ? launch_half_as<OT, true>(A, cfg, stream) : launch_half_as<OT, false>(A, cfg, stream);
}

extern "C" {

int ed_owner_map(const ed_plan_t* plan, int R1, const uint8_t* idx, uint8_t* owner, void* stream_) {
  if (!plan || !idx || !owner || R1 <= 0 || R1 > 255) return ED_ERR_INVALID;
  const ed_plan_t& P = *plan;
  if (!P.mrow_lo || !P.mrow_n || !P.mcol_lo || !P.mcol_n) return ED_ERR_INVALID;
  const dim3 block(32, 8), grid((P.W + 31) / 32, (P.H + 7) / 8);
  owner_map_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream_)>>>(P, R1, idx, owner);
  ED_LAUNCH_CHECK();
  return ED_OK;
}

static int launch_epilogue(const ed_plan_t* plan, const ed_step_params_t* d_params, int R1, const float* latent,
                           const void* unet_out, const void* const* peers, int world, int per, int out_dtype,
                           const uint8_t* idx, const uint8_t* owner, const float* noise, float* out_latent, float* out_x0,
                           void* stream_) {
  if (!plan || !d_params || !latent || (!unet_out && !peers) || !idx || !owner || !out_latent) return ED_ERR_INVALID;
  if (R1 <= 0 || R1 > 255) return ED_ERR_INVALID;
  if (peers && (world <= 0 || per <= 0)) return ED_ERR_INVALID;
  const ed_plan_t& P = *plan;
  if (!P.mrow_lo || !P.mrow_n || !P.mcol_lo || !P.mcol_n || !P.up_row || !P.up_col || !P.down_row || !P.down_col ||
      !P.views || !P.vrow_first || !P.vrow_cnt || !P.vcol_first || !P.vcol_cnt || !P.row_src || !P.col_src ||
      !P.pix_ref || !P.cell_cand || !P.cell_down || !P.vrow_off || !P.vcol_off)
    return ED_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // channels per thread: 4 when the batch alone fills the GPU, 1 (channels spread over gridDim.z) for small batches
  const int cpt = (P.C % 4 == 0 && P.B >= 8 && !peers) ? 4 : 1;
  const int c_split = P.C / cpt;
  EpiArgs A{P, d_params, latent, unet_out, peers, world, per, idx, owner, noise, out_latent, out_x0, R1};
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = (P.W % 4 == 0) && aligned(latent) && aligned(out_latent) && (!out_x0 || aligned(out_x0)) &&
                   (!noise || aligned(noise));
  const int mode = g_epilogue_mode.load();
  if (mode == ED_EPILOGUE_AUTO || mode == ED_EPILOGUE_HALF) {
    int rc = ED_ERR_UNSUPPORTED;
    switch (out_dtype) {
      case ED_F32: rc = launch_half<float>(A, stream); break;
      case ED_F16: rc = launch_half<__half>(A, stream); break;
      case ED_BF16: rc = launch_half<__nv_bfloat16>(A, stream); break;
      default: return ED_ERR_INVALID;
    }
    if (rc == ED_OK) g_launches_half.fetch_add(1);
    if (rc != ED_ERR_UNSUPPORTED || mode == ED_EPILOGUE_HALF) return rc;   // forced: no silent fall-back
  }
  if (mode != ED_EPILOGUE_DIRECT) {
    int rc = ED_ERR_UNSUPPORTED;
    // AUTO: the staged kernel is the throughput shape (a thread carries 4 pixels x 4 channels behind a TMA wait); small
    // launches - one or a few latents, everything L2-resident, latency-bound - keep the direct kernel, whose channels are
    // spread over more and shorter threads (measured in the cfg3 pipeline at B = 1: 10 us vs 17 us per launch)
    const bool big = (long long)(P.W / 4) * P.H * P.B >= 2LL * 148 * 256;
    if (!peers && vec && (big || mode == ED_EPILOGUE_STAGED) && (reinterpret_cast<uintptr_t>(owner) & 3) == 0) {
      switch (out_dtype) {
        case ED_F32: rc = launch_staged<float>(A, out_dtype, stream); break;
        case ED_F16: rc = launch_staged<__half>(A, out_dtype, stream); break;
        case ED_BF16: rc = launch_staged<__nv_bfloat16>(A, out_dtype, stream); break;
        default: return ED_ERR_INVALID;
      }
    }
    if (rc == ED_OK) g_launches_staged.fetch_add(1);
    if (rc != ED_ERR_UNSUPPORTED) return rc;
    if (mode == ED_EPILOGUE_STAGED) return ED_ERR_UNSUPPORTED;   // forced: do not silently take the other kernel
  }
  const int wv = vec ? P.W / 4 : P.W;
  int bx = 32;
  while (bx > 1 && bx / 2 >= wv) bx /= 2;                 // narrow rows: fewer idle lanes
  const dim3 block(bx, 256 / bx);
  const dim3 grid((wv + block.x - 1) / block.x, (P.H + block.y - 1) / block.y,
                  P.B * c_split > 65535 ? 65535 : P.B * c_split);
#define ED_EPI(T)                                                                                   \
  if (peers) {                                                                                      \
    if (vec) wave_epilogue_kernel<T, 4, 1, true><<<grid, block, 0, stream>>>(A);                     \
    else wave_epilogue_kernel<T, 1, 1, true><<<grid, block, 0, stream>>>(A);                         \
  } else if (vec && cpt == 4) wave_epilogue_kernel<T, 4, 4, false><<<grid, block, 0, stream>>>(A);   \
  else if (vec) wave_epilogue_kernel<T, 4, 1, false><<<grid, block, 0, stream>>>(A);                 \
  else if (cpt == 4) wave_epilogue_kernel<T, 1, 4, false><<<grid, block, 0, stream>>>(A);            \
  else wave_epilogue_kernel<T, 1, 1, false><<<grid, block, 0, stream>>>(A);
  switch (out_dtype) {
    case ED_F32: ED_EPI(float) break;
    case ED_F16: ED_EPI(__half) break;
    case ED_BF16: ED_EPI(__nv_bfloat16) break;
    default: return ED_ERR_INVALID;
  }
#undef ED_EPI
  ED_LAUNCH_CHECK();
  g_launches_direct.fetch_add(1);
  return ED_OK;
}

int ed_wave_epilogue(const ed_plan_t* plan, const ed_step_params_t* d_params, int R1, const float* latent,
                     const void* unet_out, int out_dtype, const uint8_t* idx, const uint8_t* owner, const float* noise,
                     float* out_latent, float* out_x0, void* stream_) {
  return launch_epilogue(plan, d_params, R1, latent, unet_out, nullptr, 0, 0, out_dtype, idx, owner, noise, out_latent,
                         out_x0, stream_);
}

int ed_wave_epilogue_peer(const ed_plan_t* plan, const ed_step_params_t* d_params, int R1, const float* latent,
                          const void* const* d_peer_out, int world, int per, int out_dtype, const uint8_t* idx,
                          const uint8_t* owner, const float* noise, float* out_latent, float* out_x0, void* stream_) {
  return launch_epilogue(plan, d_params, R1, latent, nullptr, d_peer_out, world, per, out_dtype, idx, owner, noise,
                         out_latent, out_x0, stream_);
}

int ed_epilogue_launch_counts(int64_t* direct, int64_t* staged) {
  if (direct) *direct = g_launches_direct.load();
  if (staged) *staged = g_launches_staged.load();
  return ED_OK;
}

int ed_epilogue_launch_counts3(int64_t* direct, int64_t* staged, int64_t* half) {
  if (half) *half = g_launches_half.load();
  return ed_epilogue_launch_counts(direct, staged);
}

int ed_set_epilogue_mode(int mode) {
  if (mode < ED_EPILOGUE_AUTO || mode > ED_EPILOGUE_HALF) return ED_ERR_INVALID;
  g_epilogue_mode.store(mode);
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
