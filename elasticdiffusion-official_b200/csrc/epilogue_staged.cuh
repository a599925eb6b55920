// Tile-staged fused wave epilogue (K2 + K5 + K6 [+ K7] [+ K8]) - the throughput shape of ed_wave_epilogue.
//
// Same contract as the direct kernel in epilogue.cu (reference elastic_diffusion.py "ed:N": scatter ed:852-861, direction
// fills ed:439-440/634-647, CFG + DDIM ed:1031-1035, undo_step ed:692-704, RRG ed:886-940 + ed:1078), different data
// movement.  Which resampling iteration "owns" a full-res pixel is random per pixel, so the direct kernel issues one
// scattered 2/4-byte load per (pixel, channel, cond/uncond) into R+1 different UNet output samples - every 32-byte sector
// of every iteration's output is touched anyway, by a different lane each time.  Here a CTA owns a tile of
// `by` rows x `4*bx` columns of the latent; the low-res cells that nearest-upsampling reads for that tile form a rectangle
// (up_row / up_col are non-decreasing), and ONE elected thread pulls that rectangle of ALL 2(R+1) global-pass outputs
// into shared memory with TMA box loads (cp.async.bulk.tensor.3d: box = cells x rows x channels; the box starts at a
// 16-byte aligned column - measured: UTMALDG faults on an unaligned start address - and is shifted back inside the canvas
// plane).  While the boxes are in flight every thread issues its own coalesced loads (latent, per-pixel references, view
// windows); after the mbarrier flips, the per-pixel picks are shared-memory reads.  With RRG the low-res reference x0 is
// evaluated once per (cell, channel) of the rectangle and shared through shared memory.
//
// This header is compiled twice: by nvcc into libelastic_b200.so, and by g++ with tests/emu/emu_shim.h (ED_HOST_EMU) into
// a host emulation that the CPU test-suite checks against oracle/wave_spec.py - test infrastructure only, never loaded by
// the package.
#pragma once
#include <type_traits>
#ifdef ED_HOST_EMU
#include "emu_shim.h"
#else
#include "common.cuh"
#define ED_EMU_COUNT(i)   // host emulation only: path-coverage counters
#endif

#ifndef ED_STAGED_MINB
#define ED_STAGED_MINB 2   // CTAs of 256 threads per SM the staged kernel is compiled for (register cap 65536 / (256 * MINB))
#endif
#ifndef ED_STAGED_MINB_PLAIN
#define ED_STAGED_MINB_PLAIN 3
#endif

namespace ed {

struct EpiArgs {
  ed_plan_t P;
  const ed_step_params_t* prm;
  const float* latent;
  const void* unet_out;
  const void* const* peers;   // multi-GPU: device array of `world` base pointers (peer-mapped), sample s lives on rank
  int world, per;             //   s / per at local index s % per; NULL: all samples in unet_out
  const uint8_t* idx;
  const uint8_t* owner;
  const float* noise;
  float* out_latent;
  float* out_x0;
  int R1;                     // resampling iterations of the wave (host copy of ed_step_params_t.R1)
};

// Launch geometry of the staged kernel, chosen on the host by staged_config().
struct StagedGeom {
  int bx, by;             // CTA = bx * by threads; tile = by latent rows x 4*bx latent columns
  int bw, bh;             // TMA box in low-res cells: columns (padded to 16 bytes), rows
  unsigned stage_bytes;   // shared-memory stride between the boxes of consecutive samples (multiple of 128)
  int vec_views;          // unet_out is 16-byte aligned: 4 consecutive view elements may be loaded with one instruction
  int origin;             // box origin policy, ED_BOX_*
  int col_align;          // elements per 16 bytes of the UNet output dtype
};
#define ED_BOX_ALIGN 1   // box starts at a 16-byte aligned canvas column
#define ED_BOX_CLAMP 2   // box shifted back inside the canvas plane: no out-of-bounds (zero-filled) part

struct StagedCfg {
  bool ok;
  StagedGeom g;
  size_t smem;
  int grid_x, grid_y, grid_z;
};

// Largest number of source indices that nearest-upsampling (src = min(floor(dst * in/out), in-1), float scale like
// F.interpolate, ed:876) reads for `tile` consecutive outputs starting at a multiple of `tile`.  An estimate only: the
// kernel re-derives the rectangle from the plan's tables and falls back to global loads where a tile needs more.
// `offset` / `align`: the span is measured from the source index + offset rounded down to a multiple of `align`.
static inline int nearest_span_max(int n_out, int n_in, int tile, int offset = 0, int align = 1) {
  const float s = (float)n_in / (float)n_out;
  int best = 1;
  for (int o0 = 0; o0 < n_out; o0 += tile) {
    const int o1 = (o0 + tile < n_out ? o0 + tile : n_out) - 1;
    int a = (int)floorf((float)o0 * s), b = (int)floorf((float)o1 * s);
    a = a < n_in - 1 ? a : n_in - 1;
    b = b < n_in - 1 ? b : n_in - 1;
    a = (a + offset) / align * align - offset;
    best = b - a + 1 > best ? b - a + 1 : best;
  }
  return best;
}

// so = bytes per UNet output element.  Preference: the largest CTA whose boxes fit twice per SM (<= 100 KB) and whose grid
// fills the GPU at least twice; otherwise the smallest CTA that fits (more CTAs for small batches); otherwise one CTA
// per SM (<= 200 KB); otherwise not applicable (the direct kernel runs).
static inline StagedCfg staged_config(const ed_plan_t& P, int R1, int so, int sms, int origin = ED_BOX_ALIGN | ED_BOX_CLAMP,
                                      int cpt = 4, int max_threads = 256) {
  StagedCfg best{};
  if (P.C != 4 || (P.W & 3) || R1 <= 0 || P.B * (P.C / cpt) > 65535 || P.B <= 0) return best;
  const int wv = P.W / 4;
  int bx = 32;
  while (bx > 1 && bx / 2 >= wv) bx /= 2;
  const int align = 16 / so;
  int best_rank = 0;   // 3: fits twice + fills GPU, 2: fits twice, 1: fits once
  for (int threads = max_threads; threads >= 64; threads >>= 1) {
    const int by = threads / bx;
    if (by < 1) break;
    int bw = nearest_span_max(P.W, P.lw, bx * 4, P.g_lp, (origin & ED_BOX_ALIGN) ? align : 1);
    bw = (bw + align - 1) / align * align;
    const int bh = nearest_span_max(P.H, P.lh, by);
    if (bw > 256 || bh > 256) continue;
    const unsigned stage = ((unsigned)(bw * bh * cpt * so) + 127u) & ~127u;
    const size_t smem = (size_t)R1 * 2 * stage + (size_t)bw * bh * cpt * 4;   // boxes + the RRG low-res reference
    if (smem > 200 * 1024) continue;
    const int gx = (P.W + bx * 4 - 1) / (bx * 4), gy = (P.H + by - 1) / by;
    const long long ctas = (long long)gx * gy * P.B * (P.C / cpt);
    const int rank = smem <= 100 * 1024 ? (ctas >= 2LL * sms ? 3 : 2) : 1;
    // rank 3 stops the search (largest CTA wins); among rank <= 2 the later (smaller) CTA wins ties
    if (rank >= best_rank) {
      best.ok = true;
      best.g = StagedGeom{bx, by, bw, bh, stage, 0, origin, align};
      best.smem = smem;
      best.grid_x = gx;
      best.grid_y = gy;
      best.grid_z = P.B * (P.C / cpt);
      best_rank = rank;
      if (rank == 3) break;
    }
  }
  return best;
}

// 4 consecutive elements with one read-only load (p aligned to 4 elements)
ED_DEVICE void ld4_ro(const float* p, float out[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}
ED_DEVICE void ld4_ro(const __half* p, float out[4]) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  __half h[4];
  memcpy(h, &t, 8);
#pragma unroll
  for (int e = 0; e < 4; ++e) out[e] = __half2float(h[e]);
}
ED_DEVICE void ld4_ro(const __nv_bfloat16* p, float out[4]) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat16 h[4];
  memcpy(h, &t, 8);
#pragma unroll
  for (int e = 0; e < 4; ++e) out[e] = __bfloat162float(h[e]);
}

// ed:692-704: x <- a_k x + b_k eps_k, sequential in k like the reference.  CPT channels x 4 pixels per thread; the noise
// is streamed once (evict-first), KB steps x CPT channels = 16 float4 loads in flight before the dependent chain.
template <int CPT>
ED_DEVICE void renoise_stream(float res[CPT][4], const float* nz, long long numel, long long hw, const ed_step_params_t& S,
                              int n_re) {
  constexpr int KB = 16 / CPT;
  int k = 0;
  for (; k + KB <= n_re; k += KB) {
    float4 t[CPT][KB];
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
      for (int j = 0; j < KB; ++j) t[cc][j] = __ldcs(reinterpret_cast<const float4*>(nz + (long long)(k + j) * numel + cc * hw));
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
  for (; k < n_re; ++k) {
    const float a = S.renoise_a[k], bb = S.renoise_b[k];
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(nz + (long long)k * numel + cc * hw));
      res[cc][0] = __fadd_rn(__fmul_rn(a, res[cc][0]), __fmul_rn(bb, t.x));
      res[cc][1] = __fadd_rn(__fmul_rn(a, res[cc][1]), __fmul_rn(bb, t.y));
      res[cc][2] = __fadd_rn(__fmul_rn(a, res[cc][2]), __fmul_rn(bb, t.z));
      res[cc][3] = __fadd_rn(__fmul_rn(a, res[cc][3]), __fmul_rn(bb, t.w));
    }
  }
}

// Correctly rounded a / b for ONE divisor shared by every element (b = sqrt(abar_t), uniform over the launch): the
// reciprocal y = RN(1/b) is taken once, then two Markstein corrections q <- q + (a - b q) y with exact FMA remainders.  The
// second correction starts from a faithful quotient, so its rounding is the IEEE quotient RN(a / b) (Markstein 1990;
// tests/emu/div_check.c compares 10^9 random quotients, and every sqrt(abar_t) of the schedule, with the hardware
// division bit for bit).  5 FMA-pipe instructions instead of __fdiv_rn's ~14 with its slow-path branch; inf / nan /
// overflowing quotients take __fdiv_rn.
struct DivBy {
  float b, nb, y;
};
ED_DEVICE DivBy make_div_by(float b) {
  DivBy d;
  d.b = b;
  d.nb = -b;
  d.y = __frcp_rn(b);
  return d;
}
ED_DEVICE float div_by(const DivBy& d, float a) {
  float q = __fmul_rn(a, d.y);
  float r = __fmaf_rn(d.nb, q, a);
  q = __fmaf_rn(r, d.y, q);
  r = __fmaf_rn(d.nb, q, a);
  q = __fmaf_rn(r, d.y, q);
  if (!(fabsf(q) <= 3.0e38f)) q = __fdiv_rn(a, d.b);
  return q;
}

// RENOISE = false: instantiation without the re-noise stream (launches that pass no noise buffer: wave 2, no-repaint
// steps); its register footprint is smaller (no 16 x float4 prefetch), so it is compiled for more resident CTAs.
// CPT = channels per thread (4 or 2): CPT = 2 spreads the channel pairs over gridDim.z - half the registers and half the
// shared memory per CTA, twice the resident warps; what a latency-bound launch (no noise stream to hide behind) needs.
template <typename OT, bool RENOISE, int CPT>
__global__ void __launch_bounds__(256, CPT == 2 ? 4 : (RENOISE ? ED_STAGED_MINB : ED_STAGED_MINB_PLAIN)) wave_epilogue_staged_kernel(const __grid_constant__ ED_TMAP tm, const EpiArgs A,
                                                                     const StagedGeom G) {
  ED_DYN_SMEM(smem_raw);
  __shared__ uint64_t bar;
  const ed_plan_t& P = A.P;
  const ed_step_params_t& S = *A.prm;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  constexpr int CG = 4 / CPT;                       // channel groups, spread over gridDim.z with the batch
  const int b = (int)blockIdx.z / CG, c_lo = ((int)blockIdx.z - b * CG) * CPT;
  const int R1 = A.R1;
  const int X0 = blockIdx.x * (G.bx * 4), Y0 = blockIdx.y * G.by;
  const int X1 = X0 + G.bx * 4 < P.W ? X0 + G.bx * 4 : P.W;
  const int Y1 = Y0 + G.by < P.H ? Y0 + G.by : P.H;
  // low-res cells nearest-upsampling reads for this tile (ed:636); the tables are non-decreasing
  const int rlo = __ldg(P.up_row + Y0), rhi = __ldg(P.up_row + Y1 - 1);
  const int clo = __ldg(P.up_col + X0), chi = __ldg(P.up_col + X1 - 1);
  // box origin in canvas coordinates (the low-res latent sits at (g_tp, g_lp) inside a canvas plane, ed:405-406)
  int bx0 = P.g_lp + clo, by0 = P.g_tp + rlo;
  if (G.origin & ED_BOX_ALIGN) bx0 &= ~(G.col_align - 1);   // col_align = 16 / sizeof(OT): a power of two
  if (G.origin & ED_BOX_CLAMP) {
    if (bx0 + G.bw > P.dW && G.bw <= P.dW) bx0 = P.dW - G.bw;
    if (by0 + G.bh > P.dH && G.bh <= P.dH) by0 = P.dH - G.bh;
  }
  const int rb = by0 - P.g_tp, cb = bx0 - P.g_lp;   // box origin in low-res cell coordinates (<= rlo, clo)
  const bool staged = (rhi - rb < G.bh) && (chi - cb < G.bw);   // CTA-uniform; false: this tile reads global memory
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (staged && tid == 0) {
    const unsigned box_bytes = (unsigned)(G.bw * G.bh * CPT) * (unsigned)sizeof(OT);
    mbar_expect_tx(&bar, 2u * (unsigned)R1 * box_bytes);
    for (int ks = 0; ks < 2 * R1; ++ks)   // sample (k, s, b) = (2k + s) * B + b; planes of a sample are its C channels
      tma_load_3d(smem_raw + (size_t)ks * G.stage_bytes, &tm, bx0, by0, (ks * P.B + b) * P.C + c_lo, &bar);
  }
  // threads beyond the latent's edge stay alive (the per-cell RRG phase below uses every thread and a CTA barrier): they
  // redo the work of the last valid position and skip the stores
  int x = X0 + (int)threadIdx.x * 4, y = Y0 + (int)threadIdx.y;
  const bool active = x < P.W && y < P.H;
  x = x < P.W ? x : P.W - 4;
  y = y < P.H ? y : P.H - 1;

  const OT* __restrict__ out = static_cast<const OT*>(A.unet_out);
  const int flags = S.flags;
  const bool fp16sem = (flags & ED_FLAG_FP16_SEM) != 0;
  const bool rrg = (flags & ED_FLAG_RRG) != 0;
  const float g = S.guidance, sb = S.sqrt_beta_t, sa = S.sqrt_alpha_t, sap = S.sqrt_alpha_prev, sd = S.sqrt_dir;
  const int n_re = (RENOISE && (flags & ED_FLAG_RENOISE)) ? S.n_renoise : 0;
  const int first_view_sample = 2 * P.B * R1;
  const long long hw = (long long)P.H * P.W;
  const long long plane = (long long)P.dH * P.dW;
  const long long sample_stride = (long long)P.C * plane;
  auto sample = [&](int sidx) -> const OT* { return out + (long long)sidx * sample_stride; };

  // ---- phase A: per-thread global loads, all independent of the boxes in flight ---------------------------------------
  const int pix = y * P.W + x;
  int4 pr[4];   // dir_off, view, view_off, cell (static per-pixel references of the plan)
#pragma unroll
  for (int e = 0; e < 4; ++e) pr[e] = __ldg(reinterpret_cast<const int4*>(P.pix_ref) + pix + e);
  const uchar4 ow = __ldg(reinterpret_cast<const uchar4*>(A.owner + pix));
  const int own[4] = {ow.x, ow.y, ow.z, ow.w};
  const int ur = __ldg(P.up_row + y);
  const DivBy div_sa = make_div_by(sa);
  const long long base0 = (((long long)b * P.C + c_lo) * P.H + y) * P.W + x;   // channel c_lo; + cc * hw per channel
  float xin[CPT][4], uu[CPT][4];
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(A.latent + base0 + cc * hw));
    xin[cc][0] = t.x; xin[cc][1] = t.y; xin[cc][2] = t.z; xin[cc][3] = t.w;
  }
  const bool one_view = pr[0].y >= 0 && pr[1].y == pr[0].y && pr[2].y == pr[0].y && pr[3].y == pr[0].y;
  if (one_view && G.vec_views && (pr[0].z & 3) == 0 && pr[3].z == pr[0].z + 3) {
    ED_EMU_COUNT(4);
    const OT* vs = sample(first_view_sample + pr[0].y * P.B + b) + pr[0].z;
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) ld4_ro(vs + (c_lo + cc) * plane, uu[cc]);
  } else {
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int v_ = pr[e].y >= 0 ? pr[e].y : 0;   // several covering windows (view < 0): patched below
        uu[cc][e] = ld_ro<OT>(sample(first_view_sample + v_ * P.B + b) + (c_lo + cc) * plane + pr[e].z);
      }
    ED_EMU_COUNT(one_view ? 5 : 6);
    if (!one_view) {   // rare: first-writer-wins walk over the covering windows where the value is non-zero (ed:852-861)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (pr[e].y < 0) {
          const int xe = x + e;
          const int r0 = __ldg(P.vrow_first + y), rn = __ldg(P.vrow_cnt + y);
          const int c0 = __ldg(P.vcol_first + xe), cn = __ldg(P.vcol_cnt + xe);
          for (int cc = 0; cc < CPT; ++cc) {
            float u = 0.f;
            bool done = false;
            for (int a = 0; a < rn && !done; ++a)
              for (int q = 0; q < cn && !done; ++q) {
                const int v = (r0 + a) * P.nvc + (c0 + q);
                const int32_t* vt = P.views + v * 8;
                const int yy = P.v_tp + __ldg(vt + 6) + (y - __ldg(vt + 0));
                const int xx = P.v_lp + __ldg(vt + 7) + (xe - __ldg(vt + 2));
                u = ld_ro<OT>(sample(first_view_sample + v * P.B + b) + (c_lo + cc) * plane + (long long)yy * P.dW + xx);
                done = (u != 0.f);
              }
            uu[cc][e] = u;
          }
        }
    }
  }
  // what survives phase A per pixel: offset of its low-res cell inside a box plane (the pix_ref entries are dead from here)
  int so[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) so[e] = (ur - rb) * G.bw + (pr[e].w - ur * P.lw - cb);

  // ---- phase B: picks from the staged boxes (ST = true) or from global memory (ST = false: tile larger than the boxes) --
  if (staged) mbar_wait_bounded(&bar, 0);
  const OT* sm = reinterpret_cast<const OT*>(smem_raw);
  const int stage_el = (int)(G.stage_bytes / sizeof(OT));
  const int plane_el = G.bh * G.bw;
  // RRG low-res reference x0 (ed:909-921) of every cell behind the tile, channel-major in the box layout
  float* rx0s = reinterpret_cast<float*>(smem_raw + (size_t)2 * R1 * G.stage_bytes);

  // the reference's low-res DDIM x0 of one (cell, channel): xl = low-res latent of the last iteration, ul = its uncond
  // score, (lun, lco) = uncond / cond scores behind downsampled_direction (ed:688, 909-921)
  auto low_res_x0 = [&](float xl, float ul, float lun, float lco) -> float {
    float dl = __fsub_rn(lco, lun);
    if (fp16sem) dl = __half2float(__float2half_rn(dl));
    float gl = __fmul_rn(g, dl);
    float el, t1;
    if (fp16sem) {
      gl = __half2float(__float2half_rn(gl));
      el = __half2float(__float2half_rn(__fadd_rn(ul, gl)));         // fp16 + fp16 (ed:918)
      t1 = __half2float(__float2half_rn(__fmul_rn(sb, el)));         // 0-dim fp32 tensor * fp16 tensor -> fp16
    } else {
      el = __fadd_rn(ul, gl);
      t1 = __fmul_rn(sb, el);
    }
    return div_by(div_sa, __fsub_rn(xl, t1));                                 // ed:920-921
  };

  // ---- phase B1 (staged tiles with RRG): the low-res reference depends on the CELL, not on the pixel - one evaluation per
  // (cell, channel) of the tile's rectangle instead of one per pixel (4x fewer at ratio 1/2), shared through smem ---------
  if (staged && rrg) {
    const int nrc = chi - clo + 1, ncell = (rhi - rlo + 1) * nrc;
    const int cells = P.lh * P.lw;
    for (int i = tid; i < ncell; i += (int)(blockDim.x * blockDim.y)) {
      const int rq = i / nrc;
      const int r = rlo + rq, c = clo + (i - rq * nrc);
      const int cell = r * P.lw + c;
      const int pk = __ldg(A.idx + (long long)(R1 - 1) * cells + cell) & 3;
      const int lo = __ldg(P.cell_cand + cell * 4 + pk);
      const int2 d = __ldg(reinterpret_cast<const int2*>(P.cell_down) + cell);
      const int kdn = __ldg(A.owner + d.x);
      const int rr = d.y / P.dW;
      const int rd = rr - P.g_tp, cd = d.y - rr * P.dW - P.g_lp;       // cell nearest-DOWNsampling reads (ed:688)
      // inside the boxes at every exact ratio; general ratios may step outside at a tile border -> global load
      const bool in_box = (unsigned)(rd - rb) < (unsigned)G.bh && (unsigned)(cd - cb) < (unsigned)G.bw;
      ED_EMU_COUNT(in_box ? 2 : 3);
      const int sc = (r - rb) * G.bw + (c - cb), so2 = (rd - rb) * G.bw + (cd - cb);
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) {
        const float xl = __ldg(A.latent + ((long long)b * P.C + c_lo + cc) * hw + lo);                 // ed:910
        const float ul = to_f32<OT>(sm[(size_t)(2 * (R1 - 1)) * stage_el + cc * plane_el + sc]);
        float lun, lco;
        if (in_box) {
          lun = to_f32<OT>(sm[(size_t)(2 * kdn) * stage_el + cc * plane_el + so2]);
          lco = to_f32<OT>(sm[(size_t)(2 * kdn + 1) * stage_el + cc * plane_el + so2]);
        } else {
          lun = ld_ro<OT>(sample((2 * kdn) * P.B + b) + (c_lo + cc) * plane + d.y);
          lco = ld_ro<OT>(sample((2 * kdn + 1) * P.B + b) + (c_lo + cc) * plane + d.y);
        }
        rx0s[cc * plane_el + sc] = low_res_x0(xl, ul, lun, lco);
      }
    }
    __syncthreads();
  }

  float res[CPT][4];
  auto finish = [&](auto staged_tag) {
    constexpr bool ST = decltype(staged_tag)::value;
    // ST: byte offset of (owner iteration, uncond, channel 0, own cell) inside the boxes; (s, cc) add uniform offsets.
    // !ST: canvas-plane offset of the own cell (what pix_ref.x held)
    const char* smb = reinterpret_cast<const char*>(smem_raw);
    const int s_off = (int)G.stage_bytes, c_off = plane_el * (int)sizeof(OT);
    int pb[4], go[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      pb[e] = 2 * own[e] * s_off + so[e] * (int)sizeof(OT);
      go[e] = ST ? 0 : (P.g_tp + ur) * P.dW + P.g_lp + (so[e] - (ur - rb) * G.bw + cb);
    }
    auto own_cell = [&](int e, int k, int s, int cc) -> float {
      float v;
      if constexpr (ST) v = to_f32<OT>(*reinterpret_cast<const OT*>(smb + pb[e] + (2 * (k - own[e]) + s) * s_off + cc * c_off));
      else v = ld_ro<OT>(sample((2 * k + s) * P.B + b) + (c_lo + cc) * plane + go[e]);
      return v;
    };
    // unstaged tiles: per-pixel RRG references (ed:886-940): the pixel the LAST iteration picked for the pixel's low-res
    // cell, and owner + canvas offset of the full-res pixel nearest-DOWNsampling reads for that cell (ed:688)
    int lat_off[4], kd[4], doff[4];
    if constexpr (!ST) {
      if (rrg) {
        const int cells = P.lh * P.lw;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cell = ur * P.lw + (so[e] - (ur - rb) * G.bw + cb);
          const int pk = __ldg(A.idx + (long long)(R1 - 1) * cells + cell) & 3;
          lat_off[e] = __ldg(P.cell_cand + cell * 4 + pk);
          const int2 d = __ldg(reinterpret_cast<const int2*>(P.cell_down) + cell);
          kd[e] = __ldg(A.owner + d.x);
          doff[e] = d.y;
        }
      }
    }
    const float rrg_norm = S.rrg_norm, rrg_w = S.rrg_weight;
    // channel by channel (the x0 of only one channel is live at a time)
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      float x0v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float un = own_cell(e, own[e], 0, cc), co = own_cell(e, own[e], 1, cc);
        float d = __fsub_rn(co, un);                                       // ed:440
        if (fp16sem) d = __half2float(__float2half_rn(d));                 // fp16 tensor under CUDA autocast / ed:655
        float gd = __fmul_rn(g, d);
        if (fp16sem) gd = __half2float(__float2half_rn(gd));               // python float * fp16 tensor -> fp16
        const float eps = __fadd_rn(uu[cc][e], gd);                        // ed:1031
        const float x0 = div_by(div_sa, __fsub_rn(xin[cc][e], __fmul_rn(sb, eps)));   // DDIM "predicted x_0"
        x0v[e] = x0;
        res[cc][e] = __fadd_rn(__fmul_rn(sap, x0), __fmul_rn(sd, eps));    // x_{t-1}, eta = 0
      }
      if (A.out_x0 && active)
        *reinterpret_cast<float4*>(A.out_x0 + base0 + cc * hw) = make_float4(x0v[0], x0v[1], x0v[2], x0v[3]);
      if (rrg) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float rx0;
          if constexpr (ST) {
            rx0 = rx0s[cc * plane_el + so[e]];
          } else {
            const float xl = __ldg(A.latent + ((long long)b * P.C + c_lo + cc) * hw + lat_off[e]);
            const float ul = own_cell(e, R1 - 1, 0, cc);
            const float lun = ld_ro<OT>(sample((2 * kd[e]) * P.B + b) + (c_lo + cc) * plane + doff[e]);
            const float lco = ld_ro<OT>(sample((2 * kd[e] + 1) * P.B + b) + (c_lo + cc) * plane + doff[e]);
            rx0 = low_res_x0(xl, ul, lun, lco);
          }
          // -d/dx0 [ w * mse(ref_up, x0) ] = -( (2/N) * (x0 - ref) * w )   (mse_loss backward, ed:932-935)
          const float grad = __fmul_rn(__fmul_rn(rrg_norm, __fsub_rn(x0v[e], rx0)), rrg_w);
          res[cc][e] = __fadd_rn(res[cc][e], -grad);                        // ed:1078
        }
      }
    }
  };
  if (staged) finish(std::true_type{});
  else finish(std::false_type{});
  ED_EMU_COUNT(staged ? 0 : 1);
  if (!active) return;
  if constexpr (RENOISE)
    if (n_re > 0) renoise_stream<CPT>(res, A.noise + base0, (long long)P.B * P.C * hw, hw, S, n_re);
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc)
    *reinterpret_cast<float4*>(A.out_latent + base0 + cc * hw) = make_float4(res[cc][0], res[cc][1], res[cc][2], res[cc][3]);
}

}  // namespace ed
