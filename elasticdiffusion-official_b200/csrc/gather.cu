// K1 (view gather, TMA), K3/K9 (random-pick gather + background pad), K10a (tile gather, TMA with OOB zero fill).
//
// Reference op chains replaced (elastic_diffusion.py, "ed:N"):
//   K1  crop_with_context x views + torch.cat                         ed:706-757, 834-845
//   K3  random_nearest_downsample / random_downsample value path      ed:523-548, 561-616
//   K9  background_pad's torch.cat of strips around the low-res latent ed:366-391, 405-408
//   K10 F.pad + per-tile slicing / cat of tiled_decode                 ed:287-300
#include "common.cuh"

namespace ed {

// ----------------------------------------------------------------------------------------------------------------
// TMA box copy: one elected thread per CTA drives a ring of STAGES shared-memory boxes; LAG loads and STAGES-LAG
// stores are in flight per CTA, no register traffic at all (global -> smem -> global through the async proxy).
// ----------------------------------------------------------------------------------------------------------------
constexpr int kStages = 6;
constexpr int kLag = 3;

struct BoxJobs {
  int n_jobs;        // total boxes
  int boxes_per_img; // row-boxes per (item, b, c) image
  int bh, bw;        // box rows / cols (elements)
  int B, C;
  const int32_t* tab;  // per-item origin table (device)
  int tab_stride, off_r, off_c;  // origin = (tab[item*stride+off_r] + add_r, tab[item*stride+off_c] + add_c)
  int add_r, add_c;
  int dst_r, dst_c;      // destination offset inside the destination plane
  int dst_first_sample;  // destination plane = (dst_first_sample + item*B + b)*C + c
};

__global__ void __launch_bounds__(32) tma_box_copy_kernel(const __grid_constant__ CUtensorMap src,
                                                          const __grid_constant__ CUtensorMap dst,
                                                          const BoxJobs J) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[kStages];
  if (threadIdx.x != 0) return;
  const uint32_t box_bytes = (uint32_t)J.bh * J.bw * 4u;
  const uint32_t stage_bytes = (box_bytes + 127u) & ~127u;
  tma_prefetch_desc(&src);
  tma_prefetch_desc(&dst);
  for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
  fence_mbar_init();

  const int first = blockIdx.x, step = gridDim.x;
  const int n_mine = first < J.n_jobs ? (J.n_jobs - first + step - 1) / step : 0;
  for (int j = 0; j < n_mine + kLag; ++j) {
    if (j >= kLag) {  // retire box j-kLag: wait for its load, store it
      const int jj = j - kLag, st = jj % kStages;
      mbar_wait(&bars[st], (uint32_t)(jj / kStages) & 1u);
      int job = first + jj * step;
      int t = job % J.boxes_per_img;
      int rest = job / J.boxes_per_img;
      int c = rest % J.C;
      rest /= J.C;
      int b = rest % J.B;
      int item = rest / J.B;
      fence_proxy_async_smem();
      tma_store_3d(&dst, smem + (size_t)st * stage_bytes, J.dst_c, J.dst_r + t * J.bh,
                   (J.dst_first_sample + item * J.B + b) * J.C + c);
      tma_commit();
    }
    if (j < n_mine) {  // issue load j
      const int st = j % kStages;
      if (j >= kStages) tma_wait_read<kStages - kLag>();  // store of box j-kStages has drained its smem stage
      int job = first + j * step;
      int t = job % J.boxes_per_img;
      int rest = job / J.boxes_per_img;
      int c = rest % J.C;
      rest /= J.C;
      int b = rest % J.B;
      int item = rest / J.B;
      int r0 = J.tab[item * J.tab_stride + J.off_r] + J.add_r;
      int c0 = J.tab[item * J.tab_stride + J.off_c] + J.add_c;
      mbar_expect_tx(&bars[st], box_bytes);
      tma_load_3d(smem + (size_t)st * stage_bytes, &src, c0, r0 + t * J.bh, b * J.C + c, &bars[st]);
    }
  }
  tma_wait_all<0>();
}

static int pick_box_rows(int rows, int cols) {
  // largest divisor of `rows` with box <= 16 KiB and <= 256 rows
  int best = 0;
  for (int bh = 1; bh <= rows && bh <= 256; ++bh)
    if (rows % bh == 0 && (long long)bh * cols * 4 <= 16384) best = bh;
  return best;
}

static int launch_tma_box_copy(const float* src, uint64_t sW, uint64_t sH, uint64_t sP, float* dst, uint64_t dW,
                               uint64_t dH, uint64_t dP, BoxJobs J, int rows, int cols, cudaStream_t stream) {
  J.bw = cols;
  J.bh = pick_box_rows(rows, cols);
  if (J.bh < 4) return ED_ERR_UNSUPPORTED;
  J.boxes_per_img = rows / J.bh;
  CUtensorMap ms, md;
  int rc = encode_tmap_3d_f32(&ms, src, sW, sH, sP, cols, J.bh, 1);
  if (rc != ED_OK) return rc;
  rc = encode_tmap_3d_f32(&md, dst, dW, dH, dP, cols, J.bh, 1);
  if (rc != ED_OK) return rc;
  int sms = 0;
  if (int rc2 = current_sm_count(&sms)) return rc2;
  const uint32_t stage_bytes = ((uint32_t)J.bh * cols * 4u + 127u) & ~127u;
  const size_t smem = (size_t)stage_bytes * kStages;
  // per launch: the opt-in is per DEVICE (a process-wide "already set" flag breaks the second GPU of a process); cheap
  ED_CUDA_CHECK(cudaFuncSetAttribute(tma_box_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * kStages + 1024));
  int grid = J.n_jobs < 2 * sms ? J.n_jobs : 2 * sms;  // persistent: 2 CTAs (1 driving thread each) per SM
  tma_box_copy_kernel<<<grid, 32, smem, stream>>>(ms, md, J);
  ED_LAUNCH_CHECK();
  return ED_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// generic (LDG/STG) view gather: output dtype conversion or TMA-incompatible strides
// ----------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gather_views_generic(const ed_plan_t P, const float* __restrict__ latent,
                                                            T* __restrict__ canvas, int first_sample) {
  const long long total = (long long)P.nv * P.B * P.C * P.vh * P.vw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % P.vw);
    long long r = i / P.vw;
    int y = (int)(r % P.vh);
    r /= P.vh;
    int c = (int)(r % P.C);
    r /= P.C;
    int b = (int)(r % P.B);
    int v = (int)(r / P.B);
    const int32_t* vt = P.views + v * 8;
    float val = latent[(((long long)b * P.C + c) * P.H + vt[4] + y) * P.W + vt[5] + x];
    canvas[(((long long)(first_sample + v * P.B + b) * P.C + c) * P.dH + P.v_tp + y) * P.dW + P.v_lp + x] = from_f32<T>(val);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// K3 + K9: every element of the (R+1)*2*B canvases of the global passes in one launch
// ----------------------------------------------------------------------------------------------------------------
struct Strips {
  const float* p[4];  // left (C,lh,l_p), right (C,lh,r_p), top (C,t_p,dW), bottom (C,b_p,dW)
};

__device__ __forceinline__ float pad_value(const Strips& S, int c, int Y, int X, int tp, int lp, int ih, int iw,
                                           int dH, int dW) {
  const int ly = Y - tp, lx = X - lp;
  if (ly < 0) return S.p[2][((long long)c * tp + Y) * dW + X];
  if (ly >= ih) return S.p[3][((long long)c * (dH - tp - ih) + (ly - ih)) * dW + X];
  if (lx < 0) return S.p[0][((long long)c * ih + ly) * lp + X];
  return S.p[1][((long long)c * ih + ly) * (dW - lp - iw) + (lx - iw)];
}

// grid: x over vector-columns of the canvas, y over canvas rows, z over (b, c) planes.  A thread owns VEC canvas
// columns of one row of plane (b, c) for ALL R1 resampling iterations: inside the low-res box it reads the 2x2
// candidate pixels of its VEC cells ONCE (the latent is read once per wave, not once per iteration) and emits the
// picked value per iteration; outside it reads the background strip value once (strips are shared by every sample).
// Either way it then issues 2*R1 vector stores (uncond + cond copies, ed:436).
template <typename T, int VEC>
__global__ void __launch_bounds__(256) pick_gather_kernel(const ed_plan_t P, int R1, const float* __restrict__ latent,
                                                          const uint8_t* __restrict__ idx, const Strips S,
                                                          T* __restrict__ canvas) {
  const int xv = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y * blockDim.y + threadIdx.y;
  if (Y >= P.dH || xv * VEC >= P.dW) return;
  const int n_planes = P.B * P.C;
  const long long plane = (long long)P.dH * P.dW;
  const long long kstride = (long long)n_planes * plane;               // one block of B samples
  const int cells = P.lh * P.lw;
  const int ly = Y - P.g_tp, lx0 = xv * VEC - P.g_lp;
  const bool inner = ly >= 0 && ly < P.lh && lx0 >= 0 && lx0 < P.lw;     // VEC-aligned by construction
  for (int z = blockIdx.z; z < n_planes; z += gridDim.z) {
    T* o = canvas + (long long)z * plane + (long long)Y * P.dW + xv * VEC;   // sample (k=0, s=0, b), channel c
    if (inner) {
      const float* src = latent + (long long)z * P.H * P.W;
      const int r0 = __ldg(P.row_src + 2 * ly) * P.W, r1 = __ldg(P.row_src + 2 * ly + 1) * P.W;
      float cand[VEC][4];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int lx = lx0 + e;
        const int c0 = __ldg(P.col_src + 2 * lx), c1 = __ldg(P.col_src + 2 * lx + 1);
        cand[e][0] = __ldg(src + r0 + c0);
        cand[e][1] = __ldg(src + r0 + c1);
        cand[e][2] = __ldg(src + r1 + c0);
        cand[e][3] = __ldg(src + r1 + c1);
      }
      const int cell0 = ly * P.lw + lx0;
#pragma unroll 2
      for (int k = 0; k < R1; ++k) {
        float v[VEC];
        if constexpr (VEC == 4) {
          const uint32_t pk = __ldg(reinterpret_cast<const uint32_t*>(idx + k * cells + cell0));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int p = (pk >> (8 * e)) & 3;
            v[e] = p == 0 ? cand[e][0] : p == 1 ? cand[e][1] : p == 2 ? cand[e][2] : cand[e][3];
          }
          store4<T>(o + (2 * k) * kstride, v);
          store4<T>(o + (2 * k + 1) * kstride, v);
        } else {
          const int p = __ldg(idx + k * cells + cell0) & 3;
          v[0] = p == 0 ? cand[0][0] : p == 1 ? cand[0][1] : p == 2 ? cand[0][2] : cand[0][3];
          o[(2 * k) * kstride] = from_f32<T>(v[0]);
          o[(2 * k + 1) * kstride] = from_f32<T>(v[0]);
        }
      }
    } else {
      const int c = z % P.C;
      float v[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = pad_value(S, c, Y, xv * VEC + e, P.g_tp, P.g_lp, P.lh, P.lw, P.dH, P.dW);
      for (int k = 0; k < 2 * R1; ++k) {
        if constexpr (VEC == 4) store4<T>(o + k * kstride, v);
        else o[k * kstride] = from_f32<T>(v[0]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pad_views_kernel(const ed_plan_t P, const Strips S, T* __restrict__ canvas,
                                                        int first_sample) {
  const long long plane = (long long)P.dH * P.dW;
  const long long total = (long long)P.nv * P.B * P.C * plane;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int X = (int)(i % P.dW);
    long long r = i / P.dW;
    int Y = (int)(r % P.dH);
    r /= P.dH;
    int c = (int)(r % P.C);
    const int ly = Y - P.v_tp, lx = X - P.v_lp;
    if (ly >= 0 && ly < P.vh && lx >= 0 && lx < P.vw) continue;
    canvas[(long long)first_sample * P.C * plane + i] =
        from_f32<T>(pad_value(S, c, Y, X, P.v_tp, P.v_lp, P.vh, P.vw, P.dH, P.dW));
  }
}

static int grid_for(long long threads) {
  long long g = (threads + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return g < 1 ? 1 : (int)g;
}

static bool check_strips(const float* const strips[4], int tp, int lp, int ih, int iw, int dH, int dW) {
  if (lp > 0 && !strips[0]) return false;
  if (dW - lp - iw > 0 && !strips[1]) return false;
  if (tp > 0 && !strips[2]) return false;
  if (dH - tp - ih > 0 && !strips[3]) return false;
  return true;
}

}  // namespace ed

using namespace ed;

extern "C" {

int ed_gather_views(const ed_plan_t* plan, const float* latent, void* canvas, int canvas_dtype, int first_sample,
                    void* stream_) {
  if (!plan || !latent || !canvas || !plan->views || first_sample < 0) return ED_ERR_INVALID;
  const ed_plan_t& P = *plan;
  if (P.nv <= 0 || P.vh <= 0 || P.vw <= 0 || P.v_tp + P.vh > P.dH || P.v_lp + P.vw > P.dW) return ED_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (canvas_dtype == ED_F32) {
    BoxJobs J{};
    J.n_jobs = 0;
    J.B = P.B;
    J.C = P.C;
    J.tab = P.views;
    J.tab_stride = 8;
    J.off_r = 4;
    J.off_c = 5;
    J.add_r = J.add_c = 0;
    J.dst_r = P.v_tp;
    J.dst_c = P.v_lp;
    J.dst_first_sample = first_sample;
    int bh = pick_box_rows(P.vh, P.vw);
    if (bh >= 4) {
      J.n_jobs = P.nv * P.B * P.C * (P.vh / bh);
      // destination tensor: all canvases up to the last view sample, as (dW, dH, samples*C) planes
      int rc = launch_tma_box_copy(latent, P.W, P.H, (uint64_t)P.B * P.C, static_cast<float*>(canvas), P.dW, P.dH,
                                   (uint64_t)(first_sample + P.nv * P.B) * P.C, J, P.vh, P.vw, stream);
      if (rc != ED_ERR_UNSUPPORTED) return rc;  // alignment not TMA-able: fall through to LDG/STG
    }
  }
  const long long total = (long long)P.nv * P.B * P.C * P.vh * P.vw;
  switch (canvas_dtype) {
    case ED_F32: gather_views_generic<float><<<grid_for(total), 256, 0, stream>>>(P, latent, (float*)canvas, first_sample); break;
    case ED_F16: gather_views_generic<__half><<<grid_for(total), 256, 0, stream>>>(P, latent, (__half*)canvas, first_sample); break;
    case ED_BF16: gather_views_generic<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(P, latent, (__nv_bfloat16*)canvas, first_sample); break;
    default: return ED_ERR_INVALID;
  }
  ED_LAUNCH_CHECK();
  return ED_OK;
}

int ed_random_pick_gather(const ed_plan_t* plan, int R1, const float* latent, const uint8_t* idx,
                          const float* const strips[4], void* canvas, int canvas_dtype, void* stream_) {
  if (!plan || !latent || !idx || !canvas || R1 <= 0 || !plan->row_src || !plan->col_src) return ED_ERR_INVALID;
  const ed_plan_t& P = *plan;
  if (P.g_tp + P.lh > P.dH || P.g_lp + P.lw > P.dW) return ED_ERR_INVALID;
  Strips S{};
  static const float* none[4] = {nullptr, nullptr, nullptr, nullptr};
  if (!strips) strips = none;
  if (!check_strips(strips, P.g_tp, P.g_lp, P.lh, P.lw, P.dH, P.dW)) return ED_ERR_INVALID;
  for (int i = 0; i < 4; ++i) S.p[i] = strips[i];
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool vec = (P.dW % 4 == 0) && (P.lw % 4 == 0) && (P.g_lp % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(canvas) & 15) == 0) && ((reinterpret_cast<uintptr_t>(idx) & 3) == 0);
  const int cols = P.dW / (vec ? 4 : 1), rows = P.dH;
  int bx = 32;
  while (bx > 1 && bx / 2 >= cols) bx /= 2;
  const dim3 block(bx, 256 / bx);
  const long long planes = (long long)P.B * P.C;
  const dim3 grid((cols + block.x - 1) / block.x, (rows + block.y - 1) / block.y, planes > 65535 ? 65535 : (int)planes);
#define ED_PICK(T)                                                                              \
  if (vec) pick_gather_kernel<T, 4><<<grid, block, 0, stream>>>(P, R1, latent, idx, S, (T*)canvas); \
  else pick_gather_kernel<T, 1><<<grid, block, 0, stream>>>(P, R1, latent, idx, S, (T*)canvas);
  switch (canvas_dtype) {
    case ED_F32: ED_PICK(float) break;
    case ED_F16: ED_PICK(__half) break;
    case ED_BF16: ED_PICK(__nv_bfloat16) break;
    default: return ED_ERR_INVALID;
  }
#undef ED_PICK
  ED_LAUNCH_CHECK();
  return ED_OK;
}

int ed_pad_views(const ed_plan_t* plan, const float* const strips[4], void* canvas, int canvas_dtype, int first_sample,
                 void* stream_) {
  if (!plan || !canvas || !strips || first_sample < 0) return ED_ERR_INVALID;
  const ed_plan_t& P = *plan;
  if (!check_strips(strips, P.v_tp, P.v_lp, P.vh, P.vw, P.dH, P.dW)) return ED_ERR_INVALID;
  Strips S{};
  for (int i = 0; i < 4; ++i) S.p[i] = strips[i];
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long total = (long long)P.nv * P.B * P.C * P.dH * P.dW;
  switch (canvas_dtype) {
    case ED_F32: pad_views_kernel<float><<<grid_for(total), 256, 0, stream>>>(P, S, (float*)canvas, first_sample); break;
    case ED_F16: pad_views_kernel<__half><<<grid_for(total), 256, 0, stream>>>(P, S, (__half*)canvas, first_sample); break;
    case ED_BF16: pad_views_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(P, S, (__nv_bfloat16*)canvas, first_sample); break;
    default: return ED_ERR_INVALID;
  }
  ED_LAUNCH_CHECK();
  return ED_OK;
}

}  // extern "C"

// ---- K12: ControlNet condition batch --------------------------------------------------------------------------
namespace ed {
template <typename T>
__global__ void __launch_bounds__(256) gather_cond_kernel(const ed_plan_t P, int R1, const float* __restrict__ cond, int CH,
                                                          int scale, const int32_t* __restrict__ row_map,
                                                          const int32_t* __restrict__ col_map,
                                                          const int32_t* __restrict__ vorigin, T* __restrict__ out) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y * blockDim.y + threadIdx.y;
  const int pW = P.dW * scale, pH = P.dH * scale;
  if (X >= pW || Y >= pH) return;
  const int n_global = 2 * P.B * R1, n_all = n_global + P.nv * P.B;
  const int ch_ = P.lh * scale, cw_ = P.lw * scale;                    // prepared condition size
  for (int z = blockIdx.z; z < n_all * CH; z += gridDim.z) {
    const int n = z / CH, c = z - n * CH;
    float v = 0.f;
    if (n < n_global) {                                               // zero-padded cond[s] (cn:457-461)
      const int s = (n / P.B) & 1;
      const int y = Y - P.g_tp * scale, x = X - P.g_lp * scale;
      if (y >= 0 && y < ch_ && x >= 0 && x < cw_) v = __ldg(cond + (((long long)s * CH + c) * ch_ + y) * cw_ + x);
    } else {                                                          // nearest-upsampled cond[0], view box (cn:933-949)
      const int vi = (n - n_global) / P.B;
      const int y = Y - P.v_tp * scale, x = X - P.v_lp * scale;
      if (y >= 0 && y < P.vh * scale && x >= 0 && x < P.vw * scale) {
        const int sy = __ldg(row_map + __ldg(vorigin + 2 * vi) + y), sx = __ldg(col_map + __ldg(vorigin + 2 * vi + 1) + x);
        v = __ldg(cond + ((long long)c * ch_ + sy) * cw_ + sx);
      }
    }
    out[((long long)z * pH + Y) * pW + X] = from_f32<T>(v);
  }
}
}  // namespace ed

extern "C" int ed_gather_cond(const ed_plan_t* plan, int R1, const float* cond, int CH, int scale, const int32_t* row_map,
                              const int32_t* col_map, const int32_t* vorigin, void* out, int out_dtype, void* stream_) {
  if (!plan || !cond || !row_map || !col_map || !vorigin || !out || R1 <= 0 || CH <= 0 || scale <= 0) return ED_ERR_INVALID;
  const ed_plan_t& P = *plan;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const dim3 block(32, 8);
  const long long planes = (long long)(2 * P.B * R1 + P.nv * P.B) * CH;
  const dim3 grid((P.dW * scale + 31) / 32, (P.dH * scale + 7) / 8, planes > 65535 ? 65535 : (int)planes);
  switch (out_dtype) {
    case ED_F32: gather_cond_kernel<float><<<grid, block, 0, stream>>>(P, R1, cond, CH, scale, row_map, col_map, vorigin, (float*)out); break;
    case ED_F16: gather_cond_kernel<__half><<<grid, block, 0, stream>>>(P, R1, cond, CH, scale, row_map, col_map, vorigin, (__half*)out); break;
    case ED_BF16: gather_cond_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(P, R1, cond, CH, scale, row_map, col_map, vorigin, (__nv_bfloat16*)out); break;
    default: return ED_ERR_INVALID;
  }
  ED_LAUNCH_CHECK();
  return ED_OK;
}

// ---- K10a: tile gather ---------------------------------------------------------------------------------------

namespace ed {
__global__ void __launch_bounds__(256) tile_gather_generic(const float* __restrict__ latent, int B, int C, int H, int W,
                                                           const int32_t* __restrict__ tiles, int ntiles, int T, int pad,
                                                           float* __restrict__ out) {
  const long long total = (long long)ntiles * B * C * T * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % T);
    long long r = i / T;
    int y = (int)(r % T);
    r /= T;
    int c = (int)(r % C);
    r /= C;
    int b = (int)(r % B);
    int j = (int)(r / B);
    int sy = tiles[j * 4 + 0] - pad + y, sx = tiles[j * 4 + 2] - pad + x;
    float v = 0.f;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = latent[(((long long)b * C + c) * H + sy) * W + sx];
    out[i] = v;
  }
}
}  // namespace ed

extern "C" {

int ed_tile_gather(const float* latent, int B, int C, int H, int W, const int32_t* tiles_dev, int ntiles, int core,
                   int pad, float* out, void* stream_) {
  if (!latent || !tiles_dev || !out || B <= 0 || C <= 0 || ntiles <= 0 || core <= 0 || pad < 0) return ED_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int T = core + 2 * pad;
  BoxJobs J{};
  J.B = B;
  J.C = C;
  J.tab = tiles_dev;
  J.tab_stride = 4;
  J.off_r = 0;
  J.off_c = 2;
  J.add_r = -pad;
  J.add_c = -pad;
  J.dst_r = J.dst_c = 0;
  J.dst_first_sample = 0;
  int bh = pick_box_rows(T, T);
  if (bh >= 4) {
    J.n_jobs = ntiles * B * C * (T / bh);
    int rc = launch_tma_box_copy(latent, W, H, (uint64_t)B * C, out, T, T, (uint64_t)ntiles * B * C, J, T, T, stream);
    if (rc != ED_ERR_UNSUPPORTED) return rc;
  }
  const long long total = (long long)ntiles * B * C * T * T;
  tile_gather_generic<<<grid_for(total), 256, 0, stream>>>(latent, B, C, H, W, tiles_dev, ntiles, T, pad, out);
  ED_LAUNCH_CHECK();
  return ED_OK;
}

}  // extern "C"
