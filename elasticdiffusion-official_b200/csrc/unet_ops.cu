// Fused element-wise / normalisation ops INSIDE the UNet forward (opt-in, `unet_ops.py`): the two patterns that the round-2
// profile of the SDXL-shaped UNet showed as the largest non-GEMM costs (profiles/r2_probe_unet.json):
//
//   ed_geglu            out = a * gelu(g) for the two halves (a | g) of a GEGLU projection  (diffusers GEGLU.forward:
//                       `hidden, gate = proj(x).chunk(2, -1); hidden * gelu(gate)`).  PyTorch runs it as two kernels over
//                       STRIDED views (the unvectorised elementwise_kernel<128,4>): 18 % of a batch-20 forward, 24 % of a
//                       batch-3 one.  Here: one pass, 16-byte loads of both halves, one 16-byte store; the intermediate
//                       rounding of gelu(g) to the tensor dtype is kept, so the result is bit-identical to the two torch ops.
//   ed_groupnorm_silu   y = silu?(group_norm(x)) for contiguous NCHW.  PyTorch's statistics kernel launches N*G CTAs - 32 CTAs
//                       on 148 SMs at batch 1 (56 us per call, 46 calls per forward, independent of the batch) - followed by a
//                       parameter kernel, an apply kernel and a separate SiLU kernel.  Here: a split statistics kernel
//                       (N*G*S CTAs, shifted sums for stability) + one apply(+SiLU) kernel that folds the partials.
//
// Both are HBM-bound byte movers like the rest of the library: no tensor cores involved.
#include "common.cuh"

namespace ed {

template <typename T> struct Pack8;   // 8 consecutive elements of T as 16-byte (fp32: 2 x 16-byte) vectors
template <> struct Pack8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = __ldg(reinterpret_cast<const float4*>(p)); b = __ldg(reinterpret_cast<const float4*>(p) + 1); }
  __device__ __forceinline__ void get(float v[8]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
};
template <> struct Pack8<__half> {
  uint4 u;
  __device__ __forceinline__ void load(const __half* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float v[8]) const {
    const __half* h = reinterpret_cast<const __half*>(&u);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __half2float(h[e]);
  }
};
template <> struct Pack8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float v[8]) const {
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&u);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __bfloat162float(h[e]);
  }
};
template <typename T> __device__ __forceinline__ void store8(T* p, const float v[8]) {
  store4<T>(p, v);
  store4<T>(p + 4, v + 4);
}
// value -> tensor dtype -> float: the rounding a torch op applies when it writes its result
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f32<T>(from_f32<T>(v)); }

// x: (M, 2N) contiguous, out: (M, N).  One thread = 8 consecutive columns of one row.
template <typename T>
__global__ void __launch_bounds__(256) geglu_kernel(const T* __restrict__ x, T* __restrict__ out, long long M, int N) {
  const int n8 = N / 8;
  const long long total = M * n8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / n8;
    const int n = (int)(i - m * n8) * 8;
    Pack8<T> pa, pg;
    pa.load(x + m * 2 * N + n);
    pg.load(x + m * 2 * N + N + n);
    float a[8], g[8], o[8];
    pa.get(a);
    pg.get(g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      // torch GeluCUDAKernelImpl (approximate='none'), opmath float: x * 0.5 * (1 + erf(x * M_SQRT1_2)), written as dtype
      const float ge = round_to<T>(__fmul_rn(__fmul_rn(g[e], 0.5f), __fadd_rn(1.0f, erff(__fmul_rn(g[e], 0.70710678118654752440f)))));
      o[e] = __fmul_rn(a[e], ge);                                  // torch mul: float product, written as dtype
    }
    store8<T>(out + m * N + n, o);
  }
}

// ---- GroupNorm statistics: grid (N*G, S); CTA (ng, s) reduces slice s of the group's contiguous chunk of length L ----------
// partial[(ng*S + s)*2 + {0,1}] = sum(x - K), sum((x - K)^2) with K = first element of the group (shifted sums: no
// cancellation when |mean| >> std).
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, float* __restrict__ partial, long long L, int S) {
  const long long ng = blockIdx.x;
  const int s = blockIdx.y;
  const T* base = x + ng * L;
  const float K = to_f32<T>(__ldg(base));
  const long long v8 = L / 8;                                      // L % 8 == 0 (checked by the launcher)
  const long long per = (v8 + S - 1) / S;
  const long long lo = s * per, hi = (lo + per < v8) ? lo + per : v8;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    Pack8<T> p;
    p.load(base + i * 8);
    float v[8];
    p.get(v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = v[e] - K;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  __shared__ float w1[8], w2[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    w1[warp] = s1;
    w2[warp] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += w1[w];
      b += w2[w];
    }
    partial[(ng * S + s) * 2 + 0] = a;
    partial[(ng * S + s) * 2 + 1] = b;
  }
}

// ---- apply (+ SiLU): grid (N*C, chunks of the HW plane); folds the S partials of the plane's group ---------------------------
template <typename T, bool SILU>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta,
                                                       const float* __restrict__ partial, T* __restrict__ out, int C, int HW, int G,
                                                       int S, float eps) {
  const long long nc = blockIdx.x;
  const int c = (int)(nc % C);
  const long long n = nc / C;
  const int cpg = C / G;
  const long long ng = n * G + c / cpg;
  const long long L = (long long)cpg * HW;
  float s1 = 0.f, s2 = 0.f;
  for (int s = 0; s < S; ++s) {
    s1 += __ldg(partial + (ng * S + s) * 2);
    s2 += __ldg(partial + (ng * S + s) * 2 + 1);
  }
  const float K = to_f32<T>(__ldg(x + ng * L));
  const float md = s1 / (float)L;                                   // mean - K
  const float var = fmaxf(s2 / (float)L - md * md, 0.f);
  const float mean = md + K;
  const float rstd = rsqrtf(var + eps);
  // torch ComputeFusedParams: a = rstd * gamma, b = -a * mean + beta; apply: a * x + b in float, written as dtype
  const float ga = gamma ? to_f32<T>(__ldg(gamma + c)) : 1.f, be = beta ? to_f32<T>(__ldg(beta + c)) : 0.f;
  const float a = rstd * ga;
  const float b = fmaf(-a, mean, be);
  const T* src = x + nc * HW;
  T* dst = out + nc * HW;
  const int v8 = HW / 8;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < v8; i += gridDim.y * blockDim.x) {
    Pack8<T> p;
    p.load(src + (long long)i * 8);
    float v[8];
    p.get(v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = fmaf(a, v[e], b);
      if constexpr (SILU) {
        y = round_to<T>(y);                                        // group_norm's output as the SiLU kernel reads it
        y = y / (1.0f + expf(-y));                                 // torch silu: x / (1 + exp(-x)) in float
      }
      v[e] = y;
    }
    store8<T>(dst + (long long)i * 8, v);
  }
}

// ---- conv epilogue: y += bias[c] (+ per_nc[n, c]) (+ residual), each step rounded to the tensor dtype like the separate torch
// ops (cuDNN conv output -> add_(bias) -> + time-embedding broadcast -> + residual): grid (N*C, chunks of the HW plane) --------
template <typename T>
__global__ void __launch_bounds__(256) bias_add_kernel(T* __restrict__ y, const T* __restrict__ bias, const T* __restrict__ per_nc,
                                                       const T* __restrict__ residual, int C, int HW) {
  const long long nc = blockIdx.x;
  const int c = (int)(nc % C);
  const float b = bias ? to_f32<T>(__ldg(bias + c)) : 0.f;
  const float e = per_nc ? to_f32<T>(__ldg(per_nc + nc)) : 0.f;
  T* dst = y + nc * HW;
  const T* res = residual ? residual + nc * HW : nullptr;
  const int v8 = HW / 8;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < v8; i += gridDim.y * blockDim.x) {
    Pack8<T> p, r;
    p.load(dst + (long long)i * 8);
    float v[8], rv[8];
    p.get(v);
    if (res) {
      r.load(res + (long long)i * 8);
      r.get(rv);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = v[k];
      if (bias) t = round_to<T>(__fadd_rn(t, b));
      if (per_nc) t = round_to<T>(__fadd_rn(t, e));
      if (res) t = __fadd_rn(t, rv[k]);
      v[k] = t;
    }
    store8<T>(dst + (long long)i * 8, v);
  }
}

// ---- the same conv epilogue reading the conv output in NHWC (what cuDNN's kernels produce natively) and writing NCHW: fuses
// the layout transform cuDNN otherwise runs as a separate nhwcToNchw pass.  src element (n, c, p) at (n*HW + p)*C + c, dst
// (n*C + c)*HW + p.  One CTA = a 64 (pixels) x 64 (channels) tile through shared memory; C % 64 == 0, HW % 64 == 0. ---------
template <typename T>
__global__ void __launch_bounds__(256) bias_add_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, const T* __restrict__ bias,
                                                            const T* __restrict__ per_nc, const T* __restrict__ residual, int C, int HW) {
  __shared__ float tile[64][65];
  const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  // load: thread -> (pixel, 8-channel vector); 64 x 8 vectors per tile, 2 per thread
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int v = tid + 256 * k, p = v >> 3, cv = v & 7;
    Pack8<T> pk;
    pk.load(src + ((long long)n * HW + p0 + p) * C + c0 + cv * 8);
    float f[8];
    pk.get(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[p][cv * 8 + j] = f[j];
  }
  __syncthreads();
  // store: thread -> (channel, 8-pixel vector); a warp covers 32 consecutive channels of one pixel vector (conflict-free reads)
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int v = tid + 256 * k, c = v & 63, pv = v >> 6;
    const long long nc = (long long)n * C + c0 + c;
    const float b = bias ? to_f32<T>(__ldg(bias + c0 + c)) : 0.f;
    const float e = per_nc ? to_f32<T>(__ldg(per_nc + nc)) : 0.f;
    float o[8], rv[8];
    if (residual) {
      Pack8<T> r;
      r.load(residual + nc * HW + p0 + pv * 8);
      r.get(rv);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = tile[pv * 8 + j][c];
      if (bias) t = round_to<T>(__fadd_rn(t, b));
      if (per_nc) t = round_to<T>(__fadd_rn(t, e));
      if (residual) t = __fadd_rn(t, rv[j]);
      o[j] = t;
    }
    store8<T>(dst + nc * HW + p0 + pv * 8, o);
  }
}

// ---- LayerNorm over the last dimension: one warp per row, the row lives in registers (two-pass mean / variance) -------------
// D % 8 == 0 and D <= 32 * 8 * VPL.  out = (x - mean) * rstd * gamma + beta in float, written as dtype (torch's formula).
template <typename T, int VPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta,
                                                        T* __restrict__ out, long long M, int D, float eps) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nv = D / 8;                                           // 8-element vectors per row
  const T* src = x + row * D;
  float v[VPL][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    if (vi < nv) {
      Pack8<T> p;
      p.load(src + vi * 8);
      p.get(v[j]);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[j][k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j)
    if (lane + 32 * j < nv) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = v[j][k] - mean;
        q = fmaf(d, d, q);
      }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)D + eps);
  T* dst = out + row * D;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    if (vi < nv) {
      float g[8], b[8];
      if (gamma) {
        Pack8<T> pg;
        pg.load(gamma + vi * 8);
        pg.get(g);
      }
      if (beta) {
        Pack8<T> pb;
        pb.load(beta + vi * 8);
        pb.get(b);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float y = (v[j][k] - mean) * rstd;
        if (gamma) y *= g[k];
        if (beta) y += b[k];
        v[j][k] = y;
      }
      store8<T>(dst + vi * 8, v[j]);
    }
  }
}

}  // namespace ed

using namespace ed;

extern "C" int ed_bias_add(void* y, const void* bias, const void* per_nc, const void* residual, int N, int C, int HW, int dtype,
                           void* stream_) {
  if (!y || N <= 0 || C <= 0 || HW <= 0) return ED_ERR_INVALID;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (HW % 8 || !al(y) || (residual && !al(residual))) return ED_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int chunks = (HW / 8 + 255) / 256;
  if (chunks > 64) chunks = 64;
  const dim3 g((unsigned)((long long)N * C), (unsigned)chunks);
  switch (dtype) {
    case ED_F32: bias_add_kernel<float><<<g, 256, 0, stream>>>((float*)y, (const float*)bias, (const float*)per_nc, (const float*)residual, C, HW); break;
    case ED_F16: bias_add_kernel<__half><<<g, 256, 0, stream>>>((__half*)y, (const __half*)bias, (const __half*)per_nc, (const __half*)residual, C, HW); break;
    case ED_BF16: bias_add_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>((__nv_bfloat16*)y, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)per_nc, (const __nv_bfloat16*)residual, C, HW); break;
    default: return ED_ERR_INVALID;
  }
  ED_LAUNCH_CHECK();
  return ED_OK;
}

extern "C" int ed_bias_add_nhwc(const void* y_nhwc, void* out_nchw, const void* bias, const void* per_nc, const void* residual, int N,
                                int C, int HW, int dtype, void* stream_) {
  if (!y_nhwc || !out_nchw || N <= 0 || C <= 0 || HW <= 0) return ED_ERR_INVALID;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (C % 64 || HW % 64 || N > 65535 || !al(y_nhwc) || !al(out_nchw) || (residual && !al(residual))) return ED_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const dim3 g((unsigned)(HW / 64), (unsigned)(C / 64), (unsigned)N);
  switch (dtype) {
    case ED_F32: bias_add_nhwc_kernel<float><<<g, 256, 0, stream>>>((const float*)y_nhwc, (float*)out_nchw, (const float*)bias, (const float*)per_nc, (const float*)residual, C, HW); break;
    case ED_F16: bias_add_nhwc_kernel<__half><<<g, 256, 0, stream>>>((const __half*)y_nhwc, (__half*)out_nchw, (const __half*)bias, (const __half*)per_nc, (const __half*)residual, C, HW); break;
    case ED_BF16: bias_add_nhwc_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)y_nhwc, (__nv_bfloat16*)out_nchw, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)per_nc, (const __nv_bfloat16*)residual, C, HW); break;
    default: return ED_ERR_INVALID;
  }
  ED_LAUNCH_CHECK();
  return ED_OK;
}

template <typename T>
static int launch_layernorm(const void* x, const void* gamma, const void* beta, void* out, long long M, int D, float eps,
                            cudaStream_t stream) {
  const int nv = D / 8;
  const unsigned grid = (unsigned)((M + 7) / 8);
  const T *xx = (const T*)x, *g = (const T*)gamma, *b = (const T*)beta;
  if (nv <= 32) layernorm_kernel<T, 1><<<grid, 256, 0, stream>>>(xx, g, b, (T*)out, M, D, eps);
  else if (nv <= 64) layernorm_kernel<T, 2><<<grid, 256, 0, stream>>>(xx, g, b, (T*)out, M, D, eps);
  else if (nv <= 96) layernorm_kernel<T, 3><<<grid, 256, 0, stream>>>(xx, g, b, (T*)out, M, D, eps);
  else if (nv <= 160) layernorm_kernel<T, 5><<<grid, 256, 0, stream>>>(xx, g, b, (T*)out, M, D, eps);
  else if (nv <= 256) layernorm_kernel<T, 8><<<grid, 256, 0, stream>>>(xx, g, b, (T*)out, M, D, eps);
  else return ED_ERR_UNSUPPORTED;
  ED_LAUNCH_CHECK();
  return ED_OK;
}

extern "C" int ed_layernorm(const void* x, const void* gamma, const void* beta, void* out, int64_t M, int D, float eps, int dtype,
                            void* stream_) {
  if (!x || !out || M <= 0 || D <= 0) return ED_ERR_INVALID;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (D % 8 || !al(x) || !al(out) || (gamma && !al(gamma)) || (beta && !al(beta))) return ED_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  switch (dtype) {
    case ED_F32: return launch_layernorm<float>(x, gamma, beta, out, M, D, eps, stream);
    case ED_F16: return launch_layernorm<__half>(x, gamma, beta, out, M, D, eps, stream);
    case ED_BF16: return launch_layernorm<__nv_bfloat16>(x, gamma, beta, out, M, D, eps, stream);
    default: return ED_ERR_INVALID;
  }
}

extern "C" int ed_geglu(const void* x, void* out, int64_t M, int N, int dtype, void* stream_) {
  if (!x || !out || M <= 0 || N <= 0) return ED_ERR_INVALID;
  if (N % 8 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return ED_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int sms = 0;
  if (int rc = current_sm_count(&sms)) return rc;
  const long long total = M * (N / 8);
  long long grid = (total + 255) / 256;
  if (grid > (long long)sms * 16) grid = (long long)sms * 16;
  switch (dtype) {
    case ED_F32: geglu_kernel<float><<<(int)grid, 256, 0, stream>>>((const float*)x, (float*)out, M, N); break;
    case ED_F16: geglu_kernel<__half><<<(int)grid, 256, 0, stream>>>((const __half*)x, (__half*)out, M, N); break;
    case ED_BF16: geglu_kernel<__nv_bfloat16><<<(int)grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, M, N); break;
    default: return ED_ERR_INVALID;
  }
  ED_LAUNCH_CHECK();
  return ED_OK;
}

extern "C" int ed_groupnorm_split(int N, int C, int HW, int G) {
  // slices per group so that the statistics kernel launches >= ~4 CTAs per SM without slices shorter than 2048 elements
  if (N <= 0 || C <= 0 || HW <= 0 || G <= 0 || C % G) return 0;
  const long long L = (long long)(C / G) * HW;
  long long want = (4LL * 148 + (long long)N * G - 1) / ((long long)N * G);
  const long long cap = L / 2048 > 0 ? L / 2048 : 1;
  if (want > cap) want = cap;
  if (want > 64) want = 64;
  return want < 1 ? 1 : (int)want;
}

extern "C" int ed_groupnorm_silu(const void* x, const void* gamma, const void* beta, void* out, float* workspace, int N, int C,
                                 int HW, int G, float eps, int silu, int dtype, void* stream_) {
  if (!x || !out || !workspace || N <= 0 || C <= 0 || HW <= 0 || G <= 0 || C % G) return ED_ERR_INVALID;
  if (HW % 8 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return ED_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int S = ed_groupnorm_split(N, C, HW, G);
  const long long L = (long long)(C / G) * HW;
  const dim3 gs((unsigned)(N * G), (unsigned)S);
  int chunks = (HW / 8 + 255) / 256;
  if (chunks > 64) chunks = 64;
  const dim3 ga((unsigned)((long long)N * C), (unsigned)chunks);
#define ED_GN(T)                                                                                                             \
  gn_stats_kernel<T><<<gs, 256, 0, stream>>>((const T*)x, workspace, L, S);                                                    \
  if (silu) gn_apply_kernel<T, true><<<ga, 256, 0, stream>>>((const T*)x, (const T*)gamma, (const T*)beta, workspace, (T*)out, C, HW, G, S, eps); \
  else gn_apply_kernel<T, false><<<ga, 256, 0, stream>>>((const T*)x, (const T*)gamma, (const T*)beta, workspace, (T*)out, C, HW, G, S, eps);
  switch (dtype) {
    case ED_F32: ED_GN(float) break;
    case ED_F16: ED_GN(__half) break;
    case ED_BF16: ED_GN(__nv_bfloat16) break;
    default: return ED_ERR_INVALID;
  }
#undef ED_GN
  ED_LAUNCH_CHECK();
  return ED_OK;
}
