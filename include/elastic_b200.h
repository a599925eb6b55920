/*
 * elastic_b200.h - C ABI of the B200-native ElasticDiffusion hot path (libelastic_b200.so, sm_100a).
 *
 * The reference (MoayedHajiAli/ElasticDiffusion-official) is 100 % Python/PyTorch: it has NO plugin / FFI / C-ABI
 * (SURVEY.md section 8b).  The drop-in boundary is the Python class `ElasticDiffusion.generate_image`
 * (reference elastic_diffusion.py:952-1130, "ed:N" below); this header defines the native op set that the new
 * `generate_image` calls underneath it.  Every entry point below names the chain of reference ops it replaces.
 *
 * Conventions
 *   - plain C linkage, raw DEVICE pointers + sizes, no torch types; `stream` is a cudaStream_t passed as void*.
 *   - no allocation, no synchronisation, no host<->device copies inside (CUDA-graph capturable), except the
 *     explicit ed_upload_* helper which is a cudaMemcpyAsync.
 *   - returns ED_OK (0) or a negative ed_status; ed_strerror() gives the text.  CUDA launch errors are returned as
 *     ED_ERR_CUDA with the cudaError_t retrievable through ed_last_cuda_error().
 *   - all latents are contiguous NCHW; "canvas" = one UNet sample of native size (C, dH, dW).
 *
 * Wave sample layout (one UNet batch per wave):
 *     sample(k, s, b) = (k*2 + s)*B + b      k = resampling iteration 0..R, s = 0 uncond / 1 cond   (ed:436-439)
 *     sample(view v, b) = 2*B*(R+1) + v*B + b                                                         (ed:845)
 */
#ifndef ELASTIC_B200_H_
#define ELASTIC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ED_ABI_VERSION 4
#define ED_MAX_RENOISE 1000

typedef enum {
  ED_OK = 0,
  ED_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unsupported dtype ...) */
  ED_ERR_UNSUPPORTED = -2, /* shape outside what the kernel handles (message says which) */
  ED_ERR_CUDA = -3,        /* CUDA runtime / driver error, see ed_last_cuda_error() */
  ED_ERR_NO_DEVICE = -4    /* no sm_100 device / driver present */
} ed_status;

typedef enum { ED_F32 = 0, ED_F16 = 1, ED_BF16 = 2 } ed_dtype;

/* Static geometry of one generate_image call ("plan"): built once on the host from the reference's integer logic
 * (get_views ed:198-229, crop_with_context ed:706-757, random_nearest_downsample tables ed:568-611,
 * restore_mask_shape ed:446-465, F.interpolate(nearest) index maps ed:636,688,922).  All pointers are DEVICE
 * pointers to int32 tables that stay alive for the whole call. */
typedef struct {
  int32_t B, C, H, W;       /* latent (B,C,H,W) fp32 */
  int32_t dH, dW;           /* UNet native canvas (64 or 128, ed:398-400) */
  int32_t lh, lw;           /* low-res (resampled) latent size */
  int32_t g_tp, g_lp;       /* top/left offset of the low-res latent inside the canvas (ed:405-406) */
  int32_t nv, nvr, nvc;     /* number of views, view-grid rows, view-grid cols (view i = r*nvc + c) */
  int32_t vh, vw;           /* view crop size (window + context), same for all views */
  int32_t v_tp, v_lp;       /* top/left offset of a view crop inside the canvas (0 unless view < native) */
  int32_t flags;            /* ED_PLAN_* : properties of the geometry the CALLER has verified on the host (0 = none) */
  const int32_t* row_src;   /* [2*lh] latent row feeding each row of the 2x-resized grid (ed:585-589,612) */
  const int32_t* col_src;   /* [2*lw] */
  const int32_t* mrow_lo;   /* [H] first resized row OR-ed into latent row y of the restored mask (ed:446-465) */
  const int32_t* mrow_n;    /* [H] how many (0,1,2) */
  const int32_t* mcol_lo;   /* [W] */
  const int32_t* mcol_n;    /* [W] */
  const int32_t* up_row;    /* [H]  low-res row read by nearest-upsampling to row y (ed:636,922) */
  const int32_t* up_col;    /* [W] */
  const int32_t* down_row;  /* [lh] full-res row read by nearest-downsampling (ed:688) */
  const int32_t* down_col;  /* [lw] */
  const int32_t* views;     /* [nv*8] h0,h1,w0,w1 (window), r0,c0 (crop origin), n_t,n_l (window offset in crop) */
  const int32_t* vrow_first;/* [H] first view-grid row whose window covers latent row y */
  const int32_t* vrow_cnt;  /* [H] number of consecutive covering view-grid rows */
  const int32_t* vcol_first;/* [W] */
  const int32_t* vcol_cnt;  /* [W] */
  /* static per-pixel / per-cell references derived from the tables above (host-built once per call) so that the
   * epilogue's per-pixel prologue is three loads instead of ~40 dependent table walks: */
  const int32_t* pix_ref;   /* [H*W*4] dir_off (offset of the low-res cell nearest-up reads, inside a canvas plane),
                               view (single covering view or -1), view_off (offset of the pixel in that view's canvas
                               plane), cell (low-res cell index ur*lw+uc) */
  const int32_t* cell_cand; /* [lh*lw*4] latent-plane offsets of the 4 candidate pixels of each 2x2 cell (ed:612-613) */
  const int32_t* cell_down; /* [lh*lw*2] pixel index Y*W+X that nearest-DOWNsampling reads for the cell (ed:688), and
                               the dir_off of that pixel */
  /* per-row / per-column view offsets (the windows form a grid): canvas row of latent row y inside the view of the
   * first covering grid row = v_tp + n_t + (y - h0), and the same per column; offset of pixel (y, x) in that view's
   * canvas plane = vrow_off[y]*dW + vcol_off[x] */
  const int32_t* vrow_off;  /* [H] */
  const int32_t* vcol_off;  /* [W] */
} ed_plan_t;

/* ED_PLAN_HALF_FAST - the exact 1/2-ratio geometry with tiling views (every tiled BASELINE config: SD2.1 512x1024, SDXL
 * 1024x2048, SDXL 2048x2048).  The caller asserts ALL of:
 *   C == 4, H even, W % 8 == 0, lh*2 == H, lw*2 == W, dW % 8 == 0, g_lp % 4 == 0;
 *   up_row[y] == y/2, up_col[x] == x/2; row_src[i] == i, col_src[i] == i (the 2x-resized grid is the latent itself);
 *   down_row[r] == 2r, down_col[c] == 2c; mrow_lo[y] == y, mrow_n[y] == 1, mcol_lo[x] == x, mcol_n[x] == 1;
 *   every pixel is covered by exactly one view window (vrow_cnt == vcol_cnt == 1 everywhere);
 *   rows 2r and 2r+1 lie in the same view on consecutive canvas rows; every aligned group of 8 columns lies in one view,
 *   contiguous, starting at a canvas column that is a multiple of 8 (vcol_off[8j] % 8 == 0).
 * With it, ed_wave_epilogue[_peer] derives every index arithmetically from (y, x) instead of loading the per-pixel
 * reference tables (the "half" kernels, csrc/epilogue_half.cuh).  Results are bit-identical with and without the flag. */
#define ED_PLAN_HALF_FAST 1

/* Per-wave scalars, read by the epilogue kernel from DEVICE memory so that a captured CUDA graph can be replayed
 * with new values (upload with ed_upload_step_params or any memcpy). */
typedef struct {
  float guidance;           /* g of eps = uncond + g*direction (ed:1031,1053) */
  float sqrt_beta_t;        /* (1 - abar_t)^0.5   DDIM, diffusers 0.21.4 scheduling_ddim.step */
  float sqrt_alpha_t;       /* abar_t^0.5 */
  float sqrt_alpha_prev;    /* abar_prev^0.5 */
  float sqrt_dir;           /* (1 - abar_prev)^0.5  (eta = 0) */
  float rrg_weight;         /* fp32(rrg_scheduler(i))  (ed:1062-1071) */
  float rrg_norm;           /* fp32(2 / (C*H*W))  mse_loss backward, reduction=mean (ed:932) */
  int32_t flags;            /* ED_FLAG_* */
  int32_t n_renoise;        /* forward steps of undo_step (ed:693), <= ED_MAX_RENOISE */
  int32_t R1;               /* resampling iterations in this wave = resampling_steps+1 (ed:661) */
  int32_t reserved[2];
  float renoise_a[ED_MAX_RENOISE]; /* (1-beta_{t+i})^0.5 (ed:702) */
  float renoise_b[ED_MAX_RENOISE]; /* beta_{t+i}^0.5 */
} ed_step_params_t;

#define ED_FLAG_RENOISE 1   /* out_latent = undo_step(x_prev) (ed:1039-1040) */
#define ED_FLAG_RRG 2       /* out_latent = x_prev + reduced_resolution_guidance(...) (ed:1062-1078) */
#define ED_FLAG_FP16_SEM 4  /* emulate the fp16 roundings of the reference's CUDA-autocast path (ed:655,1031) */

/* ---- library / device ------------------------------------------------------------------------------------- */
int ed_abi_version(void);
const char* ed_strerror(int status);
int ed_last_cuda_error(void);
/* ED_OK when device `dev` exists and is compute capability 10.x; writes the SM count. */
int ed_device_check(int dev, int* sm_count);
/* cudaMemcpyAsync(host -> device) of one ed_step_params_t (host buffer must stay valid until the copy ran). */
int ed_upload_step_params(void* d_params, const ed_step_params_t* h_params, void* stream);

/* ---- K1: view gather (TMA)  -- replaces crop_with_context x views + torch.cat, ed:834-845, 706-757 ------------
 * Copies, for every view v, batch b, channel c, the contiguous box latent[b,c, r0:r0+vh, c0:c0+vw] into
 * canvas sample(view v, b) at offset (v_tp, v_lp).  fp32 -> fp32 goes through TMA tensor-map box loads/stores
 * (UTMALDG/UTMASTG); other output dtypes or TMA-incompatible strides use a vectorised LDG/STG kernel.
 * `first_sample` = index of sample(view 0, b 0) inside `canvas`. */
int ed_gather_views(const ed_plan_t* plan, const float* latent, void* canvas, int canvas_dtype,
                    int first_sample, void* stream);

/* ---- K3/K9: random-pick gather + background pad -- replaces random_nearest_downsample / random_downsample
 * (ed:523-630: nearest 2x upsample, row/col select, 4x F.unfold, advanced index) and the torch.cat of
 * background_pad (ed:366-391) for all R+1 resampling iterations of a wave at once.
 *   idx      [R1][lh*lw] uint8 : which of the 4 pixels of each 2x2 cell was drawn (ed:534-544)
 *   strips   4 device pointers (left,right,top,bottom) to fp32 (1,C,h,w) background strips or NULL (ed:379-387);
 *            left/right are (C, lh, l_p / r_p), top/bottom are (C, t_p / b_p, dW)
 * Writes sample(k,0,b) and sample(k,1,b) (the reference feeds cat([x]*2), ed:436). */
int ed_random_pick_gather(const ed_plan_t* plan, int R1, const float* latent, const uint8_t* idx,
                          const float* const strips[4], void* canvas, int canvas_dtype, void* stream);

/* background strips for view samples smaller than the native size (window collapse, ed:820-825 + ed:405-408) */
int ed_pad_views(const ed_plan_t* plan, const float* const strips[4], void* canvas, int canvas_dtype,
                 int first_sample, void* stream);

/* ---- owner map: which resampling iteration's masked fill is the last to touch each full-res pixel -------------
 * Replaces the sequence of `torch.where(mask_k, up_k, target)` over k = 0..R and the NaN back-fill (ed:637, 643-644)
 * plus the mask restoration of restore_mask_shape (ed:446-465, 622-628): owner[y*W+x] = max{k : mask_k(y,x)} or R if no
 * iteration sampled the pixel.  owner: (H*W) uint8, recomputed per wave from idx (R1 <= 255). */
int ed_owner_map(const ed_plan_t* plan, int R1, const uint8_t* idx, uint8_t* owner, void* stream);

/* ---- K2+K5+K6(+K7)(+K8): fused wave epilogue ---------------------------------------------------------------
 * One pass over the full-resolution latent that replaces
 *   - first-writer-wins view scatter                        compute_local_uncond_signal ed:852-861
 *   - direction = cond - uncond, nearest-up, masked fill over R+1 iterations, NaN back-fill
 *                                                           ed:439-440, 634-647, 661-681
 *   - eps = uncond + g*direction, DDIM x0 / x_prev          ed:1031-1035, 1053-1056 (+ diffusers step)
 *   - flags & ED_FLAG_RENOISE: undo_step                    ed:692-704 (noise = n_renoise torch-drawn tensors)
 *   - flags & ED_FLAG_RRG: reduced-resolution guidance      ed:886-940 and global_latent = nxt + cascade ed:1078
 * R1        resampling iterations of the wave (= d_params->R1; the host needs it to size the launch)
 * unet_out  (n_samples, C, dH, dW) of dtype out_dtype in the wave sample layout
 * owner     (H*W) uint8 from ed_owner_map for the same idx
 * noise     [n_renoise][B*C*H*W] fp32 or NULL
 * out_latent / out_x0 fp32 (B,C,H,W); out_x0 may be NULL.
 * Two kernels implement it (bit-identical results): the tile-staged one (a CTA pulls the low-res rectangle of all 2*R1
 * global-pass outputs behind its latent tile into shared memory with TMA box loads, then picks per pixel from shared
 * memory) whenever C == 4, W % 4 == 0, 16-byte aligned buffers and the boxes fit in shared memory; otherwise the direct
 * kernel (scattered read-only loads). */
int ed_wave_epilogue(const ed_plan_t* plan, const ed_step_params_t* d_params, int R1, const float* latent,
                     const void* unet_out, int out_dtype, const uint8_t* idx, const uint8_t* owner,
                     const float* noise, float* out_latent, float* out_x0, void* stream);

/* Kernel selection of ed_wave_epilogue, process-wide (tests / A-B measurements): AUTO = the half kernels when the plan
 * carries ED_PLAN_HALF_FAST (and the launch is in their domain), else staged / direct as described above; DIRECT and
 * STAGED never take the half kernels; STAGED and HALF return ED_ERR_UNSUPPORTED instead of falling back. */
typedef enum { ED_EPILOGUE_AUTO = 0, ED_EPILOGUE_DIRECT = 1, ED_EPILOGUE_STAGED = 2, ED_EPILOGUE_HALF = 3 } ed_epilogue_mode;
int ed_set_epilogue_mode(int mode);
/* Process-wide number of ed_wave_epilogue / ed_wave_epilogue_peer launches that took the direct and the staged kernel
 * (diagnostics: which kernel AUTO chose; either pointer may be NULL); ed_epilogue_launch_counts3 adds the half kernels. */
int ed_epilogue_launch_counts(int64_t* direct, int64_t* staged);
int ed_epilogue_launch_counts3(int64_t* direct, int64_t* staged, int64_t* half);

/* ---- C1: the same epilogue fused with the multi-GPU exchange ---------------------------------
 * With wave samples sharded over `world` ranks (rank r holds samples [r*per, (r+1)*per) of the wave layout in its own
 * buffer), no all-gather is run: `d_peer_out` is a DEVICE array of `world` pointers to the ranks' buffers, all mapped
 * into this process (symmetric memory / CUDA IPC), and the kernel reads every sample it needs straight from its owner
 * over NVLink (P2P ld.global).  The caller orders "all ranks finished their UNet" before the launch (device-side
 * symmetric-memory barrier).  Every rank runs the (replicated, microsecond) epilogue itself. */
int ed_wave_epilogue_peer(const ed_plan_t* plan, const ed_step_params_t* d_params, int R1, const float* latent,
                          const void* const* d_peer_out, int world, int per, int out_dtype, const uint8_t* idx,
                          const uint8_t* owner, const float* noise, float* out_latent, float* out_x0, void* stream);

/* ---- K7 alone: x <- a_k*x + b_k*eps_k, k = 0..n-1 in sequence (ed:692-704) ---------------------------------- */
int ed_renoise(const ed_step_params_t* d_params, const float* x, const float* noise, float* out,
               int64_t numel, void* stream);

/* ---- K12: ControlNet condition batch of a wave (elastic_diffusion_w_controlnet.py, "cn:N") ---------------------------
 * Replaces F.pad of the condition for padded global passes (cn:457-461) and, for the local views, the nearest upsample
 * of condition_image[0:1] to full pixel size + crop_with_context at 8x coordinates + cat (cn:933, 946-949, 959-961).
 * The condition never changes during a call, so this runs ONCE per wave shape, not per step.
 *   cond     (2, CH, ch, cw) fp32 : prepared, CFG-doubled condition image, ch = lh*scale, cw = lw*scale (cn:1183-1193)
 *   row_map  [H*scale], col_map [W*scale] : source row / col of the nearest upsample to (H*scale, W*scale)
 *   vorigin  [nv*2] : pixel origin (row, col) of every view's context box in the upsampled image
 *   out      (2*B*R1 + nv*B, CH, dH*scale, dW*scale) of out_dtype: sample (k,s,b) = zero-padded cond[s];
 *            sample (view v, b) = upsampled cond[0] box, zero-padded when the view is smaller than native */
int ed_gather_cond(const ed_plan_t* plan, int R1, const float* cond, int CH, int scale, const int32_t* row_map,
                   const int32_t* col_map, const int32_t* vorigin, void* out, int out_dtype, void* stream);

/* ---- K10: tiled decode glue -- replaces F.pad + per-tile slicing/cat (ed:287-300) and the accumulate /
 * count / divide blend (ed:303-308) -----------------------------------------------------------------------------
 * Tile table (DEVICE int32): tiles[j*4 .. j*4+3] = h0,h1,w0,w1 of core tile j in latent units, row-major tile grid
 * (get_views(core, core, stride), ed:287).
 * ed_tile_gather: out[(j*B+b), c, :, :] = zero-padded latent box of side T = core+2*pad around tile j.
 *                 fp32 TMA box loads with out-of-bounds zero fill (== F.pad(..., 'constant', 0), ed:289). */
int ed_tile_gather(const float* latent, int B, int C, int H, int W, const int32_t* tiles_dev, int ntiles, int core,
                   int pad, float* out, void* stream);

typedef struct {
  int32_t ntiles, ntc;        /* number of tiles, tile-grid columns (tile j = r*ntc + c) */
  int32_t core, pad, scale;   /* latent units; scale = vae_scale_factor (8) */
  int32_t B, CH, H, W;        /* image batch, image channels (3), latent H, W */
  int32_t reserved;
  const int32_t* tiles;       /* [ntiles*4] */
  const int32_t* trow_first;  /* [H] first tile-grid row covering latent row y, and count (consecutive) */
  const int32_t* trow_cnt;
  const int32_t* tcol_first;  /* [W] */
  const int32_t* tcol_cnt;
} ed_tiles_t;

/* ed_tile_blend: image[b,ch,y,x] = ( sum over covering tiles j, ascending, of clamp(patch_j/2 + 0.5, 0, 1) ) / count
 *   patches (ntiles*B, CH, T*scale, T*scale) raw VAE decoder output of dtype patch_dtype (decode_latents' /2+.5
 *   and clamp, ed:271, are fused here); image (B, CH, H*scale, W*scale) fp32. */
int ed_tile_blend(const ed_tiles_t* tiles, const void* patches, int patch_dtype, float* image, void* stream);

/* The same blend with the decode tiles sharded over `world` ranks ("tiled decode shards by tile"):
 * rank r decoded patches [r*per, (r+1)*per) of the (j*B + b) order into its own buffer; `d_peer_patches` is a DEVICE array
 * of `world` peer-mapped base pointers (symmetric memory) and every rank's kernel reads the centre crops it needs straight
 * from their owners over NVLink - no gathered copy of the 16x larger padded patches, no collective.  The caller orders
 * "all ranks finished decoding" before the launch (device-side symmetric-memory barrier). */
int ed_tile_blend_peer(const ed_tiles_t* tiles, const void* const* d_peer_patches, int world, int per, int patch_dtype,
                       float* image, void* stream);

/* ---- opt-in: fused element-wise / normalisation ops INSIDE the UNet forward (unet_ops.py) -----------
 * Not part of the reference's loop (its UNet is diffusers'); they replace the two largest non-GEMM costs of the SDXL-shaped
 * UNet forward that every wave of the loop spends its time in.
 * ed_geglu: x (M, 2N) contiguous of `dtype`, out (M, N) = x[:, :N] * gelu(x[:, N:]) (erf form) with the intermediate rounding
 *           of gelu to `dtype` that torch's two kernels apply -> bit-identical to `a, g = x.chunk(2, -1); a * F.gelu(g)`.
 *           N % 8 == 0, 16-byte aligned pointers (else ED_ERR_UNSUPPORTED). */
int ed_geglu(const void* x, void* out, int64_t M, int N, int dtype, void* stream);
/* ed_groupnorm_silu: out = [silu](group_norm(x, G, gamma, beta, eps)) for contiguous NCHW x (N, C, HW) of `dtype` (gamma /
 *           beta of the same dtype or NULL).  `workspace`: N*G*ed_groupnorm_split(N,C,HW,G)*2 floats.  Two launches: split
 *           statistics (N*G*S CTAs, shifted sums) + apply.  HW % 8 == 0, 16-byte aligned pointers. */
/* ed_bias_add: conv epilogue in place on contiguous NCHW y (N, C, HW): y += bias[c] (+ per_nc[n*C + c]) (+ residual[n, c, :]), each
 *           step rounded to `dtype` like the separate torch ops (cuDNN conv -> add_(bias) -> + time-embedding -> + residual):
 *           bit-identical, one vectorised pass instead of two or three unvectorised broadcast adds.  Any of the three may be NULL. */
int ed_bias_add(void* y, const void* bias, const void* per_nc, const void* residual, int N, int C, int HW, int dtype, void* stream);
/* ed_bias_add_nhwc: the same epilogue for a conv output left in NHWC memory order (what cuDNN's tensor-core kernels produce when the
 *           weights are kept channels-last, so that no per-call weight transform runs): reads (N, HW, C), writes contiguous NCHW,
 *           i.e. it also replaces cuDNN's separate nhwcToNchw pass.  C % 64 == 0, HW % 64 == 0. */
int ed_bias_add_nhwc(const void* y_nhwc, void* out_nchw, const void* bias, const void* per_nc, const void* residual, int N, int C,
                     int HW, int dtype, void* stream);
/* ed_layernorm: LayerNorm over the last dimension of contiguous x (M, D): one warp per row, the row in registers, two-pass
 *           statistics.  D % 8 == 0, D <= 2048, 16-byte aligned pointers; gamma / beta of the same dtype or NULL. */
int ed_layernorm(const void* x, const void* gamma, const void* beta, void* out, int64_t M, int D, float eps, int dtype, void* stream);
int ed_groupnorm_split(int N, int C, int HW, int G);
int ed_groupnorm_silu(const void* x, const void* gamma, const void* beta, void* out, float* workspace, int N, int C, int HW,
                      int G, float eps, int silu, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ELASTIC_B200_H_ */
