// TEST INFRASTRUCTURE - host emulation driver of the staged wave epilogue (see emu_shim.h).  Exports
// emu_wave_epilogue(...) with the argument list of ed_wave_epilogue, all pointers HOST pointers; launch geometry comes from
// the same staged_config() the CUDA launcher uses.  Built by tests/test_kernel_emu.py with
//   g++ -O1 -std=c++17 -ffp-contract=off -DED_HOST_EMU -I tests/emu -I include -I <pkg>/csrc -shared -fPIC -pthread
#include <condition_variable>
#include <mutex>
#include <vector>

#include "epilogue_half.cuh"
#include "epilogue_staged.cuh"

thread_local dim3 threadIdx, blockIdx;
dim3 blockDim, gridDim;
std::atomic<long long> ed_emu_counters[8];
namespace ed {
uint8_t* emu_dyn_smem = nullptr;
}

namespace {
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  unsigned n = 0, waiting = 0, gen = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++waiting == n) {
      waiting = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};
Barrier g_barrier;
}  // namespace

void __syncthreads() { g_barrier.wait(); }

template <typename OT, bool RENOISE, int CPT>
static int run_as(const ed::EpiArgs& A, int so, int sms, int* info) {
  const ed_plan_t& P = A.P;
  ed::StagedCfg cfg = ed::staged_config(P, A.R1, so, sms, info ? info[5] : 3, CPT);
  if (!cfg.ok) return ED_ERR_UNSUPPORTED;
  cfg.g.vec_views = (info && info[7]) ? 0 : 1;   // test hook: scalar view loads
  if (info && info[6] > 0) {   // test hook: provision boxes `info[6]` rows too small -> tiles must take the global path
    cfg.g.bh = cfg.g.bh - info[6] > 0 ? cfg.g.bh - info[6] : 1;
    cfg.g.stage_bytes = ((unsigned)(cfg.g.bw * cfg.g.bh * CPT * so) + 127u) & ~127u;
    cfg.smem = (size_t)A.R1 * 2 * cfg.g.stage_bytes + (size_t)cfg.g.bw * cfg.g.bh * CPT * 4;
  }
  for (auto& c : ed_emu_counters) c = 0;
  if (info) {
    info[0] = cfg.g.bx; info[1] = cfg.g.by; info[2] = cfg.g.bw; info[3] = cfg.g.bh;
    info[4] = (int)cfg.smem; info[5] = cfg.grid_x * cfg.grid_y * cfg.grid_z;
  }
  const long long n_samples = 2LL * P.B * A.R1 + (long long)P.nv * P.B;
  ed::EmuTensorMap tm{static_cast<const uint8_t*>(A.unet_out), P.dW, P.dH, n_samples * P.C, cfg.g.bw, cfg.g.bh, CPT, so};
  std::vector<uint8_t> smem(cfg.smem + 128);
  ed::emu_dyn_smem = smem.data();
  blockDim = dim3(cfg.g.bx, cfg.g.by, 1);
  gridDim = dim3(cfg.grid_x, cfg.grid_y, cfg.grid_z);
  const unsigned nthreads = blockDim.x * blockDim.y;
  g_barrier.n = nthreads;
  for (unsigned bz = 0; bz < gridDim.z; ++bz)
    for (unsigned by = 0; by < gridDim.y; ++by)
      for (unsigned bx = 0; bx < gridDim.x; ++bx) {
        memset(smem.data(), 0xA5, smem.size());   // stale shared memory must never be read
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
          th.emplace_back([&, t] {
            threadIdx = dim3(t % blockDim.x, t / blockDim.x, 0);
            blockIdx = dim3(bx, by, bz);
            ed::wave_epilogue_staged_kernel<OT, RENOISE, CPT>(tm, A, cfg.g);
          });
        for (auto& x : th) x.join();
      }
  if (info)
    for (int i = 0; i < 7; ++i) info[8 + i] = (int)ed_emu_counters[i];
  return ED_OK;
}

// same dispatch as launch_staged() in csrc/epilogue.cu; info[9] (in) = channels per thread of the no-noise instantiation
template <typename OT>
static int run(const ed::EpiArgs& A, int so, int sms, int* info) {
  if (A.noise) return run_as<OT, true, 4>(A, so, sms, info);
  if (info && info[9] == 4) return run_as<OT, false, 4>(A, so, sms, info);
  return run_as<OT, false, 2>(A, so, sms, info);
}

extern "C" int emu_wave_epilogue(const ed_plan_t* plan, const ed_step_params_t* params, int R1, const float* latent,
                                 const void* unet_out, int out_dtype, const uint8_t* idx, const uint8_t* owner,
                                 const float* noise, float* out_latent, float* out_x0, int sms, int* info) {
  ed::EpiArgs A{*plan, params, latent, unet_out, nullptr, 0, 0, idx, owner, noise, out_latent, out_x0, R1};
  switch (out_dtype) {
    case ED_F32: return run<float>(A, 4, sms, info);
    case ED_F16: return run<__half>(A, 2, sms, info);
    case ED_BF16: return run<__nv_bfloat16>(A, 2, sms, info);
  }
  return ED_ERR_INVALID;
}

// ---- half kernels (csrc/epilogue_half.cuh): no barrier, no TMA -> the threads of a CTA run one after the other --------------
template <typename OT, bool MULTI, bool PEER>
static int run_half_as(const ed::EpiArgs& A, int so, int* info) {
  const ed::HalfCfg cfg = ed::half_config(A.P, A.R1, so);
  if (!cfg.ok || A.noise) return ED_ERR_UNSUPPORTED;
  std::vector<uint8_t> smem(cfg.smem + 128);
  ed::emu_dyn_smem = smem.data();
  blockDim = dim3(cfg.bx, cfg.by, 1);
  gridDim = dim3(cfg.grid_x, cfg.grid_y, info && info[0] > 0 ? (unsigned)info[0] : (unsigned)cfg.grid_z);   // test hook: short grid z
  if (info) { info[1] = cfg.bx; info[2] = cfg.by; info[3] = (int)cfg.smem; info[4] = cfg.grid_x * cfg.grid_y * cfg.grid_z; }
  for (unsigned bz = 0; bz < gridDim.z; ++bz)
    for (unsigned by = 0; by < gridDim.y; ++by)
      for (unsigned bx = 0; bx < gridDim.x; ++bx) {
        memset(smem.data(), 0xA5, smem.size());
        blockIdx = dim3(bx, by, bz);
        for (unsigned t = 0; t < blockDim.x * blockDim.y; ++t) {
          threadIdx = dim3(t % blockDim.x, t / blockDim.x, 0);
          ed::wave_epilogue_half_kernel<OT, MULTI, PEER>(A);
        }
      }
  return ED_OK;
}

template <typename OT>
static int run_half(const ed::EpiArgs& A, int so, int* info) {
  if (A.peers) return A.R1 > 1 ? run_half_as<OT, true, true>(A, so, info) : run_half_as<OT, false, true>(A, so, info);
  return A.R1 > 1 ? run_half_as<OT, true, false>(A, so, info) : run_half_as<OT, false, false>(A, so, info);
}

// peers != NULL: HOST array of `world` buffer pointers (the multi-GPU entry point's sample -> rank mapping)
extern "C" int emu_wave_epilogue_half(const ed_plan_t* plan, const ed_step_params_t* params, int R1, const float* latent,
                                      const void* unet_out, const void* const* peers, int world, int per, int out_dtype,
                                      const uint8_t* idx, const uint8_t* owner, float* out_latent, float* out_x0, int* info) {
  ed::EpiArgs A{*plan, params, latent, unet_out, peers, world, per, idx, owner, nullptr, out_latent, out_x0, R1};
  switch (out_dtype) {
    case ED_F32: return run_half<float>(A, 4, info);
    case ED_F16: return run_half<__half>(A, 2, info);
    case ED_BF16: return run_half<__nv_bfloat16>(A, 2, info);
  }
  return ED_ERR_INVALID;
}
