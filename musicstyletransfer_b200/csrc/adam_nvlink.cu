// K5 — data-parallel optimiser step as ONE kernel over NVLink peer memory: reduce-scatter of the gradient arenas,
// Adam on the owned slice, all-gather of the updated parameters.
//
// The reference has no multi-GPU path; its averaging rule is fixed by loss.backward() + gluon.Trainer.step(batch_size)
// (/root/reference/music_style_transfer/VarAutoEncoder/trainer.py:176-177): gradients are SUMMED over the global batch
// and rescaled by 1 / batch_size.  The baseline realisation is ncclAllReduce(sum) on the flat gradient arena followed by
// msx_adam_step on every rank (each rank redoing the whole 2 M-parameter update).  Here every rank's gradient and
// parameter arenas are peer-mapped (torch symmetric memory -> plain device pointers in this ABI) and rank r
//   1. waits on an in-kernel flag barrier until every rank's backward has produced its gradients,
//   2. pulls slice r of every rank's gradient arena over NVLink (P2P loads), sums in rank order (deterministic and
//      identical on every rank), applies the MXNet-1.3 Adam rule to slice r with its LOCAL moments (the moments of the
//      other slices are never touched: optimiser state is sharded 1 / world),
//   3. pushes the updated parameters of slice r into every rank's parameter arena (P2P stores),
//   4. the last CTA signals "slice r published, my reads of your gradients are finished" and waits for the same signal
//      from every peer, so that when the kernel retires the local parameter arena is complete and the local gradient
//      arena may be overwritten.
// Per rank this moves (world-1)/world * 2 * 4 B per parameter over NVLink (14 MB at world = 8) instead of the ring
// all-reduce's traffic plus a full-arena Adam pass, in one launch that is CUDA-graph capturable.
#include "msx_common.cuh"

namespace {

constexpr int kMaxWorld = 16;

struct NvlArgs {
  float* w; float* m; float* v;
  const float* peer_g[kMaxWorld];
  float* peer_w[kMaxWorld];
  unsigned long long* peer_flags[kMaxWorld];   // per rank: [2][kMaxWorld] u64 (entry / exit epoch written by rank src)
  unsigned* done_counter;                      // local, zero between launches
  const float* state;                          // [t, lr_t] (adam_tick)
  long long lo4, hi4;                          // owned slice in float4 units
  int rank, world;
  unsigned long long* epoch_ctr;               // local device counter: this launch is epoch *ctr + 1, the last CTA increments it
  float b1, b2, eps, wd, rescale, clip;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Spin until *flag >= epoch.  A peer that never launches (crashed rank) must not hang the job: after 30 s the kernel
// traps, which surfaces as a CUDA error on this rank instead of a silent dead-lock.
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch) {
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (ld_acquire_sys(flag) < epoch) {
    if ((++spins & 0xFFFu) == 0) {
      __nanosleep(200);
      if (globaltimer_ns() - t0 > 30000000000ull) __trap();
    }
  }
}

__global__ void __launch_bounds__(256) adam_nvlink_kernel(const NvlArgs a) {
  unsigned long long* my_flags = a.peer_flags[a.rank];
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(a.epoch_ctr) + 1ull;
  // ---- 1. entry barrier: the kernel of rank p is stream-ordered after rank p's backward
  if (blockIdx.x == 0 && threadIdx.x < a.world) st_release_sys(a.peer_flags[threadIdx.x] + a.rank, epoch);
  if (threadIdx.x < a.world) {
    wait_flag(my_flags + threadIdx.x, epoch);
  }
  __syncthreads();
  // ---- 2 + 3. pull + reduce + Adam + push on the owned slice
  const float lr_t = a.state[1];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = a.lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.hi4; i += stride) {
    float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < a.world; ++p) {
      // peer gradients are written by another GPU's kernels: read them through the coherent path, not the read-only cache
      const float4 gp = __ldcv(reinterpret_cast<const float4*>(a.peer_g[p]) + i);
      gs.x += gp.x; gs.y += gp.y; gs.z += gp.z; gs.w += gp.w;
    }
    float4 wv = reinterpret_cast<float4*>(a.w)[i];
    float4 mv = reinterpret_cast<float4*>(a.m)[i], vv = reinterpret_cast<float4*>(a.v)[i];
    float* wp = &wv.x; float* gp = &gs.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg = gp[j] * a.rescale + a.wd * wp[j];
      if (a.clip > 0.f) gg = fminf(fmaxf(gg, -a.clip), a.clip);
      mp[j] = a.b1 * mp[j] + (1.f - a.b1) * gg;
      vp[j] = a.b2 * vp[j] + (1.f - a.b2) * gg * gg;
      wp[j] -= lr_t * mp[j] / (sqrtf(vp[j]) + a.eps);
    }
    reinterpret_cast<float4*>(a.m)[i] = mv;
    reinterpret_cast<float4*>(a.v)[i] = vv;
    for (int p = 0; p < a.world; ++p) reinterpret_cast<float4*>(a.peer_w[p])[i] = wv;
  }
  // ---- 4. exit barrier by the last CTA of this rank
  __threadfence_system();
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(a.done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) *a.done_counter = 0;
  __threadfence_system();
  if (threadIdx.x < a.world) {
    st_release_sys(a.peer_flags[threadIdx.x] + kMaxWorld + a.rank, epoch);
    wait_flag(my_flags + kMaxWorld + threadIdx.x, epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.epoch_ctr = epoch;   // every other CTA of this launch has read the counter before arriving at done_counter
}

}  // namespace

// Bytes of the per-rank flag block that must be peer-mapped and zero-initialised once.
extern "C" int msx_adam_nvlink_flag_bytes(void) { return 2 * kMaxWorld * (int)sizeof(unsigned long long); }

// peer_g / peer_w / peer_flags: HOST arrays of `world` device pointers (entry `rank` is the local buffer).
// done_counter: local device u32, zero.  epoch_counter: local device u64, zero at start; the kernel numbers its launches
// with it (launch k uses epoch k on every rank), which keeps the launch CUDA-graph replayable.
// The gradient arena is zeroed after the kernel (its completion implies that no peer reads it any more).
extern "C" int msx_adam_nvlink_step(float* w, float* g, float* m, float* v, long long n, float* state,
                                    const float* const* peer_g, float* const* peer_w,
                                    unsigned long long* const* peer_flags, unsigned* done_counter, int rank, int world,
                                    unsigned long long* epoch_counter, float lr, float beta1, float beta2, float eps, float wd,
                                    float rescale, float clip, int zero_grad, int max_ctas, void* stream);

extern "C" void msx_adam_tick_launch(float* state, float lr, float b1, float b2, void* stream);

extern "C" int msx_adam_nvlink_step(float* w, float* g, float* m, float* v, long long n, float* state,
                                    const float* const* peer_g, float* const* peer_w,
                                    unsigned long long* const* peer_flags, unsigned* done_counter, int rank, int world,
                                    unsigned long long* epoch_counter, float lr, float beta1, float beta2, float eps, float wd,
                                    float rescale, float clip, int zero_grad, int max_ctas, void* stream) {
  MSX_REQUIRE(w && g && m && v && state && peer_g && peer_w && peer_flags && done_counter && epoch_counter, "msx_adam_nvlink_step: null pointer");
  MSX_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "msx_adam_nvlink_step: bad rank / world (%d / %d)", rank, world);
  MSX_REQUIRE((n & 3) == 0, "msx_adam_nvlink_step: the arena length must be a multiple of 4 (ParamArena pads every tensor)");
  MSX_REQUIRE(peer_g[rank] == g && peer_w[rank] == w, "msx_adam_nvlink_step: peer table entry `rank` must be the local arena");
  if (n == 0) return MSX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  msx_adam_tick_launch(state, lr, beta1, beta2, stream);
  NvlArgs a;
  a.w = w; a.m = m; a.v = v;
  for (int p = 0; p < world; ++p) {
    MSX_REQUIRE(peer_g[p] && peer_w[p] && peer_flags[p], "msx_adam_nvlink_step: null peer pointer (rank %d)", p);
    a.peer_g[p] = peer_g[p]; a.peer_w[p] = peer_w[p]; a.peer_flags[p] = peer_flags[p];
  }
  a.done_counter = done_counter; a.state = state; a.rank = rank; a.world = world; a.epoch_ctr = epoch_counter;
  const long long n4 = n / 4, per = (n4 + world - 1) / world;
  a.lo4 = per * rank < n4 ? per * rank : n4;
  a.hi4 = per * (rank + 1) < n4 ? per * (rank + 1) : n4;
  a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = wd; a.rescale = rescale; a.clip = clip;
  // all CTAs of all ranks spin in the entry barrier: keep the grid well inside one wave
  int grid = (int)((a.hi4 - a.lo4 + 255) / 256);
  const int cap = max_ctas > 0 ? max_ctas : msx_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  adam_nvlink_kernel<<<grid, 256, 0, st>>>(a);
  MSX_LAUNCH_CHECK();
  if (zero_grad) MSX_CUDA(cudaMemsetAsync(g, 0, (size_t)n * sizeof(float), st));
  return MSX_OK;
}
