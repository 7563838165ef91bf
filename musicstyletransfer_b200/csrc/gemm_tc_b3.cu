// K2 (forward GEMMs of the tf32x3f / bf16x3f modes) — "bf16x3" GEMM on tcgen05 for sm_100a.
//
// Y = X W^T + b with fp32 X, W, Y in HBM (the gluon.nn.Dense forwards of
// /root/reference/music_style_transfer/VarAutoEncoder/transformer.py:36-40,65-68,88-93,104 and model.py:100), products to
// ~2^-17 relative: every fp32 operand value is split inside the kernel into two bfloat16 parts
//     hi = bf16_trunc(x),   lo = bf16_trunc(x - hi)            (hi + lo carries 16 mantissa bits)
// and a k-step contributes  A_lo B_hi + A_hi B_lo + A_hi B_hi  on `tcgen05.mma.kind::f16`, which runs at twice the
// `kind::tf32` rate: the three MMAs cost 1.5 single-pass TF32 MMAs instead of 3 (gemm_tc_x3.cu), and the kernel goes back
// to being bound by the fp32 operand / result bytes instead of the tensor pipe.  The latent-means error of the whole step
// with these GEMMs is the same 1e-4 as with 3xTF32 (the compensated attention scores dominate it); single-pass TF32 gives
// 1.3e-3.  Accumulation is fp32 in TMEM.  K-major operands only (X [M,K], W [N,K]: every forward GEMM of the step).
//
// Pipeline per CTA of a cta_group::2 pair (256 x BN2 tiles as in gemm_tc.cu / gemm_tc_x3.cu):
//   warp 0        TMA producer: raw fp32 k-blocks of 32 (A 128 x 128 B, B BN2/2 x 128 B, SWIZZLE_128B) into a ring of SLOTS;
//   warps 2..5    CONVERTER: slot (32 k of fp32) -> one UNIT = one half of the 128-byte rows of a 16-bit buffer (a bf16 row of
//                 the SWIZZLE_128B K-major layout holds 64 k): for row r and output chunk j the two fp32 chunks 2j, 2j+1
//                 (at (c ^ r%8) * 16 B of the row) become the hi and lo bf16 chunks at ((4 half + j) ^ r%8) * 16 B of the
//                 hi / lo tiles, written directly in the swizzled layout; the slot is handed back (sfree), then
//                 `fence.proxy.async` and a cluster-scope arrive on the leader's conv[unit];
//   warp 1        MMA issuer (leader CTA): per unit 2 k-steps x 3 tcgen05.mma.cta_group::2.kind::f16, commit empty[unit]
//                 (the two halves of a buffer are independent pipeline stages: one is converted while the other multiplies);
//   warps 6..13   epilogue (shared with the other tensor GEMMs: bias / ReLU / dropout / ReLU bit mask / TMA store).
#include "gemm_tc_common.cuh"

using namespace msx_tc;

namespace {

constexpr int kConvWarpsB = 4;
constexpr int kEpiWarpsB = 8;
constexpr int kThreadsB = 32 * (2 + kConvWarpsB + kEpiWarpsB);
constexpr int kMaxSlotsB = 4, kMaxUnitsB = 4;

template <int BN2>
struct B3Cfg {
  static constexpr int kBRows = BN2 / 2;
  static constexpr int kRows = BM + kBRows;            // operand rows per CTA and k-block (A rows, then B rows)
  static constexpr int kSlot = kRows * 128;            // fp32: 32 k x 4 B = 128 B per row
  static constexpr int kHalf16 = kRows * 128;          // bf16: 64 k x 2 B = 128 B per row; hi tiles (A | B), then lo tiles
  static constexpr int kStage = 2 * kHalf16;           // one 16-bit BUFFER = 64 k = two pipeline UNITS of 32 k (row halves)
  static constexpr int kBuffers = BN2 == 256 ? 1 : 2;
  static constexpr int kUnits = 2 * kBuffers;
  static constexpr int kSlots = 4;                     // fp32 k-blocks in flight: what hides the TMA round trip
  static constexpr int kTmem = 2 * BN2;
  static constexpr int kChunks = BN2 / 32;
};

struct __align__(8) BarriersB3 {
  unsigned long long full[kMaxSlotsB], sfree[kMaxSlotsB], conv[kMaxUnitsB], empty[kMaxUnitsB], tmem_full[2], tmem_empty[2];
  unsigned tmem_base;
};

// eight fp32 values (two 16-byte chunks) -> their bf16 hi parts and the bf16 of the remainders.  Both parts are taken by
// TRUNCATION (the upper 16 bits of the fp32 word): hi = x & 0xFFFF0000 is exact to subtract, lo = upper half of (x - hi), so
// the split costs one PRMT per packed pair and two LOP + two FADD per pair instead of four quarter-rate F2F conversions
// (the converter warps were the pipeline's bottleneck with cvt.rn).  hi + lo then carries 16 mantissa bits truncated:
// <= 2^-16 relative, one-sided; measured against float64 on the step's shapes it stays below 1e-5 of the output scale.
__device__ __forceinline__ unsigned hi16_pair(unsigned a, unsigned b) {      // {upper half of b, upper half of a}
  unsigned r;
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void split8(const uint4& v0, const uint4& v1, uint4& hi, uint4& lo) {
  const unsigned x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  unsigned h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned a = x[2 * i], b = x[2 * i + 1];
    h[i] = hi16_pair(a, b);                                              // low half = element 2i, high half = element 2i + 1
    const float ra = __uint_as_float(a) - __uint_as_float(a & 0xFFFF0000u);
    const float rb = __uint_as_float(b) - __uint_as_float(b & 0xFFFF0000u);
    l[i] = hi16_pair(__float_as_uint(ra), __float_as_uint(rb));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int BN2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsB, 1)
    gemm_tc2b3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const TcParams p) {
  pdl_entry();
  using Cfg = B3Cfg<BN2>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stages = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* slots = stages + Cfg::kBuffers * Cfg::kStage;
  unsigned char* stage_out = slots + Cfg::kSlots * Cfg::kSlot;                   // [kEpiWarpsB][32 rows][128 B]
  BarriersB3* bars = reinterpret_cast<BarriersB3*>(stage_out + kEpiWarpsB * kOutBoxBytes);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const unsigned rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int items = p.m_tiles * p.n_tiles * p.splitk;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kSlots; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->sfree[s], kConvWarpsB); }
    for (int s = 0; s < Cfg::kUnits; ++s) { mbar_init(&bars->conv[s], 2 * kConvWarpsB); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], 2 * kEpiWarpsB); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "n"(Cfg::kTmem)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================================ TMA producer: raw fp32 k-blocks into the slot ring ================================
    int slot = 0;
    unsigned phase = 0;
    for (int it = pair; it < items; it += npairs) {
      const int nt = it % p.n_tiles, mt = (it / p.n_tiles) % p.m_tiles, ks = it / (p.n_tiles * p.m_tiles);
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      const int m0 = mt * (2 * BM) + (int)rank * BM;
      const int n0 = nt * BN2 + (int)rank * Cfg::kBRows;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bars->sfree[slot], phase ^ 1);
        unsigned char* sa = slots + slot * Cfg::kSlot;
        if (elect_one()) {
          mbar_expect_tx(&bars->full[slot], Cfg::kSlot);
          tma_load_2d(sa, &tmA, &bars->full[slot], kb * 32, m0);                 // box {32 k (128 B), 128 rows}
          tma_load_2d(sa + BM * 128, &tmB, &bars->full[slot], kb * 32, n0);      // box {32 k, BN2/2 rows}
        }
        __syncwarp();
        if (++slot == Cfg::kSlots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA) ================================
    if (leader) {
      // D = F32, A = B = BF16, both K-major, N >> 3 @17, M >> 4 @24 with M = 256 across the pair
      const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(BN2 >> 3) << 17) | ((unsigned)((2 * BM) >> 4) << 24);
      const unsigned long long ad0 = make_desc(smem_u32(stages), 16, 1024, 2);
      const unsigned long long bd0 = make_desc(smem_u32(stages) + BM * 128, 16, 1024, 2);
      constexpr unsigned long long kLo = (unsigned long long)(Cfg::kHalf16 >> 4);
      int unit = 0;
      unsigned phase = 0;
      int local = 0;
      for (int it = pair; it < items; it += npairs, ++local) {
        const int ks = it / (p.n_tiles * p.m_tiles);
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int buf = local & 1;
        const unsigned use = (unsigned)(local >> 1);
        mbar_wait(&bars->tmem_empty[buf], (use & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned tmem_d = tmem_base + buf * BN2;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_cluster(&bars->conv[unit], phase);           // both CTAs' 16-bit half rows written and fenced
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // unit -> buffer unit / 2, row half unit % 2: the half's two k-steps sit 64 B into the 128 B swizzle row
          const unsigned long long off = (unsigned long long)((unit >> 1) * (Cfg::kStage >> 4) + (unit & 1) * 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {                         // 16 bf16 of k = 32 B per step
              const unsigned long long a = ad0 + off + 2 * k, b = bd0 + off + 2 * k;
              umma_ss_pair<true>(tmem_d, a + kLo, b, idesc, (kb > kb0 || k > 0) ? 1u : 0u);   // A_lo B_hi
              umma_ss_pair<true>(tmem_d, a, b + kLo, idesc, 1u);                               // A_hi B_lo
              umma_ss_pair<true>(tmem_d, a, b, idesc, 1u);                                     // A_hi B_hi
            }
            umma_commit_pair(&bars->empty[unit]);
            if (kb == kb1 - 1) umma_commit_pair(&bars->tmem_full[buf]);
          }
          __syncwarp();
          if (++unit == Cfg::kUnits) { unit = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 2 + kConvWarpsB) {
    // ================================ converter (both CTAs) ================================
    const int ctid = threadIdx.x - 64;                    // 0 .. 127
    const unsigned conv_leader = mapa_shared(smem_u32(&bars->conv[0]), 0);
    int slot = 0, unit = 0;
    unsigned sphase = 0, uphase = 0;
    for (int it = pair; it < items; it += npairs) {
      const int ks = it / (p.n_tiles * p.m_tiles);
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bars->empty[unit], uphase ^ 1);        // the MMAs that read this unit's previous contents have retired
        mbar_wait(&bars->full[slot], sphase);             // this CTA's fp32 k-block has landed
        unsigned char* hi_t = stages + (unit >> 1) * Cfg::kStage;
        unsigned char* lo_t = hi_t + Cfg::kHalf16;
        const unsigned char* src = slots + slot * Cfg::kSlot;
        const int half = unit & 1;
#pragma unroll 2
        for (int i = ctid; i < Cfg::kRows * 4; i += 32 * kConvWarpsB) {
          const int r = i >> 2, j = i & 3, sw = r & 7;
          const uint4 v0 = *reinterpret_cast<const uint4*>(src + r * 128 + (((2 * j) ^ sw) << 4));
          const uint4 v1 = *reinterpret_cast<const uint4*>(src + r * 128 + (((2 * j + 1) ^ sw) << 4));
          uint4 h, l;
          split8(v0, v1, h, l);
          const int off = r * 128 + (((4 * half + j) ^ sw) << 4);
          *reinterpret_cast<uint4*>(hi_t + off) = h;
          *reinterpret_cast<uint4*>(lo_t + off) = l;
        }
        // generic-proxy writes of this thread -> visible to the async proxy (the tensor core's operand reads, also the
        // ones the pair leader issues against this CTA's shared memory).  The .shared::cta form is a FENCE.VIEW.ASYNC.S; the
        // unqualified fence.proxy.async compiles to MEMBAR.ALL.GPU + CCTL.IVALL + ERRBAR per stage (ncu source page: ~12 %
        // of all stall samples and an L1 invalidation per stage)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->sfree[slot]);                // this warp's reads of the slot are done
          mbar_arrive_cluster(conv_leader + (unsigned)(unit * sizeof(unsigned long long)));
        }
        if (++slot == Cfg::kSlots) { slot = 0; sphase ^= 1; }
        if (++unit == Cfg::kUnits) { unit = 0; uphase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue (8 warps, both CTAs) ================================
    const int ew = warp - (2 + kConvWarpsB);
    const int lg = warp & 3;
    const int chalf = ew >> 2;
    unsigned char* st = stage_out + ew * kOutBoxBytes;
    const bool reduce = p.accumulate || p.splitk > 1;
    int local = 0, sbuf = 0, pending = 0;
    for (int it = pair; it < items; it += npairs, ++local) {
      const int nt = it % p.n_tiles, mt = (it / p.n_tiles) % p.m_tiles;
      const int buf = local & 1;
      const unsigned use = (unsigned)(local >> 1);
      const int row0 = mt * (2 * BM) + (int)rank * BM + lg * 32;
      const int my_row = row0 + lane;
      constexpr int kChPerWarp = Cfg::kChunks / 2;
      AuxPref apre;
      if (p.aux) aux_prefetch(p, my_row, nt * BN2 + chalf * kChPerWarp * 32, apre);
      mbar_wait(&bars->tmem_full[buf], use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int ch = chalf * kChPerWarp; ch < (chalf + 1) * kChPerWarp; ++ch) {
        const int col0 = nt * BN2 + ch * 32;
        unsigned amask = 0u;
        if (p.aux) {
          amask = aux_mask(p, apre);
          if (ch + 1 < (chalf + 1) * kChPerWarp) aux_prefetch(p, my_row, col0 + 32, apre);
        }
        float v[32];
        tmem_ld32(tmem_base + ((unsigned)(lg * 32) << 16) + buf * BN2 + ch * 32, v);
        if (col0 < p.N && row0 < p.M) {
          epilogue_chunk<kEpiWarpsB, 1>(p, &tmC, v, row0, my_row, col0, lane, st, sbuf, pending, reduce, amask);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&bars->tmem_empty[buf]), 0));
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmem) : "memory");
  }
}

template <int BN2>
constexpr size_t b3_smem_bytes() {
  return 1024 + (size_t)B3Cfg<BN2>::kBuffers * B3Cfg<BN2>::kStage + (size_t)B3Cfg<BN2>::kSlots * B3Cfg<BN2>::kSlot +
         (size_t)kEpiWarpsB * kOutBoxBytes + sizeof(BarriersB3);
}

template <int BN2>
int launch_b3(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcParams& p, cudaStream_t st) {
  constexpr size_t smem = b3_smem_bytes<BN2>();
  static_assert(smem <= 232448, "bf16x3 kernel exceeds the 227 KB shared-memory limit");
  MSX_CUDA(cudaFuncSetAttribute(gemm_tc2b3_kernel<BN2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = p.m_tiles * p.n_tiles * p.splitk;
  const int max_pairs = msx_num_sms() / 2;
  const int pairs = items < max_pairs ? items : max_pairs;
  MSX_CUDA(msx_launch(gemm_tc2b3_kernel<BN2>, dim3(2 * pairs), dim3(kThreadsB), smem, st, ta, tb, tc, p));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

}  // namespace

// 1 when msx_gemm_tc_b3 takes the problem (forward form only: A [M,K] and B [N,K] row-major, i.e. transA = 0, transB = 1):
// TMA alignment rules of msx_gemm_tc and N >= 64; any M.
extern "C" int msx_gemm_tc_b3_supported(const float* A, int lda, const float* B, int ldb, const float* C, int ldc, int M,
                                        int N, int K) {
  if (!A || !B || !C || M < 1 || N < 64 || K <= 0) return 0;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15) || (lda & 3) || (ldb & 3) || (ldc & 3)) return 0;
  return 1;
}

extern "C" int msx_gemm_tc_b3(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                              const float* bias, int relu, float drop_p, unsigned long long seed, unsigned site,
                              int accumulate, uint32_t* mask_out, int ldmask, void* stream) {
  MSX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "msx_gemm_tc_b3: negative dimension");
  if (M == 0 || N == 0) return MSX_OK;
  MSX_REQUIRE(A && B && C, "msx_gemm_tc_b3: null operand");
  MSX_REQUIRE(K > 0, "msx_gemm_tc_b3: K must be > 0");
  if (!msx_gemm_tc_b3_supported(A, lda, B, ldb, C, ldc, M, N, K)) {
    msx_set_error("msx_gemm_tc_b3: needs N >= 64, 16-byte aligned operands and leading dimensions %% 4 == 0");
    return MSX_ERR_UNSUPPORTED;
  }
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_gemm_tc_b3: dropout probability must be in [0,1)");
  MSX_REQUIRE(!(mask_out && ((N & 31) || accumulate || ldmask < N / 32)), "msx_gemm_tc_b3: mask_out needs N %% 32 == 0, a plain store and ldmask >= N / 32");
  const int pad256 = msx_ceil_div(N, 256) * 256, pad128 = msx_ceil_div(N, 128) * 128;
  const int bn2 = (N > 128 && pad256 <= pad128) ? 256 : 128;
  CUtensorMap ta, tb, tc;
  int rc = make_map(&tc, C, M, N, ldc, 32, 32, false, kMapC32);
  if (rc) return rc;
  if ((rc = make_map(&ta, A, M, K, lda, 32, BM, false, kMapF32Op))) return rc;
  if ((rc = make_map(&tb, B, N, K, ldb, 32, bn2 / 2, false, kMapF32Op))) return rc;
  TcParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.relu = relu; p.drop_p = drop_p;
  p.inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.seed = seed; p.seed_ctr = msx_step_counter(); p.site = site; p.aux = nullptr; p.ldaux = 0; p.aux_scale = 1.f;
  p.accumulate = accumulate; p.out_colsum = nullptr; p.c_bf16 = 0; p.aux_bf16 = 0; p.mask_out = mask_out; p.ldmask = ldmask;
  p.dbg = 0;
  p.kb_total = msx_ceil_div(K, 32);
  p.m_tiles = msx_ceil_div(M, 2 * BM); p.n_tiles = msx_ceil_div(N, bn2);
  p.kb_per_split = p.kb_total;
  p.splitk = 1;
  cudaStream_t st = (cudaStream_t)stream;
  return bn2 == 256 ? launch_b3<256>(ta, tb, tc, p, st) : launch_b3<128>(ta, tb, tc, p, st);
}
