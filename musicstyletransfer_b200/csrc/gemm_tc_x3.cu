// K2 (strict-fp32 tensor-core path) — 3xTF32 GEMM on tcgen05 for sm_100a.
//
// Same contract as msx_gemm_tc / msx_gemm_f32 (the gluon.nn.Dense forward / backward GEMMs of
// /root/reference/music_style_transfer/VarAutoEncoder/{transformer.py:36-40,65-68,88-93,104, model.py:70-71,139-157,
// 214-227}), but the products carry fp32 precision: the reference step is fp32 end to end (trainer.py:155-179) and a
// single TF32 pass rounds each operand to 11 significant bits.  Every fp32 operand value x is split exactly into
//     hi = x with the low 13 mantissa bits cleared (what kind::tf32 reads of a raw fp32 word),  lo = x - hi  (exact),
// and a k-block contributes  A_lo B_hi + A_hi B_lo + A_hi B_hi  (three tcgen05.mma per k-step, small terms first); the
// dropped A_lo B_lo term and the rounding of lo to TF32 are each <= 2^-21 relative.  Accumulation is fp32 in TMEM.
//
// Structure = the cta_group::2 pair kernel of gemm_tc.cu (256 x BN2 tiles, TMA ring, double-buffered TMEM accumulators,
// epilogue through swizzled staging boxes + TMA store / reduce-add) plus one stage in the pipeline:
//   warp 0        TMA producer.  Operands are loaded as FLOAT32 (no rounding in the TMA unit), each CTA's bytes complete
//                 on its OWN full[s] barrier;
//   warps 2..9    CONVERTER (both CTAs; 8 warps: the LDS -> cvt -> STS rounds are latency-bound per warp, 4 warps measured
//                 6 % slower): wait full[s], read the stage's hi region (raw fp32 tiles in the swizzled layout
//                 TMA wrote), write lo = rna_tf32(x - trunc_tf32(x)) at the same offsets of the stage's lo region — the
//                 split is element-wise, so the swizzle is preserved without being decoded — then fence.proxy.async and
//                 arrive (cluster scope) on the leader's conv[s];
//   warp 1        MMA issuer (leader): waits conv[s], issues 3 x 4 tcgen05.mma.cta_group::2.kind::tf32, commits empty[s]
//                 to both CTAs;
//   warps 10..17  epilogue (8 warps, one staging box each).
// Shared memory per stage: hi (A 16 KB + B BN2/2 x 128 B) + the same again for lo; 3 stages (BN2 = 256) or 4 (BN2 = 128).
// The kernel is tensor-pipe bound (3 MMAs per operand byte), which also hides the converter: per stage it moves 64 KB
// through shared memory while the 12 MMAs take ~1536 cycles.
#include "gemm_tc_common.cuh"

using namespace msx_tc;

namespace {

#ifndef MSX_X3_CONV_WARPS
#define MSX_X3_CONV_WARPS 8
#endif
constexpr int kConvWarps = MSX_X3_CONV_WARPS;
constexpr int kEpiWarpsX3 = 8;
constexpr int kThreadsX3 = 32 * (2 + kConvWarps + kEpiWarpsX3);
constexpr int kMaxStagesX3 = 4;

template <int BN2>
struct X3Cfg {
  static constexpr int kBRows = BN2 / 2;
  static constexpr int kATile = BM * BK * 4;
  static constexpr int kBTile = kBRows * BK * 4;
  static constexpr int kHi = kATile + kBTile;          // raw fp32 tiles (read by the MMA as hi)
  static constexpr int kStage = 2 * kHi;               // + the lo tiles at the same offsets
  static constexpr int kStages = BN2 == 256 ? 3 : 4;
  static constexpr int kTmem = 2 * BN2;
  static constexpr int kChunks = BN2 / 32;
};

struct __align__(8) BarriersX3 {
  unsigned long long full[kMaxStagesX3], conv[kMaxStagesX3], empty[kMaxStagesX3], tmem_full[2], tmem_empty[2];
  unsigned tmem_base;
};

// lo part of four packed fp32 values: x - (x with the 13 low mantissa bits cleared), rounded to TF32 (rna)
__device__ __forceinline__ unsigned lo_tf32(unsigned xb) {
  const float x = __uint_as_float(xb);
  const float hi = __uint_as_float(xb & 0xFFFFE000u);
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x - hi));
  return r;
}

template <int BN2, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsX3, 1)
    gemm_tc2x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const TcParams p) {
  pdl_entry();
  using Cfg = X3Cfg<BN2>;
  using Op = OpCfg<false>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* stage_out = ring + Cfg::kStages * Cfg::kStage;                  // [kEpiWarpsX3][32 rows][128 B]
  BarriersX3* bars = reinterpret_cast<BarriersX3*>(stage_out + kEpiWarpsX3 * kOutBoxBytes);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const unsigned rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int items = p.m_tiles * p.n_tiles * p.splitk;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->conv[s], 2);                 // one arrival per CTA of the pair (after the converter warps' own barrier)
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], 2 * kEpiWarpsX3); }
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
    // ================================ TMA producer (both CTAs, each on its own full barrier) ================================
    int stage = 0;
    unsigned phase = 0;
    for (int it = pair; it < items; it += npairs) {
      const int nt = it % p.n_tiles, mt = (it / p.n_tiles) % p.m_tiles, ks = it / (p.n_tiles * p.m_tiles);
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      const int m0 = mt * (2 * BM) + (int)rank * BM;
      const int n0 = nt * BN2 + (int)rank * Cfg::kBRows;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        unsigned char* sa = ring + stage * Cfg::kStage;
        unsigned char* sb = sa + Cfg::kATile;
        if (elect_one()) {
          mbar_expect_tx(&bars->full[stage], Cfg::kHi);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &bars->full[stage], kb * Op::kBKE, m0);
          } else {
#pragma unroll
            for (int s = 0; s < BM / Op::kSlabMN; ++s)
              tma_load_2d(sa + s * Op::kSlabBytes, &tmA, &bars->full[stage], m0 + s * Op::kSlabMN, kb * Op::kBKE);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &bars->full[stage], kb * Op::kBKE, n0);
          } else {
#pragma unroll
            for (int s = 0; s < Cfg::kBRows / Op::kSlabMN; ++s)
              tma_load_2d(sb + s * Op::kSlabBytes, &tmB, &bars->full[stage], n0 + s * Op::kSlabMN, kb * Op::kBKE);
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA) ================================
    if (leader) {
      const unsigned idesc = (1u << 4) | (Op::kFmt << 7) | (Op::kFmt << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((unsigned)(BN2 >> 3) << 17) | ((unsigned)((2 * BM) >> 4) << 24);
      const unsigned long long ad0 = A_MN ? make_desc(smem_u32(ring), Op::kSlabBytes, Op::kMnSbo, Op::kMnLayout)
                                          : make_desc(smem_u32(ring), 16, 1024, 2);
      const unsigned long long bd0 = B_MN ? make_desc(smem_u32(ring) + Cfg::kATile, Op::kSlabBytes, Op::kMnSbo, Op::kMnLayout)
                                          : make_desc(smem_u32(ring) + Cfg::kATile, 16, 1024, 2);
      constexpr unsigned kAStep = A_MN ? Op::kMnStep : 2, kBStep = B_MN ? Op::kMnStep : 2;
      constexpr unsigned long long kLo = (unsigned long long)(Cfg::kHi >> 4);      // lo tiles: same layout, kHi bytes further
      int stage = 0;
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
          mbar_wait_cluster(&bars->conv[stage], phase);          // both CTAs' lo tiles written and fenced
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned long long ad = ad0 + (unsigned long long)(stage * (Cfg::kStage >> 4));
          const unsigned long long bd = bd0 + (unsigned long long)(stage * (Cfg::kStage >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const unsigned long long a = ad + kAStep * k, b = bd + kBStep * k;
              if (p.dbg & 2) {
                umma_ss_pair<false>(tmem_d, a, b, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                continue;
              }
              umma_ss_pair<false>(tmem_d, a + kLo, b, idesc, (kb > kb0 || k > 0) ? 1u : 0u);   // A_lo B_hi
              umma_ss_pair<false>(tmem_d, a, b + kLo, idesc, 1u);                               // A_hi B_lo
              umma_ss_pair<false>(tmem_d, a, b, idesc, 1u);                                     // A_hi B_hi
            }
            umma_commit_pair(&bars->empty[stage]);
            if (kb == kb1 - 1) umma_commit_pair(&bars->tmem_full[buf]);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 2 + kConvWarps) {
    // ================================ converter (both CTAs): lo tiles of every stage ================================
    const int ctid = threadIdx.x - 64;                    // 0 .. 32 * kConvWarps - 1
    const unsigned conv_leader = mapa_shared(smem_u32(&bars->conv[0]), 0);
    int stage = 0;
    unsigned phase = 0;
    for (int it = pair; it < items; it += npairs) {
      const int ks = it / (p.n_tiles * p.m_tiles);
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bars->full[stage], phase);             // this CTA's TMA bytes have landed
        const uint4* hi = reinterpret_cast<const uint4*>(ring + stage * Cfg::kStage);
        uint4* lo = reinterpret_cast<uint4*>(ring + stage * Cfg::kStage + Cfg::kHi);
#pragma unroll 4
        for (int i = ctid; i < ((p.dbg & 1) ? 0 : Cfg::kHi / 16); i += 32 * kConvWarps) {
          const uint4 v = hi[i];
          lo[i] = make_uint4(lo_tf32(v.x), lo_tf32(v.y), lo_tf32(v.z), lo_tf32(v.w));
        }
        // generic-proxy writes of this thread -> visible to the async proxy (the tensor core's operand reads, also the
        // ones the pair leader issues against this CTA's shared memory).  The .shared::cta form is a FENCE.VIEW.ASYNC.S; the
        // unqualified fence.proxy.async compiles to MEMBAR.ALL.GPU + CCTL.IVALL + ERRBAR per stage (ncu source page: ~12 %
        // of all stall samples and an L1 invalidation per stage)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // the converter warps meet on a named barrier and ONE thread signals the pair leader: 2 cluster-scope arrivals per
        // stage instead of 16 (each remote release-arrive is a DSMEM round trip on the stage's critical path)
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvWarps) : "memory");
        if (ctid == 0) mbar_arrive_cluster(conv_leader + (unsigned)(stage * sizeof(unsigned long long)));
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue (8 warps, both CTAs) ================================
    const int ew = warp - (2 + kConvWarps);
    const int lg = warp & 3;                 // TMEM lane group of this warp
    const int chalf = ew >> 2;               // warps 6..9 drain the first half of the columns, 10..13 the second
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
          epilogue_chunk<kEpiWarpsX3, 1>(p, &tmC, v, row0, my_row, col0, lane, st, sbuf, pending, reduce, amask);
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
constexpr size_t x3_smem_bytes() {
  return 1024 + (size_t)X3Cfg<BN2>::kStages * X3Cfg<BN2>::kStage + (size_t)kEpiWarpsX3 * kOutBoxBytes + sizeof(BarriersX3);
}

template <int BN2, bool A_MN, bool B_MN>
int launch_x3(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcParams& p, cudaStream_t st) {
  constexpr size_t smem = x3_smem_bytes<BN2>();
  static_assert(smem <= 232448, "3xTF32 kernel exceeds the 227 KB shared-memory limit");
  MSX_CUDA(cudaFuncSetAttribute(gemm_tc2x3_kernel<BN2, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = p.m_tiles * p.n_tiles * p.splitk;
  const int max_pairs = msx_num_sms() / 2;
  const int pairs = items < max_pairs ? items : max_pairs;
  MSX_CUDA(msx_launch(gemm_tc2x3_kernel<BN2, A_MN, B_MN>, dim3(2 * pairs), dim3(kThreadsX3), smem, st, ta, tb, tc, p));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}

}  // namespace

// 1 when msx_gemm_tc_x3 takes the problem: the TMA alignment rules of msx_gemm_tc and N >= 64.  Any M: a problem with fewer
// rows than a pair tile (the B SOS rows of the top encoder layer at small batch sizes) still runs as one 256-row tile whose
// out-of-range rows TMA zero-fills on load and clips on store — far faster than the FFMA kernel, which has two CTAs' worth
// of parallelism there (measured at M = 32, K = 1024: 150 us FFMA).  Narrower outputs belong to msx_gemm_f32.
extern "C" int msx_gemm_tc_x3_supported(const float* A, int lda, const float* B, int ldb, const float* C, int ldc, int M,
                                        int N, int K) {
  if (!A || !B || !C || M < 1 || N < 64 || K <= 0) return 0;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15) || (lda & 3) || (ldb & 3) || (ldc & 3)) return 0;
  return 1;
}

extern "C" int msx_gemm_tc_x3(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc,
                              int M, int N, int K, const float* bias, int relu, float drop_p, unsigned long long seed,
                              unsigned site, const void* aux, int ldaux, int aux_kind, float aux_scale, int accumulate,
                              int splitk, float* out_colsum, uint32_t* mask_out, int ldmask, void* stream) {
  MSX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "msx_gemm_tc_x3: negative dimension");
  MSX_REQUIRE(!(out_colsum && (accumulate || splitk > 1)), "msx_gemm_tc_x3: out_colsum needs a plain (non-accumulating) store");
  if (M == 0 || N == 0) return MSX_OK;
  MSX_REQUIRE(A && B && C, "msx_gemm_tc_x3: null operand");
  MSX_REQUIRE(K > 0, "msx_gemm_tc_x3: K must be > 0");
  if (!msx_gemm_tc_x3_supported(A, lda, B, ldb, C, ldc, M, N, K)) {
    msx_set_error("msx_gemm_tc_x3: needs N >= 64, 16-byte aligned operands and leading dimensions %% 4 == 0");
    return MSX_ERR_UNSUPPORTED;
  }
  MSX_REQUIRE(aux_kind == 0 || aux_kind == 2, "msx_gemm_tc_x3: aux_kind must be 0 (fp32 matrix) or 2 (bit mask)");
  MSX_REQUIRE(!(transA == 1 && transB == 1), "msx_gemm_tc_x3: A^T B^T is not used on this path");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_gemm_tc_x3: dropout probability must be in [0,1)");
  MSX_REQUIRE(!(splitk > 1 && (bias || relu || drop_p > 0.f || aux || accumulate || mask_out)),
              "msx_gemm_tc_x3: split-K only supports the plain atomic-add epilogue");
  MSX_REQUIRE(!(mask_out && ((N & 31) || accumulate || ldmask < N / 32)), "msx_gemm_tc_x3: mask_out needs N %% 32 == 0, a plain store and ldmask >= N / 32");
  MSX_REQUIRE(!(aux && aux_kind == 2 && ((N & 31) || ldaux < N / 32)), "msx_gemm_tc_x3: a bit-mask aux needs N %% 32 == 0 and ldaux >= N / 32 words");
  using Op = OpCfg<false>;
  if (splitk < 1) splitk = 1;
  const bool a_mn = transA == 1, b_mn = transB == 0;
  if (a_mn && !b_mn) {
    msx_set_error("msx_gemm_tc_x3: operand major combination (A MN-major, B K-major) is not instantiated");
    return MSX_ERR_UNSUPPORTED;
  }
  // tile width: the one that pads N the least (N = 293: three 128-wide tiles, not two 256-wide ones)
  const int pad256 = msx_ceil_div(N, 256) * 256, pad128 = msx_ceil_div(N, 128) * 128;
  static const int force_bn = [] { const char* e = getenv("MSX_X3_BN"); return e ? atoi(e) : 0; }();
  const int bn2 = force_bn == 128 ? 128 : (N > 128 && pad256 <= pad128) ? 256 : 128;
  CUtensorMap ta, tb, tc;
  int rc = make_map(&tc, C, M, N, ldc, 32, 32, false, kMapC32);
  if (rc) return rc;
  if (!a_mn) rc = make_map(&ta, A, M, K, lda, Op::kBKE, BM, false, kMapF32Op); else rc = make_map(&ta, A, K, M, lda, Op::kSlabMN, Op::kBKE, true, kMapF32Op);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&tb, B, N, K, ldb, Op::kBKE, bn2 / 2, false, kMapF32Op); else rc = make_map(&tb, B, K, N, ldb, Op::kSlabMN, Op::kBKE, true, kMapF32Op);
  if (rc) return rc;
  TcParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.relu = relu; p.drop_p = drop_p;
  p.inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.seed = seed; p.seed_ctr = msx_step_counter(); p.site = site; p.aux = (const float*)aux; p.ldaux = ldaux;
  p.aux_scale = aux_scale; p.accumulate = accumulate; p.out_colsum = out_colsum; p.c_bf16 = 0; p.aux_bf16 = aux_kind;
  p.mask_out = mask_out; p.ldmask = ldmask;
  static const int dbg = [] { const char* e = getenv("MSX_X3_DEBUG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  p.kb_total = msx_ceil_div(K, Op::kBKE);
  p.m_tiles = msx_ceil_div(M, 2 * BM); p.n_tiles = msx_ceil_div(N, bn2);
  if (splitk > 1) {                         // about two waves of pairs, as in msx_gemm_tc
    const int tiles = p.m_tiles * p.n_tiles, pairs = msx_num_sms() / 2;
    splitk = (2 * pairs) / tiles;
    if (splitk < 2) splitk = 2;
  }
  if (splitk > p.kb_total) splitk = p.kb_total;
  p.kb_per_split = msx_ceil_div(p.kb_total, splitk);
  p.splitk = msx_ceil_div(p.kb_total, p.kb_per_split);
  if (p.splitk == 1 && splitk > 1) { p.splitk = 2; p.kb_per_split = p.kb_total; }
  cudaStream_t st = (cudaStream_t)stream;
  if (bn2 == 256) {
    if (!a_mn && !b_mn) return launch_x3<256, false, false>(ta, tb, tc, p, st);
    if (!a_mn && b_mn) return launch_x3<256, false, true>(ta, tb, tc, p, st);
    return launch_x3<256, true, true>(ta, tb, tc, p, st);
  }
  if (!a_mn && !b_mn) return launch_x3<128, false, false>(ta, tb, tc, p, st);
  if (!a_mn && b_mn) return launch_x3<128, false, true>(ta, tb, tc, p, st);
  return launch_x3<128, true, true>(ta, tb, tc, p, st);
}
