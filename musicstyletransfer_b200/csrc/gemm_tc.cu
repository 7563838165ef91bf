// K2 (tensor-core path) — persistent, warp-specialised tcgen05 + TMA GEMM for sm_100a.
//
// Same contract as msx_gemm_f32 (gemm_simt.cu): replaces the gluon.nn.Dense forward / backward GEMMs of
// /root/reference/music_style_transfer/VarAutoEncoder/{transformer.py:36-40,65-68,88-93,104, model.py:70-71,
// 139-157,214-227}.  Two operand types, one kernel body (template BF):
//   TF32  operands stay fp32 in HBM and are consumed by `tcgen05.mma.kind::tf32` (TMA rounds them to nearest while
//         loading), so the tensor path drops into the fp32 step without conversion passes;
//   BF16  operands are bfloat16 in HBM (`kind::f16`); a stage row is 128 bytes either way, so the ring, the K-major
//         descriptors and the 4 MMAs per stage are byte-identical (OpCfg holds what differs for MN-major operands).
// Accumulation is fp32 in TMEM in both cases.
//
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B boxes) into a 4-6 stage smem ring
//   warp 1      TMEM allocator + MMA issuer (one elected lane): 4 tcgen05.mma per stage, tcgen05.commit releases the
//               smem stage / publishes the accumulator
//   warps 2..   epilogue, EW = 8 or 16 warps (EpiCfg: EW / 4 per TMEM lane group, each draining 4 / EW of the tile's
//               columns): tcgen05.ld (32 lanes x 32 columns) -> bias / ReLU / dropout / ReLU bit mask out / aux mask in /
//               bias-gradient column sums in the row-owner layout -> swizzled smem box -> TMA store (fp32 or bf16 C),
//               or TMA reduce-add (.add.f32) for accumulate and split-K, so C is never read by the SM and OOB rows /
//               columns are clipped by TMA
// Two TMEM accumulators are double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1; CTAs are
// persistent (one per SM) and walk output tiles N-fastest so the A row-block of a wave is shared through L2.
// gemm_tc_kernel: 128 x 128 tiles on one CTA; gemm_tc2_kernel: 256 x 256 / 256 x 128 tiles on a CTA pair
// (cta_group::2).  Operand majors: K-major (reduction dim contiguous: X in X W^T, W in X W^T, dY in dY W) and MN-major
// (output dim contiguous: W in dY W, dY^T and X in dY^T X) are both fed by TMA; only the shared-memory descriptor and
// the box geometry differ.
#include "gemm_tc_common.cuh"

using namespace msx_tc;

namespace {

template <bool A_MN, bool B_MN, bool BF, int EW>
__global__ void __launch_bounds__(EpiCfg<EW>::kThreads, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ P3Maps mx, const TcParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms repeat every 1024 B)
  unsigned char* ring = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kEpiWarps = EW, kColSplit = EpiCfg<EW>::kColSplit, kBoxes = EpiCfg<EW>::kBoxes;
  unsigned char* stage_out = ring + kStages * kStageBytes;                        // [kEpiWarps][kBoxes][32 rows][128 B]
  Barriers* bars = reinterpret_cast<Barriers*>(stage_out + kEpiWarps * kBoxes * kOutBoxBytes);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int items = p.m_tiles * p.n_tiles * p.splitk;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================================ TMA producer (whole warp waits, one elected lane issues) ================================
    int stage = 0;
    unsigned phase = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int nt = it % p.n_tiles, mt = (it / p.n_tiles) % p.m_tiles, ks = it / (p.n_tiles * p.m_tiles);
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        unsigned char* sa = ring + stage * kStageBytes;
        unsigned char* sb = sa + kTileBytes;
        if (elect_one()) {
          mbar_expect_tx(&bars->full[stage], kStageBytes);
          using Op = OpCfg<BF>;
          if (BF && !A_MN && !B_MN && p.x3_kb) {                                      // p3: hi*hi | hi*lo | lo*hi walks
            const int seg = kb / p.x3_kb, kk = (kb - seg * p.x3_kb) * Op::kBKE;
            tma_load_2d(sa, seg == 2 ? &mx.a_lo : &tmA, &bars->full[stage], kk, mt * BM);
            tma_load_2d(sb, seg == 1 ? &mx.b_lo : &tmB, &bars->full[stage], kk, nt * BN);
          } else {
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &bars->full[stage], kb * Op::kBKE, mt * BM);     // box {128 B of k, 128 rows}
          } else {
#pragma unroll
            for (int s = 0; s < BM / Op::kSlabMN; ++s)                              // slabs {128 B of m, kBKE k-rows}
              tma_load_2d(sa + s * Op::kSlabBytes, &tmA, &bars->full[stage], mt * BM + s * Op::kSlabMN, kb * Op::kBKE);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &bars->full[stage], kb * Op::kBKE, nt * BN);
          } else {
#pragma unroll
            for (int s = 0; s < BN / Op::kSlabMN; ++s)
              tma_load_2d(sb + s * Op::kSlabBytes, &tmB, &bars->full[stage], nt * BN + s * Op::kSlabMN, kb * Op::kBKE);
          }
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (whole warp waits, one elected lane issues) ================================
    using Op = OpCfg<BF>;
    // instruction descriptor: D=F32, A=B=TF32 / BF16, majors, N>>3 @17, M>>4 @24
    const unsigned idesc = (1u << 4) | (Op::kFmt << 7) | (Op::kFmt << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                           ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
    // descriptors of stage 0; K-major: +32 B per MMA inside the 128 B swizzle row, SBO = 8 rows * 128 B.
    // MN-major: LBO = slab stride (all k-rows of the stage * 128 B), SBO / step per MMA: see OpCfg.
    const unsigned long long ad0 = A_MN ? make_desc(smem_u32(ring), Op::kSlabBytes, Op::kMnSbo, Op::kMnLayout)
                                        : make_desc(smem_u32(ring), 16, 1024, 2);
    const unsigned long long bd0 = B_MN ? make_desc(smem_u32(ring) + kTileBytes, Op::kSlabBytes, Op::kMnSbo, Op::kMnLayout)
                                        : make_desc(smem_u32(ring) + kTileBytes, 16, 1024, 2);
    constexpr unsigned kAStep = A_MN ? Op::kMnStep : 2, kBStep = B_MN ? Op::kMnStep : 2;      // descriptor address units of 16 B
    int stage = 0;
    unsigned phase = 0;
    int local = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++local) {
      const int ks = it / (p.n_tiles * p.m_tiles);
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      const int buf = local & 1;
      const unsigned use = (unsigned)(local >> 1);
      mbar_wait(&bars->tmem_empty[buf], (use & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned tmem_d = tmem_base + buf * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bars->full[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned long long ad = ad0 + (unsigned long long)(stage * (kStageBytes >> 4));
        const unsigned long long bd = bd0 + (unsigned long long)(stage * (kStageBytes >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) umma_ss<BF>(tmem_d, ad + kAStep * k, bd + kBStep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&bars->empty[stage]);                 // frees the smem stage when these MMAs retire
          if (kb == kb1 - 1) umma_commit(&bars->tmem_full[buf]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue (warps 2..5) ================================
    const int ew = warp - 2;                 // staging slot
    const int lg = warp & 3;                 // TMEM lane group this warp may access (warps w and w+4 share one)
    const int chalf = ew >> 2;               // which 1 / kColSplit of the BN columns this warp drains
    unsigned char* st = stage_out + ew * kBoxes * kOutBoxBytes;
    const bool reduce = p.accumulate || p.splitk > 1;
    int local = 0, sbuf = 0, pending = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++local) {
      const int nt = it % p.n_tiles, mt = (it / p.n_tiles) % p.m_tiles;
      const int buf = local & 1;
      const unsigned use = (unsigned)(local >> 1);
      const int row0 = mt * BM + lg * 32;
      const int my_row = row0 + lane;
      constexpr int kChPerWarp = BN / 32 / kColSplit;
      AuxPref apre;
      if (p.aux) aux_prefetch(p, my_row, nt * BN + chalf * kChPerWarp * 32, apre);      // before waiting for the accumulator
      mbar_wait(&bars->tmem_full[buf], use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int ch = chalf * kChPerWarp; ch < (chalf + 1) * kChPerWarp; ++ch) {
        const int col0 = nt * BN + ch * 32;
        unsigned amask = 0u;
        if (p.aux) {
          amask = aux_mask(p, apre);
          if (ch + 1 < (chalf + 1) * kChPerWarp) aux_prefetch(p, my_row, col0 + 32, apre);
        }
        float v[32];
        tmem_ld32(tmem_base + ((unsigned)(lg * 32) << 16) + buf * BN + ch * 32, v);
        if (col0 < p.N && row0 < p.M) {          // warp-uniform
          epilogue_chunk<EW>(p, &tmC, v, row0, my_row, col0, lane, st, sbuf, pending, reduce, amask, &mx.c_lo);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x BN2 output tile.  Each CTA stages
// 128 rows of A and BN2/2 rows of B per k-block, the leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (UMMA M = 256) which reads both CTAs' shared memory, and each CTA's TMEM receives
// its own 128 x BN2 half of the accumulator.  Per output element the operand traffic L2 -> SM is half of the
// 1-CTA 128 x 128 kernel's, which is what bounds these K <= 1024 fp32-I/O GEMMs (the chip-wide L2 -> SM cap
// is ~2x HBM bandwidth and the 1-CTA kernel re-reads a 128 x K weight panel for every 128 x 128 tile).
//   full[s]        leader only; count 1 (leader's arrive.expect_tx for BOTH CTAs' bytes); the peer's TMA
//                  completes its bytes on the leader's barrier (cp.async.bulk.tensor .cta_group::2)
//   empty[s]       both CTAs; tcgen05.commit .multicast::cluster arrives on both
//   tmem_full[b]   both CTAs; multicast commit
//   tmem_empty[b]  leader only; count 2 x epilogue warps, the peer's warps arrive remotely
// ================================================================================================
template <int BN2, bool A_MN, bool B_MN, bool BF, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EpiCfg<EW>::kThreads, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ P3Maps mx, const TcParams p) {
  pdl_entry();
  using Cfg = PairCfg<BN2>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kEpiWarps = EW, kColSplit = EpiCfg<EW>::kColSplit, kBoxes = EpiCfg<EW>::kBoxes;
  unsigned char* stage_out = ring + Cfg::kStages2 * Cfg::kStage;                  // [kEpiWarps][kBoxes][32 rows][128 B]
  Barriers2* bars = reinterpret_cast<Barriers2*>(stage_out + kEpiWarps * kBoxes * kOutBoxBytes);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const unsigned rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int items = p.m_tiles * p.n_tiles * p.splitk;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages2; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], 2 * kEpiWarps); }
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
  cluster_sync_all();                       // barriers of both CTAs initialised, TMEM allocated in both
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs; whole warp waits, one elected lane issues) ================================
    int stage = 0;
    unsigned phase = 0;
    for (int it = pair; it < items; it += npairs) {
      const int nt = it % p.n_tiles, mt = (it / p.n_tiles) % p.m_tiles, ks = it / (p.n_tiles * p.m_tiles);
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      const int m0 = mt * (2 * BM) + (int)rank * BM;               // this CTA's 128 rows of A
      const int n0 = nt * BN2 + (int)rank * Cfg::kBRows;           // this CTA's BN2/2 rows of B
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        unsigned char* sa = ring + stage * Cfg::kStage;
        unsigned char* sb = sa + Cfg::kATile;
        const unsigned fb = mapa_shared(smem_u32(&bars->full[stage]), 0);
        if (elect_one()) {
          using Op = OpCfg<BF>;
          if (leader) mbar_expect_tx(&bars->full[stage], 2 * Cfg::kStage);
          if (BF && !A_MN && !B_MN && p.x3_kb) {                                      // p3: hi*hi | hi*lo | lo*hi walks
            const int seg = kb / p.x3_kb, kk = (kb - seg * p.x3_kb) * Op::kBKE;
            tma_load_2d_pair(sa, seg == 2 ? &mx.a_lo : &tmA, fb, kk, m0);
            tma_load_2d_pair(sb, seg == 1 ? &mx.b_lo : &tmB, fb, kk, n0);
          } else {
          if (!A_MN) {
            tma_load_2d_pair(sa, &tmA, fb, kb * Op::kBKE, m0);                     // box {128 B of k, 128 rows}
          } else {
#pragma unroll
            for (int s = 0; s < BM / Op::kSlabMN; ++s)                              // slabs {128 B of m, kBKE k-rows}
              tma_load_2d_pair(sa + s * Op::kSlabBytes, &tmA, fb, m0 + s * Op::kSlabMN, kb * Op::kBKE);
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, &tmB, fb, kb * Op::kBKE, n0);                     // box {128 B of k, BN2/2 rows}
          } else {
#pragma unroll
            for (int s = 0; s < Cfg::kBRows / Op::kSlabMN; ++s)
              tma_load_2d_pair(sb + s * Op::kSlabBytes, &tmB, fb, n0 + s * Op::kSlabMN, kb * Op::kBKE);
          }
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only; whole warp waits, one elected lane issues) ================================
    if (leader) {
      using Op = OpCfg<BF>;
      // instruction descriptor: D=F32, A=B=TF32 / BF16, majors, N>>3 @17, M>>4 @24 with M = 256 across the pair
      const unsigned idesc = (1u << 4) | (Op::kFmt << 7) | (Op::kFmt << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((unsigned)(BN2 >> 3) << 17) | ((unsigned)((2 * BM) >> 4) << 24);
      const unsigned long long ad0 = A_MN ? make_desc(smem_u32(ring), Op::kSlabBytes, Op::kMnSbo, Op::kMnLayout)
                                          : make_desc(smem_u32(ring), 16, 1024, 2);
      const unsigned long long bd0 = B_MN ? make_desc(smem_u32(ring) + Cfg::kATile, Op::kSlabBytes, Op::kMnSbo, Op::kMnLayout)
                                          : make_desc(smem_u32(ring) + Cfg::kATile, 16, 1024, 2);
      constexpr unsigned kAStep = A_MN ? Op::kMnStep : 2, kBStep = B_MN ? Op::kMnStep : 2;    // descriptor address units of 16 B
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
          mbar_wait(&bars->full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned long long ad = ad0 + (unsigned long long)(stage * (Cfg::kStage >> 4));
          const unsigned long long bd = bd0 + (unsigned long long)(stage * (Cfg::kStage >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k)
              umma_ss_pair<BF>(tmem_d, ad + kAStep * k, bd + kBStep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&bars->empty[stage]);            // frees this stage in BOTH CTAs
            if (kb == kb1 - 1) umma_commit_pair(&bars->tmem_full[buf]);   // accumulator halves complete in both CTAs
          }
          __syncwarp();
          if (++stage == Cfg::kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================================ epilogue (warps 2..9, both CTAs) ================================
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int chalf = ew >> 2;
    unsigned char* st = stage_out + ew * kBoxes * kOutBoxBytes;
    const bool reduce = p.accumulate || p.splitk > 1;
    int local = 0, sbuf = 0, pending = 0;
    for (int it = pair; it < items; it += npairs, ++local) {
      const int nt = it % p.n_tiles, mt = (it / p.n_tiles) % p.m_tiles;
      const int buf = local & 1;
      const unsigned use = (unsigned)(local >> 1);
      const int row0 = mt * (2 * BM) + (int)rank * BM + lg * 32;
      const int my_row = row0 + lane;
      constexpr int kChPerWarp = Cfg::kChunks / kColSplit;
      AuxPref apre;
      if (p.aux) aux_prefetch(p, my_row, nt * BN2 + chalf * kChPerWarp * 32, apre);     // before waiting for the accumulator
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
        if (col0 < p.N && row0 < p.M) {          // warp-uniform
          epilogue_chunk<EW>(p, &tmC, v, row0, my_row, col0, lane, st, sbuf, pending, reduce, amask, &mx.c_lo);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&bars->tmem_empty[buf]), 0));
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                       // no CTA leaves (or frees TMEM) while its pair may still touch it
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmem) : "memory");
  }
}

// -------------------------------------------------------------------------------- host side
template <int EW>
constexpr size_t smem_bytes_1cta() {
  return 1024 + (size_t)kStages * kStageBytes + (size_t)EW * EpiCfg<EW>::kBoxes * kOutBoxBytes + sizeof(Barriers);
}

template <bool A_MN, bool B_MN, bool BF, int EW>
int launch_ew(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const P3Maps& mx, const TcParams& p, cudaStream_t st) {
  constexpr size_t smem = smem_bytes_1cta<EW>();
  MSX_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, BF, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = p.m_tiles * p.n_tiles * p.splitk;
  const int grid = items < msx_num_sms() ? items : msx_num_sms();
  MSX_CUDA(msx_launch(gemm_tc_kernel<A_MN, B_MN, BF, EW>, dim3(grid), dim3(EpiCfg<EW>::kThreads), smem, st, ta, tb, tc, mx, p));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
template <bool A_MN, bool B_MN, bool BF>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const P3Maps& mx, const TcParams& p, cudaStream_t st) {
  return launch_ew<A_MN, B_MN, BF, 8>(ta, tb, tc, mx, p, st);        // K < 256 here: mainloop too short for pair tiles, 8 warps
}

template <int BN2, int EW>
constexpr size_t pair_smem_bytes() {
  return 1024 + (size_t)PairCfg<BN2>::kStages2 * PairCfg<BN2>::kStage + (size_t)EW * EpiCfg<EW>::kBoxes * kOutBoxBytes + sizeof(Barriers2);
}

template <int BN2, bool A_MN, bool B_MN, bool BF, int EW>
int launch_pair_ew(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const P3Maps& mx, const TcParams& p, cudaStream_t st) {
  constexpr size_t smem = pair_smem_bytes<BN2, EW>();
  static_assert(smem <= 232448, "pair kernel exceeds the 227 KB shared-memory limit");
  MSX_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN2, A_MN, B_MN, BF, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = p.m_tiles * p.n_tiles * p.splitk;
  const int max_pairs = msx_num_sms() / 2;
  const int pairs = items < max_pairs ? items : max_pairs;
  MSX_CUDA(msx_launch(gemm_tc2_kernel<BN2, A_MN, B_MN, BF, EW>, dim3(2 * pairs), dim3(EpiCfg<EW>::kThreads), smem, st, ta, tb, tc, mx, p));   // static cluster dims (2,1,1)
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
// epilogue-paced shapes (short K, wide N, plain stores) take 16 epilogue warps: see EpiCfg

template <int BN2, bool A_MN, bool B_MN, bool BF>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const P3Maps& mx, const TcParams& p, cudaStream_t st) {
  static const int forced = [] { const char* e = getenv("MSX_GEMM_EPI_WARPS"); return e ? atoi(e) : 0; }();
  const bool wide = forced == 16 || (forced != 8 && p.K <= 256 && p.N >= 512 && !p.accumulate && p.splitk == 1);
  if (wide) return launch_pair_ew<BN2, A_MN, B_MN, BF, 16>(ta, tb, tc, mx, p, st);
  return launch_pair_ew<BN2, A_MN, B_MN, BF, 8>(ta, tb, tc, mx, p, st);
}

// 2-CTA path switch: MSX_GEMM_PAIR=0 in the environment or msx_gemm_tc_set_pair(0) forces the 1-CTA kernel
// (A/B comparisons, bisecting); default on.
int g_pair_mode = -1;
bool pair_enabled() {
  if (g_pair_mode < 0) {
    const char* e = getenv("MSX_GEMM_PAIR");
    g_pair_mode = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pair_mode == 1;
}

// hi / lo planes of a p3 launch: A / B / C of gemm_tc_impl are the hi planes, these the lo planes (C_lo optional)
struct P3Args {
  const void* A_lo;
  const void* B_lo;
  void* C_lo;
};

template <bool BF>
int gemm_tc_impl(const void* A, int lda, int transA, const void* B, int ldb, int transB, void* C, int ldc, int c_bf16,
                 int M, int N, int K, const float* bias, int relu, float drop_p, unsigned long long seed, unsigned site,
                 const void* aux, int ldaux, int aux_bf16, float aux_scale, int accumulate, int splitk, float* out_colsum,
                 void* stream, unsigned* mask_out = nullptr, int ldmask = 0, const P3Args* p3 = nullptr) {
  using Op = OpCfg<BF>;
  constexpr MapKind kOp = BF ? kMapBf16 : kMapTf32;
  if (splitk < 1) splitk = 1;
  // operand majors: A is K-major when stored [M,K] (transA=0), MN-major when stored [K,M] (transA=1);
  //                 B is K-major when stored [N,K] (transB=1), MN-major when stored [K,N] (transB=0).
  const bool a_mn = transA == 1, b_mn = transB == 0;
  CUtensorMap ta, tb, tc;
  int rc = make_map(&tc, C, M, N, ldc, 32, 32, false, c_bf16 ? kMapC16 : kMapC32);   // epilogue box: 32 rows x 32 columns
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  TcParams p;
  p.C = (float*)C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.relu = relu; p.drop_p = drop_p;
  p.inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.seed = seed; p.seed_ctr = msx_step_counter(); p.site = site; p.aux = (const float*)aux; p.ldaux = ldaux;
  p.aux_scale = aux_scale; p.accumulate = accumulate; p.out_colsum = out_colsum; p.c_bf16 = c_bf16; p.aux_bf16 = aux_bf16;
  p.mask_out = mask_out; p.ldmask = ldmask; p.dbg = 0;
  p.kb_total = msx_ceil_div(K, Op::kBKE);
  P3Maps mx;
  if (p3) {                                   // three walks over the reduction: hi*hi | hi*lo | lo*hi
    p.x3_kb = p.kb_total;
    p.kb_total *= 3;
    p.c_planes = p3->C_lo ? 1 : 0;
  }
  mx.c_lo = tc;
  if (p3 && p3->C_lo && (rc = make_map(&mx.c_lo, p3->C_lo, M, N, ldc, 32, 32, false, kMapC16))) return rc;
  // pair tiles pay off once the mainloop is long enough to hide the 128 x 256 epilogue (measured on the step's
  // shapes: K = 128 forward GEMMs are faster on 128 x 128 tiles, K >= 256 ones 1.2-1.35x faster on pair tiles)
  // ... and once there are enough of them: with less than a wave of pair tiles (small batches: M = 2 080 rows at the
  // reference's batch of 32) the 128 x 128 tiles of the 1-CTA kernel put 2-4x as many SMs on the problem.  Per-SM work of a
  // tile: 128 x 128 x K (1-CTA) vs 128 x bn2 x K (pair); the pair kernel earns its 1.25x only on full waves.
  bool use_pair = pair_enabled() && M > BM && N >= 64 && K >= 256;
  if (use_pair && splitk <= 1) {
    const int bn2 = N > 128 ? 256 : 128;
    const long long t1 = (long long)msx_ceil_div(M, BM) * msx_ceil_div(N, BN), tp = (long long)msx_ceil_div(M, 2 * BM) * msx_ceil_div(N, bn2);
    const long long w1 = (t1 + msx_num_sms() - 1) / msx_num_sms(), wp = (tp + msx_num_sms() / 2 - 1) / (msx_num_sms() / 2);
    if (5 * w1 * BN < 4 * wp * bn2) use_pair = false;
  }
  if (use_pair) {
    // ---- 2-CTA path: 256 x 256 (N > 128) or 256 x 128 pair tiles
    const int bn2 = N > 128 ? 256 : 128;
    if (!a_mn) rc = make_map(&ta, A, M, K, lda, Op::kBKE, BM, false, kOp); else rc = make_map(&ta, A, K, M, lda, Op::kSlabMN, Op::kBKE, true, kOp);
    if (rc) return rc;
    if (!b_mn) rc = make_map(&tb, B, N, K, ldb, Op::kBKE, bn2 / 2, false, kOp); else rc = make_map(&tb, B, K, N, ldb, Op::kSlabMN, Op::kBKE, true, kOp);
    if (rc) return rc;
    mx.a_lo = ta; mx.b_lo = tb;
    if (p3 && ((rc = make_map(&mx.a_lo, p3->A_lo, M, K, lda, Op::kBKE, BM, false, kOp)) ||
               (rc = make_map(&mx.b_lo, p3->B_lo, N, K, ldb, Op::kBKE, bn2 / 2, false, kOp)))) return rc;
    p.m_tiles = msx_ceil_div(M, 2 * BM); p.n_tiles = msx_ceil_div(N, bn2);
    if (splitk > 1) {                         // re-derive the split for pair tiles: about two waves of pairs
      const int tiles = p.m_tiles * p.n_tiles, pairs = msx_num_sms() / 2;
      splitk = (2 * pairs) / tiles;
      if (splitk < 2) splitk = 2;
    }
    if (splitk > p.kb_total) splitk = p.kb_total;
    p.kb_per_split = msx_ceil_div(p.kb_total, splitk);
    p.splitk = msx_ceil_div(p.kb_total, p.kb_per_split);
    if (p.splitk == 1 && splitk > 1) { p.splitk = 2; p.kb_per_split = p.kb_total; }   // keep the atomic-add contract
    if (bn2 == 256) {
      if (!a_mn && !b_mn) return launch_pair<256, false, false, BF>(ta, tb, tc, mx, p, st);
      if (!a_mn && b_mn) return launch_pair<256, false, true, BF>(ta, tb, tc, mx, p, st);
      if (a_mn && b_mn) return launch_pair<256, true, true, BF>(ta, tb, tc, mx, p, st);
    } else {
      if (!a_mn && !b_mn) return launch_pair<128, false, false, BF>(ta, tb, tc, mx, p, st);
      if (!a_mn && b_mn) return launch_pair<128, false, true, BF>(ta, tb, tc, mx, p, st);
      if (a_mn && b_mn) return launch_pair<128, true, true, BF>(ta, tb, tc, mx, p, st);
    }
    msx_set_error("msx_gemm_tc: operand major combination (A MN-major, B K-major) is not instantiated");
    return MSX_ERR_UNSUPPORTED;
  }
  if (!a_mn) rc = make_map(&ta, A, M, K, lda, Op::kBKE, BM, false, kOp); else rc = make_map(&ta, A, K, M, lda, Op::kSlabMN, Op::kBKE, true, kOp);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&tb, B, N, K, ldb, Op::kBKE, BN, false, kOp); else rc = make_map(&tb, B, K, N, ldb, Op::kSlabMN, Op::kBKE, true, kOp);
  if (rc) return rc;
  mx.a_lo = ta; mx.b_lo = tb;
  if (p3 && ((rc = make_map(&mx.a_lo, p3->A_lo, M, K, lda, Op::kBKE, BM, false, kOp)) ||
             (rc = make_map(&mx.b_lo, p3->B_lo, N, K, ldb, Op::kBKE, BN, false, kOp)))) return rc;
  p.m_tiles = msx_ceil_div(M, BM); p.n_tiles = msx_ceil_div(N, BN);
  if (splitk > p.kb_total) splitk = p.kb_total;
  p.kb_per_split = msx_ceil_div(p.kb_total, splitk);
  p.splitk = msx_ceil_div(p.kb_total, p.kb_per_split);
  if (p.splitk == 1 && splitk > 1) { p.splitk = 2; p.kb_per_split = p.kb_total; }   // keep the atomic-add contract
  if (!a_mn && !b_mn) return launch<false, false, BF>(ta, tb, tc, mx, p, st);
  if (!a_mn && b_mn) return launch<false, true, BF>(ta, tb, tc, mx, p, st);
  if (a_mn && b_mn) return launch<true, true, BF>(ta, tb, tc, mx, p, st);
  msx_set_error("msx_gemm_tc: operand major combination (A MN-major, B K-major) is not instantiated");
  return MSX_ERR_UNSUPPORTED;
}

}  // namespace

// Returns 1 when msx_gemm_tc can take this problem (TMA needs 16-byte aligned bases and row pitches).
extern "C" int msx_gemm_tc_supported(const float* A, int lda, const float* B, int ldb, const float* C, int ldc, int M,
                                     int N, int K) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return 0;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15) || (lda & 3) || (ldb & 3) || (ldc & 3)) return 0;
  return 1;
}

// Same for msx_gemm_tc_bf16: A and B are bf16 (leading dimensions % 8), C is bf16 (ldc % 8) or fp32 (ldc % 4).
extern "C" int msx_gemm_tc_bf16_supported(const void* A, int lda, const void* B, int ldb, const void* C, int ldc,
                                          int c_bf16, int M, int N, int K) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return 0;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15) || (lda & 7) || (ldb & 7)) return 0;
  if (c_bf16 ? (ldc & 7) : (ldc & 3)) return 0;
  return 1;
}

extern "C" int msx_gemm_tc_set_pair(int enable) {
  const int prev = pair_enabled() ? 1 : 0;
  g_pair_mode = enable ? 1 : 0;
  return prev;
}

extern "C" int msx_gemm_tc(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc,
                           int M, int N, int K, const float* bias, int relu, float drop_p, unsigned long long seed,
                           unsigned site, const float* aux, int ldaux, float aux_scale, int accumulate, int splitk,
                           float* out_colsum, void* stream) {
  MSX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "msx_gemm_tc: negative dimension");
  MSX_REQUIRE(!(out_colsum && (accumulate || splitk > 1)), "msx_gemm_tc: out_colsum needs a plain (non-accumulating) store");
  if (M == 0 || N == 0) return MSX_OK;
  MSX_REQUIRE(A && B && C, "msx_gemm_tc: null operand");
  MSX_REQUIRE(K > 0, "msx_gemm_tc: K must be > 0");
  MSX_REQUIRE(msx_gemm_tc_supported(A, lda, B, ldb, C, ldc, M, N, K),
              "msx_gemm_tc: A, B and C must be 16-byte aligned with leading dimensions %% 4 == 0");
  MSX_REQUIRE(!(transA == 1 && transB == 1), "msx_gemm_tc: A^T B^T is not used on this path");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_gemm_tc: dropout probability must be in [0,1)");
  MSX_REQUIRE(!(splitk > 1 && (bias || relu || drop_p > 0.f || aux || accumulate)),
              "msx_gemm_tc: split-K only supports the plain atomic-add epilogue");
  return gemm_tc_impl<false>(A, lda, transA, B, ldb, transB, C, ldc, 0, M, N, K, bias, relu, drop_p, seed, site, aux, ldaux,
                             0, aux_scale, accumulate, splitk, out_colsum, stream);
}

// bf16 variant: A and B are bf16 in HBM and feed tcgen05.mma.kind::f16 (fp32 accumulation in TMEM); C is fp32, or bf16
// (c_bf16, plain store only: accumulation and split-K stay fp32 TMA reduce-adds); the epilogue arithmetic is fp32 and
// identical to msx_gemm_tc's.  aux (the ReLU mask source) may be fp32 or bf16.
extern "C" int msx_gemm_tc_bf16(const void* A, int lda, int transA, const void* B, int ldb, int transB, void* C, int ldc,
                                int c_bf16, int M, int N, int K, const float* bias, int relu, float drop_p,
                                unsigned long long seed, unsigned site, const void* aux, int ldaux, int aux_bf16,
                                float aux_scale, int accumulate, int splitk, float* out_colsum, void* stream) {
  MSX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "msx_gemm_tc_bf16: negative dimension");
  MSX_REQUIRE(!(out_colsum && (accumulate || splitk > 1)), "msx_gemm_tc_bf16: out_colsum needs a plain (non-accumulating) store");
  if (M == 0 || N == 0) return MSX_OK;
  MSX_REQUIRE(A && B && C, "msx_gemm_tc_bf16: null operand");
  MSX_REQUIRE(K > 0, "msx_gemm_tc_bf16: K must be > 0");
  MSX_REQUIRE(msx_gemm_tc_bf16_supported(A, lda, B, ldb, C, ldc, c_bf16, M, N, K),
              "msx_gemm_tc_bf16: operands must be 16-byte aligned, bf16 leading dimensions %% 8 == 0, fp32 ones %% 4 == 0");
  MSX_REQUIRE(!(transA == 1 && transB == 1), "msx_gemm_tc_bf16: A^T B^T is not used on this path");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_gemm_tc_bf16: dropout probability must be in [0,1)");
  MSX_REQUIRE(!(splitk > 1 && (bias || relu || drop_p > 0.f || aux || accumulate)),
              "msx_gemm_tc_bf16: split-K only supports the plain atomic-add epilogue");
  MSX_REQUIRE(!(c_bf16 && (accumulate || splitk > 1)), "msx_gemm_tc_bf16: a bf16 C takes plain stores only");
  return gemm_tc_impl<true>(A, lda, transA, B, ldb, transB, C, ldc, c_bf16 ? 1 : 0, M, N, K, bias, relu, drop_p, seed, site,
                            aux, ldaux, aux_bf16 ? 1 : 0, aux_scale, accumulate, splitk, out_colsum, stream);
}

// General entry point: the two above plus the ReLU bit mask.  ab_bf16 selects the operand type (0: fp32 memory read as
// TF32, 1: bfloat16), c_bf16 the C type; aux_kind: 0 fp32 matrix, 1 bfloat16 matrix, 2 bit mask (uint32 [M, ldaux words],
// bit j of word [m, n / 32] <=> element [m, 32 * (n / 32) + j] > 0).  mask_out (optional, N % 32 == 0) receives that bit
// mask of the values this launch writes (after bias / ReLU / dropout): the FF1 forward emits it and the FF2 dgrad reads
// 4 bytes per 32 elements instead of re-reading the hidden activation.
extern "C" int msx_gemm_tc_ex(const void* A, int lda, int transA, const void* B, int ldb, int transB, void* C, int ldc, int M,
                              int N, int K, int ab_bf16, int c_bf16, const float* bias, int relu, float drop_p,
                              unsigned long long seed, unsigned site, const void* aux, int ldaux, int aux_kind,
                              float aux_scale, int accumulate, int splitk, float* out_colsum, unsigned* mask_out, int ldmask,
                              void* stream) {
  MSX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "msx_gemm_tc_ex: negative dimension");
  MSX_REQUIRE(!(out_colsum && (accumulate || splitk > 1)), "msx_gemm_tc_ex: out_colsum needs a plain (non-accumulating) store");
  if (M == 0 || N == 0) return MSX_OK;
  MSX_REQUIRE(A && B && C, "msx_gemm_tc_ex: null operand");
  MSX_REQUIRE(K > 0, "msx_gemm_tc_ex: K must be > 0");
  MSX_REQUIRE(aux_kind >= 0 && aux_kind <= 2, "msx_gemm_tc_ex: aux_kind must be 0 (fp32), 1 (bf16) or 2 (bit mask)");
  if (ab_bf16)
    MSX_REQUIRE(msx_gemm_tc_bf16_supported(A, lda, B, ldb, C, ldc, c_bf16, M, N, K),
                "msx_gemm_tc_ex: operands must be 16-byte aligned, bf16 leading dimensions %% 8 == 0, fp32 ones %% 4 == 0");
  else
    MSX_REQUIRE(!c_bf16 && msx_gemm_tc_supported((const float*)A, lda, (const float*)B, ldb, (const float*)C, ldc, M, N, K),
                "msx_gemm_tc_ex: TF32 operands need fp32 A, B, C, 16-byte aligned, leading dimensions %% 4 == 0");
  MSX_REQUIRE(!(transA == 1 && transB == 1), "msx_gemm_tc_ex: A^T B^T is not used on this path");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_gemm_tc_ex: dropout probability must be in [0,1)");
  MSX_REQUIRE(!(splitk > 1 && (bias || relu || drop_p > 0.f || aux || accumulate || mask_out)),
              "msx_gemm_tc_ex: split-K only supports the plain atomic-add epilogue");
  MSX_REQUIRE(!(c_bf16 && (accumulate || splitk > 1)), "msx_gemm_tc_ex: a bf16 C takes plain stores only");
  MSX_REQUIRE(!(mask_out && ((N & 31) || accumulate || ldmask < N / 32)), "msx_gemm_tc_ex: mask_out needs N %% 32 == 0, a plain store and ldmask >= N / 32");
  MSX_REQUIRE(!(aux && aux_kind == 2 && ((N & 31) || ldaux < N / 32)), "msx_gemm_tc_ex: a bit-mask aux needs N %% 32 == 0 and ldaux >= N / 32 words");
  if (ab_bf16)
    return gemm_tc_impl<true>(A, lda, transA, B, ldb, transB, C, ldc, c_bf16 ? 1 : 0, M, N, K, bias, relu, drop_p, seed, site, aux,
                              ldaux, aux_kind, aux_scale, accumulate, splitk, out_colsum, stream, mask_out, ldmask);
  return gemm_tc_impl<false>(A, lda, transA, B, ldb, transB, C, ldc, 0, M, N, K, bias, relu, drop_p, seed, site, aux, ldaux,
                             aux_kind, aux_scale, accumulate, splitk, out_colsum, stream, mask_out, ldmask);
}

// "p3" forward GEMM: Y = X W^T (+ bias, ReLU, dropout, ReLU bit mask) with BOTH operands given as bfloat16 hi / lo planes
// (hi = rn_bf16(x), lo = rn_bf16(x - hi), written by the kernels that produce x: msx_embed_fwd_p, msx_add_ln_fwd_p,
// msx_attention_tc_fwd_p, this function's own C planes, msx_split_planes for the weights).  The kernel walks the
// reduction three times, A_hi B_hi + A_hi B_lo + A_lo B_hi, on kind::f16 with fp32 accumulation in TMEM: products to
// ~2^-17 relative, the precision class of 3xTF32 for this step, at 1.5 single-pass TF32 MMAs and with no conversion stage
// between TMA and the tensor pipe.  A [M, K] / B [N, K] K-major, K % 64 == 0.  c_kind: 0 = fp32 C; 2 = C as bf16 hi / lo
// planes C / C_lo (the next p3 GEMM's operand, e.g. the FF hidden activation, which then never exists in fp32).
extern "C" int msx_gemm_tc_p3_supported(const void* A_hi, int lda, const void* B_hi, int ldb, const void* C, int ldc, int c_kind,
                                        int M, int N, int K) {
  if (K % 64 != 0) return 0;
  return msx_gemm_tc_bf16_supported(A_hi, lda, B_hi, ldb, C, ldc, c_kind == 2 ? 1 : 0, M, N, K);
}

extern "C" int msx_gemm_tc_p3(const void* A_hi, const void* A_lo, int lda, const void* B_hi, const void* B_lo, int ldb, void* C,
                              void* C_lo, int ldc, int c_kind, int M, int N, int K, const float* bias, int relu, float drop_p,
                              unsigned long long seed, unsigned site, int accumulate, unsigned* mask_out, int ldmask,
                              void* stream) {
  MSX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "msx_gemm_tc_p3: negative dimension");
  if (M == 0 || N == 0) return MSX_OK;
  MSX_REQUIRE(A_hi && A_lo && B_hi && B_lo && C, "msx_gemm_tc_p3: null operand");
  MSX_REQUIRE(c_kind == 0 || c_kind == 2, "msx_gemm_tc_p3: c_kind must be 0 (fp32) or 2 (bf16 hi / lo planes)");
  MSX_REQUIRE((c_kind == 2) == (C_lo != nullptr), "msx_gemm_tc_p3: C_lo goes with c_kind == 2");
  MSX_REQUIRE(msx_gemm_tc_p3_supported(A_hi, lda, B_hi, ldb, C, ldc, c_kind, M, N, K) &&
                  (((uintptr_t)A_lo | (uintptr_t)B_lo | (uintptr_t)C_lo) & 15) == 0,
              "msx_gemm_tc_p3: planes must be 16-byte aligned, leading dimensions %% 8 == 0, K %% 64 == 0");
  MSX_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "msx_gemm_tc_p3: dropout probability must be in [0,1)");
  MSX_REQUIRE(!(c_kind == 2 && accumulate), "msx_gemm_tc_p3: plane outputs take plain stores only");
  MSX_REQUIRE(!(mask_out && ((N & 31) || accumulate || ldmask < N / 32)), "msx_gemm_tc_p3: mask_out needs N %% 32 == 0, a plain store and ldmask >= N / 32");
  const P3Args p3{A_lo, B_lo, C_lo};
  return gemm_tc_impl<true>(A_hi, lda, 0, B_hi, ldb, 1, C, ldc, c_kind == 2 ? 1 : 0, M, N, K, bias, relu, drop_p, seed, site,
                            nullptr, 0, 0, 1.f, accumulate, 1, nullptr, stream, mask_out, ldmask, &p3);
}

namespace {
__global__ void __launch_bounds__(256) split_planes_kernel(const float4* __restrict__ src, uint2* __restrict__ hi,
                                                           uint2* __restrict__ lo, long long n4) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(src + i);
    uint2 h, l;
    split4_bf16(v.x, v.y, v.z, v.w, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
}  // namespace

// fp32 -> bf16 hi / lo planes (n % 4 == 0, 16-byte aligned src, 8-byte aligned planes): the weight arena once per step,
// and the few small activations no kernel produces as planes
extern "C" int msx_split_planes(const float* src, void* hi, void* lo, long long n, void* stream) {
  MSX_REQUIRE(src && hi && lo, "msx_split_planes: null pointer");
  MSX_REQUIRE(n >= 0 && (n & 3) == 0 && ((uintptr_t)src & 15) == 0 && (((uintptr_t)hi | (uintptr_t)lo) & 7) == 0,
              "msx_split_planes: n %% 4 == 0, src 16-byte and planes 8-byte aligned");
  if (n == 0) return MSX_OK;
  const long long n4 = n / 4;
  const long long want = (n4 + 255) / 256, cap = (long long)msx_num_sms() * 8;
  MSX_CUDA(msx_launch(split_planes_kernel, dim3((int)(want < cap ? want : cap)), dim3(256), 0, (cudaStream_t)stream, 
      reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo), n4));
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
