// Shared pieces of the tcgen05 / TMA GEMM kernels (gemm_tc.cu: single-pass TF32 / BF16; gemm_tc_x3.cu: 3xTF32 split):
// tile constants, launch parameters, mbarrier / TMA / tcgen05 wrappers, the epilogue of one 32 x 32 chunk, the
// cta_group::2 helpers and the host-side tensor-map encoder.  Everything lives in namespace msx_tc (internal linkage
// through `static` / templates), both translation units say `using namespace msx_tc`.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "msx_common.cuh"

namespace msx_tc {

constexpr int BM = 128, BN = 128, BK = 32;          // BK fp32 = 128 B = one swizzle row
constexpr int kStages = 4;
constexpr int kTileBytes = BM * BK * 4;             // 16 KB per operand per stage
constexpr int kStageBytes = 2 * kTileBytes;
// Epilogue warps: EW / 4 warps per TMEM lane group (a warp may only touch lanes 32 * (warp % 4) ..), each draining
// 4 / EW of the tile's columns.  EW = 8 (two staging boxes per warp) is the default; the wide-N, short-K GEMMs of the
// step (QKV, FF1 forward, FF2 dgrad: K <= 256, N >= 512) are paced by the epilogue's instruction issue (ncu: 2 epilogue
// warps per scheduler reach ~50 % issue utilisation) and run with EW = 16 (one box per warp): measured 202 -> 162 us
// (FF1 forward, bf16), 155 -> 137 us (FF2 dgrad), 112 -> 103 us (QKV); the mainloop-paced shapes are 1-3 % slower with
// 16 and keep 8.
template <int EW>
struct EpiCfg {
  static constexpr int kColSplit = EW / 4;
  static constexpr int kBoxes = EW == 8 ? 2 : 1;     // staging boxes per epilogue warp
  static constexpr int kThreads = 32 * (2 + EW);     // warp 0 TMA, warp 1 MMA, then the epilogue warps
};
constexpr int kOutBoxBytes = 32 * 128;               // epilogue staging box: 32 rows x 32 fp32
constexpr int kTmemCols = 256;                      // 2 accumulators x 128 fp32 columns

struct TcParams {
  float* C;
  int ldc, M, N, K;
  const float* bias;
  int relu;
  float drop_p, inv_keep;
  unsigned long long seed;
  const unsigned long long* seed_ctr;   // optional device-side step counter added to seed
  unsigned site;
  const float* aux;
  int ldaux;
  float aux_scale;
  int accumulate, splitk;
  float* out_colsum;   // out_colsum[n] += sum_m C[m,n] of the values this launch writes (bias gradient of the producer)
  int m_tiles, n_tiles, kb_total, kb_per_split;
  int c_bf16;          // C is bf16 [M, ldc] (plain store only); the staging box is 32 rows x 64 B, SWIZZLE_64B
  int aux_bf16;        // aux storage: 0 fp32 [M, ldaux], 1 bf16 [M, ldaux], 2 bit mask uint32 [M, ldaux words] (bit j of word
                       // [m, n / 32] <=> element [m, n] > 0, n = 32 * (n / 32) + j), as written through mask_out
  int dbg;             // experiments only (MSX_X3_DEBUG): 1 = converter skips its work, 2 = one MMA per k-step instead of three
  unsigned* mask_out;  // optional: bit mask of (C > 0) after the epilogue, [M, ldmask words]; needs N % 32 == 0
  int ldmask;
  // "p3" products (bf16 hi / lo PLANES of both operands, split by the kernels that produce them): the k-block sequence
  // walks the reduction three times, hi*hi | hi*lo | lo*hi, as if the operands were [A_hi A_hi A_lo] and [B_hi B_lo B_hi]
  // (K-major operands, kind::f16: ~2^-17 per product at 1.5 single-pass TF32 tensor time, no conversion stage).
  int x3_kb = 0;       // k-blocks of ONE walk (kb_total = 3 * x3_kb); 0 = plain GEMM
  int c_planes = 0;    // C leaves as bf16 hi / lo planes (tmC / P3Maps::c_lo), c_bf16 must be set
};

// extra tensor maps of a p3 launch (copies of the plain maps otherwise)
struct P3Maps {
  CUtensorMap a_lo, b_lo, c_lo;
};

struct __align__(8) Barriers {
  unsigned long long full[kStages], empty[kStages], tmem_full[2], tmem_empty[2];
  unsigned tmem_base;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// One lane of a converged warp (deterministic for a given member mask).  tcgen05 / TMA instructions take
// uniform-register operands: issued from `if (lane == 0)` code the compiler wraps each one in an ELECT + R2UR
// waterfall loop (~90 cycles per instruction), issued under elect.sync from warp-uniform values it does not.
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), version 1.
// layout 2 = SWIZZLE_128B (16-byte units, 8-row atoms): K-major operands.
// layout 1 = SWIZZLE_128B_BASE32B (32-byte units, 4-row atoms): the only layout tcgen05 accepts for
//            MN-major 32-bit (TF32) operands; TMA produces it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes,
                                                        unsigned long long layout) {
  return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (layout << 61);
}

// BF = false: kind::tf32 (fp32 in memory, K = 8 per instruction); BF = true: kind::f16 with bf16 operands (K = 16).
// Either way one instruction consumes 32 bytes of the reduction dimension per operand row.
template <bool BF>
__device__ __forceinline__ void umma_ss(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                        unsigned idesc, unsigned accumulate) {
  if (BF)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Operand geometry of one pipeline stage.  A stage row is always 128 bytes of the contiguous dimension:
//   TF32: 32 elements;  MN-major slabs are 32 mn x 32 k-rows (SWIZZLE_128B_BASE32B, 4-row atoms, SBO 512 B)
//   BF16: 64 elements;  MN-major slabs are 64 mn x 64 k-rows (SWIZZLE_128B, 8-row atoms, SBO 1024 B)
template <bool BF>
struct OpCfg {
  static constexpr int kBKE = BF ? 64 : 32;                 // elements of K per stage
  static constexpr int kSlabMN = BF ? 64 : 32;              // mn elements per MN-major slab
  static constexpr int kSlabBytes = kBKE * 128;             // k-rows per stage x 128 B
  static constexpr unsigned kMnStep = BF ? 128 : 64;        // descriptor units (16 B) per MMA: 16 / 8 k-rows x 128 B
  static constexpr unsigned kMnSbo = BF ? 1024 : 512;
  static constexpr unsigned long long kMnLayout = BF ? 2 : 1;
  static constexpr unsigned kFmt = BF ? 1u : 2u;            // instruction descriptor operand format: BF16 / TF32
};
__device__ __forceinline__ float bf16_bits_to_float(unsigned short b) { return __uint_as_float((unsigned)b << 16); }
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float v[32]) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// aux (ReLU-mask source) of one 32-column chunk of this lane's row, fetched one chunk AHEAD of its use: the loads do not
// depend on the accumulator, so they are issued before the warp waits for the tile / while it works on the previous
// chunk, which takes the DRAM round trip out of the per-chunk dependency chain (the dgrad-with-mask GEMM was paced by it).
struct AuxPref { uint4 r[8]; };     // bf16 aux: r[0..3] (64 B); fp32 aux: r[0..7] (128 B)
__device__ __forceinline__ bool aux_fast(const TcParams& p, int col0) {       // warp-uniform
  if (p.aux_bf16 == 2) return p.aux != nullptr;                               // bit mask: one word per lane and chunk
  return p.aux && col0 + 32 <= p.N && (p.aux_bf16 ? (p.ldaux & 7) == 0 : (p.ldaux & 3) == 0) && ((uintptr_t)p.aux & 15) == 0;
}
__device__ __forceinline__ void aux_prefetch(const TcParams& p, int my_row, int col0, AuxPref& a) {
  if (!aux_fast(p, col0) || my_row >= p.M) return;
  if (p.aux_bf16 == 2) {
    a.r[0].x = col0 < p.N ? __ldg(reinterpret_cast<const unsigned*>(p.aux) + (size_t)my_row * p.ldaux + (col0 >> 5)) : 0u;
  } else if (p.aux_bf16) {
    const uint4* ax = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(p.aux) + (size_t)my_row * p.ldaux + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) a.r[j] = __ldg(ax + j);
  } else {
    const uint4* ax = reinterpret_cast<const uint4*>(p.aux + (size_t)my_row * p.ldaux + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.r[j] = __ldg(ax + j);
  }
}
// bit j set <=> aux[row, col0 + j] > 0 (bf16: 0 < bits < 0x8000 tested on the raw halves; fp32: sign clear and non-zero)
__device__ __forceinline__ unsigned aux_mask(const TcParams& p, const AuxPref& a) {
  unsigned m = 0u;
  if (p.aux_bf16 == 2) {
    m = a.r[0].x;
  } else if (p.aux_bf16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned w[4] = {a.r[j].x, a.r[j].y, a.r[j].z, a.r[j].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        m |= ((int)(w[q] << 16) > 0 ? 1u : 0u) << (8 * j + 2 * q);          // low half in [0x0001, 0x7FFF]
        m |= ((int)w[q] > 0xFFFF ? 1u : 0u) << (8 * j + 2 * q + 1);         // high half in [0x0001, 0x7FFF]
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m |= (__uint_as_float(a.r[j].x) > 0.f ? 1u : 0u) << (4 * j);
      m |= (__uint_as_float(a.r[j].y) > 0.f ? 1u : 0u) << (4 * j + 1);
      m |= (__uint_as_float(a.r[j].z) > 0.f ? 1u : 0u) << (4 * j + 2);
      m |= (__uint_as_float(a.r[j].w) > 0.f ? 1u : 0u) << (4 * j + 3);
    }
  }
  return m;
}

// Epilogue for one 32-row x 32-column chunk held in the row-owner layout (lane = row, v[j] = column col0 + j):
// bias / ReLU / dropout / aux mask / bias-gradient column sums, then a SWIZZLE_128B staging box that the TMA
// engine stores (or reduce-adds) into C.
template <int EW, int kBoxesT = EpiCfg<EW>::kBoxes>
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, const CUtensorMap* tmc, float (&v)[32], int row0,
                                               int my_row, int col0, int lane, unsigned char* st, int& sbuf,
                                               int& pending, bool reduce, unsigned amask,
                                               const CUtensorMap* tmc_lo = nullptr) {
    // ---- row-owner layout: this lane holds 32 consecutive columns of row my_row
    if (p.bias) {
      if (col0 + 32 <= p.N && (((uintptr_t)(p.bias + col0)) & 15) == 0) {   // warp-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
          v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += (col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.f;
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.drop_p > 0.f) {
      const unsigned long long eff_seed = msx_eff_seed(p.seed, p.seed_ctr);
      dropout_apply32(eff_seed, p.site, (unsigned long long)my_row * p.N + col0, p.drop_p, p.inv_keep, v);
    }
    if (p.mask_out && my_row < p.M) {         // N % 32 == 0 (checked on the host): every chunk is a whole mask word
      unsigned m = 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) m |= v[j] > 0.f ? (1u << j) : 0u;
      p.mask_out[(size_t)my_row * p.ldmask + (col0 >> 5)] = m;
    }
    if (p.aux && my_row < p.M) {
      if (aux_fast(p, col0)) {                // mask prefetched by the caller (aux_prefetch / aux_mask)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (amask & (1u << j)) ? v[j] * p.aux_scale : 0.f;
      } else if (p.aux_bf16) {
        const unsigned short* ax = reinterpret_cast<const unsigned short*>(p.aux) + (size_t)my_row * p.ldaux + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] *= (col0 + j < p.N && bf16_bits_to_float(__ldg(ax + j)) > 0.f) ? p.aux_scale : 0.f;
      } else {
        const float* ax = p.aux + (size_t)my_row * p.ldaux + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= (col0 + j < p.N && __ldg(ax + j) > 0.f) ? p.aux_scale : 0.f;
      }
    }
    if (p.out_colsum) {
      float t[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = my_row < p.M ? v[j] : 0.f;
      const float cs = warp_colsum32(t, lane);
      if (col0 + lane < p.N) atomicAdd(p.out_colsum + col0 + lane, cs);
    }
    // ---- stage as a SWIZZLE_128B box (row = lane, 8 x 16 B chunks XOR-ed with row % 8) and let TMA write it
    constexpr int kBoxes = kBoxesT;
    unsigned char* box = st + sbuf * kOutBoxBytes;
    if (pending >= kBoxes) {                // the box we are about to overwrite must have been read
      if (elect_one()) {
        if (kBoxes == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    }
    if (p.c_planes) {
      // hi box in the first, lo box in the second half of the staging box (each 32 rows x 64 B, SWIZZLE_64B)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint2 h0, l0, h1, l1;
        split4_bf16(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3], h0, l0);
        split4_bf16(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7], h1, l1);
        const int off = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(box + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(box + 2048 + off) = make_uint4(l0.x, l0.y, l1.x, l1.y);
      }
    } else if (p.c_bf16) {
      // 32 rows x 64 B, SWIZZLE_64B: 16-byte chunk index XOR-ed with (row / 2) % 4
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(box + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
            make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (elect_one()) {
      if (reduce)
        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmc),
                     "r"(smem_u32(box)), "r"(col0), "r"(row0)
                     : "memory");
      else
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmc),
                     "r"(smem_u32(box)), "r"(col0), "r"(row0)
                     : "memory");
      if (p.c_planes)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmc_lo),
                     "r"(smem_u32(box) + 2048), "r"(col0), "r"(row0)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (kBoxes == 2) sbuf ^= 1;
    if (pending < kBoxes) ++pending;
}

template <int BN2>
struct PairCfg {
  static constexpr int kBRows = BN2 / 2;
  static constexpr int kATile = BM * BK * 4;
  static constexpr int kBTile = kBRows * BK * 4;
  static constexpr int kStage = kATile + kBTile;
  static constexpr int kStages2 = BN2 == 256 ? 5 : 6;
  static constexpr int kTmem = 2 * BN2;
  static constexpr int kChunks = BN2 / 32;            // 32-column epilogue chunks per tile
};
constexpr int kMaxStages2 = 6;

struct __align__(8) Barriers2 {
  unsigned long long full[kMaxStages2], empty[kMaxStages2], tmem_full[2], tmem_empty[2];
  unsigned tmem_base;
};

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned mapa_shared(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, unsigned parity) {   // acquire at cluster scope
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITC_%=:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONEC_%=;\n"
      "bra WAITC_%=;\n"
      "DONEC_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(unsigned cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are counted on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, unsigned bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <bool BF>
__device__ __forceinline__ void umma_ss_pair(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                             unsigned idesc, unsigned accumulate) {
  if (BF)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((unsigned short)3)
      : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D row-major matrix [rows, cols] with leading dimension ld (elements); box = {box_cols (contiguous), box_rows}.
// kind: kMapC32 plain fp32 (the C map), kMapTf32 fp32 in memory read as TFLOAT32 (TMA rounds fp32 -> tf32 to nearest
// while loading; the MMA itself would truncate the low 13 mantissa bits), kMapBf16 bf16 operand, kMapC16 bf16 C map
// (32 x 32 box = 64-byte rows, SWIZZLE_64B).
// kMapF32Op: fp32 operand loaded WITHOUT rounding (3xTF32: the kernel splits the raw value into hi + lo itself).
enum MapKind { kMapC32 = 0, kMapTf32 = 1, kMapBf16 = 2, kMapC16 = 3, kMapF32Op = 4 };
static int make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
             int box_rows, bool mn_major, MapKind kind) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { msx_set_error("msx_gemm_tc: cuTensorMapEncodeTiled is not available from the driver"); return MSX_ERR_CUDA; }
  const bool b16 = kind == kMapBf16 || kind == kMapC16;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (b16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = kind == kMapTf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                 : b16            ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                  : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = kind == kMapC16                ? CU_TENSOR_MAP_SWIZZLE_64B
                                : (mn_major && (kind == kMapTf32 || kind == kMapF32Op)) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                                 : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { msx_set_error("msx_gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return MSX_ERR_CUDA; }
  return MSX_OK;
}

}  // namespace msx_tc
