// Shared pieces of the tcgen05 attention kernels (attention_tc.cu: T <= 128, one key tile;
// attention_tc_long.cu: 128 < T <= 384, key / query tiles of 128): PTX wrappers for mbarriers, TMA, tcgen05
// MMA / TMEM loads and stores, UMMA shared-memory descriptors, the MN-major TF32 swizzle, and the host-side
// TFLOAT32 tensor-map builder.  Everything sits in an anonymous namespace (one private copy per translation unit).
#pragma once
#include <cuda.h>

#include "msx_common.cuh"

namespace {

constexpr int DH = 32;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ unsigned long long make_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes,
                                                        unsigned long long layout) {
  return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float v[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// tcgen05.ld split into issue + wait so that several loads are in flight; the wait names the destination registers
// as read-write operands so that the compiler cannot move their consumers above it.
__device__ __forceinline__ void tmem_ld16_issue(unsigned taddr, float* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_wait(float* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                 "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float* v) {       // caller issues tcgen05.wait::st
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
      "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
// A operand from TMEM (lanes = M rows, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long bdesc, unsigned idesc,
                                             unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// one 32-element head row of an output tensor: fp32 (128 B, eight float4) or bf16 (64 B, four uint4); `elem` is the
// element offset of the row start, the same in both storage types
// (ncols = 16 writes the first half only: 16-wide heads)
__device__ __forceinline__ void store_row32(void* base, bool bf16, size_t elem, const float* o, int ncols = 32) {
  if (bf16) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(base) + elem);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (8 * j >= ncols) break;
      uint4 w;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(o[8 * j + 1]), "f"(o[8 * j + 0]));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(o[8 * j + 3]), "f"(o[8 * j + 2]));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.z) : "f"(o[8 * j + 5]), "f"(o[8 * j + 4]));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.w) : "f"(o[8 * j + 7]), "f"(o[8 * j + 6]));
      dst[j] = w;
    }
  } else {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (4 * j < ncols) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
}

// hi / lo bf16 planes of one row (operands of the p3 GEMMs): hi = rn_bf16(o) at hi_base, rn_bf16(o - hi) at lo_base
__device__ __forceinline__ void store_row32_planes(void* hi_base, void* lo_base, size_t elem, const float* o, int ncols = 32) {
  uint4* dh = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(hi_base) + elem);
  uint4* dl = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(lo_base) + elem);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (8 * j >= ncols) break;
    uint2 h0, l0, h1, l1;
    split4_bf16(o[8 * j], o[8 * j + 1], o[8 * j + 2], o[8 * j + 3], h0, l0);
    split4_bf16(o[8 * j + 4], o[8 * j + 5], o[8 * j + 6], o[8 * j + 7], h1, l1);
    dh[j] = make_uint4(h0.x, h0.y, h1.x, h1.y);
    dl[j] = make_uint4(l0.x, l0.y, l1.x, l1.y);
  }
}

// Coalesced form of store_row32 for a whole warp: lane l holds row l of a 32-row x 32-column tile; the tile goes through
// a 4 KB shared-memory staging area `wst` (private to the warp, swizzled so that both directions are conflict-free) and
// leaves with every quarter-warp writing one full 128-byte (bf16: eighth-warp, 64-byte) row segment.  With one row per
// lane a store instruction touches 32 lines at 16 bytes each; this way it touches 4 full lines.  elem0 = element offset
// of (row 0, column 0) of the tile, pitch = elements between rows, rows_valid = rows of the tile to write (<= 0: none).
__device__ __forceinline__ void store_tile32_coalesced(unsigned char* wst, void* base, bool bf16, size_t elem0, size_t pitch,
                                                       int rows_valid, const float* o, int lane) {
  if (bf16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 w;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(o[8 * j + 1]), "f"(o[8 * j + 0]));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(o[8 * j + 3]), "f"(o[8 * j + 2]));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.z) : "f"(o[8 * j + 5]), "f"(o[8 * j + 4]));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.w) : "f"(o[8 * j + 7]), "f"(o[8 * j + 6]));
      *reinterpret_cast<uint4*>(wst + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = w;
    }
    __syncwarp();
    unsigned short* dst = reinterpret_cast<unsigned short*>(base) + elem0;
    const int c = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 8 * i + (lane >> 2);
      const uint4 w = *reinterpret_cast<const uint4*>(wst + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
      if (r < rows_valid) *reinterpret_cast<uint4*>(dst + (size_t)r * pitch + 8 * c) = w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(wst + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    __syncwarp();
    float* dst = reinterpret_cast<float*>(base) + elem0;
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3);
      const float4 w = *reinterpret_cast<const float4*>(wst + r * 128 + ((c ^ (r & 7)) << 4));
      if (r < rows_valid) *reinterpret_cast<float4*>(dst + (size_t)r * pitch + 4 * c) = w;
    }
  }
  __syncwarp();
}

// MN-major operand, SWIZZLE_128B_BASE32B: element (mn, k) of a slab-structured tile with `krows` rows per slab
__device__ __forceinline__ unsigned mn_major_off(int mn, int k, int krows) {
  return (unsigned)((mn >> 5) * (krows * 128) + k * 128 + ((((mn & 31) >> 3) ^ (k & 3)) << 5) + ((mn & 7) << 2));
}


// -------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
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
// raw: the tile is loaded as FLOAT32 (bits unchanged) instead of TFLOAT32 (rounded to nearest TF32 by the TMA unit) —
// for kernels that split the value into hi + lo TF32 parts themselves.
int make_map(CUtensorMap* map, const float* ptr, long long rows, long long cols, long long ld, int box_cols,
             int box_rows, bool mn_major, bool raw = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { msx_set_error("msx_attention_tc: cuTensorMapEncodeTiled unavailable"); return MSX_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, raw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { msx_set_error("msx_attention_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return MSX_ERR_CUDA; }
  return MSX_OK;
}


}  // namespace
