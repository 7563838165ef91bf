// K1 — note-event rasteriser: MIDI note events -> event-token row + piano-roll window.
//
// Replaces (reference, /root/reference/music_style_transfer):
//   MIDIUtil/midi_io.py:70-93   EventBasedMIDIReader._parse_track   (tokens, incl. the modulo quirk)
//   MIDIUtil/Melody.py:109-126  create_note_on/off/timeshift_event  (id arithmetic)
//   MIDIUtil/midi_io.py:119-127 MelodyWriter._write_track           (clock semantics used for the roll)
// Spec of the roll: oracle/featurise.py:rasterize_sequence (derived).
//
// Design (HBM-write-bound, 8 KB of roll per 192 B of input): one WARP per sequence, persistent grid.
// The 32 lanes hold 32 consecutive events; clocks and token positions come from warp scans; the
// note that an event closes is found with match.any on the pitch (previous same-pitch lane) or in a
// per-warp 128-entry pitch table for notes opened in an earlier chunk.  Notes are painted as byte
// runs into a per-warp shared-memory tile [S][128] which leaves the SM as ONE bulk async copy
// (cp.async.bulk.global.shared::cta, the TMA engine) per sequence, so the LSU only sees the 192 B
// of input and the 260 B token row.  Same-pitch paints inside a chunk are ordered by rank rounds so
// the result equals the sequential oracle bit for bit (matters for the velocity roll).
#include "msx_common.cuh"

namespace {

constexpr int kPitches = 128;
constexpr int kNoteOnFirst = 3;      // MIDIUtil/defaults.py:51
constexpr int kNoteOffFirst = 131;   // MIDIUtil/defaults.py:53
constexpr int kShiftFirst = 259;     // MIDIUtil/defaults.py:55
constexpr int kMaxTicks = 1000;      // MIDIUtil/defaults.py:38
constexpr int kTicksPerBin = 30;     // MIDIUtil/defaults.py:40
constexpr int kPad = 0, kSos = 1;    // MIDIUtil/defaults.py:44-45

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ long long scan_incl_ll(long long v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(MSX_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ int scan_incl_i(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(MSX_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// per-warp shared memory: tile [S*128] bytes | token row [(L+1)] int32 | on_slice [128] int32 | on_val [128] u8
struct WarpSmem {
  uint8_t* tile;
  int* row;
  int* on_slice;
  uint8_t* on_val;
};

template <bool kVelocity>
__global__ void __launch_bounds__(256) rasterize_kernel(const int* __restrict__ dtick, const uint8_t* __restrict__ pitch,
                                                        const uint8_t* __restrict__ vel,
                                                        const int* __restrict__ seq_offsets, int n_seq, int res, int spq,
                                                        int S, int L, int* __restrict__ tokens,
                                                        uint8_t* __restrict__ roll, int* __restrict__ n_tokens,
                                                        int warp_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  uint8_t* base = smem + (size_t)warp * warp_bytes;
  WarpSmem ws;
  ws.tile = base;
  const int tile_bytes = S * kPitches;
  ws.row = reinterpret_cast<int*>(base + tile_bytes);
  ws.on_slice = ws.row + ((L + 1 + 3) & ~3);
  ws.on_val = reinterpret_cast<uint8_t*>(ws.on_slice + kPitches);

  const long long total_warps = (long long)gridDim.x * warps_per_block;
  long long seq = (long long)blockIdx.x * warps_per_block + warp;
  bool store_pending = false;

  for (; seq < n_seq; seq += total_warps) {
    const int e0 = __ldg(seq_offsets + seq), e1 = __ldg(seq_offsets + seq + 1);
    // issue the first chunk's loads before we wait for the previous tile to drain
    int e = e0 + lane;
    bool valid = e < e1;
    int d = valid ? __ldg(dtick + e) : 0;
    // MIDI data bytes are 7-bit; a malformed byte >= 0x80 is masked (as the oracle does) so that the 128-entry pitch
    // table / tile column indexed below can never be addressed out of range and token ids stay below the vocabulary
    int p = valid ? (__ldg(pitch + e) & 0x7F) : 0;
    int v = valid ? (__ldg(vel + e) & 0x7F) : 0;

    if (store_pending) {
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
    }
    // zero tile, init row + pitch table
    {
      uint4 z = make_uint4(0, 0, 0, 0);
      uint4* t4 = reinterpret_cast<uint4*>(ws.tile);
      for (int i = lane; i < tile_bytes / 16; i += 32) t4[i] = z;
      for (int i = lane; i <= L; i += 32) ws.row[i] = (i == 0) ? kSos : kPad;
      for (int i = lane; i < kPitches; i += 32) ws.on_slice[i] = -1;
    }
    __syncwarp();

    long long clock = 0;   // played clock (ticks)
    int tokpos = 0;        // tokens emitted so far (untruncated)
    bool roll_open = true;  // false once an event fell beyond the window (all later ones do too)

    for (int c = e0; c < e1; c += 32) {
      if (c != e0) {
        e = c + lane;
        valid = e < e1;
        d = valid ? __ldg(dtick + e) : 0;
        p = valid ? (__ldg(pitch + e) & 0x7F) : 0;
        v = valid ? (__ldg(vel + e) & 0x7F) : 0;
      }
      // ---- tokens (midi_io.py:81-89): ceil(d/1000) shift tokens of bin (d%1000)/30, then the note token
      const int n_shift = (valid && d > 0) ? (int)(((long long)d + kMaxTicks - 1) / kMaxTicks) : 0;
      const int bin = d > 0 ? (d % kMaxTicks) / kTicksPerBin : 0;
      const int ntok = valid ? n_shift + 1 : 0;
      const int incl = scan_incl_i(ntok, lane);
      int pos = tokpos + incl - ntok;
      if (valid && pos < L) {
        const int lim = min(n_shift, L - pos);
        for (int k = 0; k < lim; ++k) ws.row[1 + pos + k] = kShiftFirst + bin;
        if (pos + n_shift < L) ws.row[1 + pos + n_shift] = (v > 0 ? kNoteOnFirst : kNoteOffFirst) + p;
      }
      tokpos += __shfl_sync(MSX_FULL, incl, 31);

      // ---- played clock: every shift token advances 30*bin ticks (Melody.py:82-83)
      const long long pd = (long long)n_shift * (kTicksPerBin * bin);
      const long long cincl = scan_incl_ll(pd, lane);
      const long long myclock = clock + cincl;
      clock += __shfl_sync(MSX_FULL, cincl, 31);

      if (roll_open) {
        const long long sl = (myclock * spq) / res;
        const bool in_win = valid && sl < S;
        const int s = in_win ? (int)sl : 0;
        // previous same-pitch event inside this chunk
        const unsigned grp = __match_any_sync(MSX_FULL, in_win ? p : (0x100 + lane));
        const unsigned below = grp & ((1u << lane) - 1u);
        const int j = below ? 31 - __clz(below) : lane;
        const int pv = __shfl_sync(MSX_FULL, v, j);
        const int psl = __shfl_sync(MSX_FULL, s, j);
        const int rank = __popc(below);
        int a = -1, val = 0;
        if (in_win) {
          if (below) {
            if (pv > 0) { a = psl; val = kVelocity ? pv : 1; }
          } else {
            a = ws.on_slice[p];
            val = ws.on_val[p];
          }
        }
        __syncwarp();
        // last event of each pitch in the chunk publishes the pitch state
        if (in_win && (grp >> lane) == 1u) {
          ws.on_slice[p] = v > 0 ? s : -1;
          ws.on_val[p] = kVelocity ? (uint8_t)v : (uint8_t)1;
        }
        int end = -1;
        if (in_win && a >= 0) {
          if (v > 0) {
            end = max(a, s - 1);                                      // re-trigger closes the old note
          } else {
            const long long end_excl = (myclock * spq + res - 1) / res;  // ceil
            end = max(a, (int)min(end_excl - 1, (long long)S));
          }
          end = min(end, S - 1);
        }
        const int max_rank = __reduce_max_sync(MSX_FULL, (in_win && a >= 0) ? rank : -1);
        for (int r = 0; r <= max_rank; ++r) {
          if (in_win && a >= 0 && rank == r) {
            uint8_t* col = ws.tile + p;
            for (int t = a; t <= end; ++t) col[t * kPitches] = (uint8_t)val;
          }
          __syncwarp();
        }
        // once any valid lane is beyond the window the rest of the sequence is too
        if (__any_sync(MSX_FULL, valid && !in_win)) roll_open = false;
      }
    }
    __syncwarp();
    // notes still sounding run to the end of the window
    for (int q = lane; q < kPitches; q += 32) {
      const int a = ws.on_slice[q];
      if (a >= 0) {
        const uint8_t val = ws.on_val[q];
        for (int t = a; t < S; ++t) ws.tile[t * kPitches + q] = val;
      }
    }
    __syncwarp();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) bulk_store_s2g(roll + (size_t)seq * tile_bytes, ws.tile, (uint32_t)tile_bytes);
    store_pending = true;
    // token row + count (4-byte aligned rows: plain coalesced stores)
    int* trow = tokens + (size_t)seq * (L + 1);
    for (int i = lane; i <= L; i += 32) trow[i] = ws.row[i];
    if (lane == 0) n_tokens[seq] = tokpos;
    __syncwarp();
  }
  if (store_pending && lane == 0) bulk_wait_all0();
}

}  // namespace

extern "C" int msx_rasterize(const int32_t* dtick, const uint8_t* pitch, const uint8_t* vel, const int32_t* seq_offsets,
                             int n_seq, int resolution, int slices_per_quarter, int n_slices, int max_seq_len,
                             int velocity_roll, int32_t* tokens, uint8_t* roll, int32_t* n_tokens, void* stream) {
  MSX_REQUIRE(n_seq >= 0, "msx_rasterize: n_seq < 0");
  if (n_seq == 0) return MSX_OK;
  MSX_REQUIRE(dtick && pitch && vel && seq_offsets && tokens && roll && n_tokens, "msx_rasterize: null pointer");
  MSX_REQUIRE(resolution > 0 && slices_per_quarter > 0, "msx_rasterize: resolution and slices_per_quarter must be > 0");
  MSX_REQUIRE(n_slices > 0 && n_slices <= 1024, "msx_rasterize: n_slices must be in [1,1024]");
  MSX_REQUIRE(max_seq_len > 0 && max_seq_len <= 49152, "msx_rasterize: max_seq_len must be in [1,49152]");
  MSX_REQUIRE(((uintptr_t)roll & 15) == 0, "msx_rasterize: roll must be 16-byte aligned");
  const int tile_bytes = n_slices * kPitches;
  int warp_bytes = tile_bytes + ((max_seq_len + 1 + 3) & ~3) * 4 + kPitches * 4 + kPitches;
  warp_bytes = (warp_bytes + 127) & ~127;
  const int smem_cap = 227 * 1024;
  MSX_REQUIRE(warp_bytes <= smem_cap, "msx_rasterize: n_slices*128 + row does not fit in shared memory");
  // resident warps per SM bounded by shared memory; 8 warps per CTA when they fit
  int warps_per_block = 8;
  while (warps_per_block > 1 && warps_per_block * warp_bytes > smem_cap / 3) warps_per_block >>= 1;
  if (warps_per_block * warp_bytes > smem_cap) warps_per_block = 1;
  const int smem_bytes = warps_per_block * warp_bytes;
  const int ctas_per_sm = max(1, min(8, smem_cap / smem_bytes));
  auto kern = velocity_roll ? rasterize_kernel<true> : rasterize_kernel<false>;
  MSX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  long long want = ((long long)n_seq + warps_per_block - 1) / warps_per_block;
  int grid = (int)min((long long)msx_num_sms() * ctas_per_sm, want);
  kern<<<grid, warps_per_block * 32, smem_bytes, (cudaStream_t)stream>>>(dtick, pitch, vel, seq_offsets, n_seq,
                                                                         resolution, slices_per_quarter, n_slices,
                                                                         max_seq_len, tokens, roll, n_tokens, warp_bytes);
  MSX_LAUNCH_CHECK();
  return MSX_OK;
}
