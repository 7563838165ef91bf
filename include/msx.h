/* libmsx.so — C ABI of the B200-native (sm_100a) hot path of slyforce/MusicStyleTransfer.
 *
 * The reference (pure Python on MXNet 1.3) has no FFI layer; its boundary for this path is the L3
 * Python surface.  Each entry point below names the reference code it replaces
 * (paths relative to /root/reference/music_style_transfer).  The Python host side
 * (musicstyletransfer_b200/ops.py, ctypes) binds exactly these symbols; INTEGRATION.md shows the stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (fp32 unless typed otherwise); kernels
 *     never allocate, never synchronise, and are enqueued on `stream` (a cudaStream_t) — so a whole
 *     step is CUDA-graph capturable;
 *   - matrices are row-major, leading dimensions in elements;
 *   - return value 0 = ok, <0 = error (MSX_ERR_*); msx_last_error() holds the message (thread-local);
 *   - no C++ types or exceptions cross this boundary.
 */
#ifndef MSX_H_
#define MSX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSX_OK 0
#define MSX_ERR_ARG (-1)
#define MSX_ERR_CUDA (-2)
#define MSX_ERR_UNSUPPORTED (-3)

int msx_version(void);
const char* msx_last_error(void);
int msx_device_sm_count(void);
/* CUDA-graph support.  Every dropout site and the eps ~ N(0,1) fill derive their stream from (seed, site, element); a
 * graph freezes the host-side `seed` argument, so a device-side counter can be registered that every such kernel
 * launched afterwards adds to its seed (NULL = off).  msx_step_counter_tick increments it on the stream (one node at
 * the end of the captured step).  The gluon loop this replaces draws fresh masks per call (trainer.py:166-168). */
int msx_set_step_counter(unsigned long long* dev_counter);
int msx_step_counter_tick(unsigned long long* dev_counter, void* stream);
/* Programmatic dependent launch of the library's kernels (default on; MSX_PDL=0 in the environment also turns it off):
 * a kernel's grid is scheduled while its stream predecessor still runs and waits on the device for it to complete, in
 * eager streams and captured graphs alike.  Results are identical either way; the gluon loop this replaces has no
 * counterpart (one engine push per operator, trainer.py:155-179). */
int msx_set_pdl(int on);
int msx_get_pdl(void);

/* Keep-mask of one dropout site exactly as the step's kernels draw it: out[e] = 1 when element e (row-major index into
 * the site's [rows, width] activation) is kept, for (seed [+ registered step counter], site, drop_p).  Sites of Transformer
 * layer l (encoder: l, decoder: 8 + l): 16 l + {0: attention output [M, D], 1: FF hidden [M, 4D], 2: FF output [M, D]}
 * (gluon.nn.Dropout, transformer.py:44,155,158,200).  Used to replay a dropout step in a checker. */
int msx_dropout_mask(uint8_t* out, long long n, float drop_p, unsigned long long seed, unsigned site, void* stream);

/* (f3) — Standard MIDI File bytes -> note-event SoA for K1.  HOST function (plain C++ inside libmsx.so, no device work,
 * host pointers).  Replaces python-midi's midi.read_midifile + the event walk of EventBasedMIDIReader.read_file /
 * _parse_track (MIDIUtil/midi_io.py:35-93) and _extract_bpm (:16-25): every event's delta time advances the clock, Note-On /
 * Note-Off events emit (ticks since the previous note event, data[0], data[1]); bpm = first Set-Tempo event, else 120.
 * Events of track t occupy [track_offsets[t], track_offsets[t+1]); track_tokens[t] = number of tokens A1 makes of the track
 * (the caller's "< 10 tokens: discard" filter, midi_io.py:60-63).  Capacities: note events <= n_bytes / 3, tracks = the
 * header's count (read it with event_capacity = 0: the call then fails with info filled in).  Errors: MSX_ERR_ARG for
 * malformed files (bad chunk ids, truncated events, running status without a status byte, data bytes >= 0x80),
 * MSX_ERR_UNSUPPORTED for SMPTE time division. */
typedef struct msx_smf_info {
  int resolution;      /* ticks per quarter note */
  int format;          /* SMF format 0 / 1 / 2 */
  int n_tracks;
  double bpm;
  long long n_events;  /* note events in the file (also when they did not fit) */
} msx_smf_info;
int msx_smf_parse(const uint8_t* bytes, long long n_bytes, long long event_capacity, int track_capacity, int32_t* dtick,
                  uint8_t* pitch, uint8_t* vel, int32_t* track_offsets, int32_t* track_tokens, msx_smf_info* info);

/* K1 — note-event rasteriser.  Replaces EventBasedMIDIReader._parse_track (MIDIUtil/midi_io.py:70-93),
 * create_{note_on,note_off,timeshift}_event (MIDIUtil/Melody.py:109-126) and the clock semantics of
 * MelodyWriter._write_track (MIDIUtil/midi_io.py:119-127) for N independent sequences.
 * dtick int32[E] = ticks since the previous note event; pitch, vel uint8[E]; seq_offsets int32[N+1].
 * tokens int32[N, max_seq_len+1] (SOS + first L ids, PAD-filled); roll uint8[N, n_slices, 128]
 * (16-byte aligned); n_tokens int32[N] = untruncated token count. */
int msx_rasterize(const int32_t* dtick, const uint8_t* pitch, const uint8_t* vel, const int32_t* seq_offsets,
                  int n_seq, int resolution, int slices_per_quarter, int n_slices, int max_seq_len, int velocity_roll,
                  int32_t* tokens, uint8_t* roll, int32_t* n_tokens, void* stream);

/* K2 — dense layers.  Replaces every gluon.nn.Dense forward on the path (VarAutoEncoder/transformer.py:36-40,
 * 65-68,88-93,104; model.py:70-71,139-157,214-227) and its autograd backward (trainer.py:176).
 * C[M,N] = epilogue(opA(A)[M,K] * opB(B)[K,N]), transX=0: stored as written, transX=1: stored transposed.
 *   forward  Y = X W^T + b        : transA=0, transB=1, bias, optional relu + dropout(drop_p, seed, site)
 *   dgrad    dX = dY W            : transA=0, transB=0, optional aux mask (aux>0 ? aux_scale : 0), accumulate
 *   wgrad    dW += dY^T X, db += : transA=1, transB=0, splitk>1 (atomic adds into a zeroed C), colsum = db
 * msx_gemm_f32 is the exact-fp32 FFMA kernel; msx_gemm_tc runs the same contract on the tcgen05 tensor
 * cores (TF32 or BF16 operands, fp32 accumulate in TMEM) for the shapes it supports. */
int msx_gemm_f32(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc, int M,
                 int N, int K, const float* bias, int relu, float drop_p, unsigned long long seed, unsigned site,
                 const float* aux, int ldaux, float aux_scale, int accumulate, int splitk, float* colsum,
                 void* stream);

int msx_gemm_tc(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc, int M,
                int N, int K, const float* bias, int relu, float drop_p, unsigned long long seed, unsigned site,
                const float* aux, int ldaux, float aux_scale, int accumulate, int splitk, float* out_colsum,
                void* stream);
/* out_colsum (optional, plain stores only): out_colsum[n] += sum_m C[m,n] — the bias gradient of the layer whose
 * pre-activation gradient this dgrad produces, folded into the epilogue.
 * msx_gemm_tc writes C through TMA in 16-byte chunks: padding columns N .. roundup4(N)-1 of C may be overwritten.
 * 1 when msx_gemm_tc accepts the operands (TMA: 16-byte aligned bases, leading dimensions multiple of 4). */
int msx_gemm_tc_supported(const float* A, int lda, const float* B, int ldb, const float* C, int ldc, int M, int N,
                          int K);
/* msx_gemm_tc has two kernels: 128x128 tiles on one CTA, and 256x256 / 256x128 tiles on a CTA pair
 * (tcgen05 cta_group::2, used when M > 128, N >= 64 and K >= 256).  msx_gemm_tc_set_pair(0) forces the 1-CTA kernel,
 * (1) re-enables the pair kernel; returns the previous setting.  MSX_GEMM_PAIR=0 in the environment does the same. */
int msx_gemm_tc_set_pair(int enable);

/* bf16 variant of msx_gemm_tc (BASELINE config 4): A and B are bfloat16 in HBM (leading dimensions in elements,
 * multiples of 8) and feed tcgen05.mma.kind::f16 with fp32 accumulation in TMEM; the epilogue arithmetic is fp32 and
 * the same as msx_gemm_tc's.  C is fp32 (c_bf16=0; may accumulate / split-K through fp32 TMA reduce-adds) or bfloat16
 * (c_bf16=1, plain stores only, ldc multiple of 8).  aux (the ReLU-mask source) is fp32 (aux_bf16=0) or bfloat16. */
int msx_gemm_tc_bf16(const void* A, int lda, int transA, const void* B, int ldb, int transB, void* C, int ldc, int c_bf16,
                     int M, int N, int K, const float* bias, int relu, float drop_p, unsigned long long seed,
                     unsigned site, const void* aux, int ldaux, int aux_bf16, float aux_scale, int accumulate,
                     int splitk, float* out_colsum, void* stream);
int msx_gemm_tc_bf16_supported(const void* A, int lda, const void* B, int ldb, const void* C, int ldc, int c_bf16, int M,
                               int N, int K);
/* General form of msx_gemm_tc / msx_gemm_tc_bf16 (ab_bf16 / c_bf16 select the storage types) plus the ReLU bit mask:
 * aux_kind 0 = fp32 matrix, 1 = bfloat16 matrix, 2 = bit mask uint32 [M, ldaux words] (bit j of word [m, n/32] <=>
 * element [m, 32*(n/32)+j] > 0); mask_out (optional, N % 32 == 0, plain stores) receives that mask of the values this
 * launch writes after bias / ReLU / dropout.  FF1 forward emits the mask, FF2 dgrad reads 4 bytes per 32 elements instead
 * of re-reading the hidden activation (transformer.py:36-46 backward). */
int msx_gemm_tc_ex(const void* A, int lda, int transA, const void* B, int ldb, int transB, void* C, int ldc, int M, int N,
                   int K, int ab_bf16, int c_bf16, const float* bias, int relu, float drop_p, unsigned long long seed,
                   unsigned site, const void* aux, int ldaux, int aux_kind, float aux_scale, int accumulate, int splitk,
                   float* out_colsum, uint32_t* mask_out, int ldmask, void* stream);
/* "p3" forward GEMM (the gluon.nn.Dense forwards of transformer.py:36-40,65-68,88-93,104): Y = X W^T + b with both operands
 * given as bfloat16 hi / lo PLANES (hi = rn_bf16(x), lo = rn_bf16(x - hi)) written by the kernels that produce x
 * (msx_embed_fwd_p, msx_add_ln_fwd_p, msx_attention_tc_fwd_p, this function's C planes, msx_split_planes for weights).
 * The reduction is walked three times, A_hi B_hi + A_hi B_lo + A_lo B_hi, on kind::f16 with fp32 accumulation: products
 * to ~2^-17 relative (the accuracy class of 3xTF32 for this step) with no in-kernel conversion stage.  A [M,K], B [N,K]
 * K-major, K % 64 == 0.  c_kind 0: fp32 C; 2: C / C_lo receive the bf16 hi / lo planes of the result (plain store). */
int msx_gemm_tc_p3_supported(const void* A_hi, int lda, const void* B_hi, int ldb, const void* C, int ldc, int c_kind, int M,
                             int N, int K);
int msx_gemm_tc_p3(const void* A_hi, const void* A_lo, int lda, const void* B_hi, const void* B_lo, int ldb, void* C,
                   void* C_lo, int ldc, int c_kind, int M, int N, int K, const float* bias, int relu, float drop_p,
                   unsigned long long seed, unsigned site, int accumulate, uint32_t* mask_out, int ldmask, void* stream);
/* fp32 -> bf16 hi / lo planes, n % 4 == 0 (the weight arena once per step; small activations no kernel emits as planes). */
int msx_split_planes(const float* src, void* hi, void* lo, long long n, void* stream);
/* Strict-fp32 tensor-core GEMM (the reference step is fp32 end to end, trainer.py:155-179): same contract and epilogues as
 * msx_gemm_tc_ex with fp32 operands, but every operand value is split into hi + lo TF32 parts inside the kernel and a
 * k-block contributes A_lo B_hi + A_hi B_lo + A_hi B_hi ("3xTF32"), fp32 accumulation in TMEM: products carry ~2^-21
 * relative error instead of 2^-11.  cta_group::2 pair tiles, any M, N >= 64 (msx_gemm_tc_x3_supported; the call
 * returns MSX_ERR_UNSUPPORTED otherwise and the caller uses msx_gemm_f32).  aux_kind 0 (fp32 matrix) or 2 (bit mask). */
int msx_gemm_tc_x3_supported(const float* A, int lda, const float* B, int ldb, const float* C, int ldc, int M, int N,
                             int K);
int msx_gemm_tc_x3(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C, int ldc, int M,
                   int N, int K, const float* bias, int relu, float drop_p, unsigned long long seed, unsigned site,
                   const void* aux, int ldaux, int aux_kind, float aux_scale, int accumulate, int splitk,
                   float* out_colsum, uint32_t* mask_out, int ldmask, void* stream);
/* Forward Dense GEMM Y = X W^T (+ bias / ReLU / dropout / ReLU bit mask) with fp32 X [M,K], W [N,K], Y in HBM and "bf16x3"
 * products: each operand value is split inside the kernel into bf16 hi + lo parts (16 mantissa bits) and a k-step
 * contributes A_lo B_hi + A_hi B_lo + A_hi B_hi on tcgen05 kind::f16 (twice the TF32 rate: 1.5 single-pass MMAs instead of
 * the 3 of msx_gemm_tc_x3).  ~2^-17 relative product error instead of TF32's 2^-11.  Any M, N >= 64; accumulate adds into
 * Y.  Used for the encoder's forward GEMMs (transformer.py:36-40,65-68,88-93,104; model.py:100) in the default mode. */
int msx_gemm_tc_b3_supported(const float* A, int lda, const float* B, int ldb, const float* C, int ldc, int M, int N,
                             int K);
int msx_gemm_tc_b3(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                   const float* bias, int relu, float drop_p, unsigned long long seed, unsigned site, int accumulate,
                   uint32_t* mask_out, int ldmask, void* stream);
/* dst[i] = bfloat16(src[i]) (round to nearest even), i < n: builds bf16 operands from fp32 tensors (weights shadow,
 * tests).  src and dst 16-byte aligned. */
int msx_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);

/* out[n] += sum_m X[m,n]: bias gradient of a Dense layer when the wgrad runs on the tensor path. */
int msx_colsum(const float* X, int ld, long long M, int N, float* out, void* stream);

/* K2c — attention in the reference's convention (softmax over the QUERY axis, additive -1e9 on padded keys,
 * O = P^T V).  Replaces MultiHeadDotAttention.hybrid_forward lines 91-103 and _mask_logits
 * (VarAutoEncoder/transformer.py:91-126).  qkv [B*T, 3*H*dh] rows = [K | Q | V]; mask [B*T] (1 = real key). */
int msx_attention_fwd(const float* qkv, const float* mask, float* ctx, int B, int T, int H, int dh, void* stream);
int msx_attention_bwd(const float* qkv, const float* mask, const float* dctx, float* dqkv, int B, int T, int H,
                      int dh, void* stream);

/* Any row length (--max-seq-len is a free flag, VarAutoEncoder/config.py:36): key-tiled exact-fp32 kernels that keep nothing
 * larger than a 32 x 32 tile on chip; the context / the dQ part of dqkv are accumulated with atomics over the key tiles
 * (both outputs are zero-filled by the call).  msx_attention_fwd / _bwd route here when T x T does not fit one SM's shared
 * memory (T > ~180 at d_h = 32); the engine uses them for T > 768, beyond the tcgen05 kernels, and in the exact modes.  d_h <= 64, B * H <= 65535. */
int msx_attention_tiled_fwd(const float* qkv, const float* mask, float* ctx, int B, int T, int H, int dh, void* stream);
int msx_attention_tiled_bwd(const float* qkv, const float* mask, const float* dctx, float* dqkv, int B, int T, int H,
                            int dh, void* stream);

/* Tensor-core attention (same contract, d_h == 32 and T <= 128): S = K Q^T and O = P^T V on tcgen05, keys on the TMEM
 * lanes so the query-axis softmax is thread-local. */
int msx_attention_tc_supported(const float* qkv, int T, int dh);
int msx_attention_tc_fwd(const float* qkv, const float* mask, float* ctx, int B, int T, int H, int dh, void* stream);
int msx_attention_tc_bwd(const float* qkv, const float* mask, const float* dctx, float* dqkv, float* dbias, int B,
                         int T, int H, int dh, void* stream);
/* Profiling hook for the pipelined tensor-core backward: device buffer of 2 x 16 x 9 int64 clock64 stamps (block 0,
 * per pipeline group, first 16 items); NULL disables. */
/* bf16 variant: ctx (forward) / dqkv (backward) are written as bfloat16 when the flag is set (same element layout). */
int msx_attention_tc_fwd_ex(const float* qkv, const float* mask, void* ctx, int ctx_bf16, int B, int T, int H, int dh,
                            void* stream);
/* x3_scores: S = K Q^T accumulates K_lo Q_hi + K_hi Q_lo + K_hi Q_hi (operands split inside the kernel): fp32-equivalent
 * scores.  The softmax (transformer.py:100) turns an absolute score error into a relative error of P, which makes the
 * single-pass TF32 scores the largest forward error of the step. */
int msx_attention_tc_fwd_ex2(const float* qkv, const float* mask, void* ctx, int ctx_bf16, int x3_scores, int B, int T, int H,
                             int dh, void* stream);
/* p3 planes: ctx (bfloat16, ctx_bf16 != 0) receives the hi plane rn_bf16(o) and ctx_lo (optional) rn_bf16(o - hi): the two
 * operands of the p3 W_proj GEMM (msx_gemm_tc_p3); the context then never exists in fp32.
 * q0_only (d_h == 32): only the context row of query 0 of every sequence is computed and written (the encoder's top layer is
 * read at position 0 only, model.py:97-100): scores and the query-axis softmax as before, then O[0] = sum_k P[k][0] V[k]
 * as a reduction over the key rows; the other rows of ctx are left untouched. */
int msx_attention_tc_fwd_p(const float* qkv, const float* mask, void* ctx, void* ctx_lo, int ctx_bf16, int x3_scores,
                           int q0_only, int B, int T, int H, int dh, void* stream);
int msx_attention_tc_bwd_ex(const float* qkv, const float* mask, const float* dctx, void* dqkv, int dqkv_bf16, float* dbias,
                            int B, int T, int H, int dh, void* stream);
/* q0_only: dctx is non-zero in the row of query 0 of every sequence only (the encoder's top layer under SOS-rows-only,
 * model.py:97-100).  For T <= 80, d_h == 32 a specialised pipeline runs: dV and dS are thread-local (no dP / dV MMAs, no
 * dO tiles); other shapes take the general kernels, which give the same result. */
int msx_attention_tc_bwd_q0(const float* qkv, const float* mask, const float* dctx, void* dqkv, int dqkv_bf16, float* dbias,
                            int q0_only, int B, int T, int H, int dh, void* stream);
int msx_attention_tc_set_trace(long long* buf);
/* Long rows, 128 < T <= 768 (the L = 128 / 256 sweep points and beyond; T > 384: two-sweep forward): key tiles and query
 * chunks of 128, flash-attention style backward.  stats [B*H*T, 2] fp32 (row max * log2 e, 1 / row sum of every key row) is written by the forward and read
 * by the backward; ctx / dqkv are fp32 or bfloat16 by flag; dbias as msx_attention_tc_bwd. */
int msx_attention_tcl_supported(const float* qkv, int T, int dh);
int msx_attention_tcl_fwd(const float* qkv, const float* mask, void* ctx, int ctx_bf16, float* stats, int B, int T, int H,
                          int dh, void* stream);
int msx_attention_tcl_bwd(const float* qkv, const float* mask, const float* dctx, const float* stats, void* dqkv,
                          int dqkv_bf16, float* dbias, int B, int T, int H, int dh, void* stream);   /* dbias (optional) [3*H*dh] += column sums of dqkv */
/* q0_only as msx_attention_tc_fwd_p / msx_attention_tc_bwd_q0 (the encoder's top layer is read at position 0 only,
 * model.py:97-100).  Forward: only the context row of query 0 of every sequence is computed and written (a column sum over
 * the key rows: no P staging, no second MMA); stats cover every key row as before.  Backward: the caller guarantees that
 * dctx is zero outside the row of query 0 of every sequence; dV and dS are then thread-local (no dP / dV MMAs, no dO
 * tiles) and only that row of dctx is read. */
int msx_attention_tcl_fwd_q0(const float* qkv, const float* mask, void* ctx, int ctx_bf16, float* stats, int q0_only, int B,
                             int T, int H, int dh, void* stream);
int msx_attention_tcl_bwd_q0(const float* qkv, const float* mask, const float* dctx, const float* stats, void* dqkv,
                             int dqkv_bf16, float* dbias, int q0_only, int B, int T, int H, int dh, void* stream);
/* p3 planes (as msx_attention_tc_fwd_p): ctx (bfloat16) receives the hi plane rn_bf16(o), ctx_lo (optional) rn_bf16(o - hi) —
 * the operands of the p3 W_proj GEMM; the context then never exists in fp32 and no msx_split_planes pass follows. */
int msx_attention_tcl_fwd_p(const float* qkv, const float* mask, void* ctx, void* ctx_lo, int ctx_bf16, float* stats,
                            int q0_only, int B, int T, int H, int dh, void* stream);

/* K2d — out = LayerNorm(x + dropout(y)).  Replaces transformer.py:155,158,200 (gluon Dropout + add +
 * gluon.nn.LayerNorm, eps 1e-5).  Backward: dres = ds, dy = ds*keep (dy may be NULL when drop_p == 0);
 * fuse_xy: x and y alias (decoder's ln3(f + drop(f))) and dres receives ds*(1+keep). */
int msx_add_ln_fwd(const float* x, const float* y, const float* gamma, const float* beta, float* out, float* mean,
                   float* rstd, long long M, int D, float eps, float drop_p, unsigned long long seed, unsigned site,
                   void* stream);
int msx_add_ln_bwd(const float* x, const float* y, const float* gamma, const float* mean, const float* rstd,
                   const float* dout, float* dres, float* dy, float* dgamma, float* dbeta, float* dybias, long long M,
                   int D, float drop_p, unsigned long long seed, unsigned site, int accumulate_dres, int fuse_xy,
                   void* stream);   /* dybias (optional) [D] += column sums of the y-gradient */

/* bf16 variant: the same kernels with an additional bfloat16 copy of the output that the bf16 GEMMs read
 * (out_bf16: LN output; dy_bf16: gradient of the Dense output y, or the combined gradient under fuse_xy).  Either may be
 * NULL; dy may be NULL when only the bf16 gradient is wanted.  Needs D % 128 == 0. */
int msx_add_ln_fwd_ex(const float* x, const void* y, int y_bf16, const float* gamma, const float* beta, float* out,
                      void* out_bf16, float* mean, float* rstd, long long M, int D, float eps, float drop_p,
                      unsigned long long seed, unsigned site, void* stream);
/* p3 planes: out_bf16 / out_bf16_lo receive rn_bf16(out) and rn_bf16(out - hi) (operands of msx_gemm_tc_p3). */
int msx_add_ln_fwd_p(const float* x, const void* y, int y_bf16, const float* gamma, const float* beta, float* out,
                     void* out_bf16, void* out_bf16_lo, float* mean, float* rstd, long long M, int D, float eps, float drop_p,
                     unsigned long long seed, unsigned site, void* stream);
int msx_add_ln_bwd_ex(const float* x, const void* y, int y_bf16, const float* gamma, const float* mean, const float* rstd,
                      const float* dout, float* dres, float* dy, void* dy_bf16, float* dgamma, float* dbeta, float* dybias,
                      long long M, int D, float drop_p, unsigned long long seed, unsigned site, int accumulate_dres,
                      int fuse_xy, void* stream);   /* y_bf16: the Dense output y is bfloat16 (read only by this LayerNorm) */

/* A2 — dataset rows on the device.  Replace MelodyDataset._get_token_arrays / _count_sequence_length and the per-batch part
 * of _preprocess_batch (VarAutoEncoder/data.py:133-198).  Input: K1's token streams, one per track — ids[t * ld + col0 + j],
 * j < n_tokens[t] — with the tracks grouped by class in the reference's iteration order (class_start[c] .. class_start[c+1]).
 * msx_rows_plan (one CTA) -> row_start[n_tracks + 1] (row_start[n_tracks] = total rows R), dup_row / dup_src [n_classes]
 * (the duplicated last row per class, data.py:152-155) and len_present[L + 1] (the columns `labels[:, seq_lens] = EOS`
 * hits, data.py:166-168).  msx_rows_build -> tokens / labels int32 [R, L + 1], classes [R], seq_lens [R] (non-PAD count
 * including SOS).  msx_rows_gather_batch gathers `batch` rows by index, trimmed to t_out columns (data.py:196-198). */
int msx_rows_plan(const int32_t* n_tokens, const int32_t* class_start, int n_tracks, int n_classes, int max_seq_len,
                  int32_t* row_start, int32_t* dup_row, int32_t* dup_src, int32_t* len_present, void* stream);
int msx_rows_build(const int32_t* ids, long long ld, int col0, const int32_t* n_tokens, const int32_t* track_class,
                   int n_tracks, int n_classes, int max_seq_len, const int32_t* row_start, const int32_t* dup_row,
                   const int32_t* dup_src, const int32_t* len_present, int32_t* tokens, int32_t* labels, int32_t* classes,
                   int32_t* seq_lens, void* stream);
int msx_rows_gather_batch(const int32_t* tokens, const int32_t* labels, const int32_t* classes, const int32_t* seq_lens,
                          const int32_t* index, int batch, int ld, int t_out, int32_t* b_tokens, int32_t* b_labels,
                          int32_t* b_classes, int32_t* b_seq_lens, void* stream);

/* out[b, 0] = PAD, out[b, 1 + t] = labels[b, t]: labels for the Transformer decoder's rows, which carry the latent prefix
 * position in front (model.py:244-253). */
int msx_prefix_labels(const int32_t* labels, int32_t* out, int B, int T, void* stream);

/* Strided row copy (add = 0) / accumulate (add = 1): out[r, :width] (+)= in[r, :width], row r of X at X + r * ldX.  The
 * encoder output is only read at the SOS position (model.py:97-100), so the top encoder layer runs on one row per sequence
 * after its attention; this moves those rows between the [B*T, D] and [B, D] layouts.  width, ld % 4 == 0. */
int msx_rows_strided(const float* in, long long ld_in, float* out, long long ld_out, int rows, int width, int add,
                     void* stream);

/* K2a — embedding front end.  Replaces model.py:81-91 + transformer.py:270 (encoder), model.py:241-247 +
 * transformer.py:237 (Transformer decoder, prefix = 1 latent-state row), model.py:176 (LSTM decoder).
 * out[b,pos,:] = scale*((pos<prefix ? prefix_vec[b] : tok_emb[tokens[b,pos-prefix]]) + cls_emb[classes[b]]) + pe[pos];
 * mask[b,pos] = seq_lens ? pos < seq_lens[b]+prefix : token != 0.  NULL disables a term. */
int msx_embed_fwd(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens, const float* tok_emb,
                  const float* cls_emb, const float* prefix_vec, const float* pe, float* out, float* mask, int B, int T,
                  int D, int prefix, float scale, int vocab, void* stream);
/* bf16 variant: also (out may be NULL: only) writes a bfloat16 copy of the rows. */
int msx_embed_fwd_ex(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens, const float* tok_emb,
                     const float* cls_emb, const float* prefix_vec, const float* pe, float* out, void* out_bf16, float* mask,
                     int B, int T, int D, int prefix, float scale, int vocab, void* stream);
/* p3 planes: as msx_embed_fwd_ex plus the lo plane rn_bf16(row - hi) next to the bfloat16 (hi) copy. */
int msx_embed_fwd_p(const int32_t* tokens, const int32_t* classes, const int32_t* seq_lens, const float* tok_emb,
                    const float* cls_emb, const float* prefix_vec, const float* pe, float* out, void* out_bf16,
                    void* out_bf16_lo, float* mask, int B, int T, int D, int prefix, float scale, int vocab, void* stream);
/* Per-token row sums: out[v, :] += scale * sum of X[r, :] over the rows with tokens[r] == v — the Embedding backward
 * (model.py:141,175) as sort + segmented sum, D up to 1024.  msx_token_sort: counting sort of the row indices by token
 * (workspace int32 [3 V]; perm / sorted_tok int32 [M]); msx_rows_sum_by_token: one coalesced read of every row of X [M, ld],
 * one 16-byte red per column group and token change of a 128-row slice of the sorted order.  In the LSTM decoder's table
 * mode (msx_lstm_tc_fwd_tab) X = d(pre-activations) [B*T, 4H] and out is the gradient of the [V, 4H] i2h table. */
int msx_token_sort(const int32_t* tokens, long long M, int V, int32_t* perm, int32_t* sorted_tok, int32_t* workspace,
                   void* stream);
int msx_rows_sum_by_token(const float* X, int ld, int D, const int32_t* perm, const int32_t* sorted_tok, long long M,
                          float scale, float* out, void* stream);
/* First encoder layer's K|Q|V projection from tables: x0 = scale * (emb[token] + class2hid[class]) + pe[position]
 * (model.py:84-93, transformer.py:204-231), hence x0 W^T + b = scale * (tab[C + token] + tab[class]) + postab[position] with
 * tab [C + V, N] = [class2hid; emb] W^T and postab [T, N] = pe[:T] W^T + b computed once per step.  out [B*T, N]. */
int msx_rows_from_tables(const int32_t* tokens, const int32_t* classes, const float* tab, const float* postab, float* out,
                         int B, int T, int N, int C, int V, float scale, void* stream);
int msx_embed_bwd(const int32_t* tokens, const int32_t* classes, const float* dout, float* d_tok_emb, float* d_cls_emb,
                  float* d_prefix, int B, int T, int D, int prefix, float scale, int vocab, void* stream);

/* num_classes = rows of the class table (0 = unknown).  Large problems without prefix rows and with <= 2 classes run a
 * gather kernel (one CTA per vocabulary row and position segment, one atomic per column and CTA) instead of one
 * red.global.add per element. */
int msx_embed_bwd_ex(const int32_t* tokens, const int32_t* classes, const float* dout, float* d_tok_emb, float* d_cls_emb,
                     float* d_prefix, int B, int T, int D, int prefix, float scale, int vocab, int num_classes, void* stream);

/* Piano-roll front end (`--featurisation roll`, the north star's "step over piano-roll tensors ... sigmoid-BCE"; model side:
 * derived spec oracle/roll_model.py).  msx_roll_features: uint8 roll [B, S, 128] (K1 output) -> fp32 GEMM operands
 * renc [B, S+1, 132] (row 0 = start flag in column 128, rows 1..S = (roll > 0)) and rdec [B, S, 132] (row t = renc row t,
 * i.e. the previous slice: teacher forcing).  msx_embed_dense_*: out = scale * (E + cls_emb[classes]) + pe — the
 * Encoder front end (model.py:89-91, transformer.py:270) when the token embedding is a Dense over the multi-hot slice;
 * backward: dE = scale * dout, d_cls_emb[classes[b]] += scale * sum_t dout[b, t]. */
int msx_roll_features(const uint8_t* roll, float* renc, float* rdec, int B, int S, void* stream);
int msx_embed_dense_fwd(const float* E, const int32_t* classes, const float* cls_emb, const float* pe, float* out, int B, int T,
                        int D, float scale, void* stream);
int msx_embed_dense_bwd(const float* dout, const int32_t* classes, float* dE, float* d_cls_emb, int B, int T, int D, float scale,
                        void* stream);

/* out[e] = x[e] * keep(e) / (1 - drop_p) with the step's counter-based masks (seed [+ step counter], site, element): the
 * dropout between the layers of a stacked LSTM decoder (gluon.rnn.LSTM(dropout=...), model.py:148-153).  The backward is the
 * same call on the gradient.  n % 4 == 0; in place when out == x. */
int msx_dropout(const float* x, float* out, long long n, float drop_p, unsigned long long seed, unsigned site, void* stream);

/* K2g — persistent LSTM recurrence.  Replaces the fused gluon.rnn.LSTM call in LSTMDecoder.forward_train
 * (model.py:148-153,179); gx_inout [B,T,4H] holds x W_i2h^T + b_i2h on entry and the gate activations on exit;
 * h0/c0 are rows of a [B, ld0] buffer (model.py:159-167).  Backward turns the saved activations into
 * d(pre-activations) for the surrounding wgrad/dgrad GEMMs and returns dh0/dc0.  H in {32,64,128}. */
int msx_lstm_fwd(float* gx_inout, const float* w_h2h, const float* b_h2h, const float* h0, const float* c0, int ld0,
                 float* hs, float* hprev, float* cs, int B, int T, int H, void* stream);
int msx_lstm_bwd(float* gates_inout, const float* w_h2h, const float* cs, const float* c0, int ld0, const float* dhs,
                 float* dh0, float* dc0, float* db_i2h, float* db_h2h, int B, int T, int H, void* stream);
/* Tensor-core variant (same contract, H == 128, even ld0, 8-byte aligned h0 / c0 / dh0 / dc0): W_h2h resident in
 * registers as TF32 mma.sync fragments, fp32 accumulation, fp32 gate arithmetic. */
int msx_lstm_tc_supported(int H, int ld0, const float* h0, const float* c0);
int msx_lstm_tc_fwd(float* gx_inout, const float* w_h2h, const float* b_h2h, const float* h0, const float* c0, int ld0,
                    float* hs, float* hprev, float* cs, int B, int T, int H, void* stream);
/* Table mode: the layer's input is an embedding lookup (LSTMDecoder.forward_train, model.py:175-179: embedding -> LSTM), so
 * x_t W_i2h^T + b_i2h is row tokens[b, t] of table [V, 4H] = emb W_i2h^T + b_i2h.  The kernel fetches the rows itself (the
 * [B*T, 4H] input pre-activations never exist); gates_out [B*T, 4H] receives the gate activations.  tokens in [0, V). */
int msx_lstm_tc_fwd_tab(float* gates_out, const int32_t* tokens, const float* table, const float* w_h2h, const float* b_h2h,
                        const float* h0, const float* c0, int ld0, float* hs, float* hprev, float* cs, int B, int T, int H,
                        void* stream);
int msx_lstm_tc_bwd(float* gates_inout, const float* w_h2h, const float* cs, const float* c0, int ld0, const float* dhs,
                    float* dh0, float* dc0, float* db_i2h, float* db_h2h, int B, int T, int H, void* stream);

/* K3 — losses.  msx_reparam_kl_*: z = m + eps*s and VariationalKLLoss (model.py:292, loss.py:8-12);
 * msx_ce_*: softmax(output_layer) + SoftmaxCrossEntropy (model.py:182/256, loss.py:16-23) fused on logits
 * [B*T, ld], ce[b] = sum_t mask*nll / denom, metrics[4] += {sum nll (clamped like mx.metric.Perplexity),
 * tokens, top-1 hits, top-k hits} (trainer.py:107-120,181-186, metrics.py); msx_ce_bwd overwrites the logits
 * with the gradient; msx_softmax_rows / msx_ce_from_probs keep the probability-based API (model.py:296,
 * loss.py:16-23); msx_bce: BinaryCrossEntropy on uint8 piano rolls (loss.py:27-81), value and/or gradient; the label of a
 * cell is (roll != 0), so binary and velocity rolls give the same loss. */
int msx_reparam_kl_fwd(const float* lat, const float* eps, float* z, float* kl, int B, int Z, void* stream);
int msx_reparam_kl_bwd(const float* lat, const float* eps, const float* dz, const float* gkl, float kl_weight,
                       float* dlat, int B, int Z, void* stream);
int msx_normal_fill(float* out, long long n, unsigned long long seed, unsigned long long offset, void* stream);
/* Running sums behind the reference's per-step CustomMetric means of the KL and total loss (trainer.py:107-120,181-186):
 * sums[0] += sum_b kl[b], sums[1] += sum_b (ce[b] + kl_weight * kl[b]), sums[2] += B.  One launch, no host sync. */
int msx_loss_sums(const float* ce, const float* kl, float kl_weight, float* sums, int B, void* stream);
int msx_ce_fwd(const float* logits, int ld, const int32_t* labels, float* ce, float* lse, float* metrics, int B, int T,
               int V, int denom, int top_k, void* stream);
int msx_ce_bwd(float* logits_inout, int ld, const int32_t* labels, const float* lse, const float* gout, int B, int T,
               int V, int denom, float* dbias, void* stream);   /* dbias (optional) [V] += column sums of the gradient */
/* msx_ce_fwd followed by msx_ce_bwd with head gradient 1 in one pass over the logits (training step): logits are read
 * once and overwritten with the gradient.  V <= 512, 16-byte aligned rows; MSX_ERR_UNSUPPORTED otherwise. */
int msx_ce_fwd_bwd(float* logits_inout, int ld, const int32_t* labels, float* ce, float* lse, float* metrics, int B, int T,
                   int V, int denom, int top_k, float* dbias, void* stream);
int msx_softmax_rows(const float* logits, int ld, float* probs, long long rows, int V, void* stream);
int msx_ce_from_probs(const float* probs, const int32_t* labels, float* ce, int B, int T, int V, void* stream);
int msx_bce(const float* pred, const uint8_t* label, float* out, const float* gout, float* dpred, int B,
            int n_per_sample, int from_sigmoid, float label_smoothing, int downweight, void* stream);

/* A12 — next-token sampling.  Replaces mx.nd.random.multinomial(probs) and the score update of Sampling.sample
 * (sampler.py:181-184) on softmax(logits): next[b] = first v with cdf(v) > u_b (u from `uniforms` or Philox(seed,
 * step, b)); score[b] += -log p(next) when given; out_seq[b*out_ld + out_col] = next[b] when given. */
int msx_sample_multinomial(const float* logits, int ld, int V, const float* uniforms, unsigned long long seed,
                           unsigned long long step, int32_t* next, float* score, int32_t* out_seq, int out_ld,
                           int out_col, int B, void* stream);

/* A12, beam search (sampler.py:192-257, LSTM decoder): one step over logits [B*beam, ld].  Candidates
 * score_in[hyp] - log softmax(logits[hyp])[v]; a hypothesis whose last token is EOS / PAD is frozen (single candidate
 * PAD at its own score); at step 1 only beam 0 of every row is expanded; the `beam` smallest win (ties: smaller
 * hyp * V + v).  Writes the reordered sequences (seq_out[:, :step] = seq_in[parent], seq_out[:, step] = token), the new
 * scores, parent[B*beam] (flat source hypothesis) and next_tok; unfinished[step] += number of winners that are neither
 * EOS nor PAD (optional; the host finds the reference's stop step, sampler.py:250, from it).  msx_gather_rows reorders
 * the recurrent states. */
int msx_beam_step(const float* logits, int ld, int V, int B, int beam, const int32_t* seq_in, int32_t* seq_out, int seq_ld,
                  int step, const float* score_in, float* score_out, int32_t* parent, int32_t* next_tok, int32_t* unfinished,
                  void* stream);
int msx_gather_rows(const float* in, float* out, const int32_t* parent, int rows, int width, void* stream);

/* K4 — fused multi-tensor Adam over flat arenas.  Replaces gluon.Trainer('adam', ...).step(batch_size)
 * (trainer.py:94-101,177; MXNet 1.3 Adam: eps outside the bias correction, element-wise clip, rescale = 1/batch).
 * state[0] = step count t (device-resident), state[1] = lr_t; zero_grad clears g in the same pass. */
int msx_adam_step(float* w, float* g, float* m, float* v, long long n, float* state, float lr, float beta1, float beta2,
                  float eps, float wd, float rescale, float clip, int zero_grad, void* stream);
/* K5 — data-parallel optimiser step in one kernel over NVLink peer memory: reduce-scatter of the peer-mapped gradient
 * arenas (sum in rank order, the trainer.py:176-177 rule), Adam on the owned slice with local (1/world-sharded)
 * moments, all-gather of the updated parameters into every rank's arena.  peer_g / peer_w / peer_flags are HOST arrays
 * of `world` device pointers (entry `rank` = the local buffer; flags: msx_adam_nvlink_flag_bytes() zeroed bytes per rank),
 * done_counter a zeroed local u32, epoch_counter a zeroed local u64 (the kernel numbers its launches with it), max_ctas (0 = one per SM) bounds
 * the grid.  Every rank must launch it; it completes when the local parameters are complete and the local gradient
 * arena is no longer read by any peer (it is then zeroed if zero_grad). */
int msx_adam_nvlink_flag_bytes(void);
int msx_adam_nvlink_step(float* w, float* g, float* m, float* v, long long n, float* state, const float* const* peer_g,
                         float* const* peer_w, unsigned long long* const* peer_flags, unsigned* done_counter, int rank,
                         int world, unsigned long long* epoch_counter, float lr, float beta1, float beta2, float eps, float wd,
                         float rescale, float clip, int zero_grad, int max_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSX_H_ */
