"""Functional wrappers over the libmsx C ABI (include/msx.h).  No autograd, no allocation policy:
callers own every buffer (torch CUDA tensors), kernels run on torch's current stream."""
import ctypes as C

import torch

from . import lib

_i, _f, _ll, _u64, _u32 = C.c_int, C.c_float, C.c_longlong, C.c_ulonglong, C.c_uint
P = lib.ptr


def _chk(t, dtype=torch.float32):
    assert t.is_cuda and t.dtype == dtype, (t.device, t.dtype)
    return t


def _gemm_tag(M, N, K, accumulate, aux):
    """(flops, algorithmic HBM bytes) of one GEMM launch for bench.py's roofline: every fp32 operand element read once,
    C written once (+ read once when accumulating, + the aux mask read once)."""
    return (2.0 * M * N * K, 4.0 * (M * K + N * K + M * N * (1 + (1 if accumulate else 0) + (1 if aux is not None else 0))),
            "f32 M=%d N=%d K=%d%s%s" % (M, N, K, " acc" if accumulate else "", " aux" if aux is not None else ""))


def gemm(A, lda, transA, B, ldb, transB, Cm, ldc, M, N, K, bias=None, relu=False, drop_p=0.0, seed=0, site=0,
         aux=None, ldaux=0, aux_scale=1.0, accumulate=False, splitk=1, colsum=None):
    """C[M,N] = epilogue(opA(A)[M,K] @ opB(B)[K,N]); leading dimensions in elements (row-major storage)."""
    assert aux is None or aux.dtype == torch.float32, "the FFMA kernel takes an fp32 aux matrix (bit masks: tensor path only)"
    lib.call("msx_gemm_f32", P(A), _i(lda), _i(transA), P(B), _i(ldb), _i(transB), P(Cm), _i(ldc), _i(M), _i(N),
             _i(K), P(bias), _i(1 if relu else 0), _f(drop_p), _u64(seed), _u32(site), P(aux), _i(ldaux),
             _f(aux_scale), _i(1 if accumulate else 0), _i(splitk), P(colsum), lib.stream_ptr(),
             tag=_gemm_tag(M, N, K, accumulate, aux)[:2] + ("ffma M=%d N=%d K=%d tA=%d tB=%d sk=%d" % (M, N, K, transA, transB, splitk),))


def _aux_kind(aux):
    """0 fp32 matrix, 1 bfloat16 matrix, 2 ReLU bit mask (int32 words, msx_gemm_tc_ex)."""
    if aux is None or aux.dtype == torch.float32:
        return 0
    if aux.dtype == torch.bfloat16:
        return 1
    assert aux.dtype == torch.int32, aux.dtype
    return 2


def _gemm_tc_ex(ab16, A, lda, transA, B, ldb, transB, Cm, ldc, M, N, K, bias, relu, drop_p, seed, site, aux, ldaux, aux_scale,
                accumulate, splitk, out_colsum, mask_out, ldmask):
    c16 = Cm.dtype == torch.bfloat16
    ak = _aux_kind(aux)
    eb = 2 if ab16 else 4
    flops = 2.0 * M * N * K
    aux_bytes = 0 if aux is None else (M * N / 8.0 if ak == 2 else M * N * (2 if ak == 1 else 4))
    nbytes = eb * (M * K + N * K) + M * N * (2 if c16 else 4) * (1 + (1 if accumulate else 0)) + aux_bytes + \
        (M * N / 8.0 if mask_out is not None else 0)
    lib.call("msx_gemm_tc_ex", P(A), _i(lda), _i(transA), P(B), _i(ldb), _i(transB), P(Cm), _i(ldc), _i(M), _i(N), _i(K),
             _i(1 if ab16 else 0), _i(1 if c16 else 0), P(bias), _i(1 if relu else 0), _f(drop_p), _u64(seed), _u32(site),
             P(aux), _i(ldaux), _i(ak), _f(aux_scale), _i(1 if accumulate else 0), _i(splitk), P(out_colsum), P(mask_out),
             _i(ldmask), lib.stream_ptr(),
             tag=(flops, nbytes, "%s M=%d N=%d K=%d tA=%d tB=%d sk=%d%s%s%s%s" % (
                 "bf16" if ab16 else "tf32", M, N, K, transA, transB, splitk, " acc" if accumulate else "",
                 (" aux%d" % ak) if aux is not None else "", " c16" if c16 else "", " mask" if mask_out is not None else "")))


def gemm_tc_b3_supported(A, lda, B, ldb, Cm, ldc, M, N, K):
    return bool(lib.load().msx_gemm_tc_b3_supported(P(A), _i(lda), P(B), _i(ldb), P(Cm), _i(ldc), _i(M), _i(N), _i(K)))


def gemm_tc_b3(A, lda, B, ldb, Cm, ldc, M, N, K, bias=None, relu=False, drop_p=0.0, seed=0, site=0, accumulate=False,
               mask_out=None, ldmask=0):
    """Forward Dense GEMM Cm = A[M,K] @ B[N,K]^T with bf16x3 products (msx_gemm_tc_b3): fp32 in HBM, operands split into
    bf16 hi + lo inside the kernel, three kind::f16 MMAs per k-step."""
    nbytes = 4.0 * (M * K + N * K) + M * N * 4.0 * (1 + (1 if accumulate else 0)) + (M * N / 8.0 if mask_out is not None else 0)
    lib.call("msx_gemm_tc_b3", P(A), _i(lda), P(B), _i(ldb), P(Cm), _i(ldc), _i(M), _i(N), _i(K), P(bias),
             _i(1 if relu else 0), _f(drop_p), _u64(seed), _u32(site), _i(1 if accumulate else 0), P(mask_out), _i(ldmask),
             lib.stream_ptr(),
             tag=(2.0 * M * N * K, nbytes, "bf16x3 M=%d N=%d K=%d tA=0 tB=1 sk=1%s%s" % (
                 M, N, K, " acc" if accumulate else "", " mask" if mask_out is not None else "")))


def gemm_tc_p3_supported(A_hi, lda, B_hi, ldb, Cm, ldc, M, N, K):
    return bool(lib.load().msx_gemm_tc_p3_supported(P(A_hi), _i(lda), P(B_hi), _i(ldb), P(Cm), _i(ldc),
                                                    _i(2 if Cm.dtype == torch.bfloat16 else 0), _i(M), _i(N), _i(K)))


def gemm_tc_p3(A_hi, A_lo, lda, B_hi, B_lo, ldb, Cm, ldc, M, N, K, bias=None, relu=False, drop_p=0.0, seed=0, site=0,
               accumulate=False, mask_out=None, ldmask=0, C_lo=None):
    """Forward Dense GEMM Cm = A[M,K] @ B[N,K]^T from bf16 hi / lo planes of both operands (msx_gemm_tc_p3: three walks
    hi*hi + hi*lo + lo*hi on kind::f16).  Cm fp32, or bfloat16 together with C_lo: the result leaves as planes."""
    assert A_hi.dtype == A_lo.dtype == B_hi.dtype == B_lo.dtype == torch.bfloat16
    planes = C_lo is not None
    assert planes == (Cm.dtype == torch.bfloat16)
    nbytes = 4.0 * (M * K + N * K) + M * N * 4.0 * (1 + (1 if accumulate else 0)) + (M * N / 8.0 if mask_out is not None else 0)
    lib.call("msx_gemm_tc_p3", P(A_hi), P(A_lo), _i(lda), P(B_hi), P(B_lo), _i(ldb), P(Cm), P(C_lo), _i(ldc),
             _i(2 if planes else 0), _i(M), _i(N), _i(K), P(bias), _i(1 if relu else 0), _f(drop_p), _u64(seed), _u32(site),
             _i(1 if accumulate else 0), P(mask_out), _i(ldmask), lib.stream_ptr(),
             tag=(2.0 * M * N * K, nbytes, "bf16p3 M=%d N=%d K=%d tA=0 tB=1 sk=1%s%s%s" % (
                 M, N, K, " acc" if accumulate else "", " mask" if mask_out is not None else "", " cplanes" if planes else "")))


def split_planes(src, hi, lo, n=None):
    """hi = rn_bf16(src), lo = rn_bf16(src - hi) (bfloat16 planes of an fp32 tensor, n % 4 == 0)."""
    assert src.dtype == torch.float32 and hi.dtype == lo.dtype == torch.bfloat16
    lib.call("msx_split_planes", P(src), P(hi), P(lo), _ll(src.numel() if n is None else n), lib.stream_ptr())


def gemm_tc_x3_supported(A, lda, B, ldb, Cm, ldc, M, N, K):
    return bool(lib.load().msx_gemm_tc_x3_supported(P(A), _i(lda), P(B), _i(ldb), P(Cm), _i(ldc), _i(M), _i(N), _i(K)))


def gemm_tc(A, lda, transA, B, ldb, transB, Cm, ldc, M, N, K, bias=None, relu=False, drop_p=0.0, seed=0, site=0,
            aux=None, ldaux=0, aux_scale=1.0, accumulate=False, splitk=1, out_colsum=None, mask_out=None, ldmask=0, x3=False):
    """Same contract as gemm() on the tcgen05 tensor cores (TF32 operands, fp32 accumulate).  aux: fp32 matrix or an int32
    ReLU bit mask (ldaux in words); mask_out: optional int32 [M, ldmask] bit mask of (C > 0), N % 32 == 0.
    x3: 3xTF32 operand splitting (msx_gemm_tc_x3): fp32-equivalent products at three MMAs per k-block."""
    if x3:
        ak = _aux_kind(aux)
        aux_bytes = 0 if aux is None else (M * N / 8.0 if ak == 2 else M * N * 4)
        nbytes = 4.0 * (M * K + N * K) + M * N * 4.0 * (1 + (1 if accumulate else 0)) + aux_bytes + \
            (M * N / 8.0 if mask_out is not None else 0)
        lib.call("msx_gemm_tc_x3", P(A), _i(lda), _i(transA), P(B), _i(ldb), _i(transB), P(Cm), _i(ldc), _i(M), _i(N), _i(K),
                 P(bias), _i(1 if relu else 0), _f(drop_p), _u64(seed), _u32(site), P(aux), _i(ldaux), _i(ak), _f(aux_scale),
                 _i(1 if accumulate else 0), _i(splitk), P(out_colsum), P(mask_out), _i(ldmask), lib.stream_ptr(),
                 tag=(2.0 * M * N * K, nbytes, "tf32x3 M=%d N=%d K=%d tA=%d tB=%d sk=%d%s%s%s" % (
                     M, N, K, transA, transB, splitk, " acc" if accumulate else "", (" aux%d" % ak) if aux is not None else "",
                     " mask" if mask_out is not None else "")))
        return
    _gemm_tc_ex(False, A, lda, transA, B, ldb, transB, Cm, ldc, M, N, K, bias, relu, drop_p, seed, site, aux, ldaux,
                aux_scale, accumulate, splitk, out_colsum, mask_out, ldmask)


def gemm_tc_bf16(A, lda, transA, B, ldb, transB, Cm, ldc, M, N, K, bias=None, relu=False, drop_p=0.0, seed=0, site=0,
                 aux=None, ldaux=0, aux_scale=1.0, accumulate=False, splitk=1, out_colsum=None, mask_out=None, ldmask=0):
    """bf16 variant: A, B torch.bfloat16; Cm fp32 or bfloat16 (plain stores only); aux fp32 / bfloat16 / int32 bit mask."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16, (A.dtype, B.dtype)
    _gemm_tc_ex(True, A, lda, transA, B, ldb, transB, Cm, ldc, M, N, K, bias, relu, drop_p, seed, site, aux, ldaux,
                aux_scale, accumulate, splitk, out_colsum, mask_out, ldmask)


def gemm_tc_bf16_supported(A, lda, B, ldb, Cm, ldc, M, N, K):
    return bool(lib.load().msx_gemm_tc_bf16_supported(P(A), _i(lda), P(B), _i(ldb), P(Cm), _i(ldc),
                                                      _i(1 if Cm.dtype == torch.bfloat16 else 0), _i(M), _i(N), _i(K)))


def cast_bf16(src, dst, n=None):
    """dst (bfloat16) = src (fp32), round to nearest even."""
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16
    lib.call("msx_cast_f32_bf16", P(src), P(dst), _ll(src.numel() if n is None else n), lib.stream_ptr())


def gemm_tc_supported(A, lda, B, ldb, Cm, ldc, M, N, K):
    return bool(lib.load().msx_gemm_tc_supported(P(A), _i(lda), P(B), _i(ldb), P(Cm), _i(ldc), _i(M), _i(N), _i(K)))


def gemm_tc_set_pair(enable):
    """Selects the 2-CTA (cta_group::2) tensor GEMM kernel (default) or forces the 1-CTA one; returns the previous setting."""
    return int(lib.load().msx_gemm_tc_set_pair(_i(1 if enable else 0)))


def set_step_counter(counter):
    """Registers (tensor) / clears (None) the device-side step counter that dropout / eps seeds add (CUDA-graph replay)."""
    lib.check(lib.load().msx_set_step_counter(P(counter)), "msx_set_step_counter")


def set_pdl(enable):
    """Programmatic dependent launch of the library's kernels on / off; returns the previous setting."""
    prev = int(lib.load().msx_get_pdl())
    lib.load().msx_set_pdl(_i(1 if enable else 0))
    return prev


def step_counter_tick(counter):
    lib.call("msx_step_counter_tick", P(counter), lib.stream_ptr())


def dropout_mask(out, drop_p, seed, site):
    """out (uint8, flat view of the site's activation matrix) = keep mask the step's kernels draw for (seed, site)."""
    assert out.dtype == torch.uint8 and out.is_cuda
    lib.call("msx_dropout_mask", P(out), _ll(out.numel()), _f(drop_p), _u64(seed), _u32(site), lib.stream_ptr())


def prefix_labels(labels, out, B, T):
    lib.call("msx_prefix_labels", P(labels), P(out), _i(B), _i(T), lib.stream_ptr())


def rows_strided(src, ld_src, dst, ld_dst, rows, width, add=False):
    """dst[r, :width] (+)= src[r, :width] with row strides ld_src / ld_dst (elements)."""
    lib.call("msx_rows_strided", P(src), _ll(ld_src), P(dst), _ll(ld_dst), _i(rows), _i(width), _i(1 if add else 0),
             lib.stream_ptr())


def dropout(x, out, drop_p, seed, site):
    """out = x * keep / (1 - p) with the kernels' counter-based mask for (seed, site); in place when out is x."""
    lib.call("msx_dropout", P(x), P(out), _ll(x.numel()), _f(drop_p), _u64(seed), _u32(site), lib.stream_ptr())


def colsum(X, ld, M, N, out):
    lib.call("msx_colsum", P(X), _i(ld), _ll(M), _i(N), P(out), lib.stream_ptr())


def wgrad_splitk(M_out, N_out, K_red, sms=148):
    tiles = ((M_out + 127) // 128) * ((N_out + 127) // 128)
    want = max(1, (3 * sms) // max(tiles, 1))
    return max(1, min(want, (K_red + 511) // 512, 128))


def attention_fwd(qkv, mask, ctx, B, T, H, dh):
    lib.call("msx_attention_fwd", P(qkv), P(mask), P(ctx), _i(B), _i(T), _i(H), _i(dh), lib.stream_ptr())


def attention_tiled_fwd(qkv, mask, ctx, B, T, H, dh):
    lib.call("msx_attention_tiled_fwd", P(qkv), P(mask), P(ctx), _i(B), _i(T), _i(H), _i(dh), lib.stream_ptr())


def attention_tiled_bwd(qkv, mask, dctx, dqkv, B, T, H, dh):
    lib.call("msx_attention_tiled_bwd", P(qkv), P(mask), P(dctx), P(dqkv), _i(B), _i(T), _i(H), _i(dh), lib.stream_ptr())


def attention_tc_supported(qkv, T, dh):
    return bool(lib.load().msx_attention_tc_supported(P(qkv), _i(T), _i(dh)))


def attention_tc_fwd(qkv, mask, ctx, B, T, H, dh, x3_scores=False, ctx_lo=None, q0_only=False):
    """ctx: fp32, or bfloat16 (bf16 variant: the context only feeds the W_proj GEMMs).  x3_scores: S = K Q^T with 3xTF32
    operand splitting (fp32-equivalent scores; the softmax turns their absolute error into a relative error of P).
    ctx_lo: bfloat16 lo plane next to a bfloat16 ctx (= hi plane), the operands of the p3 W_proj GEMM.
    q0_only: compute / write the context row of query 0 of every sequence only (d_h == 32)."""
    lib.call("msx_attention_tc_fwd_p", P(qkv), P(mask), P(ctx), P(ctx_lo), _i(1 if ctx.dtype == torch.bfloat16 else 0),
             _i(1 if x3_scores else 0), _i(1 if q0_only else 0), _i(B), _i(T), _i(H), _i(dh), lib.stream_ptr())


def attention_tc_bwd(qkv, mask, dctx, dqkv, B, T, H, dh, dbias=None, q0_only=False):
    """dqkv: fp32, or bfloat16 (bf16 variant: it only feeds the K|Q|V dgrad / wgrad GEMMs).
    q0_only: the caller guarantees dctx == 0 outside the row of query 0 of every sequence."""
    lib.call("msx_attention_tc_bwd_q0", P(qkv), P(mask), P(dctx), P(dqkv), _i(1 if dqkv.dtype == torch.bfloat16 else 0),
             P(dbias), _i(1 if q0_only else 0), _i(B), _i(T), _i(H), _i(dh), lib.stream_ptr())


def attention_tcl_supported(qkv, T, dh):
    """tcgen05 attention for long rows (128 < T <= 768)."""
    return bool(lib.load().msx_attention_tcl_supported(P(qkv), _i(T), _i(dh)))


def attention_tcl_fwd(qkv, mask, ctx, stats, B, T, H, dh, q0_only=False, ctx_lo=None):
    """q0_only: compute / write the context row of query 0 of every sequence only (the other rows of ctx stay untouched).
    ctx_lo: bfloat16 lo plane next to a bfloat16 ctx (= hi plane), the operands of the p3 W_proj GEMM."""
    lib.call("msx_attention_tcl_fwd_p", P(qkv), P(mask), P(ctx), P(ctx_lo), _i(1 if ctx.dtype == torch.bfloat16 else 0), P(stats),
             _i(1 if q0_only else 0), _i(B), _i(T), _i(H), _i(dh), lib.stream_ptr())


def attention_tcl_bwd(qkv, mask, dctx, stats, dqkv, B, T, H, dh, dbias=None, q0_only=False):
    """q0_only: the caller guarantees dctx == 0 outside the row of query 0 of every sequence."""
    lib.call("msx_attention_tcl_bwd_q0", P(qkv), P(mask), P(dctx), P(stats), P(dqkv),
             _i(1 if dqkv.dtype == torch.bfloat16 else 0), P(dbias), _i(1 if q0_only else 0), _i(B), _i(T), _i(H), _i(dh),
             lib.stream_ptr())


def attention_bwd(qkv, mask, dctx, dqkv, B, T, H, dh):
    lib.call("msx_attention_bwd", P(qkv), P(mask), P(dctx), P(dqkv), _i(B), _i(T), _i(H), _i(dh), lib.stream_ptr())


def add_ln_fwd(x, y, gamma, beta, out, mean, rstd, M, D, eps=1e-5, drop_p=0.0, seed=0, site=0, out16=None, out16lo=None):
    """out16: optional bfloat16 copy of the output (operand of the bf16 GEMMs); with out16lo: its hi / lo planes."""
    lib.call("msx_add_ln_fwd_p", P(x), P(y), _i(1 if y.dtype == torch.bfloat16 else 0), P(gamma), P(beta), P(out), P(out16),
             P(out16lo), P(mean), P(rstd), _ll(M), _i(D), _f(eps),
             _f(drop_p), _u64(seed), _u32(site), lib.stream_ptr())


def add_ln_bwd(x, y, gamma, mean, rstd, dout, dres, dy, dgamma, dbeta, M, D, drop_p=0.0, seed=0, site=0,
               accumulate_dres=False, fuse_xy=False, dybias=None, dy16=None):
    """dy16: optional bfloat16 copy of the y-gradient (of the combined gradient under fuse_xy)."""
    lib.call("msx_add_ln_bwd_ex", P(x), P(y), _i(1 if y.dtype == torch.bfloat16 else 0), P(gamma), P(mean), P(rstd), P(dout),
             P(dres), P(dy), P(dy16), P(dgamma),
             P(dbeta), P(dybias), _ll(M), _i(D), _f(drop_p), _u64(seed), _u32(site), _i(1 if accumulate_dres else 0),
             _i(1 if fuse_xy else 0), lib.stream_ptr())


def embed_fwd(tokens, classes, seq_lens, tok_emb, cls_emb, prefix_vec, pe, out, mask, B, T, D, prefix, scale, vocab,
              out16=None, out16lo=None):
    lib.call("msx_embed_fwd_p", P(tokens), P(classes), P(seq_lens), P(tok_emb), P(cls_emb), P(prefix_vec), P(pe),
             P(out), P(out16), P(out16lo), P(mask), _i(B), _i(T), _i(D), _i(prefix), _f(scale), _i(vocab), lib.stream_ptr())


def embed_bwd(tokens, classes, dout, d_tok_emb, d_cls_emb, d_prefix, B, T, D, prefix, scale, vocab):
    ncls = int(d_cls_emb.shape[0]) if d_cls_emb is not None else 0
    lib.call("msx_embed_bwd_ex", P(tokens), P(classes), P(dout), P(d_tok_emb), P(d_cls_emb), P(d_prefix), _i(B), _i(T),
             _i(D), _i(prefix), _f(scale), _i(vocab), _i(ncls), lib.stream_ptr())


def rows_from_tables(tokens, classes, tab, postab, out, B, T, N, C, V, scale):
    """out[b*T + t] = scale * (tab[C + tokens[b, t]] + tab[classes[b]]) + postab[t] (rows of N floats)."""
    lib.call("msx_rows_from_tables", P(tokens), P(classes), P(tab), P(postab), P(out), _i(B), _i(T), _i(N), _i(C), _i(V),
             _f(scale), lib.stream_ptr())


def token_sort(tokens, V, perm, sorted_tok, workspace, M=None):
    """Counting sort of the row indices 0..M-1 by tokens[r] (int32 tensors; workspace int32 [3 V])."""
    lib.call("msx_token_sort", P(tokens), _ll(tokens.numel() if M is None else M), _i(V), P(perm), P(sorted_tok), P(workspace),
             lib.stream_ptr())


def rows_sum_by_token(X, ld, D, perm, sorted_tok, out, M=None, scale=1.0):
    """out [V, D] += scale * per-token sums of the rows of X [M, ld] (perm / sorted_tok from token_sort)."""
    lib.call("msx_rows_sum_by_token", P(X), _i(ld), _i(D), P(perm), P(sorted_tok), _ll(perm.numel() if M is None else M),
             _f(scale), P(out), lib.stream_ptr())


def roll_features(roll, renc, rdec, B, S):
    assert roll.dtype == torch.uint8 and roll.is_contiguous()
    lib.call("msx_roll_features", P(roll), P(_chk(renc)), P(_chk(rdec)), _i(B), _i(S), lib.stream_ptr())


def embed_dense_fwd(E, classes, cls_emb, pe, out, B, T, D, scale):
    lib.call("msx_embed_dense_fwd", P(E), P(classes), P(cls_emb), P(pe), P(out), _i(B), _i(T), _i(D), _f(scale), lib.stream_ptr())


def embed_dense_bwd(dout, classes, dE, d_cls_emb, B, T, D, scale):
    lib.call("msx_embed_dense_bwd", P(dout), P(classes), P(dE), P(d_cls_emb), _i(B), _i(T), _i(D), _f(scale), lib.stream_ptr())


def reparam_kl_fwd(lat, eps, z, kl, B, Z):
    lib.call("msx_reparam_kl_fwd", P(lat), P(eps), P(z), P(kl), _i(B), _i(Z), lib.stream_ptr())


def loss_sums(ce, kl, kl_weight, sums):
    """sums (fp32 [3]) += [sum kl, sum (ce + kl_weight * kl), B]."""
    lib.call("msx_loss_sums", P(ce), P(kl), _f(kl_weight), P(_chk(sums)), _i(kl.numel()), lib.stream_ptr())


def reparam_kl_bwd(lat, eps, dz, gkl, kl_weight, dlat, B, Z):
    lib.call("msx_reparam_kl_bwd", P(lat), P(eps), P(dz), P(gkl), _f(kl_weight), P(dlat), _i(B), _i(Z),
             lib.stream_ptr())


def normal_fill(out, seed, offset):
    lib.call("msx_normal_fill", P(out), _ll(out.numel()), _u64(seed), _u64(offset), lib.stream_ptr())


def ce_fwd(logits, ld, labels, ce, lse, metrics, B, T, V, denom, top_k=5):
    lib.call("msx_ce_fwd", P(logits), _i(ld), P(labels), P(ce), P(lse), P(metrics), _i(B), _i(T), _i(V), _i(denom),
             _i(top_k), lib.stream_ptr())


def ce_fwd_bwd(logits, ld, labels, ce, lse, metrics, B, T, V, denom, dbias=None, top_k=5):
    """Training path: msx_ce_fwd + msx_ce_bwd (head gradient 1) in one pass; the logits are overwritten with the gradient."""
    lib.call("msx_ce_fwd_bwd", P(logits), _i(ld), P(labels), P(ce), P(lse), P(metrics), _i(B), _i(T), _i(V), _i(denom),
             _i(top_k), P(dbias), lib.stream_ptr())


def ce_fwd_bwd_supported(logits, ld, V):
    return V <= 512 and ld % 4 == 0 and logits.data_ptr() % 16 == 0


def ce_bwd(logits, ld, labels, lse, gout, B, T, V, denom, dbias=None):
    lib.call("msx_ce_bwd", P(logits), _i(ld), P(labels), P(lse), P(gout), _i(B), _i(T), _i(V), _i(denom), P(dbias),
             lib.stream_ptr())


def softmax_rows(logits, ld, probs, rows, V):
    lib.call("msx_softmax_rows", P(logits), _i(ld), P(probs), _ll(rows), _i(V), lib.stream_ptr())


def ce_from_probs(probs, labels, ce, B, T, V):
    lib.call("msx_ce_from_probs", P(probs), P(labels), P(ce), _i(B), _i(T), _i(V), lib.stream_ptr())


def bce(pred, label, out, gout, dpred, B, n, from_sigmoid=False, label_smoothing=0.0, downweight=True):
    lib.call("msx_bce", P(pred), P(label), P(out), P(gout), P(dpred), _i(B), _i(n), _i(1 if from_sigmoid else 0),
             _f(label_smoothing), _i(1 if downweight else 0), lib.stream_ptr())


def lstm_fwd(gx, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, H):
    lib.call("msx_lstm_fwd", P(gx), P(w_h2h), P(b_h2h), P(h0), P(c0), _i(ld0), P(hs), P(hprev), P(cs), _i(B), _i(T),
             _i(H), lib.stream_ptr())


def lstm_bwd(gates, w_h2h, cs, c0, ld0, dhs, dh0, dc0, B, T, H, db_i2h=None, db_h2h=None):
    lib.call("msx_lstm_bwd", P(gates), P(w_h2h), P(cs), P(c0), _i(ld0), P(dhs), P(dh0), P(dc0), P(db_i2h), P(db_h2h),
             _i(B), _i(T), _i(H), lib.stream_ptr())


def lstm_tc_supported(H, ld0, h0, c0):
    return bool(lib.load().msx_lstm_tc_supported(_i(H), _i(ld0), P(h0), P(c0)))


def lstm_tc_fwd(gx, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, H):
    lib.call("msx_lstm_tc_fwd", P(gx), P(w_h2h), P(b_h2h), P(h0), P(c0), _i(ld0), P(hs), P(hprev), P(cs), _i(B), _i(T),
             _i(H), lib.stream_ptr())


def lstm_tc_fwd_tab(gates, tokens, table, w_h2h, b_h2h, h0, c0, ld0, hs, hprev, cs, B, T, H):
    """lstm_tc_fwd whose input pre-activations are rows tokens[b, t] of table [V, 4H] (= emb W_i2h^T + b_i2h)."""
    lib.call("msx_lstm_tc_fwd_tab", P(gates), P(tokens), P(table), P(w_h2h), P(b_h2h), P(h0), P(c0), _i(ld0), P(hs), P(hprev),
             P(cs), _i(B), _i(T), _i(H), lib.stream_ptr())


def lstm_tc_bwd(gates, w_h2h, cs, c0, ld0, dhs, dh0, dc0, B, T, H, db_i2h=None, db_h2h=None):
    lib.call("msx_lstm_tc_bwd", P(gates), P(w_h2h), P(cs), P(c0), _i(ld0), P(dhs), P(dh0), P(dc0), P(db_i2h), P(db_h2h),
             _i(B), _i(T), _i(H), lib.stream_ptr())


def adam_step(w, g, m, v, n, state, lr, beta1, beta2, eps, wd, rescale, clip, zero_grad=True):
    lib.call("msx_adam_step", P(w), P(g), P(m), P(v), _ll(n), P(state), _f(lr), _f(beta1), _f(beta2), _f(eps), _f(wd),
             _f(rescale), _f(clip if clip is not None else 0.0), _i(1 if zero_grad else 0), lib.stream_ptr())


def adam_nvlink_flag_bytes():
    return int(lib.load().msx_adam_nvlink_flag_bytes())


def adam_nvlink_step(w, g, m, v, n, state, peer_g, peer_w, peer_flags, done_counter, rank, world, epoch_counter, lr, beta1, beta2,
                     eps, wd, rescale, clip, zero_grad=True, max_ctas=0):
    """Fused reduce-scatter + Adam + all-gather over peer memory.  peer_* are sequences of `world` device addresses (ints);
    entry `rank` must be the local arena / flag block."""
    arr = lambda ptrs: (C.c_void_p * world)(*[C.c_void_p(int(x)) for x in ptrs])
    lib.call("msx_adam_nvlink_step", P(w), P(g), P(m), P(v), _ll(n), P(state), arr(peer_g), arr(peer_w), arr(peer_flags),
             P(done_counter), _i(rank), _i(world), P(epoch_counter), _f(lr), _f(beta1), _f(beta2), _f(eps), _f(wd), _f(rescale),
             _f(clip if clip is not None else 0.0), _i(1 if zero_grad else 0), _i(max_ctas), lib.stream_ptr())


def sample_multinomial(logits, ld, V, uniforms, seed, step, nxt, score, out_seq, out_ld, out_col, B):
    lib.call("msx_sample_multinomial", P(logits), _i(ld), _i(V), P(uniforms), _u64(seed), _u64(step), P(nxt), P(score),
             P(out_seq), _i(out_ld), _i(out_col), _i(B), lib.stream_ptr())


def beam_step(logits, ld, V, B, beam, seq_in, seq_out, seq_ld, step, score_in, score_out, parent, next_tok, unfinished=None):
    lib.call("msx_beam_step", P(logits), _i(ld), _i(V), _i(B), _i(beam), P(seq_in), P(seq_out), _i(seq_ld), _i(step),
             P(score_in), P(score_out), P(parent), P(next_tok), P(unfinished), lib.stream_ptr())


def gather_rows(src, dst, parent, rows, width):
    lib.call("msx_gather_rows", P(src), P(dst), P(parent), _i(rows), _i(width), lib.stream_ptr())
