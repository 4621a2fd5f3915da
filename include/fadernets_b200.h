/*
 * fadernets_b200 -- C ABI of the B200 (sm_100a) GM-VAE / VAE hot path.
 *
 * The reference (gudgud96/music-fader-nets) has NO native / FFI interface: the path lives in
 * Python objects.  This header is therefore the boundary a maintainer
 * would bind from the reference's own files -- every entry point cites the reference lines
 * (paths relative to the reference checkout) whose arithmetic it replaces.  INTEGRATION.md
 * shows the ctypes stubs.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is DEVICE memory unless marked "host";
 *   - the CALLER owns every buffer (including scratch); the library allocates nothing that
 *     outlives a call and keeps no mutable global state besides cached device attributes;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the host;
 *   - return value: 0 = OK, negative = FN_ERR_*; fn_last_error() gives the message;
 *   - activations inside the path are TIME-MAJOR: [T][B][...]; API-facing tensors keep the
 *     reference's batch-major layout ([B][T][...]) and are converted by the kernels that
 *     produce / consume them;
 *   - GRU gate order is PyTorch's (r, z, n); weights keep the reference's state_dict layout.
 */
#ifndef FADERNETS_B200_H
#define FADERNETS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FN_OK 0
#define FN_ERR_ARG (-1)      /* bad argument / unsupported shape */
#define FN_ERR_CUDA (-2)     /* a CUDA runtime call failed        */
#define FN_ERR_UNSUPPORTED (-3)

#define FN_ABI_VERSION 3      /* 3: + bf16x3 entry points (fn_gru_seq_*_bf16x3, fn_tc_gemm_bf16x3, fn_split_bf16, fn_time_sum_bf16x3), fn_gemm_f32_splitk; additive */

const char* fn_last_error(void);
int fn_abi_version(void);
/* sha256 prefix of the sources this binary was built from (__graft_entry__.source_hash()); the Python binding refuses a
 * library whose hash differs from the sources it ships with. */
const char* fn_source_hash(void);
/* sm count / compute capability of the current device (host out-params). */
int fn_device_info(int* sm_count, int* cc_major, int* cc_minor, int* max_smem_optin);

/* ------------------------------------------------------------------------------------------
 * Dense fp32 GEMM with arbitrary operand strides (SIMT FMA path -- exact fp32 parity mode).
 *   C[m,n] (ldc) = (accumulate ? C[m,n] : 0) + sum_k A(m,k) * B(k,n) + (bias ? bias[n] : 0)
 *   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn];  one of each stride pair must be 1.
 * Replaces every nn.Linear / addmm on the path: gmm_model.py:86,91 (latent heads), :107,:112
 * (linear_init_*), :110,:115 (linear_out_r/n), :124 (linear_init_global), :137 (linear_out_g),
 * the W_ih * x products inside nn.GRU / nn.GRUCell (:33-56) and their autograd transposes.
 * ---------------------------------------------------------------------------------------- */
int fn_gemm_f32(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                float* C, long long ldc, const float* bias, int M, int N, int K, int accumulate, void* stream);

/* Split-K variant for products with a long K and few output tiles (latent heads [B x 2H] x [2H x Z], the z-projection
 * gradients [B x 3H] x [3H x G]): `splits` CTAs per 64x64 tile each reduce a K range into `workspace`
 * (fn_gemm_f32_splitk_ws_bytes), then a fixed-order reduction applies bias / accumulate -- deterministic. */
size_t fn_gemm_f32_splitk_ws_bytes(int M, int N, int splits);
int fn_gemm_f32_splitk(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                       float* C, long long ldc, const float* bias, int M, int N, int K, int accumulate, int splits,
                       void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * bf16 tensor-core GEMM (tcgen05.mma, fp32 accumulation in TMEM, TMA-fed 128B-swizzled smem ring).
 *   C[M][N] (ldc; fp32 or bf16) = (accumulate ? C : 0) + A * B + (bias ? bias[n] : 0)
 *   a_mn_major = 0: A is stored [M][K] (lda);  1: A is stored transposed, [K][M] (lda)
 *   b_mn_major = 0: B is stored [N][K] (ldb) (nn.Linear weight layout);  1: stored [K][N] (ldb)
 * lda/ldb are in elements and must be multiples of 8 (16-byte TMA pitch); bases 16-byte aligned.
 * Same call sites as fn_gemm_f32, used by the bf16 configurations (BASELINE configs 3-5).
 * ---------------------------------------------------------------------------------------- */
int fn_tc_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                    void* C, long long ldc, int c_bf16, const float* bias, int M, int N, int K, int accumulate,
                    void* stream);

/* Split-K variant for products with few output tiles and a very long K (the T*B-row weight gradients):
 * `splits` CTAs per output tile each reduce a K range into `workspace` (fn_tc_gemm_splitk_ws_bytes), then a
 * fixed-order reduction applies bias / accumulate -- deterministic, no float atomics. */
size_t fn_tc_gemm_splitk_ws_bytes(int M, int N, int splits);
int fn_tc_gemm_bf16_splitk(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                           void* C, long long ldc, int c_bf16, const float* bias, int M, int N, int K, int accumulate,
                           int splits, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Persistent time-loop GRU ("gate block": recurrent GEMM + sigma/tanh/Hadamard per step).
 * One launch runs `n_chains` independent recurrences; each chain is split over hidden-unit
 * slices (one CTA per slice, W_hh slice resident in shared memory for all T steps) with a
 * per-chain step barrier in `barrier_ws` (>= 64 * n_chains bytes, zeroed by the call).
 *
 * Step s processes time tau = reverse ? T-1-s : s:
 *   gi = emb[ids[tau][b]] + proj[b] + dense[tau][b]        (each term optional)
 *   gh = h_prev W_hh^T + b_hh
 *   r = sigma(gi_r+gh_r); z = sigma(gi_z+gh_z); n = tanh(gi_n + r*gh_n); h = (1-z)*n + z*h_prev
 * (torch nn.GRU / nn.GRUCell semantics; reference call sites gmm_model.py:84,89 (encoders),
 *  :109,:114 (sub-decoders), :133,:136 (global decoder cells, teacher-forced: layer 1 consumes
 *  the token gather, layer 2 consumes `dense` = hx0 W_ih2^T + b_ih2).)
 * ---------------------------------------------------------------------------------------- */
typedef struct FnGruChain {
    /* parameters */
    const float* w_hh;        /* [3H][H]                                                  */
    const float* b_hh;        /* [3H]                                                     */
    /* input side */
    const float* emb;         /* [Vin][3H] = W_ih[:, :Vin]^T, or NULL                     */
    const int32_t* ids;       /* [T][B] token ids (by time), or NULL                      */
    const float* proj;        /* [B][3H] (row stride proj_ld; 0 = one broadcast row) / NULL */
    long long proj_ld;
    const float* dense;       /* [T][B][3H] (by time) or NULL                             */
    const float* h0;          /* [B][H] or NULL (zeros)                                   */
    int32_t reverse;          /* 1: run time backwards (bidirectional encoder, 2nd dir)   */
    int32_t _pad0;
    /* forward outputs */
    float* hs;                /* [T][B][H]  (by time) hidden state after that time step   */
    float* gates;             /* [T][B][4H] (r, z, n, gh_n) saved for BPTT, or NULL       */
    float* h_final;           /* [B] rows of stride h_final_ld: last state, or NULL       */
    long long h_final_ld;
    /* backward inputs */
    const float* dhs;         /* [T][B][H] (by time) grad wrt hs, or NULL                 */
    const float* dh_final;    /* grad wrt h_final (row stride dh_final_ld), or NULL       */
    long long dh_final_ld;
    /* backward outputs */
    float* dgh;               /* [T][B][3H] grad wrt gh = (dr_pre, dz_pre, dn_pre*r)      */
    float* dgin;              /* [T][B][H]  grad wrt gi_n = dn_pre  (gi_r,gi_z share dgh) */
    float* dh0;               /* [B][H] grad wrt h0 (always written)                      */
    float* dh_carry;          /* [B][H] scratch                                           */
} FnGruChain;

/* `chains` is a HOST array.  All chains of one call share B, T, H. */
int fn_gru_seq_fwd_f32(const FnGruChain* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                       size_t barrier_ws_bytes, void* stream);
/* BPTT of the above (autograd of the reference's nn.GRU calls and of the Python GRUCell loop,
 * trainer_gmm.py:249 loss.backward()).  Weight/bias/embedding gradients are finished by
 * fn_gemm_f32 / fn_emb_grad_f32 / fn_time_sum_f32 on the dgh/dgin streams. */
int fn_gru_seq_bwd_f32(const FnGruChain* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                       size_t barrier_ws_bytes, void* stream);
/* How many CTAs one chain needs for hidden size H (host query; <=0 if H unsupported). */
int fn_gru_seq_ctas_per_chain(int H);

/* ------------------------------------------------------------------------------------------
 * The same gate block on the 5th-generation tensor cores (tcgen05.mma, bf16 operands, fp32
 * accumulation in TMEM, weight slice resident in shared memory, state slabs streamed by TMA).
 * Everything is indexed BY TIME; step s of a chain works on time tau = s, or tau = T-1-s for a
 * reverse chain.  hsx has T+1 slabs of [B][H]: a forward chain reads its initial state from slab 0
 * and writes h_tau to slab tau+1; a reverse chain reads slab T and writes h_tau to slab tau.  The
 * caller fills the initial-state slab.  B <= 256 rows per chain (larger batches: several chains).
 * Same reference call sites as fn_gru_seq_*_f32; used by the bf16 configurations (BASELINE 3-5).
 * ---------------------------------------------------------------------------------------- */
typedef struct FnGruChainBf16 {
    const void* w_hh;         /* bf16 [3H][H]            (forward)                            */
    const void* w_hh_t;       /* bf16 [H][3H] = W_hh^T   (backward)                           */
    const float* b_hh;        /* fp32 [3H]                                                    */
    const void* emb;          /* bf16 [Vin][3H] = W_ih[:, :Vin]^T, or NULL (input = token gather) */
    const int32_t* ids;       /* [T][B] token ids by time, or NULL                            */
    const float* proj;        /* fp32 [B][3H] (row stride proj_ld; 0 = one broadcast row) / NULL */
    long long proj_ld;
    const void* dense;        /* bf16 [T][B][3H] by time, or NULL (input = dense stream; excludes emb) */
    int32_t reverse;
    int32_t dhs_f32;          /* dtype of dhs: 0 = bf16, 1 = fp32                             */
    void* hsx;                /* bf16 [T+1][B][H] (see above)                                 */
    void* gates;              /* bf16, T * ceil32(B) * 4H elements: r, z, n, W_hn h + b_hn saved for BPTT (NULL =
                               * inference).  OPAQUE to the caller: written by fn_gru_seq_fwd_bf16, read by
                               * fn_gru_seq_bwd_bf16, stored in [32 rows][16 columns] blocks (block order: time slab,
                               * 32-row block, 16-column block) so that the gate epilogues access it coalesced.  A time
                               * segment [t0, t0+L) starts at element offset t0 * ceil32(B) * 4H.               */
    float* h_final;           /* fp32 rows of stride h_final_ld: state after the last step / NULL */
    long long h_final_ld;
    const void* dhs;          /* [T][B][H] by time: grad wrt h_tau, or NULL                   */
    const float* dh_final;    /* fp32 grad wrt h_final (row stride dh_final_ld), or NULL      */
    long long dh_final_ld;
    void* dg;                 /* bf16 [T][B][4H]: dr_pre, dz_pre, dn_pre, dn_pre*r (cols [0,3H) = grad wrt
                                 the input-side pre-activations, [0,2H) + [3H,4H) = grad wrt W_hh h + b_hh) */
    float* dh0;               /* fp32 [B][H] grad wrt the initial state                       */
} FnGruChainBf16;
int fn_gru_seq_fwd_bf16(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                        size_t barrier_ws_bytes, void* stream);
int fn_gru_seq_bwd_bf16(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                        size_t barrier_ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * bf16x3 mode: the fp32 parity bar (1e-3 relative, BASELINE config 2) ON the tensor cores.  Every fp32 value that feeds a
 * T-scale product is carried as TWO bf16 planes, hi = bf16(x) and lo = bf16(x - hi) (16 mantissa bits together), and every
 * product is evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (the dropped lo*lo term is ~2^-16 relative).
 * Same kernels, same call sites as the bf16 entry points above; what changes is the storage:
 *   hsx    bf16 [T+1][B][2H]  columns [0,H) = hi, [H,2H) = lo;
 *   gates  bf16, T * ceil32(B) * 8H elements (opaque: the 4H hi columns then the 4H lo columns of a row, 32x16 blocks);
 *   dg     bf16 [T][B][8H]    columns [0,4H) = hi of (dr, dz, dn, dn*r), [4H,8H) = lo;
 *   w_hh   bf16 [3H][3H]  = [hi | hi | lo] of W_hh along K (fn_split_bf16 with hi2_off);  w_hh_t bf16 [H][9H] likewise;
 *   emb, dense, dhs are FP32 ([Vin][3H], [T][B][3H], [T][B][H]); sigma / tanh are evaluated with exp, not tanh.approx.
 * B <= 256 rows per chain, H % 64 == 0.
 * ---------------------------------------------------------------------------------------- */
int fn_gru_seq_fwd_bf16x3(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                          size_t barrier_ws_bytes, void* stream);
int fn_gru_seq_bwd_bf16x3(const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws,
                          size_t barrier_ws_bytes, void* stream);
/* Tensor-core GEMM over hi / lo planes: C fp32|bf16 = (A_hi + A_lo)(B_hi + B_lo) - A_lo B_lo, one fp32 accumulator.
 * a_lo_off / b_lo_off: element offset of the operand's lo plane from its hi plane (same leading dimension); 0 = the
 * operand is exact in bf16 (e.g. a one-hot) and contributes one plane.  Other arguments as fn_tc_gemm_bf16_splitk. */
int fn_tc_gemm_bf16x3(const void* A, long long lda, long long a_lo_off, int a_mn_major, const void* B, long long ldb,
                      long long b_lo_off, int b_mn_major, void* C, long long ldc, int c_bf16, const float* bias, int M, int N,
                      int K, int accumulate, int splits, void* workspace, size_t ws_bytes, void* stream);
/* fp32 -> (hi, lo) bf16 planes with arbitrary source strides: dst[r*ld_dst + c] = hi, dst[r*ld_dst + lo_off + c] = lo and,
 * if hi2_off >= 0, a second copy of hi at dst[r*ld_dst + hi2_off + c] (the [hi | hi | lo] weight layout above). */
int fn_split_bf16(const float* src, long long s_r, long long s_c, void* dst, long long ld_dst, long long rows, long long cols,
                  long long lo_off, long long hi2_off, void* stream);
/* fn_time_sum_bf16 over the split stream dg [T][B][8H] (sums hi + lo). */
int fn_time_sum_bf16x3(const void* dg, int B, int T, int H, float* dproj, float* dghsum, void* stream);

/* ------------------------------------------------------------------------------------------
 * Greedy decode of the global decoder as ONE persistent kernel (eval-mode global_decoder, gmm_model.py:119-149
 * with _sampling :73-80; drivers test_class.py:233-254, arousal_transfer.ipynb cells 15/17): per step cell 1
 * (token gather + z projection) -> cell 2 -> vocabulary projection -> first arg-max -> next token, all `steps` on
 * the device.  bf16 weights: w_hh1 / w_ih2 / w_hh2 [3H][H], w_out [V][H], emb1 [V][3H] = W_ih1[:, :V]^T;
 * proj1 fp32 [B][3H] = z W_ih1[:, V:]^T + b_ih1.  hs1 / hs2: bf16 [steps+1][B][H] state slabs (hs1 slab 0 = the
 * initial state linear_init_global(z), caller-filled; slab s+1 = state after step s); gi2: bf16 [steps][B][3H]
 * scratch.  tokens: int32 [steps][B].  logits_out: fp32 [steps][B][V] (pre-softmax) or NULL.  B <= 256.
 * ---------------------------------------------------------------------------------------- */
size_t fn_decode_greedy_ws_bytes(int B, int steps, int H, int V);
int fn_decode_greedy_bf16(const void* w_hh1, const float* b_hh1, const void* emb1, const float* proj1, const void* w_ih2,
                          const float* b_ih2, const void* w_hh2, const float* b_hh2, const void* w_out, const float* b_out,
                          void* hs1, void* hs2, void* gi2, int B, int steps, int H, int V, int start_token,
                          int32_t* tokens, float* logits_out, void* workspace, size_t ws_bytes, void* stream);

/* Profiling aid: CTA 0 of the following fn_gru_seq_*_bf16 launches writes clock64 stamps of its pipeline
 * events into `device_buffer` ((T+1)*2*16 int64); NULL switches it off. */
int fn_gru_debug_timeline(void* device_buffer);

/* Element-wise pieces of the sibling models built on the same blocks: MusicAttrFaderNets' discriminator
 * heads `dropout(relu(linear(reverse(z))))` (model_v2.py:426-435, 574-575: relu x keep-mask/(1-p); the gradient reversal
 * is fn_scale_f32 with alpha = -1 in backward) and the adversarial MSE of trainer_fader.py:105-110 (mean over the batch). */
int fn_relu_mask_fwd(const float* x, const float* mask, float* y, long long n, void* stream);
int fn_relu_mask_bwd(const float* x, const float* mask, const float* dy, float* dx, long long n, void* stream);
int fn_scale_f32(const float* src, float* dst, float alpha, long long n, void* stream);
int fn_mse_mean_fwd(const float* x, const float* y, long long n, float* loss, void* stream);
int fn_mse_mean_bwd(const float* x, const float* y, long long n, const float* dloss, float* dx, void* stream);

/* Index validation.  The reference's F.nll_loss / nn.Embedding / Embedding lookups (trainer_gmm.py:131-136,156,182;
 * gmm_model.py:84,109,132) raise on an out-of-range index.  Here every kernel that consumes an index CLAMPS it (no
 * out-of-bounds access is possible); these two entry points COUNT offending entries into `bad_count` (device int32,
 * accumulated, never reset by the library) so that the host mirror raises IndexError at its next synchronisation.
 * fn_check_index_i64 only counts (user-owned targets / labels); fn_clamp_index_i32 also clamps in place (the library's own
 * time-major id buffers that feed the embedding gathers). */
int fn_check_index_i64(const int64_t* idx, long long n, long long hi, int32_t* bad_count, void* stream);
int fn_clamp_index_i32(int32_t* idx, long long n, int hi, int32_t* bad_count, void* stream);

/* fp32 -> bf16 with arbitrary strides: dst[r*ld_dst + c] = bf16(src[r*s_r + c*s_c]) (s_r/s_c in
 * elements; s_c != 1 gives the transposed copy W_hh^T used by BPTT). */
int fn_cast_bf16(const float* src, long long s_r, long long s_c, void* dst, long long ld_dst, long long rows,
                 long long cols, void* stream);
/* int32 ids [rows] -> bf16 one-hot [rows][ld] (columns >= V zero): the MN-major operand of the
 * tensor-core form of autograd's `onehot^T @ dgi` (gmm_model.py:84,109,132-133). */
int fn_ids_to_onehot_bf16(const int32_t* ids, long long rows, int V, long long ld, void* onehot, void* stream);
/* bf16 variant of fn_time_sum_f32 over dg [T][B][4H]: dproj = sum_t dg[:, :3H],
 * dghsum = sum_t [dg[:, :2H] | dg[:, 3H:]]. */
int fn_time_sum_bf16(const void* dg, int B, int T, int H, float* dproj, float* dghsum, void* stream);
/* dst[i] = bf16(float(dst[i]) + src[i])  (bf16 twin of fn_add_f32) */
int fn_add_f32_to_bf16(void* dst, const float* src, long long n, void* stream);
/* out[c] (+)= sum_r x[r][c] for a bf16 matrix (bias gradients of the tensor-core linears). */
int fn_col_sum_bf16(const void* x, long long ld, long long rows, int cols, float* out, int accumulate,
                    void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Token plumbing.
 * ---------------------------------------------------------------------------------------- */
/* one-hot (B,T,V) fp32 -> first-max index, time-major int32 [T][B]  (inverse of
 * convert_to_one_hot, trainer_gmm.py:296-303; the reference feeds the dense one-hot to nn.GRU) */
int fn_onehot_to_ids(const float* onehot, int B, int T, int V, int32_t* ids_tm, void* stream);
/* int64 ids (B,T) -> dense one-hot (B,T,V) fp32  (convert_to_one_hot, trainer_gmm.py:296-303) */
int fn_ids_to_onehot(const int64_t* ids, int B, int T, int V, float* onehot, void* stream);
/* int64 ids (B,T) -> int32 time-major [T][B]; shift=1 gives the teacher-forced decoder input
 * stream (start token at t=0, then ids[:, t-1]; gmm_model.py:120-121,139-142). */
int fn_ids_to_time_major(const int64_t* ids, int B, int T, int shift, int start_token, int32_t* ids_tm,
                         void* stream);
/* Post-processing of decoded tokens, batched: clean_output (test_class.py:44-50) keeps, per row of tokens
 * (rows, steps) int64, the span after trimming leading / trailing zeros and cutting at the first EOS (1):
 * start[row], len[row] index the original row. */
int fn_clean_tokens(const int64_t* tokens, int rows, int steps, int32_t* start, int32_t* len, void* stream);
/* dst[c][r] = src[r][c] (+= if accumulate)  -- W_ih[:, :V] <-> embedding-table layout */
int fn_transpose_f32(const float* src, long long ld_src, float* dst, long long ld_dst, int rows, int cols,
                     int accumulate, void* stream);

/* dst[i] += src[i]  (folds dh0 of the second decoder cell into the first cell's state gradient) */
int fn_add_f32(float* dst, const float* src, long long n, void* stream);

/* Embedding-table gradient: demb[v][c] = sum over (t,b) with ids[t][b]==v of dgi[t][b][c], where
 * dgi = [dgh[:, :2H] | dgin].  Deterministic two-stage segmented reduction.  `scratch` needs
 * fn_emb_grad_scratch_bytes().  (autograd of `onehot @ W_ih^T`, gmm_model.py:84,109,132-133) */
size_t fn_emb_grad_scratch_bytes(int B, int T, int H, int V);
int fn_emb_grad_f32(const int32_t* ids_tm, const float* dgh, const float* dgin, int B, int T, int H, int V,
                    float* demb, void* scratch, size_t scratch_bytes, void* stream);
/* Time sums of the BPTT gate gradients: dproj[b][c] = sum_t dgi[t][b][c] (gradient of the
 * time-invariant input-side term: z-projection + b_ih) and dghsum[b][c] = sum_t dgh[t][b][c]
 * (column-summed afterwards into the b_hh gradient).  Either output may be NULL. */
int fn_time_sum_f32(const float* dgh, const float* dgin, int B, int T, int H, float* dproj, float* dghsum,
                    void* stream);
/* out[c] (+)= sum_r x[r][c]  (bias gradients); deterministic two-stage reduction. */
size_t fn_col_sum_scratch_bytes(long long rows, int cols);
int fn_col_sum_f32(const float* x, long long ld, long long rows, int cols, float* out, int accumulate,
                   void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Soft-max heads and NLL.
 * ---------------------------------------------------------------------------------------- */
/* logits [T][B][V] (time-major) -> log-probs out (B,T,V) batch-major: F.log_softmax over the
 * vocabulary, gmm_model.py:137 + torch.stack(x,1) :149. */
int fn_vocab_logsoftmax_fwd(const float* logits_tm, int B, int T, int V, float* out_bm, void* stream);
/* dlogits_tm = dout - exp(out) * rowsum(dout) */
int fn_vocab_logsoftmax_bwd(const float* out_bm, const float* dout_bm, int B, int T, int V, float* dlogits_tm,
                            void* stream);
/* Fused train-step variant: mean NLL of the vocabulary head without materialising dout:
 *   fwd: loss_rows[t*B+b] = -logp[target]; out_bm written only if non-NULL;
 *   bwd: dlogits_tm = scale * (softmax - onehot(target))   (F.nll_loss, trainer_gmm.py:131-132) */
int fn_vocab_nll_fwd(const float* logits_tm, const int64_t* target_bm, int B, int T, int V, float* out_bm,
                     float* lse_tm, float* loss_rows, void* stream);
int fn_vocab_nll_bwd(const float* logits_tm, const float* lse_tm, const int64_t* target_bm, const float* scale_dev,
                     float scale_host, int B, int T, int V, float* dlogits_tm, void* stream);
/* Sub-decoder heads: log-softmax over the TIME axis (dim=1 of (B,T,C); gmm_model.py:110,115).
 * logits [T][B][C] -> out (B,T,C). */
int fn_time_logsoftmax_fwd(const float* logits_tm, int B, int T, int C, float* out_bm, void* stream);
int fn_time_logsoftmax_bwd(const float* out_bm, const float* dout_bm, int B, int T, int C, float* dlogits_tm,
                           void* stream);
/* F.nll_loss(logp.view(-1,C), target.view(-1), 'mean') (trainer_gmm.py:131-136): gather + mean. */
int fn_nll_mean_fwd(const float* logp, const int64_t* target, long long rows, int C, float* loss, void* scratch,
                    size_t scratch_bytes, void* stream);
/* dlogp[row][target] (+)= dloss / rows ; all other entries zero (zero_fill=1 clears first). */
int fn_nll_mean_bwd(const int64_t* target, long long rows, int C, const float* dloss, float* dlogp, int zero_fill,
                    void* stream);
size_t fn_reduce_scratch_bytes(long long n);
/* out[0] = scale * sum(x[0..n)) , deterministic */
int fn_sum_f32(const float* x, long long n, float scale, float* out, void* scratch, size_t scratch_bytes,
               void* stream);

/* ------------------------------------------------------------------------------------------
 * Latent block.
 * ---------------------------------------------------------------------------------------- */
/* Caller-owned scratch of the reductions below (two fixed-order stages: per-CTA partials, then one sum), for a
 * latent block of B rows x Z dims and K mixture components.  The reductions stream at HBM bandwidth for large B
 * and are run-to-run deterministic.  scratch == NULL selects the single-CTA kernels (small B only). */
size_t fn_latent_scratch_bytes(int B, int Z, int K);
/* scale = exp(pre)  (var_r(x).exp_(), gmm_model.py:86,91) and z = mu + scale*eps (repar, :229-235) */
int fn_reparam_fwd(const float* mu, const float* pre_scale, const float* eps, long long n, float* scale, float* z,
                   void* stream);
/* dmu = dz + dmu_in ; dpre = (dz*eps + dscale_in) * scale */
int fn_reparam_bwd(const float* dz, const float* dmu_in, const float* dscale_in, const float* eps,
                   const float* scale, long long n, float* dmu, float* dpre, void* stream);
/* approx_qy_x (gmm_model.py:194-218): logLogit[b][k] = -0.5*sum_d((z-mu_k)^2/exp(lv_k)+lv_k+ln2pi)+ln(1/K),
 * qy = softmax_k, y = first argmax. */
int fn_qy_x_fwd(const float* z, const float* mu_lookup, const float* logvar_lookup, int B, int Z, int K,
                float* logLogit, float* qy, int64_t* y, void* stream);
/* given dlogLogit, dqy: dz (B,Z), dmu_lookup (K,Z) (+= if accumulate) */
int fn_qy_x_bwd(const float* z, const float* mu_lookup, const float* logvar_lookup, const float* qy,
                const float* dlogLogit, const float* dqy, int B, int Z, int K, float* dz, float* dmu_lookup,
                void* scratch, size_t scratch_bytes, void* stream);
/* GM-VAE KL block (trainer_gmm.py:140-194).  mode 0 = unsupervised, 1 = supervised (y_label).
 * out[0] = kld_lat (sum_k mean_b(mean_z KL(q||p_k) * qy[b,k]))   [sup: mean_b mean_z KL(q||p_y)]
 * out[1] = kld_cls ((mean_k(qy*log_softmax(logLogit)) - ln(1/K)).mean())  [sup: 0]
 * out[2] = label_clf (sup only: CrossEntropyLoss applied to the probabilities qy)  [unsup: 0]
 * p_k = Normal(mu_k, scale = exp(logvar_k))  (sic: exp(logvar) used as scale, :156-157). */
int fn_gm_kl_fwd(const float* mu, const float* scale, const float* mu_lookup, const float* logvar_lookup,
                 const float* qy, const float* logLogit, const int64_t* y_label, int mode, int B, int Z, int K,
                 float* out3, void* scratch, size_t scratch_bytes, void* stream);
/* grads wrt (mu, scale, qy, logLogit, mu_lookup) given dout3 (device, 3 floats). */
int fn_gm_kl_bwd(const float* mu, const float* scale, const float* mu_lookup, const float* logvar_lookup,
                 const float* qy, const float* logLogit, const int64_t* y_label, int mode, int B, int Z, int K,
                 const float* dout3, float* dmu, float* dscale, float* dqy, float* dlogLogit, float* dmu_lookup,
                 void* scratch, size_t scratch_bytes, void* stream);
/* vanilla VAE: out[0] = mean_{B,Z} KL(N(mu,scale) || N(0,1))  (trainer.py:104-109) */
int fn_std_kl_fwd(const float* mu, const float* scale, long long n, float* out, void* scratch, size_t scratch_bytes,
                  void* stream);
int fn_std_kl_bwd(const float* mu, const float* scale, long long n, const float* dout, float* dmu, float* dscale,
                  void* stream);
/* Pati et al. latent regularisation (trainer_gmm.py:199-217): l = mean_{i,j}(tanh(z0_i - z0_j) -
 * sign(a_i - a_j))^2 on latent dim 0; attribute differences are taken in float64 as the reference's
 * host numpy does.  z0 has element stride z_ld.  Also returns dl/dz0 (B floats) for the backward. */
int fn_latent_reg_fwd(const float* z, long long z_ld, const double* attr, int B, float* loss, float* dz0,
                      float* row_scratch /* B floats */, void* stream);
/* dz (B,Z) = 0 except column 0 = dloss[0] * dz0 */
int fn_latent_reg_bwd(const float* dz0, const float* dloss, int B, int Z, float* dz, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser: clip_grad_norm_(params, max_norm) + Adam.step() on flat buffers
 * (trainer_gmm.py:250-251; torch clip_grad.py: coef = max_norm/(norm+1e-6) clamped to 1;
 *  optim.Adam defaults betas (0.9,0.999) eps 1e-8, no weight decay / amsgrad).
 * ---------------------------------------------------------------------------------------- */
/* norm_out[0] = ||g||_2 over n elements (fp32 accumulate in double). */
int fn_grad_norm(const float* g, long long n, float* norm_out, void* scratch, size_t scratch_bytes, void* stream);
/* p,m,v updated in place; g scaled by the clip coefficient computed from *norm (device). step >= 1. */
int fn_clip_adam(float* p, const float* g, float* m, float* v, long long n, const float* norm, float max_norm,
                 float lr, float beta1, float beta2, float eps, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FADERNETS_B200_H */
