/*
 * libemmax — C ABI of the B200-native `generate_actions` hot path (Emma-X / OpenVLA).
 *
 * The reference (/root/reference) has NO native code and no FFI: every FLOP of its hot path runs inside un-vendored
 * third-party packages (timm ViT, transformers Llama, flash-attn, ATen). The drop-in boundary a user sees is the Python
 * class surface (emmax_b200.modeling), and THIS header is the boundary between that Python host code and the
 * hand-written sm_100a kernels: plain pointers and sizes, no torch types, loaded with ctypes (emmax_b200/_lib.py).
 * Each entry point cites the reference call site whose third-party kernel(s) it replaces.
 *
 * Conventions: every function returns 0 on success, <0 on error (emx_last_error() gives the text); all pointers are
 * DEVICE pointers unless stated; bf16 = 16-bit brain float, row-major, `ld*` = leading dimension in elements;
 * nothing allocates, nothing synchronises; work is enqueued on `stream`.
 */
#ifndef EMMAX_H_
#define EMMAX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* emx_stream_t; /* == cudaStream_t */

const char* emx_last_error(void);
int emx_abi_version(void);
/* "sm_100a" — the only architecture the library is built for. */
const char* emx_arch(void);

/* ---- GEMM (tcgen05 / TMEM / TMA) -----------------------------------------------------------------------------
 * C[M,N] = epilogue(A[M,K] * W[N,K]^T), bf16 in/out, fp32 accumulate.
 * Replaces the cuBLASLt GEMMs behind: timm ViT Linears (modeling_prismatic.py:121), PrismaticProjector.fc1-3
 * (modeling_prismatic.py:152-156), Llama q/k/v/o/gate/up/down at prefill (modeling_prismatic.py:404-415).
 * Epilogue order (each step rounded to bf16 like the unfused reference ops):
 *   +bias -> [GELU(erf)] -> [*layerscale] -> [+resid] -> [SwiGLU over interleaved (gate,up) column pairs]
 * resid_mod > 0: residual row index is (m % resid_mod) (position-embedding broadcast over the batch). */
#define EMX_EPI_GELU 1
#define EMX_EPI_SWIGLU 2
int emx_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, const void* bias,
                  const void* layerscale, const void* resid, int ldr, int resid_mod, int flags, emx_stream_t stream);
/* Same, with a caller-owned scratch buffer that allows split-K for the small-M problems of a bs = 1 request whose output tiles do not
 * fill the machine (Llama o_proj / down_proj at M = 296, ViT proj / fc2 at M ~ 260): 16-byte aligned device memory, ZERO-INITIALISED once
 * (its first 4096 bytes are self-resetting per-tile arrival counters; the rest holds fp32 partial tiles), at least
 * 4096 + tiles * splits * 128 * 128 * 4 bytes or the call does not split; not to be shared by GEMMs that may run concurrently (one per
 * stream / graph branch). The partials are added in a fixed order: results do not depend on which CTA finishes last. */
int emx_gemm_bf16_ws(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, const void* bias,
                     const void* layerscale, const void* resid, int ldr, int resid_mod, int flags, void* workspace, long workspace_bytes,
                     emx_stream_t stream);

/* ---- normalisation ------------------------------------------------------------------------------------------
 * LayerNorm (ViT blocks, eps 1e-6; ATen vectorized_layer_norm in the reference, via timm Block.norm1/norm2) and
 * Llama RMSNorm (transformers LlamaRMSNorm: fp32 normalise -> bf16 -> * weight). */
int emx_layernorm(const void* x, const void* weight, const void* bias, void* y, int rows, int dim, float eps, emx_stream_t stream);
int emx_rmsnorm(const void* x, const void* weight, void* y, int rows, int dim, float eps, emx_stream_t stream);

/* ---- ViT front end ------------------------------------------------------------------------------------------
 * Patch-embed conv as im2col + GEMM (timm PatchEmbed.proj = Conv2d(3,D,14,14), invoked at modeling_prismatic.py:121).
 * pixels: [B, C_total, H, W] bf16; channels [chan0, chan0+3) are gathered into rows of (c,ky,kx)-ordered patches,
 * zero-padded to `kpad` columns: out [B*gh*gw, kpad]. */
/* GPU twin of PrismaticImageProcessor.apply_transform for frames that already have the model's input size
 * (processing_prismatic.py:128-145: to_tensor -> normalize per backbone -> channel stack; then `.to(device, bf16)`):
 * hwc uint8 [B, H, W, 3] (device) -> out bf16 [B, 3*n_backbones, H, W]; mean/stdv: device fp32 [n_backbones*3]. Bit-exact with the
 * host transform followed by the bf16 cast. */
int emx_preprocess_u8(const void* hwc, int B, int H, int W, int n_backbones, const float* mean, const float* stdv, void* out,
                      emx_stream_t stream);
/* The same with the `resize-naive` bicubic-antialias resize in front (torchvision resize of a PIL image == Pillow's ImagingResample,
 * processing_prismatic.py:133): two separable integer passes, bit-exact with Pillow. hwc uint8 [B, Hin, Win, 3]; kk_* int32
 * [out_size][ksize_*] 22-bit fixed-point coefficients and bounds_* int32 [out_size][2] = (first input index, count), built on the host as
 * Pillow's precompute_coeffs + normalize_coeffs_8bpc do; tmp uint8 [B, Hin, Wout, 3] scratch; out bf16 [B, 3*n_backbones, Hout, Wout]. */
int emx_resize_preprocess_u8(const void* hwc, int B, int Hin, int Win, int Hout, int Wout, const int32_t* kk_h, const int32_t* bounds_h,
                             int ksize_h, const int32_t* kk_v, const int32_t* bounds_v, int ksize_v, void* tmp, int n_backbones,
                             const float* mean, const float* stdv, void* out, emx_stream_t stream);
/* GPU twins of the robot loop's TensorFlow image steps (bit-exact with the numpy float32 restatements in emmax_b200/robot_utils.py):
 * the 0.9-area centre crop + bilinear resize of get_vla_action / get_seq_action (experiments/robot/openvla_utils.py:81-124, :136-156:
 * uint8 -> /255 -> tf.image.crop_and_resize(box y1,x1,y2,x2 normalised) -> clip -> * 255.5, saturate, truncate), uint8 [B,H,W,3] -> [B,Ho,Wo,3] */
int emx_crop_resize_u8(const void* hwc, int B, int H, int W, float y1, float x1, float y2, float x2, void* out, int Ho, int Wo,
                       emx_stream_t stream);
/* and `resize_image`'s tf.image.resize(method="lanczos3", antialias=True) + round + clip + uint8 (experiments/robot/bridge/
 * bridgev2_utils.py:152-166): separable, span starts (int32 [out]) and normalised fp32 weights ([out][ks]) built on the host as TF's
 * ComputeSpansCore does; tmp: fp32 [H, Wo, 3] scratch; uint8 [H,W,3] -> [Ho,Wo,3]. Launches two kernels. */
int emx_lanczos3_resize_u8(const void* hwc, int H, int W, int Ho, int Wo, const int32_t* start_h, const float* w_h, int ks_h,
                           const int32_t* start_v, const float* w_v, int ks_v, void* tmp, void* out, emx_stream_t stream);
int emx_patch_im2col(const void* pixels, int B, int c_total, int chan0, int H, int W, int patch, void* out, int kpad, emx_stream_t stream);
/* tokens[b, 0:prefix] = prefix_tokens; tokens[b, prefix+i] = bf16(patch_out[b,i] + pos[i])  (timm _pos_embed) */
int emx_vit_assemble(const void* patch_out, const void* pos, const void* prefix_tokens, void* tokens, int B, int n_patches, int prefix,
                     int D, emx_stream_t stream);
/* features[b, i, col0:col0+D] = tokens[b, prefix+i, :]   (prefix strip + torch.cat(dim=2), modeling_prismatic.py:123) */
int emx_vit_gather_features(const void* tokens, void* features, int B, int n_patches, int prefix, int D, int ld_features, int col0,
                            emx_stream_t stream);

/* ---- attention, full sequence (ViT: non-causal d=64/72; Llama prefill: causal d=128) ------------------------
 * qkv: [B*T, 3*heads*hd] packed as the fused timm `qkv` Linear / the concatenated q,k,v Llama projections produce it
 * (q | k | v, head-major inside each). out: [B*T, heads*hd]. Replaces F.scaled_dot_product_attention (timm Attention)
 * and flash_attn_varlen_func (transformers LlamaFlashAttention2, selected at openvla_utils.py:45). */
int emx_attn_fwd(const void* qkv, void* out, int B, int T, int heads, int head_dim, int causal, float scale, emx_stream_t stream);

/* ---- Llama prefill helpers ----------------------------------------------------------------------------------
 * RoPE (rotate_half form, bf16 cos/sin tables [max_pos, hd/2]) applied in place to q and k of a packed qkv buffer,
 * then K and V rows are stored into the paged KV cache. positions: pos0 + t for t in [0,T).
 * KV cache page layout: [page][head][page_size][hd] bf16; block_table[b*max_pages + p/page_size] = page id. */
int emx_rope_kvstore(void* qkv, int B, int T, int heads, int head_dim, const void* cos_tab, const void* sin_tab, int pos0, void* k_cache,
                     void* v_cache, const int32_t* block_table, int max_pages, int page_size, emx_stream_t stream);

/* The Llama q|k|v projection of a prefill with emx_rope_kvstore FUSED into the GEMM epilogue (ABI 5): qkv[B*T, 3*heads*head_dim] =
 * A[B*T, K] . W[3*heads*head_dim, K]^T, RoPE applied to the q and k heads straight out of the accumulator, K / V rows appended to the paged
 * cache — one pass less over the 3*hidden-wide rows per layer. Same arithmetic and rounding points as emx_gemm_bf16 followed by
 * emx_rope_kvstore (bit-identical results); shapes the fused epilogue does not cover (head_dim != 128, problems too small for 256-column tiles) run
 * exactly those two kernels. Replaces q_proj / k_proj / v_proj + apply_rotary_pos_emb + DynamicCache.update of LlamaFlashAttention2
 * (transformers 4.40.1, reached from modeling_prismatic.py:404-415). */
int emx_gemm_qkv_rope(const void* A, int lda, const void* W, int ldw, void* qkv, int B, int T, int heads, int head_dim, int K,
                      const void* cos_tab, const void* sin_tab, int pos0, void* k_cache, void* v_cache, const int32_t* block_table,
                      int max_pages, int page_size, emx_stream_t stream);
/* x[b, 0] = E[ids[b,0]]; x[b, 1:1+P] = patches[b]; x[b, 1+P+j] = E[ids[b,1+j]]   (modeling_prismatic.py:380-385) */
int emx_embed_assemble(const int64_t* ids, int n_ids, const void* embed, const void* patches, int n_patches, void* x, int B, int H,
                       emx_stream_t stream);
/* h[m, i] = bf16(bf16(silu(gu[m, 2i])) * gu[m, 2i+1]) — unfused SwiGLU (transformers LlamaMLP) for interleaved gate/up */
int emx_swiglu(const void* gate_up, void* h, int M, int I, emx_stream_t stream);

/* ---- single-token kernels (building blocks; the decode product path is emx_decode_step) ---------------------- */
/* y[N] = bf16(W[N,K] x[K]) (+ resid). */
int emx_gemv_bf16(const void* W, int ldw, const void* x, void* y, const void* resid, int N, int K, emx_stream_t stream);
/* logits = bf16(W x), argmax with lowest-index tie break (GenerationMixin greedy). logits_out (fp32 [N]) optional. */
int emx_lmhead_argmax(const void* W, int ldw, const void* x, int N, int K, float* logits_out, int32_t* token_out, void* scratch,
                      emx_stream_t stream);

/* First tokens of a BATCHED prefill (ABI 5): logits [rows][ld] bf16 (= emx_gemm_bf16 of the rows' last hidden states against lm_head: one pass
 * over lm_head for the whole batch instead of one GEMV per row) -> optional fp32 copy [rows][n] + per-row argmax, lowest index wins ties.
 * Replaces logits[:, -1].argmax(-1) of the first GenerationMixin step (called at modeling_prismatic.py:519) for B > 1. */
int emx_argmax_rows_bf16(const void* logits, int ld, int rows, int n, float* logits_out, int32_t* tokens_out, emx_stream_t stream);

/* ---- persistent decode step ---------------------------------------------------------------------------------
 * ONE launch = one new token for one sequence: embedding gather, 32 x (RMSNorm + q|k|v GEMV + RoPE + paged-KV append +
 * split-KV attention + o_proj + residual + RMSNorm + gate/up GEMV + SwiGLU + down GEMV + residual), final norm,
 * lm_head GEMV + greedy argmax, all inside one persistent kernel (1 CTA/SM, 512 threads: consumer, attention, producer and
 * L2-prefetch warps) that streams the 13.2 GB of weights through a bulk-async (TMA) shared-memory ring; the attention warps
 * stage the cached K/V rows of their (head, split) item in tensor memory a layer ahead. Contexts up to 2048 (kv_splits = 4). There are no grid barriers: CTAs exchange activation vectors as 8-byte
 * "LL units" {2 x bf16 payload | 32-bit tag} (single 64-bit stores, polling 64-bit loads; tag = launch epoch x layer).
 * Replaces the cached branch of PrismaticForConditionalGeneration.forward (modeling_prismatic.py:325-341) + one
 * iteration of GenerationMixin's greedy loop (called at modeling_prismatic.py:519). */
typedef struct emx_decode_state {
  int32_t cur_token; /* token fed to this step (written by the previous step / prefill) */
  int32_t pos;       /* number of tokens already in the KV cache == position id of cur_token */
  int32_t n_generated;
  int32_t finished;  /* set when EOS was produced; later launches return immediately */
  /* kernel-private, zero once at allocation and never touched by the host afterwards: */
  uint32_t barrier;  /* unused since ABI 2 (was: grid-barrier ticket counter) */
  uint32_t epoch;    /* completed launches (monotonic); the LL tags are derived from it */
  uint32_t pad0_[2];
  uint32_t head_ticket[64]; /* unused since ABI 2 */
} emx_decode_state;

typedef struct emx_decode_params {
  /* model */
  int32_t hidden, inter, heads, head_dim, layers, vocab;
  float rms_eps;
  const void* embed;     /* [vocab, hidden] */
  const void* w_qkv;     /* [layers][3*hidden, hidden]  (q rows | k rows | v rows) */
  const void* w_o;       /* [layers][hidden, hidden] */
  const void* w_gateup;  /* [layers][2*inter, hidden]   row 2i = gate_i, row 2i+1 = up_i */
  const void* w_down;    /* [layers][hidden, inter] */
  const void* ln1;       /* [layers][hidden] input_layernorm */
  const void* ln2;       /* [layers][hidden] post_attention_layernorm */
  const void* final_norm;/* [hidden] */
  const void* lm_head;   /* [vocab, hidden] */
  const void* cos_tab;   /* [max_pos, head_dim/2] bf16 */
  const void* sin_tab;
  /* KV cache (paged) */
  void* k_cache;         /* [layers][n_pages][heads][page_size][head_dim] */
  void* v_cache;
  const int32_t* block_table; /* [max_pages] page ids of this sequence */
  int32_t page_size, n_pages, max_pages;
  /* exchange buffers (device, 8-byte aligned, zeroed once at allocation, then owned by the kernel). One LL unit = 8 bytes. */
  void* x;        /* [hidden/2] units: residual stream after down_proj */
  void* xo;       /* [hidden/2] units: residual stream after o_proj */
  void* qkv;      /* [3*hidden/2] units */
  void* attn;     /* [hidden/2] units */
  void* h;        /* [inter/2] units: SwiGLU output */
  void* part;     /* [heads][kv_splits][head_dim + 2] units (fp32 payload): split-KV partials */
  void* argmax_part; /* [grid][2] units: per-CTA argmax candidate (value bits, index) */
  /* outputs */
  int32_t* out_tokens;  /* out_tokens[n_generated] = new token */
  float* logits_out;    /* optional [vocab] fp32 (bf16-rounded values), for parity tests */
  int32_t eos_token;    /* -1 disables EOS handling */
  int32_t kv_splits;
  emx_decode_state* state;
  /* optional profiling buffer (device, >= 15*layers + 16 + 2*grid + 8 int64; selects the instrumented twin of the kernel): CTA 0 stores
   * %globaltimer at every phase boundary (12 per layer), then [15*layers + 4 ..] = gather/RMSNorm cycle counters, cycles the consumers
   * waited for weights / the producers for a free ring slot, bytes prefetched to L2, per-CTA end-of-layer-1 times, and the attention
   * warps' per-layer timings of CTA 0 (tools/decode_probe.py decodes it) */
  int64_t* dbg;
  /* per-CTA look-ahead (KiB) of cp.async.bulk.prefetch.L2 beyond the shared-memory ring; 0 disables (148 CTAs x 256 KiB
   * = 37 MB of the 126 MB L2 keeps HBM streaming while the consumers wait for an exchange and the ring is full) */
  int32_t l2_lookahead_kb;
  /* profiling only (results become garbage for 1, 2, 16, 64): 1 = do not wait for LL tags, 2 = skip attention (needs 1), 4 = no L2
   * evict-first hint, 8 = L2 prefetch also in catch-up mode, 16 = drop LL stores, 32 = stage the next layer's K/V into TMEM immediately,
   * 64 = skip the MMAs, bits 8..19 / 20.. = idle / catch-up prefetch pace in 10 ns per 64 KB (defaults 700 / 1300 ns) */
  int32_t debug_flags;
} emx_decode_params;

int emx_decode_step(const emx_decode_params* params, emx_stream_t stream);
/* number of CTAs emx_decode_step launches (== SM count) and its dynamic shared memory, for sizing scratch buffers */
int emx_decode_grid(void);
/* host-side view of the kernel's static row partition: rows [*r_begin, *r_end) of an n_rows-row weight phase belong to CTA `cta` of `grid`
 * (granule 2 = row pairs; 4 = gate/up quads). Runs on the host; used by the CPU tests. */
int emx_decode_phase_rows(int n_rows, int granule, int cta, int grid, int* r_begin, int* r_end);

/* ---- persistent decode step, up to 8 sequences per launch ------------------------------------------------------
 * ONE launch = one new token for EACH active sequence of a batch of <= 8 (BASELINE.json configs[4]: bs=64 over 8 GPUs = 8 per GPU,
 * mixed max_new_tokens). The reference has no counterpart: its cached branch asserts batch size 1 (modeling_prismatic.py:326,
 * :460-463), i.e. N requests cost N passes over the 13.2 GB of weights per token; here one pass serves 8 sequences (they are the 8
 * columns of the mma.m16n8k16 B operand). Same machinery as emx_decode_step (weight ring, LL exchanges, no grid barriers), plus:
 * activation vectors of all sequences parked in tensor memory in B-fragment order; attention over a linear list of K/V page-pair items
 * split evenly over the CTAs for any mix of context lengths, the pages streamed through the weight ring. page_size must be 64,
 * contexts up to 2048. Per sequence i: state->cur_token[i], pos[i], n_generated[i], finished[i] as in emx_decode_state, plus
 * limit[i] = the sequence's max_new_tokens (it goes inactive when n_generated[i] reaches it). Inactive sequences (i >= batch, finished,
 * at their limit, out of cache) cost nothing and nothing of theirs is written. */
#define EMX_DECODE_MAX_BATCH 8
#define EMX_DECODE_ATT_MAX_SEGMENTS 16
typedef struct emx_decode_batch_state {
  int32_t cur_token[EMX_DECODE_MAX_BATCH];
  int32_t pos[EMX_DECODE_MAX_BATCH];
  int32_t n_generated[EMX_DECODE_MAX_BATCH];
  int32_t finished[EMX_DECODE_MAX_BATCH];
  int32_t limit[EMX_DECODE_MAX_BATCH];
  uint32_t epoch; /* kernel-private: completed launches; zero once at allocation */
  uint32_t pad_[7];
} emx_decode_batch_state;

typedef struct emx_decode_batch_params {
  int32_t hidden, inter, heads, head_dim, layers, vocab;
  float rms_eps;
  int32_t batch;         /* sequences in use, 1..8 */
  const void* embed;     /* weights: layouts as in emx_decode_params */
  const void* w_qkv;
  const void* w_o;
  const void* w_gateup;
  const void* w_down;
  const void* ln1;
  const void* ln2;
  const void* final_norm;
  const void* lm_head;
  const void* cos_tab;
  const void* sin_tab;
  void* k_cache;         /* [layers][n_pages][heads][64][head_dim] */
  void* v_cache;
  const int32_t* block_table; /* [batch][max_pages] */
  int32_t page_size, n_pages, max_pages, out_stride;
  /* exchange buffers (8-byte LL units, zeroed once at allocation, then owned by the kernel); R(n) = n rounded up to a multiple of 16 */
  void* x;        /* [8][R(hidden/2)]: residual stream after down_proj */
  void* xo;       /* [8][R(hidden/2)]: residual stream after o_proj */
  void* attn;     /* [8][R(hidden/2)] */
  void* qkv;      /* [8][3*hidden/2] */
  void* h;        /* [8][R(inter/2)]: SwiGLU output */
  void* part;     /* [8][heads][EMX_DECODE_ATT_MAX_SEGMENTS][head_dim + 2]: split-KV partials (fp32 payloads) */
  void* argmax_part; /* [grid][8][2] */
  void* sync;        /* one uint64 (zeroed once at allocation, kernel-owned): arrival counter of the exchanges */
  int32_t* out_tokens;  /* out_tokens[i * out_stride + n_generated[i]] = new token of sequence i */
  float* logits_out;    /* optional [8][vocab] fp32 (bf16-rounded values), for parity tests */
  emx_decode_batch_state* state;
  int64_t* dbg;         /* optional (>= 2 * (7 * layers + 1) + 8 + 8 * grid + 8 int64, zeroed by the caller): CTA 0 stores %globaltimer before / after every phase's gather,
                         * every CTA the entry / exit times of layer 1's four gathers; selects the instrumented twin */
  int32_t eos_token;    /* -1 disables EOS handling */
  /* look-ahead (ring stages of 64 KB) of the idle-triggered cp.async.bulk.prefetch.L2 warp beyond the shared-memory ring; 0 disables */
  int32_t l2_lookahead_stages;
} emx_decode_batch_params;

int emx_decode_batch_step(const emx_decode_batch_params* params, emx_stream_t stream);
/* dynamic shared memory of the batched decode kernel (host-side query, no CUDA call) */
int emx_decode_batch_smem(void);

/* ---- action de-tokeniser (device twin of ActionTokenizer.decode_token_ids_to_actions + un-normalise) ---------
 * ids [n] int32 -> normalized[n], actions[n] fp64:  k = clip(vocab - id - 1, 0, n_bins-2); c = centres[k];
 * actions = mask ? 0.5*(c+1)*(q99-q01)+q01 : c.   action_tokenizer.py:49-68; modeling_prismatic.py:522-535.
 * stats arrays (q01,q99 fp64; mask uint8; length action_dim) may be null -> only `normalized` is written. */
int emx_detokenize_actions(const int32_t* ids, int n, int vocab_size, int n_bins, const double* q01, const double* q99,
                           const uint8_t* mask, int action_dim, double* normalized, double* actions, emx_stream_t stream);

/* ---- bandwidth probe (profiling aid, not on the product path) -------------------------------------------------
 * Streams `bytes` from `src` through a cp.async.bulk shared-memory ring exactly like emx_decode_step's producer, with
 * configurable stage geometry; returns the number of (rows x row_stride) blocks each CTA streamed, or <0 on error. */
int emx_debug_stream(const void* src, long bytes, int rows, int seg, long row_stride, int stages, int evict_first, int grid,
                     int npairs, emx_stream_t stream);

/* Dataflow skeleton of emx_decode_step (profiling aid): per CTA a static byte schedule of `n_phases` phases
 * (`phase_stages[i]` ring stages of rows x seg bytes each) streamed `reps` times through a `stages`-deep ring, a grid barrier
 * of protocol `barrier_variant` after every phase followed by a consumer stall of `stall_ns[i]`. `sync`: >= 32 * (2*148 + 1)
 * zeroed uint32; `out`: 32 + 2*148 int64 (CTA 0: per-phase ns, per-barrier ns, total ns; per-CTA total ns and %smid). */
int emx_debug_skeleton(const void* src, long region_bytes, int n_phases, const int* phase_stages, const int* stall_ns, int reps,
                       int rows, int seg, long row_stride, int stages, int consume_cycles, int barrier_variant, int n_prod, int n_cons,
                       const float* weight, int timers, int pf_stages, int pf_mode, int pf_pace_ns, int launch_mode, void* sync, void* out,
                       emx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EMMAX_H_ */
