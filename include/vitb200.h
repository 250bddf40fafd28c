/* vitb200.h — C ABI of the B200-native Vision Transformer forward engine (libvitb200.so).
 *
 * This is the drop-in boundary for the arithmetic behind interactive-vit's server-side node graph.  In the
 * reference every FLOP of a network node is spent inside
 *     Model.compute(node_name, pinin) -> sub(x) under torch.no_grad()        main/context.py:79-88
 * reached from Context.compute (main/context.py:143-147) through ModelNode.compute (main/context.py:119-121).
 * For a ViT those `sub(x)` calls are torchvision's VisionTransformer stages (an un-vendored dependency of the
 * reference: torchvision/models/vision_transformer.py, "TV" below).  Each entry point names the reference-side
 * call it replaces.  The reference has no FFI of its own (it is pure Python); the binding a maintainer would
 * add is the ctypes stub shown in INTEGRATION.md (shipped as interactive-vit_b200/engine.py).
 *
 * Conventions: plain pointers and sizes only; `*_host` pointers are CPU memory (pinned memory makes the
 * copies asynchronous), `*_dev` pointers are CUDA device memory of the engine's device; all tensors are
 * dense, row-major, fp32 at the boundary (the reference's wire format is fp32: main/message.py:53-59,
 * 111-121).  Every function returns VITB200_OK (0) or a negative error code; vitb200_last_error() returns
 * the message for the calling thread (the Python shim raises it, which the reference turns into HTTP 400,
 * main/views.py:40-42).  There is no CPU fallback: without a CUDA device vitb200_create fails.
 */
#ifndef VITB200_H_
#define VITB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITB200_OK 0
#define VITB200_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define VITB200_ERR_CUDA (-2)    /* CUDA runtime or driver error (message has the CUDA string) */
#define VITB200_ERR_STATE (-3)   /* call out of order (e.g. forward before all weights are loaded) */

/* Output selection flags for vitb200_forward_* (bit-or). */
#define VITB200_EMIT_AVG 1u     /* head-averaged attention probabilities per layer  [L, B, N, N]    */
#define VITB200_EMIT_CLS 2u     /* per-head class-token attention rows per layer     [L, B, H, N]    */
#define VITB200_EMIT_ROLLOUT 4u /* attention rollout of the class token              [B, N-1]        */
#define VITB200_EMIT_HEADS 8u   /* full per-head probabilities per layer (large)     [L, B, H, N, N] */
#define VITB200_EMIT_HIDDEN 16u /* residual stream after every layer                 [L, B, N, d]    */

typedef struct vitb200_engine vitb200_engine;

/* Architecture of torchvision.models.vision_transformer.VisionTransformer (TV:160-266). */
typedef struct vitb200_config {
  int image_size;  /* S: square input side, multiple of patch_size                    */
  int patch_size;  /* p: 16 (multiple of 8)                                            */
  int num_layers;  /* L                                                                */
  int num_heads;   /* H; head dim hidden_dim / H in 64..128, multiple of 16 (64 or 80)  */
  int hidden_dim;  /* d: multiple of 128                                               */
  int mlp_dim;     /* multiple of 64                                                   */
  int num_classes; /* multiple of 8                                                    */
  int max_batch;   /* workspace is sized for this many images (grows on demand)        */
  int device;      /* CUDA device ordinal                                              */
  int precision;   /* 0: bf16 operands, fp32 accumulation (<= 2e-2 of the fp32 reference)
                      1: "fp32x3": every matmul operand is carried as hi + lo bf16 and multiplied as
                         hi*hi + lo*hi + hi*lo with fp32 accumulation (<= 1e-3; ~3x tensor work; attention at head dims other than 64
                         runs in fp32 on the CUDA cores) */
} vitb200_config;

/* Host-side result pointers for vitb200_forward_host; any pointer may be NULL (output skipped). */
typedef struct vitb200_host_outputs {
  float* logits;   /* [B, num_classes]                                  heads(x[:,0])      TV:302-304 */
  float* avg_maps; /* [L, B, N, N]   mean over heads of softmax(QK^T/sqrt(D))  torch/nn/functional.py:6659 */
  float* cls_maps; /* [L, B, H, N]   softmax row of query token 0, per head                */
  float* rollout;  /* [B, N-1]       attention rollout, class-token row (oracle/vit_oracle.py) */
  float* heads;    /* [L, B, H, N, N] per-head probabilities           torch/nn/functional.py:6645-6653 */
  float* hidden;   /* [L, B, N, d]   EncoderBlock outputs              TV:110-119          */
} vitb200_host_outputs;

const char* vitb200_last_error(void);
int vitb200_version(void);

/* Replaces: constructing the torchvision module in the plugin's __init__ (static/models/vgg16.py:11-14 is the
 * in-tree template).  Creates the CUDA context objects, the stream and the workspace. */
int vitb200_create(const vitb200_config* cfg, vitb200_engine** out);
void vitb200_destroy(vitb200_engine* e);

/* Replaces: nn.Module.load_state_dict.  `name` is the torchvision state-dict key (conv_proj.weight,
 * class_token, encoder.pos_embedding, encoder.layers.encoder_layer_{i}.{ln_1,ln_2}.{weight,bias},
 * ...self_attention.in_proj_{weight,bias}, ...self_attention.out_proj.{weight,bias}, ...mlp.{0,3}.{weight,bias},
 * encoder.ln.{weight,bias}, heads.head.{weight,bias}); `data_host` holds `count` fp32 values in the
 * state-dict layout.  GEMM weights are packed to bf16 on the device. */
int vitb200_load_weight(vitb200_engine* e, const char* name, const float* data_host, size_t count);
/* VITB200_OK once every tensor of the architecture has been loaded. */
int vitb200_weights_ready(vitb200_engine* e);

/* Replaces: VisionTransformer.forward(x) (TV:289-306) for a whole batch, plus attention-map extraction
 * (need_weights=True in EncoderBlock, TV:113).  images: fp32 [B, 3, S, S].  Host variant: H2D copy, forward,
 * D2H copies of the requested outputs, synchronous on return.  Device variant: enqueues on `stream`, a cudaStream_t
 * taken literally -- NULL is the legacy default stream, as in the op_* entry points (and the handle torch reports for
 * its default stream) -- and returns without synchronising; results stay in the engine's device buffers
 * (vitb200_device_output).  vitb200_engine_stream returns the engine's own (non-blocking) stream, on which the host,
 * pipelined and node-granular entry points run; pass it to run a device forward there.
 * From the second call with the same (batch, flags, images_dev, bound outputs) the launch sequence is replayed as ONE
 * captured CUDA graph (any stream but the legacy default one; vitb200_set_graphs(e, 0) or VITB200_GRAPHS=0 disables). */
int vitb200_forward_host(vitb200_engine* e, const float* images_host, int batch, uint32_t flags,
                         const vitb200_host_outputs* out);
int vitb200_forward_device(vitb200_engine* e, const float* images_dev, int batch, uint32_t flags, void* stream);
int vitb200_engine_stream(vitb200_engine* e, void** stream);

/* Workspace management.  The activation buffers grow on demand (any call with a larger batch or a new output flag);
 * growth RE-ALLOCATES them, so device-resident state a caller relies on between calls -- the token stream between two
 * node calls, the maps a rollout node will read, preprocessed images -- is lost.  vitb200_workspace_generation returns a
 * counter that changes whenever that happened (the plugin drops its residency shortcuts when it moves);
 * vitb200_reserve grows the workspace up front for `batch` images and the outputs in `flags`, so that no later call of
 * the same request can.  vitb200_set_graphs switches CUDA-graph replay of forwards / stages off (0) or on;
 * vitb200_graph_replays counts replays (harness evidence that the captured path is the one that ran). */
uint64_t vitb200_workspace_generation(vitb200_engine* e);
int vitb200_reserve(vitb200_engine* e, int batch, uint32_t flags);
int vitb200_set_graphs(vitb200_engine* e, int on);
uint64_t vitb200_graph_replays(vitb200_engine* e);

/* Multi-GPU result exchange without a collective (the reference is single-process, so this
 * replaces nothing there).  Routes the small results of vitb200_forward_device into caller-owned device memory --
 * typically this rank's slice of rank 0's receive buffer, mapped into this process over NVLink (CUDA peer / symmetric
 * memory): the kernels that PRODUCE the results (head GEMM epilogue -> logits, the attention kernel's CLS-row writer,
 * the rollout kernel) store there directly, so the "gather" has no pack kernel, no staging copy and no NCCL call;
 * the caller only signals completion (dist.PeerPush).
 *   logits_dev : [batch, classes] rows (16-byte aligned), or NULL = engine buffer
 *   cls_dev    : layer l, image b, head h at cls_dev + l * cls_layer_stride + (b * heads + h) * tokens  (floats), or NULL
 *   rollout_dev: [batch, tokens - 1] rows, or NULL
 * All NULL restores the engine's own buffers.  vitb200_forward_device and vitb200_submit_host honour the binding (a
 * bound output must then not also be requested on the host: it is neither staged nor copied there); the synchronous
 * host and the node-granular entry points keep using the engine's buffers.  Completion signalling must be enqueued on
 * the stream the forward runs on (the `stream` argument, or the engine's own stream for submit_host). */
int vitb200_bind_outputs(vitb200_engine* e, float* logits_dev, float* cls_dev, long cls_layer_stride, float* rollout_dev);

/* Plumbing for that exchange, free of any framework dependency (nothing in the reference):
 * vitb200_peer_alloc allocates zeroed device memory on `device` and returns its 64-byte CUDA IPC handle, which the
 * caller ships to the other ranks' processes (torch.distributed is only the courier); vitb200_peer_open maps it there
 * (peer access over NVLink is enabled on demand).  vitb200_flag_signal / vitb200_flag_wait enqueue one-thread kernels on
 * `stream`: a system-scope release store of a monotone counter value, and a back-off spin until the counter has
 * reached `value` (watchdog: 30 s, then the kernel traps).  csrc/peer.cuh has the protocol built from them. */
int vitb200_peer_alloc(int device, size_t bytes, void** ptr_dev, void* handle64);
int vitb200_peer_open(int device, const void* handle64, void** ptr_dev);
int vitb200_peer_close(int device, void* ptr_dev);
int vitb200_peer_free(int device, void* ptr_dev);
int vitb200_flag_signal(void* flag_dev, uint32_t value, void* stream);
int vitb200_flag_wait(const void* flag_dev, uint32_t value, void* stream);

/* Pipelined variant of vitb200_forward_host for request streams: returns once the work is enqueued.  Up to two
 * requests are in flight; the host-to-device copy of request i+1 and the device-to-host copies of request i-1
 * overlap the forward of request i (separate copy streams, double-buffered device inputs, per-request staging of
 * logits / CLS maps / rollout / head-averaged maps).  `images_host` and the `out` pointers must stay valid (and should
 * be pinned) until vitb200_wait(ticket) returns; per-head maps and hidden states are not available on this path.
 * vitb200_staged_output: device copy of one output of an in-flight or just-completed ticket (for a device-side
 * gather): which = 0 (logits) or VITB200_EMIT_{AVG,CLS,ROLLOUT}; dense layouts [B,classes] / [L,B,N,N] / [L,B,H,N] /
 * [B,N-1]. */
int vitb200_submit_host(vitb200_engine* e, const float* images_host, int batch, uint32_t flags,
                        const vitb200_host_outputs* out, uint64_t* ticket);
int vitb200_wait(vitb200_engine* e, uint64_t ticket);
int vitb200_staged_output(vitb200_engine* e, uint64_t ticket, uint32_t which, float** ptr_dev);

/* Measurement aid: one forward on the engine's own stream with a CUDA event in front of every kernel launch.
 * `report` receives text lines "kernel,launches,total_ms" (event-to-event times, so each kernel's figure
 * includes the gap to the next launch) and a final "total,<kinds>,<ms>" line.  Synchronous. */
int vitb200_profile_forward(vitb200_engine* e, const float* images_dev, int batch, uint32_t flags, char* report,
                            size_t report_cap);

/* Device-resident results of the last forward/stage call.  which: one of the VITB200_EMIT_* flags, or 0 for
 * logits.  Returns the base pointer and the row pitch (in floats) of the innermost matrix: attention maps
 * are stored with their rows padded to `pitch` >= N floats. */
int vitb200_device_output(vitb200_engine* e, uint32_t which, float** ptr_dev, int* pitch);
int vitb200_synchronize(vitb200_engine* e);

/* Node-granular entry points: one per block-granular graph node of the ViT plugin (embed / layer.i / head /
 * rollout), i.e. one per Model.compute call of the reference (main/context.py:79-88).  The token stream
 * [B, N, d] fp32 lives in the engine between calls; set/get move it across the boundary when a node's input
 * did not come from (or its output must leave) the engine. */
/* Deferred mode for the node-granular calls below (one request = ~15 node calls on one stream): with on != 0 the calls
 * that take no host INPUT (stage_embed_resident / stage_layer / stage_attn_block / stage_mlp_block / stage_head /
 * stage_rollout / get_*) enqueue their kernels and device-to-host copies and return WITHOUT synchronising; the host
 * outputs (which should be pinned, and must stay allocated) are valid after vitb200_synchronize.  Calls with a host
 * input (stage_embed, stage_transform, set_*) still wait, so the caller's source buffer is free on return.  The plugin
 * (vit_plugin.py) hands such outputs out as tensors that synchronise on first access. */
int vitb200_set_deferred(vitb200_engine* e, int on);
int vitb200_stage_embed(vitb200_engine* e, const float* images_host, int batch);         /* TV:268-287,295-296 + pos add TV:155 */
/* `<model>:transform` node: torchvision's ImageClassification preset (transforms/_presets.py:58-65; the reference's
 * VggModel runs weights.transforms() on the CPU, static/models/vgg16.py:40-42): antialiased bilinear resize of the
 * shorter side to `resize`, centre crop to the model's image_size, ImageNet normalisation.  images_host: fp32
 * [B,3,H,W] in [0,1]; out_host (may be NULL): fp32 [B,3,S,S].  The result also stays in the engine's image buffer:
 * vitb200_stage_embed_resident embeds it without another upload. */
int vitb200_stage_transform(vitb200_engine* e, const float* images_host, int batch, int H, int W, int resize,
                            float* out_host);
int vitb200_stage_embed_resident(vitb200_engine* e, int batch);
int vitb200_stage_layer(vitb200_engine* e, int layer, int batch, uint32_t flags);        /* TV:110-119 */
/* One layer node in one call: vitb200_stage_layer (half == 0) or vitb200_stage_attn_block (half == 1) followed by the
 * copies of its outputs -- the token stream [batch, tokens, width], the head-averaged map [batch, tokens, tokens] and the
 * class token's per-head attention to the patch tokens [batch, heads, tokens - 1] (vitb200_get_tokens / _get_avg_map /
 * _get_cls_grid; any pointer may be NULL).  Saves three calls and their stream bookkeeping per node of a request. */
int vitb200_stage_layer_fetch(vitb200_engine* e, int layer, int batch, uint32_t flags, int half, float* tokens_host,
                              float* avg_host, float* cls_grid_host);
/* The two halves of an EncoderBlock as separate nodes (`<model>:layer.<i>.attn`, `<model>:layer.<i>.mlp`; SURVEY.md
 * section 8f-4, finer-grained graphs): attn = x + out_proj(MHA(LN1 x)) with the maps `flags` ask for (TV:112-116),
 * mlp = x + MLP(LN2 x) (TV:118-119).  stage_attn_block followed by stage_mlp_block is stage_layer, bit for bit. */
int vitb200_stage_attn_block(vitb200_engine* e, int layer, int batch, uint32_t flags);
int vitb200_stage_mlp_block(vitb200_engine* e, int layer, int batch);
int vitb200_stage_head(vitb200_engine* e, int batch, float* logits_host);                /* TV:157,302-304 */
int vitb200_stage_rollout(vitb200_engine* e, int batch, float* rollout_host);            /* needs avg maps of all layers */
int vitb200_set_tokens(vitb200_engine* e, const float* tokens_host, int batch);          /* [B, N, d] */
int vitb200_get_tokens(vitb200_engine* e, float* tokens_host, int batch);
int vitb200_set_avg_map(vitb200_engine* e, int layer, const float* map_host, int batch); /* [B, N, N] */
int vitb200_get_avg_map(vitb200_engine* e, int layer, float* map_host, int batch);
int vitb200_get_cls_map(vitb200_engine* e, int layer, float* map_host, int batch);       /* [B, H, N] */
int vitb200_get_head_map(vitb200_engine* e, int layer, float* map_host, int batch);      /* [B, H, N, N] */
int vitb200_get_cls_grid(vitb200_engine* e, int layer, float* map_host, int batch);      /* [B, H, N-1]: CLS rows without the class column (the UI's [H, g, g] view) */

/* Counters for the harness: kernels launched by this engine since creation. */
uint64_t vitb200_launch_count(vitb200_engine* e);

/* Single-kernel entry points (device pointers, `stream` = cudaStream_t or NULL for the default stream).
 * They exist so the parity tests can check each kernel against the oracle in isolation. */
int vitb200_op_gemm(const void* a_bf16_dev, const void* w_bf16_dev, const float* bias_dev, const float* resid_dev,
                    void* out_dev, int M, int N, int K, int gelu, int out_f32, void* stream);
/* GEMM with the LayerNorm-folding epilogues (see the header comment of csrc/engine.cu).  Producer side (needs resid_dev):
 * xb_out_dev receives the bf16 copy of the fp32 result and stats_out_dev [M, N/w, 2] the partial sums (sum, sum of
 * squares) of every row per group of w columns, w = 128 if N is a multiple of 256, else 64 (N a multiple of 128).
 * Consumer side: stats_in_dev [M, K/w, 2] (same rule for K) + colsum_dev [N] turn the GEMM into
 * LayerNorm(x) W^T + b for weights prepared by vitb200_op_fold_ln.  Unused pointers are NULL. */
int vitb200_op_gemm_ex(const void* a_bf16_dev, const void* w_bf16_dev, const float* bias_dev, const float* resid_dev,
                       void* out_dev, int M, int N, int K, int gelu, int out_f32, void* xb_out_dev, float* stats_out_dev,
                       const float* stats_in_dev, const float* colsum_dev, float ln_eps, void* stream);
/* fp32x3 precision mode building blocks: split an fp32 array into hi = bf16(x), lo = bf16(x - hi); GEMM over split
 * operands (A_hi W_hi + A_lo W_hi + A_hi W_lo, fp32 accumulate).  out_lo_dev (bf16 outputs only, may be NULL) receives
 * the low half of the result. */
int vitb200_op_split_bf16(const float* in_dev, void* hi_dev, void* lo_dev, size_t count, void* stream);
int vitb200_op_gemm_split(const void* a_hi_dev, const void* a_lo_dev, const void* w_hi_dev, const void* w_lo_dev,
                          const float* bias_dev, const float* resid_dev, void* out_dev, void* out_lo_dev, int M, int N, int K,
                          int gelu, int out_f32, void* stream);
/* W'[n,k] = bf16(gamma[k] W[n,k]); colsum[n] = sum_k W'[n,k]; bias_out[n] = bias[n] + sum_k beta[k] W[n,k]. */
int vitb200_op_fold_ln(const float* w_dev, const float* gamma_dev, const float* beta_dev, const float* bias_dev,
                       void* wq_bf16_dev, float* colsum_dev, float* bias_out_dev, int N, int K, void* stream);
int vitb200_op_layernorm(const float* x_dev, const float* gamma_dev, const float* beta_dev, void* y_bf16_dev,
                         int rows, int d, float eps, void* stream);
int vitb200_op_attention(const void* qkv_bf16_dev, void* ctx_bf16_dev, float* avg_dev, float* cls_dev,
                         float* heads_dev, int batch, int tokens, int heads, int pitch, void* stream);
/* Same with an explicit head dimension (64..128 in steps of 16; ViT-H uses 80).  Token counts above 208 or head
 * dimensions other than 64 take the key-blocked two-kernel path (attention_long.cuh).  The fused kernels multiply FP16
 * probabilities with FP16 values: v_is_f16 != 0 says the V third of the input already is fp16 (as the forward's qkv
 * GEMM writes it); with 0 the entry point converts a scratch copy first. */
int vitb200_op_attention_ex(const void* qkv_bf16_dev, void* ctx_bf16_dev, float* avg_dev, float* cls_dev,
                            float* heads_dev, int batch, int tokens, int heads, int head_dim, int pitch, int v_is_f16,
                            void* stream);
int vitb200_op_preprocess(const float* images_dev, float* out_dev, int batch, int H, int W, int resize, int crop,
                          void* stream);
int vitb200_op_patchify(const float* images_dev, void* patches_bf16_dev, int batch, int image_size, int patch,
                        void* stream);
int vitb200_op_rollout(const float* maps_dev, long layer_stride, int layers, int batch, int tokens, int pitch,
                       float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITB200_H_ */
