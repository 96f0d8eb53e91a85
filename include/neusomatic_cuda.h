/*
 * libneusomatic_cuda -- C ABI of the B200-native NeuSomatic candidate-scoring path.
 *
 * The reference (bioinform/neusomatic) has no FFI: its seam for this path is the Python
 * class neusomatic/python/network.py:39-78 (NeuSomaticNet) driven by call.py:54-118
 * (call_variants) and fed by dataloader.py:245-417 (NeuSomaticDataset.__getitem__).
 * Every entry point below names the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero NSC_E_* code otherwise;
 *     nsc_last_error() then returns a thread-local, NUL-terminated description;
 *   - plain pointers and sizes only; the caller owns every buffer it passes in; the
 *     library owns only its packed weights and its per-model device workspace;
 *   - "dev" pointers are CUDA device pointers on the model's device, "host" pointers are
 *     host memory (pinned memory makes the copies asynchronous, pageable also works);
 *   - nsc_forward_* enqueue all work on `stream` (a cudaStream_t passed as void*, NULL =
 *     the legacy default stream) and do not synchronise; nsc_score_host_* synchronise
 *     before returning because their results land in host memory;
 *   - one thread at a time per nsc_model handle;
 *   - there is NO CPU fallback: without a CUDA device nsc_model_create fails.
 */
#ifndef NEUSOMATIC_CUDA_H_
#define NEUSOMATIC_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSC_ABI_VERSION 1

/* error codes */
#define NSC_OK 0
#define NSC_E_INVALID 1   /* bad argument / unknown parameter name / wrong size            */
#define NSC_E_STATE 2     /* model not finalized, missing parameters ...                   */
#define NSC_E_CUDA 3      /* a CUDA runtime / driver call failed (text in nsc_last_error)  */
#define NSC_E_NOMEM 4

/* arithmetic of the convolution stack (nsc_model_set_precision) */
#define NSC_PREC_SPLIT16 0    /* tcgen05 implicit GEMM, 16-bit hi+lo operands (fp16), 3 MMAs per MAC, fp32 accumulate, the
                                 accumulator's round-toward-zero shrink compensated in the folded weights: fp32-level
                                 results (<= 1 % of the 4-decimal probabilities differ from the reference's).  DEFAULT */
#define NSC_PREC_SINGLE16 1   /* tcgen05, single-pass 16-bit operands: fast, FAILS the parity bar */
#define NSC_PREC_SPLIT_BF16 NSC_PREC_SPLIT16   /* historical names */
#define NSC_PREC_BF16 NSC_PREC_SINGLE16
#define NSC_PREC_FP32 2       /* fp32 FFMA kernels (CUDA cores): bit-stable reference path    */
#define NSC_PREC_SPLIT8 3     /* tcgen05: fp16 main product + the two correction products on the fp8 (e4m3) path
                                 with power-of-two scaling, 2 instruction slots per MAC: 1.25 x the throughput of
                                 SPLIT16; meets the parity bar (argmax / 2e-3 / 1e-3) with 10 x margin, but the 4-bit
                                 correction operands leave ~1e-4 of noise in the probabilities (1-8 % of the 4-decimal
                                 values differ).  Opt-in fast mode; needs |folded w| <= 28                      */

typedef struct nsc_model nsc_model;

int nsc_abi_version(void);
const char* nsc_last_error(void);

/* NeuSomaticNet.__init__ (network.py:41-64) + `.cuda()` (call.py:411-413).
 * num_channels: 26 (standalone) or 119 (ensemble) (call.py:409-410); any value in [1,128]
 * is accepted.  `device` is a CUDA ordinal. */
int nsc_model_create(nsc_model** out, int device, int num_channels);

/* net.load_state_dict (call.py:441-460): one call per tensor of the reference state_dict.
 * `name` is the reference key WITHOUT a "module." prefix ("conv1.weight",
 * "res_layers.2.bn_r1.running_var", "fc4.bias", ...; network.py:41-64); `data` is a host
 * pointer to `numel` contiguous fp32 values in PyTorch layout.  `*.num_batches_tracked`
 * is accepted and ignored.  Unknown names / wrong sizes fail with NSC_E_INVALID. */
int nsc_model_set_param(nsc_model* m, const char* name, const float* data, int64_t numel);

/* Folds every BatchNorm2d (eval mode, network.py:24,27,46; eps as nn.BatchNorm2d) into the
 * preceding convolution, packs the weights for the kernels (tap-major slices, hi/lo
 * split) and uploads them.  Idempotent; must be called after the last set_param and
 * before the first forward.  Fails with NSC_E_STATE if a parameter is missing. */
int nsc_model_finalize(nsc_model* m, float bn_eps);

/* Selects the arithmetic of the convolution stack (NSC_PREC_*). */
int nsc_model_set_precision(nsc_model* m, int precision);
int nsc_model_get_precision(const nsc_model* m);

/* checkpoint["normalize_channels"] (call.py:437-439; dataloader.py:386-388): scale the odd /
 * even auxiliary channels 3..22 by the tumor / normal count channel.  Default 0 (every
 * bundled model).  Only affects nsc_forward_u8 / nsc_score_host_u8. */
int nsc_model_set_normalize_channels(nsc_model* m, int on);

/* Keep fp32 copies of the internal activations of the first `n` candidates of each forward
 * chunk for nsc_debug_internal (0 switches it off). */
int nsc_model_set_debug(nsc_model* m, int64_t n);

/* NeuSomaticNet.forward (network.py:66-78) on the assembled fp32 input.
 *   x_dev        [B, C, 5, 32] fp32, contiguous NCHW (what call.py:74-77 feeds the net)
 *   type_logits  [B,4]  o1 = fc2 (DEL, INS, NONE, SNP; call.py:407)
 *   pos          [B,1]  o2 = fc3
 *   len_logits   [B,4]  o3 = fc4
 * Optional (may be NULL) fused extras replacing the per-candidate softmax / torch.max of
 * call.py:80-83,97-100:
 *   type_prob [B,4], len_prob [B,4]   softmax over the 4 logits
 *   type_argmax [B], len_argmax [B]   first maximal index (torch.max semantics)
 * All outputs are device pointers.  B may be 0. */
int nsc_forward_f32(nsc_model* m, const float* x_dev, int64_t B,
                    float* type_logits, float* pos, float* len_logits,
                    float* type_prob, float* len_prob,
                    int32_t* type_argmax, int32_t* len_argmax, void* stream);

/* NeuSomaticDataset.__getitem__ (is_test branch, dataloader.py:383-417, including the
 * matrix_transform of dataloader.py:25-36 / call.py:408) fused in front of forward():
 *   matrices_dev [B,5,32,23] uint8  the stored pileup image (dataloader.py:38-39)
 *   meta_dev     [B,3] int32        (center, tumor_cov, normal_cov) parsed from the tag
 *                                   (dataloader.py:267-274)
 *   anns255_dev  [B, C-26] fp32     ensemble annotation channels exactly as the reference stores
 *                                   them in its float64 matrix, cast to fp32:
 *                                   float(double(ann) * 255.0) (dataloader.py:395-396); NULL
 *                                   when C == 26
 *   coverage_thr                    checkpoint["coverage_thr"] (call.py:434-436)
 * Channel assembly happens on the device; outputs as nsc_forward_f32. */
int nsc_forward_u8(nsc_model* m, const uint8_t* matrices_dev, const int32_t* meta_dev,
                   const float* anns255_dev, float coverage_thr, int64_t B,
                   float* type_logits, float* pos, float* len_logits,
                   float* type_prob, float* len_prob,
                   int32_t* type_argmax, int32_t* len_argmax, void* stream);

/* The batch step of call_variants (call.py:68-83) with HOST buffers: copies the inputs to the
 * device in chunks (double-buffered on internal streams), runs forward, copies the
 * results back and synchronises.  Output pointers are host pointers with the shapes of
 * nsc_forward_f32; the optional ones may be NULL. */
int nsc_score_host_f32(nsc_model* m, const float* x_host, int64_t B,
                       float* type_logits, float* pos, float* len_logits,
                       float* type_prob, float* len_prob,
                       int32_t* type_argmax, int32_t* len_argmax);
int nsc_score_host_u8(nsc_model* m, const uint8_t* matrices_host, const int32_t* meta_host,
                      const float* anns255_host, float coverage_thr, int64_t B,
                      float* type_logits, float* pos, float* len_logits,
                      float* type_prob, float* len_prob,
                      int32_t* type_argmax, int32_t* len_argmax);
/* The same batch step fed with the TSV's own image fields: extract_zlib (dataloader.py:38-39,
 * zlib.decompress(base64.b64decode(field))) runs ON THE DEVICE, so a candidate crosses PCIe as ~1.3 KB of text.
 * text_host: bytes holding the fields (nsc_tsv_stage copies the file's lines there verbatim; pinned memory keeps the
 * copies asynchronous); spans_host [B][2] uint32 = (offset in text_host, characters) of every candidate's
 * base64(zlib(uint8[5,32,23])) field, pairwise disjoint (the device decodes every field in place in ITS copy of the text;
 * text_host itself is only read); meta / anns255 / outputs as in nsc_score_host_u8.  A field the device decoder
 * declines (malformed or unusual stream) is decoded by zlib on the host before its slice is scored: results never depend
 * on where a candidate was inflated, and a field zlib rejects fails the call (NSC_E_INVALID) like the reference's worker.
 * host_decoded (optional): how many candidates of this call went that way. */
int nsc_score_host_b64(nsc_model* m, const uint8_t* text_host, int64_t text_bytes, const uint32_t* spans_host,
                       const int32_t* meta_host, const float* anns255_host, float coverage_thr, int64_t B,
                       float* type_logits, float* pos, float* len_logits,
                       float* type_prob, float* len_prob,
                       int32_t* type_argmax, int32_t* len_argmax, int64_t* host_decoded);

/* Test / profiling hooks (no reference equivalent; network.py:78 returns internal_outs that
 * every caller discards).  Copies an internal activation of the LAST forward chunk to a
 * device buffer as fp32 NCHW: which = 0 -> pool1 output [B,64,5,32]; 1..4 -> output of
 * res_layers[which-1]; 5 -> relu(fc1) [B,240].  `n` candidates starting at 0. */
int nsc_debug_internal(nsc_model* m, int which, float* out_dev, int64_t n, void* stream);

/* Test hook: runs only the channel-assembly kernel of nsc_forward_u8 and writes the fp32
 * [B,C,5,32] tensor it feeds the network to x_out_dev (checked bit-for-bit against the
 * reference's NeuSomaticDataset output in tests/). */
int nsc_debug_assemble(nsc_model* m, const uint8_t* matrices_dev, const int32_t* meta_dev,
                       const float* anns255_dev, float coverage_thr, int64_t B, float* x_out_dev,
                       void* stream);

/* Per-kernel device timing (bench.py `roofline`): while switched on, every kernel launch of
 * the forward path is bracketed by CUDA events on the launching stream.
 * nsc_model_get_profile synchronises the device, adds the elapsed milliseconds and launch
 * counts per slot to ms[] / launches[] (NSC_PROF_SLOTS entries each) and resets the record.
 * Slots: 0 channel assembly, 1 conv1 stage, 2+2*i+j = res_layers[i].conv_r{j+1} stage
 * (i = 0..3, j = 0..1, epilogues fused), 10 fc1+heads, 11 input packing, 12 max-pools (tensor path). */
#define NSC_PROF_SLOTS 16
int nsc_model_set_profile(nsc_model* m, int on);
int nsc_model_get_profile(nsc_model* m, double* ms, int64_t* launches);
const char* nsc_profile_slot_name(int slot);

/* Number of kernel launches issued by this model since creation (bench.py `gpu_launches`). */
int64_t nsc_model_launch_count(const nsc_model* m);
/* Name of the dominant (conv) kernel family used by the current precision mode. */
const char* nsc_model_kernel_name(const nsc_model* m);

int nsc_model_destroy(nsc_model* m);

/* ---- training step -------------------------------------------------------------------
 * The 64 -> 64 convolutions (forward, data gradient, weight gradient; the stem's weight gradient too) run on tcgen05 in
 * split fp16 with fp32 accumulation; BatchNorm, pooling, loss and the linear layers are fp32 CUDA-core kernels.
 * Environment, read by nsc_trainer_create: NSC_TRAIN_FP32=1 -> every convolution on the fp32 CUDA-core kernels (the
 * bit-stable reference mode of the tests); NSC_TRAIN_TENSOR=<subset of "fwd","bwd","wgrad"> -> only those on tcgen05.
 * One optimisation step of train.py:379-383,416-441: training-mode forward (BatchNorm2d batch statistics and
 * running-stat update), loss = CrossEntropyLoss(w_type)(o1,y) + SmoothL1Loss()(o2[:,0],center) +
 * CrossEntropyLoss(w_len)(o3,y_len), loss.backward(), optim.SGD(lr, momentum).step().
 * Parameters / gradients / momentum are CALLER-OWNED flat fp32 device vectors in nn.Module.named_parameters()
 * order (network.py:41-64; nsc_trainer_param_slice gives each tensor's offset), `running` is a flat fp32 device
 * vector [9 BatchNorm layers][mean 64 | var 64] in module order (bn1, res_layers.0.bn_r1, .bn_r2, ...).
 * A data-parallel driver all-reduces `grads` (one NCCL call on the flat vector) between the two calls. */
typedef struct nsc_trainer nsc_trainer;
int nsc_trainer_create(nsc_trainer** out, int device, int num_channels, int64_t max_batch);
int64_t nsc_trainer_num_params(const nsc_trainer* t);
int nsc_trainer_param_slice(const nsc_trainer* t, const char* name, int64_t* offset, int64_t* numel);
/* x_dev [B,C,5,32] fp32; type_labels / len_labels [B] int32; centers [B] fp32 (var_pos_s[:,1]); w_type_dev /
 * w_len_dev [4] fp32 class weights (train.py:364-381); grads [num_params] is overwritten; losses_dev [4] =
 * (total, CE type, SmoothL1, CE length) or NULL.  2 <= B <= max_batch.  Work is enqueued on `stream`. */
int nsc_train_forward_backward(nsc_trainer* t, const float* params, float* running, const float* x_dev,
                               const int32_t* type_labels, const float* centers, const int32_t* len_labels,
                               const float* w_type_dev, const float* w_len_dev, int64_t B, float bn_momentum,
                               float* grads, float* losses_dev, void* stream);
/* Data-parallel loss normalisation.  The reference trains under nn.DataParallel, which computes ONE loss over the gathered
 * global batch (train.py:436-441): both weighted cross entropies divide by the class-weight sum of the global batch and
 * SmoothL1 by its size.  norm_dev [3] fp32 (device) = (sum of w_type[label], sum of w_len[len_label], candidates) over ALL
 * ranks makes nsc_train_forward_backward use those normalisers; the per-rank gradients then SUM (not average) to the
 * gradient of the global loss.  NULL (default): normalise by this batch.  The pointer is read when the step runs. */
int nsc_trainer_set_loss_norm(nsc_trainer* t, const float* norm_dev);
/* The same step in two halves, for callers that compute the loss themselves (the reference's loop: outputs = net(x);
 * loss = criterion(...); loss.backward() -- train.py:426-441): the forward leaves the logits in logits9 [B,9] = (type 4,
 * position 1, length 4 per candidate) and the activations in the trainer; the backward takes d(loss)/d(logits9) and writes
 * every parameter gradient.  One forward may be outstanding per trainer; x_dev must be the same tensor in both calls. */
int nsc_train_forward(nsc_trainer* t, const float* params, float* running, const float* x_dev, int64_t B, float bn_momentum,
                      float* logits9, void* stream);
int nsc_train_backward(nsc_trainer* t, const float* params, const float* x_dev, const float* dlogits9, int64_t B,
                       float* grads, void* stream);
/* optim.SGD: buf = momentum * buf + grad_scale * grad ; p -= lr * buf   (grad_scale = 1 / world size after a
 * sum all-reduce; momentum buffers start at zero). */
int nsc_sgd_momentum_update(float* params, float* momentum_buf, const float* grads, int64_t n, float lr,
                            float momentum, float grad_scale, void* stream);
/* Debug: device pointer of an internal fp32 NCHW tensor of the last step: "u0", "a0", "x0".."x4", "u1.<i>", "a1.<i>",
 * "u2.<i>", "z.<i>" (block i = 0..3: conv_r1 output, its BN+ReLU, conv_r2 output, residual sum before the pool),
 * "g", "gz", "gu" (backward scratch in its final state).  Valid until the next call on the handle. */
int nsc_trainer_debug_tensor(nsc_trainer* t, const char* name, float** ptr);
int nsc_trainer_destroy(nsc_trainer* t);

/* ---- candidate TSV decoder (host side; no GPU involved) -------------------------------------------------
 * Replaces extract_zlib / candidate_loader_tsv / extract_info_tsv (dataloader.py:38-123) and the tag parsing
 * of NeuSomaticDataset.__getitem__ (dataloader.py:267-274) for inference: mmaps a candidates_*.tsv
 * (generate_dataset.py:1569-1581), indexes its lines itself (the pickled .idx is not needed) and inflates a
 * range of candidates, multi-threaded, straight into the batch buffers nsc_score_host_u8 consumes. */
typedef struct nsc_tsv nsc_tsv;
int nsc_tsv_open(nsc_tsv** out, const char* path);
int64_t nsc_tsv_count(const nsc_tsv* t);
/* matrices [n,5,32,23] uint8; meta [n,3] int32 = (center, tumor_cov, normal_cov); anns255 [n,num_anns] fp32 =
 * float(double(text) * 255.0) or NULL when num_anns == 0; labels [n,2] int32 = (type index in
 * DEL,INS,NONE,SNP, min(length, 3) -- the two class labels of dataloader.py:55,415) or NULL; an unknown type fails like
 * the reference's KeyError.  num_threads <= 0: all hardware threads.  Fails (NSC_E_INVALID,
 * text in nsc_last_error) on malformed lines, like the reference's worker returning None. */
int nsc_tsv_decode(nsc_tsv* t, int64_t first, int64_t n, int num_anns, uint8_t* matrices, int32_t* meta,
                   float* anns255, int32_t* labels, int num_threads);
/* Staging for nsc_score_host_b64 -- everything of nsc_tsv_decode except the inflate: the lines of candidates
 * [first, first + n) are copied verbatim to text (text_cap >= nsc_tsv_range_bytes(t, first, n) + 16), spans [n,2] uint32
 * receive (offset in text, characters) of each image field, meta / anns255 / labels as above.  tag_buf (optional, with
 * tag_needed and tag_lens [n,2]): what nsc_tsv_tags_lens returns, gathered by the same threads in the same pass; when
 * tag_cap is too small the call fails with *tag_needed set.  tag_hashes (optional, [n]): 64-bit FNV-1a of each tag --
 * pairwise distinct hashes prove that no tag repeats, which lets the caller skip building a dictionary to find out. */
int nsc_tsv_stage(nsc_tsv* t, int64_t first, int64_t n, int num_anns, uint8_t* text, int64_t text_cap, uint32_t* spans,
                  int32_t* meta, float* anns255, int32_t* labels, int num_threads,
                  char* tag_buf, int64_t tag_cap, int64_t* tag_needed, int32_t* tag_lens, uint64_t* tag_hashes);
/* bytes of the lines of candidates [first, first + n) (-1: range out of bounds) */
int64_t nsc_tsv_range_bytes(const nsc_tsv* t, int64_t first, int64_t n);
/* matrices [n,5,32,23] of the listed candidates only (the images call.py:89-95 keeps for the called variants) */
int nsc_tsv_decode_rows(nsc_tsv* t, const int64_t* rows, int64_t n, uint8_t* matrices, int num_threads);
/* pointer into the mapped file + length of candidate i's tag (field 3 of the line; not NUL-terminated) */
int nsc_tsv_tag(const nsc_tsv* t, int64_t i, const char** tag, int32_t* len);
/* tags of candidates [first, first + n), each followed by '\n', copied into buf (cap bytes); *needed = bytes required
 * (call with buf = NULL to size the buffer).  One call instead of n: the tag strings are the keys of call.py's dictionaries. */
int nsc_tsv_tags(const nsc_tsv* t, int64_t first, int64_t n, char* buf, int64_t cap, int64_t* needed);
/* lens [n,2] int32 = (len(ref), len(alt)) of the tags of candidates [first, first + n): what call.py:343-350
 * (pred_vcf_records_none) branches on, so that the none.vcf filter can run on whole arrays */
int nsc_tsv_allele_lens(const nsc_tsv* t, int64_t first, int64_t n, int32_t* lens);
/* both of the above in one pass over the lines (the filing stage of the call pipeline); the tags are copied without their
 * directory part (path_.split("/")[-1], call.py:88) */
int nsc_tsv_tags_lens(const nsc_tsv* t, int64_t first, int64_t n, char* buf, int64_t cap, int64_t* needed, int32_t* lens);
/* Test hook: inflates ONE zlib stream (RFC 1950) of exactly out_len bytes with the decoder's own fast inflate
 * (csrc/nsc_inflate.h; NSC_E_STATE when it declines the stream and the decoder would fall back to zlib) or, use_zlib != 0,
 * with zlib itself (NSC_E_INVALID on a malformed stream). */
int nsc_debug_inflate(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t out_len, int use_zlib);
/* Test hook: the device decoder of nsc_score_host_b64 alone -- base64 + zlib inflate of n image fields on `device`;
 * matrices_host [n,3680] (rows of declined candidates are zero), declined_host [n] = 1 where the device decoder left the
 * candidate to the host. */
int nsc_debug_inflate_b64(int device, const uint8_t* text_host, int64_t text_bytes, const uint32_t* spans_host, int64_t n,
                          uint8_t* matrices_host, int32_t* declined_host);
int nsc_tsv_close(nsc_tsv* t);

#ifdef __cplusplus
}
#endif
#endif /* NEUSOMATIC_CUDA_H_ */
