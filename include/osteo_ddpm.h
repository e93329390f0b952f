/*
 * osteo_ddpm.h — C-ABI of the B200-native conditional-DDPM hot path.
 *
 * The reference (rare-resilience-ai/Osteosarcoma_DiffusionModel) is pure Python and has no
 * FFI / plugin layer; its boundary for this path is the class
 * `BiologyAwareDiffusionModel` (models/diffusion.py:259-449) and three validators
 * (utils/validation.py:125-175, :177-223, :273-298).  Each entry point below names the
 * reference interface it replaces.  The Python mirror of that class
 * (osteosarcoma_diffusionmodel_b200/diffusion.py) binds these symbols with ctypes.
 *
 * Conventions
 *   - extern "C", plain ints / floats / pointers; no torch types.
 *   - every `*_dev` pointer is a DEVICE pointer (row-major, dense unless a pitch is given);
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - returns 0 on success, negative on error; osteo_last_error() gives the message.
 *   - calls enqueue work on `stream` and do not synchronise unless stated.
 *   - the library owns its workspace (repacked weights, activations, padded state);
 *     it never allocates or frees caller-visible memory and never keeps caller pointers
 *     beyond the call, except weights which are repacked (copied) in set_weights.  Cached executable
 *     graphs (set_weights, train_step) record the addresses they were captured with and are replayed
 *     only by a later call that passes exactly those addresses again; any other call runs eagerly.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef OSTEO_DDPM_H
#define OSTEO_DDPM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct osteo_ddpm_ctx osteo_ddpm_ctx;

/* Precision of the tensor-core contractions. */
enum {
    OSTEO_PREC_BF16 = 0,   /* bf16 operands, fp32 accumulate (throughput mode)                     */
    OSTEO_PREC_FP32X3 = 1  /* split-bf16 hi/lo, 3 tcgen05 passes, ~fp32 accuracy (parity mode)     */
};

/* Number of weight tensors expected by osteo_ddpm_set_weights for `n_hidden` hidden dims:
 * 4 (condition_embed) + 8 (input/cond/time/output proj) + 8 per block, blocks = 2*(n_hidden-1)+1. */
int osteo_ddpm_num_weight_tensors(int n_hidden);

const char* osteo_last_error(void);
int osteo_version(void);
/* Number of CUDA devices visible (0 when none; never fails). */
int osteo_device_count(void);

/* ---- model context: replaces BiologyAwareDiffusionModel.__init__ (models/diffusion.py:264-310)
 * data_dim = mutation_dim + expression_dim + pathway_dim; cond_dim = condition_dim;
 * time_dim = config.model.latent_dim; cond_embed_dim = 64 (models/diffusion.py:285);
 * hidden_dims = config.model.hidden_dims (each a multiple of 128, <= 1024); num_steps = T. */
int osteo_ddpm_create(osteo_ddpm_ctx** out, int device, int data_dim, int cond_dim, int time_dim,
                      int cond_embed_dim, int n_hidden, const int* hidden_dims, int num_steps,
                      float dropout_p, int precision);
int osteo_ddpm_destroy(osteo_ddpm_ctx* ctx);

/* Make room for `rows` patients in the workspace (grows only; synchronises the device). */
int osteo_ddpm_reserve(osteo_ddpm_ctx* ctx, long long rows);
long long osteo_ddpm_capacity(const osteo_ddpm_ctx* ctx);
/* Bytes of device memory the context owns. */
long long osteo_ddpm_workspace_bytes(const osteo_ddpm_ctx* ctx);
/* Rows per internal pass (activations of one pass stay L2-resident). 0 = all rows at once. */
int osteo_ddpm_set_chunk_rows(osteo_ddpm_ctx* ctx, int chunk_rows);
/* Sampling graphs split the batch into `branches` (1..4, default 2) contiguous row ranges captured as parallel graph branches
 * (rows are independent: GroupNorm is per row, models/diffusion.py:202), so one branch's kernels fill the tail wave and launch gaps
 * of the other's. Results do not depend on the setting. Batches with fewer than two waves of row tiles per branch use fewer. */
int osteo_ddpm_set_branches(osteo_ddpm_ctx* ctx, int branches);
int osteo_ddpm_set_precision(osteo_ddpm_ctx* ctx, int precision);
/* bf16 mode with hidden_dims[0] <= 256 runs the FUSED reverse step by default: output_proj, the reverse update
 * (models/diffusion.py:400-423) and the NEXT step's input_proj + embedding add (models/diffusion.py:229-232) in one
 * kernel, so the fp32 state is read once and written once per step and no bf16 copy of it goes through HBM.
 * enable = 0 selects the unfused kernels (always used in FP32X3 mode). Takes effect at the next load_state /
 * init_noise. osteo_ddpm_step_is_fused reports which path the current settings select. */
int osteo_ddpm_set_fused(osteo_ddpm_ctx* ctx, int enable);
int osteo_ddpm_step_is_fused(const osteo_ddpm_ctx* ctx);
/* Row branches of the cached sampling graph (0 = no graph has been captured yet): how the last graph-replayed
 * osteo_ddpm_sample_loop actually ran (tests assert the benchmarked configuration was exercised). */
int osteo_ddpm_graph_branches(const osteo_ddpm_ctx* ctx);

/* ---- parameters: replaces nn.Module.load_state_dict / optimizer updates.
 * `weights_dev[i]` are fp32 device tensors in state_dict order (SURVEY.md §8a layer table):
 *   condition_embed.mlp.0.{weight,bias}, condition_embed.mlp.2.{weight,bias},
 *   unet.input_proj.{weight,bias}, unet.cond_proj.{weight,bias}, unet.time_proj.{weight,bias},
 *   for each block in encoder.0.., bottleneck, decoder.0..: {0.weight,0.bias,1.weight,1.bias,
 *   4.weight,4.bias,5.weight,5.bias}, unet.output_proj.{weight,bias}.
 * Repacks to bf16 [hi|lo], K padded to 64, rows padded to 128, and rebuilds the 1000x256
 * time_proj table (models/diffusion.py:222-223 hoisted out of the loop). */
int osteo_ddpm_set_weights(osteo_ddpm_ctx* ctx, const float* const* weights_dev, int n_tensors, void* stream);

/* Schedule tables, all fp32 [num_steps] on the HOST (models/diffusion.py:299-310, :401-423):
 * sqrt_ab / sqrt_1mab are the registered buffers; coef_x / coef_eps / sigma are the collapsed
 * reverse-step coefficients derived from those buffers in fp64 by the host mirror. */
int osteo_ddpm_set_schedule(osteo_ddpm_ctx* ctx, const float* sqrt_ab, const float* sqrt_1mab,
                            const float* coef_x, const float* coef_eps, const float* sigma);
/* Sinusoidal embedding table fp32 [num_steps, time_dim] on the HOST: TimeEmbedding.forward
 * (models/diffusion.py:124-139) evaluated at t/num_steps for every integer t. */
int osteo_ddpm_set_time_embedding(osteo_ddpm_ctx* ctx, const float* emb_host);

/* ---- state I/O */
/* x_dev [n, data_dim] fp32 -> internal padded fp32 state + bf16 shadow. */
int osteo_ddpm_load_state(osteo_ddpm_ctx* ctx, const float* x_dev, long long n, void* stream);
/* internal state -> out_dev [n, data_dim] fp32. */
int osteo_ddpm_store_state(osteo_ddpm_ctx* ctx, float* out_dev, long long n, void* stream);
/* internal state -> the pieces SyntheticPatientGenerator.generate makes on the host (utils/generate.py:130-135), any of them NULL to
 * skip: calls_dev uint8 [n, mutation_dim] = (x[:, :mutation_dim] > threshold); call_bits_dev uint8 [n, ceil(mutation_dim / 8)] the
 * same calls one bit per gene (LSB first); rest_dev fp32 [n, data_dim - mutation_dim] = expression | pathway scores. The comparison
 * is made on the fp32 state, i.e. on exactly the values osteo_ddpm_store_state returns. */
int osteo_ddpm_store_split(osteo_ddpm_ctx* ctx, long long n, int mutation_dim, float threshold, uint8_t* calls_dev,
                           uint8_t* call_bits_dev, float* rest_dev, void* stream);
/* x_T ~ N(0, I) from Philox4x32-10(seed; row_base + row, column): models/diffusion.py:443. */
int osteo_ddpm_init_noise(osteo_ddpm_ctx* ctx, long long n, uint64_t seed, long long row_base, void* stream);
/* ConditionalEmbedding + cond_proj hoisted out of the step loop (models/diffusion.py:107-114,
 * :226, :395): cond_dev [n, cond_dim] fp32. */
int osteo_ddpm_set_conditions(osteo_ddpm_ctx* ctx, const float* cond_dev, long long n, void* stream);

/* ---- reverse process: replaces p_sample (models/diffusion.py:382-425) and the loop of sample (:446-447).
 * One step at integer timestep t on the internal state. noise_dev: optional injected z [n, data_dim]
 * (parity runs); NULL = in-kernel Philox keyed by (seed, row_base + row, t, column).
 * eps_out_dev: optional fp32 [n, data_dim] copy of the predicted noise. */
int osteo_ddpm_reverse_step(osteo_ddpm_ctx* ctx, long long n, int t, const float* noise_dev,
                            float* eps_out_dev, uint64_t seed, long long row_base, void* stream);
/* Steps t_start, t_start-1, ..., t_end (inclusive) with in-kernel noise; one CUDA graph per step shape,
 * replayed. noise_dev: optional injected z for ALL steps, [t_start - t_end + 1, n, data_dim] in loop order
 * (the z of step t == 0 is never read). */
int osteo_ddpm_sample_loop(osteo_ddpm_ctx* ctx, long long n, int t_start, int t_end, const float* noise_dev,
                           uint64_t seed, long long row_base, int use_graph, void* stream);

/* ---- denoiser forward: replaces DiffusionUNet.forward via BiologyAwareDiffusionModel.forward(
 * return_loss=False) (models/diffusion.py:210-256, :367-380) on caller data:
 * xt_dev [n, data_dim] fp32, t_idx_dev [n] int32 (integer timesteps), conditions already set;
 * eps_out_dev [n, data_dim] fp32. Eval mode (no dropout). */
int osteo_ddpm_denoise(osteo_ddpm_ctx* ctx, const float* xt_dev, const int* t_idx_dev, long long n,
                       float* eps_out_dev, void* stream);

/* ---- forward process: replaces q_sample (models/diffusion.py:328-342).
 * xt = sqrt_ab[t]*x0 + sqrt_1mab[t]*noise per row; noise_dev in/out: if gen_noise != 0 it is FILLED
 * from Philox (stream QNOISE, step = `salt`) first. All [n, data_dim] fp32, t_idx_dev [n] int32. */
int osteo_ddpm_q_sample(osteo_ddpm_ctx* ctx, const float* x0_dev, const int* t_idx_dev, float* noise_dev,
                        float* xt_dev, long long n, int gen_noise, uint64_t seed, long long row_base,
                        uint32_t salt, void* stream);

/* ---- standalone elementwise reverse update (the epilogue of reverse_step as its own kernel):
 * x <- coef_x[t]*x - coef_eps[t]*eps + sigma[t]*z, dense [n, data_dim] fp32 tensors. z_dev may be NULL
 * (Philox). Replaces models/diffusion.py:400-423. */
int osteo_ddpm_reverse_update(osteo_ddpm_ctx* ctx, float* x_dev, const float* eps_dev, const float* z_dev,
                              long long n, int t, uint64_t seed, long long row_base, void* stream);

/* ---- training step: replaces forward(return_loss=True) + loss.backward() (models/diffusion.py:344-378,
 * utils/train.py:236-239).  x0_dev [n, data_dim], cond_dev [n, cond_dim] fp32.
 * t_idx_dev / noise_dev / drop_masks_dev are optional injections (NULL = Philox streams keyed by seed).
 * drop_masks_dev: array of n_blocks uint8 keep-masks [n, block_out_dim] (device pointers, host array).
 * loss_dev: fp32 scalar on device. grads_dev: n_tensors fp32 device tensors shaped like the weights
 * (overwritten, not accumulated); NULL = forward only. train != 0 enables dropout. */
int osteo_ddpm_train_step(osteo_ddpm_ctx* ctx, const float* x0_dev, const float* cond_dev, long long n,
                          const int* t_idx_dev, const float* noise_dev, const uint8_t* const* drop_masks_dev,
                          int train, uint64_t seed, long long row_base, float* loss_dev,
                          float* const* grads_dev, int n_tensors, void* stream);

/* The same step in two halves, for auxiliary losses on the PREDICTED CLEAN SAMPLE x0hat = (x_t - sqrt(1-ab_t) eps_hat) / sqrt(ab_t)
 * (models/diffusion.py:401-403) -- the diffusion analogue of the multi-task terms of models/cvae.py:304-341:
 *   train_forward   forward + MSE loss, keeps what the backward pass needs (arguments as osteo_ddpm_train_step);
 *   train_x0hat     out_dev [n, n_cols] = x0hat[:, cols_dev[0..n_cols)]  (x0_dev / t_idx_dev: the tensors given to train_forward);
 *   train_inject    d(loss)/d(eps_hat) += chain rule of g_dev [n, n_cols] = d(aux loss)/d(x0hat[:, cols]); columns must be distinct;
 *   train_backward  the backward pass into grads_dev (same train / seed / row_base / masks as train_forward). */
int osteo_ddpm_train_forward(osteo_ddpm_ctx* ctx, const float* x0_dev, const float* cond_dev, long long n, const int* t_idx_dev,
                             const float* noise_dev, const uint8_t* const* drop_masks_dev, int train, uint64_t seed,
                             long long row_base, float* loss_dev, void* stream);
int osteo_ddpm_train_x0hat(osteo_ddpm_ctx* ctx, const float* x0_dev, const int* t_idx_dev, long long n, const int* cols_dev, int n_cols,
                           float* out_dev, void* stream);
int osteo_ddpm_train_inject(osteo_ddpm_ctx* ctx, const int* t_idx_dev, long long n, const int* cols_dev, int n_cols,
                            const float* g_dev, void* stream);
int osteo_ddpm_train_backward(osteo_ddpm_ctx* ctx, const float* cond_dev, long long n, const int* t_idx_dev,
                              const uint8_t* const* drop_masks_dev, int train, uint64_t seed, long long row_base,
                              float* const* grads_dev, int n_tensors, void* stream);

/* The backward pass cut in two launches, so that a data-parallel caller can all-reduce the gradients of the first part while the second
 * runs (the reference has one loss.backward() and no data parallelism: utils/train.py:236-244). part 1: loss gradient, output_proj and
 * the Linear+GroupNorm half blocks j >= cut = the gradient tensors [10 + 4 * cut, n_tensors); part 2: the rest. Both after
 * osteo_ddpm_train_forward, part 1 first; together they enqueue the kernels of osteo_ddpm_train_backward in the same order. */
int osteo_ddpm_train_backward_part(osteo_ddpm_ctx* ctx, const float* cond_dev, long long n, const int* t_idx_dev,
                                   const uint8_t* const* drop_masks_dev, int train, uint64_t seed, long long row_base,
                                   float* const* grads_dev, int n_tensors, int part, int cut, void* stream);

/* With nothing injected (noise_dev == NULL, drop_masks_dev == NULL) and gradients requested, everything of the step after the two
 * kernels that read x0_dev / cond_dev is replayed as ONE executable graph from the second call with the same n and gradient
 * addresses on (keep the gradient tensors persistent to benefit); enable = 0 keeps every call eager. Default 1. */
int osteo_ddpm_set_train_graph(osteo_ddpm_ctx* ctx, int enable);

/* ---- optimiser step: replaces torch.nn.utils.clip_grad_norm_(params, max_norm) + torch.optim.AdamW.step() (utils/train.py:169-173,
 * :242-244) for a list of fp32 tensors with two launches. create takes the element counts (host array); step takes HOST arrays of
 * DEVICE pointers (params, grads, exp_avg, exp_avg_sq; re-uploaded only when an address changed), torch's hyper-parameters (doubles, like torch's Python scalars: the derived constants are rounded to fp32 once), the
 * 1-based step count, and max_norm (<= 0: no clipping). The gradients are scaled in place by min(1, max_norm / (||g||_2 + 1e-6))
 * exactly as clip_grad_norm_ does; norm_out_dev (optional, fp32 scalar on the device) receives the total norm before clipping. */
typedef struct osteo_adamw osteo_adamw;
int osteo_adamw_create(osteo_adamw** out, int n_tensors, const long long* numel_host);
int osteo_adamw_destroy(osteo_adamw* h);
int osteo_adamw_step(osteo_adamw* h, float* const* params_dev, float* const* grads_dev, float* const* exp_avg_dev,
                     float* const* exp_avg_sq_dev, double lr, double beta1, double beta2, double eps, double weight_decay,
                     long long step, double max_norm, float* norm_out_dev, void* stream);

/* Allocate the transposed weight copies the backward pass contracts against. Must be followed by
 * osteo_ddpm_set_weights (which fills them) before osteo_ddpm_train_step is asked for gradients. */
int osteo_ddpm_enable_training(osteo_ddpm_ctx* ctx, int enable);

/* ---- diagnostics */
/* One reverse step (in-kernel noise) with a CUDA event recorded on `stream` after every GEMM launch.
 * Writes the per-launch durations in ms to ms_out_host[0..] in launch order (per row chunk: input_proj,
 * the 2*blocks Linear+GroupNorm+SiLU launches, output_proj+update) and RETURNS the number of launches
 * (>= 0) or a negative error. Synchronises `stream`. Used by bench.py for the live roofline numbers. */
int osteo_ddpm_profile_step(osteo_ddpm_ctx* ctx, long long n, int t, uint64_t seed, long long row_base,
                            float* ms_out_host, int max_out, void* stream);
/* Sticky kernel status word (0 = ok; see GemmError). Synchronises `stream`. */
int osteo_ddpm_status(osteo_ddpm_ctx* ctx, void* stream);
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
long long osteo_ddpm_launch_count(const osteo_ddpm_ctx* ctx);

/* ---- building blocks exposed for unit tests / the validators (context-free) */
/* out[m, n] = a[m, :k] . w[n, :k] + bias[n]  on tcgen05; a, w fp32 device tensors (converted to bf16,
 * or split-bf16 when precision = FP32X3). out fp32 [m, n]. */
int osteo_linear_tc(const float* a_dev, const float* w_dev, const float* bias_dev, float* out_dev,
                    int m, int n, int k, int precision, void* stream);
/* Same contraction followed by GroupNorm(8) + affine + SiLU (n multiple of 128, n/8 in {16,32,64}). */
int osteo_linear_gn_silu_tc(const float* a_dev, const float* w_dev, const float* bias_dev,
                            const float* gamma_dev, const float* beta_dev, float* out_dev,
                            int m, int n, int k, int precision, void* stream);
/* dw[n_out, k_in] = dy^T x over `rows` batch rows (dy fp32 [rows, n_out], x fp32 [rows, k_in]) on tcgen05 with
 * both operands MN-major and the batch split over CTAs — the weight-gradient contraction of train_step.
 * Synchronises `stream`. */
int osteo_wgrad_tc(const float* dy_dev, const float* x_dev, float* dw_dev, long long rows, int n_out, int k_in,
                   int precision, void* stream);
/* Fill out[n, d] fp32 with Philox normals (stream id, step) — test hook for the RNG. */
int osteo_philox_normal(float* out_dev, long long n, int d, uint64_t seed, long long row_base,
                        uint32_t stream_id, uint32_t step, void* stream);
/* Raw Philox words: out[n, 4*ncol4] uint32. */
int osteo_philox_words(uint32_t* out_dev, long long n, int ncol4, uint64_t seed, long long row_base,
                       uint32_t stream_id, uint32_t step, void* stream);

/* ---- validators
 * RBF-MMD partial sums: replaces the three cdist Grams of BiologicalValidator.compute_mmd
 * (utils/validation.py:284-296).  x_dev [n, d], y_dev [m, d] fp32 device; this call handles Gram ROWS
 * [row_begin, row_end) of K(X,X) and K(X,Y) and rows [yrow_begin, yrow_end) of K(Y,Y) (row sharding
 * across GPUs) and writes sums_dev[3] = {sum Kxx, sum Kyy, sum Kxy} (fp64, device,
 * overwritten).  Operands are centred by `center_dev` [d] (RBF is translation invariant). */
int osteo_mmd_partial(const float* x_dev, long long n, const float* y_dev, long long m, int d, float gamma,
                      const float* center_dev, long long row_begin, long long row_end,
                      long long yrow_begin, long long yrow_end, int precision, double* sums_dev, void* stream);
/* Same reduction with the Gram rows sharded BLOCK-CYCLICALLY over `world` ranks (BASELINE.json configs[4]: Gram rows over 8 GPUs):
 * this rank reduces the 128-row blocks b with b % world == rank of K(X,X), K(Y,Y) and K(X,Y), and K(X,X) / K(Y,Y) as symmetric
 * half-Grams (tiles below the diagonal skipped, off-diagonal ones counted twice), so the triangular work is balanced over the ranks and
 * the partial sums of all ranks add up (NCCL all-reduce of 3 fp64 on the caller's side) to the sums of the whole Grams. */
int osteo_mmd_partial_cyclic(const float* x_dev, long long n, const float* y_dev, long long m, int d, float gamma,
                             const float* center_dev, int rank, int world, int precision, double* sums_dev, void* stream);
/* Column-gathered moment blocks for Pearson correlations: replaces DataFrame.corr / Series.corr
 * (utils/validation.py:152,156,206).  data_dev [n, ld] fp32; cols_dev [k] int32 column indices (k <= 32);
 * shift_dev [k] per-column shift (e.g. a first-row estimate; improves conditioning);
 * out_dev fp64 [1 + k + k*k] = {count, sum (x-s), sum (x-s)(x-s)^T} over rows [row_begin,row_end),
 * overwritten. */
int osteo_corr_moments(const float* data_dev, long long n, int ld, const int* cols_dev, int k,
                       const float* shift_dev, long long row_begin, long long row_end,
                       double* out_dev, void* stream);
/* The same moment blocks for up to 32 column sets in ONE pass over the rows (validate_pathway_coherence, utils/validation.py:144-157:
 * ten pathways x ~15 genes out of the same 371 columns).  data_dev [n, ld] fp32 of which the first `ncols` columns are staged;
 * cols_dev int32 [n_sets, 32], each row = the set's column indices (< ncols) packed at the front, -1 padded; shift_dev fp32 [n_sets, 32];
 * out_dev fp64 [n_sets, 1057] = per set {count, sum (x-s) [32], sum (x-s)(x-s)^T [32][32]} over rows [row_begin,row_end), overwritten. */
int osteo_corr_moments_batched(const float* data_dev, long long n, int ld, int ncols, const int* cols_dev, int n_sets,
                               const float* shift_dev, long long row_begin, long long row_end, double* out_dev, void* stream);
/* Same moment blocks (same output layout), register-tiled: the HBM-rate path behind validate_pathway_coherence
 * (utils/validation.py:144-157). max_set_size = the largest number of columns in a set (<= 32; sets of <= 16 columns take the 10-block
 * form, 25 sets per pass over the rows; larger ones the 36-block form, 7 sets per pass). shift_dev may be NULL: a column's shift is then
 * its value in row 0 of data_dev (every rank holds the whole cohort). cols_dev int32 [n_sets, 32], -1 padded, valid entries first. */
int osteo_corr_moments_tiled(const float* data_dev, long long n, int ld, int ncols, const int* cols_dev, int n_sets, int max_set_size,
                             const float* shift_dev, long long row_begin, long long row_end, double* out_dev, void* stream);
/* Per-set pathway-coherence score = mean over i < j of the Pearson correlation, float64, from moment blocks [n_sets, 1 + 32 + 32*32]
 * (the upper-triangle mean of DataFrame.corr(), utils/validation.py:152-157) -> scores_dev fp64 [n_sets]. Moment blocks of several
 * ranks are summed (all-reduced) BEFORE this call. */
int osteo_coherence_finish(const double* moments_dev, const int* cols_dev, int n_sets, double* scores_dev, void* stream);

/* ---- training ingress (utils/train.py:22-126, :204-227): one batch of a DEVICE-RESIDENT dataset, gathered by row index,
 * with MixupAugmentation fused in:  out[i, :] = lam * src[idx_a[i], :] + one_minus_lam * src[idx_b[i], :]   (src_dev [src_rows, d] fp32,
 * idx_*_dev int64 [n] device, out_dev [n, d]).  idx_a_dev == NULL: rows 0..n-1; idx_b_dev == NULL: plain gather, lam ignored.
 * lam / one_minus_lam are the fp32 roundings of the Python scalars lam and 1 - lam; products and sum are rounded separately, so the
 * result is bit-identical to the reference's `lam * data + (1 - lam) * data[index]` (utils/train.py:118). */
int osteo_mixup_rows(const float* src_dev, long long src_rows, int d, const long long* idx_a_dev, const long long* idx_b_dev,
                     long long n, float lam, float one_minus_lam, float* out_dev, void* stream);

/* ---- differentiable biology losses (the reference has only stubs: models/cvae.py:262-302).
 * Per column set s of data_dev [n, ld] (cols_dev [n_sets][32] int32, -1 padded; the first `ncols` columns are staged):
 *   modes_dev[s] == 0      loss_s = 1 - mean_{i<j} Pearson(x_i, x_j)   = 1 - the per-pathway score of validate_pathway_coherence
 *                                                                        (utils/validation.py:150-157)
 *   modes_dev[s] == +1/-1  loss_s = max(0, -mode * Pearson(x_0, x_1))  > 0 exactly when validate_mutation_expression_correlation
 *                                                                        (utils/validation.py:206-214) counts a violation
 * finish: moments_dev fp64 [n_sets * 1057] = the output of osteo_corr_moments_batched over the same sets (all-reduced over the ranks
 * of a data-parallel job, so that the loss is the GLOBAL batch's); loss_out_dev fp32 [n_sets]; coef_out_dev fp32 [n_sets * 32 * 4]
 * (saved for backward).  backward: grad_dev [n, ld] += upstream_dev[s] * d loss_s / d data over this rank's rows (accumulates). */
int osteo_corr_loss_finish(const double* moments_dev, const int* cols_dev, const float* shift_dev, const int* modes_dev, int n_sets,
                           float* loss_out_dev, float* coef_out_dev, void* stream);
int osteo_corr_loss_backward(const float* data_dev, long long n, int ld, const int* cols_dev, int n_sets,
                             const float* coef_dev, const float* upstream_dev, float* grad_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSTEO_DDPM_H */
