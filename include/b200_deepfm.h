/*
 * b200_deepfm.h -- C ABI of the fused DeepFM tower (part of libb200ps.so).
 *
 * Replaces, for the dac_ctr DeepFM workload of BASELINE.json configs[1], what the
 * reference runs in TensorFlow between the pull and the push of one minibatch
 * (paths relative to /root/reference/):
 *   - tf.gather(batch_embedding, idx)      elasticdl/python/elasticdl/embedding_delegate.py:95
 *   - model.call                           model_zoo/dac_ctr/deepfm_model.py:61-109 (DNN[16,4], FM, linear)
 *   - tape.gradient(loss, vars + BETs)     elasticdl/python/worker/ps_trainer.py:391-400
 *   - the sum of deduplicate_indexed_slices elasticdl/python/common/tensor_utils.py:39-60
 *     (the per-occurrence embedding gradients are reduced per unique id on the fly)
 * in two launches that read every embedding row twice (L2) and write every
 * per-unique-id gradient once.
 *
 * Fixed architecture (dac_ctr): 13 dense features, deep dim 8, DNN 16 -> 4 -> 1,
 * G id groups (38 for dac_ctr), one id per group per sample.
 * Parameter / gradient layout (floats), all torch.nn.Linear layouts:
 *   [ w_dense 13 (+3 pad) | w1 16 x (13 + 8G) | b1 16 | w2 4 x 16 | b2 4 | w3 4 ]
 */
#ifndef B200_DEEPFM_H_
#define B200_DEEPFM_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200_DEEPFM_NDENSE 13
#define B200_DEEPFM_DIM 8
#define B200_DEEPFM_H1 16
#define B200_DEEPFM_H2 4
#define B200_DEEPFM_SCRATCH 44 /* floats of per-sample backward state */

typedef struct {
  int32_t G;               /* id groups */
  int32_t B;               /* samples in the batch */
  const int32_t* inv;      /* [G][B]  rank of each sample's id among its group's unique ids */
  const int32_t* n_unique; /* [G]     distinct ids per group (device) */
  const float* bet_wide;   /* [G][B]      pulled wide rows (first n_unique[g] valid) */
  const float* bet_deep;   /* [G][B][8]   pulled deep rows */
  const float* dense;      /* [B][13] */
  const float* labels;     /* [B] */
  const float* params;     /* flat parameter buffer, layout above */
  float* grads;            /* flat gradient buffer, same layout (overwritten) */
  float* gsum_wide;        /* [G][B]     d loss / d pulled rows, per unique id (overwritten) */
  float* gsum_deep;        /* [G][B][8] */
  float* loss;             /* [1] mean binary cross entropy with logits */
  float* logits;           /* [B] or NULL */
  float* scratch;          /* B*B200_DEEPFM_SCRATCH + 16*(13+8G) floats: per-sample backward state, then W1^T */
} b200_deepfm_args_t;

size_t b200_deepfm_param_count(int G);
/* forward + backward of one batch; asynchronous on `stream` (cudaStream_t). */
int b200_deepfm_fwd_bwd(const b200_deepfm_args_t* args, void* stream);
/* Same contract as b200_deepfm_fwd_bwd, computed by the tensor-core tower (csrc/deepfm_tower_mma.cu): the
 * rows of a chunk of samples are gathered once into a shared-memory tile and the three first-layer
 * contractions run as 3xTF32 mma.sync.  `scratch` is not used. */
int b200_deepfm_fwd_bwd_mma(const b200_deepfm_args_t* args, void* stream);
int64_t b200_deepfm_mma_launch_count(void);
/* Same contract, computed by the TILE tower (csrc/deepfm_tower2.cu, round 2, the default): a CTA gathers the
 * rows of 32 samples once into shared memory (cp.async) and runs forward, backward and the
 * parameter-gradient contraction from that tile -- two launches (prep + tile kernel).  `scratch` needs
 * 320*16 + 16 floats (W1 in tile column order, then the dynamic tile counter); at most 38 id groups. */
int b200_deepfm_fwd_bwd_tile(const b200_deepfm_args_t* args, void* stream);
int b200_deepfm_forward_tile(const b200_deepfm_args_t* args, void* stream);
/* b200_deepfm_fwd_bwd_tile in its two halves, so that a caller can put the prologue on another stream, beside
 * the row pull: the prologue (W1 into tile order; zero the gradient buffers, the loss and the live rows of the
 * per-unique-id sums) reads only `params` and `n_unique`; the main kernel needs the prologue and the pulled rows. */
/* The step's loss to the host without a copy-engine operation: ring_host_mapped is pinned host memory as the
 * device sees it (cudaHostAlloc / UVA); entry cursor % ring_len receives *loss_dev, then the cursor advances. */
int b200_deepfm_publish_loss(const float* loss_dev, float* ring_host_mapped, int ring_len, unsigned* cursor_dev, void* stream);
int b200_deepfm_tile_prologue(const b200_deepfm_args_t* args, void* stream);
int b200_deepfm_tile_main(const b200_deepfm_args_t* args, void* stream);
int64_t b200_deepfm_tile_launch_count(void);
const char* b200_deepfm_tile_last_error(void);
/* forward only (logits), for evaluation. */
int b200_deepfm_forward(const b200_deepfm_args_t* args, void* stream);
/* kernels launched by the two calls above so far */
int64_t b200_deepfm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
