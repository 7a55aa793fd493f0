/*
 * taco_b200.h -- C ABI of the B200-native Tacotron synthesis forward path.
 *
 * The reference (Jim-Song/tacotron_multispeaker) is pure Python on TensorFlow
 * 1.x and has no FFI of its own; the boundary it offers for this path is the
 * Python contract
 *     Tacotron.initialize(inputs, input_lengths, mel_targets, linear_targets,
 *                         identities, id_num)          models/tacotron.py:18
 *     Synthesizer.load / Synthesizer.synthesize       synthesizer.py:14,37
 * whose body is one tf.Session.run (synthesizer.py:47).  This header is what a
 * binding for that body links against: each entry point below names the
 * reference lines it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - plain C, no torch / CUDA types in signatures; `stream` is a cudaStream_t
 *    passed as void* (NULL = legacy default stream).
 *  - every data pointer is a DEVICE pointer unless the name ends in `_host`.
 *  - all work is enqueued on the caller's stream; the only host syncs are the
 *    ones documented (`steps_out_host` read-back, workspace growth).
 *  - return value: 0 = TACO_OK, negative = error; taco_last_error() gives text.
 *  - one handle per GPU; a handle is not thread-safe.
 *  - tensors are dense row-major float32 / int32 with the reference's layouts:
 *      inputs [N,T_in] i32, input_lengths [N] i32, identities [N] i32,
 *      mel_targets [N,T_tgt,num_mels], mel_outputs [N,max_steps*r,num_mels],
 *      linear_outputs [N,max_steps*r,num_freq], alignments [N,T_in,max_steps]
 *    where max_steps = taco_max_steps(...).  Only the first `steps` decoder
 *    steps (steps*r frames) of each utterance are written.
 */
#ifndef TACO_B200_H_
#define TACO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct taco_handle taco_handle;

/* Subset of reference hparams.py:5-53 that the forward path reads. */
typedef struct taco_hparams {
  int32_t num_mels;                /* hparams.py:11  (80)   */
  int32_t num_freq;                /* hparams.py:12  (1025) */
  int32_t outputs_per_step;        /* hparams.py:22  r      */
  int32_t max_iters;               /* hparams.py:34         */
  int32_t embedding_text_channels; /* hparams.py:39  (256)  */
  int32_t embedding_id_channels;   /* hparams.py:40  (64)   */
  int32_t num_symbols;             /* len(symbols2), models/tacotron.py:40 (7352) */
  int32_t id_num;                  /* rows of embedding_id; <=1 = single speaker (tacotron.py:48) */
} taco_hparams;

enum {
  TACO_OK = 0,
  TACO_ERR_INVALID = -1,        /* bad argument / shape                         */
  TACO_ERR_CUDA = -2,           /* a CUDA runtime call failed                   */
  TACO_ERR_MISSING_WEIGHT = -3, /* finalize: a variable was never set           */
  TACO_ERR_STATE = -4,          /* call order (e.g. forward before finalize)    */
  TACO_ERR_OOB_ID = -5,         /* symbol / speaker id outside its table        */
  TACO_ERR_UNSUPPORTED = -6     /* device is not sm_100 / shape beyond limits   */
};

enum { TACO_BN_MOVING = 0, TACO_BN_BATCH = 1 };   /* tf.layers.batch_normalization(training=) */
enum { TACO_ACT_NONE = 0, TACO_ACT_RELU = 1, TACO_ACT_SIGMOID = 2, TACO_ACT_TANH = 3 };
enum { TACO_CBHG_ENCODER = 0, TACO_CBHG_POST = 1 };

/* ---- lifetime ----------------------------------------------------------- */
/* Replaces create_model()+Tacotron.__init__ (models/__init__.py:4-8,
 * models/tacotron.py:14-15).  Fails with TACO_ERR_UNSUPPORTED when `device`
 * is not a compute-capability-10.x GPU: there is no CPU fallback. */
int taco_create(const taco_hparams* hp, int device, taco_handle** out);
int taco_destroy(taco_handle* h);
const char* taco_last_error(const taco_handle* h);
const char* taco_version(void);

/* ---- weights ------------------------------------------------------------ */
/* Replaces Saver.restore (synthesizer.py:33-34): one call per variable, named
 * as in the TF checkpoint with or without the "model/inference/" prefix
 * (SURVEY.md Appendix A).  `data_host` is HOST float32, copied immediately. */
int taco_set_weight(taco_handle* h, const char* tf_name, const float* data_host,
                    const int64_t* shape, int ndim);
/* Packs the variables into kernel layouts (BN folded to scale/shift, GRU
 * kernels split into input/recurrent parts, decoder matrices sliced per CTA of
 * the decode cluster) and uploads them.  Must be called once after all
 * taco_set_weight calls and again after any later taco_set_weight. */
int taco_finalize_weights(taco_handle* h);
/* Number of variables the path expects / name of the i-th one. */
int taco_num_weights(const taco_handle* h);
const char* taco_weight_name(const taco_handle* h, int i);

/* ---- shapes ------------------------------------------------------------- */
/* Decoder steps the loop can take: min(max_iters, T_tgt / r) when teacher
 * forcing (helpers.py:48-55,73), else max_iters (tacotron.py:94). */
int taco_max_steps(const taco_handle* h, int teacher_force, int T_tgt);

/* ---- the whole path: body of Tacotron.initialize (tacotron.py:35-104) ---- */
/* spk may be NULL (single-speaker branch, tacotron.py:56-58); mel_targets may
 * be NULL unless teacher_force.  bn_mode / teacher_force are separate because
 * the reference ties both to `linear_targets is not None` (tacotron.py:36) and
 * the host layer decides.  linear_out / align_out may be NULL to skip the
 * post-net / the alignment stack.  Writes the step count to *steps_out_host
 * (one stream sync). */
int taco_forward(taco_handle* h, const int32_t* ids, const int32_t* lengths, const int32_t* spk,
                 const float* mel_targets, int N, int T_in, int T_tgt, int bn_mode,
                 int teacher_force, float* mel_out, float* linear_out, float* align_out,
                 int32_t* steps_out_host, void* stream);

/* Same call with HOST buffers (pinned or pageable): copies inputs to the
 * device, runs taco_forward, copies the outputs back and synchronises.  This
 * is the shape of session.run(feed_dict) at synthesizer.py:42-47. */
int taco_forward_host(taco_handle* h, const int32_t* ids_host, const int32_t* lengths_host,
                      const int32_t* spk_host, const float* mel_targets_host, int N, int T_in,
                      int T_tgt, int bn_mode, int teacher_force, float* mel_out_host,
                      float* linear_out_host, float* align_out_host, int32_t* steps_out_host,
                      void* stream);

/* The same call in two halves, for callers that keep several batches in flight
 * (one handle + stream each): _begin ENQUEUES the input copies, the forward and
 * the output copies and returns without waiting for anything -- a free-running
 * decode is enqueued for max_iters steps and its step count is read at the end
 * (only an early stop, i.e. an exactly-zero frame, makes _end redo the post-net
 * on the shorter length); _wait blocks until the decoder loop of that forward
 * (TACO_STAGE_DECODER) or its last kernel (TACO_STAGE_COMPUTE) has finished --
 * the post-net resp. the output copies may still be running -- so that a caller
 * can start another handle's forward while this one's post-net fills the SMs
 * the next decoder leaves free and its 144 MB of D2H traffic drains; _end
 * waits for the copies and returns the step count.  The
 * host buffers must stay valid until _end returns.  taco_forward_host == _begin
 * followed by _end.  _wait and _end spin by default; with TACO_BLOCKING_SYNC=1 in the
 * environment at taco_create they sleep on blocking-sync events (for hosts with more
 * waiting threads than cores). */
int taco_forward_host_begin(taco_handle* h, const int32_t* ids_host, const int32_t* lengths_host,
                            const int32_t* spk_host, const float* mel_targets_host, int N,
                            int T_in, int T_tgt, int bn_mode, int teacher_force,
                            float* mel_out_host, float* linear_out_host, float* align_out_host,
                            void* stream);
enum { TACO_STAGE_DECODER = 0, TACO_STAGE_COMPUTE = 1 };
int taco_forward_host_wait(taco_handle* h, int stage);
int taco_forward_host_end(taco_handle* h, int32_t* steps_out_host, void* stream);

/* ---- stage-level entry points (unit parity against the oracle) ----------- */
/* tacotron.py:46-55: out [N,T_in,E(+E_id)].  OOB ids write a zero row and make
 * the call return TACO_ERR_OOB_ID at the next taco_check_ids(). */
int taco_embed(taco_handle* h, const int32_t* ids, const int32_t* spk, int N, int T_in,
               float* out, void* stream);
int taco_check_ids(taco_handle* h, void* stream); /* syncs; 0 or TACO_ERR_OOB_ID */
/* tacotron.py:46-63: embeddings -> prenet -> encoder CBHG.  memory_out [N,T_in,256]. */
int taco_encoder(taco_handle* h, const int32_t* ids, const int32_t* lengths, const int32_t* spk,
                 int N, int T_in, int bn_mode, float* memory_out, void* stream);
/* tacotron.py:66-97,104: attention decoder loop.  dec_out [N,max_steps,num_mels*r]
 * (== mel_outputs [N,max_steps*r,num_mels]), align_out [N,T_in,max_steps] or NULL. */
int taco_decode(taco_handle* h, const float* memory, int N, int T_in, const float* mel_targets,
                int T_tgt, int teacher_force, float* dec_out, float* align_out,
                int32_t* steps_out_host, void* stream);
/* modules.py:35-74 on arbitrary input: which = TACO_CBHG_ENCODER ([N,T,128] in)
 * or TACO_CBHG_POST ([N,T,num_mels] in); lengths may be NULL; out [N,T,256].
 * x_batch_stride / out_batch_stride are in floats (0 = dense). */
int taco_cbhg(taco_handle* h, int which, const float* x, const int32_t* lengths, int N, int T,
              int bn_mode, int64_t x_batch_stride, float* out, void* stream);
/* tacotron.py:100-101: post CBHG + dense(num_freq).  mel [N,T,num_mels] with
 * batch stride mel_batch_stride floats, linear_out likewise. */
int taco_postnet(taco_handle* h, const float* mel, int N, int T, int bn_mode,
                 int64_t mel_batch_stride, float* linear_out, int64_t linear_batch_stride,
                 void* stream);
/* modules.py:68-74 alone: x [N,T,128] -> out [N,T,256]. */
int taco_bigru(taco_handle* h, int which, const float* x, const int32_t* lengths, int N, int T,
               float* out, void* stream);
/* tf.layers.conv1d(padding='same') + bias + activation on caller-supplied
 * device weights (modules.py:95-100; k=1 is tf.layers.dense).  x [N,T,Cin],
 * kernel [k,Cin,Cout], bias [Cout] or NULL, out [N,T,Cout]. */
int taco_conv1d(taco_handle* h, const float* x, int N, int T, int Cin, const float* kernel,
                const float* bias, int k, int Cout, int act, float* out, void* stream);

/* The two helper operators inside cbhg() that have no stage of their own, for unit parity:
 * tf.layers.max_pooling1d(pool_size=2, strides=1, padding='same') (modules.py:45-49) applied to scale[c] * x + shift[c]
 * (the folded batch-norm affine of the conv bank; NULL scale / shift = identity): x, out [N,T,C], C % 4 == 0; */
int taco_maxpool_affine(taco_handle* h, const float* x, int N, int T, int C, const float* scale,
                        const float* shift, float* out, void* stream);
/* and the training-mode statistics of tf.layers.batch_normalization (modules.py:101, is_training): biased moments of x [N,T,C]
 * over (N,T), folded with gamma / beta / epsilon 1e-3 into scale_out[c], shift_out[c] so that BN(x) = scale x + shift. C <= 2048. */
int taco_bn_batch_stats(taco_handle* h, const float* x, int N, int T, int C, const float* gamma,
                        const float* beta, float* scale_out, float* shift_out, void* stream);

/* ---- vocoder: the step right after the path ----------- */
/* Fields of reference hparams.py:13-18,35-36 that util/audio.py reads. */
typedef struct taco_audio_params {
  int32_t sample_rate;        /* hparams.py:13 (20000) */
  int32_t griffin_lim_iters;  /* hparams.py:35 (100)   */
  double frame_length_ms;     /* hparams.py:14 (50)    */
  double frame_shift_ms;      /* hparams.py:15 (12.5)  */
  double preemphasis;         /* hparams.py:16 (0.97); 0 skips the inverse pre-emphasis */
  double min_level_db;        /* hparams.py:17 (-100)  */
  double ref_level_db;        /* hparams.py:18 (20)    */
  double power;               /* hparams.py:36 (1.5)   */
} taco_audio_params;
/* Samples per utterance that taco_griffin_lim writes for T frames: (T-1)*hop + win with
 * hop = int(frame_shift_ms/1000*sample_rate), win = int(frame_length_ms/1000*sample_rate)
 * (util/audio.py:114-118; tf.contrib.signal.inverse_stft's overlap_and_add). < 0 on bad arguments. */
int64_t taco_wav_length(const taco_audio_params* ap, int T);
/* audio.inv_spectrogram_tensorflow(linear_outputs) followed by audio.inv_preemphasis
 * (synthesizer.py:27,50; util/audio.py:23-24,39-46,78-91,105-112) for a batch: linear [N,T,num_freq]
 * (batch stride linear_batch_stride floats, 0 = dense) -> wav_out [N, taco_wav_length(ap,T)] float32.
 * Needs num_freq == 1025 (n_fft 2048) and win <= 2048; TACO_ERR_UNSUPPORTED otherwise.  Does not need weights. */
int taco_griffin_lim(taco_handle* h, const taco_audio_params* ap, const float* linear, int N, int T,
                     int64_t linear_batch_stride, float* wav_out, void* stream);

/* ---- arithmetic mode of the dense layers ----------------------------------- */
/* The reference computes in fp32.  0 = fp32 FFMA kernels; 1 (default) = tcgen05 tensor
 * cores with every operand split into bf16 hi+lo and three products per k-step
 * (fp32-class accuracy, meets the 1e-3 parity bound); 2 = plain bf16 operands with fp32
 * accumulation in the dense layers AND in the decoder loop's mat-vecs (one product per
 * k-step instead of three; separately stated tolerance 5e-2, tests/test_gpu_parity.py).
 * The BiGRU recurrences always run in fp32. */
enum { TACO_GEMM_FFMA = 0, TACO_GEMM_BF16X3 = 1, TACO_GEMM_BF16 = 2 };
int taco_set_gemm_mode(taco_handle* h, int mode);

/* ---- introspection for bench / tests ------------------------------------- */
/* Kernel launches issued by this handle since creation. */
int64_t taco_launch_count(const taco_handle* h);
/* Decoder launch geometry chosen at finalize: CTAs per cluster, samples per
 * cluster used for batch N. */
int taco_decoder_geometry(const taco_handle* h, int N, int* cluster_size, int* samples_per_cluster,
                          int* num_clusters);
/* Decoder geometry: 0 (default) picks the number of 16-CTA clusters for the shortest decode (batch 32: 7 clusters of
 * 5/4 utterances, 112 SMs for 2.2 ms); n > 0 uses n clusters of up to 8 utterances each -- a throughput setting for
 * callers that keep several batches in flight (batch 32, n = 4: 64 SMs for 2.9 ms; +10 % mel frames/s with four batches
 * in flight, +16 % latency of a single one). */
int taco_set_decoder_clusters(taco_handle* h, int n);
/* Device time (ms, CUDA events on `stream`) of the stages of the last
 * taco_forward when profiling is on: [0]=encoder [1]=decoder stage (memory
 * layer + loop + step count) [2]=postnet [3]=the decoder loop kernel alone. */
int taco_set_profiling(taco_handle* h, int on);
/* The forward (taco_forward, taco_forward_host*) is ~49 dependent kernel launches; the reference runs its graph with one
 * session.run (synthesizer.py:47).  When the same call (same device pointers, shapes and modes) arrives twice in a row on a
 * non-default stream, the launches are captured into a CUDA graph and replayed from then on.  On by default
 * (TACO_GRAPHS=0 in the environment, or on = 0 here, keeps plain launches). */
int taco_set_cuda_graphs(taco_handle* h, int on);
int taco_last_stage_ms(const taco_handle* h, float* ms4_host);

#ifdef __cplusplus
}
#endif
#endif /* TACO_B200_H_ */
