/*
 * b200mel.h — C ABI of libb200mel.so: the B200 (sm_100a) fused
 * STFT -> magnitude -> mel filterbank -> log feature extractor that sits
 * behind pytorch_sound's spectral operators.
 *
 * The reference (AppleHolic/pytorch_sound) has no FFI layer: its operator
 * surface is a set of Python nn.Modules.  Every entry point below therefore
 * cites the reference Python symbol it replaces (paths relative to
 * /root/reference/pytorch_sound/); INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA runtime types in signatures
 *     (a CUDA stream is passed as void*, 0 == legacy default stream);
 *   - every function returns 0 (B200MEL_OK) or a negative B200MEL_E* code and
 *     never throws / aborts; the message is in b200mel_last_error()
 *     (thread-local);
 *   - the caller owns every data buffer; a plan owns only its device-side
 *     tables (window, twiddles, banded filterbank) and is immutable after
 *     creation, so b200mel_forward is re-entrant;
 *   - no host synchronisation and no allocation inside b200mel_forward;
 *   - there is NO CPU fallback: without a CUDA device plan_create fails.
 */
#ifndef B200MEL_H
#define B200MEL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MEL_VERSION 210 /* 0.2.1: b200mel_io gained dct_mat / out_mfcc / n_mfcc (struct_size 112), b200mel_gather_tma */

/* error codes */
#define B200MEL_OK 0
#define B200MEL_EINVAL (-1)  /* bad config / shape / null pointer            */
#define B200MEL_ECUDA (-2)   /* CUDA runtime error (message has the string)  */
#define B200MEL_ENODEV (-3)  /* no CUDA device / not an sm_100 device         */
#define B200MEL_ENOMEM (-4)  /* host or device allocation failed              */
#define B200MEL_EUNSUP (-5)  /* valid in the reference, not built here (yet)  */

/* framing: where frame 0 starts relative to sample 0 */
#define B200MEL_PAD_CENTER 0 /* reflect n_fft/2 each side  (STFT.transform models/transforms.py:55-60;
                                torch.stft(center=True) models/transforms.py:298-301)            */
#define B200MEL_PAD_HIFI 1   /* reflect (n_fft-hop)/2 each side, then no centring
                                (Audio2Mel models/transforms.py:352-360; interface/hifi_gan.py:48-54) */

/* mel scale / normalisation of the filterbank */
#define B200MEL_MEL_SLANEY 0 /* librosa.filters.mel default (htk=False)                     */
#define B200MEL_MEL_HTK 1    /* torchaudio 0.7 MelScale (LogMelSpectrogramTorchAudio :383-385) */
#define B200MEL_NORM_NONE 0
#define B200MEL_NORM_SLANEY 1 /* librosa default: rows scaled by 2/(f[i+2]-f[i]) */

/* log epilogue */
#define B200MEL_LOG_NONE 0        /* linear mel                                               */
#define B200MEL_LOG_LN_OFFSET 1   /* ln(mel + arg)        LogMelSpectrogram.forward :238      */
#define B200MEL_LOG_LN_FLOOR 2    /* ln(max(mel, arg))    interface/hifi_gan.py:61            */
#define B200MEL_LOG_LOG10_FLOOR 3 /* log10(max(mel, arg)) Audio2Mel.forward :365              */

/* optional spectrum outputs (B, n_fft/2+1, T) */
#define B200MEL_SPEC_NONE 0
#define B200MEL_SPEC_MAG_PHASE 1 /* out_a = |X|, out_b = atan2(im, re)   STFT.transform :69, STFTTorchAudio.transform :311 */
#define B200MEL_SPEC_RE_IM 2     /* out_a = re,  out_b = im              STFTTorchAudio.forward :297-303                  */
#define B200MEL_SPEC_MAG 3       /* out_a = |X| only                                                                     */

/* Geometry of one extractor.  Mirrors the constructor arguments of
 * LogMelSpectrogram (models/transforms.py:211-213), Audio2Mel (:327-336),
 * interface.hifi_gan.MelSpectrogram (interface/hifi_gan.py:34-35) and the
 * constants of settings.py:9-22. */
typedef struct b200mel_config {
    int32_t struct_size; /* = sizeof(b200mel_config), for forward compatibility */
    int32_t sample_rate;
    int32_t n_fft;      /* a power of two in [32, 2048]: sizes <= 1024 run the 1024-point pair kernel on the zero-extended
                           frame (bin k of an N-point DFT is bin k * 1024 / N of its 1024-point DFT), 2048 the split kernel */
    int32_t win_length; /* <= n_fft; periodic Hann, centre-padded to n_fft (models/transforms.py:30-31) */
    int32_t hop_length;
    int32_t n_mels; /* 0 = spectrum-only plan (STFT / STFTTorchAudio) */
    float fmin;
    float fmax; /* <= 0 means sample_rate / 2 (librosa default) */
    int32_t pad_mode;  /* B200MEL_PAD_*  */
    int32_t mel_scale; /* B200MEL_MEL_*  */
    int32_t mel_norm;  /* B200MEL_NORM_* */
    int32_t power;     /* 1 = magnitude (all pytorch_sound modules), 2 = power (torchaudio variant) */
    float mag_eps;     /* added under the sqrt: 0 or 1e-9 (interface/hifi_gan.py:55) */
} b200mel_config;

/* Per-call epilogue.  log_arg is a forward() argument in the reference
 * (LogMelSpectrogram.forward(wav, log_offset=1e-6)), so it is per call here. */
typedef struct b200mel_epilogue {
    int32_t struct_size;
    int32_t log_kind; /* B200MEL_LOG_* */
    float log_arg;    /* offset or floor */
    int32_t has_clamp_lo;
    float clamp_lo; /* natural-log units, db2log(min_db) utils/calculate.py:10-19 */
    int32_t has_clamp_hi;
    float clamp_hi;
    int32_t norm_mel; /* 1: (clamp(y, lo, hi) - lo)/(hi - lo)*2 - 1   utils/calculate.py:32-43 (needs both clamps) */
} b200mel_epilogue;

typedef struct b200mel_plan b200mel_plan; /* opaque */

int b200mel_version(void);
const char *b200mel_last_error(void);

/* Host helpers (no GPU needed).  They restate the third-party arithmetic the
 * reference calls so that the Python shim can keep `mel_filter` / `window`
 * buffers with the reference's names and values.
 *   b200mel_mel_filterbank  <- librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) (librosa 0.8.0;
 *                              call sites models/transforms.py:220,339-341, interface/hifi_gan.py:42)
 *   b200mel_hann_window     <- scipy.signal.get_window('hann', win, fftbins=True) + librosa.util.pad_center
 *                              (models/transforms.py:30-31) == torch.hann_window(win) (:287) */
int b200mel_mel_filterbank(int32_t sample_rate, int32_t n_fft, int32_t n_mels, double fmin, double fmax,
                           int32_t mel_scale, int32_t mel_norm, float *out /* [n_mels][n_fft/2+1] */);
int b200mel_hann_window(int32_t win_length, int32_t n_fft, float *out /* [n_fft] */);

/* Number of frames for clips of L samples:
 * centre: 1 + L/hop; hifi: (L + 2*((n_fft-hop)/2) - n_fft)/hop + 1.  */
int b200mel_out_frames(const b200mel_plan *plan, int64_t L, int64_t *T);

/* Create a plan on the CURRENT CUDA device (cudaGetDevice).  Builds window,
 * twiddle tables and the banded (CSR-like) filterbank and uploads them. */
int b200mel_plan_create(const b200mel_config *cfg, b200mel_plan **out);
/* Replace the plan's filterbank by caller-provided weights (HOST pointer,
 * row-major [n_mels][n_fft/2+1]) — used when a state_dict supplied a
 * different `mel_filter`.  Not thread-safe against concurrent forward calls. */
int b200mel_plan_set_filterbank(b200mel_plan *plan, const float *weights, int32_t n_mels, int32_t n_freq);
int b200mel_plan_destroy(b200mel_plan *plan);

/* One launch of the fused kernel over a clip batch.
 *   wav        device pointer, float32, B rows of L valid samples, row r at wav + r*row_stride
 *   lengths    nullable device pointer int32[B]: valid samples per clip (<= L).  Reflection happens
 *              at the clip's own end and frames t >= frames(lengths[b]) are written as 0
 *              (SpeechDataLoader.pad_collate_fn semantics, data/dataset.py:196-250).  A clip with
 *              lengths[b] <= pad cannot be reflect-padded (the reference raises): all its frames are 0.
 *   out_mel    nullable device pointer float32 (B, n_mels, T) contiguous, T from b200mel_out_frames(L)
 *   out_a/out_b nullable, (B, n_fft/2+1, T), meaning given by spec_kind
 *   stream     CUDA stream handle (cudaStream_t) or NULL
 * Replaces: LogMelSpectrogram.forward (models/transforms.py:231-244), STFT.transform (:53-69),
 * STFTTorchAudio.forward/transform (:297-311), Audio2Mel.forward (:351-366),
 * interface.hifi_gan.MelSpectrogram.forward (interface/hifi_gan.py:46-63). */
int b200mel_forward(const b200mel_plan *plan, const float *wav, int64_t B, int64_t L, int64_t row_stride,
                    const int32_t *lengths, const b200mel_epilogue *epi, float *out_mel, int32_t spec_kind,
                    float *out_a, float *out_b, void *stream);

/* Extensible form of b200mel_forward: every pointer of one call in a struct (struct_size guards the layout), plus
 * the outputs added after version 100:
 *   out_frame_mask  nullable device pointer float32 (B, T): frame-level validity mask with the semantics of
 *                   SpectrogramMasker.forward (models/transforms.py:397-416) applied to the wave-level mask of ones
 *                   that SpeechDataset appends (data/dataset.py:73-74,92-93): frame t is 1 iff its window
 *                   [t*hop - win/2, t*hop + win/2) contains a valid sample or left padding, i.e.
 *                   t*hop - win_length/2 < lengths[b] (all ones without `lengths`).  Written by the same launch
 *                   as the mel frames (centre framing only).
 *   out_mfcc        nullable device pointer float32 (B, n_mfcc, T): MelToMFCC.forward / MFCC.forward
 *                   (models/transforms.py:419-455) fused as an epilogue of the mel launch — `dct_mat` (n_mfcc, n_mels)
 *                   row-major, the module's registered buffer, is applied to the log-mel column of every frame while
 *                   it is still on chip.  out_mel may then be NULL (the mel frames are never written).  Served by the
 *                   compile-time specialised kernel only (n_fft = win_length = 1024, hop 256, no `lengths`, no frame
 *                   mask / pre-emphasis / spectrum outputs, a log epilogue, n_mfcc <= 64): any other call returns
 *                   B200MEL_EUNSUP and the caller runs b200mel_mel_to_mfcc on the mel output instead.
 * b200mel_forward(plan, wav, B, L, row_stride, lengths, epi, out_mel, spec_kind, out_a, out_b, stream) is exactly
 * b200mel_forward_io with out_frame_mask = NULL, reserve_sms = 0, preemphasis = 0, out_mfcc = NULL. */
typedef struct b200mel_io {
    int32_t struct_size; /* = sizeof(b200mel_io) */
    int32_t spec_kind;   /* B200MEL_SPEC_* */
    const float *wav;
    int64_t B, L, row_stride;
    const int32_t *lengths;
    float *out_mel, *out_a, *out_b;
    float *out_frame_mask;
    int32_t reserve_sms; /* the persistent mel kernel launches on (SM count - reserve_sms) SMs (0 = all).  It owns every
                            register of an SM it runs on, so a concurrent kernel on another stream — e.g. the one-CTA
                            barrier of b200mel_gather_copy, which gates the copy-engine transfers — only gets an SM
                            when one is left free. */
    float preemphasis;   /* != 0: y[n] = x[n] - preemphasis * x[n-1] with the reference's 1-sample reflect pad (y[0] = x[0] -
                            preemphasis * x[1]) applied to the samples as they are staged, before framing — PreEmphasis.forward
                            (models/sound.py:66-81) fused as a prologue of the mel launch */
    const float *dct_mat; /* (n_mfcc, n_mels) row-major, device; read when out_mfcc != NULL */
    float *out_mfcc;
    int32_t n_mfcc;
    int32_t reserved0;    /* must be 0 */
} b200mel_io;
int b200mel_forward_io(const b200mel_plan *plan, const b200mel_io *io, const b200mel_epilogue *epi, void *stream);

/* The mel filterbank as a tensor-core GEMM on magnitudes that already sit in device memory:
 *   out_mel (B, n_mels, T) = epilogue( W (n_mels, n_fft/2+1) @ mag (B, n_fft/2+1, T) )
 * Replaces `torch.matmul(self.mel_filter, magnitude)` + log + clamp of LogMelScale.forward
 * (models/transforms.py:261-268; the same contraction as LogMelSpectrogram.forward :235-243 when the caller keeps the
 * magnitudes, e.g. next to multi_stft_loss).  One launch of mel_tc_kernel: tcgen05.mma (UMMA 128 x n_mels x 16, bf16
 * hi/lo split of both operands, three MMAs per K step) with the fp32 accumulator in TMEM, log epilogue on tcgen05.ld.
 * `plan` supplies the filterbank (its bf16 limbs are built on first use, hence non-const); geometry limits: n_mels <=
 * 128 and a filterbank that fits in shared memory (every BASELINE config with n_fft 1024), otherwise B200MEL_EUNSUP. */
int b200mel_logmel_from_magnitude(b200mel_plan *plan, const float *mag, int64_t B, int64_t T,
                                  const b200mel_epilogue *epi, float *out_mel, void *stream);

/* Same, but wav_host / out_mel_host are HOST pointers (pinned memory gives
 * asynchronous copies): H2D copy -> kernel -> D2H copy on `stream` using
 * plan-owned device staging that grows on demand.  The staging is kept PER
 * STREAM (up to 8 streams per plan), so calls on different streams overlap —
 * the copy of one batch under the kernel and the read-back of another — and
 * calls on one stream are ordered by the stream.  It returns after enqueuing;
 * the caller synchronises the stream(s).  This is the "host buffers in, host
 * buffers out" call the reference's CPU modules correspond to (wav on CPU in,
 * mel on CPU out; DataLoader pin_memory data/dataset.py:180 + to_device
 * utils/tensor.py:15 on the way in). */
int b200mel_forward_host(b200mel_plan *plan, const float *wav_host, int64_t B, int64_t L, int64_t row_stride,
                         const b200mel_epilogue *epi, float *out_mel_host, void *stream);

/* ---- the small operators either side of the spectral path (device pointers, one launch each unless noted) ----
 *
 * b200mel_preemphasis  <- models/sound.py:66-81 PreEmphasis.forward: y[n] = x[n] - coef * x[n-1] with the
 *                         reference's 1-sample reflect pad on the left (y[0] = x[0] - coef * x[1]).  Rows of L
 *                         samples at x + r * x_row_stride -> y + r * y_row_stride; y must not alias x.
 * b200mel_volume_norm  <- utils/calculate.py:56-63 volume_norm_log_torch: y = x / (std(x) / 10^(target_db/10)),
 *                         std = unbiased standard deviation over all n elements (torch.std).  `scratch` is a
 *                         device buffer of 2 doubles owned by the caller; two launches (moments, scale).
 * b200mel_mel_to_mfcc  <- models/transforms.py:419-430 MelToMFCC.forward: out (B, n_mfcc, T) =
 *                         dct (n_mfcc, n_mels) @ mel (B, n_mels, T); dct is a device pointer (the module's
 *                         `dct_mat` buffer, torchaudio.functional.create_dct transposed, :427).
 * b200mel_stft_loss_terms <- models/sound.py:139-141, the two reductions of one resolution of multi_stft_loss over
 *                         magnitude tensors (B, F, T) (n_per_clip = F * T, contiguous):
 *                           out2[0] += mean_b ||t_b - p_b||_F / ||t_b||_F          (spectral convergence)
 *                           out2[1] += mean_b sum |ln(t_b + eps) - ln(p_b + eps)| / n_per_clip
 *                         out2 (device, 2 floats) ACCUMULATES so the resolutions of one loss add up in place (zero it
 *                         first); scratch is a caller-owned device buffer of 3 * B doubles.  Memset + two launches. */
int b200mel_stft_loss_terms(const float *pred_mag, const float *target_mag, int64_t B, int64_t n_per_clip, float eps,
                            double *scratch, float *out2, void *stream);
int b200mel_preemphasis(const float *x, int64_t B, int64_t L, int64_t x_row_stride, float coef, float *y,
                        int64_t y_row_stride, void *stream);
int b200mel_volume_norm(const float *x, int64_t n, float target_db, float *y, double *scratch, void *stream);
int b200mel_mel_to_mfcc(const float *mel, const float *dct, int64_t B, int32_t n_mels, int32_t n_mfcc, int64_t T,
                        float *out, void *stream);

/* Multi-GPU: all-gather of the ranks' mel blocks by pulling from peer memory over NVLink (the reference
 * has no distributed code, trainer.py:269-272 only strips nn.DataParallel prefixes).  Every rank holds a buffer with
 * the same layout — `world` blocks back to back, block r = elements [block_offsets[r], block_offsets[r+1]) — and has
 * just written ITS block into ITS buffer (b200mel_forward with out_mel pointing at the block).  peer_bufs is a HOST
 * array of `world` device pointers: rank r's buffer as mapped into this process (e.g. the buffer_ptrs of
 * torch.distributed._symmetric_memory; entry `rank` is ignored).  One launch copies every peer's own block into
 * local_buf with 16-byte peer loads.
 * Ordering across ranks: peer_sync (nullable) is a HOST array of `world` device pointers to each rank's sync words
 * (int32[world + 2], zero-initialised once, symmetric like the buffers): with it the kernel itself runs the barrier —
 * it publishes "my block of this step is written" to every peer (system-scope release stores) and waits for all
 * peers' flags of the same step (bounded spin, then a trap) before pulling; the step number lives in device memory,
 * so the launch can be captured in a CUDA graph and replayed.  Every rank must launch once per step.  Without
 * peer_sync the caller orders the launch after its own cross-rank barrier on `stream`.  Either way a rank must not
 * overwrite its block while peers may still be pulling it: rotate three buffers and order the extraction of step i
 * after this rank's gather of step i-2 (pytorch_sound_b200/distributed.py). */
int b200mel_gather_pull(float *local_buf, const float *const *peer_bufs, int32_t *const *peer_sync, int32_t world,
                        int32_t rank, const int64_t *block_offsets /* [world + 1] */,
                        int32_t n_ctas /* 512-thread CTAs of the pull kernel; 0 = two per SM */, void *stream);
/* b200mel_gather_pull with the TMA engine: one thread per CTA streams the peers' blocks through a ring of
 * shared-memory stages with bulk async copies (peer -> shared over NVLink, shared -> local), 128 KB in flight per CTA and
 * no load / store instructions, so n_ctas = 16 (0 = 32) CTAs cover the link and the concurrent extraction kernel keeps
 * the other SMs (io.reserve_sms = n_ctas).  Same arguments, barrier and result as b200mel_gather_pull. */
int b200mel_gather_tma(float *local_buf, const float *const *peer_bufs, int32_t *const *peer_sync, int32_t world,
                       int32_t rank, const int64_t *block_offsets, int32_t n_ctas, void *stream);

/* SM partitioning: the extraction kernel is persistent and owns every register of the SMs it runs on, so a pull
 * kernel launched on another stream only runs beside it on SMs the extraction left free — launch the extraction
 * with b200mel_io.reserve_sms = R and the pull with n_ctas = 2 R and the two overlap (R ~ 32 of 148 SMs keeps
 * NVLink busy while the extraction runs on the rest).
 *
 * Same gather with the COPY ENGINES doing the transfers: a one-CTA barrier kernel (peer_sync is required), then one
 * device-to-device copy per peer on internal streams forked from and joined back into `stream`.  The extraction
 * kernel is persistent and fills every SM, so an SM-resident pull cannot run beside it; the copy engines can — use
 * this variant when the gather of step i is overlapped with the extraction of step i+1 on another stream, and
 * b200mel_gather_pull (higher peak bandwidth, 16-byte peer loads) when the gather runs alone. */
int b200mel_gather_copy(float *local_buf, const float *const *peer_bufs, int32_t *const *peer_sync, int32_t world,
                        int32_t rank, const int64_t *block_offsets /* [world + 1] */, void *stream);

/* Number of kernel launches issued through this library since load (all plans;
 * used by bench.py's gpu_launches claim). */
int64_t b200mel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200MEL_H */
