/*
 * vapb200.h -- C ABI of the B200-native VAP streaming step (libvapb200.so).
 *
 * The reference (inokoj/VAP-Realtime) is pure Python/PyTorch and has no FFI
 * layer; its boundary for this path is the class surface of
 *     VAPRealTime.__init__ / VAPRealTime.process_vap
 *         (rvap/vap_main/vap_main.py:192-247, 249-335;
 *          twin rvap/vap_bc/vap_bc_main.py:241-300, vap_realtime/model.py:126-257)
 * and, in functional form, tools/vap_static.py:235-304.  Each entry point below
 * names the reference code it replaces.  All functions return 0 on success and
 * a negative VAPB_E* code on failure (never throw); vapb_last_error() gives
 * the text.
 *
 * Threading contract: a handle is NOT thread-safe; call it from one thread at
 * a time.  Different handles (same or different devices) may be driven from
 * different threads concurrently: the library keeps no process-global mutable
 * state (per-call launch flags are thread-local).  vapb_reset_streams,
 * vapb_export_state and vapb_import_state synchronise the whole device before
 * they return, so a following vapb_step may use any CUDA stream.
 *
 * Signatures carry plain pointers and sizes only (no torch / C++ types).
 */
#ifndef VAPB200_H_
#define VAPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAPB_OK            0
#define VAPB_EINVAL       -1   /* bad argument (id out of range, duplicate id in a batch, B > max) */
#define VAPB_ECUDA        -2   /* a CUDA runtime / driver call failed */
#define VAPB_EWEIGHTS     -3   /* weight blob malformed or tensor missing / wrong shape */
#define VAPB_ENOMEM       -4
#define VAPB_EUNSUPPORTED -5   /* not an sm_100 device, unsupported frame rate ... */

#define VAPB_HEAD_VAP 0        /* vap_head + va_classifier   (vap_main.py:131,142,290-317) */
#define VAPB_HEAD_BC  1        /* bc_head                    (vap_bc_main.py:137,272-277)  */

#define VAPB_OUT_STRIDE 6      /* floats per stream in the result rows, see vapb_step */

#if defined(__GNUC__)
#define VAPB_API __attribute__((visibility("default")))
#else
#define VAPB_API
#endif

typedef struct vapb_ctx* vapb_handle;

/*
 * Replaces VAPRealTime.__init__ (vap_main.py:192-247): builds the model from
 * a VAPW weight blob (vap_realtime_b200/weights.py; the merged VAP + CPC
 * state-dicts the reference loads at vap_main.py:199-212), allocates the
 * per-stream state for `max_streams` concurrent dialogues (LSTM h,c that the
 * reference keeps inside CPCAR, encoder_components.py:148-153, and the ring
 * of the last `ctx_frames` embeddings, vap_main.py:243-244,274-280) and all
 * workspaces for batches of up to `max_batch` streams per step.
 *   weights_blob : HOST pointer, nbytes long, copied during the call
 *   frame_hz     : 20, 10 or 5 (audio_frame_size = 16000/frame_hz + 320, vap_main.py:230)
 *   ctx_frames   : T = int(context_len_sec * frame_rate) (vap_main.py:221), 1..128
 *   device       : CUDA device ordinal (must be sm_100)
 */
VAPB_API int vapb_create(const void* weights_blob, size_t nbytes, int frame_hz, int ctx_frames,
                int max_streams, int max_batch, int head_kind, int device, vapb_handle* out);

/* Frees everything owned by the handle. */
VAPB_API int vapb_destroy(vapb_handle h);

/*
 * Zeroes ring length and LSTM (h,c) of the given streams (a fresh
 * VAPRealTime instance per dialogue).  stream_ids is a HOST array; n <= 0 or
 * stream_ids == NULL resets every stream.
 */
VAPB_API int vapb_reset_streams(vapb_handle h, const int* stream_ids, int n);

/*
 * One VAPRealTime.process_vap call (vap_main.py:249-335) for B independent
 * stereo streams.
 *   audio      : DEVICE pointer, [B, 2, chunk] fp32, chunk = vapb_chunk_samples();
 *                the first 320 samples are the previous chunk's tail
 *                (vap_main.py:224, 408-409)
 *   stream_ids : HOST array [B] of distinct state slots in [0, max_streams)
 *   out        : DEVICE pointer, [B, 6] fp32:
 *                  VAPB_HEAD_VAP: p_now[0], p_now[1], p_future[0], p_future[1], vad[0], vad[1]
 *                                 (result_p_now / result_p_future / result_vad, vap_main.py:313-320)
 *                  VAPB_HEAD_BC : p_bc_react, p_bc_emo, 0, 0, 0, 0 (vap_bc_main.py:276-284)
 *   cuda_stream: cudaStream_t the work is enqueued on (asynchronous; the
 *                caller synchronises).  May be NULL (legacy default stream).
 * The step replays one CUDA graph per batch size B (captured on first use, LRU over 96 sizes); `audio` and
 * `out` are read through a small device record refreshed by the same host->device copy that carries the
 * stream ids, so callers may pass different buffers on every call at no cost.
 */
VAPB_API int vapb_step(vapb_handle h, const float* audio, const int* stream_ids, int B, float* out,
              void* cuda_stream);

/*
 * Same step with HOST buffers, the shape of the call the reference's server
 * makes (numpy in, python floats out: vap_main.py:262-270, 310-317): copies
 * audio host->device, runs the step, copies the [B,6] result back and waits
 * for it.  `audio`/`out` should be page-locked for full copy speed.
 */
VAPB_API int vapb_step_host(vapb_handle h, const float* audio, const int* stream_ids, int B, float* out,
                   void* cuda_stream);

/*
 * Bulk offline scoring of one two-channel recording: the result of replaying it frame by frame through
 * process_vap the way rvap/vap_main/vap_offline.py:51-73 does (chunk n = samples [shift*n, shift*n + chunk),
 * shift = 16000/frame_hz, from a fresh model state), computed in two batched passes instead of N batch-1 steps:
 * the conv stack over many chunks at once (chunks are convolved in isolation), the LSTM sequentially in time, then
 * one transformer window per frame, max_batch windows per launch.  Does not touch any stream's state.
 *   audio     : DEVICE pointer, planar [2][n_samples] fp32
 *   out       : DEVICE pointer, [max_frames][6] fp32, same columns as vapb_step
 *   n_frames  : HOST, receives the number of frames ((n_samples - chunk) / shift + 1)
 * Synchronous (returns when `out` is complete).
 */
VAPB_API int vapb_score_offline(vapb_handle h, const float* audio, long long n_samples, float* out, long long max_frames,
                       long long* n_frames, void* cuda_stream);

/* Samples per channel per step: 16000/frame_hz + 320 (vap_main.py:230). */
VAPB_API int vapb_chunk_samples(vapb_handle h);

/*
 * Per-stream state for migration / parity checks (the reference never
 * serialises it).  Layout of one record, `vapb_state_floats()`
 * floats long:
 *   [0] frames seen so far (as float), [1] ring length t,
 *   h[2][256], c[2][256], ring[2][T][256] in logical order (oldest first,
 *   rows >= t zero).
 * `state` is a HOST pointer.  Both calls synchronise the device.
 */
VAPB_API size_t vapb_state_floats(vapb_handle h);
VAPB_API int vapb_export_state(vapb_handle h, int stream_id, float* state);
VAPB_API int vapb_import_state(vapb_handle h, int stream_id, const float* state);

/*
 * Options (debug / measurement):
 *   "graph"     0/1   replay the step as a CUDA graph (default 1)
 *   "gemm"      0/1   0 = fp32 CUDA-core GEMMs everywhere, 1 = tcgen05 bf16 hi/lo x3
 *                     tensor-core GEMMs for conv1-4, downsample and the transformer
 *   "keep_taps" 0/1   keep intermediates readable through vapb_debug_tensor
 *   "fused"     0/1/2 transformer stack as ONE per-stream persistent kernel (a cluster of two CTAs per stream):
 *                     0 = never (batched per-op kernels), 1 = while all clusters are co-resident, i.e. 2B <= 148
 *                     (default), 2 = always
 *   "fused_dbg" n     n > 0: clock64 stamps per op of that kernel (fine stamps for op n - 1), read back through
 *                     vapb_debug_tensor("fused_clocks")
 *   tuning / ablation switches, each measured in profiles/: "prune", "splitk", "conv4p", "lstm_fused", "fuse_ln",
 *   "k256", "tile_n", "cluster2", "attn_rk", "fork", "pdl", "timing"
 */
VAPB_API int vapb_set_option(vapb_handle h, const char* key, int value);
VAPB_API int vapb_get_option(vapb_handle h, const char* key, int* value);

/*
 * Copies a named intermediate of the LAST step to `host_out` (at most `cap`
 * floats); *n receives the element count.  Names mirror the oracle taps:
 * conv0..conv4, lstm_out, e, x_in, chan_out, cross0_out..cross2_out, comb, logits.
 */
VAPB_API int vapb_debug_tensor(vapb_handle h, const char* name, float* host_out, size_t cap, size_t* n);

/* Kernels launched by the most recent vapb_step (graph nodes count as launches). */
VAPB_API int vapb_last_launch_count(vapb_handle h);

/* Device time of the most recent vapb_step in ms (events on the caller's
 * stream; blocks until that step has finished).  Requires option "timing"=1. */
VAPB_API int vapb_last_step_ms(vapb_handle h, float* ms);

/*
 * Measurement aid: runs ONE step eagerly (no graph) with a CUDA event behind every kernel and
 * writes a CSV ("tag,launches,ms" per kernel class, then "total") into `report`.  The step
 * counts like any other step (state advances).  Synchronises the stream.
 */
VAPB_API int vapb_profile_step(vapb_handle h, const float* audio, const int* stream_ids, int B, float* out,
                      void* cuda_stream, char* report, size_t cap);

/* Text of the last error on this handle (or of the last failed vapb_create
 * when h is NULL).  Valid until the next call on the handle. */
VAPB_API const char* vapb_last_error(vapb_handle h);

/* Library version / build string, e.g. "vapb200 0.1 sm_100a". */
VAPB_API const char* vapb_version(void);

/* Stand-alone self test of the tcgen05 GEMM building block against an fp32
 * CUDA-core product (used by the GPU tests). Returns 0 and writes the max
 * relative error to *max_rel_err. `variant` selects the operand shape. */
VAPB_API int vapb_selftest_gemm(int device, int variant, double* max_rel_err);

#ifdef __cplusplus
}
#endif
#endif /* VAPB200_H_ */
