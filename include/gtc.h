/*
 * gtc.h -- C ABI of libgtc.so, the B200 (sm_100a) hot path of the guitar-tablature feature front-end.
 *
 * The reference (AshishBhardwaj01/Guitar-Tablature-Classification) is pure Python and has no FFI; its
 * "operator interface" for this path is the Python module API of cqt.py / new_cqt.py / jam_to_tablature.py /
 * my_dataloader.py / ViT_dataloader.py.  Each entry point below replaces the *library arithmetic* behind one of
 * those call sites; the Python drop-in modules in guitar-tablature-classification_b200/ bind them with ctypes
 * (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all `d_*` pointers are DEVICE pointers owned by the caller,
 *     all `h_*` pointers are HOST pointers.  The library never allocates or frees caller-visible memory;
 *     scratch memory is passed in as a caller-owned workspace.
 *   - every function returns 0 on success, <0 on error (GTC_E_*); gtc_last_error() gives a thread-local message.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it.
 *   - a plan is immutable after creation (thread-safe to share), one plan per device.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns GTC_E_CUDA.
 */
#ifndef GTC_H_
#define GTC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTC_VERSION 102

#define GTC_OK          0
#define GTC_E_ARG      -1   /* invalid argument                      */
#define GTC_E_CUDA     -2   /* CUDA runtime / driver error           */
#define GTC_E_NOMEM    -3   /* workspace too small / allocation      */
#define GTC_E_UNSUP    -4   /* unsupported configuration             */

/* GEMM engines for the segment operator contraction (the Python host's default is GTC_GEMM_TCGEN05_FP16X2) */
#define GTC_GEMM_TCGEN05_FP16X2  2   /* DEFAULT.  TMA + tcgen05.mma kind::f16 on power-of-two scaled fp16 hi/lo pairs (22 mantissa bits):
                                        half the tensor time and operand bytes of 3xTF32; needs |audio| < 256                        */
#define GTC_GEMM_TCGEN05_3XTF32  0   /* TMA + tcgen05.mma kind::tf32, tf32 hi + fp32 lo split, fp32 TMEM accumulators                */
#define GTC_GEMM_SIMT_FP32       1   /* CUDA-core fp32 FMA tiles (validation engine for the tensor-core paths)                       */

/* patch modes */
#define GTC_PATCH_VIT  0   /* ViT_dataloader.py:31-51  : (x+120)/120, clip, bicubic, 3 identical channels            */
#define GTC_PATCH_CNN  1   /* my_dataloader.py:17-21   : grey picture (top row = highest bin), bilinear, ImageNet normalise */
#define GTC_PATCH_VIT_PRENORM 2 /* "tablature-generator (1).py":349-368 prepare_for_vit: input already normalised to [0,1]; bicubic, 3 channels */

typedef struct gtc_plan gtc_plan;
typedef void* gtc_stream_t;

int         gtc_version(void);
const char* gtc_last_error(void);
/* sm count, compute capability and memory of `device`; any out pointer may be NULL. */
int         gtc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ------------------------------------------------------------------------------------------------------------
 * CQT of fixed-length segments  -- replaces librosa.cqt + np.abs()**4 + librosa.amplitude_to_db(ref=np.amax)
 * + cqt_lim at /root/reference/cqt.py:55-58 (and new_cqt.py:25-30, tablature-generator (1).py:326-331).
 *
 * For a fixed segment length the whole librosa.cqt call is a linear map; the host designs it once
 * (gtc_b200/cqt_design.py) and hands it over as `h_operator`:
 *      h_operator[(t*n_bins + bin)*2 + c][j],  c = 0 real / 1 imag, j = 0..seg_len-1   (row-major, fp32)
 * ------------------------------------------------------------------------------------------------------------ */
int gtc_cqt_plan_create(gtc_plan** out, int device, int seg_len, int seg_hop, int n_bins, int n_frames,
                        const float* h_operator, int gemm_engine);
int gtc_cqt_plan_destroy(gtc_plan* plan);
/* tuning knobs; call before the plan is shared between threads.  Returns GTC_E_ARG for unknown options. */
#define GTC_OPT_TC_KSPLIT      1   /* k-blocks (32 fp32 each) accumulated inside the tensor core before an fp32 add; default 8 */
#define GTC_OPT_GEMM_MAX_CTAS  2   /* limit of the persistent GEMM grid (0 = one CTA per SM); lets other kernels share the GPU */
#define GTC_OPT_FUSE_FINISH    3   /* 1: |C|^power -> dB -> cut done inside the tcgen05 GEMM epilogue (bit-identical); 0 (default): separate 19.5 us finish pass, measured faster */
int gtc_cqt_plan_configure(gtc_plan* plan, int option, int value);
/* number of operator rows sharing one audio row (P = seg_len/seg_hop when it divides, else 1) */
int gtc_cqt_plan_parts(const gtc_plan* plan);
/* bytes of scratch gtc_cqt_segments_db needs for `n_seg` segments spread over `n_clips` clips */
int gtc_cqt_workspace_bytes(const gtc_plan* plan, int64_t n_seg, int64_t n_clips, size_t* bytes);

/*
 * All clips of a shard concatenated in d_audio (mono fp32).  d_clip_off[n_clips+1]: sample offset of each clip;
 * d_seg_off[n_clips+1]: index of each clip's first segment in the output; segment s of clip c covers samples
 * [d_clip_off[c] + s*seg_hop, +seg_len)  (window arithmetic of cqt.py:26-45; the caller computes the counts).
 * d_out_db: [n_seg, n_bins, n_frames] fp32, C-order  == np.stack of the arrays cqt.py:58 produces.
 * power: exponent applied to |C| (4 at cqt.py:56; 1 at tablature_generator.py:620); amin/top_db as librosa;
 * values < cut_db are replaced by floor_db (cqt_lim: -60 -> -120); pass cut_db = -INFINITY to disable.
 */
int gtc_cqt_segments_db(const gtc_plan* plan, const float* d_audio, const int64_t* d_clip_off,
                        const int64_t* d_seg_off, int64_t n_clips, int64_t n_seg, float* d_out_db,
                        void* d_workspace, size_t workspace_bytes,
                        float power, float amin, float top_db, float cut_db, float floor_db,
                        gtc_stream_t stream);
/* The two stages of gtc_cqt_segments_db as separate calls (same workspace, same stream order), so that a caller can
 * run them on different streams: the pipeline frames chunk k+1 beside the tensor-core contraction of chunk k and keeps
 * GEMM -> finish -> patches of one chunk back to back on the compute stream (gtc_b200/pipeline.py). */
int gtc_cqt_frame(const gtc_plan* plan, const float* d_audio, const int64_t* d_clip_off, const int64_t* d_seg_off,
                  int64_t n_clips, int64_t n_seg, void* d_workspace, size_t workspace_bytes, gtc_stream_t stream);
/* gtc_cqt_frame for the WAV file's own 16-bit PCM samples (mono, already channel-averaged if needed): the kernel converts
 * x / 32768 in fp32 exactly as librosa.load(sr=None) does for PCM_16 files (/root/reference/cqt.py:23), so the result is
 * identical to uploading the fp32 array at half the host->device bytes. */
int gtc_cqt_frame_pcm16(const gtc_plan* plan, const int16_t* d_pcm, const int64_t* d_clip_off, const int64_t* d_seg_off,
                        int64_t n_clips, int64_t n_seg, void* d_workspace, size_t workspace_bytes, gtc_stream_t stream);
int gtc_cqt_contract_db(const gtc_plan* plan, const int64_t* d_clip_off, const int64_t* d_seg_off, int64_t n_clips,
                        int64_t n_seg, float* d_out_db, void* d_workspace, size_t workspace_bytes,
                        float power, float amin, float top_db, float cut_db, float floor_db, gtc_stream_t stream);
/* process-wide tunables */
#define GTC_OPT_PATCH_MAX_CTAS 16  /* grid limit of the patch kernel (0 = SMs x resident CTAs) */
#define GTC_OPT_PATCH_CTAS_PER_SM 17 /* resident patch CTAs per SM (default 2: measured fastest on B200; 0 restores the default).  The launch
                                       requests 228 KB / k of shared memory so that exactly k CTAs fit and the persistent grid is placed evenly */
int gtc_set_option(int option, int value);

/* Same contraction, complex output before |.|: d_out_c [n_seg, n_bins, n_frames, 2] fp32 (== librosa.cqt). */
int gtc_cqt_segments_complex(const gtc_plan* plan, const float* d_audio, const int64_t* d_clip_off,
                             const int64_t* d_seg_off, int64_t n_clips, int64_t n_seg, float* d_out_c,
                             void* d_workspace, size_t workspace_bytes, gtc_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Structured (multirate) CQT of variable-length segments -- replaces librosa.cqt + np.abs + amplitude_to_db where the
 * collapsed segment operator does not apply: the inference front-end of /root/reference/tablature_generator.py:616-620
 * (sr 22 050, hop 512, fmin C2, 84 bins, ref=np.max, 3 s segments cut by segment_audio :637-666), whole-clip CQTs and
 * any sample rate.  Evaluated the way librosa does: a chain of soxr-HQ 2:1 decimations (shared-memory FIR) and, per
 * octave, the wavelet filters applied to centred zero-padded frames.
 *
 * h_filters [n_octaves][2*filters_per_octave][n_fft] fp32: time-domain filters of octave i (0 = top octave),
 *           row = filter*2 + {re, im}, all scalings folded in (gtc_b200.cqt_design.structured_filters).
 * h_taps    [n_taps] fp32: the 2:1 decimator with librosa's sqrt(2) gain folded in; n_taps == 1 (mod 4).
 * Segment s is the signal  x[j] = audio[d_seg_start[s] + j] for j < d_seg_valid[s], 0 for d_seg_valid[s] <= j < d_seg_len[s]
 * (np.pad of the tail, tablature_generator.py:658-660); its frame count is gtc_scqt_frames(d_seg_len[s], hop, n_octaves).
 * Outputs are [n_seg, n_bins, t_max] with t_max = gtc_scqt_frames(max_len, ...), frames past a segment's own count hold
 * the value of an all-zero frame; max_len >= every d_seg_len[s].
 * ------------------------------------------------------------------------------------------------------------ */
#define GTC_SAMPLES_F32    0
#define GTC_SAMPLES_PCM16  1   /* int16 PCM, converted x/32768 on the device (== librosa.load of a PCM_16 file) */
typedef struct gtc_splan gtc_splan;
int gtc_scqt_frames(int64_t seg_len, int hop_length, int n_octaves);
int gtc_scqt_plan_create(gtc_splan** out, int device, int n_octaves, int n_fft, int hop_length, int n_bins,
                         int filters_per_octave, const float* h_filters, const float* h_taps, int n_taps);
int gtc_scqt_plan_destroy(gtc_splan* plan);
int gtc_scqt_workspace_bytes(const gtc_splan* plan, int64_t n_seg, int64_t max_len, size_t* bytes);
int gtc_scqt_segments_db(const gtc_splan* plan, const void* d_audio, int sample_format, const int64_t* d_seg_start,
                         const int32_t* d_seg_valid, const int32_t* d_seg_len, int64_t n_seg, int64_t max_len,
                         float* d_out_db, void* d_workspace, size_t workspace_bytes,
                         float power, float amin, float top_db, float cut_db, float floor_db, gtc_stream_t stream);
/* One 2:1 stage on its own: d_out[s][k], k < ceil(d_seg_len[s]/2), row stride out_stride floats.  gain = 1 is the CQT's
 * stage (librosa.resample scale=True); gain = 1/sqrt(2) is librosa.load(path, sr=native/2) -- tablature_generator.py:613,650
 * on a 44.1 kHz file. */
int gtc_scqt_decimate(const gtc_splan* plan, const void* d_audio, int sample_format, const int64_t* d_seg_start,
                      const int32_t* d_seg_valid, const int32_t* d_seg_len, int64_t n_seg, int64_t max_len,
                      float* d_out, int64_t out_stride, float gain, gtc_stream_t stream);
/* d_out_c [n_seg, n_bins, t_max, 2] fp32 (== librosa.cqt of every segment) */
int gtc_scqt_segments_complex(const gtc_splan* plan, const void* d_audio, int sample_format, const int64_t* d_seg_start,
                              const int32_t* d_seg_valid, const int32_t* d_seg_len, int64_t n_seg, int64_t max_len,
                              float* d_out_c, void* d_workspace, size_t workspace_bytes, gtc_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Label rasterisation -- replaces GuitarTablatureExtractor.extract_tablature_from_jams + midi_to_tablature +
 * extract_tablature_from_pitch_contour and the stats of process_file (/root/reference/jam_to_tablature.py:55-178,
 * 303-331).  Bit-exact integer output.
 *
 * note events of clip c: indices [d_evt_off[c], d_evt_off[c+1]) of d_onset/d_dur/d_pitch (fp64, as JAMS stores them)
 * contour observations : indices [d_con_off[c], d_con_off[c+1]) of d_con_time/d_con_midi/d_con_conf/d_con_kind
 *                        (d_con_midi = librosa.hz_to_midi(frequency) computed by the host in fp64; kind 1 marks an
 *                        observation whose confidence is None -- the reference raises on it and keeps zeros).
 *                        All contour pointers may be NULL (no fallback data).
 * d_seg_time[n_seg]    : fp64 segment times (jam_to_tablature.py:273-274), d_seg_off as above.
 * d_out [n_seg,6,19] int8 multi-hot; d_stats[3] += {total, with_notes, with_first_string} (int64, caller zeroes).
 * ------------------------------------------------------------------------------------------------------------ */
int gtc_rasterize_tabs(const double* d_onset, const double* d_dur, const double* d_pitch, const int64_t* d_evt_off,
                       const double* d_con_time, const double* d_con_midi, const double* d_con_conf,
                       const int8_t* d_con_kind, const int64_t* d_con_off,
                       const double* d_seg_time, const int64_t* d_seg_off, int64_t n_clips, int64_t n_seg,
                       int8_t* d_out, int64_t* d_stats, gtc_stream_t stream);

/* d_index[n] (may be NULL = identity) selects and orders the batch items out of d_tabs [n_total,6,19].
 * my_dataloader.py:40-44 : int8 (6,19) -> int64 (6,) argmax (first 1 wins; all-zero row -> 0); d_out [n,6]. */
int gtc_labels_argmax(const int8_t* d_tabs, const int64_t* d_index, int64_t n, int64_t* d_out, gtc_stream_t stream);
/* ViT_dataloader.py:54 + default collate: six contiguous (n,19) int64 heads: d_out[6][n][19]. */
int gtc_labels_vit_heads(const int8_t* d_tabs, const int64_t* d_index, int64_t n, int64_t* d_out, gtc_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Patch assembly -- replaces GuitarTabDataset.__getitem__ + default collate of
 * /root/reference/ViT_dataloader.py:27-51 (mode VIT) and the tensor contract of my_dataloader.py:17-21 (mode CNN).
 * d_db [n_total, n_bins, n_frames] fp32 dB features; d_index[n] (may be NULL = identity) selects and orders the
 * items of the batch (the DataLoader's sampler); d_out [n, 3, out_h, out_w] fp32.
 * ------------------------------------------------------------------------------------------------------------ */
int gtc_patches(const float* d_db, const int64_t* d_index, int64_t n, int n_bins, int n_frames,
                int out_h, int out_w, int mode, float* d_out, gtc_stream_t stream);

/* The CNN loader's PICTURE route (my_dataloader.py:10,17-21,29-33): the reference decodes a PNG rendered by new_cqt.py,
 * resizes it to 224 x 224 with PIL (host work, kept on the host: gtc_b200.loaders.load_png_dir does exactly that once per
 * dataset) and applies ToTensor + Normalize per item.  This kernel is the per-batch part: d_rgb [n_total, h, w, 3] uint8
 * (PIL's HWC bytes) -> d_out [n, 3, h, w] fp32 = ((x / 255) - mean[c]) / std[c], the same three correctly rounded fp32
 * operations torchvision performs, for the items d_index[n] selects (NULL = identity).  w must be a multiple of 4. */
int gtc_patches_rgb8(const uint8_t* d_rgb, const int64_t* d_index, int64_t n, int h, int w,
                     float mean_r, float mean_g, float mean_b, float std_r, float std_g, float std_b,
                     float* d_out, gtc_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Batch augmentation and dB normalisation -- replaces the torch op chains of /root/reference/ViT_engine.py:28-117
 * (time_shift :28-42, add_noise :44-47, frequency_mask :49-63, time_mask :65-79, composed by augment_batch :81-93,
 * then db_normalize :112-117) with one read + one write of the batch.
 * d_in/d_out [batch, channels, dim2, dim3] fp32 (dim3 % 4 == 0).  h_ops[n_ops]: DIFFERENT GTC_AUG_* ops in the order
 * the reference would apply them (host array).  Parameters are the values the reference draws with `random`:
 *   shift       out[:, :, h, :] = in[:, :, h + shift, :], zero filled           (int(uniform(-r, r) * dim2), :34)
 *   freq0/width in[:, :, :, freq0:freq0+width] = 0                              (:60-62)
 *   time0/width in[:, :, time0:time0+width, :] = 0                              (:76-78)
 *   noise       + noise_level * N(0,1), Philox4x32-10 keyed by (noise_seed, element) -- same distribution as
 *               torch.randn_like (:46), not the same stream
 * normalize != 0 applies db_normalize(ref_db) last: clamp((x - ref_db) / -ref_db, 0, 1).
 * A non-zero shift cannot run in place (d_in == d_out).
 * ------------------------------------------------------------------------------------------------------------ */
#define GTC_AUG_TIME_SHIFT 1
#define GTC_AUG_NOISE      2
#define GTC_AUG_FREQ_MASK  3
#define GTC_AUG_TIME_MASK  4
int gtc_augment_batch(const float* d_in, float* d_out, int64_t batch, int channels, int dim2, int dim3,
                      const int* h_ops, int n_ops, int shift, int freq0, int freq_width, int time0, int time_width,
                      float noise_level, uint64_t noise_seed, int normalize, float ref_db, gtc_stream_t stream);
/* ViT_engine.py:112-117 and "tablature-generator (1).py":334-335: d_out[i] = clamp((d_in[i] - ref_db) / -ref_db, 0, 1);
 * in place allowed. */
int gtc_db_normalize(const float* d_in, int64_t n, float ref_db, float* d_out, gtc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GTC_H_ */
