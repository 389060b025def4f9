/* fmgpu.h -- C-ABI of the B200-native broadcast-FM demodulation chain.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: the
 * seam is the C++ class Broadcast_FM_Demod (src/fm_demod/broadcast_fm_demod.h:91-299) built by
 * App (src/app.cpp:19) and fed by App::Run (src/app.cpp:56-65).  Every entry point below names
 * the reference interface it replaces; fm_radio_b200/csrc/shim/ holds a header-compatible
 * Broadcast_FM_Demod whose methods forward to these functions.
 *
 * Plain pointers and sizes only.  There is NO CPU fallback: every call fails with
 * FMGPU_ERR_CUDA if no sm_100 device is usable.
 *
 * Threading: one thread at a time per handle (as the reference: Process is called from exactly
 * one runner thread, fm_demod_no_tuner.cpp:179-189).  Different handles are independent.
 */
#ifndef FMGPU_H
#define FMGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmgpu_demod fmgpu_demod;
typedef struct fmgpu_rds fmgpu_rds;
/* one RDS group as RDS_Group_Sync hands it to RDS_Decoder (rds_decoder/rds_group.h) */
typedef struct fmgpu_rds_group {
    uint16_t data[4];
    uint8_t  valid[4];
    uint8_t  type[4];        /* BlockOffsetID: A=0 B=1 C=2 C1=3 D=4 (rds_constants.h:29)       */
} fmgpu_rds_group;

/* The RDS_Database fields beyond PI / PTY / PS / RT (rds_decoder/rds_database.h:26-53) that
 * RDS_Database_Decoder_Handler fills (rds_database_decoder_handler.cpp:30-138) from groups
 * 0A (TA/TP, M/S, DI bits; rds_decoder.cpp:159-245), 4A (clock-time and date, :363-405; the
 * Modified Julian Day conversion is modified_julian_date.h:8-24) and 10A (programme type name,
 * :407-441).  Alternative frequencies are a TODO in the reference handler (:117-119) and are not
 * kept here either. */
typedef struct fmgpu_rds_db_ext {
    char programme_type_name[8];
    int32_t year;                    /* datetime: zero until a 4A group with a valid block C arrives */
    uint8_t day, month, hour, minute;
    int8_t local_time_offset;        /* half hours, signed */
    uint8_t traffic_announcement;    /* rds_database.h:19-24: 0 NONE, 1 EON_INFO, 2 AWAIT_EON_ANNOUNCE, 3 NOW_EON_ANNOUNCE */
    uint8_t is_stereo, is_music, is_artificial_head, is_compressed, is_dynamic_program_type;
    uint8_t ptyn_ab_flag;            /* decoder state (rds_database_decoder_handler.h:12): last 10A A/B flag, 4 = none yet */
} fmgpu_rds_db_ext;

enum {
    FMGPU_OK = 0,
    FMGPU_ERR_ARG = -1,      /* bad argument (null, wrong size, unknown id)            */
    FMGPU_ERR_CUDA = -2,     /* CUDA runtime error or no device; see fmgpu_last_error  */
    FMGPU_ERR_STATE = -3,    /* call not valid in the handle's current state           */
    FMGPU_ERR_SIZE = -4      /* input length != block_size: the reference silently returns
                                (broadcast_fm_demod.cpp:311-313); the shim swallows this  */
};

/* Broadcast_FM_Demod(const int block_size) (broadcast_fm_demod.cpp:59) generalised to a batch of
 * n_streams independent demodulators that share one launch (all state is per stream, exactly as
 * all state is per object in the reference, broadcast_fm_demod.h:143-166). */
typedef struct fmgpu_config {
    int block_size;          /* IQ samples per stream per Process; power of two, >= 1024          */
    int n_streams;           /* >= 1                                                              */
    int device;              /* CUDA ordinal, -1 = current device                                 */
    int keep_intermediates;  /* 1: also produce every GUI-visible buffer (pilot, pll, errors,
                                lpr/lmr, BPSK display arrays) -- what the reference always does.
                                0: lean mode, only audio + RDS symbols leave the kernels          */
    int pipeline_depth;      /* slots of inter-stage buffers for the asynchronous path (>= 1);
                                0 = default (4); forced to 1 when keep_intermediates is set       */
} fmgpu_config;

int  fmgpu_create(const fmgpu_config* cfg, fmgpu_demod** out);
void fmgpu_destroy(fmgpu_demod* h);

/* ---- synchronous block interface (host buffers) ------------------------------------------
 * fmgpu_process_u8 replaces App::Run + Broadcast_FM_Demod::Process (src/app.cpp:56-65,
 * broadcast_fm_demod.cpp:309-328): iq holds n_streams consecutive blocks of block_size (I,Q) u8
 * pairs as rtl-sdr emits them; the (float)u8 - 127.0f unpack is fused into the first kernel.
 * fmgpu_process_cf32 replaces Broadcast_FM_Demod::Process(span<const complex<float>>) itself.
 * Both copy host->device, run the chain, copy audio + RDS symbols (+ intermediates) back into
 * the handle's pinned host mirrors and return when those are valid (the reference notifies its
 * observers synchronously inside Process, broadcast_fm_demod.cpp:326-327).
 * n_samples is the number of complex samples per stream and must equal block_size. */
int fmgpu_process_u8(fmgpu_demod* h, const uint8_t* iq_host, size_t n_samples);
int fmgpu_process_cf32(fmgpu_demod* h, const float* iq_host, size_t n_samples);

/* ---- asynchronous batch interface (device buffers) ---------------------------------------
 * New with the batch driver (no reference counterpart).  iq_dev is a DEVICE pointer to
 * [n_streams][block_size][2] u8.  The block is enqueued on the handle's stage streams and the
 * call returns immediately; up to pipeline_depth blocks may be in flight.  The input buffer must
 * stay valid and unmodified until fmgpu_sync (or until pipeline_depth further enqueues).
 * Outputs of enqueue number k land in ring slot k % pipeline_depth (device), see
 * fmgpu_get_device_buffer. */
int fmgpu_enqueue_u8_device(fmgpu_demod* h, const uint8_t* iq_dev);
/* Same, but iq_host is a (preferably pinned) HOST pointer: the host->device copy is queued on the
 * handle's copy stream and overlaps with the kernels of the blocks already in flight. */
int fmgpu_enqueue_u8_host(fmgpu_demod* h, const uint8_t* iq_host);
/* Same for cf32 input that is already on the device ([n_streams][block_size] complex float), e.g. the
 * channelizer's output.  after_stream (cudaStream_t as void*, may be NULL): work queued on it produces
 * iq_dev. */
int fmgpu_enqueue_cf32_device(fmgpu_demod* h, const float* iq_dev, void* after_stream);
/* Makes cuda_stream wait until the input buffer of the block enqueued pipeline_depth enqueues ago has
 * been consumed, so a producer that recycles pipeline_depth output buffers may overwrite the oldest. */
int fmgpu_stream_wait_input_free(fmgpu_demod* h, void* cuda_stream);
int fmgpu_sync(fmgpu_demod* h);
/* Copies slot's audio + symbols to the pinned host mirrors (asynchronously on the output stream;
 * valid after fmgpu_sync).  What travels is chosen by fmgpu_set_fetch_mask (default: everything):
 * the observers of the reference receive the 32 kHz float frames and the soft symbols
 * (OnAudioOut / OnRDSOut, broadcast_fm_demod.h:226-227); a bulk consumer that wants int16 PCM
 * (fm_scraper.cpp:74-78) or decodes RDS on the device (K6) need not pay for the others.  Symbols travel
 * compacted: only the part of each stream's row that can hold symbols (the timing clock runs at
 * <= 3500 Hz, ted_clock.cpp:31-44).  The symbol counts always travel.  fmgpu_get_buffer refuses a
 * buffer the last fetch left on the device (FMGPU_ERR_STATE). */
#define FMGPU_FETCH_AUDIO_F32   1u   /* GetAudioOut(): 32 kHz float frames */
#define FMGPU_FETCH_PCM_S16     2u   /* int16 PCM at FMGPU_CTL_AUDIO_PCM_RATE_HZ (when that stage is on) */
#define FMGPU_FETCH_RDS_SYMBOLS 4u   /* GetRDSPredSymbols() */
#define FMGPU_FETCH_ALL         7u
int fmgpu_set_fetch_mask(fmgpu_demod* h, unsigned mask);
int fmgpu_fetch_outputs(fmgpu_demod* h, int slot);
/* Makes the handle's first kernel wait for work already queued on an external CUDA stream
 * (cudaStream_t passed as void*, e.g. torch's current stream that produced iq_dev). */
int fmgpu_wait_external_stream(fmgpu_demod* h, void* cuda_stream);
/* Records completion of everything enqueued so far into the external stream (so that stream's
 * events time the whole chain). */
int fmgpu_signal_external_stream(fmgpu_demod* h, void* cuda_stream);

/* ---- buffers: the getters of broadcast_fm_demod.h:242-256 and bpsk_synchroniser.h:76-85 ---- */
typedef enum fmgpu_buffer {
    FMGPU_BUF_AUDIO_OUT = 0,        /* GetAudioOut: Frame<float>[B/32] = L,R interleaved f32   */
    FMGPU_BUF_RDS_PRED_SYM,         /* GetRDSPredSymbols: f32[count]                            */
    FMGPU_BUF_RDS_SYM_COUNT,        /* rds_total_symbols: i32[1]                                */
    /* the rest need keep_intermediates = 1 */
    FMGPU_BUF_FM_DEMOD,             /* fm_demod_buf f32[B/4] (private in the reference)         */
    FMGPU_BUF_FM_OUT_IQ,            /* GetFMOutIQ cf32[B/8]                                     */
    FMGPU_BUF_PILOT,                /* GetPilotOutput cf32[B/8] (after AGC)                     */
    FMGPU_BUF_PLL_DT,               /* pll_dt_buf f32[B/8] (private in the reference)           */
    FMGPU_BUF_PLL,                  /* GetPLLOutput cf32[B/8]                                   */
    FMGPU_BUF_PLL_RAW_PHASE_ERROR,  /* Get_PLL_Raw_Phase_Error_Output f32[B/8]                  */
    FMGPU_BUF_PLL_LPF_PHASE_ERROR,  /* Get_PLL_LPF_Phase_Error_Output f32[B/8]                  */
    FMGPU_BUF_AUDIO_LPR,            /* GetLPRAudioOutput f32[B/32]                              */
    FMGPU_BUF_AUDIO_LMR,            /* GetLMRAudioOutput f32[B/32]                              */
    FMGPU_BUF_RDS,                  /* GetRDSOutput cf32[B/64] (after AGC)                      */
    FMGPU_BUF_RDS_RAW_SYM,          /* GetRDSRawSymbols cf32[count]                             */
    FMGPU_BUF_BPSK_PLL_SYM,         /* BPSK_Synchroniser::GetPLLSymbols cf32[B/64]              */
    FMGPU_BUF_BPSK_ZCD,             /* GetZeroCrossings u8[B/64]                                */
    FMGPU_BUF_BPSK_INT_DUMP_TRIGGER,/* GetIntDumpTriggers u8[B/64]                              */
    FMGPU_BUF_BPSK_TED_RAW_PHASE_ERROR, /* f32[B/64]                                            */
    FMGPU_BUF_BPSK_TED_PI_PHASE_ERROR,  /* f32[B/64]                                            */
    FMGPU_BUF_BPSK_PLL_RAW_PHASE_ERROR, /* f32[B/64]                                            */
    FMGPU_BUF_BPSK_PLL_PI_PHASE_ERROR,  /* f32[B/64]                                            */
    FMGPU_BUF_BPSK_INT_DUMP_FILTER, /* cf32[B/64]                                               */
    /* audio output stage K7 (needs FMGPU_CTL_AUDIO_PCM_RATE_HZ > 0), M = (int)((rate / 32000.f) * (B/32)) */
    FMGPU_BUF_AUDIO_PCM_F32,        /* Frame<float>[M]: Resample() of GetAudioOut (audio/resampled_pcm_player.cpp:37-54) */
    FMGPU_BUF_AUDIO_PCM_S16,        /* Frame<int16_t>[M]: the same frames as Audio_Scraper writes them (fm_scraper.cpp:74-78); pinned mirror */
    FMGPU_BUF_FM_IN,                /* fm_in_buf cf32[B/4]: output of filt_poly_ds_lpf_fm_in (private in the reference; feeds its
                                       FM-in spectrum, broadcast_fm_demod.cpp:414).  keep_intermediates only: the fused kernel
                                       otherwise never stores it */
    FMGPU_BUF_AUDIO_LPR_IQ,         /* temp_audio_buf after the L+R decimator, cf32[B/32] (broadcast_fm_demod.cpp:475; its real part is
                                       GetLPRAudioOutput) -- source of the L+R spectrum (:481).  keep_intermediates only */
    FMGPU_BUF_AUDIO_LMR_IQ,         /* temp_audio_buf after the L-R decimator, cf32[B/32] (:490; its imaginary part is
                                       GetLMRAudioOutput) -- source of the L-R spectrum (:523).  keep_intermediates only */
    FMGPU_BUF__COUNT
} fmgpu_buffer;

/* Host-readable view of `buf` for `stream`, valid until the next process/fetch call on the handle
 * (same lifetime rule as the reference's spans).  *n_elems is in elements of the buffer's type
 * (complex and Frame count as one element). */
int fmgpu_get_buffer(fmgpu_demod* h, int stream, fmgpu_buffer buf, const void** host_ptr, size_t* n_elems);
/* Device pointer of ring slot `slot` for the whole batch ([n_streams][...] stream-major). */
int fmgpu_get_device_buffer(fmgpu_demod* h, int slot, fmgpu_buffer buf, void** dev_ptr, size_t* n_elems_per_stream);

/* Scalars: GetAudioLMRPhaseError (broadcast_fm_demod.h:291) and the two AGC gains. */
typedef enum fmgpu_scalar {
    FMGPU_SCALAR_AUDIO_LMR_PHASE_ERROR = 0,
    FMGPU_SCALAR_AGC_PILOT_GAIN,
    FMGPU_SCALAR_AGC_RDS_GAIN
} fmgpu_scalar;
int fmgpu_get_scalar(fmgpu_demod* h, int stream, fmgpu_scalar which, float* out);

/* ---- controls: Broadcast_FM_Demod_Controls (broadcast_fm_demod.h:64-89) ------------------- */
typedef enum fmgpu_control {
    FMGPU_CTL_AUDIO_OUT = 0,            /* 0 LPR, 1 LMR, 2 STEREO (enum AudioOut, :80)          */
    FMGPU_CTL_AUDIO_STEREO_MIX_FACTOR,  /* float, default 1                                     */
    FMGPU_CTL_USE_DEEMPHASIS,           /* bool, default 0                                      */
    FMGPU_CTL_DEEMPHASIS_TUS,           /* int microseconds; redesigns the 1-pole IIR (:337-352)*/
    FMGPU_CTL_AUDIO_LPR_CUTOFF_HZ,      /* int Hz; redesigns the L+R FIR (:355-370)             */
    FMGPU_CTL_AUDIO_LMR_CUTOFF_HZ,      /* int Hz; redesigns the L-R FIR (:373-388)             */
    FMGPU_CTL_AUDIO_PCM_RATE_HZ         /* int Hz, default 0 = off.  > 0 switches on the audio output stage (kernel K7): every
                                           block's GetAudioOut is resampled from 32 kHz to this rate exactly as
                                           Resampled_PCM_Player::ConsumeBuffer does for the sound device
                                           (audio/resampled_pcm_player.cpp:15-28, 37-54; the drivers use 48000) and converted
                                           to int16 as Audio_Scraper::on_audio_data (fm_scraper.cpp:74-78).  8000..192000. */
} fmgpu_control;
/* Latched at the next process/enqueue call (the reference reads controls at Process entry,
 * UpdateFilters, broadcast_fm_demod.cpp:316).  Applies to every stream of the handle. */
int fmgpu_set_control(fmgpu_demod* h, fmgpu_control which, double value);

/* ---- filter taps: the get_b()/get_a() arrays the reference designers fill ------------------ */
typedef enum fmgpu_filter {
    FMGPU_FILT_FM_IN = 0,     /* filt_poly_ds_lpf_fm_in   64 taps (broadcast_fm_demod.cpp:133-144) */
    FMGPU_FILT_FM_OUT,        /* filt_poly_ds_lpf_fm_out  64 taps (:146-157)                       */
    FMGPU_FILT_HILBERT,       /* filt_hilbert_transform   65 taps (:192-195)                       */
    FMGPU_FILT_AUDIO_LPR,     /* filt_poly_ds_lpf_audio_lpr 128 taps (:240-249)                    */
    FMGPU_FILT_AUDIO_LMR,     /* filt_poly_ds_lpf_audio_lmr 128 taps (:251-261)                    */
    FMGPU_FILT_RDS,           /* filt_poly_ds_lpf_rds     128 taps (:263-274)                      */
    FMGPU_FILT_DEEMPHASIS,    /* b[2], a[2] (:184-190, 337-352)                                    */
    FMGPU_FILT_PEAK_PILOT,    /* b[3], a[3] (:200-213)                                             */
    FMGPU_FILT_PLL_LPF,       /* b[2], a[2] (:215-224)                                             */
    FMGPU_FILT_BPSK_TED_LPF,  /* b[2], a[2] (bpsk_synchroniser.cpp:27-36)                          */
    FMGPU_FILT_BPSK_PLL_LPF,  /* b[2], a[2] (bpsk_synchroniser.cpp:39-48)                          */
    FMGPU_FILT__COUNT
} fmgpu_filter;
/* Arrays are in the reference's memory order (ReverseArray layout, filter_designer.cpp:27-39).
 * `a` may be NULL for FIR filters.  n must equal the filter's length. */
int fmgpu_upload_taps(fmgpu_demod* h, fmgpu_filter which, const float* b, const float* a, int n);
int fmgpu_download_taps(fmgpu_demod* h, fmgpu_filter which, float* b, float* a, int n);

/* Sample rates (broadcast_fm_demod.h:284-288): baseband, fm_in, fm_out, rds, audio. */
int fmgpu_get_rates(fmgpu_demod* h, int rates_hz[5]);
int fmgpu_get_config(fmgpu_demod* h, fmgpu_config* out);
/* Measurement aid: runs n_blocks blocks ONE AT A TIME (no overlap between blocks) on the device
 * input iq_dev with CUDA events recorded on the launching streams around each kernel, and returns
 * the average device time in ms of K1, K2, K3, K4(+K4b), K5, K6.  n_blocks < 0: |n_blocks| blocks
 * enqueued back to back as in production, i.e. the kernels' times while the stages overlap.
 * Advances the demodulator state. */
int fmgpu_profile_stages(fmgpu_demod* h, const uint8_t* iq_dev, int n_blocks, float ms[6]);
/* The same with ms[6] = K7 (audio output stage; 0 when it is off). */
int fmgpu_profile_stages7(fmgpu_demod* h, const uint8_t* iq_dev, int n_blocks, float ms[7]);
/* SM partition of the handle: sms[0] = SMs reserved for the per-stream recurrences (pilot PLL, BPSK
 * synchroniser, RDS bit path), sms[1] = SMs of the FIR stages; {0, 0} when the device is not
 * partitioned (FMGPU_NO_PARTITION=1 in the environment, or the driver has no green contexts). */
int fmgpu_get_partition(fmgpu_demod* h, int sms[2]);
/* Implementation switches for A/B measurements and parity cross-checks (no reference counterpart; the
 * defaults are the production path).  name = "k1_fp32": value 1 runs the u8 FIR + discriminator
 * (app.cpp:56-65, dsp/polyphase_filter.h:41-64, fm_demod/fm_demod.cpp:30-45) on the FP32 FMA pipe instead of
 * the tensor cores (choose before the first block); name = "k5_literal": value 1 runs the BPSK synchroniser
 * (fm_demod/bpsk_synchroniser.cpp:94-186) sample by sample in the reference's order instead of symbol by symbol
 * (identical bits). */
int fmgpu_set_option(fmgpu_demod* h, const char* name, int value);
/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
long long fmgpu_launch_count(fmgpu_demod* h);

/* ---- RDS bit path on the device (kernel K6) -------------------------------------------------
 * Every enqueue/process call also runs, per stream on the device, what App does on the host after
 * each block (app.cpp:66-77): DifferentialManchesterDecoder::Process
 * (rds_decoder/differential_manchester_decoder.h:25-59) -> RDS_Decoding_Chain::Process
 * (rds_decoding_chain.h:24) -> RDS_Group_Sync / CalculateCRC10 / RDS_Decoder::ProcessGroup and the
 * RDS_Database fields (PI/PTY/PS/RT, and fmgpu_rds_db_ext).  Results stay on the device: per stream the running totals,
 * the database, and rings of the most recent groups and 16-byte packets (the rings hold at least
 * what two blocks produce).  fmgpu_rds_device_fetch synchronises the handle and copies them to
 * the host once; the getters then read that copy.  Indices are absolute (0 = first group / byte
 * since creation); a range that has left the ring fails with FMGPU_ERR_STATE.
 * The getters return the number of items copied (>= 0) or a negative error. */
int fmgpu_rds_device_fetch(fmgpu_demod* h);
int fmgpu_rds_device_counts(fmgpu_demod* h, int stream, unsigned long long* n_groups, unsigned long long* n_bytes,
                            int ring_caps[2] /* out, may be NULL: groups, bytes */);
int fmgpu_rds_device_get_groups(fmgpu_demod* h, int stream, unsigned long long first, fmgpu_rds_group* out, int max_groups);
int fmgpu_rds_device_get_bytes(fmgpu_demod* h, int stream, unsigned long long first, uint8_t* out, int max_bytes);
int fmgpu_rds_device_get_db(fmgpu_demod* h, int stream, uint16_t* pi, char ps8[8], char rt64[64], uint8_t* pty);
int fmgpu_rds_device_get_db_ext(fmgpu_demod* h, int stream, fmgpu_rds_db_ext* out);

/* ---- host-side filter designers: src/dsp/filter_designer.h:8-35, same signatures ----------- */
void fmgpu_create_fir_lpf(float* b, int N, float k);
void fmgpu_create_fir_hpf(float* b, int N, float k);
void fmgpu_create_fir_bpf(float* b, int N, float k1, float k2);
void fmgpu_create_fir_hilbert(float* b, int N);
void fmgpu_create_iir_single_pole_lpf(float* b, float* a, float k);
void fmgpu_create_iir_notch_filter(float* b, float* a, float k, float r);
void fmgpu_create_iir_peak_1_filter(float* b, float* a, float k, float r);
/* Method 2 (zero and pole placement, A_db = attenuation outside the peak): dsp/filter_designer.h:27,
 * filter_designer.cpp:312-367.  (The reference's normalisation memoises the first call's parameters in a
 * `static` lambda, :347; this one uses every call's own.) */
void fmgpu_create_iir_peak_2_filter(float* b, float* a, float k, float r, float A_db);
/* The designers' window argument (dsp/filter_designer.h:3,9-11: `const window_func_t window = window_hamming`) and
 * the windows of dsp/window_functions.h:11-38; window = NULL selects Hamming like the reference's default. */
typedef float (*fmgpu_window_func_t)(float);
float fmgpu_window_hamming(float x);
float fmgpu_window_hann(float x);
float fmgpu_window_blackman(float x);
float fmgpu_window_blackman_harris(float x);
void fmgpu_create_fir_lpf_window(float* b, int N, float k, fmgpu_window_func_t window);
void fmgpu_create_fir_hpf_window(float* b, int N, float k, fmgpu_window_func_t window);
void fmgpu_create_fir_bpf_window(float* b, int N, float k1, float k2, fmgpu_window_func_t window);
/* dsp/filter_designer.h:34-36 */
#define FMGPU_TOTAL_TAPS_IIR_SINGLE_POLE_LPF 2
#define FMGPU_TOTAL_TAPS_IIR_SECOND_ORDER_NOTCH_FILTER 3
#define FMGPU_TOTAL_TAPS_IIR_SECOND_ORDER_PEAK_FILTER 3

/* ---- stand-alone resamplers: src/dsp/polyphase_filter.h (PolyphaseDownsampler<T>::process,
 * :41-64) on the GPU.  Host pointers; state (the last M*K inputs) lives in the object. ---------- */
typedef struct fmgpu_polyphase fmgpu_polyphase;
int  fmgpu_polyphase_ds_create(int M, int K, int is_complex, fmgpu_polyphase** out);
void fmgpu_polyphase_destroy(fmgpu_polyphase* f);
float* fmgpu_polyphase_get_b(fmgpu_polyphase* f);       /* host array of M*K taps, like get_b() */
int  fmgpu_polyphase_ds_process(fmgpu_polyphase* f, const float* x_host, float* y_host, int n_out);
/* PolyphaseUpsampler<T>(const float* b, L, K) and ::process(x, y, N) (dsp/polyphase_filter.h:90-185): b holds L*K
 * prototype taps in the designers' order (the constructor's repack and gain L are applied inside); y receives
 * n_in*L samples.  State: the last K inputs.  Destroy with fmgpu_polyphase_destroy. */
int  fmgpu_polyphase_us_create(const float* b, int L, int K, int is_complex, fmgpu_polyphase** out);
int  fmgpu_polyphase_us_process(fmgpu_polyphase* f, const float* x_host, float* y_host, int n_in);

/* ---- the other src/dsp filter classes, stand-alone (host buffers in / out, the work on the device) ----------------
 * FIR_Filter<T>(K)            dsp/fir_filter.h:9-88     y[i] = sum_k b[k] x[i - (K-1) + k]; state: the last K inputs
 * Hilbert_FIR_Filter<float>(K) dsp/hilbert_fir_filter.h:13-47  taps by create_fir_hilbert; y[i] = { x[i - (K-1)/2], FIR(x)[i] } (cf32 out)
 * IIR_Filter<T>(K)            dsp/iir_filter.h:5-89     direct form I, y = sum_i (xn[i] b[i] + yn[i] a[i]) in the reference's
 *                                                       operation order (bit-identical to the scalar program); K >= 2
 * AGC_Filter<cf32>            dsp/agc.h:6-31            block-wise gain: target_power, current_gain, beta are plain fields
 * kind: FMGPU_FILTER_*; is_complex selects T = std::complex<float> (interleaved re, im) over float.  get_b / get_a return
 * host arrays of K floats the caller fills, like get_b() / get_a() (a is NULL for the FIR kinds). */
#define FMGPU_FILTER_FIR     0
#define FMGPU_FILTER_HILBERT 1
#define FMGPU_FILTER_IIR     2
typedef struct fmgpu_dsp_filter fmgpu_dsp_filter;
int  fmgpu_dsp_filter_create(int kind, int K, int is_complex, fmgpu_dsp_filter** out);
void fmgpu_dsp_filter_destroy(fmgpu_dsp_filter* f);
float* fmgpu_dsp_filter_get_b(fmgpu_dsp_filter* f);
float* fmgpu_dsp_filter_get_a(fmgpu_dsp_filter* f);
int  fmgpu_dsp_filter_get_K(const fmgpu_dsp_filter* f);
int  fmgpu_dsp_filter_process(fmgpu_dsp_filter* f, const float* x_host, float* y_host, int n);   /* Hilbert: y_host holds n cf32 */
typedef struct fmgpu_agc { float target_power, current_gain, beta; } fmgpu_agc;          /* defaults 1.0, 0.1, 0.2 (agc.h:9-11) */
void fmgpu_agc_init(fmgpu_agc* g);
int  fmgpu_agc_process(fmgpu_agc* g, const float* x_cf32_host, float* y_cf32_host, int n);

/* ---- audio output helpers, stand-alone (host buffers; the in-chain version is FMGPU_CTL_AUDIO_PCM_RATE_HZ) ----
 * fmgpu_resample_linear = Resample(buf_in, buf_out) of audio/resampled_pcm_player.cpp:37-54 on stereo
 * Frame<float> arrays (2 floats per frame): linear interpolation at read positions j += (float)n_in/(float)n_out.
 * fmgpu_frames_to_s16 = Audio_Scraper::on_audio_data's conversion (fm_scraper.cpp:74-78):
 * (int16_t)(channel * (32767 * 0.95f)), truncating, wrapping like the reference's x86 build when out of range. */
int  fmgpu_resample_linear(const float* frames_in_host, int n_in, float* frames_out_host, int n_out);
int  fmgpu_frames_to_s16(const float* frames_host, size_t n_frames, int16_t* out_host);

/* ---- display spectra: the device half of UpdateFFTCalc
 * (fm_demod/broadcast_fm_demod.cpp:27-40).  fmgpu_calculate_fft = CalculateFFT (dsp/calculate_fft.cpp:43-50; FFTW3f
 * in the reference: the forward DFT X[k] = sum_n x[n] exp(-2 pi i n k / N), unnormalised), optionally followed by
 * InplaceFFTShift (dsp/fftshift.h:21-33), on n complex floats in host memory; n must be a power of two (every block
 * size of the chain is).  fmgpu_get_fft does the same on one stream's copy of a signal buffer of the last processed
 * block where it lies on the device (complex buffers, or real f32 ones taken as complex with zero imaginary part);
 * y_host receives *n_out complex floats.  The dB / averaging step (Calculate_FFT_Mag::Process,
 * dsp/calculate_fft_mag.cpp:11-45) is host code in the reference and stays on the host (the shim runs the
 * reference's own class on these results).  Off the hot path: only when a GUI window raises a trigger. */
int fmgpu_calculate_fft(const float* x_host, float* y_host, int n, int fftshift);
int fmgpu_get_fft(fmgpu_demod* h, int stream, fmgpu_buffer buf, int fftshift, float* y_host, size_t* n_out);

/* ---- RDS bit path on the host (differential_manchester_decoder.h:25-59, rds_group_sync.cpp,
 * crc10.cpp, and the database-filling part of rds_decoder.cpp: groups 0A, 2A, 4A, 10A) ---------------------------------- */
fmgpu_rds* fmgpu_rds_create(void);
void fmgpu_rds_destroy(fmgpu_rds* r);
void fmgpu_rds_push_symbols(fmgpu_rds* r, const float* sym, size_t n);
int  fmgpu_rds_n_groups(const fmgpu_rds* r);
int  fmgpu_rds_get_groups(const fmgpu_rds* r, fmgpu_rds_group* out, int max_groups);
int  fmgpu_rds_n_bytes(const fmgpu_rds* r);
int  fmgpu_rds_get_bytes(const fmgpu_rds* r, uint8_t* out, int max_bytes);
void fmgpu_rds_get_db(const fmgpu_rds* r, uint16_t* pi, char ps8[8], char rt64[64], uint8_t* pty);
void fmgpu_rds_get_db_ext(const fmgpu_rds* r, fmgpu_rds_db_ext* out);

/* ---- wideband channelizer (BASELINE config 4) --------------------------
 * New component: the reference tunes ONE station in hardware and has no channelizer (its TODO on a
 * configurable front end is broadcast_fm_demod.cpp:67).  One wideband u8 IQ capture at fs_in_hz is
 * split into n_channels complex channels at fs_in_hz / decimation, channel c centred on centre_hz[c]:
 *   y_c[i] = sum_{k < n_taps} b[k] * xs_c[(i+1) * decimation - n_taps + k],
 *   xs_c[n] = ((float)u8[n] - 127) * exp(-j 2 pi f_c n / fs)
 * i.e. the reference's unpack (app.cpp:56-65) and PolyphaseDownsampler convention
 * (dsp/polyphase_filter.h:41-64) around a per-channel frequency shift; b is create_fir_lpf
 * (dsp/filter_designer.cpp:84-107).  f_c is quantised to fs / 2^32 so the phase is exact integer
 * arithmetic over any capture length (fmgpu_chan_get_freqs returns the quantised values).
 * Output [n_channels][block_out] complex float on the device is exactly the cf32 input layout of a
 * demodulator handle with n_streams = n_channels, block_size = block_out. */
typedef struct fmgpu_chan fmgpu_chan;
enum { FMGPU_CHAN_MODE_AUTO = 0,     /* tensor cores when the shape allows, else FP32                    */
       FMGPU_CHAN_MODE_TENSOR = 1,   /* tcgen05.mma kind::i8 on the raw u8 capture, exact int32 sums     */
       FMGPU_CHAN_MODE_FP32 = 2 };   /* direct form on the FP32 FMA pipe                                 */
typedef struct fmgpu_chan_config {
    double fs_in_hz;         /* wideband sample rate, e.g. 20 480 000                                   */
    int decimation;          /* D: channel rate = fs_in_hz / D (20 -> the chain's 1.024 MS/s)           */
    int n_taps;              /* prototype low-pass length (tensor mode: multiple of 64, <= 512)          */
    float cutoff_k;          /* prototype cutoff relative to the input Nyquist; 0 = 0.95 / D            */
    int n_channels;
    int block_out;           /* outputs per channel per call (multiple of 128) = the demodulator's block_size;
                                one call consumes block_out * D wideband samples                       */
    int device;              /* CUDA ordinal, -1 = current                                               */
    int mode;                /* FMGPU_CHAN_MODE_*                                                        */
    int ring_depth;          /* output buffers recycled round-robin (0 = 4); >= the demodulator's pipeline_depth */
} fmgpu_chan_config;
int  fmgpu_chan_create(const fmgpu_chan_config* cfg, const double* centre_hz, fmgpu_chan** out);
void fmgpu_chan_destroy(fmgpu_chan* c);
/* Host array of n_taps prototype taps in the reference's order (like get_b()); edits take effect at the next block. */
float* fmgpu_chan_get_b(fmgpu_chan* c);
int  fmgpu_chan_get_config(fmgpu_chan* c, fmgpu_chan_config* out);
/* Quantised centre frequencies (Hz) and their phase increments per input sample (turns * 2^32); either may be NULL. */
int  fmgpu_chan_get_freqs(fmgpu_chan* c, double* centre_hz, uint32_t* phase_inc);
/* Host in, host out (out_host may be NULL); synchronous.  n_in_samples must be block_out * decimation. */
int  fmgpu_chan_process_u8(fmgpu_chan* c, const uint8_t* iq_host, size_t n_in_samples, float* out_host);
/* Device in; asynchronous on the channelizer's stream.  *out_dev (may be NULL) receives the device
 * pointer of the output buffer this block was written to. */
int  fmgpu_chan_enqueue_u8_device(fmgpu_chan* c, const uint8_t* iq_dev, float** out_dev);
/* One wideband block through channelizer and demodulators (demod: n_streams = n_channels,
 * block_size = block_out, pipeline_depth <= ring_depth); asynchronous. */
int  fmgpu_chan_feed_device(fmgpu_chan* c, fmgpu_demod* demod, const uint8_t* iq_dev);
int  fmgpu_chan_wait_external_stream(fmgpu_chan* c, void* cuda_stream);
int  fmgpu_chan_sync(fmgpu_chan* c);
void* fmgpu_chan_stream(fmgpu_chan* c);
long long fmgpu_chan_launch_count(fmgpu_chan* c);

const char* fmgpu_last_error(void);
const char* fmgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FMGPU_H */
