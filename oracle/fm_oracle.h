/* TEST INFRASTRUCTURE ONLY -- the CPU restatement of the reference FM demodulation chain.
 *
 * Plain scalar C restatement of williamyang98/FM-Radio's hot path (App::Run -> Broadcast_FM_Demod::
 * Process -> DifferentialManchesterDecoder -> RDS_Group_Sync -> RDS_Decoder), every function citing
 * the reference file:line it follows.  It is the checker for the CUDA path; it is NEVER the thing
 * shipped or measured: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * PARITY PIN: the reference has no golden vectors of its own, so this
 * restatement is pinned against the unmodified reference compiled in place (oracle/_ref/libfmref.so,
 * see oracle/ref_harness.cpp) by tests/test_oracle.py, and against the golden fixtures
 * under tests/golden/ that were generated from that library by tests/golden/make_golden.py.
 *
 * The exported functions mirror the harness (prefix fmo_ instead of fmref_) so one Python binding
 * serves both.
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rds_database.h:26-53 beyond PI / PTY / PS / RT, flattened (traffic_announcement = the enum's ordinal, :19-24;
 * ptyn_ab_flag = the handler's AB_flag_programme_type_name, rds_database_decoder_handler.h:12).
 * The harness fills the same layout from the reference's own RDS_Database. */
typedef struct fmo_db_ext {
    char programme_type_name[8];
    int32_t year;
    uint8_t day, month, hour, minute;
    int8_t local_time_offset;
    uint8_t traffic_announcement;
    uint8_t is_stereo, is_music, is_artificial_head, is_compressed, is_dynamic_program_type;
    uint8_t ptyn_ab_flag;
} fmo_db_ext;

void* fmo_create(int block_size);
void fmo_destroy(void* h);
int fmo_process_u8(void* h, const uint8_t* iq);
int fmo_process_cf32(void* h, const float* iq);
int fmo_set_control(void* h, const char* name, double value);
int fmo_get(void* h, const char* name, const void** ptr, size_t* n_elems);
float fmo_get_scalar(void* h, const char* name);
int fmo_get_taps(void* h, const char* name, float* b, float* a);
int fmo_set_taps(void* h, const char* name, const float* b, const float* a);
int fmo_n_groups(void* h);
void fmo_get_groups(void* h, uint16_t* data, uint8_t* valid, uint8_t* type);
int fmo_n_rds_bytes(void* h);
void fmo_get_rds_bytes(void* h, uint8_t* out);
void fmo_get_db(void* h, uint16_t* pi, char* ps8, char* rt64, uint8_t* pty);
void fmo_get_db_ext(void* h, fmo_db_ext* out);

void* fmo_rds_create(void);
void fmo_rds_destroy(void* r);
void fmo_rds_push_symbols(void* r, const float* sym, size_t n);
int fmo_rds_n_groups(void* r);
void fmo_rds_get_groups(void* r, uint16_t* data, uint8_t* valid, uint8_t* type);
int fmo_rds_n_bytes(void* r);
void fmo_rds_get_bytes(void* r, uint8_t* out);
void fmo_rds_get_db(void* r, uint16_t* pi, char* ps8, char* rt64, uint8_t* pty);
void fmo_rds_get_db_ext(void* r, fmo_db_ext* out);

void fmo_create_fir_lpf(float* b, int N, float k);
void fmo_create_fir_hpf(float* b, int N, float k);
void fmo_create_fir_bpf(float* b, int N, float k1, float k2);
void fmo_create_fir_hilbert(float* b, int N);
void fmo_create_iir_single_pole_lpf(float* b, float* a, float k);
void fmo_create_iir_notch_filter(float* b, float* a, float k, float r);
void fmo_create_iir_peak_1_filter(float* b, float* a, float k, float r);
void fmo_create_iir_peak_2_filter(float* b, float* a, float k, float r, float A_db);
void fmo_create_fir_lpf_window(float* b, int N, float k, int window_id);   /* 0 hamming, 1 hann, 2 blackman, 3 blackman-harris */

void fmo_polyphase_ds_f32(int M, int K, const float* b, const float* x, float* y, int N_out, int n_calls);
void fmo_polyphase_ds_cf32(int M, int K, const float* b, const float* x, float* y, int N_out, int n_calls);
void fmo_polyphase_us_f32(int L, int K, const float* b, const float* x, float* y, int N_in, int n_calls);
void fmo_resample_linear(const float* in, int n_in, float* out, int n_out);
/* stand-alone src/dsp filter classes, n_calls consecutive blocks of N samples through one object (state carried) */
void fmo_fir_f32(int K, const float* b, const float* x, float* y, int N, int n_calls);
void fmo_fir_cf32(int K, const float* b, const float* x, float* y, int N, int n_calls);
void fmo_hilbert_f32(int K, const float* x, float* y_cf32, int N, int n_calls);
void fmo_iir_f32(int K, const float* b, const float* a, const float* x, float* y, int N, int n_calls);
void fmo_iir_cf32(int K, const float* b, const float* a, const float* x, float* y, int N, int n_calls);
void fmo_agc_cf32(float target_power, float beta, float gain0, const float* x, float* y, int N, int n_calls, float* gains_out);
void fmo_frames_to_s16(const float* frames, size_t n_frames, int16_t* out);
void fmo_fft_f64(const float* x_cf32, double* y_c64, int n, int fftshift);
void fmo_fft_mag_process(int mode, float beta, const float* x_cf32, float* y, int n);

/* wideband channelizer oracle (config 4; no reference counterpart): float64 shift + decimating FIR */
void fmo_channelize_f64(const uint8_t* iq, size_t n_in, const uint8_t* hist, uint64_t n0, int D, int NN,
                        const float* b, const uint32_t* inc, int n_ch, double* out);

#ifdef __cplusplus
}
#endif
#endif
