// Front end: zero-phase IIR (K1), FIR decimation (K2), multitaper PSD (K6).
#include "hs_internal.h"

using namespace hs;

extern "C" {

size_t hs_filtfilt_ws_bytes(int n_sig, int64_t n) { (void)n_sig; (void)n; return 0; }
int hs_iir_filtfilt_f64(double*, int, int64_t, int64_t, int64_t, const double*, const double*, int, int, int, void*, void*) {
    return set_error(HS_ERR_UNSUPPORTED, "hs_iir_filtfilt_f64: not built yet");
}
int hs_fir_decimate_f64(const double*, int, int64_t, int64_t, int, const double*, int, double*, int64_t, void*) {
    return set_error(HS_ERR_UNSUPPORTED, "hs_fir_decimate_f64: not built yet");
}
size_t hs_mt_psd_ws_bytes(int, int64_t, int) { return 0; }
int hs_mt_psd_f64(const double*, int, int64_t, const double*, const double*, int, int, int, double*, void*, void*) {
    return set_error(HS_ERR_UNSUPPORTED, "hs_mt_psd_f64: not built yet");
}

}
