// Host build of fm_radio_b200/csrc/k5_core.h (test infrastructure): runs the literal per-sample loop and the
// symbol-wise loop of the K5 kernel on the same input and reports whether symbols and final state agree bit for bit.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o libk5check.so k5_core_check.cpp
#include "../../fm_radio_b200/csrc/k5_core.h"
#include <cstring>

using namespace fm;

struct Row { const float* x; };
struct Fetch { const float* x; K5Sample operator()(int i) const { return K5Sample{ x[2 * i], x[2 * i + 1] }; } };

static K5Coef coef(const float* ted_ba, const float* pll_ba) {
    K5Coef c{};
    const float Fs = 16e3f, Fsym = 2e3f, Ts = 1.0f / Fs;
    c.ted_b0 = ted_ba[0]; c.ted_b1 = ted_ba[1]; c.ted_a0 = ted_ba[2];
    c.pll_b0 = pll_ba[0]; c.pll_b1 = pll_ba[1]; c.pll_a0 = pll_ba[2];
    c.cooldown_N = 4; c.dump_KTs = 1.0f / 4.0f;
    c.ted_KTs = Ts; c.ted_fcenter = Fsym; c.ted_fgain = 1.5e3f; c.mixer_KTs = Ts; c.mixer_fgain = 10.0f;
    c.int_ted_KTs = 10.0f * Ts * (Fsym / Fs); c.int_pll_KTs = c.int_ted_KTs; c.ted_Kp = 0.3f; c.pll_Kp = 0.3f;
    return c;
}

template <int NB>
static int run_step(const K5Coef& c, K5Lane& L, const float* x, int n, float* sym) {
    int pos = 0, total = 0;
    Fetch f{ x };
    K5NoDebug dbg;
    while (pos < n) {
        float sr = 0, si = 0;
        if (k5_symbol_step<NB>(c, L, f, pos, n, true, sr, si, dbg)) { sym[2 * total] = sr; sym[2 * total + 1] = si; total++; }
    }
    return total;
}

extern "C" int k5_check(const float* x, int n, int n_blocks, const float* ted_ba, const float* pll_ba, float gain, int nb,
                        float* sym_lit, int* cnt_lit, float* sym_step, int* cnt_step, float* state_diff)
{
    const K5Coef c = coef(ted_ba, pll_ba);
    K5Lane A{}, Bn{};
    A.gain = Bn.gain = gain;
    int mismatches = 0;
    K5NoDebug dbg;
    for (int b = 0; b < n_blocks; b++) {
        const float* xb = x + (size_t)b * n * 2;
        int ta = 0;
        for (int i = 0; i < n; i++) {
            float sr = 0, si = 0;
            if (k5_sample(c, A, i, xb[2 * i], xb[2 * i + 1], sr, si, dbg)) { sym_lit[((size_t)b * n + ta) * 2] = sr; sym_lit[((size_t)b * n + ta) * 2 + 1] = si; ta++; }
        }
        int tb = 0;
        float* ss = sym_step + (size_t)b * n * 2;
        switch (nb) {
        case 1: tb = run_step<1>(c, Bn, xb, n, ss); break;
        case 3: tb = run_step<3>(c, Bn, xb, n, ss); break;
        case 7: tb = run_step<7>(c, Bn, xb, n, ss); break;
        case 8: tb = run_step<8>(c, Bn, xb, n, ss); break;
        default: return -1;
        }
        cnt_lit[b] = ta; cnt_step[b] = tb;
        if (ta != tb) mismatches++;
        else if (std::memcmp(sym_lit + (size_t)b * n * 2, ss, sizeof(float) * 2 * ta) != 0) mismatches++;
        if (std::memcmp(&A, &Bn, sizeof(K5Lane)) != 0) mismatches++;
    }
    const float* pa = (const float*)&A; const float* pb = (const float*)&Bn;
    for (size_t i = 0; i < sizeof(K5Lane) / 4; i++) state_diff[i] = pa[i] - pb[i];
    return mismatches;
}
