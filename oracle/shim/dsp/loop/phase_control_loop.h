// TEST INFRASTRUCTURE (oracle/_ref build only): stand-in for SDR++ core's <dsp/loop/phase_control_loop.h>
// (AlexandreRouma/SDRPlusPlus core/src/dsp/loop/phase_control_loop.h, restated from its published behaviour:
// second-order loop, frequency clamped, phase wrapped into [minPhase, maxPhase] by whole turns).  Unpinned.
#pragma once
#include <math.h>
namespace dsp::loop {
    template <class T, bool CLAMP_PHASE = true>
    class PhaseControlLoop {
    public:
        PhaseControlLoop() {}
        void init(T alpha, T beta, T phase, T minPhase, T maxPhase, T freq, T minFreq, T maxFreq) {
            setCoefficients(alpha, beta);
            setPhaseLimits(minPhase, maxPhase);
            setFreqLimits(minFreq, maxFreq);
            this->phase = phase;
            this->freq = freq;
        }
        static inline void criticallyDamped(T bandwidth, T& alpha, T& beta) {
            T dampningFactor = sqrt(2.0) / 2.0;
            T denominator = (1.0 + 2.0 * dampningFactor * bandwidth + bandwidth * bandwidth);
            alpha = (4 * dampningFactor * bandwidth) / denominator;
            beta = (4 * bandwidth * bandwidth) / denominator;
        }
        void setCoefficients(T alpha, T beta) { _alpha = alpha; _beta = beta; }
        void setPhaseLimits(T minPhase, T maxPhase) { _minPhase = minPhase; _maxPhase = maxPhase; _phaseDelta = _maxPhase - _minPhase; }
        void setFreqLimits(T minFreq, T maxFreq) { _minFreq = minFreq; _maxFreq = maxFreq; }
        inline void advance(T error) {
            freq += _beta * error;
            if (freq > _maxFreq) { freq = _maxFreq; }
            else if (freq < _minFreq) { freq = _minFreq; }
            phase += freq + (_alpha * error);
            if constexpr (CLAMP_PHASE) { clampPhase(); }
        }
        inline void clampPhase() {
            while (phase > _maxPhase) { phase -= _phaseDelta; }
            while (phase < _minPhase) { phase += _phaseDelta; }
        }
        T freq;
        T phase;
    protected:
        T _alpha, _beta, _minPhase, _maxPhase, _phaseDelta, _minFreq, _maxFreq;
    };
}
