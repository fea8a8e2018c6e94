// fp/ir.hpp -- drop-in for the reference's fp/ir.hpp:17-38 (impulse-response manipulation).
#pragma once
#include "tools.hpp"

namespace fp {
namespace ir {

AudioBuffer<float> invertFilter(AudioBuffer<float>& buffer, int samplerate);       // deconvolve(pulse, buffer) on the GPU
// peak search with wrap-around, threshold run-length chop, quarter-length fades (host: a short sequential scan)
AudioBuffer<float> IRchop(AudioBuffer<float>& buffer, int IRlength, float thresholdLeveldB, int consecutiveSamplesBelowThreshold);
void shifteroo(AudioBuffer<float>* buffer);                                        // second half in front of the first
// partitioned spectra {re0, re(N/2), re1, im1, ...} per partition of irPartSize samples -- the engine's own packed
// FDL/IR format, computed by the block-FFT kernel.  (The reference also dumps them to a hard-coded text file; not done.)
AudioBuffer<float> IRtoRealFFTRaw(AudioBuffer<float>& buffer, int irPartSize);

}  // namespace ir
}  // namespace fp
