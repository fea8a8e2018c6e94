// fp/convolution.hpp -- drop-in for the reference header of the same name (fp/convolution.hpp:14-48):
// same namespace, function names, argument meaning, default arguments and error behaviour; the work is
// done on a B200 through the C ABI in include/irb_b200.h.  No host compute path exists behind these calls.
#pragma once
#ifdef IRB_USE_REAL_JUCE
#include <JuceHeader.h>
#else
#include "juce_stub/JuceHeader.h"
#endif

namespace fp {

// same enumerators, same order as fp/convolution.hpp:18-24
enum ChannelLayout { unknown, IRMonoAudioMono, IRMonoAudioStereo, IRStereoAudioMono, IRStereoAudioStereo };

namespace convolution {

// Uniformly partitioned FFT convolution of whole buffers; buffer1 = audio, buffer2 = impulse response.
// Returns [audio channels][len1 + len2 - 1]; mono/stereo x mono/stereo only -- anything else prints a debug
// line and returns the cleared buffer, as fp/convolution.cpp:39-42 does.
AudioBuffer<float> convolvePeriodic(AudioBuffer<float>& buffer1, AudioBuffer<float>& buffer2, int processBlockSize = 256);

}  // namespace convolution
}  // namespace fp
