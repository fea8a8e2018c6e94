// fp/fp_include_all.hpp -- umbrella header, as the reference's fp/fp_include_all.hpp:23-30 (without Boost and
// without ParallelBufferPrinter, which is a debug dumper outside the convolution path).
#pragma once
#include "tools.hpp"
#include "ir.hpp"
#include "convolution.hpp"
#include "CircularBufferArray.hpp"
#include "ExpSineSweep.hpp"
#include "StreamingConvolver.hpp"
#include "PluginConvolver.hpp"
#include "Formats.hpp"
#include "CaptureChain.hpp"

using namespace fp;
#ifndef NOT
#define NOT not
#endif
