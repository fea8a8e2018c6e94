/* oracle/juce_shim: stand-in for boost/filesystem.hpp -- TEST INFRASTRUCTURE ONLY.
 * The fp/ sources on the convolution path never call into boost::filesystem (only
 * ParallelBufferPrinter.cpp and PluginProcessor.cpp do, and those are not compiled). */
#pragma once
#include <string>
namespace boost { namespace filesystem {
    inline bool remove(const std::string&) { return false; }
    inline bool exists(const std::string&) { return false; }
} }
