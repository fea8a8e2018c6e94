/* oracle/juce_shim: stand-in for boost/algorithm/string.hpp -- TEST INFRASTRUCTURE ONLY.
 * Only trim_right is used on a compiled path (fp/tools.cpp:314, DescribeIosFailure). */
#pragma once
#include <string>
#include <cctype>
namespace boost {
namespace algorithm {
    inline void trim_right(std::string& s) {
        while (! s.empty() && std::isspace((unsigned char) s.back())) s.pop_back();
    }
}
using algorithm::trim_right;
}
