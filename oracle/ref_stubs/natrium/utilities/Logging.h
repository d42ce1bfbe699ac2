// Stand-in for L/utilities/Logging.h (test infrastructure, see ../../README.md): LOG(level) << ... goes nowhere.
#pragma once
#include <ostream>
namespace natrium {
enum LogLevel { SILENT, ERROR, WARNING, WELCOME, BASIC, DETAILED, ALL, DEBUG };
struct NullLog {
    template <class T>
    NullLog& operator<<(const T&) { return *this; }
    NullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
}  // namespace natrium
#define LOG(level) natrium::NullLog()
