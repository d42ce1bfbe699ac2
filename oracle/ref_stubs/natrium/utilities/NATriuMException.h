// Stand-in for L/utilities/NATriuMException.h (test infrastructure, see ../../README.md): same classes, no logging / MPI.
#pragma once
#include <exception>
#include <string>
#include "BasicNames.h"
#include "Logging.h"
namespace natrium {
class NATriuMException : public std::exception {
    std::string message;
public:
    NATriuMException(const char* msg) : message(msg) {}
    NATriuMException(const std::string& msg) : message(msg) {}
    ~NATriuMException() throw() {}
    virtual const char* what() const throw() { return message.c_str(); }
};
class NotImplementedException : public NATriuMException {
public:
    NotImplementedException(const char* msg) : NATriuMException(msg) {}
    NotImplementedException(const std::string& msg) : NATriuMException(msg) {}
    ~NotImplementedException() throw() {}
};
}  // namespace natrium
