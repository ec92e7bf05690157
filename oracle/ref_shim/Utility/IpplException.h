#pragma once
#include <stdexcept>
#include <string>
class IpplException : public std::runtime_error {
public:
    IpplException(const std::string& where, const std::string& what) : std::runtime_error(where + ": " + what) {}
};
