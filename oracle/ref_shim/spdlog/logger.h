// Stand-in for <spdlog/logger.h>: the reference only creates a named logger (constraints.hpp:20-24).
#pragma once
#include <memory>
#include <string>
namespace spdlog {
class logger {
public:
    explicit logger(std::string n) : m_name(std::move(n)) {}
    template <typename... A> void trace(A&&...) {}
    template <typename... A> void debug(A&&...) {}
    template <typename... A> void info(A&&...) {}
    template <typename... A> void warn(A&&...) {}
    template <typename... A> void error(A&&...) {}
    template <typename... A> void critical(A&&...) {}
private:
    std::string m_name;
};
}  // namespace spdlog
