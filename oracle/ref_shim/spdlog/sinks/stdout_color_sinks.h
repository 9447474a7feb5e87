#pragma once
#include <spdlog/logger.h>
namespace spdlog {
inline std::shared_ptr<logger> stdout_color_mt(const std::string& name) { return std::make_shared<logger>(name); }
}
