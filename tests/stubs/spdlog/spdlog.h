// stand-in for spdlog::error(fmt, args...) -- test scaffolding only
#pragma once
namespace spdlog {
template <typename... A>
void error(const char*, A&&...) {}
} // namespace spdlog
