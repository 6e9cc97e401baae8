// Test-infrastructure shim (NOT product code, NOT reference code).
// Lets the unmodified reference match/match.cpp compile without Boost by mapping the
// handful of boost::filesystem names it uses (match.cpp:27,351,434-454,473,662) onto
// std::filesystem.
#pragma once
#include <filesystem>
namespace boost { namespace filesystem {
using namespace std::filesystem;
inline path system_complete(const path& p) { return std::filesystem::absolute(p); }
}}
