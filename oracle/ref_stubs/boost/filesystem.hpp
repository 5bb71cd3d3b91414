// Stand-in: boost::filesystem is std::filesystem here (test infrastructure).
#pragma once
#include <filesystem>
namespace boost {
namespace filesystem = std::filesystem;
}
