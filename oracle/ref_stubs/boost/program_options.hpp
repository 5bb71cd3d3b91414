// Stand-in for the slice of Boost.Program_options that modules/io/track_mem.{h,cpp} touches (test infrastructure):
// the --max-mem option is declared there but oracle/_ref never parses a command line.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "boost/lexical_cast.hpp"
namespace boost {
struct any {};
template <class E> [[noreturn]] void throw_exception(const E& e) { throw e; }
namespace program_options {
struct invalid_option_value : public std::runtime_error {
  explicit invalid_option_value(const std::string& s) : std::runtime_error("invalid option value: " + s) {}
};
namespace validators {
inline void check_first_occurrence(const any&) {}
inline std::string get_single_string(const std::vector<std::string>& xs) { return xs.empty() ? std::string() : xs[0]; }
}  // namespace validators
struct value_semantic {};
template <class T> value_semantic* value(T*) { return nullptr; }
class options_description {
 public:
  struct adder {
    adder& operator()(const char*, const value_semantic*, const char*) { return *this; }
    adder& operator()(const char*, const char*) { return *this; }
  };
  adder add_options() { return adder(); }
};
}  // namespace program_options
}  // namespace boost
