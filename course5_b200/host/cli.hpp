// cli.hpp — the `course` command line (main.cpp:17-69), re-hosted without Boost.ProgramOptions.
// Same option names, short forms, defaults and help text; attached short values (-j16), separate
// values (-j 16), --long value, --long=value and unambiguous long prefixes are accepted, like
// Boost's default style.
#pragma once

#include <iosfwd>
#include <string>

#include "config.hpp"

namespace c5host {

enum class cli_result { run, exit_ok, exit_error };

// Fills cfg. Prints help / errors to `out` exactly where the reference does:
//   --help                      -> usage, exit_ok            (main.cpp:39-42,75-78: returns 0)
//   missing -f or -d            -> "Error! Source filename ..." + usage, exit_ok (main.cpp:47-51)
//   unknown option / bad value  -> message, exit_error (the reference dies on an uncaught Boost exception)
cli_result program_options(int argc, char** argv, config_str& cfg, std::ostream& out);

void print_usage(std::ostream& out);

} // namespace c5host
