// tiramisu/debug.h — the generators only use these as no-op tracing hooks
// (reference: include/tiramisu/debug.h:20-119).
#ifndef TIRAMISU_B200_FRONTEND_DEBUG_H
#define TIRAMISU_B200_FRONTEND_DEBUG_H
#include <iostream>
#include <string>
#define ENABLE_DEBUG false
#define DEBUG(level, stmt) do { } while (0)
#define DEBUG_NO_NEWLINE(level, stmt) do { } while (0)
#define DEBUG_FCT_NAME(level) do { } while (0)
#define DEBUG_INDENT(x) do { } while (0)
#define DEBUG_NEWLINE(level) do { } while (0)
#define DEBUG_NO_NEWLINE_NO_INDENT(level, stmt) do { } while (0)
#define ERROR(message, exit_program) do { std::cerr << "Error in " << __FILE__ << ":" << __LINE__ << " - " << (message) << "\n"; if (exit_program) std::exit(1); } while (0)
namespace tiramisu { void str_dump(const std::string &s); }
#endif
