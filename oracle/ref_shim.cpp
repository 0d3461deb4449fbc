// ref_shim.cpp -- the few out-of-line symbols the reference's header templates
// reference when compiled standalone.  TEST INFRASTRUCTURE
// ONLY.  The reference defines these in src/backend/common/err_common.cpp, which
// drags in spdlog/boost and is not buildable offline; only construction and
// destruction are needed because no error path is exercised by the checker.
#include <common/err_common.hpp>

AfError::AfError(const char *const func, const char *const file, const int line, const char *const message,
                 af_err err, boost::stacktrace::stacktrace st)
    : std::logic_error(message), functionName(func), fileName(file), st_(st), lineNumber(line), error(err) {}

AfError::AfError(std::string func, std::string file, const int line, const std::string &message, af_err err,
                 boost::stacktrace::stacktrace st)
    : std::logic_error(message), functionName(func), fileName(file), st_(st), lineNumber(line), error(err) {}

AfError::~AfError() noexcept {}

DimensionError::DimensionError(const char *const func, const char *const file, const int line, const int index,
                               const char *const expectString, const boost::stacktrace::stacktrace &st)
    : AfError(func, file, line, "invalid dimension", AF_ERR_SIZE, st), expected(expectString), argIndex(index) {}
