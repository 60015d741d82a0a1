// Tahoe/Base/Config.h -- logging front-end with the reference's macro names (reference:
// Tahoe/Base/Config.h:6-29, Config.inl:25-114).  The reference appends to ./tahoe.log through a
// LogWriter singleton whose filter mask is zero in release builds, i.e. it prints nothing unless
// _DEBUG is defined; the same policy is kept here, writing to stderr instead of a file.
#pragma once

#include <stdarg.h>
#include <stdio.h>

#define TH_MEM_DEBUG_LEVEL 0

#if !defined(TH_LOG_LEVEL)
#if defined(_DEBUG)
#define TH_LOG_LEVEL 2
#else
#define TH_LOG_LEVEL 1
#endif
#endif
#define TH_LOG_FILE "tahoe.log"

namespace Tahoe {

enum LogCategory {
    LOG_BASE = 1 << 0, LOG_ERROR = 1 << 1, LOG_DEBUG = 1 << 2, LOG_IO = 1 << 3, LOG_GPU = 1 << 4,
    LOG_MATERIAL = 1 << 5, LOG_GEOMETRY = 1 << 6, LOG_TEXTURE = 1 << 7, LOG_LIGHT = 1 << 8, LOG_VOLUME = 1 << 9,
};

class LogWriter {
public:
    static LogWriter& getInstance() {
        static LogWriter s_writer;
        return s_writer;
    }
    void setFilter(unsigned mask) { m_filter = mask; }
    void print(LogCategory category, const char* fmt, ...) {
        if (!(m_filter & (unsigned)category)) return;
        va_list ap;
        va_start(ap, fmt);
        vfprintf(stderr, fmt, ap);
        va_end(ap);
    }

private:
#if defined(_DEBUG)
    LogWriter() : m_filter(LOG_BASE | LOG_ERROR | LOG_DEBUG) {}
#else
    LogWriter() : m_filter(0) {}
#endif
    unsigned m_filter;
};

}  // namespace Tahoe

#define TH_LOG_BASE(...) LogWriter::getInstance().print(Tahoe::LOG_BASE, __VA_ARGS__)
#define TH_LOG_ERROR(...) LogWriter::getInstance().print(Tahoe::LOG_ERROR, __VA_ARGS__)
#define TH_LOG_DEBUG(...) LogWriter::getInstance().print(Tahoe::LOG_DEBUG, __VA_ARGS__)
#define TH_LOG_IO(...) LogWriter::getInstance().print(Tahoe::LOG_IO, __VA_ARGS__)
#define TH_LOG_GPU(...) LogWriter::getInstance().print(Tahoe::LOG_GPU, __VA_ARGS__)
#define TH_LOG_MATERIAL(...) LogWriter::getInstance().print(Tahoe::LOG_MATERIAL, __VA_ARGS__)
#define TH_LOG_GEOMETRY(...) LogWriter::getInstance().print(Tahoe::LOG_GEOMETRY, __VA_ARGS__)
#define TH_LOG_TEXTURE(...) LogWriter::getInstance().print(Tahoe::LOG_TEXTURE, __VA_ARGS__)
#define TH_LOG_LIGHT(...) LogWriter::getInstance().print(Tahoe::LOG_LIGHT, __VA_ARGS__)
#define TH_LOG_VOLUME(...) LogWriter::getInstance().print(Tahoe::LOG_VOLUME, __VA_ARGS__)
