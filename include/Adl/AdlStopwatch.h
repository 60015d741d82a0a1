// Adl/AdlStopwatch.h -- adl::Stopwatch with the reference's interface (Adl/AdlStopwatch.h:27-83:
// start / split / stop / getMs / getMs(times, capacity) / getNIntervals, CAPACITY = 64 marks).
//
// The reference picks StopwatchHost -- a host clock -- even for the CL device (Adl/AdlStopwatch.inl:16-19), so a
// caller has to waitForCompletion around what it times.  Here a stopwatch bound to the GPU device records CUDA events
// on the device's in-order stream (b200rs_event_*): start/split/stop never block, intervals are DEVICE time, and
// getMs() waits only for the marks it reads.  Without a device (Stopwatch sw; or init(0)) it is the host clock.
#pragma once

#include <chrono>

namespace adl {

struct Stopwatch {
    enum { CAPACITY = 64 };

    Stopwatch(const Device* deviceData = 0) : m_device(0), m_idx(0), m_inited(false) {
        for (int i = 0; i < CAPACITY; ++i) m_events[i] = 0;
        if (deviceData) init(deviceData);
    }
    ~Stopwatch() {
        if (m_device && m_device->getHandle())
            for (int i = 0; i < CAPACITY; ++i)
                if (m_events[i]) b200rs_event_destroy(m_device->getHandle(), m_events[i]);
    }

    void init(const Device* deviceData) {
        ADLASSERT(!m_inited);
        m_device = deviceData;
        m_inited = true;
    }
    void start() {
        if (!m_inited) init(0);
        m_idx = 0;
        mark();
    }
    void split() { mark(); }
    void stop() { mark(); }
    float getMs() { return interval(0); }  // first interval, like StopwatchHost::getMs (AdlStopwatchHost.inl:68-73)
    void getMs(float* times, int capacity) {
        for (int i = 0; i < capacity; ++i) times[i] = i < m_idx - 1 ? interval(i) : 0.f;
    }
    int getNIntervals() const { return m_idx - 1; }

    const Device* m_device;
    int m_idx;

private:
    Stopwatch(const Stopwatch&);
    Stopwatch& operator=(const Stopwatch&);

    bool onDevice() const { return m_device && m_device->getHandle(); }
    void mark() {
        ADLASSERT(m_idx < CAPACITY);
        if (m_idx >= CAPACITY) return;
        if (onDevice()) {
            if (!m_events[m_idx]) adlCheck(b200rs_event_create(m_device->getHandle(), &m_events[m_idx]), "b200rs_event_create");
            adlCheck(b200rs_event_record(m_device->getHandle(), m_events[m_idx]), "b200rs_event_record");
        } else {
            m_host[m_idx] = std::chrono::steady_clock::now();
        }
        ++m_idx;
    }
    float interval(int i) {
        if (i < 0 || i + 1 >= m_idx) return 0.f;
        if (onDevice()) {
            float ms = 0.f;
            adlCheck(b200rs_event_elapsed_ms(m_device->getHandle(), m_events[i], m_events[i + 1], &ms), "b200rs_event_elapsed_ms");
            return ms;
        }
        return std::chrono::duration<float, std::milli>(m_host[i + 1] - m_host[i]).count();
    }

    void* m_events[CAPACITY];
    std::chrono::steady_clock::time_point m_host[CAPACITY];
    bool m_inited;
};

}  // namespace adl
