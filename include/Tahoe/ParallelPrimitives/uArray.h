// Tahoe/ParallelPrimitives/uArray.h -- host array mirrored lazily by a device buffer
// (reference: Tahoe/ParallelPrimitives/uArray.h:13-228).  Same public interface and the same four-state
// coherence rule: touching the CPU side marks the device copy stale, handing out the device buffer marks
// the CPU copy stale, and a stale side is refreshed (whole-array copy + wait) the next time it is used.
// Pprims does not use it for device scratch any more; it remains the caller-facing container
// (UnitTest/main.cpp:146 holds the CPU reference data in a uArray<SortData>).
#pragma once

#include <Adl/Adl.h>
#include <Tahoe/Math/Math.h>
#include <Tahoe/Math/Array.h>

namespace Tahoe {

template <class T>
class uArray : public Array<T> {
public:
    TH_DECLARE_ALLOCATOR(uArray);

    uArray() : Array<T>(), m_gpuBuff(0), m_status(STATUS_CLEAN) {}
    explicit uArray(int size) : Array<T>((u64)size), m_gpuBuff(0), m_status(STATUS_UNINITIALIZED) {}
    ~uArray() {
        if (m_gpuBuff) {
            adl::DeviceUtils::waitForCompletion(m_gpuBuff->m_device);
            delete m_gpuBuff;
        }
    }

    T& operator[](int idx) { touchCpu(); return Array<T>::operator[]((u64)idx); }
    const T& operator[](int idx) const { const_cast<uArray<T>*>(this)->prepareAccessCpu(); return Array<T>::operator[]((u64)idx); }
    void pushBack(const T& elem) { touchCpu(); Array<T>::pushBack(elem); }
    void clear() { touchCpu(); Array<T>::clear(); }
    void setSize(int size) {
        if ((u64)size == Array<T>::getSize()) return;
        touchCpu();
        Array<T>::setSize((u64)size);
    }
    int getSize() const { return (int)Array<T>::getSize(); }
    T* begin() { touchCpu(); return Array<T>::begin(); }
    const T* begin() const { const_cast<uArray<T>*>(this)->prepareAccessCpu(); return Array<T>::begin(); }

    // The device copy, brought up to date first; the CPU copy is considered stale afterwards.
    const adl::Buffer<T>* getGpuBuffer(const adl::Device* device) {
        prepareAccessGpu(device);
        m_status = STATUS_CPU_DIRTY;
        return m_gpuBuff;
    }
    const adl::Buffer<T>* getGpuBuffer(const adl::Device* device) const { return const_cast<uArray<T>*>(this)->getGpuBuffer(device); }

    void setToLauncher(adl::Launcher& launcher) {
        adl::Launcher::BufferInfo info(getGpuBuffer(launcher.m_deviceData));
        launcher.setBuffers(&info, 1);
    }
    void setToLauncher(adl::Launcher& launcher) const { const_cast<uArray<T>*>(this)->setToLauncher(launcher); }

    void setDataIsClean() { m_status = STATUS_CLEAN; }

protected:
    enum UStatus { STATUS_CPU_DIRTY, STATUS_GPU_DIRTY, STATUS_CLEAN, STATUS_UNINITIALIZED };

    void touchCpu() {
        prepareAccessCpu();
        m_status = STATUS_GPU_DIRTY;
    }
    void prepareAccessCpu() {
        if (m_status == STATUS_CPU_DIRTY && m_gpuBuff) {
            m_gpuBuff->read(Array<T>::begin(), Array<T>::getSize());
            adl::DeviceUtils::waitForCompletion(m_gpuBuff->m_device);
        }
        if (m_status == STATUS_CPU_DIRTY || m_status == STATUS_UNINITIALIZED) m_status = STATUS_CLEAN;
    }
    void prepareAccessGpu(const adl::Device* device) {
        const bool stale = m_status == STATUS_GPU_DIRTY || m_status == STATUS_UNINITIALIZED;
        if (m_gpuBuff && !stale) return;
        if (!m_gpuBuff) m_gpuBuff = new adl::Buffer<T>(device, Array<T>::getSize());
        else if (m_gpuBuff->getSize() < Array<T>::getSize()) m_gpuBuff->setSize(Array<T>::getSize());
        if (m_status != STATUS_UNINITIALIZED) {
            m_gpuBuff->write(Array<T>::begin(), Array<T>::getSize());
            adl::DeviceUtils::waitForCompletion(m_gpuBuff->m_device);
        }
        m_status = STATUS_CLEAN;
    }

    adl::Buffer<T>* m_gpuBuff;
    UStatus m_status;
};

}  // namespace Tahoe
