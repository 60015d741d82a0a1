// Tahoe/Math/Array.h -- growable host array with the reference's interface
// (reference: Tahoe/Math/Array.h:22-98): Array<T>(n), operator[], begin/end, pushBack, setSize ...
// Storage comes from the ALLOCATOR singleton (default TH_MEM_ALLOCATOR = malloc/free); growth
// at least doubles the capacity and moves the old contents with memcpy, as the reference does,
// so T is expected to be trivially relocatable (u32, int, SortData ...).
#pragma once

#include <string.h>
#include <new>

#include <Tahoe/Math/Error.h>
#include <Tahoe/Math/Math.h>
#include <Tahoe/Base/Memory/AllocatorBase.h>

namespace Tahoe {

template <typename T, typename ALLOCATOR = TH_MEM_ALLOCATOR>
class Array {
public:
    Array() : m_data(0), m_size(0), m_capacity(0) { reserve(DEFAULT_SIZE); }
    explicit Array(u64 size) : m_data(0), m_size(0), m_capacity(0) {
        reserve(size);
        m_size = size;
    }
    virtual ~Array() {
        if (m_data) ALLOCATOR::getInstance().deallocate(m_data);
        m_data = 0;
    }

    T& operator[](u64 idx) {
        ADLASSERT(idx < m_size);
        return m_data[idx];
    }
    const T& operator[](u64 idx) const {
        ADLASSERT(idx < m_size);
        return m_data[idx];
    }

    void pushBack(const T& elem) {
        if (m_size == m_capacity) reserve(m_capacity ? 2 * m_capacity : (u64)DEFAULT_SIZE);
        m_data[m_size++] = elem;
    }
    T pop() {
        ADLASSERT(m_size > 0);
        return m_data[--m_size];
    }
    void popBack() {
        ADLASSERT(m_size > 0);
        --m_size;
    }
    T& expandOne() {
        setSize(m_size + 1);
        return *new (&m_data[m_size - 1]) T;
    }
    void clear() { m_size = 0; }
    void setSize(u64 size) {
        if (size > m_capacity) reserve(max2(size, 2 * m_capacity));
        m_size = size;
    }
    void removeAt(u64 idx) {  // order is not preserved: the last element fills the hole
        ADLASSERT(idx < m_size);
        m_data[idx] = m_data[--m_size];
    }
    u64 indexOf(const T& value) const {
        for (u64 i = 0; i < m_size; ++i)
            if (m_data[i] == value) return i;
        return (u64)-1;
    }

    bool isEmpty() const { return m_size == 0; }
    u64 getSize() const { return m_size; }
    T* begin() { return m_data; }
    const T* begin() const { return m_data; }
    T* end() { return m_data + m_size; }
    const T* end() const { return m_data + m_size; }

protected:
    enum { DEFAULT_SIZE = 128, INCREASE_SIZE = 128 };

    void reserve(u64 capacity) {
        if (capacity <= m_capacity) return;
        T* fresh = (T*)ALLOCATOR::getInstance().allocate(sizeof(T) * (capacity ? capacity : 1), "Array", __LINE__);
        ADLASSERT(fresh != 0);
        if (m_data) {
            memcpy((void*)fresh, (const void*)m_data, sizeof(T) * m_size);
            ALLOCATOR::getInstance().deallocate(m_data);
        }
        for (u64 i = m_size; i < capacity; ++i) new (&fresh[i]) T;
        m_data = fresh;
        m_capacity = capacity;
    }

    T* m_data;
    u64 m_size;
    u64 m_capacity;

private:
    Array(const Array&);             // not copyable (as in the reference)
    Array& operator=(const Array&);
};

template <typename T>
class GlobalArray : public Array<T, DefaultAllocator> {
public:
    virtual ~GlobalArray() {}
};

}  // namespace Tahoe
