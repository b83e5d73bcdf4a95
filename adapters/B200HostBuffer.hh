/*
 * B200::HostBuffer -- growable f32 array in page-locked host memory (rb_host_alloc, include/rasr_b200.h).
 *
 * The adapters keep what the reference's scorers keep in `std::vector<f32>` members (the buffered features and the
 * score cache, src/Mm/BatchFeatureScorer.hh:164-166) in this class instead: from pageable memory every copy of
 * rb_gmm_score / rb_nn_score is staged by the driver and synchronous, from page-locked memory the slab pipeline of
 * the library (H2D of slab i+1, kernels of slab i, D2H of slab i-1) runs at the PCIe rate.
 */
#ifndef _B200_HOST_BUFFER_HH
#define _B200_HOST_BUFFER_HH

#include <Core/Types.hh>
#include <algorithm>
#include <cstring>

#include "rasr_b200.h"

namespace B200 {

class HostBuffer {
public:
    HostBuffer()
            : data_(0), size_(0), capacity_(0), pinned_(true) {}
    ~HostBuffer() {
        release(data_, pinned_);
    }
    HostBuffer(const HostBuffer&)            = delete;
    HostBuffer& operator=(const HostBuffer&) = delete;

    size_t size() const {
        return size_;
    }
    f32* data() {
        return data_;
    }
    const f32* data() const {
        return data_;
    }
    f32 operator[](size_t i) const {
        return data_[i];
    }
    bool pinned() const {
        return pinned_;
    }
    void clear() {
        size_ = 0;
    }
    /** contents up to min(old size, n) are kept, new elements are uninitialised */
    void resize(size_t n) {
        reserve(n);
        size_ = n;
    }
    void append(const f32* begin, const f32* end) {
        const size_t n = end - begin;
        reserve(size_ + n);
        std::memcpy(data_ + size_, begin, n * sizeof(f32));
        size_ += n;
    }
    void reserve(size_t n) {
        if (n <= capacity_)
            return;
        const size_t cap = std::max(n, std::max<size_t>(2 * capacity_, 1 << 16));
        void*        p   = 0;
        bool         pin = true;
        if (rb_host_alloc(cap * sizeof(f32), &p) != RB_OK) {  // page-locking failed (ulimit): pageable memory still works
            p   = ::operator new(cap * sizeof(f32));
            pin = false;
        }
        if (size_)
            std::memcpy(p, data_, size_ * sizeof(f32));
        release(data_, pinned_);
        data_     = static_cast<f32*>(p);
        capacity_ = cap;
        pinned_   = pin;
    }

private:
    static void release(f32* p, bool pinned) {
        if (!p)
            return;
        if (pinned)
            rb_host_free(p);
        else
            ::operator delete(p);
    }
    f32*   data_;
    size_t size_, capacity_;
    bool   pinned_;
};

}  // namespace B200

#endif
