// introsort.cuh -- libstdc++'s std::sort (bits/stl_algo.h: __sort -> __introsort_loop + __final_insertion_sort, median
// of three to the front, unguarded Hoare partition, heap sort below the depth limit 2 * floor(log2 n), insertion sort
// for ranges of at most 16) for plain arrays, callable from host and device.
//
// Why reproduce a particular sort: Mm::DensityClustering::selectClusters (src/Mm/DensityClustering.tcc:164-189) sorts
// (distance, cluster) pairs with a comparison that looks at the distance only and takes the first `select` entries.
// std::sort is not stable, so which of several equally distant clusters ends up in front of the boundary is decided
// by the exact sequence of swaps this algorithm performs.  The int preselection scorer's s32 distances tie often;
// bit-identical scores need the same permutation.  The recursion (right part first, loop on the left part) is
// replaced by an explicit stack: the sub-ranges are disjoint, so the order they are finished in does not matter.
#pragma once

#ifdef __CUDACC__
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline
#endif

namespace rb {
namespace introsort {

template<class T>
RB_HD void exchange(T& a, T& b) {
    T t = a;
    a   = b;
    b   = t;
}

// __adjust_heap followed by __push_heap
template<class T, class Less>
RB_HD void adjust_heap(T* first, long hole, long len, T value, Less less) {
    const long top   = hole;
    long       child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(first[child], first[child - 1]))
            --child;
        first[hole] = first[child];
        hole        = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child       = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole        = child - 1;
    }
    long parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) {
        first[hole] = first[parent];
        hole        = parent;
        parent      = (hole - 1) / 2;
    }
    first[hole] = value;
}

// __partial_sort(first, last, last): __make_heap, then __sort_heap
template<class T, class Less>
RB_HD void heap_sort(T* first, T* last, Less less) {
    const long len = last - first;
    if (len >= 2) {
        for (long parent = (len - 2) / 2;; --parent) {
            adjust_heap(first, parent, len, first[parent], less);
            if (parent == 0)
                break;
        }
    }
    while (last - first > 1) {
        --last;
        T value = *last;
        *last   = *first;
        adjust_heap(first, 0L, (long)(last - first), value, less);
    }
}

// __move_median_to_first
template<class T, class Less>
RB_HD void median_to_first(T* result, T* a, T* b, T* c, Less less) {
    if (less(*a, *b)) {
        if (less(*b, *c))
            exchange(*result, *b);
        else if (less(*a, *c))
            exchange(*result, *c);
        else
            exchange(*result, *a);
    }
    else if (less(*a, *c))
        exchange(*result, *a);
    else if (less(*b, *c))
        exchange(*result, *c);
    else
        exchange(*result, *b);
}

// __unguarded_linear_insert
template<class T, class Less>
RB_HD void linear_insert(T* last, Less less) {
    T  value = *last;
    T* next  = last - 1;
    while (less(value, *next)) {
        *last = *next;
        last  = next;
        --next;
    }
    *last = value;
}

// __insertion_sort
template<class T, class Less>
RB_HD void insertion_sort(T* first, T* last, Less less) {
    if (first == last)
        return;
    for (T* i = first + 1; i != last; ++i) {
        if (less(*i, *first)) {
            T value = *i;
            for (T* p = i; p != first; --p)
                *p = *(p - 1);
            *first = value;
        }
        else
            linear_insert(i, less);
    }
}

// kStack: capacity of the stack of waiting right parts; their depth budgets decrease strictly from bottom to top, so
// 2 * floor(log2 n) + 1 entries are enough (130 covers every n a long can hold, 18 covers n <= 256)
template<int kStack = 130, class T, class Less>
RB_HD void sort(T* first, T* last, Less less) {
    const long n = last - first;
    if (n <= 0)
        return;
    long lg = 0;
    while ((n >> (lg + 1)) > 0)
        ++lg;
    struct Range {
        T*  first;
        T*  last;
        int depth;
    };
    Range pending[kStack];
    int   top      = 0;
    pending[top++] = Range{first, last, (int)(2 * lg)};
    while (top > 0) {
        Range r = pending[--top];
        while (r.last - r.first > 16) {
            if (r.depth == 0) {
                heap_sort(r.first, r.last, less);
                break;
            }
            --r.depth;
            T* mid = r.first + (r.last - r.first) / 2;
            median_to_first(r.first, r.first + 1, mid, r.last - 1, less);
            // __unguarded_partition(first + 1, last, pivot = first)
            T* lo = r.first + 1;
            T* hi = r.last;
            for (;;) {
                while (less(*lo, *r.first))
                    ++lo;
                --hi;
                while (less(*r.first, *hi))
                    --hi;
                if (!(lo < hi))
                    break;
                exchange(*lo, *hi);
                ++lo;
            }
            pending[top++] = Range{lo, r.last, r.depth};
            r.last         = lo;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        insertion_sort(first, first + 16, less);
        for (T* i = first + 16; i != last; ++i)
            linear_insert(i, less);
    }
    else
        insertion_sort(first, last, less);
}

}  // namespace introsort
}  // namespace rb
