/*
 * std_sort_restated.h -- libstdc++'s std::sort (bits/stl_algo.h: __sort, __introsort_loop, __unguarded_partition_pivot,
 * __move_median_to_first, __unguarded_partition, __final_insertion_sort; bits/stl_heap.h for the depth-limit fallback),
 * restated for plain arrays without iterators, recursion on the smaller side replaced by an explicit stack, so that
 * the same code can run on the device.
 *
 * TEST INFRASTRUCTURE (this round): Mm::DensityClustering::selectClusters sorts (distance, cluster) pairs by distance
 * ONLY (src/Mm/DensityClustering.tcc:164-189); std::sort is not stable, so which of several clusters at exactly the
 * same distance ends up among the selected ones is whatever this algorithm does.  Reproducing the reference's choice
 * bit for bit (needed for the int preselection scorer, whose s32 distances tie often) means reproducing this
 * permutation.  tests/test_oracle_gmm.py checks the restatement against the real std::sort on arrays full of ties.
 */
#ifndef ORC_STD_SORT_RESTATED_H
#define ORC_STD_SORT_RESTATED_H

namespace stdsort {

const int kThreshold = 16; /* _S_threshold */

/* test statistic: how often the depth limit was hit (the tests want to see this branch taken) */
inline long& heapSortCalls() {
    static long calls = 0;
    return calls;
}

template<class T, class Less>
inline void adjustHeap(T* first, long holeIndex, long len, T value, Less comp) {
    const long topIndex    = holeIndex;
    long       secondChild = holeIndex;
    while (secondChild < (len - 1) / 2) {
        secondChild = 2 * (secondChild + 1);
        if (comp(first[secondChild], first[secondChild - 1]))
            secondChild--;
        first[holeIndex] = first[secondChild];
        holeIndex        = secondChild;
    }
    if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
        secondChild      = 2 * (secondChild + 1);
        first[holeIndex] = first[secondChild - 1];
        holeIndex        = secondChild - 1;
    }
    /* __push_heap */
    long parent = (holeIndex - 1) / 2;
    while (holeIndex > topIndex && comp(first[parent], value)) {
        first[holeIndex] = first[parent];
        holeIndex        = parent;
        parent           = (holeIndex - 1) / 2;
    }
    first[holeIndex] = value;
}

/* __partial_sort(first, last, last) = __heap_select (only make_heap when middle == last) + __sort_heap */
template<class T, class Less>
inline void heapSort(T* first, T* last, Less comp) {
    ++heapSortCalls();
    const long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        while (true) {
            T value = first[parent];
            adjustHeap(first, parent, len, value, comp);
            if (parent == 0)
                break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        T value = *last; /* __pop_heap(first, last, last) */
        *last   = *first;
        adjustHeap(first, 0L, (long)(last - first), value, comp);
    }
}

template<class T>
inline void swapT(T& a, T& b) {
    T t = a;
    a   = b;
    b   = t;
}

template<class T, class Less>
inline void moveMedianToFirst(T* result, T* a, T* b, T* c, Less comp) {
    if (comp(*a, *b)) {
        if (comp(*b, *c))
            swapT(*result, *b);
        else if (comp(*a, *c))
            swapT(*result, *c);
        else
            swapT(*result, *a);
    }
    else if (comp(*a, *c))
        swapT(*result, *a);
    else if (comp(*b, *c))
        swapT(*result, *c);
    else
        swapT(*result, *b);
}

template<class T, class Less>
inline T* unguardedPartition(T* first, T* last, T* pivot, Less comp) {
    while (true) {
        while (comp(*first, *pivot))
            ++first;
        --last;
        while (comp(*pivot, *last))
            --last;
        if (!(first < last))
            return first;
        swapT(*first, *last);
        ++first;
    }
}

template<class T, class Less>
inline void unguardedLinearInsert(T* last, Less comp) {
    T  val  = *last;
    T* next = last - 1;
    while (comp(val, *next)) {
        *last = *next;
        last  = next;
        --next;
    }
    *last = val;
}

template<class T, class Less>
inline void insertionSort(T* first, T* last, Less comp) {
    if (first == last)
        return;
    for (T* i = first + 1; i != last; ++i) {
        if (comp(*i, *first)) {
            T val = *i;
            for (T* p = i; p != first; --p) /* move_backward(first, i, i + 1) */
                *p = *(p - 1);
            *first = val;
        }
        else
            unguardedLinearInsert(i, comp);
    }
}

template<class T, class Less>
inline void sort(T* first, T* last, Less comp) {
    if (first == last)
        return;
    /* __introsort_loop(first, last, __lg(n) * 2): the recursive call handles [cut, last), the loop continues with
     * [first, cut) -- here the pending right parts wait on a stack (at most one per level of the depth limit) */
    long n = last - first, lg = 0;
    while ((n >> (lg + 1)) > 0)
        ++lg;
    struct Range {
        T*   first;
        T*   last;
        long depth;
    } stack[2 * 64 + 2];
    int top        = 0;
    stack[top++]   = Range{first, last, lg * 2};
    while (top > 0) {
        Range r = stack[--top];
        while (r.last - r.first > kThreshold) {
            if (r.depth == 0) {
                heapSort(r.first, r.last, comp);
                break;
            }
            --r.depth;
            T* mid = r.first + (r.last - r.first) / 2;
            moveMedianToFirst(r.first, r.first + 1, mid, r.last - 1, comp);
            T* cut = unguardedPartition(r.first + 1, r.last, r.first, comp);
            /* the reference recurses into [cut, last) FIRST, then continues with [first, cut): ranges are disjoint, so
             * the order in which they are finished does not change the result */
            stack[top++] = Range{cut, r.last, r.depth};
            r.last       = cut;
        }
    }
    /* __final_insertion_sort */
    if (last - first > kThreshold) {
        insertionSort(first, first + kThreshold, comp);
        for (T* i = first + kThreshold; i != last; ++i)
            unguardedLinearInsert(i, comp);
    }
    else
        insertionSort(first, last, comp);
}

}  // namespace stdsort

#endif
