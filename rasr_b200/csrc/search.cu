// search.cu -- the score consumer of config C5: Search::LinearSearch (time-synchronous Viterbi over the linear HMMs of
// all pronunciations, unigram LM, one book-keeping entry per time frame) fed from the dense score matrix on the device.
//
//   LinearSearch::feed            src/Search/LinearSearch.cc:233-379
//   LinearSearch::bookKeeping     src/Search/LinearSearch.cc:381-432
//   restart / Book                src/Search/LinearSearch.cc:145-152,224-231,487-493
//   getCurrentBestSentence        src/Search/LinearSearch.cc:438-468
//
// The reference calls  emissionScores->score(mixture)  through two virtual calls per HMM state and frame; here the
// scores of a whole segment already sit in HBM ([T x nEmissions], written by the scorers of this library) and one
// CTA walks one segment: a thread owns a word, keeps the recurrence of feed() for its states in place (descending
// state order, so the previous frame's values of states s, s-1, s-2 are still intact), then the warp 0 replays the
// book-keeping scan over the word ends IN WORD ORDER -- its acceptance test compares against (score - lm) + lm of the
// current best, which is not an associative minimum, so a tree reduction would not be bit-identical -- skipping 32
// words at a time when none of them beats the current threshold.  Segments are independent (one CTA each); the time
// loop is inherently sequential.  All arithmetic is the reference's f32 arithmetic in its order: word sequences, word
// end times and both scores are bit-identical to the CPU path.
#include <cfloat>
#include <cmath>
#include <algorithm>
#include <array>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kThreads = 1024;

struct SearchParams {
    // lexicon
    const uint32_t* wordOff;    // [W+1]
    const uint32_t* stateEmis;  // [nStates]
    const uint32_t* stateTdp;   // [nStates]
    const float*    tdp;        // [nModels * 4]
    const float*    unigram;    // [W]
    uint32_t        W, nStates, entryModel, nModels;
    // input
    const float*   scores;    // [frames * nEmis]
    const int64_t* frameOff;  // [U+1]
    int            nEmis;
    // per segment state, stride nStates + W entries (index 0 of a word = entry hypothesis)
    float* hypScore;
    float* hypLm;
    int*   hypBkp;
    // per segment book (capacity = frames of the segment, laid out at frameOff[u])
    int4*  books;  // [frames][2]: {score, lmScore, word, bkp}, {time, -, -, -}; entries of a segment start at its first frame
    int*   nBooks;  // [U]
    // word-end candidates of the current frame, per segment [W]
    float* endScore;
    int    useSmem;  // hypotheses, lexicon tables and word-end candidates of the segment live in shared memory
    // single-word recognition (:248-258, 355-370, 390-395): the "words" above are then the ENTRIES of the search
    // (an irregular pronunciation is followed by its irregular-chain copy)
    int             single;
    const uint8_t*  flags;      // [W] bit 0: regular word, bit 1: irregular-chain entry
    const uint32_t* entryWord;  // [W] word of the caller's lexicon, written into the book
    const uint32_t* irrList;    // [nIrr] the entries that are not regular words, ascending
    uint32_t        nIrr;
    int4*           irrBooks;   // the second book: best sequences of irregular words only; same layout as books
    int*            nIrrBooks;  // [U]
};

// back pointers name an entry of either book: >= 0 the main book, <= -2 entry (-2 - bkp) of the irregular book
__device__ __forceinline__ int irr_ref(int i) {
    return -2 - i;
}

__global__ void __launch_bounds__(kThreads) linear_search_kernel(const SearchParams p) {
    const int      u      = blockIdx.x;
    const int64_t  f0     = p.frameOff[u];
    const int      T      = (int)(p.frameOff[u + 1] - f0);
    const size_t   stride = (size_t)p.nStates + p.W;
    float*          hs     = p.hypScore + (size_t)u * stride;
    float*          hl     = p.hypLm + (size_t)u * stride;
    int*            hb     = p.hypBkp + (size_t)u * stride;
    float*          es     = p.endScore + (size_t)u * p.W;
    const uint32_t* wordOff   = p.wordOff;
    const uint32_t* stateEmis = p.stateEmis;
    const uint32_t* stateTdp  = p.stateTdp;
    const float*    tdp       = p.tdp;
    const float*    unigram   = p.unigram;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    if (p.useSmem) {
        // the working set of a segment (a few hundred KB at most for a 1000-word lexicon) is latency critical: every
        // frame is a dependent step
        float* f = reinterpret_cast<float*>(smemRaw);
        hs       = f;
        hl       = hs + stride;
        hb       = reinterpret_cast<int*>(hl + stride);
        es       = reinterpret_cast<float*>(hb + stride);
        float*    sUni = es + p.W;
        uint32_t* sOff = reinterpret_cast<uint32_t*>(sUni + p.W);
        uint32_t* sEm  = sOff + p.W + 1;
        uint32_t* sTd  = sEm + p.nStates;
        for (uint32_t i = threadIdx.x; i < p.W; i += kThreads)
            sUni[i] = p.unigram[i];
        for (uint32_t i = threadIdx.x; i <= p.W; i += kThreads)
            sOff[i] = p.wordOff[i];
        for (uint32_t i = threadIdx.x; i < p.nStates; i += kThreads) {
            sEm[i] = p.stateEmis[i];
            sTd[i] = p.stateTdp[i];
        }
        wordOff   = sOff;
        stateEmis = sEm;
        stateTdp  = sTd;
        unigram   = sUni;
    }
    int4*          books  = p.books + 2 * f0;
    int4*          irrBooks = p.single ? p.irrBooks + 2 * f0 : nullptr;
    __shared__ int   sLast;                  // index of the newest book entry, -1 if none
    __shared__ float sLastScore, sLastLm;
    __shared__ int   sLastHad;               // Book::hadRegularWord of the newest entry
    __shared__ int   sIrr;                   // newest entry of the irregular book, -1 if none
    __shared__ float sIrrScore, sIrrLm;
    __shared__ float sTdp[64 * 4];           // transition models (when there are at most 64)
    if (p.nModels <= 64) {
        for (uint32_t i = threadIdx.x; i < p.nModels * 4; i += kThreads)
            sTdp[i] = p.tdp[i];
        tdp = sTdp;
    }

    // restart(): every hypothesis at FLT_MAX
    for (size_t i = threadIdx.x; i < stride; i += kThreads) {
        hs[i] = FLT_MAX;
        hl[i] = 0.0f;
        hb[i] = -1;
    }
    if (threadIdx.x == 0) {
        sLast      = -1;
        sLastScore = 0.0f;
        sLastLm    = 0.0f;
        sLastHad   = 0;
        sIrr       = -1;
        sIrrScore  = 0.0f;
        sIrrLm     = 0.0f;
    }
    __syncthreads();
    // Book::hadRegularWord of the entry a back pointer names (entries of the irregular book never have one, :368)
    auto had_regular = [&](int bkp) { return bkp >= 0 ? books[2 * bkp + 1].y : 0; };

    for (int t = 1; t <= T; ++t) {
        const float* sc   = p.scores + (size_t)(f0 + t - 1) * p.nEmis;
        const int    mainLast = sLast, mainHad = sLastHad, irrLast = sIrr;
        const float  mainScore = sLastScore, mainLm = sLastLm, irrScore = sIrrScore, irrLm = sIrrLm;
        for (uint32_t w = threadIdx.x; w < p.W; w += kThreads) {
            const uint32_t s0 = wordOff[w], S = wordOff[w + 1] - s0;
            float*         ws = hs + s0 + w;  // [0..S]
            float*         wl = hl + s0 + w;
            int*           wb = hb + s0 + w;
            // word start (:248-290): from the newest book entry, or from scratch; in single-word recognition a
            // regular word behind a regular word, and every irregular-chain entry, starts from the irregular book
            int   last      = mainLast;
            float lastScore = mainScore, lastLm = mainLm;
            if (p.single) {
                const uint32_t fl = p.flags[w];
                if ((mainLast >= 0 && mainHad && (fl & 1u)) || (fl & 2u)) {
                    last      = irrLast >= 0 ? irr_ref(irrLast) : -1;
                    lastScore = irrScore;
                    lastLm    = irrLm;
                }
            }
            float h0lm, h0s;
            if (last != -1) {
                h0lm = __fadd_rn(unigram[w], lastLm);
                h0s  = lastScore;
            }
            else {
                h0lm = unigram[w];
                h0s  = 0.0f;
            }
            h0s = __fadd_rn(h0s, h0lm);
            // states in descending order: the old values of sta, sta-1, sta-2 are still in place
            for (uint32_t sta = S; sta >= 1; --sta) {
                float bestS = FLT_MAX, bestL = 0.0f;
                int   bestB = -1;
                for (uint32_t pre = sta >= 2 ? sta - 2 : 0; pre <= sta; ++pre) {
                    float    ps, pl;
                    int      pb;
                    uint32_t model;
                    if (pre == 0) {
                        ps    = h0s;
                        pl    = h0lm;
                        pb    = last;
                        model = p.entryModel;
                    }
                    else {
                        ps    = ws[pre];
                        pl    = wl[pre];
                        pb    = wb[pre];
                        model = stateTdp[s0 + pre - 1];
                    }
                    const float sco = __fadd_rn(ps, tdp[model * 4 + (sta - pre)]);
                    if (sco < bestS) {
                        bestS = sco;
                        bestL = pl;
                        bestB = pb;
                    }
                }
                ws[sta] = __fadd_rn(bestS, __ldg(sc + stateEmis[s0 + sta - 1]));
                wl[sta] = bestL;
                wb[sta] = bestB;
            }
            // word end candidate (:400-404)
            es[w] = __fadd_rn(ws[S], tdp[stateTdp[s0 + S - 1] * 4 + 3]);
        }
        __syncthreads();
        // book keeping (:381-432): sequential scan over the words, replayed by warp 0
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            float     nbScore, nbLm;
            int       nbWord, nbBkp;
            // irregular = false: over all entries.  irregular = true (:390-395): over the entries that are not regular
            // words and whose history holds no regular word either
            auto scan = [&](bool irregular) {
                nbScore = FLT_MAX;
                nbLm    = 0.0f;
                nbWord  = -1;
                nbBkp   = -1;
                const uint32_t n = irregular ? p.nIrr : p.W;
                for (uint32_t base = 0; base < n; base += 32) {
                    const uint32_t i  = base + lane;
                    uint32_t       w  = i;
                    bool           ok = i < n;
                    if (irregular && ok) {
                        w                 = p.irrList[i];
                        const uint32_t s0 = wordOff[w], S = wordOff[w + 1] - s0;
                        ok                = !had_regular(hb[s0 + w + S]);
                    }
                    const float    cand = ok ? es[w] : FLT_MAX;
                    uint32_t       todo = 0xffffffffu;
                    while (true) {
                        const float    thr  = __fadd_rn(nbScore, nbLm);
                        const uint32_t hits = __ballot_sync(0xffffffffu, ok && cand < thr) & todo;
                        if (!hits)
                            break;
                        const int      first = __ffs(hits) - 1;  // lowest entry of the chunk that beats the current best
                        const uint32_t ww    = __shfl_sync(0xffffffffu, w, first);
                        const uint32_t s0 = wordOff[ww], S = wordOff[ww + 1] - s0;
                        const float    tmpScore = __shfl_sync(0xffffffffu, cand, first);
                        const float    lmw = hl[s0 + ww + S];
                        nbScore = __fsub_rn(tmpScore, lmw);
                        nbLm    = lmw;
                        nbBkp   = hb[s0 + ww + S];
                        nbWord  = (int)ww;
                        todo    = first == 31 ? 0u : (0xffffffffu << (first + 1));  // entries before it were already rejected
                    }
                }
            };
            scan(false);
            const float mScore = nbScore, mLm = nbLm;
            const int   mWord = nbWord, mBkp = nbBkp;
            if (p.single) {
                scan(true);
                if (lane == 0 && nbScore != FLT_MAX) {
                    const int b = sIrr + 1;
                    irrBooks[2 * b]     = make_int4(__float_as_int(nbScore), __float_as_int(nbLm), (int)p.entryWord[nbWord], nbBkp);
                    irrBooks[2 * b + 1] = make_int4(t, 0, 0, 0);
                    sIrr      = b;
                    sIrrScore = nbScore;
                    sIrrLm    = nbLm;
                }
            }
            if (lane == 0 && mScore != FLT_MAX) {
                const int b = sLast + 1;  // entries are only ever appended: the newest is the last
                int       had = 0, word = mWord;
                if (p.single) {
                    had  = (p.flags[mWord] & 1u) ? 1 : had_regular(mBkp);
                    word = (int)p.entryWord[mWord];
                }
                books[2 * b]     = make_int4(__float_as_int(mScore), __float_as_int(mLm), word, mBkp);
                books[2 * b + 1] = make_int4(t, had, 0, 0);
                sLast       = b;
                sLastScore  = mScore;
                sLastLm     = mLm;
                sLastHad    = had;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.nBooks[u] = sLast + 1;
        if (p.single)
            p.nIrrBooks[u] = sIrr + 1;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Register-resident variant.  The per-word kernel above is bound by the latency of its dependent shared-memory chains
// (13.8 us per frame for 1000 words).  Here a thread owns NPT CONSECUTIVE HMM states of the concatenated lexicon and
// keeps their hypotheses in registers across all frames; the per-state code is branch-free and the kernel is bound by
// shared-memory wavefronts, so every table is laid out to be read in one wavefront.
//  * feed() reads only the PREVIOUS frame's values of a state and its two predecessors, so all states update
//    independently; within a thread they are walked in descending order, in place.  The last two states of the left
//    neighbour arrive by warp shuffle, across warps through a double-buffered shared-memory slot per warp.
//  * The transition penalties a state needs -- its own loop, the forward of state-1 and the skip of state-2, with the
//    entry model standing in at the word start -- are three 5-bit indices into a table of the (at most 32) DISTINCT
//    penalty values of the lexicon: 32 words in 32 banks, conflict-free for any access pattern.  A word's first state
//    has skip-in = +inf: the candidate is formed and never wins, like the reference never forming it.
//  * A hypothesis carries (score, back pointer) only.  Its lmScore is a function of (word, back pointer): it was set
//    at the word start to unigram[w] (+ lmScore of the book entry it starts from, :271-290) and is copied unchanged
//    along the word, so the book keeping recomputes it with the same single rounding for the words it accepts.
//    (Holds for every hypothesis below FLT_MAX; needs |scores| < 1e30 so that nothing unreached ever gets below it.)
//  * The score row of the NEXT frame is copied to shared memory with cp.async while this frame is processed.
//  * Book keeping (:381-432) is a sequential scan "accept w if cand_w < nbScore + nbLm; nbScore = cand_w - lm_w" whose
//    threshold after accepting w is g_w = fl(fl(cand_w - lm_w) + lm_w), within 2^-24 (2|cand_w| + |lm_w|) of cand_w.
//    After its state update every warp reduces the words whose LAST state it owns (a contiguous range) with REDUX to
//    (min, first word holding it, second smallest); after the one barrier of the frame EVERY warp combines the <= 32
//    per-warp results and keeps the newest book entry in registers (thread 0 writes it out), so the next frame needs
//    no second barrier.  If the second smallest candidate exceeds the smallest M by more than 2^-21 (|M| + max|lm|),
//    every word accepted before the first argmin j leaves a threshold above M, so j is accepted, and nothing after j
//    beats g_j: the scan ends at j, and its result is taken directly.  Otherwise (ties, near ties) warp 0 replays the
//    scan, 32 words per ballot, and broadcasts the result.
//  * What is written after the barrier and read before the next one is double-buffered by frame parity (exchange
//    slots, per-warp results, the lm bound); the lmScore of the newest book entry comes from registers.
// ------------------------------------------------------------------------------------------------------------------
// meta: first | second << 1 | loop index << 2 | forward-in index << 7 | skip-in index << 12 | last << 17 | word << 18
constexpr uint32_t kFirst = 1u, kSecond = 2u, kLast = 1u << 17, kWordShift = 18, kMaxValues = 32;

struct SmemLayout {  // byte offsets into the dynamic shared memory of the register-resident kernel
    uint32_t xch, end, rows, uni, exit, bkLm, total;
    uint32_t flags, irrSlot, irrEnd, irrList, irrLm, bkHad;  // single-word recognition only
    uint32_t endBuffers;  // 2: the word ends are double-buffered by frame parity (any warp may read them after the barrier)
    SmemLayout(uint32_t nWarps, uint32_t W, uint32_t rowFloats, uint32_t maxT, bool single = false, uint32_t nIrr = 0,
               uint32_t endBuf = 1) {
        uint32_t o = 128;
        auto take = [&](uint32_t bytes) {
            const uint32_t at = o;
            o += (bytes + 15) & ~15u;
            return at;
        };
        xch   = take(2 * nWarps * 16);
        endBuffers = endBuf;
        end   = take(W * 8 * endBuf);
        rows  = take(2 * rowFloats * 4);
        uni   = take(W * 4);
        exit  = take(W * 4);
        bkLm  = take(maxT * 4);
        flags = irrSlot = irrEnd = irrList = irrLm = bkHad = 0;
        if (single) {
            flags   = take(W);
            irrSlot = take(W * 2);
            irrEnd  = take(2 * (nIrr + 1) * 8);
            irrList = take((nIrr + 1) * 4);
            irrLm   = take(maxT * 4);
            bkHad   = take(maxT);
        }
        total = o;
    }
    SmemLayout() = default;
};

struct SearchParams2 {
    SmemLayout      lay;
    const uint32_t* stMeta;    // [threads * NPT]
    const uint32_t* stEmOff;   // [threads * NPT / 2] byte offset of the state's emission in a score row, 16 bits each
    const float*    values;    // [32] distinct transition penalties
    const float*    unigram;   // [W]
    const float*    wordExit;  // [W] exit penalty of the word's last state
    const uint32_t* warpWords; // [nWarps + 1] words whose last state lives in warp i: [warpWords[i], warpWords[i + 1])
    float           maxAbsUni;
    uint32_t        W, maxT, rowFloats, nStates;  // rowFloats: nEmis rounded up to 4
    int             forceScan;           // test hook: 1 = never take the unique-minimum shortcut, 2 = and use the warp-0 replay
    const float*    scores;
    const int64_t*  frameOff;
    int             nEmis;
    int4*           books;
    int*            nBooks;
    // single-word recognition (the "words" are then the entries of the search, see SearchParams)
    const uint8_t*  flags;
    const uint32_t* entryWord;
    const uint32_t* irrList;
    uint32_t        nIrr;
    int4*           irrBooks;
    int*            nIrrBooks;
};

__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// order-preserving map of f32 onto u32 (and back): a < b  <=>  key(a) < key(b) for non-NaN values
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float(k ^ ((k & 0x80000000u) ? 0x80000000u : 0xffffffffu));
}

//  * SINGLE (single-word recognition, :248-258, 355-370, 390-395): every warp also tracks the newest entry of the
//    irregular book in registers; a word start picks its predecessor book by the entry's flags; the word ends of the
//    (few) irregular entries are mirrored into a slot array double-buffered by frame parity, over which every warp
//    replays the second, irregular-only book-keeping scan exactly after the frame's barrier.
template<int NPT, bool SINGLE>
__global__ void __launch_bounds__(kThreads, 1) linear_search_reg_kernel(const SearchParams2 p) {
    extern __shared__ __align__(128) unsigned char smemReg[];
    const uint32_t nThreads = blockDim.x, nWarps = nThreads / 32;
    // the carve-up is computed on the host (SmemLayout): one constant-bank operand per table address
    float*  sT    = reinterpret_cast<float*>(smemReg);                   // [32] penalty values, one per bank
    float4* xch   = reinterpret_cast<float4*>(smemReg + p.lay.xch);      // [2][nWarps] {s[n-2], s[n-1], b[n-2], b[n-1]} of lane 31
    float2* sEnd  = reinterpret_cast<float2*>(smemReg + p.lay.end);      // [W] word end {candidate score, bkp}
    float*  rows  = reinterpret_cast<float*>(smemReg + p.lay.rows);      // [2][rowFloats], 16-byte aligned
    float*  sUni  = reinterpret_cast<float*>(smemReg + p.lay.uni);       // [W]
    float*  sExit = reinterpret_cast<float*>(smemReg + p.lay.exit);      // [W]
    float*  sBkLm = reinterpret_cast<float*>(smemReg + p.lay.bkLm);      // [maxT] lmScore of the book entries
    uint8_t*  sFlags   = smemReg + p.lay.flags;                               // [W] bit 0 regular, bit 1 irregular chain
    uint16_t* sIrrSlot = reinterpret_cast<uint16_t*>(smemReg + p.lay.irrSlot);  // [W] position in irrList, 0xffff: regular
    float2*   sIrrEnd  = reinterpret_cast<float2*>(smemReg + p.lay.irrEnd);   // [2][nIrr + 1] word ends of the irregular entries
    uint32_t* sIrrList = reinterpret_cast<uint32_t*>(smemReg + p.lay.irrList);  // [nIrr]
    float*    sIrrLm   = reinterpret_cast<float*>(smemReg + p.lay.irrLm);     // [maxT] lmScore of the irregular book's entries
    uint8_t*  sBkHad   = smemReg + p.lay.bkHad;                               // [maxT] Book::hadRegularWord
    // per warp and frame parity: best word end of the warp's words {min key, second key, word, lmScore}, its bkp
    __shared__ uint4 wInfo[2][32];
    __shared__ int   wBkp[2][32];
    __shared__ float sMaxLm[2];   // bound on |lmScore| of the book entries so far, by frame parity (written by thread 0)
    __shared__ float sReplay[4];  // result of the sequential replay: score, lmScore, word, bkp

    const int     u  = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t f0 = p.frameOff[u];
    const int     T  = (int)(p.frameOff[u + 1] - f0);
    const bool    vec16 = (p.nEmis & 3) == 0 && (reinterpret_cast<uintptr_t>(p.scores) & 15) == 0;
    auto prefetch_row = [&](int t) {  // row of frame t (1-based) -> rows[t & 1]
        const float* src = p.scores + (size_t)(f0 + t - 1) * p.nEmis;
        float*       dst = rows + (t & 1) * p.rowFloats;
        if (vec16)
            for (uint32_t i = tid * 4; i < (uint32_t)p.nEmis; i += nThreads * 4)
                cp_async16(dst + i, src + i);
        else
            for (uint32_t i = tid; i < (uint32_t)p.nEmis; i += nThreads)
                cp_async4(dst + i, src + i);
    };
    if (T > 0)
        prefetch_row(1);
    if (tid < 32)
        sT[tid] = p.values[tid];
    for (uint32_t i = tid; i < p.W; i += nThreads) {
        sUni[i]  = p.unigram[i];
        sExit[i] = p.wordExit[i];
        if (SINGLE) {
            sFlags[i]   = p.flags[i];
            sIrrSlot[i] = 0xffffu;
        }
    }
    if (SINGLE) {
        __syncthreads();
        for (uint32_t i = tid; i < p.nIrr; i += nThreads) {
            const uint32_t w = p.irrList[i];
            sIrrList[i] = w;
            sIrrSlot[w] = (uint16_t)i;
        }
    }
    // this thread's states; the words whose LAST state lives in this warp are a contiguous range
    const uint32_t i0 = (uint32_t)tid * NPT;
    const uint32_t wordA = p.warpWords[warp], wordB = p.warpWords[warp + 1];
    float    hs[NPT];
    int      hb[NPT];
    uint32_t meta[NPT], emOff[NPT / 2];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        hs[k]   = FLT_MAX;  // restart()
        hb[k]   = -1;
        meta[k] = p.stMeta[i0 + k];
    }
#pragma unroll
    for (int k = 0; k < NPT / 2; ++k)
        emOff[k] = p.stEmOff[i0 / 2 + k];
    if (lane == 31) {
        xch[warp]          = make_float4(FLT_MAX, FLT_MAX, __int_as_float(-1), __int_as_float(-1));
        xch[nWarps + warp] = make_float4(FLT_MAX, FLT_MAX, __int_as_float(-1), __int_as_float(-1));
    }
    if (tid == 0)
        sMaxLm[0] = sMaxLm[1] = 0.0f;
    // the newest book entry: every warp tracks it in registers (the book keeping below is evaluated by all warps)
    int   last      = -1;
    float lastScore = 0.0f, lastLm = 0.0f;  // 0 before the first entry
    int   lastHad   = 0;                    // Book::hadRegularWord of the newest entry
    int   irrLast   = -1;                   // the newest entry of the irregular book
    float irrScore = 0.0f, irrLm = 0.0f;
    cp_async_wait_all();
    __syncthreads();
    const unsigned char* sTb = reinterpret_cast<const unsigned char*>(sT);
    // lmScore of a hypothesis that came from book entry bk (the newest entry is not read from shared memory: it was
    // written after the last barrier); bk <= -2 names entry -2 - bk of the irregular book
    auto lm_of = [&](uint32_t w, int bk) {
        const float un = sUni[w];
        if (SINGLE && bk <= -2)
            return __fadd_rn(un, -2 - bk == irrLast ? irrLm : sIrrLm[-2 - bk]);
        return bk >= 0 ? __fadd_rn(un, bk == last ? lastLm : sBkLm[bk]) : un;
    };
    auto had_of = [&](int bk) -> int { return bk >= 0 ? (bk == last ? lastHad : (int)sBkHad[bk]) : 0; };

    for (int t = 1; t <= T; ++t) {
        const int par = t & 1;  // score row and boundary states of this frame are in buffer par
        if (t < T)
            prefetch_row(t + 1);
        const unsigned char* row = reinterpret_cast<const unsigned char*>(rows + par * p.rowFloats);
        float2* sEndT = sEnd + (p.lay.endBuffers == 2 ? par * p.W : 0u);  // this frame's word ends
        // single-word recognition, uniform over the frame: the flag bits that send a word start to the irregular book
        // (bit 1: irregular-chain entry; bit 0: regular word, once the main book's newest entry holds a regular word)
        const uint32_t irrSelect = 2u | ((last >= 0 && lastHad) ? 1u : 0u);
        const int      irrRef    = irrLast >= 0 ? -2 - irrLast : -1;
        {
            // previous-frame values of the two states left of my block
            float pS2 = __shfl_up_sync(0xffffffffu, hs[NPT - 2], 1), pS1 = __shfl_up_sync(0xffffffffu, hs[NPT - 1], 1);
            int   pB2 = __shfl_up_sync(0xffffffffu, hb[NPT - 2], 1), pB1 = __shfl_up_sync(0xffffffffu, hb[NPT - 1], 1);
            if (lane == 0) {
                const float4 nb = xch[par * nWarps + (warp > 0 ? warp - 1 : 0)];
                pS2 = warp > 0 ? nb.x : FLT_MAX;
                pS1 = warp > 0 ? nb.y : FLT_MAX;
                pB2 = warp > 0 ? __float_as_int(nb.z) : -1;
                pB1 = warp > 0 ? __float_as_int(nb.w) : -1;
            }
#pragma unroll
            for (int k = NPT - 1; k >= 0; --k) {
                const uint32_t m = meta[k];
                const uint32_t w = m >> kWordShift;
                const float    tLoop = *reinterpret_cast<const float*>(sTb + (m & 0x7cu));
                const float    tFwd  = *reinterpret_cast<const float*>(sTb + ((m >> 5) & 0x7cu));
                const float    tSkip = *reinterpret_cast<const float*>(sTb + ((m >> 10) & 0x7cu));
                // word start (:248-290): from the newest book entry, or from scratch; in single-word recognition a
                // regular word behind a regular word, and every irregular-chain entry, starts from the irregular book
                const float uni = sUni[w];
                int         from = last;
                float       fromScore = lastScore, fromLm = lastLm;
                // byte offset of the state's emission in the score row; its two low bits carry the entry's flags
                const uint32_t eo = (k & 1) ? emOff[k / 2] >> 16 : emOff[k / 2] & 0xffffu;
                if (SINGLE) {
                    const bool useIrr = (eo & irrSelect) != 0;  // chain entry, or regular word behind a regular word
                    from      = useIrr ? irrRef : last;
                    fromScore = useIrr ? irrScore : lastScore;
                    fromLm    = useIrr ? irrLm : lastLm;
                }
                const float h0lm = from != -1 ? __fadd_rn(uni, fromLm) : uni;
                const float h0s  = __fadd_rn(fromScore, h0lm);
                const bool  first = m & kFirst, second = m & kSecond;
                // predecessors in the reference's order pre = sta-2, sta-1, sta (the first strictly smaller one wins)
                float s1 = k >= 1 ? hs[k >= 1 ? k - 1 : 0] : pS1;
                int   b1 = k >= 1 ? hb[k >= 1 ? k - 1 : 0] : pB1;
                float s2 = k >= 2 ? hs[k >= 2 ? k - 2 : 0] : (k == 1 ? pS1 : pS2);
                int   b2 = k >= 2 ? hb[k >= 2 ? k - 2 : 0] : (k == 1 ? pB1 : pB2);
                s1 = first ? h0s : s1;
                b1 = first ? from : b1;
                s2 = second ? h0s : s2;
                b2 = second ? from : b2;
                const float c2 = __fadd_rn(s2, tSkip), c1 = __fadd_rn(s1, tFwd), c0 = __fadd_rn(hs[k], tLoop);
                const bool  t2 = c2 < FLT_MAX;
                float       bestS = t2 ? c2 : FLT_MAX;
                int         bestB = t2 ? b2 : -1;
                const bool  t1 = c1 < bestS;
                bestS = t1 ? c1 : bestS;
                bestB = t1 ? b1 : bestB;
                const bool t0 = c0 < bestS;
                bestS = t0 ? c0 : bestS;
                bestB = t0 ? hb[k] : bestB;
                const float e = *reinterpret_cast<const float*>(row + (SINGLE ? eo & 0xfffcu : eo));
                hs[k] = __fadd_rn(bestS, e);
                hb[k] = bestB;
                if (m & kLast) {  // word end candidate (:400-404)
                    const float2 en = make_float2(__fadd_rn(hs[k], sExit[w]), __int_as_float(bestB));
                    sEndT[w] = en;
                    if (SINGLE && !(eo & 1u))  // an irregular entry: mirror its word end into its slot
                        sIrrEnd[par * (p.nIrr + 1) + sIrrSlot[w]] = en;
                }
            }
            if (lane == 31)  // publish my last two states for the next warp's next frame
                xch[(par ^ 1) * nWarps + warp] =
                        make_float4(hs[NPT - 2], hs[NPT - 1], __int_as_float(hb[NPT - 2]), __int_as_float(hb[NPT - 1]));
        }
        // best word end among this warp's words: smallest candidate, the first word holding it, the second smallest
        __syncwarp();
        {
            uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu, bw = 0;
            int      bbk = -1;
#pragma unroll 1
            for (uint32_t c = wordA; c < wordB; c += 32) {
                const uint32_t w  = c + lane;
                float2         en = make_float2(FLT_MAX, __int_as_float(-1));
                if (w < wordB)
                    en = sEndT[w];
                const uint32_t key = w < wordB ? float_key(en.x) : 0xffffffffu;
                const uint32_t m1  = __reduce_min_sync(0xffffffffu, key);
                const int      i1  = __ffs(__ballot_sync(0xffffffffu, key == m1)) - 1;
                const uint32_t m2  = __reduce_min_sync(0xffffffffu, lane == i1 ? 0xffffffffu : key);
                const int      bk1 = __shfl_sync(0xffffffffu, __float_as_int(en.y), i1);
                if (m1 < k1) {  // strictly smaller: later words only replace on a real improvement
                    k2  = min(k1, m2);
                    k1  = m1;
                    bw  = c + i1;
                    bbk = bk1;
                }
                else
                    k2 = min(k2, m1);
            }
            if (lane == 0) {
                wInfo[par][warp] = make_uint4(k1, k2, bw, __float_as_uint(wordA < wordB ? lm_of(bw, bbk) : 0.0f));
                wBkp[par][warp]  = bbk;
            }
        }
        cp_async_wait_all();  // next frame's score row has landed (made visible by the barrier)
        __syncthreads();      // the only barrier of a frame
        // book keeping (:381-432), evaluated by every warp: combine the warps' best word ends
        {
            float nbScore = FLT_MAX, nbLm = 0.0f;
            int   nbWord = -1, nbBkp = -1;
            uint4 best = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
            int   bestBk = -1;
            if ((uint32_t)lane < nWarps) {
                best   = wInfo[par][lane];
                bestBk = wBkp[par][lane];
            }
            const uint32_t kM  = __reduce_min_sync(0xffffffffu, best.x);
            const int      lj  = __ffs(__ballot_sync(0xffffffffu, best.x == kM)) - 1;
            const uint32_t kMo = __reduce_min_sync(0xffffffffu, lane == lj ? 0xffffffffu : best.x);  // best of the other warps
            const uint32_t kM2 = min(kMo, __shfl_sync(0xffffffffu, best.y, lj));
            const float    M = key_float(kM), M2 = key_float(kM2), Mo = key_float(kMo);
            const float    slack = __fmul_rn(__fadd_rn(fabsf(M), __fadd_rn(p.maxAbsUni, sMaxLm[par])), 4.76837158203125e-07f);
            // the words of warps [first, ...) selected by warpMask, exactly as the reference scans them
            uint32_t warpMask = 0u;
            if (!p.forceScan && M2 > __fadd_rn(M, slack)) {
                if (M < FLT_MAX) {  // the scan accepts the unique minimum last
                    nbWord  = (int)__shfl_sync(0xffffffffu, best.z, lj);
                    nbLm    = __uint_as_float(__shfl_sync(0xffffffffu, best.w, lj));
                    nbBkp   = __shfl_sync(0xffffffffu, bestBk, lj);
                    nbScore = __fsub_rn(M, nbLm);
                }
            }
            else if (p.lay.endBuffers == 2 && p.forceScan < 2) {
                // Ties / near ties.  If every word OUTSIDE warp lj's range of n words is above M + K slack, K = n / 4 + 2,
                // the scan over that range alone, started from scratch, ends in the reference's state.  The reference
                // enters the range with a threshold t_in > M + K slack - e (only far words were accepted; e <= slack / 4
                // bounds the rounding of g_w above).  While the two scans disagree, a word accepted by one of them only
                // was rejected by the lower threshold, so the lower of the two thresholds falls by at most e per word
                // and stays above M + e over the n words; a word accepted by both makes them identical from there on,
                // and the first word holding M is such a word.  Behind the range the threshold is at most M + e and no
                // far word beats it.  (Typical: an irregular word and its irregular-chain copy with equal scores.)
                const float nLj = (float)(p.warpWords[lj + 1] - p.warpWords[lj]);
                const float far = __fmul_rn(slack, __fadd_rn(__fmul_rn(nLj, 0.25f), 2.0f));
                warpMask = (p.forceScan < 1 && Mo > __fadd_rn(M, far)) ? (1u << lj) : 0xffffffffu;
            }
            if (warpMask) {
                // ties / near ties: the sequential scan, exactly, by EVERY warp and without a second barrier (the word
                // ends are double-buffered).  A warp's words can only be accepted if its smallest candidate beats the
                // running threshold (cand >= min >= thr rejects them all and leaves the threshold as it is), so the
                // scan visits only the "record-breaking" warps in order -- a handful -- and replays their words.
                const float myMin = (uint32_t)lane < nWarps ? key_float(best.x) : FLT_MAX;  // warps without words: NaN
                uint32_t    nextWarps = warpMask;
                while (true) {
                    const float    thr0  = __fadd_rn(nbScore, nbLm);
                    const uint32_t hitsW = __ballot_sync(0xffffffffu, myMin < thr0) & nextWarps;
                    if (!hitsW)
                        break;
                    const int wi = __ffs(hitsW) - 1;
                    nextWarps    = wi == 31 ? 0u : (0xffffffffu << (wi + 1));
                    const uint32_t a = p.warpWords[wi], b = p.warpWords[wi + 1];
#pragma unroll 1
                    for (uint32_t base = a; base < b; base += 32) {
                        const uint32_t w    = base + lane;
                        float          cand = FLT_MAX, lmw = 0.0f;
                        int            bk   = -1;
                        if (w < b) {
                            const float2 en = sEndT[w];
                            bk   = __float_as_int(en.y);
                            cand = en.x;
                            lmw  = lm_of(w, bk);
                        }
                        uint32_t todo = 0xffffffffu;
                        while (true) {
                            const float    thr  = __fadd_rn(nbScore, nbLm);
                            const uint32_t hits = __ballot_sync(0xffffffffu, cand < thr) & todo;
                            if (!hits)
                                break;
                            const int   first    = __ffs(hits) - 1;
                            const float tmpScore = __shfl_sync(0xffffffffu, cand, first);
                            nbLm    = __shfl_sync(0xffffffffu, lmw, first);
                            nbBkp   = __shfl_sync(0xffffffffu, bk, first);
                            nbScore = __fsub_rn(tmpScore, nbLm);
                            nbWord  = (int)(base + first);
                            todo    = first == 31 ? 0u : (0xffffffffu << (first + 1));
                        }
                    }
                }
            }
            else if (p.forceScan || !(M2 > __fadd_rn(M, slack))) {
                // (lexicons whose word ends do not fit twice) warp 0 replays the sequential scan over the words, 32 at a time
                if (warp == 0) {
#pragma unroll 1
                    for (uint32_t base = 0; base < p.W; base += 32) {
                        const uint32_t w    = base + lane;
                        float          cand = FLT_MAX, lmw = 0.0f;
                        int            bk   = -1;
                        if (w < p.W) {
                            const float2 en = sEndT[w];
                            bk   = __float_as_int(en.y);
                            cand = en.x;
                            lmw  = lm_of(w, bk);
                        }
                        uint32_t todo = 0xffffffffu;
                        while (true) {
                            const float    thr  = __fadd_rn(nbScore, nbLm);
                            const uint32_t hits = __ballot_sync(0xffffffffu, cand < thr) & todo;
                            if (!hits)
                                break;
                            const int   first    = __ffs(hits) - 1;  // lowest word of the chunk that beats the current best
                            const float tmpScore = __shfl_sync(0xffffffffu, cand, first);
                            nbLm    = __shfl_sync(0xffffffffu, lmw, first);
                            nbBkp   = __shfl_sync(0xffffffffu, bk, first);
                            nbScore = __fsub_rn(tmpScore, nbLm);
                            nbWord  = (int)(base + first);
                            todo    = first == 31 ? 0u : (0xffffffffu << (first + 1));  // words before it were rejected
                        }
                    }
                    if (lane == 0) {
                        sReplay[0] = nbScore;
                        sReplay[1] = nbLm;
                        sReplay[2] = __int_as_float(nbWord);
                        sReplay[3] = __int_as_float(nbBkp);
                    }
                }
                __syncthreads();  // the condition is the same in every warp
                nbScore = sReplay[0];
                nbLm    = sReplay[1];
                nbWord  = __float_as_int(sReplay[2]);
                nbBkp   = __float_as_int(sReplay[3]);
            }
            // the second book (:355-370): the exact scan over the irregular entries whose history holds no regular
            // word, replayed by every warp (they all keep the newest entries in registers); evaluated on the state
            // BEFORE the main book's new entry is appended, like the reference (both scans see this frame's hypotheses)
            float ibScore = FLT_MAX, ibLm = 0.0f;
            int   ibWord = -1, ibBkp = -1;
            if (SINGLE) {
#pragma unroll 1
                for (uint32_t base = 0; base < p.nIrr; base += 32) {
                    const uint32_t i    = base + lane;
                    float          cand = FLT_MAX, lmw = 0.0f;
                    int            bk   = -1;
                    uint32_t       w    = 0;
                    if (i < p.nIrr) {
                        const float2 en = sIrrEnd[par * (p.nIrr + 1) + i];
                        w  = sIrrList[i];
                        bk = __float_as_int(en.y);
                        if (!had_of(bk)) {
                            cand = en.x;
                            lmw  = lm_of(w, bk);
                        }
                    }
                    uint32_t todo = 0xffffffffu;
                    while (true) {
                        const float    thr  = __fadd_rn(ibScore, ibLm);
                        const uint32_t hits = __ballot_sync(0xffffffffu, cand < thr) & todo;
                        if (!hits)
                            break;
                        const int   first    = __ffs(hits) - 1;
                        const float tmpScore = __shfl_sync(0xffffffffu, cand, first);
                        ibLm    = __shfl_sync(0xffffffffu, lmw, first);
                        ibBkp   = __shfl_sync(0xffffffffu, bk, first);
                        ibWord  = (int)__shfl_sync(0xffffffffu, w, first);
                        ibScore = __fsub_rn(tmpScore, ibLm);
                        todo    = first == 31 ? 0u : (0xffffffffu << (first + 1));
                    }
                }
            }
            if (nbScore != FLT_MAX) {
                int had = 0;
                if (SINGLE)
                    had = (sFlags[nbWord] & 1u) ? 1 : had_of(nbBkp);  // before `last` moves on
                ++last;  // entries are only ever appended: the newest is the last
                if (tid == 0) {
                    int4* books = p.books + 2 * (f0 + last);
                    books[0] = make_int4(__float_as_int(nbScore), __float_as_int(nbLm),
                                         SINGLE ? (int)p.entryWord[nbWord] : nbWord, nbBkp);
                    books[1] = make_int4(t, had, 0, 0);
                    sBkLm[last] = nbLm;
                    if (SINGLE)
                        sBkHad[last] = (uint8_t)had;
                }
                lastScore = nbScore;
                lastLm    = nbLm;
                lastHad   = had;
            }
            if (SINGLE && ibScore != FLT_MAX) {
                ++irrLast;
                if (tid == 0) {
                    int4* books = p.irrBooks + 2 * (f0 + irrLast);
                    books[0] = make_int4(__float_as_int(ibScore), __float_as_int(ibLm), (int)p.entryWord[ibWord], ibBkp);
                    books[1] = make_int4(t, 0, 0, 0);
                    sIrrLm[irrLast] = ibLm;
                }
                irrScore = ibScore;
                irrLm    = ibLm;
            }
            if (tid == 0) {  // read by every warp after the next barrier
                float m = nbScore != FLT_MAX ? fmaxf(sMaxLm[par], fabsf(nbLm)) : sMaxLm[par];
                if (SINGLE && ibScore != FLT_MAX)
                    m = fmaxf(m, fabsf(ibLm));
                sMaxLm[par ^ 1] = m;
            }
        }
    }
    if (tid == 0) {
        p.nBooks[u] = last + 1;
        if (SINGLE)
            p.nIrrBooks[u] = irrLast + 1;
    }
}

}  // namespace

struct rb_search {
    rb::DeviceInfo dev;
    uint32_t       W = 0, nStates = 0, nModels = 0, entryModel = 0;
    cudaStream_t   stream = nullptr;
    rb::DevBuf<uint32_t> dWordOff, dStateEmis, dStateTdp, dStateMeta, dStateEmOff, dWarpWords;
    rb::DevBuf<float>    dValues, dWordExit;
    float                maxAbsUni = 0.0f;
    int                  npt = 0, regThreads = 0;  // states per thread / threads of the register-resident kernel, 0: per-word kernel
    uint32_t             maxEmis = 0;
    rb::DevBuf<float>    dTdp, dUnigram, dHypScore, dHypLm, dEndScore, dScores;
    rb::DevBuf<int>      dHypBkp, dBooks, dNBooks;  // dBooks: 8 words per frame (two int4 per book entry)
    rb::PinnedBuf<int>   hostBooks;  // page-locked: the copy back runs at the PCIe rate and does not stage
    // single-word recognition: W / nStates count the ENTRIES of the search (irregular words twice)
    bool                  single = false;
    uint32_t              nIrr   = 0;
    rb::DevBuf<uint8_t>   dFlags;
    rb::DevBuf<uint32_t>  dEntryWord, dIrrList;
    rb::DevBuf<int>       dIrrBooks, dNIrrBooks;
    rb::PinnedBuf<int>    hostIrrBooks;
    std::vector<int>      nIrrBooks;
    rb::DevBuf<int64_t>  dFrameOff;
    // results of the last decode, on the host
    std::vector<int64_t> frameOff;
    std::vector<int>     nBooks;
    // the record a back pointer names (>= 0: main book, <= -2: entry -2 - bkp of the irregular book) in segment f0
    const int* record(int64_t f0, int ref) const {
        return ref >= 0 ? hostBooks.p + (f0 + ref) * 8 : hostIrrBooks.p + (f0 + (-2 - ref)) * 8;
    }
    ~rb_search() {
        if (stream)
            cudaStreamDestroy(stream);
    }
};

extern "C" int rb_search_create(const rb_lexicon* lx, int device, rb_search** out) {
    RB_REQUIRE(lx && out, "NULL argument");
    *out = nullptr;
    RB_REQUIRE(lx->n_words >= 1 && lx->word_offsets && lx->state_emission && lx->state_tdp_model && lx->tdp && lx->unigram,
               "lexicon has NULL tables or no words");
    RB_REQUIRE(lx->n_models >= 1 && lx->entry_model < lx->n_models, "entry model %u of %u", lx->entry_model, lx->n_models);
    const uint32_t nStates = lx->word_offsets[lx->n_words];
    for (uint32_t w = 0; w < lx->n_words; ++w)
        RB_REQUIRE(lx->word_offsets[w + 1] > lx->word_offsets[w], "word %u has no HMM state", w);
    for (uint32_t s = 0; s < nStates; ++s)
        RB_REQUIRE(lx->state_tdp_model[s] < lx->n_models, "state %u refers to transition model %u >= %u", s,
                   lx->state_tdp_model[s], lx->n_models);
    rb_search* h = new (std::nothrow) rb_search();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    // single-word recognition: the search runs over ENTRIES -- every pronunciation, an irregular one followed by its
    // irregular-chain copy (addPronunciations, :472-487); the tables below are those of the entries
    std::vector<uint32_t> eOff, eEmis, eTdp, eWord, irrList;
    std::vector<float>    eUni;
    std::vector<uint8_t>  eFlags;
    rb_lexicon            entries = *lx;
    if (lx->single_word) {
        eOff.push_back(0);
        for (uint32_t w = 0; w < lx->n_words; ++w) {
            const bool regular = !lx->word_regular || lx->word_regular[w];
            for (int copy = 0; copy < (regular ? 1 : 2); ++copy) {
                if (!regular)
                    irrList.push_back((uint32_t)eWord.size());
                eWord.push_back(w);
                eFlags.push_back((uint8_t)((regular ? 1 : 0) | (copy ? 2 : 0)));
                eUni.push_back(lx->unigram[w]);
                eEmis.insert(eEmis.end(), lx->state_emission + lx->word_offsets[w], lx->state_emission + lx->word_offsets[w + 1]);
                eTdp.insert(eTdp.end(), lx->state_tdp_model + lx->word_offsets[w], lx->state_tdp_model + lx->word_offsets[w + 1]);
                eOff.push_back((uint32_t)eEmis.size());
            }
        }
        entries.n_words         = (uint32_t)eWord.size();
        entries.word_offsets    = eOff.data();
        entries.state_emission  = eEmis.data();
        entries.state_tdp_model = eTdp.data();
        entries.unigram         = eUni.data();
        lx                      = &entries;
        h->single               = true;
        h->nIrr                 = (uint32_t)irrList.size();
        if (irrList.empty())
            irrList.push_back(0);  // never read (nIrr = 0); keeps the table uploads uniform
    }
    const uint32_t nStatesAll = lx->word_offsets[lx->n_words];  // of the entries, from here on
    auto fail = [&](int code) {
        delete h;
        return code;
    };
    int rc = rb::use_device(device, &h->dev);
    if (rc != RB_OK)
        return fail(rc);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rb::set_error("cudaStreamCreate failed");
        return fail(RB_ERR_CUDA);
    }
    h->W          = lx->n_words;
    h->nStates    = nStatesAll;
    h->nModels    = lx->n_models;
    h->entryModel = lx->entry_model;
    // register-resident kernel: per-state descriptor and the table of distinct transition penalties
    {
        const float           inf = std::numeric_limits<float>::infinity();
        std::vector<float>    values;
        bool                  fits = lx->n_words <= (1u << (32 - kWordShift)) && h->nIrr <= 1024;
        auto valueOf = [&](float v) -> uint32_t {
            for (size_t i = 0; i < values.size(); ++i)
                if (memcmp(&values[i], &v, 4) == 0)
                    return (uint32_t)i;
            values.push_back(v);
            if (values.size() > kMaxValues)
                fits = false;
            return (uint32_t)values.size() - 1;
        };
        const uint32_t infIdx  = valueOf(inf);
        uint32_t       maxEmis = 0;
        for (uint32_t s = 0; s < nStatesAll; ++s)
            maxEmis = std::max(maxEmis, lx->state_emission[s]);
        h->maxEmis = maxEmis;  // kept for every lexicon: decode checks it against the width of the score rows
        fits = fits && maxEmis < 16384 && nStatesAll <= (uint32_t)kThreads * 16;
        const float*          tdp = lx->tdp;
        const uint32_t        em  = lx->entry_model;
        const uint32_t*       mo  = lx->state_tdp_model;
        std::vector<uint32_t> meta((size_t)kThreads * 16, infIdx << 2 | infIdx << 7 | infIdx << 12);  // stays at FLT_MAX
        std::vector<uint16_t> emOff((size_t)kThreads * 16, 0);
        std::vector<float>    wordExit(lx->n_words);
        for (uint32_t w = 0; w < lx->n_words && fits; ++w) {
            for (uint32_t i = lx->word_offsets[w]; i < lx->word_offsets[w + 1] && fits; ++i) {
                const uint32_t j     = i - lx->word_offsets[w];
                const uint32_t oLoop = valueOf(tdp[mo[i] * 4]);
                const uint32_t oFwd  = valueOf(j == 0 ? tdp[em * 4 + 1] : tdp[mo[i - 1] * 4 + 1]);
                const uint32_t oSkip = j == 0 ? infIdx : valueOf(j == 1 ? tdp[em * 4 + 2] : tdp[mo[i - 2] * 4 + 2]);
                meta[i]  = (j == 0 ? kFirst : 0u) | (j == 1 ? kSecond : 0u) | oLoop << 2 | oFwd << 7 | oSkip << 12 |
                           (i + 1 == lx->word_offsets[w + 1] ? kLast : 0u) | (w << kWordShift);
                emOff[i] = (uint16_t)(lx->state_emission[i] * 4 | (h->single ? eFlags[w] : 0));
            }
            wordExit[w]  = tdp[mo[lx->word_offsets[w + 1] - 1] * 4 + 3];
            h->maxAbsUni = std::max(h->maxAbsUni, std::fabs(lx->unigram[w]));
        }
        if (fits && getenv("RB_SEARCH_PER_WORD") == nullptr) {
            // few states per thread while that still fills the SM's four schedulers with several warps each
            h->npt = nStatesAll <= 512 * 2 ? 2 : (nStatesAll <= (uint32_t)kThreads * 4 ? 4 : (nStatesAll <= (uint32_t)kThreads * 8 ? 8 : 16));
            if (const char* e = getenv("RB_SEARCH_NPT"))
                if (atoi(e) * (uint32_t)kThreads >= nStatesAll && (atoi(e) == 2 || atoi(e) == 4 || atoi(e) == 8 || atoi(e) == 16))
                    h->npt = atoi(e);
            h->regThreads = (int)(((nStatesAll + h->npt - 1) / h->npt + 31) / 32 * 32);
            // words whose last state lives in warp i (contiguous: states are in lexicon order)
            std::vector<uint32_t> warpWords(h->regThreads / 32 + 1, lx->n_words);
            warpWords[0] = 0;
            for (size_t i = 1; i < warpWords.size(); ++i) {  // first word whose last state is at or behind warp i
                uint32_t lo = 0, hi = lx->n_words;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) / 2;
                    if ((lx->word_offsets[mid + 1] - 1) / (32u * h->npt) < i)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                warpWords[i] = lo;
            }
            if (h->dWarpWords.upload(warpWords.data(), warpWords.size(), h->stream) != RB_OK) {
                rb::set_error("lexicon upload failed");
                return fail(RB_ERR_CUDA);
            }
            values.resize(kMaxValues, inf);
            if (h->dStateMeta.upload(meta.data(), meta.size(), h->stream) != RB_OK ||
                h->dStateEmOff.upload(reinterpret_cast<const uint32_t*>(emOff.data()), emOff.size() / 2, h->stream) != RB_OK ||
                h->dValues.upload(values.data(), values.size(), h->stream) != RB_OK ||
                h->dWordExit.upload(wordExit.data(), wordExit.size(), h->stream) != RB_OK) {
                rb::set_error("lexicon upload failed");
                return fail(RB_ERR_CUDA);
            }
        }
    }
    if (h->single && (h->dFlags.upload(eFlags.data(), eFlags.size(), h->stream) != RB_OK ||
                      h->dEntryWord.upload(eWord.data(), eWord.size(), h->stream) != RB_OK ||
                      h->dIrrList.upload(irrList.data(), irrList.size(), h->stream) != RB_OK)) {
        rb::set_error("lexicon upload failed");
        return fail(RB_ERR_CUDA);
    }
    if (h->dWordOff.upload(lx->word_offsets, lx->n_words + 1, h->stream) != RB_OK ||
        h->dStateEmis.upload(lx->state_emission, nStatesAll, h->stream) != RB_OK ||
        h->dStateTdp.upload(lx->state_tdp_model, nStatesAll, h->stream) != RB_OK ||
        h->dTdp.upload(lx->tdp, (size_t)lx->n_models * 4, h->stream) != RB_OK ||
        h->dUnigram.upload(lx->unigram, lx->n_words, h->stream) != RB_OK ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
        rb::set_error("lexicon upload failed");
        return fail(RB_ERR_CUDA);
    }
    *out = h;
    return RB_OK;
}

extern "C" void rb_search_destroy(rb_search* h) {
    if (!h)
        return;
    cudaSetDevice(h->dev.ordinal);
    delete h;
}

// d_scores [frames x n_emissions] on the device; results are kept in the handle (rb_search_traceback)
extern "C" int rb_search_decode_dev(rb_search* h, const float* d_scores, int n_emissions, const int64_t* frame_offsets,
                                    int n_utt, void* stream) {
    RB_REQUIRE(h && frame_offsets && n_utt >= 0 && n_emissions >= 1, "bad argument");
    // require(emissionScores->nEmissions() >= acousticModel_->nEmissions()), src/Search/LinearSearch.cc:235
    RB_REQUIRE(h->nStates == 0 || h->maxEmis < (uint32_t)n_emissions,
               "the lexicon refers to emission %u, the score rows have %d emissions", h->maxEmis, n_emissions);
    h->frameOff.assign(frame_offsets, frame_offsets + n_utt + 1);
    const int64_t base = frame_offsets[0], T = frame_offsets[n_utt] - base;
    for (auto& f : h->frameOff)
        f -= base;
    h->nBooks.assign(n_utt, 0);
    if (n_utt == 0 || T <= 0)
        return RB_OK;
    RB_REQUIRE(d_scores != nullptr, "NULL score buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    cudaStream_t s      = stream ? (cudaStream_t)stream : h->stream;
    const size_t stride = (size_t)h->nStates + h->W;
    RB_CHECK(h->dBooks.reserve((size_t)T * 8));
    RB_CHECK(h->dNBooks.reserve((size_t)n_utt));
    RB_CHECK(h->dFrameOff.reserve((size_t)n_utt + 1));
    RB_CUDA(cudaMemcpyAsync(h->dFrameOff.p, h->frameOff.data(), sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, s));
    int64_t maxT = 0;
    for (int u = 0; u < n_utt; ++u)
        maxT = std::max(maxT, h->frameOff[u + 1] - h->frameOff[u]);
    const uint32_t rowFloats = ((uint32_t)n_emissions + 3) & ~3u;
    SmemLayout lay(h->regThreads / 32, h->W, rowFloats, (uint32_t)maxT, h->single, h->nIrr, 2);
    if (lay.total + 4096 > h->dev.smem_optin)
        lay = SmemLayout(h->regThreads / 32, h->W, rowFloats, (uint32_t)maxT, h->single, h->nIrr, 1);
    const size_t smem2 = lay.total;
    if (h->npt && smem2 + 4096 <= h->dev.smem_optin) {
        SearchParams2 q;
        q.lay       = lay;
        q.warpWords = h->dWarpWords.p;
        q.nStates   = h->nStates;
        q.stMeta    = h->dStateMeta.p;
        q.stEmOff   = h->dStateEmOff.p;
        q.values    = h->dValues.p;
        q.unigram   = h->dUnigram.p;
        q.wordExit  = h->dWordExit.p;
        q.maxAbsUni = h->maxAbsUni;
        q.W         = h->W;
        q.maxT      = (uint32_t)maxT;
        q.rowFloats = rowFloats;
        q.forceScan = getenv("RB_SEARCH_FORCE_SCAN") ? std::max(1, atoi(getenv("RB_SEARCH_FORCE_SCAN"))) : 0;
        q.scores    = d_scores;
        q.frameOff  = h->dFrameOff.p;
        q.nEmis     = n_emissions;
        q.books     = reinterpret_cast<int4*>(h->dBooks.p);
        q.nBooks    = h->dNBooks.p;
        q.flags     = h->dFlags.p;
        q.entryWord = h->dEntryWord.p;
        q.irrList   = h->dIrrList.p;
        q.nIrr      = h->nIrr;
        q.irrBooks  = nullptr;
        q.nIrrBooks = nullptr;
        if (h->single) {
            RB_CHECK(h->dIrrBooks.reserve((size_t)T * 8));
            RB_CHECK(h->dNIrrBooks.reserve((size_t)n_utt));
            q.irrBooks  = reinterpret_cast<int4*>(h->dIrrBooks.p);
            q.nIrrBooks = h->dNIrrBooks.p;
        }
        auto pick = [&](auto one, auto two, auto three, auto four) {
            return h->npt == 2 ? one : (h->npt == 4 ? two : (h->npt == 8 ? three : four));
        };
        auto k = h->single ? pick(linear_search_reg_kernel<2, true>, linear_search_reg_kernel<4, true>,
                                  linear_search_reg_kernel<8, true>, linear_search_reg_kernel<16, true>)
                           : pick(linear_search_reg_kernel<2, false>, linear_search_reg_kernel<4, false>,
                                  linear_search_reg_kernel<8, false>, linear_search_reg_kernel<16, false>);
        RB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        k<<<n_utt, h->regThreads, smem2, s>>>(q);
    }
    else {
        // per-word kernel: hypotheses in global memory unless they fit into shared memory
        RB_CHECK(h->dHypScore.reserve(stride * n_utt));
        RB_CHECK(h->dHypLm.reserve(stride * n_utt));
        RB_CHECK(h->dHypBkp.reserve(stride * n_utt));
        RB_CHECK(h->dEndScore.reserve((size_t)h->W * n_utt));
        SearchParams p;
        p.wordOff    = h->dWordOff.p;
        p.stateEmis  = h->dStateEmis.p;
        p.stateTdp   = h->dStateTdp.p;
        p.tdp        = h->dTdp.p;
        p.unigram    = h->dUnigram.p;
        p.W          = h->W;
        p.nStates    = h->nStates;
        p.entryModel = h->entryModel;
        p.nModels    = h->nModels;
        p.scores     = d_scores;
        p.frameOff   = h->dFrameOff.p;
        p.nEmis      = n_emissions;
        p.hypScore   = h->dHypScore.p;
        p.hypLm      = h->dHypLm.p;
        p.hypBkp     = h->dHypBkp.p;
        p.books      = reinterpret_cast<int4*>(h->dBooks.p);
        p.nBooks     = h->dNBooks.p;
        p.endScore   = h->dEndScore.p;
        p.single     = h->single ? 1 : 0;
        p.flags      = h->dFlags.p;
        p.entryWord  = h->dEntryWord.p;
        p.irrList    = h->dIrrList.p;
        p.nIrr       = h->nIrr;
        p.irrBooks   = nullptr;
        p.nIrrBooks  = nullptr;
        if (h->single) {
            RB_CHECK(h->dIrrBooks.reserve((size_t)T * 8));
            RB_CHECK(h->dNIrrBooks.reserve((size_t)n_utt));
            p.irrBooks  = reinterpret_cast<int4*>(h->dIrrBooks.p);
            p.nIrrBooks = h->dNIrrBooks.p;
        }
        const size_t smem = (stride * 3 + (size_t)h->W * 2 + (h->W + 1) + (size_t)h->nStates * 2) * 4;
        p.useSmem         = smem <= h->dev.smem_optin - 1024 ? 1 : 0;
        if (p.useSmem)
            RB_CUDA(cudaFuncSetAttribute(linear_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        linear_search_kernel<<<n_utt, kThreads, p.useSmem ? smem : 0, s>>>(p);
    }
    RB_LAUNCH_CHECK();
    RB_CHECK(h->hostBooks.reserve((size_t)T * 8));
    RB_CUDA(cudaMemcpyAsync(h->hostBooks.p, h->dBooks.p, sizeof(int) * 8 * T, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaMemcpyAsync(h->nBooks.data(), h->dNBooks.p, sizeof(int) * n_utt, cudaMemcpyDeviceToHost, s));
    if (h->single) {
        RB_CHECK(h->hostIrrBooks.reserve((size_t)T * 8));
        h->nIrrBooks.assign(n_utt, 0);
        RB_CUDA(cudaMemcpyAsync(h->hostIrrBooks.p, h->dIrrBooks.p, sizeof(int) * 8 * T, cudaMemcpyDeviceToHost, s));
        RB_CUDA(cudaMemcpyAsync(h->nIrrBooks.data(), h->dNIrrBooks.p, sizeof(int) * n_utt, cudaMemcpyDeviceToHost, s));
    }
    RB_CUDA(cudaStreamSynchronize(s));
    return RB_OK;
}

extern "C" int rb_search_decode(rb_search* h, const float* scores, int n_emissions, const int64_t* frame_offsets,
                                int n_utt) {
    RB_REQUIRE(h && frame_offsets && n_utt >= 0 && n_emissions >= 1, "bad argument");
    const int64_t base = frame_offsets[0], T = frame_offsets[n_utt] - base;
    if (n_utt == 0 || T <= 0)
        return rb_search_decode_dev(h, nullptr, n_emissions, frame_offsets, n_utt, nullptr);
    RB_REQUIRE(scores != nullptr, "NULL score buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    RB_CHECK(h->dScores.reserve((size_t)T * n_emissions));
    RB_CUDA(cudaMemcpyAsync(h->dScores.p, scores + base * n_emissions, (size_t)T * n_emissions * 4,
                            cudaMemcpyHostToDevice, h->stream));
    return rb_search_decode_dev(h, h->dScores.p, n_emissions, frame_offsets, n_utt, h->stream);
}

// getCurrentBestSentence of segment `utt` of the last decode: word ends in chronological order.  Any output pointer
// may be NULL; capacity = frames of the segment.  Returns the number of words, < 0 on error.
extern "C" long rb_search_traceback(const rb_search* h, int utt, uint32_t* words, int32_t* times, float* am_scores,
                                    float* lm_scores) {
    if (!h || utt < 0 || utt >= (int)h->nBooks.size()) {
        rb::set_error("no such segment in the last decode");
        return RB_ERR_INVALID;
    }
    const int64_t           f0 = h->frameOff[utt];
    std::vector<const int*> chain;  // records {score, lmScore, word, bkp, time, hadRegularWord, -, -}
    for (int b = h->nBooks[utt] - 1; b != -1; b = h->record(f0, b)[3])
        chain.push_back(h->record(f0, b));
    long n = 0;
    for (auto it = chain.rbegin(); it != chain.rend(); ++it, ++n) {
        const int* r = *it;
        if (words)
            words[n] = (uint32_t)r[2];
        if (times)
            times[n] = r[4];
        if (am_scores)
            memcpy(&am_scores[n], r, 4);
        if (lm_scores)
            memcpy(&lm_scores[n], r + 1, 4);
    }
    return n;
}

// every segment of the last decode in one call: word_offsets [n_utt + 1] are prefix counts into the flat arrays
// (capacity entries each; any may be NULL).  Returns the total number of words, < 0 on error.
extern "C" long rb_search_traceback_all(const rb_search* h, int64_t* word_offsets, uint32_t* words, int32_t* times,
                                        float* am_scores, float* lm_scores, long capacity) {
    if (!h || !word_offsets) {
        rb::set_error("NULL argument");
        return RB_ERR_INVALID;
    }
    const int n_utt  = (int)h->nBooks.size();
    long      total  = 0;
    word_offsets[0]  = 0;
    std::vector<const int*> chain;
    for (int u = 0; u < n_utt; ++u) {
        const int64_t f0 = h->frameOff[u];
        chain.clear();
        for (int b = h->nBooks[u] - 1; b != -1; b = h->record(f0, b)[3])
            chain.push_back(h->record(f0, b));
        for (auto it = chain.rbegin(); it != chain.rend(); ++it, ++total) {
            if (total >= capacity)
                continue;
            const int* r = *it;
            if (words)
                words[total] = (uint32_t)r[2];
            if (times)
                times[total] = r[4];
            if (am_scores)
                memcpy(&am_scores[total], r, 4);
            if (lm_scores)
                memcpy(&lm_scores[total], r + 1, 4);
        }
        word_offsets[u + 1] = total;
    }
    return total;
}
