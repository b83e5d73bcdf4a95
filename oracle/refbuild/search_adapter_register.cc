/*
 * search_adapter_register.cc -- makes the Search::SearchAlgorithm adapter (adapters/B200LinearSearch.cc) known to the
 * test host of ref_search.cc; a RASR checkout adds one case to Search::Module_::createRecognizer instead
 * (src/Search/Module.cc:88-110).  TEST INFRASTRUCTURE ONLY; contains no reference code.
 */
#include "../../adapters/B200LinearSearch.hh"

extern "C" void ref_search_set_factory(void* factory);

static B200::LinearSearch* newest = 0;

static Search::SearchAlgorithm* makeB200LinearSearch(const Core::Configuration& c) {
    return newest = new B200::LinearSearch(c);
}

/* of the adapter created last: frames whose scores arrived as whole rows (out[0]) and score(e) calls made (out[1]) */
extern "C" void b200_search_adapter_statistics(unsigned long long* out) {
    out[0] = newest ? newest->nDenseRows() : 0;
    out[1] = newest ? newest->nScoreCalls() : 0;
}

extern "C" void b200_search_adapter_register() {
    ref_search_set_factory((void*)&makeB200LinearSearch);
}
