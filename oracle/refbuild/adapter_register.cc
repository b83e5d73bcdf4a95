/*
 * adapter_register.cc -- the one line a RASR tool adds to see the adapters: INIT_MODULE(B200)
 * (src/Core/Application.hh:285-286), behind a C entry point so that the tests can do it after
 * loading oracle/_ref/libb200_adapters.so.  TEST INFRASTRUCTURE ONLY; contains no reference code.
 */
#include <Core/Application.hh>

#include "../../adapters/Module.hh"

extern "C" void b200_adapters_register() {
    INIT_MODULE(B200);
}
