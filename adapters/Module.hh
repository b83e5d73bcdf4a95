/* Module "B200": registers the adapters (pattern: src/Onnx/Module.{hh,cc}); a tool activates it with
 * INIT_MODULE(B200) (src/Core/Application.hh:285-286). */
#ifndef _B200_MODULE_HH
#define _B200_MODULE_HH

#include <Core/Singleton.hh>

namespace B200 {

class Module_ {
public:
    Module_();
    ~Module_() = default;
};

typedef Core::SingletonHolder<Module_> Module;

}  // namespace B200

#endif
