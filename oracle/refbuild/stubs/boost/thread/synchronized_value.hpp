#pragma once
#include <mutex>
namespace boost { template<class T> class synchronized_value { T v_; public: synchronized_value() {} synchronized_value(const T& v):v_(v){} T* operator->(){return &v_;} const T* operator->() const {return &v_;} T get() const {return v_;} T& value(){return v_;} template<class F> auto operator()(F f){return f(v_);} struct ptr { T* p; T* operator->(){return p;} T& operator*(){return *p;} }; ptr synchronize(){return ptr{&v_};} }; }
