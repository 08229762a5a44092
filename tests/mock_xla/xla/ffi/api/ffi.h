// Minimal stand-in for jaxlib's xla/ffi/api/ffi.h: just enough surface for tests/test_capi_host.py to type-check
// integration/jax_ffi/ecnf_jax_ffi.cc (every call into include/ecnf_b200.h) with g++ -fsyntax-only.  Not XLA.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
namespace xla { namespace ffi {
enum DataType { F32, S32, U8 };
template <DataType> struct NativeT;
template <> struct NativeT<F32> { using type = float; };
template <> struct NativeT<S32> { using type = int32_t; };
template <> struct NativeT<U8> { using type = uint8_t; };
template <DataType D> struct Buffer {
  using T = typename NativeT<D>::type;
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  std::vector<int64_t> dimensions() const { return {1}; }
  size_t size_bytes() const { return 0; }
  size_t element_count() const { return 0; }
};
template <DataType D> struct ResultBufferT { Buffer<D> b; Buffer<D>* operator->() { return &b; } };
template <DataType D> using ResultBuffer = ResultBufferT<D>;
struct Error { static Error Success() { return {}; } static Error Internal(std::string) { return {}; } static Error InvalidArgument(std::string) { return {}; } };
template <typename T> struct PlatformStream {};
struct Binding {
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
  template <typename T> Binding& Arg() { return *this; }
  template <typename T> Binding& Ret() { return *this; }
};
struct Ffi { static Binding Bind() { return {}; } };
}}
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) extern "C" void* name() { auto b = binding; (void)b; return (void*)&impl; }
