// MOCK of xla/ffi/api/ffi.h for tests/test_ffi_shim.py -- NOT the real XLA header (JAX is not installable in this image).
// It models just enough of the binding API for `g++ -fsyntax-only ffi/dynax_ffi.cc` to type-check every handler:
// the argument list implied by Ffi::Bind().Ctx<>().Arg<>().Attr<>().Ret<>() must be what the handler function takes,
// and the handler bodies must call the C-ABI of include/dynax_b200.h with the right argument types and counts.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { F32, U8, S32, S64 };
template <DataType>
struct NativeOf;
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<S64> { using type = int64_t; };

struct Span {
    std::vector<int64_t> v;
    size_t size() const { return v.size(); }
    int64_t operator[](size_t i) const { return v[i]; }
};

template <DataType D>
struct Buffer {
    using T = typename NativeOf<D>::type;
    T *typed_data() const { return nullptr; }
    Span dimensions() const { return {}; }
    size_t size_bytes() const { return 0; }
    size_t element_count() const { return 0; }
};

template <DataType D>
struct ResultBuffer {
    Buffer<D> b;
    Buffer<D> *operator->() { return &b; }
};

struct Error {
    static Error Success() { return {}; }
    static Error Internal(const std::string &) { return {}; }
    static Error InvalidArgument(const std::string &) { return {}; }
};

template <class T>
struct PlatformStream {};

template <class... Ts>
struct Binding {
    template <class C> struct CtxType;
    template <class T> struct CtxType<PlatformStream<T>> { using type = T; };
    template <class C> Binding<Ts..., typename CtxType<C>::type> Ctx() const { return {}; }
    template <class B> Binding<Ts..., B> Arg() const { return {}; }
    template <class A> Binding<Ts..., A> Attr(const char *) const { return {}; }
    template <class B> struct RetType;
    template <DataType D> struct RetType<Buffer<D>> { using type = ResultBuffer<D>; };
    template <class B> Binding<Ts..., typename RetType<B>::type> Ret() const { return {}; }
    template <class F> static constexpr bool matches = std::is_invocable_r_v<Error, F, Ts...>;
};

struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, fn, binding)                                                  \
    static_assert(decltype(binding)::template matches<decltype(&fn)>, #name ": binding does not match " #fn); \
    extern "C" void *name() { return (void *)&fn; }
