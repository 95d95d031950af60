// SYNTAX-CHECK STAND-IN for "xla/ffi/api/ffi.h" - TEST INFRASTRUCTURE, not XLA.
//
// JAX / jaxlib are not installable in the build image, so cmad_b200/xla/cmad_b200_xla.cc
// cannot be compiled against the real header here.  This file declares the slice of the
// xla::ffi C++ API that source uses (Buffer / ResultBuffer / Span / Error / the Ffi::Bind()
// chain / XLA_FFI_DEFINE_HANDLER_SYMBOL) with the same names and call shapes, so that
// tests/test_xla_ffi_source.py can run `g++ -fsyntax-only` over the handlers: the C-ABI
// struct fields and entry-point signatures they use are checked against include/cmad_b200.h,
// and every handler's parameter list is checked against its binding (contexts, arguments,
// attributes and results in binding order, as the real binder passes them).
// The real build (cmad_b200/xla/__init__.py:build) uses jax.ffi.include_dir().
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum DataType { F64, F32, S32, S64, U8 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<U8> { using type = uint8_t; };

template <class T>
class Span {
  public:
    Span() = default;
    Span(T* d, size_t n) : d_(d), n_(n) {}
    T* data() const { return d_; }
    size_t size() const { return n_; }
    T& operator[](size_t i) const { return d_[i]; }
  private:
    T* d_ = nullptr;
    size_t n_ = 0;
};

template <DataType DT>
class Buffer {
  public:
    using T = typename NativeOf<DT>::type;
    Span<const int64_t> dimensions() const { return {}; }
    T* typed_data() const { return nullptr; }
    size_t element_count() const { return 0; }
};

template <class B>
class Result {
  public:
    B* operator->() { return &b_; }
    const B* operator->() const { return &b_; }
  private:
    B b_;
};
template <DataType DT> using ResultBuffer = Result<Buffer<DT>>;

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kUnimplemented };
class Error {
  public:
    Error() = default;
    Error(ErrorCode c, std::string m) : c_(c), m_(std::move(m)) {}
    static Error Success() { return Error(); }
    static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
    static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
    bool failure() const { return c_ != ErrorCode::kOk; }
    bool success() const { return c_ == ErrorCode::kOk; }
  private:
    ErrorCode c_ = ErrorCode::kOk;
    std::string m_;
};

template <class T> struct PlatformStream {};
template <class C> struct CtxType;
template <class T> struct CtxType<PlatformStream<T>> { using type = T; };
template <class R> struct RetType { using type = Result<R>; };

template <class... Ts>
struct Binding {
    template <class C> Binding<Ts..., typename CtxType<C>::type> Ctx() const { return {}; }
    template <class A> Binding<Ts..., A> Arg() const { return {}; }
    template <class A> Binding<Ts..., A> Attr(const char*) const { return {}; }
    template <class R> Binding<Ts..., typename RetType<R>::type> Ret() const { return {}; }
    template <class Fn> static constexpr bool matches() { return std::is_invocable_r<Error, Fn, Ts...>::value; }
};

struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                              \
    static_assert(decltype(binding)::template matches<decltype(&impl)>(),                               \
                  #impl ": parameter list does not match its binding (ctx, args, attrs, rets in order)"); \
    extern "C" void* name(void*) { return nullptr; }
