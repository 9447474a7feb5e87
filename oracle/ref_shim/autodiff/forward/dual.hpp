// Stand-in for autodiff v1.1.2 <autodiff/forward/dual.hpp> (https://github.com/autodiff/autodiff,
// pinned by the reference's CMakeLists.txt:22-34; NOT vendored under /root/reference).
//
// TEST INFRASTRUCTURE: lets the reference's own newton_raphson.hpp / equation_primitives.hpp
// compile here.  It restates the published evaluation rules of autodiff's first-order forward
// `dual` expression templates for the operators the reference uses (+, -, *, unary -, pow with
// an arithmetic exponent, derivative/wrt/at): lazy expression nodes, the operator rewrite rules,
// and the assign / assignAdd / assignSub / assignMul / assignPow evaluation order.  It is a
// restatement from knowledge of the library, not a copy of its source.
#pragma once

#include <cmath>
#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

namespace autodiff {
namespace detail {

template <typename T, typename G>
struct Dual;

struct NegOp {};
struct AddOp {};
struct MulOp {};
struct PowOp {};
struct NumberDualMulOp {};

template <typename Op, typename R>
struct UnaryExpr {
    R r;
};
template <typename Op, typename L, typename R>
struct BinaryExpr {
    L l;
    R r;
};

template <typename R>
using NegExpr = UnaryExpr<NegOp, R>;
template <typename L, typename R>
using AddExpr = BinaryExpr<AddOp, L, R>;
template <typename L, typename R>
using MulExpr = BinaryExpr<MulOp, L, R>;
template <typename L, typename R>
using PowExpr = BinaryExpr<PowOp, L, R>;
template <typename L, typename R>
using NumberDualMulExpr = BinaryExpr<NumberDualMulOp, L, R>;

template <typename T>
using Plain = std::remove_cv_t<std::remove_reference_t<T>>;

template <typename T>
struct IsDual : std::false_type {};
template <typename T, typename G>
struct IsDual<Dual<T, G>> : std::true_type {};
template <typename T>
struct IsUnary : std::false_type {};
template <typename Op, typename R>
struct IsUnary<UnaryExpr<Op, R>> : std::true_type {};
template <typename T>
struct IsBinary : std::false_type {};
template <typename Op, typename L, typename R>
struct IsBinary<BinaryExpr<Op, L, R>> : std::true_type {};
template <typename T, typename Op>
struct HasOp : std::false_type {};
template <typename Op, typename R>
struct HasOp<UnaryExpr<Op, R>, Op> : std::true_type {};
template <typename Op, typename L, typename R>
struct HasOp<BinaryExpr<Op, L, R>, Op> : std::true_type {};

template <typename T>
constexpr bool isArithmetic = std::is_arithmetic_v<Plain<T>>;
template <typename T>
constexpr bool isDual = IsDual<Plain<T>>::value;
template <typename T>
constexpr bool isUnaryExpr = IsUnary<Plain<T>>::value;
template <typename T>
constexpr bool isBinaryExpr = IsBinary<Plain<T>>::value;
template <typename T>
constexpr bool isExpr = isDual<T> || isUnaryExpr<T> || isBinaryExpr<T>;
template <typename T>
constexpr bool isNegExpr = HasOp<Plain<T>, NegOp>::value;
template <typename T>
constexpr bool isAddExpr = HasOp<Plain<T>, AddOp>::value;
template <typename T>
constexpr bool isMulExpr = HasOp<Plain<T>, MulOp>::value;
template <typename T>
constexpr bool isPowExpr = HasOp<Plain<T>, PowOp>::value;
template <typename T>
constexpr bool isNumberDualMulExpr = HasOp<Plain<T>, NumberDualMulOp>::value;
template <typename L, typename R>
constexpr bool isOperable = (isExpr<L> && isExpr<R>) || (isArithmetic<L> && isExpr<R>) || (isExpr<L> && isArithmetic<R>);

template <bool B>
using Requires = std::enable_if_t<B, bool>;

// ---- evaluation ---------------------------------------------------------------------------
template <typename T, typename G, typename U> constexpr void assign(Dual<T, G>& self, const U& other);
template <typename T, typename G, typename U> constexpr void assignAdd(Dual<T, G>& self, const U& other);
template <typename T, typename G, typename U> constexpr void assignSub(Dual<T, G>& self, const U& other);
template <typename T, typename G, typename U> constexpr void assignMul(Dual<T, G>& self, const U& other);
template <typename T, typename G, typename U> constexpr void assignPow(Dual<T, G>& self, const U& other);

template <typename T, typename G>
struct Dual {
    T val {};
    G grad {};

    constexpr Dual() = default;
    template <typename U, Requires<isArithmetic<U>> = true>
    constexpr Dual(U v) : val(static_cast<T>(v)), grad() {}
    template <typename U, Requires<isExpr<U> && !isDual<U>> = true>
    constexpr Dual(const U& e) { assign(*this, e); }

    template <typename U, Requires<isArithmetic<U> || (isExpr<U> && !isDual<U>)> = true>
    constexpr Dual& operator=(const U& o) { Dual tmp; assign(tmp, o); val = tmp.val; grad = tmp.grad; return *this; }
    template <typename U, Requires<isArithmetic<U> || isExpr<U>> = true>
    constexpr Dual& operator+=(const U& o) { assignAdd(*this, o); return *this; }
    template <typename U, Requires<isArithmetic<U> || isExpr<U>> = true>
    constexpr Dual& operator-=(const U& o) { assignSub(*this, o); return *this; }
    template <typename U, Requires<isArithmetic<U> || isExpr<U>> = true>
    constexpr Dual& operator*=(const U& o) { assignMul(*this, o); return *this; }
    explicit constexpr operator T() const { return val; }
};

template <typename T, typename G>
constexpr void negate(Dual<T, G>& self) { self.val = -self.val; self.grad = -self.grad; }

template <typename T, typename G, typename U>
constexpr void assign(Dual<T, G>& self, const U& other)
{
    static_assert(isExpr<U> || isArithmetic<U>);
    if constexpr (isArithmetic<U>) { self.val = other; self.grad = G(); }
    else if constexpr (isDual<U>) { self.val = other.val; self.grad = other.grad; }
    else if constexpr (isNumberDualMulExpr<U>) { assign(self, other.r); self.val *= other.l; self.grad *= other.l; }
    else if constexpr (isNegExpr<U>) { assign(self, other.r); negate(self); }
    else if constexpr (isAddExpr<U>) { assign(self, other.r); assignAdd(self, other.l); }
    else if constexpr (isMulExpr<U>) { assign(self, other.r); assignMul(self, other.l); }
    else if constexpr (isPowExpr<U>) { assign(self, other.l); assignPow(self, other.r); }
}

template <typename T, typename G, typename U>
constexpr void assignAdd(Dual<T, G>& self, const U& other)
{
    if constexpr (isArithmetic<U>) { self.val += other; }
    else if constexpr (isDual<U>) { self.val += other.val; self.grad += other.grad; }
    else if constexpr (isNegExpr<U>) { assignSub(self, other.r); }
    else if constexpr (isNumberDualMulExpr<U>) { self.val += other.l * other.r.val; self.grad += other.l * other.r.grad; }
    else if constexpr (isAddExpr<U>) { assignAdd(self, other.l); assignAdd(self, other.r); }
    else { Dual<T, G> tmp; assign(tmp, other); assignAdd(self, tmp); }
}

template <typename T, typename G, typename U>
constexpr void assignSub(Dual<T, G>& self, const U& other)
{
    if constexpr (isArithmetic<U>) { self.val -= other; }
    else if constexpr (isDual<U>) { self.val -= other.val; self.grad -= other.grad; }
    else if constexpr (isNegExpr<U>) { assignAdd(self, other.r); }
    else if constexpr (isNumberDualMulExpr<U>) { self.val -= other.l * other.r.val; self.grad -= other.l * other.r.grad; }
    else if constexpr (isAddExpr<U>) { assignSub(self, other.l); assignSub(self, other.r); }
    else { Dual<T, G> tmp; assign(tmp, other); assignSub(self, tmp); }
}

template <typename T, typename G, typename U>
constexpr void assignMul(Dual<T, G>& self, const U& other)
{
    if constexpr (isArithmetic<U>) { self.val *= other; self.grad *= other; }
    else if constexpr (isDual<U>) {
        const G aux = other.grad;  // avoid aliasing when self is other
        self.grad *= other.val;
        self.grad += self.val * aux;
        self.val *= other.val;
    }
    else if constexpr (isNegExpr<U>) { assignMul(self, other.r); negate(self); }
    else if constexpr (isNumberDualMulExpr<U>) { assignMul(self, other.r); assignMul(self, other.l); }
    else if constexpr (isMulExpr<U>) { assignMul(self, other.l); assignMul(self, other.r); }
    else { Dual<T, G> tmp; assign(tmp, other); assignMul(self, tmp); }
}

template <typename T, typename G, typename U>
constexpr void assignPow(Dual<T, G>& self, const U& other)
{
    using std::pow;
    static_assert(isArithmetic<U>, "the stand-in supports pow(expr, number) only");
    const T aux = pow(self.val, other - 1);
    self.grad *= other * aux;
    self.val = aux * self.val;
}

// ---- operators ----------------------------------------------------------------------------
template <typename R, Requires<isExpr<R>> = true>
constexpr auto operator-(const R& r)
{
    if constexpr (isNegExpr<R>) return r.r;                                    // -(-x) => x
    else if constexpr (isNumberDualMulExpr<R>) return (-r.l) * r.r;            // -(number * dual)
    else return NegExpr<R> { r };
}

template <typename L, typename R, Requires<isOperable<L, R>> = true>
constexpr auto operator+(const L& l, const R& r)
{
    if constexpr (isNegExpr<L> && isNegExpr<R>) return -(l.r + r.r);           // (-x) + (-y) => -(x + y)
    else if constexpr (isExpr<L> && isArithmetic<R>) return r + l;             // expr + number => number + expr
    else return AddExpr<L, R> { l, r };
}

template <typename L, typename R, Requires<isOperable<L, R>> = true>
constexpr auto operator*(const L& l, const R& r)
{
    if constexpr (isNegExpr<L> && isNegExpr<R>) return l.r * r.r;              // (-x) * (-y) => x * y
    else if constexpr (isExpr<L> && isArithmetic<R>) return r * l;             // expr * number => number * expr
    else if constexpr (isArithmetic<L> && isNegExpr<R>) return (-l) * r.r;     // number * (-expr)
    else if constexpr (isArithmetic<L> && isNumberDualMulExpr<R>) return (l * r.l) * r.r;
    else if constexpr (isArithmetic<L> && isDual<R>) return NumberDualMulExpr<L, R> { l, r };
    else return MulExpr<L, R> { l, r };
}

template <typename L, typename R, Requires<isOperable<L, R>> = true>
constexpr auto operator-(const L& l, const R& r)
{
    return l + (-r);                                                            // a - b => a + (-b)
}

template <typename L, typename R, Requires<isExpr<L> && isArithmetic<R>> = true>
constexpr auto pow(const L& l, const R& r)
{
    return PowExpr<L, R> { l, r };
}

// ---- derivative(f, wrt(x), at(x, y)) -------------------------------------------------------
template <typename... Vars>
struct Wrt { std::tuple<Vars&...> args; };
template <typename... Args>
struct At { std::tuple<Args&...> args; };

template <typename... Vars>
auto wrt(Vars&... v) { return Wrt<Vars...> { std::tuple<Vars&...>(v...) }; }
template <typename... Args>
auto at(Args&... a) { return At<Args...> { std::tuple<Args&...>(a...) }; }

template <typename Fun, typename Var, typename... Args>
auto derivative(const Fun& f, const Wrt<Var>& w, const At<Args...>& a)
{
    auto& x = std::get<0>(w.args);
    x.grad = 1.0;                                  // seed
    auto u = std::apply(f, a.args);                // the callable takes its arguments by value
    x.grad = 0.0;                                  // unseed
    return u.grad;
}

}  // namespace detail

using dual = detail::Dual<double, double>;
using detail::at;
using detail::derivative;
using detail::wrt;

}  // namespace autodiff
