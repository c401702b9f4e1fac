// TEST INFRASTRUCTURE ONLY (oracle/): a minimal, eagerly evaluated stand-in for the subset of the Eigen 3 API that the
// reference's hot-path sources touch, so that those sources compile UNMODIFIED from /root/reference into oracle/_ref/
// (Eigen itself is a third-party dependency that is absent from this image and from the reference tree).
// Every expression is evaluated immediately into a plain Matrix; blocks, maps and quaternion vec() are views.
// Coefficient formulas that define results (quaternion product / rotation / toRotationMatrix, LLT) follow Eigen's
// documented definitions; reductions run in index order (Eigen's SIMD order differs at the last ulp).
// Nothing under lf-vio_b200/ includes this file.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <type_traits>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { ComputeFullU = 4, ComputeThinU = 8, ComputeFullV = 16, ComputeThinV = 32 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

template <class S, int R, int C, int O = ((R == 1 && C != 1) ? RowMajor : ColMajor), int MR = R, int MC = C>
class Matrix;
template <class X, int BR, int BC>
class Block;
template <class T, int MapOptions = 0, class Stride = void>
class Map;
template <class S>
class Quaternion;
template <class D>
class ArrayWrapper;
template <class S>
class DiagonalWrapper;
template <class M>
class LLT;
template <class M>
class LDLT;
template <class M>
class JacobiSVD;
template <class Q>
class QuatVecView;

namespace internal {
template <class T>
struct traits;
template <class T>
struct traits<const T> : traits<T> {};
template <class S, int R, int C, int O, int MR, int MC>
struct traits<Matrix<S, R, C, O, MR, MC>> {
    typedef S Scalar;
    enum { Rows = R, Cols = C };
};
template <class X, int BR, int BC>
struct traits<Block<X, BR, BC>> {
    typedef typename traits<X>::Scalar Scalar;
    enum { Rows = BR, Cols = BC };
};
template <class T, int MO, class St>
struct traits<Map<T, MO, St>> : traits<T> {};
template <class S>
struct traits<Quaternion<S>> {
    typedef S Scalar;
};
template <class Q>
struct traits<QuatVecView<Q>> {
    typedef typename traits<Q>::Scalar Scalar;
    enum { Rows = 3, Cols = 1 };
};
}  // namespace internal

namespace internal {
constexpr int pick(int a, int b) { return a != Dynamic ? a : b; }
template <class T>
struct is_arith : std::integral_constant<bool, std::is_arithmetic<T>::value> {};
}  // namespace internal

// ------------------------------------------------------------------------------------------------ MatrixBase (CRTP)
template <class Derived>
class MatrixBase {
  public:
    typedef typename internal::traits<Derived>::Scalar Scalar;
    Derived &derived() { return *static_cast<Derived *>(this); }
    const Derived &derived() const { return *static_cast<const Derived *>(this); }
    Index size() const { return derived().rows() * derived().cols(); }

    // ---- element access
    auto operator()(Index i, Index j) const { return derived().coeff(i, j); }
    decltype(auto) operator()(Index i, Index j) { return derived().coeffRef(i, j); }
    auto operator()(Index i) const { return vcoeff(i); }
    decltype(auto) operator()(Index i) { return vcoeffRef(i); }
    auto operator[](Index i) const { return vcoeff(i); }
    decltype(auto) operator[](Index i) { return vcoeffRef(i); }
    auto x() const { return vcoeff(0); }
    auto y() const { return vcoeff(1); }
    auto z() const { return vcoeff(2); }
    auto w() const { return vcoeff(3); }
    decltype(auto) x() { return vcoeffRef(0); }
    decltype(auto) y() { return vcoeffRef(1); }
    decltype(auto) z() { return vcoeffRef(2); }
    decltype(auto) w() { return vcoeffRef(3); }
    auto vcoeff(Index i) const { return derived().cols() == 1 ? derived().coeff(i, 0) : derived().coeff(0, i); }
    decltype(auto) vcoeffRef(Index i) { return derived().cols() == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }

    // ---- plain evaluation
    auto eval() const;
    auto transpose() const;
    auto normalized() const;
    auto inverse() const;
    auto cwiseSqrt() const;
    auto cwiseAbs() const;
    auto array() const;
    auto asDiagonal() const;
    auto norm() const {
        using std::sqrt;
        return sqrt(squaredNorm());
    }
    auto squaredNorm() const {
        Scalar s(0);
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) s += derived().coeff(i, j) * derived().coeff(i, j);
        return s;
    }
    auto sum() const {
        Scalar s(0);
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) s += derived().coeff(i, j);
        return s;
    }
    auto trace() const {
        Scalar s(0);
        for (Index i = 0; i < derived().rows(); i++) s += derived().coeff(i, i);
        return s;
    }
    auto maxCoeff() const {
        auto m = derived().coeff(0, 0);
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) m = std::max(m, derived().coeff(i, j));
        return m;
    }
    auto minCoeff() const {
        auto m = derived().coeff(0, 0);
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) m = std::min(m, derived().coeff(i, j));
        return m;
    }
    auto determinant() const;
    template <class O>
    auto dot(const MatrixBase<O> &o) const {
        Scalar s(0);
        for (Index i = 0; i < size(); i++) s += vcoeff(i) * o.vcoeff(i);
        return s;
    }
    template <class O>
    auto cross(const MatrixBase<O> &o) const;
    void normalize() { *this /= norm(); }
    auto ldlt() const;
    auto llt() const;
    auto jacobiSvd(unsigned opts = 0) const;

    // ---- writers
    Derived &setZero() { return setConstant(Scalar(0)); }
    Derived &setConstant(const Scalar &v) {
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) = v;
        return derived();
    }
    Derived &setIdentity() {
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) = Scalar(i == j ? 1 : 0);
        return derived();
    }
    template <class O>
    Derived &assign(const MatrixBase<O> &o) {
        derived().resizeLike(o.derived().rows(), o.derived().cols());
        if (derived().rows() != o.derived().rows() && derived().rows() == o.derived().cols() && derived().cols() == o.derived().rows() &&
            (derived().rows() == 1 || derived().cols() == 1)) {  // Eigen transposes vectors implicitly on assignment
            for (Index j = 0; j < derived().cols(); j++)
                for (Index i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) = o.derived().coeff(j, i);
            return derived();
        }
        assert(derived().rows() == o.derived().rows() && derived().cols() == o.derived().cols());
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) = o.derived().coeff(i, j);
        return derived();
    }
    template <class O>
    Derived &operator+=(const MatrixBase<O> &o) {
        assert(derived().rows() == o.derived().rows() && derived().cols() == o.derived().cols());
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) += o.derived().coeff(i, j);
        return derived();
    }
    template <class O>
    Derived &operator-=(const MatrixBase<O> &o) {
        assert(derived().rows() == o.derived().rows() && derived().cols() == o.derived().cols());
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) -= o.derived().coeff(i, j);
        return derived();
    }
    template <class T, class = typename std::enable_if<internal::is_arith<T>::value>::type>
    Derived &operator*=(const T &s) {
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) *= s;
        return derived();
    }
    template <class T, class = typename std::enable_if<internal::is_arith<T>::value>::type>
    Derived &operator/=(const T &s) {
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++) derived().coeffRef(i, j) /= s;
        return derived();
    }
    template <class O>
    Derived &operator*=(const MatrixBase<O> &o);

    // ---- comma initialiser
    struct CommaInit {
        Derived &m;
        Index k;
        template <class T>
        CommaInit &operator,(const T &v) {
            put(v);
            return *this;
        }
        template <class T>
        void put(const T &v) {
            Index r = k / m.cols(), c = k % m.cols();
            m.coeffRef(r, c) = Scalar(v);
            k++;
        }
    };
    template <class T, class = typename std::enable_if<internal::is_arith<T>::value>::type>
    CommaInit operator<<(const T &v) {
        CommaInit ci{derived(), 0};
        ci.put(v);
        return ci;
    }

    // ---- blocks (views)
    template <int BR, int BC>
    Block<Derived, BR, BC> block(Index i, Index j) { return Block<Derived, BR, BC>(derived(), i, j, BR, BC); }
    template <int BR, int BC>
    Block<const Derived, BR, BC> block(Index i, Index j) const { return Block<const Derived, BR, BC>(derived(), i, j, BR, BC); }
    Block<Derived, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) { return Block<Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
    Block<const Derived, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) const { return Block<const Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
    template <int BR, int BC>
    Block<Derived, BR, BC> block(Index i, Index j, Index, Index) { return block<BR, BC>(i, j); }
    template <int BR, int BC>
    Block<const Derived, BR, BC> block(Index i, Index j, Index, Index) const { return block<BR, BC>(i, j); }

#define LFVIO_EIGEN_VIEW2(NAME, I, J, NR, NC)                                                                          \
    template <int BR, int BC>                                                                                          \
    Block<Derived, BR, BC> NAME() {                                                                                    \
        const Index R_ = derived().rows(), C_ = derived().cols();                                                      \
        (void)R_; (void)C_;                                                                                            \
        return Block<Derived, BR, BC>(derived(), I, J, BR, BC);                                                        \
    }                                                                                                                  \
    template <int BR, int BC>                                                                                          \
    Block<const Derived, BR, BC> NAME() const {                                                                        \
        const Index R_ = derived().rows(), C_ = derived().cols();                                                      \
        (void)R_; (void)C_;                                                                                            \
        return Block<const Derived, BR, BC>(derived(), I, J, BR, BC);                                                  \
    }                                                                                                                  \
    Block<Derived, Dynamic, Dynamic> NAME(Index BR, Index BC) {                                                        \
        const Index R_ = derived().rows(), C_ = derived().cols();                                                      \
        (void)R_; (void)C_;                                                                                            \
        return Block<Derived, Dynamic, Dynamic>(derived(), I, J, BR, BC);                                              \
    }                                                                                                                  \
    Block<const Derived, Dynamic, Dynamic> NAME(Index BR, Index BC) const {                                            \
        const Index R_ = derived().rows(), C_ = derived().cols();                                                      \
        (void)R_; (void)C_;                                                                                            \
        return Block<const Derived, Dynamic, Dynamic>(derived(), I, J, BR, BC);                                        \
    }
    LFVIO_EIGEN_VIEW2(topLeftCorner, 0, 0, BR, BC)
    LFVIO_EIGEN_VIEW2(topRightCorner, 0, C_ - BC, BR, BC)
    LFVIO_EIGEN_VIEW2(bottomLeftCorner, R_ - BR, 0, BR, BC)
    LFVIO_EIGEN_VIEW2(bottomRightCorner, R_ - BR, C_ - BC, BR, BC)
#undef LFVIO_EIGEN_VIEW2

    // column / row groups: full extent in the other dimension
#define LFVIO_EIGEN_COLS(NAME, J0)                                                                                     \
    template <int N>                                                                                                   \
    Block<Derived, internal::traits<Derived>::Rows, N> NAME() {                                                             \
        const Index C_ = derived().cols(); (void)C_;                                                                   \
        return Block<Derived, internal::traits<Derived>::Rows, N>(derived(), 0, J0, derived().rows(), N);                   \
    }                                                                                                                  \
    template <int N>                                                                                                   \
    Block<const Derived, internal::traits<Derived>::Rows, N> NAME() const {                                                 \
        const Index C_ = derived().cols(); (void)C_;                                                                   \
        return Block<const Derived, internal::traits<Derived>::Rows, N>(derived(), 0, J0, derived().rows(), N);             \
    }                                                                                                                  \
    Block<Derived, internal::traits<Derived>::Rows, Dynamic> NAME(Index N) {                                                \
        const Index C_ = derived().cols(); (void)C_;                                                                   \
        return Block<Derived, internal::traits<Derived>::Rows, Dynamic>(derived(), 0, J0, derived().rows(), N);             \
    }                                                                                                                  \
    Block<const Derived, internal::traits<Derived>::Rows, Dynamic> NAME(Index N) const {                                    \
        const Index C_ = derived().cols(); (void)C_;                                                                   \
        return Block<const Derived, internal::traits<Derived>::Rows, Dynamic>(derived(), 0, J0, derived().rows(), N);       \
    }
    LFVIO_EIGEN_COLS(leftCols, 0)
    LFVIO_EIGEN_COLS(rightCols, C_ - N)
#undef LFVIO_EIGEN_COLS
#define LFVIO_EIGEN_ROWS(NAME, I0)                                                                                     \
    template <int N>                                                                                                   \
    Block<Derived, N, internal::traits<Derived>::Cols> NAME() {                                                             \
        const Index R_ = derived().rows(); (void)R_;                                                                   \
        return Block<Derived, N, internal::traits<Derived>::Cols>(derived(), I0, 0, N, derived().cols());                   \
    }                                                                                                                  \
    template <int N>                                                                                                   \
    Block<const Derived, N, internal::traits<Derived>::Cols> NAME() const {                                                 \
        const Index R_ = derived().rows(); (void)R_;                                                                   \
        return Block<const Derived, N, internal::traits<Derived>::Cols>(derived(), I0, 0, N, derived().cols());             \
    }                                                                                                                  \
    Block<Derived, Dynamic, internal::traits<Derived>::Cols> NAME(Index N) {                                                \
        const Index R_ = derived().rows(); (void)R_;                                                                   \
        return Block<Derived, Dynamic, internal::traits<Derived>::Cols>(derived(), I0, 0, N, derived().cols());             \
    }                                                                                                                  \
    Block<const Derived, Dynamic, internal::traits<Derived>::Cols> NAME(Index N) const {                                    \
        const Index R_ = derived().rows(); (void)R_;                                                                   \
        return Block<const Derived, Dynamic, internal::traits<Derived>::Cols>(derived(), I0, 0, N, derived().cols());       \
    }
    LFVIO_EIGEN_ROWS(topRows, 0)
    LFVIO_EIGEN_ROWS(bottomRows, R_ - N)
#undef LFVIO_EIGEN_ROWS
    Block<Derived, internal::traits<Derived>::Rows, Dynamic> middleCols(Index j, Index n) { return Block<Derived, internal::traits<Derived>::Rows, Dynamic>(derived(), 0, j, derived().rows(), n); }
    Block<const Derived, internal::traits<Derived>::Rows, Dynamic> middleCols(Index j, Index n) const { return Block<const Derived, internal::traits<Derived>::Rows, Dynamic>(derived(), 0, j, derived().rows(), n); }
    Block<Derived, Dynamic, internal::traits<Derived>::Cols> middleRows(Index i, Index n) { return Block<Derived, Dynamic, internal::traits<Derived>::Cols>(derived(), i, 0, n, derived().cols()); }
    Block<const Derived, Dynamic, internal::traits<Derived>::Cols> middleRows(Index i, Index n) const { return Block<const Derived, Dynamic, internal::traits<Derived>::Cols>(derived(), i, 0, n, derived().cols()); }
    Block<Derived, internal::traits<Derived>::Rows, 1> col(Index j) { return Block<Derived, internal::traits<Derived>::Rows, 1>(derived(), 0, j, derived().rows(), 1); }
    Block<const Derived, internal::traits<Derived>::Rows, 1> col(Index j) const { return Block<const Derived, internal::traits<Derived>::Rows, 1>(derived(), 0, j, derived().rows(), 1); }
    Block<Derived, 1, internal::traits<Derived>::Cols> row(Index i) { return Block<Derived, 1, internal::traits<Derived>::Cols>(derived(), i, 0, 1, derived().cols()); }
    Block<const Derived, 1, internal::traits<Derived>::Cols> row(Index i) const { return Block<const Derived, 1, internal::traits<Derived>::Cols>(derived(), i, 0, 1, derived().cols()); }

    // vector segments (column or row vectors)
    template <int N>
    auto segment(Index i) { return vseg<N>(i, N); }
    template <int N>
    auto segment(Index i) const { return vseg<N>(i, N); }
    template <int N>
    auto segment(Index i, Index) { return vseg<N>(i, N); }
    template <int N>
    auto segment(Index i, Index) const { return vseg<N>(i, N); }
    auto segment(Index i, Index n) { return vseg<Dynamic>(i, n); }
    auto segment(Index i, Index n) const { return vseg<Dynamic>(i, n); }
    template <int N>
    auto head() { return vseg<N>(0, N); }
    template <int N>
    auto head() const { return vseg<N>(0, N); }
    auto head(Index n) { return vseg<Dynamic>(0, n); }
    auto head(Index n) const { return vseg<Dynamic>(0, n); }
    template <int N>
    auto tail() { return vseg<N>(size() - N, N); }
    template <int N>
    auto tail() const { return vseg<N>(size() - N, N); }
    auto tail(Index n) { return vseg<Dynamic>(size() - n, n); }
    auto tail(Index n) const { return vseg<Dynamic>(size() - n, n); }

    template <class O>
    bool operator==(const MatrixBase<O> &o) const {
        if (derived().rows() != o.derived().rows() || derived().cols() != o.derived().cols()) return false;
        for (Index j = 0; j < derived().cols(); j++)
            for (Index i = 0; i < derived().rows(); i++)
                if (!(derived().coeff(i, j) == o.derived().coeff(i, j))) return false;
        return true;
    }
    template <class O>
    bool operator!=(const MatrixBase<O> &o) const { return !(*this == o); }
    template <class O>
    bool isApprox(const MatrixBase<O> &o, double prec = 1e-12) const;

  private:
    template <int N>
    auto vseg(Index i, Index n) {
        constexpr bool colvec = internal::traits<Derived>::Cols == 1;
        using B = typename std::conditional<colvec, Block<Derived, N, 1>, Block<Derived, 1, N>>::type;
        return colvec ? B(derived(), i, 0, n, 1) : B(derived(), 0, i, 1, n);
    }
    template <int N>
    auto vseg(Index i, Index n) const {
        constexpr bool colvec = internal::traits<Derived>::Cols == 1;
        using B = typename std::conditional<colvec, Block<const Derived, N, 1>, Block<const Derived, 1, N>>::type;
        return colvec ? B(derived(), i, 0, n, 1) : B(derived(), 0, i, 1, n);
    }
};

// ------------------------------------------------------------------------------------------------ storage
namespace internal {
template <class S, int R, int C>
struct Storage {
    S d[R * C];
    Storage() {
        for (int i = 0; i < R * C; i++) d[i] = S();
    }
    S *data() { return d; }
    const S *data() const { return d; }
    Index rows() const { return R; }
    Index cols() const { return C; }
    void resize(Index r, Index c) {
        assert(r == R && c == C);
        (void)r;
        (void)c;
    }
};
template <class S, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)>
struct StorageSel {
    typedef Storage<S, R, C> type;
};
template <class S, int R, int C>
struct DynStorage {
    std::vector<S> d;
    Index r_ = (R == Dynamic ? 0 : R), c_ = (C == Dynamic ? 0 : C);
    S *data() { return d.data(); }
    const S *data() const { return d.data(); }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    void resize(Index r, Index c) {
        assert((R == Dynamic || r == R) && (C == Dynamic || c == C));
        if (r != r_ || c != c_ || (Index)d.size() != r * c) {
            r_ = r;
            c_ = c;
            d.assign((size_t)(r * c), S());
        }
    }
};
template <class S, int R, int C>
struct StorageSel<S, R, C, true> {
    typedef DynStorage<S, R, C> type;
};
}  // namespace internal

// ------------------------------------------------------------------------------------------------ Matrix
template <class S, int R, int C, int O, int MR, int MC>
class Matrix : public MatrixBase<Matrix<S, R, C, O, MR, MC>> {
    typename internal::StorageSel<S, R, C>::type st;

  public:
    typedef S Scalar;
    typedef MatrixBase<Matrix> Base;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C, IsRowMajor = (O & RowMajor) ? 1 : 0, SizeAtCompileTime = (R == Dynamic || C == Dynamic) ? Dynamic : R * C };

    Matrix() {}
    Matrix(const Matrix &) = default;
    Matrix(Matrix &&) = default;
    Matrix &operator=(const Matrix &) = default;
    Matrix &operator=(Matrix &&) = default;
    template <class D>
    Matrix(const MatrixBase<D> &o) { this->assign(o); }
    template <class D>
    Matrix &operator=(const MatrixBase<D> &o) { return this->assign(o); }
    // one argument: a size for vectors (ignored for fixed sizes), or the single coefficient of a 1x1
    template <class T, class = typename std::enable_if<internal::is_arith<T>::value>::type>
    explicit Matrix(const T &n) {
        if (R == 1 && C == 1)
            st.data()[0] = S(n);
        else if (C == 1 || (R == Dynamic && C == Dynamic))
            st.resize(R == Dynamic ? (Index)n : R, C == Dynamic ? 1 : C);
        else
            st.resize(R == Dynamic ? 1 : R, C == Dynamic ? (Index)n : C);
    }
    // two arguments: coefficients of a fixed 2-vector, else (rows, cols)
    template <class T0, class T1, class = typename std::enable_if<internal::is_arith<T0>::value && internal::is_arith<T1>::value>::type>
    Matrix(const T0 &a, const T1 &b) {
        if (R != Dynamic && C != Dynamic && R * C == 2) {
            st.data()[0] = S(a);
            st.data()[1] = S(b);
        } else
            st.resize((Index)a, (Index)b);
    }
    template <class T0, class T1, class T2>
    Matrix(const T0 &a, const T1 &b, const T2 &c) {
        static_assert(R * C == 3, "3-coefficient constructor on a non-3-vector");
        st.data()[0] = S(a);
        st.data()[1] = S(b);
        st.data()[2] = S(c);
    }
    template <class T0, class T1, class T2, class T3>
    Matrix(const T0 &a, const T1 &b, const T2 &c, const T3 &d) {
        static_assert(R * C == 4, "4-coefficient constructor on a non-4-vector");
        st.data()[0] = S(a);
        st.data()[1] = S(b);
        st.data()[2] = S(c);
        st.data()[3] = S(d);
    }
    explicit Matrix(const S *p) {
        for (Index i = 0; i < rows() * cols(); i++) st.data()[i] = p[i];
    }

    Index rows() const { return st.rows(); }
    Index cols() const { return st.cols(); }
    Index innerStride() const { return 1; }
    S *data() { return st.data(); }
    const S *data() const { return st.data(); }
    const S &coeff(Index i, Index j) const {
        assert(i >= 0 && i < rows() && j >= 0 && j < cols());
        return IsRowMajor ? st.data()[i * cols() + j] : st.data()[j * rows() + i];
    }
    S &coeffRef(Index i, Index j) {
        assert(i >= 0 && i < rows() && j >= 0 && j < cols());
        return IsRowMajor ? st.data()[i * cols() + j] : st.data()[j * rows() + i];
    }
    void resize(Index r, Index c) { st.resize(r, c); }
    void resize(Index n) {
        if (C == 1 || (R == Dynamic && C == Dynamic && cols() <= 1))
            st.resize(n, 1);
        else
            st.resize(1, n);
    }
    void conservativeResize(Index r, Index c) {
        Matrix old = *this;
        st.resize(r, c);
        for (Index j = 0; j < std::min(c, old.cols()); j++)
            for (Index i = 0; i < std::min(r, old.rows()); i++) coeffRef(i, j) = old.coeff(i, j);
    }
    void resizeLike(Index r, Index c) {
        if (R == Dynamic || C == Dynamic) {
            if ((C == 1 && r == 1 && c != 1) || (R == 1 && c == 1 && r != 1)) std::swap(r, c);  // vector = transposed vector
            st.resize(r, c);
        }
    }
    void swap(Matrix &o) { std::swap(st, o.st); }

    static Matrix Zero() {
        Matrix m;
        m.setZero();
        return m;
    }
    static Matrix Zero(Index r, Index c) {
        Matrix m;
        m.resize(r, c);
        m.setZero();
        return m;
    }
    static Matrix Zero(Index n) {
        Matrix m;
        m.resize(n);
        m.setZero();
        return m;
    }
    static Matrix Ones() {
        Matrix m;
        m.setConstant(S(1));
        return m;
    }
    static Matrix Ones(Index r, Index c) {
        Matrix m;
        m.resize(r, c);
        m.setConstant(S(1));
        return m;
    }
    static Matrix Constant(const S &v) {
        Matrix m;
        m.setConstant(v);
        return m;
    }
    static Matrix Identity() {
        Matrix m;
        m.setIdentity();
        return m;
    }
    static Matrix Identity(Index r, Index c) {
        Matrix m;
        m.resize(r, c);
        m.setIdentity();
        return m;
    }
    static Matrix UnitX() {
        Matrix m;
        m.setZero();
        m(0) = S(1);
        return m;
    }
    static Matrix UnitY() {
        Matrix m;
        m.setZero();
        m(1) = S(1);
        return m;
    }
    static Matrix UnitZ() {
        Matrix m;
        m.setZero();
        m(2) = S(1);
        return m;
    }
};

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 1, Dynamic> RowVectorXd;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<int, 3, 1> Vector3i;

// ------------------------------------------------------------------------------------------------ Block (view)
template <class X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC>> {
    X *x;
    Index r0, c0, nr, nc;

  public:
    typedef typename std::remove_const<X>::type::Scalar Scalar;
    typedef MatrixBase<Block> Base;
    enum { RowsAtCompileTime = BR, ColsAtCompileTime = BC, IsRowMajor = 0, SizeAtCompileTime = (BR == Dynamic || BC == Dynamic) ? Dynamic : BR * BC };
    Block(X &xpr, Index i, Index j, Index r, Index c) : x(&xpr), r0(i), c0(j), nr(r), nc(c) {
        assert(i >= 0 && j >= 0 && i + r <= xpr.rows() && j + c <= xpr.cols());
    }
    Block(const Block &) = default;
    Index rows() const { return nr; }
    Index cols() const { return nc; }
    auto coeff(Index i, Index j) const { return static_cast<const typename std::remove_const<X>::type *>(x)->coeff(r0 + i, c0 + j); }
    decltype(auto) coeffRef(Index i, Index j) {
        if constexpr (std::is_const<X>::value)
            return x->coeff(r0 + i, c0 + j);
        else
            return x->coeffRef(r0 + i, c0 + j);
    }
    void resizeLike(Index, Index) {}
    Block &operator=(const Block &o) {
        Matrix<Scalar, BR, BC> tmp(o);
        return this->assign(tmp);
    }
    template <class D>
    Block &operator=(const MatrixBase<D> &o) {
        Matrix<Scalar, internal::pick(BR, D::RowsAtCompileTime), internal::pick(BC, D::ColsAtCompileTime)> tmp(o);  // alias-safe
        return this->assign(tmp);
    }
};

// ------------------------------------------------------------------------------------------------ Map (view of raw memory)
namespace internal {
template <class PS, class S, int R, int C, int O>
class MapImpl : public MatrixBase<Map<typename std::conditional<std::is_const<PS>::value, const Matrix<S, R, C, O>, Matrix<S, R, C, O>>::type>> {
    PS *p;
    Index nr, nc;

  public:
    typedef S Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C, IsRowMajor = (O & RowMajor) ? 1 : 0, SizeAtCompileTime = (R == Dynamic || C == Dynamic) ? Dynamic : R * C };
    MapImpl(PS *ptr) : p(ptr), nr(R), nc(C) { static_assert(R != Dynamic && C != Dynamic, "dynamic Map needs sizes"); }
    MapImpl(PS *ptr, Index n) : p(ptr), nr(C == 1 ? n : (R == Dynamic ? 1 : R)), nc(C == 1 ? 1 : n) {}
    MapImpl(PS *ptr, Index r, Index c) : p(ptr), nr(r), nc(c) {}
    Index rows() const { return nr; }
    Index cols() const { return nc; }
    PS *data() const { return p; }
    const S &coeff(Index i, Index j) const {
        assert(i >= 0 && i < nr && j >= 0 && j < nc);
        return IsRowMajor ? p[i * nc + j] : p[j * nr + i];
    }
    PS &coeffRef(Index i, Index j) {
        assert(i >= 0 && i < nr && j >= 0 && j < nc);
        return IsRowMajor ? p[i * nc + j] : p[j * nr + i];
    }
    void resizeLike(Index, Index) {}
};
}  // namespace internal

template <class S, int R, int C, int O, int MO, class St>
class Map<Matrix<S, R, C, O>, MO, St> : public internal::MapImpl<S, S, R, C, O> {
  public:
    typedef internal::MapImpl<S, S, R, C, O> Impl;
    using Impl::Impl;
    Map(const Map &) = default;
    Map &operator=(const Map &o) {
        Matrix<S, R, C> tmp(o);
        return this->assign(tmp);
    }
    template <class D>
    Map &operator=(const MatrixBase<D> &o) {
        Matrix<S, internal::pick(R, D::RowsAtCompileTime), internal::pick(C, D::ColsAtCompileTime)> tmp(o);
        return this->assign(tmp);
    }
};
template <class S, int R, int C, int O, int MO, class St>
class Map<const Matrix<S, R, C, O>, MO, St> : public internal::MapImpl<const S, S, R, C, O> {
  public:
    typedef internal::MapImpl<const S, S, R, C, O> Impl;
    using Impl::Impl;
};

// ------------------------------------------------------------------------------------------------ results of operations
namespace internal {
template <class A>
using Plain = Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime>;
template <class A, class B>
using Sum = Matrix<typename A::Scalar, pick(A::RowsAtCompileTime, B::RowsAtCompileTime), pick(A::ColsAtCompileTime, B::ColsAtCompileTime)>;
template <class A, class B>
using Prod = Matrix<typename A::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime>;
template <class A>
Plain<A> make_like(const A &a) {
    Plain<A> r;
    r.resize(a.rows(), a.cols());
    return r;
}
}  // namespace internal

template <class D>
auto MatrixBase<D>::eval() const {
    return internal::Plain<D>(*this);
}
template <class D>
auto MatrixBase<D>::transpose() const {
    Matrix<typename D::Scalar, D::ColsAtCompileTime, D::RowsAtCompileTime> r;
    r.resize(derived().cols(), derived().rows());
    for (Index j = 0; j < derived().cols(); j++)
        for (Index i = 0; i < derived().rows(); i++) r.coeffRef(j, i) = derived().coeff(i, j);
    return r;
}
template <class D>
auto MatrixBase<D>::normalized() const {
    internal::Plain<D> r(*this);
    auto n = r.norm();
    r /= n;
    return r;
}
template <class D>
auto MatrixBase<D>::cwiseSqrt() const {
    using std::sqrt;
    internal::Plain<D> r(*this);
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = sqrt(r.coeff(i, j));
    return r;
}
template <class D>
auto MatrixBase<D>::cwiseAbs() const {
    using std::abs;
    internal::Plain<D> r(*this);
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = abs(r.coeff(i, j));
    return r;
}
template <class D>
template <class O>
auto MatrixBase<D>::cross(const MatrixBase<O> &o) const {
    Matrix<typename D::Scalar, 3, 1> r;
    const auto a0 = vcoeff(0), a1 = vcoeff(1), a2 = vcoeff(2);
    const auto b0 = o.vcoeff(0), b1 = o.vcoeff(1), b2 = o.vcoeff(2);
    r(0) = a1 * b2 - a2 * b1;
    r(1) = a2 * b0 - a0 * b2;
    r(2) = a0 * b1 - a1 * b0;
    return r;
}

template <class A, class B>
internal::Sum<A, B> operator+(const MatrixBase<A> &a, const MatrixBase<B> &b) {
    assert(a.derived().rows() == b.derived().rows() && a.derived().cols() == b.derived().cols());
    internal::Sum<A, B> r;
    r.resize(a.derived().rows(), a.derived().cols());
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) + b.derived().coeff(i, j);
    return r;
}
template <class A, class B>
internal::Sum<A, B> operator-(const MatrixBase<A> &a, const MatrixBase<B> &b) {
    assert(a.derived().rows() == b.derived().rows() && a.derived().cols() == b.derived().cols());
    internal::Sum<A, B> r;
    r.resize(a.derived().rows(), a.derived().cols());
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) - b.derived().coeff(i, j);
    return r;
}
template <class A>
internal::Plain<A> operator-(const MatrixBase<A> &a) {
    auto r = internal::make_like(a.derived());
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = -a.derived().coeff(i, j);
    return r;
}
template <class A, class T, class = typename std::enable_if<internal::is_arith<T>::value>::type>
internal::Plain<A> operator*(const MatrixBase<A> &a, const T &s) {
    auto r = internal::make_like(a.derived());
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) * s;
    return r;
}
template <class A, class T, class = typename std::enable_if<internal::is_arith<T>::value>::type>
internal::Plain<A> operator*(const T &s, const MatrixBase<A> &a) {
    auto r = internal::make_like(a.derived());
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = s * a.derived().coeff(i, j);
    return r;
}
template <class A, class T, class = typename std::enable_if<internal::is_arith<T>::value>::type>
internal::Plain<A> operator/(const MatrixBase<A> &a, const T &s) {
    auto r = internal::make_like(a.derived());
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) / s;
    return r;
}
template <class A, class B>
internal::Prod<A, B> operator*(const MatrixBase<A> &a, const MatrixBase<B> &b) {
    const Index n = a.derived().rows(), m = b.derived().cols(), k = a.derived().cols();
    assert(k == b.derived().rows());
    internal::Prod<A, B> r;
    r.resize(n, m);
    for (Index j = 0; j < m; j++)
        for (Index i = 0; i < n; i++) {
            typename A::Scalar s(0);
            for (Index l = 0; l < k; l++) s += a.derived().coeff(i, l) * b.derived().coeff(l, j);
            r.coeffRef(i, j) = s;
        }
    return r;
}
template <class D>
template <class O>
D &MatrixBase<D>::operator*=(const MatrixBase<O> &o) {
    internal::Plain<D> t = (*this) * o;
    return this->assign(t);
}
template <class D>
template <class O>
bool MatrixBase<D>::isApprox(const MatrixBase<O> &o, double prec) const {
    return ((*this) - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
}
template <class D>
std::ostream &operator<<(std::ostream &os, const MatrixBase<D> &m) {
    for (Index i = 0; i < m.derived().rows(); i++) {
        for (Index j = 0; j < m.derived().cols(); j++) os << (j ? " " : "") << m.derived().coeff(i, j);
        if (i + 1 < m.derived().rows()) os << "\n";
    }
    return os;
}

// ------------------------------------------------------------------------------------------------ inverse / determinant (partial-pivot LU)
template <class D>
auto MatrixBase<D>::inverse() const {
    typedef typename D::Scalar S;
    const Index n = derived().rows();
    assert(n == derived().cols());
    internal::Plain<D> a(*this), inv = internal::make_like(derived());
    inv.setIdentity();
    for (Index k = 0; k < n; k++) {
        Index p = k;
        for (Index i = k + 1; i < n; i++)
            if (std::abs(a.coeff(i, k)) > std::abs(a.coeff(p, k))) p = i;
        if (p != k)
            for (Index j = 0; j < n; j++) {
                std::swap(a.coeffRef(k, j), a.coeffRef(p, j));
                std::swap(inv.coeffRef(k, j), inv.coeffRef(p, j));
            }
        const S d = S(1) / a.coeff(k, k);
        for (Index j = 0; j < n; j++) {
            a.coeffRef(k, j) *= d;
            inv.coeffRef(k, j) *= d;
        }
        for (Index i = 0; i < n; i++) {
            if (i == k) continue;
            const S f = a.coeff(i, k);
            if (f == S(0)) continue;
            for (Index j = 0; j < n; j++) {
                a.coeffRef(i, j) -= f * a.coeff(k, j);
                inv.coeffRef(i, j) -= f * inv.coeff(k, j);
            }
        }
    }
    return inv;
}
template <class D>
auto MatrixBase<D>::determinant() const {
    typedef typename D::Scalar S;
    const Index n = derived().rows();
    internal::Plain<D> a(*this);
    S det(1);
    for (Index k = 0; k < n; k++) {
        Index p = k;
        for (Index i = k + 1; i < n; i++)
            if (std::abs(a.coeff(i, k)) > std::abs(a.coeff(p, k))) p = i;
        if (a.coeff(p, k) == S(0)) return S(0);
        if (p != k) {
            for (Index j = 0; j < n; j++) std::swap(a.coeffRef(k, j), a.coeffRef(p, j));
            det = -det;
        }
        det *= a.coeff(k, k);
        for (Index i = k + 1; i < n; i++) {
            const S f = a.coeff(i, k) / a.coeff(k, k);
            for (Index j = k; j < n; j++) a.coeffRef(i, j) -= f * a.coeff(k, j);
        }
    }
    return det;
}

// ------------------------------------------------------------------------------------------------ array(), select(), asDiagonal()
template <class S>
struct BoolArray {
    std::vector<char> v;
    Index r, c;
    template <class T, class E>
    Matrix<S, Dynamic, Dynamic> select2(const T &thenv, const E &elsev) const {
        Matrix<S, Dynamic, Dynamic> o(r, c);
        for (Index j = 0; j < c; j++)
            for (Index i = 0; i < r; i++) o(i, j) = v[(size_t)(j * r + i)] ? thenv(i, j) : elsev(i, j);
        return o;
    }
};
template <class D>
class ArrayWrapper {
  public:
    internal::Plain<D> m;
    typedef typename D::Scalar Scalar;
    explicit ArrayWrapper(const internal::Plain<D> &mm) : m(mm) {}
    ArrayWrapper inverse() const {
        ArrayWrapper r(m);
        for (Index j = 0; j < m.cols(); j++)
            for (Index i = 0; i < m.rows(); i++) r.m.coeffRef(i, j) = Scalar(1) / m.coeff(i, j);
        return r;
    }
    ArrayWrapper sqrt() const {
        ArrayWrapper r(m);
        for (Index j = 0; j < m.cols(); j++)
            for (Index i = 0; i < m.rows(); i++) r.m.coeffRef(i, j) = std::sqrt(m.coeff(i, j));
        return r;
    }
    ArrayWrapper abs() const {
        ArrayWrapper r(m);
        for (Index j = 0; j < m.cols(); j++)
            for (Index i = 0; i < m.rows(); i++) r.m.coeffRef(i, j) = std::abs(m.coeff(i, j));
        return r;
    }
    const internal::Plain<D> &matrix() const { return m; }
    Scalar operator()(Index i, Index j) const { return m.coeff(i, j); }

    struct Mask {
        std::vector<char> v;
        Index r, c;
        // select(array, scalar) and select(array, array)
        template <class T>
        internal::Plain<D> select(const ArrayWrapper &thenv, const T &elsev,
                                  typename std::enable_if<internal::is_arith<T>::value>::type * = 0) const {
            internal::Plain<D> o = thenv.m;
            for (Index j = 0; j < c; j++)
                for (Index i = 0; i < r; i++)
                    if (!v[(size_t)(j * r + i)]) o.coeffRef(i, j) = Scalar(elsev);
            return o;
        }
        internal::Plain<D> select(const ArrayWrapper &thenv, const ArrayWrapper &elsev) const {
            internal::Plain<D> o = thenv.m;
            for (Index j = 0; j < c; j++)
                for (Index i = 0; i < r; i++)
                    if (!v[(size_t)(j * r + i)]) o.coeffRef(i, j) = elsev.m.coeff(i, j);
            return o;
        }
    };
    template <class Cmp>
    Mask mask(Cmp cmp) const {
        Mask k;
        k.r = m.rows();
        k.c = m.cols();
        k.v.resize((size_t)(k.r * k.c));
        for (Index j = 0; j < k.c; j++)
            for (Index i = 0; i < k.r; i++) k.v[(size_t)(j * k.r + i)] = cmp(m.coeff(i, j)) ? 1 : 0;
        return k;
    }
    template <class T>
    Mask operator>(const T &s) const {
        return mask([&](Scalar a) { return a > s; });
    }
    template <class T>
    Mask operator>=(const T &s) const {
        return mask([&](Scalar a) { return a >= s; });
    }
    template <class T>
    Mask operator<(const T &s) const {
        return mask([&](Scalar a) { return a < s; });
    }
    template <class T>
    Mask operator<=(const T &s) const {
        return mask([&](Scalar a) { return a <= s; });
    }
};
template <class D>
auto MatrixBase<D>::array() const {
    return ArrayWrapper<D>(internal::Plain<D>(*this));
}

template <class S>
class DiagonalWrapper {
  public:
    Matrix<S, Dynamic, 1> d;
};
template <class D>
auto MatrixBase<D>::asDiagonal() const {
    DiagonalWrapper<typename D::Scalar> w;
    w.d.resize(size());
    for (Index i = 0; i < size(); i++) w.d(i) = vcoeff(i);
    return w;
}
template <class A, class S>
internal::Plain<A> operator*(const MatrixBase<A> &a, const DiagonalWrapper<S> &w) {
    assert(a.derived().cols() == w.d.size());
    auto r = internal::make_like(a.derived());
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) * w.d(j);
    return r;
}
template <class A, class S>
internal::Plain<A> operator*(const DiagonalWrapper<S> &w, const MatrixBase<A> &a) {
    assert(a.derived().rows() == w.d.size());
    auto r = internal::make_like(a.derived());
    for (Index j = 0; j < r.cols(); j++)
        for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = w.d(i) * a.derived().coeff(i, j);
    return r;
}

// ------------------------------------------------------------------------------------------------ LLT / LDLT
template <class M>
class LLT {
    M L;
    ComputationInfo info_ = Success;

  public:
    LLT() {}
    template <class D>
    explicit LLT(const MatrixBase<D> &a) { compute(a); }
    template <class D>
    LLT &compute(const MatrixBase<D> &a) {
        typedef typename M::Scalar S;
        const Index n = a.derived().rows();
        L = a;
        info_ = Success;
        // right-looking, column by column (lower triangle of A is read, as in Eigen's default UpLo = Lower)
        for (Index k = 0; k < n; k++) {
            S x = L.coeff(k, k);
            for (Index l = 0; l < k; l++) x -= L.coeff(k, l) * L.coeff(k, l);
            if (!(x > S(0))) {
                info_ = NumericalIssue;
                x = std::numeric_limits<S>::quiet_NaN();
            }
            const S d = std::sqrt(x);
            L.coeffRef(k, k) = d;
            for (Index i = k + 1; i < n; i++) {
                S s = L.coeff(i, k);
                for (Index l = 0; l < k; l++) s -= L.coeff(i, l) * L.coeff(k, l);
                L.coeffRef(i, k) = s / d;
            }
        }
        for (Index j = 0; j < n; j++)
            for (Index i = 0; i < j; i++) L.coeffRef(i, j) = S(0);
        return *this;
    }
    const M &matrixL() const { return L; }
    auto matrixU() const { return L.transpose(); }
    ComputationInfo info() const { return info_; }
    template <class D>
    internal::Plain<D> solve(const MatrixBase<D> &b) const {
        typedef typename M::Scalar S;
        const Index n = L.rows();
        internal::Plain<D> x(b);
        for (Index c = 0; c < x.cols(); c++) {
            for (Index i = 0; i < n; i++) {
                S s = x.coeff(i, c);
                for (Index l = 0; l < i; l++) s -= L.coeff(i, l) * x.coeff(l, c);
                x.coeffRef(i, c) = s / L.coeff(i, i);
            }
            for (Index i = n - 1; i >= 0; i--) {
                S s = x.coeff(i, c);
                for (Index l = i + 1; l < n; l++) s -= L.coeff(l, i) * x.coeff(l, c);
                x.coeffRef(i, c) = s / L.coeff(i, i);
            }
        }
        return x;
    }
};

// A.ldlt().solve(b): Eigen's LDLT is a pivoted (Bunch-Kaufman-free, diagonal pivoting) LDLᵀ for semidefinite matrices.
// Stand-in: symmetric diagonal pivoting LDLᵀ (largest remaining diagonal first), same solution for SPD systems.
template <class M>
class LDLT {
    M A;
    std::vector<Index> perm;
    ComputationInfo info_ = Success;

  public:
    template <class D>
    explicit LDLT(const MatrixBase<D> &a) {
        typedef typename M::Scalar S;
        A = a;
        const Index n = A.rows();
        perm.resize((size_t)n);
        for (Index i = 0; i < n; i++) perm[(size_t)i] = i;
        for (Index k = 0; k < n; k++) {
            Index p = k;
            for (Index i = k + 1; i < n; i++)
                if (std::abs(A.coeff(i, i)) > std::abs(A.coeff(p, p))) p = i;
            if (p != k) {
                for (Index j = 0; j < n; j++) std::swap(A.coeffRef(k, j), A.coeffRef(p, j));
                for (Index i = 0; i < n; i++) std::swap(A.coeffRef(i, k), A.coeffRef(i, p));
                std::swap(perm[(size_t)k], perm[(size_t)p]);
            }
            const S d = A.coeff(k, k);
            if (d == S(0)) continue;
            for (Index i = k + 1; i < n; i++) A.coeffRef(i, k) /= d;
            for (Index j = k + 1; j < n; j++)
                for (Index i = j; i < n; i++) {
                    A.coeffRef(i, j) -= A.coeff(i, k) * d * A.coeff(j, k);
                    A.coeffRef(j, i) = A.coeff(i, j);
                }
        }
    }
    ComputationInfo info() const { return info_; }
    template <class D>
    internal::Plain<D> solve(const MatrixBase<D> &b) const {
        typedef typename M::Scalar S;
        const Index n = A.rows();
        internal::Plain<D> x = internal::make_like(b.derived());
        for (Index c = 0; c < x.cols(); c++) {
            std::vector<S> y((size_t)n);
            for (Index i = 0; i < n; i++) y[(size_t)i] = b.derived().coeff(perm[(size_t)i], c);
            for (Index i = 0; i < n; i++)
                for (Index l = 0; l < i; l++) y[(size_t)i] -= A.coeff(i, l) * y[(size_t)l];
            for (Index i = 0; i < n; i++) y[(size_t)i] = A.coeff(i, i) != S(0) ? y[(size_t)i] / A.coeff(i, i) : S(0);
            for (Index i = n - 1; i >= 0; i--)
                for (Index l = i + 1; l < n; l++) y[(size_t)i] -= A.coeff(l, i) * y[(size_t)l];
            for (Index i = 0; i < n; i++) x.coeffRef(perm[(size_t)i], c) = y[(size_t)i];
        }
        return x;
    }
};
template <class D>
auto MatrixBase<D>::ldlt() const {
    return LDLT<internal::Plain<D>>(*this);
}
template <class D>
auto MatrixBase<D>::llt() const {
    return LLT<internal::Plain<D>>(*this);
}

// ------------------------------------------------------------------------------------------------ symmetric eigensolver (cyclic Jacobi)
template <class M>
class SelfAdjointEigenSolver {
    typedef typename M::Scalar S;
    Matrix<S, M::RowsAtCompileTime, 1> w;
    M V;

  public:
    SelfAdjointEigenSolver() {}
    template <class D>
    explicit SelfAdjointEigenSolver(const MatrixBase<D> &a) { compute(a); }
    template <class D>
    SelfAdjointEigenSolver &compute(const MatrixBase<D> &a0) {
        const Index n = a0.derived().rows();
        Matrix<S, Dynamic, Dynamic> a(n, n);
        for (Index j = 0; j < n; j++)
            for (Index i = 0; i < n; i++) a(i, j) = i >= j ? a0.derived().coeff(i, j) : a0.derived().coeff(j, i);  // lower triangle, as Eigen
        Matrix<S, Dynamic, Dynamic> v = Matrix<S, Dynamic, Dynamic>::Identity(n, n);
        for (int sweep = 0; sweep < 100; sweep++) {
            S off(0), diag(0);
            for (Index i = 0; i < n; i++) {
                diag += a(i, i) * a(i, i);
                for (Index j = 0; j < i; j++) off += a(i, j) * a(i, j);
            }
            if (off == S(0) || off <= S(1e-60) * diag) break;
            for (Index p = 0; p < n - 1; p++)
                for (Index q = p + 1; q < n; q++) {
                    const S apq = a(p, q);
                    if (apq == S(0)) continue;
                    const S theta = (a(q, q) - a(p, p)) / (S(2) * apq);
                    const S t = (theta >= 0 ? S(1) : S(-1)) / (std::abs(theta) + std::sqrt(theta * theta + S(1)));
                    const S c = S(1) / std::sqrt(t * t + S(1)), s = t * c;
                    for (Index k = 0; k < n; k++) {
                        const S akp = a(k, p), akq = a(k, q);
                        a(k, p) = c * akp - s * akq;
                        a(k, q) = s * akp + c * akq;
                    }
                    for (Index k = 0; k < n; k++) {
                        const S apk = a(p, k), aqk = a(q, k);
                        a(p, k) = c * apk - s * aqk;
                        a(q, k) = s * apk + c * aqk;
                    }
                    for (Index k = 0; k < n; k++) {
                        const S vkp = v(k, p), vkq = v(k, q);
                        v(k, p) = c * vkp - s * vkq;
                        v(k, q) = s * vkp + c * vkq;
                    }
                }
        }
        std::vector<Index> ord((size_t)n);
        for (Index i = 0; i < n; i++) ord[(size_t)i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](Index x, Index y) { return a(x, x) < a(y, y); });
        w.resize(n);
        V.resize(n, n);
        for (Index j = 0; j < n; j++) {
            w(j) = a(ord[(size_t)j], ord[(size_t)j]);
            for (Index i = 0; i < n; i++) V(i, j) = v(i, ord[(size_t)j]);
        }
        return *this;
    }
    const Matrix<S, M::RowsAtCompileTime, 1> &eigenvalues() const { return w; }
    const M &eigenvectors() const { return V; }
    ComputationInfo info() const { return Success; }
};

// ------------------------------------------------------------------------------------------------ SVD (one-sided Jacobi on A, or on Aᵀ when wide)
template <class M>
class JacobiSVD {
    typedef typename M::Scalar S;
    mutable Matrix<S, Dynamic, Dynamic> left;  // singular vectors of the long side (completed lazily)
    Matrix<S, Dynamic, Dynamic> Rt;            // singular vectors of the short side
    Matrix<S, Dynamic, 1> sv;
    Index rank = 0;
    bool want_full_left = false, is_wide = false;
    mutable bool left_done = false;

    static void hestenes(Matrix<S, Dynamic, Dynamic> &a, Matrix<S, Dynamic, Dynamic> &v) {  // a: m x n, m >= n; columns orthogonalised in place
        const Index m = a.rows(), n = a.cols();
        v = Matrix<S, Dynamic, Dynamic>::Identity(n, n);
        for (int sweep = 0; sweep < 100; sweep++) {
            bool rotated = false;
            for (Index p = 0; p < n - 1; p++)
                for (Index q = p + 1; q < n; q++) {
                    S alpha(0), beta(0), gamma(0);
                    for (Index k = 0; k < m; k++) {
                        alpha += a(k, p) * a(k, p);
                        beta += a(k, q) * a(k, q);
                        gamma += a(k, p) * a(k, q);
                    }
                    if (gamma == S(0) || std::abs(gamma) <= S(1.8e-15) * std::sqrt(alpha * beta)) continue;  // orthogonal to working precision
                    rotated = true;
                    const S zeta = (beta - alpha) / (S(2) * gamma);
                    const S t = (zeta >= 0 ? S(1) : S(-1)) / (std::abs(zeta) + std::sqrt(S(1) + zeta * zeta));
                    const S c = S(1) / std::sqrt(S(1) + t * t), s = c * t;
                    for (Index k = 0; k < m; k++) {
                        const S x = a(k, p), y = a(k, q);
                        a(k, p) = c * x - s * y;
                        a(k, q) = s * x + c * y;
                    }
                    for (Index k = 0; k < n; k++) {
                        const S x = v(k, p), y = v(k, q);
                        v(k, p) = c * x - s * y;
                        v(k, q) = s * x + c * y;
                    }
                }
            if (!rotated) break;
        }
    }
    // complete the first `have` orthonormal columns of q (m x m) to a full orthonormal basis
    static void complete(Matrix<S, Dynamic, Dynamic> &q, Index have) {
        const Index m = q.rows();
        for (Index j = have; j < m; j++) {
            S best(-1);
            Matrix<S, Dynamic, 1> bv(m);
            for (Index e = 0; e < m; e++) {
                Matrix<S, Dynamic, 1> c = Matrix<S, Dynamic, 1>::Zero(m);
                c(e) = S(1);
                for (int rep = 0; rep < 2; rep++)
                    for (Index l = 0; l < j; l++) {
                        S d(0);
                        for (Index k = 0; k < m; k++) d += q(k, l) * c(k);
                        for (Index k = 0; k < m; k++) c(k) -= d * q(k, l);
                    }
                const S nn = c.norm();
                if (nn > best) {
                    best = nn;
                    bv = c / nn;
                }
            }
            for (Index k = 0; k < m; k++) q(k, j) = bv(k);
        }
    }

  public:
    JacobiSVD() {}
    template <class D>
    explicit JacobiSVD(const MatrixBase<D> &a, unsigned opts = 0) { compute(a, opts); }
    template <class D>
    JacobiSVD &compute(const MatrixBase<D> &a0, unsigned opts = 0) {
        const Index m = a0.derived().rows(), n = a0.derived().cols();
        const bool wide = m < n;
        Matrix<S, Dynamic, Dynamic> a = wide ? Matrix<S, Dynamic, Dynamic>(a0.transpose()) : Matrix<S, Dynamic, Dynamic>(a0);
        Matrix<S, Dynamic, Dynamic> v;
        hestenes(a, v);
        const Index mm = a.rows(), nn = a.cols();  // mm >= nn
        std::vector<S> nrm((size_t)nn);
        std::vector<Index> ord((size_t)nn);
        for (Index j = 0; j < nn; j++) {
            S s(0);
            for (Index k = 0; k < mm; k++) s += a(k, j) * a(k, j);
            nrm[(size_t)j] = std::sqrt(s);
            ord[(size_t)j] = j;
        }
        std::stable_sort(ord.begin(), ord.end(), [&](Index x, Index y) { return nrm[(size_t)x] > nrm[(size_t)y]; });
        sv.resize(nn);
        const bool fullL = wide ? (opts & ComputeFullV) : (opts & ComputeFullU);
        left.resize(mm, nn);
        left.setZero();
        Matrix<S, Dynamic, Dynamic> right(nn, nn);
        rank = 0;
        const S tiny = nn ? nrm[(size_t)ord[0]] * S(1e-300) : S(0);
        for (Index j = 0; j < nn; j++) {
            const Index o = ord[(size_t)j];
            sv(j) = nrm[(size_t)o];
            for (Index k = 0; k < nn; k++) right(k, j) = v(k, o);
            if (nrm[(size_t)o] > tiny) {
                for (Index k = 0; k < mm; k++) left(k, j) = a(k, o) / nrm[(size_t)o];
                rank = j + 1;
            }
        }
        // The long side's orthonormal completion (the m x m "full U" of a tall matrix) is only built when it is asked for:
        // the reference requests ComputeFullU | ComputeFullV everywhere but reads matrixV() only (feature_tracker.cpp:202-206).
        want_full_left = fullL;
        left_done = false;
        is_wide = wide;
        Rt = right;
        return *this;
    }
    void finish_left() const {
        if (left_done) return;
        const Index mm = left.rows(), nn = left.cols();
        if (want_full_left || rank < nn) {
            Matrix<S, Dynamic, Dynamic> full(mm, want_full_left ? mm : nn);
            full.setZero();
            for (Index j = 0; j < rank; j++)
                for (Index k = 0; k < mm; k++) full(k, j) = left(k, j);
            if (want_full_left) {
                complete(full, rank);
                left = full;
            } else {
                Matrix<S, Dynamic, Dynamic> sq(mm, mm);
                sq.setZero();
                for (Index j = 0; j < rank; j++)
                    for (Index k = 0; k < mm; k++) sq(k, j) = left(k, j);
                complete(sq, rank);
                for (Index j = rank; j < nn; j++)
                    for (Index k = 0; k < mm; k++) left(k, j) = sq(k, j);
            }
        }
        left_done = true;
    }
    const Matrix<S, Dynamic, Dynamic> &matrixU() const {
        if (is_wide) return Rt;
        finish_left();
        return left;
    }
    const Matrix<S, Dynamic, Dynamic> &matrixV() const {
        if (!is_wide) return Rt;
        finish_left();
        return left;
    }
    const Matrix<S, Dynamic, 1> &singularValues() const { return sv; }
    template <class D>
    Matrix<S, Dynamic, Dynamic> solve(const MatrixBase<D> &b) const {
        Matrix<S, Dynamic, Dynamic> y = matrixU().transpose() * b;
        for (Index i = 0; i < sv.size(); i++)
            for (Index j = 0; j < y.cols(); j++) y(i, j) = sv(i) > sv(0) * S(1e-15) ? y(i, j) / sv(i) : S(0);
        return matrixV().leftCols(sv.size()) * y.topRows(sv.size());
    }
};
template <class D>
auto MatrixBase<D>::jacobiSvd(unsigned opts) const {
    return JacobiSVD<Matrix<typename D::Scalar, Dynamic, Dynamic>>(*this, opts);
}

// ------------------------------------------------------------------------------------------------ quaternions
template <class Q>
class QuatVecView : public MatrixBase<QuatVecView<Q>> {
    Q *q;

  public:
    typedef typename std::remove_const<Q>::type::Scalar Scalar;
    enum { RowsAtCompileTime = 3, ColsAtCompileTime = 1, IsRowMajor = 0, SizeAtCompileTime = 3 };
    explicit QuatVecView(Q &qq) : q(&qq) {}
    QuatVecView(const QuatVecView &) = default;
    Index rows() const { return 3; }
    Index cols() const { return 1; }
    Scalar coeff(Index i, Index) const { return static_cast<const typename std::remove_const<Q>::type *>(q)->c(i); }
    decltype(auto) coeffRef(Index i, Index) { return q->c(i); }  // const Q: c() is the const accessor
    void resizeLike(Index, Index) {}
    template <class D>
    QuatVecView &operator=(const MatrixBase<D> &o) {
        Matrix<Scalar, 3, 1> t(o);
        return this->assign(t);
    }
    QuatVecView &operator=(const QuatVecView &o) {
        Matrix<Scalar, 3, 1> t(o);
        return this->assign(t);
    }
};

template <class Derived>
class QuaternionBase {
  public:
    Derived &derived() { return *static_cast<Derived *>(this); }
    const Derived &derived() const { return *static_cast<const Derived *>(this); }
    // coefficient order in memory is (x, y, z, w), as in Eigen
    auto x() const { return derived().c(0); }
    auto y() const { return derived().c(1); }
    auto z() const { return derived().c(2); }
    auto w() const { return derived().c(3); }
    decltype(auto) x() { return derived().c(0); }
    decltype(auto) y() { return derived().c(1); }
    decltype(auto) z() { return derived().c(2); }
    decltype(auto) w() { return derived().c(3); }

    QuatVecView<Derived> vec() { return QuatVecView<Derived>(derived()); }
    QuatVecView<const Derived> vec() const { return QuatVecView<const Derived>(derived()); }

    auto coeffs() const {
        Matrix<typename internal::traits<Derived>::Scalar, 4, 1> r;
        for (int i = 0; i < 4; i++) r(i) = derived().c(i);
        return r;
    }
    auto squaredNorm() const { return x() * x() + y() * y() + z() * z() + w() * w(); }
    auto norm() const {
        using std::sqrt;
        return sqrt(squaredNorm());
    }
    void normalize() {
        const auto n = norm();
        for (int i = 0; i < 4; i++) derived().c(i) /= n;
    }
    Quaternion<typename internal::traits<Derived>::Scalar> normalized() const;
    Quaternion<typename internal::traits<Derived>::Scalar> conjugate() const;
    Quaternion<typename internal::traits<Derived>::Scalar> inverse() const;
    Derived &setIdentity() {
        typedef typename internal::traits<Derived>::Scalar S;
        derived().c(0) = S(0);
        derived().c(1) = S(0);
        derived().c(2) = S(0);
        derived().c(3) = S(1);
        return derived();
    }
    Matrix<typename internal::traits<Derived>::Scalar, 3, 3> toRotationMatrix() const {
        typedef typename internal::traits<Derived>::Scalar S;
        Matrix<S, 3, 3> res;
        const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
        const S twx = tx * w(), twy = ty * w(), twz = tz * w();
        const S txx = tx * x(), txy = ty * x(), txz = tz * x();
        const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res.coeffRef(0, 0) = S(1) - (tyy + tzz);
        res.coeffRef(0, 1) = txy - twz;
        res.coeffRef(0, 2) = txz + twy;
        res.coeffRef(1, 0) = txy + twz;
        res.coeffRef(1, 1) = S(1) - (txx + tzz);
        res.coeffRef(1, 2) = tyz - twx;
        res.coeffRef(2, 0) = txz - twy;
        res.coeffRef(2, 1) = tyz + twx;
        res.coeffRef(2, 2) = S(1) - (txx + tyy);
        return res;
    }
    template <class O>
    Quaternion<typename internal::traits<Derived>::Scalar> operator*(const QuaternionBase<O> &b) const;
    template <class D>
    Matrix<typename internal::traits<Derived>::Scalar, 3, 1> operator*(const MatrixBase<D> &v) const {
        // v' = v + 2w (q_v × v) + 2 q_v × (q_v × v), the form Eigen's _transformVector uses
        typedef typename internal::traits<Derived>::Scalar S;
        Matrix<S, 3, 1> qv(x(), y(), z()), vv(v);
        Matrix<S, 3, 1> uv = qv.cross(vv);
        uv += uv;
        return vv + w() * uv + qv.cross(uv);
    }
    template <class O>
    auto dot(const QuaternionBase<O> &o) const { return x() * o.x() + y() * o.y() + z() * o.z() + w() * o.w(); }
    template <class O>
    auto angularDistance(const QuaternionBase<O> &o) const {
        auto d = (*this) * o.conjugate();
        return typename internal::traits<Derived>::Scalar(2) * std::atan2(d.vec().norm(), std::abs(d.w()));
    }
    template <class O>
    Derived &assignQ(const QuaternionBase<O> &o) {
        const auto a = o.x(), b = o.y(), c = o.z(), d = o.w();
        derived().c(0) = a;
        derived().c(1) = b;
        derived().c(2) = c;
        derived().c(3) = d;
        return derived();
    }
};

template <class S>
class Quaternion : public QuaternionBase<Quaternion<S>> {
    S q[4];

  public:
    typedef S Scalar;
    Quaternion() { q[0] = q[1] = q[2] = q[3] = S(); }
    Quaternion(const S &w, const S &x, const S &y, const S &z) {
        q[0] = x;
        q[1] = y;
        q[2] = z;
        q[3] = w;
    }
    explicit Quaternion(const S *p) {
        for (int i = 0; i < 4; i++) q[i] = p[i];
    }
    Quaternion(const Quaternion &) = default;
    Quaternion &operator=(const Quaternion &) = default;
    template <class O>
    Quaternion(const QuaternionBase<O> &o) { this->assignQ(o); }
    template <class O>
    Quaternion &operator=(const QuaternionBase<O> &o) { return this->assignQ(o); }
    // from a rotation matrix (3x3) or a coefficient 4-vector (x,y,z,w)
    template <class D>
    explicit Quaternion(const MatrixBase<D> &m) { *this = m; }
    template <class D>
    Quaternion &operator=(const MatrixBase<D> &m) {
        using std::sqrt;
        if (m.derived().rows() == 4 && m.derived().cols() == 1) {
            for (int i = 0; i < 4; i++) q[i] = m.derived().coeff(i, 0);
            return *this;
        }
        // Shepperd's method as Eigen implements it
        S t = m.derived().coeff(0, 0) + m.derived().coeff(1, 1) + m.derived().coeff(2, 2);
        if (t > S(0)) {
            t = sqrt(t + S(1.0));
            q[3] = S(0.5) * t;
            t = S(0.5) / t;
            q[0] = (m.derived().coeff(2, 1) - m.derived().coeff(1, 2)) * t;
            q[1] = (m.derived().coeff(0, 2) - m.derived().coeff(2, 0)) * t;
            q[2] = (m.derived().coeff(1, 0) - m.derived().coeff(0, 1)) * t;
        } else {
            Index i = 0;
            if (m.derived().coeff(1, 1) > m.derived().coeff(0, 0)) i = 1;
            if (m.derived().coeff(2, 2) > m.derived().coeff(i, i)) i = 2;
            Index j = (i + 1) % 3, k = (j + 1) % 3;
            t = sqrt(m.derived().coeff(i, i) - m.derived().coeff(j, j) - m.derived().coeff(k, k) + S(1.0));
            q[i] = S(0.5) * t;
            t = S(0.5) / t;
            q[3] = (m.derived().coeff(k, j) - m.derived().coeff(j, k)) * t;
            q[j] = (m.derived().coeff(j, i) + m.derived().coeff(i, j)) * t;
            q[k] = (m.derived().coeff(k, i) + m.derived().coeff(i, k)) * t;
        }
        return *this;
    }
    const S &c(Index i) const { return q[i]; }
    S &c(Index i) { return q[i]; }
    static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }
    template <class A, class B>
    static Quaternion FromTwoVectors(const MatrixBase<A> &a, const MatrixBase<B> &b) {
        Quaternion r;
        r.setFromTwoVectors(a, b);
        return r;
    }
    template <class A, class B>
    Quaternion &setFromTwoVectors(const MatrixBase<A> &a, const MatrixBase<B> &b) {
        using std::sqrt;
        Matrix<S, 3, 1> v0 = a.normalized(), v1 = b.normalized();
        S cc = v1.dot(v0);
        if (cc < S(-1) + std::numeric_limits<S>::epsilon()) {
            // antiparallel: any axis orthogonal to v0 (Eigen takes the smallest right singular vector of [v0ᵀ; v1ᵀ])
            cc = std::max(cc, S(-1));
            Matrix<S, 3, 1> e = std::abs(v0(0)) < S(0.9) ? Matrix<S, 3, 1>(1, 0, 0) : Matrix<S, 3, 1>(0, 1, 0);
            Matrix<S, 3, 1> axis = v0.cross(e).normalized();
            const S w2 = (S(1) + cc) * S(0.5);
            q[3] = sqrt(w2);
            const S s = sqrt(S(1) - w2);
            q[0] = axis(0) * s;
            q[1] = axis(1) * s;
            q[2] = axis(2) * s;
            return *this;
        }
        Matrix<S, 3, 1> axis = v0.cross(v1);
        const S s = sqrt((S(1) + cc) * S(2));
        const S invs = S(1) / s;
        q[0] = axis(0) * invs;
        q[1] = axis(1) * invs;
        q[2] = axis(2) * invs;
        q[3] = s * S(0.5);
        return *this;
    }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class S, int MO, class St>
class Map<Quaternion<S>, MO, St> : public QuaternionBase<Map<Quaternion<S>, MO, St>> {
    S *p;

  public:
    typedef S Scalar;
    explicit Map(S *ptr) : p(ptr) {}
    Map(const Map &) = default;
    const S &c(Index i) const { return p[i]; }
    S &c(Index i) { return p[i]; }
    template <class O>
    Map &operator=(const QuaternionBase<O> &o) { return this->assignQ(o); }
    Map &operator=(const Map &o) { return this->assignQ(o); }
};
template <class S, int MO, class St>
class Map<const Quaternion<S>, MO, St> : public QuaternionBase<Map<const Quaternion<S>, MO, St>> {
    const S *p;

  public:
    typedef S Scalar;
    explicit Map(const S *ptr) : p(ptr) {}
    const S &c(Index i) const { return p[i]; }
};

template <class D>
Quaternion<typename internal::traits<D>::Scalar> QuaternionBase<D>::conjugate() const {
    return Quaternion<typename D::Scalar>(w(), -x(), -y(), -z());
}
template <class D>
Quaternion<typename internal::traits<D>::Scalar> QuaternionBase<D>::inverse() const {
    typedef typename D::Scalar S;
    const S n2 = squaredNorm();
    if (n2 > S(0)) return Quaternion<S>(w() / n2, -x() / n2, -y() / n2, -z() / n2);
    return Quaternion<S>(S(0), S(0), S(0), S(0));
}
template <class D>
Quaternion<typename internal::traits<D>::Scalar> QuaternionBase<D>::normalized() const {
    typedef typename D::Scalar S;
    const S n = norm();
    return Quaternion<S>(w() / n, x() / n, y() / n, z() / n);
}
template <class D>
template <class O>
Quaternion<typename internal::traits<D>::Scalar> QuaternionBase<D>::operator*(const QuaternionBase<O> &b) const {
    typedef typename D::Scalar S;
    const QuaternionBase &a = *this;
    return Quaternion<S>(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                         a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                         a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                         a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
}

template <class S>
class AngleAxis {
    S ang;
    Matrix<S, 3, 1> ax;

  public:
    template <class D>
    AngleAxis(const S &a, const MatrixBase<D> &axis) : ang(a), ax(axis) {}
    Matrix<S, 3, 3> toRotationMatrix() const {
        Matrix<S, 3, 3> res;
        const S s = std::sin(ang), c = std::cos(ang);
        Matrix<S, 3, 1> sin_axis = s * ax, cos1_axis = (S(1) - c) * ax;
        S tmp = cos1_axis.x() * ax.y();
        res.coeffRef(0, 1) = tmp - sin_axis.z();
        res.coeffRef(1, 0) = tmp + sin_axis.z();
        tmp = cos1_axis.x() * ax.z();
        res.coeffRef(0, 2) = tmp + sin_axis.y();
        res.coeffRef(2, 0) = tmp - sin_axis.y();
        tmp = cos1_axis.y() * ax.z();
        res.coeffRef(1, 2) = tmp - sin_axis.x();
        res.coeffRef(2, 1) = tmp + sin_axis.x();
        res.coeffRef(0, 0) = cos1_axis.x() * ax.x() + c;
        res.coeffRef(1, 1) = cos1_axis.y() * ax.y() + c;
        res.coeffRef(2, 2) = cos1_axis.z() * ax.z() + c;
        return res;
    }
};
typedef AngleAxis<double> AngleAxisd;

}  // namespace Eigen
