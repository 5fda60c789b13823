// Test infrastructure: a minimal stand-in for the parts of Rcpp that ggdmc_b200/r/ggdmc_b200_glue.cpp uses, so that the
// R glue can be compiled and driven in an image without R.  R objects are reference-counted nodes (numeric / integer /
// logical / character vectors, lists, S4 objects) with attributes; S4 slots live in the attribute map like in R.
// Not a re-implementation of Rcpp: only the calls the glue makes, with Rcpp's spelling and semantics.
#pragma once
#include <cmath>
#include <cstdint>
#include <initializer_list>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef long R_xlen_t;
inline bool R_finite(double x) { return std::isfinite(x); }

namespace Rcpp {

struct Node;
typedef std::shared_ptr<Node> NodeP;
struct Node {
    enum Kind { NIL, REAL, INT, LGL, STR, LIST, S4OBJ } kind = NIL;
    std::vector<double> real;
    std::vector<int> ints; // INT and LGL
    std::vector<std::string> str;
    std::vector<NodeP> list;
    std::map<std::string, NodeP> attrs; // names, dim, class and the slots of an S4 object
    R_xlen_t length() const
    {
        switch (kind) {
        case REAL: return (R_xlen_t)real.size();
        case INT: case LGL: return (R_xlen_t)ints.size();
        case STR: return (R_xlen_t)str.size();
        case LIST: return (R_xlen_t)list.size();
        default: return 0;
        }
    }
};

[[noreturn]] inline void stop(const std::string &msg) { throw std::runtime_error(msg); }
static std::ostream &Rcout = std::cerr;

class RObject {
public:
    NodeP p;
    RObject() : p(std::make_shared<Node>()) {}
    RObject(NodeP q) : p(q ? q : std::make_shared<Node>()) {}
    RObject(int v) : RObject() { p->kind = Node::INT; p->ints = {v}; }
    RObject(double v) : RObject() { p->kind = Node::REAL; p->real = {v}; }
    RObject(const std::vector<std::string> &v) : RObject() { p->kind = Node::STR; p->str = v; }
    bool isNULL() const { return p->kind == Node::NIL; }
    double scalar() const
    {
        if (p->kind == Node::REAL && !p->real.empty()) return p->real[0];
        if ((p->kind == Node::INT || p->kind == Node::LGL) && !p->ints.empty()) return (double)p->ints[0];
        stop("not a numeric scalar");
    }
};

// element / slot / attribute of a parent node: readable as an RObject, assignable
class Proxy : public RObject {
    NodeP parent;
    std::string key; // attribute or slot name, or list element name
    long index;      // list position (>= 0) when key is empty
public:
    Proxy(NodeP par, const std::string &k, NodeP cur) : RObject(cur), parent(par), key(k), index(-1) {}
    Proxy(NodeP par, long i, NodeP cur) : RObject(cur), parent(par), index(i) {}
    Proxy &operator=(const RObject &v)
    {
        p = v.p;
        if (index >= 0) parent->list.at((size_t)index) = v.p;
        else parent->attrs[key] = v.p;
        return *this;
    }
    Proxy &operator=(int v) { return *this = RObject(v); }
    Proxy &operator=(double v) { return *this = RObject(v); }
    Proxy &operator=(const std::vector<std::string> &v) { return *this = RObject(v); }
    operator int() const { return (int)scalar(); }
    operator double() const { return scalar(); }
};

inline Proxy attr_of(const NodeP &n, const std::string &name)
{
    auto it = n->attrs.find(name);
    return Proxy(n, name, it == n->attrs.end() ? NodeP() : it->second);
}

template <typename T> T as(const RObject &o);
template <> inline double as<double>(const RObject &o) { return o.scalar(); }
template <> inline int as<int>(const RObject &o) { return (int)o.scalar(); }
template <> inline bool as<bool>(const RObject &o) { return o.scalar() != 0.0; }
template <> inline std::string as<std::string>(const RObject &o)
{
    if (o.p->kind != Node::STR || o.p->str.empty()) stop("not a string");
    return o.p->str[0];
}
template <> inline std::vector<std::string> as<std::vector<std::string>>(const RObject &o)
{
    if (o.p->kind == Node::NIL) return {};
    if (o.p->kind != Node::STR) stop("not a character vector");
    return o.p->str;
}

class S4 : public RObject {
public:
    S4(const RObject &o) : RObject(o.p)
    {
        if (p->kind != Node::S4OBJ) stop("not an S4 object");
    }
    explicit S4(const std::string &klass) : RObject()
    {
        p->kind = Node::S4OBJ;
        p->attrs["class"] = RObject(std::vector<std::string>{klass}).p;
    }
    Proxy slot(const std::string &name) const
    {
        return attr_of(p, name); // reading a missing slot yields NULL here; assigning creates it (new("posterior") has them all)
    }
};
template <> inline S4 as<S4>(const RObject &o) { return S4(o); }

template <Node::Kind K, typename T> class Vector : public RObject {
protected:
    std::vector<T> &store() const;
public:
    Vector() : RObject() { p->kind = K; }
    Vector(const RObject &o) : RObject(o.p)
    {
        if (p->kind == Node::NIL) { p = std::make_shared<Node>(); p->kind = K; }
        // R coerces between integer / logical / double on the way into a typed vector: copy-convert
        if (p->kind != K) {
            NodeP q = std::make_shared<Node>();
            q->kind = K;
            q->attrs = p->attrs;
            Vector tmp(q, 0);
            if (p->kind == Node::REAL) for (double v : p->real) tmp.store().push_back((T)v);
            else if (p->kind == Node::INT || p->kind == Node::LGL) for (int v : p->ints) tmp.store().push_back((T)v);
            else stop("cannot coerce to a numeric vector");
            p = q;
        }
    }
    explicit Vector(R_xlen_t n) : RObject() { p->kind = K; store().assign((size_t)n, T()); }
    Vector(const T *first, const T *last) : RObject() { p->kind = K; store().assign(first, last); }
    R_xlen_t size() const { return (R_xlen_t)store().size(); }
    T &operator[](R_xlen_t i) { return store().at((size_t)i); }
    const T &operator[](R_xlen_t i) const { return store().at((size_t)i); }
    typename std::vector<T>::iterator begin() { return store().begin(); }
    typename std::vector<T>::iterator end() { return store().end(); }
    typename std::vector<T>::const_iterator begin() const { return store().begin(); }
    typename std::vector<T>::const_iterator end() const { return store().end(); }
    RObject names() const { return attr_of(p, "names"); }
    Proxy attr(const std::string &name) const { return attr_of(p, name); }
    static Vector create(T a, T b, T c)
    {
        Vector v;
        v.store() = {a, b, c};
        return v;
    }
private:
    Vector(NodeP q, int) : RObject(q) {}
};
template <> inline std::vector<double> &Vector<Node::REAL, double>::store() const { return p->real; }
template <> inline std::vector<int> &Vector<Node::INT, int>::store() const { return p->ints; }
template <> inline std::vector<int> &Vector<Node::LGL, int>::store() const { return p->ints; }
typedef Vector<Node::REAL, double> NumericVector;
typedef Vector<Node::INT, int> IntegerVector;
typedef Vector<Node::LGL, int> LogicalVector;

template <typename V, typename T> class Matrix : public V {
    int nr = 0, nc = 0;
    void read_dim()
    {
        IntegerVector d(this->attr("dim"));
        if (d.size() != 2) stop("not a matrix");
        nr = d[0]; nc = d[1];
    }
public:
    Matrix(const RObject &o) : V(o) { read_dim(); }
    Matrix(int nrow, int ncol, const T *src) : V(src, src + (size_t)nrow * ncol), nr(nrow), nc(ncol)
    {
        IntegerVector d(2);
        d[0] = nrow; d[1] = ncol;
        this->attr("dim") = d;
    }
    int nrow() const { return nr; }
    int ncol() const { return nc; }
    T &operator()(int i, int j) { return (*this)[(R_xlen_t)i + (R_xlen_t)nr * j]; }          // column-major like R
    const T &operator()(int i, int j) const { return (*this)[(R_xlen_t)i + (R_xlen_t)nr * j]; }
};
typedef Matrix<NumericVector, double> NumericMatrix;
typedef Matrix<IntegerVector, int> IntegerMatrix;

struct NamedValue { std::string name; RObject value; };
struct Named {
    std::string name;
    explicit Named(const std::string &n) : name(n) {}
    NamedValue operator=(const RObject &v) const { return NamedValue{name, v}; }
};

class List : public RObject {
public:
    List(const RObject &o) : RObject(o.p)
    {
        if (p->kind != Node::LIST) stop("not a list");
    }
    explicit List(R_xlen_t n) : RObject()
    {
        p->kind = Node::LIST;
        p->list.assign((size_t)n, std::make_shared<Node>());
    }
    R_xlen_t size() const { return (R_xlen_t)p->list.size(); }
    RObject names() const { return attr_of(p, "names"); }
    Proxy operator[](R_xlen_t i) const { return Proxy(p, (long)i, p->list.at((size_t)i)); }
    Proxy operator[](int i) const { return (*this)[(R_xlen_t)i]; }
    Proxy operator[](size_t i) const { return (*this)[(R_xlen_t)i]; }
    Proxy operator[](const char *name) const
    {
        std::vector<std::string> nm = as<std::vector<std::string>>(names());
        for (size_t i = 0; i < nm.size(); ++i)
            if (nm[i] == name) return (*this)[(R_xlen_t)i];
        stop(std::string("no list element named ") + name);
    }
    static List create(const NamedValue &a, const NamedValue &b)
    {
        List l(2);
        l.p->list[0] = a.value.p;
        l.p->list[1] = b.value.p;
        l.p->attrs["names"] = RObject(std::vector<std::string>{a.name, b.name}).p;
        return l;
    }
};

// options() of the mock R session (tests set them through the harness); getOption(name, default) is the one R function
// the glue calls
inline std::map<std::string, std::string> &mock_options()
{
    static std::map<std::string, std::string> o;
    return o;
}
class Function {
    std::string fn;
public:
    Function(const RObject &o) : fn(as<std::string>(o)) {}
    RObject operator()(const char *name, const char *dflt) const
    {
        if (fn != "getOption") stop("the Rcpp stand-in only knows getOption()");
        auto it = mock_options().find(name);
        return RObject(std::vector<std::string>{it == mock_options().end() ? std::string(dflt) : it->second});
    }
};
class Environment {
public:
    static Environment base_env() { return Environment(); }
    RObject operator[](const char *name) const { return RObject(std::vector<std::string>{name}); }
};

} // namespace Rcpp
