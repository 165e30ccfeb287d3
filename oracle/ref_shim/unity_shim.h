/*
 * unity_shim.h -- the slice of UnityEngine / System that the reference's battle code touches, restated in C++ so that
 * the MECHANICAL transliteration of /root/reference/Assets/Script/*.cs (tools/cs2cpp.py -> oracle/_ref/) compiles.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/footsies_oracle.h): loaded by tests/ as the second, independent checker of the
 * oracle.  Nothing here contains battle logic -- the battle logic is the reference's own source text, transliterated.
 *
 * Conventions the transliterator relies on:
 *   - every C# reference type is held as Ref<T> (intrusive count); every member access is written `->`; the value types
 *     below (Rect, Vector2, Vector2Int) define operator-> returning `this`, so `a->b` reads the same for both kinds;
 *   - computed C# properties become methods: Rect.xMin -> xMin(), List.Count -> Count(), array.Length -> Length().
 *
 * Third-party behaviour restated here (closed-source UnityEngine, NOT under /root/reference => "parity unpinned"):
 *   - UnityEngine.Random: Marsaglia xorshift128 as publicly described (seeding s_i = 1812433253 * s_{i-1} + 1;
 *     Range(min,max) = min + next % (max - min));
 *   - UnityEngine.Rect: x is the left edge, xMax = width + x, Overlaps is strict (documented behaviour);
 *   - Time.deltaTime inside FixedUpdate = fixedDeltaTime = 0.02 (ProjectSettings/TimeManager.asset:6).
 */
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

[[noreturn]] inline void cs_throw(const char *what) { std::fprintf(stderr, "unity_shim: %s\n", what); std::abort(); }

/* ---- object model: System.Object + garbage collection by intrusive reference count ---- */
struct Object {
    long _rc = 1;                       /* construction reference, dropped by New<T>() */
    Object() {}
    Object(const Object &) {}           /* MemberwiseClone: the copy starts its own count */
    Object &operator=(const Object &) { return *this; }
    virtual ~Object() {}
};

template <class T> struct Ref {
    typedef T element_type;
    T *p = nullptr;
    Ref() {}
    Ref(std::nullptr_t) {}
    Ref(T *q) : p(q) { retain(); }
    Ref(const Ref &o) : p(o.p) { retain(); }
    template <class U, class = typename std::enable_if<std::is_convertible<U *, T *>::value>::type>
    Ref(const Ref<U> &o) : p(o.p) { retain(); }
    Ref &operator=(const Ref &o) { T *q = o.p; if (q) q->_rc++; release(); p = q; return *this; }   /* self-assignment safe */
    ~Ref() { release(); }
    void retain() { if (p) p->_rc++; }
    void release() { if (p && --p->_rc == 0) delete p; p = nullptr; }
    T *operator->() const { if (!p) cs_throw("NullReferenceException"); return p; }
    T &operator*() const { if (!p) cs_throw("NullReferenceException"); return *p; }
    template <class I, class U = T> auto operator[](I i) const -> decltype(std::declval<U &>()[i]) { return (**this)[i]; }
    template <class U = T> auto begin() const -> decltype(std::declval<U &>().begin()) { return (**this).begin(); }
    template <class U = T> auto end() const -> decltype(std::declval<U &>().end()) { return (**this).end(); }
    friend bool operator==(const Ref &a, const Ref &b) { return a.p == b.p; }
    friend bool operator!=(const Ref &a, const Ref &b) { return a.p != b.p; }
};
template <class T, class... A> Ref<T> New(A &&...a) { T *p = new T(std::forward<A>(a)...); Ref<T> r(p); p->_rc--; return r; }
template <class T, class U> bool Is(const Ref<U> &x) { return dynamic_cast<T *>(x.p) != nullptr; }   /* C# `x is T` */

/* ---- System.Collections.Generic ---- */
template <class T> struct CsArray : Object {      /* T[] */
    std::vector<T> v;
    explicit CsArray(size_t n) : v(n) {}           /* elements start as default(T): 0 / false / null */
    CsArray(std::initializer_list<T> init) : v(init) {}
    int Length() const { return (int)v.size(); }
    T &operator[](long i) { if (i < 0 || (size_t)i >= v.size()) cs_throw("IndexOutOfRangeException"); return v[(size_t)i]; }
    void CopyTo(const Ref<CsArray<T>> &dst, int index) { for (size_t i = 0; i < v.size(); i++) dst[(long)(index + i)] = v[i]; }
    typename std::vector<T>::iterator begin() { return v.begin(); }
    typename std::vector<T>::iterator end() { return v.end(); }
};
template <class T> Ref<CsArray<T>> NewArray(size_t n) { return New<CsArray<T>>(n); }
template <class T> Ref<CsArray<T>> NewArray(size_t n, std::initializer_list<T> init) {
    if (init.size() != n) cs_throw("array initializer length"); return New<CsArray<T>>(init); }
struct Array { template <class A, class V> static void Fill(const A &a, const V &value) { for (auto &e : a->v) e = value; } };

template <class T> struct List : Object {
    std::vector<T> v;
    void Add(const T &x) { v.push_back(x); }
    void Clear() { v.clear(); }
    int Count() const { return (int)v.size(); }
    T &operator[](long i) { if (i < 0 || (size_t)i >= v.size()) cs_throw("ArgumentOutOfRangeException"); return v[(size_t)i]; }
    bool Contains(const T &x) const { for (const T &e : v) if (e == x) return true; return false; }
    Ref<CsArray<T>> ToArray() const { Ref<CsArray<T>> a = NewArray<T>(v.size()); a->v = v; return a; }
    template <class F> void ForEach(F f) { for (size_t i = 0; i < v.size(); i++) f(v[i]); }
    template <class F> T Find(F f) { for (size_t i = 0; i < v.size(); i++) if (f(v[i])) return v[i]; return T(); }
    template <class F> Ref<List<T>> FindAll(F f) { Ref<List<T>> r = New<List<T>>(); for (size_t i = 0; i < v.size(); i++) if (f(v[i])) r->Add(v[i]); return r; }
    typename std::vector<T>::iterator begin() { return v.begin(); }
    typename std::vector<T>::iterator end() { return v.end(); }
};
template <class T> struct Queue : Object {
    std::deque<T> q;
    void Enqueue(const T &x) { q.push_back(x); }
    T Dequeue() { if (q.empty()) cs_throw("InvalidOperationException: Queue empty"); T x = q.front(); q.pop_front(); return x; }
    int Count() const { return (int)q.size(); }
    void Clear() { q.clear(); }
};
template <class K, class V> struct Dictionary : Object {
    std::vector<std::pair<K, V>> kv;                /* insertion-ordered; the battle code only looks keys up */
    void Add(const K &k, const V &v) { if (ContainsKey(k)) cs_throw("ArgumentException: duplicate key"); kv.emplace_back(k, v); }
    bool ContainsKey(const K &k) const { for (auto &e : kv) if (e.first == k) return true; return false; }
    V &operator[](const K &k) { for (auto &e : kv) if (e.first == k) return e.second; cs_throw("KeyNotFoundException"); }
};

/* ---- UnityEngine value types ---- */
struct Vector2 {
    float x = 0, y = 0;
    Vector2() {}
    Vector2(float x_, float y_) : x(x_), y(y_) {}
    Vector2 *operator->() { return this; }
    static const Vector2 zero;
};
struct Vector2Int { int x = 0, y = 0; Vector2Int *operator->() { return this; } const Vector2Int *operator->() const { return this; } };
struct Rect {                                       /* x = LEFT edge (unlike the game's own BoxBase, Fighter.cs:8-26) */
    float x = 0, y = 0, width = 0, height = 0;
    Rect *operator->() { return this; }
    const Rect *operator->() const { return this; }
    float xMin() const { return x; }
    float yMin() const { return y; }
    float xMax() const { return width + x; }
    float yMax() const { return height + y; }
    void Set(float x_, float y_, float w, float h) { x = x_; y = y_; width = w; height = h; }
    bool Overlaps(const Rect &other) const {
        return other.xMax() > xMin() && other.xMin() < xMax() && other.yMax() > yMin() && other.yMin() < yMax();
    }
};

/* ---- UnityEngine statics; all per-thread so that independent games can run on independent threads ---- */
struct Mathf {
    static int Abs(int v) { return v < 0 ? -v : v; }
    static float Abs(float v) { return v < 0 ? -v : v; }
    static float Min(float a, float b) { return a < b ? a : b; }
    static float Max(float a, float b) { return a > b ? a : b; }
};
struct Time {
    static thread_local float deltaTime;            /* = fixedDeltaTime inside FixedUpdate */
    static thread_local float fixedTime;
};
struct RandomState { uint32_t s[4]; const uint32_t *tape; int tape_n, tape_pos; long draws; };
struct Random {                                      /* UnityEngine.Random (process-global in the game) */
    static thread_local RandomState *cur;           /* bound by the harness to the game being stepped */
    static void InitState(RandomState *r, int seed) {
        uint32_t x = (uint32_t)seed;
        for (int i = 0; i < 4; i++) { r->s[i] = x; x = 1812433253u * x + 1u; }
    }
    static void InitState(int seed) { InitState(cur, seed); cur->tape = nullptr; }
    static uint32_t Next(RandomState *r) {
        if (r->tape) { if (r->tape_pos >= r->tape_n) cs_throw("rng tape exhausted"); return r->tape[r->tape_pos++]; }
        uint32_t t = r->s[0] ^ (r->s[0] << 11);
        r->s[0] = r->s[1]; r->s[1] = r->s[2]; r->s[2] = r->s[3];
        r->s[3] = r->s[3] ^ (r->s[3] >> 19) ^ t ^ (t >> 8);
        return r->s[3];
    }
    static int Range(int min, int max) {            /* max exclusive */
        cur->draws++;
        if (max <= min) return min;
        return min + (int)(Next(cur) % (uint32_t)(max - min));
    }
};

/* ---- engine objects the battle code only passes around ---- */
struct AudioClip : Object {};
struct Sprite : Object {};
struct ScriptableObject : Object {};
struct MonoBehaviour : Object {};
struct Animator : Object { void SetTrigger(const char *) {} };
struct GameObject : Object { template <class R> R GetComponent() { return R(New<typename R::element_type>()); } };
struct Task { static Task CompletedTask; };
