// sora_b200/brick/brick.hpp — a from-scratch, ISO C++17 rendering of Sora's BRICK contract.
//
// Purpose: the reference's framework header (kernel/brick/inc/brick.h) only compiles with MSVC/WDK.  So that the GPU
// adaptor bricks (b200_bricks.hpp) can be built and tested here, this header re-creates the *interface* the
// reference bricks are written against — same names, same argument meaning, same call protocol:
//
//   IReferenceCounting / IControlPoint / ISource            brick.h:27-87, 336-353
//   TBrick / TSink / TFilter / TSource                       brick.h:152-172, 246-410
//   DEFINE_IPORT / DEFINE_OPORT, opin(), pin queues          brick.h:182-238, pinqueue.h:6-150
//   BOOL_FUNC_PROCESS, STD_T*_CONSTRUCTOR/RESET/FLUSH        brick.h:174-180, 240-335
//   CREATE_BRICK_SOURCE/FILTER/SINK                          brick.h:411-424
//   DEFINE_LOCAL_CONTEXT / LOCAL_CONTEXT / BIND_CONTEXT /
//   CTX_VAR_RW / CTX_VAR_RO / FACADE_FIELD / RaiseEvent      brick.h:426-470
//
// In an MSVC build of the reference this file is NOT used: b200_bricks.hpp is included after the reference's own
// brick.h.  Nothing here is copied from the reference; it is the minimal conforming subset.
#pragma once
#include <cassert>
#include <climits>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>

#ifndef FINL
#define FINL inline
#endif

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
#include <sys/types.h>                // glibc's `ulong` (64-bit on LP64; the reference's is 32-bit on Windows LLP64 — the
                                      // facade fields only ever hold 32-bit codes, so the wider type is harmless)
struct COMPLEX16 { short re, im; };   // kernel/core/inc/complex.h
struct COMPLEX8 { signed char re, im; };

class CF_VOID {};

// ---- pin queues (pinqueue.h:6-150): producer appends bursts of N, consumer pops bursts of M, capacity lcm(N,M) ----
template <class T, size_t N, size_t M, size_t NSTREAM = 1>
class TPinQueue {
public:
    typedef T DataType;
    static const size_t rsize = M, wsize = N, qsize = (N / std::gcd(N, M)) * M, nstream = NSTREAM;
    TPinQueue() : w_cnt(0), r_cnt(0) {}
    bool check_read() const { return w_cnt - r_cnt >= M; }
    const T* peek(size_t iss = 0) const { return buf_[iss] + r_cnt; }
    void pop() { r_cnt += M; assert(r_cnt <= w_cnt); if (r_cnt == w_cnt) r_cnt = w_cnt = 0; }
    void clear() { w_cnt = r_cnt = 0; }
    size_t count() const { return w_cnt - r_cnt; }
    T* write(size_t iss = 0) { assert(w_cnt < qsize); return buf_[iss] + w_cnt; }
    T* append() { T* p = buf_[0] + w_cnt; w_cnt += N; assert(w_cnt <= qsize); return p; }
    bool pad(const T& item = T()) {
        if (wsize % rsize == 0) return false;
        size_t c = count(), npad = (rsize - c % rsize) % rsize;
        for (size_t i = 0; i < npad; i++) { for (size_t s = 0; s < NSTREAM; s++) buf_[s][w_cnt] = item; w_cnt++; }
        return true;
    }
    void zerobuf() { memset(buf_, 0, sizeof(buf_)); }
private:
    alignas(16) T buf_[NSTREAM][qsize];
    unsigned w_cnt, r_cnt;
};
template <class TYPE, size_t BURST = 1, size_t NSTREAM = 1> struct iport_traits { typedef TYPE type; static const size_t burst = BURST, nstream = NSTREAM; };
template <class TYPE, size_t BURST = 1, size_t NSTREAM = 1> struct oport_traits { typedef TYPE type; static const size_t burst = BURST, nstream = NSTREAM; };
template <class O, class I> struct DeducedPinQueue { typedef TPinQueue<typename O::type, O::burst, I::burst, O::nstream> type; };
template <class O> struct DeducedPinQueue<O, iport_traits<void> > { typedef TPinQueue<typename O::type, O::burst, O::burst, O::nstream> type; };

// ---- object model ----
struct ISource;
class IReferenceCounting {
    mutable unsigned cnt;
protected:
    IReferenceCounting() : cnt(0) {}
public:
    virtual ~IReferenceCounting() {}
    template <class T> static void AddRef(T* p) { assert(p); ++((const IReferenceCounting*)p)->cnt; }
    static void Release(const IReferenceCounting* self) { assert(self); if (--self->cnt == 0) delete self; }
    static void Release(ISource*& self);
};
struct IControlPoint { virtual ~IControlPoint() {} virtual void Reset() = 0; virtual void Flush() = 0; };
struct ISource : public IReferenceCounting, public IControlPoint {
    virtual bool Process() = 0;
    virtual int Seek(int offset) = 0;
    static const int START_POS = INT_MIN, END_POS = INT_MAX;
};
inline void IReferenceCounting::Release(ISource*& self) { if (!self) return; Release((const IReferenceCounting*)self); self = nullptr; }

template <class T_CTX> class TBrick {
protected:
    T_CTX& __ctx_;
public:
    explicit TBrick(T_CTX& ctx) : __ctx_(ctx) {}
    T_CTX& ctx() { return __ctx_; }
};

#define BOOL_FUNC_PROCESS(ipin) template <class T_IPIN> FINL bool Process(T_IPIN& ipin)

#define SB_PORT_PICK(_1, _2, _3, _4, NAME, ...) NAME
#define DEFINE_IPORT(...) SB_PORT_PICK(__VA_ARGS__, SB_X, SB_IPORT3, SB_IPORT2, SB_X)(__VA_ARGS__)
#define SB_IPORT2(TYPE, BURST) typedef ::iport_traits<TYPE, BURST, 1> iport_traits;
#define SB_IPORT3(TYPE, BURST, NS) typedef ::iport_traits<TYPE, BURST, NS> iport_traits;
#define DEFINE_OPORT(...) SB_PORT_PICK(__VA_ARGS__, SB_X, SB_OPORT3, SB_OPORT2, SB_X)(__VA_ARGS__)
#define SB_OPORT2(TYPE, BURST) SB_OPORT3(TYPE, BURST, 1)
#define SB_OPORT3(TYPE, BURST, NS)                                                                         \
    typedef ::oport_traits<TYPE, BURST, NS> oport_traits;                                                  \
    typedef typename DeducedPinQueue<oport_traits, typename T_NEXT::iport_traits>::type opin_type;         \
    opin_type __opin_;                                                                                     \
    opin_type& opin() { return __opin_; }

#define TSINK_ARGS class T_CTX
#define TSINK_PARAMS T_CTX
#define STD_TSINK_CONSTRUCTOR(name) name(T_CTX& ctx) : TSink<T_CTX>(ctx)
#define STD_TSINK_RESET() FINL void Reset()
#define STD_TSINK_FLUSH() FINL void Flush()
template <class T_CTX> class TSink : public TBrick<T_CTX>, public IReferenceCounting, public IControlPoint {
public:
    explicit TSink(T_CTX& ctx) : TBrick<T_CTX>(ctx) {}
    void Reset() override {}
    void Flush() override {}
};

#define TFILTER_ARGS class T_CTX, class T_NEXT
#define TFILTER_PARAMS T_CTX, T_NEXT
#define STD_TFILTER_CONSTRUCTOR(name) name(T_CTX& ctx, T_NEXT* n) : TFilter<T_CTX, T_NEXT>(ctx, n)
#define STD_TFILTER_RESET() FINL void Reset() { this->Next()->Reset(); opin().clear(); __ResetSelf(); } FINL void __ResetSelf()
#define STD_TFILTER_FLUSH() FINL void Flush() { __FlushSelf(); FlushPort(); } FINL void __FlushSelf()
#define FlushPort() do { if (opin().pad()) this->Next()->Process(opin()); this->Next()->Flush(); } while (false)
template <class T_CTX, class T_NEXT> class TFilter : public TBrick<T_CTX>, public IReferenceCounting, public IControlPoint {
protected:
    T_NEXT* __next_;
    T_NEXT* Next() { return __next_; }
public:
    TFilter(T_CTX& ctx, T_NEXT* n) : TBrick<T_CTX>(ctx), __next_(n) { AddRef(n); }
    ~TFilter() { IReferenceCounting::Release(__next_); }
    void Reset() override { __next_->Reset(); }
    void Flush() override { __next_->Flush(); }
};

#define TSOURCE_ARGS class T_CTX, class T_NEXT
#define TSOURCE_PARAMS T_CTX, T_NEXT
#define STD_TSOURCE_CONSTRUCTOR(name) name(T_CTX& ctx, T_NEXT* n) : TSource<T_CTX, T_NEXT>(ctx, n)
#define STD_TSOURCE_RESET() FINL void Reset() { this->Next()->Reset(); opin().clear(); __ResetSelf(); } FINL void __ResetSelf()
#define STD_TSOURCE_FLUSH() FINL void Flush() { __FlushSelf(); FlushPort(); } FINL void __FlushSelf()
template <class T_CTX, class T_NEXT> class TSource : public TBrick<T_CTX>, public ISource {
protected:
    T_NEXT* __next_;
    T_NEXT* Next() { return __next_; }
public:
    TSource(T_CTX& ctx, T_NEXT* n) : TBrick<T_CTX>(ctx), __next_(n) { AddRef(this); AddRef(n); }
    ~TSource() { IReferenceCounting::Release(__next_); }
    void Reset() override { __next_->Reset(); }
    void Flush() override { __next_->Flush(); }
    bool Process() override { return true; }
    int Seek(int) override { return 0; }
};

// ---- graph construction (brick.h:411-424): objects live on the heap, released through the source ----
#define CREATE_BRICK_SINK(NAME, TPL_BRICK, CONTEXT) \
    typedef TPL_BRICK<std::remove_reference<decltype(CONTEXT)>::type> __C_##NAME; __C_##NAME* NAME = new __C_##NAME(CONTEXT);
#define CREATE_BRICK_FILTER(NAME, TPL_BRICK, CONTEXT, NEXT_BRICK) \
    typedef TPL_BRICK<std::remove_reference<decltype(CONTEXT)>::type, std::remove_reference<decltype(*NEXT_BRICK)>::type> __C_##NAME; \
    __C_##NAME* NAME = new __C_##NAME(CONTEXT, NEXT_BRICK);
#define CREATE_BRICK_SOURCE(NAME, TPL_BRICK, CONTEXT, NEXT_BRICK) CREATE_BRICK_FILTER(NAME, TPL_BRICK, CONTEXT, NEXT_BRICK)

// ---- contexts and facades (brick.h:426-470) ----
#define SB_VBASES_1(a) virtual public a
#define SB_VBASES_2(a, b) virtual public a, virtual public b
#define SB_VBASES_3(a, b, c) virtual public a, virtual public b, virtual public c
#define SB_VBASES_4(a, b, c, d) virtual public a, virtual public b, virtual public c, virtual public d
#define SB_VBASES_5(a, b, c, d, e) SB_VBASES_4(a, b, c, d), virtual public e
#define SB_VBASES_6(a, b, c, d, e, f) SB_VBASES_5(a, b, c, d, e), virtual public f
#define SB_VB_PICK(_1, _2, _3, _4, _5, _6, NAME, ...) NAME
#define DEFINE_LOCAL_CONTEXT(TPL_BRICK, ...) \
    struct __LocalContext_##TPL_BRICK : SB_VB_PICK(__VA_ARGS__, SB_VBASES_6, SB_VBASES_5, SB_VBASES_4, SB_VBASES_3, SB_VBASES_2, SB_VBASES_1)(__VA_ARGS__) {};
#define REFERENCE_LOCAL_CONTEXT(TPL_BRICK) typedef struct __LocalContext_##TPL_BRICK __lctx_type;
#define LOCAL_CONTEXT(TPL_BRICK) __LocalContext_##TPL_BRICK
#define RefCtxFunc(fn) static_cast<__lctx_type&>(this->__ctx_).fn
#define RaiseEvent(fn) static_cast<__lctx_type&>(this->__ctx_).fn
#define BIND_CONTEXT(CONTEXT_FUNC, LOCAL_VARIABLE) , LOCAL_VARIABLE(static_cast<__lctx_type&>(ctx).CONTEXT_FUNC())
#define CTX_VAR_RW(TYPE, NAME) TYPE& NAME;
#define CTX_VAR_RO(TYPE, NAME) TYPE const& NAME;
#define FACADE_FIELD(TYPE, NAME) private: TYPE __##NAME##__; public: TYPE& NAME() { return __##NAME##__; }
